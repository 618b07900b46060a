// crp_cuda.cu - runtime wrappers and data-movement kernels of the thin C-ABI
// CUDA layer (include/crp_cuda.h lists what each entry point replaces).
// sm_100a only; no other architecture is built.
#include <atomic>
#include <cstring>

#include "crp_cuda_internal.cuh"

static std::atomic<unsigned long long> g_launches{0};
void crp_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" unsigned long long crp_kernel_launch_count(void) { return g_launches.load(); }

// ------------------------------------------------------------------ devices
extern "C" int crp_cuda_device_count(void)
{
    int n = 0;
    cudaError_t err = cudaGetDeviceCount(&n);
    if (err != cudaSuccess)
    {
        (void) cudaGetLastError();
        return 0;
    }
    return n;
}

static int env_int(const char *name, int *out)
{
    const char *v = getenv(name);
    if (v == NULL || v[0] == 0) return 0;
    *out = atoi(v);
    return 1;
}

extern "C" void crp_cuda_select_device_by_local_rank(void)
{
    // same idea as the reference's launcher-variable probing (cuda_proxy.cu:11-46), plus minimpi / torchrun
    static const char *names[] = {
        "MINIMPI_LOCAL_RANK", "LOCAL_RANK", "MPI_LOCALRANKID", "MV2_COMM_WORLD_LOCAL_RANK",
        "OMPI_COMM_WORLD_LOCAL_RANK", "OMPI_COMM_WORLD_NODE_RANK", "SLURM_LOCALID", "PBS_O_VNODENUM"
    };
    int local_rank = 0;
    for (size_t i = 0; i < sizeof(names) / sizeof(names[0]); i++)
        if (env_int(names[i], &local_rank)) break;
    int ngpu = crp_cuda_device_count();
    if (ngpu <= 0)
    {
        fprintf(stderr, "[FATAL] no CUDA device visible; CRP-SpMM has no CPU fallback\n");
        abort();
    }
    CRP_CUDA_CHECK(cudaSetDevice(local_rank % ngpu));
}

extern "C" void crp_cuda_set_device(const int dev_id) { CRP_CUDA_CHECK(cudaSetDevice(dev_id)); }

extern "C" int crp_cuda_get_device(void)
{
    int d = 0;
    CRP_CUDA_CHECK(cudaGetDevice(&d));
    return d;
}

extern "C" int crp_cuda_sm_count(void)
{
    int d = crp_cuda_get_device(), n = 0;
    CRP_CUDA_CHECK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d));
    return n;
}

extern "C" int crp_cuda_ptr_is_device(const void *ptr)
{
    if (ptr == NULL) return 0;
    cudaPointerAttributes attr;
    cudaError_t err = cudaPointerGetAttributes(&attr, ptr);
    if (err != cudaSuccess)
    {
        (void) cudaGetLastError();
        return 0;
    }
    return (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) ? 1 : 0;
}

// ------------------------------------------------------------------- memory
extern "C" void crp_cuda_malloc_dev(void **dptr_, const size_t bytes)
{
    *dptr_ = NULL;
    if (bytes == 0) return;
    CRP_CUDA_CHECK(cudaMalloc(dptr_, bytes));
}

extern "C" void crp_cuda_malloc_host(void **hptr_, const size_t bytes)
{
    *hptr_ = NULL;
    if (bytes == 0) return;
    CRP_CUDA_CHECK(cudaMallocHost(hptr_, bytes));
}

extern "C" void crp_cuda_free_dev(void *dptr) { if (dptr) CRP_CUDA_CHECK(cudaFree(dptr)); }
extern "C" void crp_cuda_free_host(void *hptr) { if (hptr) CRP_CUDA_CHECK(cudaFreeHost(hptr)); }

extern "C" void crp_cuda_memset_dev(void *dptr, const int value, const size_t bytes)
{
    if (bytes) CRP_CUDA_CHECK(cudaMemset(dptr, value, bytes));
}

extern "C" void crp_cuda_memset_async(void *dptr, const int value, const size_t bytes, void *stream)
{
    if (bytes) CRP_CUDA_CHECK(cudaMemsetAsync(dptr, value, bytes, as_stream(stream)));
}

extern "C" void crp_cuda_memcpy_h2d(const void *hptr, void *dptr, const size_t bytes)
{
    if (bytes) CRP_CUDA_CHECK(cudaMemcpy(dptr, hptr, bytes, cudaMemcpyHostToDevice));
}

extern "C" void crp_cuda_memcpy_d2h(const void *dptr, void *hptr, const size_t bytes)
{
    if (bytes) CRP_CUDA_CHECK(cudaMemcpy(hptr, dptr, bytes, cudaMemcpyDeviceToHost));
}

extern "C" void crp_cuda_memcpy_d2d(const void *dptr_src, void *dptr_dst, const size_t bytes)
{
    if (bytes) CRP_CUDA_CHECK(cudaMemcpy(dptr_dst, dptr_src, bytes, cudaMemcpyDeviceToDevice));
}

extern "C" void crp_cuda_memcpy_auto(const void *src, void *dst, const size_t bytes)
{
    if (bytes) CRP_CUDA_CHECK(cudaMemcpy(dst, src, bytes, cudaMemcpyDefault));
}

extern "C" void crp_cuda_memcpy_async(const void *src, void *dst, const size_t bytes, void *stream)
{
    if (bytes) CRP_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, as_stream(stream)));
}

extern "C" void crp_cuda_memcpy2d_async(const void *src, size_t src_pitch, void *dst, size_t dst_pitch, size_t row_bytes, size_t nrow, void *stream)
{
    if (row_bytes == 0 || nrow == 0) return;
    CRP_CUDA_CHECK(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, row_bytes, nrow, cudaMemcpyDefault, as_stream(stream)));
}

extern "C" int crp_cuda_host_register(const void *hptr, const size_t bytes)
{
    if (hptr == NULL || bytes == 0) return 0;
    cudaError_t err = cudaHostRegister((void *) hptr, bytes, cudaHostRegisterDefault);
    if (err != cudaSuccess)
    {
        (void) cudaGetLastError();
        return 0;
    }
    return 1;
}

extern "C" void crp_cuda_host_unregister(const void *hptr)
{
    if (hptr == NULL) return;
    if (cudaHostUnregister((void *) hptr) != cudaSuccess) (void) cudaGetLastError();
}

// -------------------------------------------------------- streams and events
extern "C" void crp_cuda_device_sync(void) { CRP_CUDA_CHECK(cudaDeviceSynchronize()); }
extern "C" void crp_cuda_stream_sync(void *stream) { CRP_CUDA_CHECK(cudaStreamSynchronize(as_stream(stream))); }

extern "C" void *crp_cuda_stream_create(void)
{
    cudaStream_t s;
    CRP_CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    return (void *) s;
}

extern "C" void *crp_cuda_stream_create_high_priority(void)
{
    int lo = 0, hi = 0;
    CRP_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    cudaStream_t s;
    CRP_CUDA_CHECK(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, hi));
    return (void *) s;
}

extern "C" void crp_cuda_stream_destroy(void *stream) { if (stream) CRP_CUDA_CHECK(cudaStreamDestroy(as_stream(stream))); }

extern "C" void *crp_cuda_event_create(void)
{
    cudaEvent_t e;
    CRP_CUDA_CHECK(cudaEventCreate(&e));
    return (void *) e;
}

extern "C" void crp_cuda_event_destroy(void *event) { if (event) CRP_CUDA_CHECK(cudaEventDestroy((cudaEvent_t) event)); }
extern "C" void crp_cuda_event_record(void *event, void *stream) { CRP_CUDA_CHECK(cudaEventRecord((cudaEvent_t) event, as_stream(stream))); }
extern "C" void crp_cuda_event_sync(void *event) { CRP_CUDA_CHECK(cudaEventSynchronize((cudaEvent_t) event)); }
extern "C" int crp_cuda_event_done(void *event)
{
    cudaError_t err = cudaEventQuery((cudaEvent_t) event);
    if (err == cudaSuccess) return 1;
    if (err == cudaErrorNotReady) return 0;
    CRP_CUDA_CHECK(err);
    return 0;
}

extern "C" void crp_cuda_stream_wait_event(void *stream, void *event) { CRP_CUDA_CHECK(cudaStreamWaitEvent(as_stream(stream), (cudaEvent_t) event, 0)); }

extern "C" float crp_cuda_event_elapsed_ms(void *start, void *stop)
{
    float ms = 0.f;
    CRP_CUDA_CHECK(cudaEventElapsedTime(&ms, (cudaEvent_t) start, (cudaEvent_t) stop));
    return ms;
}

// ---------------------------------------------------- data-movement kernels
// All of these are pure HBM-bound byte movers: 128-bit accesses whenever every
// address involved is 16-byte aligned, one element of VB bytes per thread step,
// rows mapped onto consecutive threads so that warps read and write whole lines.

template <typename V>
__device__ __forceinline__ void copy_rows_body(
    const char *__restrict__ src, size_t src_pitch, char *__restrict__ dst, size_t dst_pitch,
    uint32_t nrow, uint32_t row_bytes, const int *__restrict__ ridx, size_t tid, size_t nthreads
)
{
    const uint32_t vpr = row_bytes / (uint32_t) sizeof(V);
    const size_t total = (size_t) nrow * vpr;
    for (size_t t = tid; t < total; t += nthreads)
    {
        const uint32_t r = (uint32_t) (t / vpr), v = (uint32_t) (t - (size_t) r * vpr);
        const size_t sr = ridx ? (size_t) ridx[r] : (size_t) r;
        const V val = *reinterpret_cast<const V *>(src + sr * src_pitch + (size_t) v * sizeof(V));
        *reinterpret_cast<V *>(dst + (size_t) r * dst_pitch + (size_t) v * sizeof(V)) = val;
    }
}

__device__ __forceinline__ int common_alignment(uintptr_t a, uintptr_t b, size_t c, size_t d, size_t e)
{
    const uintptr_t all = a | b | (uintptr_t) c | (uintptr_t) d | (uintptr_t) e;
    if ((all & 15) == 0) return 16;
    if ((all & 7) == 0) return 8;
    return 4;
}

// dst[r, :] = src[ridx ? ridx[r] : r, :]
__global__ void __launch_bounds__(256) copy_rows_kernel(
    const char *__restrict__ src, size_t src_pitch, char *__restrict__ dst, size_t dst_pitch,
    uint32_t nrow, uint32_t row_bytes, const int *__restrict__ ridx
)
{
    const size_t tid = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nth = (size_t) gridDim.x * blockDim.x;
    const int al = common_alignment((uintptr_t) src, (uintptr_t) dst, src_pitch, dst_pitch, row_bytes);
    if (al == 16)     copy_rows_body<uint4>(src, src_pitch, dst, dst_pitch, nrow, row_bytes, ridx, tid, nth);
    else if (al == 8) copy_rows_body<uint2>(src, src_pitch, dst, dst_pitch, nrow, row_bytes, ridx, tid, nth);
    else              copy_rows_body<uint32_t>(src, src_pitch, dst, dst_pitch, nrow, row_bytes, ridx, tid, nth);
}

static int copy_grid(size_t work_items)
{
    // enough CTAs to cover the work once, capped at a few waves of the 148 SMs
    size_t blocks = (work_items + 255) / 256;
    const size_t cap = (size_t) 148 * 8 * 4;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int) blocks;
}

extern "C" void crp_cuda_copy_matrix_async(size_t dt_size, const int nrow, const int ncol, const void *src, const int lds, void *dst, const int ldd, void *stream)
{
    if (nrow <= 0 || ncol <= 0) return;
    if (dt_size != 4 && dt_size != 8) { fprintf(stderr, "[FATAL] crp_cuda_copy_matrix: dt_size must be 4 or 8\n"); abort(); }
    const size_t row_bytes = dt_size * (size_t) ncol;
    copy_rows_kernel<<<copy_grid((size_t) nrow * (row_bytes / 4)), 256, 0, as_stream(stream)>>>(
        (const char *) src, dt_size * (size_t) lds, (char *) dst, dt_size * (size_t) ldd, (uint32_t) nrow, (uint32_t) row_bytes, NULL);
    CRP_LAUNCH_CHECK();
}

extern "C" void crp_cuda_copy_matrix(size_t dt_size, const int nrow, const int ncol, const void *src, const int lds, void *dst, const int ldd)
{
    crp_cuda_copy_matrix_async(dt_size, nrow, ncol, src, lds, dst, ldd, NULL);
    CRP_CUDA_CHECK(cudaStreamSynchronize(0));
}

extern "C" void crp_cuda_gather_rows(size_t dt_size, const int nrow, const int ncol, const void *src, const int lds, const int *ridx_d, void *dst, const int ldd, void *stream)
{
    if (nrow <= 0 || ncol <= 0) return;
    if (dt_size != 4 && dt_size != 8) { fprintf(stderr, "[FATAL] crp_cuda_gather_rows: dt_size must be 4 or 8\n"); abort(); }
    const size_t row_bytes = dt_size * (size_t) ncol;
    copy_rows_kernel<<<copy_grid((size_t) nrow * (row_bytes / 4)), 256, 0, as_stream(stream)>>>(
        (const char *) src, dt_size * (size_t) lds, (char *) dst, dt_size * (size_t) ldd, (uint32_t) nrow, (uint32_t) row_bytes, ridx_d);
    CRP_LAUNCH_CHECK();
}

// blockIdx.y = block descriptor; blockIdx.x strides over that block's bytes
__global__ void __launch_bounds__(256) copy_blocks_kernel(const crp_copy_block *__restrict__ blocks, const char *__restrict__ src_base, char *__restrict__ dst_base)
{
    const crp_copy_block b = blocks[blockIdx.y];
    const char *src = src_base + b.src_off;
    char *dst = dst_base + b.dst_off;
    const size_t tid = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nth = (size_t) gridDim.x * blockDim.x;
    const int al = common_alignment((uintptr_t) src, (uintptr_t) dst, b.src_pitch, b.dst_pitch, b.row_bytes);
    if (al == 16)     copy_rows_body<uint4>(src, b.src_pitch, dst, b.dst_pitch, b.nrow, b.row_bytes, NULL, tid, nth);
    else if (al == 8) copy_rows_body<uint2>(src, b.src_pitch, dst, b.dst_pitch, b.nrow, b.row_bytes, NULL, tid, nth);
    else              copy_rows_body<uint32_t>(src, b.src_pitch, dst, b.dst_pitch, b.nrow, b.row_bytes, NULL, tid, nth);
}

extern "C" void crp_cuda_copy_blocks(const crp_copy_block *blocks_d, const int nblk, const void *src_base, void *dst_base, void *stream)
{
    if (nblk <= 0) return;
    if (nblk > 65535) { fprintf(stderr, "[FATAL] crp_cuda_copy_blocks: too many blocks (%d)\n", nblk); abort(); }
    int gx = (148 * 8 + nblk - 1) / nblk;         // about 8 CTAs per SM in total
    if (gx < 1) gx = 1;
    dim3 grid((unsigned) gx, (unsigned) nblk, 1);
    copy_blocks_kernel<<<grid, 256, 0, as_stream(stream)>>>(blocks_d, (const char *) src_base, (char *) dst_base);
    CRP_LAUNCH_CHECK();
}

// ------------------------------------------------------- peer-memory exchange
extern "C" int crp_cuda_ipc_get_handle(void *dptr, void *handle64)
{
    static_assert(sizeof(cudaIpcMemHandle_t) == CRP_IPC_HANDLE_BYTES, "IPC handle size");
    cudaError_t err = cudaIpcGetMemHandle((cudaIpcMemHandle_t *) handle64, dptr);
    if (err != cudaSuccess) { (void) cudaGetLastError(); return 0; }
    return 1;
}

extern "C" void *crp_cuda_ipc_open(const void *handle64)
{
    void *p = NULL;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    cudaError_t err = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (err != cudaSuccess) { (void) cudaGetLastError(); return NULL; }
    return p;
}

extern "C" void crp_cuda_ipc_close(void *peer_ptr)
{
    if (peer_ptr && cudaIpcCloseMemHandle(peer_ptr) != cudaSuccess) (void) cudaGetLastError();
}

// Gather rows of the local B and store them straight into the receive buffers of the ranks that need them:
// the "pack" and the "send" of the reference in one kernel, NVLink stores issued by the SMs (128-bit when aligned).
template <typename V>
__global__ void __launch_bounds__(256) put_rows_kernel(
    const char *__restrict__ src, const size_t src_pitch, const int *__restrict__ ridx, char *const *__restrict__ dst_rows,
    const uint32_t nrow, const uint32_t row_bytes
)
{
    const uint32_t vpr = row_bytes / (uint32_t) sizeof(V);
    const size_t total = (size_t) nrow * vpr;
    for (size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t) gridDim.x * blockDim.x)
    {
        const uint32_t r = (uint32_t) (t / vpr), v = (uint32_t) (t - (size_t) r * vpr);
        const V val = *reinterpret_cast<const V *>(src + (size_t) ridx[r] * src_pitch + (size_t) v * sizeof(V));
        *reinterpret_cast<V *>(dst_rows[r] + (size_t) v * sizeof(V)) = val;
    }
}

extern "C" void crp_cuda_put_rows(size_t dt_size, const int nrow, const int ncol, const void *src, const int lds, const int *ridx_d, void *const *dst_rows_d, void *stream)
{
    if (nrow <= 0 || ncol <= 0) return;
    const size_t row_bytes = dt_size * (size_t) ncol, pitch = dt_size * (size_t) lds;
    const int grid = copy_grid((size_t) nrow * (row_bytes / 4));
    // destination rows are row_bytes apart inside 256-byte aligned buffers: 16-byte stores need row_bytes % 16 == 0
    if (((uintptr_t) src & 15) == 0 && pitch % 16 == 0 && row_bytes % 16 == 0)
        put_rows_kernel<uint4><<<grid, 256, 0, as_stream(stream)>>>((const char *) src, pitch, ridx_d, (char *const *) dst_rows_d, (uint32_t) nrow, (uint32_t) row_bytes);
    else if (((uintptr_t) src & 7) == 0 && pitch % 8 == 0 && row_bytes % 8 == 0)
        put_rows_kernel<uint2><<<grid, 256, 0, as_stream(stream)>>>((const char *) src, pitch, ridx_d, (char *const *) dst_rows_d, (uint32_t) nrow, (uint32_t) row_bytes);
    else
        put_rows_kernel<uint32_t><<<grid, 256, 0, as_stream(stream)>>>((const char *) src, pitch, ridx_d, (char *const *) dst_rows_d, (uint32_t) nrow, (uint32_t) row_bytes);
    CRP_LAUNCH_CHECK();
}

// put + signal in one launch: the gather / NVLink stores of put_rows_kernel, then the LAST block to finish publishes this rank's
// arrival flag on every neighbour (all blocks fence their stores system-wide before they count themselves done).
template <typename V>
__global__ void __launch_bounds__(256) put_rows_signal_kernel(
    const char *__restrict__ src, const size_t src_pitch, const int *__restrict__ ridx, char *const *__restrict__ dst_rows,
    const uint32_t nrow, const uint32_t row_bytes, unsigned int *const *__restrict__ flag_ptrs, const int nflag, const unsigned int epoch,
    unsigned int *__restrict__ done_counter, const size_t dst_off
)
{
    const uint32_t vpr = row_bytes / (uint32_t) sizeof(V);
    const size_t total = (size_t) nrow * vpr;
    for (size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t) gridDim.x * blockDim.x)
    {
        const uint32_t r = (uint32_t) (t / vpr), v = (uint32_t) (t - (size_t) r * vpr);
        const V val = *reinterpret_cast<const V *>(src + (size_t) ridx[r] * src_pitch + (size_t) v * sizeof(V));
        *reinterpret_cast<V *>(dst_rows[r] + dst_off + (size_t) v * sizeof(V)) = val;
    }
    __threadfence_system();
    __syncthreads();
    __shared__ unsigned int s_last;
    if (threadIdx.x == 0) s_last = (atomicAdd(done_counter, 1u) == gridDim.x - 1u) ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();                            // the other blocks' stores (fenced before their atomicAdd) are ordered before the flags
    for (int j = threadIdx.x; j < nflag; j += blockDim.x)
        asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(flag_ptrs[j]), "r"(epoch) : "memory");
    if (threadIdx.x == 0) *done_counter = 0u;   // ready for the next launch on this stream
}

extern "C" void crp_cuda_put_rows_signal(
    size_t dt_size, const int nrow, const int ncol, const void *src, const int lds, const int *ridx_d, void *const *dst_rows_d,
    unsigned int *const *flag_ptrs_d, const int nflag, const unsigned int epoch, unsigned int *done_counter_d, const size_t dst_off_bytes, void *stream
)
{
    if (nflag <= 0 && (nrow <= 0 || ncol <= 0)) return;
    const size_t row_bytes = dt_size * (size_t) (ncol > 0 ? ncol : 0), pitch = dt_size * (size_t) lds;
    const uint32_t rows = (nrow > 0 && ncol > 0) ? (uint32_t) nrow : 0u;
    const int grid = rows ? copy_grid((size_t) rows * (row_bytes / 4)) : 1;
#define CRP_PUT(V) put_rows_signal_kernel<V><<<grid, 256, 0, as_stream(stream)>>>((const char *) src, pitch, ridx_d, (char *const *) dst_rows_d, rows, (uint32_t) row_bytes, flag_ptrs_d, nflag, epoch, done_counter_d, dst_off_bytes)
    if (((uintptr_t) src & 15) == 0 && pitch % 16 == 0 && row_bytes % 16 == 0 && dst_off_bytes % 16 == 0) CRP_PUT(uint4);
    else if (((uintptr_t) src & 7) == 0 && pitch % 8 == 0 && row_bytes % 8 == 0 && dst_off_bytes % 8 == 0) CRP_PUT(uint2);
    else CRP_PUT(uint32_t);
#undef CRP_PUT
    CRP_LAUNCH_CHECK();
}

__global__ void signal_peers_kernel(unsigned int *const *__restrict__ flag_ptrs, const int nflag, const unsigned int epoch)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nflag) return;
    __threadfence_system();                     // the puts of the previous kernel on this stream are visible system-wide first
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(flag_ptrs[j]), "r"(epoch) : "memory");
}

extern "C" void crp_cuda_signal_peers(unsigned int *const *flag_ptrs_d, const int nflag, const unsigned int epoch, void *stream)
{
    if (nflag <= 0) return;
    signal_peers_kernel<<<(nflag + 63) / 64, 64, 0, as_stream(stream)>>>(flag_ptrs_d, nflag, epoch);
    CRP_LAUNCH_CHECK();
}

// The waiting side never waits for a kernel of its own GPU: the flags are written by the peers' GPUs.
__global__ void wait_flags_kernel(const unsigned int *flags, const int *__restrict__ wait_idx, const int nwait, const unsigned int epoch, const long long timeout_ns, int *err)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nwait) return;
    const unsigned int *f = flags + wait_idx[j];
    long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;)
    {
        unsigned int v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
        if ((int) (v - epoch) >= 0) break;
        long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > timeout_ns) { *err = 1; break; }
        __nanosleep(200);
    }
}

extern "C" void crp_cuda_wait_flags(const unsigned int *flags_d, const int *wait_idx_d, const int nwait, const unsigned int epoch, const double timeout_s, int *err_d, void *stream)
{
    if (nwait <= 0) return;
    wait_flags_kernel<<<(nwait + 63) / 64, 64, 0, as_stream(stream)>>>(flags_d, wait_idx_d, nwait, epoch, (long long) (timeout_s * 1e9), err_d);
    CRP_LAUNCH_CHECK();
}

// 32 x 32 shared-memory tile transpose (+1 padding: conflict-free column reads)
template <typename T>
__global__ void __launch_bounds__(256) transpose_kernel(const T *__restrict__ src, size_t lds, T *__restrict__ dst, size_t ldd, int nrow, int ncol)
{
    __shared__ T tile[32][33];
    const int tiles_x = (ncol + 31) / 32;
    const long long ntiles = (long long) tiles_x * ((nrow + 31) / 32);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x)
    {
        const int r0 = (int) (t / tiles_x) * 32, c0 = (int) (t % tiles_x) * 32;
        #pragma unroll
        for (int j = 0; j < 32; j += 8)
        {
            const int r = r0 + ty + j, c = c0 + tx;
            if (r < nrow && c < ncol) tile[ty + j][tx] = src[(size_t) r * lds + c];
        }
        __syncthreads();
        #pragma unroll
        for (int j = 0; j < 32; j += 8)
        {
            const int c = c0 + ty + j, r = r0 + tx;
            if (r < nrow && c < ncol) dst[(size_t) c * ldd + r] = tile[tx][ty + j];
        }
        __syncthreads();
    }
}

extern "C" void crp_cuda_transpose(size_t dt_size, const int nrow, const int ncol, const void *src, const int lds, void *dst, const int ldd, void *stream)
{
    if (nrow <= 0 || ncol <= 0) return;
    const long long ntiles = (long long) ((ncol + 31) / 32) * ((nrow + 31) / 32);
    long long blocks = ntiles < 148LL * 16 ? ntiles : 148LL * 16;
    if (dt_size == 8)
        transpose_kernel<double><<<(unsigned) blocks, 256, 0, as_stream(stream)>>>((const double *) src, (size_t) lds, (double *) dst, (size_t) ldd, nrow, ncol);
    else if (dt_size == 4)
        transpose_kernel<float><<<(unsigned) blocks, 256, 0, as_stream(stream)>>>((const float *) src, (size_t) lds, (float *) dst, (size_t) ldd, nrow, ncol);
    else { fprintf(stderr, "[FATAL] crp_cuda_transpose: dt_size must be 4 or 8\n"); abort(); }
    CRP_LAUNCH_CHECK();
}

// ------------------------------------------------------------- fp64 FMA peak (measurement aid)
// SURVEY.md §8(d) asks for the fp64 roofline next to the HBM one because the headline configuration sits on the
// ridge (5.8 flop / byte).  8 independent FMA chains per thread, 16 warps per scheduler: pure DFMA issue.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, const int iters, const double b, const double c)
{
    double a[8];
    #pragma unroll
    for (int j = 0; j < 8; j++) a[j] = (double) (threadIdx.x + j);
    for (int i = 0; i < iters; i++)
    {
        #pragma unroll
        for (int j = 0; j < 8; j++) a[j] = fma(a[j], b, c);
    }
    double sum = 0.0;
    #pragma unroll
    for (int j = 0; j < 8; j++) sum += a[j];
    out[(size_t) blockIdx.x * blockDim.x + threadIdx.x] = sum;
}

extern "C" double crp_cuda_measure_dfma_tflops(void)
{
    const int nsm = crp_cuda_sm_count(), blocks = nsm * 8, iters = 4096;
    double *out = NULL;
    CRP_CUDA_CHECK(cudaMalloc((void **) &out, sizeof(double) * (size_t) blocks * 256));
    cudaEvent_t e0, e1;
    CRP_CUDA_CHECK(cudaEventCreate(&e0));
    CRP_CUDA_CHECK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++)
    {
        CRP_CUDA_CHECK(cudaEventRecord(e0, 0));
        dfma_peak_kernel<<<blocks, 256>>>(out, iters, 0.999999, 1e-9);
        CRP_LAUNCH_CHECK();
        CRP_CUDA_CHECK(cudaEventRecord(e1, 0));
        CRP_CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CRP_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    CRP_CUDA_CHECK(cudaEventDestroy(e0));
    CRP_CUDA_CHECK(cudaEventDestroy(e1));
    CRP_CUDA_CHECK(cudaFree(out));
    const double flops = 2.0 * 8.0 * (double) iters * 256.0 * (double) blocks;
    return flops / ((double) best * 1e-3) / 1e12;
}
