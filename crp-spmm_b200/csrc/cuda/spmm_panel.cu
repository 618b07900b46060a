// spmm_panel.cu - row-group CSR x dense kernel with the B rows of a tile staged ONCE per thread block
// in shared memory by the TMA engine (cp.async.bulk + mbarrier pipeline), warp-specialised.
//
// Replaces the mkl_sparse_d_mm call at reference src/rowpara_spmm.c:404-407 for matrices with
// row-group structure (FEM / multi-dof), like spmm_rowgroup.cu, and removes what limited that kernel
// (round-1 ncu: 3.0x the algorithmic bytes moved L2 -> L1 because every warp fetched "its" B rows
// itself, 12 warps / SM waiting on those gathers, fp64 pipe 45 % busy):
//
//   * a TILE of K consecutive row groups is owned by one persistent thread block; the union of the
//     groups' column lists - the tile's B row panel - is copied global -> shared exactly once, one
//     cp.async.bulk per B row slice (UBLKCP), by a PRODUCER warp that runs NSTAGE - 1 chunks ahead of
//     the math (mbarrier full / empty ring, expect_tx byte counts: SYNCS.ARRIVE.TRANS64);
//   * the chunk's block values / row slots arrive in the same stage as one contiguous record
//     (panel_build.hpp), so the K CONSUMER warps touch global memory only to store C: their inner
//     loop is LDS (29 cycles) + R * U * VEC FMAs per block, no long-scoreboard dependency at all;
//   * tiles are dealt round-robin to the blocks, so at any time the whole grid works on a contiguous
//     window of rows and the panel rows shared between neighbouring tiles are L2 hits;
//   * (multi-GPU) the producer is the only warp that touches B: it can wait for a neighbour's arrival
//     flag right before the first chunk that needs a received row, so the product of the own rows
//     overlaps the exchange inside one kernel (see crp_cuda_spmm_exec_wait).
//
// Bound: HBM for A (values travel inside the meta records), first touch of B and C; shared-memory
// bandwidth (fill + R-fold reuse reads) and fp64 issue inside the SM.  DESIGN.md has the numbers.
#include <cstring>
#include <vector>
#include <unistd.h>

#include "crp_cuda_internal.cuh"
#include "panel_build.hpp"

// ---------------------------------------------------------------------------------- PTX helpers
namespace {

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, const unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, const unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, const unsigned parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy (TMA, 1-D), completion counted in bytes on `bar`; size and both addresses multiples of 16
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, const unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned lds_u32(const unsigned addr)
{
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

template <typename T, int VEC> struct pvec;
template <> struct pvec<double, 2>
{
    static __device__ __forceinline__ void lds_s(const unsigned addr, double (&v)[2]) { asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "r"(addr)); }
    static __device__ __forceinline__ void lds(const double *p, double (&v)[2]) { const double2 t = *reinterpret_cast<const double2 *>(p); v[0] = t.x; v[1] = t.y; }
    static __device__ __forceinline__ void ldc(const double *p, double (&v)[2]) { const double2 t = *reinterpret_cast<const double2 *>(p); v[0] = t.x; v[1] = t.y; }
    static __device__ __forceinline__ void ldg(const double *p, double (&v)[2]) { const double2 t = __ldg(reinterpret_cast<const double2 *>(p)); v[0] = t.x; v[1] = t.y; }
    static __device__ __forceinline__ void st(double *p, const double (&v)[2]) { __stcs(reinterpret_cast<double2 *>(p), make_double2(v[0], v[1])); }
};
template <> struct pvec<float, 4>
{
    static __device__ __forceinline__ void lds_s(const unsigned addr, float (&v)[4]) { asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(addr)); }
    static __device__ __forceinline__ void lds(const float *p, float (&v)[4]) { const float4 t = *reinterpret_cast<const float4 *>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    static __device__ __forceinline__ void ldc(const float *p, float (&v)[4]) { const float4 t = *reinterpret_cast<const float4 *>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    static __device__ __forceinline__ void ldg(const float *p, float (&v)[4]) { const float4 t = __ldg(reinterpret_cast<const float4 *>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    static __device__ __forceinline__ void st(float *p, const float (&v)[4]) { __stcs(reinterpret_cast<float4 *>(p), make_float4(v[0], v[1], v[2], v[3])); }
};

// the R values of one entry from shared memory (broadcast reads), widest access the entry stride allows
template <typename T, int R>
__device__ __forceinline__ void lds_vals(const T *p, T (&a)[R])
{
    constexpr int BYTES = R * (int) sizeof(T);
    if constexpr (BYTES % 16 == 0)
    {
        #pragma unroll
        for (int i = 0; i < BYTES / 16; i++)
        {
            const uint4 t = reinterpret_cast<const uint4 *>(p)[i];
            memcpy(reinterpret_cast<char *>(a) + 16 * i, &t, 16);
        }
    } else if constexpr (BYTES % 8 == 0) {
        #pragma unroll
        for (int i = 0; i < BYTES / 8; i++)
        {
            const uint2 t = reinterpret_cast<const uint2 *>(p)[i];
            memcpy(reinterpret_cast<char *>(a) + 8 * i, &t, 8);
        }
    } else {
        #pragma unroll
        for (int i = 0; i < R; i++) a[i] = p[i];
    }
}

// the R values of one entry at a 32-bit shared address
template <typename T, int R>
__device__ __forceinline__ void lds_vals_s(const unsigned addr, T (&a)[R])
{
    constexpr int BYTES = R * (int) sizeof(T);
    if constexpr (BYTES % 16 == 0)
    {
        #pragma unroll
        for (int i = 0; i < BYTES / 16; i++)
        {
            uint4 t;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "r"(addr + 16u * i));
            memcpy(reinterpret_cast<char *>(a) + 16 * i, &t, 16);
        }
    } else if constexpr (BYTES % 8 == 0) {
        #pragma unroll
        for (int i = 0; i < BYTES / 8; i++)
        {
            uint2 t;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(t.x), "=r"(t.y) : "r"(addr + 8u * i));
            memcpy(reinterpret_cast<char *>(a) + 8 * i, &t, 8);
        }
    } else {
        #pragma unroll
        for (int i = 0; i < BYTES / 4; i++)
        {
            unsigned t = lds_u32(addr + 4u * i);
            memcpy(reinterpret_cast<char *>(a) + 4 * i, &t, 4);
        }
    }
}

}   // namespace

enum { CRP_PANEL_MAXSTAGE = 8, CRP_PANEL_BAR_BYTES = 128 };

__device__ __forceinline__ long long gtime_ns()
{
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Per-block time stamps are a development aid compiled in only with -DCRP_PANEL_TRACE_BUILD (make CRP_NVCC_EXTRA=-DCRP_PANEL_TRACE_BUILD):
// the hooks cost two instructions in the consumers' entry loop through register allocation (141 instead of 139), about 2 % of the
// headline kernel.  The traces under profiles/r02_trace_*.txt were taken with such a build.
#ifdef CRP_PANEL_TRACE_BUILD
#define CRP_TRACE_ON(a) ((a).trace != NULL)
#else
#define CRP_TRACE_ON(a) false
#endif
#define CRP_TRACE(slot) do { if (CRP_TRACE_ON(a) && lane == 0) a.trace[(size_t) (blockIdx.y * gridDim.x + blockIdx.x) * 8 + (slot)] = gtime_ns(); } while (0)

template <typename T>
struct panel_args
{
    const int2 *tile_ranges;            // ntiles: chunk range of the tile at each position of the processing order
    const crp_panel_chunk *chunks;      // nchunks + 1 (stop record last)
    const int *ucol;
    const unsigned char *meta;
    int ntiles, nchunks;
    int CR, nstage;
    unsigned meta_max;                  // bytes reserved per stage for a meta record
    const char *X0;  size_t ldx0;       // bytes
    const char *X1;  size_t ldx1;
    int x0_rows;
    int n;                              // dense columns
    T alpha, beta;
    T *C;  size_t ldc;                  // elements
    // multi-GPU: chunks that read received rows wait for the arrival flags of the ranks that sent them
    const unsigned *chunk_need;         // per chunk: bit j = needs wait slot j (NULL: nothing to wait for)
    const unsigned *flags;              // arrival flags (one 32-bit word per rank)
    const int *wait_idx;                // wait slot -> flag index
    // rows that are in no group ("rest" rows: the cut nodes at a rank's first / last row, irregular rows): multiplied inside
    // this kernel by one extra warp per block straight from the CSR arrays, so that a handful of rows does not cost a launch
    int nrest;                          // 0: none (or the caller runs the row-split kernel for them)
    const int *rest_rows;
    const int *rowptr;  const int *colidx;  const T *val;
    // multi-GPU, fused exchange: every block first stores its share of the B rows the neighbours need straight into their
    // receive buffers (NVLink), the last block to finish publishes this rank's arrival flag on every neighbour
    int put_nrow;                       // 0: nothing to put
    unsigned put_row_bytes;             // multiple of 16
    const int *put_ridx;                // rows of X0 to send
    char *const *put_dst_rows;          // destination of each (peer memory)
    unsigned int *const *put_flag_ptrs; // this rank's flag on each neighbour
    int put_nflag;
    unsigned int *put_counter;          // blocks that have finished their share (reset by the last one)
    size_t put_dst_off;                 // bytes added to every destination
    int nwait;
    int wait_all_first;                 // no wait map for this neighbour list: wait for everybody before the first chunk
    unsigned epoch;
    long long timeout_ns;
    int *err;
    long long *trace;                   // development aid (CRP_PANEL_TRACE): 8 time stamps per block, NULL in production
};

// FAST: every column slice is full (n is a multiple of the slice width) and all groups are exact - no bounds predicates,
// no mask tests in the inner loop.  The other instantiation handles partial slices and masked (relaxed-group) entries.
// Warp layout of a block: NP producer warps (chunk i of the block's stream is issued by producer i % NP - one warp needs
// ~1000 cycles of dependent instructions per chunk, which bounded the pipeline before), K consumer warps (one row group of
// the tile each; 168 registers: the register file of a scheduler holds three such warps), one warp for the rows that are in
// no group.  Measured and dropped in round 2 (profiles/r02_kernel_sweep.md): K = 11 / 12 consumers (with setmaxnreg register
// rebalancing) and two half-width warps per group - more warps per scheduler did not help, the shared-memory pipe is the limit.
template <int K> struct panel_layout
{
    static constexpr int NP = 2;                            // producer warps = warps before the first consumer
    static constexpr int THREADS = (NP + K + 1) * 32;
};

template <typename T, int VEC, int R, int U, int K, bool FAST>
__global__ void __launch_bounds__(panel_layout<K>::THREADS, 1) spmm_panel_kernel(const panel_args<T> a)
{
    constexpr int PW = panel_layout<K>::NP;
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int W = 32 * U * VEC;                         // dense columns per block
    constexpr int RBW = W * (int) sizeof(T);                // bytes per staged row slice
    constexpr int HDR = ((3 + 2 * K + 3) / 4) * 16;
    constexpr unsigned FULLMASK = (1u << R) - 1u;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem);
    uint64_t *empty = full + CRP_PANEL_MAXSTAGE;
    unsigned char *rows = smem + CRP_PANEL_BAR_BYTES;
    unsigned char *metas = rows + (size_t) a.nstage * a.CR * RBW;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nstage = a.nstage;
    const int col0 = blockIdx.y * W;
    const int ncols = min(a.n - col0, W);

    if (warp == 0) CRP_TRACE(0);                            // block start
    if (threadIdx.x == 0)
    {
        for (int s = 0; s < nstage; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], K); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();                                        // barriers initialised
    if (a.put_nflag > 0 && warp >= PW)
    {
        // Fused "pack + send" (reference src/rowpara_spmm.c:232-301): 128-bit loads of the own B rows, 128-bit stores over NVLink.
        // The producers never take part - they start their descriptor chain at once.  A small share (the usual case: a few KB
        // per block) is put by the REST warp alone, four independent 16-byte pieces per lane in flight, so that the consumers
        // start multiplying as soon as the first chunk has landed; only a large share is spread over the consumer warps too,
        // which then meet at a named barrier behind the system fence.  (Round-2 traces, profiles/r02_trace_n2_*.txt: with all
        // warps putting behind a block-wide barrier the first chunk was consumed 15 us after the block started, with the
        // consumers putting 9.7 us, the producers alone need 6.)
        constexpr unsigned NPUT = (K + 1) * 32;
        const unsigned nblocks = gridDim.x * gridDim.y, bid = blockIdx.y * gridDim.x + blockIdx.x;
        const unsigned vpr = a.put_row_bytes / 16u;
        const size_t total = (size_t) a.put_nrow * vpr;
        // more than 2 KB per block: spread over the consumer warps too.  (N = 8 trace, profiles/r02_trace_n8.txt: the rest warp alone
        // needed 30 - 42 us for 11.6 KB per block, so the flags reached the neighbours after their consumers had finished, and their
        // rest warps - which wait for every flag before they multiply the rows at the slice boundaries - ended the kernel late.)
        const bool by_all = total > (size_t) nblocks * 128;
        auto piece_src = [&](const size_t t) -> const uint4 * {
            const unsigned r = (unsigned) (t / vpr), v = (unsigned) (t - (size_t) r * vpr);
            return reinterpret_cast<const uint4 *>(a.X0 + (size_t) a.put_ridx[r] * a.ldx0 + (size_t) v * 16);
        };
        auto piece_dst = [&](const size_t t) -> uint4 * {
            const unsigned r = (unsigned) (t / vpr), v = (unsigned) (t - (size_t) r * vpr);
            return reinterpret_cast<uint4 *>(a.put_dst_rows[r] + a.put_dst_off + (size_t) v * 16);
        };
        if (by_all || warp == PW + K)
        {
            const unsigned nthr = by_all ? NPUT : 32u;
            const unsigned tid = by_all ? threadIdx.x - PW * 32 : (unsigned) lane;
            const size_t stride = (size_t) nblocks * nthr;
            for (size_t t = (size_t) bid * nthr + tid; t < total; t += 4 * stride)
            {
                uint4 v4[4];
                #pragma unroll
                for (int j = 0; j < 4; j++) if (t + j * stride < total) v4[j] = *piece_src(t + j * stride);
                #pragma unroll
                for (int j = 0; j < 4; j++) if (t + j * stride < total) *piece_dst(t + j * stride) = v4[j];
            }
            __threadfence_system();
            if (by_all) asm volatile("bar.sync 1, %0;" :: "r"(NPUT) : "memory");        // the putting warps only
            else __syncwarp();
            if (tid == 0)
            {
                if (atomicAdd(a.put_counter, 1u) == nblocks - 1u)
                {
                    __threadfence();            // the other blocks' stores (fenced before their atomicAdd) come before the flags
                    for (int j = 0; j < a.put_nflag; j++)
                        asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(a.put_flag_ptrs[j]), "r"(a.epoch) : "memory");
                    *a.put_counter = 0u;
                }
                if (CRP_TRACE_ON(a)) a.trace[(size_t) bid * 8 + 1] = gtime_ns();        // this block's share of the put is done (and counted)
            }
        }
    }

    if (warp < PW)
    {
        constexpr int NP = panel_layout<K>::NP;
        // ------------------------------------------------------------------ producer
        const unsigned rb = (unsigned) ncols * (unsigned) sizeof(T);
        const char *x0 = a.X0 + (size_t) col0 * sizeof(T);
        const char *x1 = a.X1 + (size_t) col0 * sizeof(T);
        unsigned seen = 0;                                  // wait slots whose flag has been observed
        int s = warp;                                       // this producer issues stream positions warp, warp + NP, ...
        unsigned ph = 1;                                    // parity the empty barrier of stage s is waited on
        while (s >= nstage) { s -= nstage; ph ^= 1u; }

        // The chunk stream of this block: the chunks of tiles blockIdx.x, blockIdx.x + gridDim.x, ... in order, then the stop
        // record.  Descriptors are fetched TWO chunks ahead and the panel's column ids ONE chunk ahead of the chunk being
        // issued, so that the dependent global loads (tile range -> descriptor -> column ids -> row address) are off the
        // critical path: while the producer sleeps on an empty barrier they complete.  (Round-2 ncu of the first version, which
        // loaded them in the iteration that used them: consumers spent 29 % of their samples waiting for the full barrier.)
        int it_t = blockIdx.x, it_c = -1, it_cend = -1, nx_c = -1, nx_cend = -1;
        if (it_t < a.ntiles) { const int2 tr = __ldg(a.tile_ranges + it_t); it_c = tr.x; it_cend = tr.y; }
        if (it_t + (int) gridDim.x < a.ntiles) { const int2 tr = __ldg(a.tile_ranges + it_t + gridDim.x); nx_c = tr.x; nx_cend = tr.y; }
        bool stop_given = false;
        auto next_chunk = [&]() -> int {                    // next chunk id of the stream; a.nchunks = stop record; -1 = past the end
            if (it_c < 0)
            {
                if (stop_given) return -1;
                stop_given = true;
                return a.nchunks;
            }
            const int id = it_c;
            if (++it_c >= it_cend)
            {
                it_t += gridDim.x;
                if (it_t < a.ntiles)
                {
                    it_c = nx_c;  it_cend = nx_cend;
                    if (it_t + (int) gridDim.x < a.ntiles) { const int2 tr = __ldg(a.tile_ranges + it_t + gridDim.x); nx_c = tr.x; nx_cend = tr.y; }
                } else it_c = -1;
            }
            return id;
        };
        const int4 *descs = reinterpret_cast<const int4 *>(a.chunks);
        auto next_mine = [&]() -> int {                     // the next stream position that belongs to this producer
            int id = next_chunk();
            #pragma unroll
            for (int j = 1; j < NP; j++) next_chunk();
            return id;
        };
        for (int j = 0; j < warp; j++) next_chunk();        // skip to this producer's first position
        int id0 = next_mine(), id1 = next_mine(), id2 = next_mine(), id3 = next_mine();
        if (id0 < 0) return;
        const int4 zero4 = make_int4(0, 0, 0, 0);
        int4 d0 = __ldg(descs + id0);
        int4 d1 = (id1 >= 0) ? __ldg(descs + id1) : zero4;
        int4 d2 = (id2 >= 0) ? __ldg(descs + id2) : zero4;
        int col0 = (lane < d0.y) ? __ldg(a.ucol + d0.x + lane) : 0;
        int col1 = (id1 >= 0 && lane < d1.y) ? __ldg(a.ucol + d1.x + lane) : 0;
        // rows [base, base + 32) of a chunk: one operation per run of rows that are consecutive in B and in memory
        auto for_runs = [&](const int nrows_, const int base, const int mycol, auto &&op) {
            const int r = base + lane;
            const bool active = r < nrows_;
            const int prev = __shfl_up_sync(0xffffffffu, mycol, 1);
            const bool in0 = mycol < a.x0_rows;
            const bool contig = (rb == (unsigned) RBW) && ((in0 ? a.ldx0 : a.ldx1) == (size_t) rb);
            const bool head = active && (lane == 0 || !contig || mycol != prev + 1 || mycol == a.x0_rows);
            const unsigned H = __ballot_sync(0xffffffffu, head);
            const unsigned A = __ballot_sync(0xffffffffu, active);
            if (head)
            {
                const unsigned above = (lane == 31) ? 0u : ((H >> (lane + 1)) << (lane + 1));
                const int end = above ? (__ffs(above) - 1) : __popc(A);
                const char *src = in0 ? x0 + (size_t) mycol * a.ldx0 : x1 + (size_t) (mycol - a.x0_rows) * a.ldx1;
                op(r, src, (unsigned) (end - lane) * rb);
            }
        };
        while (id0 >= 0)
        {
            const int4 d3 = (id3 >= 0) ? __ldg(descs + id3) : zero4;
            const int col2 = (id2 >= 0 && lane < d2.y) ? __ldg(a.ucol + d2.x + lane) : 0;      // column ids two turns ahead
            const int uo0 = d0.x, nrows = d0.y;
            const unsigned mo16 = (unsigned) d0.z, mbytes = (unsigned) d0.w * 16u;
            const bool is_stop = (id0 == a.nchunks);
            if (a.nwait > 0)
            {
                // chunks that read received rows wait for the ranks that send them; the stop record waits for every neighbour
                // (a rank that has seen all flags of epoch e knows that nobody reads the buffer half epoch e + 1 overwrites)
                unsigned need = 0;
                if (is_stop || a.wait_all_first) need = (a.nwait >= 32) ? 0xffffffffu : ((1u << a.nwait) - 1u);
                else if (a.chunk_need != NULL) need = __ldg(a.chunk_need + id0);
                need &= ~seen;
                if (need)
                {
                    const long long tw0 = CRP_TRACE_ON(a) ? gtime_ns() : 0;
                    // one lane per missing neighbour spins on its arrival flag; the rows were written by the peer's stores
                    // before the flag (fence + st.release.sys on the sender)
                    if (lane < a.nwait && ((need >> lane) & 1u))
                    {
                        const unsigned *f = a.flags + a.wait_idx[lane];
                        long long t0;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
                        while ((int) (ld_acquire_sys(f) - a.epoch) < 0)
                        {
                            long long t1;
                            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                            if (t1 - t0 > a.timeout_ns) { *a.err = 1; break; }
                            __nanosleep(100);
                        }
                    }
                    __syncwarp();
                    asm volatile("fence.proxy.async;" ::: "memory");     // generic-proxy acquire -> async-proxy (TMA) reads
                    seen |= need;
                    if (warp == 0 && CRP_TRACE_ON(a) && lane == 0)
                    {
                        const size_t tb = (size_t) (blockIdx.y * gridDim.x + blockIdx.x) * 8;
                        a.trace[tb + 4] += gtime_ns() - tw0;                    // time this producer spent waiting for flags
                        if (a.trace[tb + 6] == 0) a.trace[tb + 6] = tw0;        // when it first had to wait
                    }
                }
            }
            mbar_wait(&empty[s], ph);
            if (warp == 0 && CRP_TRACE_ON(a) && lane == 0)
            {
                const size_t tb = (size_t) (blockIdx.y * gridDim.x + blockIdx.x) * 8;
                if (a.trace[tb + 2] == 0) a.trace[tb + 2] = gtime_ns();         // first chunk issued
                if (is_stop) a.trace[tb + 7] = gtime_ns();                      // stop record issued (all flags seen)
            }
            if (lane == 0) mbar_arrive_expect_tx(&full[s], (unsigned) nrows * rb + mbytes);
            __syncwarp();
            if (lane == 0) bulk_g2s(metas + (size_t) s * a.meta_max, a.meta + (size_t) mo16 * 16, mbytes, &full[s]);
            unsigned char *dst = rows + (size_t) s * a.CR * RBW;
            // One bulk copy per RUN of panel rows that are consecutive in B and in memory (leading dimension == slice width):
            // the panel of an FEM matrix is a few long runs (the B rows of neighbouring nodes), so a chunk needs 1 - 3 copies
            // instead of one per row.  Rows that are not contiguous in memory are copied one by one.
            for (int base = 0; base < nrows; base += 32)
            {
                const int mycol = (base == 0) ? col0 : ((base + lane < nrows) ? __ldg(a.ucol + uo0 + base + lane) : 0);
                for_runs(nrows, base, mycol, [&](const int r, const char *src, const unsigned bytes) { bulk_g2s(dst + (size_t) r * RBW, src, bytes, &full[s]); });
            }
            s += NP;
            while (s >= nstage) { s -= nstage; ph ^= 1u; }
            id0 = id1;  id1 = id2;  id2 = id3;  id3 = next_mine();
            d0 = d1;  d1 = d2;  d2 = d3;  col0 = col1;  col1 = col2;
        }
        return;
    }

    // ---------------------------------------------------------------------- rest rows
    if (warp == PW + K)
    {
        if (a.nrest <= 0) return;
        if (a.nwait > 0)
        {
            // a rest row may read received rows: all neighbours first (plain loads follow, no async-proxy fence needed)
            if (lane < a.nwait)
            {
                const unsigned *f = a.flags + a.wait_idx[lane];
                long long t0;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
                while ((int) (ld_acquire_sys(f) - a.epoch) < 0)
                {
                    long long t1;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                    if (t1 - t0 > a.timeout_ns) { *a.err = 1; break; }
                    __nanosleep(100);
                }
            }
            __syncwarp();
        }
        constexpr int UT = W / (32 * VEC);                  // this warp covers the block's whole column slice
        constexpr int NZ = (UT * VEC * (int) sizeof(T) <= 32) ? 8 : 4;     // nonzeros in flight (register budget: NZ * UT * VEC values)
        const T *X0 = reinterpret_cast<const T *>(a.X0) + col0 + lane * VEC;
        const T *X1 = reinterpret_cast<const T *>(a.X1) + col0 + lane * VEC;
        const size_t ld0 = a.ldx0 / sizeof(T), ld1 = a.ldx1 / sizeof(T);
        bool ok[UT];
        #pragma unroll
        for (int u = 0; u < UT; u++) ok[u] = (lane * VEC + u * 32 * VEC) < ncols;
        for (int i = blockIdx.x; i < a.nrest; i += gridDim.x)
        {
            const int row = __ldg(a.rest_rows + i);
            const int pb = __ldg(a.rowptr + row), pe = __ldg(a.rowptr + row + 1);
            T acc[UT][VEC];
            #pragma unroll
            for (int u = 0; u < UT; u++)
                #pragma unroll
                for (int q = 0; q < VEC; q++) acc[u][q] = (T) 0;
            for (int p = pb; p < pe; p += NZ)
            {
                int c[NZ];  T v[NZ];  T x[NZ][UT][VEC];
                #pragma unroll
                for (int j = 0; j < NZ; j++)
                {
                    const bool in = p + j < pe;
                    c[j] = in ? __ldg(a.colidx + p + j) : 0;
                    v[j] = in ? __ldg(a.val + p + j) : (T) 0;
                }
                #pragma unroll
                for (int j = 0; j < NZ; j++)
                {
                    const T *xr = (c[j] < a.x0_rows) ? X0 + (size_t) c[j] * ld0 : X1 + (size_t) (c[j] - a.x0_rows) * ld1;
                    #pragma unroll
                    for (int u = 0; u < UT; u++)
                    {
                        if (ok[u] && p + j < pe) pvec<T, VEC>::ldg(xr + u * 32 * VEC, x[j][u]);
                        else { for (int q = 0; q < VEC; q++) x[j][u][q] = (T) 0; }
                    }
                }
                #pragma unroll
                for (int j = 0; j < NZ; j++)
                    if (p + j < pe)
                    {
                        #pragma unroll
                        for (int u = 0; u < UT; u++)
                            #pragma unroll
                            for (int q = 0; q < VEC; q++) acc[u][q] = fma(v[j], x[j][u][q], acc[u][q]);
                    }
            }
            if (pb == pe && a.beta == (T) 1) continue;
            T *crow = a.C + (size_t) row * a.ldc + col0 + lane * VEC;
            #pragma unroll
            for (int u = 0; u < UT; u++)
            {
                if (!ok[u]) continue;
                T out[VEC];
                if (a.beta == (T) 0)
                {
                    #pragma unroll
                    for (int q = 0; q < VEC; q++) out[q] = a.alpha * acc[u][q];
                } else {
                    T old[VEC];
                    pvec<T, VEC>::ldc(crow + u * 32 * VEC, old);
                    #pragma unroll
                    for (int q = 0; q < VEC; q++) out[q] = fma(a.alpha, acc[u][q], a.beta * old[q]);
                }
                pvec<T, VEC>::st(crow + u * 32 * VEC, out);
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers
    // Everything a consumer reads in its inner loop is shared memory addressed with 32-bit addresses and immediate offsets:
    // per entry one LDS (slot word, fetched two entries ahead), U 128-bit LDS of the B row slice, the R values (broadcast),
    // R * U * VEC FMAs.  The entry lists are padded with four dummy entries (panel_build.hpp), so the software pipeline
    // never needs a bounds check before it prefetches.
    const int w = warp - PW;                                // group of the tile this warp works for
    constexpr int UB = 32 * VEC * (int) sizeof(T);          // bytes between a lane's consecutive column vectors
    const int cbase = lane * VEC;                           // this lane's first column inside the slice
    bool valid[U];
    #pragma unroll
    for (int u = 0; u < U; u++) valid[u] = FAST || (cbase + u * 32 * VEC < ncols);
    const unsigned rows_s = smem_u32(rows) + (unsigned) cbase * (unsigned) sizeof(T);
    const unsigned metas_s = smem_u32(metas);

    T acc[R][U][VEC];
    int row0 = -1;
    int s = 0;
    unsigned ph = 0;                                        // parity the full barrier of stage s is waited on
    for (;; s = (s + 1 == nstage) ? 0 : s + 1, ph ^= (s == 0) ? 1u : 0u)
    {
        mbar_wait(&full[s], ph);
        if (w == 0 && CRP_TRACE_ON(a) && lane == 0)
        {
            const size_t tb = (size_t) (blockIdx.y * gridDim.x + blockIdx.x) * 8;
            if (a.trace[tb + 3] == 0) a.trace[tb + 3] = gtime_ns();             // first chunk landed
            a.trace[tb + 5] = gtime_ns();                                       // last chunk (or the stop record) seen
        }
        const unsigned mrec = metas_s + (unsigned) s * a.meta_max;
        // all header words at once (one shared-memory round trip), then the stop test
        const int flags = (int) lds_u32(mrec + 4);
        const int e0 = (int) lds_u32(mrec + 4 * (2 + K + w)), e1 = (int) lds_u32(mrec + 4 * (3 + K + w)), ne = (int) lds_u32(mrec + 4 * (2 + 2 * K));
        const int row0_new = (int) lds_u32(mrec + 4 * (2 + w));
        if (flags & CRP_PANEL_STOP) break;
        if (flags & CRP_PANEL_FIRST)
        {
            row0 = row0_new;
            #pragma unroll
            for (int r = 0; r < R; r++)
                #pragma unroll
                for (int u = 0; u < U; u++)
                    #pragma unroll
                    for (int q = 0; q < VEC; q++) acc[r][u][q] = (T) 0;
        }
        const unsigned slot_a = mrec + HDR;
        const unsigned val_a = slot_a + ((((unsigned) ne + 4u) * 4u + 15u) & ~15u);
        const unsigned x_a = rows_s + (unsigned) s * (unsigned) a.CR * RBW;

        auto ldx = [&](const unsigned sm, T (&xv)[U][VEC]) {
            const unsigned xr = x_a + (sm & 0xffffu) * RBW;
            #pragma unroll
            for (int u = 0; u < U; u++)
            {
                if (FAST || valid[u]) pvec<T, VEC>::lds_s(xr + u * UB, xv[u]);
                else { for (int q = 0; q < VEC; q++) xv[u][q] = (T) 0; }
            }
        };
        auto lda = [&](const int e, T (&av)[R]) { lds_vals_s<T, R>(val_a + (unsigned) e * (R * (unsigned) sizeof(T)), av); };
        auto fm = [&](const unsigned sm, const T (&av)[R], const T (&xv)[U][VEC]) {
            if (FAST || (sm >> 16) == FULLMASK)
            {
                #pragma unroll
                for (int r = 0; r < R; r++)
                    #pragma unroll
                    for (int u = 0; u < U; u++)
                        #pragma unroll
                        for (int q = 0; q < VEC; q++) acc[r][u][q] = fma(av[r], xv[u][q], acc[r][u][q]);
            } else {
                // near-identical group: rows that do not have this column are skipped, never multiplied by a stored zero
                #pragma unroll
                for (int r = 0; r < R; r++)
                    if ((sm >> (16 + r)) & 1u)
                    {
                        #pragma unroll
                        for (int u = 0; u < U; u++)
                            #pragma unroll
                            for (int q = 0; q < VEC; q++) acc[r][u][q] = fma(av[r], xv[u][q], acc[r][u][q]);
                    }
            }
        };

        {
            // software pipeline over the entries: the loads of entry e + 1 are issued before the FMAs of entry e, the slot words
            // two entries ahead (measured on B200: 0.325 ms; the variant that refills a register set right after its own FMAs
            // - no rotation - was slower, 0.343 ms: ptxas sinks those loads below the other set's FMAs anyway)
            int e = e0;
            unsigned s0 = lds_u32(slot_a + 4u * (unsigned) e), s1 = lds_u32(slot_a + 4u * (unsigned) e + 4u);
            T aa[R], ab[R], xa[U][VEC], xb[U][VEC];
            ldx(s0, xa);
            lda(e, aa);
            while (e + 2 <= e1)
            {
                const unsigned s2 = lds_u32(slot_a + 4u * (unsigned) e + 8u);
                ldx(s1, xb);
                lda(e + 1, ab);
                fm(s0, aa, xa);
                const unsigned s3 = lds_u32(slot_a + 4u * (unsigned) e + 12u);
                ldx(s2, xa);
                lda(e + 2, aa);
                fm(s1, ab, xb);
                s0 = s2;  s1 = s3;  e += 2;
            }
            if (e < e1) fm(s0, aa, xa);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);              // this warp is done with stage s

        if ((flags & CRP_PANEL_LAST) && row0 >= 0)
        {
            #pragma unroll
            for (int r = 0; r < R; r++)
            {
                T *crow = a.C + (size_t) (row0 + r) * a.ldc + col0 + cbase;
                #pragma unroll
                for (int u = 0; u < U; u++)
                {
                    if (!(FAST || valid[u])) continue;
                    T out[VEC];
                    if (a.beta == (T) 0)
                    {
                        #pragma unroll
                        for (int q = 0; q < VEC; q++) out[q] = (a.alpha == (T) 1) ? acc[r][u][q] : a.alpha * acc[r][u][q];
                    } else {
                        T old[VEC];
                        pvec<T, VEC>::ldc(crow + u * 32 * VEC, old);
                        #pragma unroll
                        for (int q = 0; q < VEC; q++) out[q] = fma(a.alpha, acc[r][u][q], a.beta * old[q]);
                    }
                    pvec<T, VEC>::st(crow + u * 32 * VEC, out);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------ plan side (host)

static int panel_env_int(const char *name, const int dflt)
{
    const char *e = getenv(name);
    return (e && e[0]) ? atoi(e) : dflt;
}

static void *panel_upload(const void *src, const size_t bytes)
{
    void *d = NULL;
    if (bytes == 0) return d;
    CRP_CUDA_CHECK(cudaMalloc(&d, bytes));
    CRP_CUDA_CHECK(cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice));
    return d;
}

// Build the panel form for the plan's row groups (called once at plan creation; the meta records of a value type
// are made on the first exec with that type).  Parameters: CRP_PANEL_CR rows and
// CRP_PANEL_EMAX blocks per chunk; defaults sized so that three to four stages of 2 KB row slices fit into 227 KB.
void crp_panel_build(crp_spmm_plan *plan)
{
    crp_panel *pn = &plan->pn;
    memset(pn, 0, sizeof(*pn));
    const crp_rowgroup_host *rg = plan->rg_host;
    if (rg == NULL || rg->R < 2 || rg->g_row.empty() || plan->n_hint < 64) return;
    if (panel_env_int("CRP_SPMM_PANEL", 1) == 0) return;
    const int K = 8;        // groups per tile = consumer warps (two per scheduler)
    int CR = panel_env_int("CRP_PANEL_CR", 32);
    if (CR < 4) CR = 4;
    if (CR > 64) CR = 64;
    int EMAX = panel_env_int("CRP_PANEL_EMAX", 4 * CR);
    if (EMAX < K) EMAX = K;
    crp_panel_host *ph = new crp_panel_host();
    crp_panel_build_structure(*rg, K, CR, EMAX, ph);
    // Tiles as patches of K groups with the largest column overlap (crp_panel_cluster_tiles) instead of K consecutive groups:
    // kept only where the panels shrink by more than a third.  Measured on B200 (profiles/r02_kbench_cluster.log): the 27-point
    // stencil (staged rows -51 %) gains 13 % (12.1 -> 10.5 ms), the headline FEM matrix (-17 %) LOSES 7 % (0.320 -> 0.344 ms) -
    // its consecutive-node tiles give every consumer warp the same number of entries in every chunk, the 2-D patches do not,
    // and the chunks are consumed in lockstep.  CRP_PANEL_CLUSTER = 0 / 1 forces the choice.
    const int want_cluster = panel_env_int("CRP_PANEL_CLUSTER", -1);
    if (want_cluster != 0)
    {
        int ncols = 0;
        for (size_t b = 0; b < rg->b_col.size(); b++) if (rg->b_col[b] >= ncols) ncols = rg->b_col[b] + 1;
        std::vector<int> order;
        crp_panel_cluster_tiles(*rg, K, ncols, &order);
        crp_panel_host *pc = new crp_panel_host();
        crp_panel_build_structure(*rg, K, CR, EMAX, pc, &order);
        if (want_cluster == 1 || 3 * pc->ucol.size() < 2 * ph->ucol.size()) { delete ph; ph = pc; pn->clustered = 1; }
        else delete pc;
    }
    pn->host = ph;
    pn->K = K;  pn->CR = CR;  pn->EMAX = EMAX;  pn->R = rg->R;
    pn->ntiles = ph->ntiles;  pn->nchunks = ph->nchunks();  pn->union_rows = (long long) ph->ucol.size();
    std::vector<int> ranges((size_t) 2 * ph->ntiles);
    for (int t = 0; t < ph->ntiles; t++) { ranges[2 * t] = ph->tile_chunk_ptr[t]; ranges[2 * t + 1] = ph->tile_chunk_ptr[t + 1]; }
    pn->d_tile_chunk_ptr = (int *) panel_upload(ranges.data(), sizeof(int) * ranges.size());
    pn->d_ucol = (int *) panel_upload(ph->ucol.data(), sizeof(int) * ph->ucol.size());
}

template <typename T>
static void panel_make_meta(crp_spmm_plan *plan)
{
    crp_panel *pn = &plan->pn;
    const int slot = (sizeof(T) == 8) ? 0 : 1;
    if (pn->d_meta[slot] != NULL) return;
    crp_panel_host *ph = (crp_panel_host *) pn->host;
    std::vector<unsigned char> meta;
    crp_panel_fill_meta<T>(*plan->rg_host, ph, &meta);
    pn->d_meta[slot] = (unsigned char *) panel_upload(meta.data(), meta.size());
    pn->d_chunks[slot] = panel_upload(ph->chunks.data(), sizeof(crp_panel_chunk) * ph->chunks.size());
    pn->meta_bytes[slot] = meta.size();
}

void crp_panel_destroy(crp_spmm_plan *plan)
{
    crp_panel *pn = &plan->pn;
    if (pn->d_trace != NULL)
    {
        std::vector<long long> t((size_t) 8 * 1024);
        CRP_CUDA_CHECK(cudaDeviceSynchronize());
        CRP_CUDA_CHECK(cudaMemcpy(t.data(), pn->d_trace, sizeof(long long) * t.size(), cudaMemcpyDeviceToHost));
        char path[512];
        snprintf(path, sizeof(path), "%s.%d", getenv("CRP_PANEL_TRACE"), (int) getpid());
        FILE *f = fopen(path, "w");
        if (f != NULL)
        {
            fprintf(f, "# block start put_done first_issue first_landed flag_wait_ns last_seen first_wait_at stop_issued   (ns, relative to the earliest block start)\n");
            long long t0 = 0;
            for (size_t b = 0; b < 1024; b++) if (t[8 * b] != 0 && (t0 == 0 || t[8 * b] < t0)) t0 = t[8 * b];
            for (size_t b = 0; b < 1024; b++)
            {
                if (t[8 * b] == 0) continue;
                fprintf(f, "%zu", b);
                for (int j = 0; j < 8; j++) fprintf(f, " %lld", (j == 4 || t[8 * b + j] == 0) ? t[8 * b + j] : t[8 * b + j] - t0);
                fprintf(f, "\n");
            }
            fclose(f);
        }
    }
    if (pn->d_tile_chunk_ptr) CRP_CUDA_CHECK(cudaFree(pn->d_tile_chunk_ptr));
    if (pn->d_ucol) CRP_CUDA_CHECK(cudaFree(pn->d_ucol));
    if (pn->d_chunk_need) CRP_CUDA_CHECK(cudaFree(pn->d_chunk_need));
    for (int i = 0; i < 2; i++)
    {
        if (pn->d_meta[i]) CRP_CUDA_CHECK(cudaFree(pn->d_meta[i]));
        if (pn->d_chunks[i]) CRP_CUDA_CHECK(cudaFree(pn->d_chunks[i]));
    }
    delete (crp_panel_host *) pn->host;
    memset(pn, 0, sizeof(*pn));
}

// Per chunk, which neighbours' rows it reads: bit j set = a row received from the rank of wait slot j
// (recv_off[j] .. recv_off[j + 1] are that rank's rows in the receive buffer, i.e. virtual ids x0_rows + ...).
void crp_panel_set_wait_map(crp_spmm_plan *plan, const int nslot, const int *recv_off)
{
    crp_panel *pn = &plan->pn;
    if (pn->d_chunk_need) { CRP_CUDA_CHECK(cudaFree(pn->d_chunk_need)); pn->d_chunk_need = NULL; }
    pn->nslot = 0;
    if (pn->host == NULL || nslot <= 0 || nslot > 32) return;      // more than 32 neighbours: the caller keeps the separate wait kernel
    const crp_panel_host *ph = (const crp_panel_host *) pn->host;
    std::vector<unsigned> need((size_t) ph->nchunks() + 1, 0u);
    bool any = false;
    for (int c = 0; c < ph->nchunks(); c++)
    {
        unsigned m = 0;
        for (int r = 0; r < ph->chunks[c].nrows; r++)
        {
            const int v = ph->ucol[(size_t) ph->chunks[c].uo0 + r] - plan->x0_rows;
            if (v < 0) continue;
            int j = 0;
            while (j + 1 < nslot && v >= recv_off[j + 1]) j++;
            m |= 1u << j;
        }
        need[c] = m;
        any = any || m != 0;
    }
    if (any) pn->d_chunk_need = (unsigned *) panel_upload(need.data(), sizeof(unsigned) * need.size());
    pn->nslot = nslot;
    // processing order: tiles that only read the rank's own B rows first, tiles that read received rows last - the exchange
    // is then hidden behind the local part of the product inside the one kernel
    std::vector<int> ranges;
    ranges.reserve((size_t) 2 * ph->ntiles);
    for (int pass = 0; pass < 2; pass++)
        for (int t = 0; t < ph->ntiles; t++)
        {
            bool remote = false;
            for (int c = ph->tile_chunk_ptr[t]; c < ph->tile_chunk_ptr[t + 1]; c++) remote = remote || need[c] != 0;
            if ((int) remote == pass) { ranges.push_back(ph->tile_chunk_ptr[t]); ranges.push_back(ph->tile_chunk_ptr[t + 1]); }
        }
    CRP_CUDA_CHECK(cudaMemcpy(pn->d_tile_chunk_ptr, ranges.data(), sizeof(int) * ranges.size(), cudaMemcpyHostToDevice));
}

// ------------------------------------------------------------------------------------- launch

template <typename T, int VEC, int R, int U, int K, bool FAST>
static bool panel_launch_cfg(crp_spmm_plan *plan, const panel_args<T> &args0, cudaStream_t s)
{
    crp_panel *pn = &plan->pn;
    panel_args<T> args = args0;
    constexpr int W = 32 * U * VEC, RBW = W * (int) sizeof(T);
    const crp_panel_host *ph = (const crp_panel_host *) pn->host;
    args.meta_max = (unsigned) ph->meta_max(sizeof(T));
    const size_t stage = (size_t) pn->CR * RBW + args.meta_max;
    static int smem_max = 0, nsm = 0;          // one device per process
    if (nsm == 0)
    {
        int dev = 0;
        CRP_CUDA_CHECK(cudaGetDevice(&dev));
        CRP_CUDA_CHECK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        CRP_CUDA_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    }
    int nstage = (int) (((size_t) smem_max - CRP_PANEL_BAR_BYTES) / stage);
    if (nstage > CRP_PANEL_MAXSTAGE) nstage = CRP_PANEL_MAXSTAGE;
    const int want = panel_env_int("CRP_PANEL_STAGES", 0);
    if (want >= 2 && want < nstage) nstage = want;
    if (nstage < 2) return false;
    args.nstage = nstage;
    const size_t smem = CRP_PANEL_BAR_BYTES + (size_t) nstage * stage;
    auto kern = spmm_panel_kernel<T, VEC, R, U, K, FAST>;
    static bool attr_set = false;
    if (!attr_set) { CRP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max)); attr_set = true; }
    const int nslice = (args.n + W - 1) / W;
    int gx = nsm / nslice;
    const int gx_env = panel_env_int("CRP_PANEL_GRID", 0);
    if (gx_env > 0) gx = gx_env;
    if (gx < 1) gx = 1;
    if (gx > pn->ntiles) gx = pn->ntiles;
    kern<<<dim3((unsigned) gx, (unsigned) nslice), panel_layout<K>::THREADS, smem, s>>>(args);
    CRP_LAUNCH_CHECK();
    return true;
}

template <typename T, int VEC, int R>
static bool panel_launch_R(crp_spmm_plan *plan, const panel_args<T> &args, cudaStream_t s)
{
    const int nv = args.n / VEC;
    constexpr int UMAX = (R * VEC * (int) sizeof(T) <= 6 * 16) ? 4 : 2;      // accumulator tile <= 96 registers, as in spmm_rowgroup.cu
    const int U = (nv >= 128 && UMAX >= 4) ? 4 : (nv >= 64 ? 2 : 1);
    const bool fast = plan->rg.exact && (args.n % (32 * U * VEC) == 0);
#define CRP_PN(U_) (fast ? panel_launch_cfg<T, VEC, R, U_, 8, true>(plan, args, s) : panel_launch_cfg<T, VEC, R, U_, 8, false>(plan, args, s))
    if (U == 4) { if constexpr (UMAX >= 4) return CRP_PN(4); else return false; }
    if (U == 2) return CRP_PN(2);
    return CRP_PN(1);
#undef CRP_PN
}

// false: the panel form does not apply to this call (no panel, misaligned operands, narrow n) - the caller falls back
// to the row-group kernel.  wait != NULL: the kernel itself waits for the neighbours' arrival flags (peer-memory transport).
// *rest_done (out): the kernel also multiplied the rows that are in no group (true) or the caller still has to (false)
template <typename T, int VEC>
bool crp_launch_panel(
    crp_spmm_plan *plan, const T *val, const int n, const T *X0, size_t ldx0, const T *X1, size_t ldx1, T alpha, T beta, T *C, size_t ldc,
    const crp_spmm_wait *wait, const crp_spmm_put *put, bool *rest_done, cudaStream_t s
)
{
    *rest_done = false;
    crp_panel *pn = &plan->pn;
    if (pn->host == NULL || pn->ntiles == 0) return false;
    if (put != NULL && (put->row_bytes % 16 != 0 || put->dst_off % 16 != 0 || ((uintptr_t) X0 & 15) != 0)) return false;       // the fused put moves 16-byte units
    const size_t es = sizeof(T);
    if (n < 64 || (n * es) % 16 != 0 || (ldx0 * es) % 16 != 0 || (X1 != NULL && (ldx1 * es) % 16 != 0) || (ldc * es) % 16 != 0) return false;
    if ((((uintptr_t) X0 | (uintptr_t) X1 | (uintptr_t) C) & 15) != 0) return false;
    panel_make_meta<T>(plan);
    const int slot = (es == 8) ? 0 : 1;
    panel_args<T> a;
    memset(&a, 0, sizeof(a));
    a.tile_ranges = (const int2 *) pn->d_tile_chunk_ptr;
    a.chunks = (const crp_panel_chunk *) pn->d_chunks[slot];
    a.ucol = pn->d_ucol;
    a.meta = pn->d_meta[slot];
    a.ntiles = pn->ntiles;  a.nchunks = pn->nchunks;
    a.CR = pn->CR;
    a.X0 = (const char *) X0;  a.ldx0 = ldx0 * es;
    a.X1 = (const char *) X1;  a.ldx1 = ldx1 * es;
    a.x0_rows = plan->x0_rows;
    a.n = n;
    a.alpha = alpha;  a.beta = beta;
    a.C = C;  a.ldc = ldc;
    // few rest rows without very long ones: the kernel's extra warp takes them (one row per block and turn)
    if (plan->rg.nrest > 0 && plan->rg.nrest <= 2048 && plan->lr.nlong == 0 && panel_env_int("CRP_PANEL_REST", 1))
    {
        a.nrest = plan->rg.nrest;  a.rest_rows = plan->rg.d_rest;
        a.rowptr = plan->d_rowptr;  a.colidx = plan->d_colidx;  a.val = val;
        *rest_done = true;
    }
    if (wait != NULL && wait->nwait > 0)
    {
        if (wait->nwait > 32) return false;                 // the caller waits with the separate kernel
        a.chunk_need = (pn->nslot == wait->nwait) ? pn->d_chunk_need : NULL;
        a.wait_all_first = (pn->nslot == wait->nwait) ? 0 : 1;
        a.flags = wait->flags;  a.wait_idx = wait->wait_idx;  a.nwait = wait->nwait;
        a.epoch = wait->epoch;  a.timeout_ns = wait->timeout_ns;  a.err = wait->err;
    }
    if (put != NULL && put->nflag > 0)
    {
        a.put_nrow = put->nrow;  a.put_row_bytes = (unsigned) put->row_bytes;  a.put_ridx = put->ridx;
        a.put_dst_rows = (char *const *) put->dst_rows;  a.put_flag_ptrs = put->flag_ptrs;  a.put_nflag = put->nflag;
        a.put_counter = put->counter;  a.epoch = put->epoch;  a.put_dst_off = put->dst_off;
    }
    {
        // development aid (builds with -DCRP_PANEL_TRACE_BUILD only): CRP_PANEL_TRACE=<file prefix> records 8 time stamps per block
        // of every launch (the last one is written to <prefix>.<pid> when the plan is destroyed)
        static long long *d_trace = NULL;
        static int trace_on = -1;
        if (trace_on < 0) { const char *e = getenv("CRP_PANEL_TRACE"); trace_on = (e && e[0]) ? 1 : 0; }
#ifndef CRP_PANEL_TRACE_BUILD
        if (trace_on == 1) { fprintf(stderr, "[crpspmm] CRP_PANEL_TRACE is set but this library was built without -DCRP_PANEL_TRACE_BUILD: no trace\n"); trace_on = 0; }
#endif
        if (trace_on)
        {
            if (d_trace == NULL) CRP_CUDA_CHECK(cudaMalloc((void **) &d_trace, sizeof(long long) * 8 * 1024));
            CRP_CUDA_CHECK(cudaMemsetAsync(d_trace, 0, sizeof(long long) * 8 * 1024, s));
            a.trace = d_trace;
            pn->d_trace = d_trace;
        }
    }
    switch (pn->R)
    {
        case 2: return panel_launch_R<T, VEC, 2>(plan, a, s);
        case 3: return panel_launch_R<T, VEC, 3>(plan, a, s);
        case 4: return panel_launch_R<T, VEC, 4>(plan, a, s);
        case 6: return panel_launch_R<T, VEC, 6>(plan, a, s);
        case 8: return panel_launch_R<T, VEC, 8>(plan, a, s);
        default: return false;
    }
}

template bool crp_launch_panel<double, 2>(crp_spmm_plan *, const double *, const int, const double *, size_t, const double *, size_t, double, double, double *, size_t, const crp_spmm_wait *, const crp_spmm_put *, bool *, cudaStream_t);
template bool crp_launch_panel<float, 4>(crp_spmm_plan *, const float *, const int, const float *, size_t, const float *, size_t, float, float, float *, size_t, const crp_spmm_wait *, const crp_spmm_put *, bool *, cudaStream_t);
