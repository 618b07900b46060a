// spmm_mergepath.cu - nnz-balanced CSR x dense kernel for matrices with skewed row lengths (RMAT, power-law graphs).
//
// Replaces the mkl_sparse_d_mm call at reference src/rowpara_spmm.c:404-407 where one-row-per-warp (spmm_rowsplit.cu)
// leaves warps with 1 nonzero next to warps with 10^5.
//
// Partition (plan time, host): the merge path over (row ends, nonzeros) is cut into chunks of at most ITEMS merge
// items (one item per nonzero + one per row end), as in merge-based CSR SpMV.  A cut that would fall inside a row
// shorter than a chunk is snapped back to the row's start, so only rows with more than ITEMS - 1 nonzeros are ever
// split: those are cut into segments of ITEMS nonzeros that write partial results into a scratch matrix, and a
// fix-up kernel adds the segments of each row in ascending order.  Every warp therefore does the same amount of
// work (+- one short row), the summation order is fixed (bit-reproducible run to run) and no floating-point atomics
// are used.  Rows without nonzeros are items too: their C rows are written (beta = 0 -> zeros).
//
// Kernel: one warp per chunk.  The chunk's column indices, values and row pointers are staged into the warp's
// shared-memory slice with cp.async (coalesced, no registers held) - the per-nonzero reads are then LDS broadcasts
// instead of dependent global loads; the B row segments are gathered with 128-bit read-only loads, four nonzeros in
// flight; C is written with streaming stores.  Bound: the B gather (L2 / HBM) - see DESIGN.md.
#include <cstring>
#include <vector>

#include "crp_cuda_internal.cuh"
#include "mergepath_build.hpp"

namespace {

template <typename T, int VEC> struct mvec;
template <> struct mvec<double, 2>
{
    static __device__ __forceinline__ void ld(const double *p, double (&v)[2]) { const double2 t = __ldg(reinterpret_cast<const double2 *>(p)); v[0] = t.x; v[1] = t.y; }
    static __device__ __forceinline__ void ldrw(const double *p, double (&v)[2]) { const double2 t = *reinterpret_cast<const double2 *>(p); v[0] = t.x; v[1] = t.y; }
    static __device__ __forceinline__ void st(double *p, const double (&v)[2]) { __stcs(reinterpret_cast<double2 *>(p), make_double2(v[0], v[1])); }
};
template <> struct mvec<float, 4>
{
    static __device__ __forceinline__ void ld(const float *p, float (&v)[4]) { const float4 t = __ldg(reinterpret_cast<const float4 *>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    static __device__ __forceinline__ void ldrw(const float *p, float (&v)[4]) { const float4 t = *reinterpret_cast<const float4 *>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    static __device__ __forceinline__ void st(float *p, const float (&v)[4]) { __stcs(reinterpret_cast<float4 *>(p), make_float4(v[0], v[1], v[2], v[3])); }
};
template <typename T> struct mvec<T, 1>
{
    static __device__ __forceinline__ void ld(const T *p, T (&v)[1]) { v[0] = __ldg(p); }
    static __device__ __forceinline__ void ldrw(const T *p, T (&v)[1]) { v[0] = *p; }
    static __device__ __forceinline__ void st(T *p, const T (&v)[1]) { __stcs(p, v[0]); }
};

__device__ __forceinline__ void cpa4(void *smem, const void *gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((unsigned) __cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cpa8(void *smem, const void *gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((unsigned) __cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}

}   // namespace

template <typename T, int VEC, int U, int ITEMS, int WPB>
__global__ void __launch_bounds__(WPB * 32) spmm_mergepath_kernel(
    const int nchunks, const int4 *__restrict__ desc, const int *__restrict__ rowptr, const int *__restrict__ colidx, const T *__restrict__ val,
    const int nv, const T *__restrict__ X0, const size_t ldx0, const int x0_rows, const T *__restrict__ X1, const size_t ldx1,
    const T alpha, const T beta, T *__restrict__ C, const size_t ldc, T *__restrict__ scratch, const size_t lds
)
{
    constexpr int NZ = 4;
    __shared__ int s_col[WPB][ITEMS];
    __shared__ T   s_val[WPB][ITEMS];
    __shared__ int s_rp[WPB][ITEMS + 1];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * WPB + wib;
    if (c >= nchunks) return;                       // warps are independent: no block-level barrier below
    const int4 d = __ldg(desc + c);                 // x: first row, y: rows (> 0) or -(scratch slot + 1) for a segment of row x, z / w: nonzero range
    const int p0 = d.z, cnt = d.w - d.z;
    const int nrows = d.y > 0 ? d.y : 0;
    for (int j = lane; j < cnt; j += 32)
    {
        cpa4(&s_col[wib][j], colidx + p0 + j);
        if (sizeof(T) == 8) cpa8(&s_val[wib][j], val + p0 + j); else cpa4(&s_val[wib][j], val + p0 + j);
    }
    for (int j = lane; j <= nrows; j += 32) cpa4(&s_rp[wib][j], rowptr + d.x + j);
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncwarp();

    const int *sc = s_col[wib];
    const T *sv = s_val[wib];
    for (int v0 = 0; v0 < nv; v0 += 32 * U)
    {
        const int nr = nrows > 0 ? nrows : 1;
        for (int r = 0; r < nr; r++)
        {
            const int jb = nrows > 0 ? s_rp[wib][r] - p0 : 0;
            const int je = nrows > 0 ? s_rp[wib][r + 1] - p0 : cnt;
            T acc[U][VEC];
            #pragma unroll
            for (int u = 0; u < U; u++)
                #pragma unroll
                for (int e = 0; e < VEC; e++) acc[u][e] = (T) 0;
            int j = jb;
            for (; j + NZ <= je; j += NZ)
            {
                T x[NZ][U][VEC], a[NZ];
                #pragma unroll
                for (int q = 0; q < NZ; q++)
                {
                    const int col = sc[j + q];
                    a[q] = sv[j + q];
                    const T *xr = (col < x0_rows) ? X0 + (size_t) col * ldx0 : X1 + (size_t) (col - x0_rows) * ldx1;
                    #pragma unroll
                    for (int u = 0; u < U; u++)
                    {
                        const int v = v0 + u * 32 + lane;
                        if (v < nv) mvec<T, VEC>::ld(xr + (size_t) v * VEC, x[q][u]);
                        else { for (int e = 0; e < VEC; e++) x[q][u][e] = (T) 0; }
                    }
                }
                #pragma unroll
                for (int q = 0; q < NZ; q++)
                    #pragma unroll
                    for (int u = 0; u < U; u++)
                        #pragma unroll
                        for (int e = 0; e < VEC; e++) acc[u][e] = fma(a[q], x[q][u][e], acc[u][e]);
            }
            for (; j < je; j++)
            {
                const int col = sc[j];
                const T a = sv[j];
                const T *xr = (col < x0_rows) ? X0 + (size_t) col * ldx0 : X1 + (size_t) (col - x0_rows) * ldx1;
                #pragma unroll
                for (int u = 0; u < U; u++)
                {
                    const int v = v0 + u * 32 + lane;
                    if (v >= nv) continue;
                    T x[VEC];
                    mvec<T, VEC>::ld(xr + (size_t) v * VEC, x);
                    #pragma unroll
                    for (int e = 0; e < VEC; e++) acc[u][e] = fma(a, x[e], acc[u][e]);
                }
            }
            if (nrows > 0)
            {
                if (jb == je && beta == (T) 1) continue;        // C row unchanged (received-rows pass of the overlap mode)
                T *crow = C + (size_t) (d.x + r) * ldc;
                #pragma unroll
                for (int u = 0; u < U; u++)
                {
                    const int v = v0 + u * 32 + lane;
                    if (v >= nv) continue;
                    T out[VEC];
                    if (beta == (T) 0)
                    {
                        #pragma unroll
                        for (int e = 0; e < VEC; e++) out[e] = alpha * acc[u][e];
                    } else {
                        T old[VEC];
                        mvec<T, VEC>::ldrw(crow + (size_t) v * VEC, old);
                        #pragma unroll
                        for (int e = 0; e < VEC; e++) out[e] = fma(alpha, acc[u][e], beta * old[e]);
                    }
                    mvec<T, VEC>::st(crow + (size_t) v * VEC, out);
                }
            } else {
                // segment of a long row: raw partial sum, combined by the fix-up kernel in segment order
                T *srow = scratch + (size_t) (-d.y - 1) * lds;
                #pragma unroll
                for (int u = 0; u < U; u++)
                {
                    const int v = v0 + u * 32 + lane;
                    if (v >= nv) continue;
                    T *q = srow + (size_t) v * VEC;
                    #pragma unroll
                    for (int e = 0; e < VEC; e++) q[e] = acc[u][e];
                }
            }
        }
    }
}

// C[row, :] = alpha * (sum of the row's segments, ascending) + beta * C[row, :]
template <typename T>
__global__ void __launch_bounds__(256) mergepath_fixup_kernel(
    const int nlong, const int *__restrict__ long_row, const int *__restrict__ long_sptr,
    const T *__restrict__ scratch, const size_t lds, const int n, const T alpha, const T beta, T *__restrict__ C, const size_t ldc
)
{
    const int i = blockIdx.y;
    if (i >= nlong) return;
    const int row = long_row[i], s0 = long_sptr[i], s1 = long_sptr[i + 1];
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
    {
        T acc = (T) 0;
        for (int s = s0; s < s1; s++) acc += scratch[(size_t) s * lds + j];
        T *c = C + (size_t) row * ldc + j;
        *c = (beta == (T) 0) ? alpha * acc : fma(alpha, acc, beta * *c);
    }
}

// ------------------------------------------------------------------------------ plan side (host)

void crp_mergepath_build(crp_spmm_plan *plan, const int *rowptr)
{
    crp_mergepath *mp = &plan->mp;
    memset(mp, 0, sizeof(*mp));
    if (plan->m == 0) return;
    crp_mergepath_host h;
    crp_mergepath_partition(plan->m, rowptr, CRP_MP_ITEMS, &h);
    auto upload = [](const void *src, size_t bytes) -> void * {
        void *d = NULL;
        if (bytes == 0) return d;
        CRP_CUDA_CHECK(cudaMalloc(&d, bytes));
        CRP_CUDA_CHECK(cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice));
        return d;
    };
    mp->nchunks = (int) (h.desc.size() / 4);
    mp->nlong = (int) h.long_row.size();
    mp->nseg = h.nseg;
    mp->d_desc = upload(h.desc.data(), sizeof(int) * h.desc.size());
    mp->d_long_row = (int *) upload(h.long_row.data(), sizeof(int) * h.long_row.size());
    mp->d_long_sptr = (int *) upload(h.long_sptr.data(), sizeof(int) * h.long_sptr.size());
}

void crp_mergepath_destroy(crp_spmm_plan *plan)
{
    crp_mergepath *mp = &plan->mp;
    if (mp->d_desc) CRP_CUDA_CHECK(cudaFree(mp->d_desc));
    if (mp->d_long_row) CRP_CUDA_CHECK(cudaFree(mp->d_long_row));
    if (mp->d_long_sptr) CRP_CUDA_CHECK(cudaFree(mp->d_long_sptr));
    if (mp->d_scratch) CRP_CUDA_CHECK(cudaFree(mp->d_scratch));
    memset(mp, 0, sizeof(*mp));
}

// ------------------------------------------------------------------------------------- launch

template <typename T, int VEC, int U>
static void mp_launch(
    crp_spmm_plan *plan, const T *val, const int nv, const T *X0, size_t ldx0, const T *X1, size_t ldx1, T alpha, T beta, T *C, size_t ldc, cudaStream_t s
)
{
    crp_mergepath *mp = &plan->mp;
    constexpr int WPB = 8;
    const size_t lds = (size_t) nv * VEC;
    if (mp->nseg > 0)
    {
        const size_t need = sizeof(T) * (size_t) mp->nseg * lds;
        if (need > mp->scratch_bytes)
        {
            CRP_CUDA_CHECK(cudaStreamSynchronize(s));
            if (mp->d_scratch) CRP_CUDA_CHECK(cudaFree(mp->d_scratch));
            CRP_CUDA_CHECK(cudaMalloc(&mp->d_scratch, need));
            mp->scratch_bytes = need;
        }
    }
    const unsigned blocks = (unsigned) ((mp->nchunks + WPB - 1) / WPB);
    spmm_mergepath_kernel<T, VEC, U, CRP_MP_ITEMS, WPB><<<blocks, WPB * 32, 0, s>>>(
        mp->nchunks, (const int4 *) mp->d_desc, plan->d_rowptr, plan->d_colidx, val, nv, X0, ldx0, plan->x0_rows, X1, ldx1,
        alpha, beta, C, ldc, (T *) mp->d_scratch, lds);
    CRP_LAUNCH_CHECK();
    const int n = nv * VEC;
    for (int base = 0; base < mp->nlong; base += 65535)
    {
        const int cnt = (mp->nlong - base < 65535) ? mp->nlong - base : 65535;
        mergepath_fixup_kernel<T><<<dim3((unsigned) ((n + 255) / 256), (unsigned) cnt), 256, 0, s>>>(
            cnt, mp->d_long_row + base, mp->d_long_sptr + base, (const T *) mp->d_scratch, lds, n, alpha, beta, C, ldc);
        CRP_LAUNCH_CHECK();
    }
}

// false: not applicable (no chunks / dense matrix too narrow for one warp per chunk) - the caller uses the row-split kernel
template <typename T, int VECN>
bool crp_launch_mergepath(
    crp_spmm_plan *plan, const T *val, const int n, const T *X0, size_t ldx0, const T *X1, size_t ldx1, T alpha, T beta, T *C, size_t ldc, cudaStream_t s
)
{
    if (plan->mp.nchunks == 0) return false;
    const uintptr_t ptrs = (uintptr_t) X0 | (uintptr_t) X1 | (uintptr_t) C;
    const bool vec_ok = (n % VECN == 0) && (ldx0 % VECN == 0) && (X1 == NULL || ldx1 % VECN == 0) && (ldc % VECN == 0) && ((ptrs & 15) == 0);
    if (vec_ok)
    {
        const int nv = n / VECN;
        if (nv >= 128)     mp_launch<T, VECN, 4>(plan, val, nv, X0, ldx0, X1, ldx1, alpha, beta, C, ldc, s);
        else if (nv >= 64) mp_launch<T, VECN, 2>(plan, val, nv, X0, ldx0, X1, ldx1, alpha, beta, C, ldc, s);
        else if (nv >= 16) mp_launch<T, VECN, 1>(plan, val, nv, X0, ldx0, X1, ldx1, alpha, beta, C, ldc, s);
        else return false;
    } else {
        if (n >= 64)      mp_launch<T, 1, 2>(plan, val, n, X0, ldx0, X1, ldx1, alpha, beta, C, ldc, s);
        else if (n >= 16) mp_launch<T, 1, 1>(plan, val, n, X0, ldx0, X1, ldx1, alpha, beta, C, ldc, s);
        else return false;
    }
    return true;
}

template bool crp_launch_mergepath<double, 2>(crp_spmm_plan *, const double *, const int, const double *, size_t, const double *, size_t, double, double, double *, size_t, cudaStream_t);
template bool crp_launch_mergepath<float, 4>(crp_spmm_plan *, const float *, const int, const float *, size_t, const float *, size_t, float, float, float *, size_t, cudaStream_t);
