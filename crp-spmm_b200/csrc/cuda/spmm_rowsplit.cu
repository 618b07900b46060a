// spmm_rowsplit.cu - general CSR x dense kernel: one row per (sub-)warp.
//
// Replaces the mkl_sparse_d_mm call at reference src/rowpara_spmm.c:404-407
// (C := 1.0 * A * rB + 0.0 * C, A general 0-based CSR, dense row-major).
//
// Mapping: a row of C is owned by LPR consecutive lanes (LPR = 32 for wide
// dense matrices, fewer when n is small so that a warp covers 32 / LPR rows);
// every lane keeps U 128-bit accumulators, so one pass over the row's nonzeros
// produces LPR * U * 16 bytes of the C row.  For each nonzero the column index
// and value are fetched with a broadcast load and the B row segment with U
// fully coalesced 128-bit read-only loads; four nonzeros are in flight at once.
// C is written with streaming stores, never read when beta == 0.
// Bound: HBM for A and C, L2/L1 gather bandwidth for B (no reuse across rows in
// this variant - that is what the row-block variant adds).
#include "crp_cuda_internal.cuh"

template <typename T, int VEC> struct vec_io;

template <> struct vec_io<double, 2>
{
    static __device__ __forceinline__ void load(const double *p, double (&v)[2]) { const double2 t = __ldg(reinterpret_cast<const double2 *>(p)); v[0] = t.x; v[1] = t.y; }
    static __device__ __forceinline__ void load_rw(const double *p, double (&v)[2]) { const double2 t = *reinterpret_cast<const double2 *>(p); v[0] = t.x; v[1] = t.y; }
    static __device__ __forceinline__ void store(double *p, const double (&v)[2]) { __stcs(reinterpret_cast<double2 *>(p), make_double2(v[0], v[1])); }
};
template <> struct vec_io<float, 4>
{
    static __device__ __forceinline__ void load(const float *p, float (&v)[4]) { const float4 t = __ldg(reinterpret_cast<const float4 *>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    static __device__ __forceinline__ void load_rw(const float *p, float (&v)[4]) { const float4 t = *reinterpret_cast<const float4 *>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    static __device__ __forceinline__ void store(float *p, const float (&v)[4]) { __stcs(reinterpret_cast<float4 *>(p), make_float4(v[0], v[1], v[2], v[3])); }
};
template <typename T> struct vec_io<T, 1>
{
    static __device__ __forceinline__ void load(const T *p, T (&v)[1]) { v[0] = __ldg(p); }
    static __device__ __forceinline__ void load_rw(const T *p, T (&v)[1]) { v[0] = *p; }
    static __device__ __forceinline__ void store(T *p, const T (&v)[1]) { __stcs(p, v[0]); }
};

template <typename T, int VEC, int LPR, int U>
__global__ void __launch_bounds__(256) spmm_rowsplit_kernel(
    const int m, const int *__restrict__ row_list, const int *__restrict__ pbeg, const int *__restrict__ pend, const int *__restrict__ colidx, const T *__restrict__ val,
    const int nv,                                   // 128-bit (or scalar) column groups per row
    const T *__restrict__ X0, const size_t ldx0, const int x0_rows,
    const T *__restrict__ X1, const size_t ldx1,
    const T alpha, const T beta, T *__restrict__ C, const size_t ldc
)
{
    constexpr int RW = 32 / LPR;                    // rows per warp
    constexpr int NZ = 4;                           // nonzeros in flight
    const int warp = (int) (((size_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    const int ridx = warp * RW + lane / LPR;        // position in the row list (or the row itself)
    const int l    = lane % LPR;
    int row = ridx, p_beg = 0, p_end = 0;
    if (ridx < m)
    {
        if (row_list != NULL) row = __ldg(row_list + ridx);
        p_beg = __ldg(pbeg + row);
        p_end = __ldg(pend + row);
    }

    for (int v0 = 0; v0 < nv; v0 += LPR * U)
    {
        T acc[U][VEC];
        #pragma unroll
        for (int u = 0; u < U; u++)
            #pragma unroll
            for (int e = 0; e < VEC; e++) acc[u][e] = (T) 0;

        int p = p_beg;
        for (; p + NZ <= p_end; p += NZ)
        {
            int c[NZ]; T a[NZ];
            #pragma unroll
            for (int q = 0; q < NZ; q++) { c[q] = __ldg(colidx + p + q); a[q] = __ldg(val + p + q); }
            T x[NZ][U][VEC];
            #pragma unroll
            for (int q = 0; q < NZ; q++)
            {
                const T *xr = (c[q] < x0_rows) ? X0 + (size_t) c[q] * ldx0 : X1 + (size_t) (c[q] - x0_rows) * ldx1;
                #pragma unroll
                for (int u = 0; u < U; u++)
                {
                    const int v = v0 + u * LPR + l;
                    if (v < nv) vec_io<T, VEC>::load(xr + (size_t) v * VEC, x[q][u]);
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
        for (; p < p_end; p++)
        {
            const int c = __ldg(colidx + p);
            const T a = __ldg(val + p);
            const T *xr = (c < x0_rows) ? X0 + (size_t) c * ldx0 : X1 + (size_t) (c - x0_rows) * ldx1;
            #pragma unroll
            for (int u = 0; u < U; u++)
            {
                const int v = v0 + u * LPR + l;
                if (v < nv)
                {
                    T x[VEC];
                    vec_io<T, VEC>::load(xr + (size_t) v * VEC, x);
                    #pragma unroll
                    for (int e = 0; e < VEC; e++) acc[u][e] = fma(a, x[e], acc[u][e]);
                }
            }
        }

        // beta == 1 and no nonzeros: C row unchanged, do not touch it (the received-rows pass of the overlap mode)
        if (ridx < m && !(p_beg == p_end && beta == (T) 1))
        {
            T *crow = C + (size_t) row * ldc;
            #pragma unroll
            for (int u = 0; u < U; u++)
            {
                const int v = v0 + u * LPR + l;
                if (v >= nv) continue;
                T out[VEC];
                if (beta == (T) 0)
                {
                    #pragma unroll
                    for (int e = 0; e < VEC; e++) out[e] = alpha * acc[u][e];
                } else {
                    T old[VEC];
                    vec_io<T, VEC>::load_rw(crow + (size_t) v * VEC, old);
                    #pragma unroll
                    for (int e = 0; e < VEC; e++) out[e] = fma(alpha, acc[u][e], beta * old[e]);
                }
                vec_io<T, VEC>::store(crow + (size_t) v * VEC, out);
            }
        }
    }
}

template <typename T, int VEC, int LPR, int U>
static void launch_one(
    const crp_spmm_plan *plan, const int nrows, const int *row_list, const int *pbeg, const int *pend, const T *val, const int nv, const T *X0, size_t ldx0, int x0_rows, const T *X1, size_t ldx1,
    T alpha, T beta, T *C, size_t ldc, cudaStream_t stream
)
{
    constexpr int RW = 32 / LPR;
    const long long warps  = ((long long) nrows + RW - 1) / RW;
    const long long blocks = (warps + 7) / 8;
    if (blocks == 0) return;
    spmm_rowsplit_kernel<T, VEC, LPR, U><<<(unsigned) blocks, 256, 0, stream>>>(
        nrows, row_list, pbeg, pend, plan->d_colidx, val, nv, X0, ldx0, x0_rows, X1, ldx1, alpha, beta, C, ldc);
    CRP_LAUNCH_CHECK();
}

template <typename T, int VEC>
static void launch_width(
    const crp_spmm_plan *plan, const int nrows, const int *row_list, const int *pbeg, const int *pend, const T *val, const int nv, const T *X0, size_t ldx0, int x0_rows, const T *X1, size_t ldx1,
    T alpha, T beta, T *C, size_t ldc, cudaStream_t stream
)
{
#define CRP_RS(LPR, U) launch_one<T, VEC, LPR, U>(plan, nrows, row_list, pbeg, pend, val, nv, X0, ldx0, x0_rows, X1, ldx1, alpha, beta, C, ldc, stream)
    if (nv >= 128)     CRP_RS(32, 4);
    else if (nv >= 64) CRP_RS(32, 2);
    else if (nv > 16)  CRP_RS(32, 1);
    else if (nv > 8)   CRP_RS(16, 1);
    else if (nv > 4)   CRP_RS(8, 1);
    else if (nv > 2)   CRP_RS(4, 1);
    else               CRP_RS(2, 1);
#undef CRP_RS
}

template <typename T, int VECN>
void crp_launch_rowsplit(
    const crp_spmm_plan *plan, const int nrows, const int *row_list, const int *pbeg, const int *pend, const T *val, const int n, const T *X0, size_t ldx0, int x0_rows, const T *X1, size_t ldx1,
    T alpha, T beta, T *C, size_t ldc, cudaStream_t stream
)
{
    const uintptr_t ptrs = (uintptr_t) X0 | (uintptr_t) X1 | (uintptr_t) C;
    const bool vec_ok = (n % VECN == 0) && (ldx0 % VECN == 0) && (X1 == NULL || ldx1 % VECN == 0) && (ldc % VECN == 0) && ((ptrs & 15) == 0);
    if (vec_ok) launch_width<T, VECN>(plan, nrows, row_list, pbeg, pend, val, n / VECN, X0, ldx0, x0_rows, X1, ldx1, alpha, beta, C, ldc, stream);
    else        launch_width<T, 1>(plan, nrows, row_list, pbeg, pend, val, n, X0, ldx0, x0_rows, X1, ldx1, alpha, beta, C, ldc, stream);
}

template void crp_launch_rowsplit<double, 2>(const crp_spmm_plan *, const int, const int *, const int *, const int *, const double *, const int, const double *, size_t, int, const double *, size_t, double, double, double *, size_t, cudaStream_t);
template void crp_launch_rowsplit<float, 4>(const crp_spmm_plan *, const int, const int *, const int *, const int *, const float *, const int, const float *, size_t, int, const float *, size_t, float, float, float *, size_t, cudaStream_t);
