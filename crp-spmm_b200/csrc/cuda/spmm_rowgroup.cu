// spmm_rowgroup.cu - register-blocked CSR x dense kernel for matrices whose consecutive
// rows share their column pattern (FEM / multi-dof matrices such as pwtk: 6 unknowns per node).
//
// Replaces the mkl_sparse_d_mm call at reference src/rowpara_spmm.c:404-407 for such matrices.
//
// Why: the plain row-split kernel (spmm_rowsplit.cu) issues one B-row load per nonzero, so it
// moves nnz * n * s bytes through L1 (24.4 GB for the pwtk-shaped n = 256 case, 23x the
// compulsory HBM traffic) and is bound by L1/L2 throughput at a few % of the HBM roofline.
// Here a group of R consecutive rows whose column sets are identical is stored as a list of
// R x 1 column blocks; one B-row segment is loaded ONCE per block and reused from registers for
// all R rows: L1 traffic drops by R and the inner loop becomes R FMAs per loaded value.
// Groups that do not have this structure are left to the row-split kernel (no explicit zeros
// are ever multiplied, so results on Inf / NaN inputs are the reference's).
//
// Mapping: a group is owned by LPR lanes (a full warp for wide dense matrices); each lane keeps
// R x U 128-bit accumulators (U * LPR * 16 bytes of every C row of the group).  Per block the
// lanes issue U coalesced 128-bit read-only loads of the B row and broadcast loads of the R
// values; two blocks are in flight per iteration and the column indices of the next iteration
// are prefetched so that the index -> address -> load chain is off the critical path.
// C is written once with streaming stores.  Bound: fp64 FMA issue and L1/L2 gather bandwidth
// for fp64 n = 256, HBM for A / first touch of B / C.
#include <algorithm>
#include <cstring>
#include <vector>

#include "crp_cuda_internal.cuh"
#include "rowgroup_build.hpp"

// ------------------------------------------------------------------ plan-time analysis (host)
// The analysis itself is rowgroup_build.hpp (pure C++, also exercised by the CPU tests).

static double rg_min_fill()
{
    // relaxed (masked) groups: CRP_SPMM_RG_FILL = minimum fraction of real nonzeros in a group's R x |union| slots;
    // 1.0 = exact groups only
    const char *e = getenv("CRP_SPMM_RG_FILL");
    double f = (e && e[0]) ? atof(e) : 0.75;
    if (f < 0.05) f = 0.05;
    if (f > 1.0) f = 1.0;
    return f;
}

static crp_rg_choice rg_decide(const int m, const int *rowptr, const int *colidx, const crp_rg_rowinfo &ri, const int forced, const int n_hint)
{
    // exact groups first (cheap: flags only); relaxed groups are considered when the exact ones leave more than 10 % of the
    // nonzeros to the row-split kernel, on a sample of the rows, and only where the panel kernel can run them (n >= 64)
    crp_rg_choice ex = crp_rg_choose(m, rowptr, colidx, ri, forced, 1.0, 0);
    const long long nnz = (long long) rowptr[m] - rowptr[0];
    const double fill = rg_min_fill();
    if (fill >= 1.0 || n_hint < 64 || (ex.R > 1 && ex.grouped_nnz * 10 >= nnz * 9)) return ex;
    const int sample = 1 << 16;
    crp_rg_choice rx = crp_rg_choose(m, rowptr, colidx, ri, forced, fill, sample);
    if (rx.R <= 1) return ex;
    const int mm = (sample < m) ? sample : m;
    const long long nnz_s = (long long) rowptr[mm] - rowptr[0];
    const double frac_rx = nnz_s > 0 ? (double) rx.grouped_nnz / (double) nnz_s : 0.0;
    const double frac_ex = nnz > 0 ? (double) ex.grouped_nnz / (double) nnz : 0.0;
    if (ex.R > 1 && frac_rx < frac_ex + 0.05) return ex;
    rx.all_exact = false;
    return rx;
}

// Host-only view of the plan-time decision (no device needed): which group size / alignment would be used for this CSR
// pattern and how many R x 1 blocks it yields.  R = 1 means "row-split kernel only".
extern "C" void crp_cuda_spmm_analyse(const int m, const int *rowptr, const int *colidx, int *R, int *offset, long long *nblk, int *n_long_rows)
{
    *R = 1;  *offset = 0;  *nblk = 0;  *n_long_rows = 0;
    if (m <= 0 || rowptr[m] == rowptr[0]) return;
    int forced = 0;
    if (const char *e = getenv("CRP_SPMM_ROWGROUP_R")) forced = atoi(e);
    if (forced != 1)
    {
        crp_rg_rowinfo ri;
        crp_rg_scan_rows(m, rowptr, colidx, &ri);
        const crp_rg_choice c = crp_rg_choose(m, rowptr, colidx, ri, forced, 1.0, 0);
        *R = c.R;  *offset = c.off;  *nblk = c.nblk;
    }
    for (int i = 0; i < m; i++) if (rowptr[i + 1] - rowptr[i] > CRP_LONG_ROW) (*n_long_rows)++;
}

void crp_rowgroup_build(crp_spmm_plan *plan, const int *rowptr, const int *colidx, const double *val, std::vector<int> *rest_out)
{
    if (rest_out) rest_out->clear();
    crp_rowgroup *rg = &plan->rg;
    memset(rg, 0, sizeof(*rg));
    const int m = plan->m;
    if (m == 0 || plan->nnz == 0) return;
    int forced = 0;
    if (const char *e = getenv("CRP_SPMM_ROWGROUP_R")) forced = atoi(e);      // 0 auto, 1 disable, else force R
    if (forced == 1) return;
    crp_rg_rowinfo ri;
    crp_rg_scan_rows(m, rowptr, colidx, &ri);
    const crp_rg_choice ch = rg_decide(m, rowptr, colidx, ri, forced, plan->n_hint);
    if (ch.R == 1) return;

    crp_rowgroup_host *rh = new crp_rowgroup_host();
    std::vector<int> rest_rows;
    long long rest = 0;
    crp_rg_build(m, rowptr, colidx, val, ri, ch.R, ch.off, ch.all_exact ? 1.0 : rg_min_fill(), rh, &rest_rows, &rest);
    if (rh->g_row.empty()) { delete rh; return; }
    rg->R = ch.R;
    rg->exact = rh->b_mask.empty() ? 1 : 0;
    rg->ngroups = (int) rh->g_row.size();
    rg->nblk = (long long) rh->b_col.size();
    rg->nrest = (int) rest_rows.size();
    rg->rest_nnz = rest;
    auto upload = [](const void *src, size_t bytes) -> void * {
        void *d = NULL;
        if (bytes == 0) return d;
        CRP_CUDA_CHECK(cudaMalloc(&d, bytes));
        CRP_CUDA_CHECK(cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice));
        return d;
    };
    if (rg->exact)
    {
        // arrays of the register-blocked fallback kernels below (the panel kernel has its own records)
        rg->d_grow = (int *) upload(rh->g_row.data(), sizeof(int) * rh->g_row.size());
        rg->d_gptr = (int *) upload(rh->g_ptr.data(), sizeof(int) * rh->g_ptr.size());
        rg->d_bcol = (int *) upload(rh->b_col.data(), sizeof(int) * rh->b_col.size());
        rg->d_bval = (double *) upload(rh->b_val.data(), sizeof(double) * rh->b_val.size());
    }
    rg->d_rest = (int *) upload(rest_rows.data(), sizeof(int) * rest_rows.size());
    if (rest_out) rest_out->swap(rest_rows);
    plan->rg_host = rh;
    crp_panel_build(plan);
}

void crp_rowgroup_destroy(crp_spmm_plan *plan)
{
    crp_rowgroup *rg = &plan->rg;
    crp_panel_destroy(plan);
    delete plan->rg_host;
    plan->rg_host = NULL;
    if (rg->d_grow) CRP_CUDA_CHECK(cudaFree(rg->d_grow));
    if (rg->d_gptr) CRP_CUDA_CHECK(cudaFree(rg->d_gptr));
    if (rg->d_bcol) CRP_CUDA_CHECK(cudaFree(rg->d_bcol));
    if (rg->d_bval) CRP_CUDA_CHECK(cudaFree(rg->d_bval));
    if (rg->d_bval32) CRP_CUDA_CHECK(cudaFree(rg->d_bval32));
    if (rg->d_rest) CRP_CUDA_CHECK(cudaFree(rg->d_rest));
    memset(rg, 0, sizeof(*rg));
}

// ------------------------------------------------------------------------------- kernel

template <typename T> struct vec128;
template <> struct vec128<double> { typedef double2 type; static constexpr int N = 2; };
template <> struct vec128<float>  { typedef float4  type; static constexpr int N = 4; };

template <typename T, int VEC> struct xload;
template <> struct xload<double, 2>
{
    static __device__ __forceinline__ void ld(const double *p, double (&v)[2]) { const double2 t = __ldg(reinterpret_cast<const double2 *>(p)); v[0] = t.x; v[1] = t.y; }
    static __device__ __forceinline__ void st(double *p, const double (&v)[2]) { __stcs(reinterpret_cast<double2 *>(p), make_double2(v[0], v[1])); }
};
template <> struct xload<float, 4>
{
    static __device__ __forceinline__ void ld(const float *p, float (&v)[4]) { const float4 t = __ldg(reinterpret_cast<const float4 *>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    static __device__ __forceinline__ void st(float *p, const float (&v)[4]) { __stcs(reinterpret_cast<float4 *>(p), make_float4(v[0], v[1], v[2], v[3])); }
};
template <typename T> struct xload<T, 1>
{
    static __device__ __forceinline__ void ld(const T *p, T (&v)[1]) { v[0] = __ldg(p); }
    static __device__ __forceinline__ void st(T *p, const T (&v)[1]) { __stcs(p, v[0]); }
};

// the R values of one block: 128-bit broadcast loads where the block stride allows it
template <typename T, int R>
__device__ __forceinline__ void load_block_vals(const T *__restrict__ p, T (&a)[R])
{
    constexpr int PER16 = 16 / (int) sizeof(T);
    if constexpr ((R % PER16) == 0)
    {
        #pragma unroll
        for (int i = 0; i < R / PER16; i++)
        {
            const typename vec128<T>::type t = __ldg(reinterpret_cast<const typename vec128<T>::type *>(p) + i);
            const T *tp = reinterpret_cast<const T *>(&t);
            #pragma unroll
            for (int e = 0; e < PER16; e++) a[i * PER16 + e] = tp[e];
        }
    } else if constexpr (sizeof(T) == 4 && (R % 2) == 0) {
        #pragma unroll
        for (int i = 0; i < R / 2; i++)
        {
            const float2 t = __ldg(reinterpret_cast<const float2 *>(p) + i);
            a[2 * i] = t.x; a[2 * i + 1] = t.y;
        }
    } else {
        #pragma unroll
        for (int i = 0; i < R; i++) a[i] = __ldg(p + i);
    }
}

// Register-only variant for narrow dense matrices (a group is owned by LPR < 32 lanes): NB blocks in flight per step,
// the column indices of the next step prefetched.  beta == 0 only (the caller routes beta != 0 to the staged variant).
template <typename T, int VEC, int R, int LPR, int U, int NB, int BS>
__global__ void __launch_bounds__(BS) spmm_rowgroup_kernel(
    const int ngroups, const int *__restrict__ grow, const int *__restrict__ gptr,
    const int *__restrict__ bcol, const T *__restrict__ bval,
    const int nv,                                   // VEC-wide column groups per dense row
    const T *__restrict__ X0, const size_t ldx0, const int x0_rows,
    const T *__restrict__ X1, const size_t ldx1,
    const T alpha, T *__restrict__ C, const size_t ldc
)
{
    constexpr int GW = 32 / LPR;                    // groups per warp
    const int warp = (int) ((blockIdx.x * (unsigned) blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    const int g = warp * GW + lane / LPR;
    const int l = lane % LPR;
    const int v0 = blockIdx.y * (LPR * U);          // first column group of this CTA's column chunk
    int p = 0, p_end = 0, row0 = 0;
    if (g < ngroups) { p = __ldg(gptr + g); p_end = __ldg(gptr + g + 1); row0 = __ldg(grow + g); }

    int voff[U];                                    // element offset of this lane's column groups; -1: beyond n
    #pragma unroll
    for (int u = 0; u < U; u++) { const int v = v0 + u * LPR + l; voff[u] = (v < nv) ? v * VEC : -1; }

    T acc[R][U][VEC];
    #pragma unroll
    for (int r = 0; r < R; r++)
        #pragma unroll
        for (int u = 0; u < U; u++)
            #pragma unroll
            for (int e = 0; e < VEC; e++) acc[r][u][e] = (T) 0;

    int cn[NB];
    #pragma unroll
    for (int q = 0; q < NB; q++) cn[q] = (p + q < p_end) ? __ldg(bcol + p + q) : 0;
    for (; p + NB <= p_end; p += NB)
    {
        int c[NB];
        #pragma unroll
        for (int q = 0; q < NB; q++) c[q] = cn[q];
        #pragma unroll
        for (int q = 0; q < NB; q++) cn[q] = (p + NB + q < p_end) ? __ldg(bcol + p + NB + q) : 0;
        T x[NB][U][VEC], a[NB][R];
        #pragma unroll
        for (int q = 0; q < NB; q++)
        {
            const T *xr = (c[q] < x0_rows) ? X0 + (size_t) c[q] * ldx0 : X1 + (size_t) (c[q] - x0_rows) * ldx1;
            #pragma unroll
            for (int u = 0; u < U; u++)
            {
                if (voff[u] >= 0) xload<T, VEC>::ld(xr + voff[u], x[q][u]);
                else { for (int e = 0; e < VEC; e++) x[q][u][e] = (T) 0; }
            }
            load_block_vals<T, R>(bval + (size_t) (p + q) * R, a[q]);
        }
        #pragma unroll
        for (int q = 0; q < NB; q++)
            #pragma unroll
            for (int r = 0; r < R; r++)
                #pragma unroll
                for (int u = 0; u < U; u++)
                    #pragma unroll
                    for (int e = 0; e < VEC; e++) acc[r][u][e] = fma(a[q][r], x[q][u][e], acc[r][u][e]);
    }
    // fewer than NB blocks left; cn[] holds their columns
    #pragma unroll
    for (int q = 0; q < NB - 1; q++)
    {
        if (p + q >= p_end) break;
        const int c = cn[q];
        const T *xr = (c < x0_rows) ? X0 + (size_t) c * ldx0 : X1 + (size_t) (c - x0_rows) * ldx1;
        T a[R];
        load_block_vals<T, R>(bval + (size_t) (p + q) * R, a);
        #pragma unroll
        for (int u = 0; u < U; u++)
        {
            if (voff[u] < 0) continue;
            T x[VEC];
            xload<T, VEC>::ld(xr + voff[u], x);
            #pragma unroll
            for (int r = 0; r < R; r++)
                #pragma unroll
                for (int e = 0; e < VEC; e++) acc[r][u][e] = fma(a[r], x[e], acc[r][u][e]);
        }
    }

    if (g < ngroups)
    {
        #pragma unroll
        for (int r = 0; r < R; r++)
        {
            T *crow = C + (size_t) (row0 + r) * ldc;
            #pragma unroll
            for (int u = 0; u < U; u++)
            {
                if (voff[u] < 0) continue;
                T out[VEC];
                #pragma unroll
                for (int e = 0; e < VEC; e++) out[e] = alpha * acc[r][u][e];
                xload<T, VEC>::st(crow + voff[u], out);
            }
        }
    }
}

// ---- variant with the block values and columns staged through shared memory ("sv") ----
// A warp owns one group.  It copies the values and columns of the next CB blocks of its group into its private
// shared-memory ring with cp.async (LDGSTS, no registers held) while it works on the current CB blocks, and reads
// them back with broadcast LDS; only the B-row gathers remain long-latency loads.  Round-1 kernel of the headline
// configuration (0.42 ms); since round 2 the fallback of spmm_panel.cu for operands the bulk copies cannot take
// (unaligned pointers / leading dimensions, n < 64).
__device__ __forceinline__ void cp_async_bytes16(void *smem, const void *gmem)
{
    const unsigned s = (unsigned) __cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_bytes8(void *smem, const void *gmem)
{
    const unsigned s = (unsigned) __cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_bytes4(void *smem, const void *gmem)
{
    const unsigned s = (unsigned) __cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <typename T, int R>
__device__ __forceinline__ void lds_block_vals(const T *p, T (&a)[R])
{
    constexpr int PER16 = 16 / (int) sizeof(T);
    if constexpr ((R % PER16) == 0)
    {
        #pragma unroll
        for (int i = 0; i < R / PER16; i++)
        {
            const typename vec128<T>::type t = reinterpret_cast<const typename vec128<T>::type *>(p)[i];
            const T *tp = reinterpret_cast<const T *>(&t);
            #pragma unroll
            for (int e = 0; e < PER16; e++) a[i * PER16 + e] = tp[e];
        }
    } else {
        #pragma unroll
        for (int i = 0; i < R; i++) a[i] = p[i];
    }
}

template <typename T, int VEC, int R, int U, int NB, int BS, int CB, int MINB>
__global__ void __launch_bounds__(BS, MINB) spmm_rowgroup_sv_kernel(
    const int ngroups, const int *__restrict__ grow, const int *__restrict__ gptr,
    const int *__restrict__ bcol, const T *__restrict__ bval,
    const int nv,
    const T *__restrict__ X0, const size_t ldx0, const int x0_rows,
    const T *__restrict__ X1, const size_t ldx1,
    const T alpha, const T beta, T *__restrict__ C, const size_t ldc
)
{
    constexpr int WPB = BS / 32;
    constexpr int VAL_BYTES = CB * R * (int) sizeof(T);
    constexpr int GRAN = (R * (int) sizeof(T)) % 16 == 0 ? 16 : ((R * (int) sizeof(T)) % 8 == 0 ? 8 : 4);
    __shared__ __align__(16) unsigned char s_val[WPB][2][VAL_BYTES];
    __shared__ int s_col[WPB][2][CB];

    const int wib  = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * WPB + wib;            // one group per warp
    const int v0 = blockIdx.y * (32 * U);
    if (g >= ngroups) return;                        // whole warp leaves together: no block-level barrier below
    const int p_beg = __ldg(gptr + g), p_end = __ldg(gptr + g + 1), row0 = __ldg(grow + g);

    int voff[U];
    #pragma unroll
    for (int u = 0; u < U; u++) { const int v = v0 + u * 32 + lane; voff[u] = (v < nv) ? v * VEC : -1; }

    T acc[R][U][VEC];
    #pragma unroll
    for (int r = 0; r < R; r++)
        #pragma unroll
        for (int u = 0; u < U; u++)
            #pragma unroll
            for (int e = 0; e < VEC; e++) acc[r][u][e] = (T) 0;

    // copy values and columns of blocks [pc, pc + nb) into ring slot `buf`
    auto stage = [&](const int buf, const int pc, const int nb) {
        const char *src = (const char *) (bval + (size_t) pc * R);
        const int nbytes = nb * R * (int) sizeof(T);
        for (int off = lane * GRAN; off < nbytes; off += 32 * GRAN)
        {
            if (GRAN == 16) cp_async_bytes16(&s_val[wib][buf][off], src + off);
            else if (GRAN == 8) cp_async_bytes8(&s_val[wib][buf][off], src + off);
            else cp_async_bytes4(&s_val[wib][buf][off], src + off);
        }
        for (int j = lane; j < nb; j += 32) cp_async_bytes4(&s_col[wib][buf][j], bcol + pc + j);
        cp_async_commit();
    };

    int buf = 0;
    stage(0, p_beg, min(CB, p_end - p_beg));
    for (int pc = p_beg; pc < p_end; pc += CB, buf ^= 1)
    {
        const int nb = min(CB, p_end - pc);
        cp_async_wait_all();
        __syncwarp();
        if (pc + CB < p_end) stage(buf ^ 1, pc + CB, min(CB, p_end - pc - CB));
        const T *sv = reinterpret_cast<const T *>(&s_val[wib][buf][0]);
        const int *sc = &s_col[wib][buf][0];
        int j = 0;
        for (; j + NB <= nb; j += NB)
        {
            T x[NB][U][VEC];
            #pragma unroll
            for (int q = 0; q < NB; q++)
            {
                const int c = sc[j + q];
                const T *xr = (c < x0_rows) ? X0 + (size_t) c * ldx0 : X1 + (size_t) (c - x0_rows) * ldx1;
                #pragma unroll
                for (int u = 0; u < U; u++)
                {
                    if (voff[u] >= 0) xload<T, VEC>::ld(xr + voff[u], x[q][u]);
                    else { for (int e = 0; e < VEC; e++) x[q][u][e] = (T) 0; }
                }
            }
            #pragma unroll
            for (int q = 0; q < NB; q++)
            {
                T a[R];
                lds_block_vals<T, R>(sv + (size_t) (j + q) * R, a);
                #pragma unroll
                for (int r = 0; r < R; r++)
                    #pragma unroll
                    for (int u = 0; u < U; u++)
                        #pragma unroll
                        for (int e = 0; e < VEC; e++) acc[r][u][e] = fma(a[r], x[q][u][e], acc[r][u][e]);
            }
        }
        for (; j < nb; j++)
        {
            const int c = sc[j];
            const T *xr = (c < x0_rows) ? X0 + (size_t) c * ldx0 : X1 + (size_t) (c - x0_rows) * ldx1;
            T a[R];
            lds_block_vals<T, R>(sv + (size_t) j * R, a);
            #pragma unroll
            for (int u = 0; u < U; u++)
            {
                if (voff[u] < 0) continue;
                T x[VEC];
                xload<T, VEC>::ld(xr + voff[u], x);
                #pragma unroll
                for (int r = 0; r < R; r++)
                    #pragma unroll
                    for (int e = 0; e < VEC; e++) acc[r][u][e] = fma(a[r], x[e], acc[r][u][e]);
            }
        }
        __syncwarp();       // everybody is done with slot `buf` before it is refilled two chunks later
    }

    #pragma unroll
    for (int r = 0; r < R; r++)
    {
        T *crow = C + (size_t) (row0 + r) * ldc;
        #pragma unroll
        for (int u = 0; u < U; u++)
        {
            if (voff[u] < 0) continue;
            T out[VEC];
            if (beta == (T) 0)
            {
                #pragma unroll
                for (int e = 0; e < VEC; e++) out[e] = alpha * acc[r][u][e];
            } else {
                T old[VEC];
                #pragma unroll
                for (int e = 0; e < VEC; e++) old[e] = crow[voff[u] + e];
                #pragma unroll
                for (int e = 0; e < VEC; e++) out[e] = fma(alpha, acc[r][u][e], beta * old[e]);
            }
            xload<T, VEC>::st(crow + voff[u], out);
        }
    }
}

// MINB: 128-thread CTAs are meant to run 3 per SM (<= 170 registers); without the bound ptxas spends 188 and drops to 2
template <typename T, int VEC, int R, int U, int NB, int BS, int CB, int MINB = (BS == 128 ? 3 : 1)>
static void rg_launch_sv(const crp_rowgroup *rg, const T *bval, int nv, const T *X0, size_t ldx0, int x0_rows, const T *X1, size_t ldx1, T alpha, T beta, T *C, size_t ldc, cudaStream_t s)
{
    const unsigned blocks = (unsigned) ((rg->ngroups + BS / 32 - 1) / (BS / 32));
    const unsigned chunks = (unsigned) ((nv + 32 * U - 1) / (32 * U));
    if (blocks == 0) return;
    spmm_rowgroup_sv_kernel<T, VEC, R, U, NB, BS, CB, MINB><<<dim3(blocks, chunks), BS, 0, s>>>(
        rg->ngroups, rg->d_grow, rg->d_gptr, rg->d_bcol, bval, nv, X0, ldx0, x0_rows, X1, ldx1, alpha, beta, C, ldc);
    CRP_LAUNCH_CHECK();
}

template <typename T, int VEC, int R, int LPR, int U, int NB = 2, int BS = 256>
static void rg_launch_one(const crp_rowgroup *rg, const T *bval, int nv, const T *X0, size_t ldx0, int x0_rows, const T *X1, size_t ldx1, T alpha, T *C, size_t ldc, cudaStream_t s)
{
    constexpr int GW = 32 / LPR;
    const long long warps = ((long long) rg->ngroups + GW - 1) / GW;
    const unsigned blocks = (unsigned) ((warps + BS / 32 - 1) / (BS / 32));
    const unsigned chunks = (unsigned) ((nv + LPR * U - 1) / (LPR * U));
    if (blocks == 0) return;
    spmm_rowgroup_kernel<T, VEC, R, LPR, U, NB, BS><<<dim3(blocks, chunks), BS, 0, s>>>(
        rg->ngroups, rg->d_grow, rg->d_gptr, rg->d_bcol, bval, nv, X0, ldx0, x0_rows, X1, ldx1, alpha, C, ldc);
    CRP_LAUNCH_CHECK();
}

template <typename T, int VEC, int R>
static void rg_launch_R(const crp_rowgroup *rg, const T *bval, int nv, const T *X0, size_t ldx0, int x0_rows, const T *X1, size_t ldx1, T alpha, T beta, T *C, size_t ldc, cudaStream_t s)
{
    constexpr int UMAX = (R * VEC * (int) sizeof(T) <= 6 * 16) ? 4 : 2;       // keep the accumulator tile <= 96 registers
#define CRP_SV(U, NB, BS) rg_launch_sv<T, VEC, R, U, NB, BS, 32>(rg, bval, nv, X0, ldx0, x0_rows, X1, ldx1, alpha, beta, C, ldc, s)
#define CRP_RG(LPR, U) rg_launch_one<T, VEC, R, LPR, U>(rg, bval, nv, X0, ldx0, x0_rows, X1, ldx1, alpha, C, ldc, s)
    if (nv >= 128 && UMAX >= 4) CRP_SV(4, 2, 128);
    else if (nv >= 64)          CRP_SV(2, 2, 128);
    else if (nv > 16 || beta != (T) 0) CRP_SV(1, 4, 256);
    else if (nv > 8)            CRP_RG(16, 1);
    else if (nv > 4)            CRP_RG(8, 1);
    else                        CRP_RG(4, 1);
#undef CRP_RG
#undef CRP_SV
}

template <typename T, int VEC>
void crp_launch_rowgroup(const crp_rowgroup *rg, const T *bval, int nv, const T *X0, size_t ldx0, int x0_rows, const T *X1, size_t ldx1, T alpha, T beta, T *C, size_t ldc, cudaStream_t s)
{
#define CRP_RGR(R) rg_launch_R<T, VEC, R>(rg, bval, nv, X0, ldx0, x0_rows, X1, ldx1, alpha, beta, C, ldc, s)
    switch (rg->R)
    {
        case 2: CRP_RGR(2); break;
        case 3: CRP_RGR(3); break;
        case 4: CRP_RGR(4); break;
        case 6: CRP_RGR(6); break;
        case 8: CRP_RGR(8); break;
        default: fprintf(stderr, "[FATAL] crp_launch_rowgroup: unsupported group size %d\n", rg->R); abort();
    }
#undef CRP_RGR
}

template void crp_launch_rowgroup<double, 2>(const crp_rowgroup *, const double *, int, const double *, size_t, int, const double *, size_t, double, double, double *, size_t, cudaStream_t);
template void crp_launch_rowgroup<double, 1>(const crp_rowgroup *, const double *, int, const double *, size_t, int, const double *, size_t, double, double, double *, size_t, cudaStream_t);
template void crp_launch_rowgroup<float, 4>(const crp_rowgroup *, const float *, int, const float *, size_t, int, const float *, size_t, float, float, float *, size_t, cudaStream_t);
template void crp_launch_rowgroup<float, 1>(const crp_rowgroup *, const float *, int, const float *, size_t, int, const float *, size_t, float, float, float *, size_t, cudaStream_t);
