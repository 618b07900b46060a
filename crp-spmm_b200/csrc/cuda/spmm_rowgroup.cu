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

// ------------------------------------------------------------------ plan-time analysis (host)

static const int kCandR[] = { 8, 6, 4, 3, 2 };

// Row i "continues" row i - 1 when both have exactly the same column list; a group of R rows starting
// at r0 is usable ("perfect") when rows r0 + 1 .. r0 + R - 1 all continue their predecessor, the list is
// non-empty and strictly increasing.  Groups start at row `off` + multiples of R: the first local row of a
// rank is generally not aligned with the matrix's own block structure, so every offset is tried.
// Cost unit: L1 wavefronts per 64 B of C row (4 per B-row load + 1 per row FMA'd), see DESIGN.md.
struct rg_rowinfo
{
    std::vector<unsigned char> cont;    // row i has the same columns as row i - 1
    std::vector<unsigned char> incr;    // row i's columns are strictly increasing and the row is not empty
};

static void rg_scan_rows(const int m, const int *rowptr, const int *colidx, rg_rowinfo *ri)
{
    ri->cont.assign((size_t) m, 0);
    ri->incr.assign((size_t) m, 0);
    for (int i = 0; i < m; i++)
    {
        const int b = rowptr[i], len = rowptr[i + 1] - b;
        bool inc = len > 0;
        for (int p = b + 1; inc && p < b + len; p++) if (colidx[p] <= colidx[p - 1]) inc = false;
        ri->incr[(size_t) i] = inc;
        if (i > 0 && len > 0 && rowptr[i] - rowptr[i - 1] == len)
            ri->cont[(size_t) i] = (memcmp(colidx + b, colidx + rowptr[i - 1], sizeof(int) * (size_t) len) == 0);
    }
}

static inline bool rg_group_ok(const rg_rowinfo &ri, const int r0, const int R, const int m)
{
    if (r0 + R > m || !ri.incr[(size_t) r0]) return false;
    for (int r = r0 + 1; r < r0 + R; r++) if (!ri.cont[(size_t) r]) return false;
    return true;
}

static double analyse_R(const int m, const int *rowptr, const rg_rowinfo &ri, const int R, const int off, long long *nblk_out, long long *rest_nnz_out)
{
    long long nblk = 0;
    for (int r0 = off; r0 + R <= m; r0 += R)
        if (rg_group_ok(ri, r0, R, m)) nblk += rowptr[r0 + 1] - rowptr[r0];
    const long long rest = (long long) rowptr[m] - nblk * R;
    *nblk_out = nblk;
    *rest_nnz_out = rest;
    return (double) nblk * (4.0 + R) + (double) rest * 5.0;
}

// the group size and alignment with the lowest modelled cost (R = 1: keep the row-split kernel); forced: CRP_SPMM_ROWGROUP_R
static void rg_choose(const int m, const int *rowptr, const rg_rowinfo &ri, const int forced, int *best_R_, int *best_off_, long long *nblk_)
{
    double best_cost = (double) rowptr[m] * 5.0;      // everything in the row-split kernel
    int best_R = 1, best_off = 0;
    long long best_nblk = 0;
    for (int R : kCandR)
    {
        if (forced > 1 && R != forced) continue;
        for (int off = 0; off < R && off < m; off++)
        {
            long long nblk, rest;
            const double cost = analyse_R(m, rowptr, ri, R, off, &nblk, &rest);
            const bool take = (forced > 1) ? (nblk > 0 && (best_R == 1 || cost < best_cost)) : (cost < 0.9 * best_cost || (best_R == R && cost < best_cost));
            if (take) { best_cost = cost; best_R = R; best_off = off; best_nblk = nblk; }
        }
    }
    *best_R_ = best_R;  *best_off_ = best_off;  *nblk_ = best_nblk;
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
        rg_rowinfo ri;
        rg_scan_rows(m, rowptr, colidx, &ri);
        rg_choose(m, rowptr, ri, forced, R, offset, nblk);
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
    rg_rowinfo ri;
    rg_scan_rows(m, rowptr, colidx, &ri);
    int best_R = 1, best_off = 0;
    long long best_nblk = 0;
    rg_choose(m, rowptr, ri, forced, &best_R, &best_off, &best_nblk);
    if (best_R == 1) return;

    const int R = best_R;
    std::vector<int> g_row, g_ptr, b_col, rest_rows;
    std::vector<double> b_val;
    g_ptr.push_back(0);
    long long rest = 0;
    for (int r = 0; r < best_off && r < m; r++) { rest_rows.push_back(r); rest += rowptr[r + 1] - rowptr[r]; }
    for (int r0 = best_off; r0 < m; r0 += R)
    {
        const int r1 = std::min(m, r0 + R);
        if (!rg_group_ok(ri, r0, R, m))
        {
            for (int r = r0; r < r1; r++) { rest_rows.push_back(r); rest += rowptr[r + 1] - rowptr[r]; }
            continue;
        }
        g_row.push_back(r0);
        const int len = rowptr[r0 + 1] - rowptr[r0];
        for (int j = 0; j < len; j++)
        {
            b_col.push_back(colidx[rowptr[r0] + j]);
            for (int r = 0; r < R; r++) b_val.push_back(val[rowptr[r0 + r] + j]);
        }
        g_ptr.push_back((int) b_col.size());
    }
    rg->R = R;
    rg->ngroups = (int) g_row.size();
    rg->nblk = (long long) b_col.size();
    rg->nrest = (int) rest_rows.size();
    rg->rest_nnz = rest;
    auto upload = [](const void *src, size_t bytes) -> void * {
        void *d = NULL;
        if (bytes == 0) return d;
        CRP_CUDA_CHECK(cudaMalloc(&d, bytes));
        CRP_CUDA_CHECK(cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice));
        return d;
    };
    rg->d_grow = (int *) upload(g_row.data(), sizeof(int) * g_row.size());
    rg->d_gptr = (int *) upload(g_ptr.data(), sizeof(int) * g_ptr.size());
    rg->d_bcol = (int *) upload(b_col.data(), sizeof(int) * b_col.size());
    rg->d_bval = (double *) upload(b_val.data(), sizeof(double) * b_val.size());
    rg->d_rest = (int *) upload(rest_rows.data(), sizeof(int) * rest_rows.size());
    if (rest_out) rest_out->swap(rest_rows);
}

void crp_rowgroup_destroy(crp_spmm_plan *plan)
{
    crp_rowgroup *rg = &plan->rg;
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

// one stage of the software pipeline: NB blocks' worth of B-row segments and values
template <typename T, int VEC, int R, int U, int NB>
struct rg_stage
{
    T x[NB][U][VEC];
    T a[NB][R];
};

template <typename T, int VEC, int R, int U, int NB>
__device__ __forceinline__ void rg_load_stage(
    rg_stage<T, VEC, R, U, NB> &st, const int (&c)[NB], const int p, const T *__restrict__ bval, const int (&voff)[U],
    const T *__restrict__ X0, const size_t ldx0, const int x0_rows, const T *__restrict__ X1, const size_t ldx1
)
{
    #pragma unroll
    for (int q = 0; q < NB; q++)
    {
        const T *xr = (c[q] < x0_rows) ? X0 + (size_t) c[q] * ldx0 : X1 + (size_t) (c[q] - x0_rows) * ldx1;
        #pragma unroll
        for (int u = 0; u < U; u++)
        {
            if (voff[u] >= 0) xload<T, VEC>::ld(xr + voff[u], st.x[q][u]);
            else { for (int e = 0; e < VEC; e++) st.x[q][u][e] = (T) 0; }
        }
        load_block_vals<T, R>(bval + (size_t) (p + q) * R, st.a[q]);
    }
}

template <typename T, int VEC, int R, int U, int NB>
__device__ __forceinline__ void rg_fma_stage(const rg_stage<T, VEC, R, U, NB> &st, T (&acc)[R][U][VEC])
{
    #pragma unroll
    for (int q = 0; q < NB; q++)
        #pragma unroll
        for (int r = 0; r < R; r++)
            #pragma unroll
            for (int u = 0; u < U; u++)
                #pragma unroll
                for (int e = 0; e < VEC; e++) acc[r][u][e] = fma(st.a[q][r], st.x[q][u][e], acc[r][u][e]);
}

// PIPE = 1: the loads of the next NB blocks are issued before the FMAs of the current ones (register
// double buffering).  PF > 0: the B rows of the blocks PF positions ahead are prefetched into L1
// (CCTL.E.PF1, one 128-byte line per lane), which keeps gathers in flight without holding registers.
template <typename T, int VEC, int R, int LPR, int U, int NB, int PIPE, int PF, int BS>
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

    // L1 prefetch: lane l covers bytes [128 l, 128 l + 128) of the CTA's column chunk of a B row
    constexpr int CHUNK_BYTES = LPR * U * VEC * (int) sizeof(T);
    const int pf_off = (v0 * VEC * (int) sizeof(T)) + l * 128;
    const bool pf_lane = (PF > 0) && (l * 128 < CHUNK_BYTES) && (v0 * VEC + l * (128 / (int) sizeof(T)) < nv * VEC);
    auto prefetch_row = [&](const int pp) {
        if (PF > 0 && pp < p_end && pf_lane)
        {
            const int c = __ldg(bcol + pp);
            const char *xr = (c < x0_rows) ? (const char *) (X0 + (size_t) c * ldx0) : (const char *) (X1 + (size_t) (c - x0_rows) * ldx1);
            asm volatile("prefetch.global.L1 [%0];" :: "l"(xr + pf_off));
        }
    };
    if (PF > 0)
    {
        #pragma unroll 1
        for (int d = 0; d < PF; d++) prefetch_row(p + d);
    }

    typedef rg_stage<T, VEC, R, U, NB> stage_t;
    int cn[NB];
    #pragma unroll
    for (int q = 0; q < NB; q++) cn[q] = (p + q < p_end) ? __ldg(bcol + p + q) : 0;

    if (PIPE == 0)
    {
        for (; p + NB <= p_end; p += NB)
        {
            int c[NB];
            #pragma unroll
            for (int q = 0; q < NB; q++) c[q] = cn[q];
            #pragma unroll
            for (int q = 0; q < NB; q++) cn[q] = (p + NB + q < p_end) ? __ldg(bcol + p + NB + q) : 0;
            #pragma unroll
            for (int q = 0; q < NB; q++) prefetch_row(p + PF + q);
            stage_t st;
            rg_load_stage<T, VEC, R, U, NB>(st, c, p, bval, voff, X0, ldx0, x0_rows, X1, ldx1);
            rg_fma_stage<T, VEC, R, U, NB>(st, acc);
        }
    } else {
        stage_t sa, sb;
        bool have_a = false;
        if (p + NB <= p_end)
        {
            rg_load_stage<T, VEC, R, U, NB>(sa, cn, p, bval, voff, X0, ldx0, x0_rows, X1, ldx1);
            have_a = true;
            #pragma unroll
            for (int q = 0; q < NB; q++) cn[q] = (p + NB + q < p_end) ? __ldg(bcol + p + NB + q) : 0;
        }
        // invariant: sa holds blocks [p, p + NB), cn the columns of [p + NB, p + 2 NB)
        while (have_a)
        {
            const bool next_b = (p + 2 * NB <= p_end);
            if (next_b)
            {
                rg_load_stage<T, VEC, R, U, NB>(sb, cn, p + NB, bval, voff, X0, ldx0, x0_rows, X1, ldx1);
                #pragma unroll
                for (int q = 0; q < NB; q++) cn[q] = (p + 2 * NB + q < p_end) ? __ldg(bcol + p + 2 * NB + q) : 0;
            }
            #pragma unroll
            for (int q = 0; q < NB; q++) prefetch_row(p + PF + q);
            rg_fma_stage<T, VEC, R, U, NB>(sa, acc);
            p += NB;
            if (!next_b) break;
            const bool next_a = (p + 2 * NB <= p_end);
            if (next_a)
            {
                rg_load_stage<T, VEC, R, U, NB>(sa, cn, p + NB, bval, voff, X0, ldx0, x0_rows, X1, ldx1);
                #pragma unroll
                for (int q = 0; q < NB; q++) cn[q] = (p + 2 * NB + q < p_end) ? __ldg(bcol + p + 2 * NB + q) : 0;
            }
            #pragma unroll
            for (int q = 0; q < NB; q++) prefetch_row(p + PF + q);
            rg_fma_stage<T, VEC, R, U, NB>(sb, acc);
            p += NB;
            have_a = next_a;
        }
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
// In the kernel above every iteration waits for the block values: they stream from HBM (first
// touch, no reuse), so each NB-block step pays a full DRAM round trip (~1400 cycles, ncu:
// long_scoreboard 2.9 of 4.0 stall slots) before its FMAs.  Here a warp copies the values and
// columns of the next CB blocks of its group into its private shared-memory ring with cp.async
// (LDGSTS, no registers held) while it works on the current CB blocks, and reads them back with
// broadcast LDS (29 cycles).  Only the B-row gathers remain long-latency loads.
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

template <typename T, int VEC, int R, int U, int NB, int BS, int CB, int PIPE, int PF, int MINB>
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

    if constexpr (PIPE == 2)
    {
        // Rolling prefetch: the registers of an (block, column-group) slot are reloaded with the data of the
        // same slot NB blocks ahead right after the FMAs that consumed them, so the gathers of the next step
        // are in flight during the whole FMA phase of the current one - without a second set of registers.
        static_assert(CB % NB == 0, "chunk size must be a multiple of the blocks per step");
        const int total = p_end - p_beg;
        const int nchunk = (total + CB - 1) / CB;
        stage(0, p_beg, min(CB, total));
        cp_async_wait_all();
        __syncwarp();
        if (nchunk > 1) stage(1, p_beg + CB, min(CB, total - CB));
        auto col_at = [&](const int jj) -> int { return s_col[wib][(jj / CB) & 1][jj % CB]; };
        auto row_at = [&](const int jj) -> const T * {
            const int c = col_at(jj);
            return (c < x0_rows) ? X0 + (size_t) c * ldx0 : X1 + (size_t) (c - x0_rows) * ldx1;
        };
        T x[NB][U][VEC];
        const int nfull = total - total % NB;
        if (nfull > 0)
        {
            #pragma unroll
            for (int q = 0; q < NB; q++)
            {
                const T *xr = row_at(q);
                #pragma unroll
                for (int u = 0; u < U; u++)
                {
                    if (voff[u] >= 0) xload<T, VEC>::ld(xr + voff[u], x[q][u]);
                    else { for (int e = 0; e < VEC; e++) x[q][u][e] = (T) 0; }
                }
            }
        }
        for (int j = 0; j < nfull; j += NB)
        {
            const int chunk = j / CB;
            const bool last_of_chunk = ((j % CB) == CB - NB);
            if (last_of_chunk && chunk + 1 < nchunk) { cp_async_wait_all(); __syncwarp(); }     // next chunk readable for the look-ahead
            const T *sv = reinterpret_cast<const T *>(&s_val[wib][chunk & 1][0]);
            #pragma unroll
            for (int q = 0; q < NB; q++)
            {
                T a[R];
                lds_block_vals<T, R>(sv + (size_t) ((j % CB) + q) * R, a);
                const int jn = j + NB + q;                  // the block that takes this slot next
                const T *xn = row_at(jn < nfull ? jn : nfull - 1);
                #pragma unroll
                for (int u = 0; u < U; u++)
                {
                    #pragma unroll
                    for (int r = 0; r < R; r++)
                        #pragma unroll
                        for (int e = 0; e < VEC; e++) acc[r][u][e] = fma(a[r], x[q][u][e], acc[r][u][e]);
                    if (voff[u] >= 0) xload<T, VEC>::ld(xn + voff[u], x[q][u]);
                }
            }
            if (last_of_chunk)
            {
                __syncwarp();                               // everybody is done with this chunk's buffer
                if (chunk + 2 < nchunk) stage(chunk & 1, p_beg + (chunk + 2) * CB, min(CB, total - (chunk + 2) * CB));
            }
        }
        if (nfull < total)                                  // fewer than NB blocks left
        {
            const int chunk = nfull / CB;
            if (nfull > 0 && (nfull % CB) == 0) { cp_async_wait_all(); __syncwarp(); }
            const T *sv = reinterpret_cast<const T *>(&s_val[wib][chunk & 1][0]);
            for (int j = nfull; j < total; j++)
            {
                const T *xr = row_at(j);
                T a[R];
                lds_block_vals<T, R>(sv + (size_t) (j % CB) * R, a);
                #pragma unroll
                for (int u = 0; u < U; u++)
                {
                    if (voff[u] < 0) continue;
                    T xx[VEC];
                    xload<T, VEC>::ld(xr + voff[u], xx);
                    #pragma unroll
                    for (int r = 0; r < R; r++)
                        #pragma unroll
                        for (int e = 0; e < VEC; e++) acc[r][u][e] = fma(a[r], xx[e], acc[r][u][e]);
                }
            }
        }
    }

    int buf = 0;
    if (PIPE != 2) stage(0, p_beg, min(CB, p_end - p_beg));
    for (int pc = p_beg; PIPE != 2 && pc < p_end; pc += CB, buf ^= 1)
    {
        const int nb = min(CB, p_end - pc);
        cp_async_wait_all();
        __syncwarp();
        if (pc + CB < p_end) stage(buf ^ 1, pc + CB, min(CB, p_end - pc - CB));
        const T *sv = reinterpret_cast<const T *>(&s_val[wib][buf][0]);
        const int *sc = &s_col[wib][buf][0];
        int j = 0;
        auto row_of = [&](const int jj) -> const T * {
            const int c = sc[jj];
            return (c < x0_rows) ? X0 + (size_t) c * ldx0 : X1 + (size_t) (c - x0_rows) * ldx1;
        };
        auto load_x = [&](const int jj, T (&x)[NB][U][VEC]) {
            #pragma unroll
            for (int q = 0; q < NB; q++)
            {
                const T *xr = row_of(jj + q);
                #pragma unroll
                for (int u = 0; u < U; u++)
                {
                    if (voff[u] >= 0) xload<T, VEC>::ld(xr + voff[u], x[q][u]);
                    else { for (int e = 0; e < VEC; e++) x[q][u][e] = (T) 0; }
                }
            }
        };
        auto fma_x = [&](const int jj, const T (&x)[NB][U][VEC]) {
            #pragma unroll
            for (int q = 0; q < NB; q++)
            {
                T a[R];
                lds_block_vals<T, R>(sv + (size_t) (jj + q) * R, a);
                #pragma unroll
                for (int r = 0; r < R; r++)
                    #pragma unroll
                    for (int u = 0; u < U; u++)
                        #pragma unroll
                        for (int e = 0; e < VEC; e++) acc[r][u][e] = fma(a[r], x[q][u][e], acc[r][u][e]);
            }
        };
        // L1 prefetch of the B rows PF blocks ahead (lane l covers the l-th 128-byte line of the column chunk)
        constexpr int CHUNK_BYTES = 32 * U * VEC * (int) sizeof(T);
        const bool pf_lane = (PF > 0) && (lane * 128 < CHUNK_BYTES) && (v0 * VEC + lane * (128 / (int) sizeof(T)) < nv * VEC);
        auto prefetch = [&](const int jj) {
            if (PF > 0 && jj < nb && pf_lane)
                asm volatile("prefetch.global.L1 [%0];" :: "l"((const char *) row_of(jj) + (size_t) v0 * VEC * sizeof(T) + lane * 128));
        };
        if (PF > 0) { for (int d = 0; d < PF; d++) prefetch(d); }
        if (PIPE == 0)
        {
            for (; j + NB <= nb; j += NB)
            {
                T x[NB][U][VEC];
                #pragma unroll
                for (int q = 0; q < NB; q++) prefetch(j + PF + q);
                load_x(j, x);
                fma_x(j, x);
            }
        } else {
            T xa[NB][U][VEC], xb[NB][U][VEC];
            if (j + NB <= nb) load_x(j, xa);
            while (j + NB <= nb)
            {
                const bool nb1 = (j + 2 * NB <= nb);
                if (nb1) load_x(j + NB, xb);
                fma_x(j, xa);
                j += NB;
                if (!nb1) break;
                const bool nb2 = (j + 2 * NB <= nb);
                if (nb2) load_x(j + NB, xa);
                fma_x(j, xb);
                j += NB;
                if (!nb2) break;
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
template <typename T, int VEC, int R, int U, int NB, int BS, int CB, int PIPE = 0, int PF = 0, int MINB = (BS == 128 ? 3 : 1)>
static void rg_launch_sv(const crp_rowgroup *rg, const T *bval, int nv, const T *X0, size_t ldx0, int x0_rows, const T *X1, size_t ldx1, T alpha, T beta, T *C, size_t ldc, cudaStream_t s)
{
    const unsigned blocks = (unsigned) ((rg->ngroups + BS / 32 - 1) / (BS / 32));
    const unsigned chunks = (unsigned) ((nv + 32 * U - 1) / (32 * U));
    if (blocks == 0) return;
    spmm_rowgroup_sv_kernel<T, VEC, R, U, NB, BS, CB, PIPE, PF, MINB><<<dim3(blocks, chunks), BS, 0, s>>>(
        rg->ngroups, rg->d_grow, rg->d_gptr, rg->d_bcol, bval, nv, X0, ldx0, x0_rows, X1, ldx1, alpha, beta, C, ldc);
    CRP_LAUNCH_CHECK();
}

template <typename T, int VEC, int R, int LPR, int U, int NB = 2, int PIPE = 0, int PF = 0, int BS = 256>
static void rg_launch_one(const crp_rowgroup *rg, const T *bval, int nv, const T *X0, size_t ldx0, int x0_rows, const T *X1, size_t ldx1, T alpha, T *C, size_t ldc, cudaStream_t s)
{
    constexpr int GW = 32 / LPR;
    const long long warps = ((long long) rg->ngroups + GW - 1) / GW;
    const unsigned blocks = (unsigned) ((warps + BS / 32 - 1) / (BS / 32));
    const unsigned chunks = (unsigned) ((nv + LPR * U - 1) / (LPR * U));
    if (blocks == 0) return;
    spmm_rowgroup_kernel<T, VEC, R, LPR, U, NB, PIPE, PF, BS><<<dim3(blocks, chunks), BS, 0, s>>>(
        rg->ngroups, rg->d_grow, rg->d_gptr, rg->d_bcol, bval, nv, X0, ldx0, x0_rows, X1, ldx1, alpha, C, ldc);
    CRP_LAUNCH_CHECK();
}

bool crp_launch_rowgroup_x(
    const int cfg, const crp_rowgroup *rg, const double *bval, const int n, const double *X0, const size_t ldx0,
    const double *X1, const double alpha, const double beta, double *C, const size_t ldc, cudaStream_t s
);

// development switch (CRP_SPMM_RG_CFG = index): fp64, 128-bit, R = 6, full-warp groups only.
// Indices 0..49 are the configurations measured in round 1 (compiled only with -DCRP_DEV_SWEEP); 60.. are the lean variants of
// spmm_rowgroup_x.cu.  Without the variable, or for an index that is not compiled in, the shipped kernel runs.
template <typename T, int VEC, int R>
static bool rg_launch_experiment(const crp_rowgroup *rg, const T *bval, int nv, const T *X0, size_t ldx0, int x0_rows, const T *X1, size_t ldx1, T alpha, T *C, size_t ldc, cudaStream_t s)
{
    if constexpr (sizeof(T) == 8 && VEC == 2 && R == 6)
    {
        const char *e = getenv("CRP_SPMM_RG_CFG");
        if (e == NULL || nv < 128) return false;
        if (atoi(e) >= 60) return crp_launch_rowgroup_x(atoi(e), rg, bval, nv * VEC, X0, ldx0, X1, alpha, 0.0, C, ldc, s);   // spmm_rowgroup_x.cu
#ifdef CRP_DEV_SWEEP     /* the round-1 sweep (profiles/kbench*.log): build with CRP_NVCC_EXTRA=-DCRP_DEV_SWEEP to get these ~40 instantiations */
#define CRP_RGX(U, NB, PIPE, PF, BS) rg_launch_one<T, VEC, R, 32, U, NB, PIPE, PF, BS>(rg, bval, nv, X0, ldx0, x0_rows, X1, ldx1, alpha, C, ldc, s); return true
        switch (atoi(e))
        {
            case 0:  CRP_RGX(4, 2, 0, 0, 256);
            case 1:  CRP_RGX(4, 2, 0, 0, 128);
            case 2:  CRP_RGX(2, 2, 0, 0, 256);
            case 3:  CRP_RGX(2, 4, 0, 0, 256);
            case 4:  CRP_RGX(4, 1, 1, 0, 128);
            case 5:  CRP_RGX(4, 2, 0, 8, 128);
            case 6:  CRP_RGX(2, 2, 1, 0, 128);
            case 7:  CRP_RGX(2, 2, 0, 8, 256);
            case 8:  CRP_RGX(1, 4, 0, 0, 256);
            case 9:  CRP_RGX(4, 1, 1, 8, 128);
            case 10: CRP_RGX(2, 4, 0, 16, 256);
            case 11: CRP_RGX(1, 4, 0, 16, 256);
            case 12: CRP_RGX(4, 1, 0, 8, 128);
            case 13: CRP_RGX(2, 1, 1, 0, 256);
#define CRP_RGSV(U, NB, BS, CB) rg_launch_sv<T, VEC, R, U, NB, BS, CB>(rg, bval, nv, X0, ldx0, x0_rows, X1, ldx1, alpha, (T) 0, C, ldc, s); return true
            case 20: CRP_RGSV(4, 2, 128, 32);
            case 21: CRP_RGSV(4, 2, 256, 32);
            case 22: CRP_RGSV(4, 1, 128, 32);
            case 23: CRP_RGSV(4, 4, 128, 32);
            case 24: CRP_RGSV(2, 2, 128, 32);
            case 25: CRP_RGSV(2, 4, 128, 32);
            case 26: CRP_RGSV(2, 4, 256, 32);
            case 27: CRP_RGSV(4, 2, 128, 64);
            case 28: CRP_RGSV(4, 3, 128, 32);
            case 29: CRP_RGSV(1, 4, 256, 32);
#define CRP_RGSV2(U, NB, BS, CB, PIPE, PF) rg_launch_sv<T, VEC, R, U, NB, BS, CB, PIPE, PF>(rg, bval, nv, X0, ldx0, x0_rows, X1, ldx1, alpha, (T) 0, C, ldc, s); return true
            case 30: CRP_RGSV2(4, 1, 128, 32, 1, 0);
            case 31: CRP_RGSV2(4, 2, 128, 32, 1, 0);
            case 32: CRP_RGSV2(2, 2, 128, 32, 1, 0);
            case 33: CRP_RGSV2(4, 2, 128, 32, 0, 4);
            case 34: CRP_RGSV2(4, 2, 128, 32, 0, 8);
            case 35: CRP_RGSV2(2, 2, 128, 32, 0, 8);
            case 36: CRP_RGSV2(4, 1, 128, 32, 0, 8);
            case 37: CRP_RGSV2(2, 1, 128, 32, 1, 0);
            case 38: CRP_RGSV2(4, 2, 64, 32, 0, 0);
            case 39: CRP_RGSV2(4, 2, 96, 32, 0, 0);
            case 40: CRP_RGSV2(4, 2, 128, 32, 2, 0);
            case 41: CRP_RGSV2(4, 1, 128, 32, 2, 0);
            case 42: CRP_RGSV2(2, 2, 128, 32, 2, 0);
            case 43: CRP_RGSV2(2, 4, 128, 32, 2, 0);
            case 44: CRP_RGSV2(4, 4, 128, 32, 2, 0);
            case 45: CRP_RGSV2(2, 2, 256, 32, 2, 0);
#define CRP_RGSV3(U, NB, BS, CB, MINB) rg_launch_sv<T, VEC, R, U, NB, BS, CB, 0, 0, MINB>(rg, bval, nv, X0, ldx0, x0_rows, X1, ldx1, alpha, (T) 0, C, ldc, s); return true
            case 46: CRP_RGSV3(4, 2, 128, 32, 4);
            case 47: CRP_RGSV3(4, 1, 128, 32, 4);
            case 48: CRP_RGSV3(4, 2, 64, 32, 7);
            case 49: CRP_RGSV3(4, 1, 64, 32, 8);
#undef CRP_RGSV3
#undef CRP_RGSV2
#undef CRP_RGSV
            default: return false;
        }
#undef CRP_RGX
#endif
    }
    return false;
}

template <typename T, int VEC, int R>
static void rg_launch_R(const crp_rowgroup *rg, const T *bval, int nv, const T *X0, size_t ldx0, int x0_rows, const T *X1, size_t ldx1, T alpha, T beta, T *C, size_t ldc, cudaStream_t s)
{
    if (beta == (T) 0 && rg_launch_experiment<T, VEC, R>(rg, bval, nv, X0, ldx0, x0_rows, X1, ldx1, alpha, C, ldc, s)) return;
    constexpr int UMAX = (R * VEC * (int) sizeof(T) <= 6 * 16) ? 4 : 2;       // keep the accumulator tile <= 96 registers
    // Full-warp groups: values / columns staged through shared memory (measured on the pwtk-shaped
    // n = 256 fp64 case: 0.60 -> 0.42 ms); 128-thread CTAs so that 168 registers still give 12 warps per SM.
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
