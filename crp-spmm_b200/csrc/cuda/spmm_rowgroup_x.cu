// spmm_rowgroup_x.cu - EXPERIMENTAL lean variants of the row-group kernel (fp64, 128-bit, full column chunks only).
//
// NOT on the default path: reached only through the development switch CRP_SPMM_RG_CFG = 60..64
// (spmm_rowgroup.cu, rg_launch_experiment).  Written at the end of round 1, after the GPU budget was spent, from the
// SASS of the shipped kernel (spmm_rowgroup_sv_kernel<double,2,6,4,2,128,32>): its hot loop issues 203 instructions per
// 96 DFMA - 16 CS2R (zero fill for column groups beyond n), 11 ISETP / 4 SEL / 4 BRA (bounds and X0 / X1 selection),
// 29 IMAD + 8 LEA (one 64-bit address per 128-bit load), 9 LDC / LDCU (kernel parameters re-read every step).
// These variants remove what the launch conditions make unnecessary:
//   - every lane's column groups are inside n  (n * 8 bytes is a multiple of the 32 * U * 16-byte chunk),
//   - the dense operand is one piece (X1 unused: single rank, or a plan without received rows),
//   - one 64-bit address per block, the U loads use immediate offsets (u * 512 bytes),
// and (PERSIST) let a warp walk over several groups with the next group's first value chunk staged while the current
// group finishes, so the per-group start-up (row pointers -> cp.async -> first gathers, ~3000 cycles of ~45000) is hidden
// and the last partial wave disappears.
// They must be validated against the row-split result (tools/kbench.py --check) before they replace anything.
#include "crp_cuda_internal.cuh"

namespace {

__device__ __forceinline__ void cpa16(void *smem, const void *gmem)
{
    const unsigned s = (unsigned) __cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cpa4(void *smem, const void *gmem)
{
    const unsigned s = (unsigned) __cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cpa_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// R even (16-byte value rows), U 128-bit column groups per lane, NB blocks per step, CB blocks per staged chunk
template <int R, int U, int NB, int BS, int CB, bool PERSIST>
__global__ void __launch_bounds__(BS, (BS == 128 ? 3 : 1)) spmm_rowgroup_x_kernel(
    const int ngroups, const int *__restrict__ grow, const int *__restrict__ gptr,
    const int *__restrict__ bcol, const double *__restrict__ bval,
    const double2 *__restrict__ X, const size_t ldx2,          // leading dimension in double2 units
    const double alpha, double2 *__restrict__ C, const size_t ldc2
)
{
    static_assert(R % 2 == 0 && CB % NB == 0, "layout assumptions");
    constexpr int WPB = BS / 32;
    __shared__ __align__(16) double s_val[WPB][2][CB * R];
    __shared__ int s_col[WPB][2][CB];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warp0 = blockIdx.x * WPB + wib;
    const int nwarp = gridDim.x * WPB;
    const double2 *xlane = X + (size_t) blockIdx.y * (32 * U) + lane;
    double2 *clane = C + (size_t) blockIdx.y * (32 * U) + lane;

    auto stage = [&](const int buf, const int pc, const int nb) {
        const char *src = (const char *) (bval + (size_t) pc * R);
        const int nbytes = nb * R * 8;
        for (int off = lane * 16; off < nbytes; off += 32 * 16) cpa16((char *) &s_val[wib][buf][0] + off, src + off);
        for (int j = lane; j < nb; j += 32) cpa4(&s_col[wib][buf][j], bcol + pc + j);
        cpa_commit();
    };

    int g = warp0;
    if (g >= ngroups) return;
    int p_beg = __ldg(gptr + g), p_end = __ldg(gptr + g + 1), row0 = __ldg(grow + g);
    int buf = 0;                                    // ring slot that holds (or will hold) the current chunk
    stage(buf, p_beg, min(CB, p_end - p_beg));
    for (;;)
    {
        double2 acc[R][U];
        #pragma unroll
        for (int r = 0; r < R; r++)
            #pragma unroll
            for (int u = 0; u < U; u++) acc[r][u] = make_double2(0.0, 0.0);

        // next group of this warp (PERSIST): its row data is fetched early, its first chunk is staged during the last chunk
        const int gn = PERSIST ? g + nwarp : ngroups;
        int n_beg = 0, n_end = 0, n_row0 = 0;
        if (gn < ngroups) { n_beg = __ldg(gptr + gn); n_end = __ldg(gptr + gn + 1); n_row0 = __ldg(grow + gn); }

        for (int pc = p_beg; pc < p_end; pc += CB, buf ^= 1)
        {
            const int nb = min(CB, p_end - pc);
            cpa_wait_all();
            __syncwarp();
            if (pc + CB < p_end) stage(buf ^ 1, pc + CB, min(CB, p_end - pc - CB));
            else if (gn < ngroups) stage(buf ^ 1, n_beg, min(CB, n_end - n_beg));
            const double *sv = &s_val[wib][buf][0];
            const int *sc = &s_col[wib][buf][0];
            int j = 0;
            for (; j + NB <= nb; j += NB)
            {
                double2 x[NB][U];
                #pragma unroll
                for (int q = 0; q < NB; q++)
                {
                    const double2 *xr = xlane + (size_t) sc[j + q] * ldx2;
                    #pragma unroll
                    for (int u = 0; u < U; u++) x[q][u] = __ldg(xr + u * 32);
                }
                #pragma unroll
                for (int q = 0; q < NB; q++)
                {
                    double a[R];
                    #pragma unroll
                    for (int i = 0; i < R / 2; i++)
                    {
                        const double2 t = reinterpret_cast<const double2 *>(sv + (size_t) (j + q) * R)[i];
                        a[2 * i] = t.x; a[2 * i + 1] = t.y;
                    }
                    #pragma unroll
                    for (int r = 0; r < R; r++)
                        #pragma unroll
                        for (int u = 0; u < U; u++)
                        {
                            acc[r][u].x = fma(a[r], x[q][u].x, acc[r][u].x);
                            acc[r][u].y = fma(a[r], x[q][u].y, acc[r][u].y);
                        }
                }
            }
            for (; j < nb; j++)
            {
                const double2 *xr = xlane + (size_t) sc[j] * ldx2;
                double a[R];
                #pragma unroll
                for (int i = 0; i < R / 2; i++)
                {
                    const double2 t = reinterpret_cast<const double2 *>(sv + (size_t) j * R)[i];
                    a[2 * i] = t.x; a[2 * i + 1] = t.y;
                }
                #pragma unroll
                for (int u = 0; u < U; u++)
                {
                    const double2 xx = __ldg(xr + u * 32);
                    #pragma unroll
                    for (int r = 0; r < R; r++)
                    {
                        acc[r][u].x = fma(a[r], xx.x, acc[r][u].x);
                        acc[r][u].y = fma(a[r], xx.y, acc[r][u].y);
                    }
                }
            }
            __syncwarp();
        }
        // after the chunk loop `buf` names the slot of the chunk staged last = the next group's first chunk (if any)
        #pragma unroll
        for (int r = 0; r < R; r++)
        {
            double2 *crow = clane + (size_t) (row0 + r) * ldc2;
            #pragma unroll
            for (int u = 0; u < U; u++) __stcs(crow + u * 32, make_double2(alpha * acc[r][u].x, alpha * acc[r][u].y));
        }
        if (gn >= ngroups) break;
        g = gn; p_beg = n_beg; p_end = n_end; row0 = n_row0;
    }
}

// Two warps per group ("split-2"): each takes one half of the group's blocks, the second warp hands its partial sums over
// through shared memory and the first adds them (first half + second half: a fixed order, but not the row-split kernel's
// left-to-right order - results agree to rounding, not bit for bit).  Meant for small per-rank problems (8 GPUs: 4540
// groups on 1776 resident warps = 2.6 waves of 27-step dependency chains): twice the warps, half the chain length.
template <int R, int U, int NB, int CB>
__global__ void __launch_bounds__(128, 3) spmm_rowgroup_x2_kernel(
    const int ngroups, const int *__restrict__ grow, const int *__restrict__ gptr,
    const int *__restrict__ bcol, const double *__restrict__ bval,
    const double2 *__restrict__ X, const size_t ldx2, const double alpha, double2 *__restrict__ C, const size_t ldc2
)
{
    static_assert(R % 2 == 0 && CB % NB == 0, "layout assumptions");
    __shared__ __align__(16) double s_val[4][2][CB * R];
    __shared__ int s_col[4][2][CB];
    __shared__ double2 s_part[2][R][U][32];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pair = wib >> 1, half = wib & 1;
    const int g = blockIdx.x * 2 + pair;
    const bool valid = g < ngroups;
    const double2 *xlane = X + (size_t) blockIdx.y * (32 * U) + lane;
    double2 *clane = C + (size_t) blockIdx.y * (32 * U) + lane;
    int p_beg = 0, p_end = 0, row0 = 0;
    if (valid)
    {
        const int b = __ldg(gptr + g), e = __ldg(gptr + g + 1);
        row0 = __ldg(grow + g);
        int mid = b + (((e - b) / 2 + NB - 1) / NB) * NB;          // first half: a multiple of NB blocks
        if (mid > e) mid = e;
        p_beg = half ? mid : b;
        p_end = half ? e : mid;
    }
    double2 acc[R][U];
    #pragma unroll
    for (int r = 0; r < R; r++)
        #pragma unroll
        for (int u = 0; u < U; u++) acc[r][u] = make_double2(0.0, 0.0);

    auto stage = [&](const int buf, const int pc, const int nb) {
        const char *src = (const char *) (bval + (size_t) pc * R);
        const int nbytes = nb * R * 8;
        for (int off = lane * 16; off < nbytes; off += 32 * 16) cpa16((char *) &s_val[wib][buf][0] + off, src + off);
        for (int j = lane; j < nb; j += 32) cpa4(&s_col[wib][buf][j], bcol + pc + j);
        cpa_commit();
    };
    int buf = 0;
    if (p_beg < p_end) stage(0, p_beg, min(CB, p_end - p_beg));
    for (int pc = p_beg; pc < p_end; pc += CB, buf ^= 1)
    {
        const int nb = min(CB, p_end - pc);
        cpa_wait_all();
        __syncwarp();
        if (pc + CB < p_end) stage(buf ^ 1, pc + CB, min(CB, p_end - pc - CB));
        const double *sv = &s_val[wib][buf][0];
        const int *sc = &s_col[wib][buf][0];
        int j = 0;
        for (; j + NB <= nb; j += NB)
        {
            double2 x[NB][U];
            #pragma unroll
            for (int q = 0; q < NB; q++)
            {
                const double2 *xr = xlane + (size_t) sc[j + q] * ldx2;
                #pragma unroll
                for (int u = 0; u < U; u++) x[q][u] = __ldg(xr + u * 32);
            }
            #pragma unroll
            for (int q = 0; q < NB; q++)
            {
                double a[R];
                #pragma unroll
                for (int i = 0; i < R / 2; i++)
                {
                    const double2 t = reinterpret_cast<const double2 *>(sv + (size_t) (j + q) * R)[i];
                    a[2 * i] = t.x; a[2 * i + 1] = t.y;
                }
                #pragma unroll
                for (int r = 0; r < R; r++)
                    #pragma unroll
                    for (int u = 0; u < U; u++)
                    {
                        acc[r][u].x = fma(a[r], x[q][u].x, acc[r][u].x);
                        acc[r][u].y = fma(a[r], x[q][u].y, acc[r][u].y);
                    }
            }
        }
        for (; j < nb; j++)
        {
            const double2 *xr = xlane + (size_t) sc[j] * ldx2;
            double a[R];
            #pragma unroll
            for (int i = 0; i < R / 2; i++)
            {
                const double2 t = reinterpret_cast<const double2 *>(sv + (size_t) j * R)[i];
                a[2 * i] = t.x; a[2 * i + 1] = t.y;
            }
            #pragma unroll
            for (int u = 0; u < U; u++)
            {
                const double2 xx = __ldg(xr + u * 32);
                #pragma unroll
                for (int r = 0; r < R; r++)
                {
                    acc[r][u].x = fma(a[r], xx.x, acc[r][u].x);
                    acc[r][u].y = fma(a[r], xx.y, acc[r][u].y);
                }
            }
        }
        __syncwarp();
    }
    if (half == 1)
    {
        #pragma unroll
        for (int r = 0; r < R; r++)
            #pragma unroll
            for (int u = 0; u < U; u++) s_part[pair][r][u][lane] = acc[r][u];
    }
    __syncthreads();                                // every thread of the CTA gets here (no early return above)
    if (half == 0 && valid)
    {
        #pragma unroll
        for (int r = 0; r < R; r++)
        {
            double2 *crow = clane + (size_t) (row0 + r) * ldc2;
            #pragma unroll
            for (int u = 0; u < U; u++)
            {
                const double2 o = s_part[pair][r][u][lane];
                __stcs(crow + u * 32, make_double2(alpha * (acc[r][u].x + o.x), alpha * (acc[r][u].y + o.y)));
            }
        }
    }
}

}  // namespace

// returns false when the launch conditions of the lean variants do not hold (caller falls back to the shipped kernel)
bool crp_launch_rowgroup_x(
    const int cfg, const crp_rowgroup *rg, const double *bval, const int n, const double *X0, const size_t ldx0,
    const double *X1, const double alpha, const double beta, double *C, const size_t ldc, cudaStream_t s
)
{
    if (rg->R != 6 || beta != 0.0) return false;
    if (X1 != NULL) return false;                                   // two-piece operand (received rows): not handled here
    if ((n % 256) != 0 || (ldx0 % 2) != 0 || (ldc % 2) != 0) return false;
    if ((((uintptr_t) X0 | (uintptr_t) C) & 15) != 0) return false;
    const unsigned chunks = (unsigned) (n / 256);                   // U = 4: 32 lanes x 4 x double2
    const int sm = 148;
#define CRP_X(NB, BS, CB, PERSIST, GRID)                                                                                     \
    spmm_rowgroup_x_kernel<6, 4, NB, BS, CB, PERSIST><<<dim3((unsigned) (GRID), chunks), BS, 0, s>>>(                         \
        rg->ngroups, rg->d_grow, rg->d_gptr, rg->d_bcol, bval, (const double2 *) X0, ldx0 / 2, alpha, (double2 *) C, ldc / 2)
    const int per_cta = 4;                                          // warps per 128-thread CTA
    const int all_ctas = (rg->ngroups + per_cta - 1) / per_cta;
    const int resident = sm * 3 < all_ctas ? sm * 3 : all_ctas;     // 3 CTAs per SM at <= 170 registers
    switch (cfg)
    {
        case 60: CRP_X(2, 128, 32, false, all_ctas); break;
        case 61: CRP_X(2, 128, 32, true, resident); break;
        case 62: CRP_X(1, 128, 32, true, resident); break;
        case 63: CRP_X(2, 128, 64, true, resident); break;
        case 64:
            spmm_rowgroup_x2_kernel<6, 4, 2, 32><<<dim3((unsigned) ((rg->ngroups + 1) / 2), chunks), 128, 0, s>>>(
                rg->ngroups, rg->d_grow, rg->d_gptr, rg->d_bcol, bval, (const double2 *) X0, ldx0 / 2, alpha, (double2 *) C, ldc / 2);
            break;
        default: return false;
    }
#undef CRP_X
    CRP_LAUNCH_CHECK();
    return true;
}
