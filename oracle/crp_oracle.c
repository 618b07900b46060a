/*
 * crp_oracle.c - CPU restatement of the CRP-SpMM hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing in the product (crp-spmm_b200/, include/) includes, links or calls this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg do,
 * and only as the checker.
 *
 * It restates, in one process and without MPI, what the reference computes on
 * P ranks.  "Collectives" are loops over the simulated ranks.  Every function
 * names the reference lines it follows (paths relative to the reference root).
 * Parity pin: tests/test_oracle_golden.py checks this file against dumps of the
 * UNMODIFIED reference sources run under oracle/_ref (tests/golden/, made by
 * tests/golden/make_golden.py) - grids, splits, every plan array and C bit for bit
 * (C bit for bit because both use the same left-to-right accumulation; the
 * reference's MKL is unavailable offline, see oracle/stubs/mkl_standin.c).
 */
#include <limits.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ a5: utils.c:26-48 */
ORC_API void orc_block_spos_size(int len, int nblk, int iblk, int *spos, int *size)
{
    if (iblk < 0 || iblk > nblk) { *spos = -1; *size = 0; return; }
    int rem = len % nblk, bs0 = len / nblk, bs1 = bs0 + 1;
    if (iblk < rem) { *spos = bs1 * iblk; *size = bs1; }
    else            { *spos = bs0 * iblk + rem; *size = bs0; }
}

/* ------------------------------------------------------------------ a1: spmat_part.c:12-35 */
ORC_API void orc_row_partition(int nrow, const int *row_ptr, int nblk, int *rblk_ptr)
{
    int nnz = row_ptr[nrow];
    rblk_ptr[0] = 0;
    for (int i = 0; i < nblk; i++)
    {
        int target = (nnz / nblk) * (i + 1);            /* :19 integer division first */
        if (i == nblk - 1) target = nnz;                /* :20 */
        int st = 0, en = nrow;
        while (st < en)                                 /* :22-33 halving search with equality exit */
        {
            int mid = (st + en) / 2;
            if (row_ptr[mid] == target) { st = mid; break; }
            if (row_ptr[mid] < target) st = mid + 1; else en = mid;
        }
        rblk_ptr[i + 1] = st;
    }
}

/* ------------------------------------------------------------------ a2: spmat_part.c:38-64 */
ORC_API void orc_comm_size(
    int nrow, int ncol, const int *row_ptr, const int *col_idx, int nblk,
    const int *rblk_ptr, const int *x_displs, int *comm_sizes, int *total
)
{
    (void) nrow;
    char *seen = (char *) malloc((size_t) (ncol > 0 ? ncol : 1));
    *total = 0;
    for (int b = 0; b < nblk; b++)
    {
        memset(seen, 0, (size_t) ncol);
        for (int p = row_ptr[rblk_ptr[b]]; p < row_ptr[rblk_ptr[b + 1]]; p++) seen[col_idx[p]] = 1;
        int cnt = 0;
        for (int c = 0; c < ncol; c++) cnt += seen[c];                              /* distinct columns touched */
        for (int c = x_displs[b]; c < x_displs[b + 1]; c++) cnt -= seen[c];         /* minus the ones it owns   */
        comm_sizes[b] = cnt;
        *total += cnt;
    }
    free(seen);
}

/* ------------------------------------------------------------------ a3: spmat_part.c:66-81 */
ORC_API int orc_prime_factors(int n, int *out)
{
    int cnt = 0;
    for (int d = 2; d * d <= n; d++)
        while (n % d == 0) { out[cnt++] = d; n /= d; }
    if (n > 1) out[cnt++] = n;
    return cnt;
}

/* ------------------------------------------------------------------ a4: spmat_part.c:85-210
 * Outputs go into caller arrays: A0_rowptr[nproc + 1], B_rowptr / AC_rowptr[nproc + 1], BC_colptr[nproc + 1]. */
ORC_API void orc_part2d(
    int nproc, int m, int n, int k, const int *rb_displs0, const int *rowptr, const int *colidx, int rA,
    int *pm_, int *pn_, uint64_t *comm_cost, int *A0_rowptr, int *B_rowptr, int *AC_rowptr, int *BC_colptr
)
{
    int *k_displs = (int *) malloc(sizeof(int) * (size_t) (nproc + 1));
    int *m_displs = (int *) malloc(sizeof(int) * (size_t) (nproc + 1));
    int *m_try    = (int *) malloc(sizeof(int) * (size_t) (nproc + 1));
    int *sizes    = (int *) malloc(sizeof(int) * (size_t) (nproc + 1));
    int dummy, tot;
    /* :98-112 cost of the 1-D layout */
    for (int i = 0; i <= nproc; i++)
    {
        if (m == k) k_displs[i] = rb_displs0[i];
        else orc_block_spos_size(k, nproc, i, &k_displs[i], &dummy);
        m_displs[i] = rb_displs0[i];
    }
    orc_comm_size(m, k, rowptr, colidx, nproc, rb_displs0, k_displs, sizes, &tot);
    size_t best = (size_t) tot * (size_t) n;
    int pm = nproc, pn = 1, failed = -1;
    /* :114-161 greedy: largest prime factor first */
    int fac[32];
    int nfac = orc_prime_factors(nproc, fac);
    int A_nnz = rowptr[m];
    for (int f = nfac - 1; f >= 0; f--)
    {
        int p = fac[f];
        if (p == failed) continue;                                          /* :123 */
        int pn2 = pn * p, pm2 = nproc / pn2;
        for (int i = 0; i <= pm2; i++) m_try[i] = rb_displs0[i * pn2];      /* :127 */
        for (int i = 0; i <= pm2; i++)
        {
            if (m == k) k_displs[i] = m_try[i];
            else orc_block_spos_size(k, pm2, i, &k_displs[i], &dummy);
        }
        orc_comm_size(m, k, rowptr, colidx, pm2, m_try, k_displs, sizes, &tot);
        size_t A_cost = (size_t) ((double) A_nnz * (double) (pn2 - 1) * 1.5);   /* :143 */
        size_t B_cost = (size_t) rA * (size_t) tot * (size_t) n;                /* :144 full n */
        if (A_cost + B_cost < best)
        {
            best = A_cost + B_cost;  pn = pn2;  pm = pm2;  failed = -1;
            memcpy(m_displs, m_try, sizeof(int) * (size_t) (pm2 + 1));
        } else failed = p;
    }
    *pm_ = pm;  *pn_ = pn;  *comm_cost = (uint64_t) best;
    /* :166-186 */
    for (int i = 0; i <= pm; i++)
    {
        AC_rowptr[i] = m_displs[i];
        if (m == k) B_rowptr[i] = m_displs[i];
        else orc_block_spos_size(k, pm, i, &B_rowptr[i], &dummy);
    }
    for (int j = 0; j <= pn; j++) orc_block_spos_size(n, pn, j, &BC_colptr[j], &dummy);
    /* :188-202 nnz-balanced re-split of each panel into pn pieces */
    for (int ip = 0; ip < pm; ip++)
    {
        int srow = m_displs[ip], nrow = m_displs[ip + 1] - srow;
        int *shifted = (int *) malloc(sizeof(int) * (size_t) (nrow + 1));
        for (int i = 0; i <= nrow; i++) shifted[i] = rowptr[srow + i] - rowptr[srow];
        int *piece = A0_rowptr + ip * pn;
        orc_row_partition(nrow, shifted, pn, piece);
        for (int j = 0; j <= pn; j++) piece[j] += srow;
        free(shifted);
    }
    free(k_displs); free(m_displs); free(m_try); free(sizes);
}

/* ------------------------------------------------------------------ a6/a7: rowpara_spmm.h:8-40, rowpara_spmm.c:20-190 */
typedef struct orc_rp
{
    int    nproc, my_rank, glb_n, A_nrow, rB_nrow;
    int    rB_self_src_offset, rB_self_dst_offset, rB_self_nrow, rB_reidx;
    int    *A_rowptr, *A_colidx, *rB_self_src_ridxs;
    int    *rB_scnts, *rB_sridxs, *rB_sdispls, *rB_rcnts, *rB_rridxs, *rB_rdispls;
    double *A_val;
    uint64_t rB_recv_size;
    /* scratch between orc_rp_create and orc_rp_link */
    int    rB_srow;
    int    *rowmap;
} orc_rp;

/* steps 1-3 of the init for one simulated rank (:46-149); requests stay GLOBAL row ids until orc_rp_link */
ORC_API orc_rp *orc_rp_create(
    int nproc, int me, int A_nrow, const int *A_rowptr, const int *A_colidx, const double *A_val,
    const int *B_row_displs, int glb_n, int reidx
)
{
    orc_rp *r = (orc_rp *) calloc(1, sizeof(orc_rp));
    r->nproc = nproc;  r->my_rank = me;  r->glb_n = glb_n;  r->A_nrow = A_nrow;  r->rB_reidx = reidx;
    int nnz = A_rowptr[A_nrow] - A_rowptr[0], base = A_rowptr[0];
    int glb_k = B_row_displs[nproc];
    r->A_rowptr = (int *) malloc(sizeof(int) * (size_t) (A_nrow + 1));
    r->A_colidx = (int *) malloc(sizeof(int) * (size_t) (nnz + 1));
    r->A_val    = (double *) malloc(sizeof(double) * (size_t) (nnz + 1));
    int srow = INT_MAX, erow = 0;
    for (int i = 0; i < nnz; i++)
    {
        if (A_colidx[i] < srow) srow = A_colidx[i];
        if (A_colidx[i] > erow) erow = A_colidx[i];
    }
    for (int i = 0; i <= A_nrow; i++) r->A_rowptr[i] = A_rowptr[i] - base;
    for (int i = 0; i < nnz; i++) { r->A_colidx[i] = A_colidx[i] - srow; r->A_val[i] = A_val[i]; }
    int *flag = (int *) calloc((size_t) (glb_k > 0 ? glb_k : 1), sizeof(int));
    for (int i = 0; i < nnz; i++) flag[A_colidx[i]] = 1;
    r->rB_srow = srow;
    r->rB_nrow = (nnz > 0) ? erow - srow + 1 : 0;       /* the reference's value is garbage when nnz == 0 */
    r->rowmap = (int *) malloc(sizeof(int) * (size_t) (r->rB_nrow + 1));
    for (int i = 0; i < r->rB_nrow; i++) r->rowmap[i] = i;
    if (reidx)                                          /* :78-86 */
    {
        int cnt = 0;
        for (int g = 0; g < glb_k; g++) if (flag[g]) r->rowmap[g - srow] = cnt++;
        for (int i = 0; i < nnz; i++) r->A_colidx[i] = r->rowmap[r->A_colidx[i]];
        r->rB_nrow = cnt;
    }
    /* :88-117 own rows */
    r->rB_self_src_ridxs = (int *) malloc(sizeof(int) * (size_t) (B_row_displs[me + 1] - B_row_displs[me] + 1));
    for (int g = B_row_displs[me]; g < B_row_displs[me + 1]; g++)
    {
        if (!flag[g]) continue;
        if (r->rB_self_nrow == 0)
        {
            r->rB_self_src_offset = g - B_row_displs[me];
            r->rB_self_dst_offset = reidx ? r->rowmap[g - srow] : g - srow;
        }
        r->rB_self_src_ridxs[r->rB_self_nrow++] = g;
        flag[g] = 0;
    }
    /* :119-149 requests per owner */
    r->rB_rcnts   = (int *) calloc((size_t) nproc, sizeof(int));
    r->rB_rdispls = (int *) calloc((size_t) (nproc + 1), sizeof(int));
    r->rB_rridxs  = (int *) malloc(sizeof(int) * (size_t) (r->rB_nrow + 1));
    int cnt = 0;
    for (int p = 0; p < nproc; p++)
    {
        for (int g = B_row_displs[p]; g < B_row_displs[p + 1]; g++)
            if (flag[g]) { r->rB_rridxs[cnt++] = g; r->rB_rcnts[p]++; }
        r->rB_rdispls[p + 1] = r->rB_rdispls[p] + r->rB_rcnts[p];
    }
    r->rB_recv_size = (uint64_t) (r->rB_rdispls[nproc] - r->rB_rcnts[me]);
    free(flag);
    return r;
}

/* steps 4-5 for all ranks of a communicator (:151-184): the Alltoall / Alltoallv become loops */
ORC_API void orc_rp_link(orc_rp **rk, int nproc, const int *B_row_displs)
{
    for (int me = 0; me < nproc; me++)
    {
        orc_rp *r = rk[me];
        r->rB_scnts   = (int *) calloc((size_t) nproc, sizeof(int));
        r->rB_sdispls = (int *) calloc((size_t) (nproc + 1), sizeof(int));
        for (int p = 0; p < nproc; p++)
        {
            r->rB_scnts[p] = rk[p]->rB_rcnts[me];
            r->rB_sdispls[p + 1] = r->rB_sdispls[p] + r->rB_scnts[p];
        }
        r->rB_sridxs = (int *) malloc(sizeof(int) * (size_t) (r->rB_sdispls[nproc] + 1));
        for (int p = 0; p < nproc; p++)
            for (int i = 0; i < r->rB_scnts[p]; i++)
                r->rB_sridxs[r->rB_sdispls[p] + i] = rk[p]->rB_rridxs[rk[p]->rB_rdispls[me] + i] - B_row_displs[me];
    }
    for (int me = 0; me < nproc; me++)
    {
        orc_rp *r = rk[me];
        int n = r->glb_n;
        for (int i = 0; i < r->rB_rdispls[nproc]; i++)
        {
            int v = r->rB_rridxs[i] - r->rB_srow;
            r->rB_rridxs[i] = r->rB_reidx ? r->rowmap[v] : v;
        }
        for (int p = 0; p < nproc; p++)
        {
            r->rB_rcnts[p] *= n;  r->rB_rdispls[p] *= n;
            r->rB_scnts[p] *= n;  r->rB_sdispls[p] *= n;
        }
        r->rB_rdispls[nproc] *= n;
        r->rB_sdispls[nproc] *= n;
    }
}

ORC_API void orc_rp_free(orc_rp *r)
{
    if (r == NULL) return;
    free(r->A_rowptr); free(r->A_colidx); free(r->A_val); free(r->rB_self_src_ridxs);
    free(r->rB_scnts); free(r->rB_sridxs); free(r->rB_sdispls);
    free(r->rB_rcnts); free(r->rB_rridxs); free(r->rB_rdispls); free(r->rowmap);
    free(r);
}

/* element (row i, col j) of a dense block in either layout */
static inline size_t at(int layout, size_t ld, size_t i, size_t j) { return layout == 0 ? i * ld + j : j * ld + i; }

/* ------------------------------------------------------------------ a8-a11: rowpara_spmm.c:212-422
 * One exec of all ranks of a communicator: pack -> exchange -> unpack -> self copy -> local product.
 * The local product is the textbook CSR loop (what the MKL call at :398-408 computes):
 * per output element, products added left to right in stored nonzero order, plain mul + add. */
ORC_API void orc_rp_exec(orc_rp **rk, int nproc, int layout, const double **B, const int *ldB, double **C, const int *ldC)
{
    for (int me = 0; me < nproc; me++)
    {
        orc_rp *r = rk[me];
        int n = r->glb_n, rBn = r->rB_nrow;
        size_t ldr = (layout == 0) ? (size_t) n : (size_t) rBn;
        double *rB = (double *) malloc(sizeof(double) * ((size_t) rBn * (size_t) n + 1));
        /* rows other ranks packed for me, in the order their send lists hold them (:232-262, :314-344) */
        for (int p = 0; p < nproc; p++)
        {
            int nrecv = n ? r->rB_rcnts[p] / n : 0;
            const int *dstpos = r->rB_rridxs + (n ? r->rB_rdispls[p] / n : 0);
            const int *srcrow = rk[p]->rB_sridxs + (n ? rk[p]->rB_sdispls[me] / n : 0);
            for (int i = 0; i < nrecv; i++)
                for (int j = 0; j < n; j++)
                    rB[at(layout, ldr, (size_t) dstpos[i], (size_t) j)] = B[p][at(layout, (size_t) ldB[p], (size_t) srcrow[i], (size_t) j)];
        }
        /* own rows (:348-384) */
        for (int i = 0; i < r->rB_self_nrow; i++)
        {
            int step = r->rB_self_src_ridxs[i] - r->rB_self_src_ridxs[0];
            int src = r->rB_self_src_offset + step;
            int dst = r->rB_self_dst_offset + (r->rB_reidx ? i : step);
            for (int j = 0; j < n; j++)
                rB[at(layout, ldr, (size_t) dst, (size_t) j)] = B[me][at(layout, (size_t) ldB[me], (size_t) src, (size_t) j)];
        }
        /* C = A * rB */
        for (int i = 0; i < r->A_nrow; i++)
            for (int j = 0; j < n; j++)
            {
                double acc = 0.0;
                for (int p = r->A_rowptr[i]; p < r->A_rowptr[i + 1]; p++)
                    acc += r->A_val[p] * rB[at(layout, ldr, (size_t) r->A_colidx[p], (size_t) j)];
                C[me][at(layout, (size_t) ldC[me], (size_t) i, (size_t) j)] = 1.0 * acc;
            }
        free(rB);
    }
}

/* plain CSR x dense, row-major, fp64 or fp32 accumulation: the single-process check the drivers run
 * (examples/test_utils.c:157-178) */
ORC_API void orc_csr_spmm(int m, int n, const int *rowptr, const int *colidx, const double *val, const double *B, int ldB, double *C, int ldC)
{
    #pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < m; i++)
    {
        double *c = C + (size_t) i * (size_t) ldC;
        for (int j = 0; j < n; j++) c[j] = 0.0;
        for (int p = rowptr[i]; p < rowptr[i + 1]; p++)
        {
            const double a = val[p];
            const double *x = B + (size_t) colidx[p] * (size_t) ldB;
            for (int j = 0; j < n; j++) c[j] += a * x[j];
        }
    }
}

/* ------------------------------------------------------------------ a13: para2d_spmm.c:49-98
 * Rows of grid row pi pooled from its pn owners: row pointers keep the first owner's offset. */
ORC_API void orc_para2d_panel(
    int pn, int pi, const int *A0_rowptr, const int *const *rowptrs, const int *const *colidxs, const double *const *vals,
    int *panel_rowptr, int *panel_colidx, double *panel_val
)
{
    int r = 0, z = 0;
    for (int j = 0; j < pn; j++)
    {
        int rank = pi * pn + j;
        int nrow = A0_rowptr[rank + 1] - A0_rowptr[rank];
        int nnz = rowptrs[j][nrow] - rowptrs[j][0];
        for (int i = 0; i < nrow; i++) panel_rowptr[r + i] = rowptrs[j][i];
        memcpy(panel_colidx + z, colidxs[j], sizeof(int) * (size_t) nnz);
        memcpy(panel_val + z, vals[j], sizeof(double) * (size_t) nnz);
        r += nrow;  z += nnz;
    }
    panel_rowptr[r] = panel_rowptr[0] + z;      /* :77 */
}

/* rA_cost as the last rank computes it (:100-109) */
ORC_API uint64_t orc_para2d_rA_cost(int glb_nnz, int pn) { return (uint64_t) (size_t) ((double) glb_nnz * (double) (pn - 1) * 1.5); }

/* ------------------------------------------------------------------ a15/a16: mat_redist.c:9-41, 79-204, 298-419 */
static int seg_isect(int s0, int e0, int s1, int e1, int *is, int *ie)
{
    if (s0 > s1) { int t = s0; s0 = s1; s1 = t; t = e0; e0 = e1; e1 = t; }
    if (s1 > e0 || s1 > e1 || s0 > e0) return 0;
    *is = s1;  *ie = (e0 < e1) ? e0 : e1;
    return 1;
}

/* rects: per rank 8 ints {src_srow, src_scol, src_nrow, src_ncol, req_srow, req_scol, req_nrow, req_ncol}.
 * side 0: what `me` sends (its src against everyone's req); side 1: what it receives.
 * Outputs sized nproc: ranks, sizes, blks[4 * nproc] (srow, scol, nrow, ncol); displs[nproc + 1].  Returns #pieces. */
ORC_API int orc_redist_plan(int nproc, const int *rects, int me, int side, int *ranks, int *sizes, int *displs, int *blks, int *total)
{
    const int *mine = rects + 8 * me + (side == 0 ? 0 : 4);
    int n = 0, cnt = 0;
    for (int p = 0; p < nproc; p++)
    {
        const int *o = rects + 8 * p + (side == 0 ? 4 : 0);
        int rs, re, cs, ce;
        if (!seg_isect(mine[0], mine[0] + mine[2] - 1, o[0], o[0] + o[2] - 1, &rs, &re)) continue;
        if (!seg_isect(mine[1], mine[1] + mine[3] - 1, o[1], o[1] + o[3] - 1, &cs, &ce)) continue;
        blks[4 * n] = rs;  blks[4 * n + 1] = cs;  blks[4 * n + 2] = re - rs + 1;  blks[4 * n + 3] = ce - cs + 1;
        ranks[n] = p;  displs[n] = cnt;  sizes[n] = blks[4 * n + 2] * blks[4 * n + 3];
        cnt += sizes[n];
        n++;
    }
    displs[n] = cnt;
    *total = cnt;
    return n;
}

/* the whole exchange for all ranks: dst blocks receive the wanted rectangles (row-major, any element size) */
ORC_API void orc_redist_exec(int nproc, const int *rects, int dt_size, const char **src, const int *src_ld, char **dst, const int *dst_ld)
{
    int *ranks = (int *) malloc(sizeof(int) * (size_t) nproc), *sizes = (int *) malloc(sizeof(int) * (size_t) nproc);
    int *displs = (int *) malloc(sizeof(int) * (size_t) (nproc + 1)), *blks = (int *) malloc(sizeof(int) * 4 * (size_t) nproc);
    for (int me = 0; me < nproc; me++)
    {
        int total;
        int n = orc_redist_plan(nproc, rects, me, 1, ranks, sizes, displs, blks, &total);
        const int *req = rects + 8 * me + 4;
        for (int b = 0; b < n; b++)
        {
            int p = ranks[b];
            const int *ps = rects + 8 * p;
            for (int i = 0; i < blks[4 * b + 2]; i++)
            {
                size_t so = ((size_t) (blks[4 * b] + i - ps[0]) * (size_t) src_ld[p] + (size_t) (blks[4 * b + 1] - ps[1])) * (size_t) dt_size;
                size_t d_o = ((size_t) (blks[4 * b] + i - req[0]) * (size_t) dst_ld[me] + (size_t) (blks[4 * b + 1] - req[1])) * (size_t) dt_size;
                memcpy(dst[me] + d_o, src[p] + so, (size_t) blks[4 * b + 3] * (size_t) dt_size);
            }
        }
    }
    free(ranks); free(sizes); free(displs); free(blks);
}
