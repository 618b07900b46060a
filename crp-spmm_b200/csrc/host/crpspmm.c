/*
 * crpspmm.c - the composite engine of include/crpspmm.h: A, B and C in the caller's layouts.
 *
 * What the reference's deprecated engine does in one 786-line file with its own exchange code
 * (deprecated/src/crpspmm.c) is a composition here:
 *   init : gather the sparsity pattern's row lengths on every rank and the column indices on rank 0,
 *          let the LIVE cost model pick the grid (csr_mat_row_partition + calc_spmm_part2d_from_1d,
 *          src/spmat_part.c), plan the row-interval exchange that moves A from the caller's 1-D row
 *          layout to the model's initial ownership A0_rowptr, move the column indices, and set up two
 *          mat_redist engines (B in, C out) on device memory;
 *   exec : move A's values with the same plan; (re)build the para2d engine if the values changed;
 *          redistribute B on the device; para2d_spmm_exec on device-resident blocks; redistribute C.
 * Statistics and the "Communicated Matrix Elements" table follow deprecated/src/crpspmm.c:715-772.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mpi.h>

#include "utils.h"
#include "spmat_part.h"
#include "crpspmm.h"
#include "crp_ext.h"
#include "crp_internal.h"

struct crp_composite
{
    int    *glb_rowptr;                     /* m + 1, on every rank                                   */
    int    *A0_rowptr, *B_rowptr, *AC_rowptr, *BC_colptr;
    int    src_A_srow, src_A_nrow, src_nnz;
    int    src_B[4], dst_C[4];              /* srow, scol, nrow, ncol                                 */
    int    *s_cnt, *s_dsp, *r_cnt, *r_dsp;  /* nonzero-level exchange plan of the A redistribution     */
    double *val_cache;                      /* values the current para2d engine was built from         */
    int    have_engine;
    double prev_a2a, prev_spmm, prev_pack;  /* rp counters at the last read                            */
    void   *dB_user, *dB_loc, *dC_loc, *dC_user;
};

static void *xmalloc(size_t bytes)
{
    void *p = malloc(bytes > 0 ? bytes : 1);
    ASSERT_PRINTF(p != NULL, "Failed to allocate %zu bytes of work memory for crpspmm_engine\n", bytes);
    return p;
}

/* overlap of the row intervals [a0, a1) and [b0, b1) */
static int rows_overlap(int a0, int a1, int b0, int b1, int *lo, int *hi)
{
    *lo = a0 > b0 ? a0 : b0;
    *hi = a1 < b1 ? a1 : b1;
    if (*hi < *lo) *hi = *lo;
    return *hi > *lo;
}

void crpspmm_engine_init(
    const int m, const int n, const int k,
    const int src_A_srow, const int src_A_nrow,
    const int *src_A_rowptr, const int *src_A_colidx,
    const int src_B_srow, const int src_B_nrow,
    const int src_B_scol, const int src_B_ncol,
    const int dst_C_srow, const int dst_C_nrow,
    const int dst_C_scol, const int dst_C_ncol,
    MPI_Comm comm, int use_CUDA, crpspmm_engine_p *engine_, size_t *workbuf_bytes
)
{
    const double t0 = get_wtime_sec();
    crpspmm_engine_p e = (crpspmm_engine_p) calloc(1, sizeof(crpspmm_engine_s));
    struct crp_composite *c = (struct crp_composite *) calloc(1, sizeof(struct crp_composite));
    ASSERT_PRINTF(e != NULL && c != NULL, "Failed to allocate crpspmm_engine\n");
    e->priv = c;
    e->comm_glb = comm;
    e->use_CUDA = use_CUDA;
    e->alloc_workbuf = 1;
    e->glb_m = m;  e->glb_n = n;  e->glb_k = k;
    MPI_Comm_size(comm, &e->np_glb);
    MPI_Comm_rank(comm, &e->rank_glb);
    const int nproc = e->np_glb, me = e->rank_glb;
    c->src_A_srow = src_A_srow;  c->src_A_nrow = src_A_nrow;
    c->src_nnz = src_A_rowptr[src_A_nrow] - src_A_rowptr[0];
    c->src_B[0] = src_B_srow;  c->src_B[1] = src_B_scol;  c->src_B[2] = src_B_nrow;  c->src_B[3] = src_B_ncol;
    c->dst_C[0] = dst_C_srow;  c->dst_C[1] = dst_C_scol;  c->dst_C[2] = dst_C_nrow;  c->dst_C[3] = dst_C_ncol;

    /* 1. global row pointers on every rank, column indices on rank 0 */
    int mine[2] = { src_A_srow, src_A_nrow };
    int *ivals = (int *) xmalloc(sizeof(int) * 2 * (size_t) nproc);
    MPI_Allgather(mine, 2, MPI_INT, ivals, 2, MPI_INT, comm);
    int *srow_all = (int *) xmalloc(sizeof(int) * (size_t) nproc), *nrow_all = (int *) xmalloc(sizeof(int) * (size_t) nproc);
    long long covered = 0;
    for (int p = 0; p < nproc; p++) { srow_all[p] = ivals[2 * p]; nrow_all[p] = ivals[2 * p + 1]; covered += nrow_all[p]; }
    free(ivals);
    ASSERT_PRINTF(covered == m, "crpspmm_engine_init: the processes' A row ranges hold %lld rows, the matrix has %d\n", covered, m);
    int *rowlen = (int *) xmalloc(sizeof(int) * (size_t) src_A_nrow);
    for (int i = 0; i < src_A_nrow; i++) rowlen[i] = src_A_rowptr[i + 1] - src_A_rowptr[i];
    int *G = (int *) xmalloc(sizeof(int) * ((size_t) m + 1));
    MPI_Allgatherv(rowlen, src_A_nrow, MPI_INT, G + 1, nrow_all, srow_all, MPI_INT, comm);
    free(rowlen);
    G[0] = 0;
    for (int i = 0; i < m; i++) G[i + 1] += G[i];
    c->glb_rowptr = G;

    int *nnz_cnt = (int *) xmalloc(sizeof(int) * (size_t) nproc), *nnz_dsp = (int *) xmalloc(sizeof(int) * (size_t) nproc);
    for (int p = 0; p < nproc; p++) { nnz_dsp[p] = G[srow_all[p]]; nnz_cnt[p] = G[srow_all[p] + nrow_all[p]] - G[srow_all[p]]; }
    int *glb_colidx = (me == 0) ? (int *) xmalloc(sizeof(int) * (size_t) G[m]) : NULL;
    MPI_Gatherv(src_A_colidx, c->src_nnz, MPI_INT, glb_colidx, nnz_cnt, nnz_dsp, MPI_INT, 0, comm);
    free(nnz_cnt);
    free(nnz_dsp);

    /* 2. the live cost model chooses the grid and the splits (rank 0), everybody learns them */
    int hdr[2] = { nproc, 1 };
    int *A0 = NULL, *Br = NULL, *AC = NULL, *BC = NULL;
    if (me == 0)
    {
        int *rb = (int *) xmalloc(sizeof(int) * ((size_t) nproc + 1));
        size_t cost = 0;
        csr_mat_row_partition(m, G, nproc, rb);
        calc_spmm_part2d_from_1d(nproc, m, n, k, rb, G, glb_colidx, 1, &hdr[0], &hdr[1], &cost, &A0, &Br, &AC, &BC, 0);
        free(rb);
        free(glb_colidx);
    }
    MPI_Bcast(hdr, 2, MPI_INT, 0, comm);
    const int pm = hdr[0], pn = hdr[1];
    if (me != 0)
    {
        A0 = (int *) xmalloc(sizeof(int) * ((size_t) nproc + 1));
        Br = (int *) xmalloc(sizeof(int) * ((size_t) pm + 1));
        AC = (int *) xmalloc(sizeof(int) * ((size_t) pm + 1));
        BC = (int *) xmalloc(sizeof(int) * ((size_t) pn + 1));
    }
    MPI_Bcast(A0, nproc + 1, MPI_INT, 0, comm);
    MPI_Bcast(Br, pm + 1, MPI_INT, 0, comm);
    MPI_Bcast(AC, pm + 1, MPI_INT, 0, comm);
    MPI_Bcast(BC, pn + 1, MPI_INT, 0, comm);
    c->A0_rowptr = A0;  c->B_rowptr = Br;  c->AC_rowptr = AC;  c->BC_colptr = BC;
    e->np_row = pm;  e->np_col = pn;
    e->rank_row = me / pn;  e->rank_col = me % pn;
    const int pi = e->rank_row, pj = e->rank_col;
    e->loc_A_srow = A0[me];  e->loc_A_erow = A0[me + 1];
    e->loc_A_nrow = A0[me + 1] - A0[me];
    e->loc_A_nnz  = G[A0[me + 1]] - G[A0[me]];
    e->loc_B_srow = Br[pi];  e->loc_B_erow = Br[pi + 1];  e->loc_B_nrow = Br[pi + 1] - Br[pi];
    e->loc_B_scol = BC[pj];  e->loc_B_ecol = BC[pj + 1];  e->loc_B_ncol = BC[pj + 1] - BC[pj];
    e->loc_C_srow = AC[pi];  e->loc_C_nrow = AC[pi + 1] - AC[pi];

    /* 3. nonzero-level plan of the A redistribution: row-interval overlaps, nothing to negotiate */
    c->s_cnt = (int *) xmalloc(sizeof(int) * (size_t) nproc);  c->s_dsp = (int *) xmalloc(sizeof(int) * (size_t) nproc);
    c->r_cnt = (int *) xmalloc(sizeof(int) * (size_t) nproc);  c->r_dsp = (int *) xmalloc(sizeof(int) * (size_t) nproc);
    for (int p = 0; p < nproc; p++)
    {
        int lo, hi;
        rows_overlap(src_A_srow, src_A_srow + src_A_nrow, A0[p], A0[p + 1], &lo, &hi);         /* what I hold and p will own */
        c->s_cnt[p] = G[hi] - G[lo];
        c->s_dsp[p] = G[lo] - G[src_A_srow];
        rows_overlap(srow_all[p], srow_all[p] + nrow_all[p], A0[me], A0[me + 1], &lo, &hi);    /* what p holds and I will own */
        c->r_cnt[p] = G[hi] - G[lo];
        c->r_dsp[p] = G[lo] - G[A0[me]];
    }
    free(srow_all);
    free(nrow_all);
    e->loc_A_rowptr = (int *) xmalloc(sizeof(int) * ((size_t) e->loc_A_nrow + 1));
    memcpy(e->loc_A_rowptr, G + A0[me], sizeof(int) * ((size_t) e->loc_A_nrow + 1));             /* global nnz offsets, as scatter_csr_rows leaves them */
    e->loc_A_colidx = (int *) xmalloc(sizeof(int) * (size_t) e->loc_A_nnz);
    e->loc_A_val    = (double *) xmalloc(sizeof(double) * (size_t) e->loc_A_nnz);
    c->val_cache    = (double *) xmalloc(sizeof(double) * (size_t) e->loc_A_nnz);
    MPI_Alltoallv(src_A_colidx, c->s_cnt, c->s_dsp, MPI_INT, e->loc_A_colidx, c->r_cnt, c->r_dsp, MPI_INT, comm);

    /* 4. dense redistributions on device memory (skipped in plan-only mode: no device, no exec) */
    if (crp_device_ready())
    {
        mat_redist_engine_init(src_B_srow, src_B_scol, src_B_nrow, src_B_ncol, e->loc_B_srow, e->loc_B_scol, e->loc_B_nrow, e->loc_B_ncol,
                               comm, MPI_DOUBLE, sizeof(double), DEV_TYPE_CUDA, &e->rd_B, NULL);
        mat_redist_engine_init(e->loc_C_srow, e->loc_B_scol, e->loc_C_nrow, e->loc_B_ncol, dst_C_srow, dst_C_scol, dst_C_nrow, dst_C_ncol,
                               comm, MPI_DOUBLE, sizeof(double), DEV_TYPE_CUDA, &e->rd_C, NULL);
        ASSERT_PRINTF(e->rd_B != NULL && e->rd_C != NULL, "crpspmm_engine_init: cannot set up the B / C redistributions\n");
        crp_cuda_malloc_dev(&c->dB_loc, sizeof(double) * (size_t) e->loc_B_nrow * (size_t) e->loc_B_ncol);
        crp_cuda_malloc_dev(&c->dC_loc, sizeof(double) * (size_t) e->loc_C_nrow * (size_t) e->loc_B_ncol);
    }
    if (workbuf_bytes != NULL) *workbuf_bytes = 0;
    e->t_init = get_wtime_sec() - t0;
    *engine_ = e;
}

void crpspmm_engine_attach_workbuf(crpspmm_engine_p engine, double *workbuf)
{
    (void) engine;
    (void) workbuf;
}

void crpspmm_engine_redist_A_values(crpspmm_engine_p e, const double *src_A_val)
{
    if (e == NULL) return;
    struct crp_composite *c = (struct crp_composite *) e->priv;
    MPI_Alltoallv(src_A_val, c->s_cnt, c->s_dsp, MPI_DOUBLE, e->loc_A_val, c->r_cnt, c->r_dsp, MPI_DOUBLE, e->comm_glb);
}

void crpspmm_engine_exec(
    crpspmm_engine_p e,
    const int *src_A_rowptr, const int *src_A_colidx, const double *src_A_val,
    const double *src_B, const int ldB, double *dst_C, const int ldC
)
{
    (void) src_A_rowptr;
    (void) src_A_colidx;        /* the pattern was fixed at init, as in the reference */
    if (e == NULL) return;
    struct crp_composite *c = (struct crp_composite *) e->priv;
    if (!crp_device_ready())
    {
        fprintf(stderr, "[FATAL] crpspmm_engine_exec: no CUDA device (plan-only mode); there is no CPU path\n");
        fflush(stderr);
        abort();
    }
    const int pn = e->np_col, pi = e->rank_row;
    const double t_start = get_wtime_sec();
    double t0, t1;
    /* The three stages run on the engines' own streams and hand their results over through host synchronisation, so the
     * composite always runs them with blocking semantics, whatever the caller set with crp_set_blocking(). */
    const int caller_blocking = crp_opt_blocking();
    crp_set_blocking(1);

    /* 1. A's values into the owned-rows layout; rebuild the 2-D engine only if they changed */
    t0 = get_wtime_sec();
    crpspmm_engine_redist_A_values(e, src_A_val);
    t1 = get_wtime_sec();
    e->t_rd_A += t1 - t0;
    /* the counters follow the reference's definitions (deprecated/src/crpspmm.c:449-456, 587-596): what each rank HOLDS after
     * a phase, its own share included - not what crossed a link */
    e->nelem_A_rd = (size_t) e->loc_A_nnz;
    int rebuild = !c->have_engine || memcmp(c->val_cache, e->loc_A_val, sizeof(double) * (size_t) e->loc_A_nnz) != 0;
    int any = 0;
    MPI_Allreduce(&rebuild, &any, 1, MPI_INT, MPI_MAX, e->comm_glb);        /* para2d_spmm_init is collective */
    if (any)
    {
        if (e->p2d != NULL) para2d_spmm_free(&e->p2d);
        para2d_spmm_init(e->comm_glb, e->np_row, e->np_col, c->A0_rowptr, c->B_rowptr, c->AC_rowptr, c->BC_colptr,
                         e->loc_A_rowptr, e->loc_A_colidx, e->loc_A_val, &e->p2d);
        memcpy(c->val_cache, e->loc_A_val, sizeof(double) * (size_t) e->loc_A_nnz);
        c->have_engine = 1;
        c->prev_a2a = c->prev_spmm = c->prev_pack = 0.0;
        e->t_agv_A += e->p2d->t_ag_A;
        const int *G = c->glb_rowptr;
        const size_t panel = (size_t) (G[c->A0_rowptr[(pi + 1) * pn]] - G[c->A0_rowptr[pi * pn]]);
        e->nelem_A_agv = (pn > 1) ? panel : 0;
    }

    /* 2. B: caller's block -> the grid's block, on the device */
    t0 = get_wtime_sec();
    const size_t es = sizeof(double);
    const void *Bsrc = src_B;
    int ldBsrc = ldB;
    if (c->src_B[2] > 0 && c->src_B[3] > 0 && !crp_cuda_ptr_is_device(src_B))
    {
        if (c->dB_user == NULL) crp_cuda_malloc_dev(&c->dB_user, es * (size_t) c->src_B[2] * (size_t) c->src_B[3]);
        crp_cuda_memcpy2d_async(src_B, es * (size_t) ldB, c->dB_user, es * (size_t) c->src_B[3], es * (size_t) c->src_B[3], (size_t) c->src_B[2], NULL);
        crp_cuda_stream_sync(NULL);
        Bsrc = c->dB_user;
        ldBsrc = c->src_B[3];
    }
    mat_redist_engine_exec(e->rd_B, Bsrc, ldBsrc, c->dB_loc, e->loc_B_ncol);
    t1 = get_wtime_sec();
    e->t_rd_B += t1 - t0;
    e->nelem_B_rd = (size_t) e->rd_B->recv_cnt;

    /* 3. replicate B + local SpMM on device-resident blocks */
    t0 = get_wtime_sec();
    para2d_spmm_exec(e->p2d, 0, (const double *) c->dB_loc, e->loc_B_ncol, (double *) c->dC_loc, e->loc_B_ncol);
    t1 = get_wtime_sec();
    e->t_exec_nr += t1 - t0;
    rp_spmm_p rp = e->p2d->rp_spmm;
    rp_spmm_sync_stats(rp);                     /* fold this exec's CUDA-event times before reading the counters */
    e->t_a2a_B += (rp->t_a2a + rp->t_pack) - (c->prev_a2a + c->prev_pack);
    e->t_spmm  += rp->t_spmm - c->prev_spmm;
    c->prev_a2a = rp->t_a2a;  c->prev_pack = rp->t_pack;  c->prev_spmm = rp->t_spmm;
    /* only the needed rows travel (the reference's A2A_B_FINEGRAIN=1 mode): own rows + received rows, times the local width */
    e->nelem_B_a2av_min = ((size_t) rp->rB_self_nrow + rp->rB_recv_size) * (size_t) rp->glb_n;
    e->nelem_B_a2av = (e->np_row > 1) ? e->nelem_B_a2av_min : 0;

    /* 4. C: the grid's block -> caller's block */
    t0 = get_wtime_sec();
    void *Cdst = dst_C;
    int ldCdst = ldC;
    const int C_host = (c->dst_C[2] > 0 && c->dst_C[3] > 0 && !crp_cuda_ptr_is_device(dst_C));
    if (C_host)
    {
        if (c->dC_user == NULL) crp_cuda_malloc_dev(&c->dC_user, es * (size_t) c->dst_C[2] * (size_t) c->dst_C[3]);
        Cdst = c->dC_user;
        ldCdst = c->dst_C[3];
    }
    mat_redist_engine_exec(e->rd_C, c->dC_loc, e->loc_B_ncol, Cdst, ldCdst);
    if (C_host)
    {
        crp_cuda_memcpy2d_async(c->dC_user, es * (size_t) c->dst_C[3], dst_C, es * (size_t) ldC, es * (size_t) c->dst_C[3], (size_t) c->dst_C[2], NULL);
        crp_cuda_stream_sync(NULL);
    }
    t1 = get_wtime_sec();
    e->t_rd_C += t1 - t0;

    crp_set_blocking(caller_blocking);
    e->t_exec += get_wtime_sec() - t_start;
    e->n_exec++;
}

void crpspmm_engine_free(crpspmm_engine_p *engine_)
{
    crpspmm_engine_p e = *engine_;
    if (e == NULL) return;
    struct crp_composite *c = (struct crp_composite *) e->priv;
    if (e->p2d != NULL) para2d_spmm_free(&e->p2d);
    mat_redist_engine_free(&e->rd_B);
    mat_redist_engine_free(&e->rd_C);
    if (c->dB_user) crp_cuda_free_dev(c->dB_user);
    if (c->dB_loc)  crp_cuda_free_dev(c->dB_loc);
    if (c->dC_loc)  crp_cuda_free_dev(c->dC_loc);
    if (c->dC_user) crp_cuda_free_dev(c->dC_user);
    free(c->glb_rowptr);
    free(c->A0_rowptr);  free(c->B_rowptr);  free(c->AC_rowptr);  free(c->BC_colptr);
    free(c->s_cnt);  free(c->s_dsp);  free(c->r_cnt);  free(c->r_dsp);
    free(c->val_cache);
    free(c);
    free(e->loc_A_rowptr);
    free(e->loc_A_colidx);
    free(e->loc_A_val);
    free(e);
    *engine_ = NULL;
}

/* Same two tables, same row labels as deprecated/src/crpspmm.c:715-772. */
void crpspmm_engine_print_stat(crpspmm_engine_p e)
{
    if (e == NULL) return;
    if (e->rank_glb == 0) printf("crpspmm_engine init time: %.3f s\n", e->t_init);
    const int n_exec = e->n_exec;
    if (n_exec == 0) return;
    double raw[8] = { e->t_rd_A, e->t_rd_B, e->t_agv_A, e->t_a2a_B, e->t_spmm, e->t_exec_nr, e->t_rd_C, e->t_exec };
    double tmin[8], tmax[8], tavg[8];
    unsigned long long cs[5] = { e->nelem_A_rd, e->nelem_A_agv, e->nelem_B_rd, e->nelem_B_a2av, e->nelem_B_a2av_min };
    unsigned long long cmin[5], cmax[5], csum[5];
    MPI_Reduce(raw, tmin, 8, MPI_DOUBLE, MPI_MIN, 0, e->comm_glb);
    MPI_Reduce(raw, tmax, 8, MPI_DOUBLE, MPI_MAX, 0, e->comm_glb);
    MPI_Reduce(raw, tavg, 8, MPI_DOUBLE, MPI_SUM, 0, e->comm_glb);
    MPI_Reduce(cs, cmin, 5, MPI_UNSIGNED_LONG_LONG, MPI_MIN, 0, e->comm_glb);
    MPI_Reduce(cs, cmax, 5, MPI_UNSIGNED_LONG_LONG, MPI_MAX, 0, e->comm_glb);
    MPI_Reduce(cs, csum, 5, MPI_UNSIGNED_LONG_LONG, MPI_SUM, 0, e->comm_glb);
    if (e->rank_glb != 0) return;
    for (int i = 0; i < 8; i++)
    {
        tmin[i] /= (double) n_exec;
        tmax[i] /= (double) n_exec;
        tavg[i] /= (double) e->np_glb * n_exec;
    }
    printf("-------------------------- Runtime (s) -------------------------\n");
    printf("                                   min         avg         max\n");
    printf("Redist A to internal 1D layout  %6.3f      %6.3f      %6.3f\n", tmin[0], tavg[0], tmax[0]);
    printf("Redist B to internal 2D layout  %6.3f      %6.3f      %6.3f\n", tmin[1], tavg[1], tmax[1]);
    printf("Replicate A with allgatherv     %6.3f      %6.3f      %6.3f\n", tmin[2], tavg[2], tmax[2]);
    printf("Replicate B with alltoallv      %6.3f      %6.3f      %6.3f\n", tmin[3], tavg[3], tmax[3]);
    printf("Local SpMM                      %6.3f      %6.3f      %6.3f\n", tmin[4], tavg[4], tmax[4]);
    printf("SpMM w/o Redist                 %6.3f      %6.3f      %6.3f\n", tmin[5], tavg[5], tmax[5]);
    printf("Redist C to user's 2D layout    %6.3f      %6.3f      %6.3f\n", tmin[6], tavg[6], tmax[6]);
    printf("SpMM total (avg of %3d runs)    %6.3f      %6.3f      %6.3f\n", n_exec, tmin[7], tavg[7], tmax[7]);
    printf("----------------------------------------------------------------\n");
    printf("------------------ Communicated Matrix Elements -----------------\n");
    printf("                               min           max            sum\n");
    printf("Redist A                %10zu    %10zu    %11zu\n", (size_t) cmin[0], (size_t) cmax[0], (size_t) csum[0]);
    printf("Allgatherv A            %10zu    %10zu    %11zu\n", (size_t) cmin[1], (size_t) cmax[1], (size_t) csum[1]);
    printf("Redist B                %10zu    %10zu    %11zu\n", (size_t) cmin[2], (size_t) cmax[2], (size_t) csum[2]);
    printf("Alltoallv B             %10zu    %10zu    %11zu\n", (size_t) cmin[3], (size_t) cmax[3], (size_t) csum[3]);
    printf("Alltoallv B necessary   %10zu    %10zu    %11zu\n", (size_t) cmin[4], (size_t) cmax[4], (size_t) csum[4]);
    printf("----------------------------------------------------------------\n");
    printf("\n");
    fflush(stdout);
}

void crpspmm_engine_clear_stat(crpspmm_engine_p e)
{
    if (e == NULL) return;
    e->n_exec    = 0;
    e->t_exec    = 0.0;
    e->t_rd_A    = 0.0;
    e->t_agv_A   = 0.0;
    e->t_rd_B    = 0.0;
    e->t_a2a_B   = 0.0;
    e->t_spmm    = 0.0;
    e->t_rd_C    = 0.0;
    e->t_exec_nr = 0.0;
}
