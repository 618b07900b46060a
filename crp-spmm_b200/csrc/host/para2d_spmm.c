/*
 * para2d_spmm.c - 2-D (pm x pn) SpMM engine (include/para2d_spmm.h).
 *
 * Same contract as reference src/para2d_spmm.c:20-205: rank r is grid point
 * (r / pn, r % pn); the pn ranks of a grid row pool their A rows ("replicate
 * A", reference lines 49-98), then every grid column runs a row-parallel engine
 * on its column slice of B and C (lines 111-117).
 *
 * What differs underneath: the nonzeros of the panel are pooled on the GPUs -
 * every rank uploads its own slice once and the slices are exchanged with one
 * grouped NCCL send/recv over NVLink (the reference's two concurrent
 * MPI_Iallgatherv, lines 81-83) - and come back to the host only because the
 * row-parallel plan is built there.  When ranks share a GPU (NCCL cannot be
 * used) the pooling runs over MPI exactly like the reference.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mpi.h>

#include "utils.h"
#include "para2d_spmm.h"
#include "crp_ext.h"
#include "crp_internal.h"

static void *xmalloc(size_t bytes)
{
    void *p = malloc(bytes > 0 ? bytes : 1);
    ASSERT_PRINTF(p != NULL, "Failed to allocate %zu bytes of work memory for para2d_spmm\n", bytes);
    return p;
}

/* pool colidx / val of the grid row on the devices; results land in the host arrays */
static void pool_panel_nccl(
    MPI_Comm comm, const int pi, const int pj, const int pn, const int *nnz_displs,
    const int *A_colidx, const double *A_val, int *panel_colidx, double *panel_val, double *t_dev, size_t *recv_bytes
)
{
    const size_t tot = (size_t) nnz_displs[pn];
    const size_t mine = (size_t) (nnz_displs[pj + 1] - nnz_displs[pj]);
    crp_nccl_comm *nc = crp_nccl_get(comm);
    if (tot == 0) return;
    void *d_col = NULL, *d_val = NULL;
    crp_cuda_malloc_dev(&d_col, sizeof(int) * tot);
    crp_cuda_malloc_dev(&d_val, sizeof(double) * tot);
    void *stream = crp_cuda_stream_create();
    if (mine > 0)
    {
        crp_cuda_memcpy_async(A_colidx, (char *) d_col + sizeof(int) * (size_t) nnz_displs[pj], sizeof(int) * mine, stream);
        crp_cuda_memcpy_async(A_val, (char *) d_val + sizeof(double) * (size_t) nnz_displs[pj], sizeof(double) * mine, stream);
    }
    /* a tiny exchange with the same peers first: NCCL sets its connections up lazily on the first transfer between two ranks
     * (about a second for a fresh communicator) - that is not replication time */
    {
        void *d_warm = NULL;
        crp_cuda_malloc_dev(&d_warm, 64 * (size_t) pn);
        crp_nccl_group_start();
        for (int j = 0; j < pn; j++)
        {
            if (j == pj) continue;
            crp_nccl_send(nc, (char *) d_warm + 64 * (size_t) pj, 32, pi * pn + j, stream);
            crp_nccl_recv(nc, (char *) d_warm + 64 * (size_t) j + 32, 32, pi * pn + j, stream);
        }
        crp_nccl_group_end();
        crp_cuda_stream_sync(stream);
        crp_cuda_free_dev(d_warm);
    }
    void *e0 = crp_cuda_event_create(), *e1 = crp_cuda_event_create();
    crp_cuda_event_record(e0, stream);
    crp_nccl_group_start();
    for (int j = 0; j < pn; j++)
    {
        if (j == pj) continue;
        const int peer = pi * pn + j;       /* rank in comm == rank in its NCCL communicator */
        const size_t theirs = (size_t) (nnz_displs[j + 1] - nnz_displs[j]);
        if (mine > 0)
        {
            crp_nccl_send(nc, (char *) d_col + sizeof(int) * (size_t) nnz_displs[pj], sizeof(int) * mine, peer, stream);
            crp_nccl_send(nc, (char *) d_val + sizeof(double) * (size_t) nnz_displs[pj], sizeof(double) * mine, peer, stream);
        }
        if (theirs > 0)
        {
            crp_nccl_recv(nc, (char *) d_col + sizeof(int) * (size_t) nnz_displs[j], sizeof(int) * theirs, peer, stream);
            crp_nccl_recv(nc, (char *) d_val + sizeof(double) * (size_t) nnz_displs[j], sizeof(double) * theirs, peer, stream);
        }
    }
    crp_nccl_group_end();
    crp_cuda_event_record(e1, stream);
    crp_cuda_memcpy_async(d_col, panel_colidx, sizeof(int) * tot, stream);
    crp_cuda_memcpy_async(d_val, panel_val, sizeof(double) * tot, stream);
    crp_cuda_stream_sync(stream);
    *t_dev += 1e-3 * crp_cuda_event_elapsed_ms(e0, e1);
    *recv_bytes += (sizeof(int) + sizeof(double)) * (tot - mine);
    crp_cuda_event_destroy(e0);
    crp_cuda_event_destroy(e1);
    crp_cuda_stream_destroy(stream);
    crp_cuda_free_dev(d_col);
    crp_cuda_free_dev(d_val);
}

void para2d_spmm_init(
    MPI_Comm comm, const int pm, const int pn, const int *A0_rowptr,
    const int *B_rowptr, const int *AC_rowptr, const int *BC_colptr,
    const int *A_rowptr, const int *A_colidx, const double *A_val,
    para2d_spmm_p *para2d_spmm
)
{
    (void) AC_rowptr;   /* implied by A0_rowptr, as in the reference */
    para2d_spmm_p eng = (para2d_spmm_p) calloc(1, sizeof(para2d_spmm_s));
    ASSERT_PRINTF(eng != NULL, "Failed to allocate para2d_spmm\n");
    eng->comm_glb = comm;

    /* 1. grid coordinates and the two sub-communicators */
    double t0 = get_wtime_sec();
    int rank;
    MPI_Comm_rank(comm, &rank);
    const int pi = rank / pn, pj = rank % pn;
    MPI_Comm comm_row;
    MPI_Comm_split(comm, pi, pj, &comm_row);
    MPI_Comm_split(comm, pj, pi, &eng->comm_col);
    eng->t_init += get_wtime_sec() - t0;

    /* 2. pool the A rows of this grid row */
    t0 = get_wtime_sec();
    const int my_nrow = A0_rowptr[rank + 1] - A0_rowptr[rank];
    const int my_nnz = A_rowptr[my_nrow] - A_rowptr[0];
    const int panel_srow = A0_rowptr[pi * pn];
    const int panel_nrow = A0_rowptr[(pi + 1) * pn] - panel_srow;
    int *panel_rowptr = (int *) xmalloc(sizeof(int) * ((size_t) panel_nrow + 1));
    int *panel_colidx = NULL;
    double *panel_val = NULL;
    if (pn > 1)
    {
        int *cnts   = (int *) xmalloc(sizeof(int) * (size_t) pn);
        int *displs = (int *) xmalloc(sizeof(int) * ((size_t) pn + 1));
        int *nnzs   = (int *) xmalloc(sizeof(int) * (size_t) pn);
        MPI_Allgather(&my_nnz, 1, MPI_INT, nnzs, 1, MPI_INT, comm_row);
        displs[0] = 0;
        for (int j = 0; j < pn; j++)
        {
            cnts[j] = A0_rowptr[pi * pn + j + 1] - A0_rowptr[pi * pn + j];
            displs[j + 1] = displs[j] + cnts[j];
        }
        /* row pointers keep the caller's (global) nnz offsets; the last entry is patched */
        MPI_Allgatherv(A_rowptr, cnts[pj], MPI_INT, panel_rowptr, cnts, displs, MPI_INT, comm_row);
        displs[0] = 0;
        for (int j = 0; j < pn; j++)
        {
            cnts[j] = nnzs[j];
            displs[j + 1] = displs[j] + cnts[j];
        }
        const int panel_nnz = displs[pn];
        panel_rowptr[panel_nrow] = panel_rowptr[0] + panel_nnz;
        panel_colidx = (int *) xmalloc(sizeof(int) * (size_t) panel_nnz);
        panel_val    = (double *) xmalloc(sizeof(double) * (size_t) panel_nnz);

        int wsize = 1, transport;
        MPI_Comm_size(MPI_COMM_WORLD, &wsize);
        GET_ENV_INT_VAR(transport, "CRP_SPMM_TRANSPORT", "transport", -1, 0, 1, 0);
        if (transport < 0) transport = (crp_device_ready() && wsize <= crp_cuda_device_count()) ? 0 : 1;
        if (transport == 0 && crp_device_ready())
        {
            pool_panel_nccl(comm, pi, pj, pn, displs, A_colidx, A_val, panel_colidx, panel_val, &eng->t_ag_A_dev, &eng->ag_A_recv_bytes);
        } else {
            MPI_Allgatherv(A_colidx, my_nnz, MPI_INT, panel_colidx, cnts, displs, MPI_INT, comm_row);
            MPI_Allgatherv(A_val, my_nnz, MPI_DOUBLE, panel_val, cnts, displs, MPI_DOUBLE, comm_row);
        }
        free(cnts);
        free(displs);
        free(nnzs);
    } else {
        panel_colidx = (int *) xmalloc(sizeof(int) * (size_t) my_nnz);
        panel_val    = (double *) xmalloc(sizeof(double) * (size_t) my_nnz);
        memcpy(panel_rowptr, A_rowptr, sizeof(int) * ((size_t) panel_nrow + 1));
        memcpy(panel_colidx, A_colidx, sizeof(int) * (size_t) my_nnz);
        memcpy(panel_val, A_val, sizeof(double) * (size_t) my_nnz);
    }
    eng->t_ag_A += get_wtime_sec() - t0;

    /* modelled replication volume: the last rank knows the global nnz (its row pointer end) */
    if (rank == pm * pn - 1)
    {
        const int glb_nnz = A_rowptr[my_nrow];
        unsigned long long cost = (unsigned long long) (size_t) ((double) glb_nnz * (double) (pn - 1) * 1.5);
        MPI_Send(&cost, 1, MPI_UNSIGNED_LONG_LONG, 0, 0, comm);
    }
    if (rank == 0)
    {
        unsigned long long cost = 0;
        MPI_Recv(&cost, 1, MPI_UNSIGNED_LONG_LONG, pm * pn - 1, 0, comm, MPI_STATUS_IGNORE);
        eng->rA_cost = (size_t) cost;
    }

    /* 3. the row-parallel engine of this grid column, on this rank's column slice of B and C */
    t0 = get_wtime_sec();
    const int loc_n = BC_colptr[pj + 1] - BC_colptr[pj];
    rp_spmm_init_on(panel_srow, panel_nrow, panel_rowptr, panel_colidx, panel_val, B_rowptr, loc_n, eng->comm_col, comm, &eng->rp_spmm);
    eng->t_init += get_wtime_sec() - t0;

    MPI_Comm_free(&comm_row);
    free(panel_rowptr);
    free(panel_colidx);
    free(panel_val);
    *para2d_spmm = eng;
}

void para2d_spmm_free(para2d_spmm_p *para2d_spmm)
{
    para2d_spmm_p eng = *para2d_spmm;
    if (eng == NULL) return;
    rp_spmm_free(&eng->rp_spmm);
    MPI_Comm_free(&eng->comm_col);
    free(eng);
    *para2d_spmm = NULL;
}

void para2d_spmm_exec(para2d_spmm_p para2d_spmm, const int BC_layout, const double *B, const int ldB, double *C, const int ldC)
{
    if (para2d_spmm == NULL) return;
    rp_spmm_exec(para2d_spmm->rp_spmm, BC_layout, B, ldB, C, ldC);
}

void para2d_spmm_exec_f32(para2d_spmm_p para2d_spmm, const int BC_layout, const float *B, const int ldB, float *C, const int ldC)
{
    if (para2d_spmm == NULL) return;
    rp_spmm_exec_f32(para2d_spmm->rp_spmm, BC_layout, B, ldB, C, ldC);
}

/* Same table, same row labels as the reference (src/para2d_spmm.c:151-198). */
void para2d_spmm_print_stat(para2d_spmm_p eng)
{
    if (eng == NULL) return;
    rp_spmm_p rp = eng->rp_spmm;
    double dummy0, dummy1;
    rp_spmm_device_times(rp, &dummy0, &dummy1);      /* folds the last exec's events into the counters */
    int rank, nproc;
    MPI_Comm_rank(eng->comm_glb, &rank);
    MPI_Comm_size(eng->comm_glb, &nproc);
    const int n_exec = rp->n_exec;
    if (n_exec == 0) return;
    unsigned long long recv = (unsigned long long) rp->rB_recv_size * (unsigned long long) rp->glb_n, recv_max = 0, recv_sum = 0;
    double raw[7] = { eng->t_init, eng->t_ag_A, rp->t_pack, rp->t_a2a, rp->t_unpack, rp->t_spmm, rp->t_exec };
    double tmax[7], tavg[7];
    MPI_Reduce(&recv, &recv_max, 1, MPI_UNSIGNED_LONG_LONG, MPI_MAX, 0, eng->comm_glb);
    MPI_Reduce(&recv, &recv_sum, 1, MPI_UNSIGNED_LONG_LONG, MPI_SUM, 0, eng->comm_glb);
    MPI_Reduce(raw, tmax, 7, MPI_DOUBLE, MPI_MAX, 0, eng->comm_glb);
    MPI_Reduce(raw, tavg, 7, MPI_DOUBLE, MPI_SUM, 0, eng->comm_glb);
    if (rank != 0) return;
    for (int i = 2; i < 7; i++)
    {
        tmax[i] /= n_exec;
        tavg[i] /= (double) n_exec * nproc;
    }
    tavg[1] /= nproc;
    printf("para2d_spmm_init() time = %.2f s\n", tmax[0]);
    printf("Total comm size for replicating A = %zu\n", eng->rA_cost);
    printf("Total comm size for replicating B = %zu\n", (size_t) recv_sum);
    printf("Total comm size for SpMM          = %zu\n", eng->rA_cost + (size_t) recv_sum);
    printf("-------------------- Runtime (s) --------------------\n");
    printf("                                     avg         max\n");
    printf("Replicate A matrix (once)         %6.3f      %6.3f\n", tavg[1], tmax[1]);
    printf("Pack B matrix for redistribution  %6.3f      %6.3f\n", tavg[2], tmax[2]);
    printf("Redistribute B matrix             %6.3f      %6.3f\n", tavg[3], tmax[3]);
    printf("Unpack received B matrix data     %6.3f      %6.3f\n", tavg[4], tmax[4]);
    printf("Local SpMM                        %6.3f      %6.3f\n", tavg[5], tmax[5]);
    printf("Total para2d_spmm_exec()          %6.3f      %6.3f\n", tavg[6], tmax[6]);
    printf("Replicate A + para2d_spmm_exec()  %6.3f      %6.3f\n", tavg[1] + tavg[6], tmax[1] + tmax[6]);
    printf("\n");
    fflush(stdout);
}

void para2d_spmm_clear_stat(para2d_spmm_p para2d_spmm)
{
    if (para2d_spmm == NULL) return;
    rp_spmm_clear_stat(para2d_spmm->rp_spmm);
}
