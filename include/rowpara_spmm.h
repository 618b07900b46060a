/*
 * rowpara_spmm.h - 1-D row-parallel SpMM engine, C := A * B.
 *
 * Replaces reference src/rowpara_spmm.h (same entry points, argument meaning
 * and struct fields):
 *   rp_spmm_init        <- src/rowpara_spmm.c:20-190   (host plan, bit-exact index lists and counts)
 *   rp_spmm_free        <- src/rowpara_spmm.c:193-209
 *   rp_spmm_exec        <- src/rowpara_spmm.c:212-422  (pack -> exchange -> [unpack, self copy] -> local SpMM)
 *   rp_spmm_print_stat  <- src/rowpara_spmm.c:425-464
 *   rp_spmm_clear_stat  <- src/rowpara_spmm.c:467-476
 * What changes underneath: the local product (MKL mkl_sparse_d_mm in the
 * reference) is a hand-written sm_100a CSR x dense kernel, the pack loop is a
 * gather kernel, the B-row exchange is grouped NCCL send/recv on device
 * buffers, and unpack / self copy disappear because the kernel reads the
 * caller's B and the receive buffer in place.  There is no CPU fallback:
 * rp_spmm_exec aborts if no CUDA device is available.
 *
 * B and C may be device pointers (zero-copy; e.g. from dev_type_malloc with
 * DEV_TYPE_CUDA) or host pointers (staged through device memory every call).
 *
 * The leading fields of struct rowpara_spmm are the reference's, in the
 * reference's order, and stay valid host arrays; device state hangs off `dev`.
 */
#ifndef CRPSPMM_ROWPARA_SPMM_H
#define CRPSPMM_ROWPARA_SPMM_H

#include <stddef.h>
#include <stdlib.h>
#include <mpi.h>

struct rowpara_spmm
{
    int    nproc, my_rank;      /* size of comm and this process's rank in it                   */
    int    glb_n;               /* number of columns of B and C                                 */
    int    A_nrow;              /* rows of the local A block                                    */
    int    rB_nrow;             /* rows of the gathered ("redistributed") B the local A indexes */
    int    rB_self_src_offset;  /* first needed row inside this rank's own B block              */
    int    rB_self_dst_offset;  /* where the own rows start inside rB                           */
    int    rB_self_nrow;        /* how many own B rows are needed                               */
    int    rB_p2p;              /* RP_SPMM_P2P: point-to-point (1) or all-to-all (0) exchange   */
    int    rB_reidx;            /* RP_SPMM_REIDX: compact rB to the needed rows only (1) or not */
    int    *A_rowptr;           /* A_nrow + 1, 0-based row pointers of the local A              */
    int    *A_colidx;           /* nnz, column indices into rB                                  */
    int    *rB_self_src_ridxs;  /* rB_self_nrow, global ids of the needed own rows              */
    int    *rB_scnts;           /* nproc, elements (rows * glb_n) sent to each rank             */
    int    *rB_sridxs;          /* local B row ids to send, grouped by destination              */
    int    *rB_sdispls;         /* nproc + 1, element offsets into the send buffer              */
    int    *rB_rcnts;           /* nproc, elements received from each rank                      */
    int    *rB_rridxs;          /* rB row positions of received rows, grouped by source         */
    int    *rB_rdispls;         /* nproc + 1, element offsets into the receive buffer           */
    double *A_val;              /* nnz, values of the local A                                   */
    MPI_Comm comm;              /* not duplicated; the caller keeps ownership                   */

    /* statistics */
    size_t rB_recv_size;        /* remote B rows received per exec                              */
    int    n_exec;              /* rp_spmm_exec calls since the last clear                      */
    double t_init;              /* seconds in rp_spmm_init                                      */
    double t_pack;              /* seconds packing B rows (device time)                         */
    double t_a2a;               /* seconds exchanging B rows (device time)                      */
    double t_unpack;            /* always 0: received rows are consumed in place                */
    double t_spmm;              /* seconds in the local SpMM kernel (device time)               */
    double t_exec;              /* seconds in rp_spmm_exec (host wall clock)                    */

    /* ---- fields below are additions of this implementation ---- */
    void   *dev;                /* device-side state (struct crp_rp_dev), NULL in plan-only mode */
};
typedef struct rowpara_spmm  rp_spmm_s;
typedef struct rowpara_spmm *rp_spmm_p;

#ifdef __cplusplus
extern "C" {
#endif

/* Build the engine for this rank's rows of A.
 *   A_srow, A_nrow : first global row (unused, as in the reference) and row count of the local A
 *   A_rowptr       : A_nrow + 1 row pointers; may carry a global nnz offset (A_rowptr[0] != 0)
 *   A_colidx/A_val : A_rowptr[A_nrow] - A_rowptr[0] global column indices / values (copied)
 *   B_row_displs   : nproc + 1, first B row owned by each rank
 *   glb_n          : columns of B and C
 *   comm           : communicator of the participating ranks (collective call)
 *   rp_spmm        : out, the engine */
void rp_spmm_init(
    const int A_srow, const int A_nrow, const int *A_rowptr, const int *A_colidx,
    const double *A_val, const int *B_row_displs, const int glb_n, MPI_Comm comm,
    rp_spmm_p *rp_spmm
);

/* Release an engine and set the pointer to NULL (NULL-safe). */
void rp_spmm_free(rp_spmm_p *rp_spmm);

/* C := A * B (collective over comm).
 *   BC_layout : 0 row-major, 1 column-major (both B and C)
 *   B, ldB    : this rank's B rows; ldB >= glb_n (row-major) or >= local B rows (column-major)
 *   C, ldC    : A_nrow x glb_n result, fully overwritten; ldC >= glb_n or >= A_nrow */
void rp_spmm_exec(
    rp_spmm_p rp_spmm, const int BC_layout, const double *B, const int ldB,
    double *C, const int ldC
);

/* Rank 0 prints the timing table (collective); clear resets the per-exec counters. */
void rp_spmm_print_stat(rp_spmm_p rp_spmm);
void rp_spmm_clear_stat(rp_spmm_p rp_spmm);

#ifdef __cplusplus
}
#endif

#endif
