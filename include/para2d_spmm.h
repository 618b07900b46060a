/*
 * para2d_spmm.h - 2-D (pm x pn) SpMM engine: pn column groups, each running a
 * pm-way row-parallel SpMM on its slice of the columns of B and C.
 *
 * Replaces reference src/para2d_spmm.h:
 *   para2d_spmm_init        <- src/para2d_spmm.c:20-127  (comm split, replicate A inside a grid row, rp_spmm_init on the grid column)
 *   para2d_spmm_free        <- src/para2d_spmm.c:130-138
 *   para2d_spmm_exec        <- src/para2d_spmm.c:141-148
 *   para2d_spmm_print_stat  <- src/para2d_spmm.c:151-198
 *   para2d_spmm_clear_stat  <- src/para2d_spmm.c:201-205
 * The replicate-A allgather (MPI_Iallgatherv x2 in the reference) runs as
 * grouped NCCL send/recv between the GPUs of a grid row.
 */
#ifndef CRPSPMM_PARA2D_SPMM_H
#define CRPSPMM_PARA2D_SPMM_H

#include "rowpara_spmm.h"

struct para2d_spmm
{
    rp_spmm_p rp_spmm;      /* the row-parallel engine of this rank's grid column        */
    MPI_Comm  comm_glb;     /* caller's communicator (borrowed)                          */
    MPI_Comm  comm_col;     /* ranks of the same grid column (owned, freed in _free)     */
    size_t    rA_cost;      /* modelled volume of replicating A (valid on rank 0)        */
    double    t_init;       /* seconds in para2d_spmm_init outside the A replication     */
    double    t_ag_A;       /* seconds replicating A                                     */
    /* ---- additions of this implementation ---- */
    double    t_ag_A_dev;   /* of which: the NCCL exchange between the GPUs of the grid row (device time, CUDA events) */
    size_t    ag_A_recv_bytes;  /* bytes of colidx / val this rank received in it                */
};
typedef struct para2d_spmm  para2d_spmm_s;
typedef struct para2d_spmm *para2d_spmm_p;

#ifdef __cplusplus
extern "C" {
#endif

/* Collective over comm.  Rank r is grid point (r / pn, r % pn).
 *   A0_rowptr : nproc + 1, rows of A initially owned by each rank
 *   B_rowptr  : pm + 1, row split of B          AC_rowptr : pm + 1, row split of A (replicated) and C
 *   BC_colptr : pn + 1, column split of B and C
 *   A_rowptr  : this rank's row pointers (may keep global nnz offsets), A_colidx / A_val its entries
 * P(i, j) afterwards holds A rows AC_rowptr[i..i+1), expects the B block
 * (B_rowptr[i..i+1), BC_colptr[j..j+1)) and produces the C block (AC_rowptr[i..i+1), BC_colptr[j..j+1)). */
void para2d_spmm_init(
    MPI_Comm comm, const int pm, const int pn, const int *A0_rowptr,
    const int *B_rowptr, const int *AC_rowptr, const int *BC_colptr,
    const int *A_rowptr, const int *A_colidx, const double *A_val,
    para2d_spmm_p *para2d_spmm
);

void para2d_spmm_free(para2d_spmm_p *para2d_spmm);

/* C := A * B on this rank's blocks; same layout / leading-dimension rules as rp_spmm_exec. */
void para2d_spmm_exec(
    para2d_spmm_p para2d_spmm, const int BC_layout, const double *B, const int ldB,
    double *C, const int ldC
);

void para2d_spmm_print_stat(para2d_spmm_p para2d_spmm);
void para2d_spmm_clear_stat(para2d_spmm_p para2d_spmm);

#ifdef __cplusplus
}
#endif

#endif
