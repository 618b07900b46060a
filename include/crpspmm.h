/*
 * crpspmm.h - the composite CRP-SpMM engine: C := A * B with A, B and C in the CALLER's layouts.
 *
 * Same entry points and argument lists as the reference's (deprecated) monolithic engine
 * deprecated/src/crpspmm.h:89-130, so deprecated/examples/test_crpspmm.c compiles unchanged:
 *   crpspmm_engine_init            <- deprecated/src/crpspmm.c:63-472
 *   crpspmm_engine_attach_workbuf  <- deprecated/src/crpspmm.c:475-519
 *   crpspmm_engine_exec            <- deprecated/src/crpspmm.c:522-689
 *   crpspmm_engine_free / print_stat / clear_stat <- deprecated/src/crpspmm.c:692-786
 * It is a composition of the live pieces instead of a second implementation:
 *   redistribute A (1-D row layout of the caller -> the cost model's initial ownership A0_rowptr)   MPI, host
 *   replicate A + build the row-parallel plans                                                    para2d_spmm_init
 *   redistribute B (caller's 2-D block -> the grid's B block)                                      mat_redist_engine (device, NCCL)
 *   replicate B + local SpMM                                                                       para2d_spmm_exec (device)
 *   redistribute C (the grid's C block -> caller's 2-D block)                                      mat_redist_engine (device, NCCL)
 * The process grid is the one calc_spmm_part2d_from_1d chooses (src/spmat_part.c:85-210), not the
 * deprecated engine's own heuristic (deprecated/src/crpspmm.c:133-195).
 * A's values are taken at exec time, as in the reference; the device plans are rebuilt only when the values changed.
 * B and C may be host or device pointers.  `use_CUDA` is accepted for compatibility and ignored: there is no CPU path.
 */
#ifndef CRPSPMM_CRPSPMM_H
#define CRPSPMM_CRPSPMM_H

#include <mpi.h>
#include "dev_type.h"
#include "mat_redist.h"
#include "para2d_spmm.h"

struct crpspmm_engine
{
    int    np_glb, rank_glb;            /* size of comm and this process's rank                                  */
    int    np_row, np_col;              /* process grid pm x pn chosen by the cost model                         */
    int    rank_row, rank_col;          /* this process's grid coordinates (rank / np_col, rank % np_col)        */
    int    glb_m, glb_n, glb_k;         /* A is m x k, B is k x n, C is m x n                                    */
    int    loc_A_srow, loc_A_erow;      /* rows [srow, erow) of A owned after the redistribution (A0 layout)     */
    int    loc_A_nrow, loc_A_nnz;
    int    loc_B_srow, loc_B_erow;      /* rows / columns of the B block the grid expects on this process        */
    int    loc_B_scol, loc_B_ecol;
    int    loc_B_nrow, loc_B_ncol;
    int    loc_C_srow, loc_C_nrow;      /* rows of the C block the grid produces on this process (columns = B's) */
    int    alloc_workbuf;               /* always 1: buffers are owned by the engine                             */
    int    use_CUDA;                    /* as passed by the caller (ignored)                                     */
    int    *loc_A_rowptr;               /* loc_A_nrow + 1, row pointers of the owned rows with GLOBAL nnz offsets */
    int    *loc_A_colidx;               /* loc_A_nnz global column indices                                        */
    double *loc_A_val;                  /* loc_A_nnz values of the last exec                                      */
    MPI_Comm comm_glb;                  /* the caller's communicator (borrowed)                                   */
    mat_redist_engine_p rd_B, rd_C;     /* redistribution engines for B and C (device)                            */
    para2d_spmm_p p2d;                  /* the 2-D engine (created at the first exec, when A's values are known)  */

    /* statistics (same meaning as the reference's) */
    int    n_exec;
    double t_init, t_exec;
    double t_rd_A, t_agv_A;
    double t_rd_B, t_a2a_B;
    double t_spmm, t_rd_C;
    double t_exec_nr;
    size_t nelem_A_rd, nelem_A_agv;
    size_t nelem_B_rd, nelem_B_a2av;
    size_t nelem_B_a2av_min;

    void   *priv;                       /* private state (struct crp_composite) */
};
typedef struct crpspmm_engine  crpspmm_engine_s;
typedef struct crpspmm_engine *crpspmm_engine_p;

#ifdef __cplusplus
extern "C" {
#endif

/* Collective over comm.  Arguments as in deprecated/src/crpspmm.h:62-96:
 *   m, n, k                  global sizes
 *   src_A_srow, src_A_nrow   rows of A this process holds (the processes' row ranges tile [0, m))
 *   src_A_rowptr             src_A_nrow + 1 row pointers (any base: differences are used)
 *   src_A_colidx             global column indices of those rows
 *   src_B_{srow,nrow,scol,ncol}  block of B this process holds
 *   dst_C_{srow,nrow,scol,ncol}  block of C this process wants
 *   workbuf_bytes            if not NULL it receives 0: this engine owns its (device) buffers */
void crpspmm_engine_init(
    const int m, const int n, const int k,
    const int src_A_srow, const int src_A_nrow,
    const int *src_A_rowptr, const int *src_A_colidx,
    const int src_B_srow, const int src_B_nrow,
    const int src_B_scol, const int src_B_ncol,
    const int dst_C_srow, const int dst_C_nrow,
    const int dst_C_scol, const int dst_C_ncol,
    MPI_Comm comm, int use_CUDA, crpspmm_engine_p *engine_, size_t *workbuf_bytes
);

/* Kept for source compatibility; the engine owns its buffers, the argument is ignored. */
void crpspmm_engine_attach_workbuf(crpspmm_engine_p engine, double *workbuf);

/* C := A * B (collective).  src_B: src_B_nrow x src_B_ncol row-major, leading dimension ldB; dst_C likewise with ldC. */
void crpspmm_engine_exec(
    crpspmm_engine_p engine,
    const int *src_A_rowptr, const int *src_A_colidx, const double *src_A_val,
    const double *src_B, const int ldB, double *dst_C, const int ldC
);

void crpspmm_engine_free(crpspmm_engine_p *engine_);
void crpspmm_engine_print_stat(crpspmm_engine_p engine);
void crpspmm_engine_clear_stat(crpspmm_engine_p engine);

/* Step 1 of exec on its own (host only, no GPU needed): move A's values into the owned-rows layout
 * (engine->loc_A_val).  Exposed so that the redistribution plan can be checked on a box without GPU. */
void crpspmm_engine_redist_A_values(crpspmm_engine_p engine, const double *src_A_val);

#ifdef __cplusplus
}
#endif

#endif
