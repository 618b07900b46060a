/*
 * spmat_part.h - nnz-balanced 1-D row partition and the CRP 2-D grid / cost model.
 *
 * Replaces reference src/spmat_part.h; results are bit-exact with it:
 *   csr_mat_row_partition        <- src/spmat_part.c:12-35
 *   csr_mat_row_part_comm_size   <- src/spmat_part.c:38-64
 *   prime_factorization          <- src/spmat_part.c:66-81
 *   calc_spmm_part2d_from_1d     <- src/spmat_part.c:85-210
 * Pure integer host code (no MPI, no device).
 */
#ifndef CRPSPMM_SPMAT_PART_H
#define CRPSPMM_SPMAT_PART_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Split rows [0, nrow) into nblk contiguous blocks of roughly equal nnz.
 *   row_ptr  : nrow + 1 CSR row pointers, row_ptr[0] == 0
 *   rblk_ptr : out, nblk + 1 block boundaries
 * Block i ends at the row found by a binary search for (nnz / nblk) * (i + 1)
 * (nnz for the last block) that stops early on an exact hit. */
void csr_mat_row_partition(const int nrow, const int *row_ptr, const int nblk, int *rblk_ptr);

/* Ascending prime factors of n; *factors is malloc'd, the count is returned. */
int prime_factorization(int n, int **factors);

/* For every row block: number of distinct columns its rows touch, minus those
 * that fall in the block's own x range [x_displs[b], x_displs[b+1]).
 *   comm_sizes : out, nblk entries;  *total_size : out, their (int) sum */
void csr_mat_row_part_comm_size(
    const int nrow, const int ncol, const int *row_ptr, const int *col_idx,
    const int nblk, const int *rblk_ptr, const int *x_displs,
    int *comm_sizes, int *total_size
);

/* Choose the pm x pn process grid and all splits from a 1-D row partition.
 *   nproc, m, n, k : process count; A is m x k, B is k x n
 *   rb_displs0     : nproc + 1, 1-D row partition of A
 *   rowptr, colidx : global CSR pattern of A
 *   rA             : how many times A is reused (weights the B cost)
 * Outputs (arrays are malloc'd here, caller frees):
 *   *pm, *pn       : grid; rank r sits at (r / pn, r % pn)
 *   *comm_cost     : modelled communication volume of the chosen grid
 *   *A0_rowptr     : nproc + 1, initial 1-D ownership of A rows
 *   *B_rowptr      : pm + 1, row split of B
 *   *AC_rowptr     : pm + 1, row split of replicated A and of C
 *   *BC_colptr     : pn + 1, column split of B and C
 * P(i, j) starts with A rows A0_rowptr[i*pn+j .. i*pn+j+1) and B block
 * (B_rowptr[i..i+1), BC_colptr[j..j+1)), and computes the C block
 * (AC_rowptr[i..i+1), BC_colptr[j..j+1)). */
void calc_spmm_part2d_from_1d(
    const int nproc, const int m, const int n, const int k, const int *rb_displs0,
    const int *rowptr, const int *colidx, const int rA, int *pm, int *pn, size_t *comm_cost,
    int **A0_rowptr, int **B_rowptr, int **AC_rowptr, int **BC_colptr, int dbg_print
);

#ifdef __cplusplus
}
#endif

#endif
