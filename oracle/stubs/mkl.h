/*
 * Stand-in for <mkl.h> used ONLY to build the unmodified reference sources
 * into oracle/_ref (test infrastructure; the product never includes this).
 *
 * The reference delegates its local product to Intel MKL's inspector-executor
 * sparse BLAS (src/rowpara_spmm.c:398-408, examples/test_utils.c:157-178),
 * version unpinned (compiler flag -mkl).  MKL is not available offline, so the
 * three entry points the reference calls are restated in mkl_standin.c as the
 * textbook CSR x dense loop with the documented semantics
 *     C := alpha * A * X + beta * C,  A general, non-transposed, 0-based,
 *     4-array CSR (rows_start / rows_end), row- or column-major X and C.
 * Only the enumerators and types the reference names are declared.
 */
#ifndef ORACLE_MKL_STANDIN_H
#define ORACLE_MKL_STANDIN_H

#ifdef __cplusplus
extern "C" {
#endif

typedef int MKL_INT;

typedef enum { SPARSE_STATUS_SUCCESS = 0, SPARSE_STATUS_NOT_SUPPORTED = 6 } sparse_status_t;
typedef enum { SPARSE_INDEX_BASE_ZERO = 0, SPARSE_INDEX_BASE_ONE = 1 } sparse_index_base_t;
typedef enum { SPARSE_OPERATION_NON_TRANSPOSE = 10, SPARSE_OPERATION_TRANSPOSE = 11 } sparse_operation_t;
typedef enum { SPARSE_MATRIX_TYPE_GENERAL = 20, SPARSE_MATRIX_TYPE_SYMMETRIC = 21 } sparse_matrix_type_t;
typedef enum { SPARSE_FILL_MODE_LOWER = 40, SPARSE_FILL_MODE_UPPER = 41, SPARSE_FILL_MODE_FULL = 42 } sparse_fill_mode_t;
typedef enum { SPARSE_DIAG_NON_UNIT = 50, SPARSE_DIAG_UNIT = 51 } sparse_diag_type_t;
typedef enum { SPARSE_LAYOUT_ROW_MAJOR = 101, SPARSE_LAYOUT_COLUMN_MAJOR = 102 } sparse_layout_t;

struct matrix_descr
{
    sparse_matrix_type_t type;
    sparse_fill_mode_t   mode;
    sparse_diag_type_t   diag;
};

struct oracle_mkl_csr;
typedef struct oracle_mkl_csr *sparse_matrix_t;

sparse_status_t mkl_sparse_d_create_csr(
    sparse_matrix_t *A, const sparse_index_base_t indexing, const MKL_INT rows, const MKL_INT cols,
    MKL_INT *rows_start, MKL_INT *rows_end, MKL_INT *col_indx, double *values
);

sparse_status_t mkl_sparse_d_mm(
    const sparse_operation_t operation, const double alpha, const sparse_matrix_t A,
    const struct matrix_descr descr, const sparse_layout_t layout, const double *B,
    const MKL_INT columns, const MKL_INT ldb, const double beta, double *C, const MKL_INT ldc
);

sparse_status_t mkl_sparse_destroy(sparse_matrix_t A);

#ifdef __cplusplus
}
#endif

#endif
