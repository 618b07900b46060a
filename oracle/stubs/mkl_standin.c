/*
 * Textbook restatement of the three MKL sparse-BLAS entry points the reference
 * calls (see mkl.h in this directory).  Test infrastructure only.
 *
 * Arithmetic: for every output row, products are accumulated in ascending
 * nonzero order into a zero-initialised accumulator, then combined as
 * alpha * acc + beta * C (beta == 0 overwrites without reading C, as MKL does).
 * Compiled with -ffp-contract=off so the summation is plain IEEE mul + add.
 */
#include <stdlib.h>
#include <string.h>
#include "mkl.h"

struct oracle_mkl_csr
{
    int rows, cols, base;
    const int *rs, *re, *ci;
    const double *v;
};

sparse_status_t mkl_sparse_d_create_csr(
    sparse_matrix_t *A, const sparse_index_base_t indexing, const MKL_INT rows, const MKL_INT cols,
    MKL_INT *rows_start, MKL_INT *rows_end, MKL_INT *col_indx, double *values
)
{
    struct oracle_mkl_csr *h = (struct oracle_mkl_csr *) malloc(sizeof(*h));
    h->rows = rows; h->cols = cols; h->base = (indexing == SPARSE_INDEX_BASE_ONE) ? 1 : 0;
    h->rs = rows_start; h->re = rows_end; h->ci = col_indx; h->v = values;
    *A = h;
    return SPARSE_STATUS_SUCCESS;
}

sparse_status_t mkl_sparse_d_mm(
    const sparse_operation_t operation, const double alpha, const sparse_matrix_t A,
    const struct matrix_descr descr, const sparse_layout_t layout, const double *B,
    const MKL_INT columns, const MKL_INT ldb, const double beta, double *C, const MKL_INT ldc
)
{
    if (operation != SPARSE_OPERATION_NON_TRANSPOSE || descr.type != SPARSE_MATRIX_TYPE_GENERAL)
        return SPARSE_STATUS_NOT_SUPPORTED;
    const int base = A->base;
    if (layout == SPARSE_LAYOUT_ROW_MAJOR)
    {
        #pragma omp parallel
        {
            double *acc = (double *) malloc(sizeof(double) * (size_t) (columns > 0 ? columns : 1));
            #pragma omp for schedule(dynamic, 64)
            for (int i = 0; i < A->rows; i++)
            {
                for (int j = 0; j < columns; j++) acc[j] = 0.0;
                for (int p = A->rs[i] - base; p < A->re[i] - base; p++)
                {
                    const double a = A->v[p];
                    const double *x = B + (size_t) (A->ci[p] - base) * (size_t) ldb;
                    for (int j = 0; j < columns; j++) acc[j] += a * x[j];
                }
                double *c = C + (size_t) i * (size_t) ldc;
                if (beta == 0.0) for (int j = 0; j < columns; j++) c[j] = alpha * acc[j];
                else             for (int j = 0; j < columns; j++) c[j] = alpha * acc[j] + beta * c[j];
            }
            free(acc);
        }
    } else {
        #pragma omp parallel for schedule(static) collapse(1)
        for (int j = 0; j < columns; j++)
        {
            const double *x = B + (size_t) j * (size_t) ldb;
            double *c = C + (size_t) j * (size_t) ldc;
            for (int i = 0; i < A->rows; i++)
            {
                double acc = 0.0;
                for (int p = A->rs[i] - base; p < A->re[i] - base; p++)
                    acc += A->v[p] * x[A->ci[p] - base];
                c[i] = (beta == 0.0) ? alpha * acc : alpha * acc + beta * c[i];
            }
        }
    }
    return SPARSE_STATUS_SUCCESS;
}

sparse_status_t mkl_sparse_destroy(sparse_matrix_t A)
{
    free(A);
    return SPARSE_STATUS_SUCCESS;
}
