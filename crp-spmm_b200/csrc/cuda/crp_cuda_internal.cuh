// Internal declarations shared by the .cu files of the thin CUDA layer.
#ifndef CRP_CUDA_INTERNAL_CUH
#define CRP_CUDA_INTERNAL_CUH

#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#include "crp_cuda.h"

// abort-on-error, like the reference proxy's CUDA_RUNTIME_CHECK (deprecated/src/cuda_utils.h:23-34)
#define CRP_CUDA_CHECK(call)                                                              \
    do {                                                                                  \
        cudaError_t crp_err_ = (call);                                                    \
        if (crp_err_ != cudaSuccess)                                                      \
        {                                                                                 \
            fprintf(stderr, "[FATAL] CUDA error %s (%d) at %s:%d: %s\n",                  \
                    cudaGetErrorName(crp_err_), (int) crp_err_, __FILE__, __LINE__,       \
                    cudaGetErrorString(crp_err_));                                        \
            fflush(stderr);                                                               \
            abort();                                                                      \
        }                                                                                 \
    } while (0)

#define CRP_LAUNCH_CHECK() do { crp_count_launch(); CRP_CUDA_CHECK(cudaGetLastError()); } while (0)

void crp_count_launch();

static inline cudaStream_t as_stream(void *s) { return (cudaStream_t) s; }

// ---- SpMM plan: device CSR + auxiliary structures of the kernel variants ----
enum { CRP_VARIANT_AUTO = 0, CRP_VARIANT_ROWSPLIT = 1, CRP_VARIANT_ROWGROUP = 2, CRP_VARIANT_MERGEPATH = 3 };

// row-group (register-blocked) decomposition, see spmm_rowgroup.cu
struct crp_rowgroup
{
    int       R;                // rows per group (0: not built / not worthwhile)
    int       ngroups;          // groups stored as R x 1 column blocks
    long long nblk;             // blocks in total
    int       nrest;            // rows left to the row-split kernel
    long long rest_nnz;
    int       *d_grow;          // ngroups: first row of each group
    int       *d_gptr;          // ngroups + 1: block range of each group
    int       *d_bcol;          // nblk: column of each block
    double    *d_bval;          // nblk * R: values, the R rows of a block contiguous
    float     *d_bval32;        // fp32 copy, made on the first fp32 exec
    int       *d_rest;          // nrest: row ids for the row-split kernel
};

struct crp_spmm_plan
{
    int       m, k;
    int       x0_rows;          // X rows [0, x0_rows) come from X0, the rest from X1
    long long nnz;
    int       n_hint;
    int       variant;          // CRP_VARIANT_*
    const char *last_kernel;
    int       *d_rowptr;        // m + 1
    int       *d_colidx;        // nnz
    double    *d_val;           // nnz
    float     *d_val32;         // nnz, created on the first fp32 exec
    int       max_row_nnz;
    double    avg_row_nnz;
    // merge-path (nnz-balanced) decomposition, built on demand
    int       *d_mp_rowstart;   // per work item: first row
    int       mp_items, mp_chunk;
    crp_rowgroup rg;
    char      kernel_name[64];
};

void crp_rowgroup_build(crp_spmm_plan *plan, const int *rowptr, const int *colidx, const double *val);
void crp_rowgroup_destroy(crp_spmm_plan *plan);

#endif
