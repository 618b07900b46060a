// Internal declarations shared by the .cu files of the thin CUDA layer.
#ifndef CRP_CUDA_INTERNAL_CUH
#define CRP_CUDA_INTERNAL_CUH

#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

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
enum { CRP_VARIANT_AUTO = 0, CRP_VARIANT_ROWSPLIT = 1, CRP_VARIANT_ROWGROUP = 2, CRP_VARIANT_MERGEPATH = 3, CRP_VARIANT_PANEL = 4 };

// Rows of the row-split kernel with more than CRP_LONG_ROW nonzeros are cut into segments of
// CRP_LONG_SEG nonzeros that are multiplied as independent "virtual rows" into a scratch matrix and
// summed per row in segment order by a second kernel (deterministic, no atomics).
enum { CRP_LONG_ROW = 1024, CRP_LONG_SEG = 512 };
struct crp_longrows
{
    int       nlong;            // rows that are split (0: nothing to do)
    int       nseg;             // segments in total
    int       nshort;           // rows left to the plain row-split launch
    int       *d_short;         // nshort row ids (NULL when nlong == 0)
    int       *d_seg_beg;       // nseg: first nonzero of each segment
    int       *d_seg_end;       // nseg: one past its last nonzero
    int       *d_long_row;      // nlong: row id
    int       *d_long_sptr;     // nlong + 1: segment range of each long row
    void      *d_scratch;       // nseg x n partial results
    size_t    scratch_bytes;
};

// row-group (register-blocked) decomposition, see spmm_rowgroup.cu
struct crp_rowgroup
{
    int       R;                // rows per group (0: not built / not worthwhile)
    int       exact;            // 1: every group's rows share one column list (the register-blocked kernels need this); 0: masked blocks, panel kernel only
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

// nnz-balanced chunks of spmm_mergepath.cu (see mergepath_build.hpp)
struct crp_mergepath
{
    int       nchunks, nlong, nseg;
    void      *d_desc;          // int4 per chunk
    int       *d_long_row;      // nlong
    int       *d_long_sptr;     // nlong + 1
    void      *d_scratch;       // nseg partial rows
    size_t    scratch_bytes;
};

// B-row-panel form of the row groups (spmm_panel.cu / panel_build.hpp): tiles of K groups whose B rows are staged
// once per thread block in shared memory by cp.async.bulk
struct crp_rowgroup_host;
struct crp_panel
{
    void      *host;            // crp_panel_host (structure kept for the lazily built fp32 records and the wait map)
    int       K, CR, EMAX, R;
    int       clustered;        // tiles formed by column overlap (1) or K consecutive groups (0)
    int       ntiles, nchunks;
    long long union_rows;       // B rows staged per pass over the matrix
    int       *d_tile_chunk_ptr;
    int       *d_ucol;
    void      *d_chunks[2];     // chunk descriptors, [0] fp64 records, [1] fp32 records
    unsigned char *d_meta[2];
    size_t    meta_bytes[2];
    unsigned  *d_chunk_need;    // multi-GPU: wait slots each chunk depends on (NULL: none)
    long long *d_trace;         // development aid (CRP_PANEL_TRACE), not owned
    int       nslot;
};

// what a kernel needs to wait for the neighbours' arrival flags itself (peer-memory transport)
struct crp_spmm_wait
{
    const unsigned *flags;      // this rank's arrival flags, one word per rank
    const int *wait_idx;        // device: wait slot -> flag index
    int       nwait;
    unsigned  epoch;
    long long timeout_ns;
    int       *err;             // pinned host word set on timeout
};

// the B rows this rank owes its neighbours (peer-memory transport): stored by the SpMM kernel itself when it can
struct crp_spmm_put
{
    int       nrow;             // rows to send (may be 0: only the flags are published)
    size_t    row_bytes;
    const int *ridx;            // device: row of X0 for each
    void *const *dst_rows;      // device: destination address of each (peer memory)
    unsigned int *const *flag_ptrs;   // device: this rank's arrival flag on each neighbour
    int       nflag;
    unsigned  epoch;
    unsigned int *counter;      // device word, zero between launches
    size_t    dst_off;          // bytes added to every destination (column block of wider rows)
};

// Reuse profile of the B rows over the row sweep (built once at plan creation, O(nnz)): the sweep is cut into blocks of
// CRP_REUSE_TILE rows; every use of a B row by a block other than the one that used it last is a "reuse" at distance
// d = blocks in between.  With it the exec decides whether the product is made in several column passes so that the
// window of B / C rows between a row's uses stays inside the 126 MB L2 (spmm_plan.cu: choose_passes).
enum { CRP_REUSE_TILE = 32, CRP_REUSE_BUCKETS = 32 };
struct crp_reuse
{
    long long ntile;                        // row blocks
    long long first_uses;                   // B rows touched for the first time (compulsory reads)
    long long reuses[CRP_REUSE_BUCKETS];    // bucket b: reuses at distance [2^b, 2^(b+1)) blocks
    double    union_per_tile;               // distinct B rows a block touches, on average
    double    new_per_tile;                 // first_uses / ntile
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
    // nnz-balanced handling of very long rows (power-law matrices), see spmm_longrow.cu
    crp_longrows lr;
    crp_rowgroup rg;
    crp_mergepath mp;
    int       mp_tried;         // the merge-path partition has been built (lazily, on the first launch that wants it)
    int       *h_rowptr;        // host copy of the row pointers (m + 1) for lazily built auxiliary structures
    crp_rowgroup_host *rg_host; // host copy of the row-group arrays (panel construction)
    crp_panel pn;
    crp_reuse reuse;
    int       passes_forced;    // > 0: number of column passes set by crp_cuda_spmm_set_passes (experiments / tests)
    int       last_passes;      // column passes of the last exec
    char      kernel_name[80];
};

void crp_rowgroup_build(crp_spmm_plan *plan, const int *rowptr, const int *colidx, const double *val, std::vector<int> *rest_out);
void crp_rowgroup_destroy(crp_spmm_plan *plan);
void crp_mergepath_build(crp_spmm_plan *plan, const int *rowptr);
void crp_mergepath_destroy(crp_spmm_plan *plan);
void crp_panel_build(crp_spmm_plan *plan);
void crp_panel_destroy(crp_spmm_plan *plan);
// recv_off[j] .. recv_off[j + 1]: rows of the receive buffer that come from the rank of wait slot j (nslot <= 32)
void crp_panel_set_wait_map(crp_spmm_plan *plan, const int nslot, const int *recv_off);
// rows: the row ids the row-split kernel is responsible for (NULL = all m rows)
void crp_longrows_build(crp_spmm_plan *plan, const int *rowptr, const int *rows, const int nrows);
void crp_longrows_destroy(crp_spmm_plan *plan);

#endif
