/*
 * crp_cuda.h - the thin C-ABI CUDA layer under the CRP-SpMM host code.
 *
 * Successor of the reference's (deprecated, host-only) CUDA proxy
 * deprecated/src/cuda_proxy.h:13-66: plain C linkage, plain pointers and sizes,
 * void returns, abort-on-error (like CUDA_RUNTIME_CHECK, deprecated/src/cuda_utils.h:23-34).
 * Where the proxy wrapped cudaMemcpy2D and one cusparseSpMM call, this layer
 * owns hand-written sm_100a kernels.  Entry point -> what it replaces:
 *   crp_cuda_select_device_by_local_rank <- select_cuda_device_by_mpi_local_rank  cuda_proxy.cu:38-46
 *   crp_cuda_set_device                  <- cuda_set_rt_dev_id                    cuda_proxy.cu:48-51
 *   crp_cuda_memcpy_{h2d,d2h,d2d,auto}   <- cuda_memcpy_*                         cuda_proxy.cu:53-72
 *   crp_cuda_malloc_{dev,host}, free_*   <- cuda_malloc_*, cuda_free_*            cuda_proxy.cu:74-97
 *   crp_cuda_memset_dev                  <- cuda_memset_dev
 *   crp_cuda_device_sync / stream_sync   <- cuda_device_sync / cuda_stream_sync   cuda_proxy.cu:99-106
 *   crp_cuda_copy_matrix                 <- cuda_copy_matrix (cudaMemcpy2D)       cuda_proxy.cu:108-118
 *   crp_cuda_csr_spmm_host               <- cuda_cusparse_csr_spmm                cuda_proxy.cu:122-182
 *   crp_cuda_spmm_plan_* / spmm_exec     <- mkl_sparse_d_create_csr / _mm / _destroy at src/rowpara_spmm.c:398-408
 *   crp_cuda_gather_rows                 <- the OpenMP pack loop                  src/rowpara_spmm.c:232-262
 *   crp_cuda_copy_blocks                 <- the per-block dev_type_copy_matrix loops src/mat_redist.c:327-348, 395-416
 *   crp_cuda_transpose                   <- (new) column-major <-> row-major staging for BC_layout = 1
 * `stream` arguments are cudaStream_t passed as void* (NULL = the legacy default stream).
 */
#ifndef CRPSPMM_CRP_CUDA_H
#define CRPSPMM_CRP_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- device management ---- */
int  crp_cuda_device_count(void);                 /* 0 when no driver / no GPU; never aborts       */
void crp_cuda_select_device_by_local_rank(void);  /* MINIMPI_LOCAL_RANK / LOCAL_RANK / ... % #GPUs  */
void crp_cuda_set_device(const int dev_id);
int  crp_cuda_get_device(void);
int  crp_cuda_sm_count(void);
int  crp_cuda_ptr_is_device(const void *ptr);     /* 1 if ptr is device (or managed) memory         */

/* ---- memory ---- */
void crp_cuda_malloc_dev(void **dptr_, const size_t bytes);
void crp_cuda_malloc_host(void **hptr_, const size_t bytes);   /* pinned */
void crp_cuda_free_dev(void *dptr);
void crp_cuda_free_host(void *hptr);
void crp_cuda_memset_dev(void *dptr, const int value, const size_t bytes);
void crp_cuda_memset_async(void *dptr, const int value, const size_t bytes, void *stream);
void crp_cuda_memcpy_h2d(const void *hptr, void *dptr, const size_t bytes);
void crp_cuda_memcpy_d2h(const void *dptr, void *hptr, const size_t bytes);
void crp_cuda_memcpy_d2d(const void *dptr_src, void *dptr_dst, const size_t bytes);
void crp_cuda_memcpy_auto(const void *src, void *dst, const size_t bytes);
void crp_cuda_memcpy_async(const void *src, void *dst, const size_t bytes, void *stream);
/* strided host<->device / device<->device copy of nrow rows of row_bytes bytes (cudaMemcpy2DAsync) */
void crp_cuda_memcpy2d_async(const void *src, size_t src_pitch, void *dst, size_t dst_pitch, size_t row_bytes, size_t nrow, void *stream);
/* pin an existing host range so async copies from / to it run at full PCIe speed; returns 1 on success */
int  crp_cuda_host_register(const void *hptr, const size_t bytes);
void crp_cuda_host_unregister(const void *hptr);

/* ---- streams and events ---- */
void  crp_cuda_device_sync(void);
void  crp_cuda_stream_sync(void *stream);
void *crp_cuda_stream_create(void);               /* non-blocking stream */
void *crp_cuda_stream_create_high_priority(void); /* non-blocking, highest priority: its CTAs are placed first */
void  crp_cuda_stream_destroy(void *stream);
void *crp_cuda_event_create(void);
void  crp_cuda_event_destroy(void *event);
void  crp_cuda_event_record(void *event, void *stream);
void  crp_cuda_event_sync(void *event);
int   crp_cuda_event_done(void *event);            /* 1 if the event has completed (cudaEventQuery) */
void  crp_cuda_stream_wait_event(void *stream, void *event);
float crp_cuda_event_elapsed_ms(void *start, void *stop);

/* ---- data-movement kernels (all row-major, element size 4 or 8 unless noted) ---- */

/* dst[0:nrow, 0:ncol] := src[0:nrow, 0:ncol]; blocking like the reference proxy */
void crp_cuda_copy_matrix(size_t dt_size, const int nrow, const int ncol, const void *src, const int lds, void *dst, const int ldd);
void crp_cuda_copy_matrix_async(size_t dt_size, const int nrow, const int ncol, const void *src, const int lds, void *dst, const int ldd, void *stream);

/* dst[i, 0:ncol] := src[ridx[i], 0:ncol], i < nrow; ridx is a device array */
void crp_cuda_gather_rows(size_t dt_size, const int nrow, const int ncol, const void *src, const int lds, const int *ridx_d, void *dst, const int ldd, void *stream);

/* one launch copying a list of rectangular byte blocks (device array of descriptors) */
typedef struct
{
    uint64_t src_off;     /* byte offset of the block's first element from src_base */
    uint64_t dst_off;     /* byte offset from dst_base                              */
    uint64_t src_pitch;   /* bytes between consecutive rows in the source           */
    uint64_t dst_pitch;   /* bytes between consecutive rows in the destination      */
    uint32_t nrow;        /* rows                                                   */
    uint32_t row_bytes;   /* bytes per row                                          */
} crp_copy_block;
void crp_cuda_copy_blocks(const crp_copy_block *blocks_d, const int nblk, const void *src_base, void *dst_base, void *stream);

/* dst (ncol x nrow, leading dimension ldd) := transpose of src (nrow x ncol, leading dimension lds) */
void crp_cuda_transpose(size_t dt_size, const int nrow, const int ncol, const void *src, const int lds, void *dst, const int ldd, void *stream);

/* ---- peer-memory exchange over NVLink (one process per GPU; replaces the MPI P2P ring / MPI_Alltoallv of
 * B rows at reference src/rowpara_spmm.c:280-308 with direct stores into the peers' receive buffers) ---- */
enum { CRP_IPC_HANDLE_BYTES = 64 };
int   crp_cuda_ipc_get_handle(void *dptr, void *handle64);      /* 1 on success (dptr must come from crp_cuda_malloc_dev) */
void *crp_cuda_ipc_open(const void *handle64);                  /* NULL on failure; enables peer access lazily            */
void  crp_cuda_ipc_close(void *peer_ptr);
/* dst_rows_d[i] (a device array of peer addresses) := src[ridx_d[i], 0:ncol]; one launch for all peers */
void  crp_cuda_put_rows(size_t dt_size, const int nrow, const int ncol, const void *src, const int lds, const int *ridx_d, void *const *dst_rows_d, void *stream);
/* crp_cuda_put_rows followed by crp_cuda_signal_peers in ONE launch (the last thread block to finish publishes the flags);
 * done_counter_d: one zero-initialised device word owned by the caller, reset by the kernel.  nrow == 0: only signals. */
void  crp_cuda_put_rows_signal(
    size_t dt_size, const int nrow, const int ncol, const void *src, const int lds, const int *ridx_d, void *const *dst_rows_d,
    unsigned int *const *flag_ptrs_d, const int nflag, const unsigned int epoch, unsigned int *done_counter_d, const size_t dst_off_bytes, void *stream
);
/* after the puts (same stream): *flag_ptrs_d[j] := epoch for j < nflag, with system-scope release */
void  crp_cuda_signal_peers(unsigned int *const *flag_ptrs_d, const int nflag, const unsigned int epoch, void *stream);
/* spin until every flags_d[wait_idx_d[j]] has reached epoch, j < nwait; on timeout (seconds) *err_d is set to 1 and the kernel returns */
void  crp_cuda_wait_flags(const unsigned int *flags_d, const int *wait_idx_d, const int nwait, const unsigned int epoch, const double timeout_s, int *err_d, void *stream);

/* ---- local SpMM ---- */
typedef struct crp_spmm_plan crp_spmm_plan;

/* Upload a 0-based CSR matrix (host arrays; m rows, column indices < k) and
 * prepare whatever auxiliary structures the kernels want.
 * The dense operand X of the product is given in two pieces at exec time:
 * rows [0, x0_rows) live in X0, rows [x0_rows, k) in X1 (x0_rows >= k: one piece).
 * n_hint is the expected dense width (0 = unknown). */
crp_spmm_plan *crp_cuda_spmm_plan_create(const int m, const int k, const int x0_rows, const int *rowptr_h, const int *colidx_h, const double *val_h, const int n_hint);
void crp_cuda_spmm_plan_destroy(crp_spmm_plan *plan);

/* C (m x n, row-major, ldc) := alpha * A * X + beta * C   (beta == 0: C is not read).
 * X row r is X0 + r * ldx0 for r < x0_rows, else X1 + (r - x0_rows) * ldx1 (both row-major).
 * elem_size 8: X, C are fp64.  elem_size 4: X, C are fp32 and A's values are used as fp32. */
void crp_cuda_spmm_exec(
    crp_spmm_plan *plan, const int n, const int elem_size, const double alpha,
    const void *X0, const int ldx0, const void *X1, const int ldx1,
    const double beta, void *C, const int ldc, void *stream
);
/* Same product, for the peer-memory transport of rp_spmm_exec (successor of the MPI_Waitall at reference
 * src/rowpara_spmm.c:301): X1 is a receive buffer that the neighbours' GPUs fill with NVLink stores; the
 * kernel waits for flags_d[wait_idx_d[j]] >= epoch (j < nwait) itself, and - when the plan has a wait map -
 * only right before the first piece of work that reads a row of that neighbour, so everything that needs
 * the rank's own B rows only runs while the exchange is still in flight.  nwait == 0: plain exec.
 * On timeout *err (pinned host memory) is set to 1. */
void crp_cuda_spmm_exec_wait(
    crp_spmm_plan *plan, const int n, const int elem_size, const double alpha,
    const void *X0, const int ldx0, const void *X1, const int ldx1,
    const double beta, void *C, const int ldc,
    const unsigned int *flags_d, const int *wait_idx_d, const int nwait, const unsigned int epoch, const double timeout_s, int *err,
    void *stream
);
/* The whole exchange + product of one rp_spmm_exec in the peer-memory transport (successor of pack + MPI_Isend / Irecv /
 * Waitall + mkl_sparse_d_mm at reference src/rowpara_spmm.c:232-408) as ONE fused kernel where the plan has the B-row-panel
 * form: every thread block first stores its share of the rows the neighbours need (rows send_ridx_d of X0) straight into
 * their receive buffers over NVLink, the last block to finish publishes this rank's arrival flag (value epoch) on every
 * neighbour, and the product then runs as crp_cuda_spmm_exec_wait describes.  Plans without that form get the same effect
 * from separate launches (crp_cuda_put_rows_signal, wait, SpMM). */
typedef struct crp_exchange
{
    int       n_send_rows;                  /* rows this rank sends (0: it only publishes flags)            */
    const int *send_ridx_d;                 /* device: their row numbers in X0                               */
    void *const *dst_rows_d;                /* device: destination address of each row (peer memory)        */
    unsigned int *const *flag_ptrs_d;       /* device: this rank's arrival flag on each neighbour            */
    int       nflag;
    unsigned int *done_counter_d;           /* device word, zero between launches                            */
    const unsigned int *flags_d;            /* this rank's own flag words, one per rank                      */
    const int *wait_idx_d;                  /* device: ranks whose flag is waited for                        */
    int       nwait;
    unsigned int epoch;
    double    timeout_s;
    int       *err;                         /* pinned host word set to 1 on timeout                          */
    size_t    dst_off_bytes;                /* added to every destination address (a column block of wider rows) */
} crp_exchange;
void crp_cuda_spmm_exec_exchange(
    crp_spmm_plan *plan, const int n, const int elem_size, const double alpha,
    const void *X0, const int ldx0, const void *X1, const int ldx1, const double beta, void *C, const int ldc,
    const crp_exchange *xc, void *stream
);
/* wait map: rows [recv_off[j], recv_off[j + 1]) of X1 are written by the neighbour of wait slot j (nslot <= 32) */
void crp_cuda_spmm_set_wait_map(crp_spmm_plan *plan, const int nslot, const int *recv_off);
/* name of the kernel variant the last crp_cuda_spmm_exec on this plan launched (static string) */
const char *crp_cuda_spmm_last_kernel(const crp_spmm_plan *plan);
/* force a kernel variant for experiments: "auto", "rowsplit", "rowgroup", "panel", "mergepath" */
void crp_cuda_spmm_set_variant(crp_spmm_plan *plan, const char *name);
/* Column passes (opt-in): the product is made in several passes over column blocks of the dense operands (a multiple of 64
 * columns each), so that the window of live B / C rows fits a small L2.  On B200 one pass measured fastest everywhere, so
 * the default is 1; set_passes forces a count (0: default; CRP_SPMM_PASSES=<count> does the same for every plan,
 * CRP_SPMM_PASSES=model decides from the plan's reuse profile of the B rows), last_passes reports what the last exec used. */
void crp_cuda_spmm_set_passes(crp_spmm_plan *plan, const int passes);
int crp_cuda_spmm_last_passes(const crp_spmm_plan *plan);
/* host-only (no device needed): the number of passes the model picks for a CSR pattern, n columns of elem_size bytes, an L2 of l2_bytes */
int crp_cuda_spmm_model_passes(const int m, const int k, const int *rowptr_h, const int *colidx_h, const int n, const int elem_size, const double l2_bytes);

/* Device-side construction of the O(nnz) parts of the host plans (plan_build.cu); integers only, bit-identical to the host code.
 * Both return 1 when they did the work, 0 when the caller must run its host loop (empty input, bitmaps too large).
 *   crp_cuda_plan_needed_rows : colidx_out[i] = position of column colidx[i] among the distinct columns (reidx) or colidx[i] - lo;
 *                               *needed_rows = malloc'ed ascending list of the distinct columns (caller frees), lo / hi = their range
 *   crp_cuda_part_comm_size   : csr_mat_row_part_comm_size of include/spmat_part.h; the pattern is cached on the device between
 *                               calls with the same arrays until crp_cuda_part_cache_release() */
int crp_cuda_plan_needed_rows(const int *colidx_h, const long long nnz, const int glb_k, const int reidx,
                              int *colidx_out_h, int *lo, int *hi, int *n_needed, int **needed_rows);
int crp_cuda_part_comm_size(const int nrow, const int ncol, const int *row_ptr_h, const int *col_idx_h,
                            const int nblk, const int *rblk_ptr_h, const int *x_displs_h, int *comm_sizes_h, int *total_size);
void crp_cuda_part_cache_release(void);

/* what the plan holds: out[0] group size R (1: none), [1] groups, [2] R x 1 blocks, [3] rows left to the row-split kernel,
 * [4] their nonzeros, [5] panel tiles, [6] panel chunks, [7] B rows staged per pass (sum of the tiles' unions), [8] 1 if all groups
 * are exact, [9] long rows cut into segments, [10] nnz, [11] merge-path chunks */
void crp_cuda_spmm_plan_info(const crp_spmm_plan *plan, long long out[12]);

/* measured fp64 FMA throughput of the current device in TFLOP/s (pure DFMA issue; for the fp64 roofline of bench.py) */
double crp_cuda_measure_dfma_tflops(void);

/* Host-only view of the plan-time kernel selection (needs no device): the row-group size R (1 = row-split kernel only), the
 * alignment of the first group, the number of R x 1 blocks, and how many rows are long enough to be cut into segments. */
void crp_cuda_spmm_analyse(const int m, const int *rowptr, const int *colidx, int *R, int *offset, long long *nblk, int *n_long_rows);

/* Host-in / host-out convenience with the deprecated proxy's argument list:
 * C_h := alpha * A * B_h + beta * C_h, everything on the host, row-major fp64. */
void crp_cuda_csr_spmm_host(
    const int m, const int n, const int k, const double alpha,
    const int A_nnz, const int *A_rowptr_h, const int *A_colidx_h, const double *A_val_h,
    const double *B_h, const int ldB, const double beta, double *C_h, const int ldC
);

#ifdef __cplusplus
}
#endif

#endif
