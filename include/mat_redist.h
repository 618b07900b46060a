/*
 * mat_redist.h - redistribution of a dense matrix between two arbitrary
 * non-overlapping rectangular block layouts.
 *
 * Replaces reference src/mat_redist.h:
 *   mat_redist_engine_init            <- src/mat_redist.c:44-236  (rectangle-intersection plan, bit-exact)
 *   mat_redist_engine_attach_workbuf  <- src/mat_redist.c:239-267
 *   mat_redist_engine_free            <- src/mat_redist.c:270-295
 *   mat_redist_engine_exec            <- src/mat_redist.c:298-419 (pack -> exchange -> unpack)
 * DEV_TYPE_HOST moves host blocks through MPI (what the drivers' result check
 * uses); DEV_TYPE_CUDA / DEV_TYPE_CUDA_MPI_DIRECT move device blocks with one
 * batched pack kernel, grouped NCCL send/recv and one batched unpack kernel.
 * The struct keeps the reference's fields (the swapped comments on
 * src_scol/src_nrow/req_scol/req_nrow in the reference header are not repeated).
 */
#ifndef CRPSPMM_MAT_REDIST_H
#define CRPSPMM_MAT_REDIST_H

#include "mpi.h"
#include "dev_type.h"

struct mat_redist_engine
{
    MPI_Comm graph_comm;    /* neighbourhood communicator (host path)                    */
    MPI_Datatype dtype;     /* element MPI datatype                                      */
    size_t  dt_size;        /* element size in bytes                                     */
    int     nproc;          /* size of comm                                              */
    int     rank;           /* rank in comm                                              */
    int     src_srow;       /* first row of the block this rank holds                    */
    int     src_scol;       /* first column of the block this rank holds                 */
    int     src_nrow;       /* rows of the block this rank holds                         */
    int     src_ncol;       /* columns of the block this rank holds                      */
    int     req_srow;       /* first row of the block this rank wants                    */
    int     req_scol;       /* first column of the block this rank wants                 */
    int     req_nrow;       /* rows of the block this rank wants                         */
    int     req_ncol;       /* columns of the block this rank wants                      */
    int     n_proc_send;    /* ranks this rank sends a piece to (may include itself)     */
    int     n_proc_recv;    /* ranks this rank receives a piece from                     */
    int     send_cnt;       /* elements sent                                             */
    int     recv_cnt;       /* elements received                                         */
    int     alloc_workbuf;  /* 1 if the work buffer is owned by the engine               */
    int     *send_ranks;    /* n_proc_send, destination ranks (ascending)                */
    int     *send_sizes;    /* n_proc_send, elements per destination                     */
    int     *send_displs;   /* n_proc_send + 1, element offsets in the send buffer       */
    int     *sblk_sizes;    /* n_proc_send x 4: srow, scol, nrow, ncol of each piece     */
    int     *recv_ranks;    /* n_proc_recv, source ranks (ascending)                     */
    int     *recv_sizes;    /* n_proc_recv, elements per source                          */
    int     *recv_displs;   /* n_proc_recv + 1, element offsets in the receive buffer    */
    int     *rblk_sizes;    /* n_proc_recv x 4: srow, scol, nrow, ncol of each piece     */
    int     *send_info0;    /* scratch during init (NULL afterwards)                     */
    int     *recv_info0;    /* scratch during init (NULL afterwards)                     */
    void    *sendbuf_h;     /* host send buffer   (alias into workbuf_h)                 */
    void    *recvbuf_h;     /* host receive buffer (alias into workbuf_h)                */
    void    *sendbuf_d;     /* device send buffer  (alias into workbuf_d)                */
    void    *recvbuf_d;     /* device receive buffer (alias into workbuf_d)              */
    void    *workbuf_h;     /* host work buffer: [send | recv]                           */
    void    *workbuf_d;     /* device work buffer: [send | recv]                         */
    double  hd_trans_ms;    /* host<->device staging time of the last exec (always 0)    */
    dev_type_t dev_type;    /* where src / dst blocks live                               */

    /* ---- additions of this implementation ---- */
    void    *dev;           /* device-side plan (struct crp_redist_dev), NULL for DEV_TYPE_HOST */
};
typedef struct mat_redist_engine  mat_redist_engine_s;
typedef struct mat_redist_engine* mat_redist_engine_p;

#ifdef __cplusplus
extern "C" {
#endif

/* Collective over comm.
 *   src_{srow,scol,nrow,ncol} : the block this rank holds (blocks of different ranks must not overlap)
 *   req_{srow,scol,nrow,ncol} : the block this rank wants
 *   dtype, dt_size            : element type and its size in bytes
 *   dev_type                  : where the blocks live
 *   engine_                   : out; left untouched if dev_type is invalid
 *   workbuf_bytes             : NULL -> the engine allocates its own work buffer;
 *                               else the required size is returned and the caller attaches one */
void mat_redist_engine_init(
    const int src_srow, const int src_scol, const int src_nrow, const int src_ncol,
    const int req_srow, const int req_scol, const int req_nrow, const int req_ncol,
    MPI_Comm comm, MPI_Datatype dtype, const size_t dt_size, dev_type_t dev_type,
    mat_redist_engine_p *engine_, size_t *workbuf_bytes
);

/* Use caller-provided work buffers (workbuf_d may be NULL for DEV_TYPE_HOST,
 * workbuf_h may be NULL for DEV_TYPE_CUDA_MPI_DIRECT). */
void mat_redist_engine_attach_workbuf(mat_redist_engine_p engine, void *workbuf_h, void *workbuf_d);

/* dst_blk (leading dimension dst_ld) := the wanted block, assembled from the
 * src_blk (leading dimension src_ld) pieces of all ranks.  Row-major.  Collective. */
void mat_redist_engine_exec(
    mat_redist_engine_p engine, const void *src_blk, const int src_ld,
    void *dst_blk, const int dst_ld
);

void mat_redist_engine_free(mat_redist_engine_p *engine_);

#ifdef __cplusplus
}
#endif

#endif
