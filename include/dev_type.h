/*
 * dev_type.h - memory-space abstraction (host / CUDA device).
 *
 * Replaces reference src/dev_type.h.  In the reference the CUDA branches are
 * dead code (src/dev_type.c:10 includes a header that only exists under
 * deprecated/); here DEV_TYPE_CUDA and DEV_TYPE_CUDA_MPI_DIRECT are live and
 * backed by the thin C-ABI layer in crp_cuda.h:
 *   is_dev_type_valid      <- src/dev_type.c:13-22
 *   dev_type_malloc        <- src/dev_type.c:25-51   (HOST allocations are pinned)
 *   dev_type_free          <- src/dev_type.c:54-74
 *   dev_type_realloc       <- src/dev_type.c:77-84
 *   dev_type_memset        <- src/dev_type.c:87-101
 *   dev_type_memcpy        <- src/dev_type.c:104-124
 *   dev_type_copy_matrix   <- src/dev_type.c:127-150
 *   MALLOC_ATTACH_WORKBUF  <- src/dev_type.h:63-88
 * Both CUDA types mean "device-resident"; the data plane is NCCL in either case.
 */
#ifndef CRPSPMM_DEV_TYPE_H
#define CRPSPMM_DEV_TYPE_H

#include <stdint.h>
#include "utils.h"

#include <mpi.h>

typedef enum
{
    DEV_TYPE_HOST = 0,          /* host memory                                   */
    DEV_TYPE_CUDA,              /* device memory                                 */
    DEV_TYPE_CUDA_MPI_DIRECT    /* device memory, transfers go device-to-device  */
} dev_type_t;

#ifdef __cplusplus
extern "C" {
#endif

/* 1 if dev_type can be used in this process (CUDA types need a visible GPU), else 0. */
int is_dev_type_valid(dev_type_t dev_type);

/* Allocate / release `bytes` bytes in the given memory space. */
void *dev_type_malloc(size_t bytes, dev_type_t dev_type);
void dev_type_free(void *mem, dev_type_t dev_type);

/* Grow-only reallocation: no-op when *curr_bytes >= req_bytes; contents are not preserved. */
void dev_type_realloc(size_t *curr_bytes, size_t req_bytes, dev_type_t dev_type, void **mem);

void dev_type_memset(void *mem, int value, size_t bytes, dev_type_t dev_type);

void dev_type_memcpy(void *dst, const void *src, size_t bytes, dev_type_t dst_dev_type, dev_type_t src_dev_type);

/* Copy an nrow x ncol block between row-major matrices living in `dev_type`
 * memory (lds / ldd in elements of dt_size bytes; dt_size must be 4 or 8 on CUDA). */
void dev_type_copy_matrix(
    size_t dt_size, const int nrow, const int ncol,
    const void *src, const int lds, void *dst, const int ldd,
    dev_type_t dev_type
);

#ifdef __cplusplus
}
#endif

/* Host and / or device work buffer for an engine of the given memory space: DEV_TYPE_HOST wants a host buffer,
 * DEV_TYPE_CUDA both (host mirror + device), DEV_TYPE_CUDA_MPI_DIRECT a device buffer only.
 * Returns 0 on success, 1 if the host allocation failed, 2 if the device allocation failed (nothing stays allocated). */
#ifdef __cplusplus
extern "C"
#endif
int dev_type_alloc_workbufs(dev_type_t dev_type, size_t bytes, void **workbuf_h, void **workbuf_d);

/* Same contract as the reference's macro of this name (src/dev_type.h:63-88): allocate what `dev_type` needs, hand it to
 * attach_func(engine, host, device); on failure report, release the engine with free_func(&engine) and attach nothing. */
#define MALLOC_ATTACH_WORKBUF(attach_func, free_func, engine, dev_type, workbuf_bytes, workbuf_h, workbuf_d)              \
    do {                                                                                                                  \
        void *crp_wb_h_ = NULL, *crp_wb_d_ = NULL;                                                                        \
        const int crp_wb_rc_ = dev_type_alloc_workbufs((dev_type), (workbuf_bytes), &crp_wb_h_, &crp_wb_d_);              \
        workbuf_h = crp_wb_h_;                                                                                            \
        workbuf_d = crp_wb_d_;                                                                                            \
        if (crp_wb_rc_ == 0) attach_func(engine, workbuf_h, workbuf_d);                                                   \
        else                                                                                                              \
        {                                                                                                                 \
            ERROR_PRINTF("Allocate %s workbuf failed\n", crp_wb_rc_ == 1 ? "host" : "CUDA");                              \
            free_func(&engine);                                                                                           \
        }                                                                                                                 \
    } while (0)

#endif
