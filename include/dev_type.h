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

/* Allocate the host and/or device work buffer an engine asked for and attach it. */
#define MALLOC_ATTACH_WORKBUF(attach_func, free_func, engine, dev_type, workbuf_bytes, workbuf_h, workbuf_d) \
    do {                                                                                \
        int crp_need_h_ = ((dev_type) == DEV_TYPE_HOST) || ((dev_type) == DEV_TYPE_CUDA);               \
        int crp_need_d_ = ((dev_type) == DEV_TYPE_CUDA) || ((dev_type) == DEV_TYPE_CUDA_MPI_DIRECT);    \
        workbuf_h = NULL;                                                               \
        workbuf_d = NULL;                                                               \
        if (crp_need_h_)                                                                \
        {                                                                               \
            workbuf_h = dev_type_malloc(workbuf_bytes, DEV_TYPE_HOST);                  \
            if (workbuf_h == NULL)                                                      \
            {                                                                           \
                ERROR_PRINTF("Allocate host workbuf failed\n");                         \
                free_func(&engine);                                                     \
                break;                                                                  \
            }                                                                           \
        }                                                                               \
        if (crp_need_d_)                                                                \
        {                                                                               \
            workbuf_d = dev_type_malloc(workbuf_bytes, DEV_TYPE_CUDA);                  \
            if (workbuf_d == NULL)                                                      \
            {                                                                           \
                ERROR_PRINTF("Allocate CUDA workbuf failed\n");                         \
                free_func(&engine);                                                     \
                break;                                                                  \
            }                                                                           \
        }                                                                               \
        attach_func(engine, workbuf_h, workbuf_d);                                      \
    } while (0)

#endif
