/*
 * dev_type.c - memory-space dispatch (include/dev_type.h lists the reference
 * lines).  The CUDA branches call the C-ABI layer in crp_cuda.h; host
 * allocations are pinned when a GPU is in use, because host buffers of this
 * library are staging mirrors of device buffers (as the reference intends,
 * src/dev_type.c:38-41).
 */
#include <stdlib.h>
#include <string.h>

#include "dev_type.h"
#include "crp_cuda.h"

int crp_device_ready(void);     /* crp_common.c: binds this process to its GPU on first use */

static int have_gpu(void)
{
    static int cached = -1;
    if (cached < 0)
    {
        cached = (crp_cuda_device_count() > 0) ? 1 : 0;
        if (cached) cached = crp_device_ready();
    }
    return cached;
}

static int is_cuda_type(dev_type_t t) { return (t == DEV_TYPE_CUDA) || (t == DEV_TYPE_CUDA_MPI_DIRECT); }

int is_dev_type_valid(dev_type_t dev_type)
{
    if (dev_type == DEV_TYPE_HOST) return 1;
    if (is_cuda_type(dev_type)) return have_gpu();
    return 0;
}

void *dev_type_malloc(size_t bytes, dev_type_t dev_type)
{
    void *mem = NULL;
    if (!is_dev_type_valid(dev_type))
    {
        ERROR_PRINTF("Invalid device type %d\n", dev_type);
        return NULL;
    }
    if (dev_type == DEV_TYPE_HOST)
    {
        if (have_gpu()) crp_cuda_malloc_host(&mem, bytes);
        else mem = malloc(bytes);
    } else {
        crp_cuda_malloc_dev(&mem, bytes);
    }
    if (bytes > 0 && mem == NULL) ERROR_PRINTF("Failed to malloc %zu bytes on device type %d\n", bytes, dev_type);
    return mem;
}

void dev_type_free(void *mem, dev_type_t dev_type)
{
    if (!is_dev_type_valid(dev_type))
    {
        ERROR_PRINTF("Invalid device type %d\n", dev_type);
        return;
    }
    if (mem == NULL) return;
    if (dev_type == DEV_TYPE_HOST)
    {
        if (have_gpu()) crp_cuda_free_host(mem);
        else free(mem);
    } else {
        crp_cuda_free_dev(mem);
    }
}

int dev_type_alloc_workbufs(dev_type_t dev_type, size_t bytes, void **workbuf_h, void **workbuf_d)
{
    *workbuf_h = NULL;
    *workbuf_d = NULL;
    const int want_h = (dev_type == DEV_TYPE_HOST) || (dev_type == DEV_TYPE_CUDA);
    const int want_d = is_cuda_type(dev_type);
    if (want_h)
    {
        *workbuf_h = dev_type_malloc(bytes, DEV_TYPE_HOST);
        if (*workbuf_h == NULL && bytes > 0) return 1;
    }
    if (want_d)
    {
        *workbuf_d = dev_type_malloc(bytes, DEV_TYPE_CUDA);
        if (*workbuf_d == NULL && bytes > 0)
        {
            if (*workbuf_h != NULL) dev_type_free(*workbuf_h, DEV_TYPE_HOST);
            *workbuf_h = NULL;
            return 2;
        }
    }
    return 0;
}

void dev_type_realloc(size_t *curr_bytes, size_t req_bytes, dev_type_t dev_type, void **mem)
{
    if (req_bytes <= *curr_bytes) return;
    dev_type_free(*mem, dev_type);
    *mem = dev_type_malloc(req_bytes, dev_type);
    *curr_bytes = (*mem != NULL) ? req_bytes : 0;
}

void dev_type_memset(void *mem, int value, size_t bytes, dev_type_t dev_type)
{
    if (!is_dev_type_valid(dev_type))
    {
        ERROR_PRINTF("Invalid device type %d\n", dev_type);
        return;
    }
    if (dev_type == DEV_TYPE_HOST) memset(mem, value, bytes);
    else crp_cuda_memset_dev(mem, value, bytes);
}

void dev_type_memcpy(void *dst, const void *src, size_t bytes, dev_type_t dst_dev_type, dev_type_t src_dev_type)
{
    if (!is_dev_type_valid(dst_dev_type) || !is_dev_type_valid(src_dev_type))
    {
        ERROR_PRINTF("Invalid dst device type %d or src device type %d\n", dst_dev_type, src_dev_type);
        return;
    }
    const int d = is_cuda_type(dst_dev_type), s = is_cuda_type(src_dev_type);
    if (!d && !s) memcpy(dst, src, bytes);
    else if (d && !s) crp_cuda_memcpy_h2d(src, dst, bytes);
    else if (!d && s) crp_cuda_memcpy_d2h(src, dst, bytes);
    else crp_cuda_memcpy_d2d(src, dst, bytes);
}

void dev_type_copy_matrix(
    size_t dt_size, const int nrow, const int ncol,
    const void *src, const int lds, void *dst, const int ldd,
    dev_type_t dev_type
)
{
    if (!is_dev_type_valid(dev_type))
    {
        ERROR_PRINTF("Invalid device type %d\n", dev_type);
        return;
    }
    if (dev_type == DEV_TYPE_HOST)
    {
        copy_matrix(dt_size, nrow, ncol, src, lds, dst, ldd, 1);
        return;
    }
    ASSERT_PRINTF(dt_size == 4 || dt_size == 8, "dt_size == 4 or 8 required for CUDA memory\n");
    crp_cuda_copy_matrix(dt_size, nrow, ncol, src, lds, dst, ldd);
}
