/*
 * utils.c - host helpers of the public surface (see include/utils.h for the
 * reference lines each one stands in for).  calc_block_spos_size is the only
 * one with a bit-exactness contract: the partitioner's even splits use it.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "utils.h"

double get_wtime_sec()
{
    struct timespec ts;
    clock_gettime(CLOCK_REALTIME, &ts);
    return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

void calc_block_spos_size(const int len, const int nblk, const int iblk, int *blk_spos, int *blk_size)
{
    if (iblk < 0 || iblk > nblk)
    {
        *blk_spos = -1;
        *blk_size = 0;
        return;
    }
    /* the first `extra` blocks are one element longer than the rest */
    const int base  = len / nblk;
    const int extra = len % nblk;
    const int longer = (iblk < extra);
    *blk_size = base + (longer ? 1 : 0);
    *blk_spos = longer ? iblk * (base + 1) : iblk * base + extra;
}

void *malloc_aligned(size_t size, size_t alignment)
{
    void *p = NULL;
    if (posix_memalign(&p, alignment, size) != 0) return NULL;
    return p;
}

void free_aligned(void *mem)
{
    free(mem);
}

double calc_2norm(const int len, const double *x)
{
    double s = 0.0;
    for (int i = 0; i < len; i++) s += x[i] * x[i];
    return sqrt(s);
}

void calc_err_2norm(const int len, const double *x0, const double *x1, double *x0_2norm_, double *err_2norm_)
{
    double ref2 = 0.0, err2 = 0.0;
    for (int i = 0; i < len; i++)
    {
        const double d = x0[i] - x1[i];
        ref2 += x0[i] * x0[i];
        err2 += d * d;
    }
    *x0_2norm_  = sqrt(ref2);
    *err_2norm_ = sqrt(err2);
}

void copy_matrix(
    const size_t dt_size, const int nrow, const int ncol,
    const void *src, const int lds, void *dst, const int ldd, const int use_omp
)
{
    const size_t row_bytes = dt_size * (size_t) ncol;
    const size_t sp = dt_size * (size_t) lds, dp = dt_size * (size_t) ldd;
    const char *s = (const char *) src;
    char *d = (char *) dst;
    if (row_bytes == 0 || nrow <= 0) return;
    if (sp == row_bytes && dp == row_bytes && !use_omp)
    {
        memcpy(d, s, row_bytes * (size_t) nrow);
        return;
    }
    #pragma omp parallel for schedule(static) if (use_omp)
    for (int r = 0; r < nrow; r++)
        memcpy(d + (size_t) r * dp, s + (size_t) r * sp, row_bytes);
}

void print_matrix(
    const int dtype, const int stype, const void *mat, const int ldm,
    const int nrow, const int ncol, const char *fmt, const char *name
)
{
    const size_t rs = (stype == 0) ? (size_t) ldm : 1, cs = (stype == 0) ? 1 : (size_t) ldm;
    printf("%s:\n", name);
    for (int i = 0; i < nrow; i++)
    {
        for (int j = 0; j < ncol; j++)
        {
            const size_t at = (size_t) i * rs + (size_t) j * cs;
            if (dtype == 0) printf(fmt, ((const int *) mat)[at]);
            if (dtype == 1) printf(fmt, ((const double *) mat)[at]);
        }
        printf("\n");
    }
}

void dump_binary(const char *fname, void *data, const size_t bytes)
{
    FILE *fp = fopen(fname, "wb");
    if (fp == NULL)
    {
        ERROR_PRINTF("Cannot open %s for writing\n", fname);
        return;
    }
    fwrite(data, 1, bytes, fp);
    fclose(fp);
}
