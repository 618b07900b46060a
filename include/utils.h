/*
 * utils.h - small host helpers that are part of the CRP-SpMM public surface.
 *
 * Replaces reference src/utils.h (same names, argument meaning and behaviour;
 * the drivers examples/test_*.c call get_wtime_sec, calc_block_spos_size and
 * calc_err_2norm and use the *_PRINTF / GET_ENV_INT_VAR macros).
 *   get_wtime_sec          <- src/utils.c:15-22
 *   calc_block_spos_size   <- src/utils.c:26-48   (bit-exact: the partitioner depends on it)
 *   malloc_aligned / free  <- src/utils.c:51-62
 *   calc_2norm             <- src/utils.c:66-71
 *   calc_err_2norm         <- src/utils.c:75-89
 *   copy_matrix            <- src/utils.c:92-119
 *   print_matrix           <- src/utils.c:122-155
 *   dump_binary            <- src/utils.c:158-163
 */
#ifndef CRPSPMM_UTILS_H
#define CRPSPMM_UTILS_H

#include <stdio.h>
#include <stdlib.h>
#include <stddef.h>

#ifdef __cplusplus
#include <cassert>
extern "C" {
#else
#include <assert.h>
#endif

#define INT_MSIZE sizeof(int)
#define DBL_MSIZE sizeof(double)

#ifndef MIN
#define MIN(a, b)  ((a) < (b) ? (a) : (b))
#endif
#ifndef MAX
#define MAX(a, b)  ((a) > (b) ? (a) : (b))
#endif

/* "[LEVEL] file, line: message" on stdout (INFO) or stderr (everything else), flushed. */
#define CRP_LOG_(stream, level, fmt, ...) \
    do { fprintf(stream, "[" level "] %s, %d: " fmt, __FILE__, __LINE__, ##__VA_ARGS__); fflush(stream); } while (0)

#define INFO_PRINTF(fmt, ...)    CRP_LOG_(stdout, "INFO",    fmt, ##__VA_ARGS__)
#define DEBUG_PRINTF(fmt, ...)   CRP_LOG_(stderr, "DEBUG",   fmt, ##__VA_ARGS__)
#define WARNING_PRINTF(fmt, ...) CRP_LOG_(stderr, "WARNING", fmt, ##__VA_ARGS__)
#define ERROR_PRINTF(fmt, ...)   CRP_LOG_(stderr, "ERROR",   fmt, ##__VA_ARGS__)

/* Fatal check: message on stderr, then assert (abort). */
#define ASSERT_PRINTF(expr, fmt, ...)                       \
    do {                                                    \
        if (!(expr))                                        \
        {                                                   \
            CRP_LOG_(stderr, "FATAL", fmt, ##__VA_ARGS__);  \
            assert(expr);                                   \
            abort();                                        \
        }                                                   \
    } while (0)

/* var = atoi(getenv(env_str)) if set and inside [min_val, max_val], else default_val. */
#define GET_ENV_INT_VAR(var, env_str, var_str, default_val, min_val, max_val, print_info)   \
    do {                                                                                    \
        const char *crp_env_val_ = getenv(env_str);                                         \
        var = default_val;                                                                  \
        if (crp_env_val_ != NULL)                                                           \
        {                                                                                   \
            var = atoi(crp_env_val_);                                                       \
            if (var < (min_val) || var > (max_val)) var = default_val;                      \
            if ((print_info) && var != (default_val))                                       \
                INFO_PRINTF("Overriding parameter %s: %d (default) --> %d (runtime)\n",     \
                            var_str, default_val, var);                                     \
        }                                                                                   \
    } while (0)

/* Grow-only host buffer: reallocates (contents NOT preserved) when new_bytes > curr_bytes. */
#define REALLOC_BUFFER(ptr, ptr_type, curr_bytes, new_bytes)    \
    do {                                                        \
        if ((new_bytes) > (curr_bytes))                         \
        {                                                       \
            free(ptr);                                          \
            curr_bytes = (new_bytes);                           \
            ptr = (ptr_type) malloc(curr_bytes);                \
        }                                                       \
    } while (0)

/* Wall-clock time in seconds. */
double get_wtime_sec();

/* Even split of `len` items into `nblk` blocks: the first len % nblk blocks get
 * len / nblk + 1 items.  Returns start and size of block `iblk` (0 <= iblk <= nblk;
 * iblk == nblk gives start == len).  Out-of-range iblk: *blk_spos = -1, *blk_size = 0. */
void calc_block_spos_size(const int len, const int nblk, const int iblk, int *blk_spos, int *blk_size);

/* posix_memalign wrapper and its release. */
void *malloc_aligned(size_t size, size_t alignment);
void free_aligned(void *mem);

/* sqrt(sum x[i]^2), plain accumulation. */
double calc_2norm(const int len, const double *x);

/* *x0_2norm_ = ||x0||_2 and *err_2norm_ = ||x0 - x1||_2, plain accumulation in index order. */
void calc_err_2norm(const int len, const double *x0, const double *x1, double *x0_2norm_, double *err_2norm_);

/* Copy an nrow x ncol block between two row-major host matrices with leading
 * dimensions lds / ldd (in elements of dt_size bytes); use_omp != 0 threads the row loop. */
void copy_matrix(
    const size_t dt_size, const int nrow, const int ncol,
    const void *src, const int lds, void *dst, const int ldd, const int use_omp
);

/* Debug print: dtype 0 = int, 1 = double; stype 0 = row-major, 1 = column-major. */
void print_matrix(
    const int dtype, const int stype, const void *mat, const int ldm,
    const int nrow, const int ncol, const char *fmt, const char *name
);

/* Write `bytes` bytes of `data` to file `fname`. */
void dump_binary(const char *fname, void *data, const size_t bytes);

#ifdef __cplusplus
}
#endif

#endif
