/*
 * mmio_fast.c - fast ingest for the drivers: drop-in replacements of the two functions the reference's
 * drivers read their matrix with,
 *     mm_read_sparse_RPI   (reference examples/mmio_utils.c:11-125: fscanf of one entry per call)
 *     coo2csr              (reference examples/mmio_utils.c:148-190)
 * with the reference's signatures (examples/mmio_utils.h) and the reference's results - same entry order
 * in the COO arrays (file order, mirrored entries of a symmetric matrix appended in file order), same
 * CSR (rows sorted by column) - so examples/test_utils.c:read_mtx_csr and the drivers' main files are
 * compiled unchanged against it (Makefile: DRV_HELP).  What changes is the speed:
 *   * the file is mmap'ed and cut into one byte range per thread at line boundaries; every thread counts
 *     its lines, a prefix sum gives each range its place in the output, then the ranges are parsed
 *     concurrently (strtol / strtod on the mapped text: same correctly rounded doubles as fscanf("%lf"));
 *   * a file that starts with the 8-byte magic of the binary CSR format (pycrp/gen.py:write_csr_bin -
 *     int64 m, k, nnz, then rowptr, colidx, val) is not parsed at all: the arrays are copied out in
 *     parallel;
 *   * coo2csr counts and scatters with per-thread row-range ownership and sorts the rows in parallel;
 *     input that is already row-major sorted (what a binary CSR file yields) is detected and copied.
 * Not part of libcrpspmm: built as libcrpingest.so and linked by the drivers only.
 */
#define _GNU_SOURCE
#include <ctype.h>
#include <errno.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <omp.h>

static const char CRP_BIN_MAGIC[8] = { 'C', 'R', 'P', 'C', 'S', 'R', '1', '\0' };

int mm_read_sparse_RPI(const char *fname, const int need_symm, int *nrow_, int *ncol_, int *nnz_, int **row_, int **col_, double **val_);
void coo2csr(const int nrow, const int ncol, const int nnz, const int *row, const int *col, const double *val, int **row_ptr_, int **col_idx_, double **csr_val_);

static const char *skip_ws(const char *p, const char *end)
{
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\r')) p++;
    return p;
}

static const char *next_line(const char *p, const char *end)
{
    while (p < end && *p != '\n') p++;
    return (p < end) ? p + 1 : end;
}

/* a line that carries an entry: anything but blank */
static int line_has_entry(const char *p, const char *end)
{
    p = skip_ws(p, end);
    return p < end && *p != '\n';
}

static int word_is(const char *w, const char *s)
{
    for (; *s; w++, s++) if (tolower((unsigned char) *w) != *s) return 0;
    return *w == '\0';
}

/* binary CSR -> COO arrays in row-major order (what coo2csr then recognises as already sorted) */
static int read_binary(const char *map, const size_t size, const int need_symm, int *nrow_, int *ncol_, int *nnz_, int **row_, int **col_, double **val_)
{
    if (need_symm) { fprintf(stderr, "binary CSR files carry no symmetry flag; need_symm is not supported for them\n"); return -1; }
    if (size < 32) return -1;
    int64_t hdr[3];
    memcpy(hdr, map + 8, sizeof(hdr));
    const int64_t m = hdr[0], k = hdr[1], nnz = hdr[2];
    if (m < 0 || k < 0 || nnz < 0 || m > INT32_MAX - 1 || k > INT32_MAX || nnz > INT32_MAX) return -1;
    const size_t need = 32 + 4 * (size_t) (m + 1) + 4 * (size_t) nnz + 8 * (size_t) nnz;
    if (size < need) { fprintf(stderr, "binary CSR file is truncated (%zu of %zu bytes)\n", size, need); return -1; }
    const int32_t *rowptr = (const int32_t *) (map + 32);
    const int32_t *colidx = rowptr + (m + 1);
    const char *valp = (const char *) (colidx + nnz);       /* 8-byte aligned only if m + 1 + nnz is even: copy bytewise */
    int *row = (int *) malloc(sizeof(int) * (size_t) (nnz > 0 ? nnz : 1));
    int *col = (int *) malloc(sizeof(int) * (size_t) (nnz > 0 ? nnz : 1));
    double *val = (double *) malloc(sizeof(double) * (size_t) (nnz > 0 ? nnz : 1));
    if (!row || !col || !val) { free(row); free(col); free(val); return -1; }
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < m; i++)
        for (int32_t p = rowptr[i]; p < rowptr[i + 1]; p++) row[p] = (int) i;
    #pragma omp parallel
    {
        const int t = omp_get_thread_num(), nt = omp_get_num_threads();
        const size_t a = (size_t) nnz * (size_t) t / (size_t) nt, b = (size_t) nnz * (size_t) (t + 1) / (size_t) nt;
        memcpy(col + a, colidx + a, sizeof(int) * (b - a));
        memcpy(val + a, valp + 8 * a, sizeof(double) * (b - a));
    }
    *nrow_ = (int) m;  *ncol_ = (int) k;  *nnz_ = (int) nnz;
    *row_ = row;  *col_ = col;  *val_ = val;
    return 0;
}

int mm_read_sparse_RPI(
    const char *fname, const int need_symm, int *nrow_, int *ncol_, int *nnz_,
    int **row_, int **col_, double **val_
)
{
    const int fd = open(fname, O_RDONLY);
    if (fd < 0) return -1;
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size <= 0) { close(fd); return -1; }
    const size_t size = (size_t) st.st_size;
    const char *map = (const char *) mmap(NULL, size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map == MAP_FAILED) return -1;
    madvise((void *) map, size, MADV_SEQUENTIAL | MADV_WILLNEED);
    int rc = -1;
    const char *end = map + size;

    if (size >= 8 && memcmp(map, CRP_BIN_MAGIC, 8) == 0)
    {
        rc = read_binary(map, size, need_symm, nrow_, ncol_, nnz_, row_, col_, val_);
        munmap((void *) map, size);
        return rc;
    }

    /* ---- banner: %%MatrixMarket matrix coordinate {real|integer|pattern} {general|symmetric} ---- */
    char w[5][64];
    memset(w, 0, sizeof(w));
    {
        char line[1100];
        const char *le = next_line(map, end);
        size_t len = (size_t) (le - map);
        if (len >= sizeof(line)) len = sizeof(line) - 1;
        memcpy(line, map, len);
        line[len] = '\0';
        if (sscanf(line, "%63s %63s %63s %63s %63s", w[0], w[1], w[2], w[3], w[4]) != 5 || strcmp(w[0], "%%MatrixMarket") != 0)
        {
            printf("Could not process Matrix Market banner in file [%s]\n", fname);
            munmap((void *) map, size);
            return -1;
        }
    }
    const int is_real = word_is(w[3], "real"), is_int = word_is(w[3], "integer"), is_pat = word_is(w[3], "pattern");
    const int is_general = word_is(w[4], "general"), is_symm = word_is(w[4], "symmetric");
    int valid = (is_real || is_int || is_pat) && word_is(w[1], "matrix") && word_is(w[2], "coordinate") && (is_general || is_symm);
    if (need_symm && !is_symm)
    {
        fprintf(stderr, "The matrix is not symmetric.\n");
        munmap((void *) map, size);
        return -1;
    }
    if (!valid)
    {
        fprintf(stderr, "Does not support Market Market type: [%s %s %s %s]\n", w[1], w[2], w[3], w[4]);
        munmap((void *) map, size);
        return -1;
    }
    /* ---- comments, then the size line ---- */
    const char *p = next_line(map, end);
    while (p < end && (*p == '%' || !line_has_entry(p, end))) p = next_line(p, end);
    long nrow = 0, ncol = 0, nnz = 0;
    {
        char *q;
        nrow = strtol(p, &q, 10);  ncol = strtol(q, &q, 10);  nnz = strtol(q, &q, 10);
        if (nrow <= 0 || ncol <= 0 || nnz < 0 || nnz > INT32_MAX / 2)
        {
            fprintf(stderr, "Could not parse matrix size.\n");
            munmap((void *) map, size);
            return -1;
        }
        p = next_line(q, end);
    }
    const size_t cap = (size_t) nnz * (is_symm ? 2 : 1) + 1;
    int *row = (int *) malloc(sizeof(int) * cap), *col = (int *) malloc(sizeof(int) * cap);
    double *val = (double *) malloc(sizeof(double) * cap);
    if (!row || !col || !val) { free(row); free(col); free(val); munmap((void *) map, size); return -1; }

    /* ---- entries: one byte range per thread, cut at line starts ---- */
    int nt = omp_get_max_threads();
    if (nt > 256) nt = 256;
    if (nt < 1) nt = 1;
    const char *cut[257];
    long cnt[257];
    const size_t body = (size_t) (end - p);
    cut[0] = p;
    for (int t = 1; t < nt; t++)
    {
        const char *c = p + body * (size_t) t / (size_t) nt;
        if (c < cut[t - 1]) c = cut[t - 1];
        cut[t] = (c > p) ? next_line(c - 1, end) : p;      /* c - 1: a cut that already sits on a line start stays there */
        if (cut[t] < cut[t - 1]) cut[t] = cut[t - 1];
    }
    cut[nt] = end;
    long bad = 0;
    #pragma omp parallel num_threads(nt) reduction(+:bad)
    {
        const int t = omp_get_thread_num();
        long c = 0;
        for (const char *q = cut[t]; q < cut[t + 1]; q = next_line(q, cut[t + 1])) if (line_has_entry(q, cut[t + 1])) c++;
        cnt[t + 1] = c;
        #pragma omp barrier
        #pragma omp single
        {
            cnt[0] = 0;
            for (int i = 1; i <= nt; i++) cnt[i] += cnt[i - 1];
        }
        long o = cnt[t];
        for (const char *q = cut[t]; q < cut[t + 1] && o < nnz; q = next_line(q, cut[t + 1]))
        {
            if (!line_has_entry(q, cut[t + 1])) continue;
            char *e;
            const long r = strtol(q, &e, 10);
            const long cc = strtol(e, &e, 10);
            double v = 1.0;
            if (is_real) v = strtod(e, &e);
            else if (is_int) v = (double) (int) strtol(e, &e, 10);
            if (r < 1 || r > nrow || cc < 1 || cc > ncol) bad++;
            row[o] = (int) r - 1;  col[o] = (int) cc - 1;  val[o] = v;
            o++;
        }
    }
    munmap((void *) map, size);
    if (cnt[nt] < nnz || bad > 0)
    {
        fprintf(stderr, "Matrix Market file [%s]: %ld of %ld entries found, %ld with indices out of range\n", fname, cnt[nt] < nnz ? cnt[nt] : nnz, nnz, bad);
        free(row); free(col); free(val);
        return -1;
    }
    /* ---- symmetric: the mirrored entries follow, in file order (serial: a running output position) ---- */
    long total = nnz;
    if (is_symm)
    {
        for (long i = 0; i < nnz; i++)
            if (row[i] != col[i]) { row[total] = col[i]; col[total] = row[i]; val[total] = val[i]; total++; }
    }
    *nrow_ = (int) nrow;  *ncol_ = (int) ncol;  *nnz_ = (int) total;
    *row_ = row;  *col_ = col;  *val_ = val;
    return 0;
}

static int cmp_ci(const void *a, const void *b)
{
    const int x = *(const int *) a, y = *(const int *) b;
    return (x > y) - (x < y);
}

void coo2csr(
    const int nrow, const int ncol, const int nnz,
    const int *row, const int *col, const double *val,
    int **row_ptr_, int **col_idx_, double **csr_val_
)
{
    (void) ncol;
    int *row_ptr = (int *) malloc(sizeof(int) * ((size_t) nrow + 1));
    int *col_idx = (int *) malloc(sizeof(int) * (size_t) (nnz > 0 ? nnz : 1));
    double *csr_val = (double *) malloc(sizeof(double) * (size_t) (nnz > 0 ? nnz : 1));
    if (row_ptr == NULL || col_idx == NULL || csr_val == NULL)
    {
        fprintf(stderr, "Failed to allocate work arrays for %s\n", __func__);
        abort();
    }
    /* already row-major with ascending columns (binary CSR input, sorted Matrix Market files)? then it is a copy */
    int sorted = 1;
    #pragma omp parallel for reduction(&&:sorted) schedule(static)
    for (int i = 1; i < nnz; i++)
        if (row[i] < row[i - 1] || (row[i] == row[i - 1] && col[i] <= col[i - 1])) sorted = 0;
    memset(row_ptr, 0, sizeof(int) * ((size_t) nrow + 1));
    if (sorted)
    {
        #pragma omp parallel for schedule(static)
        for (int i = 0; i < nnz; i++)
        {
            col_idx[i] = col[i];
            csr_val[i] = val[i];
            if (i == 0 || row[i] != row[i - 1])
                for (int r = (i == 0 ? 0 : row[i - 1] + 1); r <= row[i]; r++) row_ptr[r] = i;
        }
        for (int r = (nnz > 0 ? row[nnz - 1] + 1 : 0); r <= nrow; r++) row_ptr[r] = nnz;
        *row_ptr_ = row_ptr;  *col_idx_ = col_idx;  *csr_val_ = csr_val;
        return;
    }
    /* counting sort by row; every thread owns a contiguous range of ROWS and scans all entries for them - no atomics, and
     * entries of a row keep their input order, as in the reference's serial bucket pass */
    int nt = omp_get_max_threads();
    if (nt > nnz / 65536 + 1) nt = nnz / 65536 + 1;
    for (int i = 0; i < nnz; i++) row_ptr[row[i] + 1]++;
    for (int i = 1; i <= nrow; i++) row_ptr[i] += row_ptr[i - 1];
    int *fill = (int *) malloc(sizeof(int) * ((size_t) nrow + 1));
    memcpy(fill, row_ptr, sizeof(int) * ((size_t) nrow + 1));
    #pragma omp parallel num_threads(nt)
    {
        const int t = omp_get_thread_num(), n_t = omp_get_num_threads();
        const int r0 = (int) ((long long) nrow * t / n_t), r1 = (int) ((long long) nrow * (t + 1) / n_t);
        for (int i = 0; i < nnz; i++)
        {
            const int r = row[i];
            if (r < r0 || r >= r1) continue;
            const int idx = fill[r]++;
            col_idx[idx] = col[i];
            csr_val[idx] = val[i];
        }
    }
    free(fill);
    /* sort every row by column (ties cannot occur in a valid file; the pairs are sorted through an index permutation) */
    #pragma omp parallel
    {
        int cap = 256;
        int *key = (int *) malloc(sizeof(int) * 2 * (size_t) cap);
        double *tmp = (double *) malloc(sizeof(double) * (size_t) cap);
        #pragma omp for schedule(dynamic, 512)
        for (int i = 0; i < nrow; i++)
        {
            const int b = row_ptr[i], len = row_ptr[i + 1] - b;
            int ok = 1;
            for (int j = 1; j < len; j++) if (col_idx[b + j] < col_idx[b + j - 1]) { ok = 0; break; }
            if (ok) continue;
            if (len > cap)
            {
                cap = len;
                key = (int *) realloc(key, sizeof(int) * 2 * (size_t) cap);
                tmp = (double *) realloc(tmp, sizeof(double) * (size_t) cap);
            }
            for (int j = 0; j < len; j++) { key[2 * j] = col_idx[b + j]; key[2 * j + 1] = j; tmp[j] = csr_val[b + j]; }
            qsort(key, (size_t) len, 2 * sizeof(int), cmp_ci);
            for (int j = 0; j < len; j++) { col_idx[b + j] = key[2 * j]; csr_val[b + j] = tmp[key[2 * j + 1]]; }
        }
        free(key);
        free(tmp);
    }
    *row_ptr_ = row_ptr;
    *col_idx_ = col_idx;
    *csr_val_ = csr_val;
}
