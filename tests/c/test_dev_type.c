/*
 * test_dev_type.c - behavioural test of include/dev_type.h and of the external work-buffer path of
 * include/mat_redist.h, written the way a caller of the reference would use them:
 *   dev_type_malloc / free / realloc / memset / memcpy / copy_matrix   (reference src/dev_type.c:13-150)
 *   MALLOC_ATTACH_WORKBUF                                              (reference src/dev_type.h:63-88)
 *   mat_redist_engine_init(..., &workbuf_bytes) + mat_redist_engine_attach_workbuf   (reference src/mat_redist.c:44-267)
 * Usage: minimpirun -np P test_dev_type.exe <dev_type: 0 host | 1 cuda | 2 cuda-direct>
 * Prints "DEVTYPE OK" on every rank on success; any mismatch aborts with a message.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mpi.h>

#include "dev_type.h"
#include "mat_redist.h"

#define CHECK(cond, ...) do { if (!(cond)) { fprintf(stderr, "CHECK failed %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); exit(3); } } while (0)

static int g_attached = 0, g_freed = 0;
static void fake_attach(void *engine, void *h, void *d) { (void) engine; (void) h; (void) d; g_attached++; }
static void fake_free(void **engine) { *engine = NULL; g_freed++; }

int main(int argc, char **argv)
{
    MPI_Init(&argc, &argv);
    int rank, nproc;
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    MPI_Comm_size(MPI_COMM_WORLD, &nproc);
    const dev_type_t dt = (argc > 1) ? (dev_type_t) atoi(argv[1]) : DEV_TYPE_HOST;

    CHECK(is_dev_type_valid(DEV_TYPE_HOST) == 1, "host type must be valid");
    CHECK(is_dev_type_valid((dev_type_t) 17) == 0, "unknown type must be invalid");
    CHECK(is_dev_type_valid(dt) == 1, "dev_type %d not usable here", (int) dt);

    /* malloc / memset / memcpy round trip through the space under test */
    const size_t N = 1000;
    double *h0 = (double *) dev_type_malloc(sizeof(double) * N, DEV_TYPE_HOST);
    double *h1 = (double *) dev_type_malloc(sizeof(double) * N, DEV_TYPE_HOST);
    void *m = dev_type_malloc(sizeof(double) * N, dt);
    CHECK(h0 && h1 && m, "allocation failed");
    for (size_t i = 0; i < N; i++) { h0[i] = 0.5 * (double) i + rank; h1[i] = -1.0; }
    dev_type_memcpy(m, h0, sizeof(double) * N, dt, DEV_TYPE_HOST);
    dev_type_memcpy(h1, m, sizeof(double) * N, DEV_TYPE_HOST, dt);
    CHECK(memcmp(h0, h1, sizeof(double) * N) == 0, "memcpy round trip differs");
    dev_type_memset(m, 0, sizeof(double) * N, dt);
    dev_type_memcpy(h1, m, sizeof(double) * N, DEV_TYPE_HOST, dt);
    for (size_t i = 0; i < N; i++) CHECK(h1[i] == 0.0, "memset: element %zu = %g", i, h1[i]);
    /* same-space copy */
    void *m2 = dev_type_malloc(sizeof(double) * N, dt);
    dev_type_memcpy(m, h0, sizeof(double) * N, dt, DEV_TYPE_HOST);
    dev_type_memcpy(m2, m, sizeof(double) * N, dt, dt);
    dev_type_memcpy(h1, m2, sizeof(double) * N, DEV_TYPE_HOST, dt);
    CHECK(memcmp(h0, h1, sizeof(double) * N) == 0, "same-space memcpy differs");

    /* copy_matrix: a 7 x 5 block between matrices with different leading dimensions, 8- and 4-byte elements */
    for (int es = 8; es >= 4; es -= 4)
    {
        const int nr = 7, nc = 5, lds = 9, ldd = 6;
        unsigned char *hs = (unsigned char *) h0, *hd = (unsigned char *) h1;
        for (size_t i = 0; i < (size_t) nr * lds * es; i++) hs[i] = (unsigned char) (i * 7 + 3);
        memset(hd, 0xEE, (size_t) nr * ldd * es);
        dev_type_memcpy(m, hs, (size_t) nr * lds * es, dt, DEV_TYPE_HOST);
        dev_type_memcpy(m2, hd, (size_t) nr * ldd * es, dt, DEV_TYPE_HOST);
        dev_type_copy_matrix((size_t) es, nr, nc, m, lds, m2, ldd, dt);
        dev_type_memcpy(hd, m2, (size_t) nr * ldd * es, DEV_TYPE_HOST, dt);
        for (int i = 0; i < nr; i++)
            for (int j = 0; j < ldd * es; j++)
            {
                const unsigned char want = (j < nc * es) ? hs[(size_t) i * lds * es + j] : 0xEE;
                CHECK(hd[(size_t) i * ldd * es + j] == want, "copy_matrix es=%d (%d, byte %d)", es, i, j);
            }
    }

    /* realloc: grow-only, keeps the pointer when the request fits */
    size_t cur = sizeof(double) * N;
    void *before = m;
    dev_type_realloc(&cur, cur / 2, dt, &m);
    CHECK(m == before && cur == sizeof(double) * N, "realloc must not shrink");
    dev_type_realloc(&cur, 4 * sizeof(double) * N, dt, &m);
    CHECK(m != NULL && cur == 4 * sizeof(double) * N, "realloc must grow");
    dev_type_memset(m, 0, cur, dt);      /* the whole new range is usable */

    /* MALLOC_ATTACH_WORKBUF: what each memory space gets, and the attach callback */
    {
        void *eng = (void *) &g_attached, *wh = NULL, *wd = NULL;
        MALLOC_ATTACH_WORKBUF(fake_attach, fake_free, eng, dt, 4096, wh, wd);
        CHECK(g_attached == 1 && g_freed == 0 && eng != NULL, "attach callback not called");
        CHECK((wh != NULL) == (dt == DEV_TYPE_HOST || dt == DEV_TYPE_CUDA), "host work buffer presence wrong for type %d", (int) dt);
        CHECK((wd != NULL) == (dt != DEV_TYPE_HOST), "device work buffer presence wrong for type %d", (int) dt);
        if (wh) { memset(wh, 1, 4096); dev_type_free(wh, DEV_TYPE_HOST); }
        if (wd) { dev_type_memset(wd, 1, 4096, DEV_TYPE_CUDA); dev_type_free(wd, DEV_TYPE_CUDA); }
    }

    /* mat_redist with a caller-provided work buffer: a G x G matrix from row blocks to column blocks (G[i][j] = 1000 i + j) */
    {
        const int G = 6 * nproc + 5;
        int rs, rn, cs, cn;
        calc_block_spos_size(G, nproc, rank, &rs, &rn);
        calc_block_spos_size(G, nproc, nproc - 1 - rank, &cs, &cn);
        mat_redist_engine_p eng = NULL;
        size_t wb = 0;
        mat_redist_engine_init(rs, 0, rn, G, 0, cs, G, cn, MPI_COMM_WORLD, MPI_DOUBLE, sizeof(double), dt, &eng, &wb);
        CHECK(eng != NULL, "engine init failed");
        CHECK(wb >= sizeof(double) * ((size_t) rn * G + (size_t) G * cn), "workbuf_bytes %zu too small", wb);
        void *wh = NULL, *wd = NULL;
        MALLOC_ATTACH_WORKBUF(mat_redist_engine_attach_workbuf, mat_redist_engine_free, eng, dt, wb, wh, wd);
        CHECK(eng != NULL, "attach released the engine");
        const int lds = G + 2, ldd = cn + 1;
        double *hs = (double *) malloc(sizeof(double) * (size_t) rn * lds), *hd = (double *) malloc(sizeof(double) * (size_t) G * ldd);
        for (int i = 0; i < rn; i++) for (int j = 0; j < lds; j++) hs[(size_t) i * lds + j] = (j < G) ? 1000.0 * (rs + i) + j : -5.0;
        for (size_t i = 0; i < (size_t) G * ldd; i++) hd[i] = -9.0;
        void *ds = dev_type_malloc(sizeof(double) * (size_t) rn * lds, dt), *dd = dev_type_malloc(sizeof(double) * (size_t) G * ldd, dt);
        dev_type_memcpy(ds, hs, sizeof(double) * (size_t) rn * lds, dt, DEV_TYPE_HOST);
        dev_type_memcpy(dd, hd, sizeof(double) * (size_t) G * ldd, dt, DEV_TYPE_HOST);
        for (int rep = 0; rep < 2; rep++) mat_redist_engine_exec(eng, ds, lds, dd, ldd);
        dev_type_memcpy(hd, dd, sizeof(double) * (size_t) G * ldd, DEV_TYPE_HOST, dt);
        for (int i = 0; i < G; i++)
            for (int j = 0; j < ldd; j++)
            {
                const double want = (j < cn) ? 1000.0 * i + (cs + j) : -9.0;
                CHECK(hd[(size_t) i * ldd + j] == want, "redist (%d, %d): %g != %g", i, j, hd[(size_t) i * ldd + j], want);
            }
        mat_redist_engine_free(&eng);       /* must not free the caller's work buffers */
        CHECK(eng == NULL, "free must clear the handle");
        if (wh) { memset(wh, 2, wb); dev_type_free(wh, DEV_TYPE_HOST); }
        if (wd) { dev_type_memset(wd, 2, wb, DEV_TYPE_CUDA); dev_type_free(wd, DEV_TYPE_CUDA); }
        dev_type_free(ds, dt);  dev_type_free(dd, dt);
        free(hs);  free(hd);
    }

    dev_type_free(m, dt);  dev_type_free(m2, dt);
    dev_type_free(h0, DEV_TYPE_HOST);  dev_type_free(h1, DEV_TYPE_HOST);
    MPI_Barrier(MPI_COMM_WORLD);
    printf("DEVTYPE OK rank %d of %d type %d\n", rank, nproc, (int) dt);
    MPI_Finalize();
    return 0;
}
