/*
 * ref_redist_dump - run the UNMODIFIED reference mat_redist engine
 * (src/mat_redist.c:44-236 init, :298-419 exec) on a layout description and
 * dump its plan and the redistributed block.  Test infrastructure only.
 *
 * usage: minimpirun -np P ref_redist_dump <layout.txt> <dump-prefix>
 *
 * layout.txt: first line "P glb_nrow glb_ncol", then P lines
 *   src_srow src_scol src_nrow src_ncol req_srow req_scol req_nrow req_ncol
 * The global matrix is G[i,j] = i * 1000.5 + j (doubles); each rank fills its
 * source block from G and the destination block is dumped.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <mpi.h>

#include "utils.h"
#include "dev_type.h"
#include "mat_redist.h"

static void put(FILE *fp, const char *name, size_t esz, size_t cnt, const void *data)
{
    char nm[24];
    memset(nm, 0, sizeof(nm));
    strncpy(nm, name, sizeof(nm) - 1);
    int64_t hdr[2] = { (int64_t) esz, (int64_t) cnt };
    fwrite(nm, 1, sizeof(nm), fp);
    fwrite(hdr, sizeof(int64_t), 2, fp);
    if (cnt) fwrite(data, esz, cnt, fp);
}

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s <layout.txt> <dump-prefix>\n", argv[0]); return 2; }
    int nproc, rank;
    MPI_Init(&argc, &argv);
    MPI_Comm_size(MPI_COMM_WORLD, &nproc);
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);

    FILE *fin = fopen(argv[1], "r");
    int P, gr, gc;
    if (fin == NULL || fscanf(fin, "%d %d %d", &P, &gr, &gc) != 3 || P != nproc)
    { fprintf(stderr, "bad layout file or rank count\n"); MPI_Abort(MPI_COMM_WORLD, 2); }
    int r[8] = {0};
    for (int i = 0; i <= rank; i++)
        if (fscanf(fin, "%d %d %d %d %d %d %d %d", &r[0], &r[1], &r[2], &r[3], &r[4], &r[5], &r[6], &r[7]) != 8)
            MPI_Abort(MPI_COMM_WORLD, 2);
    fclose(fin);

    int src_ld = r[3] + 3, dst_ld = r[7] + 2;   /* deliberately padded leading dimensions */
    double *src = (double *) malloc(sizeof(double) * ((size_t) r[2] * src_ld + 1));
    double *dst = (double *) malloc(sizeof(double) * ((size_t) r[6] * dst_ld + 1));
    for (int i = 0; i < r[2]; i++)
        for (int j = 0; j < src_ld; j++)
            src[(size_t) i * src_ld + j] = (j < r[3]) ? (r[0] + i) * 1000.5 + (r[1] + j) : -7.0;
    for (size_t i = 0; i < (size_t) r[6] * dst_ld; i++) dst[i] = -1.0;

    mat_redist_engine_p e = NULL;
    mat_redist_engine_init(r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7],
                           MPI_COMM_WORLD, MPI_DOUBLE, sizeof(double), DEV_TYPE_HOST, &e, NULL);
    mat_redist_engine_exec(e, src, src_ld, dst, dst_ld);

    char fn[512];
    snprintf(fn, sizeof(fn), "%s.r%d.bin", argv[2], rank);
    FILE *fp = fopen(fn, "wb");
    put(fp, "n_proc_send", sizeof(int), 1, &e->n_proc_send);
    put(fp, "n_proc_recv", sizeof(int), 1, &e->n_proc_recv);
    put(fp, "send_cnt", sizeof(int), 1, &e->send_cnt);
    put(fp, "recv_cnt", sizeof(int), 1, &e->recv_cnt);
    put(fp, "send_ranks", sizeof(int), (size_t) e->n_proc_send, e->send_ranks);
    put(fp, "send_sizes", sizeof(int), (size_t) e->n_proc_send, e->send_sizes);
    put(fp, "send_displs", sizeof(int), (size_t) e->n_proc_send + 1, e->send_displs);
    put(fp, "sblk_sizes", sizeof(int), (size_t) e->n_proc_send * 4, e->sblk_sizes);
    put(fp, "recv_ranks", sizeof(int), (size_t) e->n_proc_recv, e->recv_ranks);
    put(fp, "recv_sizes", sizeof(int), (size_t) e->n_proc_recv, e->recv_sizes);
    put(fp, "recv_displs", sizeof(int), (size_t) e->n_proc_recv + 1, e->recv_displs);
    put(fp, "rblk_sizes", sizeof(int), (size_t) e->n_proc_recv * 4, e->rblk_sizes);
    put(fp, "dst_ld", sizeof(int), 1, &dst_ld);
    put(fp, "dst", sizeof(double), (size_t) r[6] * dst_ld, dst);
    fclose(fp);

    mat_redist_engine_free(&e);
    free(src); free(dst);
    MPI_Finalize();
    return 0;
}
