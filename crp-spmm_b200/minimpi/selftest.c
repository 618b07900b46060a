/* minimpi self-test: exercises every call the library, drivers and oracle use. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mpi.h"

#define CHECK(cond, msg) do { if (!(cond)) { fprintf(stderr, "rank %d FAIL: %s (line %d)\n", rank, msg, __LINE__); MPI_Abort(MPI_COMM_WORLD, 3); } } while (0)

int main(int argc, char **argv)
{
    int rank, size;
    MPI_Init(&argc, &argv);
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    MPI_Comm_size(MPI_COMM_WORLD, &size);

    /* bcast + 4 concurrent ibcast */
    int v[4] = {0, 0, 0, 0};
    if (rank == 0) { v[0] = 11; v[1] = 22; v[2] = 33; v[3] = 44; }
    MPI_Request rq[4];
    for (int i = 0; i < 4; i++) MPI_Ibcast(&v[i], 1, MPI_INT, 0, MPI_COMM_WORLD, &rq[i]);
    MPI_Waitall(4, rq, MPI_STATUSES_IGNORE);
    CHECK(v[0] == 11 && v[3] == 44, "ibcast");

    /* ring p2p with tags = sender rank, big message to force partial writes */
    size_t big = 3u << 20;
    double *sb = (double *) malloc(sizeof(double) * big), *rb = (double *) malloc(sizeof(double) * big);
    for (size_t i = 0; i < big; i++) sb[i] = rank * 1000.0 + (double) (i % 97);
    int nxt = (rank + 1) % size, prv = (rank + size - 1) % size;
    MPI_Request pr[2];
    MPI_Irecv(rb, (int) big, MPI_DOUBLE, prv, prv, MPI_COMM_WORLD, &pr[0]);
    MPI_Isend(sb, (int) big, MPI_DOUBLE, nxt, rank, MPI_COMM_WORLD, &pr[1]);
    MPI_Waitall(2, pr, MPI_STATUSES_IGNORE);
    CHECK(rb[5] == prv * 1000.0 + 5.0 && rb[big - 1] == prv * 1000.0 + (double) ((big - 1) % 97), "ring");
    free(sb); free(rb);

    /* allgather / allgatherv */
    int *all = (int *) malloc(sizeof(int) * size);
    MPI_Allgather(&rank, 1, MPI_INT, all, 1, MPI_INT, MPI_COMM_WORLD);
    for (int i = 0; i < size; i++) CHECK(all[i] == i, "allgather");
    int *cnt = (int *) malloc(sizeof(int) * size), *dsp = (int *) malloc(sizeof(int) * (size + 1));
    dsp[0] = 0;
    for (int i = 0; i < size; i++) { cnt[i] = i + 1; dsp[i + 1] = dsp[i] + cnt[i]; }
    int *mine = (int *) malloc(sizeof(int) * (rank + 1)), *gv = (int *) malloc(sizeof(int) * dsp[size]);
    for (int i = 0; i <= rank; i++) mine[i] = rank * 10 + i;
    MPI_Request ag[1];
    MPI_Iallgatherv(mine, rank + 1, MPI_INT, gv, cnt, dsp, MPI_INT, MPI_COMM_WORLD, &ag[0]);
    MPI_Waitall(1, ag, MPI_STATUSES_IGNORE);
    for (int i = 0; i < size; i++) for (int j = 0; j <= i; j++) CHECK(gv[dsp[i] + j] == i * 10 + j, "allgatherv");

    /* alltoall / alltoallv */
    int *a2s = (int *) malloc(sizeof(int) * size), *a2r = (int *) malloc(sizeof(int) * size);
    for (int i = 0; i < size; i++) a2s[i] = rank * 100 + i;
    MPI_Alltoall(a2s, 1, MPI_INT, a2r, 1, MPI_INT, MPI_COMM_WORLD);
    for (int i = 0; i < size; i++) CHECK(a2r[i] == i * 100 + rank, "alltoall");
    {
        /* rank r sends (d+1) ints to d */
        int *sc = (int *) malloc(sizeof(int) * size), *sd = (int *) malloc(sizeof(int) * (size + 1));
        int *rc = (int *) malloc(sizeof(int) * size), *rd = (int *) malloc(sizeof(int) * (size + 1));
        sd[0] = rd[0] = 0;
        for (int i = 0; i < size; i++) { sc[i] = i + 1; sd[i + 1] = sd[i] + sc[i]; rc[i] = rank + 1; rd[i + 1] = rd[i] + rc[i]; }
        int *s = (int *) malloc(sizeof(int) * sd[size]), *r = (int *) malloc(sizeof(int) * rd[size]);
        for (int i = 0; i < size; i++) for (int j = 0; j < sc[i]; j++) s[sd[i] + j] = rank * 1000 + i * 10 + j;
        MPI_Alltoallv(s, sc, sd, MPI_INT, r, rc, rd, MPI_INT, MPI_COMM_WORLD);
        for (int i = 0; i < size; i++) for (int j = 0; j < rc[i]; j++) CHECK(r[rd[i] + j] == i * 1000 + rank * 10 + j, "alltoallv");
        free(sc); free(sd); free(rc); free(rd); free(s); free(r);
    }

    /* scatterv / gatherv */
    {
        int *src = NULL;
        if (rank == 0) { src = (int *) malloc(sizeof(int) * dsp[size]); for (int i = 0; i < dsp[size]; i++) src[i] = i; }
        int *dst = (int *) malloc(sizeof(int) * (rank + 1));
        MPI_Request sr;
        MPI_Iscatterv(src, cnt, dsp, MPI_INT, dst, rank + 1, MPI_INT, 0, MPI_COMM_WORLD, &sr);
        MPI_Waitall(1, &sr, MPI_STATUSES_IGNORE);
        for (int j = 0; j <= rank; j++) CHECK(dst[j] == dsp[rank] + j, "scatterv");
        int *back = (rank == 0) ? (int *) malloc(sizeof(int) * dsp[size]) : NULL;
        MPI_Gatherv(dst, rank + 1, MPI_INT, back, cnt, dsp, MPI_INT, 0, MPI_COMM_WORLD);
        if (rank == 0) for (int i = 0; i < dsp[size]; i++) CHECK(back[i] == i, "gatherv");
        free(src); free(dst); free(back);
    }

    /* reduce */
    {
        unsigned long long x = (unsigned long long) rank + 1, mx = 0, sm = 0;
        MPI_Reduce(&x, &mx, 1, MPI_UNSIGNED_LONG_LONG, MPI_MAX, 0, MPI_COMM_WORLD);
        MPI_Reduce(&x, &sm, 1, MPI_UNSIGNED_LONG_LONG, MPI_SUM, 0, MPI_COMM_WORLD);
        if (rank == 0) CHECK(mx == (unsigned long long) size && sm == (unsigned long long) size * (size + 1) / 2, "reduce");
        double d[2] = { rank * 0.5, -rank * 1.0 }, dm[2];
        MPI_Allreduce(d, dm, 2, MPI_DOUBLE, MPI_MAX, MPI_COMM_WORLD);
        CHECK(dm[0] == (size - 1) * 0.5 && dm[1] == 0.0, "allreduce");
    }

    /* comm split: rows / cols of a 2 x (size/2) grid, two comms on the same group in flight */
    if (size % 2 == 0)
    {
        int pn = size / 2, pi = rank / pn, pj = rank % pn;
        MPI_Comm row, row2, col;
        MPI_Comm_split(MPI_COMM_WORLD, pi, pj, &row);
        MPI_Comm_split(MPI_COMM_WORLD, pi, pj, &row2);
        MPI_Comm_split(MPI_COMM_WORLD, pj, pi, &col);
        int rr, rs, cr, cs;
        MPI_Comm_rank(row, &rr); MPI_Comm_size(row, &rs);
        MPI_Comm_rank(col, &cr); MPI_Comm_size(col, &cs);
        CHECK(rr == pj && rs == pn && cr == pi && cs == 2, "split");
        int *ra = (int *) malloc(sizeof(int) * pn), *rb2 = (int *) malloc(sizeof(int) * pn);
        int *c1 = (int *) malloc(sizeof(int) * pn), *d1 = (int *) malloc(sizeof(int) * pn);
        for (int i = 0; i < pn; i++) { c1[i] = 1; d1[i] = i; }
        MPI_Request q[2];
        int m1 = rank, m2 = -rank;
        MPI_Iallgatherv(&m1, 1, MPI_INT, ra, c1, d1, MPI_INT, row, &q[0]);
        MPI_Iallgatherv(&m2, 1, MPI_INT, rb2, c1, d1, MPI_INT, row2, &q[1]);
        MPI_Waitall(2, q, MPI_STATUSES_IGNORE);
        for (int i = 0; i < pn; i++) CHECK(ra[i] == pi * pn + i && rb2[i] == -(pi * pn + i), "iallgatherv on split comms");
        int s2 = rank, r2 = -1;
        MPI_Allreduce(&s2, &r2, 1, MPI_INT, MPI_SUM, col);
        CHECK(r2 == pj + (pn + pj), "allreduce on col");
        MPI_Comm_free(&row); MPI_Comm_free(&row2); MPI_Comm_free(&col);
        free(ra); free(rb2); free(c1); free(d1);
    }

    /* dist graph + neighbor alltoallv: everyone sends to (rank+1)%size and to itself */
    if (size > 1)
    {
        int srcs[2] = { prv, rank }, dsts[2] = { rank, nxt };
        MPI_Comm g;
        MPI_Dist_graph_create_adjacent(MPI_COMM_WORLD, 2, srcs, MPI_UNWEIGHTED, 2, dsts, MPI_UNWEIGHTED, MPI_INFO_NULL, 0, &g);
        double s[3] = { rank + 0.25, rank + 0.5, rank + 0.75 }, r[3] = { 0, 0, 0 };
        int sc[2] = { 1, 2 }, sd[2] = { 0, 1 }, rc[2] = { 2, 1 }, rd[2] = { 0, 2 };
        MPI_Neighbor_alltoallv(s, sc, sd, MPI_DOUBLE, r, rc, rd, MPI_DOUBLE, g);
        CHECK(r[0] == prv + 0.5 && r[1] == prv + 0.75 && r[2] == rank + 0.25, "neighbor alltoallv");
        MPI_Barrier(g);
        MPI_Comm_free(&g);
    }

    /* blocking send/recv last -> 0 */
    {
        unsigned long long cost = 123456789ULL;
        if (rank == size - 1) MPI_Send(&cost, 1, MPI_UNSIGNED_LONG_LONG, 0, 0, MPI_COMM_WORLD);
        if (rank == 0)
        {
            unsigned long long got = 0;
            MPI_Recv(&got, 1, MPI_UNSIGNED_LONG_LONG, size - 1, 0, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
            CHECK(got == cost, "send/recv");
        }
    }

    MPI_Barrier(MPI_COMM_WORLD);
    if (rank == 0) printf("minimpi selftest OK on %d ranks\n", size);
    free(all); free(cnt); free(dsp); free(mine); free(gv); free(a2s); free(a2r);
    MPI_Finalize();
    return 0;
}
