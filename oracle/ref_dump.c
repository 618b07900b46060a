/*
 * ref_dump - drive the UNMODIFIED reference library (oracle/_ref/libcrpspmm_ref.so)
 * on a binary CSR file and dump everything the parity tests pin:
 * the 2-D grid and splits, every public field of the rowpara_spmm plan on each
 * rank, rA_cost, and the local C block.  Test infrastructure (oracle side) only.
 *
 * It follows the reference drivers' flow step by step:
 *   mode "2d": examples/test_para2d_spmm.c:41-167  (partition on rank 0, bcast,
 *              rows scattered by A0_rowptr with GLOBAL nnz offsets in rowptr,
 *              fill_B, para2d_spmm_init, exec)
 *   mode "rp": examples/test_rp_spmm.c:41-145
 * but reads a binary CSR (every rank reads its own slice) instead of a .mtx.
 *
 * usage: minimpirun -np P ref_dump <csr.bin> <n> <ntest> <2d|rp> <dump-prefix|-> [layout] [warmup]
 *
 * Binary CSR: "CRPCSR1\0", int64 m, k, nnz, int32 rowptr[m+1], int32 colidx[nnz], f64 val[nnz].
 * Dump file <prefix>.r<rank>.bin: a sequence of records
 *   char name[24]; int64 elem_size; int64 count; payload
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <mpi.h>
#include <omp.h>

#include "utils.h"
#include "spmat_part.h"
#include "rowpara_spmm.h"
#include "para2d_spmm.h"

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

static void put_i(FILE *fp, const char *name, int v) { put(fp, name, sizeof(int), 1, &v); }

static void dump_rp(FILE *fp, rp_spmm_p rp)
{
    int np = rp->nproc, n = rp->glb_n;
    int nnz = rp->A_rowptr[rp->A_nrow];
    put_i(fp, "nproc", np);
    put_i(fp, "my_rank", rp->my_rank);
    put_i(fp, "glb_n", n);
    put_i(fp, "A_nrow", rp->A_nrow);
    put_i(fp, "rB_nrow", rp->rB_nrow);
    put_i(fp, "rB_self_src_offset", rp->rB_self_src_offset);
    put_i(fp, "rB_self_dst_offset", rp->rB_self_dst_offset);
    put_i(fp, "rB_self_nrow", rp->rB_self_nrow);
    put(fp, "A_rowptr", sizeof(int), (size_t) rp->A_nrow + 1, rp->A_rowptr);
    put(fp, "A_colidx", sizeof(int), (size_t) nnz, rp->A_colidx);
    put(fp, "A_val", sizeof(double), (size_t) nnz, rp->A_val);
    put(fp, "rB_self_src_ridxs", sizeof(int), (size_t) rp->rB_self_nrow, rp->rB_self_src_ridxs);
    put(fp, "rB_scnts", sizeof(int), (size_t) np, rp->rB_scnts);
    put(fp, "rB_sdispls", sizeof(int), (size_t) np + 1, rp->rB_sdispls);
    put(fp, "rB_sridxs", sizeof(int), n ? (size_t) (rp->rB_sdispls[np] / n) : 0, rp->rB_sridxs);
    put(fp, "rB_rcnts", sizeof(int), (size_t) np, rp->rB_rcnts);
    put(fp, "rB_rdispls", sizeof(int), (size_t) np + 1, rp->rB_rdispls);
    put(fp, "rB_rridxs", sizeof(int), n ? (size_t) (rp->rB_rdispls[np] / n) : 0, rp->rB_rridxs);
    unsigned long long rs = (unsigned long long) rp->rB_recv_size;
    put(fp, "rB_recv_size", sizeof(rs), 1, &rs);
}

/* B[i,j] = 0.19 * i + 0.24 * j on global indices (examples/test_utils.c:121-154, factors test_para2d_spmm.c:124) */
static void fill_B_ref(int layout, double *B, int ldB, int srow, int nrow, int scol, int ncol)
{
    for (int i = 0; i < nrow; i++)
        for (int j = 0; j < ncol; j++)
        {
            double v = (srow + i) * 0.19 + (scol + j) * 0.24;
            if (layout == 0) B[(size_t) i * ldB + j] = v; else B[(size_t) j * ldB + i] = v;
        }
}

int main(int argc, char **argv)
{
    if (argc < 6)
    {
        fprintf(stderr, "usage: %s <csr.bin> <n> <ntest> <2d|rp> <dump-prefix|-> [layout]\n", argv[0]);
        return 2;
    }
    int glb_n = atoi(argv[2]), n_test = atoi(argv[3]);
    int mode2d = (strcmp(argv[4], "2d") == 0);
    const char *prefix = argv[5];
    int layout = (argc >= 7) ? atoi(argv[6]) : 0;
    int n_warm = (argc >= 8) ? atoi(argv[7]) : 1;      /* the reference drivers do one untimed exec */
    if (n_warm < 1) n_warm = 1;

    int nproc, rank;
    MPI_Init(&argc, &argv);
    MPI_Comm_size(MPI_COMM_WORLD, &nproc);
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);

    FILE *fin = fopen(argv[1], "rb");
    if (fin == NULL) { fprintf(stderr, "cannot open %s\n", argv[1]); MPI_Abort(MPI_COMM_WORLD, 2); }
    char magic[8];
    int64_t dims[3];
    if (fread(magic, 1, 8, fin) != 8 || memcmp(magic, "CRPCSR1", 7) != 0 || fread(dims, sizeof(int64_t), 3, fin) != 3)
    { fprintf(stderr, "bad CSR file\n"); MPI_Abort(MPI_COMM_WORLD, 2); }
    int glb_m = (int) dims[0], glb_k = (int) dims[1];
    size_t glb_nnz = (size_t) dims[2];
    int *rowptr = (int *) malloc(sizeof(int) * ((size_t) glb_m + 1));
    if (fread(rowptr, sizeof(int), (size_t) glb_m + 1, fin) != (size_t) glb_m + 1) MPI_Abort(MPI_COMM_WORLD, 2);
    long col_off = ftell(fin);
    long val_off = col_off + (long) (sizeof(int) * glb_nnz);

    /* rank 0 plans, everyone learns the result (test_para2d_spmm.c:41-83 / test_rp_spmm.c:41-74) */
    int pm = nproc, pn = 1;
    size_t comm_cost = 0;
    int *A0_rowptr = NULL, *B_rowptr = NULL, *AC_rowptr = NULL, *BC_colptr = NULL;
    int *rb = (int *) malloc(sizeof(int) * (nproc + 1));
    double t_part = 0.0;
    if (rank == 0)
    {
        int *colidx = (int *) malloc(sizeof(int) * glb_nnz);
        fseek(fin, col_off, SEEK_SET);
        if (fread(colidx, sizeof(int), glb_nnz, fin) != glb_nnz) MPI_Abort(MPI_COMM_WORLD, 2);
        double st = get_wtime_sec();
        csr_mat_row_partition(glb_m, rowptr, nproc, rb);
        if (mode2d)
        {
            calc_spmm_part2d_from_1d(
                nproc, glb_m, glb_n, glb_k, rb, rowptr, colidx, 1,
                &pm, &pn, &comm_cost, &A0_rowptr, &B_rowptr, &AC_rowptr, &BC_colptr, 0
            );
        }
        t_part = get_wtime_sec() - st;
        free(colidx);
    }
    int pmn[2] = { pm, pn };
    MPI_Bcast(pmn, 2, MPI_INT, 0, MPI_COMM_WORLD);
    pm = pmn[0]; pn = pmn[1];
    MPI_Bcast(rb, nproc + 1, MPI_INT, 0, MPI_COMM_WORLD);
    if (mode2d)
    {
        if (rank != 0)
        {
            A0_rowptr = (int *) malloc(sizeof(int) * (nproc + 1));
            B_rowptr  = (int *) malloc(sizeof(int) * (pm + 1));
            AC_rowptr = (int *) malloc(sizeof(int) * (pm + 1));
            BC_colptr = (int *) malloc(sizeof(int) * (pn + 1));
        }
        MPI_Bcast(A0_rowptr, nproc + 1, MPI_INT, 0, MPI_COMM_WORLD);
        MPI_Bcast(B_rowptr,  pm + 1,    MPI_INT, 0, MPI_COMM_WORLD);
        MPI_Bcast(AC_rowptr, pm + 1,    MPI_INT, 0, MPI_COMM_WORLD);
        MPI_Bcast(BC_colptr, pn + 1,    MPI_INT, 0, MPI_COMM_WORLD);
    }

    /* my rows of A; rowptr slice keeps GLOBAL nnz offsets like scatter_csr_rows (test_utils.c:78-91) */
    const int *split = mode2d ? A0_rowptr : rb;
    int a_srow = split[rank], a_nrow = split[rank + 1] - split[rank];
    size_t a_nnz0 = (size_t) rowptr[a_srow], a_nnz = (size_t) rowptr[a_srow + a_nrow] - a_nnz0;
    int *loc_rowptr = (int *) malloc(sizeof(int) * ((size_t) a_nrow + 1));
    int *loc_colidx = (int *) malloc(sizeof(int) * (a_nnz ? a_nnz : 1));
    double *loc_val = (double *) malloc(sizeof(double) * (a_nnz ? a_nnz : 1));
    memcpy(loc_rowptr, rowptr + a_srow, sizeof(int) * ((size_t) a_nrow + 1));
    fseek(fin, col_off + (long) (sizeof(int) * a_nnz0), SEEK_SET);
    if (fread(loc_colidx, sizeof(int), a_nnz, fin) != a_nnz) MPI_Abort(MPI_COMM_WORLD, 2);
    fseek(fin, val_off + (long) (sizeof(double) * a_nnz0), SEEK_SET);
    if (fread(loc_val, sizeof(double), a_nnz, fin) != a_nnz) MPI_Abort(MPI_COMM_WORLD, 2);
    fclose(fin);

    /* B / C blocks */
    int *x_displs = (int *) malloc(sizeof(int) * (nproc + 1));
    int b_srow, b_nrow, c_nrow, bc_scol, bc_ncol;
    if (mode2d)
    {
        int pi = rank / pn, pj = rank % pn;
        b_srow = B_rowptr[pi];  b_nrow = B_rowptr[pi + 1] - b_srow;
        c_nrow = AC_rowptr[pi + 1] - AC_rowptr[pi];
        bc_scol = BC_colptr[pj]; bc_ncol = BC_colptr[pj + 1] - bc_scol;
    } else {
        if (glb_m == glb_k) memcpy(x_displs, rb, sizeof(int) * (nproc + 1));
        else { int tmp; for (int i = 0; i <= nproc; i++) calc_block_spos_size(glb_k, nproc, i, x_displs + i, &tmp); }
        b_srow = x_displs[rank]; b_nrow = x_displs[rank + 1] - b_srow;
        c_nrow = a_nrow;
        bc_scol = 0; bc_ncol = glb_n;
    }
    int ldB = (layout == 0) ? bc_ncol : b_nrow, ldC = (layout == 0) ? bc_ncol : c_nrow;
    double *B = (double *) malloc(sizeof(double) * ((size_t) b_nrow * bc_ncol + 1));
    double *C = (double *) malloc(sizeof(double) * ((size_t) c_nrow * bc_ncol + 1));
    fill_B_ref(layout, B, ldB, b_srow, b_nrow, bc_scol, bc_ncol);

    para2d_spmm_p p2d = NULL;
    rp_spmm_p rp = NULL;
    if (mode2d)
    {
        para2d_spmm_init(MPI_COMM_WORLD, pm, pn, A0_rowptr, B_rowptr, AC_rowptr, BC_colptr, loc_rowptr, loc_colidx, loc_val, &p2d);
        rp = p2d->rp_spmm;
    } else {
        rp_spmm_init(a_srow, a_nrow, loc_rowptr, loc_colidx, loc_val, x_displs, glb_n, MPI_COMM_WORLD, &rp);
    }

    /* warm-up + timed loop (test_para2d_spmm.c:151-165) */
    for (int it = 0; it < n_warm; it++) rp_spmm_exec(rp, layout, B, ldB, C, ldC);
    rp_spmm_clear_stat(rp);
    double t_sum = 0.0, t_min = 1e30, t_max = 0.0;
    for (int it = 0; it < n_test; it++)
    {
        MPI_Barrier(MPI_COMM_WORLD);
        double st = get_wtime_sec();
        rp_spmm_exec(rp, layout, B, ldB, C, ldC);
        MPI_Barrier(MPI_COMM_WORLD);
        double ut = get_wtime_sec() - st;
        t_sum += ut; if (ut < t_min) t_min = ut; if (ut > t_max) t_max = ut;
    }
    double t_spmm_loc = (rp->n_exec > 0) ? rp->t_spmm / rp->n_exec : 0.0, t_spmm_max = 0.0;
    MPI_Reduce(&t_spmm_loc, &t_spmm_max, 1, MPI_DOUBLE, MPI_MAX, 0, MPI_COMM_WORLD);
    if (rank == 0)
    {
        printf("REFDUMP grid %d %d comm_cost %zu part_s %.6f nnz %zu m %d k %d n %d nproc %d threads %d\n",
               pm, pn, comm_cost, t_part, glb_nnz, glb_m, glb_k, glb_n, nproc, omp_get_max_threads());
        if (n_test > 0)
            printf("REFDUMP exec_s min %.6f avg %.6f max %.6f local_spmm_max_s %.6f\n", t_min, t_sum / n_test, t_max, t_spmm_max);
        fflush(stdout);
    }
    if (mode2d) para2d_spmm_print_stat(p2d); else rp_spmm_print_stat(rp);

    if (strcmp(prefix, "-") != 0)
    {
        char fn[512];
        snprintf(fn, sizeof(fn), "%s.r%d.bin", prefix, rank);
        FILE *fp = fopen(fn, "wb");
        put_i(fp, "pm", pm);
        put_i(fp, "pn", pn);
        unsigned long long cc = (unsigned long long) comm_cost;
        put(fp, "comm_cost", sizeof(cc), 1, &cc);
        put(fp, "rb_displs0", sizeof(int), (size_t) nproc + 1, rb);
        if (mode2d)
        {
            put(fp, "A0_rowptr", sizeof(int), (size_t) nproc + 1, A0_rowptr);
            put(fp, "B_rowptr", sizeof(int), (size_t) pm + 1, B_rowptr);
            put(fp, "AC_rowptr", sizeof(int), (size_t) pm + 1, AC_rowptr);
            put(fp, "BC_colptr", sizeof(int), (size_t) pn + 1, BC_colptr);
            unsigned long long ra = (unsigned long long) p2d->rA_cost;
            put(fp, "rA_cost", sizeof(ra), 1, &ra);
        } else {
            put(fp, "x_displs", sizeof(int), (size_t) nproc + 1, x_displs);
        }
        dump_rp(fp, rp);
        put_i(fp, "layout", layout);
        put_i(fp, "ldC", ldC);
        put_i(fp, "C_nrow", c_nrow);
        put_i(fp, "C_ncol", bc_ncol);
        put(fp, "C", sizeof(double), (size_t) c_nrow * bc_ncol, C);
        fclose(fp);
    }

    if (mode2d) para2d_spmm_free(&p2d); else rp_spmm_free(&rp);
    free(rowptr); free(loc_rowptr); free(loc_colidx); free(loc_val); free(B); free(C);
    free(rb); free(x_displs); free(A0_rowptr); free(B_rowptr); free(AC_rowptr); free(BC_colptr);
    MPI_Finalize();
    return 0;
}
