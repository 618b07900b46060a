/*
 * rowpara_spmm.c - 1-D row-parallel SpMM engine (include/rowpara_spmm.h).
 *
 * Host plan: a from-scratch restatement of what reference src/rowpara_spmm.c:46-184
 * computes - the same index lists, counts and displacements, element for element
 * (tests/ compare every public field against a build of the reference sources).
 *
 * Execution: nothing of the reference's exec survives.  The reference packs,
 * exchanges, unpacks into a freshly malloc'd rB, copies its own rows into rB and
 * builds an MKL handle on every call (src/rowpara_spmm.c:225-413).  Here
 *   - the device CSR is created once at init, with "virtual" column ids: an id
 *     below nB (the number of B rows this rank owns) addresses the caller's B
 *     block directly, an id >= nB addresses row (id - nB) of the receive buffer.
 *     The SpMM kernel therefore consumes its own rows and the received rows in
 *     place: there is no rB, no unpack and no self copy;
 *   - pack is one gather kernel over the flat (peer, row) list;
 *   - the exchange is one grouped NCCL send/recv (byte counts = the reference's
 *     rB_scnts / rB_rcnts), or, when several ranks share one GPU (more ranks
 *     than devices - NCCL cannot do that), the same messages staged through
 *     pinned host memory and MPI, like the reference's DEV_TYPE_CUDA staging
 *     (src/mat_redist.c:362-378);
 *   - all buffers persist between execs; phases are timed with CUDA events.
 */
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mpi.h>

#include "utils.h"
#include "rowpara_spmm.h"
#include "crp_ext.h"
#include "crp_internal.h"

/* ------------------------------------------------------------------ helpers */

static void *xmalloc(size_t bytes)
{
    void *p = malloc(bytes > 0 ? bytes : 1);
    ASSERT_PRINTF(p != NULL, "Failed to allocate %zu bytes of work memory for rp_spmm\n", bytes);
    return p;
}

static void grow_dev(void **buf, size_t *cap, size_t need)
{
    if (need <= *cap) return;
    crp_cuda_free_dev(*buf);
    crp_cuda_malloc_dev(buf, need);
    *cap = need;
}

static void grow_host(void **buf, size_t *cap, size_t need)
{
    if (need <= *cap) return;
    crp_cuda_free_host(*buf);
    crp_cuda_malloc_host(buf, need);
    *cap = need;
}

/* ------------------------------------------------------------- host planning */

/*
 * Steps 1-5 of the reference init (src/rowpara_spmm.c:46-184).  Collective over comm
 * (one MPI_Alltoall and one MPI_Alltoallv, as in the reference).
 */
static void rp_build_plan(
    rp_spmm_p rp, const int A_nrow, const int *A_rowptr, const int *A_colidx, const double *A_val,
    const int *B_row_displs, const int glb_n, MPI_Comm comm, int **send_row_cnt, int **recv_row_cnt
)
{
    const int nproc = rp->nproc, me = rp->my_rank;
    const int reidx = rp->rB_reidx;
    const int nnz_base = A_rowptr[0];
    const int nnz = A_rowptr[A_nrow] - nnz_base;
    const int glb_k = B_row_displs[nproc];

    /* 1. column range of the local A, needed-row flags, compact copy of A.  Large blocks with a usable GPU: the three O(nnz)
     *    sweeps (range, flags, re-indexing) and the prefix sum over the flags run on the device (csrc/cuda/plan_build.cu) and
     *    return the re-indexed columns and the sorted list of needed rows; the flags are rebuilt from that list. */
    int *rowptr = (int *) xmalloc(sizeof(int) * ((size_t) A_nrow + 1));
    int *colidx = (int *) xmalloc(sizeof(int) * (size_t) nnz);
    double *val = (double *) xmalloc(sizeof(double) * (size_t) nnz);
    for (int i = 0; i <= A_nrow; i++) rowptr[i] = A_rowptr[i] - nnz_base;
    memcpy(val, A_val, sizeof(double) * (size_t) nnz);
    unsigned char *needed = (unsigned char *) xmalloc((size_t) glb_k);
    memset(needed, 0, (size_t) glb_k);

    int lo = INT_MAX, hi = 0, span, rB_nrow;
    int *pos_of = NULL;                 /* reidx: position in rB of global row lo + i */
    int gpu_n_needed = 0, *gpu_rows = NULL;
    if (crp_gpu_plan_enabled((long long) nnz) &&
        crp_cuda_plan_needed_rows(A_colidx, (long long) nnz, glb_k, reidx, colidx, &lo, &hi, &gpu_n_needed, &gpu_rows))
    {
        span = hi - lo + 1;
        rB_nrow = reidx ? gpu_n_needed : span;
        if (reidx) pos_of = (int *) xmalloc(sizeof(int) * (size_t) (span > 0 ? span : 1));
        for (int j = 0; j < gpu_n_needed; j++)
        {
            needed[gpu_rows[j]] = 1;
            if (reidx) pos_of[gpu_rows[j] - lo] = j;
        }
        free(gpu_rows);
    } else {
        for (int i = 0; i < nnz; i++)
        {
            const int c = A_colidx[i];
            if (c < lo) lo = c;
            if (c > hi) hi = c;
        }
        for (int i = 0; i < nnz; i++) needed[A_colidx[i]] = 1;
        /* span == hi - lo + 1 also when nnz == 0 (the reference's INT_MAX arithmetic, kept for parity) */
        span = hi - lo + 1;
        rB_nrow = span;
        if (reidx)
        {
            pos_of = (int *) xmalloc(sizeof(int) * (size_t) (span > 0 ? span : 1));
            int cnt = 0;
            for (int g = 0; g < glb_k; g++)
                if (needed[g]) pos_of[g - lo] = cnt++;
            rB_nrow = cnt;
            for (int i = 0; i < nnz; i++) colidx[i] = pos_of[A_colidx[i] - lo];
        } else {
            for (int i = 0; i < nnz; i++) colidx[i] = A_colidx[i] - lo;
        }
    }
    rp->A_rowptr = rowptr;
    rp->A_colidx = colidx;
    rp->A_val    = val;
    rp->rB_nrow  = rB_nrow;

    /* 2. rows taken from this rank's own B block */
    const int my_lo = B_row_displs[me], my_hi = B_row_displs[me + 1];
    int self_n = 0;
    for (int g = my_lo; g < my_hi; g++) self_n += needed[g];
    int *self_rows = (int *) xmalloc(sizeof(int) * (size_t) self_n);
    self_n = 0;
    for (int g = my_lo; g < my_hi; g++)
        if (needed[g])
        {
            self_rows[self_n++] = g;
            needed[g] = 0;          /* own rows are not requested from anybody */
        }
    rp->rB_self_nrow = self_n;
    rp->rB_self_src_ridxs = self_rows;
    rp->rB_self_src_offset = 0;
    rp->rB_self_dst_offset = 0;
    if (self_n > 0)
    {
        rp->rB_self_src_offset = self_rows[0] - my_lo;
        rp->rB_self_dst_offset = reidx ? pos_of[self_rows[0] - lo] : self_rows[0] - lo;
    }

    /* 3. rows requested from every other owner, grouped by owner, ascending global id */
    int *rcnts   = (int *) xmalloc(sizeof(int) * (size_t) nproc);
    int *rdispls = (int *) xmalloc(sizeof(int) * ((size_t) nproc + 1));
    int *rridxs  = (int *) xmalloc(sizeof(int) * (size_t) (rB_nrow > 0 ? rB_nrow : 1));
    int nreq = 0;
    rdispls[0] = 0;
    for (int p = 0; p < nproc; p++)
    {
        int c = 0;
        for (int g = B_row_displs[p]; g < B_row_displs[p + 1]; g++)
            if (needed[g]) { rridxs[nreq++] = g; c++; }
        rcnts[p] = c;
        rdispls[p + 1] = rdispls[p] + c;
    }
    free(needed);
    rp->rB_recv_size = (size_t) (rdispls[nproc] - rcnts[me]);

    /* 4. tell every owner which of its rows this rank wants */
    int *scnts   = (int *) xmalloc(sizeof(int) * (size_t) nproc);
    int *sdispls = (int *) xmalloc(sizeof(int) * ((size_t) nproc + 1));
    MPI_Alltoall(rcnts, 1, MPI_INT, scnts, 1, MPI_INT, comm);
    sdispls[0] = 0;
    for (int p = 0; p < nproc; p++) sdispls[p + 1] = sdispls[p] + scnts[p];
    int *sridxs = (int *) xmalloc(sizeof(int) * (size_t) sdispls[nproc]);
    MPI_Alltoallv(rridxs, rcnts, rdispls, MPI_INT, sridxs, scnts, sdispls, MPI_INT, comm);

    /* 5. global ids -> positions (receive side: in rB; send side: in the own B block); counts in elements */
    for (int i = 0; i < nreq; i++) rridxs[i] = reidx ? pos_of[rridxs[i] - lo] : rridxs[i] - lo;
    for (int i = 0; i < sdispls[nproc]; i++) sridxs[i] -= my_lo;
    /* rows per peer, kept apart: the public counts below are rows * glb_n in int arithmetic (the reference's) and may wrap */
    *send_row_cnt = (int *) xmalloc(sizeof(int) * (size_t) nproc);
    *recv_row_cnt = (int *) xmalloc(sizeof(int) * (size_t) nproc);
    for (int p = 0; p < nproc; p++) { (*send_row_cnt)[p] = scnts[p]; (*recv_row_cnt)[p] = rcnts[p]; }
    for (int p = 0; p < nproc; p++)
    {
        rcnts[p] *= glb_n;  rdispls[p] *= glb_n;
        scnts[p] *= glb_n;  sdispls[p] *= glb_n;
    }
    rdispls[nproc] *= glb_n;
    sdispls[nproc] *= glb_n;
    free(pos_of);

    rp->rB_rcnts = rcnts;   rp->rB_rdispls = rdispls;   rp->rB_rridxs = rridxs;
    rp->rB_scnts = scnts;   rp->rB_sdispls = sdispls;   rp->rB_sridxs = sridxs;
}

/* ------------------------------------------------------------- device state */

/* Default data plane with one rank per GPU: 2 = NVLink peer stores (measured on the pwtk-shaped n = 256 case:
 * 8 GPUs 0.134 vs 0.150 ms per exec with NCCL, 2 GPUs 0.271 vs 0.277; profiles/r01_bench_n*_{p2p,nccl}.json).
 * Falls back to NCCL by itself when CUDA IPC is not available. */
#ifndef CRP_DEFAULT_TRANSPORT
#define CRP_DEFAULT_TRANSPORT 2
#endif
enum { CRP_P2P_HDR = 1024 };
#define CRP_P2P_TIMEOUT_S 20.0

/*
 * Peer-memory transport: every rank exports one allocation (arrival flags + two receive halves) with
 * CUDA IPC, maps the allocations of the ranks it sends to, and learns at which row of their receive
 * buffer its rows start.  Collective over rp->comm.  Returns 0 (and leaves d->p2p = 0 on every rank)
 * if any rank could not export or map - the exchange then falls back to NCCL.
 */
static int rp_p2p_setup(rp_spmm_p rp, struct crp_rp_dev *d)
{
    const int nproc = rp->nproc, me = rp->my_rank, n = rp->glb_n;
    int ok = (nproc <= CRP_P2P_HDR / (int) sizeof(unsigned int)) ? 1 : 0;
    if (!ok && me == 0)
        WARNING_PRINTF("peer-memory transport: %d ranks exceed the %d arrival flags of the header; the B-row exchange uses NCCL\n",
                       nproc, CRP_P2P_HDR / (int) sizeof(unsigned int));
    d->p2p_half_bytes = (((size_t) d->n_recv_rows * (size_t) n * sizeof(double)) + 255) & ~(size_t) 255;
    crp_cuda_malloc_dev(&d->p2p_mem, CRP_P2P_HDR + 2 * d->p2p_half_bytes);
    crp_cuda_memset_dev(d->p2p_mem, 0, CRP_P2P_HDR);
    crp_cuda_device_sync();
    struct { unsigned char h[CRP_IPC_HANDLE_BYTES]; unsigned long long half; int ok; int pad; } mine, *all;
    memset(&mine, 0, sizeof(mine));
    mine.ok = ok && crp_cuda_ipc_get_handle(d->p2p_mem, mine.h);
    mine.half = (unsigned long long) d->p2p_half_bytes;
    all = xmalloc(sizeof(mine) * (size_t) nproc);
    MPI_Allgather(&mine, (int) sizeof(mine), MPI_BYTE, all, (int) sizeof(mine), MPI_BYTE, rp->comm);
    /* where do my rows start in each peer's receive buffer? = that peer's recv_rows[me] */
    d->peer_recv_off = (int *) xmalloc(sizeof(int) * (size_t) nproc);
    MPI_Alltoall(d->recv_rows, 1, MPI_INT, d->peer_recv_off, 1, MPI_INT, rp->comm);
    d->peer_mem = (void **) calloc((size_t) nproc, sizeof(void *));
    d->peer_half_bytes = (size_t *) calloc((size_t) nproc, sizeof(size_t));
    ok = 1;
    for (int p = 0; p < nproc; p++)
    {
        if (!all[p].ok) ok = 0;
        d->peer_half_bytes[p] = (size_t) all[p].half;
    }
    for (int p = 0; ok && p < nproc; p++)
    {
        /* neighbours = ranks this rank exchanges rows with in either direction (the relation is symmetric) */
        if (p == me || (d->send_rows[p + 1] == d->send_rows[p] && d->recv_rows[p + 1] == d->recv_rows[p])) continue;
        d->peer_mem[p] = crp_cuda_ipc_open(all[p].h);
        if (d->peer_mem[p] == NULL) ok = 0;
    }
    free(all);
    int all_ok = 0;
    MPI_Allreduce(&ok, &all_ok, 1, MPI_INT, MPI_MIN, rp->comm);
    if (!all_ok)
    {
        for (int p = 0; p < nproc; p++) crp_cuda_ipc_close(d->peer_mem[p]);
        MPI_Barrier(rp->comm);
        crp_cuda_free_dev(d->p2p_mem);
        free(d->peer_mem); free(d->peer_half_bytes); free(d->peer_recv_off);
        d->p2p_mem = NULL; d->peer_mem = NULL; d->peer_half_bytes = NULL; d->peer_recv_off = NULL;
        if (me == 0) WARNING_PRINTF("CUDA IPC peer mapping failed; the B-row exchange falls back to NCCL\n");
        return 0;
    }
    /* arrival flags: mine on every peer I send to; the ones I wait for */
    void **fp = (void **) xmalloc(sizeof(void *) * (size_t) nproc);
    int *wi = (int *) xmalloc(sizeof(int) * (size_t) nproc);
    d->n_flag = d->n_wait = 0;
    for (int p = 0; p < nproc; p++)
    {
        if (p == me) continue;
        /* flags go both ways between neighbours even if rows go one way only: a rank that has seen epoch e - 1
         * from a neighbour knows the neighbour is done reading the buffer half that epoch e will overwrite */
        if (d->peer_mem[p] == NULL) continue;
        fp[d->n_flag++] = (char *) d->peer_mem[p] + sizeof(unsigned int) * (size_t) me;
        wi[d->n_wait++] = p;
    }
    if (d->n_flag) { crp_cuda_malloc_dev(&d->d_flag_ptrs, sizeof(void *) * (size_t) d->n_flag); crp_cuda_memcpy_h2d(fp, d->d_flag_ptrs, sizeof(void *) * (size_t) d->n_flag); }
    if (d->n_wait) { crp_cuda_malloc_dev((void **) &d->d_wait_idx, sizeof(int) * (size_t) d->n_wait); crp_cuda_memcpy_h2d(wi, d->d_wait_idx, sizeof(int) * (size_t) d->n_wait); }
    free(fp);
    int *wi_copy = wi;
    crp_cuda_malloc_host((void **) &d->h_err, sizeof(int));
    *d->h_err = 0;
    crp_cuda_malloc_dev((void **) &d->d_put_counter, sizeof(unsigned int));
    crp_cuda_memset_dev(d->d_put_counter, 0, sizeof(unsigned int));
    /* wait map of the local-SpMM plans: which neighbour's flag a piece of work depends on (rows of the receive buffer are
     * grouped by sending rank, ranks ascending - the order of the wait slots) */
    {
        int *roff = (int *) xmalloc(sizeof(int) * ((size_t) d->n_wait + 1));
        for (int j = 0; j < d->n_wait; j++) roff[j] = d->recv_rows[wi_copy[j]];
        roff[d->n_wait] = d->n_recv_rows;
        crp_cuda_spmm_set_wait_map(d->plan, d->n_wait, roff);
        if (d->plan_off) crp_cuda_spmm_set_wait_map(d->plan_off, d->n_wait, roff);
        free(roff);
    }
    free(wi_copy);
    d->epoch = 0;
    d->dst_elem_size = 0;
    MPI_Barrier(rp->comm);
    return 1;
}

/* destination address of every send row, for both receive halves, for the given element size */
static void rp_p2p_tables(rp_spmm_p rp, struct crp_rp_dev *d, const int elem_size)
{
    if (d->dst_elem_size == elem_size || d->n_send_rows == 0) { d->dst_elem_size = elem_size; return; }
    const size_t row_bytes = (size_t) elem_size * (size_t) rp->glb_n;
    void **tab = (void **) xmalloc(sizeof(void *) * (size_t) d->n_send_rows);
    for (int half = 0; half < 2; half++)
    {
        for (int p = 0; p < rp->nproc; p++)
            for (int i = d->send_rows[p]; i < d->send_rows[p + 1]; i++)
                tab[i] = (char *) d->peer_mem[p] + CRP_P2P_HDR + (size_t) half * d->peer_half_bytes[p]
                       + row_bytes * ((size_t) d->peer_recv_off[p] + (size_t) (i - d->send_rows[p]));
        if (d->d_dst_rows[half] == NULL) crp_cuda_malloc_dev(&d->d_dst_rows[half], sizeof(void *) * (size_t) d->n_send_rows);
        crp_cuda_memcpy_h2d(tab, d->d_dst_rows[half], sizeof(void *) * (size_t) d->n_send_rows);
    }
    free(tab);
    d->dst_elem_size = elem_size;
}

static void rp_p2p_teardown(rp_spmm_p rp, struct crp_rp_dev *d)
{
    if (!d->p2p) return;
    for (int p = 0; p < rp->nproc; p++) crp_cuda_ipc_close(d->peer_mem[p]);
    MPI_Barrier(rp->comm);          /* nobody frees an allocation a peer still has mapped */
    crp_cuda_free_dev(d->p2p_mem);
    crp_cuda_free_dev(d->d_dst_rows[0]);
    crp_cuda_free_dev(d->d_dst_rows[1]);
    crp_cuda_free_dev(d->d_flag_ptrs);
    crp_cuda_free_dev(d->d_wait_idx);
    crp_cuda_free_dev(d->d_put_counter);
    crp_cuda_free_host(d->h_err);
    free(d->peer_mem); free(d->peer_half_bytes); free(d->peer_recv_off);
}


/*
 * Upload the local A with virtual column ids and the send list; size the
 * persistent exchange buffers.  n_send_rows / n_recv_rows are recomputed from
 * row counts kept in 64 bits (the public int counts are rows * glb_n and may
 * wrap for very wide B, as they do in the reference).
 */
static void rp_build_device_state(rp_spmm_p rp, const int *B_row_displs, MPI_Comm nccl_parent, const int *send_row_cnt, const int *recv_row_cnt)
{
    struct crp_rp_dev *d = (struct crp_rp_dev *) calloc(1, sizeof(struct crp_rp_dev));
    ASSERT_PRINTF(d != NULL, "Failed to allocate device state for rp_spmm\n");
    rp->dev = d;
    const int nproc = rp->nproc, me = rp->my_rank, n = rp->glb_n;
    const int nnz = rp->A_rowptr[rp->A_nrow];
    d->nB = B_row_displs[me + 1] - B_row_displs[me];

    /* row counts per peer, as counted by the plan (not recovered from the element counts, which may have wrapped) */
    d->send_rows = (int *) xmalloc(sizeof(int) * ((size_t) nproc + 1));
    d->recv_rows = (int *) xmalloc(sizeof(int) * ((size_t) nproc + 1));
    d->send_rows[0] = d->recv_rows[0] = 0;
    for (int p = 0; p < nproc; p++)
    {
        d->send_rows[p + 1] = d->send_rows[p] + (n > 0 ? send_row_cnt[p] : 0);
        d->recv_rows[p + 1] = d->recv_rows[p] + (n > 0 && p != rp->my_rank ? recv_row_cnt[p] : 0);
    }
    d->n_send_rows = d->send_rows[nproc];
    d->n_recv_rows = d->recv_rows[nproc];

    /* rB position -> virtual id */
    const int rB_nrow = rp->rB_nrow > 0 ? rp->rB_nrow : 0;
    int *vid = (int *) xmalloc(sizeof(int) * (size_t) (rB_nrow > 0 ? rB_nrow : 1));
    for (int i = 0; i < rB_nrow; i++) vid[i] = -1;
    for (int i = 0; i < d->n_recv_rows; i++) vid[rp->rB_rridxs[i]] = d->nB + i;
    for (int i = 0; i < rp->rB_self_nrow; i++)
    {
        const int step = rp->rB_self_src_ridxs[i] - rp->rB_self_src_ridxs[0];
        const int dst = rp->rB_self_dst_offset + (rp->rB_reidx ? i : step);
        vid[dst] = rp->rB_self_src_offset + step;
    }
    int *vcol = (int *) xmalloc(sizeof(int) * (size_t) (nnz > 0 ? nnz : 1));
    for (int i = 0; i < nnz; i++)
    {
        vcol[i] = vid[rp->A_colidx[i]];
        ASSERT_PRINTF(vcol[i] >= 0, "rp_spmm: column %d of the local A has no source row\n", rp->A_colidx[i]);
    }
    free(vid);
    /* Rows whose virtual ids are not ascending (received rows have ids above the own ones whatever their global position)
     * are put in ascending virtual order, values along: the kernels' row-group analysis wants sorted rows - without this the
     * rows at a rank boundary fell out of the grouped form (465 "rest" rows on rank 1 of 2 on the pwtk-shaped matrix).  The
     * public A_colidx / A_val keep the reference's order; only the device copy is permuted (a row's products are then added
     * in a different order: within the 1e-12 contract, bit-identical between runs). */
    double *vval = (double *) xmalloc(sizeof(double) * (size_t) (nnz > 0 ? nnz : 1));
    memcpy(vval, rp->A_val, sizeof(double) * (size_t) nnz);
    for (int i = 0; i < rp->A_nrow; i++)
    {
        const int b = rp->A_rowptr[i], e = rp->A_rowptr[i + 1];
        int sorted = 1;
        for (int p = b + 1; p < e; p++) if (vcol[p] < vcol[p - 1]) { sorted = 0; break; }
        if (sorted) continue;
        for (int p = b + 1; p < e; p++)         /* insertion sort: rows are short and nearly sorted (two sorted runs) */
        {
            const int c = vcol[p];
            const double x = vval[p];
            int q = p - 1;
            while (q >= b && vcol[q] > c) { vcol[q + 1] = vcol[q]; vval[q + 1] = vval[q]; q--; }
            vcol[q + 1] = c;  vval[q + 1] = x;
        }
    }

    /* transport: NCCL needs one device per rank; ranks that share a GPU stage the exchange through the host */
    int wsize = 1;
    MPI_Comm_size(MPI_COMM_WORLD, &wsize);
    int transport;
    GET_ENV_INT_VAR(transport, "CRP_SPMM_TRANSPORT", "transport", -1, 0, 2, 0);   /* 0 NCCL, 1 staged MPI, 2 NVLink peer stores */
    /* Ranks that share a GPU (more ranks than devices; NCCL cannot do that): the default is the host-staged transport.
     * CRP_SPMM_TRANSPORT=2 is honoured there too - CUDA IPC maps a peer's buffer on the same device just as well - but
     * kernels of different processes on ONE GPU must never spin on each other (nothing guarantees that they run at the
     * same time), so the arrival of the rows is established by a host barrier after the put kernels have completed
     * ("hostsync"); the flags are still written, and read (already satisfied) by the SpMM kernel. */
    const int shared_gpu = (crp_ranks_on_this_node(wsize) > crp_cuda_device_count());
    if (shared_gpu && transport != 2) transport = 1;
    else if (transport < 0) transport = CRP_DEFAULT_TRANSPORT;
    d->staged = (transport == 1);
    d->p2p = (transport == 2 && nproc > 1);
    d->p2p_hostsync = (d->p2p && shared_gpu) ? 1 : 0;

    /* Overlap mode (default with NCCL): the product is split by column into the part that needs only
     * this rank's own B rows - it runs while the exchange is in flight - and the part that needs
     * received rows, accumulated afterwards (C += ...).  Rows of A stay whole in each part's CSR. */
    /* Default: only when the exchange is big enough to be worth hiding.  Measured on 8 B200s: with
     * 1.7 MB received per rank (pwtk-shaped, n = 256) the exchange is latency / skew bound and the split
     * costs more than it hides (0.213 vs 0.204 ms per exec), so the threshold is a few MB. */
    int want_overlap;
    GET_ENV_INT_VAR(want_overlap, "CRP_SPMM_OVERLAP", "overlap", -1, 0, 1, 0);
    const int auto_overlap = ((size_t) d->n_recv_rows * (size_t) n * sizeof(double) >= ((size_t) 4 << 20)) ? 1 : 0;
    /* the peer-memory transport overlaps inside the SpMM kernel (it waits for a neighbour right before the first piece of
     * work that reads its rows), so the split is only made there when asked for explicitly */
    if (want_overlap < 0 && transport == 2) want_overlap = 0;
    if (want_overlap < 0 && d->staged) want_overlap = 0;
    if (want_overlap < 0) want_overlap = auto_overlap;
    d->overlap = (want_overlap && nproc > 1 && (d->n_send_rows > 0 || d->n_recv_rows > 0)) ? 1 : 0;
    if (d->overlap && d->n_recv_rows > 0)
    {
        const int m = rp->A_nrow;
        int *rp_d = (int *) xmalloc(sizeof(int) * ((size_t) m + 1)), *rp_o = (int *) xmalloc(sizeof(int) * ((size_t) m + 1));
        int *ci_d = (int *) xmalloc(sizeof(int) * (size_t) (nnz > 0 ? nnz : 1)), *ci_o = (int *) xmalloc(sizeof(int) * (size_t) (nnz > 0 ? nnz : 1));
        double *v_d = (double *) xmalloc(sizeof(double) * (size_t) (nnz > 0 ? nnz : 1)), *v_o = (double *) xmalloc(sizeof(double) * (size_t) (nnz > 0 ? nnz : 1));
        int nd = 0, no = 0;
        rp_d[0] = rp_o[0] = 0;
        for (int i = 0; i < m; i++)
        {
            for (int p = rp->A_rowptr[i]; p < rp->A_rowptr[i + 1]; p++)
            {
                if (vcol[p] < d->nB) { ci_d[nd] = vcol[p]; v_d[nd++] = vval[p]; }
                else                 { ci_o[no] = vcol[p]; v_o[no++] = vval[p]; }
            }
            rp_d[i + 1] = nd;
            rp_o[i + 1] = no;
        }
        d->plan = crp_cuda_spmm_plan_create(m, d->nB + d->n_recv_rows, d->nB, rp_d, ci_d, v_d, n);
        d->plan_off = (no > 0) ? crp_cuda_spmm_plan_create(m, d->nB + d->n_recv_rows, d->nB, rp_o, ci_o, v_o, n) : NULL;
        free(rp_d); free(rp_o); free(ci_d); free(ci_o); free(v_d); free(v_o);
    } else {
        d->plan = crp_cuda_spmm_plan_create(rp->A_nrow, d->nB + d->n_recv_rows, d->nB, rp->A_rowptr, vcol, vval, n);
        d->plan_off = NULL;
    }
    free(vcol);
    free(vval);

    if (d->n_send_rows > 0)
    {
        crp_cuda_malloc_dev((void **) &d->d_sridxs, sizeof(int) * (size_t) d->n_send_rows);
        crp_cuda_memcpy_h2d(rp->rB_sridxs, d->d_sridxs, sizeof(int) * (size_t) d->n_send_rows);
    }
    d->stream = crp_cuda_stream_create();
    d->stream2 = crp_cuda_stream_create_high_priority();
    for (int k = 0; k < CRP_RP_RING; k++)
        for (int i = 0; i < CRP_RP_NEV; i++) d->ev[k][i] = crp_cuda_event_create();

    d->nc = NULL;
    d->peer_nc_rank = NULL;
    if (d->p2p) d->p2p = rp_p2p_setup(rp, d);
    if (!d->p2p) d->p2p_hostsync = 0;
    if (!d->p2p && shared_gpu) d->staged = 1;
    /* host-buffer pipelining across ranks (rp_e2e_panel_count): possible only if every rank runs the fused peer-memory route and
     * has rows of A and of B (an empty rank could not tell host from device buffers) - agreed once, here */
    {
        int mine = (nproc == 1) ? 1 : ((d->p2p && !d->overlap && !d->p2p_hostsync && rp->A_nrow > 0 && d->nB > 0) ? 1 : 0), all = mine;
        if (nproc > 1) MPI_Allreduce(&mine, &all, 1, MPI_INT, MPI_MIN, rp->comm);
        d->e2e_multi_ok = all;
    }
    /* NCCL for the exchange: creating the communicator is collective over nccl_parent (the whole grid), the peer-memory
     * decision above was taken per grid column - so every member learns whether ANY column needs NCCL and joins the creation */
    int need_nccl = (nproc > 1 && !d->staged && !d->p2p) ? 1 : 0, any_nccl = need_nccl;
    if (!shared_gpu && nccl_parent != MPI_COMM_NULL) MPI_Allreduce(&need_nccl, &any_nccl, 1, MPI_INT, MPI_MAX, nccl_parent);
    if (any_nccl)
    {
        crp_nccl_comm *nc = crp_nccl_get(nccl_parent);
        if (need_nccl)
        {
            d->nc = nc;
            d->peer_nc_rank = crp_comm_ranks_in_parent(rp->comm, nccl_parent);
        }
    }
}

static void rp_free_device_state(rp_spmm_p rp)
{
    struct crp_rp_dev *d = (struct crp_rp_dev *) rp->dev;
    if (d == NULL) return;
    crp_cuda_stream_sync(d->stream);
    crp_cuda_stream_sync(d->stream2);
    if (d->stream_in != NULL) { crp_cuda_stream_sync(d->stream_in); crp_cuda_stream_sync(d->stream_out); }
    rp_p2p_teardown(rp, d);
    crp_cuda_spmm_plan_destroy(d->plan);
    crp_cuda_spmm_plan_destroy(d->plan_off);
    crp_cuda_free_dev(d->d_sridxs);
    crp_cuda_free_dev(d->d_sendbuf);
    crp_cuda_free_dev(d->d_recvbuf);
    crp_cuda_free_dev(d->d_Bwork);
    crp_cuda_free_dev(d->d_Cwork);
    crp_cuda_free_dev(d->d_Lwork);
    crp_cuda_free_host(d->h_sendbuf);
    crp_cuda_free_host(d->h_recvbuf);
    for (int k = 0; k < CRP_RP_RING; k++)
        for (int i = 0; i < CRP_RP_NEV; i++) crp_cuda_event_destroy(d->ev[k][i]);
    crp_cuda_stream_destroy(d->stream);
    crp_cuda_stream_destroy(d->stream2);
    if (d->stream_in != NULL)
    {
        for (int j = 0; j < CRP_E2E_MAX_PANELS; j++) { crp_cuda_event_destroy(d->ev_in[j]); crp_cuda_event_destroy(d->ev_out[j]); }
        crp_cuda_stream_destroy(d->stream_in);
        crp_cuda_stream_destroy(d->stream_out);
    }
    free(d->send_rows);
    free(d->recv_rows);
    free(d->peer_nc_rank);
    free(d);
    rp->dev = NULL;
}

/* --------------------------------------------------------------- public API */

void rp_spmm_init_on(
    const int A_srow, const int A_nrow, const int *A_rowptr, const int *A_colidx,
    const double *A_val, const int *B_row_displs, const int glb_n, MPI_Comm comm,
    MPI_Comm nccl_parent, rp_spmm_p *rp_spmm
)
{
    (void) A_srow;      /* unused in the reference as well */
    rp_spmm_p rp = (rp_spmm_p) calloc(1, sizeof(rp_spmm_s));
    ASSERT_PRINTF(rp != NULL, "Failed to allocate rp_spmm\n");
    const double t0 = get_wtime_sec();

    MPI_Comm_size(comm, &rp->nproc);
    MPI_Comm_rank(comm, &rp->my_rank);
    rp->A_nrow = A_nrow;
    rp->glb_n  = glb_n;
    rp->comm   = comm;
    int wrank;
    MPI_Comm_rank(MPI_COMM_WORLD, &wrank);
    GET_ENV_INT_VAR(rp->rB_p2p,   "RP_SPMM_P2P",   "rB_p2p",   1, 0, 1, wrank == 0);
    GET_ENV_INT_VAR(rp->rB_reidx, "RP_SPMM_REIDX", "rB_reidx", 1, 0, 1, wrank == 0);

    int *send_row_cnt = NULL, *recv_row_cnt = NULL;
    rp_build_plan(rp, A_nrow, A_rowptr, A_colidx, A_val, B_row_displs, glb_n, comm, &send_row_cnt, &recv_row_cnt);
    if (crp_device_ready()) rp_build_device_state(rp, B_row_displs, nccl_parent, send_row_cnt, recv_row_cnt);
    free(send_row_cnt);
    free(recv_row_cnt);

    rp->t_init = get_wtime_sec() - t0;
    *rp_spmm = rp;
}

void rp_spmm_init(
    const int A_srow, const int A_nrow, const int *A_rowptr, const int *A_colidx,
    const double *A_val, const int *B_row_displs, const int glb_n, MPI_Comm comm,
    rp_spmm_p *rp_spmm
)
{
    rp_spmm_init_on(A_srow, A_nrow, A_rowptr, A_colidx, A_val, B_row_displs, glb_n, comm, comm, rp_spmm);
}

void rp_spmm_free(rp_spmm_p *rp_spmm)
{
    rp_spmm_p rp = *rp_spmm;
    if (rp == NULL) return;
    rp_free_device_state(rp);
    crp_unpin_host_all();           /* registrations of caller buffers never outlive the engine that made them */
    free(rp->A_rowptr);
    free(rp->A_colidx);
    free(rp->A_val);
    free(rp->rB_self_src_ridxs);
    free(rp->rB_rcnts);
    free(rp->rB_rdispls);
    free(rp->rB_rridxs);
    free(rp->rB_scnts);
    free(rp->rB_sdispls);
    free(rp->rB_sridxs);
    free(rp);
    *rp_spmm = NULL;
}

/* A peer-memory exec whose wait for a neighbour timed out has produced garbage: never let that pass silently,
 * whatever the blocking mode (the flag is pinned host memory written by the wait kernel / the SpMM kernel). */
static void rp_check_p2p_error(struct crp_rp_dev *d)
{
    if (d != NULL && d->p2p && d->h_err != NULL && *(volatile int *) d->h_err)
    {
        fprintf(stderr, "[FATAL] rp_spmm_exec: a neighbour's B rows did not arrive within %g s (peer-memory transport)\n", CRP_P2P_TIMEOUT_S);
        fflush(stderr);
        abort();
    }
}

/* Fold the events of finished execs into the statistics.  mode 0: only those whose events have completed
 * (never stalls the host); 1: all of them (synchronises); 2: make room in the ring - wait for the oldest
 * exec only, then fold whatever else has completed. */
static void rp_collect(rp_spmm_p rp, const int mode)
{
    struct crp_rp_dev *d = (struct crp_rp_dev *) rp->dev;
    if (d == NULL) return;
    int wait = (mode != 0);
    while (d->ring_count > 0)
    {
        const int k = (d->ring_head - d->ring_count + CRP_RP_RING) % CRP_RP_RING;      /* oldest */
        void **ev = d->mark[k];
        if (wait) crp_cuda_event_sync(ev[CRP_EV_END]);
        else if (!crp_cuda_event_done(ev[CRP_EV_END])) break;
        if (mode == 2) wait = 0;
        rp->t_pack += 1e-3 * crp_cuda_event_elapsed_ms(ev[CRP_EV_B_IN],   ev[CRP_EV_PACKED]);
        rp->t_a2a  += 1e-3 * crp_cuda_event_elapsed_ms(ev[CRP_EV_PACKED], ev[CRP_EV_XCHG]);
        if (d->ring_overlap[k])     /* own-rows product (concurrent with the exchange) + received-rows product */
            rp->t_spmm += 1e-3 * (crp_cuda_event_elapsed_ms(ev[CRP_EV_B_IN], ev[CRP_EV_DIAG]) + crp_cuda_event_elapsed_ms(ev[CRP_EV_OFF0], ev[CRP_EV_SPMM]));
        else
            rp->t_spmm += 1e-3 * crp_cuda_event_elapsed_ms(ev[CRP_EV_XCHG], ev[CRP_EV_SPMM]);
        d->t_h2d   += 1e-3 * crp_cuda_event_elapsed_ms(ev[CRP_EV_START],  ev[CRP_EV_B_IN]);
        d->t_d2h   += 1e-3 * crp_cuda_event_elapsed_ms(ev[CRP_EV_SPMM],   ev[CRP_EV_END]);
        /* blocking execs: host wall clock of the call; enqueue-only execs: device time start -> end */
        if (d->ring_host_t1[k] > 0.0) rp->t_exec += d->ring_host_t1[k] - d->ring_host_t0[k];
        else rp->t_exec += 1e-3 * crp_cuda_event_elapsed_ms(ev[CRP_EV_START], ev[CRP_EV_END]);
        d->ring_count--;
        d->n_folded++;
    }
    rp_check_p2p_error(d);
}

/* the exchange of packed rows: device send buffer -> device receive buffer */
static void rp_exchange(rp_spmm_p rp, struct crp_rp_dev *d, const size_t row_bytes, void *stream)
{
    const int nproc = rp->nproc, me = rp->my_rank;
    if (nproc == 1 || (d->n_send_rows == 0 && d->n_recv_rows == 0)) return;
    if (!d->staged)
    {
        crp_nccl_group_start();
        for (int p = 0; p < nproc; p++)
        {
            if (p == me) continue;
            const int ns = d->send_rows[p + 1] - d->send_rows[p];
            const int nr = d->recv_rows[p + 1] - d->recv_rows[p];
            if (ns > 0) crp_nccl_send(d->nc, (const char *) d->d_sendbuf + row_bytes * (size_t) d->send_rows[p], row_bytes * (size_t) ns, d->peer_nc_rank[p], stream);
            if (nr > 0) crp_nccl_recv(d->nc, (char *) d->d_recvbuf + row_bytes * (size_t) d->recv_rows[p], row_bytes * (size_t) nr, d->peer_nc_rank[p], stream);
        }
        crp_nccl_group_end();
        return;
    }
    /* several ranks on one GPU: stage through pinned host memory and MPI */
    grow_host(&d->h_sendbuf, &d->h_sendbuf_bytes, row_bytes * (size_t) d->n_send_rows);
    grow_host(&d->h_recvbuf, &d->h_recvbuf_bytes, row_bytes * (size_t) d->n_recv_rows);
    if (d->n_send_rows > 0) crp_cuda_memcpy_async(d->d_sendbuf, d->h_sendbuf, row_bytes * (size_t) d->n_send_rows, stream);
    crp_cuda_stream_sync(stream);
    MPI_Request *reqs = (MPI_Request *) xmalloc(sizeof(MPI_Request) * 2 * (size_t) nproc);
    int nreq = 0;
    const size_t chunk = (size_t) 1 << 30;      /* MPI counts are ints */
    for (int p = 0; p < nproc; p++)
    {
        if (p == me) continue;
        const size_t nr = row_bytes * (size_t) (d->recv_rows[p + 1] - d->recv_rows[p]);
        ASSERT_PRINTF(nr < chunk * 2, "rp_spmm staged exchange: message of %zu bytes is too large\n", nr);
        if (nr > 0) MPI_Irecv((char *) d->h_recvbuf + row_bytes * (size_t) d->recv_rows[p], (int) nr, MPI_BYTE, p, p, rp->comm, &reqs[nreq++]);
    }
    for (int p = 0; p < nproc; p++)
    {
        if (p == me) continue;
        const size_t ns = row_bytes * (size_t) (d->send_rows[p + 1] - d->send_rows[p]);
        ASSERT_PRINTF(ns < chunk * 2, "rp_spmm staged exchange: message of %zu bytes is too large\n", ns);
        if (ns > 0) MPI_Isend((const char *) d->h_sendbuf + row_bytes * (size_t) d->send_rows[p], (int) ns, MPI_BYTE, p, me, rp->comm, &reqs[nreq++]);
    }
    MPI_Waitall(nreq, reqs, MPI_STATUSES_IGNORE);
    free(reqs);
    if (d->n_recv_rows > 0) crp_cuda_memcpy_async(d->h_recvbuf, d->d_recvbuf, row_bytes * (size_t) d->n_recv_rows, stream);
}


/*
 * Host-resident B and C (the reference's own calling convention, src/rowpara_spmm.c:188-227): column j of C depends on column j
 * of B only, so the call is cut into column panels that flow through three streams -
 *     stream_in :  H2D panel 0 | H2D panel 1 | H2D panel 2 | ...
 *     stream    :              | exchange + product 0 | exchange + product 1 | ...
 *     stream_out:                                     | D2H panel 0 | D2H panel 1 | ...
 * - and the two PCIe directions work at the same time instead of one after the other (round 1: 8.1 + 0.4 + 8.1 ms serial).
 * Every panel is a complete exchange round of the peer-memory transport (own epoch, alternating receive halves), restricted
 * to the panel's columns of the full-width receive buffer.  Returns the number of panels (0: not applicable - the caller
 * runs the serial path).  CRP_SPMM_E2E_PANELS sets the count (default 4, 1 = off).
 */
static int rp_e2e_panel_count(rp_spmm_p rp, struct crp_rp_dev *d, const int BC_layout, const int B_on_dev, const int C_on_dev, const int elem_size)
{
    static int want = -1;
    if (want < 0) GET_ENV_INT_VAR(want, "CRP_SPMM_E2E_PANELS", "e2e_panels", 4, 1, CRP_E2E_MAX_PANELS, 0);
    const int n = rp->glb_n;
    int P = want;
    while (P > 1 && n / P < 32) P--;
    /* panels start on 64-byte boundaries: multiples of 8 columns */
    int ok = (P > 1) && BC_layout == 0 && !B_on_dev && !C_on_dev && rp->A_nrow > 0 && d->nB > 0 && (n % 8 == 0);
    /* Several ranks: every rank must cut the call the same way (each panel is an exchange round).  Whether the engine CAN do it
     * (fused peer-memory route on every rank, no rank with an empty block) was agreed once at init (e2e_multi_ok); the rest of the
     * condition is the same on every rank as long as all of them pass the same kind of buffers (host or device) to a collective
     * exec - which the reference's API implies (host only) and include/crp_ext.h now states.  No communication here: a host
     * collective per exec between the start event and the launch cost 30 us of GPU idle time at 8 ranks (profiles/r02_trace_n8.txt). */
    if (rp->nproc > 1) ok = ok && d->e2e_multi_ok && ((size_t) elem_size * (size_t) n) % 16 == 0;
    return ok ? P : 0;
}

void rp_spmm_exec_any(rp_spmm_p rp, const int BC_layout, const void *B, const int ldB, void *C, const int ldC, const int elem_size)
{
    if (rp == NULL) return;
    struct crp_rp_dev *d = (struct crp_rp_dev *) rp->dev;
    if (d == NULL)
    {
        fprintf(stderr, "[FATAL] rp_spmm_exec: this engine has no device state (plan-only mode or no GPU); there is no CPU path\n");
        fflush(stderr);
        abort();
    }
    rp_collect(rp, d->ring_count == CRP_RP_RING ? 2 : 0);
    const double host_t0 = get_wtime_sec();
    void **ev = d->ev[d->ring_head];
    void **mark = d->mark[d->ring_head];
    /* an event is recorded only after a phase that did work: timing events are not free on the device */
#define CRP_MARK_ON(i, did_work, st) do { if ((did_work) || (i) == CRP_EV_START) { crp_cuda_event_record(ev[i], st); mark[i] = ev[i]; } else mark[i] = mark[(i) - 1]; } while (0)
#define CRP_MARK(i, did_work) CRP_MARK_ON(i, did_work, stream)
    const int n = rp->glb_n, m = rp->A_nrow, nB = d->nB;
    const size_t es = (size_t) elem_size;
    const size_t row_bytes = es * (size_t) n;
    void *stream = crp_opt_stream() ? crp_opt_stream() : d->stream;
    int overlap = d->overlap;
    const int B_on_dev = (nB > 0 && n > 0) ? crp_cuda_ptr_is_device(B) : 1;
    const int C_on_dev = (m > 0 && n > 0) ? crp_cuda_ptr_is_device(C) : 1;

    CRP_MARK(CRP_EV_START, 1);

    const int npanel = rp_e2e_panel_count(rp, d, BC_layout, B_on_dev, C_on_dev, elem_size);
    d->e2e_panels = npanel;
    if (npanel > 1)
    {
        if (d->stream_in == NULL)
        {
            d->stream_in = crp_cuda_stream_create();
            d->stream_out = crp_cuda_stream_create();
            for (int j = 0; j < CRP_E2E_MAX_PANELS; j++) { d->ev_in[j] = crp_cuda_event_create(); d->ev_out[j] = crp_cuda_event_create(); }
        }
        crp_pin_host_range(B, es * ((size_t) (nB - 1) * (size_t) ldB + (size_t) n));
        crp_pin_host_range(C, es * ((size_t) (m - 1) * (size_t) ldC + (size_t) n));
        grow_dev(&d->d_Bwork, &d->Bwork_bytes, row_bytes * (size_t) nB);
        grow_dev(&d->d_Cwork, &d->Cwork_bytes, row_bytes * (size_t) m);
        if (d->p2p) rp_p2p_tables(rp, d, elem_size);
        /* panel boundaries: equal panels, multiples of 8 columns (64-byte segments).  CRP_SPMM_E2E_TAPER=1 makes the last panel
         * about half as wide (what follows the last H2D - its product and its D2H - is the only part nothing hides); measured on
         * B200 it LOSES (14.1 vs 11.6 ms at n = 256: three 576-byte and one 320-byte segment per row copy worse than four of 512). */
        int pcol[CRP_E2E_MAX_PANELS + 1];
        {
            static int taper = -1;
            if (taper < 0) GET_ENV_INT_VAR(taper, "CRP_SPMM_E2E_TAPER", "e2e_taper", 0, 0, 1, 0);
            int base = taper ? (int) ((2ll * n) / (2 * npanel - 1)) : (n + npanel - 1) / npanel;
            base = taper ? base / 8 * 8 : (base + 7) / 8 * 8;
            if (base < 8) base = 8;
            pcol[0] = 0;
            for (int j = 1; j < npanel; j++) { pcol[j] = pcol[j - 1] + base; if (pcol[j] > n) pcol[j] = n; }
            pcol[npanel] = n;
        }
        crp_cuda_stream_wait_event(d->stream_in, mark[CRP_EV_START]);
        for (int np = 0; np < npanel; np++)
        {
            const int c0 = pcol[np], w = pcol[np + 1] - c0;
            const size_t off = es * (size_t) c0;
            if (w > 0) crp_cuda_memcpy2d_async((const char *) B + off, es * (size_t) ldB, (char *) d->d_Bwork + off, row_bytes, es * (size_t) w, (size_t) nB, d->stream_in);
            crp_cuda_event_record(d->ev_in[np], d->stream_in);
        }
        for (int np = 0; np < npanel; np++)
        {
            const int c0 = pcol[np], w = pcol[np + 1] - c0;
            const size_t off = es * (size_t) c0;
            crp_cuda_stream_wait_event(stream, d->ev_in[np]);
            if (np == 0) { crp_cuda_event_record(ev[CRP_EV_B_IN], stream); mark[CRP_EV_B_IN] = ev[CRP_EV_B_IN]; }
            if (w <= 0) continue;                   /* every rank skips the same (empty) panels: no exchange round is lost */
            if (d->p2p)
            {
                d->epoch++;
                const int half = (int) (d->epoch & 1u);
                const char *X1 = (const char *) d->p2p_mem + CRP_P2P_HDR + (size_t) half * d->p2p_half_bytes + off;
                crp_exchange xc;
                memset(&xc, 0, sizeof(xc));
                xc.n_send_rows = d->n_send_rows;  xc.send_ridx_d = d->d_sridxs;  xc.dst_rows_d = (void *const *) d->d_dst_rows[half];
                xc.flag_ptrs_d = (unsigned int *const *) d->d_flag_ptrs;  xc.nflag = d->n_flag;  xc.done_counter_d = d->d_put_counter;
                xc.flags_d = (const unsigned int *) d->p2p_mem;  xc.wait_idx_d = d->d_wait_idx;  xc.nwait = d->n_wait;
                xc.epoch = d->epoch;  xc.timeout_s = CRP_P2P_TIMEOUT_S;  xc.err = d->h_err;  xc.dst_off_bytes = off;
                crp_cuda_spmm_exec_exchange(d->plan, w, elem_size, 1.0, (const char *) d->d_Bwork + off, n, X1, n, 0.0, (char *) d->d_Cwork + off, n, &xc, stream);
            }
            else crp_cuda_spmm_exec(d->plan, w, elem_size, 1.0, (const char *) d->d_Bwork + off, n, NULL, 0, 0.0, (char *) d->d_Cwork + off, n, stream);
            crp_cuda_event_record(d->ev_out[np], stream);
            crp_cuda_stream_wait_event(d->stream_out, d->ev_out[np]);
            crp_cuda_memcpy2d_async((const char *) d->d_Cwork + off, row_bytes, (char *) C + off, es * (size_t) ldC, es * (size_t) w, (size_t) m, d->stream_out);
        }
        /* phase marks of a pipelined exec: H2D = start .. first panel on the device, SpMM = that .. last product done (the
         * remaining H2D hides under it), D2H = the exposed tail after the last product */
        mark[CRP_EV_PACKED] = mark[CRP_EV_XCHG] = mark[CRP_EV_DIAG] = mark[CRP_EV_OFF0] = mark[CRP_EV_B_IN];
        crp_cuda_event_record(ev[CRP_EV_SPMM], stream);   mark[CRP_EV_SPMM] = ev[CRP_EV_SPMM];
        crp_cuda_event_record(ev[CRP_EV_END], d->stream_out);  mark[CRP_EV_END] = ev[CRP_EV_END];
        crp_cuda_stream_wait_event(stream, mark[CRP_EV_END]);          /* later work on the engine's stream follows the last D2H */
        overlap = 0;
        goto exec_enqueued;
    }

    /* ---- B as a row-major device matrix Bd (leading dimension ldBd) ---- */
    const void *Bd = B;
    size_t ldBd = (size_t) ldB;
    if (nB > 0 && n > 0)
    {
        if (BC_layout == 0)
        {
            if (!B_on_dev)
            {
                crp_pin_host_range(B, es * ((size_t) (nB - 1) * (size_t) ldB + (size_t) n));
                grow_dev(&d->d_Bwork, &d->Bwork_bytes, row_bytes * (size_t) nB);
                crp_cuda_memcpy2d_async(B, es * (size_t) ldB, d->d_Bwork, row_bytes, row_bytes, (size_t) nB, stream);
                Bd = d->d_Bwork;
                ldBd = (size_t) n;
            }
        } else {
            /* column-major nB x n with leading dimension ldB == row-major n x nB */
            const void *Bcm = B;
            size_t ldcm = (size_t) ldB;
            if (!B_on_dev)
            {
                crp_pin_host_range(B, es * ((size_t) (n - 1) * (size_t) ldB + (size_t) nB));
                grow_dev(&d->d_Lwork, &d->Lwork_bytes, es * (size_t) nB * (size_t) n);
                crp_cuda_memcpy2d_async(B, es * (size_t) ldB, d->d_Lwork, es * (size_t) nB, es * (size_t) nB, (size_t) n, stream);
                Bcm = d->d_Lwork;
                ldcm = (size_t) nB;
            }
            grow_dev(&d->d_Bwork, &d->Bwork_bytes, row_bytes * (size_t) nB);
            crp_cuda_transpose(es, n, nB, Bcm, (int) ldcm, d->d_Bwork, n, stream);
            Bd = d->d_Bwork;
            ldBd = (size_t) n;
        }
    }
    CRP_MARK(CRP_EV_B_IN, Bd != B);

    /* ---- pack the rows other ranks need, exchange: on the communication stream in overlap mode ---- */
    void *cs = overlap ? d->stream2 : stream;
    if (overlap)
    {
        CRP_MARK(CRP_EV_B_IN, 1);                               /* fork point: B is in place */
        crp_cuda_stream_wait_event(cs, mark[CRP_EV_B_IN]);
    }
    const void *X1 = d->d_recvbuf;
    /* peer-memory transport without the overlap split: the SpMM kernel itself waits for the neighbours' flags */
    int kwait = 0, fused = 0;
    if (d->p2p && n > 0)
    {
        /* one launch: gather + NVLink stores into the peers' receive halves, then this rank's arrival flag on every neighbour */
        rp_p2p_tables(rp, d, elem_size);
        d->epoch++;
        const int half = (int) (d->epoch & 1u);
        X1 = (const char *) d->p2p_mem + CRP_P2P_HDR + (size_t) half * d->p2p_half_bytes;
        fused = (!overlap && !d->p2p_hostsync) ? 1 : 0;
        if (!fused)
        {
            /* one launch: gather + NVLink stores into the peers' receive halves, then this rank's arrival flag on every neighbour */
            crp_cuda_put_rows_signal(es, d->n_send_rows, n, Bd, (int) ldBd, d->d_sridxs, (void *const *) d->d_dst_rows[half],
                                     (unsigned int *const *) d->d_flag_ptrs, d->n_flag, d->epoch, d->d_put_counter, 0, cs);
            CRP_MARK_ON(CRP_EV_PACKED, d->n_send_rows > 0 || d->n_flag > 0, cs);
        } else {
            mark[CRP_EV_PACKED] = mark[CRP_EV_B_IN];        /* the SpMM kernel stores the rows itself */
        }
        if (d->p2p_hostsync)
        {
            /* ranks sharing one GPU: no kernel may spin on another process' kernel - a host barrier after the puts have
             * completed establishes the arrival; the flags are then already set when the SpMM kernel reads them */
            crp_cuda_stream_sync(cs);
            MPI_Barrier(rp->comm);
        }
        if (overlap)
        {
            crp_cuda_wait_flags((const unsigned int *) d->p2p_mem, d->d_wait_idx, d->n_wait, d->epoch, CRP_P2P_TIMEOUT_S, d->h_err, cs);
            CRP_MARK_ON(CRP_EV_XCHG, 1, cs);
        } else {
            kwait = d->n_wait;
            mark[CRP_EV_XCHG] = mark[CRP_EV_PACKED];
        }
    } else {
        if (d->n_send_rows > 0 && n > 0)
        {
            grow_dev(&d->d_sendbuf, &d->sendbuf_bytes, row_bytes * (size_t) d->n_send_rows);
            crp_cuda_gather_rows(es, d->n_send_rows, n, Bd, (int) ldBd, d->d_sridxs, d->d_sendbuf, n, cs);
        }
        if (d->n_recv_rows > 0 && n > 0) grow_dev(&d->d_recvbuf, &d->recvbuf_bytes, row_bytes * (size_t) d->n_recv_rows);
        CRP_MARK_ON(CRP_EV_PACKED, d->n_send_rows > 0 && n > 0, cs);
        if (n > 0) rp_exchange(rp, d, row_bytes, cs);
        CRP_MARK_ON(CRP_EV_XCHG, overlap || (rp->nproc > 1 && (d->n_send_rows > 0 || d->n_recv_rows > 0) && n > 0), cs);
        X1 = d->d_recvbuf;
    }

    /* ---- local product, reading own rows from Bd and remote rows from the receive buffer ---- */
    void *Cd = C;
    size_t ldCd = (size_t) ldC;
    const int C_direct = (BC_layout == 0) && C_on_dev;
    if (m > 0 && n > 0 && !C_direct)
    {
        grow_dev(&d->d_Cwork, &d->Cwork_bytes, row_bytes * (size_t) m);
        Cd = d->d_Cwork;
        ldCd = (size_t) n;
    }
    if (!overlap)
    {
        mark[CRP_EV_DIAG] = mark[CRP_EV_OFF0] = mark[CRP_EV_XCHG];
        if (fused)
        {
            /* ONE kernel: stores the rows the neighbours need into their receive halves, publishes the arrival flags, multiplies
             * what needs only own rows, waits for a neighbour's flag right before the first tile that reads its rows */
            crp_exchange xc;
            memset(&xc, 0, sizeof(xc));
            xc.n_send_rows = d->n_send_rows;  xc.send_ridx_d = d->d_sridxs;  xc.dst_rows_d = (void *const *) d->d_dst_rows[d->epoch & 1u];
            xc.flag_ptrs_d = (unsigned int *const *) d->d_flag_ptrs;  xc.nflag = d->n_flag;  xc.done_counter_d = d->d_put_counter;
            xc.flags_d = (const unsigned int *) d->p2p_mem;  xc.wait_idx_d = d->d_wait_idx;  xc.nwait = kwait;
            xc.epoch = d->epoch;  xc.timeout_s = CRP_P2P_TIMEOUT_S;  xc.err = d->h_err;
            crp_cuda_spmm_exec_exchange(d->plan, n, elem_size, 1.0, Bd, (int) ldBd, X1, n, 0.0, Cd, (int) ldCd, &xc, stream);
        }
        else if (n > 0 && (m > 0 || kwait > 0))
            crp_cuda_spmm_exec_wait(d->plan, n, elem_size, 1.0, Bd, (int) ldBd, X1, n, 0.0, Cd, (int) ldCd,
                                    (const unsigned int *) d->p2p_mem, d->d_wait_idx, kwait, d->epoch, CRP_P2P_TIMEOUT_S, d->h_err, stream);
        CRP_MARK(CRP_EV_SPMM, n > 0 && (m > 0 || kwait > 0 || fused));
    } else {
        if (m > 0 && n > 0) crp_cuda_spmm_exec(d->plan, n, elem_size, 1.0, Bd, (int) ldBd, X1, n, 0.0, Cd, (int) ldCd, stream);
        CRP_MARK(CRP_EV_DIAG, 1);
        crp_cuda_stream_wait_event(stream, mark[CRP_EV_XCHG]);  /* join: received rows are in place, sends are done */
        CRP_MARK(CRP_EV_OFF0, 1);
        if (m > 0 && n > 0 && d->plan_off != NULL)
            crp_cuda_spmm_exec(d->plan_off, n, elem_size, 1.0, Bd, (int) ldBd, X1, n, 1.0, Cd, (int) ldCd, stream);
        CRP_MARK(CRP_EV_SPMM, 1);
    }

    /* ---- C back to where the caller wants it ---- */
    if (m > 0 && n > 0 && !C_direct)
    {
        if (BC_layout == 0)
        {
            crp_pin_host_range(C, es * ((size_t) (m - 1) * (size_t) ldC + (size_t) n));
            crp_cuda_memcpy2d_async(Cd, row_bytes, C, es * (size_t) ldC, row_bytes, (size_t) m, stream);
        } else if (C_on_dev) {
            crp_cuda_transpose(es, m, n, Cd, n, C, ldC, stream);
        } else {
            crp_pin_host_range(C, es * ((size_t) (n - 1) * (size_t) ldC + (size_t) m));
            grow_dev(&d->d_Lwork, &d->Lwork_bytes, es * (size_t) m * (size_t) n);
            crp_cuda_transpose(es, m, n, Cd, n, d->d_Lwork, m, stream);
            crp_cuda_memcpy2d_async(d->d_Lwork, es * (size_t) m, C, es * (size_t) ldC, es * (size_t) m, (size_t) n, stream);
        }
    }
    CRP_MARK(CRP_EV_END, m > 0 && n > 0 && !C_direct);
exec_enqueued: ;
#undef CRP_MARK
#undef CRP_MARK_ON

    const int k = d->ring_head;
    d->ring_head = (d->ring_head + 1) % CRP_RP_RING;
    d->ring_count++;
    d->ring_host_t0[k] = host_t0;
    d->ring_overlap[k] = overlap;
    d->ring_host_t1[k] = 0.0;
    rp->n_exec++;
    if (crp_opt_blocking() || !C_on_dev || !B_on_dev)
    {
        crp_cuda_event_sync(mark[CRP_EV_END]);
        d->ring_host_t1[k] = get_wtime_sec();
        rp_collect(rp, 1);
    }
}

void rp_spmm_exec(rp_spmm_p rp_spmm, const int BC_layout, const double *B, const int ldB, double *C, const int ldC)
{
    rp_spmm_exec_any(rp_spmm, BC_layout, B, ldB, C, ldC, 8);
}

void rp_spmm_exec_f32(rp_spmm_p rp_spmm, const int BC_layout, const float *B, const int ldB, float *C, const int ldC)
{
    rp_spmm_exec_any(rp_spmm, BC_layout, B, ldB, C, ldC, 4);
}

const char *rp_spmm_kernel_name(rp_spmm_p rp)
{
    if (rp == NULL || rp->dev == NULL) return "none";
    return crp_cuda_spmm_last_kernel(((struct crp_rp_dev *) rp->dev)->plan);
}

const char *rp_spmm_transport_name(rp_spmm_p rp)
{
    if (rp == NULL || rp->dev == NULL) return "none";
    const struct crp_rp_dev *d = (const struct crp_rp_dev *) rp->dev;
    if (rp->nproc == 1) return "single";
    if (d->p2p) return d->p2p_hostsync ? (d->overlap ? "p2p-hostsync+overlap" : "p2p-hostsync") : (d->overlap ? "p2p+overlap" : "p2p");
    if (d->staged) return d->overlap ? "staged+overlap" : "staged";
    return d->overlap ? "nccl+overlap" : "nccl";
}

void rp_spmm_set_kernel(rp_spmm_p rp, const char *name)
{
    if (rp == NULL || rp->dev == NULL) return;
    crp_cuda_spmm_set_variant(((struct crp_rp_dev *) rp->dev)->plan, name);
}

void rp_spmm_plan_info(rp_spmm_p rp, long long out[12])
{
    for (int i = 0; i < 12; i++) out[i] = 0;
    if (rp != NULL && rp->dev != NULL) crp_cuda_spmm_plan_info(((struct crp_rp_dev *) rp->dev)->plan, out);
}

int rp_spmm_is_plan_only(rp_spmm_p rp) { return (rp != NULL && rp->dev == NULL) ? 1 : 0; }

void rp_spmm_sync_stats(rp_spmm_p rp)
{
    if (rp != NULL && rp->dev != NULL) rp_collect(rp, 1);
}

void rp_spmm_device_times(rp_spmm_p rp, double *t_h2d, double *t_d2h)
{
    *t_h2d = *t_d2h = 0.0;
    if (rp == NULL || rp->dev == NULL) return;
    rp_collect(rp, 1);
    *t_h2d = ((struct crp_rp_dev *) rp->dev)->t_h2d;
    *t_d2h = ((struct crp_rp_dev *) rp->dev)->t_d2h;
}

/* Same table, same row labels as the reference (src/rowpara_spmm.c:425-464). */
void rp_spmm_print_stat(rp_spmm_p rp)
{
    if (rp == NULL) return;
    rp_collect(rp, 1);
    const int n_exec = rp->n_exec;
    if (n_exec == 0) return;
    unsigned long long recv_rows = (unsigned long long) rp->rB_recv_size, recv_max = 0, recv_sum = 0;
    double raw[6] = { rp->t_init, rp->t_pack, rp->t_a2a, rp->t_unpack, rp->t_spmm, rp->t_exec };
    double tmax[6], tavg[6];
    MPI_Reduce(&recv_rows, &recv_max, 1, MPI_UNSIGNED_LONG_LONG, MPI_MAX, 0, rp->comm);
    MPI_Reduce(&recv_rows, &recv_sum, 1, MPI_UNSIGNED_LONG_LONG, MPI_SUM, 0, rp->comm);
    MPI_Reduce(raw, tmax, 6, MPI_DOUBLE, MPI_MAX, 0, rp->comm);
    MPI_Reduce(raw, tavg, 6, MPI_DOUBLE, MPI_SUM, 0, rp->comm);
    if (rp->my_rank != 0) return;
    for (int i = 1; i < 6; i++)
    {
        tmax[i] /= n_exec;
        tavg[i] /= (double) n_exec * rp->nproc;
    }
    printf("rp_spmm_init() time = %.2f s\n", tmax[0]);
    printf("Total / rank-max SpMM comm size = %zu, %zu\n", (size_t) (recv_sum * (unsigned long long) rp->glb_n), (size_t) (recv_max * (unsigned long long) rp->glb_n));
    printf("-------------------- Runtime (s) --------------------\n");
    printf("                                     avg         max\n");
    printf("Pack B matrix for redistribution  %6.3f      %6.3f\n", tavg[1], tmax[1]);
    printf("Redistribute B matrix             %6.3f      %6.3f\n", tavg[2], tmax[2]);
    printf("Unpack received B matrix data     %6.3f      %6.3f\n", tavg[3], tmax[3]);
    printf("Local SpMM                        %6.3f      %6.3f\n", tavg[4], tmax[4]);
    printf("Total rp_spmm_exec()              %6.3f      %6.3f\n", tavg[5], tmax[5]);
    printf("\n");
    fflush(stdout);
}

void rp_spmm_clear_stat(rp_spmm_p rp)
{
    if (rp == NULL) return;
    rp_collect(rp, 1);
    rp->n_exec   = 0;
    rp->t_pack   = 0.0;
    rp->t_a2a    = 0.0;
    rp->t_unpack = 0.0;
    rp->t_spmm   = 0.0;
    rp->t_exec   = 0.0;
    if (rp->dev) ((struct crp_rp_dev *) rp->dev)->t_h2d = ((struct crp_rp_dev *) rp->dev)->t_d2h = 0.0;
}
