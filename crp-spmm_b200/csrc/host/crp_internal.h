/*
 * crp_internal.h - declarations shared by the host sources of libcrpspmm.
 * Nothing here is part of the public surface.
 */
#ifndef CRP_INTERNAL_H
#define CRP_INTERNAL_H

#include <stddef.h>
#include <stdint.h>
#include <mpi.h>

#include "crp_cuda.h"
#include "rowpara_spmm.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- process-wide options (crp_common.c) ---- */
void *crp_opt_stream(void);          /* caller-provided stream or NULL               */
int   crp_opt_blocking(void);        /* 1: exec returns after completion (default)   */
int   crp_opt_plan_only(void);       /* CRP_SPMM_PLAN_ONLY=1: host plans only        */
int   crp_opt_pin_host(void);        /* CRP_SPMM_PIN_HOST (default 1): register host B / C buffers */
/* Binds this process to "its" GPU on first use (local rank % #GPUs); aborts when
 * there is no GPU unless plan-only mode was requested.  Returns 1 if a device is in use. */
int   crp_device_ready(void);

/* cache of pinned (cudaHostRegister'ed) caller buffers, so that repeated execs on the
 * same host B / C run their PCIe copies at full speed */
void  crp_pin_host_range(const void *ptr, size_t bytes);

/* ---- NCCL layer (comm/crp_nccl.c) ---- */
typedef struct crp_nccl_comm crp_nccl_comm;
/* Communicator spanning exactly the ranks of `comm` (cached by membership; collective over comm). */
crp_nccl_comm *crp_nccl_get(MPI_Comm comm);
int   crp_nccl_rank(const crp_nccl_comm *nc);
int   crp_nccl_size(const crp_nccl_comm *nc);
void  crp_nccl_group_start(void);
void  crp_nccl_group_end(void);
void  crp_nccl_send(crp_nccl_comm *nc, const void *buf, size_t bytes, int peer, void *stream);
void  crp_nccl_recv(crp_nccl_comm *nc, void *buf, size_t bytes, int peer, void *stream);
void  crp_nccl_shutdown(void);

/* rank of every member of `sub` inside `parent` (malloc'd, size of sub); collective over sub */
int  *crp_comm_ranks_in_parent(MPI_Comm sub, MPI_Comm parent);

/* ---- device-side state of a row-parallel engine (rowpara_spmm.c) ---- */
enum { CRP_E2E_MAX_PANELS = 16 };
enum { CRP_EV_START = 0, CRP_EV_B_IN, CRP_EV_PACKED, CRP_EV_XCHG, CRP_EV_DIAG, CRP_EV_OFF0, CRP_EV_SPMM, CRP_EV_END, CRP_RP_NEV };
enum { CRP_RP_RING = 8 };

struct crp_rp_dev
{
    crp_spmm_plan *plan;        /* device CSR with "virtual" column ids (see rp_build_device_state);
                                 * in overlap mode only the entries that reference this rank's own B rows */
    crp_spmm_plan *plan_off;    /* overlap mode: the entries that reference received rows (NULL if none) */
    int     overlap;            /* 1: pack + exchange run on stream2 concurrently with the own-rows product */
    void    *stream2;           /* communication stream of the overlap mode                          */
    int     ring_overlap[CRP_RP_RING];
    int     nB;                 /* B rows owned by this rank                                         */
    int     n_send_rows;        /* rows packed per exec                                              */
    int     n_recv_rows;        /* remote rows received per exec                                     */
    int     *send_rows;         /* nproc + 1, row offsets per destination in the send buffer         */
    int     *recv_rows;         /* nproc + 1, row offsets per source in the receive buffer           */
    int     *d_sridxs;          /* device copy of rB_sridxs                                          */
    void    *d_sendbuf;  size_t sendbuf_bytes;
    void    *d_recvbuf;  size_t recvbuf_bytes;
    void    *d_Bwork;    size_t Bwork_bytes;   /* row-major device copy of B when the caller's is host / column-major */
    void    *d_Cwork;    size_t Cwork_bytes;
    void    *d_Lwork;    size_t Lwork_bytes;   /* column-major staging for BC_layout = 1             */
    void    *h_sendbuf;  size_t h_sendbuf_bytes;   /* pinned, staged-MPI transport only              */
    void    *h_recvbuf;  size_t h_recvbuf_bytes;
    void    *stream;            /* own non-blocking stream                                           */
    void    *ev[CRP_RP_RING][CRP_RP_NEV];  /* ring of event sets: stats of an exec are folded in later,     */
    double  ring_host_t0[CRP_RP_RING];     /* when its events have completed, without stalling the host     */
    double  ring_host_t1[CRP_RP_RING];
    void    *mark[CRP_RP_RING][CRP_RP_NEV]; /* event that closes each phase (a phase without work reuses the previous one) */
    int     ring_head, ring_count;
    long long n_folded;         /* execs whose event times have been folded into the statistics      */
    double  t_h2d, t_d2h;       /* staging of host B / C (seconds, device time)                      */
    int     staged;             /* 1: exchange through pinned host memory + MPI (ranks share a GPU)  */
    /* peer-memory transport (CRP_SPMM_TRANSPORT=2): rows are stored straight into the peers' receive buffers */
    int     p2p;
    void    *p2p_mem;           /* own IPC-exported allocation: [flags, 1024 B][receive half 0][receive half 1] */
    size_t  p2p_half_bytes;
    void    **peer_mem;         /* nproc: mapped allocations of the peers this rank sends to (else NULL)   */
    size_t  *peer_half_bytes;   /* nproc                                                                    */
    int     *peer_recv_off;     /* nproc: row offset of this rank's rows inside peer p's receive buffer     */
    void    *d_dst_rows[2];     /* device tables (per buffer half): destination address of every send row  */
    int     dst_elem_size;      /* element size the tables were built for                                   */
    void    *d_flag_ptrs;  int n_flag;      /* addresses of this rank's arrival flag on the peers it sends to */
    int     *d_wait_idx;   int n_wait;      /* ranks whose arrival flag this rank waits for                   */
    int     *h_err;             /* pinned: set by the wait kernel / the SpMM kernel on timeout              */
    unsigned int *d_put_counter; /* device word: blocks of the put kernel that have finished (reset by the kernel) */
    int     p2p_hostsync;       /* ranks share a GPU: arrival is established by a host barrier, kernels never spin on each other */
    unsigned int epoch;
    crp_nccl_comm *nc;          /* NCCL communicator used for the B-row exchange                     */
    int     *peer_nc_rank;      /* nproc: rank of each member of rp->comm inside nc                  */
    /* host B / C: column panels pipelined over three streams (H2D | exchange + product | D2H)       */
    void    *stream_in, *stream_out;
    void    *ev_in[CRP_E2E_MAX_PANELS], *ev_out[CRP_E2E_MAX_PANELS];
    int     e2e_panels;         /* panels of the last pipelined exec (0: the last exec was not pipelined) */
    int     e2e_multi_ok;       /* agreed at init: every rank of comm can cut a host-buffer exec into panel rounds */
};

/* rp_spmm_init with an explicit NCCL parent: the exchange runs on the NCCL
 * communicator of `nccl_parent` (collective over it must be possible for all its
 * members at this point) instead of one created for `comm`. */
void rp_spmm_init_on(
    const int A_srow, const int A_nrow, const int *A_rowptr, const int *A_colidx,
    const double *A_val, const int *B_row_displs, const int glb_n, MPI_Comm comm,
    MPI_Comm nccl_parent, rp_spmm_p *rp_spmm
);

int crp_gpu_plan_enabled(const long long nnz);
int crp_ranks_on_this_node(const int world_size);

void rp_spmm_exec_any(rp_spmm_p rp, const int BC_layout, const void *B, const int ldB, void *C, const int ldC, const int elem_size);

#ifdef __cplusplus
}
#endif

#endif
