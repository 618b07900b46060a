/*
 * minimpi.c - single-node MPI subset over unix-domain stream sockets.
 *
 * Design: one process per rank, a full mesh of non-blocking stream sockets, a
 * single-threaded progress engine (poll) that is driven from inside every
 * blocking call.  All sends are eager; a message that arrives before its
 * receive is posted is buffered ("unexpected" queue).  Matching is FIFO per
 * (context, source, tag) as MPI requires.  Collectives are linear algorithms
 * over point-to-point on a separate context, which is plenty for <= 64 ranks
 * on one box and keeps reductions deterministic (rank order).
 *
 * See mpi.h for why this exists and what it covers.
 */
#define _GNU_SOURCE
#include <errno.h>
#include <fcntl.h>
#include <poll.h>
#include <signal.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <sys/un.h>
#include <time.h>
#include <unistd.h>

#include "mpi.h"

/* ------------------------------------------------------------------ types */

typedef struct
{
    int32_t  ctx;
    int32_t  tag;
    uint64_t nbytes;
} wire_hdr;

enum { RQ_SEND = 0, RQ_RECV = 1, RQ_DONE = 2 };

struct minimpi_request
{
    int      kind;
    int      done;
    int      ctx, tag, peer;      /* peer: world rank, or MPI_ANY_SOURCE */
    char     *buf;
    size_t   cap;                 /* recv: capacity; send: payload size  */
    size_t   nbytes;              /* recv: bytes actually delivered      */
    int      src_world, src_tag;  /* recv: envelope of the matched msg   */
    MPI_Comm comm;
    wire_hdr hdr;                 /* send: header being written          */
    size_t   off;                 /* send: bytes of hdr+payload written  */
    struct minimpi_request *next;
};
typedef struct minimpi_request req_t;

typedef struct umsg
{
    int    ctx, tag, src;
    size_t nbytes;
    char   *data;
    int    complete;
    req_t  *claimed;
    struct umsg *next;
} umsg_t;

typedef struct
{
    int      fd;
    wire_hdr in_hdr;
    size_t   in_hdr_got;
    int      in_body;             /* 1 while reading a payload */
    char     *in_dst;
    size_t   in_need, in_got;
    req_t    *in_req;
    umsg_t   *in_umsg;
    req_t    *out_head, *out_tail;
} peer_t;

typedef struct
{
    int used;
    int ctx;
    int size, rank;
    int *wr;                      /* world rank of each member */
    int indeg, outdeg;
    int *srcs, *dsts;
} comm_t;

/* ---------------------------------------------------------------- globals */

static int     g_inited = 0, g_finalized = 0;
static int     g_rank = 0, g_size = 1;
static peer_t  *g_peers = NULL;
static comm_t  *g_comms = NULL;
static int     g_ncomm = 0, g_capcomm = 0;
static int     g_next_ctx = 2;
static req_t   *g_posted_head = NULL, *g_posted_tail = NULL;
static umsg_t  *g_unexp_head = NULL, *g_unexp_tail = NULL;
static char    g_dir[80];
static char    g_mypath[104];
static int     g_listen_fd = -1;

static void fatal(const char *fmt, ...)
{
    va_list ap;
    fprintf(stderr, "[minimpi rank %d] FATAL: ", g_rank);
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fprintf(stderr, "\n");
    fflush(stderr);
    if (g_mypath[0]) unlink(g_mypath);
    _exit(86);
}

static size_t dt_size(MPI_Datatype dt) { return (size_t) (dt & 0xff); }
static int    dt_kind(MPI_Datatype dt) { return dt >> 8; }

static comm_t *get_comm(MPI_Comm c)
{
    if (c < 0 || c >= g_ncomm || !g_comms[c].used) fatal("invalid communicator handle %d", c);
    return &g_comms[c];
}

static MPI_Comm new_comm(int ctx, int size, int rank, const int *wr)
{
    if (g_ncomm == g_capcomm)
    {
        g_capcomm = g_capcomm ? 2 * g_capcomm : 16;
        g_comms = (comm_t *) realloc(g_comms, sizeof(comm_t) * g_capcomm);
    }
    comm_t *c = &g_comms[g_ncomm];
    memset(c, 0, sizeof(*c));
    c->used = 1;
    c->ctx  = ctx;
    c->size = size;
    c->rank = rank;
    c->wr   = (int *) malloc(sizeof(int) * (size > 0 ? size : 1));
    memcpy(c->wr, wr, sizeof(int) * size);
    return g_ncomm++;
}

/* --------------------------------------------------------------- matching */

static int env_match(int want_ctx, int want_src, int want_tag, int ctx, int src, int tag)
{
    if (want_ctx != ctx) return 0;
    if (want_src != MPI_ANY_SOURCE && want_src != src) return 0;
    if (want_tag != MPI_ANY_TAG && want_tag != tag) return 0;
    return 1;
}

static req_t *take_posted(int ctx, int src, int tag)
{
    req_t *prev = NULL;
    for (req_t *r = g_posted_head; r != NULL; prev = r, r = r->next)
    {
        if (!env_match(r->ctx, r->peer, r->tag, ctx, src, tag)) continue;
        if (prev) prev->next = r->next; else g_posted_head = r->next;
        if (g_posted_tail == r) g_posted_tail = prev;
        r->next = NULL;
        return r;
    }
    return NULL;
}

static void unexp_remove(umsg_t *u)
{
    umsg_t *prev = NULL;
    for (umsg_t *x = g_unexp_head; x != NULL; prev = x, x = x->next)
    {
        if (x != u) continue;
        if (prev) prev->next = x->next; else g_unexp_head = x->next;
        if (g_unexp_tail == x) g_unexp_tail = prev;
        break;
    }
    free(u->data);
    free(u);
}

static void recv_finish(req_t *r, int src, int tag, size_t nbytes)
{
    r->src_world = src;
    r->src_tag   = tag;
    r->nbytes    = nbytes;
    r->done      = 1;
}

static void deliver_umsg(umsg_t *u, req_t *r)
{
    if (u->nbytes > r->cap) fatal("message truncated: %zu bytes into a %zu byte receive (tag %d)", u->nbytes, r->cap, u->tag);
    if (u->nbytes) memcpy(r->buf, u->data, u->nbytes);
    recv_finish(r, u->src, u->tag, u->nbytes);
    unexp_remove(u);
}

static umsg_t *unexp_new(int ctx, int src, int tag, size_t nbytes)
{
    umsg_t *u = (umsg_t *) calloc(1, sizeof(umsg_t));
    u->ctx = ctx; u->src = src; u->tag = tag; u->nbytes = nbytes;
    u->data = (char *) malloc(nbytes ? nbytes : 1);
    if (u->data == NULL) fatal("out of memory buffering a %zu byte unexpected message", nbytes);
    if (g_unexp_tail) g_unexp_tail->next = u; else g_unexp_head = u;
    g_unexp_tail = u;
    return u;
}

/* --------------------------------------------------------------- progress */

static void msg_complete(peer_t *p, int src)
{
    if (p->in_req) recv_finish(p->in_req, src, p->in_hdr.tag, (size_t) p->in_hdr.nbytes);
    if (p->in_umsg)
    {
        p->in_umsg->complete = 1;
        if (p->in_umsg->claimed) deliver_umsg(p->in_umsg, p->in_umsg->claimed);
    }
    p->in_req = NULL; p->in_umsg = NULL; p->in_body = 0;
    p->in_hdr_got = 0; p->in_got = 0; p->in_need = 0; p->in_dst = NULL;
}

static void do_read(int src)
{
    peer_t *p = &g_peers[src];
    for (;;)
    {
        if (!p->in_body)
        {
            ssize_t n = recv(p->fd, (char *) &p->in_hdr + p->in_hdr_got, sizeof(wire_hdr) - p->in_hdr_got, 0);
            if (n == 0)
            {
                /* orderly EOF between messages: the peer finalized (or died); only an error if someone still waits on it */
                if (p->in_hdr_got != 0) fatal("rank %d closed its connection mid-header", src);
                close(p->fd);
                p->fd = -1;
                return;
            }
            if (n < 0)
            {
                if (errno == EAGAIN || errno == EWOULDBLOCK) return;
                if (errno == EINTR) continue;
                fatal("recv from rank %d failed: %s", src, strerror(errno));
            }
            p->in_hdr_got += (size_t) n;
            if (p->in_hdr_got < sizeof(wire_hdr)) continue;
            size_t nbytes = (size_t) p->in_hdr.nbytes;
            req_t *r = take_posted(p->in_hdr.ctx, src, p->in_hdr.tag);
            if (r != NULL)
            {
                if (nbytes > r->cap) fatal("message truncated: %zu bytes into a %zu byte receive (src %d tag %d)", nbytes, r->cap, src, p->in_hdr.tag);
                p->in_req = r;
                p->in_dst = r->buf;
            } else {
                p->in_umsg = unexp_new(p->in_hdr.ctx, src, p->in_hdr.tag, nbytes);
                p->in_dst  = p->in_umsg->data;
            }
            p->in_need = nbytes;
            p->in_got  = 0;
            p->in_body = 1;
            if (nbytes == 0) msg_complete(p, src);
        } else {
            ssize_t n = recv(p->fd, p->in_dst + p->in_got, p->in_need - p->in_got, 0);
            if (n == 0) fatal("rank %d closed its connection mid-message", src);
            if (n < 0)
            {
                if (errno == EAGAIN || errno == EWOULDBLOCK) return;
                if (errno == EINTR) continue;
                fatal("recv from rank %d failed: %s", src, strerror(errno));
            }
            p->in_got += (size_t) n;
            if (p->in_got == p->in_need) msg_complete(p, src);
        }
    }
}

static void do_write(int dst)
{
    peer_t *p = &g_peers[dst];
    while (p->out_head != NULL)
    {
        req_t *r = p->out_head;
        const char *src;
        size_t len;
        if (r->off < sizeof(wire_hdr))
        {
            src = (const char *) &r->hdr + r->off;
            len = sizeof(wire_hdr) - r->off;
        } else {
            src = r->buf + (r->off - sizeof(wire_hdr));
            len = r->cap - (r->off - sizeof(wire_hdr));
        }
        if (len > 0)
        {
            ssize_t n = send(p->fd, src, len, MSG_NOSIGNAL);
            if (n < 0)
            {
                if (errno == EAGAIN || errno == EWOULDBLOCK) return;
                if (errno == EINTR) continue;
                fatal("send to rank %d failed: %s", dst, strerror(errno));
            }
            r->off += (size_t) n;
        }
        if (r->off == sizeof(wire_hdr) + r->cap)
        {
            p->out_head = r->next;
            if (p->out_head == NULL) p->out_tail = NULL;
            r->next = NULL;
            r->done = 1;
        }
    }
}

static void progress(int timeout_ms)
{
    static struct pollfd *pfds = NULL;
    static int *pidx = NULL;
    if (g_size == 1) return;
    if (pfds == NULL)
    {
        pfds = (struct pollfd *) malloc(sizeof(struct pollfd) * g_size);
        pidx = (int *) malloc(sizeof(int) * g_size);
    }
    int n = 0;
    for (int i = 0; i < g_size; i++)
    {
        if (g_peers[i].fd < 0) continue;
        pfds[n].fd = g_peers[i].fd;
        pfds[n].events = POLLIN | (g_peers[i].out_head ? POLLOUT : 0);
        pfds[n].revents = 0;
        pidx[n] = i;
        n++;
    }
    int rc = poll(pfds, (nfds_t) n, timeout_ms);
    if (rc < 0)
    {
        if (errno == EINTR) return;
        fatal("poll failed: %s", strerror(errno));
    }
    for (int j = 0; j < n; j++)
    {
        if (pfds[j].revents & POLLIN) do_read(pidx[j]);
        if (pfds[j].revents & POLLOUT) do_write(pidx[j]);
        if ((pfds[j].revents & (POLLERR | POLLHUP | POLLNVAL)) && !(pfds[j].revents & POLLIN))
            fatal("connection to rank %d lost", pidx[j]);
    }
}

static void wait_done(req_t *r)
{
    while (!r->done)
    {
        if (r->peer >= 0 && r->peer != g_rank && g_peers[r->peer].fd < 0)
            fatal("waiting on rank %d, which has already exited", r->peer);
        progress(1000);
    }
}

/* ------------------------------------------------------- low-level p2p API */

static req_t *post_send(const void *buf, size_t nbytes, int dst_world, int ctx, int tag)
{
    req_t *r = (req_t *) calloc(1, sizeof(req_t));
    r->kind = RQ_SEND;
    r->ctx = ctx; r->tag = tag; r->peer = dst_world;
    r->buf = (char *) buf;
    r->cap = nbytes;
    if (dst_world == g_rank)
    {
        req_t *pr = take_posted(ctx, g_rank, tag);
        if (pr != NULL)
        {
            if (nbytes > pr->cap) fatal("self message truncated: %zu > %zu", nbytes, pr->cap);
            if (nbytes) memcpy(pr->buf, buf, nbytes);
            recv_finish(pr, g_rank, tag, nbytes);
        } else {
            umsg_t *u = unexp_new(ctx, g_rank, tag, nbytes);
            if (nbytes) memcpy(u->data, buf, nbytes);
            u->complete = 1;
        }
        r->done = 1;
        return r;
    }
    if (dst_world < 0 || dst_world >= g_size) fatal("send to invalid world rank %d", dst_world);
    r->hdr.ctx = ctx; r->hdr.tag = tag; r->hdr.nbytes = (uint64_t) nbytes;
    peer_t *p = &g_peers[dst_world];
    if (p->out_tail) p->out_tail->next = r; else p->out_head = r;
    p->out_tail = r;
    do_write(dst_world);
    return r;
}

static req_t *post_recv(void *buf, size_t cap, int src_world, int ctx, int tag, MPI_Comm comm)
{
    req_t *r = (req_t *) calloc(1, sizeof(req_t));
    r->kind = RQ_RECV;
    r->ctx = ctx; r->tag = tag; r->peer = src_world;
    r->buf = (char *) buf;
    r->cap = cap;
    r->comm = comm;
    for (umsg_t *u = g_unexp_head; u != NULL; u = u->next)
    {
        if (u->claimed != NULL) continue;
        if (!env_match(ctx, src_world, tag, u->ctx, u->src, u->tag)) continue;
        if (u->complete) deliver_umsg(u, r);
        else u->claimed = r;
        return r;
    }
    if (g_posted_tail) g_posted_tail->next = r; else g_posted_head = r;
    g_posted_tail = r;
    return r;
}

static void finish_req(req_t *r)
{
    wait_done(r);
    free(r);
}

static void xsend(const void *buf, size_t nbytes, int dst_world, int ctx, int tag)
{
    finish_req(post_send(buf, nbytes, dst_world, ctx, tag));
}

static void xrecv(void *buf, size_t nbytes, int src_world, int ctx, int tag)
{
    finish_req(post_recv(buf, nbytes, src_world, ctx, tag, MPI_COMM_WORLD));
}

#define P2P_CTX(c)  ((c)->ctx * 2)
#define COLL_CTX(c) ((c)->ctx * 2 + 1)

enum { T_BARRIER = 1, T_BCAST, T_GATHER, T_SCATTER, T_ALLGATHER, T_ALLTOALL, T_REDUCE, T_NEIGHBOR, T_CTX };

/* ------------------------------------------------------------- rendezvous */

static void set_nonblock(int fd)
{
    int fl = fcntl(fd, F_GETFL, 0);
    fcntl(fd, F_SETFL, fl | O_NONBLOCK);
    int sz = 4 << 20;
    setsockopt(fd, SOL_SOCKET, SO_SNDBUF, &sz, sizeof(sz));
    setsockopt(fd, SOL_SOCKET, SO_RCVBUF, &sz, sizeof(sz));
}

static void full_write(int fd, const void *buf, size_t n)
{
    const char *p = (const char *) buf;
    while (n > 0)
    {
        ssize_t w = send(fd, p, n, MSG_NOSIGNAL);
        if (w < 0) { if (errno == EINTR) continue; fatal("handshake write failed: %s", strerror(errno)); }
        p += w; n -= (size_t) w;
    }
}

static void full_read(int fd, void *buf, size_t n)
{
    char *p = (char *) buf;
    while (n > 0)
    {
        ssize_t r = recv(fd, p, n, 0);
        if (r == 0) fatal("handshake: peer closed");
        if (r < 0) { if (errno == EINTR) continue; fatal("handshake read failed: %s", strerror(errno)); }
        p += r; n -= (size_t) r;
    }
}

static const char *first_env(const char *a, const char *b, const char *c)
{
    const char *v;
    if (a && (v = getenv(a)) != NULL && v[0]) return v;
    if (b && (v = getenv(b)) != NULL && v[0]) return v;
    if (c && (v = getenv(c)) != NULL && v[0]) return v;
    return NULL;
}

static void setup_mesh(void)
{
    const char *dir = getenv("MINIMPI_DIR");
    if (dir != NULL && dir[0])
    {
        snprintf(g_dir, sizeof(g_dir), "%s", dir);
    } else {
        const char *port = getenv("MASTER_PORT");
        const char *run  = getenv("TORCHELASTIC_RUN_ID");
        snprintf(g_dir, sizeof(g_dir), "/tmp/minimpi-%d-%s-%.24s", (int) getuid(), port ? port : "0", run ? run : "x");
    }
    mkdir(g_dir, 0700);
    snprintf(g_mypath, sizeof(g_mypath), "%.80s/r%d.sock", g_dir, g_rank);
    unlink(g_mypath);

    g_listen_fd = socket(AF_UNIX, SOCK_STREAM, 0);
    if (g_listen_fd < 0) fatal("socket: %s", strerror(errno));
    struct sockaddr_un addr;
    memset(&addr, 0, sizeof(addr));
    addr.sun_family = AF_UNIX;
    snprintf(addr.sun_path, sizeof(addr.sun_path), "%.104s", g_mypath);
    if (bind(g_listen_fd, (struct sockaddr *) &addr, sizeof(addr)) < 0) fatal("bind %s: %s", g_mypath, strerror(errno));
    if (listen(g_listen_fd, g_size + 8) < 0) fatal("listen: %s", strerror(errno));

    double deadline = MPI_Wtime() + 600.0;
    for (int j = 0; j < g_rank; j++)
    {
        struct sockaddr_un pa;
        memset(&pa, 0, sizeof(pa));
        pa.sun_family = AF_UNIX;
        snprintf(pa.sun_path, sizeof(pa.sun_path), "%.80s/r%d.sock", g_dir, j);
        int fd = -1;
        for (;;)
        {
            fd = socket(AF_UNIX, SOCK_STREAM, 0);
            if (fd < 0) fatal("socket: %s", strerror(errno));
            if (connect(fd, (struct sockaddr *) &pa, sizeof(pa)) == 0) break;
            close(fd);
            if (errno != ENOENT && errno != ECONNREFUSED && errno != EAGAIN && errno != EINTR)
                fatal("connect to rank %d: %s", j, strerror(errno));
            if (MPI_Wtime() > deadline) fatal("timed out connecting to rank %d at %s", j, pa.sun_path);
            usleep(5000);
        }
        int32_t hello[2] = { g_rank, g_size };
        full_write(fd, hello, sizeof(hello));
        g_peers[j].fd = fd;
    }
    for (int cnt = g_rank + 1; cnt < g_size; cnt++)
    {
        int fd = accept(g_listen_fd, NULL, NULL);
        if (fd < 0) { if (errno == EINTR) { cnt--; continue; } fatal("accept: %s", strerror(errno)); }
        int32_t hello[2];
        full_read(fd, hello, sizeof(hello));
        if (hello[1] != g_size || hello[0] <= g_rank || hello[0] >= g_size || g_peers[hello[0]].fd >= 0)
            fatal("handshake mismatch: got rank %d size %d", hello[0], hello[1]);
        g_peers[hello[0]].fd = fd;
    }
    for (int i = 0; i < g_size; i++) if (g_peers[i].fd >= 0) set_nonblock(g_peers[i].fd);
}

/* ------------------------------------------------------------ environment */

int MPI_Init(int *argc, char ***argv)
{
    (void) argc; (void) argv;
    if (g_inited) return MPI_SUCCESS;
    signal(SIGPIPE, SIG_IGN);
    const char *r = first_env("MINIMPI_RANK", "RANK", "OMPI_COMM_WORLD_RANK");
    const char *s = first_env("MINIMPI_SIZE", "WORLD_SIZE", "OMPI_COMM_WORLD_SIZE");
    g_rank = r ? atoi(r) : 0;
    g_size = s ? atoi(s) : 1;
    if (g_size < 1 || g_rank < 0 || g_rank >= g_size) fatal("bad rank/size %d/%d", g_rank, g_size);
    g_peers = (peer_t *) calloc((size_t) g_size, sizeof(peer_t));
    for (int i = 0; i < g_size; i++) g_peers[i].fd = -1;
    if (g_size > 1) setup_mesh();

    int *wr = (int *) malloc(sizeof(int) * g_size);
    for (int i = 0; i < g_size; i++) wr[i] = i;
    new_comm(0, g_size, g_rank, wr);   /* MPI_COMM_WORLD = 0 */
    new_comm(1, 1, 0, &g_rank);        /* MPI_COMM_SELF  = 1 */
    free(wr);
    g_next_ctx = 2;
    g_inited = 1;
    MPI_Barrier(MPI_COMM_WORLD);
    return MPI_SUCCESS;
}

int MPI_Finalize(void)
{
    if (!g_inited || g_finalized) return MPI_SUCCESS;
    MPI_Barrier(MPI_COMM_WORLD);
    /* drain our outgoing queues so peers never see a truncated stream */
    for (int i = 0; i < g_size; i++) while (g_peers[i].out_head) progress(100);
    for (int i = 0; i < g_size; i++) if (g_peers[i].fd >= 0) { shutdown(g_peers[i].fd, SHUT_WR); }
    if (g_listen_fd >= 0) close(g_listen_fd);
    if (g_mypath[0]) unlink(g_mypath);
    if (g_size > 1) rmdir(g_dir);
    g_finalized = 1;
    return MPI_SUCCESS;
}

int MPI_Initialized(int *flag) { *flag = g_inited; return MPI_SUCCESS; }
int MPI_Finalized(int *flag) { *flag = g_finalized; return MPI_SUCCESS; }

int MPI_Abort(MPI_Comm comm, int errorcode)
{
    (void) comm;
    fprintf(stderr, "[minimpi rank %d] MPI_Abort(%d)\n", g_rank, errorcode);
    if (g_mypath[0]) unlink(g_mypath);
    _exit(errorcode ? errorcode : 1);
    return MPI_SUCCESS;
}

double MPI_Wtime(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

int MPI_Get_processor_name(char *name, int *resultlen)
{
    if (gethostname(name, MPI_MAX_PROCESSOR_NAME) != 0) snprintf(name, MPI_MAX_PROCESSOR_NAME, "localhost");
    name[MPI_MAX_PROCESSOR_NAME - 1] = 0;
    *resultlen = (int) strlen(name);
    return MPI_SUCCESS;
}

int minimpi_local_rank(void)
{
    const char *v = first_env("MINIMPI_LOCAL_RANK", "LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK");
    if (v) return atoi(v);
    return g_rank;
}

int minimpi_comm_world_rank(MPI_Comm comm, int rank_in_comm)
{
    comm_t *c = get_comm(comm);
    if (rank_in_comm < 0 || rank_in_comm >= c->size) return -1;
    return c->wr[rank_in_comm];
}

/* ----------------------------------------------------------- communicators */

int MPI_Comm_size(MPI_Comm comm, int *size) { *size = get_comm(comm)->size; return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm comm, int *rank) { *rank = get_comm(comm)->rank; return MPI_SUCCESS; }

/* agree on a fresh context id among the members of c */
static int agree_ctx(comm_t *c)
{
    int mine = g_next_ctx, best = mine;
    int cctx = COLL_CTX(c);
    if (c->rank == 0)
    {
        for (int i = 1; i < c->size; i++)
        {
            int v;
            xrecv(&v, sizeof(int), c->wr[i], cctx, T_CTX);
            if (v > best) best = v;
        }
        for (int i = 1; i < c->size; i++) xsend(&best, sizeof(int), c->wr[i], cctx, T_CTX);
    } else {
        xsend(&mine, sizeof(int), c->wr[0], cctx, T_CTX);
        xrecv(&best, sizeof(int), c->wr[0], cctx, T_CTX);
    }
    g_next_ctx = best + 1;
    return best;
}

int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm *newcomm)
{
    comm_t *c = get_comm(comm);
    int n = c->size, me = c->rank;
    int mine[2] = { color, key };
    int *all = (int *) malloc(sizeof(int) * 2 * n);
    MPI_Allgather(mine, 2, MPI_INT, all, 2, MPI_INT, comm);
    c = get_comm(comm);
    int ctx = agree_ctx(c);
    if (color == MPI_UNDEFINED)
    {
        *newcomm = MPI_COMM_NULL;
        free(all);
        return MPI_SUCCESS;
    }
    int *members = (int *) malloc(sizeof(int) * n);
    int cnt = 0;
    for (int i = 0; i < n; i++) if (all[2 * i] == color) members[cnt++] = i;
    /* stable insertion sort by (key, old rank) */
    for (int i = 1; i < cnt; i++)
    {
        int m = members[i], j = i - 1;
        while (j >= 0 && all[2 * members[j] + 1] > all[2 * m + 1]) { members[j + 1] = members[j]; j--; }
        members[j + 1] = m;
    }
    int *wr = (int *) malloc(sizeof(int) * cnt);
    int myrank = -1;
    for (int i = 0; i < cnt; i++)
    {
        wr[i] = c->wr[members[i]];
        if (members[i] == me) myrank = i;
    }
    *newcomm = new_comm(ctx, cnt, myrank, wr);
    free(wr); free(members); free(all);
    return MPI_SUCCESS;
}

/* Balanced factorisation of nnodes over the entries of dims[] that are 0 (non-increasing order), as MPI specifies. */
int MPI_Dims_create(int nnodes, int ndims, int dims[])
{
    int fixed = 1, nfree = 0;
    for (int i = 0; i < ndims; i++) { if (dims[i] > 0) fixed *= dims[i]; else nfree++; }
    if (nfree == 0) return MPI_SUCCESS;
    if (fixed <= 0 || nnodes % fixed != 0) fatal("MPI_Dims_create: %d nodes cannot be split with the given fixed dimensions", nnodes);
    int rest = nnodes / fixed;
    int *out = (int *) malloc(sizeof(int) * (size_t) nfree);
    for (int i = 0; i < nfree; i++) out[i] = 1;
    /* hand the prime factors, largest first, to the currently smallest dimension */
    int primes[64], np = 0;
    for (int d = 2; rest > 1; ) { if (rest % d == 0) { primes[np++] = d; rest /= d; } else d++; }
    for (int i = np - 1; i >= 0; i--)
    {
        int small = 0;
        for (int j = 1; j < nfree; j++) if (out[j] < out[small]) small = j;
        out[small] *= primes[i];
    }
    for (int i = 0; i < nfree; i++)                 /* sort non-increasing */
        for (int j = i + 1; j < nfree; j++) if (out[j] > out[i]) { int t = out[i]; out[i] = out[j]; out[j] = t; }
    for (int i = 0, j = 0; i < ndims; i++) if (dims[i] <= 0) dims[i] = out[j++];
    free(out);
    return MPI_SUCCESS;
}

int MPI_Comm_dup(MPI_Comm comm, MPI_Comm *newcomm)
{
    comm_t *c = get_comm(comm);
    int ctx = agree_ctx(c);
    MPI_Comm nc = new_comm(ctx, c->size, c->rank, c->wr);
    *newcomm = nc;
    return MPI_SUCCESS;
}

int MPI_Comm_free(MPI_Comm *comm)
{
    if (comm == NULL || *comm == MPI_COMM_NULL) return MPI_SUCCESS;
    if (*comm == MPI_COMM_WORLD || *comm == MPI_COMM_SELF) fatal("cannot free a predefined communicator");
    comm_t *c = get_comm(*comm);
    free(c->wr); free(c->srcs); free(c->dsts);
    memset(c, 0, sizeof(*c));
    *comm = MPI_COMM_NULL;
    return MPI_SUCCESS;
}

int MPI_Dist_graph_create_adjacent(
    MPI_Comm comm_old, int indegree, const int *sources, const int *sourceweights,
    int outdegree, const int *destinations, const int *destweights,
    MPI_Info info, int reorder, MPI_Comm *comm_dist_graph
)
{
    (void) sourceweights; (void) destweights; (void) info; (void) reorder;
    comm_t *c = get_comm(comm_old);
    int ctx = agree_ctx(c);
    MPI_Comm nc = new_comm(ctx, c->size, c->rank, c->wr);
    comm_t *g = get_comm(nc);
    g->indeg  = indegree;
    g->outdeg = outdegree;
    g->srcs = (int *) malloc(sizeof(int) * (indegree > 0 ? indegree : 1));
    g->dsts = (int *) malloc(sizeof(int) * (outdegree > 0 ? outdegree : 1));
    memcpy(g->srcs, sources, sizeof(int) * indegree);
    memcpy(g->dsts, destinations, sizeof(int) * outdegree);
    *comm_dist_graph = nc;
    return MPI_SUCCESS;
}

/* ---------------------------------------------------------- point to point */

static int src_to_world(comm_t *c, int source)
{
    if (source == MPI_ANY_SOURCE) return MPI_ANY_SOURCE;
    if (source < 0 || source >= c->size) fatal("invalid source rank %d", source);
    return c->wr[source];
}

static int world_to_comm(MPI_Comm comm, int world)
{
    comm_t *c = get_comm(comm);
    for (int i = 0; i < c->size; i++) if (c->wr[i] == world) return i;
    return world;
}

static void fill_status(MPI_Status *st, req_t *r)
{
    if (st == MPI_STATUS_IGNORE || r->kind != RQ_RECV) return;
    st->MPI_SOURCE = world_to_comm(r->comm, r->src_world);
    st->MPI_TAG    = r->src_tag;
    st->MPI_ERROR  = MPI_SUCCESS;
    st->minimpi_nbytes = r->nbytes;
}

int MPI_Isend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm, MPI_Request *req)
{
    comm_t *c = get_comm(comm);
    if (dest < 0 || dest >= c->size) fatal("MPI_Isend: invalid destination %d", dest);
    *req = post_send(buf, (size_t) count * dt_size(dt), c->wr[dest], P2P_CTX(c), tag);
    return MPI_SUCCESS;
}

int MPI_Irecv(void *buf, int count, MPI_Datatype dt, int source, int tag, MPI_Comm comm, MPI_Request *req)
{
    comm_t *c = get_comm(comm);
    *req = post_recv(buf, (size_t) count * dt_size(dt), src_to_world(c, source), P2P_CTX(c), tag, comm);
    return MPI_SUCCESS;
}

int MPI_Wait(MPI_Request *req, MPI_Status *status)
{
    if (req == NULL || *req == MPI_REQUEST_NULL) return MPI_SUCCESS;
    req_t *r = *req;
    wait_done(r);
    fill_status(status, r);
    free(r);
    *req = MPI_REQUEST_NULL;
    return MPI_SUCCESS;
}

int MPI_Waitall(int count, MPI_Request reqs[], MPI_Status statuses[])
{
    for (int i = 0; i < count; i++)
        MPI_Wait(&reqs[i], statuses == MPI_STATUSES_IGNORE ? MPI_STATUS_IGNORE : &statuses[i]);
    return MPI_SUCCESS;
}

int MPI_Send(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm)
{
    MPI_Request r;
    MPI_Isend(buf, count, dt, dest, tag, comm, &r);
    return MPI_Wait(&r, MPI_STATUS_IGNORE);
}

int MPI_Recv(void *buf, int count, MPI_Datatype dt, int source, int tag, MPI_Comm comm, MPI_Status *status)
{
    MPI_Request r;
    MPI_Irecv(buf, count, dt, source, tag, comm, &r);
    return MPI_Wait(&r, status);
}

int MPI_Get_count(const MPI_Status *status, MPI_Datatype dt, int *count)
{
    *count = (int) (status->minimpi_nbytes / dt_size(dt));
    return MPI_SUCCESS;
}

int MPI_Type_size(MPI_Datatype dt, int *size) { *size = (int) dt_size(dt); return MPI_SUCCESS; }

/* -------------------------------------------------------------- collectives */

static MPI_Request done_request(void)
{
    req_t *r = (req_t *) calloc(1, sizeof(req_t));
    r->kind = RQ_DONE;
    r->done = 1;
    return r;
}

int MPI_Barrier(MPI_Comm comm)
{
    comm_t *c = get_comm(comm);
    int cctx = COLL_CTX(c);
    char z = 0;
    if (c->size == 1) return MPI_SUCCESS;
    if (c->rank == 0)
    {
        for (int i = 1; i < c->size; i++) xrecv(&z, 0, c->wr[i], cctx, T_BARRIER);
        for (int i = 1; i < c->size; i++) xsend(&z, 0, c->wr[i], cctx, T_BARRIER);
    } else {
        xsend(&z, 0, c->wr[0], cctx, T_BARRIER);
        xrecv(&z, 0, c->wr[0], cctx, T_BARRIER);
    }
    return MPI_SUCCESS;
}

int MPI_Bcast(void *buf, int count, MPI_Datatype dt, int root, MPI_Comm comm)
{
    comm_t *c = get_comm(comm);
    size_t nbytes = (size_t) count * dt_size(dt);
    int cctx = COLL_CTX(c);
    if (c->size == 1) return MPI_SUCCESS;
    if (c->rank == root)
    {
        req_t **rs = (req_t **) malloc(sizeof(req_t *) * c->size);
        for (int i = 0; i < c->size; i++) rs[i] = (i == root) ? NULL : post_send(buf, nbytes, c->wr[i], cctx, T_BCAST);
        for (int i = 0; i < c->size; i++) if (rs[i]) finish_req(rs[i]);
        free(rs);
    } else {
        xrecv(buf, nbytes, c->wr[root], cctx, T_BCAST);
    }
    return MPI_SUCCESS;
}

int MPI_Ibcast(void *buf, int count, MPI_Datatype dt, int root, MPI_Comm comm, MPI_Request *req)
{
    MPI_Bcast(buf, count, dt, root, comm);
    *req = done_request();
    return MPI_SUCCESS;
}

int MPI_Gatherv(const void *sbuf, int scount, MPI_Datatype sdt, void *rbuf, const int rcounts[], const int displs[], MPI_Datatype rdt, int root, MPI_Comm comm)
{
    comm_t *c = get_comm(comm);
    int cctx = COLL_CTX(c);
    if (c->rank == root)
    {
        size_t es = dt_size(rdt);
        req_t **rs = (req_t **) malloc(sizeof(req_t *) * c->size);
        for (int i = 0; i < c->size; i++)
        {
            char *dst = (char *) rbuf + (size_t) displs[i] * es;
            if (i == root)
            {
                rs[i] = NULL;
                if (sbuf != MPI_IN_PLACE) memcpy(dst, sbuf, (size_t) scount * dt_size(sdt));
            } else {
                rs[i] = post_recv(dst, (size_t) rcounts[i] * es, c->wr[i], cctx, T_GATHER, comm);
            }
        }
        for (int i = 0; i < c->size; i++) if (rs[i]) finish_req(rs[i]);
        free(rs);
    } else {
        xsend(sbuf, (size_t) scount * dt_size(sdt), c->wr[root], cctx, T_GATHER);
    }
    return MPI_SUCCESS;
}

int MPI_Gather(const void *sbuf, int scount, MPI_Datatype sdt, void *rbuf, int rcount, MPI_Datatype rdt, int root, MPI_Comm comm)
{
    comm_t *c = get_comm(comm);
    int n = c->size;
    int *cnt = (int *) malloc(sizeof(int) * n), *dsp = (int *) malloc(sizeof(int) * n);
    for (int i = 0; i < n; i++) { cnt[i] = rcount; dsp[i] = i * rcount; }
    MPI_Gatherv(sbuf, scount, sdt, rbuf, cnt, dsp, rdt, root, comm);
    free(cnt); free(dsp);
    return MPI_SUCCESS;
}

int MPI_Scatterv(const void *sbuf, const int scounts[], const int displs[], MPI_Datatype sdt, void *rbuf, int rcount, MPI_Datatype rdt, int root, MPI_Comm comm)
{
    comm_t *c = get_comm(comm);
    int cctx = COLL_CTX(c);
    if (c->rank == root)
    {
        size_t es = dt_size(sdt);
        req_t **rs = (req_t **) malloc(sizeof(req_t *) * c->size);
        for (int i = 0; i < c->size; i++)
        {
            const char *src = (const char *) sbuf + (size_t) displs[i] * es;
            if (i == root)
            {
                rs[i] = NULL;
                if (rbuf != MPI_IN_PLACE) memcpy(rbuf, src, (size_t) scounts[i] * es);
            } else {
                rs[i] = post_send(src, (size_t) scounts[i] * es, c->wr[i], cctx, T_SCATTER);
            }
        }
        for (int i = 0; i < c->size; i++) if (rs[i]) finish_req(rs[i]);
        free(rs);
    } else {
        xrecv(rbuf, (size_t) rcount * dt_size(rdt), c->wr[root], cctx, T_SCATTER);
    }
    return MPI_SUCCESS;
}

int MPI_Iscatterv(const void *sbuf, const int scounts[], const int displs[], MPI_Datatype sdt, void *rbuf, int rcount, MPI_Datatype rdt, int root, MPI_Comm comm, MPI_Request *req)
{
    MPI_Scatterv(sbuf, scounts, displs, sdt, rbuf, rcount, rdt, root, comm);
    *req = done_request();
    return MPI_SUCCESS;
}

int MPI_Allgatherv(const void *sbuf, int scount, MPI_Datatype sdt, void *rbuf, const int rcounts[], const int displs[], MPI_Datatype rdt, MPI_Comm comm)
{
    comm_t *c = get_comm(comm);
    int cctx = COLL_CTX(c), n = c->size, me = c->rank;
    size_t es = dt_size(rdt);
    char *mine = (char *) rbuf + (size_t) displs[me] * es;
    size_t mybytes;
    if (sbuf == MPI_IN_PLACE)
    {
        mybytes = (size_t) rcounts[me] * es;
    } else {
        mybytes = (size_t) scount * dt_size(sdt);
        memcpy(mine, sbuf, mybytes);
    }
    if (n == 1) return MPI_SUCCESS;
    req_t **rs = (req_t **) malloc(sizeof(req_t *) * 2 * n);
    int k = 0;
    for (int i = 0; i < n; i++)
    {
        if (i == me) continue;
        rs[k++] = post_recv((char *) rbuf + (size_t) displs[i] * es, (size_t) rcounts[i] * es, c->wr[i], cctx, T_ALLGATHER, comm);
    }
    for (int i = 0; i < n; i++)
    {
        if (i == me) continue;
        rs[k++] = post_send(mine, mybytes, c->wr[i], cctx, T_ALLGATHER);
    }
    for (int i = 0; i < k; i++) finish_req(rs[i]);
    free(rs);
    return MPI_SUCCESS;
}

int MPI_Iallgatherv(const void *sbuf, int scount, MPI_Datatype sdt, void *rbuf, const int rcounts[], const int displs[], MPI_Datatype rdt, MPI_Comm comm, MPI_Request *req)
{
    MPI_Allgatherv(sbuf, scount, sdt, rbuf, rcounts, displs, rdt, comm);
    *req = done_request();
    return MPI_SUCCESS;
}

int MPI_Allgather(const void *sbuf, int scount, MPI_Datatype sdt, void *rbuf, int rcount, MPI_Datatype rdt, MPI_Comm comm)
{
    comm_t *c = get_comm(comm);
    int n = c->size;
    int *cnt = (int *) malloc(sizeof(int) * n), *dsp = (int *) malloc(sizeof(int) * n);
    for (int i = 0; i < n; i++) { cnt[i] = rcount; dsp[i] = i * rcount; }
    MPI_Allgatherv(sbuf, scount, sdt, rbuf, cnt, dsp, rdt, comm);
    free(cnt); free(dsp);
    return MPI_SUCCESS;
}

int MPI_Alltoallv(const void *sbuf, const int scounts[], const int sdispls[], MPI_Datatype sdt, void *rbuf, const int rcounts[], const int rdispls[], MPI_Datatype rdt, MPI_Comm comm)
{
    comm_t *c = get_comm(comm);
    int cctx = COLL_CTX(c), n = c->size, me = c->rank;
    size_t ss = dt_size(sdt), rs_ = dt_size(rdt);
    req_t **rs = (req_t **) malloc(sizeof(req_t *) * 2 * n);
    int k = 0;
    for (int i = 0; i < n; i++)
    {
        if (i == me) continue;
        rs[k++] = post_recv((char *) rbuf + (size_t) rdispls[i] * rs_, (size_t) rcounts[i] * rs_, c->wr[i], cctx, T_ALLTOALL, comm);
    }
    memcpy((char *) rbuf + (size_t) rdispls[me] * rs_, (const char *) sbuf + (size_t) sdispls[me] * ss, (size_t) scounts[me] * ss);
    for (int i = 0; i < n; i++)
    {
        if (i == me) continue;
        rs[k++] = post_send((const char *) sbuf + (size_t) sdispls[i] * ss, (size_t) scounts[i] * ss, c->wr[i], cctx, T_ALLTOALL);
    }
    for (int i = 0; i < k; i++) finish_req(rs[i]);
    free(rs);
    return MPI_SUCCESS;
}

int MPI_Alltoall(const void *sbuf, int scount, MPI_Datatype sdt, void *rbuf, int rcount, MPI_Datatype rdt, MPI_Comm comm)
{
    comm_t *c = get_comm(comm);
    int n = c->size;
    int *sc = (int *) malloc(sizeof(int) * n), *sd = (int *) malloc(sizeof(int) * n);
    int *rc = (int *) malloc(sizeof(int) * n), *rd = (int *) malloc(sizeof(int) * n);
    for (int i = 0; i < n; i++) { sc[i] = scount; sd[i] = i * scount; rc[i] = rcount; rd[i] = i * rcount; }
    MPI_Alltoallv(sbuf, sc, sd, sdt, rbuf, rc, rd, rdt, comm);
    free(sc); free(sd); free(rc); free(rd);
    return MPI_SUCCESS;
}

int MPI_Neighbor_alltoallv(const void *sbuf, const int scounts[], const int sdispls[], MPI_Datatype sdt, void *rbuf, const int rcounts[], const int rdispls[], MPI_Datatype rdt, MPI_Comm comm)
{
    comm_t *c = get_comm(comm);
    int cctx = COLL_CTX(c);
    size_t ss = dt_size(sdt), rs_ = dt_size(rdt);
    int nreq = c->indeg + c->outdeg, k = 0;
    req_t **rs = (req_t **) malloc(sizeof(req_t *) * (nreq > 0 ? nreq : 1));
    for (int i = 0; i < c->indeg; i++)
        rs[k++] = post_recv((char *) rbuf + (size_t) rdispls[i] * rs_, (size_t) rcounts[i] * rs_, c->wr[c->srcs[i]], cctx, T_NEIGHBOR, comm);
    for (int i = 0; i < c->outdeg; i++)
        rs[k++] = post_send((const char *) sbuf + (size_t) sdispls[i] * ss, (size_t) scounts[i] * ss, c->wr[c->dsts[i]], cctx, T_NEIGHBOR);
    for (int i = 0; i < k; i++) finish_req(rs[i]);
    free(rs);
    return MPI_SUCCESS;
}

#define REDUCE_LOOP(T)                                                   \
    do {                                                                 \
        T *a = (T *) acc; const T *b = (const T *) in;                   \
        for (int i = 0; i < count; i++)                                  \
        {                                                                \
            if (op == MPI_SUM) a[i] = a[i] + b[i];                       \
            else if (op == MPI_MAX) a[i] = (b[i] > a[i]) ? b[i] : a[i];  \
            else a[i] = (b[i] < a[i]) ? b[i] : a[i];                     \
        }                                                                \
    } while (0)

static void reduce_into(void *acc, const void *in, int count, MPI_Datatype dt, MPI_Op op)
{
    if (op != MPI_SUM && op != MPI_MAX && op != MPI_MIN) fatal("unsupported reduction op %d", op);
    switch (dt_kind(dt))
    {
        case 1: REDUCE_LOOP(signed char); break;
        case 2: REDUCE_LOOP(unsigned char); break;
        case 3: REDUCE_LOOP(int32_t); break;
        case 4: REDUCE_LOOP(uint32_t); break;
        case 5: REDUCE_LOOP(int64_t); break;
        case 6: REDUCE_LOOP(uint64_t); break;
        case 7: REDUCE_LOOP(float); break;
        case 8: REDUCE_LOOP(double); break;
        default: fatal("unsupported reduction datatype %d", dt);
    }
}

int MPI_Reduce(const void *sbuf, void *rbuf, int count, MPI_Datatype dt, MPI_Op op, int root, MPI_Comm comm)
{
    comm_t *c = get_comm(comm);
    int cctx = COLL_CTX(c);
    size_t nbytes = (size_t) count * dt_size(dt);
    if (c->rank != root)
    {
        xsend(sbuf, nbytes, c->wr[root], cctx, T_REDUCE);
        return MPI_SUCCESS;
    }
    /* fold contributions in rank order so that results are reproducible */
    char *mine = (char *) malloc(nbytes ? nbytes : 1);
    char *tmp  = (char *) malloc(nbytes ? nbytes : 1);
    memcpy(mine, (sbuf == MPI_IN_PLACE) ? rbuf : sbuf, nbytes);
    for (int i = 0; i < c->size; i++)
    {
        const char *contrib;
        if (i == root) contrib = mine;
        else { xrecv(tmp, nbytes, c->wr[i], cctx, T_REDUCE); contrib = tmp; }
        if (i == 0) memcpy(rbuf, contrib, nbytes);
        else reduce_into(rbuf, contrib, count, dt, op);
    }
    free(mine); free(tmp);
    return MPI_SUCCESS;
}

int MPI_Allreduce(const void *sbuf, void *rbuf, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm)
{
    MPI_Reduce(sbuf, rbuf, count, dt, op, 0, comm);
    MPI_Bcast(rbuf, count, dt, 0, comm);
    return MPI_SUCCESS;
}
