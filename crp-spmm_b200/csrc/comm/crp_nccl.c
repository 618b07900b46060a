/*
 * crp_nccl.c - the NCCL data plane: communicator cache and grouped send/recv.
 *
 * NCCL has no allgatherv / alltoallv / neighbourhood collectives, so every
 * exchange of the reference (MPI_Iallgatherv of A, src/para2d_spmm.c:81-83; the
 * P2P ring / MPI_Alltoallv of B rows, src/rowpara_spmm.c:280-308;
 * MPI_Neighbor_alltoallv, src/mat_redist.c:357-360) becomes one
 * ncclGroupStart .. ncclGroupEnd of ncclSend / ncclRecv on device buffers, with
 * zero-size peers skipped and byte counts identical to the reference's
 * element counts.  On NVSwitch every peer is reachable at full NVLink
 * bandwidth, so the reference's ring-offset ordering is not needed.
 *
 * libnccl is loaded at run time (dlopen) so that single-GPU use and the CPU
 * test-suite never need it: first the copy a host process already loaded
 * (torch's), else libnccl.so.2 from the default search path.
 * Bootstrap: rank 0 of the MPI communicator creates the ncclUniqueId and
 * MPI_Bcast's it - the control plane stays on MPI.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <nccl.h>

#include "utils.h"
#include "crp_internal.h"

struct crp_nccl_comm
{
    ncclComm_t comm;
    int size, rank;
    int *world_ranks;      /* membership key */
};

static struct
{
    void *handle;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*GroupStart)(void);
    ncclResult_t (*GroupEnd)(void);
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    const char *(*GetErrorString)(ncclResult_t);
    ncclResult_t (*GetVersion)(int *);
} g_nccl;

static crp_nccl_comm **g_cache = NULL;
static int g_ncache = 0;
static unsigned long long g_groups = 0;

unsigned long long crp_nccl_group_count(void) { return g_groups; }

#define NCCL_CHECK(call)                                                                        \
    do {                                                                                        \
        ncclResult_t r_ = (call);                                                               \
        if (r_ != ncclSuccess)                                                                  \
        {                                                                                       \
            fprintf(stderr, "[FATAL] NCCL error %d (%s) at %s:%d\n", (int) r_,                  \
                    g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?", __FILE__, __LINE__); \
            fflush(stderr);                                                                     \
            abort();                                                                            \
        }                                                                                       \
    } while (0)

static void *must_sym(const char *name)
{
    void *p = dlsym(g_nccl.handle, name);
    if (p == NULL)
    {
        fprintf(stderr, "[FATAL] libnccl does not export %s\n", name);
        abort();
    }
    return p;
}

static void load_nccl(void)
{
    if (g_nccl.handle != NULL) return;
    const char *override = getenv("CRP_NCCL_LIB");
    if (override && override[0]) g_nccl.handle = dlopen(override, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle == NULL) g_nccl.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   /* already in the process? */
    if (g_nccl.handle == NULL) g_nccl.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle == NULL) g_nccl.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle == NULL)
    {
        fprintf(stderr, "[FATAL] cannot load libnccl.so.2 (%s); multi-GPU runs need NCCL (set CRP_NCCL_LIB)\n", dlerror());
        abort();
    }
    *(void **) &g_nccl.GetUniqueId    = must_sym("ncclGetUniqueId");
    *(void **) &g_nccl.CommInitRank   = must_sym("ncclCommInitRank");
    *(void **) &g_nccl.CommDestroy    = must_sym("ncclCommDestroy");
    *(void **) &g_nccl.GroupStart     = must_sym("ncclGroupStart");
    *(void **) &g_nccl.GroupEnd       = must_sym("ncclGroupEnd");
    *(void **) &g_nccl.Send           = must_sym("ncclSend");
    *(void **) &g_nccl.Recv           = must_sym("ncclRecv");
    *(void **) &g_nccl.GetErrorString = must_sym("ncclGetErrorString");
    *(void **) &g_nccl.GetVersion     = must_sym("ncclGetVersion");
}

crp_nccl_comm *crp_nccl_get(MPI_Comm comm)
{
    int n, me, wme;
    MPI_Comm_size(comm, &n);
    MPI_Comm_rank(comm, &me);
    MPI_Comm_rank(MPI_COMM_WORLD, &wme);
    int *key = (int *) malloc(sizeof(int) * (size_t) n);
    MPI_Allgather(&wme, 1, MPI_INT, key, 1, MPI_INT, comm);
    for (int i = 0; i < g_ncache; i++)
    {
        crp_nccl_comm *c = g_cache[i];
        if (c->size == n && memcmp(c->world_ranks, key, sizeof(int) * (size_t) n) == 0)
        {
            free(key);
            return c;
        }
    }
    load_nccl();
    crp_device_ready();
    ncclUniqueId id;
    memset(&id, 0, sizeof(id));
    if (me == 0) NCCL_CHECK(g_nccl.GetUniqueId(&id));
    MPI_Bcast(&id, (int) sizeof(id), MPI_BYTE, 0, comm);
    crp_nccl_comm *c = (crp_nccl_comm *) calloc(1, sizeof(crp_nccl_comm));
    c->size = n;
    c->rank = me;
    c->world_ranks = key;
    NCCL_CHECK(g_nccl.CommInitRank(&c->comm, n, id, me));
    g_cache = (crp_nccl_comm **) realloc(g_cache, sizeof(crp_nccl_comm *) * (size_t) (g_ncache + 1));
    g_cache[g_ncache++] = c;
    return c;
}

/* Creates (or finds) the NCCL communicator of MPI_COMM_WORLD and returns its size: every data-plane communicator
 * (grid rows, grid columns, redistribution groups) is a subset of it.  bench.py calls this once per multi-GPU run
 * so that the NCCL INIT lines of the log show all ranks. */
int crp_nccl_world_nranks(void)
{
    return crp_nccl_get(MPI_COMM_WORLD)->size;
}

int crp_nccl_version(void)
{
    int v = 0;
    load_nccl();
    NCCL_CHECK(g_nccl.GetVersion(&v));
    return v;
}

int crp_nccl_rank(const crp_nccl_comm *nc) { return nc->rank; }
int crp_nccl_size(const crp_nccl_comm *nc) { return nc->size; }

void crp_nccl_group_start(void)
{
    load_nccl();
    NCCL_CHECK(g_nccl.GroupStart());
}

void crp_nccl_group_end(void)
{
    NCCL_CHECK(g_nccl.GroupEnd());
    g_groups++;
}

void crp_nccl_send(crp_nccl_comm *nc, const void *buf, size_t bytes, int peer, void *stream)
{
    NCCL_CHECK(g_nccl.Send(buf, bytes, ncclInt8, peer, nc->comm, (cudaStream_t) stream));
}

void crp_nccl_recv(crp_nccl_comm *nc, void *buf, size_t bytes, int peer, void *stream)
{
    NCCL_CHECK(g_nccl.Recv(buf, bytes, ncclInt8, peer, nc->comm, (cudaStream_t) stream));
}

void crp_nccl_shutdown(void)
{
    for (int i = 0; i < g_ncache; i++)
    {
        if (g_nccl.CommDestroy) g_nccl.CommDestroy(g_cache[i]->comm);
        free(g_cache[i]->world_ranks);
        free(g_cache[i]);
    }
    free(g_cache);
    g_cache = NULL;
    g_ncache = 0;
}
