/*
 * crp_common.c - process-wide options, device binding and the pinned-buffer cache.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "utils.h"
#include "crp_ext.h"
#include "crp_internal.h"

static void *g_user_stream = NULL;
static int   g_blocking = 1;
static int   g_dev_state = -1;      /* -1 unknown, 0 no device (plan-only), 1 bound */

const char *crp_version(void) { return "crpspmm-b200 0.1 (sm_100a, NCCL data plane)"; }

void crp_set_stream(void *stream) { g_user_stream = stream; }
void crp_set_blocking(const int blocking) { g_blocking = blocking ? 1 : 0; }
void *crp_opt_stream(void) { return g_user_stream; }
int crp_opt_blocking(void) { return g_blocking; }

int crp_opt_plan_only(void)
{
    int v;
    GET_ENV_INT_VAR(v, "CRP_SPMM_PLAN_ONLY", "plan_only", 0, 0, 1, 0);
    return v;
}

int crp_opt_pin_host(void)
{
    static int v = -1;
    /* OPT-IN (default 0): a cudaHostRegister registration outlives the caller's buffer - if the caller frees B or C and a
     * later malloc returns the same address, CUDA would still DMA from / to the old physical pages.  Callers that keep their
     * buffers for the life of the engine (the reference drivers do) may set CRP_SPMM_PIN_HOST=1; registrations are dropped
     * by crp_unpin_host_all() and whenever an engine is freed. */
    if (v < 0) GET_ENV_INT_VAR(v, "CRP_SPMM_PIN_HOST", "pin_host", 0, 0, 1, 0);
    return v;
}

int crp_device_ready(void)
{
    if (g_dev_state >= 0) return g_dev_state;
    if (crp_opt_plan_only())
    {
        g_dev_state = 0;
        return 0;
    }
    if (crp_cuda_device_count() <= 0)
    {
        fprintf(stderr,
            "[FATAL] CRP-SpMM (B200 build): no CUDA device is visible and there is no CPU fallback.\n"
            "        (Host-side planning alone can be exercised with CRP_SPMM_PLAN_ONLY=1.)\n");
        fflush(stderr);
        abort();
    }
    crp_cuda_select_device_by_local_rank();
    g_dev_state = 1;
    return 1;
}

/* Ranks of this job that run on this node (they share its GPUs): what the launcher says, else the world size (single node). */
int crp_ranks_on_this_node(const int world_size)
{
    static const char *names[] = { "OMPI_COMM_WORLD_LOCAL_SIZE", "MV2_COMM_WORLD_LOCAL_SIZE", "MPI_LOCALNRANKS", "SLURM_NTASKS_PER_NODE", "LOCAL_WORLD_SIZE" };
    for (size_t i = 0; i < sizeof(names) / sizeof(names[0]); i++)
    {
        const char *e = getenv(names[i]);
        if (e != NULL && e[0] >= '1' && e[0] <= '9')
        {
            const int v = atoi(e);
            if (v >= 1 && v <= world_size) return v;
        }
    }
    return world_size;
}

/* GPU-side plan construction (csrc/cuda/plan_build.cu) for matrices of at least CRP_SPMM_GPU_PLAN_MIN_NNZ nonzeros (default 2M:
 * below that the copies cost more than the host loops), when a device is usable; CRP_SPMM_GPU_PLAN=0 keeps everything on the host.
 * Never aborts: the partitioner can be used on a machine without a GPU. */
int crp_gpu_plan_enabled(const long long nnz)
{
    static int on = -1, min_nnz = 0;
    if (on < 0)
    {
        GET_ENV_INT_VAR(on, "CRP_SPMM_GPU_PLAN", "gpu_plan", 1, 0, 1, 0);
        GET_ENV_INT_VAR(min_nnz, "CRP_SPMM_GPU_PLAN_MIN_NNZ", "gpu_plan_min_nnz", 2000000, 0, 2147483647, 0);
    }
    if (!on || nnz < (long long) min_nnz || nnz <= 0) return 0;
    if (g_dev_state == 0 || crp_opt_plan_only()) return 0;
    if (g_dev_state < 0 && crp_cuda_device_count() <= 0) return 0;
    return crp_device_ready();
}

/* ---- pinned caller buffers ---- */
typedef struct { const char *ptr; size_t bytes; } pin_entry;
static pin_entry g_pins[64];
static int g_npin = 0;

void crp_pin_host_range(const void *ptr, size_t bytes)
{
    if (!crp_opt_pin_host() || ptr == NULL || bytes < ((size_t) 1 << 20)) return;
    const char *p = (const char *) ptr;
    for (int i = 0; i < g_npin; i++)
        if (p >= g_pins[i].ptr && p + bytes <= g_pins[i].ptr + g_pins[i].bytes) return;
    /* an overlapping but different range (e.g. realloc'd buffer): drop the stale registration first */
    for (int i = 0; i < g_npin; i++)
    {
        if (p < g_pins[i].ptr + g_pins[i].bytes && g_pins[i].ptr < p + bytes)
        {
            crp_cuda_host_unregister(g_pins[i].ptr);
            g_pins[i] = g_pins[--g_npin];
            i--;
        }
    }
    if (g_npin == (int) (sizeof(g_pins) / sizeof(g_pins[0])))
    {
        crp_cuda_host_unregister(g_pins[0].ptr);
        memmove(&g_pins[0], &g_pins[1], sizeof(pin_entry) * (size_t) (g_npin - 1));
        g_npin--;
    }
    if (crp_cuda_host_register(ptr, bytes))
    {
        g_pins[g_npin].ptr = p;
        g_pins[g_npin].bytes = bytes;
        g_npin++;
    }
}

void crp_unpin_host_all(void)
{
    for (int i = 0; i < g_npin; i++) crp_cuda_host_unregister(g_pins[i].ptr);
    g_npin = 0;
}

int *crp_comm_ranks_in_parent(MPI_Comm sub, MPI_Comm parent)
{
    int n, mine;
    MPI_Comm_size(sub, &n);
    MPI_Comm_rank(parent, &mine);
    int *tab = (int *) malloc(sizeof(int) * (size_t) n);
    MPI_Allgather(&mine, 1, MPI_INT, tab, 1, MPI_INT, sub);
    return tab;
}
