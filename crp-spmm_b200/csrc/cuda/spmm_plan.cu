// spmm_plan.cu - SpMM plan (device CSR) management and kernel dispatch.
// crp_cuda_spmm_plan_create / _exec / _destroy stand where the reference calls
// mkl_sparse_d_create_csr / mkl_sparse_d_mm / mkl_sparse_destroy
// (src/rowpara_spmm.c:398-408) - but the handle is built once at init, not per exec.
#include <cstring>
#include <vector>

#include "crp_cuda_internal.cuh"

template <typename T, int VECN>
void crp_launch_rowsplit(
    const crp_spmm_plan *plan, const int nrows, const int *row_list, const int *pbeg, const int *pend, const T *val, const int n,
    const T *X0, size_t ldx0, int x0_rows, const T *X1, size_t ldx1, T alpha, T beta, T *C, size_t ldc, cudaStream_t stream
);
template <typename T>
void crp_launch_longrow_reduce(const crp_longrows *lr, const int n, T alpha, T beta, T *C, size_t ldc, cudaStream_t s);
template <typename T, int VECN>
bool crp_launch_mergepath(
    crp_spmm_plan *plan, const T *val, const int n, const T *X0, size_t ldx0, const T *X1, size_t ldx1, T alpha, T beta, T *C, size_t ldc, cudaStream_t s
);
template <typename T, int VEC>
bool crp_launch_panel(
    crp_spmm_plan *plan, const T *val, const int n, const T *X0, size_t ldx0, const T *X1, size_t ldx1, T alpha, T beta, T *C, size_t ldc,
    const crp_spmm_wait *wait, const crp_spmm_put *put, bool *rest_done, cudaStream_t s
);
template <typename T, int VEC>
void crp_launch_rowgroup(const crp_rowgroup *rg, const T *bval, int nv, const T *X0, size_t ldx0, int x0_rows, const T *X1, size_t ldx1, T alpha, T beta, T *C, size_t ldc, cudaStream_t s);

// Reuse-distance histogram of the B rows (crp_reuse in crp_cuda_internal.cuh).
static void reuse_profile(crp_reuse *ru, const int m, const int k, const int *rowptr, const int *colidx)
{
    memset(ru, 0, sizeof(*ru));
    if (m <= 0 || k <= 0 || rowptr[m] <= rowptr[0]) return;
    std::vector<int> last((size_t) k, -1);
    const long long ntile = ((long long) m + CRP_REUSE_TILE - 1) / CRP_REUSE_TILE;
    long long uses = 0;
    for (long long t = 0; t < ntile; t++)
    {
        const int r0 = (int) (t * CRP_REUSE_TILE), r1 = (int) ((t + 1) * CRP_REUSE_TILE < m ? (t + 1) * CRP_REUSE_TILE : m);
        for (int q = rowptr[r0]; q < rowptr[r1]; q++)
        {
            const int c = colidx[q];
            if (c < 0 || c >= k) continue;
            const int l = last[(size_t) c];
            if (l == (int) t) continue;
            uses++;
            if (l < 0) ru->first_uses++;
            else
            {
                unsigned d = (unsigned) ((int) t - l);
                int b = 0;
                while (d > 1u) { d >>= 1; b++; }
                ru->reuses[b]++;
            }
            last[(size_t) c] = (int) t;
        }
    }
    ru->ntile = ntile;
    ru->union_per_tile = (double) uses / (double) ntile;
    ru->new_per_tile = (double) ru->first_uses / (double) ntile;
}

// Number of column passes for a product with n dense columns of es bytes.  Traffic model: a reuse at distance d blocks is an
// L2 hit when the rows touched in between - the block's own union, (d - 1) blocks' new B rows, d blocks of C rows and of A -
// fit into half of the L2; every pass re-reads A.  P is raised (1, 2, 4, 8) only when the modelled DRAM traffic drops by
// more than 15 %.  A pass is at least 64 columns wide and a multiple of 64 columns (16-byte alignment for both types).
static int model_passes(const crp_reuse *ru, const int m, const long long nnz, const int n, const int es, const double l2_bytes)
{
    if (ru->ntile == 0 || n < 128) return 1;
    const double budget = 0.5 * l2_bytes;
    const double a_bytes = (double) nnz * (4.0 + es) + 4.0 * m;
    const double a_tile = a_bytes / (double) ru->ntile;
    double best = 0.0;
    int bestP = 1;
    for (int P = 1; P <= 8; P *= 2)
    {
        const int ns = ((n + P - 1) / P + 63) / 64 * 64;
        if (P > 1 && (ns < 64 || ns >= n)) break;
        const double row = (double) ns * es;
        double hits = 0.0;
        for (int b = 0; b < CRP_REUSE_BUCKETS; b++)
        {
            if (ru->reuses[b] == 0) continue;
            const double d = (double) (1ull << b) * 1.5;          // middle of the bucket
            const double foot = (ru->union_per_tile + (d - 1.0) * ru->new_per_tile + d * CRP_REUSE_TILE) * row + d * a_tile;
            if (foot <= budget) hits += (double) ru->reuses[b];
        }
        double all = (double) ru->first_uses;
        for (int b = 0; b < CRP_REUSE_BUCKETS; b++) all += (double) ru->reuses[b];
        const double traffic = (double) ((n + ns - 1) / ns) * a_bytes + (all - hits) * (double) n * es + (double) m * n * es;
        if (P == 1 || traffic < 0.85 * best) { best = traffic; bestP = P; }
    }
    return bestP;
}

static bool passes_env_is_model()
{
    const char *e = getenv("CRP_SPMM_PASSES");
    return e != NULL && strcmp(e, "model") == 0;
}

// Column passes are OPT-IN: measured on B200 (profiles/r02_kbench_stencil_passes.log) one pass is the fastest on every BASELINE
// shape - the 126 MB L2 and the kernels' access order already keep the re-read B rows close, and every extra pass re-reads A
// and the panel's meta records.  CRP_SPMM_PASSES=<count> / crp_cuda_spmm_set_passes force a count, CRP_SPMM_PASSES=model
// (set when the plan is created) lets the traffic model above decide, for GPUs with a smaller L2.
static int choose_passes(const crp_spmm_plan *p, const int n, const int es)
{
    if (p->passes_forced > 0) return p->passes_forced;
    const char *e = getenv("CRP_SPMM_PASSES");
    const int env = (e && e[0]) ? atoi(e) : 0;
    if (env > 0) return env;
    if (!passes_env_is_model() || p->reuse.ntile == 0) return 1;
    static double l2_bytes = 0.0;
    if (l2_bytes == 0.0)
    {
        int dev = 0, v = 0;
        CRP_CUDA_CHECK(cudaGetDevice(&dev));
        CRP_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, dev));
        l2_bytes = (v > 0) ? (double) v : 64.0e6;
    }
    return model_passes(&p->reuse, p->m, p->nnz, n, es, l2_bytes);
}

// host-only view of the decision (no device needed): column passes the exec would use for this matrix, width and L2 size
extern "C" int crp_cuda_spmm_model_passes(const int m, const int k, const int *rowptr_h, const int *colidx_h, const int n, const int elem_size, const double l2_bytes)
{
    crp_reuse ru;
    reuse_profile(&ru, m, k, rowptr_h, colidx_h);
    return model_passes(&ru, m, (m > 0) ? (long long) rowptr_h[m] - rowptr_h[0] : 0, n, elem_size, l2_bytes);
}

extern "C" void crp_cuda_spmm_set_passes(crp_spmm_plan *plan, const int passes) { if (plan != NULL) plan->passes_forced = passes; }
extern "C" int crp_cuda_spmm_last_passes(const crp_spmm_plan *plan) { return plan ? plan->last_passes : 0; }

extern "C" crp_spmm_plan *crp_cuda_spmm_plan_create(const int m, const int k, const int x0_rows, const int *rowptr_h, const int *colidx_h, const double *val_h, const int n_hint)
{
    crp_spmm_plan *p = (crp_spmm_plan *) calloc(1, sizeof(crp_spmm_plan));
    p->m = m;
    p->k = k;
    p->x0_rows = (x0_rows > k) ? k : x0_rows;
    p->nnz = (m > 0) ? (long long) rowptr_h[m] - rowptr_h[0] : 0;
    p->n_hint = n_hint;
    p->variant = CRP_VARIANT_AUTO;
    p->last_kernel = "none";
    if (m > 0 && rowptr_h[0] != 0) { fprintf(stderr, "[FATAL] crp_cuda_spmm_plan_create: rowptr must be 0-based\n"); abort(); }
    int mx = 0;
    for (int i = 0; i < m; i++) { const int d = rowptr_h[i + 1] - rowptr_h[i]; if (d > mx) mx = d; }
    p->max_row_nnz = mx;
    p->avg_row_nnz = (m > 0) ? (double) p->nnz / m : 0.0;
    CRP_CUDA_CHECK(cudaMalloc((void **) &p->d_rowptr, sizeof(int) * ((size_t) m + 1)));
    if (m > 0) CRP_CUDA_CHECK(cudaMemcpy(p->d_rowptr, rowptr_h, sizeof(int) * ((size_t) m + 1), cudaMemcpyHostToDevice));
    else CRP_CUDA_CHECK(cudaMemset(p->d_rowptr, 0, sizeof(int)));
    if (p->nnz > 0)
    {
        CRP_CUDA_CHECK(cudaMalloc((void **) &p->d_colidx, sizeof(int) * (size_t) p->nnz));
        CRP_CUDA_CHECK(cudaMalloc((void **) &p->d_val, sizeof(double) * (size_t) p->nnz));
        CRP_CUDA_CHECK(cudaMemcpy(p->d_colidx, colidx_h, sizeof(int) * (size_t) p->nnz, cudaMemcpyHostToDevice));
        CRP_CUDA_CHECK(cudaMemcpy(p->d_val, val_h, sizeof(double) * (size_t) p->nnz, cudaMemcpyHostToDevice));
    }
    if (passes_env_is_model()) reuse_profile(&p->reuse, m, k, rowptr_h, colidx_h);
    std::vector<int> rest;
    crp_rowgroup_build(p, rowptr_h, colidx_h, val_h, &rest);
    if (p->rg.R > 1) crp_longrows_build(p, rowptr_h, rest.data(), (int) rest.size());
    else crp_longrows_build(p, rowptr_h, NULL, m);
    // nnz-balanced chunks for matrices without row-group structure (kept on the host side until a launch wants them:
    // the partition is cheap, O(m)); automatic choice: skewed row lengths, see spmm_dispatch
    p->h_rowptr = (int *) malloc(sizeof(int) * ((size_t) m + 1));
    if (m > 0) memcpy(p->h_rowptr, rowptr_h, sizeof(int) * ((size_t) m + 1)); else p->h_rowptr[0] = 0;
    return p;
}

extern "C" void crp_cuda_spmm_plan_destroy(crp_spmm_plan *plan)
{
    if (plan == NULL) return;
    if (plan->d_rowptr) CRP_CUDA_CHECK(cudaFree(plan->d_rowptr));
    if (plan->d_colidx) CRP_CUDA_CHECK(cudaFree(plan->d_colidx));
    if (plan->d_val)    CRP_CUDA_CHECK(cudaFree(plan->d_val));
    if (plan->d_val32)  CRP_CUDA_CHECK(cudaFree(plan->d_val32));
    crp_longrows_destroy(plan);
    crp_rowgroup_destroy(plan);
    crp_mergepath_destroy(plan);
    free(plan->h_rowptr);
    free(plan);
}

__global__ void __launch_bounds__(256) cast_f64_to_f32_kernel(const double *__restrict__ src, float *__restrict__ dst, size_t n)
{
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
        dst[i] = (float) src[i];
}

static void cast_to_f32(const double *src, float **dst, size_t count, cudaStream_t stream)
{
    if (*dst != NULL || count == 0) return;
    CRP_CUDA_CHECK(cudaMalloc((void **) dst, sizeof(float) * count));
    size_t blocks = (count + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    cast_f64_to_f32_kernel<<<(unsigned) blocks, 256, 0, stream>>>(src, *dst, count);
    CRP_LAUNCH_CHECK();
}

// the row-split kernel over a row set, with the long rows of that set cut into segments (spmm_longrow.cu)
template <typename T, int VECN>
static void rowsplit_balanced(
    crp_spmm_plan *plan, const bool use_lr, const int nrows, const int *rows, const T *val, const int n,
    const T *X0, size_t ldx0, const T *X1, size_t ldx1, T alpha, T beta, T *C, size_t ldc, cudaStream_t s
)
{
    const int x0_rows = plan->x0_rows;
    crp_longrows *lr = &plan->lr;
    if (!use_lr || lr->nlong == 0)
    {
        crp_launch_rowsplit<T, VECN>(plan, nrows, rows, plan->d_rowptr, plan->d_rowptr + 1, val, n, X0, ldx0, x0_rows, X1, ldx1, alpha, beta, C, ldc, s);
        return;
    }
    if (lr->nshort > 0)
        crp_launch_rowsplit<T, VECN>(plan, lr->nshort, lr->d_short, plan->d_rowptr, plan->d_rowptr + 1, val, n, X0, ldx0, x0_rows, X1, ldx1, alpha, beta, C, ldc, s);
    const size_t need = sizeof(T) * (size_t) lr->nseg * (size_t) n;
    if (need > lr->scratch_bytes)
    {
        CRP_CUDA_CHECK(cudaStreamSynchronize(s));
        if (lr->d_scratch) CRP_CUDA_CHECK(cudaFree(lr->d_scratch));
        CRP_CUDA_CHECK(cudaMalloc(&lr->d_scratch, need));
        lr->scratch_bytes = need;
    }
    crp_launch_rowsplit<T, VECN>(plan, lr->nseg, NULL, lr->d_seg_beg, lr->d_seg_end, val, n, X0, ldx0, x0_rows, X1, ldx1, (T) 1, (T) 0, (T *) lr->d_scratch, (size_t) n, s);
    crp_launch_longrow_reduce<T>(lr, n, alpha, beta, C, ldc, s);
}

// A kernel that cannot wait for the neighbours' flags itself is preceded by the plain wait kernel.
static void wait_first(const crp_spmm_wait *wait, cudaStream_t s)
{
    if (wait == NULL || wait->nwait <= 0) return;
    crp_cuda_wait_flags(wait->flags, wait->wait_idx, wait->nwait, wait->epoch, 1e-9 * (double) wait->timeout_ns, wait->err, (void *) s);
}

template <typename T, int VECN>
static void spmm_dispatch(
    crp_spmm_plan *plan, const T *val, const T *bval, const int n, const T *X0, size_t ldx0, const T *X1, size_t ldx1,
    T alpha, T beta, T *C, size_t ldc, const crp_spmm_wait *wait, const crp_spmm_put *put, const int n_elem_size, cudaStream_t s, const char *tname
)
{
    const int x0_rows = plan->x0_rows;
    const crp_rowgroup *rg = &plan->rg;
    const bool want_rg = (plan->variant == CRP_VARIANT_AUTO || plan->variant == CRP_VARIANT_ROWGROUP);
    const char *lr_tag = (plan->lr.nlong > 0) ? "+longrow" : "";
    const bool want_pn = (plan->variant == CRP_VARIANT_AUTO || plan->variant == CRP_VARIANT_PANEL);
    if (want_pn && rg->R > 1 && rg->ngroups > 0)
    {
        // the panel kernel waits for the neighbours itself (and for all of them before it ends), so the rest rows,
        // which may read received rows too, follow it on the stream
        bool rest_done = false;
        if (crp_launch_panel<T, VECN>(plan, val, n, X0, ldx0, X1, ldx1, alpha, beta, C, ldc, wait, put, &rest_done, s))
        {
            const bool rest_sep = rg->nrest > 0 && !rest_done;
            if (rest_sep) rowsplit_balanced<T, VECN>(plan, true, rg->nrest, rg->d_rest, val, n, X0, ldx0, X1, ldx1, alpha, beta, C, ldc, s);
            snprintf(plan->kernel_name, sizeof(plan->kernel_name), "spmm_panel_%s_R%d_K%d%s%s", tname, rg->R, plan->pn.K, rest_sep ? "+rowsplit" : "", rest_sep ? lr_tag : "");
            plan->last_kernel = plan->kernel_name;
            return;
        }
    }
    // the other kernels neither put nor wait themselves: separate launches first
    if (put != NULL && put->nflag > 0)
        crp_cuda_put_rows_signal((size_t) n_elem_size, put->nrow, (int) (put->row_bytes / (size_t) n_elem_size), X0, (int) ldx0, put->ridx, put->dst_rows, put->flag_ptrs, put->nflag, put->epoch, put->counter, put->dst_off, (void *) s);
    wait_first(wait, s);
    // nnz-balanced kernel: forced, or chosen for matrices without row groups whose longest row is far above the average
    // (power-law graphs: one-row-per-warp leaves most warps idle behind the hubs)
    const bool skewed = (rg->R <= 1) && plan->avg_row_nnz > 0.0 && (double) plan->max_row_nnz > 16.0 * plan->avg_row_nnz + 64.0;
    if (plan->variant == CRP_VARIANT_MERGEPATH || (plan->variant == CRP_VARIANT_AUTO && skewed))
    {
        if (plan->mp.nchunks == 0 && !plan->mp_tried) { crp_mergepath_build(plan, plan->h_rowptr); plan->mp_tried = 1; }
        if (crp_launch_mergepath<T, VECN>(plan, val, n, X0, ldx0, X1, ldx1, alpha, beta, C, ldc, s))
        {
            snprintf(plan->kernel_name, sizeof(plan->kernel_name), "spmm_mergepath_%s%s", tname, plan->mp.nlong > 0 ? "+fixup" : "");
            plan->last_kernel = plan->kernel_name;
            return;
        }
    }
    if (want_rg && rg->R > 1 && rg->ngroups > 0 && rg->exact)
    {
        const uintptr_t ptrs = (uintptr_t) X0 | (uintptr_t) X1 | (uintptr_t) C;
        const bool vec_ok = (n % VECN == 0) && (ldx0 % VECN == 0) && (X1 == NULL || ldx1 % VECN == 0) && (ldc % VECN == 0) && ((ptrs & 15) == 0);
        if (vec_ok) crp_launch_rowgroup<T, VECN>(rg, bval, n / VECN, X0, ldx0, x0_rows, X1, ldx1, alpha, beta, C, ldc, s);
        else        crp_launch_rowgroup<T, 1>(rg, bval, n, X0, ldx0, x0_rows, X1, ldx1, alpha, beta, C, ldc, s);
        if (rg->nrest > 0) rowsplit_balanced<T, VECN>(plan, true, rg->nrest, rg->d_rest, val, n, X0, ldx0, X1, ldx1, alpha, beta, C, ldc, s);
        snprintf(plan->kernel_name, sizeof(plan->kernel_name), "spmm_rowgroup_%s_R%d%s%s", tname, rg->R, rg->nrest > 0 ? "+rowsplit" : "", rg->nrest > 0 ? lr_tag : "");
    } else {
        // forcing "rowsplit" on a plan that has a row-group form: the segment lists were built for the rest rows only
        const bool use_lr = (rg->R <= 1) && plan->variant != CRP_VARIANT_ROWSPLIT && plan->variant != CRP_VARIANT_MERGEPATH;
        rowsplit_balanced<T, VECN>(plan, use_lr, plan->m, NULL, val, n, X0, ldx0, X1, ldx1, alpha, beta, C, ldc, s);
        snprintf(plan->kernel_name, sizeof(plan->kernel_name), "spmm_rowsplit_%s%s", tname, use_lr ? lr_tag : "");
    }
    plan->last_kernel = plan->kernel_name;
}

extern "C" void crp_cuda_spmm_exec(
    crp_spmm_plan *plan, const int n, const int elem_size, const double alpha,
    const void *X0, const int ldx0, const void *X1, const int ldx1,
    const double beta, void *C, const int ldc, void *stream
)
{
    crp_cuda_spmm_exec_wait(plan, n, elem_size, alpha, X0, ldx0, X1, ldx1, beta, C, ldc, NULL, NULL, 0, 0, 0.0, NULL, stream);
}

static void spmm_exec_common(
    crp_spmm_plan *plan, const int n, const int elem_size, const double alpha, const void *X0, const int ldx0, const void *X1, const int ldx1,
    const double beta, void *C, const int ldc, const crp_spmm_wait *wait, const crp_spmm_put *put, cudaStream_t s
);

extern "C" void crp_cuda_spmm_exec_exchange(
    crp_spmm_plan *plan, const int n, const int elem_size, const double alpha,
    const void *X0, const int ldx0, const void *X1, const int ldx1, const double beta, void *C, const int ldc,
    const crp_exchange *xc, void *stream
)
{
    if (plan == NULL) { fprintf(stderr, "[FATAL] crp_cuda_spmm_exec: NULL plan\n"); abort(); }
    crp_spmm_wait wt, *wait = NULL;
    crp_spmm_put pt, *put = NULL;
    if (xc != NULL && xc->nwait > 0)
    {
        wt.flags = xc->flags_d;  wt.wait_idx = xc->wait_idx_d;  wt.nwait = xc->nwait;  wt.epoch = xc->epoch;
        wt.timeout_ns = (long long) (xc->timeout_s * 1e9);  wt.err = xc->err;
        wait = &wt;
    }
    if (xc != NULL && xc->nflag > 0)
    {
        pt.nrow = xc->n_send_rows;  pt.row_bytes = (size_t) elem_size * (size_t) n;  pt.ridx = xc->send_ridx_d;  pt.dst_rows = xc->dst_rows_d;
        pt.flag_ptrs = xc->flag_ptrs_d;  pt.nflag = xc->nflag;  pt.epoch = xc->epoch;  pt.counter = xc->done_counter_d;  pt.dst_off = xc->dst_off_bytes;
        put = &pt;
    }
    spmm_exec_common(plan, n, elem_size, alpha, X0, ldx0, X1, ldx1, beta, C, ldc, wait, put, as_stream(stream));
}

extern "C" void crp_cuda_spmm_set_wait_map(crp_spmm_plan *plan, const int nslot, const int *recv_off)
{
    if (plan != NULL) crp_panel_set_wait_map(plan, nslot, recv_off);
}

extern "C" void crp_cuda_spmm_exec_wait(
    crp_spmm_plan *plan, const int n, const int elem_size, const double alpha,
    const void *X0, const int ldx0, const void *X1, const int ldx1,
    const double beta, void *C, const int ldc,
    const unsigned int *flags_d, const int *wait_idx_d, const int nwait, const unsigned int epoch, const double timeout_s, int *err,
    void *stream
)
{
    if (plan == NULL) { fprintf(stderr, "[FATAL] crp_cuda_spmm_exec: NULL plan\n"); abort(); }
    crp_spmm_wait wt, *wait = NULL;
    if (nwait > 0)
    {
        wt.flags = flags_d;  wt.wait_idx = wait_idx_d;  wt.nwait = nwait;  wt.epoch = epoch;
        wt.timeout_ns = (long long) (timeout_s * 1e9);  wt.err = err;
        wait = &wt;
    }
    spmm_exec_common(plan, n, elem_size, alpha, X0, ldx0, X1, ldx1, beta, C, ldc, wait, NULL, as_stream(stream));
}

static void spmm_exec_common(
    crp_spmm_plan *plan, const int n, const int elem_size, const double alpha, const void *X0, const int ldx0, const void *X1, const int ldx1,
    const double beta, void *C, const int ldc, const crp_spmm_wait *wait, const crp_spmm_put *put, cudaStream_t s
)
{
    if (plan->m == 0 || n <= 0)
    {
        if (put != NULL && put->nflag > 0 && n > 0)
            crp_cuda_put_rows_signal((size_t) elem_size, put->nrow, n, X0, ldx0, put->ridx, put->dst_rows, put->flag_ptrs, put->nflag, put->epoch, put->counter, put->dst_off, (void *) s);
        wait_first(wait, s);
        return;
    }
    if (elem_size != 4 && elem_size != 8)
    {
        fprintf(stderr, "[FATAL] crp_cuda_spmm_exec: elem_size must be 4 or 8\n");
        abort();
    }
    if (elem_size == 4)
    {
        cast_to_f32(plan->d_val, &plan->d_val32, (size_t) plan->nnz, s);
        if (plan->rg.d_bval != NULL) cast_to_f32(plan->rg.d_bval, &plan->rg.d_bval32, (size_t) plan->rg.nblk * (size_t) plan->rg.R, s);
    }
    // Column passes sized for the L2 (choose_passes): the dense operands are cut into column blocks that are multiplied one
    // after the other on the stream.  The first pass carries the exchange (put / wait); when it has finished every neighbour's
    // rows have arrived (the kernels wait for all flags before they end), so the later passes need neither.
    int P = choose_passes(plan, n, elem_size);
    int ns = n;
    if (P > 1)
    {
        ns = ((n + P - 1) / P + 63) / 64 * 64;
        const uintptr_t al = (uintptr_t) X0 | (uintptr_t) X1 | (uintptr_t) C | ((uintptr_t) ldx0 * elem_size) | ((uintptr_t) ldx1 * elem_size) | ((uintptr_t) ldc * elem_size);
        if (ns >= n || (al & 15) != 0) { P = 1; ns = n; }
    }
    int npass = 0;
    for (int c0 = 0; c0 < n; c0 += ns, npass++)
    {
        const int nn = (n - c0 < ns) ? n - c0 : ns;
        const size_t off = (size_t) c0 * (size_t) elem_size;
        const char *x0 = (const char *) X0 + off, *x1 = (X1 != NULL) ? (const char *) X1 + off : NULL;
        char *c = (char *) C + off;
        const crp_spmm_wait *w = (c0 == 0) ? wait : NULL;
        const crp_spmm_put *pt = (c0 == 0) ? put : NULL;
        if (elem_size == 8)
            spmm_dispatch<double, 2>(plan, plan->d_val, plan->rg.d_bval, nn, (const double *) x0, (size_t) ldx0, (const double *) x1, (size_t) ldx1,
                                     alpha, beta, (double *) c, (size_t) ldc, w, pt, 8, s, "f64");
        else
            spmm_dispatch<float, 4>(plan, plan->d_val32, plan->rg.d_bval32, nn, (const float *) x0, (size_t) ldx0, (const float *) x1, (size_t) ldx1,
                                    (float) alpha, (float) beta, (float *) c, (size_t) ldc, w, pt, 4, s, "f32");
    }
    plan->last_passes = npass;
    if (npass > 1)
    {
        const size_t len = strlen(plan->kernel_name);
        snprintf(plan->kernel_name + len, sizeof(plan->kernel_name) - len, "_x%dpass", npass);
    }
}

extern "C" void crp_cuda_spmm_plan_info(const crp_spmm_plan *plan, long long out[12])
{
    for (int i = 0; i < 12; i++) out[i] = 0;
    if (plan == NULL) return;
    out[0] = plan->rg.R > 1 ? plan->rg.R : 1;
    out[1] = plan->rg.ngroups;  out[2] = plan->rg.nblk;  out[3] = plan->rg.R > 1 ? plan->rg.nrest : plan->m;
    out[4] = plan->rg.R > 1 ? plan->rg.rest_nnz : plan->nnz;
    out[5] = plan->pn.ntiles;  out[6] = plan->pn.nchunks;  out[7] = plan->pn.union_rows;
    out[8] = plan->rg.exact;  out[9] = plan->lr.nlong;  out[10] = plan->nnz;  out[11] = plan->mp.nchunks;
}

extern "C" const char *crp_cuda_spmm_last_kernel(const crp_spmm_plan *plan) { return plan ? plan->last_kernel : "none"; }

extern "C" void crp_cuda_spmm_set_variant(crp_spmm_plan *plan, const char *name)
{
    if (plan == NULL || name == NULL) return;
    if (strcmp(name, "rowsplit") == 0) plan->variant = CRP_VARIANT_ROWSPLIT;
    else if (strcmp(name, "rowgroup") == 0) plan->variant = CRP_VARIANT_ROWGROUP;
    else if (strcmp(name, "mergepath") == 0) plan->variant = CRP_VARIANT_MERGEPATH;
    else if (strcmp(name, "panel") == 0) plan->variant = CRP_VARIANT_PANEL;
    else plan->variant = CRP_VARIANT_AUTO;
}

extern "C" void crp_cuda_csr_spmm_host(
    const int m, const int n, const int k, const double alpha,
    const int A_nnz, const int *A_rowptr_h, const int *A_colidx_h, const double *A_val_h,
    const double *B_h, const int ldB, const double beta, double *C_h, const int ldC
)
{
    (void) A_nnz;
    crp_spmm_plan *plan = crp_cuda_spmm_plan_create(m, k, k, A_rowptr_h, A_colidx_h, A_val_h, n);
    double *B_d = NULL, *C_d = NULL;
    const size_t row_bytes = sizeof(double) * (size_t) n;
    CRP_CUDA_CHECK(cudaMalloc((void **) &B_d, row_bytes * (size_t) (k > 0 ? k : 1)));
    CRP_CUDA_CHECK(cudaMalloc((void **) &C_d, row_bytes * (size_t) (m > 0 ? m : 1)));
    if (k > 0 && n > 0) CRP_CUDA_CHECK(cudaMemcpy2D(B_d, row_bytes, B_h, sizeof(double) * (size_t) ldB, row_bytes, (size_t) k, cudaMemcpyHostToDevice));
    if (beta != 0.0 && m > 0 && n > 0) CRP_CUDA_CHECK(cudaMemcpy2D(C_d, row_bytes, C_h, sizeof(double) * (size_t) ldC, row_bytes, (size_t) m, cudaMemcpyHostToDevice));
    crp_cuda_spmm_exec(plan, n, 8, alpha, B_d, n, NULL, 0, beta, C_d, n, NULL);
    CRP_CUDA_CHECK(cudaStreamSynchronize(0));
    if (m > 0 && n > 0) CRP_CUDA_CHECK(cudaMemcpy2D(C_h, sizeof(double) * (size_t) ldC, C_d, row_bytes, row_bytes, (size_t) m, cudaMemcpyDeviceToHost));
    CRP_CUDA_CHECK(cudaFree(B_d));
    CRP_CUDA_CHECK(cudaFree(C_d));
    crp_cuda_spmm_plan_destroy(plan);
}
