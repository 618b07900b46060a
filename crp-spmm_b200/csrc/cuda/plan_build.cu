// plan_build.cu - device-side construction of the O(nnz) parts of the host plans.
//
//   crp_cuda_plan_needed_rows   the sweeps of rp_spmm_init over the local A (reference src/rowpara_spmm.c:46-112):
//                               column range, "is this B row needed" flags, the compact position of every needed
//                               row (a prefix sum over the flags), the re-indexed column array.  What comes back
//                               is small - lo / hi, the sorted list of needed rows - plus the re-indexed columns.
//   crp_cuda_part_comm_size     csr_mat_row_part_comm_size (reference src/spmat_part.c:38-64): for every row block
//                               the number of distinct columns outside the block's own column range - one bitmap
//                               of ncol bits per block filled with atomicOr, counted with popc.  The matrix stays
//                               on the device between the calls the 2-D cost model makes (one per candidate grid).
//
// Both produce integers only and must equal the host code bit for bit (tests/test_gpu_plan.py).  The host code
// (csrc/host/rowpara_spmm.c, spmat_part.c) decides when to call them: a device is bound, the matrix is large
// enough to pay for the copy (CRP_SPMM_GPU_PLAN_MIN_NNZ), CRP_SPMM_GPU_PLAN != 0.
#include <climits>
#include <cstring>

#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "crp_cuda_internal.cuh"

namespace {

__global__ void __launch_bounds__(256) flag_cols_kernel(const int *__restrict__ col, const long long nnz, int *__restrict__ flag, int *__restrict__ lohi)
{
    int lo = INT_MAX, hi = 0;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (long long) gridDim.x * blockDim.x)
    {
        const int c = __ldg(col + i);
        flag[c] = 1;                                // benign race: every writer stores the same word (32-bit flags: the scan sums them as ints)
        lo = min(lo, c);
        hi = max(hi, c);
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(lohi, lo); atomicMax(lohi + 1, hi); }
}

__global__ void __launch_bounds__(256) reindex_kernel(const int *__restrict__ col, const long long nnz, const int *__restrict__ pos, const int lo, const int reidx, int *__restrict__ out)
{
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (long long) gridDim.x * blockDim.x)
    {
        const int c = __ldg(col + i);
        out[i] = reidx ? __ldg(pos + c) : c - lo;
    }
}

// one warp per row: block of the row by binary search over the block boundaries, then one atomicOr per nonzero
__global__ void __launch_bounds__(256) block_bitmap_kernel(
    const int *__restrict__ rowptr, const int *__restrict__ col, const int row_lo, const int row_hi, const int *__restrict__ rblk, const int nblk,
    unsigned int *__restrict__ bitmap, const size_t words
)
{
    const int warp = (int) (((size_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    const int nwarp = (int) (((size_t) gridDim.x * blockDim.x) >> 5);
    for (int r = row_lo + warp; r < row_hi; r += nwarp)
    {
        int a = 0, b = nblk;                        // largest blk with rblk[blk] <= r
        while (b - a > 1) { const int mid = (a + b) >> 1; if (__ldg(rblk + mid) <= r) a = mid; else b = mid; }
        unsigned int *bm = bitmap + (size_t) a * words;
        for (int p = __ldg(rowptr + r) + lane; p < __ldg(rowptr + r + 1); p += 32)
        {
            const int c = __ldg(col + p);
            atomicOr(bm + (c >> 5), 1u << (c & 31));
        }
    }
}

__global__ void __launch_bounds__(256) bitmap_count_kernel(
    const unsigned int *__restrict__ bitmap, const size_t words, const int ncol, const int *__restrict__ xdispl, int *__restrict__ counts
)
{
    const int blk = blockIdx.y;
    const int own_lo = xdispl[blk], own_hi = xdispl[blk + 1];
    const unsigned int *bm = bitmap + (size_t) blk * words;
    int cnt = 0;
    for (size_t w = (size_t) blockIdx.x * blockDim.x + threadIdx.x; w < words; w += (size_t) gridDim.x * blockDim.x)
    {
        unsigned int v = bm[w];
        if (v == 0u) continue;
        const long long c0 = (long long) w * 32;
        // clear the bits of the block's own columns [own_lo, own_hi)
        if (c0 + 32 > own_lo && c0 < own_hi)
        {
            const int a = (int) max((long long) own_lo - c0, 0ll), b = (int) min((long long) own_hi - c0, 32ll);
            const unsigned int m = ((b >= 32) ? 0xffffffffu : ((1u << b) - 1u)) & ~((1u << a) - 1u);
            v &= ~m;
        }
        (void) ncol;
        cnt += __popc(v);
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(counts + blk, cnt);
}

int grid_for(const long long n)
{
    long long g = (n + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    return (int) (g < 1 ? 1 : g);
}

// the global CSR pattern kept on the device between the cost model's calls
struct pattern_cache
{
    const int *h_rowptr, *h_colidx;
    int nrow;
    long long nnz;
    int *d_rowptr, *d_colidx;
} g_pat = { NULL, NULL, 0, 0, NULL, NULL };

}   // namespace

extern "C" void crp_cuda_part_cache_release(void)
{
    if (g_pat.d_rowptr) CRP_CUDA_CHECK(cudaFree(g_pat.d_rowptr));
    if (g_pat.d_colidx) CRP_CUDA_CHECK(cudaFree(g_pat.d_colidx));
    memset(&g_pat, 0, sizeof(g_pat));
}

extern "C" int crp_cuda_part_comm_size(
    const int nrow, const int ncol, const int *row_ptr_h, const int *col_idx_h,
    const int nblk, const int *rblk_ptr_h, const int *x_displs_h, int *comm_sizes_h, int *total_size
)
{
    if (nblk <= 0 || nrow <= 0 || ncol <= 0) return 0;
    const long long nnz = (long long) row_ptr_h[nrow] - row_ptr_h[0];
    const size_t words = ((size_t) ncol + 31) / 32;
    if ((double) nblk * (double) words * 4.0 > 8.0e9) return 0;           // bitmaps would not be worth it: host path
    if (g_pat.h_rowptr != row_ptr_h || g_pat.h_colidx != col_idx_h || g_pat.nrow != nrow || g_pat.nnz != nnz)
    {
        crp_cuda_part_cache_release();
        CRP_CUDA_CHECK(cudaMalloc((void **) &g_pat.d_rowptr, sizeof(int) * ((size_t) nrow + 1)));
        CRP_CUDA_CHECK(cudaMalloc((void **) &g_pat.d_colidx, sizeof(int) * (size_t) (nnz > 0 ? nnz : 1)));
        CRP_CUDA_CHECK(cudaMemcpy(g_pat.d_rowptr, row_ptr_h, sizeof(int) * ((size_t) nrow + 1), cudaMemcpyHostToDevice));
        if (nnz > 0) CRP_CUDA_CHECK(cudaMemcpy(g_pat.d_colidx, col_idx_h + row_ptr_h[0], sizeof(int) * (size_t) nnz, cudaMemcpyHostToDevice));
        g_pat.h_rowptr = row_ptr_h;  g_pat.h_colidx = col_idx_h;  g_pat.nrow = nrow;  g_pat.nnz = nnz;
    }
    unsigned int *d_bitmap = NULL;
    int *d_small = NULL;                // rblk (nblk + 1) | xdispl (nblk + 1) | counts (nblk)
    CRP_CUDA_CHECK(cudaMalloc((void **) &d_bitmap, sizeof(unsigned int) * words * (size_t) nblk));
    CRP_CUDA_CHECK(cudaMalloc((void **) &d_small, sizeof(int) * (3 * (size_t) nblk + 2)));
    CRP_CUDA_CHECK(cudaMemsetAsync(d_bitmap, 0, sizeof(unsigned int) * words * (size_t) nblk, 0));
    CRP_CUDA_CHECK(cudaMemcpyAsync(d_small, rblk_ptr_h, sizeof(int) * ((size_t) nblk + 1), cudaMemcpyHostToDevice, 0));
    CRP_CUDA_CHECK(cudaMemcpyAsync(d_small + nblk + 1, x_displs_h, sizeof(int) * ((size_t) nblk + 1), cudaMemcpyHostToDevice, 0));
    CRP_CUDA_CHECK(cudaMemsetAsync(d_small + 2 * nblk + 2, 0, sizeof(int) * (size_t) nblk, 0));
    const int row_lo = rblk_ptr_h[0], row_hi = rblk_ptr_h[nblk];
    if (row_hi > row_lo)
    {
        // the cached column array starts at nonzero row_ptr_h[0]: shift the pointer so that rowptr values index it directly
        block_bitmap_kernel<<<grid_for(32ll * (row_hi - row_lo)), 256>>>(g_pat.d_rowptr, g_pat.d_colidx - row_ptr_h[0], row_lo, row_hi, d_small, nblk, d_bitmap, words);
        CRP_LAUNCH_CHECK();
    }
    dim3 grid((unsigned) grid_for((long long) words), (unsigned) nblk);
    bitmap_count_kernel<<<grid, 256>>>(d_bitmap, words, ncol, d_small + nblk + 1, d_small + 2 * nblk + 2);
    CRP_LAUNCH_CHECK();
    CRP_CUDA_CHECK(cudaMemcpy(comm_sizes_h, d_small + 2 * nblk + 2, sizeof(int) * (size_t) nblk, cudaMemcpyDeviceToHost));
    CRP_CUDA_CHECK(cudaFree(d_bitmap));
    CRP_CUDA_CHECK(cudaFree(d_small));
    int sum = 0;
    for (int b = 0; b < nblk; b++) sum += comm_sizes_h[b];
    *total_size = sum;
    return 1;
}

extern "C" int crp_cuda_plan_needed_rows(
    const int *colidx_h, const long long nnz, const int glb_k, const int reidx,
    int *colidx_out_h, int *lo_out, int *hi_out, int *n_needed_out, int **needed_rows_out
)
{
    if (nnz <= 0 || glb_k <= 0) return 0;
    int *d_col = NULL, *d_out = NULL, *d_pos = NULL, *d_rows = NULL, *d_meta = NULL;
    int *d_flag = NULL;
    CRP_CUDA_CHECK(cudaMalloc((void **) &d_col, sizeof(int) * (size_t) nnz));
    CRP_CUDA_CHECK(cudaMalloc((void **) &d_out, sizeof(int) * (size_t) nnz));
    CRP_CUDA_CHECK(cudaMalloc((void **) &d_flag, sizeof(int) * (size_t) glb_k));
    CRP_CUDA_CHECK(cudaMalloc((void **) &d_pos, sizeof(int) * (size_t) glb_k));
    CRP_CUDA_CHECK(cudaMalloc((void **) &d_rows, sizeof(int) * (size_t) glb_k));
    CRP_CUDA_CHECK(cudaMalloc((void **) &d_meta, sizeof(int) * 4));
    CRP_CUDA_CHECK(cudaMemcpy(d_col, colidx_h, sizeof(int) * (size_t) nnz, cudaMemcpyHostToDevice));
    CRP_CUDA_CHECK(cudaMemset(d_flag, 0, sizeof(int) * (size_t) glb_k));
    const int init[4] = { INT_MAX, 0, 0, 0 };
    CRP_CUDA_CHECK(cudaMemcpy(d_meta, init, sizeof(init), cudaMemcpyHostToDevice));
    flag_cols_kernel<<<grid_for(nnz), 256>>>(d_col, nnz, d_flag, d_meta);
    CRP_LAUNCH_CHECK();
    // position of every needed row among the needed rows (exclusive prefix sum of the flags) and their sorted list
    void *d_tmp = NULL;
    size_t tmp_scan = 0, tmp_sel = 0;
    thrust::counting_iterator<int> ids(0);
    CRP_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(NULL, tmp_scan, d_flag, d_pos, glb_k));
    CRP_CUDA_CHECK(cub::DeviceSelect::Flagged(NULL, tmp_sel, ids, d_flag, d_rows, d_meta + 2, glb_k));
    const size_t tmp_bytes = tmp_scan > tmp_sel ? tmp_scan : tmp_sel;
    CRP_CUDA_CHECK(cudaMalloc(&d_tmp, tmp_bytes > 0 ? tmp_bytes : 16));
    size_t tb = tmp_bytes;
    CRP_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_flag, d_pos, glb_k));
    tb = tmp_bytes;
    CRP_CUDA_CHECK(cub::DeviceSelect::Flagged(d_tmp, tb, ids, d_flag, d_rows, d_meta + 2, glb_k));
    crp_count_launch();
    int meta[4];
    CRP_CUDA_CHECK(cudaMemcpy(meta, d_meta, sizeof(meta), cudaMemcpyDeviceToHost));
    const int lo = meta[0], hi = meta[1], n_needed = meta[2];
    reindex_kernel<<<grid_for(nnz), 256>>>(d_col, nnz, d_pos, lo, reidx, d_out);
    CRP_LAUNCH_CHECK();
    CRP_CUDA_CHECK(cudaMemcpy(colidx_out_h, d_out, sizeof(int) * (size_t) nnz, cudaMemcpyDeviceToHost));
    int *rows = (int *) malloc(sizeof(int) * (size_t) (n_needed > 0 ? n_needed : 1));
    if (n_needed > 0) CRP_CUDA_CHECK(cudaMemcpy(rows, d_rows, sizeof(int) * (size_t) n_needed, cudaMemcpyDeviceToHost));
    *lo_out = lo;  *hi_out = hi;  *n_needed_out = n_needed;  *needed_rows_out = rows;
    CRP_CUDA_CHECK(cudaFree(d_tmp));
    CRP_CUDA_CHECK(cudaFree(d_col));  CRP_CUDA_CHECK(cudaFree(d_out));  CRP_CUDA_CHECK(cudaFree(d_flag));
    CRP_CUDA_CHECK(cudaFree(d_pos));  CRP_CUDA_CHECK(cudaFree(d_rows));  CRP_CUDA_CHECK(cudaFree(d_meta));
    return 1;
}
