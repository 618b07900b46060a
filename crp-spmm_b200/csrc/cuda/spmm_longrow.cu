// spmm_longrow.cu - nnz-balanced treatment of very long rows for the row-split kernel.
//
// A power-law matrix (RMAT scale 22: rows with 10^5 nonzeros next to rows with one) would leave one
// warp of the row-split kernel working for milliseconds after all others have finished.  Rows longer
// than CRP_LONG_ROW nonzeros are therefore cut at plan time into segments of CRP_LONG_SEG nonzeros;
// the segments run through the same row-split kernel as independent virtual rows (their (begin, end)
// pairs replace the row pointers) into a scratch matrix, and `longrow_reduce_kernel` adds the segments
// of each row in ascending order - the summation order is fixed, results are bit-reproducible, and no
// floating-point atomics are used.  Replaces, for such rows, the same
// mkl_sparse_d_mm call as the other kernels (reference src/rowpara_spmm.c:404-407).
#include <cstring>
#include <vector>

#include "crp_cuda_internal.cuh"

void crp_longrows_build(crp_spmm_plan *plan, const int *rowptr, const int *rows, const int nrows)
{
    crp_longrows *lr = &plan->lr;
    memset(lr, 0, sizeof(*lr));
    std::vector<int> shorts, seg_beg, seg_end, long_row, long_sptr;
    long_sptr.push_back(0);
    for (int i = 0; i < nrows; i++)
    {
        const int r = rows ? rows[i] : i;
        const int b = rowptr[r], e = rowptr[r + 1];
        if (e - b <= CRP_LONG_ROW) { shorts.push_back(r); continue; }
        long_row.push_back(r);
        for (int p = b; p < e; p += CRP_LONG_SEG)
        {
            seg_beg.push_back(p);
            seg_end.push_back(p + CRP_LONG_SEG < e ? p + CRP_LONG_SEG : e);
        }
        long_sptr.push_back((int) seg_beg.size());
    }
    if (long_row.empty()) return;
    auto upload = [](const std::vector<int> &v) -> int * {
        int *d = NULL;
        if (v.empty()) return d;
        CRP_CUDA_CHECK(cudaMalloc((void **) &d, sizeof(int) * v.size()));
        CRP_CUDA_CHECK(cudaMemcpy(d, v.data(), sizeof(int) * v.size(), cudaMemcpyHostToDevice));
        return d;
    };
    lr->nlong = (int) long_row.size();
    lr->nseg = (int) seg_beg.size();
    lr->nshort = (int) shorts.size();
    lr->d_short = upload(shorts);
    lr->d_seg_beg = upload(seg_beg);
    lr->d_seg_end = upload(seg_end);
    lr->d_long_row = upload(long_row);
    lr->d_long_sptr = upload(long_sptr);
}

void crp_longrows_destroy(crp_spmm_plan *plan)
{
    crp_longrows *lr = &plan->lr;
    if (lr->d_short) CRP_CUDA_CHECK(cudaFree(lr->d_short));
    if (lr->d_seg_beg) CRP_CUDA_CHECK(cudaFree(lr->d_seg_beg));
    if (lr->d_seg_end) CRP_CUDA_CHECK(cudaFree(lr->d_seg_end));
    if (lr->d_long_row) CRP_CUDA_CHECK(cudaFree(lr->d_long_row));
    if (lr->d_long_sptr) CRP_CUDA_CHECK(cudaFree(lr->d_long_sptr));
    if (lr->d_scratch) CRP_CUDA_CHECK(cudaFree(lr->d_scratch));
    memset(lr, 0, sizeof(*lr));
}

// C[row, :] = alpha * sum_{s in segments of row, ascending} scratch[s, :] + beta * C[row, :]
template <typename T>
__global__ void __launch_bounds__(256) longrow_reduce_kernel(
    const int nlong, const int *__restrict__ long_row, const int *__restrict__ long_sptr,
    const T *__restrict__ scratch, const int n, const T alpha, const T beta, T *__restrict__ C, const size_t ldc
)
{
    const int i = blockIdx.y;
    if (i >= nlong) return;
    const int row = long_row[i], s0 = long_sptr[i], s1 = long_sptr[i + 1];
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
    {
        T acc = (T) 0;
        for (int s = s0; s < s1; s++) acc += scratch[(size_t) s * n + j];
        T *c = C + (size_t) row * ldc + j;
        *c = (beta == (T) 0) ? alpha * acc : fma(alpha, acc, beta * *c);
    }
}

template <typename T>
void crp_launch_longrow_reduce(const crp_longrows *lr, const int n, T alpha, T beta, T *C, size_t ldc, cudaStream_t s)
{
    if (lr->nlong == 0) return;

    const int gx = (n + 255) / 256;
    for (int base = 0; base < lr->nlong; base += 65535)
    {
        const int cnt = (lr->nlong - base < 65535) ? lr->nlong - base : 65535;
        longrow_reduce_kernel<T><<<dim3((unsigned) gx, (unsigned) cnt), 256, 0, s>>>(
            cnt, lr->d_long_row + base, lr->d_long_sptr + base, (const T *) lr->d_scratch, n, alpha, beta, C, ldc);
        CRP_LAUNCH_CHECK();
    }
}

template void crp_launch_longrow_reduce<double>(const crp_longrows *, const int, double, double, double *, size_t, cudaStream_t);
template void crp_launch_longrow_reduce<float>(const crp_longrows *, const int, float, float, float *, size_t, cudaStream_t);
