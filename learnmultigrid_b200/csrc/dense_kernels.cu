// dense_kernels.cu -- coarsest-level dense direct solve (replaces SuperLU spsolve, Multigrid.py:106).
// The inverse is formed once at setup by Gauss-Jordan elimination with partial pivoting (cooperative
// multi-CTA kernel, one grid barrier per pivot); the per-cycle solve is one streaming GEMV.
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace mgb {

struct PivotCand { double absval; int32_t row; int32_t pad; };

// W = [A | I] is n x 2n row-major.  Rows are never swapped physically: piv_of_col[k] = the row chosen as
// pivot of column k.  Each CTA owns rows r = blockIdx.x, blockIdx.x + gridDim.x, ...
__global__ void __launch_bounds__(256)
gauss_jordan_kernel(int64_t n, double *W, int32_t *piv_of_col, int32_t *pivoted, PivotCand *cand,
                    int32_t *singular) {
    cg::grid_group grid = cg::this_grid();
    const int64_t w = 2 * n;
    __shared__ double s_abs[256];
    __shared__ int32_t s_row[256];
    __shared__ int32_t s_piv;

    // candidate search for column `col` among this CTA's unpivoted rows
    auto local_candidate = [&](int64_t col, PivotCand *out) {
        double best = -1.0;
        int32_t brow = -1;
        for (int64_t r = blockIdx.x + (int64_t)threadIdx.x * gridDim.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
            if (!pivoted[r]) {
                const double a = fabs(W[r * w + col]);
                if (a > best || (a == best && (int32_t)r < brow)) { best = a; brow = (int32_t)r; }
            }
        }
        s_abs[threadIdx.x] = best;
        s_row[threadIdx.x] = brow;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (threadIdx.x < o) {
                const double a = s_abs[threadIdx.x + o];
                const int32_t rr = s_row[threadIdx.x + o];
                if (rr >= 0 && (a > s_abs[threadIdx.x] || (a == s_abs[threadIdx.x] && (s_row[threadIdx.x] < 0 || rr < s_row[threadIdx.x])))) {
                    s_abs[threadIdx.x] = a;
                    s_row[threadIdx.x] = rr;
                }
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) { out[blockIdx.x].absval = s_abs[0]; out[blockIdx.x].row = s_row[0]; }
    };

    // candidates are double-buffered (cand + (k&1)*4096) so that one grid barrier per pivot suffices
    local_candidate(0, cand);
    grid.sync();
    for (int64_t k = 0; k < n; ++k) {
        // every CTA reduces the candidates redundantly (deterministic: largest |a|, ties -> smallest row), all of its
        // threads taking part: one thread walking the ~1200 candidates through L2 was most of a pivot step (88 us per
        // pivot at n = 2064, 0.18 s of the 0.25 s coarsest factorisation of the 8193^2 hierarchy)
        {
            double best = -1.0;
            int32_t brow = -1;
            const PivotCand *cur = cand + (k & 1) * 4096;
            for (unsigned bkt = threadIdx.x; bkt < gridDim.x; bkt += blockDim.x) {
                const double a = __ldcg(&cur[bkt].absval);
                const int32_t rr = __ldcg(&cur[bkt].row);
                if (rr >= 0 && (a > best || (a == best && rr < brow))) { best = a; brow = rr; }
            }
            s_abs[threadIdx.x] = best;
            s_row[threadIdx.x] = brow;
            __syncthreads();
            for (int o = 128; o > 0; o >>= 1) {
                if (threadIdx.x < o) {
                    const double a = s_abs[threadIdx.x + o];
                    const int32_t rr = s_row[threadIdx.x + o];
                    if (rr >= 0 && (a > s_abs[threadIdx.x] || (a == s_abs[threadIdx.x] && (s_row[threadIdx.x] < 0 || rr < s_row[threadIdx.x])))) {
                        s_abs[threadIdx.x] = a;
                        s_row[threadIdx.x] = rr;
                    }
                }
                __syncthreads();
            }
            if (threadIdx.x == 0) {
                int32_t pr = s_row[0];
                if (pr < 0 || s_abs[0] == 0.0) { *singular = 1; pr = -1; }
                s_piv = pr;
            }
        }
        __syncthreads();
        const int32_t p = s_piv;
        if (p < 0) return;   // uniform across the grid: every CTA sees the same candidates
        if (blockIdx.x == 0 && threadIdx.x == 0) { piv_of_col[k] = p; }
        if (blockIdx.x == (unsigned)(p % gridDim.x) && threadIdx.x == 0) pivoted[p] = 1;
        const double *prow = W + (int64_t)p * w;
        const double pinv = 1.0 / __ldcg(prow + k);
        // eliminate column k from this CTA's rows (all rows except p, pivoted or not: Gauss-Jordan)
        for (int64_t r = blockIdx.x; r < n; r += gridDim.x) {
            if (r == p) continue;
            double *row = W + r * w;
            const double f = row[k] * pinv;
            if (f != 0.0) {
                // columns <= k of the left block are already (or become) zero except pivots; update k..2n
                for (int64_t c = k + 1 + threadIdx.x; c < w; c += blockDim.x) row[c] -= f * __ldcg(prow + c);
            }
            __syncthreads();
            if (threadIdx.x == 0) row[k] = 0.0;
        }
        __syncthreads();
        __threadfence();
        if (k + 1 < n) local_candidate(k + 1, cand + ((k + 1) & 1) * 4096);
        grid.sync();
    }
}

// Ainv[k, :] = W[piv_of_col[k], n:2n] / W[piv_of_col[k], k]
__global__ void __launch_bounds__(256)
extract_inverse_kernel(int64_t n, const double *__restrict__ W, const int32_t *__restrict__ piv_of_col,
                       double *__restrict__ Ainv) {
    const int64_t k = blockIdx.x;
    const int64_t p = piv_of_col[k];
    const double d = W[p * 2 * n + k];
    for (int64_t c = threadIdx.x; c < n; c += blockDim.x) Ainv[k * n + c] = W[p * 2 * n + n + c] / d;
}

__global__ void __launch_bounds__(256)
build_augmented_kernel(int64_t n, const double *__restrict__ A, double *__restrict__ W) {
    const int64_t r = blockIdx.x;
    for (int64_t c = threadIdx.x; c < 2 * n; c += blockDim.x)
        W[r * 2 * n + c] = (c < n) ? A[r * n + c] : ((c - n == r) ? 1.0 : 0.0);
}

// y = M x for the rows [row0, row0+nrows) of a row-major matrix with m columns.  T threads share a row (kBlock/T rows
// per CTA), each keeps four independent loads in flight; partial sums are combined in a fixed order (four
// accumulators, warp tree, then the row's warps in order), so the result depends on T only.  The output index of
// row i is i, or ((i / bm) << shift) * bm + i % bm when bm > 0 (block-strided vectors of the BCR tail solve).
template <int T>
__global__ void __launch_bounds__(kBlock)
gemv_rows_kernel(int64_t row0, int64_t nrows, int64_t m, const double *__restrict__ M, const double *__restrict__ x,
                 double *__restrict__ y, int64_t bm, int shift) {
    pdl_prologue();
    constexpr int RPC = kBlock / T;
    __shared__ double part[kBlock / 32];
    const int sub = threadIdx.x / T, t = threadIdx.x % T;
    const int64_t r = (int64_t)blockIdx.x * RPC + sub;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    if (r < nrows) {
        const double *mr = M + (row0 + r) * m;
        int64_t c = t;
        for (; c + 3 * T < m; c += 4 * T) {
            const double v0 = __ldcs(mr + c), v1 = __ldcs(mr + c + T), v2 = __ldcs(mr + c + 2 * T),
                         v3 = __ldcs(mr + c + 3 * T);
            a0 += v0 * x[c];
            a1 += v1 * x[c + T];
            a2 += v2 * x[c + 2 * T];
            a3 += v3 * x[c + 3 * T];
        }
        for (; c < m; c += T) a0 += __ldcs(mr + c) * x[c];
    }
    double s = (a0 + a1) + (a2 + a3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (T > 32) {
        const int w = threadIdx.x >> 5;
        if ((threadIdx.x & 31) == 0) part[w] = s;
        __syncthreads();
        if (t == 0) {
            s = 0.0;
            for (int k = 0; k < T / 32; ++k) s += part[sub * (T / 32) + k];
        }
    }
    if (t == 0 && r < nrows) {
        const int64_t i = row0 + r;
        y[bm > 0 ? ((i / bm) << shift) * bm + i % bm : i] = s;
    }
}

__global__ void __launch_bounds__(kBlock)
csr_to_dense_kernel(int64_t n, const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                    const double *__restrict__ values, double *__restrict__ D) {
    const int64_t row = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (row >= n) return;
    for (int32_t p = indptr[row]; p < indptr[row + 1]; ++p) D[row * n + indices[p]] += values[p];
}

// rows [row0,row0+nrows) of y = M x; `total_rows` (the whole matrix) picks the threads per row, so that a row block
// computed on its own (split solve) gets the same bits as in the full product
int gemv_rows(int64_t total_rows, int64_t row0, int64_t nrows, int64_t m, const double *M, const double *x, double *y,
              int64_t bm, int shift, cudaStream_t st) {
    if (nrows <= 0) return MG_OK;
    const int64_t want = (int64_t)sm_count() * 2048;          // threads that fill the device
    const int T = (total_rows * 32 >= want) ? 32 : (total_rows * 128 >= want) ? 128 : 256;
    const int64_t rpc = kBlock / T;
    const unsigned grid = (unsigned)((nrows + rpc - 1) / rpc);
    if (T == 32) launch_k(gemv_rows_kernel<32>, (unsigned)(grid), (unsigned)kBlock, st, row0, nrows, m, M, x, y, bm, shift);
    else if (T == 128) launch_k(gemv_rows_kernel<128>, (unsigned)(grid), (unsigned)kBlock, st, row0, nrows, m, M, x, y, bm, shift);
    else launch_k(gemv_rows_kernel<256>, (unsigned)(grid), (unsigned)kBlock, st, row0, nrows, m, M, x, y, bm, shift);
    MG_CHECK_LAUNCH("gemv_rows");
    return MG_OK;
}

int dense_gemv(int64_t n, int64_t m, const double *M, const double *x, double *y, cudaStream_t st) {
    return gemv_rows(n, 0, n, m, M, x, y, 0, 0, st);
}

}  // namespace mgb

using namespace mgb;

extern "C" {

int64_t mg_dense_inverse_workspace(int64_t n) {
    // W (n x 2n doubles) + piv_of_col (n) + pivoted (n) int32 + candidates + flag
    return n * 2 * n * (int64_t)sizeof(double) + 2 * n * (int64_t)sizeof(int32_t) +
           2 * 4096 * (int64_t)sizeof(PivotCand) + 256;
}

int mg_dense_inverse(int64_t n, double *d_a, double *d_ainv, void *d_work, void *stream) {
    MG_REQUIRE(n > 0 && d_a && d_ainv && d_work, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    char *base = (char *)d_work;
    double *W = (double *)base;
    base += n * 2 * n * sizeof(double);
    PivotCand *cand = (PivotCand *)base;
    base += 2 * 4096 * sizeof(PivotCand);
    int32_t *piv = (int32_t *)base;
    base += n * sizeof(int32_t);
    int32_t *pivoted = (int32_t *)base;
    base += n * sizeof(int32_t);
    int32_t *singular = (int32_t *)base;
    MG_CHECK_CUDA(cudaMemsetAsync(pivoted, 0, n * sizeof(int32_t), st));
    MG_CHECK_CUDA(cudaMemsetAsync(singular, 0, sizeof(int32_t), st));
    build_augmented_kernel<<<(unsigned)n, 256, 0, st>>>(n, d_a, W);
    MG_CHECK_LAUNCH("build_augmented");
    int per_sm = 0;
    MG_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gauss_jordan_kernel, 256, 0));
    int64_t grid = (int64_t)per_sm * sm_count();
    if (grid > n) grid = n;
    if (grid > 4096) grid = 4096;
    if (grid < 1) grid = 1;
    void *args[] = {(void *)&n, (void *)&W, (void *)&piv, (void *)&pivoted, (void *)&cand, (void *)&singular};
    MG_CHECK_CUDA(cudaLaunchCooperativeKernel((void *)gauss_jordan_kernel, dim3((unsigned)grid), dim3(256), args, 0, st));
    ++g_launch_count;
    int32_t h_singular = 0;
    MG_CHECK_CUDA(cudaMemcpyAsync(&h_singular, singular, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    MG_CHECK_CUDA(cudaStreamSynchronize(st));
    if (h_singular) return set_error(MG_ERR_SINGULAR, "mg_dense_inverse", "zero pivot: coarsest operator is singular");
    extract_inverse_kernel<<<(unsigned)n, 256, 0, st>>>(n, W, piv, d_ainv);
    MG_CHECK_LAUNCH("extract_inverse");
    MG_CHECK_CUDA(cudaStreamSynchronize(st));
    return MG_OK;
}

int mg_dense_gemv(int64_t n, int64_t m, const double *d_m, const double *d_x, double *d_y, void *stream) {
    MG_REQUIRE(n >= 0 && m >= 0, "negative size");
    return dense_gemv(n, m, d_m, d_x, d_y, (cudaStream_t)stream);
}

int mg_csr_to_dense(int64_t n, const int32_t *d_indptr, const int32_t *d_indices, const double *d_values,
                    double *d_dense, void *stream) {
    MG_REQUIRE(n >= 0, "negative size");
    if (n == 0) return MG_OK;
    MG_CHECK_CUDA(cudaMemsetAsync(d_dense, 0, n * n * sizeof(double), (cudaStream_t)stream));
    csr_to_dense_kernel<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, (cudaStream_t)stream>>>(
        n, d_indptr, d_indices, d_values, d_dense);
    MG_CHECK_LAUNCH("csr_to_dense");
    return MG_OK;
}

}  // extern "C"
