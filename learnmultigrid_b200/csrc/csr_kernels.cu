// csr_kernels.cu -- CSR kernels in natural ordering: the generic entry points behind the reference's
// SciPy / PyAMG call sites (include/mgb200.h).  Thread-per-row with sequential, unfused accumulation so
// that every result is bit-identical to the CPU libraries.  The bandwidth-critical path uses SELL
// (sell_kernels.cu); these serve small levels, the parity modes and the public API.
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace mgb {

enum CsrMode { C_SPMV = 0, C_RESID = 1, C_JACOBI = 2, C_PROLONG = 3 };

template <int MODE>
__global__ void __launch_bounds__(kBlock)
csr_kernel(int64_t n, const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
           const double *__restrict__ values, const double *x, const double *__restrict__ b,
           const double *aux, double *y, double omega) {
    const int64_t row = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (row >= n) return;
    double sum = 0.0;
    const int32_t p1 = indptr[row + 1];
    for (int32_t p = indptr[row]; p < p1; ++p) sum = mul_add_unfused(sum, values[p], x[indices[p]]);
    if (MODE == C_SPMV) y[row] = sum;
    else if (MODE == C_RESID) y[row] = __dsub_rn(b[row], sum);
    else if (MODE == C_JACOBI)
        y[row] = __dadd_rn(x[row], __dmul_rn(omega, __dmul_rn(aux[row], __dsub_rn(b[row], sum))));
    else if (MODE == C_PROLONG) y[row] = __dadd_rn(aux[row], sum);
}

__device__ __forceinline__ void gs_update_row(int64_t row, const int32_t *__restrict__ indptr,
                                              const int32_t *__restrict__ indices,
                                              const double *__restrict__ values, double *x,
                                              const double *__restrict__ b) {
    double rsum = 0.0, diag = 0.0;
    const int32_t p1 = indptr[row + 1];
    for (int32_t p = indptr[row]; p < p1; ++p) {
        const int32_t j = indices[p];
        const double a = values[p];
        if (j == row) diag = a;
        else rsum = mul_add_unfused(rsum, a, __ldcg(x + j));   // L2 read: other CTAs may have written x_j
    }
    if (diag != 0.0) __stcg(x + row, __ddiv_rn(__dsub_rn(b[row], rsum), diag));
}

__global__ void __launch_bounds__(kBlock)
gs_rows_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
               const double *__restrict__ values, double *x, const double *__restrict__ b,
               const int32_t *__restrict__ rows, int64_t nrows) {
    const int64_t t = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (t < nrows) gs_update_row(rows[t], indptr, indices, values, x, b);
}

// Exact index-order Gauss-Seidel by dependency levels.  Cooperative launch; levels are separated by a grid
// barrier (a CTA barrier when one CTA suffices).
__global__ void __launch_bounds__(kBlock)
gs_lex_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
              const double *__restrict__ values, double *x, const double *__restrict__ b,
              const int64_t *__restrict__ level_ptr, const int32_t *__restrict__ level_rows,
              int64_t nlevels, int iterations) {
    cg::grid_group grid = cg::this_grid();
    const bool single = gridDim.x == 1;
    const int64_t tid = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int it = 0; it < iterations; ++it) {
        for (int64_t l = 0; l < nlevels; ++l) {
            const int64_t p0 = level_ptr[l], p1 = level_ptr[l + 1];
            for (int64_t p = p0 + tid; p < p1; p += stride)
                gs_update_row(level_rows[p], indptr, indices, values, x, b);
            if (single) __syncthreads();
            else grid.sync();
        }
    }
}

template <int MODE>
static int launch_csr(int64_t n, const int32_t *ip, const int32_t *ix, const double *v, const double *x,
                      const double *b, const double *aux, double *y, double omega, cudaStream_t st,
                      const char *name) {
    if (n <= 0) return MG_OK;
    const int64_t grid = (n + kBlock - 1) / kBlock;
    csr_kernel<MODE><<<(unsigned)grid, kBlock, 0, st>>>(n, ip, ix, v, x, b, aux, y, omega);
    MG_CHECK_LAUNCH(name);
    return MG_OK;
}

int csr_gs_lex(const int32_t *ip, const int32_t *ix, const double *v, double *x, const double *b,
               const int64_t *level_ptr, const int32_t *level_rows, int64_t nlevels, int64_t n,
               int iterations, cudaStream_t st) {
    if (n <= 0 || nlevels <= 0 || iterations <= 0) return MG_OK;
    int max_blocks_per_sm = 0;
    MG_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_blocks_per_sm, gs_lex_kernel, kBlock, 0));
    const int64_t avg_width = (n + nlevels - 1) / nlevels;
    int64_t grid = (avg_width + kBlock - 1) / kBlock;
    const int64_t cap = (int64_t)max_blocks_per_sm * sm_count();
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    void *args[] = {(void *)&ip, (void *)&ix, (void *)&v, (void *)&x, (void *)&b,
                    (void *)&level_ptr, (void *)&level_rows, (void *)&nlevels, (void *)&iterations};
    MG_CHECK_CUDA(cudaLaunchCooperativeKernel((void *)gs_lex_kernel, dim3((unsigned)grid), dim3(kBlock), args, 0, st));
    ++g_launch_count;
    return MG_OK;
}

}  // namespace mgb

using namespace mgb;

extern "C" {

int mg_spmv_csr(int64_t n, const int32_t *d_indptr, const int32_t *d_indices, const double *d_values,
                const double *d_x, double *d_y, void *stream) {
    MG_REQUIRE(n >= 0, "negative size");
    return launch_csr<C_SPMV>(n, d_indptr, d_indices, d_values, d_x, nullptr, nullptr, d_y, 0.0,
                              (cudaStream_t)stream, "mg_spmv_csr");
}
int mg_residual_csr(int64_t n, const int32_t *d_indptr, const int32_t *d_indices, const double *d_values,
                    const double *d_x, const double *d_b, double *d_r, void *stream) {
    MG_REQUIRE(n >= 0, "negative size");
    return launch_csr<C_RESID>(n, d_indptr, d_indices, d_values, d_x, d_b, nullptr, d_r, 0.0,
                               (cudaStream_t)stream, "mg_residual_csr");
}
int mg_jacobi_sweep_csr(int64_t n, const int32_t *d_indptr, const int32_t *d_indices, const double *d_values,
                        const double *d_dinv, const double *d_x, const double *d_b, double *d_x_out,
                        double omega, void *stream) {
    MG_REQUIRE(n >= 0, "negative size");
    MG_REQUIRE(d_x != d_x_out, "Jacobi is out of place: x_out must not alias x");
    return launch_csr<C_JACOBI>(n, d_indptr, d_indices, d_values, d_x, d_b, d_dinv, d_x_out, omega,
                                (cudaStream_t)stream, "mg_jacobi_sweep_csr");
}
int mg_prolong_correct_csr(int64_t n, const int32_t *d_indptr, const int32_t *d_indices,
                           const double *d_values, const double *d_e, double *d_u, void *stream) {
    MG_REQUIRE(n >= 0, "negative size");
    return launch_csr<C_PROLONG>(n, d_indptr, d_indices, d_values, d_e, nullptr, d_u, d_u, 0.0,
                                 (cudaStream_t)stream, "mg_prolong_correct_csr");
}
int mg_gs_multicolor_sweep_csr(int64_t n, const int32_t *d_indptr, const int32_t *d_indices,
                               const double *d_values, double *d_x, const double *d_b,
                               const int64_t *h_color_ptr, const int32_t *d_color_rows, int ncolors,
                               void *stream) {
    MG_REQUIRE(n >= 0 && ncolors >= 0, "negative size");
    MG_REQUIRE(ncolors == 0 || h_color_ptr, "null colour pointer");
    for (int c = 0; c < ncolors; ++c) {
        const int64_t m = h_color_ptr[c + 1] - h_color_ptr[c];
        if (m <= 0) continue;
        gs_rows_kernel<<<(unsigned)((m + kBlock - 1) / kBlock), kBlock, 0, (cudaStream_t)stream>>>(
            d_indptr, d_indices, d_values, d_x, d_b, d_color_rows + h_color_ptr[c], m);
        MG_CHECK_LAUNCH("mg_gs_multicolor_sweep_csr");
    }
    return MG_OK;
}
int mg_gs_lex_sweep_csr(int64_t n, const int32_t *d_indptr, const int32_t *d_indices, const double *d_values,
                        double *d_x, const double *d_b, const int64_t *d_level_ptr,
                        const int32_t *d_level_rows, int64_t nlevels, int iterations, void *stream) {
    MG_REQUIRE(n >= 0 && nlevels >= 0 && iterations >= 0, "negative size");
    return csr_gs_lex(d_indptr, d_indices, d_values, d_x, d_b, d_level_ptr, d_level_rows, nlevels, n,
                      iterations, (cudaStream_t)stream);
}

}  // extern "C"
