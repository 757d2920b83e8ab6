// sell_kernels.cu -- the streaming hot-path kernels on the SELL-32 format (see include/mgb200.h).
//
// One thread per row, one warp per slice.  Entry k of the 32 rows of a slice is contiguous
// (256 B of values + 128 B of columns per warp request), so the matrix streams fully coalesced; the x
// gathers go through L1/L2, which hold the few grid lines a structured stencil touches.
// Every kernel is HBM-bound: algorithmic bytes per row = 12*nnz_row + 4 (CSR yardstick of SURVEY 8d) plus
// the vector traffic listed at each launcher.
#include "common.cuh"

namespace mgb {

enum SellMode { SPMV = 0, RESID = 1, RESNORM = 2, JACOBI = 3, GS = 4, PROLONG = 5 };

struct SellArgs {
    const int64_t *__restrict__ slice_ptr;
    const int32_t *__restrict__ cols;
    const double *__restrict__ vals;
    int64_t row_begin;   // first row this launch touches
    int64_t row_end;     // one past the last row
    int64_t first_row;   // row of thread 0 of block 0 (row_begin rounded down to a slice)
};

template <int MODE>
__global__ void __launch_bounds__(kBlock)
sell_kernel(SellArgs A, const double *x, const double *__restrict__ b, const double *aux,
            double *y, double omega, double *__restrict__ partials) {
    const int64_t row = A.first_row + (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const bool active = row >= A.row_begin && row < A.row_end;
    double contrib = 0.0;
    if (row < A.row_end) {   // warp-uniform except in the last slice
        const int64_t slice = row >> 5;
        const int lane = (int)(row & 31);
        const int64_t base = A.slice_ptr[slice];
        const int len = (int)((A.slice_ptr[slice + 1] - base) >> 5);
        const double *__restrict__ v = A.vals + base + lane;
        const int32_t *__restrict__ c = A.cols + base + lane;
        double sum = 0.0, diag = 0.0;
        int k = 0;
        for (; k + 4 <= len; k += 4) {
            const int32_t c0 = ld_stream(c + (k + 0) * kSlice), c1 = ld_stream(c + (k + 1) * kSlice),
                          c2 = ld_stream(c + (k + 2) * kSlice), c3 = ld_stream(c + (k + 3) * kSlice);
            const double v0 = ld_stream(v + (k + 0) * kSlice), v1 = ld_stream(v + (k + 1) * kSlice),
                         v2 = ld_stream(v + (k + 2) * kSlice), v3 = ld_stream(v + (k + 3) * kSlice);
            const double x0 = x[c0], x1 = x[c1], x2 = x[c2], x3 = x[c3];
            if (MODE == GS) {
                if (c0 == row) { if (v0 != 0.0) diag = v0; } else sum = mul_add_unfused(sum, v0, x0);
                if (c1 == row) { if (v1 != 0.0) diag = v1; } else sum = mul_add_unfused(sum, v1, x1);
                if (c2 == row) { if (v2 != 0.0) diag = v2; } else sum = mul_add_unfused(sum, v2, x2);
                if (c3 == row) { if (v3 != 0.0) diag = v3; } else sum = mul_add_unfused(sum, v3, x3);
            } else {
                sum = mul_add_unfused(sum, v0, x0);
                sum = mul_add_unfused(sum, v1, x1);
                sum = mul_add_unfused(sum, v2, x2);
                sum = mul_add_unfused(sum, v3, x3);
            }
        }
        for (; k < len; ++k) {
            const int32_t c0 = ld_stream(c + k * kSlice);
            const double v0 = ld_stream(v + k * kSlice);
            const double x0 = x[c0];
            if (MODE == GS) {
                if (c0 == row) { if (v0 != 0.0) diag = v0; } else sum = mul_add_unfused(sum, v0, x0);
            } else {
                sum = mul_add_unfused(sum, v0, x0);
            }
        }
        if (active) {
            if (MODE == SPMV) {
                y[row] = sum;
            } else if (MODE == RESID) {
                y[row] = __dsub_rn(b[row], sum);
            } else if (MODE == RESNORM) {
                const double r = __dsub_rn(b[row], sum);
                contrib = r * r;
            } else if (MODE == JACOBI) {
                const double r = __dsub_rn(b[row], sum);
                y[row] = __dadd_rn(x[row], __dmul_rn(omega, __dmul_rn(aux[row], r)));
            } else if (MODE == GS) {
                if (diag != 0.0) y[row] = __ddiv_rn(__dsub_rn(b[row], sum), diag);
            } else if (MODE == PROLONG) {
                y[row] = __dadd_rn(aux[row], sum);   // aux = u (may alias y)
            }
        }
    }
    if (MODE == RESNORM) {
        const double s = block_sum<kBlock>(contrib);
        if (threadIdx.x == 0) partials[blockIdx.x] = s;
    }
}

// second stage of the deterministic norm / dot: one CTA sums the per-block partials in a fixed order
__global__ void __launch_bounds__(1024) reduce_partials_kernel(const double *__restrict__ partials, int64_t n,
                                                               double *__restrict__ out) {
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 1024) s += partials[i];
    s = block_sum<1024>(s);
    if (threadIdx.x == 0) *out = s;
}

template <int MODE>
static int launch_sell(const mg_sell *A, const double *x, const double *b, const double *aux, double *y,
                       double omega, double *partials, int64_t row0, int64_t row1, cudaStream_t st,
                       const char *name) {
    if (row1 <= row0) return MG_OK;
    SellArgs a;
    a.slice_ptr = A->d_slice_ptr;
    a.cols = A->d_cols;
    a.vals = A->d_vals;
    a.row_begin = row0;
    a.row_end = row1;
    a.first_row = row0 & ~(int64_t)(kSlice - 1);
    const int64_t nthreads = row1 - a.first_row;
    const int64_t grid = (nthreads + kBlock - 1) / kBlock;
    if (grid > 0x7fffffffLL) return set_error(MG_ERR_OVERFLOW, name, "grid too large");
    sell_kernel<MODE><<<(unsigned)grid, kBlock, 0, st>>>(a, x, b, aux, y, omega, partials);
    MG_CHECK_LAUNCH(name);
    return MG_OK;
}

int sell_spmv(const mg_sell *A, const double *x, double *y, cudaStream_t st) {
    return launch_sell<SPMV>(A, x, nullptr, nullptr, y, 0.0, nullptr, 0, A->nrows, st, "sell_spmv");
}
int sell_residual(const mg_sell *A, const double *x, const double *b, double *r, cudaStream_t st) {
    return launch_sell<RESID>(A, x, b, nullptr, r, 0.0, nullptr, 0, A->nrows, st, "sell_residual");
}
int sell_residual_norm2(const mg_sell *A, const double *x, const double *b, double *partials, double *out,
                        cudaStream_t st) {
    const int64_t nblocks = (A->nrows + kBlock - 1) / kBlock;
    int rc = launch_sell<RESNORM>(A, x, b, nullptr, nullptr, 0.0, partials, 0, A->nrows, st,
                                  "sell_residual_norm2");
    if (rc) return rc;
    reduce_partials_kernel<<<1, 1024, 0, st>>>(partials, nblocks, out);
    MG_CHECK_LAUNCH("reduce_partials");
    return MG_OK;
}
int sell_jacobi(const mg_sell *A, const double *dinv, const double *x, const double *b, double *xo,
                double omega, cudaStream_t st) {
    return launch_sell<JACOBI>(A, x, b, dinv, xo, omega, nullptr, 0, A->nrows, st, "sell_jacobi");
}
int sell_gs_rows(const mg_sell *A, double *x, const double *b, int64_t row0, int64_t row1, cudaStream_t st) {
    return launch_sell<GS>(A, x, b, nullptr, x, 0.0, nullptr, row0, row1, st, "sell_gs_rows");
}
int sell_prolong(const mg_sell *Q, const double *e, const double *u, double *uo, cudaStream_t st) {
    return launch_sell<PROLONG>(Q, e, nullptr, u, uo, 0.0, nullptr, 0, Q->nrows, st, "sell_prolong");
}

}  // namespace mgb

using namespace mgb;

extern "C" {

static int check_sell(const mg_sell *A) {
    if (!A || A->nrows < 0 || (A->nrows > 0 && (!A->d_slice_ptr || !A->d_cols || !A->d_vals)))
        return set_error(MG_ERR_INVALID, "mg_sell", "null or negative-sized SELL matrix");
    if (A->nslices != (A->nrows + kSlice - 1) / kSlice)
        return set_error(MG_ERR_INVALID, "mg_sell", "nslices != ceil(nrows/32)");
    return MG_OK;
}

int mg_sell_spmv(const mg_sell *A, const double *d_x, double *d_y, void *stream) {
    if (int rc = check_sell(A)) return rc;
    return sell_spmv(A, d_x, d_y, (cudaStream_t)stream);
}
int mg_sell_residual(const mg_sell *A, const double *d_x, const double *d_b, double *d_r, void *stream) {
    if (int rc = check_sell(A)) return rc;
    return sell_residual(A, d_x, d_b, d_r, (cudaStream_t)stream);
}
int64_t mg_norm_workspace_size(int64_t n) { return (n + kBlock - 1) / kBlock + 1; }
int mg_sell_residual_norm2(const mg_sell *A, const double *d_x, const double *d_b, double *d_partials,
                           double *d_norm2, void *stream) {
    if (int rc = check_sell(A)) return rc;
    if (A->nrows == 0) return mg_fill(1, 0.0, d_norm2, stream);
    return sell_residual_norm2(A, d_x, d_b, d_partials, d_norm2, (cudaStream_t)stream);
}
int mg_sell_jacobi(const mg_sell *A, const double *d_dinv, const double *d_x, const double *d_b,
                   double *d_x_out, double omega, void *stream) {
    if (int rc = check_sell(A)) return rc;
    MG_REQUIRE(d_x != d_x_out, "Jacobi is out of place: x_out must not alias x");
    return sell_jacobi(A, d_dinv, d_x, d_b, d_x_out, omega, (cudaStream_t)stream);
}
int mg_sell_gs_rows(const mg_sell *A, double *d_x, const double *d_b, int64_t row0, int64_t row1,
                    void *stream) {
    if (int rc = check_sell(A)) return rc;
    MG_REQUIRE(row0 >= 0 && row1 <= A->nrows && row0 <= row1, "row range outside the matrix");
    return sell_gs_rows(A, d_x, d_b, row0, row1, (cudaStream_t)stream);
}
int mg_sell_prolong_correct(const mg_sell *Q, const double *d_e, const double *d_u, double *d_u_out,
                            void *stream) {
    if (int rc = check_sell(Q)) return rc;
    return sell_prolong(Q, d_e, d_u, d_u_out, (cudaStream_t)stream);
}

}  // extern "C"
