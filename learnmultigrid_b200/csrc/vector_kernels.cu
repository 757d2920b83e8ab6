// vector_kernels.cu -- level-vector kernels: dot / norm (deterministic two-stage), axpby, fill, permute.
// Reference call sites: np.linalg.norm at Multigrid.py:63; dots and updates of CG.py:30-48.
#include "common.cuh"

namespace mgb {

__global__ void __launch_bounds__(kBlock)
dot_partials_kernel(int64_t n, const double *__restrict__ x, const double *__restrict__ y,
                    double *__restrict__ partials) {
    pdl_prologue();
    // each CTA owns a fixed contiguous chunk of 4*kBlock elements per pass -> fixed summation order
    double s = 0.0;
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) s += x[i] * y[i];
    s = block_sum_last<kBlock>(s);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

__global__ void __launch_bounds__(1024) reduce_partials_kernel2(const double *__restrict__ partials, int64_t n,
                                                                double *__restrict__ out) {
    pdl_prologue();
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 1024) s += partials[i];
    s = block_sum<1024>(s);
    if (threadIdx.x == 0) *out = s;
}

__global__ void __launch_bounds__(kBlock)
axpby_kernel(int64_t n, double a, const double *x, double b, const double *y, double *out) {
    pdl_prologue();
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) {
        // a*x + b*y with separately rounded products and sum; b == 0 skips y entirely (y may be null)
        const double ax = __dmul_rn(a, x[i]);
        out[i] = (b == 0.0) ? ax : __dadd_rn(ax, __dmul_rn(b, y[i]));
    }
}

// out = omega * (dinv .* b): the first damped-Jacobi sweep from a zero iterate (no matrix pass needed)
__global__ void __launch_bounds__(kBlock)
diag_scale_kernel(int64_t n, double omega, const double *__restrict__ dinv, const double *__restrict__ b,
                  double *__restrict__ out) {
    pdl_prologue();
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride)
        out[i] = __dadd_rn(0.0, __dmul_rn(omega, __dmul_rn(dinv[i], b[i])));
}

__global__ void __launch_bounds__(kBlock) fill_kernel(int64_t n, double v, double *x) {
    pdl_prologue();
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) x[i] = v;
}

__global__ void __launch_bounds__(kBlock)
gather_kernel(int64_t n, const int32_t *__restrict__ idx, const double *__restrict__ in, double *__restrict__ out) {
    pdl_prologue();
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) out[i] = in[idx[i]];
}

__global__ void __launch_bounds__(kBlock)
scatter_kernel(int64_t n, const int32_t *__restrict__ idx, const double *__restrict__ in, double *__restrict__ out) {
    pdl_prologue();
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) out[idx[i]] = in[i];
}

// ---- conjugate gradients with the scalars on the device (cycle.cu mg_pcg_*; CG.py:30-48) ----------------------------
// scalars s[]: [0] r.z  [1] p.Ap  [2] alpha  [3] beta  [4] r.r
// p = z + beta p  (first iteration: p = z); same rounding as mg_axpby(beta, p, 1, z, p)
__global__ void __launch_bounds__(kBlock)
pcg_direction_kernel(int64_t n, const double *__restrict__ z, double *__restrict__ p, const double *__restrict__ s, int first) {
    pdl_prologue();
    const double beta = first ? 0.0 : s[3];
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride)
        p[i] = first ? z[i] : __dadd_rn(__dmul_rn(beta, p[i]), z[i]);
}
// x += alpha p, r -= alpha Ap and the per-CTA partial sums of r.r in one pass (48 B per row instead of 64 + a dot)
__global__ void __launch_bounds__(kBlock)
pcg_update_kernel(int64_t n, const double *__restrict__ p, const double *__restrict__ Ap, double *__restrict__ x,
                  double *__restrict__ r, const double *__restrict__ s, double *__restrict__ partials) {
    pdl_prologue();
    const double alpha = s[2];
    double acc = 0.0;
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) {
        x[i] = __dadd_rn(__dmul_rn(alpha, p[i]), x[i]);
        const double ri = __dadd_rn(__dmul_rn(-alpha, Ap[i]), r[i]);
        r[i] = ri;
        acc += ri * ri;
    }
    acc = block_sum_last<kBlock>(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}
// what follows a reduction: op 0: s[4] = v (r.r);  op 1: beta = v / s[0], s[0] = v (v = new r.z);  op 2: s[1] = v,
// alpha = s[0] / v (v = p.Ap)
__global__ void pcg_scalar_kernel(int op, int first, double *s, const double *in) {
    pdl_prologue();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double v = *in;
    if (op == 0) {
        s[4] = v;
    } else if (op == 1) {
        s[3] = first ? 0.0 : v / s[0];
        s[0] = v;
    } else {
        s[1] = v;
        s[2] = s[0] / v;
    }
}

static inline unsigned stream_grid(int64_t n) {
    int64_t g = (n + kBlock - 1) / kBlock;
    const int64_t cap = (int64_t)sm_count() * 8;   // 8 resident CTAs of 256 threads per SM
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

int vec_dot(int64_t n, const double *x, const double *y, double *partials, double *out, cudaStream_t st) {
    const unsigned g = stream_grid(n);
    launch_k(dot_partials_kernel, (unsigned)(g), (unsigned)kBlock, st, n, x, y, partials);
    MG_CHECK_LAUNCH("dot_partials");
    launch_k(reduce_partials_kernel2, (unsigned)(1), (unsigned)1024, st, partials, g, out);
    MG_CHECK_LAUNCH("reduce_partials");
    return MG_OK;
}
int vec_dot_partials(int64_t n, const double *x, const double *y, double *partials, int *nblocks, cudaStream_t st) {
    const unsigned g = stream_grid(n);
    launch_k(dot_partials_kernel, (unsigned)(g), (unsigned)kBlock, st, n, x, y, partials);
    MG_CHECK_LAUNCH("dot_partials");
    *nblocks = (int)g;
    return MG_OK;
}
int vec_pcg_direction(int64_t n, const double *z, double *p, const double *s, int first, cudaStream_t st) {
    if (n <= 0) return MG_OK;
    launch_k(pcg_direction_kernel, (unsigned)(stream_grid(n)), (unsigned)kBlock, st, n, z, p, s, first);
    MG_CHECK_LAUNCH("pcg_direction");
    return MG_OK;
}
int vec_pcg_update(int64_t n, const double *p, const double *Ap, double *x, double *r, const double *s,
                   double *partials, int *nblocks, cudaStream_t st) {
    const unsigned g = stream_grid(n);
    launch_k(pcg_update_kernel, (unsigned)(g), (unsigned)kBlock, st, n, p, Ap, x, r, s, partials);
    MG_CHECK_LAUNCH("pcg_update");
    *nblocks = (int)g;
    return MG_OK;
}
int vec_pcg_scalar(int op, int first, double *s, const double *in, cudaStream_t st) {
    launch_k(pcg_scalar_kernel, 1u, 32u, st, op, first, s, in);
    MG_CHECK_LAUNCH("pcg_scalar");
    return MG_OK;
}
int vec_axpby(int64_t n, double a, const double *x, double b, const double *y, double *out, cudaStream_t st) {
    if (n <= 0) return MG_OK;
    launch_k(axpby_kernel, (unsigned)(stream_grid(n)), (unsigned)kBlock, st, n, a, x, b, y, out);
    MG_CHECK_LAUNCH("axpby");
    return MG_OK;
}
int vec_diag_scale(int64_t n, double omega, const double *dinv, const double *b, double *out, cudaStream_t st) {
    if (n <= 0) return MG_OK;
    launch_k(diag_scale_kernel, (unsigned)(stream_grid(n)), (unsigned)kBlock, st, n, omega, dinv, b, out);
    MG_CHECK_LAUNCH("diag_scale");
    return MG_OK;
}
int vec_fill(int64_t n, double v, double *x, cudaStream_t st) {
    if (n <= 0) return MG_OK;
    launch_k(fill_kernel, (unsigned)(stream_grid(n)), (unsigned)kBlock, st, n, v, x);
    MG_CHECK_LAUNCH("fill");
    return MG_OK;
}

int vec_scatter(int64_t n, const int32_t *idx, const double *in, double *out, cudaStream_t st) {
    if (n <= 0) return MG_OK;
    launch_k(scatter_kernel, (unsigned)(stream_grid(n)), (unsigned)kBlock, st, n, idx, in, out);
    MG_CHECK_LAUNCH("scatter");
    return MG_OK;
}

}  // namespace mgb

using namespace mgb;

extern "C" {

int mg_dot(int64_t n, const double *d_x, const double *d_y, double *d_partials, double *d_out, void *stream) {
    MG_REQUIRE(n >= 0, "negative size");
    return vec_dot(n, d_x, d_y, d_partials, d_out, (cudaStream_t)stream);
}
int mg_axpby(int64_t n, double a, const double *d_x, double b, const double *d_y, double *d_out, void *stream) {
    MG_REQUIRE(n >= 0, "negative size");
    return vec_axpby(n, a, d_x, b, d_y, d_out, (cudaStream_t)stream);
}
int mg_fill(int64_t n, double value, double *d_x, void *stream) {
    MG_REQUIRE(n >= 0, "negative size");
    return vec_fill(n, value, d_x, (cudaStream_t)stream);
}
int mg_gather(int64_t n, const int32_t *d_idx, const double *d_in, double *d_out, void *stream) {
    MG_REQUIRE(n >= 0, "negative size");
    if (n == 0) return MG_OK;
    launch_k(gather_kernel, (unsigned)(stream_grid(n)), (unsigned)kBlock, (cudaStream_t)stream, n, d_idx, d_in, d_out);
    MG_CHECK_LAUNCH("gather");
    return MG_OK;
}
int mg_scatter(int64_t n, const int32_t *d_idx, const double *d_in, double *d_out, void *stream) {
    MG_REQUIRE(n >= 0, "negative size");
    return vec_scatter(n, d_idx, d_in, d_out, (cudaStream_t)stream);
}

}  // extern "C"
