// bcr.cu -- coarsest-level direct solver for banded operators: block cyclic reduction with dense blocks.
//
// Replaces spsolve(A_coarse, res_coarse) (SuperLU, refactorised by the reference in every cycle,
// Multigrid.py:106) when the coarsest grid is too large for an explicit dense inverse (e.g. 257^2 unknowns at
// the bottom of a 6-level hierarchy on 8193^2).  A matrix with half bandwidth <= m is block tridiagonal with
// m x m blocks (L_i, D_i, U_i).  Each reduction level eliminates the odd-positioned blocks,
//     x_p = Dinv_p f_p - HL_p x_{p-1} - HU_p x_{p+1}          HL = Dinv L,  HU = Dinv U
// and updates the even ones,
//     f_p <- f_p - GL_p f_{p-1} - GU_p f_{p+1}                GL_p = L_p Dinv_{p-1},  GU_p = U_p Dinv_{p+1}
//     D_p <- D_p - GL_p U_{p-1} - GU_p L_{p+1},  L_p <- -GL_p L_{p-1},  U_p <- -GU_p U_{p+1}
// All dense factors are formed ONCE at setup (batched Gauss-Jordan with partial pivoting + batched GEMM below);
// a solve is 2*levels+1 launches of batched matrix-vector kernels that stream the stored factors once.
#include "exchange.cuh"

namespace mgb {

struct BcrHandle {            // mirrors mg_bcr in include/mgb200.h
    int64_t n, n_pad, m, nb;
    int32_t nlevels, pad_;
    const double *GL[32], *GU[32], *Dinv[32], *HL[32], *HU[32];
    int64_t na[32];
    const double *last_inv;
    double *f, *x;
    int64_t tail_na;          // blocks left after the reductions (0/1: one block); last_inv is (tail_na*m)^2
    double *tail;             // tail_na*m doubles: contiguous right-hand side of the tail system
    const int32_t *perm;      // optional: the factors belong to the rows / columns perm[0..n) of the caller's operator
};

// ---- setup kernels -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
bcr_blocks_kernel(int64_t n, int64_t n_pad, int64_t m, const int32_t *__restrict__ indptr,
                  const int32_t *__restrict__ indices, const double *__restrict__ values, double *__restrict__ D,
                  double *__restrict__ L, double *__restrict__ U, int32_t *__restrict__ bad) {
    const int64_t r = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (r >= n_pad) return;
    const int64_t bi = r / m, ri = r % m;
    if (r >= n) { D[(bi * m + ri) * m + ri] = 1.0; return; }      // identity padding rows
    for (int32_t p = indptr[r]; p < indptr[r + 1]; ++p) {
        const int64_t c = indices[p];
        const int64_t bj = c / m, ci = c % m;
        double *dst = (bj == bi) ? D : (bj == bi - 1) ? L : (bj == bi + 1) ? U : nullptr;
        if (!dst) { atomicExch(bad, 1); continue; }
        dst[(bi * m + ri) * m + ci] += values[p];
    }
}

// One CTA per matrix: Gauss-Jordan with partial pivoting on W = [A | I] (m x 2m, global / L2 resident).
constexpr int kGjThreads = 1024;
constexpr int kGjMaxM = 4096;

__global__ void __launch_bounds__(kGjThreads)
batched_gauss_jordan_kernel(int m, const double *__restrict__ A, int64_t strideA, double *__restrict__ Wall,
                            double *__restrict__ out, int64_t strideOut, int32_t *__restrict__ singular) {
    __shared__ double s_abs[kGjThreads];
    __shared__ int s_row[kGjThreads];
    __shared__ int s_piv_of_col[kGjMaxM];
    __shared__ unsigned char s_pivoted[kGjMaxM];
    const int tid = threadIdx.x;
    const int64_t w = 2 * (int64_t)m;
    const double *Ab = A + (int64_t)blockIdx.x * strideA;
    double *W = Wall + (int64_t)blockIdx.x * m * w;
    for (int64_t e = tid; e < (int64_t)m * w; e += kGjThreads) {
        const int64_t r = e / w, c = e % w;
        W[e] = (c < m) ? Ab[r * m + c] : ((c - m == r) ? 1.0 : 0.0);
    }
    for (int r = tid; r < m; r += kGjThreads) s_pivoted[r] = 0;
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31, nwarps = kGjThreads / 32;
    for (int k = 0; k < m; ++k) {
        double best = -1.0;
        int brow = -1;
        for (int r = tid; r < m; r += kGjThreads) {
            if (!s_pivoted[r]) {
                const double a = fabs(W[r * w + k]);
                if (a > best) { best = a; brow = r; }
            }
        }
        s_abs[tid] = best;
        s_row[tid] = brow;
        __syncthreads();
        for (int o = kGjThreads / 2; o > 0; o >>= 1) {
            if (tid < o) {
                const double a = s_abs[tid + o];
                const int rr = s_row[tid + o];
                if (rr >= 0 && (a > s_abs[tid] || (a == s_abs[tid] && (s_row[tid] < 0 || rr < s_row[tid])))) {
                    s_abs[tid] = a;
                    s_row[tid] = rr;
                }
            }
            __syncthreads();
        }
        const int p = s_row[0];
        const double pa = s_abs[0];
        __syncthreads();
        if (p < 0 || pa == 0.0) {
            if (tid == 0) atomicExch(singular, 1);
            return;
        }
        if (tid == 0) { s_pivoted[p] = 1; s_piv_of_col[k] = p; }
        const double *prow = W + (int64_t)p * w;
        const double pinv = 1.0 / prow[k];
        for (int r = warp; r < m; r += nwarps) {
            if (r == p) continue;
            double *row = W + (int64_t)r * w;
            const double f = row[k] * pinv;
            if (f != 0.0)
                for (int64_t c = k + 1 + lane; c < w; c += 32) row[c] -= f * prow[c];
        }
        __syncthreads();
    }
    double *ob = out + (int64_t)blockIdx.x * strideOut;
    for (int64_t e = tid; e < (int64_t)m * m; e += kGjThreads) {
        const int64_t k = e / m, c = e % m;
        const int p = s_piv_of_col[k];
        ob[e] = W[(int64_t)p * w + m + c] / W[(int64_t)p * w + k];
    }
}

// C[b] = alpha * A[b] * B[b] + beta * C[b], all m x m row-major, element strides between batch members.
constexpr int TM = 64, TN = 64, TK = 16;

__global__ void __launch_bounds__(256)
gemm_batched_kernel(int m, const double *__restrict__ A, int64_t sA, const double *__restrict__ B, int64_t sB,
                    double *__restrict__ C, int64_t sC, double alpha, double beta) {
    __shared__ double As[TK][TM + 1];
    __shared__ double Bs[TK][TN];
    const double *Ab = A + (int64_t)blockIdx.z * sA;
    const double *Bb = B + (int64_t)blockIdx.z * sB;
    double *Cb = C + (int64_t)blockIdx.z * sC;
    const int row0 = blockIdx.y * TM, col0 = blockIdx.x * TN;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;       // 16 x 16 threads, 4 x 4 outputs each
    double acc[4][4] = {};
    for (int k0 = 0; k0 < m; k0 += TK) {
        for (int e = threadIdx.x; e < TM * TK; e += 256) {        // A tile: rows row0.., cols k0..
            const int r = e / TK, c = e % TK;
            const int gr = row0 + r, gc = k0 + c;
            As[c][r] = (gr < m && gc < m) ? Ab[(int64_t)gr * m + gc] : 0.0;
        }
        for (int e = threadIdx.x; e < TK * TN; e += 256) {        // B tile: rows k0.., cols col0..
            const int r = e / TN, c = e % TN;
            const int gr = k0 + r, gc = col0 + c;
            Bs[r][c] = (gr < m && gc < m) ? Bb[(int64_t)gr * m + gc] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gr = row0 + ty * 4 + i;
        if (gr >= m) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gc = col0 + tx * 4 + j;
            if (gc >= m) continue;
            const int64_t o = (int64_t)gr * m + gc;
            Cb[o] = (beta == 0.0) ? alpha * acc[i][j] : alpha * acc[i][j] + beta * Cb[o];
        }
    }
}

// ---- solve kernels: one warp per block row ------------------------------------------------------------------------
__device__ __forceinline__ double warp_row_dot(const double *__restrict__ row, const double *__restrict__ v, int m,
                                               int lane) {
    double s = 0.0;
    for (int c = lane; c < m; c += 32) s += __ldcs(row + c) * v[c];
    return s;
}

__global__ void __launch_bounds__(kBlock)
bcr_load_kernel(int64_t n, int64_t n_pad, const double *__restrict__ b, const int32_t *__restrict__ perm,
                double *__restrict__ f) {
    pdl_prologue();
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i < n_pad) f[i] = (i < n) ? b[perm ? perm[i] : i] : 0.0;
}

// level s forward: kept blocks j in [j0, j0+nk) (position p = 2j, original block p << s)
__global__ void __launch_bounds__(kBlock)
bcr_forward_kernel(int m, int s, int64_t na, int64_t j0, int64_t nk, const double *__restrict__ GL,
                   const double *__restrict__ GU, double *f) {
    pdl_prologue();
    const int64_t wid = ((int64_t)blockIdx.x * kBlock + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= nk * m) return;
    const int64_t j = j0 + wid / m;
    const int r = (int)(wid % m);
    const int64_t p = 2 * j;
    double acc = 0.0;
    if (p >= 1) acc += warp_row_dot(GL + (j * m + r) * (int64_t)m, f + ((p - 1) << s) * m, m, lane);
    if (p + 1 < na) acc += warp_row_dot(GU + (j * m + r) * (int64_t)m, f + ((p + 1) << s) * m, m, lane);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if (lane == 0) f[(p << s) * m + r] -= acc;
}

// the system left after s reduction levels (tail_na blocks at positions p << s): gather its right-hand side into a
// contiguous vector; the dense inverse is then applied by gemv_rows (dense_kernels.cu), which scatters the result back
__global__ void __launch_bounds__(kBlock)
bcr_tail_gather_kernel(int m, int s, int64_t nt, const double *__restrict__ f, double *__restrict__ g) {
    pdl_prologue();
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i < nt) g[i] = f[((i / m) << s) * m + i % m];
}

// level s backward: eliminated blocks j in [j0, j0+nodd) (position p = 2j+1)
__global__ void __launch_bounds__(kBlock)
bcr_backward_kernel(int m, int s, int64_t na, int64_t j0, int64_t nodd, const double *__restrict__ Dinv,
                    const double *__restrict__ HL, const double *__restrict__ HU, const double *__restrict__ f,
                    double *x) {
    pdl_prologue();
    const int64_t wid = ((int64_t)blockIdx.x * kBlock + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= nodd * m) return;
    const int64_t j = j0 + wid / m;
    const int r = (int)(wid % m);
    const int64_t p = 2 * j + 1;
    const int64_t ro = (j * m + r) * (int64_t)m;
    double acc = warp_row_dot(Dinv + ro, f + (p << s) * m, m, lane);
    acc -= warp_row_dot(HL + ro, x + ((p - 1) << s) * m, m, lane);
    if (p + 1 < na) acc -= warp_row_dot(HU + ro, x + ((p + 1) << s) * m, m, lane);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if (lane == 0) x[(p << s) * m + r] = acc;
}

__global__ void __launch_bounds__(kBlock)
bcr_store_kernel(int64_t n, const double *__restrict__ x, const int32_t *__restrict__ perm, double *__restrict__ out) {
    pdl_prologue();
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i < n) out[perm ? perm[i] : i] = x[i];
}

static inline unsigned warps_grid(int64_t nwarps) { return (unsigned)((nwarps * 32 + kBlock - 1) / kBlock); }

int comm_exchange(mg_comm *, const mg_xfer *, const double *, double *, cudaStream_t);
int gemv_rows(int64_t total_rows, int64_t row0, int64_t nrows, int64_t m, const double *M, const double *x, double *y,
              int64_t bm, int shift, cudaStream_t st);

// x = A^-1 rhs.  With `dist` (and `comm`) the block rows of the large reduction levels and the rows of the dense tail
// are split over the ranks and every such step is followed by an all-gather of what the ranks computed, so f and x
// stay complete on every rank; the per-row arithmetic, hence the result, is the same as in the replicated solve.
int bcr_solve(const void *handle, const mg_bcr_dist *dist, mg_comm *comm, const double *rhs, double *x,
              cudaStream_t st) {
    const BcrHandle *H = (const BcrHandle *)handle;
    if (!H || H->m <= 0 || H->nlevels < 0 || H->nlevels > 32) return set_error(MG_ERR_INVALID, "bcr_solve", "bad handle");
    if (dist && !comm) dist = nullptr;
    const int m = (int)H->m;
    launch_k(bcr_load_kernel, (unsigned)((unsigned)((H->n_pad + kBlock - 1) / kBlock)), (unsigned)kBlock, st, H->n, H->n_pad, rhs, H->perm, H->f);
    MG_CHECK_LAUNCH("bcr_load");
    for (int s = 0; s < H->nlevels; ++s) {
        const int64_t na = H->na[s], nk = (na + 1) / 2;
        const bool split = dist && dist->fwd_xfer[s];
        const int64_t j0 = split ? dist->fwd_j0[s] : 0, j1 = split ? dist->fwd_j1[s] : nk;
        if (j1 > j0) {
            launch_k(bcr_forward_kernel, (unsigned)(warps_grid((j1 - j0) * m)), (unsigned)kBlock, st, m, s, na, j0, j1 - j0, H->GL[s], H->GU[s], H->f);
            MG_CHECK_LAUNCH("bcr_forward");
        }
        if (split) {
            int rc = comm_exchange(comm, dist->fwd_xfer[s], H->f, H->f, st);
            if (rc) return rc;
        }
    }
    {
        const int64_t tna = H->tail_na > 1 ? H->tail_na : 1;
        const bool split = dist && dist->tail_xfer;
        const int64_t i0 = split ? dist->tail_i0 : 0, i1 = split ? dist->tail_i1 : tna * m;
        if (!H->tail) return set_error(MG_ERR_INVALID, "bcr_solve", "handle has no tail work vector");
        const int64_t nt = tna * m;
        launch_k(bcr_tail_gather_kernel, (unsigned)((unsigned)((nt + kBlock - 1) / kBlock)), (unsigned)kBlock, st, m, H->nlevels, nt, H->f, H->tail);
        MG_CHECK_LAUNCH("bcr_tail_gather");
        int rc = gemv_rows(nt, i0, i1 - i0, nt, H->last_inv, H->tail, H->x, m, H->nlevels, st);
        if (rc) return rc;
        if (split) {
            int rc = comm_exchange(comm, dist->tail_xfer, H->x, H->x, st);
            if (rc) return rc;
        }
    }
    for (int s = H->nlevels - 1; s >= 0; --s) {
        const int64_t na = H->na[s], nodd = na / 2;
        const bool split = dist && dist->bwd_xfer[s];
        const int64_t j0 = split ? dist->bwd_j0[s] : 0, j1 = split ? dist->bwd_j1[s] : nodd;
        if (j1 > j0) {
            launch_k(bcr_backward_kernel, (unsigned)(warps_grid((j1 - j0) * m)), (unsigned)kBlock, st, m, s, na, j0, j1 - j0, H->Dinv[s], H->HL[s],
                                                                              H->HU[s], H->f, H->x);
            MG_CHECK_LAUNCH("bcr_backward");
        }
        if (split) {
            int rc = comm_exchange(comm, dist->bwd_xfer[s], H->x, H->x, st);
            if (rc) return rc;
        }
    }
    launch_k(bcr_store_kernel, (unsigned)((unsigned)((H->n + kBlock - 1) / kBlock)), (unsigned)kBlock, st, H->n, H->x, H->perm, x);
    MG_CHECK_LAUNCH("bcr_store");
    return MG_OK;
}

}  // namespace mgb

using namespace mgb;

extern "C" {

/* scatter a banded CSR matrix (half bandwidth <= m) into zeroed dense block arrays D, L, U ([nb][m][m] each);
 * rows >= n are identity padding.  *d_bad is set if an entry falls outside the block tridiagonal. */
int mg_bcr_blocks_from_csr(int64_t n, int64_t n_pad, int64_t m, const int32_t *d_indptr, const int32_t *d_indices,
                           const double *d_values, double *d_D, double *d_L, double *d_U, int32_t *d_bad,
                           void *stream) {
    MG_REQUIRE(n > 0 && m > 0 && n_pad >= n && n_pad % m == 0, "bad sizes");
    bcr_blocks_kernel<<<(unsigned)((n_pad + kBlock - 1) / kBlock), kBlock, 0, (cudaStream_t)stream>>>(
        n, n_pad, m, d_indptr, d_indices, d_values, d_D, d_L, d_U, d_bad);
    MG_CHECK_LAUNCH("bcr_blocks");
    return MG_OK;
}

/* batch of dense inverses (Gauss-Jordan, partial pivoting), one CTA per matrix; d_work: batch*m*2m doubles */
int mg_dense_inverse_batched(int64_t m, int64_t batch, const double *d_a, int64_t stride_a, double *d_out,
                             int64_t stride_out, double *d_work, int32_t *d_singular, void *stream) {
    MG_REQUIRE(m > 0 && m <= kGjMaxM && batch >= 0, "block size out of range");
    if (batch == 0) return MG_OK;
    batched_gauss_jordan_kernel<<<(unsigned)batch, kGjThreads, 0, (cudaStream_t)stream>>>(
        (int)m, d_a, stride_a, d_work, d_out, stride_out, d_singular);
    MG_CHECK_LAUNCH("batched_gauss_jordan");
    return MG_OK;
}

/* C[b] = alpha*A[b]*B[b] + beta*C[b] for b < batch, m x m row-major, strides in elements */
int mg_dense_gemm_batched(int64_t m, int64_t batch, const double *d_a, int64_t stride_a, const double *d_b,
                          int64_t stride_b, double *d_c, int64_t stride_c, double alpha, double beta, void *stream) {
    MG_REQUIRE(m > 0 && batch >= 0 && batch <= 65535, "bad sizes");
    if (batch == 0) return MG_OK;
    dim3 grid((unsigned)((m + TN - 1) / TN), (unsigned)((m + TM - 1) / TM), (unsigned)batch);
    gemm_batched_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((int)m, d_a, stride_a, d_b, stride_b, d_c, stride_c,
                                                                alpha, beta);
    MG_CHECK_LAUNCH("gemm_batched");
    return MG_OK;
}

/* x = A^-1 rhs with the factors of `bcr` (an mg_bcr filled by the host side) */
int mg_bcr_solve(const mg_bcr *bcr, const double *d_rhs, double *d_x, void *stream) {
    MG_REQUIRE(bcr && d_rhs && d_x, "null argument");
    return bcr_solve((const void *)bcr, nullptr, nullptr, d_rhs, d_x, (cudaStream_t)stream);
}

/* the same solve with the large steps split over the ranks of `comm` (one program of its own) */
int mg_bcr_solve_dist(mg_comm *comm, const mg_bcr *bcr, const mg_bcr_dist *dist, const double *d_rhs, double *d_x,
                      void *stream) {
    MG_REQUIRE(comm && bcr && dist && d_rhs && d_x, "null argument");
    int rc = mg_comm_begin(comm);
    if (!rc) rc = bcr_solve((const void *)bcr, dist, comm, d_rhs, d_x, (cudaStream_t)stream);
    if (!rc) rc = mg_comm_end(comm, stream);
    return rc;
}

}  // extern "C"
