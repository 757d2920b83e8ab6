// setup_kernels.cu -- hierarchy setup on the device: two-pass SpGEMM for the Galerkin products, CSR transpose,
// row/column permutation, SELL-32 build, diagonal extraction.
//
// Galerkin product.  The reference computes csr_matrix(i.T @ A @ i) (Multigrid.py:97-98) with SciPy's two-pass
// csr_matmat.  Following SciPy's own evaluation order (CSC operands are multiplied as transposed CSR), the value
// it produces is
//     T = A^T Q          T[k,r] = sum_{m ascending} A[m,k] * Q[m,r]
//     C = Q^T T          C[c,r] = sum_{k ascending} Q[k,c] * T[k,r]          A_c = C^T
// every product rounded, then added in that order, and entries whose sum is exactly 0 dropped by the numeric pass.
// spgemm_kernel reproduces exactly this: a group of G lanes owns one output row, walks the row of the left
// operand sequentially and spreads the matching row of the right operand over its lanes (distinct output columns),
// so each output entry receives its contributions in the same order as SciPy -> bit-identical values, and the
// exact-zero pruning yields bit-identical sparsity patterns.
#include <cub/cub.cuh>
#include "common.cuh"

namespace mgb {

// ------------------------------------------------------------------------------------------------------------
// scans / sorts (CUB, temp storage supplied by the caller)
struct ToI64 {
    __host__ __device__ __forceinline__ int64_t operator()(const int32_t &v) const { return (int64_t)v; }
};
struct Times32 {
    __host__ __device__ __forceinline__ int64_t operator()(const int32_t &v) const { return (int64_t)v * kSlice; }
};

__global__ void iota_kernel(int64_t n, int32_t *out) {
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) out[i] = (int32_t)i;
}

static inline unsigned grid_for(int64_t n) {
    int64_t g = (n + kBlock - 1) / kBlock;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

// ------------------------------------------------------------------------------------------------------------
// SpGEMM
__device__ __forceinline__ uint32_t hash_col(int32_t c, int log_t) {
    return (uint32_t)((uint32_t)c * 0x9E3779B1u) >> (32 - log_t);
}

template <int G, bool NUMERIC>
__global__ void __launch_bounds__(kBlock)
spgemm_kernel(int64_t nrows, const int32_t *__restrict__ a_ptr, const int32_t *__restrict__ a_idx,
              const double *__restrict__ a_val, const int32_t *__restrict__ b_ptr,
              const int32_t *__restrict__ b_idx, const double *__restrict__ b_val, int log_t,
              int32_t *__restrict__ row_count, const int32_t *__restrict__ c_ptr, int32_t *__restrict__ c_idx,
              double *__restrict__ c_val, int32_t *__restrict__ overflow) {
    extern __shared__ __align__(8) unsigned char smem[];
    const int T = 1 << log_t;
    const int groups = kBlock / G;
    const int g = threadIdx.x / G;
    const int lane = threadIdx.x % G;
    // values first (8-byte aligned), then keys
    double *vals = reinterpret_cast<double *>(smem) + (size_t)g * T;
    int32_t *keys = reinterpret_cast<int32_t *>(smem + (NUMERIC ? (size_t)groups * T * sizeof(double) : 0)) + (size_t)g * T;
    const unsigned full = 0xffffffffu;
    const unsigned gmask = (G == 32) ? full : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
    const int64_t gstride = (int64_t)gridDim.x * groups;
    for (int64_t row = (int64_t)blockIdx.x * groups + g; row < nrows; row += gstride) {
        for (int t = lane; t < T; t += G) {
            keys[t] = -1;
            if (NUMERIC) vals[t] = 0.0;
        }
        __syncwarp(gmask);
        const int32_t a0 = a_ptr[row], a1 = a_ptr[row + 1];
        for (int32_t jj = a0; jj < a1; ++jj) {
            const int32_t k = a_idx[jj];
            const double av = NUMERIC ? a_val[jj] : 0.0;
            const int32_t b0 = b_ptr[k], b1 = b_ptr[k + 1];
            for (int32_t kk = b0 + lane; kk < b1; kk += G) {
                const int32_t c = b_idx[kk];
                uint32_t slot = hash_col(c, log_t);
                int probes = 0;
                while (true) {
                    const int32_t old = atomicCAS(&keys[slot], -1, c);
                    if (old == -1 || old == c) break;
                    slot = (slot + 1) & (T - 1);
                    if (++probes >= T) { atomicExch(overflow, 1); slot = 0xffffffffu; break; }
                }
                if (NUMERIC && slot != 0xffffffffu) vals[slot] = mul_add_unfused(vals[slot], av, b_val[kk]);
            }
            __syncwarp(gmask);   // the next left-operand entry adds after this one, in order
        }
        // ---- extraction ----
        int cnt = 0;
        if (!NUMERIC) {
            for (int t = lane; t < T; t += G) cnt += (keys[t] != -1);
        } else {
            // The occupied slots are first moved to the front of the table -- in place, G slots at a time: a chunk is
            // read by all lanes before any of its entries is written, and entries only move towards the front -- and
            // then ranked among themselves: n_occ^2 / G comparisons per lane instead of T^2 / G (a 5-point operator
            // times linear interpolation fills 9 of 32 slots; ranking over the whole table was most of the kernel).
            const int gshift = (G == 32) ? 0 : ((threadIdx.x & 31) / G * G);
            int n_occ = 0;
            for (int t0 = 0; t0 < T; t0 += G) {
                const int t = t0 + lane;
                int32_t key = -1;
                double v = 0.0;
                if (t < T) {
                    key = keys[t];
                    v = vals[t];
                }
                const unsigned occ = __ballot_sync(gmask, key != -1);
                const unsigned mine = (G == 32) ? occ : ((occ >> gshift) & ((1u << G) - 1u));
                const int pos = n_occ + __popc(mine & ((1u << lane) - 1u));
                __syncwarp(gmask);          // every slot of the chunk has been read
                if (key != -1) {
                    keys[pos] = key;
                    vals[pos] = v;
                }
                n_occ += __popc(mine);
                __syncwarp(gmask);
            }
            const int32_t base = c_ptr[row];
            for (int e = lane; e < n_occ; e += G) {
                const int32_t key = keys[e];
                int rank = 0;
                for (int u = 0; u < n_occ; ++u) rank += (keys[u] < key);
                const double v = vals[e];
                c_idx[base + rank] = key;
                c_val[base + rank] = v;
                cnt += (v != 0.0);
            }
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) cnt += __shfl_xor_sync(gmask, cnt, o, G);
        if (lane == 0) row_count[row] = cnt;
        __syncwarp(gmask);
    }
}

template <int G, bool NUMERIC>
static int launch_spgemm(int64_t nrows, const int32_t *ap, const int32_t *ai, const double *av, const int32_t *bp,
                         const int32_t *bi, const double *bv, int log_t, int32_t *row_count, const int32_t *cp,
                         int32_t *ci, double *cv, int32_t *overflow, cudaStream_t st) {
    const int groups = kBlock / G;
    const size_t smem = (size_t)groups * ((size_t)1 << log_t) * (NUMERIC ? 12 : 4);
    if (smem > 200 * 1024) return set_error(MG_ERR_UNSUPPORTED, "mg_spgemm", "hash table does not fit shared memory");
    MG_CHECK_CUDA(cudaFuncSetAttribute(spgemm_kernel<G, NUMERIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t grid = (nrows + groups - 1) / groups;
    const int64_t cap = (int64_t)sm_count() * 32;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    spgemm_kernel<G, NUMERIC><<<(unsigned)grid, kBlock, smem, st>>>(nrows, ap, ai, av, bp, bi, bv, log_t, row_count,
                                                                    cp, ci, cv, overflow);
    MG_CHECK_LAUNCH("spgemm_kernel");
    return MG_OK;
}

template <bool NUMERIC>
static int dispatch_spgemm(int group, int64_t nrows, const int32_t *ap, const int32_t *ai, const double *av,
                           const int32_t *bp, const int32_t *bi, const double *bv, int log_t, int32_t *row_count,
                           const int32_t *cp, int32_t *ci, double *cv, int32_t *overflow, cudaStream_t st) {
    switch (group) {
        case 4: return launch_spgemm<4, NUMERIC>(nrows, ap, ai, av, bp, bi, bv, log_t, row_count, cp, ci, cv, overflow, st);
        case 8: return launch_spgemm<8, NUMERIC>(nrows, ap, ai, av, bp, bi, bv, log_t, row_count, cp, ci, cv, overflow, st);
        case 16: return launch_spgemm<16, NUMERIC>(nrows, ap, ai, av, bp, bi, bv, log_t, row_count, cp, ci, cv, overflow, st);
        case 32: return launch_spgemm<32, NUMERIC>(nrows, ap, ai, av, bp, bi, bv, log_t, row_count, cp, ci, cv, overflow, st);
        default: return set_error(MG_ERR_INVALID, "mg_spgemm", "group size must be 4, 8, 16 or 32");
    }
}

// copy the entries with value != 0 of every row, keeping their order (SciPy's numeric pass drops exact zeros)
__global__ void __launch_bounds__(kBlock)
compact_nonzeros_kernel(int64_t nrows, const int32_t *__restrict__ in_ptr, const int32_t *__restrict__ in_idx,
                        const double *__restrict__ in_val, const int32_t *__restrict__ out_ptr,
                        int32_t *__restrict__ out_idx, double *__restrict__ out_val) {
    const int64_t row = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (row >= nrows) return;
    int32_t o = out_ptr[row];
    for (int32_t p = in_ptr[row]; p < in_ptr[row + 1]; ++p) {
        const double v = in_val[p];
        if (v != 0.0) { out_idx[o] = in_idx[p]; out_val[o] = v; ++o; }
    }
}

// ------------------------------------------------------------------------------------------------------------
// transpose helpers
__global__ void __launch_bounds__(kBlock)
transpose_gather_kernel(int64_t nnz, int64_t nrows, const int32_t *__restrict__ indptr,
                        const double *__restrict__ values, const int32_t *__restrict__ pos_sorted,
                        int32_t *__restrict__ t_indices, double *__restrict__ t_values) {
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t p = (int64_t)blockIdx.x * kBlock + threadIdx.x; p < nnz; p += stride) {
        const int32_t src = pos_sorted[p];
        // row of entry src: last row r with indptr[r] <= src
        int64_t lo = 0, hi = nrows;
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (indptr[mid] <= src) lo = mid; else hi = mid;
        }
        t_indices[p] = (int32_t)lo;
        t_values[p] = values[src];
    }
}

// row pointer of the transposed matrix from the sorted column keys
__global__ void __launch_bounds__(kBlock)
boundaries_kernel(int64_t nnz, int64_t ncols, const int32_t *__restrict__ keys_sorted, int32_t *__restrict__ t_indptr) {
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t p = (int64_t)blockIdx.x * kBlock + threadIdx.x; p <= nnz; p += stride) {
        const int64_t prev = (p == 0) ? -1 : keys_sorted[p - 1];
        const int64_t cur = (p == nnz) ? ncols : keys_sorted[p];
        for (int64_t c = prev + 1; c <= cur; ++c) t_indptr[c] = (int32_t)p;
    }
}

// ------------------------------------------------------------------------------------------------------------
// permutation / SELL build
__global__ void __launch_bounds__(kBlock)
row_lengths_kernel(int64_t n, const int32_t *__restrict__ indptr, const int32_t *__restrict__ perm,
                   int32_t *__restrict__ lens) {
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    const int64_t r = perm ? perm[i] : i;
    lens[i] = indptr[r + 1] - indptr[r];
}

__global__ void __launch_bounds__(kBlock)
permute_copy_kernel(int64_t n, const int32_t *__restrict__ in_ptr, const int32_t *__restrict__ in_idx,
                    const double *__restrict__ in_val, const int32_t *__restrict__ perm,
                    const int32_t *__restrict__ col_iperm, const int32_t *__restrict__ out_ptr,
                    int32_t *__restrict__ out_idx, double *__restrict__ out_val) {
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    const int64_t r = perm ? perm[i] : i;
    int32_t o = out_ptr[i];
    for (int32_t p = in_ptr[r]; p < in_ptr[r + 1]; ++p, ++o) {
        const int32_t c = in_idx[p];
        out_idx[o] = col_iperm ? col_iperm[c] : c;
        out_val[o] = in_val[p];
    }
}

__global__ void __launch_bounds__(kBlock)
invert_perm_kernel(int64_t n, const int32_t *__restrict__ perm, int32_t *__restrict__ iperm) {
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i < n) iperm[perm[i]] = (int32_t)i;
}

// one warp per slice: maximum row length of the slice's 32 rows
__global__ void __launch_bounds__(kBlock)
slice_lengths_kernel(int64_t n, const int32_t *__restrict__ indptr, int32_t *__restrict__ slice_len) {
    const int64_t row = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    int len = (row < n) ? (indptr[row + 1] - indptr[row]) : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
    if ((threadIdx.x & 31) == 0 && (row >> 5) < (n + kSlice - 1) / kSlice) slice_len[row >> 5] = len;
}

__global__ void __launch_bounds__(kBlock)
sell_fill_kernel(int64_t n, const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                 const double *__restrict__ values, const int64_t *__restrict__ slice_ptr,
                 int32_t *__restrict__ cols, double *__restrict__ vals) {
    const int64_t row = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const int64_t nslices = (n + kSlice - 1) / kSlice;
    const int64_t slice = row >> 5;
    if (slice >= nslices) return;
    const int lane = (int)(row & 31);
    const int64_t base = slice_ptr[slice];
    const int len = (int)((slice_ptr[slice + 1] - base) >> 5);
    int32_t p0 = 0, p1 = 0;
    if (row < n) { p0 = indptr[row]; p1 = indptr[row + 1]; }
    const int mylen = p1 - p0;
    const int32_t padcol = (mylen > 0) ? indices[p1 - 1] : 0;
    for (int k = 0; k < len; ++k) {
        const int64_t dst = base + (int64_t)k * kSlice + lane;
        if (k < mylen) { cols[dst] = indices[p0 + k]; vals[dst] = values[p0 + k]; }
        else { cols[dst] = padcol; vals[dst] = 0.0; }
    }
}

__global__ void __launch_bounds__(kBlock)
dinv_kernel(int64_t n, const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
            const double *__restrict__ values, const int32_t *__restrict__ perm, double *__restrict__ dinv) {
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    const int64_t r = perm ? perm[i] : i;
    double d = 0.0;
    for (int32_t p = indptr[r]; p < indptr[r + 1]; ++p)
        if (indices[p] == r) d = __dadd_rn(d, values[p]);
    dinv[i] = __ddiv_rn(1.0, d);
}

}  // namespace mgb

using namespace mgb;

extern "C" {

// ---- scans -------------------------------------------------------------------------------------------------
int64_t mg_scan_workspace_size(int64_t n) {
    size_t a = 0, b = 0, c = 0;
    cub::TransformInputIterator<int64_t, ToI64, const int32_t *> it64((const int32_t *)nullptr, ToI64());
    cub::TransformInputIterator<int64_t, Times32, const int32_t *> it32((const int32_t *)nullptr, Times32());
    cub::DeviceScan::InclusiveSum(nullptr, a, (const int32_t *)nullptr, (int32_t *)nullptr, n);
    cub::DeviceReduce::Sum(nullptr, b, it64, (int64_t *)nullptr, n);
    cub::DeviceScan::InclusiveSum(nullptr, c, it32, (int64_t *)nullptr, n);
    size_t d = 0;
    cub::DeviceReduce::Max(nullptr, d, (const int32_t *)nullptr, (int32_t *)nullptr, n);
    size_t m = a > b ? a : b;
    if (c > m) m = c;
    if (d > m) m = d;
    return (int64_t)m + 1024;
}

/* d_out[0] = 0, d_out[i+1] = d_in[0] + ... + d_in[i]  (int32 row pointer); *d_total = the same sum in int64.
 * Fails with MG_ERR_OVERFLOW (after synchronising) if the total does not fit int32. */
int mg_exclusive_scan_i32(int64_t n, const int32_t *d_in, int32_t *d_out, int64_t *d_total, void *d_temp,
                          int64_t temp_bytes, void *stream) {
    MG_REQUIRE(n >= 0 && d_out && d_total && d_temp, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    MG_CHECK_CUDA(cudaMemsetAsync(d_out, 0, sizeof(int32_t), st));
    MG_CHECK_CUDA(cudaMemsetAsync(d_total, 0, sizeof(int64_t), st));
    int64_t total = 0;
    if (n > 0) {
        size_t bytes = (size_t)temp_bytes;
        MG_CHECK_CUDA(cub::DeviceScan::InclusiveSum(d_temp, bytes, d_in, d_out + 1, n, st));
        cub::TransformInputIterator<int64_t, ToI64, const int32_t *> it64(d_in, ToI64());
        bytes = (size_t)temp_bytes;
        MG_CHECK_CUDA(cub::DeviceReduce::Sum(d_temp, bytes, it64, d_total, n, st));
        g_launch_count += 2;
        MG_CHECK_CUDA(cudaMemcpyAsync(&total, d_total, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        MG_CHECK_CUDA(cudaStreamSynchronize(st));
    }
    if (total >= 2147483647LL) return set_error(MG_ERR_OVERFLOW, "mg_exclusive_scan_i32", "nnz does not fit int32");
    return MG_OK;
}

// ---- SpGEMM ---------------------------------------------------------------------------------------------------
/* symbolic pass: d_row_count[i] = number of distinct columns of row i of A*B.  log2_table = log2 of the per-row
 * hash-table size (must exceed the largest row); *d_overflow is set to 1 if a table filled up. */
int mg_spgemm_symbolic(int64_t nrows, const int32_t *d_a_indptr, const int32_t *d_a_indices,
                       const int32_t *d_b_indptr, const int32_t *d_b_indices, int group, int log2_table,
                       int32_t *d_row_count, int32_t *d_overflow, void *stream) {
    MG_REQUIRE(nrows >= 0 && log2_table >= 2 && log2_table <= 14, "bad size");
    if (nrows == 0) return MG_OK;
    return dispatch_spgemm<false>(group, nrows, d_a_indptr, d_a_indices, nullptr, d_b_indptr, d_b_indices, nullptr,
                                  log2_table, d_row_count, nullptr, nullptr, nullptr, d_overflow,
                                  (cudaStream_t)stream);
}

/* numeric pass: fills row i of C at [d_c_indptr[i], d_c_indptr[i+1]) with sorted columns and the values
 * accumulated in SciPy's order; d_row_nonzeros[i] = number of entries of the row whose value is not exactly 0. */
int mg_spgemm_numeric(int64_t nrows, const int32_t *d_a_indptr, const int32_t *d_a_indices, const double *d_a_values,
                      const int32_t *d_b_indptr, const int32_t *d_b_indices, const double *d_b_values, int group,
                      int log2_table, const int32_t *d_c_indptr, int32_t *d_c_indices, double *d_c_values,
                      int32_t *d_row_nonzeros, int32_t *d_overflow, void *stream) {
    MG_REQUIRE(nrows >= 0 && log2_table >= 2 && log2_table <= 14, "bad size");
    if (nrows == 0) return MG_OK;
    return dispatch_spgemm<true>(group, nrows, d_a_indptr, d_a_indices, d_a_values, d_b_indptr, d_b_indices,
                                 d_b_values, log2_table, d_row_nonzeros, d_c_indptr, d_c_indices, d_c_values,
                                 d_overflow, (cudaStream_t)stream);
}

/* drop the exact zeros (SciPy csr_matmat does so in its numeric pass): d_out_indptr = scan of the nonzero counts */
int mg_csr_compact_nonzeros(int64_t nrows, const int32_t *d_in_indptr, const int32_t *d_in_indices,
                            const double *d_in_values, const int32_t *d_out_indptr, int32_t *d_out_indices,
                            double *d_out_values, void *stream) {
    MG_REQUIRE(nrows >= 0, "negative size");
    if (nrows == 0) return MG_OK;
    compact_nonzeros_kernel<<<(unsigned)((nrows + kBlock - 1) / kBlock), kBlock, 0, (cudaStream_t)stream>>>(
        nrows, d_in_indptr, d_in_indices, d_in_values, d_out_indptr, d_out_indices, d_out_values);
    MG_CHECK_LAUNCH("compact_nonzeros");
    return MG_OK;
}

// ---- transpose / stable sort -------------------------------------------------------------------------------------
int64_t mg_sort_workspace_size(int64_t n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int32_t *)nullptr, (int32_t *)nullptr,
                                    (const int32_t *)nullptr, (int32_t *)nullptr, n);
    return (int64_t)bytes + 256;
}

/* stable sort of (key, index) pairs by key: d_perm_out[i] = original position of the i-th smallest key */
int mg_stable_argsort_i32(int64_t n, const int32_t *d_keys, int32_t *d_keys_sorted, int32_t *d_perm_out,
                          int32_t *d_iota_tmp, int key_bits, void *d_temp, int64_t temp_bytes, void *stream) {
    MG_REQUIRE(n >= 0 && key_bits >= 1 && key_bits <= 32, "bad argument");
    if (n == 0) return MG_OK;
    cudaStream_t st = (cudaStream_t)stream;
    iota_kernel<<<grid_for(n), kBlock, 0, st>>>(n, d_iota_tmp);
    MG_CHECK_LAUNCH("iota");
    size_t bytes = (size_t)temp_bytes;
    MG_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(d_temp, bytes, d_keys, d_keys_sorted, d_iota_tmp, d_perm_out, n, 0,
                                                  key_bits, st));
    ++g_launch_count;
    return MG_OK;
}

int64_t mg_csr_transpose_workspace(int64_t nnz) { return 3 * nnz * (int64_t)sizeof(int32_t) + mg_sort_workspace_size(nnz) + 256; }

/* CSR of A^T (ncols x nrows) whose row entries are in ascending original-row order, i.e. the order in which SciPy's
 * csc kernels visit them (Multigrid.py:93 `i.T @ res`, and the transposed operands of the Galerkin product). */
int mg_csr_transpose(int64_t nrows, int64_t ncols, int64_t nnz, const int32_t *d_indptr, const int32_t *d_indices,
                     const double *d_values, int32_t *d_t_indptr, int32_t *d_t_indices, double *d_t_values,
                     void *d_work, void *stream) {
    MG_REQUIRE(nrows >= 0 && ncols >= 0 && nnz >= 0 && d_t_indptr, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (nnz == 0) {
        MG_CHECK_CUDA(cudaMemsetAsync(d_t_indptr, 0, (ncols + 1) * sizeof(int32_t), st));
        return MG_OK;
    }
    int32_t *iota = (int32_t *)d_work;
    int32_t *keys_sorted = iota + nnz;
    int32_t *pos_sorted = keys_sorted + nnz;
    void *temp = (void *)(pos_sorted + nnz);
    int bits = 1;
    while (bits < 32 && ((int64_t)1 << bits) < ncols) ++bits;
    int rc = mg_stable_argsort_i32(nnz, d_indices, keys_sorted, pos_sorted, iota, bits, temp, mg_sort_workspace_size(nnz), stream);
    if (rc) return rc;
    transpose_gather_kernel<<<grid_for(nnz), kBlock, 0, st>>>(nnz, nrows, d_indptr, d_values, pos_sorted, d_t_indices, d_t_values);
    MG_CHECK_LAUNCH("transpose_gather");
    boundaries_kernel<<<grid_for(nnz + 1), kBlock, 0, st>>>(nnz, ncols, keys_sorted, d_t_indptr);
    MG_CHECK_LAUNCH("boundaries");
    return MG_OK;
}

// ---- permutation / SELL ----------------------------------------------------------------------------------------------
int mg_invert_permutation(int64_t n, const int32_t *d_perm, int32_t *d_iperm, void *stream) {
    MG_REQUIRE(n >= 0, "negative size");
    if (n == 0) return MG_OK;
    invert_perm_kernel<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, (cudaStream_t)stream>>>(n, d_perm, d_iperm);
    MG_CHECK_LAUNCH("invert_perm");
    return MG_OK;
}

/* d_lens[i] = length of row perm[i] (perm may be NULL = identity) */
int mg_csr_row_lengths(int64_t n, const int32_t *d_indptr, const int32_t *d_perm, int32_t *d_lens, void *stream) {
    MG_REQUIRE(n >= 0, "negative size");
    if (n == 0) return MG_OK;
    row_lengths_kernel<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, (cudaStream_t)stream>>>(n, d_indptr, d_perm, d_lens);
    MG_CHECK_LAUNCH("row_lengths");
    return MG_OK;
}

/* new row i = old row perm[i]; column j relabelled to col_iperm[j]; entries keep their order inside the row */
int mg_csr_permute(int64_t n, const int32_t *d_in_indptr, const int32_t *d_in_indices, const double *d_in_values,
                   const int32_t *d_perm, const int32_t *d_col_iperm, const int32_t *d_out_indptr,
                   int32_t *d_out_indices, double *d_out_values, void *stream) {
    MG_REQUIRE(n >= 0, "negative size");
    if (n == 0) return MG_OK;
    permute_copy_kernel<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, (cudaStream_t)stream>>>(
        n, d_in_indptr, d_in_indices, d_in_values, d_perm, d_col_iperm, d_out_indptr, d_out_indices, d_out_values);
    MG_CHECK_LAUNCH("permute_copy");
    return MG_OK;
}

__global__ void __launch_bounds__(kBlock) fill_i32_kernel(int64_t n, int32_t v, int32_t *out) {
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i < n) out[i] = v;
}

/* SELL-32 build, step 1: d_slice_ptr[nslices+1] (int64 entry offsets); on the host: padded entry count, longest
 * slice, and the uniform length (all slices are padded to the longest one when that costs <= 3 % extra entries -- 25 % for rows of <= 2 entries;
 * 0 otherwise).  d_slice_len_tmp: nslices int32.  Synchronises the stream. */
int mg_sell_layout(int64_t n, const int32_t *d_indptr, int32_t *d_slice_len_tmp, int64_t *d_slice_ptr,
                   int64_t *h_total_out, int64_t *h_max_len_out, int64_t *h_uniform_len_out, void *d_temp,
                   int64_t temp_bytes, void *stream) {
    MG_REQUIRE(n >= 0 && d_slice_ptr && h_total_out && h_max_len_out && h_uniform_len_out && temp_bytes > 512,
               "bad argument");
    *h_max_len_out = 0;
    *h_uniform_len_out = 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nslices = (n + kSlice - 1) / kSlice;
    MG_CHECK_CUDA(cudaMemsetAsync(d_slice_ptr, 0, sizeof(int64_t), st));
    *h_total_out = 0;
    if (nslices == 0) return MG_OK;
    slice_lengths_kernel<<<(unsigned)((nslices * kSlice + kBlock - 1) / kBlock), kBlock, 0, st>>>(n, d_indptr, d_slice_len_tmp);
    MG_CHECK_LAUNCH("slice_lengths");
    // the first 256 bytes of the workspace hold the reduction results, the rest is CUB scratch
    int32_t *d_max = (int32_t *)d_temp;
    int64_t *d_sum = (int64_t *)((char *)d_temp + 64);
    void *cub_temp = (void *)((char *)d_temp + 256);
    size_t bytes = (size_t)temp_bytes - 256;
    MG_CHECK_CUDA(cub::DeviceReduce::Max(cub_temp, bytes, d_slice_len_tmp, d_max, nslices, st));
    cub::TransformInputIterator<int64_t, ToI64, const int32_t *> it64(d_slice_len_tmp, ToI64());
    bytes = (size_t)temp_bytes - 256;
    MG_CHECK_CUDA(cub::DeviceReduce::Sum(cub_temp, bytes, it64, d_sum, nslices, st));
    int32_t h_max = 0;
    int64_t h_sum = 0;
    MG_CHECK_CUDA(cudaMemcpyAsync(&h_max, d_max, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    MG_CHECK_CUDA(cudaMemcpyAsync(&h_sum, d_sum, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    MG_CHECK_CUDA(cudaStreamSynchronize(st));
    // rows of one or two entries (linear transfer operators): a quarter more entries is cheaper than a dependent load of
    // the slice pointer in front of every row (sell_short_kernel)
    const int64_t allow = h_max <= 2 ? 125 : 103;
    const bool uniform = h_max > 0 && nslices * (int64_t)h_max * 100 <= allow * h_sum;
    if (uniform) {
        fill_i32_kernel<<<(unsigned)((nslices + kBlock - 1) / kBlock), kBlock, 0, st>>>(nslices, h_max, d_slice_len_tmp);
        MG_CHECK_LAUNCH("fill_i32");
    }
    cub::TransformInputIterator<int64_t, Times32, const int32_t *> it(d_slice_len_tmp, Times32());
    bytes = (size_t)temp_bytes - 256;
    MG_CHECK_CUDA(cub::DeviceScan::InclusiveSum(cub_temp, bytes, it, d_slice_ptr + 1, nslices, st));
    g_launch_count += 3;
    MG_CHECK_CUDA(cudaMemcpyAsync(h_total_out, d_slice_ptr + nslices, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    MG_CHECK_CUDA(cudaStreamSynchronize(st));
    *h_max_len_out = h_max;
    *h_uniform_len_out = uniform ? h_max : 0;
    return MG_OK;
}

/* SELL-32 build, step 2: fill cols / vals (padding: value 0, column = the row's last column) */
int mg_sell_fill(int64_t n, const int32_t *d_indptr, const int32_t *d_indices, const double *d_values,
                 const int64_t *d_slice_ptr, int32_t *d_cols, double *d_vals, void *stream) {
    MG_REQUIRE(n >= 0, "negative size");
    if (n == 0) return MG_OK;
    const int64_t nslices = (n + kSlice - 1) / kSlice;
    sell_fill_kernel<<<(unsigned)((nslices * kSlice + kBlock - 1) / kBlock), kBlock, 0, (cudaStream_t)stream>>>(
        n, d_indptr, d_indices, d_values, d_slice_ptr, d_cols, d_vals);
    MG_CHECK_LAUNCH("sell_fill");
    return MG_OK;
}

/* d_dinv[i] = 1 / A[perm[i], perm[i]] */
int mg_extract_dinv(int64_t n, const int32_t *d_indptr, const int32_t *d_indices, const double *d_values,
                    const int32_t *d_perm, double *d_dinv, void *stream) {
    MG_REQUIRE(n >= 0, "negative size");
    if (n == 0) return MG_OK;
    dinv_kernel<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, (cudaStream_t)stream>>>(n, d_indptr, d_indices, d_values, d_perm, d_dinv);
    MG_CHECK_LAUNCH("dinv");
    return MG_OK;
}

}  // extern "C"
