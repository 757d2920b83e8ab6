// valdict.cu -- value dictionary of a sparse matrix (setup time).
//
// Finite-element operators on uniform meshes and interpolation operators hold very few DISTINCT values: the 5-point
// Laplacian {4, -1, 1, 0}, linear interpolation {1, 0.5, 0}, their Galerkin products a few dozen.  When a matrix has at
// most 256 distinct bit patterns its values can be stored as one byte per entry plus a 2 KB table ("value indexing",
// Kourtis et al., CF'08): the streaming kernels then read 1 instead of 8 bytes per entry and look the value up in L1.
// Lossless -- the table holds the exact doubles -- so every result keeps its bits.  Matrices with more distinct values
// (variable coefficients, quasi-L2 / NN transfers) are detected after a few thousand entries and left alone.
#include "common.cuh"

namespace mgb {

constexpr int kDictSlots = 4096;                        // open-addressing hash table over the value bit patterns
constexpr int kDictMax = 256;                           // distinct values a byte can index
constexpr unsigned long long kDictEmpty = 0xffffffffffffffffull;   // a NaN payload no assembled matrix contains

__device__ __forceinline__ unsigned dict_hash(unsigned long long b) {
    b ^= b >> 33;
    b *= 0xff51afd7ed558ccdull;
    b ^= b >> 29;
    return (unsigned)b & (kDictSlots - 1);
}

// state: keys[kDictSlots] | count | overflow
__global__ void __launch_bounds__(kBlock)
dict_collect_kernel(int64_t n, const double *__restrict__ vals, unsigned long long *__restrict__ keys,
                    int *__restrict__ count, int *__restrict__ overflow) {
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) {
        if (*(volatile int *)overflow) return;
        const unsigned long long b = (unsigned long long)__double_as_longlong(vals[i]);
        if (b == kDictEmpty) { atomicExch(overflow, 1); return; }
        unsigned s = dict_hash(b);
        for (int probe = 0; probe < kDictSlots; ++probe, s = (s + 1) & (kDictSlots - 1)) {
            const unsigned long long cur = keys[s];
            if (cur == b) break;
            if (cur == kDictEmpty) {
                const unsigned long long old = atomicCAS(keys + s, kDictEmpty, b);
                if (old == kDictEmpty) {
                    if (atomicAdd(count, 1) >= kDictMax) atomicExch(overflow, 1);
                    break;
                }
                if (old == b) break;
            }
        }
    }
}

// one CTA: number the occupied slots (slot order), write the table
__global__ void __launch_bounds__(1024)
dict_number_kernel(const unsigned long long *__restrict__ keys, int *__restrict__ slot_index, double *__restrict__ table) {
    __shared__ int offs[1024];
    constexpr int per = kDictSlots / 1024;
    int mine = 0;
    for (int k = 0; k < per; ++k) mine += keys[threadIdx.x * per + k] != kDictEmpty;
    offs[threadIdx.x] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int t = 0; t < 1024; ++t) { const int c = offs[t]; offs[t] = run; run += c; }
    }
    __syncthreads();
    int idx = offs[threadIdx.x];
    for (int k = 0; k < per; ++k) {
        const int s = threadIdx.x * per + k;
        if (keys[s] != kDictEmpty) {
            slot_index[s] = idx;
            if (idx < kDictMax) table[idx] = __longlong_as_double((long long)keys[s]);
            ++idx;
        } else {
            slot_index[s] = -1;
        }
    }
}

__global__ void __launch_bounds__(kBlock)
dict_encode_kernel(int64_t n, const double *__restrict__ vals, const unsigned long long *__restrict__ keys,
                   const int *__restrict__ slot_index, unsigned char *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) {
        const unsigned long long b = (unsigned long long)__double_as_longlong(vals[i]);
        unsigned s = dict_hash(b);
        while (keys[s] != b) s = (s + 1) & (kDictSlots - 1);
        out[i] = (unsigned char)slot_index[s];
    }
}

}  // namespace mgb

using namespace mgb;

extern "C" {

int64_t mg_value_dict_workspace(void) { return kDictSlots * 8 + kDictSlots * 4 + 64; }

/* d_table[256], d_index[n]; d_work: mg_value_dict_workspace() bytes.  *h_count = number of distinct values, or -1 if
 * there are more than 256 (d_index / d_table are then not written).  Synchronises the stream. */
int mg_value_dict_build(int64_t n, const double *d_vals, unsigned char *d_index, double *d_table, void *d_work,
                        int *h_count, void *stream) {
    MG_REQUIRE(n >= 0 && d_work && h_count && (n == 0 || (d_vals && d_index && d_table)), "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    *h_count = 0;
    if (n == 0) return MG_OK;
    unsigned long long *keys = (unsigned long long *)d_work;
    int *slot_index = (int *)(keys + kDictSlots);
    int *state = slot_index + kDictSlots;              // count, overflow
    MG_CHECK_CUDA(cudaMemsetAsync(keys, 0xff, kDictSlots * 8, st));
    MG_CHECK_CUDA(cudaMemsetAsync(state, 0, 16, st));
    int64_t g = (n + kBlock - 1) / kBlock;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (g > cap) g = cap;
    // the first CTA-waves see most of the distinct values; a matrix with many of them overflows almost at once and the
    // remaining threads return at their first look at the flag
    dict_collect_kernel<<<(unsigned)g, kBlock, 0, st>>>(n, d_vals, keys, state, state + 1);
    MG_CHECK_LAUNCH("dict_collect");
    int host[2] = {0, 0};
    MG_CHECK_CUDA(cudaMemcpyAsync(host, state, sizeof(host), cudaMemcpyDeviceToHost, st));
    MG_CHECK_CUDA(cudaStreamSynchronize(st));
    if (host[1] || host[0] > kDictMax) {
        *h_count = -1;
        return MG_OK;
    }
    dict_number_kernel<<<1, 1024, 0, st>>>(keys, slot_index, d_table);
    MG_CHECK_LAUNCH("dict_number");
    dict_encode_kernel<<<(unsigned)g, kBlock, 0, st>>>(n, d_vals, keys, slot_index, d_index);
    MG_CHECK_LAUNCH("dict_encode");
    MG_CHECK_CUDA(cudaStreamSynchronize(st));
    *h_count = host[0];
    return MG_OK;
}

}  // extern "C"
