// comm.cu -- halo exchange between the row blocks of a distributed level over peer-mapped memory (NVLink/NVSwitch).
//
// There is no reference counterpart (the reference is single-process); the contract is SURVEY 8e.
// Mechanism: every rank owns a communication arena allocated with cudaMalloc and exported with CUDA IPC; peers map
// it and WRITE the boundary values the owner needs straight into its staging area (st.global over NVLink), then
// publish a sequence number with a system-scope release store.  The owner's consumer kernel spins on its local flag
// with acquire loads and unpacks the staging area into the halo part of its level vector.  Both kernels are ordinary
// stream work, so a whole V-cycle including its exchanges is captured in one CUDA graph; the sequence number expected
// by each exchange site is read from a device counter that the graph advances at its end.
#include "common.cuh"

namespace mgb {

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// gather src[idx[i]] (idx == NULL: src[i]) into the peer's staging area, then publish seq_base[0] + site.
// The last CTA to finish (counted in *done, which it resets) performs the release store.
__global__ void __launch_bounds__(kBlock)
halo_push_kernel(const double *__restrict__ src, const int32_t *__restrict__ idx, int64_t count,
                 double *__restrict__ peer_dst, unsigned long long *peer_flag,
                 const unsigned long long *__restrict__ seq_base, unsigned long long site, unsigned int *done) {
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < count; i += stride)
        peer_dst[i] = idx ? src[idx[i]] : src[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(done, 1u);
        if (prev == gridDim.x - 1) {
            *done = 0;
            __threadfence_system();
            st_release_sys(peer_flag, seq_base[0] + site);
        }
    }
}

// wait until *flag >= seq_base[0] + site (written by the peer), then dst[i] = staging[i]
__global__ void __launch_bounds__(kBlock)
halo_wait_unpack_kernel(const unsigned long long *flag, const unsigned long long *__restrict__ seq_base,
                        unsigned long long site, const double *staging, double *__restrict__ dst, int64_t count) {
    if (threadIdx.x == 0) {
        const unsigned long long want = seq_base[0] + site;
        while (ld_acquire_sys(flag) < want) { __nanosleep(64); }
    }
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < count; i += stride)
        dst[i] = __ldcv(staging + i);      // volatile load: never served from a stale L1 line
}

__global__ void seq_advance_kernel(unsigned long long *seq_base, unsigned long long delta) {
    if (threadIdx.x == 0 && blockIdx.x == 0) seq_base[0] += delta;
}

// column relabelling of a row block: global column -> local [owned (permuted) | halo] index through a lookup table:
// slot_of[c] = halo slot (>= 0) or -1
__global__ void __launch_bounds__(kBlock)
remap_cols_table_kernel(int64_t nnz, const int32_t *__restrict__ cols_in, int64_t c0, int64_t c1,
                        const int32_t *__restrict__ own_iperm, int64_t n_own,
                        const int32_t *__restrict__ slot_of, int32_t *__restrict__ cols_out,
                        int32_t *__restrict__ missing) {
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t p = (int64_t)blockIdx.x * kBlock + threadIdx.x; p < nnz; p += stride) {
        const int64_t c = cols_in[p];
        if (c >= c0 && c < c1) {
            cols_out[p] = own_iperm ? own_iperm[c - c0] : (int32_t)(c - c0);
        } else {
            const int32_t s = slot_of[c];
            if (s < 0) atomicExch(missing, 1);
            cols_out[p] = (int32_t)(n_own + s);
        }
    }
}

static inline unsigned small_grid(int64_t n) {
    int64_t g = (n + kBlock - 1) / kBlock;
    if (g > 64) g = 64;          // halos are small: a few CTAs keep the barrier cheap
    if (g < 1) g = 1;
    return (unsigned)g;
}

}  // namespace mgb

using namespace mgb;

extern "C" {

/* communication arena: cudaMalloc'ed (IPC needs a whole allocation), zero-initialised */
int mg_comm_alloc(int64_t bytes, void **d_ptr_out) {
    MG_REQUIRE(bytes > 0 && d_ptr_out, "bad argument");
    MG_CHECK_CUDA(cudaMalloc(d_ptr_out, (size_t)bytes));
    MG_CHECK_CUDA(cudaMemset(*d_ptr_out, 0, (size_t)bytes));
    return MG_OK;
}
int mg_comm_free(void *d_ptr) {
    if (d_ptr) MG_CHECK_CUDA(cudaFree(d_ptr));
    return MG_OK;
}
/* 64-byte CUDA IPC handle of an arena (to be sent to the peers) */
int mg_comm_export(void *d_ptr, unsigned char *h_handle64) {
    MG_REQUIRE(d_ptr && h_handle64, "null argument");
    cudaIpcMemHandle_t h;
    MG_CHECK_CUDA(cudaIpcGetMemHandle(&h, d_ptr));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(h_handle64, &h, 64);
    return MG_OK;
}
/* map a peer's arena into this process (enables peer access lazily) */
int mg_comm_import(const unsigned char *h_handle64, void **d_peer_ptr_out) {
    MG_REQUIRE(h_handle64 && d_peer_ptr_out, "null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, h_handle64, 64);
    MG_CHECK_CUDA(cudaIpcOpenMemHandle(d_peer_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return MG_OK;
}
int mg_comm_unmap(void *d_peer_ptr) {
    if (d_peer_ptr) MG_CHECK_CUDA(cudaIpcCloseMemHandle(d_peer_ptr));
    return MG_OK;
}

/* peer_dst[i] = src[idx[i]] (idx NULL: src[i]) for i < count, then *peer_flag = *d_seq_base + site (release, system
 * scope).  d_done: one zero-initialised uint32 of scratch per concurrent push. */
int mg_halo_push(const double *d_src, const int32_t *d_idx, int64_t count, double *d_peer_dst, void *d_peer_flag,
                 const void *d_seq_base, int64_t site, void *d_done, void *stream) {
    MG_REQUIRE(count >= 0 && d_peer_flag && d_seq_base && d_done, "null argument");
    halo_push_kernel<<<small_grid(count), kBlock, 0, (cudaStream_t)stream>>>(
        d_src, d_idx, count, d_peer_dst, (unsigned long long *)d_peer_flag, (const unsigned long long *)d_seq_base,
        (unsigned long long)site, (unsigned int *)d_done);
    MG_CHECK_LAUNCH("halo_push");
    return MG_OK;
}
/* spin until *d_flag >= *d_seq_base + site, then d_dst[i] = d_staging[i] */
int mg_halo_wait_unpack(const void *d_flag, const void *d_seq_base, int64_t site, const double *d_staging,
                        double *d_dst, int64_t count, void *stream) {
    MG_REQUIRE(count >= 0 && d_flag && d_seq_base, "null argument");
    halo_wait_unpack_kernel<<<small_grid(count), kBlock, 0, (cudaStream_t)stream>>>(
        (const unsigned long long *)d_flag, (const unsigned long long *)d_seq_base, (unsigned long long)site,
        d_staging, d_dst, count);
    MG_CHECK_LAUNCH("halo_wait_unpack");
    return MG_OK;
}
int mg_seq_advance(void *d_seq_base, int64_t delta, void *stream) {
    MG_REQUIRE(d_seq_base, "null argument");
    seq_advance_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((unsigned long long *)d_seq_base, (unsigned long long)delta);
    MG_CHECK_LAUNCH("seq_advance");
    return MG_OK;
}
/* relabel the columns of a row block: c in [c0,c1) -> d_own_iperm[c-c0] (NULL: c-c0), otherwise n_own + d_slot_of[c] */
int mg_csr_remap_cols(int64_t nnz, const int32_t *d_cols_in, int64_t c0, int64_t c1, const int32_t *d_own_iperm,
                      int64_t n_own, const int32_t *d_slot_of, int32_t *d_cols_out, int32_t *d_missing, void *stream) {
    MG_REQUIRE(nnz >= 0 && d_slot_of && d_missing, "null argument");
    if (nnz == 0) return MG_OK;
    int64_t g = (nnz + kBlock - 1) / kBlock;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (g > cap) g = cap;
    remap_cols_table_kernel<<<(unsigned)g, kBlock, 0, (cudaStream_t)stream>>>(nnz, d_cols_in, c0, c1, d_own_iperm, n_own,
                                                                             d_slot_of, d_cols_out, d_missing);
    MG_CHECK_LAUNCH("remap_cols");
    return MG_OK;
}

}  // extern "C"
