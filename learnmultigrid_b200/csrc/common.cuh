// common.cuh -- shared helpers of libmgb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/mgb200.h"

namespace mgb {

extern thread_local char g_last_error[512];
extern thread_local int64_t g_launch_count;

inline int set_error(int code, const char *what, const char *detail) {
    snprintf(g_last_error, sizeof(g_last_error), "%s: %s", what, detail ? detail : "");
    return code;
}

#define MG_CHECK_CUDA(expr)                                                                   \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) return mgb::set_error(MG_ERR_CUDA, #expr, cudaGetErrorString(_e)); \
    } while (0)

#define MG_CHECK_LAUNCH(name)                                                                 \
    do {                                                                                      \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess) return mgb::set_error(MG_ERR_CUDA, name, cudaGetErrorString(_e)); \
        ++mgb::g_launch_count;                                                                \
    } while (0)

#define MG_REQUIRE(cond, msg)                                                                 \
    do {                                                                                      \
        if (!(cond)) return mgb::set_error(MG_ERR_INVALID, __func__, msg);                    \
    } while (0)

// GS_RES / GS_NORM: colour sweep that also yields the residual / the squared residual norm of the swept rows
// SPMV_DOT: y = A x and per-CTA partial sums of aux . y (the p . A p of conjugate gradients from the SpMV's registers)
enum SellMode { SPMV = 0, RESID = 1, RESNORM = 2, JACOBI = 3, GS = 4, PROLONG = 5, GS_RES = 6, GS_NORM = 7, SPMV_DOT = 8 };

constexpr int kSlice = 32;      // SELL slice height = one warp
constexpr int kBlock = 256;     // threads per CTA of the streaming kernels

// IEEE multiply / add that the compiler may not contract into an FMA: the CPU oracle (gcc
// -ffp-contract=off, SciPy, PyAMG) rounds the product and the sum separately and we match it bit for bit.
__device__ __forceinline__ double mul_add_unfused(double acc, double a, double b) {
    return __dadd_rn(acc, __dmul_rn(a, b));
}

// streaming (read-once) loads of the matrix arrays: do not pollute L1, which we want for the x gathers
__device__ __forceinline__ double ld_stream(const double *p) { return __ldcs(p); }
__device__ __forceinline__ int32_t ld_stream(const int32_t *p) { return __ldcs(p); }
__device__ __forceinline__ unsigned char ld_stream(const unsigned char *p) { return __ldcs(p); }

// ---- programmatic dependent launch (PDL) --------------------------------------------------------------------------
// The V-cycle is a long chain of short dependent kernels (60-150 per cycle, many of a few microseconds on the coarse
// levels and between exchange sites).  Every kernel of the chain starts with pdl_prologue(): it waits until the
// preceding grid has completed and flushed (griddepcontrol.wait) and then allows the NEXT kernel of the stream to
// be scheduled early (griddepcontrol.launch_dependents), so that its launch latency overlaps this kernel's
// execution; launch_k() marks the launch accordingly.  Since every kernel of the chain waits unconditionally, a grid
// can never complete before all of its predecessors have, which keeps the ordering transitive.  Stream capture turns
// these launches into programmatic graph edges.  mg_set_pdl(0) falls back to plain stream order.
extern int g_pdl;
__device__ __forceinline__ void pdl_prologue() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), unsigned grid, unsigned block, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(block, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);     // errors are picked up by MG_CHECK_LAUNCH
}

inline int sm_count() {
    static int cached = 0;
    if (!cached) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
        if (cached <= 0) cached = 148;
    }
    return cached;
}

// deterministic block reduction (fixed tree), result valid in thread 0
template <int BLOCK>
__device__ __forceinline__ double block_sum(double v) {
    __shared__ double sh[BLOCK / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        v = (l < BLOCK / 32) ? sh[l] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    }
    __syncthreads();
    return v;
}

// The same sum as the LAST thing a kernel does: only warp 0 waits.  The other warps park their warp sums in shared
// memory, ARRIVE at a named barrier and are done (barrier.cta.arrive / barrier.cta.sync, the producer-consumer pattern of
// the PTX ISA): no warp stalls on the slowest one, which matters when the block holds few warps per SM to hide it
// (ncu on the sweep with a fused norm: 3.3 barrier stalls per issue, 5.5 instead of 6.8 TB/s).  Same tree, same bits.
// Every thread of the block must call it, exactly once, and nothing may follow that needs the other warps.
template <int BLOCK>
__device__ __forceinline__ double block_sum_last(double v) {
    __shared__ double sh[BLOCK / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = v;
    if (w != 0) {
        asm volatile("barrier.cta.arrive 1, %0;" ::"n"(BLOCK) : "memory");
        return 0.0;
    }
    asm volatile("barrier.cta.sync 1, %0;" ::"n"(BLOCK) : "memory");
    v = (l < BLOCK / 32) ? sh[l] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace mgb
