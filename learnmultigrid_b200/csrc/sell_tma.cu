// sell_tma.cu -- bulk-async (TMA) staged variant of the SELL-32 streaming kernels for large levels.
//
// ncu on the plain kernel (profiles/r01_ncu_full_sell_gs_v1.csv) shows it latency-bound: every thread walks
// slice_ptr -> cols/vals -> x, three dependent global loads with ~60 B per thread in flight, 84 % of the copy
// bandwidth and 38 long-scoreboard stalls per issue.  Here the matrix stream is decoupled from the threads:
// each warp owns a private ring of D shared-memory stages; one lane issues cp.async.bulk (UBLKCP) copies of the
// cols/vals of G consecutive slices (contiguous in SELL) per stage, completion is tracked by an mbarrier per
// stage, and the warp consumes stage i while stages i+1 .. i+D-1 are in flight.  Persistent grid
// (#SM x resident CTAs), stages handed to warps round-robin.  Row arithmetic is IDENTICAL to sell_kernels.cu
// (same order, no FMA), so results stay bit-identical to the CPU oracle.
#include "common.cuh"

namespace mgb {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

struct TmaArgs {
    const int64_t *__restrict__ slice_ptr;
    const int32_t *__restrict__ cols;
    const double *__restrict__ vals;
    int64_t row_begin, row_end;   // rows this launch updates
    int64_t s_first, s_end;       // slices [s_first, s_end) cover them
    int64_t ntasks;               // ceil((s_end - s_first) / G)
    int32_t depth;                // ring stages per warp
    int32_t cap;                  // entries per stage slot (>= 32 * G * max slice length)
    int32_t warp_bytes;           // shared memory per warp
};

template <int MODE, int G>
__global__ void __launch_bounds__(256)
sell_tma_kernel(TmaArgs A, const double *x, const double *__restrict__ b, const double *aux, double *y,
                double omega, double *__restrict__ partials) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nw = blockDim.x >> 5;
    unsigned char *wbase = smem_raw + (size_t)warp * A.warp_bytes;
    double *s_vals = reinterpret_cast<double *>(wbase);
    int32_t *s_cols = reinterpret_cast<int32_t *>(wbase + (size_t)A.depth * A.cap * sizeof(double));
    uint64_t *bars = reinterpret_cast<uint64_t *>(wbase + (size_t)A.depth * A.cap * 12);
    const int D = A.depth;
    if (lane == 0) {
        for (int d = 0; d < D; ++d) mbar_init(&bars[d], 1);
        fence_mbar_init();
    }
    __syncwarp();

    const int64_t gw = (int64_t)blockIdx.x * nw + warp;
    const int64_t W = (int64_t)gridDim.x * nw;

    auto issue = [&](int64_t e0, int64_t e1, int d) {   // lane 0 only
        const uint32_t nent = (uint32_t)(e1 - e0);
        if (nent == 0) {
            mbar_arrive_expect_tx(&bars[d], 0);
        } else {
            mbar_arrive_expect_tx(&bars[d], nent * 12u);
            bulk_g2s(s_vals + (size_t)d * A.cap, A.vals + e0, nent * 8u, &bars[d]);
            bulk_g2s(s_cols + (size_t)d * A.cap, A.cols + e0, nent * 4u, &bars[d]);
        }
    };
    auto task_range = [&](int64_t t, int64_t &e0, int64_t &e1) {
        const int64_t s0 = A.s_first + t * G;
        const int64_t s1 = (s0 + G < A.s_end) ? s0 + G : A.s_end;
        e0 = A.slice_ptr[s0];
        e1 = A.slice_ptr[s1];
    };

    if (lane == 0) {
        int64_t pe0[8], pe1[8];
#pragma unroll
        for (int d = 0; d < 8; ++d) {
            const int64_t t = gw + (int64_t)d * W;
            pe0[d] = pe1[d] = 0;
            if (d < D && t < A.ntasks) task_range(t, pe0[d], pe1[d]);
        }
#pragma unroll
        for (int d = 0; d < 8; ++d) {
            const int64_t t = gw + (int64_t)d * W;
            if (d < D && t < A.ntasks) issue(pe0[d], pe1[d], d);
        }
    }
    // slice offsets of a task: lane i (i <= G) holds slice_ptr[s0 + i] (clamped at the end); loaded one task ahead
    auto load_sp = [&](int64_t t) -> int64_t {
        int64_t v = 0;
        if (t < A.ntasks && lane <= G) {
            const int64_t s0 = A.s_first + t * G;
            const int64_t s = (s0 + lane < A.s_end) ? s0 + lane : A.s_end;
            v = A.slice_ptr[s];
        }
        return v;
    };
    double contrib = 0.0;
    int it = 0;
    int64_t sp_next = load_sp(gw);
    for (int64_t t = gw; t < A.ntasks; t += W, ++it) {
        const int d = it % D;
        const uint32_t parity = (uint32_t)((it / D) & 1);
        const int64_t s0 = A.s_first + t * G;
        const int64_t sp = sp_next;
        sp_next = load_sp(t + W);
        // prefetch the range of the task that will reuse this ring slot
        const int64_t tn = t + (int64_t)D * W;
        int64_t ne0 = 0, ne1 = 0;
        if (lane == 0 && tn < A.ntasks) task_range(tn, ne0, ne1);
        // per-row vector operands, issued before waiting for the matrix data
        double bv[G], av[G];
        bool active[G];
#pragma unroll
        for (int i = 0; i < G; ++i) {
            const int64_t row = ((s0 + i) << 5) + lane;
            active[i] = (s0 + i < A.s_end) && row >= A.row_begin && row < A.row_end;
            bv[i] = 0.0;
            av[i] = 0.0;
            if (active[i]) {
                if (MODE == RESID || MODE == RESNORM || MODE == JACOBI || MODE == GS) bv[i] = b[row];
                if (MODE == JACOBI || MODE == PROLONG) av[i] = aux[row];
            }
        }
        const int64_t base0 = __shfl_sync(0xffffffffu, sp, 0);
        int off[G], len[G];
#pragma unroll
        for (int i = 0; i < G; ++i) {
            const int64_t a0 = __shfl_sync(0xffffffffu, sp, i);
            const int64_t a1 = __shfl_sync(0xffffffffu, sp, i + 1);
            off[i] = (int)(a0 - base0);
            len[i] = (int)((a1 - a0) >> 5);
        }
        mbar_wait(&bars[d], parity);
        const double *sv = s_vals + (size_t)d * A.cap + lane;
        const int32_t *sc = s_cols + (size_t)d * A.cap + lane;
#pragma unroll
        for (int i = 0; i < G; ++i) {
            const int64_t row = ((s0 + i) << 5) + lane;
            const double *v = sv + off[i];
            const int32_t *c = sc + off[i];
            double sum = 0.0, diag = 0.0;
            int k = 0;
            for (; k + 4 <= len[i]; k += 4) {
                const int32_t c0 = c[(k + 0) * kSlice], c1 = c[(k + 1) * kSlice], c2 = c[(k + 2) * kSlice],
                              c3 = c[(k + 3) * kSlice];
                const double v0 = v[(k + 0) * kSlice], v1 = v[(k + 1) * kSlice], v2 = v[(k + 2) * kSlice],
                             v3 = v[(k + 3) * kSlice];
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
            for (; k < len[i]; ++k) {
                const int32_t c0 = c[k * kSlice];
                const double v0 = v[k * kSlice];
                const double x0 = x[c0];
                if (MODE == GS) {
                    if (c0 == row) { if (v0 != 0.0) diag = v0; } else sum = mul_add_unfused(sum, v0, x0);
                } else {
                    sum = mul_add_unfused(sum, v0, x0);
                }
            }
            if (active[i]) {
                if (MODE == SPMV) {
                    y[row] = sum;
                } else if (MODE == RESID) {
                    y[row] = __dsub_rn(bv[i], sum);
                } else if (MODE == RESNORM) {
                    const double r = __dsub_rn(bv[i], sum);
                    contrib += r * r;
                } else if (MODE == JACOBI) {
                    const double r = __dsub_rn(bv[i], sum);
                    y[row] = __dadd_rn(x[row], __dmul_rn(omega, __dmul_rn(av[i], r)));
                } else if (MODE == GS) {
                    if (diag != 0.0) y[row] = __ddiv_rn(__dsub_rn(bv[i], sum), diag);
                } else if (MODE == PROLONG) {
                    y[row] = __dadd_rn(av[i], sum);
                }
            }
        }
        __syncwarp();
        if (lane == 0 && tn < A.ntasks) {
            fence_proxy_async();          // order this warp's generic reads of the slot before the async refill
            issue(ne0, ne1, d);
        }
    }
    if (MODE == RESNORM) {
        __shared__ double sh[8];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_down_sync(0xffffffffu, contrib, o);
        if (lane == 0) sh[warp] = contrib;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int w2 = 0; w2 < nw; ++w2) s += sh[w2];
            partials[blockIdx.x] = s;
        }
    }
}

struct TmaPlan {
    int G, depth, warps, cap, warp_bytes, ctas_per_sm;
    size_t smem;
};

// stage geometry from the longest slice: ~3-4 KB stages, ~12 KB ring per warp, ~190 KB in flight per SM
static bool make_plan(int64_t max_len, TmaPlan &p) {
    if (max_len <= 0) return false;
    const int64_t slice_bytes = 32 * max_len * 12;
    if (slice_bytes > 24 * 1024) return false;                  // very long rows: plain kernel
    p.G = (slice_bytes * 4 <= 4096) ? 4 : (slice_bytes * 2 <= 4096) ? 2 : 1;
    const int64_t stage = slice_bytes * p.G;
    p.cap = (int)(32 * max_len * p.G);
    p.depth = (int)(12 * 1024 / stage);
    if (p.depth < 2) p.depth = 2;
    if (p.depth > 8) p.depth = 8;
    int64_t wb = (int64_t)p.depth * stage + 8 * p.depth;
    wb = (wb + 127) / 128 * 128;
    p.warp_bytes = (int)wb;
    p.warps = 8;
    while (p.warps > 1 && (int64_t)p.warps * wb > 100 * 1024) p.warps >>= 1;
    p.smem = (size_t)p.warps * wb;
    if (p.smem > 200 * 1024) return false;
    p.ctas_per_sm = (int)((200 * 1024) / p.smem);
    if (p.ctas_per_sm > 4) p.ctas_per_sm = 4;
    if (p.ctas_per_sm < 1) p.ctas_per_sm = 1;
    return true;
}

template <int MODE, int G>
static int launch_g(const TmaPlan &p, TmaArgs &a, const double *x, const double *b, const double *aux, double *y,
                    double omega, double *partials, int *grid_out, cudaStream_t st, const char *name) {
    static bool configured = false;
    if (!configured) {
        MG_CHECK_CUDA(cudaFuncSetAttribute(sell_tma_kernel<MODE, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
    }
    a.ntasks = (a.s_end - a.s_first + G - 1) / G;
    int64_t grid = (int64_t)sm_count() * p.ctas_per_sm;
    const int64_t need = (a.ntasks + p.warps - 1) / p.warps;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    *grid_out = (int)grid;
    sell_tma_kernel<MODE, G><<<(unsigned)grid, p.warps * 32, p.smem, st>>>(a, x, b, aux, y, omega, partials);
    MG_CHECK_LAUNCH(name);
    return MG_OK;
}

// returns 1 if the matrix is not eligible (caller falls back to the plain kernel), 0 on success, <0 on error
template <int MODE>
int launch_sell_tma(const mg_sell *M, int64_t max_len, const double *x, const double *b, const double *aux, double *y,
                    double omega, double *partials, int64_t row0, int64_t row1, int *grid_out, cudaStream_t st,
                    const char *name) {
    TmaPlan p;
    if (!make_plan(max_len, p)) return 1;
    TmaArgs a;
    a.slice_ptr = M->d_slice_ptr;
    a.cols = M->d_cols;
    a.vals = M->d_vals;
    a.row_begin = row0;
    a.row_end = row1;
    a.s_first = row0 >> 5;
    a.s_end = (row1 + kSlice - 1) >> 5;
    a.depth = p.depth;
    a.cap = p.cap;
    a.warp_bytes = p.warp_bytes;
    a.ntasks = 0;
    switch (p.G) {
        case 4: return launch_g<MODE, 4>(p, a, x, b, aux, y, omega, partials, grid_out, st, name);
        case 2: return launch_g<MODE, 2>(p, a, x, b, aux, y, omega, partials, grid_out, st, name);
        default: return launch_g<MODE, 1>(p, a, x, b, aux, y, omega, partials, grid_out, st, name);
    }
}

template int launch_sell_tma<SPMV>(const mg_sell *, int64_t, const double *, const double *, const double *, double *, double, double *, int64_t, int64_t, int *, cudaStream_t, const char *);
template int launch_sell_tma<RESID>(const mg_sell *, int64_t, const double *, const double *, const double *, double *, double, double *, int64_t, int64_t, int *, cudaStream_t, const char *);
template int launch_sell_tma<RESNORM>(const mg_sell *, int64_t, const double *, const double *, const double *, double *, double, double *, int64_t, int64_t, int *, cudaStream_t, const char *);
template int launch_sell_tma<JACOBI>(const mg_sell *, int64_t, const double *, const double *, const double *, double *, double, double *, int64_t, int64_t, int *, cudaStream_t, const char *);
template int launch_sell_tma<GS>(const mg_sell *, int64_t, const double *, const double *, const double *, double *, double, double *, int64_t, int64_t, int *, cudaStream_t, const char *);
template int launch_sell_tma<PROLONG>(const mg_sell *, int64_t, const double *, const double *, const double *, double *, double, double *, int64_t, int64_t, int *, cudaStream_t, const char *);

}  // namespace mgb
