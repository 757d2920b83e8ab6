// tools/sellbench.cu -- kernel-variant microbenchmark for the fine-level Gauss-Seidel colour sweep (development
// tool, not part of the product).  Builds the colour-blocked 5-point operator of an (N+1)^2 grid directly on the
// device and times several kernel variants with CUDA events.  Usage: sellbench N [reps]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ double madd(double acc, double a, double b) { return __dadd_rn(acc, __dmul_rn(a, b)); }

struct Grid { int64_t N, W, n, nred; };
__device__ __forceinline__ int64_t perm_of(const Grid g, int64_t i) { return (i & 1) ? g.nred + (i >> 1) : (i >> 1); }
__device__ __forceinline__ int64_t nat_of(const Grid g, int64_t p) { return p < g.nred ? 2 * p : 2 * (p - g.nred) + 1; }
__device__ __forceinline__ bool interior(const Grid g, int64_t i) {
    int64_t iy = i / g.W, ix = i % g.W;
    return ix > 0 && ix < g.N && iy > 0 && iy < g.N;
}
__global__ void slice_len_kernel(Grid g, int32_t *len, int uniform) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t ns = (g.n + 31) / 32;
    if (s >= ns) return;
    int l = 0;
    for (int r = 0; r < 32; ++r) { int64_t p = s * 32 + r; if (p < g.n) l = max(l, interior(g, nat_of(g, p)) ? 5 : 1); }
    len[s] = uniform ? 5 : l;
}
__global__ void fill_kernel(Grid g, const int64_t *sp, int32_t *cols, double *vals) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t ns = (g.n + 31) / 32;
    if (p >= ns * 32) return;
    int64_t s = p >> 5; int lane = p & 31;
    int64_t base = sp[s]; int len = (int)((sp[s + 1] - base) >> 5);
    int32_t c[5]; double v[5]; int m = 0;
    if (p < g.n) {
        int64_t i = nat_of(g, p);
        if (interior(g, i)) {
            int64_t nb[5] = {i - g.W, i - 1, i, i + 1, i + g.W};
            for (int k = 0; k < 5; ++k) { c[k] = (int32_t)perm_of(g, nb[k]); v[k] = (k == 2) ? 4.0 : -1.0; }
            m = 5;
        } else { c[0] = (int32_t)p; v[0] = 1.0; m = 1; }
    }
    for (int k = 0; k < len; ++k) {
        int64_t d = base + (int64_t)k * 32 + lane;
        if (k < m) { cols[d] = c[k]; vals[d] = v[k]; } else { cols[d] = m ? c[m - 1] : 0; vals[d] = 0.0; }
    }
}

struct Args { const int64_t *sp; const int32_t *cols; const double *vals; int64_t r0, r1, first; int ulen; };

// ---- V0: the product's plain kernel ------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gs_v0(Args A, double *x, const double *__restrict__ b) {
    const int64_t row = A.first + (int64_t)blockIdx.x * 256 + threadIdx.x;
    const bool active = row >= A.r0 && row < A.r1;
    if (row < A.r1) {
        const int64_t slice = row >> 5; const int lane = row & 31;
        const int64_t base = A.sp[slice];
        const int len = (int)((A.sp[slice + 1] - base) >> 5);
        const double *v = A.vals + base + lane; const int32_t *c = A.cols + base + lane;
        double sum = 0, diag = 0;
        for (int k = 0; k < len; ++k) {
            int32_t c0 = __ldcs(c + k * 32); double v0 = __ldcs(v + k * 32); double x0 = x[c0];
            if (c0 == row) { if (v0 != 0.0) diag = v0; } else sum = madd(sum, v0, x0);
        }
        if (active && diag != 0.0) x[row] = __ddiv_rn(__dsub_rn(b[row], sum), diag);
    }
}
// ---- V1: uniform slice length (no slice_ptr load), R rows per thread, b loaded first ---------------------------
template <int R>
__global__ void __launch_bounds__(256) gs_v1(Args A, double *x, const double *__restrict__ b) {
    const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t warp = t >> 5; const int lane = t & 31;
    const int len = A.ulen;
    int64_t row[R]; double bv[R]; bool act[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        row[r] = A.first + (warp * R + r) * 32 + lane;
        act[r] = row[r] >= A.r0 && row[r] < A.r1;
        bv[r] = act[r] ? b[row[r]] : 0.0;
    }
    double sum[R], diag[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { sum[r] = 0; diag[r] = 0; }
    if (row[0] >= A.r1) return;
    for (int k = 0; k < len; ++k) {
        int32_t c0[R]; double v0[R], x0[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int64_t e = (row[r] >> 5) * (int64_t)len * 32 + (int64_t)k * 32 + lane;
            const bool ok = row[r] < A.r1;
            c0[r] = ok ? __ldcs(A.cols + e) : 0; v0[r] = ok ? __ldcs(A.vals + e) : 0.0;
        }
#pragma unroll
        for (int r = 0; r < R; ++r) x0[r] = x[c0[r]];
#pragma unroll
        for (int r = 0; r < R; ++r) { if (c0[r] == row[r]) { if (v0[r] != 0.0) diag[r] = v0[r]; } else sum[r] = madd(sum[r], v0[r], x0[r]); }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) if (act[r] && diag[r] != 0.0) x[row[r]] = __ddiv_rn(__dsub_rn(bv[r], sum[r]), diag[r]);
}
// ---- V2: like V1 but fully unrolled for len == 5: all 5 cols/vals loads issued before the gathers --------------
template <int R>
__global__ void __launch_bounds__(256) gs_v2(Args A, double *x, const double *__restrict__ b) {
    const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t warp = t >> 5; const int lane = t & 31;
    int64_t row[R]; double bv[R]; bool act[R];
    int32_t c[R][5]; double v[R][5];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        row[r] = A.first + (warp * R + r) * 32 + lane;
        act[r] = row[r] >= A.r0 && row[r] < A.r1;
        const bool ok = row[r] < A.r1;
        const int64_t e = (row[r] >> 5) * 160 + lane;
#pragma unroll
        for (int k = 0; k < 5; ++k) { c[r][k] = ok ? __ldcs(A.cols + e + k * 32) : 0; v[r][k] = ok ? __ldcs(A.vals + e + k * 32) : 0.0; }
        bv[r] = act[r] ? b[row[r]] : 0.0;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        double xx[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) xx[k] = x[c[r][k]];
        double sum = 0, diag = 0;
#pragma unroll
        for (int k = 0; k < 5; ++k) { if (c[r][k] == row[r]) { if (v[r][k] != 0.0) diag = v[r][k]; } else sum = madd(sum, v[r][k], xx[k]); }
        if (act[r] && diag != 0.0) x[row[r]] = __ddiv_rn(__dsub_rn(bv[r], sum), diag);
    }
}
// ---- V4: product-style: slice_ptr + chunked register staging -------------------------------------------------
template <int CH, bool USE_SP>
__global__ void __launch_bounds__(256) gs_v4(Args A, double *x, const double *__restrict__ b) {
    const int64_t row = A.first + (int64_t)blockIdx.x * 256 + threadIdx.x;
    const bool active = row >= A.r0 && row < A.r1;
    if (row < A.r1) {
        const int64_t slice = row >> 5; const int lane = row & 31;
        int64_t base; int len;
        if (USE_SP) { base = A.sp[slice]; len = (int)((A.sp[slice + 1] - base) >> 5); }
        else { base = slice * 160; len = A.ulen; }
        const double *__restrict__ v = A.vals + base + lane; const int32_t *__restrict__ c = A.cols + base + lane;
        double bv = active ? b[row] : 0.0;
        double sum = 0, diag = 0;
        for (int k0 = 0; k0 < len; k0 += CH) {
            int32_t cc[CH]; double vv[CH], xx[CH];
#pragma unroll
            for (int j = 0; j < CH; ++j) { const bool ok = k0 + j < len; cc[j] = ok ? __ldcs(c + (k0 + j) * 32) : 0; vv[j] = ok ? __ldcs(v + (k0 + j) * 32) : 0.0; }
#pragma unroll
            for (int j = 0; j < CH; ++j) xx[j] = (k0 + j < len) ? x[cc[j]] : 0.0;
#pragma unroll
            for (int j = 0; j < CH; ++j) if (k0 + j < len) { if (cc[j] == row) { if (vv[j] != 0.0) diag = vv[j]; } else sum = madd(sum, vv[j], xx[j]); }
        }
        if (active && diag != 0.0) x[row] = __ddiv_rn(__dsub_rn(bv, sum), diag);
    }
}
// ---- V5: slice_ptr + switch on the (warp-uniform) slice length into straight-line chunks ----------------------------
template <int CNT>
__device__ __forceinline__ void chunk(const int32_t *__restrict__ c, const double *__restrict__ v, const double *x, int64_t row, double &sum, double &diag) {
    int32_t cc[CNT]; double vv[CNT], xx[CNT];
#pragma unroll
    for (int j = 0; j < CNT; ++j) { cc[j] = __ldcs(c + j * 32); vv[j] = __ldcs(v + j * 32); }
#pragma unroll
    for (int j = 0; j < CNT; ++j) xx[j] = x[cc[j]];
#pragma unroll
    for (int j = 0; j < CNT; ++j) { if (cc[j] == row) { if (vv[j] != 0.0) diag = vv[j]; } else sum = madd(sum, vv[j], xx[j]); }
}
template <bool USE_SP>
__global__ void __launch_bounds__(256) gs_v5(Args A, double *x, const double *__restrict__ b) {
    const int64_t row = A.first + (int64_t)blockIdx.x * 256 + threadIdx.x;
    const bool active = row >= A.r0 && row < A.r1;
    if (row < A.r1) {
        const int64_t slice = row >> 5; const int lane = row & 31;
        int64_t base; int len;
        if (USE_SP) { base = A.sp[slice]; len = (int)((A.sp[slice + 1] - base) >> 5); }
        else { base = slice * 160; len = A.ulen; }
        const double *__restrict__ v = A.vals + base + lane; const int32_t *__restrict__ c = A.cols + base + lane;
        double bv = active ? b[row] : 0.0;
        double sum = 0, diag = 0;
        int k0 = 0;
        for (; k0 + 8 <= len; k0 += 8) chunk<8>(c + k0 * 32, v + k0 * 32, x, row, sum, diag);
        switch (len - k0) {
            case 1: chunk<1>(c + k0 * 32, v + k0 * 32, x, row, sum, diag); break;
            case 2: chunk<2>(c + k0 * 32, v + k0 * 32, x, row, sum, diag); break;
            case 3: chunk<3>(c + k0 * 32, v + k0 * 32, x, row, sum, diag); break;
            case 4: chunk<4>(c + k0 * 32, v + k0 * 32, x, row, sum, diag); break;
            case 5: chunk<5>(c + k0 * 32, v + k0 * 32, x, row, sum, diag); break;
            case 6: chunk<6>(c + k0 * 32, v + k0 * 32, x, row, sum, diag); break;
            case 7: chunk<7>(c + k0 * 32, v + k0 * 32, x, row, sum, diag); break;
            default: break;
        }
        if (active && diag != 0.0) x[row] = __ddiv_rn(__dsub_rn(bv, sum), diag);
    }
}
// ---- V6: launch-time specialisation on the matrix' max slice length; other lengths take a rolled loop ----------------
template <int LEN, bool UNIFORM>
__global__ void __launch_bounds__(256) gs_v6(Args A, double *x, const double *__restrict__ b) {
    const int64_t row = A.first + (int64_t)blockIdx.x * 256 + threadIdx.x;
    const bool active = row >= A.r0 && row < A.r1;
    if (row < A.r1) {
        const int64_t slice = row >> 5; const int lane = row & 31;
        int64_t base; int len;
        if (!UNIFORM) { base = A.sp[slice]; len = (int)((A.sp[slice + 1] - base) >> 5); }
        else { base = slice * (32 * LEN); len = LEN; }
        const double *__restrict__ v = A.vals + base + lane; const int32_t *__restrict__ c = A.cols + base + lane;
        double sum = 0, diag = 0;
        if (UNIFORM || len == LEN) chunk<LEN>(c, v, x, row, sum, diag);
        else for (int k = 0; k < len; ++k) { int32_t c0 = __ldcs(c + k * 32); double v0 = __ldcs(v + k * 32); double x0 = x[c0];
                                             if (c0 == row) { if (v0 != 0.0) diag = v0; } else sum = madd(sum, v0, x0); }
        if (active && diag != 0.0) x[row] = __ddiv_rn(__dsub_rn(b[row], sum), diag);
    }
}
// ---- stream-only bound: read cols+vals+b, write x, no gathers --------------------------------------------------
__global__ void __launch_bounds__(256) stream_only(Args A, double *x, const double *__restrict__ b) {
    const int64_t row = A.first + (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (row >= A.r1) return;
    const int64_t e = (row >> 5) * 160 + (row & 31);
    double s = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) s += __ldcs(A.vals + e + k * 32) * (double)__ldcs(A.cols + e + k * 32);
    if (row >= A.r0) x[row] = b[row] - s;
}

// ---- V3: bulk-async ring (warp private), parameters at run time -------------------------------------------------
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mb_expect(uint64_t *b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mb_wait(uint64_t *b, uint32_t ph) {
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk(void *d, const void *s, uint32_t n, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(d)), "l"(s), "r"(n), "r"(s32(b)) : "memory");
}
// CTA-wide ring: one producer thread (warp 0 lane 0) streams chunks of SL slices (uniform len 5) into a ring of D
// stages; all warps consume a stage together (rows = SL*32 per stage), full/empty mbarriers.
template <int SL>
__global__ void __launch_bounds__(SL * 32 + 32) gs_v3(Args A, double *x, const double *__restrict__ b, int D, int64_t nchunks) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int ent = SL * 160;                       // entries per stage
    double *sv = (double *)sm; int32_t *sc = (int32_t *)(sm + (size_t)D * ent * 8);
    uint64_t *full = (uint64_t *)(sm + (size_t)D * ent * 12); uint64_t *empty = full + D;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { for (int d = 0; d < D; ++d) { mb_init(&full[d], 1); mb_init(&empty[d], SL); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const int64_t s_first = A.first >> 5;
    if (warp == SL) {                               // producer warp
        if (lane == 0) {
            int it = 0;
            for (int64_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x, ++it) {
                const int d = it % D;
                if (it >= D) mb_wait(&empty[d], ((it / D) - 1) & 1);
                const int64_t e0 = (s_first + ch * SL) * 160;
                mb_expect(&full[d], ent * 12);
                bulk(sv + (size_t)d * ent, A.vals + e0, ent * 8, &full[d]);
                bulk(sc + (size_t)d * ent, A.cols + e0, ent * 4, &full[d]);
            }
        }
        return;
    }
    int it = 0;
    for (int64_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x, ++it) {
        const int d = it % D;
        const int64_t row = ((s_first + ch * SL + warp) << 5) + lane;
        const bool act = row >= A.r0 && row < A.r1;
        const double bv = act ? b[row] : 0.0;
        mb_wait(&full[d], (it / D) & 1);
        const double *v = sv + (size_t)d * ent + warp * 160 + lane; const int32_t *c = sc + (size_t)d * ent + warp * 160 + lane;
        int32_t cc[5]; double vv[5], xx[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) { cc[k] = c[k * 32]; vv[k] = v[k * 32]; }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[d])) : "memory");
#pragma unroll
        for (int k = 0; k < 5; ++k) xx[k] = x[cc[k]];
        double sum = 0, diag = 0;
#pragma unroll
        for (int k = 0; k < 5; ++k) { if (cc[k] == row) { if (vv[k] != 0.0) diag = vv[k]; } else sum = madd(sum, vv[k], xx[k]); }
        if (act && diag != 0.0) x[row] = __ddiv_rn(__dsub_rn(bv, sum), diag);
    }
}

int main(int argc, char **argv) {
    int64_t N = argc > 1 ? atoll(argv[1]) : 4096; int reps = argc > 2 ? atoi(argv[2]) : 20;
    Grid g; g.N = N; g.W = N + 1; g.n = g.W * g.W; g.nred = (g.n + 1) / 2;
    int64_t ns = (g.n + 31) / 32;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s, %d SMs; grid %lld^2 = %lld rows\n", prop.name, prop.multiProcessorCount, (long long)g.W, (long long)g.n);
    double *x, *b; CK(cudaMalloc(&x, g.n * 8)); CK(cudaMalloc(&b, g.n * 8));
    CK(cudaMemset(x, 0, g.n * 8)); CK(cudaMemset(b, 0, g.n * 8));
    struct M { int64_t *sp; int32_t *cols; double *vals; int64_t tot; } m[2];
    for (int u = 0; u < 2; ++u) {
        int32_t *len; CK(cudaMalloc(&len, ns * 4));
        slice_len_kernel<<<(unsigned)((ns + 255) / 256), 256>>>(g, len, u);
        std::vector<int32_t> hl(ns); CK(cudaMemcpy(hl.data(), len, ns * 4, cudaMemcpyDeviceToHost));
        std::vector<int64_t> hp(ns + 1); hp[0] = 0; for (int64_t s = 0; s < ns; ++s) hp[s + 1] = hp[s] + (int64_t)hl[s] * 32;
        m[u].tot = hp[ns];
        CK(cudaMalloc(&m[u].sp, (ns + 1) * 8)); CK(cudaMemcpy(m[u].sp, hp.data(), (ns + 1) * 8, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&m[u].cols, m[u].tot * 4)); CK(cudaMalloc(&m[u].vals, m[u].tot * 8));
        fill_kernel<<<(unsigned)((ns * 32 + 255) / 256), 256>>>(g, m[u].sp, m[u].cols, m[u].vals);
        CK(cudaDeviceSynchronize()); CK(cudaFree(len));
        printf("matrix %d: %lld padded entries (%.2f GB)\n", u, (long long)m[u].tot, m[u].tot * 12 / 1e9);
    }
    const double nnz = 5.0 * (N - 1) * (N - 1) + (g.n - (double)(N - 1) * (N - 1));
    const double sweep_bytes = 12 * nnz + 4 * (g.n + 1) + 24.0 * g.n;      // S(a,n) + 24 n
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto colour = [&](int c, int64_t &r0, int64_t &r1) { r0 = c ? g.nred : 0; r1 = c ? g.n : g.nred; };
    auto report = [&](const char *name, float ms) { printf("%-44s %8.3f ms/sweep  %8.1f GB/s algorithmic\n", name, ms, sweep_bytes / (ms * 1e-3) / 1e9); fflush(stdout); };
    auto timeit = [&](const char *name, auto launch) {
        for (int i = 0; i < 3; ++i) for (int c = 0; c < 2; ++c) launch(c);
        CK(cudaDeviceSynchronize()); CK(cudaEventRecord(e0));
        for (int i = 0; i < reps; ++i) for (int c = 0; c < 2; ++c) launch(c);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize()); CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); report(name, ms / reps);
    };
    auto mkargs = [&](int u, int c) { Args a; a.sp = m[u].sp; a.cols = m[u].cols; a.vals = m[u].vals; colour(c, a.r0, a.r1); a.first = a.r0 & ~31LL; a.ulen = 5; return a; };
    timeit("v0 plain (slice_ptr)", [&](int c) { Args a = mkargs(0, c); int64_t nt = a.r1 - a.first; gs_v0<<<(unsigned)((nt + 255) / 256), 256>>>(a, x, b); });
    timeit("v0 plain on uniform matrix", [&](int c) { Args a = mkargs(1, c); int64_t nt = a.r1 - a.first; gs_v0<<<(unsigned)((nt + 255) / 256), 256>>>(a, x, b); });
    timeit("v1 uniform R=1", [&](int c) { Args a = mkargs(1, c); int64_t nt = a.r1 - a.first; gs_v1<1><<<(unsigned)((nt + 255) / 256), 256>>>(a, x, b); });
    timeit("v1 uniform R=2", [&](int c) { Args a = mkargs(1, c); int64_t nt = (a.r1 - a.first + 1) / 2 + 32; gs_v1<2><<<(unsigned)((nt + 255) / 256), 256>>>(a, x, b); });
    timeit("v1 uniform R=4", [&](int c) { Args a = mkargs(1, c); int64_t nt = (a.r1 - a.first + 3) / 4 + 32; gs_v1<4><<<(unsigned)((nt + 255) / 256), 256>>>(a, x, b); });
    timeit("v2 uniform unrolled R=1", [&](int c) { Args a = mkargs(1, c); int64_t nt = a.r1 - a.first; gs_v2<1><<<(unsigned)((nt + 255) / 256), 256>>>(a, x, b); });
    timeit("v2 uniform unrolled R=2", [&](int c) { Args a = mkargs(1, c); int64_t nt = (a.r1 - a.first + 1) / 2 + 32; gs_v2<2><<<(unsigned)((nt + 255) / 256), 256>>>(a, x, b); });
    timeit("v4 slice_ptr + CH=5", [&](int c) { Args a = mkargs(0, c); int64_t nt = a.r1 - a.first; gs_v4<5, true><<<(unsigned)((nt + 255) / 256), 256>>>(a, x, b); });
    timeit("v4 slice_ptr + CH=8", [&](int c) { Args a = mkargs(0, c); int64_t nt = a.r1 - a.first; gs_v4<8, true><<<(unsigned)((nt + 255) / 256), 256>>>(a, x, b); });
    timeit("v4 computed base, runtime len + CH=5", [&](int c) { Args a = mkargs(1, c); int64_t nt = a.r1 - a.first; gs_v4<5, false><<<(unsigned)((nt + 255) / 256), 256>>>(a, x, b); });
    timeit("v5 slice_ptr + switch(len)", [&](int c) { Args a = mkargs(0, c); int64_t nt = a.r1 - a.first; gs_v5<true><<<(unsigned)((nt + 255) / 256), 256>>>(a, x, b); });
    timeit("v5 computed base + switch(len)", [&](int c) { Args a = mkargs(1, c); int64_t nt = a.r1 - a.first; gs_v5<false><<<(unsigned)((nt + 255) / 256), 256>>>(a, x, b); });
    timeit("v6 slice_ptr, LEN=5 fast path", [&](int c) { Args a = mkargs(0, c); int64_t nt = a.r1 - a.first; gs_v6<5, false><<<(unsigned)((nt + 255) / 256), 256>>>(a, x, b); });
    timeit("v6 uniform, LEN=5", [&](int c) { Args a = mkargs(1, c); int64_t nt = a.r1 - a.first; gs_v6<5, true><<<(unsigned)((nt + 255) / 256), 256>>>(a, x, b); });
    timeit("stream-only bound (no gathers)", [&](int c) { Args a = mkargs(1, c); int64_t nt = a.r1 - a.first; stream_only<<<(unsigned)((nt + 255) / 256), 256>>>(a, x, b); });
    for (int D : {2}) for (int cps : {3}) {
        auto run = [&](auto kern, int SL) {
            size_t smem = (size_t)D * SL * 160 * 12 + 16 * D;
            if (smem * cps > 220 * 1024) return;
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            char nm[96]; snprintf(nm, 96, "v3 TMA ring SL=%d D=%d ctas/SM=%d (%.0f KB)", SL, D, cps, smem * cps / 1024.0);
            timeit(nm, [&](int c) { Args a = mkargs(1, c); int64_t nsl = ((a.r1 + 31) >> 5) - (a.first >> 5); int64_t nch = nsl / SL;   /* tail ignored */
                                    kern<<<prop.multiProcessorCount * cps, SL * 32 + 32, smem>>>(a, x, b, D, nch); });
        };
        run(gs_v3<8>, 8); run(gs_v3<16>, 16);
    }
    // plain device-to-device copy of the same byte count for reference
    { size_t nb = (size_t)(sweep_bytes / 2); void *s, *d; CK(cudaMalloc(&s, nb)); CK(cudaMalloc(&d, nb));
      CK(cudaMemcpy(d, s, nb, cudaMemcpyDeviceToDevice)); CK(cudaDeviceSynchronize()); CK(cudaEventRecord(e0));
      for (int i = 0; i < reps; ++i) CK(cudaMemcpyAsync(d, s, nb, cudaMemcpyDeviceToDevice)); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); printf("%-44s %8.3f ms        %8.1f GB/s (read+write)\n", "cudaMemcpy D2D same bytes", ms / reps, 2.0 * nb / (ms / reps * 1e-3) / 1e9); }
    return 0;
}
