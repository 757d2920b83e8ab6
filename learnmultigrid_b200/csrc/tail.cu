// tail.cu -- the latency-bound tail of the V-cycle as ONE persistent cooperative kernel ("tail program").
//
// On the coarse levels of the hierarchy (Multigrid.v_cycle, learn_multigrid/solvers/Multigrid.py:77-124, levels of a
// few 10^5 rows and below) every kernel of the cycle -- one colour of a Gauss-Seidel sweep, the residual, Q^T r,
// x += Q e -- moves a few megabytes at most and costs its launch gap, not its bytes: ~100 dependent launches of
// 3-4 us each on BASELINE configs[1] (profiles/r01_bench_c2_irregular_nn.json), ~35 on the replicated tail of the
// row-partitioned 64 M-DOF run, which is what is left of Amdahl's law at 8 GPUs.  Here cycle.cu RECORDS
// those operations instead of launching them; the record is handed to one cooperative kernel as a by-value program
// (<= 40 operations of 96 bytes, i.e. inside the 4 KB parameter space, so it is baked into the CUDA-graph node and
// needs no device buffer), whose CTAs walk the operations and separate dependent ones by a grid barrier.
//
// Arithmetic: the per-row code is the one of sell_kernels.cu -- separately rounded products added in storage order,
// the Gauss-Seidel diagonal taken from the row -- written once as __host__ __device__ code, so results are
// bit-identical to the launch-per-operation path (tests/test_gpu_tail.py) and the same code runs serially on host
// arrays for the CPU suite (mg_host_tail_vcycle, tests/test_tail_host.py), including a shuffled execution order inside
// every barrier-free group, which checks the dependence analysis that places the barriers.
// Off by default: mg_set_tail_max_rows(rows) turns it on for the levels with at most `rows` rows.
#include <cooperative_groups.h>
#include <algorithm>
#include <vector>
#include "tail.cuh"

namespace cg = cooperative_groups;

namespace mgb {

struct TailOp {
    const int64_t *slice_ptr;   // SELL matrix (T_SELL)
    const int32_t *cols;
    const double *vals;
    const double *x;            // gathered vector (T_SELL), source (T_COPY)
    const double *b;
    const double *aux;          // dinv (JACOBI, T_DIAG_SCALE), u (PROLONG)
    double *y;
    int64_t row0, row1;         // rows / elements this operation writes
    double omega;               // Jacobi damping; fill value (T_FILL)
    int32_t kind;               // TailKind
    int32_t mode;               // SellMode (T_SELL)
    int32_t uniform_len;        // > 0: every slice has this many entries per row (offsets computed)
    int32_t sync_after;         // grid barrier before the next operation
};
static_assert(sizeof(TailOp) == 96, "TailOp layout");

constexpr int kTailMaxOps = 40;
struct TailProgram {
    int32_t nops, pad_;
    TailOp ops[kTailMaxOps];
};
static_assert(sizeof(TailProgram) <= 4000, "the program must fit the 4 KB kernel parameter space");

// separately rounded product and sum on both sides (the host compiler does not contract without -mfma)
__host__ __device__ __forceinline__ double hd_mul(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    volatile double p = a * b;
    return p;
#endif
}
__host__ __device__ __forceinline__ double hd_add(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    volatile double s = a + b;
    return s;
#endif
}
__host__ __device__ __forceinline__ double hd_sub(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dsub_rn(a, b);
#else
    volatile double s = a - b;
    return s;
#endif
}
__host__ __device__ __forceinline__ double hd_div(double a, double b) {
#ifdef __CUDA_ARCH__
    return __ddiv_rn(a, b);
#else
    volatile double s = a / b;
    return s;
#endif
}

// what sell_body does with the finished row sum (sell_kernels.cu)
__host__ __device__ __forceinline__ void tail_epilogue(const TailOp &op, int64_t row, double sum, double diag) {
    switch (op.mode) {
        case SPMV: op.y[row] = sum; break;
        case RESID: op.y[row] = hd_sub(op.b[row], sum); break;
        case JACOBI:
            op.y[row] = hd_add(op.x[row], hd_mul(op.omega, hd_mul(op.aux[row], hd_sub(op.b[row], sum))));
            break;
        case GS:
            if (diag != 0.0) op.y[row] = hd_div(hd_sub(op.b[row], sum), diag);
            break;
        case PROLONG: op.y[row] = hd_add(op.aux[row], sum); break;
        default: break;
    }
}
__host__ __device__ __forceinline__ void tail_accumulate(int mode, int32_t c, double v, double xv, int64_t row,
                                                         double &sum, double &diag) {
    if (mode == GS && c == row) {
        if (v != 0.0) diag = v;
    } else {
        sum = hd_add(sum, hd_mul(v, xv));
    }
}
__host__ __device__ __forceinline__ void tail_slice_extent(const TailOp &op, int64_t slice, int64_t &base, int &len) {
    if (op.uniform_len > 0) {
        base = slice * (int64_t)(kSlice * op.uniform_len);
        len = op.uniform_len;
    } else {
        base = op.slice_ptr[slice];
        len = (int)((op.slice_ptr[slice + 1] - base) >> 5);
    }
}
__host__ __device__ __forceinline__ void tail_vector_element(const TailOp &op, int64_t i) {
    if (op.kind == T_FILL) op.y[i] = op.omega;
    else if (op.kind == T_DIAG_SCALE) op.y[i] = hd_add(0.0, hd_mul(op.omega, hd_mul(op.aux[i], op.b[i])));
    else if (op.kind == T_COPY) op.y[i] = hd_mul(1.0, op.x[i]);
}

// one row of a SELL operation, entries in storage order (host executor)
static void tail_row_host(const TailOp &op, int64_t row) {
    int64_t base;
    int len;
    tail_slice_extent(op, row >> 5, base, len);
    const int lane = (int)(row & 31);
    double sum = 0.0, diag = 0.0;
    for (int k = 0; k < len; ++k) {
        const int32_t c = op.cols[base + (int64_t)k * kSlice + lane];
        const double v = op.vals[base + (int64_t)k * kSlice + lane];
        tail_accumulate(op.mode, c, v, op.x[c], row, sum, diag);
    }
    tail_epilogue(op, row, sum, diag);
}

// one slice per warp, one row per lane; the loads of up to kTailChunk entries are issued before the first gather.
// Vectors are read with plain (coherent) loads: other CTAs wrote them earlier in this kernel, and the acquire side of
// the grid barrier is what makes those writes visible, so no non-coherent (ld.global.nc) path may be used for them.
constexpr int kTailChunk = 8;
__device__ __forceinline__ void tail_slice_device(const TailOp &op, int64_t slice, int lane) {
    const int64_t row = slice * kSlice + lane;
    if (row < op.row0 || row >= op.row1) return;
    int64_t base;
    int len;
    tail_slice_extent(op, slice, base, len);
    const double *v = op.vals + base + lane;
    const int32_t *c = op.cols + base + lane;
    const double *x = op.x;
    double sum = 0.0, diag = 0.0;
    for (int k0 = 0; k0 < len; k0 += kTailChunk) {
        int32_t cc[kTailChunk];
        double vv[kTailChunk], xx[kTailChunk];
#pragma unroll
        for (int j = 0; j < kTailChunk; ++j)
            if (k0 + j < len) {
                cc[j] = ld_stream(c + (int64_t)(k0 + j) * kSlice);
                vv[j] = ld_stream(v + (int64_t)(k0 + j) * kSlice);
            }
#pragma unroll
        for (int j = 0; j < kTailChunk; ++j)
            if (k0 + j < len) xx[j] = x[cc[j]];
#pragma unroll
        for (int j = 0; j < kTailChunk; ++j)
            if (k0 + j < len) tail_accumulate(op.mode, cc[j], vv[j], xx[j], row, sum, diag);
    }
    tail_epilogue(op, row, sum, diag);
}

__global__ void __launch_bounds__(kBlock) tail_kernel(const __grid_constant__ TailProgram prog) {
    pdl_prologue();
    cg::grid_group grid = cg::this_grid();
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (int64_t)blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kBlock / 32);
    const int64_t gthread = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * kBlock;
    for (int i = 0; i < prog.nops; ++i) {
        const TailOp &op = prog.ops[i];
        if (op.kind == T_SELL) {
            const int64_t s1 = (op.row1 + kSlice - 1) >> 5;
            for (int64_t s = (op.row0 >> 5) + gwarp; s < s1; s += nwarps) tail_slice_device(op, s, lane);
        } else {
            for (int64_t e = op.row0 + gthread; e < op.row1; e += nthreads) tail_vector_element(op, e);
        }
        if (op.sync_after && i + 1 < prog.nops) grid.sync();
    }
}

// ---- recorder ------------------------------------------------------------------------------------------------------
struct TailRecorder {
    std::vector<TailOp> ops;
    std::vector<int64_t> xlen;     // per op: entries of the gathered vector (dependence analysis)
    bool active = false;
    bool host = false;
    uint64_t shuffle = 0;          // host mode: != 0 executes every barrier-free group in a pseudo-random order
    int64_t total_ops = 0, total_syncs = 0, total_launches = 0;
};
static thread_local TailRecorder g_rec;
static int64_t g_tail_max_rows = 0;
static int g_tail_ctas_per_sm = 2;
static int g_tail_drop_barriers = 0;     // test hook: proves that the shuffled host execution detects a missing barrier

bool tail_recording() { return g_rec.active; }
bool tail_host_mode() { return g_rec.active && g_rec.host; }
int64_t tail_max_rows() { return g_tail_max_rows; }
void tail_begin(bool host, uint64_t shuffle) {
    g_rec.ops.clear();
    g_rec.xlen.clear();
    g_rec.active = true;
    g_rec.host = host;
    g_rec.shuffle = shuffle;
}
void tail_end() {
    g_rec.active = false;
    g_rec.ops.clear();
    g_rec.xlen.clear();
}
void tail_stats_reset() { g_rec.total_ops = g_rec.total_syncs = g_rec.total_launches = 0; }
void tail_stats(int64_t *ops, int64_t *syncs, int64_t *launches) {
    if (ops) *ops = g_rec.total_ops;
    if (syncs) *syncs = g_rec.total_syncs;
    if (launches) *launches = g_rec.total_launches;
}

void tail_record_sell(int mode, const mg_sell *A, const double *x, const double *b, const double *aux, double *y,
                      double omega, int64_t row0, int64_t row1) {
    if (row1 <= row0) return;
    TailOp op;
    memset(&op, 0, sizeof(op));
    op.slice_ptr = A->d_slice_ptr;
    op.cols = A->d_cols;
    op.vals = A->d_vals;
    op.x = x;
    op.b = b;
    op.aux = aux;
    op.y = y;
    op.row0 = row0;
    op.row1 = row1;
    op.omega = omega;
    op.kind = T_SELL;
    op.mode = mode;
    op.uniform_len = (A->uniform_len > 0 && A->uniform_len == A->max_slice_len) ? (int32_t)A->uniform_len : 0;
    op.sync_after = 1;
    g_rec.ops.push_back(op);
    g_rec.xlen.push_back(A->ncols);
}
void tail_record_vector(int kind, int64_t n, double value, const double *src, const double *b, const double *aux,
                        double *y) {
    if (n <= 0) return;
    TailOp op;
    memset(&op, 0, sizeof(op));
    op.x = src;
    op.b = b;
    op.aux = aux;
    op.y = y;
    op.row0 = 0;
    op.row1 = n;
    op.omega = value;
    op.kind = kind;
    op.sync_after = 1;
    g_rec.ops.push_back(op);
    g_rec.xlen.push_back(0);
}

// byte ranges an operation reads / writes
struct Span {
    const char *lo, *hi;
};
static inline bool overlap(const Span &a, const Span &b) { return a.lo < b.hi && b.lo < a.hi; }
static inline Span span_of(const double *p, int64_t i0, int64_t i1) {
    return Span{(const char *)(p + i0), (const char *)(p + i1)};
}
static void op_spans(const TailOp &op, int64_t xlen, std::vector<Span> &reads, Span &write) {
    reads.clear();
    write = span_of(op.y, op.row0, op.row1);
    if (op.kind == T_SELL) {
        reads.push_back(span_of(op.x, 0, xlen));                                   // gathers: anywhere in the vector
        if (op.mode == RESID || op.mode == JACOBI || op.mode == GS) reads.push_back(span_of(op.b, op.row0, op.row1));
        if (op.mode == JACOBI || op.mode == PROLONG) reads.push_back(span_of(op.aux, op.row0, op.row1));
    } else if (op.kind == T_DIAG_SCALE) {
        reads.push_back(span_of(op.b, op.row0, op.row1));
        reads.push_back(span_of(op.aux, op.row0, op.row1));
    } else if (op.kind == T_COPY) {
        reads.push_back(span_of(op.x, op.row0, op.row1));
    }
}

// A barrier is needed between two operations exactly when one writes what the other reads or writes.  Operations are
// grouped greedily: an operation joins the current barrier-free group unless it conflicts with ANY member.
static void place_barriers(std::vector<TailOp> &ops, const std::vector<int64_t> &xlen) {
    std::vector<Span> group_reads, group_writes, reads;
    Span write;
    for (size_t k = 0; k < ops.size(); ++k) {
        op_spans(ops[k], xlen[k], reads, write);
        bool conflict = false;
        for (const Span &w : group_writes) {
            if (overlap(w, write)) conflict = true;
            for (const Span &r : reads)
                if (overlap(w, r)) conflict = true;
        }
        for (const Span &r : group_reads)
            if (overlap(r, write)) conflict = true;
        // an operation whose own gathers overlap its own output is in place (one colour of Gauss-Seidel): its rows do
        // not read each other, which is the colouring's contract, not something to detect here
        if (k > 0) ops[k - 1].sync_after = conflict ? 1 : 0;
        if (conflict) {
            group_reads.clear();
            group_writes.clear();
        }
        group_reads.insert(group_reads.end(), reads.begin(), reads.end());
        group_writes.push_back(write);
    }
    if (!ops.empty()) ops.back().sync_after = 1;
    if (g_tail_drop_barriers)
        for (size_t k = 0; k + 1 < ops.size(); ++k) ops[k].sync_after = 0;
}

static int tail_grid() {
    static int per_sm = -1;
    if (per_sm < 0) {
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tail_kernel, kBlock, 0) != cudaSuccess || occ < 1) occ = 1;
        per_sm = occ;
    }
    int k = g_tail_ctas_per_sm < 1 ? 1 : g_tail_ctas_per_sm;
    if (k > per_sm) k = per_sm;
    return k * sm_count();
}

static void tail_execute_host(std::vector<TailOp> &ops, uint64_t shuffle) {
    size_t g0 = 0;
    uint64_t state = shuffle * 0x9E3779B97F4A7C15ull + 1;
    while (g0 < ops.size()) {
        size_t g1 = g0;
        while (g1 + 1 < ops.size() && !ops[g1].sync_after) ++g1;
        ++g1;                                                        // group = [g0, g1)
        if (!shuffle) {
            for (size_t k = g0; k < g1; ++k)
                for (int64_t i = ops[k].row0; i < ops[k].row1; ++i) {
                    if (ops[k].kind == T_SELL) tail_row_host(ops[k], i);
                    else tail_vector_element(ops[k], i);
                }
        } else {
            // all (operation, row) items of the group in a pseudo-random order: legal if the barriers are right
            std::vector<std::pair<int32_t, int64_t>> items;
            for (size_t k = g0; k < g1; ++k)
                for (int64_t i = ops[k].row0; i < ops[k].row1; ++i) items.push_back({(int32_t)k, i});
            for (size_t i = items.size(); i > 1; --i) {
                state ^= state << 13; state ^= state >> 7; state ^= state << 17;
                std::swap(items[i - 1], items[state % i]);
            }
            for (const auto &it : items) {
                if (ops[it.first].kind == T_SELL) tail_row_host(ops[it.first], it.second);
                else tail_vector_element(ops[it.first], it.second);
            }
        }
        g0 = g1;
    }
}

// run what has been recorded so far (one cooperative launch per kTailMaxOps operations) and empty the record
int tail_flush(cudaStream_t st) {
    if (!g_rec.active || g_rec.ops.empty()) return MG_OK;
    place_barriers(g_rec.ops, g_rec.xlen);
    g_rec.total_ops += (int64_t)g_rec.ops.size();
    for (size_t k = 0; k + 1 < g_rec.ops.size(); ++k) g_rec.total_syncs += g_rec.ops[k].sync_after;
    if (g_rec.host) {
        tail_execute_host(g_rec.ops, g_rec.shuffle);
        g_rec.ops.clear();
        g_rec.xlen.clear();
        return MG_OK;
    }
    const int grid = tail_grid();
    size_t k0 = 0;
    while (k0 < g_rec.ops.size()) {
        // cut at a barrier: a launch boundary is one
        size_t k1 = std::min(g_rec.ops.size(), k0 + (size_t)kTailMaxOps);
        if (k1 < g_rec.ops.size()) {
            size_t cut = k1;
            while (cut > k0 + 1 && !g_rec.ops[cut - 1].sync_after) --cut;
            if (g_rec.ops[cut - 1].sync_after) k1 = cut;
            else g_rec.ops[k1 - 1].sync_after = 1;      // a group longer than a launch: conflict-free, so any cut is legal
        }
        TailProgram prog;
        memset(&prog, 0, sizeof(prog));
        prog.nops = (int32_t)(k1 - k0);
        for (size_t k = k0; k < k1; ++k) prog.ops[k - k0] = g_rec.ops[k];
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)grid, 1, 1);
        cfg.blockDim = dim3(kBlock, 1, 1);
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, tail_kernel, prog);
        MG_CHECK_LAUNCH("tail_kernel");
        ++g_rec.total_launches;
        k0 = k1;
    }
    g_rec.ops.clear();
    g_rec.xlen.clear();
    return MG_OK;
}

}  // namespace mgb

using namespace mgb;

extern "C" {

static int64_t g_tail_epoch = 0;
int64_t mg_set_tail_max_rows(int64_t rows) {
    const int64_t prev = g_tail_max_rows;
    g_tail_max_rows = rows < 0 ? 0 : rows;
    ++g_tail_epoch;
    if (g_tail_max_rows > 0) {
        // load the kernel now (lazy module loading must not happen while a peer's exchange kernel is spinning,
        // see mg_comm in mgb200.h); without a device this fails harmlessly
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, tail_kernel) != cudaSuccess) cudaGetLastError();
    }
    return prev;
}
int mg_set_tail_ctas_per_sm(int ctas) {
    const int prev = g_tail_ctas_per_sm;
    g_tail_ctas_per_sm = ctas < 1 ? 1 : ctas;
    ++g_tail_epoch;
    return prev;
}
int64_t mg_tail_config_epoch(void) { return g_tail_epoch; }
int mg_tail_debug_drop_barriers(int on) {
    const int prev = g_tail_drop_barriers;
    g_tail_drop_barriers = on ? 1 : 0;
    return prev;
}
int mg_tail_last_stats(int64_t *ops, int64_t *barriers, int64_t *launches) {
    tail_stats(ops, barriers, launches);
    return MG_OK;
}

}  // extern "C"
