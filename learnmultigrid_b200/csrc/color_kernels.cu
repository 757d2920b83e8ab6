// color_kernels.cu -- the first-fit colouring of mg_host_greedy_color (host_helpers.cu) computed on the device.
//
// First-fit in a fixed order is inherently sequential along dependency chains, but only along them: a row can take its
// colour as soon as every neighbour that comes EARLIER in the order has one (Jones-Plassmann with the place in the
// order as priority).  The order is the host helper's: rows with off-diagonal entries by index, then the rows with
// nothing but a diagonal entry by index; neighbours are taken on the symmetrised pattern (rows of A and of A^T).  So
//     colour(i) = smallest colour no earlier neighbour has
// is evaluated in ROUNDS: a round colours every row whose earlier neighbours are all coloured (they were coloured in
// previous rounds, so their colours are final and visible), then releases the later neighbours; a neighbour whose
// counter drops to zero joins the next round's work list.  The number of rounds is the longest dependency chain -- about
// 2 W on a W x W grid in row-major numbering (anti-diagonal wavefronts), a few hundred on unstructured numberings -- and
// the total work is O(nnz).  A 1D chain numbered end to end degenerates to n rounds; the host then falls back to the
// serial helper (round limit).  The colours are those of mg_host_greedy_color, entry for entry: the per-row code is
// shared by the kernels and by a serial host emulation (mg_host_color_rounds), which tests/test_host_logic.py compares
// with the serial first-fit.
#include <cooperative_groups.h>
#include "common.cuh"

namespace mgb {

// frontiers of at most this many rows are walked by the persistent cluster kernel (mg_set_color_cluster_frontier)
int64_t g_color_cluster_frontier = 32768;

struct JpView {
    int64_t n;
    const int32_t *ip, *ix;     // pattern of A
    const int32_t *tip, *tix;   // pattern of A^T
    uint32_t *prio;             // place in the first-fit order: i, or n + i for rows with nothing but a diagonal entry
    int32_t *wait;              // earlier neighbours (with multiplicity over both patterns) not coloured yet
    int32_t *color;
};

__host__ __device__ inline int jp_lowest_zero_bit(uint64_t w) {   // w != all ones
#ifdef __CUDA_ARCH__
    return __ffsll((long long)~w) - 1;
#else
    return __builtin_ctzll(~w);
#endif
}

// Colours and work-list entries are written by one SM and read by another INSIDE one kernel when the rounds run in the
// persistent cluster kernel below: those loads bypass L1 (ld.global.cg), which is not coherent between SMs.
template <typename T>
__host__ __device__ inline T jp_load_shared(const T *p) {
#ifdef __CUDA_ARCH__
    return __ldcg(p);
#else
    return *p;
#endif
}

__host__ __device__ inline uint32_t jp_priority(const JpView &v, int64_t i) {
    bool diag_only = true;
    for (int32_t p = v.ip[i]; p < v.ip[i + 1] && diag_only; ++p) diag_only = v.ix[p] == i;
    return (uint32_t)i + (diag_only ? (uint32_t)v.n : 0u);
}

__host__ __device__ inline int jp_count_earlier(const JpView &v, int64_t i) {
    const uint32_t me = v.prio[i];
    int w = 0;
    for (int32_t p = v.ip[i]; p < v.ip[i + 1]; ++p) {
        const int32_t j = v.ix[p];
        if (j != i && v.prio[j] < me) ++w;
    }
    for (int32_t p = v.tip[i]; p < v.tip[i + 1]; ++p) {
        const int32_t j = v.tix[p];
        if (j != i && v.prio[j] < me) ++w;
    }
    return w;
}

// smallest colour none of the (already coloured) neighbours has; -1: more than 128 colours would be needed
__host__ __device__ inline int jp_first_free(const JpView &v, int64_t i) {
    uint64_t lo = 0, hi = 0;
    for (int32_t p = v.ip[i]; p < v.ip[i + 1]; ++p) {
        const int32_t j = v.ix[p];
        const int32_t c = j != i ? jp_load_shared(v.color + j) : -1;
        if (c >= 0) { if (c < 64) lo |= 1ull << c; else hi |= 1ull << (c - 64); }
    }
    for (int32_t p = v.tip[i]; p < v.tip[i + 1]; ++p) {
        const int32_t j = v.tix[p];
        const int32_t c = j != i ? jp_load_shared(v.color + j) : -1;
        if (c >= 0) { if (c < 64) lo |= 1ull << c; else hi |= 1ull << (c - 64); }
    }
    if (~lo) return jp_lowest_zero_bit(lo);
    if (~hi) return 64 + jp_lowest_zero_bit(hi);
    return -1;
}

// row i has its colour: every later neighbour waits for one row less; those that wait for nobody any more go to `next`
__host__ __device__ inline void jp_release(const JpView &v, int64_t i, int32_t *next, int32_t *next_cnt) {
    const uint32_t me = v.prio[i];
    for (int pass = 0; pass < 2; ++pass) {
        const int32_t *ptr = pass ? v.tip : v.ip, *idx = pass ? v.tix : v.ix;
        for (int32_t p = ptr[i]; p < ptr[i + 1]; ++p) {
            const int32_t j = idx[p];
            if (j == i || v.prio[j] <= me) continue;
#ifdef __CUDA_ARCH__
            if (atomicSub(&v.wait[j], 1) == 1) next[atomicAdd(next_cnt, 1)] = j;
#else
            if (--v.wait[j] == 0) next[(*next_cnt)++] = j;
#endif
        }
    }
}

// counters in the workspace: [0..2] work-list sizes (round r reads r % 3, fills (r+1) % 3, clears (r+2) % 3),
// [3] rows coloured so far, [4] error flag (a row needed a 129th colour)
// [5] rounds the last cluster launch walked, [6] the frontier it stopped at
constexpr int kJpCounters = 8;

__global__ void __launch_bounds__(kBlock) jp_priority_kernel(JpView v) {
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= v.n) return;
    v.prio[i] = jp_priority(v, i);
    v.color[i] = -1;
}

__global__ void __launch_bounds__(kBlock) jp_wait_kernel(JpView v, int32_t *list0, int32_t *counters) {
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= v.n) return;
    const int w = jp_count_earlier(v, i);
    v.wait[i] = w;
    if (w == 0) list0[atomicAdd(&counters[0], 1)] = (int32_t)i;
}

__global__ void __launch_bounds__(kBlock)
jp_round_kernel(JpView v, const int32_t *cur, int32_t *next, int32_t *counters, int r) {
    int32_t *cur_cnt = counters + r % 3, *next_cnt = counters + (r + 1) % 3;
    const int32_t m = *cur_cnt;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        counters[(r + 2) % 3] = 0;                       // read in round r - 1, filled in round r + 1
        atomicAdd(&counters[3], m);
    }
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t t = (int64_t)blockIdx.x * kBlock + threadIdx.x; t < m; t += stride) {
        const int64_t i = cur[t];
        int c = jp_first_free(v, i);
        if (c < 0) { counters[4] = 1; c = 127; }
        v.color[i] = c;
        jp_release(v, i, next, next_cnt);
    }
}

// The same rounds without a launch per round.  On a W x W grid in row-major numbering a round is one anti-diagonal --
// at most W rows -- and there are 2 W of them per level (32 k rounds for the 8193^2 hierarchy): a launch per round is
// all latency (18 us per round measured, 0.6 s).  Here ONE thread-block cluster of 8 x 1024 threads walks the rounds: a
// round is one pass over the work list, then the hardware cluster barrier (barrier.cluster, a few hundred ns) in place
// of the kernel boundary.  Everything a later round reads from an earlier one (colours, work lists, counters) is
// written through to L2 and read with L1 bypassed (jp_load_shared / atomics), with a device-scope fence before the
// barrier.  The kernel hands back to the host loop when the frontier outgrows it (unstructured numberings start with
// n/6 independent rows: the wide grid of jp_round_kernel is the right tool there), when the work lists run dry, or after
// `limit` rounds.  Round numbering, list parity and counter rotation are those of jp_round_kernel, so the two can
// alternate.
constexpr int kJpClusterCtas = 8;
constexpr int kJpClusterThreads = 1024;

__global__ void __cluster_dims__(kJpClusterCtas, 1, 1) __launch_bounds__(kJpClusterThreads)
jp_cluster_rounds_kernel(JpView v, int32_t *list0, int32_t *list1, int32_t *counters, int r0, int limit, int big) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = (int)blockIdx.x * kJpClusterThreads + (int)threadIdx.x;
    constexpr int kThreads = kJpClusterCtas * kJpClusterThreads;
    int done = 0, m = 0;
    for (int r = r0; done < limit; ++r, ++done) {      // r0 is the global round number mod 6: parity and mod 3 survive
        m = jp_load_shared(counters + r % 3);
        if (m == 0 || m > big) break;                  // the same value in every thread: written before the last barrier
        const int32_t *cur = (r & 1) ? list1 : list0;
        int32_t *next = (r & 1) ? list0 : list1;
        int32_t *next_cnt = counters + (r + 1) % 3;
        if (tid == 0) {
            atomicExch(counters + (r + 2) % 3, 0);     // read in round r - 1 (before its barrier), filled in round r + 1
            atomicAdd(counters + 3, m);
        }
        for (int t = tid; t < m; t += kThreads) {
            const int64_t i = jp_load_shared(cur + t);
            int c = jp_first_free(v, i);
            if (c < 0) { atomicExch(counters + 4, 1); c = 127; }
            v.color[i] = c;
            jp_release(v, i, next, next_cnt);
        }
        __threadfence();
        cluster.sync();
    }
    if (tid == 0) {
        counters[5] = done;
        counters[6] = m;
    }
}

}  // namespace mgb

using namespace mgb;

namespace mgb {
__global__ void __launch_bounds__(kBlock)
csr_coloring_flags_kernel(int64_t nrows, int64_t row0, const int32_t *__restrict__ indptr,
                          const int32_t *__restrict__ indices, const double *__restrict__ values,
                          const int32_t *__restrict__ color, int32_t *__restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= nrows) return;
    const int64_t gi = row0 + i;
    const int32_t ci = color[gi];
    int bad = 2;
    for (int32_t p = indptr[i]; p < indptr[i + 1]; ++p) {
        if (values[p] == 0.0) continue;
        const int64_t j = indices[p];
        if (j == gi) bad &= ~2;
        else if (color[j] == ci) bad |= 1;
    }
    if (bad) atomicOr(flags, bad);
}
}  // namespace mgb

extern "C" {

static JpView jp_view(int64_t n, const int32_t *ip, const int32_t *ix, const int32_t *tip, const int32_t *tix, char *work,
                      int32_t *colors, int32_t **list0, int32_t **list1, int32_t **counters) {
    JpView v;
    v.n = n;
    v.ip = ip; v.ix = ix; v.tip = tip; v.tix = tix;
    v.prio = (uint32_t *)work;
    v.wait = (int32_t *)(work + 4 * n);
    v.color = colors;
    *list0 = (int32_t *)(work + 8 * n);
    *list1 = (int32_t *)(work + 12 * n);
    *counters = (int32_t *)(work + 16 * n);
    return v;
}

/* Work lists of at most `rows` rows are walked by the persistent cluster kernel, longer ones by one launch per round
 * (default 32768; 0 = always one launch per round).  Returns the previous value. */
int64_t mg_set_color_cluster_frontier(int64_t rows) {
    const int64_t old = g_color_cluster_frontier;
    if (rows >= 0) g_color_cluster_frontier = rows;
    return old;
}

/* bytes of workspace for n rows (priorities, counters, two work lists) */
int64_t mg_color_workspace_size(int64_t n) { return n < 0 ? -1 : 16 * n + (int64_t)sizeof(int32_t) * kJpCounters; }

/* d_colors[i] = colour of row i under mg_host_greedy_color's rule, from the patterns of A and A^T (device CSR).
 * Setup-time call: it synchronises the stream every `batch` rounds to look at the progress counter.
 * *h_rounds = rounds used.  Returns MG_ERR_UNSUPPORTED if more than max_rounds would be needed (the caller falls back
 * to the host helper) or if a row needs more than 128 colours. */
int mg_color_first_fit(int64_t n, const int32_t *d_indptr, const int32_t *d_indices, const int32_t *d_t_indptr,
                       const int32_t *d_t_indices, int32_t *d_colors, void *d_work, int64_t work_bytes,
                       int64_t max_rounds, int64_t *h_rounds, void *stream) {
    MG_REQUIRE(n >= 0 && n < (1ll << 31) && d_indptr && d_t_indptr && d_colors && d_work, "bad argument");
    MG_REQUIRE(work_bytes >= mg_color_workspace_size(n), "workspace too small");
    if (h_rounds) *h_rounds = 0;
    if (n == 0) return MG_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int32_t *list[2], *counters;
    JpView v = jp_view(n, d_indptr, d_indices, d_t_indptr, d_t_indices, (char *)d_work, d_colors, &list[0], &list[1], &counters);
    MG_CHECK_CUDA(cudaMemsetAsync(counters, 0, sizeof(int32_t) * kJpCounters, st));
    const unsigned rows_grid = (unsigned)((n + kBlock - 1) / kBlock);
    jp_priority_kernel<<<rows_grid, kBlock, 0, st>>>(v);
    MG_CHECK_LAUNCH("jp_priority");
    jp_wait_kernel<<<rows_grid, kBlock, 0, st>>>(v, list[0], counters);
    MG_CHECK_LAUNCH("jp_wait");
    int64_t cap = (int64_t)sm_count() * 4;
    if (cap > rows_grid) cap = rows_grid;
    const int64_t batch = 16;
    const int64_t big = g_color_cluster_frontier < 0x7fffffffLL ? g_color_cluster_frontier : 0x7fffffffLL;
    int64_t r = 0;
    int32_t host[kJpCounters];
    for (;;) {
        // the frontier of round r decides who walks the next rounds: short work lists (structured grids: one
        // anti-diagonal per round) go to the persistent cluster kernel, long ones to a launch per round on a wide grid
        MG_CHECK_CUDA(cudaMemcpyAsync(host, counters, sizeof(host), cudaMemcpyDeviceToHost, st));
        MG_CHECK_CUDA(cudaStreamSynchronize(st));
        if (host[4]) return set_error(MG_ERR_UNSUPPORTED, "mg_color_first_fit", "more than 128 colours needed");
        if (host[3] >= n) break;
        if (r >= max_rounds) return set_error(MG_ERR_UNSUPPORTED, "mg_color_first_fit", "dependency chains too long for the round-based colouring");
        const int64_t frontier = host[r % 3];
        if (frontier == 0)      // rows left but nobody ready: the transposed pattern does not match the pattern
            return set_error(MG_ERR_INVALID, "mg_color_first_fit", "rows left uncoloured (inconsistent transpose pattern?)");
        if (frontier <= big) {
            int64_t limit = max_rounds - r;
            if (limit > (1 << 20)) limit = 1 << 20;
            jp_cluster_rounds_kernel<<<kJpClusterCtas, kJpClusterThreads, 0, st>>>(v, list[0], list[1], counters, (int)(r % 6),
                                                                                  (int)limit, (int)big);
            MG_CHECK_LAUNCH("jp_cluster_rounds");
            ++g_launch_count;
            MG_CHECK_CUDA(cudaMemcpyAsync(host, counters, sizeof(host), cudaMemcpyDeviceToHost, st));
            MG_CHECK_CUDA(cudaStreamSynchronize(st));
            r += host[5];
        } else {
            for (int64_t k = 0; k < batch; ++k, ++r)
                jp_round_kernel<<<(unsigned)cap, kBlock, 0, st>>>(v, list[r & 1], list[(r + 1) & 1], counters, (int)(r % 3));
            MG_CHECK_LAUNCH("jp_round");
        }
    }
    if (h_rounds) *h_rounds = r;
    return MG_OK;
}

/* Is a colouring proper for the operator, and has every row a diagonal?  Rows [0,nrows) of a CSR block whose row i is
 * global row row0 + i; d_color is indexed by GLOBAL row / column id.  *d_flags (device int32, zeroed by the caller)
 * |= 1 if a row has a non-zero entry in the column of another row of its own colour, |= 2 if a row has no non-zero
 * diagonal entry.  This is what mg_level_inspect finds on a colour-blocked SELL level, asked of the operator in its
 * natural ordering -- the form in which a partitioned level can ask it about the GLOBAL colouring. */
int mg_csr_coloring_flags(int64_t nrows, int64_t row0, const int32_t *d_indptr, const int32_t *d_indices,
                          const double *d_values, const int32_t *d_color, int32_t *d_flags, void *stream) {
    MG_REQUIRE(nrows >= 0 && d_indptr && d_flags && d_color, "bad argument");
    if (nrows == 0) return MG_OK;
    csr_coloring_flags_kernel<<<(unsigned)((nrows + kBlock - 1) / kBlock), kBlock, 0, (cudaStream_t)stream>>>(
        nrows, row0, d_indptr, d_indices, d_values, d_color, d_flags);
    MG_CHECK_LAUNCH("csr_coloring_flags");
    return MG_OK;
}

#ifdef MGB_TESTING
/* The same rounds run serially on HOST arrays with the same per-row code (CPU test-suite; not called by the product).
 * h_work: mg_color_workspace_size(n) bytes.  *h_rounds = number of non-empty rounds. */
int mg_host_color_rounds(int64_t n, const int32_t *h_indptr, const int32_t *h_indices, const int32_t *h_t_indptr,
                         const int32_t *h_t_indices, int32_t *h_colors, void *h_work, int64_t *h_rounds) {
    MG_REQUIRE(n >= 0 && h_indptr && h_t_indptr && h_colors && h_work, "bad argument");
    int32_t *list[2], *counters;
    JpView v = jp_view(n, h_indptr, h_indices, h_t_indptr, h_t_indices, (char *)h_work, h_colors, &list[0], &list[1], &counters);
    for (int64_t i = 0; i < n; ++i) {
        v.prio[i] = jp_priority(v, i);
        v.color[i] = -1;
    }
    int32_t cnt = 0;
    for (int64_t i = 0; i < n; ++i) {
        v.wait[i] = jp_count_earlier(v, i);
        if (v.wait[i] == 0) list[0][cnt++] = (int32_t)i;
    }
    int64_t r = 0, done = 0;
    while (cnt > 0) {
        int32_t next_cnt = 0;
        // the rows of a round are independent: walk them backwards to make any hidden order dependence show
        for (int32_t t = cnt - 1; t >= 0; --t) {
            const int64_t i = list[r & 1][t];
            const int c = jp_first_free(v, i);
            if (c < 0) return set_error(MG_ERR_UNSUPPORTED, "mg_host_color_rounds", "more than 128 colours needed");
            v.color[i] = c;
        }
        for (int32_t t = cnt - 1; t >= 0; --t) jp_release(v, list[r & 1][t], list[(r + 1) & 1], &next_cnt);
        done += cnt;
        cnt = next_cnt;
        ++r;
    }
    if (h_rounds) *h_rounds = r;
    if (done != n) return set_error(MG_ERR_INVALID, "mg_host_color_rounds", "rows left uncoloured (inconsistent transpose pattern?)");
    return MG_OK;
}
#endif  // MGB_TESTING

}  // extern "C"
