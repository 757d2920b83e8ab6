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
#include "common.cuh"

namespace mgb {

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

// The neighbours of row i are the entries of its row in A, then in A^T (with multiplicity, as jp_count_earlier counts
// them).  The two functions below visit entries lane, lane + G, ... of that list: the kernels give a row to a group of G
// lanes -- a serial walk is a chain of (index -> colour / priority -> atomic) round trips per neighbour, ~20 of them per
// 5-point row, and a round is nothing but that latency -- the serial host emulation calls them with lane 0 of 1.
// colours taken by the (already coloured) neighbours among this lane's entries: bit c of lo / bit c - 64 of hi
__host__ __device__ inline void jp_taken_colors(const JpView &v, int64_t i, int lane, int G, uint64_t &lo, uint64_t &hi) {
    const int32_t a0 = v.ip[i], da = v.ip[i + 1] - a0, b0 = v.tip[i], d = da + v.tip[i + 1] - b0;
    for (int32_t k = lane; k < d; k += G) {
        const int32_t j = k < da ? v.ix[a0 + k] : v.tix[b0 + k - da];
        const int32_t c = j != i ? v.color[j] : -1;
        if (c >= 0) { if (c < 64) lo |= 1ull << c; else hi |= 1ull << (c - 64); }
    }
}

// smallest colour not in the masks; -1: more than 128 colours would be needed
__host__ __device__ inline int jp_first_free(uint64_t lo, uint64_t hi) {
    if (~lo) return jp_lowest_zero_bit(lo);
    if (~hi) return 64 + jp_lowest_zero_bit(hi);
    return -1;
}

// row i has its colour: every later neighbour (among this lane's entries) waits for one row less; those that wait for
// nobody any more go to `next`
__host__ __device__ inline void jp_release(const JpView &v, int64_t i, int lane, int G, int32_t *next, int32_t *next_cnt) {
    const uint32_t me = v.prio[i];
    const int32_t a0 = v.ip[i], da = v.ip[i + 1] - a0, b0 = v.tip[i], d = da + v.tip[i + 1] - b0;
    for (int32_t k = lane; k < d; k += G) {
        const int32_t j = k < da ? v.ix[a0 + k] : v.tix[b0 + k - da];
        if (j == i || v.prio[j] <= me) continue;
#ifdef __CUDA_ARCH__
        if (atomicSub(&v.wait[j], 1) == 1) next[atomicAdd(next_cnt, 1)] = j;
#else
        if (--v.wait[j] == 0) next[(*next_cnt)++] = j;
#endif
    }
}

// counters in the workspace: [0..2] work-list sizes (round r reads r % 3, fills (r+1) % 3, clears (r+2) % 3),
// [3] rows coloured so far, [4] error flag (a row needed a 129th colour)
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

// One round: every row of the current work list takes its colour and releases its later neighbours.  A group of
// kJpGroup lanes shares a row (see above); the launches of a batch are programmatic dependents of each other, so that a
// round's launch latency hides behind the previous round.
constexpr int kJpGroup = 8;

__global__ void __launch_bounds__(kBlock)
jp_round_kernel(JpView v, const int32_t *cur, int32_t *next, int32_t *counters, int r) {
    pdl_prologue();
    constexpr int G = kJpGroup;
    int32_t *cur_cnt = counters + r % 3, *next_cnt = counters + (r + 1) % 3;
    const int32_t m = *cur_cnt;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        counters[(r + 2) % 3] = 0;                       // read in round r - 1, filled in round r + 1
        atomicAdd(&counters[3], m);
    }
    const int lane = threadIdx.x % G;
    const unsigned gmask = ((1u << G) - 1u) << ((threadIdx.x & 31) / G * G);
    const int64_t stride = (int64_t)gridDim.x * (kBlock / G);
    for (int64_t t = (int64_t)blockIdx.x * (kBlock / G) + threadIdx.x / G; t < m; t += stride) {
        const int64_t i = cur[t];
        uint64_t lo = 0, hi = 0;
        jp_taken_colors(v, i, lane, G, lo, hi);
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            lo |= __shfl_xor_sync(gmask, lo, o, G);
            hi |= __shfl_xor_sync(gmask, hi, o, G);
        }
        int c = jp_first_free(lo, hi);
        if (c < 0) { counters[4] = 1; c = 127; }
        if (lane == 0) v.color[i] = c;
        jp_release(v, i, lane, G, next, next_cnt);
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
    // kBlock / kJpGroup rows per CTA; as many CTAs as one work list can fill, a few per SM at most
    int64_t cap = (int64_t)sm_count() * 8;
    const int64_t want = (n + kBlock / kJpGroup - 1) / (kBlock / kJpGroup);
    if (cap > want) cap = want;
    const int64_t batch = 128;
    int64_t r = 0;
    int32_t host[kJpCounters];
    for (;;) {
        for (int64_t k = 0; k < batch; ++k, ++r)
            launch_k(jp_round_kernel, (unsigned)cap, (unsigned)kBlock, st, v, (const int32_t *)list[r & 1], list[(r + 1) & 1],
                     counters, (int)(r % 3));
        MG_CHECK_LAUNCH("jp_round");
        MG_CHECK_CUDA(cudaMemcpyAsync(host, counters, sizeof(host), cudaMemcpyDeviceToHost, st));
        MG_CHECK_CUDA(cudaStreamSynchronize(st));
        if (host[4]) return set_error(MG_ERR_UNSUPPORTED, "mg_color_first_fit", "more than 128 colours needed");
        if (host[3] >= n) break;
        if (r >= max_rounds) return set_error(MG_ERR_UNSUPPORTED, "mg_color_first_fit", "dependency chains too long for the round-based colouring");
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
            uint64_t lo = 0, hi = 0;
            jp_taken_colors(v, i, 0, 1, lo, hi);
            const int c = jp_first_free(lo, hi);
            if (c < 0) return set_error(MG_ERR_UNSUPPORTED, "mg_host_color_rounds", "more than 128 colours needed");
            v.color[i] = c;
        }
        for (int32_t t = cnt - 1; t >= 0; --t) jp_release(v, list[r & 1][t], 0, 1, list[(r + 1) & 1], &next_cnt);
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
