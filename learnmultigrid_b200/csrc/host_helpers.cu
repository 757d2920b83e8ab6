// host_helpers.cu -- serial host-side setup helpers (host pointers only; no device work).
#include <algorithm>
#include <vector>
#include "common.cuh"

using namespace mgb;

extern "C" {

// First-fit greedy colouring on the symmetrised pattern of A (structural entries (i,j) or (j,i), j != i).
// Rows that have off-diagonal entries are coloured first, in index order; rows with only a diagonal entry
// (row-replaced Dirichlet rows, which other rows may still reference) are coloured afterwards.  On the 5-point
// grid in row-major numbering this is exactly red-black; on the 7-point P1 pattern it yields 3 colours.
//
// The worker colours a ROW BLOCK: columns < n_own are the block's own rows, columns n_own + k are external nodes whose
// colours (-1 = not coloured yet) come in through ext_color.  forbidden_{lo,hi}[n_own + n_ext] carry, per node, the
// colours taken by already coloured rows that reach the node only through THEIR row (entry (j,i) without (i,j): the
// pattern need not be symmetric): in for the own rows (pushes received from other blocks), out for the external ones
// (pushes this block makes).  phase 0 colours the rows with off-diagonal entries, phase 1 the rest, both bits set: both.
// With n_ext = 0 and both phases this is the whole-matrix colouring; partition_setup.py drives it block by block in
// rank order and obtains the same colours without any process holding the global pattern.
static int greedy_color_worker(int64_t n_own, int64_t n_ext, const int32_t *h_indptr, const int32_t *h_indices,
                               const int32_t *ext_color, uint64_t *forbidden_lo, uint64_t *forbidden_hi,
                               int32_t *h_colors, int phases, bool fresh) {
    std::vector<int64_t> deferred;                 // rows with nothing but a diagonal entry, coloured last
    if (fresh)
        for (int64_t i = 0; i < n_own; ++i) h_colors[i] = -1;
    int ncolors = 0;
    auto color_of = [&](int32_t j) -> int32_t { return j < n_own ? h_colors[j] : ext_color[j - n_own]; };
    auto color_row = [&](int64_t i) -> int {
        const int32_t p0 = h_indptr[i], p1 = h_indptr[i + 1];
        uint64_t lo = forbidden_lo[i], hi = forbidden_hi ? forbidden_hi[i] : 0;
        for (int32_t p = p0; p < p1; ++p) {
            const int32_t j = h_indices[p];
            if (j == i) continue;
            const int32_t c = color_of(j);
            if (c >= 0) { if (c < 64) lo |= (1ull << c); else hi |= (1ull << (c - 64)); }
        }
        int c;
        if (~lo) c = __builtin_ctzll(~lo);
        else if (forbidden_hi && ~hi) c = 64 + __builtin_ctzll(~hi);
        else return -1;
        h_colors[i] = c;
        if (c + 1 > ncolors) ncolors = c + 1;
        for (int32_t p = p0; p < p1; ++p) {
            const int32_t j = h_indices[p];
            if (j != i && color_of(j) < 0) {
                if (c < 64) forbidden_lo[j] |= (1ull << c); else forbidden_hi[j] |= (1ull << (c - 64));
            }
        }
        return c;
    };
    for (int64_t i = 0; i < n_own; ++i) {
        const int32_t p0 = h_indptr[i], p1 = h_indptr[i + 1];
        bool diag_only = true;
        for (int32_t p = p0; p < p1 && diag_only; ++p) diag_only = h_indices[p] == i;
        if (diag_only) { deferred.push_back(i); continue; }
        if (!(phases & 1)) continue;
        if (color_row(i) < 0) return -1;
    }
    if (phases & 2)
        for (int64_t i : deferred)
            if (color_row(i) < 0) return -1;
    for (int64_t i = 0; i < n_own; ++i)
        if (h_colors[i] + 1 > ncolors) ncolors = h_colors[i] + 1;
    return ncolors;
}

// The whole-matrix colouring while 64 colours suffice (no external nodes, one mask word per row): the same rule as the
// worker above with the row's two loops stripped of everything that case does not need.  Returns -1 as soon as a 65th
// colour would be needed; the caller then starts over with the general worker and two mask words.
static int greedy_color_whole_64(int64_t n, const int32_t *h_indptr, const int32_t *h_indices, uint64_t *forbidden,
                                 int32_t *h_colors) {
    std::vector<int64_t> deferred;
    for (int64_t i = 0; i < n; ++i) h_colors[i] = -1;
    uint64_t used = 0;
    auto color_row = [&](int64_t i, int32_t p0, int32_t p1, uint64_t mask) -> bool {
        if (!~mask) return false;
        const int c = __builtin_ctzll(~mask);
        h_colors[i] = c;
        used |= 1ull << c;
        const uint64_t bit = 1ull << c;
        for (int32_t p = p0; p < p1; ++p) {
            const int32_t j = h_indices[p];
            if (h_colors[j] < 0) forbidden[j] |= bit;          // j == i is coloured by now
        }
        return true;
    };
    for (int64_t i = 0; i < n; ++i) {
        const int32_t p0 = h_indptr[i], p1 = h_indptr[i + 1];
        uint64_t mask = forbidden[i];
        bool offdiag = false;
        for (int32_t p = p0; p < p1; ++p) {
            const int32_t j = h_indices[p];
            const int32_t c = h_colors[j];                      // -1 for j == i (not coloured yet) and for later rows
            offdiag |= j != i;
            mask |= (uint64_t)(c >= 0) << (c & 63);
        }
        if (!offdiag) { deferred.push_back(i); continue; }
        if (!color_row(i, p0, p1, mask)) return -1;
    }
    for (int64_t i : deferred)
        if (!color_row(i, h_indptr[i], h_indptr[i + 1], forbidden[i])) return -1;
    return used ? 64 - __builtin_clzll(used) : 0;
}

// Returns the number of colours (>0) or a negative status.
int mg_host_greedy_color(int64_t n, const int32_t *h_indptr, const int32_t *h_indices, int32_t *h_colors) {
    MG_REQUIRE(n >= 0 && h_indptr && h_colors, "null argument");
    // one 64-bit word per row; the second word (colours 64..127) only if a 65th colour is ever needed
    std::vector<uint64_t> lo((size_t)n, 0), hi;
    int nc = greedy_color_whole_64(n, h_indptr, h_indices, lo.data(), h_colors);
    if (nc < 0) {
        std::fill(lo.begin(), lo.end(), 0);
        hi.assign((size_t)n, 0);
        nc = greedy_color_worker(n, 0, h_indptr, h_indices, nullptr, lo.data(), hi.data(), h_colors, 3, true);
    }
    if (nc < 0) return set_error(MG_ERR_UNSUPPORTED, "mg_host_greedy_color", "more than 128 colours needed");
    return nc;
}

// One phase of the same colouring on a row block (see greedy_color_worker): columns are block-local (own rows
// 0..n_own-1, external node k at n_own + k); h_forbidden_lo / _hi have n_own + n_ext words each (in/out); phase 0 =
// rows with off-diagonal entries (h_colors is reset first), phase 1 = the remaining rows.  Returns the largest colour
// used in the block + 1 (possibly 0), or a negative status.
int mg_host_greedy_color_block(int64_t n_own, int64_t n_ext, const int32_t *h_indptr, const int32_t *h_indices,
                               const int32_t *h_ext_color, uint64_t *h_forbidden_lo, uint64_t *h_forbidden_hi,
                               int32_t *h_colors, int phase) {
    MG_REQUIRE(n_own >= 0 && n_ext >= 0 && h_indptr && h_colors && h_forbidden_lo && h_forbidden_hi &&
                   (n_ext == 0 || h_ext_color) && (phase == 0 || phase == 1), "bad argument");
    const int nc = greedy_color_worker(n_own, n_ext, h_indptr, h_indices, h_ext_color, h_forbidden_lo, h_forbidden_hi,
                                       h_colors, phase == 0 ? 1 : 2, phase == 0);
    if (nc < 0) return set_error(MG_ERR_UNSUPPORTED, "mg_host_greedy_color_block", "more than 128 colours needed");
    return nc;
}

// Dependency level of each row for an index-order Gauss-Seidel sweep on the symmetrised pattern:
// level(i) = 1 + max level(j) over coupled j < i (0 if none).  Returns the number of levels.
int64_t mg_host_lex_levels(int64_t n, const int32_t *h_indptr, const int32_t *h_indices, int32_t *h_level) {
    if (n < 0 || !h_indptr || !h_level) return set_error(MG_ERR_INVALID, "mg_host_lex_levels", "null argument");
    std::vector<int32_t> pend((size_t)n, 0);
    int64_t nlev = 0;
    for (int64_t i = 0; i < n; ++i) {
        int32_t lv = pend[i];
        for (int32_t p = h_indptr[i]; p < h_indptr[i + 1]; ++p) {
            const int32_t j = h_indices[p];
            if (j < i && h_level[j] + 1 > lv) lv = h_level[j] + 1;
        }
        h_level[i] = lv;
        if (lv + 1 > nlev) nlev = lv + 1;
        for (int32_t p = h_indptr[i]; p < h_indptr[i + 1]; ++p) {
            const int32_t j = h_indices[p];
            if (j > i && lv + 1 > pend[j]) pend[j] = lv + 1;
        }
    }
    return nlev;
}

}  // extern "C"
