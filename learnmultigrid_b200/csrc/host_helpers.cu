// host_helpers.cu -- serial host-side setup helpers (host pointers only; no device work).
#include <vector>
#include "common.cuh"

using namespace mgb;

extern "C" {

// First-fit greedy colouring on the symmetrised pattern of A (structural entries (i,j) or (j,i), j != i).
// Rows that have off-diagonal entries are coloured first, in index order; rows with only a diagonal entry
// (row-replaced Dirichlet rows, which other rows may still reference) are coloured afterwards.  On the 5-point
// grid in row-major numbering this is exactly red-black; on the 7-point P1 pattern it yields 3 colours.
// Returns the number of colours (>0) or a negative status.
int mg_host_greedy_color(int64_t n, const int32_t *h_indptr, const int32_t *h_indices, int32_t *h_colors) {
    MG_REQUIRE(n >= 0 && h_indptr && h_colors, "null argument");
    // forbidden[i]: colours already taken by neighbours that were coloured BEFORE row i and reach it only through their
    // own row (entry (j,i) without (i,j): the pattern need not be symmetric).  One 64-bit word per row; the second word
    // (colours 64..127) is only allocated if a 65th colour is ever needed.
    std::vector<uint64_t> forbidden_lo((size_t)n, 0), forbidden_hi;
    std::vector<int64_t> deferred;                 // rows with nothing but a diagonal entry, coloured last
    for (int64_t i = 0; i < n; ++i) h_colors[i] = -1;
    int ncolors = 0;
    auto color_row = [&](int64_t i) -> int {
        const int32_t p0 = h_indptr[i], p1 = h_indptr[i + 1];
        uint64_t lo = forbidden_lo[i], hi = forbidden_hi.empty() ? 0 : forbidden_hi[i];
        for (int32_t p = p0; p < p1; ++p) {
            const int32_t j = h_indices[p];
            if (j == i) continue;
            const int32_t c = h_colors[j];
            if (c >= 0) { if (c < 64) lo |= (1ull << c); else hi |= (1ull << (c - 64)); }
        }
        int c;
        if (~lo) c = __builtin_ctzll(~lo);
        else if (~hi) c = 64 + __builtin_ctzll(~hi);
        else return -1;
        h_colors[i] = c;
        if (c + 1 > ncolors) ncolors = c + 1;
        if (c >= 64 && forbidden_hi.empty()) forbidden_hi.assign((size_t)n, 0);
        for (int32_t p = p0; p < p1; ++p) {
            const int32_t j = h_indices[p];
            if (j != i && h_colors[j] < 0) {
                if (c < 64) forbidden_lo[j] |= (1ull << c); else forbidden_hi[j] |= (1ull << (c - 64));
            }
        }
        return c;
    };
    for (int64_t i = 0; i < n; ++i) {
        const int32_t p0 = h_indptr[i], p1 = h_indptr[i + 1];
        const bool diag_only = p1 == p0 || (p1 - p0 == 1 && h_indices[p0] == i) ||
                               [&] { for (int32_t p = p0; p < p1; ++p) if (h_indices[p] != i) return false; return true; }();
        if (diag_only) { deferred.push_back(i); continue; }
        if (color_row(i) < 0) return set_error(MG_ERR_UNSUPPORTED, "mg_host_greedy_color", "more than 128 colours needed");
    }
    for (int64_t i : deferred)
        if (color_row(i) < 0) return set_error(MG_ERR_UNSUPPORTED, "mg_host_greedy_color", "more than 128 colours needed");
    return ncolors;
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
