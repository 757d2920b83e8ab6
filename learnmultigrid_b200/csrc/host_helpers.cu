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
    std::vector<uint64_t> forbidden_lo((size_t)n, 0), forbidden_hi((size_t)n, 0);
    for (int64_t i = 0; i < n; ++i) h_colors[i] = -1;
    int ncolors = 0;
    for (int pass = 0; pass < 2; ++pass) {
        for (int64_t i = 0; i < n; ++i) {
            bool has_offdiag = false;
            for (int32_t p = h_indptr[i]; p < h_indptr[i + 1]; ++p)
                if (h_indices[p] != i) { has_offdiag = true; break; }
            if (has_offdiag != (pass == 0)) continue;
            uint64_t lo = forbidden_lo[i], hi = forbidden_hi[i];
            for (int32_t p = h_indptr[i]; p < h_indptr[i + 1]; ++p) {
                const int32_t j = h_indices[p];
                const int32_t c = (j != i) ? h_colors[j] : -1;
                if (c >= 0) { if (c < 64) lo |= (1ull << c); else hi |= (1ull << (c - 64)); }
            }
            int c;
            if (~lo) c = __builtin_ctzll(~lo);
            else if (~hi) c = 64 + __builtin_ctzll(~hi);
            else return set_error(MG_ERR_UNSUPPORTED, "mg_host_greedy_color", "more than 128 colours needed");
            h_colors[i] = c;
            if (c + 1 > ncolors) ncolors = c + 1;
            for (int32_t p = h_indptr[i]; p < h_indptr[i + 1]; ++p) {
                const int32_t j = h_indices[p];
                if (j != i && h_colors[j] < 0) {
                    if (c < 64) forbidden_lo[j] |= (1ull << c); else forbidden_hi[j] |= (1ull << (c - 64));
                }
            }
        }
    }
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
