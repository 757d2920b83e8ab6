// nn_kernels.cu -- the 2D "neural" transfer-operator builder of the reference on the device
// (NeuralMG_2D.define_hierarchy and its helpers, learn_multigrid/solvers/Multigrid.py:401-765).
//
// Per level the reference does, in Python loops over a DENSE copy of the mass matrix (Multigrid.py:404):
//   coarsening (:401-426)        greedy C/F splitting: take the smallest remaining index, drop its neighbours;
//   extract_patches (:591-677)   per coarse node 43 features (node, <= 6 neighbours, their rows, "virtual" padding) and
//                                31 fill indices (node, fine neighbours, coarse neighbours-of-neighbours, 6 x 3
//                                coarse nodes "intersecting" each fine neighbour);
//   model.predict                an MLP gives 31 values per patch (torch on the device; not here);
//   fill_B (:687-732)            scatter of the predictions into B with a running pairwise mean on collisions;
//   Q = B / rowsum(B), M <- Q^T M Q, pre_process (:735-739) cuts the coarse mass rows to the predicted neighbours.
// Here: one thread per coarse node / prediction, CSR input, no dense matrix.  The per-node logic is written once as
// __host__ __device__ functions; the mg_host_nn_* entry points run the same code serially on host arrays, which is how
// the CPU test-suite checks it against the golden vectors produced by the reference itself
// (tests/golden/neural_2d_cases.npz).  A coarse node with 6 + a neighbours (a > 0; never the case for Mesh2D meshes or
// for coarser levels, whose rows pre_process cuts to 6) gets a + 1 patch variants as in the reference (:645-663):
// variant v drops the neighbours ranked v .. v+a-1 by ascending mass entry; variant 0 sits in the node's own slot, the
// others are appended behind all regular patches in node order.  a <= 6 (beyond that the reference itself fails).
#include "common.cuh"

namespace mgb {

constexpr int kMaxRow = 16;        // longest row (entries incl. diagonal) the extraction accepts
constexpr int kMaxDegree = 12;     // most neighbours of a coarse node: 6 kept + up to 6 dropped per variant
constexpr int kPatch = 43, kFill = 31;
enum NnError { NN_OK = 0, NN_DEGREE = 1, NN_NEGATIVE = 2, NN_ISOLATED = 3, NN_LONG_ROW = 4 };

// scaling_vnodes (Multigrid.py:428-436): neighbour count -> (up, down)
__host__ __device__ inline bool nn_scaling(int neighs, double &up, double &down) {
    switch (neighs) {
        case 2: up = 6.0; down = 1.0; return true;
        case 3: up = 3.0; down = 2.0; return true;
        case 4: up = 2.0; down = 3.0; return true;
        case 5: up = 1.5; down = 4.0; return true;
        default: return false;
    }
}

// remove the first occurrence of the minimum until `keep` entries are left (Multigrid.py:473-477, 505-511, 522-528)
__host__ __device__ inline void nn_trim_min(double *v, int32_t *idx, int &len, int keep) {
    while (len > keep) {
        int pos = 0;
        for (int i = 1; i < len; ++i)
            if (v[i] < v[pos]) pos = i;
        for (int i = pos; i + 1 < len; ++i) {
            v[i] = v[i + 1];
            if (idx) idx[i] = idx[i + 1];
        }
        --len;
    }
}

__host__ __device__ inline bool nn_contains(const int32_t *a, int n, int32_t x) {
    for (int i = 0; i < n; ++i)
        if (a[i] == x) return true;
    return false;
}

// number of patch variants beyond the first for coarse node c: max(0, positive off-diagonal entries - 6)
// (Multigrid.py:606-614, 631-632)
__host__ __device__ inline int nn_extra_variants(const int32_t *indptr, const int32_t *indices, const double *values,
                                                 int32_t c) {
    int deg = 0;
    for (int32_t p = indptr[c]; p < indptr[c + 1]; ++p)
        if (indices[p] != c && values[p] > 0.0) ++deg;
    return deg > 6 ? deg - 6 : 0;
}

// single_extraction (Multigrid.py:545-589) for patch variant `variant` of coarse node c (0 for a node with at most 6
// neighbours).  Returns an NnError.
__host__ __device__ inline int nn_extract_one(const int32_t *indptr, const int32_t *indices, const double *values,
                                              const int32_t *cmap, int32_t c, int variant, double *patch,
                                              int32_t *fill) {
    for (int i = 0; i < kPatch; ++i) patch[i] = -1.0;
    for (int i = 0; i < kFill; ++i) fill[i] = -1;
    // row of c: `where` = columns of the nonzero entries other than c, `row` = the positive values (:622-629)
    int32_t where[kMaxDegree];
    double row[kMaxDegree];
    int nnb = 0;
    double node_M = 0.0;
    for (int32_t p = indptr[c]; p < indptr[c + 1]; ++p) {
        const int32_t j = indices[p];
        const double v = values[p];
        if (j == c) { node_M = v; continue; }
        if (v == 0.0) continue;
        if (v < 0.0) return NN_NEGATIVE;
        if (nnb >= kMaxDegree) return NN_DEGREE;
        where[nnb] = j;
        row[nnb] = v;
        ++nnb;
    }
    if (nnb > 6) {
        // `ordered = argsort(row)`; variant v deletes positions ordered[v : v + additional] (:636-644).  Ties are ranked
        // by position (NumPy's default sort does not promise an order for equal keys).
        const int additional = nnb - 6;
        if (variant < 0 || variant > additional) return NN_DEGREE;
        bool drop[kMaxDegree];
        for (int i = 0; i < nnb; ++i) {
            int rank = 0;
            for (int k = 0; k < nnb; ++k)
                if (row[k] < row[i] || (row[k] == row[i] && k < i)) ++rank;
            drop[i] = rank >= variant && rank < variant + additional;
        }
        int kept = 0;
        for (int i = 0; i < nnb; ++i)
            if (!drop[i]) { where[kept] = where[i]; row[kept] = row[i]; ++kept; }
        nnb = kept;
    } else if (variant != 0) {
        return NN_DEGREE;
    }
    double up = 0.0, down = 1.0;
    if (nnb < 6 && !nn_scaling(nnb, up, down)) return NN_ISOLATED;
    // direct_neighs (:456-490)
    int32_t Bl[6 * kMaxRow];
    int nB = 0;
    double *neighs = patch + 7;
    for (int k = 0; k < nnb; ++k) {
        const int32_t nb = where[k];
        double vals[kMaxRow];
        int nv = 0;
        double neigh_node = 0.0;
        if (indptr[nb + 1] - indptr[nb] > kMaxRow) return NN_LONG_ROW;
        for (int32_t p = indptr[nb]; p < indptr[nb + 1]; ++p) {
            const int32_t j = indices[p];
            const double v = values[p];
            if (j == c || v == 0.0) continue;             // row_neigh[0, c_node] = 0 ; nonzero()
            if (!nn_contains(Bl, nB, j)) Bl[nB++] = j;    // B.extend(...) then "remove duplicates keeping order"
            if (j == nb) { neigh_node = v; continue; }
            if (v < 0.0) return NN_NEGATIVE;
            vals[nv++] = v;
        }
        nn_trim_min(vals, nullptr, nv, 5);
        if (nv < 5) {
            double u2, d2;
            if (!nn_scaling(nv + 1, u2, d2)) return NN_ISOLATED;
            const double pad = neigh_node / d2;
            while (nv < 5) vals[nv++] = pad;
        }
        neighs[6 * k] = neigh_node;
        for (int i = 0; i < 5; ++i) neighs[6 * k + 1 + i] = vals[i];
    }
    // create_virtual_nodes (:438-454): pad the row and append whole virtual neighbour blocks
    patch[0] = node_M;
    for (int k = 0; k < nnb; ++k) patch[1 + k] = row[k];
    for (int k = nnb; k < 6; ++k) {
        patch[1 + k] = node_M / down;
        neighs[6 * k] = node_M * up;
        for (int i = 0; i < 5; ++i) neighs[6 * k + 1 + i] = node_M / down;
    }
    // fill indices (:577-579): node, fine neighbours, coarse nodes among the neighbours of the neighbours
    fill[0] = c;
    for (int k = 0; k < nnb; ++k) fill[1 + k] = where[k];
    int pos = 7;
    for (int i = 0; i < nB; ++i) {
        const int32_t x = Bl[i];
        if (x == c || nn_contains(where, nnb, x) || cmap[x] < 0) continue;
        if (pos < kFill) fill[pos] = x;
        ++pos;
    }
    // intersecting_rows (:492-543): for every fine neighbour the (last) three coarse nodes two rings away
    pos = 13;
    for (int k = 0; k < nnb; ++k) {
        const int32_t el = where[k];
        int32_t rn[kMaxRow];
        double rv[kMaxRow];
        int nr = 0;
        for (int32_t p = indptr[el]; p < indptr[el + 1]; ++p) {
            const int32_t j = indices[p];
            const double v = values[p];
            if (j == c || j == el || v == 0.0) continue;
            if (v < 0.0) return NN_NEGATIVE;
            rn[nr] = j;
            rv[nr] = v;
            ++nr;
        }
        nn_trim_min(rv, rn, nr, 5);
        int32_t wwc[5 * kMaxRow];
        int nw = 0;
        for (int q = 0; q < nr; ++q) {
            const int32_t kk = rn[q];
            int32_t wh[kMaxRow];
            double wv[kMaxRow];
            int nh = 0;
            if (indptr[kk + 1] - indptr[kk] > kMaxRow) return NN_LONG_ROW;
            for (int32_t p = indptr[kk]; p < indptr[kk + 1]; ++p) {
                const double v = values[p];
                if (v == 0.0) continue;
                if (v < 0.0) return NN_NEGATIVE;
                wh[nh] = indices[p];
                wv[nh] = v;
                ++nh;
            }
            nn_trim_min(wv, wh, nh, 7);
            for (int i = 0; i < nh; ++i) {
                const int32_t x = wh[i];
                if (x == c || cmap[x] < 0 || nn_contains(wwc, nw, x)) continue;    // unique, coarse, not c_node
                wwc[nw++] = x;
            }
        }
        const int first = nw > 3 ? nw - 3 : 0;                                      // `while len > 3: delete [0]`
        for (int i = first; i < nw; ++i) fill[pos + (i - first)] = wwc[i];
        pos += 3;
    }
    return NN_OK;
}

// fill_B (Multigrid.py:687-732): the contributions of prediction j, in the reference's application order.  Each is
// (fine row, coarse column, value); entries a prediction does not use get row = unused_row (>= number of rows, so
// that they sort behind everything).  Also d_neighs[node_coarse].
__host__ __device__ inline void nn_contributions_one(const int32_t *fill, const double *pred, const int32_t *cmap,
                                                     int32_t unused_row, int32_t *rows, int32_t *cols, double *vals,
                                                     int32_t *dn) {
    for (int i = 0; i < kFill; ++i) { rows[i] = unused_row; cols[i] = 0; vals[i] = 0.0; }
    const int32_t node = fill[0];
    const int32_t nc = cmap[node];
    int ncol = 0;
    int32_t colB[6];
    for (int k = 1; k < 7; ++k)
        if (fill[k] >= 0) colB[ncol++] = fill[k];
    int nrow = 0;
    int32_t rowB[6];
    for (int k = 7; k < 13; ++k)
        if (fill[k] >= 0 && cmap[fill[k]] >= 0) rowB[nrow++] = cmap[fill[k]];
    for (int k = 0; k < 6; ++k) dn[k] = k < nrow ? rowB[k] : -1;
    rows[0] = node; cols[0] = nc; vals[0] = pred[0];
    for (int k = 0; k < ncol; ++k) { rows[1 + k] = colB[k]; cols[1 + k] = nc; vals[1 + k] = pred[1 + k]; }
    for (int k = 0; k < nrow; ++k) { rows[7 + k] = node; cols[7 + k] = rowB[k]; vals[7 + k] = pred[7 + k]; }
    int pos = 13;
    for (int k = 0; k < ncol; ++k) {
        int t = 0;
        for (int i = 0; i < 3; ++i) {
            const int32_t w = fill[pos + i];
            if (w < 0) continue;
            rows[pos + t] = colB[k];
            cols[pos + t] = cmap[w];
            vals[pos + t] = pred[pos + t];
            ++t;
        }
        pos += 3;
    }
}

// ---- kernels ---------------------------------------------------------------------------------------------------------
// one round of the lexicographically-first independent set: node i becomes coarse once every j < i with M[j,i] > 0 is
// decided and none of them is coarse (rows of the TRANSPOSE are scanned: the reference drops the neighbours found in the
// ROW of the chosen node, and pre_process can make the pattern unsymmetric)
__global__ void __launch_bounds__(kBlock)
nn_mis_round_kernel(int64_t n, const int32_t *__restrict__ t_indptr, const int32_t *__restrict__ t_indices,
                    const double *__restrict__ t_values, const int32_t *__restrict__ state_in,
                    int32_t *__restrict__ state_out, int32_t *__restrict__ undecided) {
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    int32_t s = state_in[i];
    if (s == 0) {
        bool wait = false, drop = false;
        for (int32_t p = t_indptr[i]; p < t_indptr[i + 1]; ++p) {
            const int32_t j = t_indices[p];
            if (j >= i || !(t_values[p] > 0.0)) continue;
            const int32_t sj = state_in[j];
            if (sj == 1) { drop = true; break; }
            if (sj == 0) wait = true;
        }
        if (drop) s = 2;
        else if (!wait) s = 1;
        else atomicAdd(undecided, 1);
    }
    state_out[i] = s;
}

__global__ void __launch_bounds__(kBlock)
nn_compact_kernel(int64_t n, const int32_t *__restrict__ state, const int32_t *__restrict__ scan,
                  int32_t *__restrict__ cmap, int32_t *__restrict__ clist) {
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    if (state[i] == 1) {
        cmap[i] = scan[i];
        clist[scan[i]] = (int32_t)i;
    } else {
        cmap[i] = -1;
    }
}

__global__ void __launch_bounds__(128)
nn_extract_kernel(int64_t nc, const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                  const double *__restrict__ values, const int32_t *__restrict__ cmap,
                  const int32_t *__restrict__ clist, double *__restrict__ patches, int32_t *__restrict__ fill,
                  int32_t *__restrict__ err) {
    const int64_t k = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (k >= nc) return;
    double patch[kPatch];
    int32_t fi[kFill];
    const int rc = nn_extract_one(indptr, indices, values, cmap, clist[k], 0, patch, fi);
    if (rc != NN_OK) atomicCAS(err, 0, rc);
    for (int i = 0; i < kPatch; ++i) patches[k * kPatch + i] = patch[i];
    for (int i = 0; i < kFill; ++i) fill[k * kFill + i] = fi[i];
}

__global__ void __launch_bounds__(kBlock)
nn_count_variants_kernel(int64_t nc, const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                         const double *__restrict__ values, const int32_t *__restrict__ clist,
                         int32_t *__restrict__ extra) {
    const int64_t k = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (k >= nc) return;
    extra[k] = nn_extra_variants(indptr, indices, values, clist[k]);
}

// variants 1 .. a of the coarse nodes that have them, written behind the nc regular patches in node order;
// not_last[j] = 1 for every patch of such a node except its last variant (whose d_neighs entry is the one fill_B keeps)
__global__ void __launch_bounds__(128)
nn_extract_variants_kernel(int64_t nc, const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                           const double *__restrict__ values, const int32_t *__restrict__ cmap,
                           const int32_t *__restrict__ clist, const int32_t *__restrict__ extra_ptr,
                           double *__restrict__ patches, int32_t *__restrict__ fill, int32_t *__restrict__ not_last,
                           int32_t *__restrict__ err) {
    const int64_t k = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (k >= nc) return;
    const int32_t first = extra_ptr[k], additional = extra_ptr[k + 1] - first;
    if (additional == 0) return;
    not_last[k] = 1;
    double patch[kPatch];
    int32_t fi[kFill];
    for (int v = 1; v <= additional; ++v) {
        const int rc = nn_extract_one(indptr, indices, values, cmap, clist[k], v, patch, fi);
        if (rc != NN_OK) atomicCAS(err, 0, rc);
        const int64_t j = nc + first + (v - 1);
        for (int i = 0; i < kPatch; ++i) patches[j * kPatch + i] = patch[i];
        for (int i = 0; i < kFill; ++i) fill[j * kFill + i] = fi[i];
        not_last[j] = v < additional ? 1 : 0;
    }
}

__global__ void __launch_bounds__(128)
nn_contrib_kernel(int64_t np_, const int32_t *__restrict__ fill, const double *__restrict__ pred,
                  const int32_t *__restrict__ cmap, const int32_t *__restrict__ not_last, int32_t unused_row,
                  int32_t *__restrict__ rows, int32_t *__restrict__ cols, double *__restrict__ vals,
                  int32_t *__restrict__ dneigh) {
    const int64_t j = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (j >= np_) return;
    int32_t r[kFill], c[kFill], dn[6];
    double v[kFill];
    nn_contributions_one(fill + j * kFill, pred + j * kFill, cmap, unused_row, r, c, v, dn);
    for (int i = 0; i < kFill; ++i) {
        rows[j * kFill + i] = r[i];
        cols[j * kFill + i] = c[i];
        vals[j * kFill + i] = v[i];
    }
    if (not_last && not_last[j]) return;       // d_neighs[node] is overwritten by every patch of the node: the last one stays
    const int32_t ncoarse = cmap[fill[j * kFill]];
    for (int i = 0; i < 6; ++i) dneigh[(int64_t)ncoarse * 6 + i] = dn[i];
}

// contributions sorted by (row, col), application order kept inside a run: the first element of every run folds it
// with the reference's rule (empty or exactly 0 -> take the value, otherwise the mean of old and new) and counts
__global__ void __launch_bounds__(kBlock)
nn_fold_kernel(int64_t m, int32_t unused_row, const int32_t *__restrict__ rows, const int32_t *__restrict__ cols,
               const double *__restrict__ vals, const int32_t *__restrict__ order, int32_t *__restrict__ head,
               double *__restrict__ folded) {
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= m) return;
    const int32_t o = order[i];
    const int32_t r = rows[o], c = cols[o];
    head[i] = 0;
    if (r >= unused_row) return;
    if (i > 0) {
        const int32_t po = order[i - 1];
        if (rows[po] == r && cols[po] == c) return;
    }
    double cur = 0.0;
    for (int64_t k = i; k < m; ++k) {
        const int32_t ok = order[k];
        if (rows[ok] != r || cols[ok] != c) break;
        const double v = vals[ok];
        cur = (cur == 0.0) ? v : (cur + v) / 2.0;
    }
    head[i] = cur != 0.0 ? 1 : 0;          // the dense B of the reference has no entry where the value is 0
    folded[i] = cur;
}

// B entries (one per run head, in (row, col) order) -> CSR arrays; row_count is incremented per entry
__global__ void __launch_bounds__(kBlock)
nn_emit_kernel(int64_t m, const int32_t *__restrict__ rows, const int32_t *__restrict__ cols,
               const int32_t *__restrict__ order, const int32_t *__restrict__ head, const int32_t *__restrict__ slot,
               const double *__restrict__ folded, int32_t *__restrict__ out_rows, int32_t *__restrict__ out_cols,
               double *__restrict__ out_vals) {
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= m || !head[i]) return;
    const int32_t o = order[i];
    const int32_t s = slot[i];
    out_rows[s] = rows[o];
    out_cols[s] = cols[o];
    out_vals[s] = folded[i];
}

// Q = B / rowsum(B) (Multigrid.py:758-759); row sums in column order
__global__ void __launch_bounds__(kBlock)
nn_row_normalise_kernel(int64_t n, const int32_t *__restrict__ indptr, double *__restrict__ values) {
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int32_t p = indptr[i]; p < indptr[i + 1]; ++p) s += values[p];
    for (int32_t p = indptr[i]; p < indptr[i + 1]; ++p) values[p] = values[p] / s;
}

// pre_process (Multigrid.py:735-739): keep of row i only the diagonal and the columns listed in dneigh[i][0..5]
__global__ void __launch_bounds__(kBlock)
nn_cut_count_kernel(int64_t n, const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                    const double *__restrict__ values, const int32_t *__restrict__ dneigh, int32_t *__restrict__ keep,
                    int32_t *__restrict__ count) {
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    int32_t cnt = 0;
    for (int32_t p = indptr[i]; p < indptr[i + 1]; ++p) {
        const int32_t j = indices[p];
        bool k = (j == i);
        for (int q = 0; q < 6 && !k; ++q) k = dneigh[i * 6 + q] == j;
        k = k && values[p] != 0.0;          // lil_matrix assignment does not store zeros
        keep[p] = k ? 1 : 0;
        cnt += k ? 1 : 0;
    }
    count[i] = cnt;
}
__global__ void __launch_bounds__(kBlock)
nn_cut_fill_kernel(int64_t n, const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                   const double *__restrict__ values, const int32_t *__restrict__ keep,
                   const int32_t *__restrict__ out_indptr, int32_t *__restrict__ out_indices,
                   double *__restrict__ out_values) {
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    int32_t o = out_indptr[i];
    for (int32_t p = indptr[i]; p < indptr[i + 1]; ++p)
        if (keep[p]) {
            out_indices[o] = indices[p];
            out_values[o] = values[p];
            ++o;
        }
}

static const char *nn_error_text(int rc) {
    switch (rc) {
        case NN_DEGREE: return "a coarse node has more than 12 neighbours (more than 6 dropped per patch variant, Multigrid.py:636-644: the reference fails there too)";
        case NN_NEGATIVE: return "negative off-diagonal entry in the mass matrix";
        case NN_ISOLATED: return "a node has fewer than 2 neighbours (scaling_vnodes has no entry; the reference fails too)";
        case NN_LONG_ROW: return "a row has more than 16 entries";
        default: return "unknown";
    }
}

}  // namespace mgb

using namespace mgb;

extern "C" {

/* Greedy C/F splitting of NeuralMG_2D.coarsening (Multigrid.py:401-426) = the lexicographically first maximal
 * independent set of the graph {(i,j): M[i,j] > 0, i != j}.  Input: CSR of M^T (rows of the transpose = columns of M).
 * d_state (n int32) receives 1 for coarse, 2 for fine nodes; d_work: n + 1 int32.  Iterates rounds until every node is
 * decided (synchronises); returns the number of rounds (> 0) or a negative status. */
int mg_nn_coarsen(int64_t n, const int32_t *d_t_indptr, const int32_t *d_t_indices, const double *d_t_values,
                  int32_t *d_state, int32_t *d_work, void *stream) {
    MG_REQUIRE(n > 0 && d_t_indptr && d_t_indices && d_t_values && d_state && d_work, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    int32_t *other = d_work, *flag = d_work + n;
    MG_CHECK_CUDA(cudaMemsetAsync(d_state, 0, n * sizeof(int32_t), st));
    int32_t *cur = d_state, *nxt = other;
    const unsigned grid = (unsigned)((n + kBlock - 1) / kBlock);
    int rounds = 0;
    for (;;) {
        int32_t undecided = 0;
        for (int k = 0; k < 16; ++k) {            // a batch of rounds per host round trip
            MG_CHECK_CUDA(cudaMemsetAsync(flag, 0, sizeof(int32_t), st));
            nn_mis_round_kernel<<<grid, kBlock, 0, st>>>(n, d_t_indptr, d_t_indices, d_t_values, cur, nxt, flag);
            MG_CHECK_LAUNCH("nn_mis_round");
            int32_t *t = cur; cur = nxt; nxt = t;
            ++rounds;
        }
        MG_CHECK_CUDA(cudaMemcpyAsync(&undecided, flag, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        MG_CHECK_CUDA(cudaStreamSynchronize(st));
        if (undecided == 0) break;
        if (rounds > n + 16) return set_error(MG_ERR_INVALID, "mg_nn_coarsen", "no progress");
    }
    if (cur != d_state) MG_CHECK_CUDA(cudaMemcpyAsync(d_state, cur, n * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    MG_CHECK_CUDA(cudaStreamSynchronize(st));
    return rounds;
}

/* d_scan = exclusive scan of (state == 1) (mg_exclusive_scan_i32): cmap[i] = coarse index of node i or -1,
 * clist[k] = node of coarse index k (ascending = the reference's selection order, map_coarse :679-684) */
int mg_nn_compact(int64_t n, const int32_t *d_state, const int32_t *d_scan, int32_t *d_cmap, int32_t *d_clist,
                  void *stream) {
    MG_REQUIRE(n > 0 && d_state && d_scan && d_cmap && d_clist, "null argument");
    nn_compact_kernel<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, (cudaStream_t)stream>>>(n, d_state, d_scan, d_cmap, d_clist);
    MG_CHECK_LAUNCH("nn_compact");
    return MG_OK;
}

/* extract_patches (Multigrid.py:591-677): d_patches [nc][43] doubles, d_fill [nc][31] int32 (-1 padded); CSR of M with
 * sorted columns.  Synchronises; MG_ERR_UNSUPPORTED (text in mg_last_error) outside the supported regime. */
int mg_nn_extract_patches(int64_t nc, const int32_t *d_indptr, const int32_t *d_indices, const double *d_values,
                          const int32_t *d_cmap, const int32_t *d_clist, double *d_patches, int32_t *d_fill,
                          int32_t *d_err, void *stream) {
    MG_REQUIRE(nc > 0 && d_indptr && d_indices && d_values && d_cmap && d_clist && d_patches && d_fill && d_err, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    MG_CHECK_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int32_t), st));
    nn_extract_kernel<<<(unsigned)((nc + 127) / 128), 128, 0, st>>>(nc, d_indptr, d_indices, d_values, d_cmap, d_clist, d_patches, d_fill, d_err);
    MG_CHECK_LAUNCH("nn_extract");
    int32_t e = 0;
    MG_CHECK_CUDA(cudaMemcpyAsync(&e, d_err, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    MG_CHECK_CUDA(cudaStreamSynchronize(st));
    if (e) return set_error(MG_ERR_UNSUPPORTED, "mg_nn_extract_patches", nn_error_text(e));
    return MG_OK;
}

/* Patch variants of coarse nodes with more than 6 neighbours (Multigrid.py:631-663).  mg_nn_count_variants:
 * d_extra[k] = max(0, neighbours of clist[k] - 6).  With d_extra_ptr = its exclusive scan (nc + 1 entries, total = T),
 * mg_nn_extract_variants writes variants 1.. of those nodes to rows nc .. nc+T-1 of d_patches / d_fill (rows 0..nc-1 =
 * variant 0, written by mg_nn_extract_patches) and sets d_not_last ([nc + T], zero-initialised by the caller) for every
 * patch that is not its node's last one.  Synchronises. */
int mg_nn_count_variants(int64_t nc, const int32_t *d_indptr, const int32_t *d_indices, const double *d_values,
                         const int32_t *d_clist, int32_t *d_extra, void *stream) {
    MG_REQUIRE(nc > 0 && d_indptr && d_indices && d_values && d_clist && d_extra, "null argument");
    nn_count_variants_kernel<<<(unsigned)((nc + kBlock - 1) / kBlock), kBlock, 0, (cudaStream_t)stream>>>(nc, d_indptr, d_indices, d_values, d_clist, d_extra);
    MG_CHECK_LAUNCH("nn_count_variants");
    return MG_OK;
}
int mg_nn_extract_variants(int64_t nc, const int32_t *d_indptr, const int32_t *d_indices, const double *d_values,
                           const int32_t *d_cmap, const int32_t *d_clist, const int32_t *d_extra_ptr,
                           double *d_patches, int32_t *d_fill, int32_t *d_not_last, int32_t *d_err, void *stream) {
    MG_REQUIRE(nc > 0 && d_indptr && d_indices && d_values && d_cmap && d_clist && d_extra_ptr && d_patches && d_fill && d_not_last && d_err, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    MG_CHECK_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int32_t), st));
    nn_extract_variants_kernel<<<(unsigned)((nc + 127) / 128), 128, 0, st>>>(nc, d_indptr, d_indices, d_values, d_cmap, d_clist, d_extra_ptr, d_patches, d_fill, d_not_last, d_err);
    MG_CHECK_LAUNCH("nn_extract_variants");
    int32_t e = 0;
    MG_CHECK_CUDA(cudaMemcpyAsync(&e, d_err, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    MG_CHECK_CUDA(cudaStreamSynchronize(st));
    if (e) return set_error(MG_ERR_UNSUPPORTED, "mg_nn_extract_variants", nn_error_text(e));
    return MG_OK;
}

/* fill_B step 1: the (row, col, value) contributions of every prediction in application order ([np][31] each, row =
 * unused_row (>= number of fine rows) where unused) and the d_neighs table ([ncoarse][6], -1 padded).  d_not_last
 * (may be null: no patch variants) marks the patches that must not write d_neighs (see mg_nn_extract_variants). */
int mg_nn_contributions(int64_t np_, const int32_t *d_fill, const double *d_pred, const int32_t *d_cmap,
                        const int32_t *d_not_last, int32_t unused_row, int32_t *d_rows, int32_t *d_cols, double *d_vals,
                        int32_t *d_dneigh, void *stream) {
    MG_REQUIRE(np_ > 0 && d_fill && d_pred && d_cmap && d_rows && d_cols && d_vals && d_dneigh, "null argument");
    nn_contrib_kernel<<<(unsigned)((np_ + 127) / 128), 128, 0, (cudaStream_t)stream>>>(np_, d_fill, d_pred, d_cmap, d_not_last, unused_row, d_rows, d_cols, d_vals, d_dneigh);
    MG_CHECK_LAUNCH("nn_contrib");
    return MG_OK;
}

/* fill_B step 2: with d_order = the contributions sorted stably by (row, col) (two mg_stable_argsort_i32 passes), fold
 * every run with the reference's running mean; d_head[i] = 1 at the first element of a run, d_folded[i] its value */
int mg_nn_fold(int64_t m, int32_t unused_row, const int32_t *d_rows, const int32_t *d_cols, const double *d_vals, const int32_t *d_order,
               int32_t *d_head, double *d_folded, void *stream) {
    MG_REQUIRE(m > 0 && d_rows && d_cols && d_vals && d_order && d_head && d_folded, "null argument");
    nn_fold_kernel<<<(unsigned)((m + kBlock - 1) / kBlock), kBlock, 0, (cudaStream_t)stream>>>(m, unused_row, d_rows, d_cols, d_vals, d_order, d_head, d_folded);
    MG_CHECK_LAUNCH("nn_fold");
    return MG_OK;
}

/* fill_B step 3: d_slot = exclusive scan of d_head; writes the COO triplets of B in (row, col) order */
int mg_nn_emit(int64_t m, const int32_t *d_rows, const int32_t *d_cols, const int32_t *d_order, const int32_t *d_head,
               const int32_t *d_slot, const double *d_folded, int32_t *d_out_rows, int32_t *d_out_cols,
               double *d_out_vals, void *stream) {
    MG_REQUIRE(m > 0 && d_rows && d_cols && d_order && d_head && d_slot && d_folded, "null argument");
    nn_emit_kernel<<<(unsigned)((m + kBlock - 1) / kBlock), kBlock, 0, (cudaStream_t)stream>>>(m, d_rows, d_cols, d_order, d_head, d_slot, d_folded, d_out_rows, d_out_cols, d_out_vals);
    MG_CHECK_LAUNCH("nn_emit");
    return MG_OK;
}

/* Q = B / rowsum(B), in place on the CSR values (Multigrid.py:758-759) */
int mg_nn_row_normalise(int64_t n, const int32_t *d_indptr, double *d_values, void *stream) {
    MG_REQUIRE(n > 0 && d_indptr && d_values, "null argument");
    nn_row_normalise_kernel<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, (cudaStream_t)stream>>>(n, d_indptr, d_values);
    MG_CHECK_LAUNCH("nn_row_normalise");
    return MG_OK;
}

/* pre_process (Multigrid.py:735-739), two passes: count (d_keep [nnz], d_count [n]) then, after a scan, fill */
int mg_nn_cut_count(int64_t n, const int32_t *d_indptr, const int32_t *d_indices, const double *d_values,
                    const int32_t *d_dneigh, int32_t *d_keep, int32_t *d_count, void *stream) {
    MG_REQUIRE(n > 0 && d_indptr && d_indices && d_values && d_dneigh && d_keep && d_count, "null argument");
    nn_cut_count_kernel<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, (cudaStream_t)stream>>>(n, d_indptr, d_indices, d_values, d_dneigh, d_keep, d_count);
    MG_CHECK_LAUNCH("nn_cut_count");
    return MG_OK;
}
int mg_nn_cut_fill(int64_t n, const int32_t *d_indptr, const int32_t *d_indices, const double *d_values,
                   const int32_t *d_keep, const int32_t *d_out_indptr, int32_t *d_out_indices, double *d_out_values,
                   void *stream) {
    MG_REQUIRE(n > 0 && d_indptr && d_indices && d_values && d_keep && d_out_indptr, "null argument");
    nn_cut_fill_kernel<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, (cudaStream_t)stream>>>(n, d_indptr, d_indices, d_values, d_keep, d_out_indptr, d_out_indices, d_out_values);
    MG_CHECK_LAUNCH("nn_cut_fill");
    return MG_OK;
}

#ifdef MGB_TESTING
/* ---- the same per-node code on HOST arrays (serial): used by the CPU test-suite to check the extraction and the
 * contribution logic against the reference's golden vectors without a GPU.  Not called by the product. ---- */
int mg_host_nn_coarsen(int64_t n, const int32_t *h_t_indptr, const int32_t *h_t_indices, const double *h_t_values,
                       int32_t *h_cmap, int32_t *h_clist) {
    MG_REQUIRE(n > 0 && h_t_indptr && h_t_indices && h_t_values && h_cmap && h_clist, "null argument");
    int32_t nc = 0;
    for (int64_t i = 0; i < n; ++i) {
        bool drop = false;
        for (int32_t p = h_t_indptr[i]; p < h_t_indptr[i + 1] && !drop; ++p) {
            const int32_t j = h_t_indices[p];
            if (j < i && h_t_values[p] > 0.0 && h_cmap[j] >= 0) drop = true;
        }
        if (drop) h_cmap[i] = -1;
        else { h_cmap[i] = nc; h_clist[nc++] = (int32_t)i; }
    }
    return nc;
}
/* h_patches / h_fill hold nc + (sum of extra variants) rows; returns that row count (>= nc) or a negative status.
 * Pass null output pointers to only count. */
int mg_host_nn_extract_patches(int64_t nc, const int32_t *h_indptr, const int32_t *h_indices, const double *h_values,
                               const int32_t *h_cmap, const int32_t *h_clist, double *h_patches, int32_t *h_fill) {
    MG_REQUIRE(nc > 0 && h_indptr && h_indices && h_values && h_cmap && h_clist, "null argument");
    int64_t next = nc;
    for (int64_t k = 0; k < nc; ++k) {
        const int additional = nn_extra_variants(h_indptr, h_indices, h_values, h_clist[k]);
        for (int v = 0; v <= additional; ++v) {
            const int64_t j = v == 0 ? k : next++;
            if (!h_patches || !h_fill) continue;
            const int rc = nn_extract_one(h_indptr, h_indices, h_values, h_cmap, h_clist[k], v, h_patches + j * kPatch, h_fill + j * kFill);
            if (rc != NN_OK) return set_error(MG_ERR_UNSUPPORTED, "mg_host_nn_extract_patches", nn_error_text(rc));
        }
    }
    return (int)next;
}
int mg_host_nn_contributions(int64_t np_, const int32_t *h_fill, const double *h_pred, const int32_t *h_cmap,
                             int32_t unused_row, int32_t *h_rows, int32_t *h_cols, double *h_vals, int32_t *h_dneigh) {
    MG_REQUIRE(np_ > 0 && h_fill && h_pred && h_cmap && h_rows && h_cols && h_vals && h_dneigh, "null argument");
    for (int64_t j = 0; j < np_; ++j) {
        int32_t dn[6];
        nn_contributions_one(h_fill + j * kFill, h_pred + j * kFill, h_cmap, unused_row, h_rows + j * kFill, h_cols + j * kFill, h_vals + j * kFill, dn);
        const int32_t c = h_cmap[h_fill[j * kFill]];
        for (int i = 0; i < 6; ++i) h_dneigh[(int64_t)c * 6 + i] = dn[i];
    }
    return MG_OK;
}
#endif  // MGB_TESTING

}  // extern "C"
