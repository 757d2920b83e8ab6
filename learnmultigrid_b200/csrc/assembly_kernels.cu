// assembly_kernels.cu -- P1 finite-element assembly on the device.
//
// Reference: MassMatrix.compute_mass_2d (learn_multigrid/assembly/MassMatrix.py:21-35, 52-59),
// StiffnessMatrix.compute_stiffness_2d (StiffnessMatrix.py:21-36, 53-59; Quadrature2D.compute_grad, Quadrature.py:72-82),
// LoadVector.compute_rhs_2d (LoadVector.py:20-51) and the Dirichlet row replacement of the drivers
// (test/thesis_structured_2d.py:407-414).  The reference loops over the elements in Python (~1 ms per element) and adds
// each 3x3 block into a lil_matrix; here one thread per element writes its nine (row, col, value) contributions, which
// are then sorted stably by (row, col) and summed run by run IN ELEMENT ORDER (the order the reference's `+=` sees
// them), exact zeros dropped as lil_matrix does.  The element arithmetic is the product's vectorised host assembly
// (learnmultigrid_b200/assembly/*.py) operation for operation without FMA contraction, so device and host results are
// bit-identical; both differ from the reference in the last bit of detJ (np.linalg.det goes through a pivoted LU).
#include "common.cuh"

namespace mgb {

__device__ __forceinline__ void element_jacobian(const double *__restrict__ p, const int32_t *__restrict__ conn,
                                                 int64_t e, int32_t node[3], double &J00, double &J01, double &J10,
                                                 double &J11, double &det) {
    node[0] = conn[3 * e];
    node[1] = conn[3 * e + 1];
    node[2] = conn[3 * e + 2];
    const double x0 = p[2 * (int64_t)node[0]], y0 = p[2 * (int64_t)node[0] + 1];
    const double x1 = p[2 * (int64_t)node[1]], y1 = p[2 * (int64_t)node[1] + 1];
    const double x2 = p[2 * (int64_t)node[2]], y2 = p[2 * (int64_t)node[2] + 1];
    J00 = __dsub_rn(x1, x0);
    J01 = __dsub_rn(x2, x0);
    J10 = __dsub_rn(y1, y0);
    J11 = __dsub_rn(y2, y0);
    det = __dsub_rn(__dmul_rn(J00, J11), __dmul_rn(J01, J10));      // signed, like the reference (MassMatrix.py:31)
}

struct MassConst { double c[9]; };
struct StiffConst { double g[6]; double w[3]; int npts; };
struct LoadConst { double c[3]; };

// loc_M[i,j] = detJ * c[i][j],   c[i][j] = sum_k phi_i(p_k) phi_j(p_k) w_k (evaluated once on the host)
__global__ void __launch_bounds__(kBlock)
assemble_mass_kernel(int64_t ne, const double *__restrict__ p, const int32_t *__restrict__ conn, MassConst K,
                     int32_t *__restrict__ rows, int32_t *__restrict__ cols, double *__restrict__ vals) {
    const int64_t e = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (e >= ne) return;
    int32_t nd[3];
    double J00, J01, J10, J11, det;
    element_jacobian(p, conn, e, nd, J00, J01, J10, J11, det);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int64_t o = 9 * e + 3 * i + j;
            rows[o] = nd[i];
            cols[o] = nd[j];
            vals[o] = __dmul_rn(det, K.c[3 * i + j]);
        }
}

// loc_A[i,j] = detJ * sum_k ((J^-T g_i)^T J^-T) g_j * w[i]  (the reference indexes w by i, Quadrature.py:80), times an
// optional per-element coefficient
__global__ void __launch_bounds__(kBlock)
assemble_stiffness_kernel(int64_t ne, const double *__restrict__ p, const int32_t *__restrict__ conn, StiffConst K,
                          const double *__restrict__ coef, int32_t *__restrict__ rows, int32_t *__restrict__ cols,
                          double *__restrict__ vals) {
    const int64_t e = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (e >= ne) return;
    int32_t nd[3];
    double J00, J01, J10, J11, det;
    element_jacobian(p, conn, e, nd, J00, J01, J10, J11, det);
    const double T00 = __ddiv_rn(J11, det), T01 = __ddiv_rn(-J10, det);
    const double T10 = __ddiv_rn(-J01, det), T11 = __ddiv_rn(J00, det);
    const double ke = coef ? coef[e] : 1.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double a0 = __dadd_rn(__dmul_rn(T00, K.g[2 * i]), __dmul_rn(T01, K.g[2 * i + 1]));
        const double a1 = __dadd_rn(__dmul_rn(T10, K.g[2 * i]), __dmul_rn(T11, K.g[2 * i + 1]));
        const double t0 = __dadd_rn(__dmul_rn(a0, T00), __dmul_rn(a1, T10));
        const double t1 = __dadd_rn(__dmul_rn(a0, T01), __dmul_rn(a1, T11));
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double s = __dmul_rn(__dadd_rn(__dmul_rn(t0, K.g[2 * j]), __dmul_rn(t1, K.g[2 * j + 1])), K.w[i]);
            double res = 0.0;
            for (int k = 0; k < K.npts; ++k) res = __dadd_rn(res, s);
            double v = __dmul_rn(det, res);
            if (coef) v = __dmul_rn(v, ke);
            const int64_t o = 9 * e + 3 * i + j;
            rows[o] = nd[i];
            cols[o] = nd[j];
            vals[o] = v;
        }
    }
}

// loc_rhs[i] = detJ * c[i]
__global__ void __launch_bounds__(kBlock)
assemble_load_kernel(int64_t ne, const double *__restrict__ p, const int32_t *__restrict__ conn, LoadConst K,
                     int32_t *__restrict__ nodes, double *__restrict__ vals) {
    const int64_t e = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (e >= ne) return;
    int32_t nd[3];
    double J00, J01, J10, J11, det;
    element_jacobian(p, conn, e, nd, J00, J01, J10, J11, det);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        nodes[3 * e + i] = nd[i];
        vals[3 * e + i] = __dmul_rn(det, K.c[i]);
    }
}

// contributions sorted stably by (row, col): the first element of a run adds the run up in order
__global__ void __launch_bounds__(kBlock)
coo_fold_sum_kernel(int64_t m, const int32_t *__restrict__ rows, const int32_t *__restrict__ cols,
                    const double *__restrict__ vals, const int32_t *__restrict__ order, int32_t *__restrict__ head,
                    double *__restrict__ folded) {
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= m) return;
    const int32_t o = order[i];
    const int32_t r = rows[o], c = cols ? cols[o] : 0;
    head[i] = 0;
    if (i > 0) {
        const int32_t po = order[i - 1];
        if (rows[po] == r && (cols ? cols[po] : 0) == c) return;
    }
    double s = 0.0;
    for (int64_t k = i; k < m; ++k) {
        const int32_t ok = order[k];
        if (rows[ok] != r || (cols ? cols[ok] : 0) != c) break;
        s = __dadd_rn(s, vals[ok]);
    }
    head[i] = s != 0.0 ? 1 : 0;
    folded[i] = s;
}

__global__ void __launch_bounds__(kBlock)
vector_from_runs_kernel(int64_t m, const int32_t *__restrict__ rows, const int32_t *__restrict__ order,
                        const int32_t *__restrict__ head, const double *__restrict__ folded, double *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= m || !head[i]) return;
    out[rows[order[i]]] = folded[i];
}

// A[nodes,:] = I[nodes,:]
__global__ void __launch_bounds__(kBlock)
dirichlet_count_kernel(int64_t n, const int32_t *__restrict__ indptr, const int32_t *__restrict__ flag,
                       int32_t *__restrict__ count) {
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i < n) count[i] = flag[i] ? 1 : indptr[i + 1] - indptr[i];
}
__global__ void __launch_bounds__(kBlock)
dirichlet_fill_kernel(int64_t n, const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                      const double *__restrict__ values, const int32_t *__restrict__ flag,
                      const int32_t *__restrict__ out_indptr, int32_t *__restrict__ out_indices,
                      double *__restrict__ out_values) {
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    int32_t o = out_indptr[i];
    if (flag[i]) {
        out_indices[o] = (int32_t)i;
        out_values[o] = 1.0;
        return;
    }
    for (int32_t p = indptr[i]; p < indptr[i + 1]; ++p, ++o) {
        out_indices[o] = indices[p];
        out_values[o] = values[p];
    }
}

// ---- semi-geometric coupling operator B[f,c] = int phi_f phi_c on triangle-triangle intersections -------------------
// (no reference counterpart: L2Projection.compute_transfer_2d is a stub, L2Projection.py:17-24; the recipe is the 1D
// one, CouplingOperator.py:31-69.  Host model: learnmultigrid_b200/L2_projection/coupling2d.py.)
// One candidate pair: clip the fine triangle by the three half-planes of the (counter-clockwise) coarse triangle
// (Sutherland-Hodgman), fan-triangulate the polygon, integrate the products of the two P1 bases with the edge-midpoint
// rule (exact for quadratics).  loc[i][j] (9 doubles) and the overlap area are returned; all zero if the triangles do
// not overlap.
__host__ __device__ inline void bary2(const double *t, double x, double y, double *l) {
    const double ax = t[0], ay = t[1], bx = t[2], by = t[3], cx = t[4], cy = t[5];
    const double det = (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);
    const double l1 = ((x - ax) * (cy - ay) - (y - ay) * (cx - ax)) / det;
    const double l2 = ((bx - ax) * (y - ay) - (by - ay) * (x - ax)) / det;
    l[0] = 1.0 - l1 - l2;
    l[1] = l1;
    l[2] = l2;
}

__host__ __device__ inline void coupling_pair(const double *tf, const double *tc_in, double *loc, double *area_out) {
    for (int i = 0; i < 9; ++i) loc[i] = 0.0;
    *area_out = 0.0;
    double tc[6];
    for (int i = 0; i < 6; ++i) tc[i] = tc_in[i];
    const double orient = (tc[2] - tc[0]) * (tc[5] - tc[1]) - (tc[3] - tc[1]) * (tc[4] - tc[0]);
    if (orient < 0) {            // clip against a counter-clockwise triangle: swap vertices 1 and 2
        double t;
        t = tc[2]; tc[2] = tc[4]; tc[4] = t;
        t = tc[3]; tc[3] = tc[5]; tc[5] = t;
    }
    double px[8], py[8], qx[8], qy[8];
    int n = 3;
    for (int i = 0; i < 3; ++i) { px[i] = tf[2 * i]; py[i] = tf[2 * i + 1]; }
    for (int e = 0; e < 3 && n > 0; ++e) {
        const double ax = tc[2 * e], ay = tc[2 * e + 1];
        const double ex = tc[2 * ((e + 1) % 3)] - ax, ey = tc[2 * ((e + 1) % 3) + 1] - ay;
        int m = 0;
        for (int k = 0; k < n; ++k) {
            const int kn = (k + 1 < n) ? k + 1 : 0;
            const double d = ex * (py[k] - ay) - ey * (px[k] - ax);
            const double dn = ex * (py[kn] - ay) - ey * (px[kn] - ax);
            const bool in = d >= 0, inn = dn >= 0;
            if (in && m < 8) { qx[m] = px[k]; qy[m] = py[k]; ++m; }
            if (in != inn && m < 8) {
                const double t = d / (d - dn);
                qx[m] = px[k] + t * (px[kn] - px[k]);
                qy[m] = py[k] + t * (py[kn] - py[k]);
                ++m;
            }
        }
        n = m;
        for (int k = 0; k < n; ++k) { px[k] = qx[k]; py[k] = qy[k]; }
    }
    double total = 0.0;
    for (int i = 1; i + 1 < n; ++i) {
        const double x0 = px[0], y0 = py[0], x1 = px[i], y1 = py[i], x2 = px[i + 1], y2 = py[i + 1];
        double area = 0.5 * ((x1 - x0) * (y2 - y0) - (y1 - y0) * (x2 - x0));
        if (area < 0) area = -area;
        if (!(area > 0)) continue;
        total += area;
        const double mx[3] = {0.5 * (x0 + x1), 0.5 * (x1 + x2), 0.5 * (x2 + x0)};
        const double my[3] = {0.5 * (y0 + y1), 0.5 * (y1 + y2), 0.5 * (y2 + y0)};
        double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int q = 0; q < 3; ++q) {
            double lf[3], lc[3];
            bary2(tf, mx[q], my[q], lf);
            bary2(tc_in, mx[q], my[q], lc);
            for (int a = 0; a < 3; ++a)
                for (int b2 = 0; b2 < 3; ++b2) acc[3 * a + b2] += lf[a] * lc[b2];
        }
        const double w = area / 3.0;
        for (int a = 0; a < 9; ++a) loc[a] += w * acc[a];
    }
    *area_out = total;
}

__global__ void __launch_bounds__(kBlock)
coupling_pairs_kernel(int64_t npairs, const int32_t *__restrict__ pair_f, const int32_t *__restrict__ pair_c,
                      const double *__restrict__ pf, const int32_t *__restrict__ tf, const double *__restrict__ pc,
                      const int32_t *__restrict__ tc, int32_t *__restrict__ rows, int32_t *__restrict__ cols,
                      double *__restrict__ vals, double *__restrict__ area) {
    const int64_t k = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (k >= npairs) return;
    const int64_t f = pair_f[k], c = pair_c[k];
    double a[6], b[6], loc[9], ar;
    int32_t nf[3], nc[3];
    for (int i = 0; i < 3; ++i) {
        nf[i] = tf[3 * f + i];
        nc[i] = tc[3 * c + i];
        a[2 * i] = pf[2 * (int64_t)nf[i]];
        a[2 * i + 1] = pf[2 * (int64_t)nf[i] + 1];
        b[2 * i] = pc[2 * (int64_t)nc[i]];
        b[2 * i + 1] = pc[2 * (int64_t)nc[i] + 1];
    }
    coupling_pair(a, b, loc, &ar);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            rows[9 * k + 3 * i + j] = nf[i];
            cols[9 * k + 3 * i + j] = nc[j];
            vals[9 * k + 3 * i + j] = loc[3 * i + j];
        }
    if (area) area[k] = ar;
}

static inline unsigned grid_for(int64_t n) { return (unsigned)((n + kBlock - 1) / kBlock); }

// ---- candidate pairs of the coupling operator: uniform-grid binning of the triangles' bounding boxes ---------------
// A (fine, coarse) pair is a candidate if the two bounding boxes overlap.  Both meshes are binned on one G x G grid;
// a fine triangle walks the cells its box covers and the coarse triangles listed there, and reports a pair only in the
// cell that holds the lower-left corner of the boxes' intersection, so every pair is reported exactly once.
struct BinGrid {
    double lo[2], inv[2];       // cell of a point p: floor((p - lo) * inv), clamped to [0, G)
    int32_t G;
};
__device__ __forceinline__ int bin_cell(const BinGrid &g, double v, int d) {
    int c = (int)floor((v - g.lo[d]) * g.inv[d]);
    return c < 0 ? 0 : (c >= g.G ? g.G - 1 : c);
}
// box[t] = (xlo, ylo, xhi, yhi); ncells[t] = cells the box covers
__global__ void __launch_bounds__(kBlock)
tri_boxes_kernel(int64_t ne, const double *__restrict__ pts, const int32_t *__restrict__ conn, BinGrid g,
                 double *__restrict__ box, int32_t *__restrict__ ncells) {
    const int64_t t = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (t >= ne) return;
    double lo[2] = {1e300, 1e300}, hi[2] = {-1e300, -1e300};
    for (int i = 0; i < 3; ++i) {
        const int64_t v = conn[3 * t + i];
        for (int d = 0; d < 2; ++d) {
            const double x = pts[2 * v + d];
            lo[d] = fmin(lo[d], x);
            hi[d] = fmax(hi[d], x);
        }
    }
    box[4 * t] = lo[0]; box[4 * t + 1] = lo[1]; box[4 * t + 2] = hi[0]; box[4 * t + 3] = hi[1];
    if (ncells)
        ncells[t] = (bin_cell(g, hi[0], 0) - bin_cell(g, lo[0], 0) + 1) * (bin_cell(g, hi[1], 1) - bin_cell(g, lo[1], 1) + 1);
}
// (cell, triangle) incidences of the coarse mesh, triangle-major: inc_cell[ptr[t] + k], inc_tri likewise
__global__ void __launch_bounds__(kBlock)
tri_incidence_kernel(int64_t ne, const double *__restrict__ box, BinGrid g, const int32_t *__restrict__ ptr,
                     int32_t *__restrict__ inc_cell, int32_t *__restrict__ inc_tri) {
    const int64_t t = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (t >= ne) return;
    const int x0 = bin_cell(g, box[4 * t], 0), x1 = bin_cell(g, box[4 * t + 2], 0);
    const int y0 = bin_cell(g, box[4 * t + 1], 1), y1 = bin_cell(g, box[4 * t + 3], 1);
    int64_t o = ptr[t];
    for (int y = y0; y <= y1; ++y)
        for (int x = x0; x <= x1; ++x, ++o) {
            inc_cell[o] = y * g.G + x;
            inc_tri[o] = (int32_t)t;
        }
}
// pass 0 (pairs == NULL): count[f] = candidate pairs of fine triangle f; pass 1: write them at ptr[f], coarse ids
// ascending (the lists of a cell are ascending, the few cells of a box are merged by insertion)
__global__ void __launch_bounds__(kBlock)
tri_pairs_kernel(int64_t nf, const double *__restrict__ box_f, const double *__restrict__ box_c, BinGrid g,
                 const int32_t *__restrict__ cell_ptr, const int32_t *__restrict__ cell_tri,
                 int32_t *__restrict__ count, const int32_t *__restrict__ ptr, int32_t *__restrict__ pair_f,
                 int32_t *__restrict__ pair_c) {
    const int64_t f = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (f >= nf) return;
    const double fx0 = box_f[4 * f], fy0 = box_f[4 * f + 1], fx1 = box_f[4 * f + 2], fy1 = box_f[4 * f + 3];
    const int x0 = bin_cell(g, fx0, 0), x1 = bin_cell(g, fx1, 0), y0 = bin_cell(g, fy0, 1), y1 = bin_cell(g, fy1, 1);
    int n = 0;
    const int64_t base = pair_c ? ptr[f] : 0;
    for (int y = y0; y <= y1; ++y)
        for (int x = x0; x <= x1; ++x) {
            const int cell = y * g.G + x;
            for (int32_t q = cell_ptr[cell]; q < cell_ptr[cell + 1]; ++q) {
                const int32_t c = cell_tri[q];
                const double cx0 = box_c[4 * c], cy0 = box_c[4 * c + 1], cx1 = box_c[4 * c + 2], cy1 = box_c[4 * c + 3];
                if (!(fx0 <= cx1 && cx0 <= fx1 && fy0 <= cy1 && cy0 <= fy1)) continue;
                // the cell of the intersection's lower-left corner owns the pair
                if (bin_cell(g, fmax(fx0, cx0), 0) != x || bin_cell(g, fmax(fy0, cy0), 1) != y) continue;
                if (pair_c) {
                    int64_t k = base + n;                      // insertion keeps the coarse ids ascending
                    while (k > base && pair_c[k - 1] > c) { pair_c[k] = pair_c[k - 1]; --k; }
                    pair_c[k] = c;
                    pair_f[base + n] = (int32_t)f;
                }
                ++n;
            }
        }
    if (!pair_c) count[f] = n;
}

}  // namespace mgb

using namespace mgb;

extern "C" {

/* the nine (row, col, value) contributions of every P1 element, element-major ([ne][3][3]).
 * kind 0: mass, h_const = c[3][3] with c[i][j] = sum_k phi_i(p_k) phi_j(p_k) w_k (MassMatrix.py:52-59);
 * kind 1: stiffness, h_const = g[3][2] (reference gradients) followed by w[3] and the number of quadrature points
 *         (StiffnessMatrix.py:53-59, Quadrature.py:72-82); d_coef: optional per-element coefficient (NULL: 1).
 * d_points: [np][2] doubles, d_conn: [ne][3] int32. */
int mg_assemble_p1_2d(int64_t ne, const double *d_points, const int32_t *d_conn, int kind, const double *h_const,
                      const double *d_coef, int32_t *d_rows, int32_t *d_cols, double *d_vals, void *stream) {
    MG_REQUIRE(ne > 0 && d_points && d_conn && h_const && d_rows && d_cols && d_vals, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (kind == 0) {
        MassConst K;
        for (int i = 0; i < 9; ++i) K.c[i] = h_const[i];
        assemble_mass_kernel<<<grid_for(ne), kBlock, 0, st>>>(ne, d_points, d_conn, K, d_rows, d_cols, d_vals);
    } else if (kind == 1) {
        StiffConst K;
        for (int i = 0; i < 6; ++i) K.g[i] = h_const[i];
        for (int i = 0; i < 3; ++i) K.w[i] = h_const[6 + i];
        K.npts = (int)h_const[9];
        assemble_stiffness_kernel<<<grid_for(ne), kBlock, 0, st>>>(ne, d_points, d_conn, K, d_coef, d_rows, d_cols, d_vals);
    } else {
        return set_error(MG_ERR_INVALID, "mg_assemble_p1_2d", "kind must be 0 (mass) or 1 (stiffness)");
    }
    MG_CHECK_LAUNCH("assemble_p1_2d");
    return MG_OK;
}

/* load vector contributions ([ne][3] nodes / values), h_c3[i] = sum_k phi_i(p_k) f(p_k) w_k (LoadVector.py:45-51) */
int mg_assemble_load_p1_2d(int64_t ne, const double *d_points, const int32_t *d_conn, const double *h_c3,
                           int32_t *d_nodes, double *d_vals, void *stream) {
    MG_REQUIRE(ne > 0 && d_points && d_conn && h_c3 && d_nodes && d_vals, "null argument");
    LoadConst K;
    for (int i = 0; i < 3; ++i) K.c[i] = h_c3[i];
    assemble_load_kernel<<<grid_for(ne), kBlock, 0, (cudaStream_t)stream>>>(ne, d_points, d_conn, K, d_nodes, d_vals);
    MG_CHECK_LAUNCH("assemble_load");
    return MG_OK;
}

/* with d_order = the stable (row, col) sort order of the contributions: d_head[i] = 1 where a run starts and its sum
 * (added in order) is not exactly zero, d_folded[i] = that sum.  d_cols may be NULL (vectors).  mg_nn_emit writes
 * the triplets. */
int mg_coo_fold_sum(int64_t m, const int32_t *d_rows, const int32_t *d_cols, const double *d_vals,
                    const int32_t *d_order, int32_t *d_head, double *d_folded, void *stream) {
    MG_REQUIRE(m > 0 && d_rows && d_vals && d_order && d_head && d_folded, "null argument");
    coo_fold_sum_kernel<<<grid_for(m), kBlock, 0, (cudaStream_t)stream>>>(m, d_rows, d_cols, d_vals, d_order, d_head, d_folded);
    MG_CHECK_LAUNCH("coo_fold_sum");
    return MG_OK;
}
/* d_out[row of run] = sum of the run, for runs marked by mg_coo_fold_sum (d_out zero-initialised by the caller) */
int mg_vector_from_runs(int64_t m, const int32_t *d_rows, const int32_t *d_order, const int32_t *d_head,
                        const double *d_folded, double *d_out, void *stream) {
    MG_REQUIRE(m > 0 && d_rows && d_order && d_head && d_folded && d_out, "null argument");
    vector_from_runs_kernel<<<grid_for(m), kBlock, 0, (cudaStream_t)stream>>>(m, d_rows, d_order, d_head, d_folded, d_out);
    MG_CHECK_LAUNCH("vector_from_runs");
    return MG_OK;
}

/* Semi-geometric coupling operator between two P1 triangle meshes (coupling2d.py is the host model): for every
 * candidate pair (fine element d_pair_f[k], coarse element d_pair_c[k]) the nine contributions
 * (fine node i, coarse node j, int_{overlap} phi_i phi_j) and, optionally, the overlap area.  Sum the runs with
 * mg_coo_fold_sum / mg_nn_emit. */
int mg_coupling_pairs_p1_2d(int64_t npairs, const int32_t *d_pair_f, const int32_t *d_pair_c, const double *d_pf,
                            const int32_t *d_tf, const double *d_pc, const int32_t *d_tc, int32_t *d_rows,
                            int32_t *d_cols, double *d_vals, double *d_area, void *stream) {
    MG_REQUIRE(npairs > 0 && d_pair_f && d_pair_c && d_pf && d_tf && d_pc && d_tc && d_rows && d_cols && d_vals, "null argument");
    coupling_pairs_kernel<<<grid_for(npairs), kBlock, 0, (cudaStream_t)stream>>>(npairs, d_pair_f, d_pair_c, d_pf, d_tf, d_pc, d_tc, d_rows, d_cols, d_vals, d_area);
    MG_CHECK_LAUNCH("coupling_pairs");
    return MG_OK;
}
/* Candidate pairs on the device (see the kernels above).  Boxes: d_box[4*ne]; with d_ncells the number of grid cells each
 * box covers.  lo / inv_size / G describe the common G x G binning grid. */
int mg_tri_boxes_2d(int64_t ne, const double *d_points, const int32_t *d_conn, const double *h_lo, const double *h_inv_size,
                    int32_t G, double *d_box, int32_t *d_ncells, void *stream) {
    MG_REQUIRE(ne > 0 && d_points && d_conn && h_lo && h_inv_size && G > 0 && d_box, "bad argument");
    BinGrid g{{h_lo[0], h_lo[1]}, {h_inv_size[0], h_inv_size[1]}, G};
    tri_boxes_kernel<<<grid_for(ne), kBlock, 0, (cudaStream_t)stream>>>(ne, d_points, d_conn, g, d_box, d_ncells);
    MG_CHECK_LAUNCH("tri_boxes");
    return MG_OK;
}
int mg_tri_incidence_2d(int64_t ne, const double *d_box, const double *h_lo, const double *h_inv_size, int32_t G,
                        const int32_t *d_ptr, int32_t *d_inc_cell, int32_t *d_inc_tri, void *stream) {
    MG_REQUIRE(ne > 0 && d_box && h_lo && h_inv_size && G > 0 && d_ptr && d_inc_cell && d_inc_tri, "bad argument");
    BinGrid g{{h_lo[0], h_lo[1]}, {h_inv_size[0], h_inv_size[1]}, G};
    tri_incidence_kernel<<<grid_for(ne), kBlock, 0, (cudaStream_t)stream>>>(ne, d_box, g, d_ptr, d_inc_cell, d_inc_tri);
    MG_CHECK_LAUNCH("tri_incidence");
    return MG_OK;
}
/* d_pair_c == NULL: d_count[f] = candidate pairs of fine triangle f.  Otherwise the pairs are written at d_ptr[f]
 * (exclusive scan of the counts): fine id, coarse ids ascending. */
int mg_tri_pairs_2d(int64_t nf, const double *d_box_f, const double *d_box_c, const double *h_lo, const double *h_inv_size,
                    int32_t G, const int32_t *d_cell_ptr, const int32_t *d_cell_tri, int32_t *d_count,
                    const int32_t *d_ptr, int32_t *d_pair_f, int32_t *d_pair_c, void *stream) {
    MG_REQUIRE(nf > 0 && d_box_f && d_box_c && h_lo && h_inv_size && G > 0 && d_cell_ptr && d_cell_tri, "bad argument");
    MG_REQUIRE((d_pair_c && d_pair_f && d_ptr) || (!d_pair_c && d_count), "count or fill arguments missing");
    BinGrid g{{h_lo[0], h_lo[1]}, {h_inv_size[0], h_inv_size[1]}, G};
    tri_pairs_kernel<<<grid_for(nf), kBlock, 0, (cudaStream_t)stream>>>(nf, d_box_f, d_box_c, g, d_cell_ptr, d_cell_tri,
                                                                          d_count, d_ptr, d_pair_f, d_pair_c);
    MG_CHECK_LAUNCH("tri_pairs");
    return MG_OK;
}
#ifdef MGB_TESTING
/* the same per-pair code on HOST arrays (serial), for the CPU test-suite */
int mg_host_coupling_pairs_p1_2d(int64_t npairs, const int32_t *h_pair_f, const int32_t *h_pair_c, const double *h_pf,
                                 const int32_t *h_tf, const double *h_pc, const int32_t *h_tc, int32_t *h_rows,
                                 int32_t *h_cols, double *h_vals, double *h_area) {
    MG_REQUIRE(npairs > 0 && h_pair_f && h_pair_c && h_pf && h_tf && h_pc && h_tc && h_rows && h_cols && h_vals, "null argument");
    for (int64_t k = 0; k < npairs; ++k) {
        const int64_t f = h_pair_f[k], c = h_pair_c[k];
        double a[6], b[6], loc[9], ar;
        for (int i = 0; i < 3; ++i) {
            a[2 * i] = h_pf[2 * (int64_t)h_tf[3 * f + i]];
            a[2 * i + 1] = h_pf[2 * (int64_t)h_tf[3 * f + i] + 1];
            b[2 * i] = h_pc[2 * (int64_t)h_tc[3 * c + i]];
            b[2 * i + 1] = h_pc[2 * (int64_t)h_tc[3 * c + i] + 1];
        }
        coupling_pair(a, b, loc, &ar);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                h_rows[9 * k + 3 * i + j] = h_tf[3 * f + i];
                h_cols[9 * k + 3 * i + j] = h_tc[3 * c + j];
                h_vals[9 * k + 3 * i + j] = loc[3 * i + j];
            }
        if (h_area) h_area[k] = ar;
    }
    return MG_OK;
}
#endif  // MGB_TESTING

/* Dirichlet rows (thesis_structured_2d.py:407-414: A[nodes,:] = I[nodes,:]): rows with d_flag != 0 become (i, 1.0).
 * Two passes around a scan of d_count. */
int mg_csr_dirichlet_count(int64_t n, const int32_t *d_indptr, const int32_t *d_flag, int32_t *d_count, void *stream) {
    MG_REQUIRE(n > 0 && d_indptr && d_flag && d_count, "null argument");
    dirichlet_count_kernel<<<grid_for(n), kBlock, 0, (cudaStream_t)stream>>>(n, d_indptr, d_flag, d_count);
    MG_CHECK_LAUNCH("dirichlet_count");
    return MG_OK;
}
int mg_csr_dirichlet_fill(int64_t n, const int32_t *d_indptr, const int32_t *d_indices, const double *d_values,
                          const int32_t *d_flag, const int32_t *d_out_indptr, int32_t *d_out_indices,
                          double *d_out_values, void *stream) {
    MG_REQUIRE(n > 0 && d_indptr && d_indices && d_values && d_flag && d_out_indptr && d_out_indices && d_out_values, "null argument");
    dirichlet_fill_kernel<<<grid_for(n), kBlock, 0, (cudaStream_t)stream>>>(n, d_indptr, d_indices, d_values, d_flag, d_out_indptr, d_out_indices, d_out_values);
    MG_CHECK_LAUNCH("dirichlet_fill");
    return MG_OK;
}

}  // extern "C"
