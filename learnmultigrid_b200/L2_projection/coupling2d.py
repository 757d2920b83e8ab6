"""Semi-geometric coupling operator between two ARBITRARY (nested or not) P1 triangle meshes of the same domain:
    B[f, c] = integral of phi_f^fine * phi_c^coarse
by intersecting every fine triangle with the coarse triangles it overlaps and integrating on the intersection polygons.

The reference only sketches this in 2D (L2Projection.compute_transfer_2d is an unfinished stub,
learn_multigrid/L2_projection/L2Projection.py:17-24; Intersection.find_intersections2d, Intersection.py:19-34, is the
nested child map); this module finishes it along the 1D recipe (CouplingOperator.compute_b_1d,
CouplingOperator.py:31-69: quadrature on every intersection segment of the two shape functions).  There is therefore no
reference output to pin it to -- parity is UNPINNED; the tests check it through exact properties instead:
on nested meshes B = M_h P to rounding, rowsum(B) = rowsum(M_fine) and colsum(B) = rowsum(M_coarse)
(partition of unity of either basis), for any pair of meshes.

Everything is vectorised NumPy over the candidate pairs (host-side setup):
  1. candidate pairs by binning the triangles' bounding boxes into a uniform grid;
  2. Sutherland-Hodgman clipping of the fine triangle by the three half-planes of the coarse triangle
     (polygons of at most 7 vertices, processed slot by slot for all pairs at once);
  3. fan triangulation of the polygon; the product of two P1 functions is quadratic, so the 3-point rule at the edge
     midpoints integrates it exactly on every sub-triangle.
"""
import numpy as np
import scipy.sparse as sp

MAXV = 8


def _signed_area2(p, tri):
    a, b, c = p[tri[:, 0]], p[tri[:, 1]], p[tri[:, 2]]
    return (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0])


def candidate_pairs(pf, tf, pc, tc):
    """(fine element, coarse element) pairs whose bounding boxes overlap"""
    lo_c, hi_c = pc[tc].min(axis=1), pc[tc].max(axis=1)
    lo_f, hi_f = pf[tf].min(axis=1), pf[tf].max(axis=1)
    lo = np.minimum(lo_c.min(axis=0), lo_f.min(axis=0))
    hi = np.maximum(hi_c.max(axis=0), hi_f.max(axis=0))
    G = max(1, int(np.sqrt(len(tc) / 2.0)))
    size = (hi - lo) / G
    size[size == 0] = 1.0

    def cells(l, h):
        i0 = np.clip(np.floor((l - lo) / size).astype(np.int64), 0, G - 1)
        i1 = np.clip(np.floor((h - lo) / size).astype(np.int64), 0, G - 1)
        return i0, i1

    def incidence(l, h, n):
        i0, i1 = cells(l, h)
        span = (i1 - i0).max(axis=0) + 1
        rows, cols = [], []
        for dx in range(int(span[0])):
            for dy in range(int(span[1])):
                ok = (i0[:, 0] + dx <= i1[:, 0]) & (i0[:, 1] + dy <= i1[:, 1])
                e = np.flatnonzero(ok)
                rows.append(e)
                cols.append((i0[e, 1] + dy) * G + (i0[e, 0] + dx))
        r, c = np.concatenate(rows), np.concatenate(cols)
        return sp.csr_matrix((np.ones(len(r), dtype=np.int8), (r, c)), shape=(n, G * G))

    pairs = sp.coo_matrix(incidence(lo_f, hi_f, len(tf)) @ incidence(lo_c, hi_c, len(tc)).T)
    f, c = pairs.row, pairs.col
    keep = np.all(lo_f[f] <= hi_c[c], axis=1) & np.all(lo_c[c] <= hi_f[f], axis=1)
    return f[keep].astype(np.int64), c[keep].astype(np.int64)


def clip_polygons(poly, count, a, b):
    """Sutherland-Hodgman step for all polygons at once: keep the part on the left of the directed line a -> b.
    poly (K, MAXV, 2), count (K,), a, b (K, 2)."""
    K = len(poly)
    ex, ey = (b - a)[:, 0], (b - a)[:, 1]
    d = ex[:, None] * (poly[:, :, 1] - a[:, None, 1]) - ey[:, None] * (poly[:, :, 0] - a[:, None, 0])   # > 0: left
    idx = np.arange(MAXV)[None, :]
    valid = idx < count[:, None]
    nxt = np.where(idx + 1 < count[:, None], idx + 1, 0)
    rows = np.arange(K)[:, None]
    d_n = d[rows, nxt]
    p_n = poly[rows, nxt]
    inside = d >= 0
    inside_n = d_n >= 0
    emit_v = valid & inside                                   # the vertex itself
    emit_x = valid & (inside != inside_n)                     # the crossing of the edge to the next vertex
    denom = d - d_n
    t = np.where(emit_x, d / np.where(denom == 0, 1.0, denom), 0.0)
    cross = poly + t[:, :, None] * (p_n - poly)
    flags = np.stack([emit_v, emit_x], axis=2).reshape(K, 2 * MAXV)
    pts = np.stack([poly, cross], axis=2).reshape(K, 2 * MAXV, 2)
    pos = np.cumsum(flags, axis=1) - 1
    new_count = flags.sum(axis=1)
    if new_count.max(initial=0) > MAXV:
        raise RuntimeError("clipped polygon with more than %d vertices" % MAXV)
    out = np.zeros_like(poly)
    kk, ss = np.nonzero(flags)
    out[kk, pos[kk, ss]] = pts[kk, ss]
    return out, new_count


def _barycentric(p, tri, elem, x):
    """barycentric coordinates (K, 3) of the points x (K, 2) in the triangles tri[elem]"""
    a, b, c = p[tri[elem, 0]], p[tri[elem, 1]], p[tri[elem, 2]]
    det = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0])
    l1 = ((x[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (x[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0])) / det
    l2 = ((b[:, 0] - a[:, 0]) * (x[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (x[:, 0] - a[:, 0])) / det
    return np.stack([1.0 - l1 - l2, l1, l2], axis=1)


def candidate_pairs_device(S, d_pf, d_tf, d_pc, d_tc, lo, hi):
    """candidate_pairs on the GPU (csrc/assembly_kernels.cu: mg_tri_boxes_2d / mg_tri_incidence_2d / mg_tri_pairs_2d):
    both meshes binned on one G x G grid, per-cell lists of the coarse triangles by a stable radix sort of the (cell,
    triangle) incidences, two passes (count, fill) over the fine triangles.  Returns device int32 arrays (f, c), sorted
    by fine id, then coarse id -- the same SET of pairs as the host function."""
    import ctypes
    from .. import _lib
    torch, lib, dev = S.torch, S.lib, S.dev
    st = S.st()
    nf, nc = d_tf.shape[0], d_tc.shape[0]
    G = max(1, int(np.sqrt(nc / 2.0)))
    size = (np.asarray(hi, dtype=np.float64) - np.asarray(lo, dtype=np.float64)) / G
    size[size == 0] = 1.0
    h_lo = np.ascontiguousarray(lo, dtype=np.float64)
    h_inv = np.ascontiguousarray(1.0 / size, dtype=np.float64)
    plo, pinv = h_lo.ctypes.data_as(ctypes.c_void_p), h_inv.ctypes.data_as(ctypes.c_void_p)
    box_f, box_c = S.empty(4 * nf, torch.float64), S.empty(4 * nc, torch.float64)
    ncells = S.empty(nc, torch.int32)
    _lib.check(lib.mg_tri_boxes_2d(nf, d_pf.data_ptr(), d_tf.data_ptr(), plo, pinv, G, box_f.data_ptr(), None, st),
               "mg_tri_boxes_2d")
    _lib.check(lib.mg_tri_boxes_2d(nc, d_pc.data_ptr(), d_tc.data_ptr(), plo, pinv, G, box_c.data_ptr(),
                                   ncells.data_ptr(), st), "mg_tri_boxes_2d")
    iptr, ninc = S.scan(ncells, nc)
    inc_cell, inc_tri = S.empty(ninc, torch.int32), S.empty(ninc, torch.int32)
    _lib.check(lib.mg_tri_incidence_2d(nc, box_c.data_ptr(), plo, pinv, G, iptr.data_ptr(), inc_cell.data_ptr(),
                                       inc_tri.data_ptr(), st), "mg_tri_incidence_2d")
    bits = max(1, int(np.ceil(np.log2(max(G * G, 2)))))
    order = S.argsort_i32(inc_cell, bits)                     # stable: triangle ids stay ascending inside a cell
    cell_tri = inc_tri[order.long()].contiguous()
    cell_sorted = inc_cell[order.long()].contiguous()
    cell_ptr = torch.searchsorted(cell_sorted, torch.arange(G * G + 1, dtype=torch.int32, device=dev)).to(torch.int32)
    count = S.empty(nf, torch.int32)
    _lib.check(lib.mg_tri_pairs_2d(nf, box_f.data_ptr(), box_c.data_ptr(), plo, pinv, G, cell_ptr.data_ptr(),
                                   cell_tri.data_ptr(), count.data_ptr(), None, None, None, st), "mg_tri_pairs_2d")
    pptr, K = S.scan(count, nf)
    pair_f, pair_c = S.empty(K, torch.int32), S.empty(K, torch.int32)
    if K:
        _lib.check(lib.mg_tri_pairs_2d(nf, box_f.data_ptr(), box_c.data_ptr(), plo, pinv, G, cell_ptr.data_ptr(),
                                       cell_tri.data_ptr(), None, pptr.data_ptr(), pair_f.data_ptr(), pair_c.data_ptr(),
                                       st), "mg_tri_pairs_2d")
    return pair_f, pair_c


def _pair_kernel_inputs(pf, tf, pc, tc, f, c):
    return (np.ascontiguousarray(f, dtype=np.int32), np.ascontiguousarray(c, dtype=np.int32),
            np.ascontiguousarray(pf, dtype=np.float64), np.ascontiguousarray(tf, dtype=np.int32),
            np.ascontiguousarray(pc, dtype=np.float64), np.ascontiguousarray(tc, dtype=np.int32))


def coupling_operator_2d_native(fine_mesh, coarse_mesh, where="device", pairs="device"):
    """The same operator through libmgb200 (csrc/assembly_kernels.cu: bounding boxes binned on a grid give the candidate
    pairs; one thread per candidate pair clips, triangulates and integrates; runs of equal (row, col) summed by the
    assembly fold).  where="device": CUDA kernels end to end, returns a device CSR (setup_device.DevCSR);
    where="host": the same per-pair C code run serially on the host with the NumPy binning (CPU tests), returns SciPy
    CSR."""
    import ctypes
    from .. import _lib
    pf = np.asarray(fine_mesh.get_points(), dtype=np.float64)
    tf = np.asarray(fine_mesh.get_connections(), dtype=np.int64)
    pc = np.asarray(coarse_mesh.get_points(), dtype=np.float64)
    tc = np.asarray(coarse_mesh.get_connections(), dtype=np.int64)
    lib = _lib.load()
    if where == "host":
        lib = _lib.load_testing()                # the serial host twin of the pair kernel is test infrastructure
        f, c = candidate_pairs(pf, tf, pc, tc)
        hf, hc, hpf, htf, hpc, htc = _pair_kernel_inputs(pf, tf, pc, tc, f, c)
        K = len(hf)
        rows, cols = np.empty(9 * K, dtype=np.int32), np.empty(9 * K, dtype=np.int32)
        vals = np.empty(9 * K)
        ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        _lib.check(lib.mg_host_coupling_pairs_p1_2d(K, ptr(hf), ptr(hc), ptr(hpf), ptr(htf), ptr(hpc), ptr(htc), ptr(rows),
                                                    ptr(cols), ptr(vals), None), "mg_host_coupling_pairs_p1_2d")
        B = sp.coo_matrix((vals, (rows, cols)), shape=(len(pf), len(pc))).tocsr()
        B.sum_duplicates()
        B.data[np.abs(B.data) < 1e-13 * np.abs(B.data).max()] = 0.0
        B.eliminate_zeros()
        B.sort_indices()
        return B
    from .. import setup_device as SD
    torch = _lib.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    S = SD.DeviceSetup(torch, dev)
    # meshes to the device once; candidate pairs by binning there (pairs="host": the NumPy binning, for cross-checks)
    d_pf, d_pc = (torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev) for a in (pf, pc))
    d_tf, d_tc = (torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev) for a in (tf, tc))
    if pairs == "host":
        f, c = candidate_pairs(pf, tf, pc, tc)
        d_f = torch.from_numpy(np.ascontiguousarray(f, dtype=np.int32)).to(dev)
        d_c = torch.from_numpy(np.ascontiguousarray(c, dtype=np.int32)).to(dev)
    else:
        lo = np.minimum(pf.min(axis=0), pc.min(axis=0))
        hi = np.maximum(pf.max(axis=0), pc.max(axis=0))
        d_f, d_c = candidate_pairs_device(S, d_pf, d_tf, d_pc, d_tc, lo, hi)
    K = int(d_f.numel())
    d = [d_f, d_c, d_pf, d_tf, d_pc, d_tc]
    rows, cols, vals = S.empty(9 * K, torch.int32), S.empty(9 * K, torch.int32), S.empty(9 * K, torch.float64)
    st = _lib.stream_handle(torch)
    _lib.check(lib.mg_coupling_pairs_p1_2d(K, *[t.data_ptr() for t in d], rows.data_ptr(), cols.data_ptr(),
                                           vals.data_ptr(), None, st), "mg_coupling_pairs_p1_2d")
    m = 9 * K
    order = S.row_col_order(rows, cols, len(pf), len(pc))
    head, folded = S.empty(m, torch.int32), S.empty(m, torch.float64)
    _lib.check(lib.mg_coo_fold_sum(m, rows.data_ptr(), cols.data_ptr(), vals.data_ptr(), order.data_ptr(),
                                   head.data_ptr(), folded.data_ptr(), st), "mg_coo_fold_sum")
    live = head.bool()
    thr = 1e-13 * float(folded[live].abs().max().item())
    head[live & (folded.abs() < thr)] = 0                       # slivers between triangles that merely touch
    slot, nnz = S.scan(head, m)
    orow, ocol, oval = S.empty(nnz, torch.int32), S.empty(nnz, torch.int32), S.empty(nnz, torch.float64)
    _lib.check(lib.mg_nn_emit(m, rows.data_ptr(), cols.data_ptr(), order.data_ptr(), head.data_ptr(), slot.data_ptr(),
                              folded.data_ptr(), orow.data_ptr(), ocol.data_ptr(), oval.data_ptr(), st), "mg_nn_emit")
    indptr = torch.searchsorted(orow, torch.arange(len(pf) + 1, dtype=torch.int32, device=dev)).to(torch.int32)
    return SD.DevCSR((len(pf), len(pc)), indptr, ocol, oval)


def coupling_operator_2d(fine_mesh, coarse_mesh, return_pairs=False):
    """B (n_fine x n_coarse CSR) with B[f, c] = int phi_f phi_c; meshes expose get_points() / get_connections().
    NumPy model of the kernel in csrc/assembly_kernels.cu (coupling_operator_2d_native)."""
    pf = np.asarray(fine_mesh.get_points(), dtype=np.float64)
    tf = np.asarray(fine_mesh.get_connections(), dtype=np.int64)
    pc = np.asarray(coarse_mesh.get_points(), dtype=np.float64)
    tc = np.asarray(coarse_mesh.get_connections(), dtype=np.int64)
    f, c = candidate_pairs(pf, tf, pc, tc)
    K = len(f)
    poly = np.zeros((K, MAXV, 2))
    poly[:, :3] = pf[tf[f]]
    count = np.full(K, 3, dtype=np.int64)
    flip = _signed_area2(pc, tc)[c] < 0                      # clip against counter-clockwise coarse triangles
    cv = pc[tc[c]]
    cv[flip] = cv[flip][:, ::-1]
    for e in range(3):
        poly, count = clip_polygons(poly, count, cv[:, e], cv[:, (e + 1) % 3])
    rows, cols, vals = [], [], []
    for i in range(1, MAXV - 1):                             # fan triangles (v0, v_i, v_{i+1})
        sel = np.flatnonzero(count >= i + 2)
        if len(sel) == 0:
            break
        v0, v1, v2 = poly[sel, 0], poly[sel, i], poly[sel, i + 1]
        area = 0.5 * np.abs((v1[:, 0] - v0[:, 0]) * (v2[:, 1] - v0[:, 1]) - (v1[:, 1] - v0[:, 1]) * (v2[:, 0] - v0[:, 0]))
        nz = area > 0
        sel, v0, v1, v2, area = sel[nz], v0[nz], v1[nz], v2[nz], area[nz]
        loc = np.zeros((len(sel), 3, 3))
        for x in (0.5 * (v0 + v1), 0.5 * (v1 + v2), 0.5 * (v2 + v0)):          # degree-2 exact, weights area/3
            lf = _barycentric(pf, tf, f[sel], x)
            lc = _barycentric(pc, tc, c[sel], x)
            loc += lf[:, :, None] * lc[:, None, :]
        loc *= (area / 3.0)[:, None, None]
        rows.append(np.repeat(tf[f[sel]], 3, axis=1).reshape(-1))
        cols.append(np.tile(tc[c[sel]], (1, 3)).reshape(-1))
        vals.append(loc.reshape(-1))
    B = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                      shape=(len(pf), len(pc))).tocsr()
    B.sum_duplicates()
    B.data[np.abs(B.data) < 1e-13 * np.abs(B.data).max()] = 0.0      # slivers between triangles that merely touch
    B.eliminate_zeros()
    B.sort_indices()
    if return_pairs:
        area = np.zeros(K)
        for i in range(1, MAXV - 1):
            sel = np.flatnonzero(count >= i + 2)
            v0, v1, v2 = poly[sel, 0], poly[sel, i], poly[sel, i + 1]
            area[sel] += 0.5 * np.abs((v1[:, 0] - v0[:, 0]) * (v2[:, 1] - v0[:, 1]) - (v1[:, 1] - v0[:, 1]) * (v2[:, 0] - v0[:, 0]))
        hit = area > 0
        return B, np.stack([f[hit], c[hit]], axis=1), area[hit]
    return B
