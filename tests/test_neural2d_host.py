"""The per-node logic of the NN transfer-operator builder (csrc/nn_kernels.cu: coarsening, patch extraction, fill_B
contributions), run serially on host arrays through the mg_host_nn_* entry points, against the golden vectors the
REFERENCE's own NeuralMG_2D methods produced (tests/golden/make_golden_neural.py -> neural_2d_cases.npz).
The CUDA kernels run the same __host__ __device__ functions; tests/test_gpu_neural2d.py checks them on the GPU."""
import ctypes

import numpy as np
import pytest
import scipy.sparse as sp

from learnmultigrid_b200 import _lib
from helpers import load_golden

CASES = ["s81", "r77", "s289", "i81", "i289", "f81", "f169"]
VARIANTS = {"f81": 6, "f169": 12}      # extra patch variants on the fine level (coarse nodes with 7 / 8 neighbours)


def level_data(g, case, l):
    key = "%s__l%d_" % (case, l)
    if key + "C" not in g.files:
        return None
    M = sp.csr_matrix((g[key + "M_data"], (g[key + "M_row"], g[key + "M_col"])), shape=tuple(g[key + "M_shape"]))
    M.sort_indices()
    return {"M": M, "C": g[key + "C"], "patches": g[key + "patches"], "fill": g[key + "fill"], "pred": g[key + "pred"],
            "B": g[key + "B"], "Q": g[key + "Q"], "dn": g[key + "dn"]}


def levels_of(g, case):
    out, l = [], 0
    while True:
        d = level_data(g, case, l)
        if d is None:
            return out
        out.append(d)
        l += 1


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def host_coarsen(M):
    lib = _lib.load_testing()
    MT = sp.csr_matrix(M.T)
    MT.sort_indices()
    n = M.shape[0]
    ip, ix, va = MT.indptr.astype(np.int32), MT.indices.astype(np.int32), MT.data.astype(np.float64)
    cmap = np.empty(n, dtype=np.int32)
    clist = np.empty(n, dtype=np.int32)
    nc = lib.mg_host_nn_coarsen(n, ptr(ip), ptr(ix), ptr(va), ptr(cmap), ptr(clist))
    assert nc > 0
    return cmap, clist[:nc].copy()


def fold_reference_rule(rows, cols, vals, n, nc, unused):
    """the running mean of fill_B applied in contribution order (tiny cases: plain Python)"""
    B = np.zeros((n, nc))
    for r, c, v in zip(rows.ravel(), cols.ravel(), vals.ravel()):
        if r >= unused:
            continue
        B[r, c] = v if B[r, c] == 0 else (B[r, c] + v) / 2.0
    return B


@pytest.mark.parametrize("case", CASES)
def test_coarsening_extraction_and_fill_match_the_reference(case):
    lib = _lib.load_testing()
    g = load_golden("neural_2d_cases.npz")
    for l, d in enumerate(levels_of(g, case)):
        M = d["M"]
        n = M.shape[0]
        cmap, clist = host_coarsen(M)
        assert np.array_equal(clist, d["C"])                         # same coarse nodes, same order
        nc = len(clist)
        ip, ix, va = M.indptr.astype(np.int32), M.indices.astype(np.int32), M.data.astype(np.float64)
        rows_p = lib.mg_host_nn_extract_patches(nc, ptr(ip), ptr(ix), ptr(va), ptr(cmap), ptr(clist), None, None)
        assert rows_p == d["patches"].shape[0]                       # nc + extra patch variants (Multigrid.py:631-663)
        if l == 0:
            assert rows_p - nc == VARIANTS.get(case, 0)
        else:
            assert rows_p == nc                                      # pre_process cuts coarse rows to 6 neighbours
        patches = np.empty((rows_p, 43))
        fill = np.empty((rows_p, 31), dtype=np.int32)
        assert lib.mg_host_nn_extract_patches(nc, ptr(ip), ptr(ix), ptr(va), ptr(cmap), ptr(clist), ptr(patches),
                                              ptr(fill)) == rows_p, lib.mg_last_error()
        assert np.array_equal(fill, d["fill"])
        assert np.array_equal(patches, d["patches"])                 # bit for bit
        pred = np.ascontiguousarray(d["pred"])
        rows = np.empty((rows_p, 31), dtype=np.int32)
        cols = np.empty((rows_p, 31), dtype=np.int32)
        vals = np.empty((rows_p, 31))
        dn = -np.ones((nc, 6), dtype=np.int32)
        _lib.check(lib.mg_host_nn_contributions(rows_p, ptr(fill), ptr(pred), ptr(cmap), n, ptr(rows), ptr(cols),
                                                ptr(vals), ptr(dn)), "mg_host_nn_contributions")
        assert np.array_equal(dn, d["dn"])
        B = fold_reference_rule(rows, cols, vals, n, nc, n)
        assert np.array_equal(B, d["B"])                             # bit for bit, order-dependent means included
        Q = B / B.sum(axis=1)[:, None]
        assert np.array_equal(Q, d["Q"])


def star(arms):
    n = arms + 1
    A = sp.lil_matrix((n, n))
    A.setdiag(4.0)
    A[0, 1:n] = 1.0 + 0.01 * np.arange(arms)
    A[1:n, 0] = 1.0
    for i in range(1, n):                 # a ring, so that no neighbour is left with fewer than 2 neighbours
        A[i, 1 + i % arms] = 0.5
        A[1 + i % arms, i] = 0.5
    M = sp.csr_matrix(A)
    M.sort_indices()
    return M


def test_variants_drop_a_sliding_window_of_the_smallest_entries():
    """a hub with 9 neighbours of increasing mass entry: variant v keeps all but the entries ranked v..v+2"""
    lib = _lib.load_testing()
    M = star(9)
    cmap, clist = host_coarsen(M)
    assert clist[0] == 0
    ip, ix, va = M.indptr.astype(np.int32), M.indices.astype(np.int32), M.data.astype(np.float64)
    rows_p = lib.mg_host_nn_extract_patches(len(clist), ptr(ip), ptr(ix), ptr(va), ptr(cmap), ptr(clist), None, None)
    assert rows_p == len(clist) + 3
    patches = np.empty((rows_p, 43))
    fill = np.empty((rows_p, 31), dtype=np.int32)
    assert lib.mg_host_nn_extract_patches(len(clist), ptr(ip), ptr(ix), ptr(va), ptr(cmap), ptr(clist), ptr(patches),
                                          ptr(fill)) == rows_p
    variants = [0] + list(range(len(clist), rows_p))
    for v, j in enumerate(variants):
        kept = [k for k in range(1, 10) if not (v <= k - 1 < v + 3)]
        assert list(fill[j, :7]) == [0] + kept
        np.testing.assert_array_equal(patches[j, 1:7], 1.0 + 0.01 * (np.array(kept) - 1))


def test_more_than_twelve_neighbours_is_reported_not_worked_around():
    lib = _lib.load_testing()
    M = star(13)
    cmap, clist = host_coarsen(M)
    ip, ix, va = M.indptr.astype(np.int32), M.indices.astype(np.int32), M.data.astype(np.float64)
    patches = np.empty((len(clist) + 7, 43))
    fill = np.empty((len(clist) + 7, 31), dtype=np.int32)
    rc = lib.mg_host_nn_extract_patches(len(clist), ptr(ip), ptr(ix), ptr(va), ptr(cmap), ptr(clist), ptr(patches), ptr(fill))
    assert rc == -3 and b"more than 12 neighbours" in lib.mg_last_error()


def test_first_reference_run_with_the_stub_predictor():
    """tests/golden/neural_2d_stub.npz: the reference's NeuralMG_2D on Mesh2D(64) (SURVEY 8c probe: the 25 coarse nodes
    (2j, 2k), patches (25, 43), Q shapes (81, 25) and (25, 9))"""
    lib = _lib.load_testing()
    g = load_golden("neural_2d_stub.npz")
    M = sp.csr_matrix((g["M_data"], (g["M_row"], g["M_col"])), shape=tuple(g["M_shape"]))
    M.sort_indices()
    cmap, clist = host_coarsen(M)
    assert np.array_equal(clist, g["C_order"]) and np.array_equal(np.sort(clist), g["C"])
    assert np.array_equal(clist, [9 * (2 * j) + 2 * k for j in range(5) for k in range(5)])
    ip, ix, va = M.indptr.astype(np.int32), M.indices.astype(np.int32), M.data.astype(np.float64)
    patches = np.empty((25, 43))
    fill = np.empty((25, 31), dtype=np.int32)
    assert lib.mg_host_nn_extract_patches(25, ptr(ip), ptr(ix), ptr(va), ptr(cmap), ptr(clist), ptr(patches),
                                          ptr(fill)) == 25
    assert np.array_equal(patches, g["patches"]) and np.array_equal(fill, g["fill"])
    h2 = (1.0 / 8) ** 2                                                   # interior patch of SURVEY 8c
    k = list(clist).index(40)
    np.testing.assert_allclose(patches[k], [h2 / 2] + [h2 / 12] * 6 + ([h2 / 2] + [h2 / 12] * 5) * 6, rtol=1e-12)
    assert list(fill[k, :13]) == [40, 30, 31, 39, 41, 49, 50, 20, 22, 38, 42, 58, 60]
    assert g["Q0"].shape == (81, 25) and g["Q1"].shape == (25, 9)


@pytest.mark.parametrize("case", CASES)
def test_next_level_mass_matrix_is_the_galerkin_product(case):
    """define_hierarchy's `mass = Q^T mass Q` (Multigrid.py:763): the stored next-level / final matrices are the SciPy
    products of the stored Q and M, pattern and values (what the device SpGEMM is checked against on the GPU)"""
    g = load_golden("neural_2d_cases.npz")
    lv = levels_of(g, case)
    for l, d in enumerate(lv):
        Q = sp.csr_matrix(d["Q"])
        prod = sp.csr_matrix(Q.T @ d["M"] @ Q)
        prod.sort_indices()
        if l + 1 < len(lv):
            # the next level's stored matrix is the product after pre_process cut its rows to the predicted neighbours
            nxt = lv[l + 1]["M"]
            dense = prod.toarray()
            keep = nxt.toarray() != 0
            assert np.all(dense[keep] != 0)
            np.testing.assert_allclose(nxt.toarray()[keep], dense[keep], rtol=1e-13)
            assert np.all(np.diff(nxt.indptr) <= 7)                  # diagonal + at most 6 predicted neighbours
        else:
            key = "%s__final_M_" % case
            fin = sp.csr_matrix((g[key + "data"], (g[key + "row"], g[key + "col"])), shape=tuple(g[key + "shape"]))
            np.testing.assert_allclose(fin.toarray(), prod.toarray(), rtol=1e-13, atol=1e-300)


def test_virtual_nodes_and_connectivity_helpers_against_the_reference_patches():
    """NeuralMG_2D.create_virtual_nodes / get_conn (Multigrid.py:391-398, 438-454): for every coarse node of the
    structured golden case with fewer than 6 neighbours, the padded neighbour entries and the virtual patch rows the
    helper returns are the ones in the patch the REFERENCE extracted"""
    from learnmultigrid_b200.solvers.Multigrid import NeuralMG_2D
    g = load_golden("neural_2d_cases.npz")
    d = level_data(g, "s81", 0)
    M = d["M"]
    nmg = NeuralMG_2D.__new__(NeuralMG_2D)                  # the helpers use no state
    conn = NeuralMG_2D.get_conn(M)
    off = M.toarray() - np.diag(M.diagonal())
    assert np.array_equal(conn, (off > 0).astype(float)) and M.diagonal().min() > 0      # argument untouched
    seen = set()
    for k, c in enumerate(d["C"]):
        nb = np.flatnonzero(conn[c])
        vals = np.sort(off[c, nb])[::-1]                    # the reference orders a node's neighbours by mass entry
        node_M = M[c, c]
        virt, row = nmg.create_virtual_nodes(vals, node_M)
        patch = d["patches"][k]
        assert patch[0] == node_M and len(row) == 6
        assert np.array_equal(np.sort(row)[::-1], np.sort(patch[1:7])[::-1])
        if len(nb) < 6:
            assert len(virt) == 6 * (6 - len(nb))
            assert np.array_equal(patch[43 - len(virt):], virt)
            seen.add(len(nb))
        else:
            assert len(virt) == 0
    assert seen >= {2, 3, 4}                                # corners of both kinds and edge nodes
