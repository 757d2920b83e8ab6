"""GPU parity of the NN transfer-operator builder (learnmultigrid_b200/neural2d.py, csrc/nn_kernels.cu) against the
golden vectors produced by the REFERENCE's NeuralMG_2D methods (tests/golden/neural_2d_cases.npz), level by level, and
of the V-cycle that runs on the resulting hierarchy (BASELINE configs[1]) against the CPU oracle.

Bars: coarse sets, fill indices and the d_neighs tables exact; patches and B bit for bit (copies, one division /
multiplication, and the ordered running mean); Q to 4 ulp (row sums are added in column order here, pairwise over the
dense row by NumPy in the reference); next-level mass matrices to 1e-13 relative (Q^T M Q evaluated in another order).
"""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import assert_history_close, load_golden
from test_neural2d_host import CASES, levels_of

pytestmark = pytest.mark.gpu


class Stub:
    """the predictor make_golden_neural.py used"""

    def predict(self, X):
        X = np.asarray(X, dtype=np.float64)
        w = np.linspace(0.5, 1.5, 31)[None, :]
        return (1.0 + np.tanh(X.sum(axis=1, keepdims=True))) * w


@pytest.fixture(scope="module")
def nb():
    import torch
    assert torch.cuda.is_available()
    from learnmultigrid_b200.neural2d import NeuralBuilder
    return NeuralBuilder()


@pytest.mark.parametrize("case", CASES)
def test_device_builder_matches_reference_level_by_level(nb, case):
    g = load_golden("neural_2d_cases.npz")
    torch = nb.torch
    for d in levels_of(g, case):
        M = nb.upload(d["M"])
        n = d["M"].shape[0]
        cmap, clist, nc = nb.coarsen(M)
        assert np.array_equal(clist.cpu().numpy(), d["C"])
        patches, fill = nb.extract(M, cmap, clist)
        assert np.array_equal(fill.cpu().numpy(), d["fill"])
        assert np.array_equal(patches.cpu().numpy(), d["patches"])
        pred = torch.from_numpy(np.ascontiguousarray(d["pred"])).to(nb.dev)
        B, dn = nb.fill_B(pred, fill, cmap, n, nc, not_last=nb.not_last)      # flags from the variant kernel
        assert np.array_equal(dn.cpu().numpy(), d["dn"])
        if fill.shape[0] > nc:                                                  # patch variants (cases f81, f169)
            assert int(nb.not_last.sum()) == fill.shape[0] - nc                 # a node with a extra variants: a flags
            B2, dn2 = nb.fill_B(pred, fill, cmap, n, nc)                        # flags derived from the fill indices
            assert np.array_equal(dn2.cpu().numpy(), d["dn"])
            assert np.array_equal(nb.download(B2).toarray(), d["B"])
        Bh = nb.download(B)
        assert np.array_equal(Bh.toarray(), d["B"])
        assert Bh.nnz == np.count_nonzero(d["B"]) and Bh.has_sorted_indices
        Qh = nb.download(nb.normalise(B)).toarray()
        np.testing.assert_allclose(Qh, d["Q"], rtol=1e-15, atol=0)
        np.testing.assert_allclose(Qh.sum(axis=1), 1.0, rtol=0, atol=1e-15)      # partition of unity


@pytest.mark.parametrize("case", CASES)
def test_device_define_hierarchy_matches_reference(nb, case):
    g = load_golden("neural_2d_cases.npz")
    lv = levels_of(g, case)
    mean, std = (g["mean2"], g["std2"]) if case in ("r77", "i289", "f169") else (np.zeros(43), np.ones(43))
    Qs = nb.define_hierarchy(lv[0]["M"], Stub(), mean, std, len(lv) + 1, keep_intermediates=True)
    assert len(Qs) == len(lv)
    for l, (Q, d) in enumerate(zip(Qs, lv)):
        tr = nb.trace[l]
        assert np.array_equal(tr["clist"].cpu().numpy(), d["C"])
        assert np.array_equal(tr["fill"].cpu().numpy(), d["fill"])
        Mh = nb.download(tr["M"])
        assert np.array_equal(Mh.indices, d["M"].indices) and np.array_equal(Mh.indptr, d["M"].indptr)
        np.testing.assert_allclose(Mh.data, d["M"].data, rtol=1e-13)
        np.testing.assert_allclose(tr["patches"].cpu().numpy(), d["patches"], rtol=1e-13)
        np.testing.assert_allclose(nb.download(Q).toarray(), d["Q"], rtol=1e-12, atol=1e-300)


def test_api_neuralmg_2d_builds_and_solves_on_an_irregular_mesh(nb):
    """configs[1] at test size: irregular refined mesh (1089 nodes), 3-level hierarchy from the mass matrix through the
    predictor interface, multicolour Gauss-Seidel V(3,3) as in the 2D scripts (thesis_structured_2d.py:457-458);
    residual history and iteration count against the CPU oracle on the same transfer operators"""
    from learnmultigrid_b200 import problems as P
    from learnmultigrid_b200.neural2d import MassSurrogate
    from learnmultigrid_b200.solvers.Multigrid import NeuralMG_2D
    from oracle.vcycle import OracleMultigrid
    pb = P.irregular_p1_2d(32, seed=3)
    A, rhs, M = pb["A"], pb["rhs"], pb["M"]
    mg = NeuralMG_2D(A, rhs, MassSurrogate(), M, np.ones(43), np.zeros(43))
    mg.define_hierarchy(levels=3)
    Qs = mg.l_hierarchy
    assert [q.shape[0] for q in Qs] == [1089, Qs[0].shape[1]] and Qs[1].shape[1] < Qs[0].shape[1] < 1089 / 3
    for q in Qs:
        np.testing.assert_allclose(np.asarray(q.sum(axis=1)).ravel(), 1.0, atol=1e-14)
    mg.solve(levels=3, smoother="GaussSeidel", smooth_steps=3, error=1e-9, max_iterations=40)
    o = OracleMultigrid(A, rhs, Qs, smoother="mcgs", colors=mg.get_hierarchy().colors, hoist_setup=True)
    o.solve(levels=3, smooth_steps=3, error=1e-9, max_iterations=40)
    assert mg.get_iterations() == len(o.track_res) < 40
    assert_history_close(mg.track_res, o.track_res, A, o.solution)
    np.testing.assert_allclose(mg.get_solution(), o.solution, rtol=0, atol=1e-12 * np.linalg.norm(o.solution))


def test_api_helpers_keep_the_reference_signatures(nb):
    from learnmultigrid_b200.solvers.Multigrid import NeuralMG_2D
    g = load_golden("neural_2d_cases.npz")
    d = levels_of(g, "i81")[0]
    n = d["M"].shape[0]
    mg = NeuralMG_2D(sp.identity(n, format="csr"), np.zeros((n, 1)), Stub(), d["M"], np.ones(43), np.zeros(43))
    C, Fn, Cn, Fnn = mg.coarsening(sp.lil_matrix(d["M"]))
    assert C == list(d["C"]) and sorted(C + Fn) == list(range(n))
    patches, fill = mg.extract_patches(C, sp.lil_matrix(d["M"]))
    assert np.array_equal(patches, d["patches"]) and np.array_equal(fill, d["fill"])
    B, dn = mg.fill_B(d["pred"], fill, n, mg.map_coarse(C), C)
    assert np.array_equal(B.toarray(), d["B"])
    assert all(np.array_equal(dn[k], d["dn"][k][d["dn"][k] >= 0]) for k in range(len(C)))
    mg.define_hierarchy(levels=3)
    np.testing.assert_allclose(mg.l_hierarchy[0].toarray(), d["Q"], rtol=1e-15)


def test_torch_mlp_predictor_runs_on_device(nb):
    from learnmultigrid_b200 import problems as P
    from learnmultigrid_b200.neural2d import TorchMLP
    pb = P.irregular_p1_2d(32, seed=5)
    Qs = nb.define_hierarchy(pb["M"], TorchMLP(hidden=(64, 64), seed=1), np.zeros(43), np.ones(43) * pb["M"].data.max(), 3)
    for Q in Qs:
        Qh = nb.download(Q)
        np.testing.assert_allclose(np.asarray(Qh.sum(axis=1)).ravel(), 1.0, atol=1e-13)
        assert Qh.data.min() > 0


@pytest.mark.parametrize("name", ["reg64", "irr128"])
def test_api_neuralmg_1d_reproduces_reference_history(name):
    """1D NeuralMG.solve (Multigrid.py:211-301) with the reference's smoother (index-order Gauss-Seidel): the
    reference's own residual history, iteration count and solution (tests/golden/neural_1d.npz)"""
    from test_neural1d_host import make
    g = load_golden("neural_1d.npz")
    mg = make(g, name)
    levels, steps = (int(v) for v in g[name + "_par"])
    mg.solve(levels=levels, smoother="GaussSeidel", smooth_steps=steps, error=1e-10, max_iterations=40,
             initial_guess=np.zeros((g[name + "_A"].shape[0], 1)), gs_order="lexicographic")
    want = g[name + "_hist"]
    assert mg.get_iterations() == len(want)
    assert_history_close(mg.track_res, want, g[name + "_A"], g[name + "_sol"])
    np.testing.assert_allclose(mg.get_solution(), g[name + "_sol"], rtol=0, atol=1e-11 * np.linalg.norm(g[name + "_sol"]))
    # and the default (multicolour) smoother converges as well
    mg2 = make(g, name)
    mg2.solve(levels=levels, smoother="GaussSeidel", smooth_steps=steps, error=1e-10, max_iterations=40)
    assert mg2.get_iterations() <= len(want) + 3
