"""1D NeuralMG (solvers/Multigrid.py::NeuralMG): features and transfer operators against the reference's own
prepare_nn_input / transfer_op (tests/golden/make_golden_neural1d.py -> neural_1d.npz).  Host logic only; the solve
runs on the GPU (tests/test_gpu_neural2d.py::test_api_neuralmg_1d_*)."""
import numpy as np
import pytest

from helpers import load_golden


class Stub:
    def __init__(self, mean, std, eps=0.05):
        self.mean, self.std, self.eps = mean, std, eps

    def predict(self, X):
        X = np.asarray(X, dtype=np.float64) * self.std + self.mean
        out = np.zeros((X.shape[0], 9))
        out[:, 2] = 0.5 * X[:, 2]
        out[:, 4] = X[:, 2] + 0.5 * X[:, 1]
        out[:, 5] = X[:, 3] + 0.5 * (X[:, 2] + X[:, 4])
        out[:, 6] = X[:, 4] + 0.5 * X[:, 5]
        out[:, 8] = 0.5 * X[:, 4]
        out *= 1.0 + self.eps * np.sin(np.arange(X.shape[0]))[:, None]
        out[:, [0, 1, 3, 7]] = -7.0
        return out


def make(g, name):
    from learnmultigrid_b200.solvers.Multigrid import NeuralMG
    std = g[name + "_std"]
    return NeuralMG(g[name + "_A"], g[name + "_rhs"], Stub(np.zeros(7), std), g[name + "_M"], std, np.zeros(7))


@pytest.mark.parametrize("name", ["reg64", "irr128"])
def test_features_and_transfer_operators_match_reference(name):
    g = load_golden("neural_1d.npz")
    mg = make(g, name)
    levels = int(g[name + "_par"][0])
    M = g[name + "_M"]
    for l in range(levels - 1):
        assert np.array_equal(mg.prepare_nn_input(M), g["%s_feat%d" % (name, l)])
        Q = mg.transfer_op(M)
        assert isinstance(Q, np.ndarray)                     # dense in, dense out, like the reference
        want = g["%s_Q%d" % (name, l)]
        assert np.array_equal(Q != 0, want != 0)
        np.testing.assert_allclose(Q, want, rtol=1e-14, atol=0)
        np.testing.assert_allclose(Q.sum(axis=1), 1.0, atol=1e-15)
        M = want.T @ M @ want                                # the reference's own coarse mass for the next level
    mg.define_hierarchy(levels)
    assert [q.shape for q in mg.l_hierarchy] == [g["%s_Q%d" % (name, l)].shape for l in range(levels - 1)]


def test_v_cycle_takes_the_reference_argument_list(monkeypatch):
    """NeuralMG.v_cycle(A, M, u0, rhs, ...) (Multigrid.py:246): the mass matrix comes second; the transfer operators
    handed to the engine are the ones predicted from THAT M.  The device part is replaced by a recorder here."""
    from learnmultigrid_b200.solvers import Multigrid as MGmod
    g = load_golden("neural_1d.npz")
    mg = make(g, "reg64")
    levels = int(g["reg64_par"][0])
    seen = {}

    class FakeHierarchy:
        def make_params(self, **kw):
            return kw

        def set_rhs(self, b):
            seen["rhs"] = b

        def set_x(self, x):
            seen["x"] = x

        def vcycle(self, params):
            seen["params"] = params

        def get_x(self):
            return seen["x"]

    def fake_build(self, A, lv, smoother, gs_order, colors, first_call):
        seen["Q"] = self._transfer_list(lv, first_call)
        return FakeHierarchy()
    monkeypatch.setattr(MGmod.Multigrid, "_build", fake_build)
    A, M = g["reg64_A"], g["reg64_M"]
    u0, rhs = np.zeros((A.shape[0], 1)), g["reg64_rhs"]
    out = mg.v_cycle(A, M, u0, rhs, "GaussSeidel", 3, 1e-10, levels)
    assert out is u0 and seen["params"]["nu_pre"] == 3 and seen["rhs"] is rhs
    assert len(seen["Q"]) == levels - 1
    for l, Q in enumerate(seen["Q"]):
        np.testing.assert_allclose(Q.toarray(), g["reg64_Q%d" % l], rtol=1e-14, atol=0)
    M2 = 2.0 * M                                            # another mass matrix: the operators are rebuilt from it
    mg.v_cycle(A, M2, u0, rhs, "GaussSeidel", 1, 1e-10, levels)
    assert mg.M is M2 and len(mg.l_hierarchy) == levels - 1
