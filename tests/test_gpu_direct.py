"""General sparse direct solve on the device (coarse.build_coarse_solver, solvers.Solver.DirectSolver): replaces
scipy.sparse.linalg.spsolve (learn_multigrid/solvers/Solver.py:56-59, Multigrid.py:106) for ANY sparse system -- dense
inverse, block cyclic reduction, or reverse Cuthill-McKee + block cyclic reduction for numberings that are not banded."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

pytestmark = pytest.mark.gpu


def test_direct_solver_on_an_unstructured_20k_system():
    import torch
    from learnmultigrid_b200 import problems as P
    from learnmultigrid_b200.solvers.Solver import DirectSolver
    torch.cuda.set_device(0)
    pb = P.irregular_p1_2d(140, seed=3)                  # 141^2 = 19 881 unknowns, parents-first numbering: not banded
    A, b = pb["A"], pb["rhs"]
    want = spla.spsolve(sp.csc_matrix(A), b).reshape(-1, 1)
    ds = DirectSolver(A, b)
    ds.solve()
    assert "Cuthill" in ds.method
    np.testing.assert_allclose(ds.get_solution(), want, rtol=0, atol=1e-12 * np.linalg.norm(want))
    assert ds.get_residual() <= 1e-12 * np.linalg.norm(b)
    # a random symmetric permutation of a structured operator: banded only after reordering
    N = 150
    S = P.structured_laplacian_2d(N)
    n = S.shape[0]
    p = np.random.default_rng(0).permutation(n)
    Sp = sp.csr_matrix(S[p][:, p])
    bs = np.random.default_rng(1).standard_normal((n, 1))
    want = spla.spsolve(sp.csc_matrix(Sp), bs).reshape(-1, 1)
    ds = DirectSolver(Sp, bs)
    ds.solve()
    assert "Cuthill" in ds.method
    np.testing.assert_allclose(ds.get_solution(), want, rtol=0, atol=1e-12 * np.linalg.norm(want))
    # small systems keep the dense inverse, banded ones plain block cyclic reduction
    small = DirectSolver(P.structured_laplacian_2d(30), np.ones((31 * 31, 1)))
    small.solve()
    assert small.method == "dense inverse"
    banded = DirectSolver(S, bs)
    banded.solve()
    assert banded.method == "block cyclic reduction"
    np.testing.assert_allclose(banded.get_solution(), spla.spsolve(sp.csc_matrix(S), bs).reshape(-1, 1), rtol=0,
                               atol=1e-12 * np.linalg.norm(bs))


def test_hierarchy_with_shuffled_numbering_solves_its_coarsest_level():
    """C2-style hierarchy (irregular mesh, NN-built transfers) whose fine numbering is shuffled: the coarsest operator
    (> 4096 unknowns, unbanded) goes through RCM + BCR inside the V-cycle; history against the CPU oracle"""
    import torch
    from learnmultigrid_b200 import problems as P
    from learnmultigrid_b200.neural2d import MassSurrogate
    from learnmultigrid_b200.solvers.Multigrid import NeuralMG_2D, SemiGeometricMG
    from oracle.vcycle import OracleMultigrid
    from helpers import assert_history_close
    torch.cuda.set_device(0)
    pb = P.irregular_p1_2d(192, seed=5)                  # 37 249 unknowns -> 2 levels: ~9.4 k coarse unknowns
    nmg = NeuralMG_2D(pb["A"], pb["rhs"], MassSurrogate(), pb["M"], np.ones(43), np.zeros(43))
    nmg.define_hierarchy(2)
    Q = sp.csr_matrix(nmg.l_hierarchy[0])
    assert Q.shape[1] > 4096
    rng = np.random.default_rng(9)
    pf, pc = rng.permutation(Q.shape[0]), rng.permutation(Q.shape[1])
    A = sp.csr_matrix(pb["A"][pf][:, pf])
    Qp = sp.csr_matrix(Q[pf][:, pc])
    rhs = pb["rhs"][pf]
    mg = SemiGeometricMG(A, rhs, Qp)
    mg.solve(levels=2, smoother="GaussSeidel", smooth_steps=2, error=1e-9, max_iterations=40)
    h = mg.get_hierarchy()
    assert getattr(h.levels[-1].coarse, "perm", None) is not None        # reordered block cyclic reduction
    o = OracleMultigrid(A, rhs, [Qp], smoother="mcgs", colors=h.colors, hoist_setup=True)
    o.solve(levels=2, smooth_steps=2, error=1e-9, max_iterations=40)
    assert mg.get_iterations() == len(o.track_res) < 40
    assert_history_close(mg.track_res, o.track_res, A, o.solution)
    np.testing.assert_allclose(mg.get_solution(), o.solution, rtol=0, atol=1e-11 * np.linalg.norm(o.solution))
