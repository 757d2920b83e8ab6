"""The 1D path of SURVEY 8a rows a16-a21 -- Mesh1D, P1 mass / stiffness / load assembly, Intersection, CouplingOperator
and L2Projection (quasi / pseudo / L2) -- against the outputs of the REFERENCE's own classes
(tests/golden/transfer_1d_small.npz and c1_1d_1024.npz, made by tests/golden/make_golden.py through oracle/refshim.py):
regular and seeded irregular nested pairs, the non-nested pair of test/testMG.py:44-49, and BASELINE configs[0]
(Mesh1D(1024) / Mesh1D(512)).  Bars: meshes, intersections, A, M, rhs, B and the quasi-L2 Q bit for bit (elementwise
arithmetic in the reference's order); "pseudo" / "L2" to 1e-13 / 1e-12 (dense LAPACK solves, which may round differently
on another host -- in the authoring container they are bit-identical too)."""
import numpy as np
import pytest

from helpers import coo_from, load_golden


def ones(x):
    return np.ones(shape=np.shape(x))


def problem_1d(ne, regular=True, seed=None):
    """tests/golden/make_golden.py::problem_1d with the product's classes (Dirichlet rows as test/test_NN.py:176-181)"""
    from learnmultigrid_b200.mesh.Mesh1D import Mesh1D
    from learnmultigrid_b200.assembly.MassMatrix import MassMatrix
    from learnmultigrid_b200.assembly.StiffnessMatrix import StiffnessMatrix
    from learnmultigrid_b200.assembly.LoadVector import LoadVector
    from learnmultigrid_b200.assembly.Quadrature import Quadrature
    from learnmultigrid_b200.assembly.ShapeFunction import Function, Gradient
    if seed is not None:
        np.random.seed(seed)
    m = Mesh1D(regular, ne)
    m.construct()
    A = StiffnessMatrix(m).compute_stiffness_1d(Gradient(2), Quadrature(3))
    M = MassMatrix(m).compute_mass_1d(Function(2), Quadrature(3))
    rhs = LoadVector(m).compute_rhs_1d(ones)
    rhs[0] = 0
    rhs[-1] = 0
    A[1, 0] = 0
    A[-2, -1] = 0
    A[0, :] = 0
    A[-1, :] = 0
    A[0, 0] = 1
    A[-1, -1] = 1
    return m, A, M, rhs


def coarse_for(tag, m):
    from learnmultigrid_b200.mesh.Mesh1D import Mesh1D
    mc = Mesh1D(True, 8)
    mc.construct()
    if tag == "irr":                 # nested coarse mesh: every other node of the irregular fine mesh
        mc.x = m.get_mesh()[::2].copy()
        mc.connection_matrix()
    return mc


@pytest.mark.parametrize("tag,regular,seed", [("reg", True, None), ("irr", False, 42)])
def test_small_1d_pipeline_equals_reference(tag, regular, seed):
    from learnmultigrid_b200.L2_projection.Intersection import Intersection
    from learnmultigrid_b200.L2_projection.CouplingOperator import CouplingOperator
    from learnmultigrid_b200.L2_projection.L2Projection import L2Projection
    from learnmultigrid_b200.assembly.Quadrature import Quadrature
    from learnmultigrid_b200.assembly.ShapeFunction import Function
    t = load_golden("transfer_1d_small.npz")
    m, A, M, rhs = problem_1d(16, regular, seed)
    mc = coarse_for(tag, m)
    assert np.array_equal(m.get_mesh(), t[tag + "_x_fine"])              # same np.random draws for the irregular mesh
    assert np.array_equal(mc.get_mesh(), t[tag + "_x_coarse"])
    assert np.array_equal(np.asarray(A), t[tag + "_A"])
    assert np.array_equal(np.asarray(M), t[tag + "_M"])
    assert np.array_equal(np.asarray(rhs), t[tag + "_rhs"])
    inter = Intersection(m, mc)
    inter.find_intersections1d()
    ints, coords, union = inter.get_info()
    assert np.array_equal(np.asarray(ints), t[tag + "_intersections"])
    assert np.array_equal(np.asarray(coords), t[tag + "_int_coord"])
    B = CouplingOperator(inter, m, mc).compute_b_1d(Quadrature(3), Function(2))
    assert np.array_equal(np.asarray(B), t[tag + "_B"])
    for typ, tol in (("quasi", 0.0), ("pseudo", 1e-13), ("L2", 1e-12)):
        Q, seconds = L2Projection(typ, m, mc).compute_transfer_1d()
        assert seconds >= 0.0
        np.testing.assert_allclose(np.asarray(Q), t[tag + "_Q_" + typ], rtol=tol, atol=tol)
    # partition of unity of every transfer type (rows sum to 1): what makes constants interpolate exactly
    Q, _ = L2Projection("quasi", m, mc).compute_transfer_1d()
    np.testing.assert_allclose(np.asarray(Q).sum(axis=1), 1.0, rtol=1e-14)


def test_non_nested_pair_equals_reference():
    """Mesh1DRefinement 2*2^3 = 16 vs 3*2^2 = 12 elements (test/testMG.py:44-49): fine nodes inside coarse elements"""
    from learnmultigrid_b200.mesh.Mesh1D import Mesh1DRefinement
    from learnmultigrid_b200.L2_projection.L2Projection import L2Projection
    t = load_golden("transfer_1d_small.npz")
    mf = Mesh1DRefinement(coarse_ne=2, n_ref=3)
    mf.construct()
    mcn = Mesh1DRefinement(coarse_ne=3, n_ref=2)
    mcn.construct()
    assert np.array_equal(mf.get_mesh(), t["nonnested_x_fine"]) and np.array_equal(mcn.get_mesh(), t["nonnested_x_coarse"])
    for typ, tol in (("quasi", 0.0), ("pseudo", 1e-13), ("L2", 1e-12)):
        Q, _ = L2Projection(typ, mf, mcn).compute_transfer_1d()
        want = t["nonnested_Q_" + typ]
        assert np.asarray(Q).shape == want.shape == (17, 13)
        np.testing.assert_allclose(np.asarray(Q), want, rtol=tol, atol=tol)


def test_c1_operators_equal_reference_at_full_size():
    """BASELINE configs[0]: Mesh1D(1024) / Mesh1D(512): A, M, rhs and the three Q of the reference"""
    from learnmultigrid_b200.mesh.Mesh1D import Mesh1D
    from learnmultigrid_b200.L2_projection.L2Projection import L2Projection
    c1 = load_golden("c1_1d_1024.npz")
    m, A, M, rhs = problem_1d(1024)
    mc = Mesh1D(True, 512)
    mc.construct()
    assert np.array_equal(m.get_mesh(), c1["x_fine"]) and np.array_equal(mc.get_mesh(), c1["x_coarse"])
    Ar, Mr = coo_from(c1, "A").toarray(), coo_from(c1, "M").toarray()
    assert np.array_equal(np.asarray(A), Ar) and np.array_equal(np.asarray(M), Mr)
    assert np.array_equal(np.asarray(rhs), c1["rhs"])
    for typ, tol in (("quasi", 0.0), ("pseudo", 1e-13)):
        Q, _ = L2Projection(typ, m, mc).compute_transfer_1d()
        want = coo_from(c1, "Q_" + typ).toarray()
        assert np.array_equal(np.asarray(Q) != 0, want != 0) and np.count_nonzero(want) == 2561
        np.testing.assert_allclose(np.asarray(Q), want, rtol=tol, atol=tol)
    Q, _ = L2Projection("L2", m, mc).compute_transfer_1d()
    np.testing.assert_allclose(np.asarray(Q), c1["Q_L2_dense"], rtol=0, atol=1e-11)
