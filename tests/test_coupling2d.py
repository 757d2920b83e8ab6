"""2D semi-geometric coupling operator on triangle-triangle intersections (L2_projection/coupling2d.py).  The reference's
2D path is an unfinished stub, so parity is UNPINNED; these are the exact properties the operator must have:

  * nested meshes: B = M_h P to rounding (the identity SURVEY 7.1 established in 1D against compute_b_1d);
  * any pair of meshes of the same domain: rowsum(B) = rowsum(M_fine) and colsum(B) = rowsum(M_coarse) (either basis
    is a partition of unity), the intersection areas add up to the domain's area, every fine element is covered;
  * the transfer operators built from it have unit row sums and make a two-level cycle converge on non-nested meshes
    (CPU oracle V-cycle)."""
import numpy as np
import pytest
import scipy.sparse as sp

from learnmultigrid_b200 import problems as P
from learnmultigrid_b200.L2_projection.coupling2d import coupling_operator_2d
from learnmultigrid_b200.L2_projection.L2Projection import L2Projection
from learnmultigrid_b200.L2_projection.Intersection import Intersection
from learnmultigrid_b200.mesh.Mesh2D import Mesh2D


def mass(mesh):
    from learnmultigrid_b200.assembly.MassMatrix import MassMatrix
    from learnmultigrid_b200.assembly.Quadrature import Quadrature2D
    from learnmultigrid_b200.assembly.ShapeFunction import FunctionTriangle
    return MassMatrix(mesh).compute_mass_2d(FunctionTriangle(1), Quadrature2D(3), format="csr")


def irregular(ne, seed):
    np.random.seed(seed)
    m = Mesh2D(ne)
    m.refine(regular=False)
    return m


@pytest.mark.parametrize("N", [8, 16])
def test_nested_meshes_give_mass_times_interpolation(N):
    fine, coarse = Mesh2D(N * N), Mesh2D((N // 2) ** 2)
    B = coupling_operator_2d(fine, coarse)
    want = sp.csr_matrix(mass(fine) @ P.linear_P_2d(N))
    want.sort_indices()
    assert np.array_equal(B.indptr, want.indptr) and np.array_equal(B.indices, want.indices)
    np.testing.assert_allclose(B.data, want.data, rtol=1e-12, atol=1e-16)
    Q = L2Projection("quasi", fine, coarse).compute_transfer_2d(B=B)       # (without B: integrated on the device)
    np.testing.assert_allclose(Q.toarray(), P.quasi_l2_Q_2d(N).toarray(), rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("fine,coarse", [("irr16", "s36"), ("s100", "irr8"), ("irr16", "irr8b"), ("s64", "s25")])
def test_partition_of_unity_on_non_nested_meshes(fine, coarse):
    meshes = {"irr16": lambda: irregular(64, 1), "irr8": lambda: irregular(16, 2), "irr8b": lambda: irregular(16, 7),
              "s36": lambda: Mesh2D(36), "s100": lambda: Mesh2D(100), "s64": lambda: Mesh2D(64), "s25": lambda: Mesh2D(25)}
    mf, mc = meshes[fine](), meshes[coarse]()
    B, pairs, area = coupling_operator_2d(mf, mc, return_pairs=True)
    Mf, Mc = mass(mf), mass(mc)
    np.testing.assert_allclose(np.asarray(B.sum(axis=1)).ravel(), np.asarray(Mf.sum(axis=1)).ravel(), rtol=1e-12)
    np.testing.assert_allclose(np.asarray(B.sum(axis=0)).ravel(), np.asarray(Mc.sum(axis=1)).ravel(), rtol=1e-12)
    np.testing.assert_allclose(area.sum(), 1.0, rtol=1e-13)
    assert set(pairs[:, 0]) == set(range(len(mf.get_connections())))          # every fine element is covered
    assert B.data.min() > 0
    inter = Intersection(mf, mc)
    inter.find_intersections2d(geometric=True)
    assert len(inter.get_intersections()) == len(pairs)
    for typ in ("quasi", "pseudo"):
        Q = L2Projection(typ, mf, mc).compute_transfer_2d(B=B)
        if typ == "quasi":
            np.testing.assert_allclose(np.asarray(Q.sum(axis=1)).ravel(), 1.0, rtol=1e-13)
        assert Q.shape == (mf.get_np(), mc.get_np())


def test_two_level_cycle_converges_on_non_nested_meshes():
    """fine = irregularly refined 17x17-node mesh, coarse = unrelated 7x7-node structured mesh (non-nested): the
    quasi-L2 operator from the intersections gives a convergent two-level cycle (oracle, the reference's smoother)"""
    from oracle.vcycle import OracleMultigrid
    pb = P.irregular_p1_2d(16, seed=4)
    coarse = Mesh2D(36)
    Q = L2Projection("quasi", pb["mesh"], coarse).compute_transfer_2d(B=coupling_operator_2d(pb["mesh"], coarse))
    # coarse boundary nodes interpolate into the Dirichlet rows only: drop their columns so that Q^T A Q is regular
    pc = np.asarray(coarse.get_points())
    interior_c = np.flatnonzero((pc[:, 0] > 1e-12) & (pc[:, 0] < 1 - 1e-12) & (pc[:, 1] > 1e-12) & (pc[:, 1] < 1 - 1e-12))
    Q = sp.csr_matrix(Q[:, interior_c])
    o = OracleMultigrid(pb["A"], pb["rhs"], [Q], smoother="gs", hoist_setup=True)
    o.solve(levels=2, smooth_steps=3, error=1e-9, max_iterations=60)
    h = o.track_res.ravel()
    assert len(h) < 40 and h[-1] <= 1e-9
    assert np.all(h[2:] < 0.8 * h[1:-1])


@pytest.mark.parametrize("fine,coarse", [("irr16", "s36"), ("s64", "s16")])
def test_per_pair_c_code_matches_numpy_model(fine, coarse):
    """csrc/assembly_kernels.cu::coupling_pair run serially on the host (the CUDA kernel runs the same function)"""
    from learnmultigrid_b200.L2_projection.coupling2d import coupling_operator_2d_native
    meshes = {"irr16": lambda: irregular(64, 1), "s36": lambda: Mesh2D(36), "s64": lambda: Mesh2D(64), "s16": lambda: Mesh2D(16)}
    mf, mc = meshes[fine](), meshes[coarse]()
    B = coupling_operator_2d(mf, mc)
    Bn = coupling_operator_2d_native(mf, mc, where="host")
    assert np.array_equal(B.indptr, Bn.indptr) and np.array_equal(B.indices, Bn.indices)
    np.testing.assert_allclose(Bn.data, B.data, rtol=1e-12, atol=1e-17)
