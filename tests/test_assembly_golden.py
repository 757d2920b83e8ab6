"""The vectorised P1 assembly (learnmultigrid_b200/assembly, learnmultigrid_b200/problems.py) against the reference's
own assembly run through oracle/refshim.py (tests/golden/assembly_2d.npz: Mesh2D(16), Mesh2D(256), Mesh2D(12);
'raw' = before the Dirichlet rows are replaced, thesis_structured_2d.py:380-414)."""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import coo_from, load_golden


class GoldenMesh:
    def __init__(self, p, conn):
        self.p, self.conn = np.asarray(p), np.asarray(conn)

    def get_np(self):
        return len(self.p)

    def get_points(self):
        return self.p

    def get_connections(self):
        return self.conn


def assemble(mesh):
    from learnmultigrid_b200.assembly.MassMatrix import MassMatrix
    from learnmultigrid_b200.assembly.StiffnessMatrix import StiffnessMatrix
    from learnmultigrid_b200.assembly.LoadVector import LoadVector
    from learnmultigrid_b200.assembly.LoadFunction import LoadFunction
    from learnmultigrid_b200.assembly.Quadrature import Quadrature2D
    from learnmultigrid_b200.assembly.ShapeFunction import FunctionTriangle, GradientTriangle
    q = Quadrature2D(3)
    A = StiffnessMatrix(mesh).compute_stiffness_2d(GradientTriangle(1), q, format="csr")
    M = MassMatrix(mesh).compute_mass_2d(FunctionTriangle(1), q, format="csr")
    rhs = LoadVector(mesh).compute_rhs_2d(LoadFunction(lambda pts: -1.0), FunctionTriangle(1), q)
    return A, M, rhs


@pytest.mark.parametrize("name", ["N4", "N16"])
def test_vectorised_assembly_matches_reference(name):
    d = load_golden("assembly_2d.npz")
    mesh = GoldenMesh(d[name + "_p"], d[name + "_conn"])
    A, M, rhs = assemble(mesh)
    Ar, Mr = sp.csr_matrix(coo_from(d, name + "_A_raw")), sp.csr_matrix(coo_from(d, name + "_M"))
    for got, want in ((A, Ar), (M, Mr)):
        want.sort_indices()
        assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)   # pattern exact
        np.testing.assert_allclose(got.data, want.data, rtol=1e-14)
    np.testing.assert_allclose(rhs, d[name + "_rhs_raw"], rtol=1e-14)


def test_structured_generator_matches_reference_operator():
    from learnmultigrid_b200 import problems as P
    d = load_golden("assembly_2d.npz")
    for N, name in ((4, "N4"), (16, "N16")):
        want = sp.csr_matrix(coo_from(d, name + "_A"))
        want.sort_indices()
        got = P.structured_laplacian_2d(N)
        assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
        np.testing.assert_allclose(got.data, want.data, rtol=1e-13)
        np.testing.assert_allclose(P.structured_rhs_2d(N), d[name + "_rhs"], rtol=1e-13)
