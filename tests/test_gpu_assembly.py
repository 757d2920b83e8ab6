"""Device P1 assembly (learnmultigrid_b200/assembly_device.py, csrc/assembly_kernels.cu) against the vectorised host
assembly (same arithmetic: bit for bit) and against the reference's own assembly (tests/golden/assembly_2d.npz:
1e-14 relative -- the reference's detJ comes out of a pivoted LU, np.linalg.det), then the whole device-resident path
mesh -> A, M -> NN-built transfers -> hierarchy -> V-cycle against the host-fed one."""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import coo_from, load_golden
from test_assembly_golden import GoldenMesh, assemble

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def asm():
    import torch
    assert torch.cuda.is_available()
    from learnmultigrid_b200.assembly_device import DeviceAssembler
    return DeviceAssembler()


def device_assemble(asm, mesh, coefficient=None):
    from learnmultigrid_b200.assembly.LoadFunction import LoadFunction
    from learnmultigrid_b200.assembly.Quadrature import Quadrature2D
    from learnmultigrid_b200.assembly.ShapeFunction import FunctionTriangle, GradientTriangle
    q = Quadrature2D(3)
    m = asm.mesh_to_device(mesh)
    A = asm.stiffness(*m, GradientTriangle(1), q, coefficient)
    M = asm.mass(*m, FunctionTriangle(1), q)
    rhs = asm.load(*m, LoadFunction(lambda pts: -1.0), FunctionTriangle(1), q)
    return A, M, rhs


@pytest.mark.parametrize("name", ["N4", "N16"])
def test_device_assembly_matches_reference_and_host(asm, name):
    d = load_golden("assembly_2d.npz")
    mesh = GoldenMesh(d[name + "_p"], d[name + "_conn"])
    A, M, rhs = device_assemble(asm, mesh)
    Ah, Mh, rhsh = assemble(mesh)
    Ar, Mr = sp.csr_matrix(coo_from(d, name + "_A_raw")), sp.csr_matrix(coo_from(d, name + "_M"))
    for got, host, ref in ((asm.S.download(A), Ah, Ar), (asm.S.download(M), Mh, Mr)):
        ref.sort_indices()
        assert np.array_equal(got.indptr, ref.indptr) and np.array_equal(got.indices, ref.indices)   # pattern exact
        assert np.array_equal(got.data, host.data)                                                   # same arithmetic
        np.testing.assert_allclose(got.data, ref.data, rtol=1e-14)
    np.testing.assert_allclose(rhs.cpu().numpy().reshape(-1, 1), d[name + "_rhs_raw"], rtol=1e-14)
    np.testing.assert_allclose(rhs.cpu().numpy().reshape(-1, 1), rhsh, rtol=1e-14)


def test_device_assembly_irregular_mesh_variable_coefficient_and_dirichlet(asm):
    from learnmultigrid_b200 import problems as P
    from learnmultigrid_b200.assembly.StiffnessMatrix import StiffnessMatrix
    from learnmultigrid_b200.assembly.Quadrature import Quadrature2D
    from learnmultigrid_b200.assembly.ShapeFunction import GradientTriangle
    pb = P.irregular_p1_2d(32, seed=9)
    mesh = pb["mesh"]
    A, M, rhs = device_assemble(asm, mesh)
    Ad = asm.dirichlet(A, pb["boundary"], rhs)
    got = asm.S.download(Ad)
    assert np.array_equal(got.indptr, pb["A"].indptr) and np.array_equal(got.indices, pb["A"].indices)
    # irregular triangles: the host path leaves the order in which duplicates are added to SciPy; the device adds them
    # in element order like the reference -> last-bit differences
    np.testing.assert_allclose(got.data, pb["A"].data, rtol=1e-13, atol=1e-16)
    np.testing.assert_allclose(asm.S.download(M).data, pb["M"].data, rtol=1e-14)
    np.testing.assert_allclose(rhs.cpu().numpy().reshape(-1, 1), pb["rhs"], rtol=1e-13, atol=1e-20)
    q = Quadrature2D(3)
    Ak = asm.S.download(asm.stiffness(*asm.mesh_to_device(mesh), GradientTriangle(1), q, P.variable_coefficient))
    Akh = StiffnessMatrix(mesh).compute_stiffness_2d(GradientTriangle(1), q, format="csr", coefficient=P.variable_coefficient)
    assert np.array_equal(Ak.indices, Akh.indices)
    np.testing.assert_allclose(Ak.data, Akh.data, rtol=1e-13, atol=1e-15 * abs(Akh.data).max())   # centroid means by torch
    assert StiffnessMatrix(mesh).compute_stiffness_2d(GradientTriangle(1), q, format="device").nnz == A.nnz


@pytest.mark.parametrize("case", ["i81", "i289", "r77"])
def test_device_assembly_matches_reference_on_irregular_meshes(asm, case):
    """the reference's own assembly on irregularly refined meshes (inputs and outputs in neural_2d_cases.npz)"""
    g = load_golden("neural_2d_cases.npz")
    mesh = GoldenMesh(g[case + "__mesh_p"], g[case + "__mesh_conn"])
    A, M, _ = device_assemble(asm, mesh)
    n = len(mesh.p)
    Ar = sp.csr_matrix((g[case + "__A_data"], (g[case + "__A_row"], g[case + "__A_col"])), shape=(n, n))
    Mr = sp.csr_matrix((g[case + "__l0_M_data"], (g[case + "__l0_M_row"], g[case + "__l0_M_col"])), shape=(n, n))
    for got, ref in ((asm.S.download(A), Ar), (asm.S.download(M), Mr)):
        ref.sort_indices()
        if case == "r77" and ref is Ar:
            # non-square right triangles: the hypotenuse couplings are 0 in exact arithmetic; the reference's LU-based
            # inverse leaves ~1e-17 there (and stores it), this arithmetic gives exact zeros (not stored)
            assert got.nnz < ref.nnz
            np.testing.assert_allclose(got.toarray(), ref.toarray(), rtol=1e-12, atol=1e-15 * abs(ref.data).max())
            continue
        assert np.array_equal(got.indptr, ref.indptr) and np.array_equal(got.indices, ref.indices)
        np.testing.assert_allclose(got.data, ref.data, rtol=1e-12, atol=1e-15 * abs(ref.data).max())


def test_device_resident_pipeline_equals_host_fed_pipeline(asm):
    """mesh -> device assembly -> NeuralBuilder (device M) -> DeviceHierarchy(device A, device Q) -> V-cycles, against
    the same pipeline fed through host SciPy matrices: identical iterates"""
    from learnmultigrid_b200 import problems as P
    from learnmultigrid_b200.engine import DeviceHierarchy
    from learnmultigrid_b200.neural2d import MassSurrogate, NeuralBuilder
    pb = P.irregular_p1_2d(32, seed=2)
    A, M, rhs = device_assemble(asm, pb["mesh"])
    A = asm.dirichlet(A, pb["boundary"], rhs)
    nb = NeuralBuilder()
    Qs = nb.define_hierarchy(M, MassSurrogate(), np.zeros(43), np.ones(43), 3)
    Qs_host = [nb.download(Q) for Q in Qs]
    h_dev = DeviceHierarchy(A, Qs, smoother="mcgs")
    h_host = DeviceHierarchy(asm.S.download(A), Qs_host, smoother="mcgs", colors=h_dev.colors)
    p = h_dev.make_params(nu_pre=2, nu_post=2)
    h_dev.set_rhs(rhs)
    h_host.set_rhs(rhs.cpu().numpy())
    for h in (h_dev, h_host):
        h.zero_x()
    for _ in range(3):
        h_dev.vcycle(p)
        h_host.vcycle(p)
        assert np.array_equal(h_dev.get_x(), h_host.get_x())
    assert h_dev.residual_norm() < 0.2 * np.linalg.norm(pb["rhs"])        # three V(2,2) cycles


def test_device_coupling_operator_on_triangle_intersections(asm):
    """B[f,c] = int phi_f phi_c between non-nested meshes by the per-pair CUDA kernel: against the NumPy model and the
    exact properties (partition of unity of both bases)"""
    from learnmultigrid_b200 import problems as P
    from learnmultigrid_b200.L2_projection.coupling2d import coupling_operator_2d, coupling_operator_2d_native
    from learnmultigrid_b200.mesh.Mesh2D import Mesh2D
    from learnmultigrid_b200.assembly.MassMatrix import MassMatrix
    from learnmultigrid_b200.assembly.Quadrature import Quadrature2D
    from learnmultigrid_b200.assembly.ShapeFunction import FunctionTriangle
    pb = P.irregular_p1_2d(32, seed=8)
    for coarse in (Mesh2D(13 * 13), Mesh2D(16 * 16)):          # unrelated and (for 16x16 parents) nested
        Bd = asm.S.download(coupling_operator_2d_native(pb["mesh"], coarse))                 # pairs binned on the device
        Bx = asm.S.download(coupling_operator_2d_native(pb["mesh"], coarse, pairs="host"))   # pairs from the NumPy binning
        assert np.array_equal(Bd.indptr, Bx.indptr) and np.array_equal(Bd.indices, Bx.indices)
        np.testing.assert_allclose(Bd.data, Bx.data, rtol=1e-13, atol=1e-18)
        Bh = coupling_operator_2d(pb["mesh"], coarse)
        assert np.array_equal(Bd.indptr, Bh.indptr) and np.array_equal(Bd.indices, Bh.indices)
        np.testing.assert_allclose(Bd.data, Bh.data, rtol=1e-12, atol=1e-17)
        Mc = MassMatrix(coarse).compute_mass_2d(FunctionTriangle(1), Quadrature2D(3), format="csr")
        np.testing.assert_allclose(np.asarray(Bd.sum(axis=1)).ravel(), np.asarray(pb["M"].sum(axis=1)).ravel(), rtol=1e-12)
        np.testing.assert_allclose(np.asarray(Bd.sum(axis=0)).ravel(), np.asarray(Mc.sum(axis=1)).ravel(), rtol=1e-12)
        # the API's 2D transfer operator is built from the device coupling operator
        from learnmultigrid_b200.L2_projection.L2Projection import L2Projection
        Qd = L2Projection("quasi", pb["mesh"], coarse).compute_transfer_2d()
        Qh = L2Projection("quasi", pb["mesh"], coarse).compute_transfer_2d(B=Bh)
        assert np.array_equal(Qd.indptr, Qh.indptr) and np.array_equal(Qd.indices, Qh.indices)
        np.testing.assert_allclose(Qd.data, Qh.data, rtol=1e-12, atol=1e-17)
