"""problems_device.py: the structured benchmark inputs generated in HBM equal the host generators of problems.py, and the
API takes operators that already live on the device (DevCSR) without a host detour."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _host(M):
    import scipy.sparse as sp
    return sp.csr_matrix((M.values.cpu().numpy(), M.indices.cpu().numpy(), M.indptr.cpu().numpy()), shape=M.shape)


@pytest.mark.parametrize("N,Ny", [(8, None), (64, None), (32, 96)])
def test_device_generators_equal_host_generators(N, Ny):
    import torch
    from learnmultigrid_b200 import problems as P, problems_device as PD
    torch.cuda.set_device(0)
    A_h = P.structured_laplacian_2d(N, Ny=Ny)
    A_d = _host(PD.structured_laplacian_2d(N, Ny=Ny))
    assert np.array_equal(A_d.indptr, A_h.indptr) and np.array_equal(A_d.indices, A_h.indices)
    assert np.array_equal(A_d.data, A_h.data)
    V_h = P.structured_laplacian_2d(N, P.variable_coefficient, Ny=Ny)
    V_d = _host(PD.structured_laplacian_2d(N, PD.variable_coefficient, Ny=Ny))
    assert np.array_equal(V_d.indptr, V_h.indptr) and np.array_equal(V_d.indices, V_h.indices)
    np.testing.assert_allclose(V_d.data, V_h.data, rtol=1e-14, atol=1e-15)          # the device's sin(): last bits
    assert np.array_equal(PD.structured_rhs_2d(N, Ny=Ny).cpu().numpy(), P.structured_rhs_2d(N, Ny=Ny).ravel())
    if Ny is None:                                   # the symmetrically eliminated operator of the PCG configuration
        S_h = P.symmetric_dirichlet(P.structured_laplacian_2d(N), P.boundary_nodes_2d(N))
        S_d = _host(PD.structured_laplacian_2d(N, symmetric=True))
        assert np.array_equal(S_d.indptr, S_h.indptr) and np.array_equal(S_d.indices, S_h.indices)
        assert np.array_equal(S_d.data, S_h.data)
        Sv_h = P.symmetric_dirichlet(V_h, P.boundary_nodes_2d(N))
        Sv_d = _host(PD.structured_laplacian_2d(N, PD.variable_coefficient, symmetric=True))
        assert np.array_equal(Sv_d.indptr, Sv_h.indptr) and np.array_equal(Sv_d.indices, Sv_h.indices)
        np.testing.assert_allclose(Sv_d.data, Sv_h.data, rtol=1e-14, atol=1e-15)
    ny = N if Ny is None else Ny
    for q_d, q_h in zip(PD.structured_hierarchy_2d(N, 3, Ny=Ny), P.structured_hierarchy_2d(N, 3, Ny=Ny)):
        q_d = _host(q_d)
        assert q_d.shape == q_h.shape
        assert np.array_equal(q_d.indptr, q_h.indptr) and np.array_equal(q_d.indices, q_h.indices)
        assert np.array_equal(q_d.data, q_h.data)


def test_api_takes_device_operators():
    """SemiGeometricMG(A_dev, rhs, [Q_dev ...]): same hierarchy, same history, same solution as with host inputs"""
    import torch
    from learnmultigrid_b200 import problems as P, problems_device as PD
    from learnmultigrid_b200.solvers.Multigrid import SemiGeometricMG
    torch.cuda.set_device(0)
    N, L = 128, 4
    rhs = P.structured_rhs_2d(N)
    kw = dict(levels=L, smoother="GaussSeidel", smooth_steps=1, error=1e-9, max_iterations=30)
    host = SemiGeometricMG(P.structured_laplacian_2d(N), rhs, P.structured_hierarchy_2d(N, L))
    host.solve(**kw)
    dev = SemiGeometricMG(PD.structured_laplacian_2d(N), rhs, PD.structured_hierarchy_2d(N, L))
    dev.solve(**kw)
    assert dev.get_iterations() == host.get_iterations() < 30
    assert np.array_equal(dev.track_res, host.track_res)
    assert np.array_equal(dev.get_solution(), host.get_solution())
    for l in range(L):
        a, b = dev.get_hierarchy().level_matrix(l), host.get_hierarchy().level_matrix(l)
        assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices) and np.array_equal(a.data, b.data)
