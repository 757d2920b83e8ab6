"""GPU tests of the device setup path: SpGEMM Galerkin products (bit-exact values AND sparsity patterns against
SciPy's csr_matrix(Q.T @ A @ Q), the reference's Multigrid.py:97-98), transposes, permutations, SELL build, and the
banded coarsest-level solver, all through the C ABI."""
import ctypes

import numpy as np
import pytest
import scipy.sparse as sp

from helpers import bilinear_P, coo_from, load_golden, poisson2d

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import torch
    from learnmultigrid_b200.setup_device import DeviceSetup
    assert torch.cuda.is_available()
    return DeviceSetup(torch, torch.device("cuda", 0))


def same_csr(dev_csr, want, S, values="exact"):
    from learnmultigrid_b200 import formats as F
    got = S.download(dev_csr)
    want = F.canonical_csr(want)
    assert got.shape == want.shape
    assert np.array_equal(got.indptr, want.indptr), "row pointers differ"
    assert np.array_equal(got.indices, want.indices), "sparsity pattern differs"
    if values == "exact":
        assert np.array_equal(got.data, want.data), "values differ (max %.3e)" % np.abs(got.data - want.data).max()
    else:
        np.testing.assert_allclose(got.data, want.data, rtol=1e-14)


@pytest.mark.parametrize("shape,density", [((1, 1), 1.0), ((50, 70), 0.1), ((1000, 300), 0.01), ((5000, 5000), 0.001)])
def test_transpose_matches_scipy_order(S, shape, density):
    from learnmultigrid_b200 import formats as F
    A = F.canonical_csr(sp.random(*shape, density=density, random_state=2, format="csr"))
    same_csr(S.transpose(S.upload(A)), F.transpose_csr(A), S)


def test_transpose_empty_and_empty_rows(S):
    from learnmultigrid_b200 import formats as F
    A = sp.csr_matrix((5, 7))
    T = S.transpose(S.upload(A))
    assert T.shape == (7, 5) and T.nnz == 0 and np.array_equal(T.indptr.cpu().numpy(), np.zeros(8, dtype=np.int32))
    A = sp.lil_matrix((6, 4))
    A[1, 3] = 2.0
    A[5, 0] = -1.0
    A[5, 3] = 4.0
    same_csr(S.transpose(S.upload(sp.csr_matrix(A))), F.transpose_csr(sp.csr_matrix(A)), S)


@pytest.mark.parametrize("n,k,m,da,db", [(40, 30, 50, 0.2, 0.2), (700, 900, 400, 0.01, 0.02), (3000, 3000, 3000, 0.003, 0.003)])
def test_spgemm_bit_exact_against_scipy(S, n, k, m, da, db):
    from learnmultigrid_b200 import formats as F
    A = F.canonical_csr(sp.random(n, k, density=da, random_state=1, format="csr"))
    B = F.canonical_csr(sp.random(k, m, density=db, random_state=2, format="csr"))
    want = sp.csr_matrix(A @ B)
    want.sort_indices()
    same_csr(S.spgemm(S.upload(A), S.upload(B)), want, S)


def test_spgemm_prunes_exact_zeros_like_scipy(S):
    """[[1,-1],[1,1]] @ [[1,0],[1,0]] has an exact cancellation: SciPy keeps 1 entry (SURVEY 7 hard parts)."""
    A = sp.csr_matrix(np.array([[1.0, -1.0], [1.0, 1.0]]))
    B = sp.csr_matrix(np.array([[1.0, 0.0], [1.0, 0.0]]))
    want = sp.csr_matrix(A @ B)
    assert want.nnz == 1
    same_csr(S.spgemm(S.upload(A), S.upload(B)), want, S)


def test_spgemm_long_rows_grow_the_hash_table(S):
    from learnmultigrid_b200 import formats as F
    A = F.canonical_csr(sp.random(64, 200, density=0.5, random_state=4, format="csr"))
    B = F.canonical_csr(sp.random(200, 600, density=0.3, random_state=5, format="csr"))
    want = sp.csr_matrix(A @ B)
    want.sort_indices()
    same_csr(S.spgemm(S.upload(A), S.upload(B)), want, S)


@pytest.mark.parametrize("transfer", ["linear", "quasi"])
def test_galerkin_hierarchy_patterns_and_values_bit_exact(S, transfer):
    """A_c = Q^T A Q on 4 levels of a 2D hierarchy: patterns and values identical to SciPy's
    csr_matrix(Q.T @ A @ Q) with A stored CSC as the reference does (Solver.py:18)."""
    from learnmultigrid_b200 import problems as P
    N = 64
    A = P.structured_laplacian_2d(N, P.variable_coefficient)
    Qs = P.structured_hierarchy_2d(N, 4, transfer=transfer)
    Ad = S.upload(A)
    Ah = sp.csc_matrix(A)
    for Q in Qs:
        Qd = S.upload(Q)
        Ad = S.galerkin(Ad, Qd, S.transpose(Qd))
        Ah = sp.csr_matrix(Q.T @ Ah @ Q)
        Ah.sort_indices()
        same_csr(Ad, Ah, S)


def test_galerkin_1d_c1_matches_scipy(S):
    from learnmultigrid_b200 import formats as F
    c1 = load_golden("c1_1d_1024.npz")
    A = coo_from(c1, "A")
    for Q in (coo_from(c1, "Q_quasi"), coo_from(c1, "Q_pseudo"), sp.csr_matrix(c1["Q_L2_dense"])):
        Qd = S.upload(Q)
        want = sp.csr_matrix(F.canonical_csr(Q).T @ sp.csc_matrix(A) @ F.canonical_csr(Q))
        want.sort_indices()
        same_csr(S.galerkin(S.upload(A), Qd, S.transpose(Qd)), want, S)


def test_permute_sell_dinv_match_host_formats(S):
    from learnmultigrid_b200 import formats as F
    A = poisson2d(20)
    colors, nc = F.greedy_colors(A)
    perm_h, cptr_h = F.color_permutation(colors)
    iperm_h = F.inverse_permutation(perm_h)
    perm, iperm, cptr = S.color_perm(colors)
    assert np.array_equal(perm.cpu().numpy(), perm_h) and np.array_equal(iperm.cpu().numpy(), iperm_h)
    assert np.array_equal(cptr, cptr_h)
    Ad = S.upload(A)
    Ap = S.permute(Ad, perm, iperm)
    want = F.permute_csr(A, perm_h, iperm_h)
    got = S.download(Ap)
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
    assert np.array_equal(got.data, want.data)
    sell = S.to_sell(Ap)
    sp_, sc_, sv_ = F.csr_to_sell(want)
    assert np.array_equal(sell.slice_ptr.cpu().numpy(), sp_)
    assert np.array_equal(sell.cols.cpu().numpy(), sc_) and np.array_equal(sell.vals.cpu().numpy(), sv_)
    assert np.array_equal(S.dinv(Ad, perm).cpu().numpy(), (1.0 / A.diagonal())[perm_h])
    # rectangular + ragged + column relabelling only
    Q = bilinear_P(20)
    Qp = S.permute(S.upload(Q), perm, None)
    wantq = F.permute_csr(F.canonical_csr(Q), perm_h, None)
    gq = S.download(Qp)
    assert np.array_equal(gq.indices, wantq.indices) and np.array_equal(gq.data, wantq.data)
    sq = S.to_sell(Qp)
    a, b, c = F.csr_to_sell(wantq)
    assert np.array_equal(sq.slice_ptr.cpu().numpy(), a) and np.array_equal(sq.cols.cpu().numpy(), b)
    assert np.array_equal(sq.vals.cpu().numpy(), c)


@pytest.mark.parametrize("smoother", ["jacobi", "mcgs", "lexgs"])
def test_device_setup_equals_host_setup(smoother):
    """the two setup paths hand identical data to the kernels: identical V-cycle iterates, bit for bit"""
    from learnmultigrid_b200 import problems as P
    from learnmultigrid_b200.engine import DeviceHierarchy
    N = 32
    A = P.structured_laplacian_2d(N)
    Qs = P.structured_hierarchy_2d(N, 3, transfer="quasi")
    rng = np.random.default_rng(0)
    b, x0 = rng.standard_normal(A.shape[0]), rng.standard_normal(A.shape[0])
    outs = []
    for setup in ("host", "device"):
        h = DeviceHierarchy(A, Qs, smoother=smoother, setup=setup)
        h.set_rhs(b)
        h.set_x(x0)
        p = h.make_params(nu_pre=2, nu_post=1, omega=0.7)
        h.vcycle(p)
        h.vcycle(p)
        outs.append(h.get_x())
        if setup == "device":
            Ah = sp.csr_matrix(Qs[0].T @ sp.csc_matrix(A) @ Qs[0])
            Ah.sort_indices()
            got = h.level_matrix(1)
            assert np.array_equal(got.indices, Ah.indices) and np.array_equal(got.data, Ah.data)
    assert np.array_equal(outs[0], outs[1])


def test_batched_inverse_and_gemm(S):
    from learnmultigrid_b200 import _lib
    torch, lib = S.torch, S.lib
    rng = np.random.default_rng(1)
    for m, batch in ((1, 3), (37, 5), (130, 4)):
        A = rng.standard_normal((batch, m, m)) + 0.3 * m * np.eye(m)
        dA = torch.from_numpy(A.copy()).to(S.dev)
        out = torch.zeros(batch * m * m, dtype=torch.float64, device=S.dev)
        work = torch.zeros(batch * m * 2 * m, dtype=torch.float64, device=S.dev)
        sing = torch.zeros(1, dtype=torch.int32, device=S.dev)
        _lib.check(lib.mg_dense_inverse_batched(m, batch, dA.data_ptr(), m * m, out.data_ptr(), m * m, work.data_ptr(),
                                                sing.data_ptr(), S.st()))
        inv = out.cpu().numpy().reshape(batch, m, m)
        assert int(sing.item()) == 0
        for b in range(batch):
            np.testing.assert_allclose(inv[b] @ A[b], np.eye(m), atol=1e-10)
        B = rng.standard_normal((batch, m, m))
        C = rng.standard_normal((batch, m, m))
        dB, dC = torch.from_numpy(B.copy()).to(S.dev), torch.from_numpy(C.copy()).to(S.dev)
        _lib.check(lib.mg_dense_gemm_batched(m, batch, dA.data_ptr(), m * m, dB.data_ptr(), m * m, dC.data_ptr(), m * m,
                                             -0.5, 2.0, S.st()))
        np.testing.assert_allclose(dC.cpu().numpy(), -0.5 * A @ B + 2.0 * C, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("N,transfer", [(40, "linear"), (24, "quasi")])
def test_bcr_solver_matches_direct_solve(S, N, transfer):
    """banded coarse operator (Galerkin of the 2D Laplacian) solved by block cyclic reduction vs SuperLU"""
    from learnmultigrid_b200 import problems as P, formats as F
    from learnmultigrid_b200.coarse import BcrCoarse, half_bandwidth
    from scipy.sparse.linalg import spsolve
    A = P.structured_laplacian_2d(2 * N)
    Q = P.structured_hierarchy_2d(2 * N, 2, transfer=transfer)[0]
    Ac = F.canonical_csr(sp.csr_matrix(Q.T @ sp.csc_matrix(A) @ Q))
    n = Ac.shape[0]
    bw = half_bandwidth(Ac.indptr, Ac.indices)
    Ad = S.upload(Ac)
    for min_block, tail in ((1, 1), (1, 5), (1, 16), (64, 1), (64, 16)):
        bcr = BcrCoarse(S.torch, S.dev, n, Ad.indptr, Ad.indices, Ad.values, bw, min_block=min_block, tail_blocks=tail)
        assert bcr.nb >= 4 and 1 <= bcr.tail_na <= max(tail, 1)
        rng = np.random.default_rng(3)
        rhs = rng.standard_normal(n)
        d_rhs = S.torch.from_numpy(rhs).to(S.dev)
        d_x = S.torch.zeros(n, dtype=S.torch.float64, device=S.dev)
        bcr.solve(S.torch, d_rhs, d_x)
        want = spsolve(sp.csc_matrix(Ac), rhs)
        np.testing.assert_allclose(d_x.cpu().numpy(), want, rtol=0, atol=1e-11 * np.linalg.norm(want))


def test_vcycle_with_bcr_coarsest_level(S):
    from learnmultigrid_b200 import problems as P
    from learnmultigrid_b200.engine import DeviceHierarchy
    from oracle.vcycle import OracleMultigrid
    N = 64
    A = P.structured_laplacian_2d(N)
    Qs = P.structured_hierarchy_2d(N, 2, transfer="linear")          # coarsest = 33^2 = 1089 unknowns
    rng = np.random.default_rng(0)
    b = rng.standard_normal(A.shape[0])
    h = DeviceHierarchy(A, Qs, smoother="mcgs", setup="device", dense_coarse_max=500)
    assert h.levels[-1].coarse_kind == 1
    o = OracleMultigrid(A, b.reshape(-1, 1), Qs, smoother="mcgs", colors=h.colors, hoist_setup=True)
    o.build_hierarchy(2)
    h.set_rhs(b)
    h.zero_x()
    p = h.make_params(nu_pre=1, nu_post=1)
    xo = np.zeros((A.shape[0], 1))
    for _ in range(3):
        h.vcycle(p)
        xo = o.v_cycle(o.matrix, xo, b.reshape(-1, 1), 1, 2)
        np.testing.assert_allclose(h.get_x(), xo, rtol=0, atol=1e-12 * np.linalg.norm(xo))
