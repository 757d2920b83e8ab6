"""GPU parity tests: every CUDA path, called through the C ABI (ctypes on libmgb200.so) or through the
reference-facing Python API, against the CPU oracle on identical matrices and transfer operators.

Bars (BASELINE.json north_star): per-kernel results of SpMV / residual / Jacobi / Gauss-Seidel / transfers are
compared BIT FOR BIT (the kernels add in the oracle's order without FMA); V-cycle iterates to <= 1e-12 relative
(the coarsest direct solve differs from SuperLU in the last bits); residual histories to the same bar and
identical iteration counts.
"""
import ctypes

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import kernels as K
from oracle.vcycle import OracleMultigrid
from helpers import (assert_history_close, bilinear_P, coo_from, load_golden, poisson2d)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    from learnmultigrid_b200 import _lib
    assert torch.cuda.is_available()
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    cc = ctypes.c_int()
    sm = ctypes.c_int()
    mem = ctypes.c_int64()
    _lib.check(lib.mg_device_info(ctypes.byref(sm), ctypes.byref(mem), ctypes.byref(cc)))
    return {"torch": torch, "lib": lib, "L": _lib, "dev": dev, "sm": sm.value, "cc": cc.value}


def up(env, a):
    return env["torch"].from_numpy(np.ascontiguousarray(a)).to(env["dev"])


def stream(env):
    return env["L"].stream_handle(env["torch"])


def random_system(n, density, seed, diag=4.0):
    from learnmultigrid_b200 import formats as F
    rng = np.random.default_rng(seed)
    A = sp.random(n, n, density=density, random_state=seed, format="csr") + diag * sp.eye(n)
    A = F.canonical_csr(A)
    return A, rng.standard_normal(n), rng.standard_normal(n)


def test_device_is_blackwell(env):
    assert env["cc"] >= 100 and env["sm"] >= 100


@pytest.mark.parametrize("n,density", [(1, 1.0), (31, 0.2), (32, 0.2), (33, 0.2), (1000, 0.01), (4097, 0.002)])
def test_csr_kernels_bit_exact(env, n, density):
    L, lib = env["L"], env["lib"]
    A, x, b = random_system(n, density, 11 + n)
    ip, ix, va = up(env, A.indptr), up(env, A.indices), up(env, A.data)
    dx, db = up(env, x), up(env, b)
    out = env["torch"].empty_like(dx)
    L.check(lib.mg_spmv_csr(n, ip.data_ptr(), ix.data_ptr(), va.data_ptr(), dx.data_ptr(), out.data_ptr(), stream(env)))
    assert np.array_equal(out.cpu().numpy(), K.spmv(A, x))
    L.check(lib.mg_residual_csr(n, ip.data_ptr(), ix.data_ptr(), va.data_ptr(), dx.data_ptr(), db.data_ptr(),
                                out.data_ptr(), stream(env)))
    assert np.array_equal(out.cpu().numpy(), K.residual(A, x, b))
    dinv = 1.0 / A.diagonal()
    dd = up(env, dinv)
    for omega in (1.0, 2.0 / 3.0):
        L.check(lib.mg_jacobi_sweep_csr(n, ip.data_ptr(), ix.data_ptr(), va.data_ptr(), dd.data_ptr(), dx.data_ptr(),
                                        db.data_ptr(), out.data_ptr(), omega, stream(env)))
        assert np.array_equal(out.cpu().numpy(), K.jacobi(A, x, b, dinv, omega, 1))
    du = up(env, b.copy())
    L.check(lib.mg_prolong_correct_csr(n, ip.data_ptr(), ix.data_ptr(), va.data_ptr(), dx.data_ptr(), du.data_ptr(),
                                       stream(env)))
    assert np.array_equal(du.cpu().numpy(), K.prolong_correct(A, x, b))


def test_jacobi_rejects_aliasing(env):
    L, lib = env["L"], env["lib"]
    A, x, b = random_system(10, 0.3, 1)
    ip, ix, va, dx, db = (up(env, a) for a in (A.indptr, A.indices, A.data, x, b))
    rc = lib.mg_jacobi_sweep_csr(10, ip.data_ptr(), ix.data_ptr(), va.data_ptr(), dx.data_ptr(), dx.data_ptr(),
                                 db.data_ptr(), dx.data_ptr(), 1.0, stream(env))
    assert rc == -1 and b"alias" in lib.mg_last_error()


@pytest.mark.parametrize("n,density", [(40, 0.2), (1000, 0.01), (5000, 0.002)])
def test_gauss_seidel_kernels_bit_exact(env, n, density):
    from learnmultigrid_b200 import formats as F
    L, lib = env["L"], env["lib"]
    A, x, b = random_system(n, density, 5 + n)
    A = F.canonical_csr(A + A.T)                      # symmetric pattern, like the FE operators
    ip, ix, va, db = (up(env, a) for a in (A.indptr, A.indices, A.data, b))
    # multicolour
    colors, nc = F.greedy_colors(A)
    perm, cptr = F.color_permutation(colors)
    rows = [perm[cptr[c]:cptr[c + 1]] for c in range(nc)]
    want = x.copy()
    K.gauss_seidel_multicolor(A, want, b, rows, iterations=2)
    dx = up(env, x.copy())
    cp = (ctypes.c_int64 * (nc + 1))(*[int(v) for v in cptr])
    dperm = up(env, perm)
    for _ in range(2):
        L.check(lib.mg_gs_multicolor_sweep_csr(n, ip.data_ptr(), ix.data_ptr(), va.data_ptr(), dx.data_ptr(),
                                               db.data_ptr(), cp, dperm.data_ptr(), nc, stream(env)))
    assert np.array_equal(dx.cpu().numpy(), want)
    # exact lexicographic (PyAMG order) through dependency levels
    lp, lr = F.lex_levels(A)
    want = x.copy()
    K.gauss_seidel(A, want, b, iterations=3)
    dx = up(env, x.copy())
    dlp, dlr = up(env, lp), up(env, lr)
    L.check(lib.mg_gs_lex_sweep_csr(n, ip.data_ptr(), ix.data_ptr(), va.data_ptr(), dx.data_ptr(), db.data_ptr(),
                                    dlp.data_ptr(), dlr.data_ptr(), len(lp) - 1, 3, stream(env)))
    assert np.array_equal(dx.cpu().numpy(), want)


def test_lex_gs_on_1d_chain_and_zero_diagonal_rows(env):
    """tridiagonal = 1025 dependency levels of one row; rows with a zero diagonal are skipped like PyAMG"""
    from learnmultigrid_b200 import formats as F
    L, lib = env["L"], env["lib"]
    c1 = load_golden("c1_1d_1024.npz")
    A = F.canonical_csr(coo_from(c1, "A"))
    A = A.tolil()
    A[7, 7] = 0.0
    A = F.canonical_csr(sp.csr_matrix(A))
    b = c1["rhs"].ravel()
    x = np.linspace(0, 1, 1025)
    want = x.copy()
    K.gauss_seidel(A, want, b, iterations=2)
    lp, lr = F.lex_levels(A)
    ip, ix, va, db, dx, dlp, dlr = (up(env, a) for a in (A.indptr, A.indices, A.data, b, x.copy(), lp, lr))
    L.check(lib.mg_gs_lex_sweep_csr(1025, ip.data_ptr(), ix.data_ptr(), va.data_ptr(), dx.data_ptr(), db.data_ptr(),
                                    dlp.data_ptr(), dlr.data_ptr(), len(lp) - 1, 2, stream(env)))
    assert np.array_equal(dx.cpu().numpy(), want)


def sell_struct(env, A):
    from learnmultigrid_b200.engine import DeviceSell
    return DeviceSell(env["torch"], A, env["dev"])


@pytest.fixture(params=["plain", "tma"])
def sell_path(request, env):
    """run the SELL tests through the plain kernel and through the bulk-async (TMA) staged kernel"""
    old = env["lib"].mg_set_tma_min_rows(1 if request.param == "tma" else 0)
    yield request.param
    env["lib"].mg_set_tma_min_rows(old)


@pytest.mark.parametrize("n,per_row", [(5000, 2), (70000, 5), (33000, 9), (4100, 40), (600, 700)])
def test_sell_tma_all_modes_and_stage_geometries(env, n, per_row):
    """max slice length 2 / 5 / ~10 / ~40 / > 500 exercises the G = 4, 2, 1 stage geometries and the fallback for
    very long rows; row ranges that start and end inside a slice exercise the colour-block masking"""
    from learnmultigrid_b200 import formats as F
    L, lib, torch = env["L"], env["lib"], env["torch"]
    rng = np.random.default_rng(n)
    rows = np.repeat(np.arange(n), per_row)
    cols = (rows + rng.integers(-3 * per_row, 3 * per_row + 1, size=rows.size)) % n
    A = F.canonical_csr(sp.csr_matrix((rng.standard_normal(rows.size), (rows, cols)), shape=(n, n)) + 4.0 * sp.eye(n))
    S = sell_struct(env, A)
    x, b = rng.standard_normal(n), rng.standard_normal(n)
    dx, db = up(env, x), up(env, b)
    old = lib.mg_set_tma_min_rows(1)
    try:
        out = torch.empty(n, dtype=torch.float64, device=env["dev"])
        L.check(lib.mg_sell_spmv(ctypes.byref(S.struct), dx.data_ptr(), out.data_ptr(), stream(env)))
        assert np.array_equal(out.cpu().numpy(), K.spmv(A, x))
        L.check(lib.mg_sell_residual(ctypes.byref(S.struct), dx.data_ptr(), db.data_ptr(), out.data_ptr(), stream(env)))
        r = K.residual(A, x, b)
        assert np.array_equal(out.cpu().numpy(), r)
        ws = torch.zeros(int(lib.mg_norm_workspace_size(n)) + 8, dtype=torch.float64, device=env["dev"])
        nrm = torch.zeros(1, dtype=torch.float64, device=env["dev"])
        L.check(lib.mg_sell_residual_norm2(ctypes.byref(S.struct), dx.data_ptr(), db.data_ptr(), ws.data_ptr(),
                                           nrm.data_ptr(), stream(env)))
        np.testing.assert_allclose(np.sqrt(nrm.item()), np.linalg.norm(r), rtol=1e-13)
        dinv = 1.0 / A.diagonal()
        dd = up(env, dinv)
        L.check(lib.mg_sell_jacobi(ctypes.byref(S.struct), dd.data_ptr(), dx.data_ptr(), db.data_ptr(), out.data_ptr(),
                                   0.8, stream(env)))
        assert np.array_equal(out.cpu().numpy(), K.jacobi(A, x, b, dinv, 0.8, 1))
        # Gauss-Seidel on an arbitrary row range [r0, r1) (Jacobi-style inside the range, like one colour block)
        r0, r1 = 37, n - 45
        want = x.copy()
        Ao = A.copy()
        acc = b[r0:r1] - (A[r0:r1] @ x - A.diagonal()[r0:r1] * x[r0:r1])
        dxs = up(env, x.copy())
        L.check(lib.mg_sell_gs_rows(ctypes.byref(S.struct), dxs.data_ptr(), db.data_ptr(), r0, r1, stream(env)))
        got = dxs.cpu().numpy()
        assert np.array_equal(got[:r0], x[:r0]) and np.array_equal(got[r1:], x[r1:])
        # rows inside the range may read rows of the same range that were already updated only if coupled; compare
        # against the oracle on the decoupled reference: one row at a time with the ORIGINAL x (what a valid colour sees)
        ref = x.copy()
        for i in (r0, r0 + 1, (r0 + r1) // 2, r1 - 1):
            xi = x.copy()
            K.gauss_seidel_multicolor(A, xi, b, [np.array([i], dtype=np.int32)])
            ref[i] = xi[i]
            if not np.any((A[i].indices >= r0) & (A[i].indices < r1) & (A[i].indices != i)):
                assert got[i] == ref[i]
        du = up(env, b.copy())
        L.check(lib.mg_sell_prolong_correct(ctypes.byref(S.struct), dx.data_ptr(), du.data_ptr(), du.data_ptr(), stream(env)))
        assert np.array_equal(du.cpu().numpy(), K.prolong_correct(A, x, b))
    finally:
        lib.mg_set_tma_min_rows(old)


@pytest.mark.parametrize("shape,density", [((1, 1), 1.0), ((33, 20), 0.3), ((1000, 1000), 0.01), ((4100, 900), 0.004)])
def test_sell_kernels_bit_exact(env, sell_path, shape, density):
    from learnmultigrid_b200 import formats as F
    L, lib, torch = env["L"], env["lib"], env["torch"]
    rng = np.random.default_rng(shape[0])
    A = F.canonical_csr(sp.random(*shape, density=density, random_state=3, format="csr"))
    S = sell_struct(env, A)
    x = rng.standard_normal(shape[1])
    u = rng.standard_normal(shape[0])
    dx, du = up(env, x), up(env, u)
    out = torch.empty(shape[0], dtype=torch.float64, device=env["dev"])
    L.check(lib.mg_sell_spmv(ctypes.byref(S.struct), dx.data_ptr(), out.data_ptr(), stream(env)))
    assert np.array_equal(out.cpu().numpy(), K.spmv(A, x))
    L.check(lib.mg_sell_prolong_correct(ctypes.byref(S.struct), dx.data_ptr(), du.data_ptr(), out.data_ptr(), stream(env)))
    assert np.array_equal(out.cpu().numpy(), K.prolong_correct(A, x, u))
    L.check(lib.mg_sell_prolong_correct(ctypes.byref(S.struct), dx.data_ptr(), du.data_ptr(), du.data_ptr(), stream(env)))
    assert np.array_equal(du.cpu().numpy(), K.prolong_correct(A, x, u))            # in place


@pytest.mark.parametrize("n,density", [(1, 1.0), (64, 0.1), (1500, 0.01)])
def test_sell_square_kernels_bit_exact(env, sell_path, n, density):
    from learnmultigrid_b200 import formats as F
    L, lib, torch = env["L"], env["lib"], env["torch"]
    A, x, b = random_system(n, density, 100 + n)
    A = F.canonical_csr(A + A.T)
    S = sell_struct(env, A)
    dx, db = up(env, x), up(env, b)
    out = torch.empty(n, dtype=torch.float64, device=env["dev"])
    L.check(lib.mg_sell_residual(ctypes.byref(S.struct), dx.data_ptr(), db.data_ptr(), out.data_ptr(), stream(env)))
    r = K.residual(A, x, b)
    assert np.array_equal(out.cpu().numpy(), r)
    ws = torch.zeros(int(lib.mg_norm_workspace_size(n)) + 8, dtype=torch.float64, device=env["dev"])
    nrm = torch.zeros(1, dtype=torch.float64, device=env["dev"])
    L.check(lib.mg_sell_residual_norm2(ctypes.byref(S.struct), dx.data_ptr(), db.data_ptr(), ws.data_ptr(),
                                       nrm.data_ptr(), stream(env)))
    np.testing.assert_allclose(np.sqrt(nrm.item()), np.linalg.norm(r), rtol=1e-14)
    dinv = 1.0 / A.diagonal()
    dd = up(env, dinv)
    L.check(lib.mg_sell_jacobi(ctypes.byref(S.struct), dd.data_ptr(), dx.data_ptr(), db.data_ptr(), out.data_ptr(),
                               2.0 / 3.0, stream(env)))
    assert np.array_equal(out.cpu().numpy(), K.jacobi(A, x, b, dinv, 2.0 / 3.0, 1))
    # Gauss-Seidel on colour-blocked ordering == multicolour GS in natural ordering, bit for bit
    colors, nc = F.greedy_colors(A)
    perm, cptr = F.color_permutation(colors)
    iperm = F.inverse_permutation(perm)
    Sp = sell_struct(env, F.permute_csr(A, perm, iperm))
    want = x.copy()
    K.gauss_seidel_multicolor(A, want, b, [perm[cptr[c]:cptr[c + 1]] for c in range(nc)])
    dxp, dbp = up(env, x[perm]), up(env, b[perm])
    for c in range(nc):
        L.check(lib.mg_sell_gs_rows(ctypes.byref(Sp.struct), dxp.data_ptr(), dbp.data_ptr(), int(cptr[c]),
                                    int(cptr[c + 1]), stream(env)))
    got = np.empty(n)
    got[perm] = dxp.cpu().numpy()
    assert np.array_equal(got, want)


def test_vector_kernels(env):
    L, lib, torch = env["L"], env["lib"], env["torch"]
    rng = np.random.default_rng(0)
    for n in (1, 255, 256, 257, 100003):
        x, y = rng.standard_normal(n), rng.standard_normal(n)
        dx, dy = up(env, x), up(env, y)
        ws = torch.zeros(8192, dtype=torch.float64, device=env["dev"])
        out = torch.zeros(1, dtype=torch.float64, device=env["dev"])
        L.check(lib.mg_dot(n, dx.data_ptr(), dy.data_ptr(), ws.data_ptr(), out.data_ptr(), stream(env)))
        np.testing.assert_allclose(out.item(), float(x @ y), rtol=1e-12, atol=1e-12)
        o = torch.empty_like(dx)
        L.check(lib.mg_axpby(n, 0.3, dx.data_ptr(), -1.7, dy.data_ptr(), o.data_ptr(), stream(env)))
        assert np.array_equal(o.cpu().numpy(), 0.3 * x + (-1.7) * y)
        perm = rng.permutation(n).astype(np.int32)
        dp = up(env, perm)
        L.check(lib.mg_gather(n, dp.data_ptr(), dx.data_ptr(), o.data_ptr(), stream(env)))
        assert np.array_equal(o.cpu().numpy(), x[perm])
        o2 = torch.empty_like(dx)
        L.check(lib.mg_scatter(n, dp.data_ptr(), o.data_ptr(), o2.data_ptr(), stream(env)))
        assert np.array_equal(o2.cpu().numpy(), x)


@pytest.mark.parametrize("n", [1, 3, 257, 700])
def test_dense_inverse_and_gemv(env, n):
    L, lib, torch = env["L"], env["lib"], env["torch"]
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n)) + n * np.eye(n) * 0.05
    if n == 3:
        A = np.array([[0.0, 2.0, 1.0], [1.0, 0.0, 3.0], [4.0, 1.0, 0.0]])      # needs pivoting
    dA = up(env, A.reshape(-1).copy())
    inv = torch.empty(n * n, dtype=torch.float64, device=env["dev"])
    work = torch.empty(int(lib.mg_dense_inverse_workspace(n)), dtype=torch.uint8, device=env["dev"])
    L.check(lib.mg_dense_inverse(n, dA.data_ptr(), inv.data_ptr(), work.data_ptr(), stream(env)))
    got = inv.cpu().numpy().reshape(n, n)
    np.testing.assert_allclose(got @ A, np.eye(n), atol=1e-9)
    b = rng.standard_normal(n)
    db = up(env, b)
    y = torch.empty(n, dtype=torch.float64, device=env["dev"])
    L.check(lib.mg_dense_gemv(n, n, inv.data_ptr(), db.data_ptr(), y.data_ptr(), stream(env)))
    np.testing.assert_allclose(y.cpu().numpy(), np.linalg.solve(A, b), rtol=1e-8, atol=1e-10)


def test_dense_inverse_reports_singular(env):
    L, lib, torch = env["L"], env["lib"], env["torch"]
    A = np.ones((4, 4))
    dA = up(env, A.reshape(-1).copy())
    inv = torch.empty(16, dtype=torch.float64, device=env["dev"])
    work = torch.empty(int(lib.mg_dense_inverse_workspace(4)), dtype=torch.uint8, device=env["dev"])
    rc = lib.mg_dense_inverse(4, dA.data_ptr(), inv.data_ptr(), work.data_ptr(), stream(env))
    assert rc == -5


# ------------------------------------------------------------------------------------------------------------
# hierarchy level: one V-cycle against the oracle
def _cycle_case(env, A, Qs, smoother, omega, nu, use_graph, seed=0):
    from learnmultigrid_b200.engine import DeviceHierarchy
    rng = np.random.default_rng(seed)
    n = A.shape[0]
    b = rng.standard_normal(n)
    x0 = rng.standard_normal(n)
    h = DeviceHierarchy(A, Qs, smoother=smoother, setup="host")
    o = OracleMultigrid(A, b.reshape(-1, 1), Qs,
                        smoother={"jacobi": "jacobi", "mcgs": "mcgs", "lexgs": "gs"}[smoother], omega=omega,
                        colors=h.colors, hoist_setup=True)
    L = len(Qs) + 1
    o.build_hierarchy(L)
    h.set_rhs(b)
    h.set_x(x0)
    params = h.make_params(nu_pre=nu, nu_post=nu, omega=omega)
    xo = x0.reshape(-1, 1).copy()
    for _ in range(2):
        h.vcycle(params, use_graph=use_graph)
        xo = o.v_cycle(o.matrix, xo, b.reshape(-1, 1), nu, L)
        got = h.get_x()
        np.testing.assert_allclose(got, xo, rtol=0, atol=1e-12 * np.linalg.norm(xo))
    # outer residual norm
    np.testing.assert_allclose(h.residual_norm(), np.linalg.norm(b.reshape(-1, 1) - o.matrix @ xo), rtol=1e-9)
    return h


@pytest.mark.parametrize("smoother,omega", [("jacobi", 2.0 / 3.0), ("jacobi", 1.0), ("mcgs", 1.0), ("lexgs", 1.0)])
@pytest.mark.parametrize("nu", [1, 2, 3])
def test_vcycle_2d_three_levels(env, sell_path, smoother, omega, nu):
    N = 32
    A = poisson2d(N)
    Qs = [bilinear_P(N), bilinear_P(N // 2)]
    _cycle_case(env, A, Qs, smoother, omega, nu, use_graph=(nu != 2))


@pytest.mark.parametrize("smoother", ["jacobi", "mcgs"])
def test_vcycle_2d_quasi_l2_transfers(env, sell_path, smoother):
    from learnmultigrid_b200 import problems as P
    N = 32
    A = P.structured_laplacian_2d(N)
    Qs = P.structured_hierarchy_2d(N, 4, transfer="quasi")
    h = _cycle_case(env, A, Qs, smoother, 2.0 / 3.0, 1, True)
    assert h.levels[1].nnz_A / h.levels[1].n > 12          # 19-point Galerkin stencil on level 1


def test_vcycle_variable_coefficient(env):
    from learnmultigrid_b200 import problems as P
    N = 32
    A = P.structured_laplacian_2d(N, P.variable_coefficient)
    Qs = P.structured_hierarchy_2d(N, 3, transfer="linear")
    _cycle_case(env, A, Qs, "mcgs", 1.0, 2, True)


def test_vcycle_1d_c1(env):
    from learnmultigrid_b200 import formats as F
    c1 = load_golden("c1_1d_1024.npz")
    A = coo_from(c1, "A")
    Qs = [coo_from(c1, "Q_quasi"), F.geometric_interpolator_csr(513)]
    for sm in ("jacobi", "mcgs", "lexgs"):
        _cycle_case(env, A, Qs, sm, 2.0 / 3.0, 1, True)


# ------------------------------------------------------------------------------------------------------------
# API level: the reference-facing classes against the reference's own known answers
@pytest.mark.parametrize("typ,levels,steps,its", [("quasi", 3, 1, 11), ("quasi", 2, 3, 5), ("quasi", 3, 3, 5),
                                                  ("pseudo", 3, 3, 5), ("L2", 2, 3, 6), ("quasi", 5, 3, 6)])
def test_api_reference_histories_lexicographic(env, typ, levels, steps, its):
    """SemiGeometricMG(A, rhs, Q).solve(levels, "GaussSeidel", steps, error=1e-10, max_iterations=40) with the
    exact index-order smoother reproduces the REFERENCE'S OWN residual history and iteration count."""
    from learnmultigrid_b200.solvers.Multigrid import SemiGeometricMG
    c1 = load_golden("c1_1d_1024.npz")
    A = coo_from(c1, "A").toarray()
    Q = c1["Q_L2_dense"] if typ == "L2" else coo_from(c1, "Q_" + typ)
    mg = SemiGeometricMG(A, c1["rhs"], Q)
    mg.solve(levels=levels, smoother="GaussSeidel", smooth_steps=steps, error=1e-10, max_iterations=40,
             gs_order="lexicographic")
    key = "ref_sgmg_%s_L%d_s%d" % (typ, levels, steps)
    assert mg.get_iterations() == its == int(c1[key + "_its"])
    assert mg.track_res.shape == (its, 1) and mg.track_res[0, 0] == np.sqrt(1025.0)
    assert_history_close(mg.track_res, c1[key + "_track"], A, c1[key + "_x"])
    np.testing.assert_allclose(mg.get_solution(), c1[key + "_x"], rtol=0, atol=1e-12 * np.linalg.norm(c1[key + "_x"]))


def test_api_geometric_mg_and_fd(env):
    from learnmultigrid_b200.solvers.Multigrid import GeometricMG, SemiGeometricMG
    c1 = load_golden("c1_1d_1024.npz")
    A = coo_from(c1, "A")
    g = GeometricMG(A, c1["rhs"])
    g.solve(levels=3, smoother="GaussSeidel", smooth_steps=3, error=1e-10, max_iterations=40, gs_order="lexicographic")
    assert_history_close(g.track_res, c1["ref_gmg_L3_s3_track"], A, c1["ref_gmg_L3_s3_x"])
    Afd = coo_from(c1, "A_fd")
    mg = SemiGeometricMG(Afd, c1["rhs_fd"], coo_from(c1, "Q_pseudo"))
    mg.solve(smoother="GaussSeidel", smooth_steps=1, levels=3, max_iterations=100, error=1e-11,
             gs_order="lexicographic")
    assert len(mg.track_res) == 12
    assert_history_close(mg.track_res, c1["ref_fd_pseudo_L3_s1_track"], Afd, c1["ref_fd_pseudo_L3_s1_x"])


def test_api_iterations_persist_across_solves(env):
    from learnmultigrid_b200.solvers.Multigrid import SemiGeometricMG
    c1 = load_golden("c1_1d_1024.npz")
    A = coo_from(c1, "A")
    mg = SemiGeometricMG(A, c1["rhs"], coo_from(c1, "Q_quasi"))
    kw = dict(levels=2, smoother="GaussSeidel", smooth_steps=1, error=1e-10, gs_order="lexicographic")
    mg.solve(max_iterations=2, **kw)
    x_after_first = mg.get_solution()
    mg.solve(max_iterations=3, initial_guess=None, **kw)
    assert mg.get_iterations() == int(c1["ref_twice_its"])
    assert_history_close(mg.track_res, c1["ref_twice_track"], A, x_after_first)


def test_api_jacobi_and_multicolor_modes_match_oracle(env):
    from learnmultigrid_b200.solvers.Multigrid import SemiGeometricMG
    c1 = load_golden("c1_1d_1024.npz")
    A = coo_from(c1, "A")
    Q = coo_from(c1, "Q_quasi")
    for levels, steps, its in ((3, 1, 10), (3, 3, 6), (2, 3, 5)):
        mg = SemiGeometricMG(A, c1["rhs"], Q)
        mg.solve(levels=levels, smoother="Jacobi", smooth_steps=steps, error=1e-10, max_iterations=40, omega=2.0 / 3.0)
        o = OracleMultigrid(A, c1["rhs"], [Q], smoother="jacobi", omega=2.0 / 3.0, geometric_below=True)
        o.solve(levels=levels, smooth_steps=steps, error=1e-10, max_iterations=40)
        assert mg.get_iterations() == its == len(o.track_res)
        assert_history_close(mg.track_res, o.track_res, A, o.solution)
    # omega = 1 (reference Jacobi.py:35): stagnates exactly like the oracle
    mg = SemiGeometricMG(A, c1["rhs"], Q)
    mg.solve(levels=3, smoother="Jacobi", smooth_steps=1, error=1e-10, max_iterations=12, omega=1.0)
    o = OracleMultigrid(A, c1["rhs"], [Q], smoother="jacobi", omega=1.0, geometric_below=True)
    o.solve(levels=3, smooth_steps=1, error=1e-10, max_iterations=12)
    assert len(mg.track_res) == 12
    assert_history_close(mg.track_res, o.track_res, A, o.solution)
    # multicolour GS against the SciPy/C oracle with the engine's colours
    mg = SemiGeometricMG(A, c1["rhs"], Q)
    mg.solve(levels=3, smoother="GaussSeidel", smooth_steps=2, error=1e-10, max_iterations=40)
    h = mg.get_hierarchy()
    o = OracleMultigrid(A, c1["rhs"], [Q], smoother="mcgs", colors=h.colors, geometric_below=True)
    o.solve(levels=3, smooth_steps=2, error=1e-10, max_iterations=40)
    assert mg.get_iterations() == len(o.track_res)
    assert_history_close(mg.track_res, o.track_res, A, o.solution)


def test_api_2d_two_level_reference_answers(env):
    """BASELINE.md section 2, 2D anchors: reference assembly + SemiGeometricMG.solve(levels=2, GS x3, 1e-9)."""
    from learnmultigrid_b200.solvers.Multigrid import SemiGeometricMG
    from learnmultigrid_b200 import problems as P
    d = load_golden("assembly_2d.npz")
    A = coo_from(d, "N16_A")
    rhs = d["N16_rhs"]
    for Q, last, its in ((P.quasi_l2_Q_2d(16), 3.2880825748e-10, 7), (P.linear_P_2d(16), 9.6242479321e-11, 7)):
        mg = SemiGeometricMG(A, rhs, Q)
        mg.solve(levels=2, smoother="GaussSeidel", smooth_steps=3, error=1e-09, max_iterations=20,
                 gs_order="lexicographic")
        assert mg.get_iterations() == its
        np.testing.assert_allclose(mg.track_res[-1, 0], last, rtol=1e-4)
        np.testing.assert_allclose(mg.get_solution()[8 * 17 + 8, 0], -7.3445766e-02, rtol=1e-6)


@pytest.mark.parametrize("name", ["N16_quasi", "N16_linear", "N32_quasi"])
def test_api_2d_two_level_reference_histories(env, name):
    """SURVEY 8c, 2D two-level answers: the reference's own residual histories, iteration counts and solutions
    (tests/golden/solve_2d.npz, thesis_structured_2d.py:457-458) through the drop-in API with the reference's
    index-order Gauss-Seidel"""
    from learnmultigrid_b200.solvers.Multigrid import SemiGeometricMG
    g = load_golden("solve_2d.npz")
    A, Q, rhs = coo_from(g, name + "_A"), coo_from(g, name + "_Q"), g[name + "_rhs"]
    mg = SemiGeometricMG(A, rhs, Q)
    mg.solve(levels=2, smoother="GaussSeidel", smooth_steps=3, error=1e-09, max_iterations=20,
             gs_order="lexicographic")
    want, x_ref = g[name + "_track"], g[name + "_x"]
    assert mg.get_iterations() == int(g[name + "_its"]) == 7
    assert_history_close(mg.track_res, want, A, x_ref)
    np.testing.assert_allclose(mg.get_solution(), x_ref, rtol=0, atol=1e-12 * np.linalg.norm(x_ref))
    # multicolour Gauss-Seidel (the engine's default smoother) converges at least as fast on this problem
    mc = SemiGeometricMG(A, rhs, Q)
    mc.solve(levels=2, smoother="GaussSeidel", smooth_steps=3, error=1e-09, max_iterations=20)
    assert mc.get_iterations() <= 8
    np.testing.assert_allclose(mc.get_solution(), x_ref, rtol=0, atol=1e-8)


def test_api_stationary_solvers_cg_direct(env):
    from learnmultigrid_b200.solvers.Jacobi import Jacobi
    from learnmultigrid_b200.solvers.GaussSeidel import GaussSeidel
    from learnmultigrid_b200.solvers.CG import CG
    from learnmultigrid_b200.solvers.Solver import DirectSolver
    s = load_golden("solvers_small.npz")
    j = Jacobi(s["A3"], s["b3"])
    j.solve(max_iterations=1000, error=1e-12)
    assert j.get_iterations() == 66
    np.testing.assert_allclose(j.track_res, s["jacobi_track"], rtol=1e-9, atol=1e-14)
    np.testing.assert_allclose(j.get_solution(), s["jacobi_x"], rtol=1e-12)
    g = GaussSeidel(s["A3"], s["b3"])
    g.solve(max_iterations=1000, error=1e-12)
    assert g.get_iterations() == 24
    np.testing.assert_allclose(g.track_res, s["gs_track"], rtol=1e-8, atol=1e-14)
    d = DirectSolver(s["A3"], s["b3"])
    d.solve()
    np.testing.assert_allclose(d.get_solution(), s["direct_x"], rtol=1e-12)
    assert d.get_residual() < 1e-14
    cg = CG(s["cg_A"], s["cg_rhs"])
    cg.solve(max_iterations=200, error=1e-10)
    assert cg.get_iterations() == int(s["cg_its"])
    np.testing.assert_allclose(cg.track_res, s["cg_track"], rtol=1e-8, atol=1e-13)
    np.testing.assert_allclose(cg.get_solution(), s["cg_x"], rtol=1e-9, atol=1e-13)


def test_api_edge_cases(env):
    from learnmultigrid_b200.solvers.Multigrid import SemiGeometricMG
    c1 = load_golden("c1_1d_1024.npz")
    A = coo_from(c1, "A")
    Q = coo_from(c1, "Q_quasi")
    mg = SemiGeometricMG(A, c1["rhs"], Q)
    mg.solve(levels=2, smoother="Jacobi", max_iterations=0)
    assert mg.track_res.shape == (0, 1) and mg.get_iterations() == 0
    with pytest.raises(TypeError):
        mg.solve(levels=2, smoother="NoSuchSmoother")
    with pytest.raises(ValueError):
        mg.solve(levels=1, smoother="Jacobi")
    with pytest.raises(ValueError):
        SemiGeometricMG(A, c1["rhs"], Q[:100]).solve(levels=2, smoother="Jacobi")
    with pytest.raises(SystemExit):
        mg.solve(levels=2, smoother="Jacobi", cycle="W")
    # rhs = 0 converges at the second residual evaluation (the first one is overwritten by sqrt(n))
    mg0 = SemiGeometricMG(A, np.zeros((1025, 1)), Q)
    mg0.solve(levels=2, smoother="GaussSeidel", error=1e-12, max_iterations=5)
    assert mg0.get_iterations() == 2 and mg0.track_res[1, 0] == 0.0


def test_mg_preconditioned_cg_matches_oracle(env):
    """BASELINE configs[4]: CG preconditioned by one symmetric V(1,1) cycle (multicolour GS, post-smoothing in reverse
    colour order) against the same algorithm on the CPU oracle: same iteration count, same history"""
    from learnmultigrid_b200 import problems as P
    from learnmultigrid_b200.solvers.CG import CG
    from learnmultigrid_b200.solvers.Multigrid import SemiGeometricMG
    from oracle.vcycle import pcg_solve
    N, L = 64, 4
    A = P.symmetric_dirichlet(P.structured_laplacian_2d(N, P.variable_coefficient), P.boundary_nodes_2d(N))
    rhs = P.structured_rhs_2d(N)
    Qs = P.structured_hierarchy_2d(N, L, transfer="linear")
    mg = SemiGeometricMG(A, rhs, Qs)
    pre = mg.as_preconditioner(levels=L, smoother="GaussSeidel", smooth_steps=1)
    cg = CG(A, rhs)
    cg.solve(max_iterations=60, error=1e-10, preconditioner=pre)
    o = OracleMultigrid(A, rhs, Qs, smoother="mcgs", colors=mg.get_hierarchy().colors, hoist_setup=True,
                        reverse_post=True)
    o.build_hierarchy(L)
    n = A.shape[0]
    x, hist, its = pcg_solve(A, rhs, lambda r: o.v_cycle(o.matrix, np.zeros((n, 1)), r, 1, L), 60, 1e-10)
    assert cg.get_iterations() == its and its < 15
    np.testing.assert_allclose(cg.track_res, hist, rtol=1e-6)
    np.testing.assert_allclose(cg.get_solution(), x, rtol=0, atol=1e-9 * np.linalg.norm(x))
    # plain CG needs many more iterations on the same problem
    cg0 = CG(A, rhs)
    cg0.solve(max_iterations=400, error=1e-10)
    assert cg0.get_iterations() > 5 * its


def test_api_semigeometric_mg_on_non_nested_2d_meshes(env):
    """the README's headline use: a transfer operator from the semi-geometric L2 projection between an irregular fine
    mesh and an unrelated (non-nested) coarse mesh (L2_projection/coupling2d.py), used by SemiGeometricMG exactly as in
    the 2D scripts (levels=2, Gauss-Seidel x3, thesis_structured_2d.py:457-458); history against the CPU oracle"""
    from learnmultigrid_b200 import problems as P
    from learnmultigrid_b200.L2_projection.L2Projection import L2Projection
    from learnmultigrid_b200.mesh.Mesh2D import Mesh2D
    from learnmultigrid_b200.solvers.Multigrid import SemiGeometricMG
    pb = P.irregular_p1_2d(32, seed=6)
    coarse = Mesh2D(13 * 13)
    Q = L2Projection("quasi", pb["mesh"], coarse).compute_transfer_2d()
    pc = np.asarray(coarse.get_points())
    inner = np.flatnonzero((pc[:, 0] > 1e-12) & (pc[:, 0] < 1 - 1e-12) & (pc[:, 1] > 1e-12) & (pc[:, 1] < 1 - 1e-12))
    Q = sp.csr_matrix(Q[:, inner])
    mg = SemiGeometricMG(pb["A"], pb["rhs"], Q)
    mg.solve(levels=2, smoother="GaussSeidel", smooth_steps=3, error=1e-9, max_iterations=40)
    o = OracleMultigrid(pb["A"], pb["rhs"], [Q], smoother="mcgs", colors=mg.get_hierarchy().colors, hoist_setup=True)
    o.solve(levels=2, smooth_steps=3, error=1e-9, max_iterations=40)
    assert mg.get_iterations() == len(o.track_res) < 40
    assert_history_close(mg.track_res, o.track_res, pb["A"], o.solution)
    np.testing.assert_allclose(mg.get_solution(), o.solution, rtol=0, atol=1e-12 * np.linalg.norm(o.solution))
