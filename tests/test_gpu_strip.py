"""GPU tests of distributed_strip.StripHierarchy (partitioned hierarchy built from row blocks only): iterates and
residual norms bit-identical to the single-GPU hierarchy, local products on the device SpGEMM identical to SciPy's.

Also: device first-fit colouring, implied columns.  First run on a B200 in round 2
(profiles/r02_pytest_gpu_unverified_first_run.log: 14 passed); the MGB_UNVERIFIED gate of round 1 is gone."""

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = [pytest.mark.gpu]


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    return torch


def test_device_ops_equal_scipy_ops(torch_mod):
    from learnmultigrid_b200 import formats as F, partition_setup as PS, problems as P, setup_device as SD
    from learnmultigrid_b200.distributed_strip import DeviceOps
    torch = torch_mod
    ops = DeviceOps(SD.DeviceSetup(torch, torch.device("cuda", 0)))
    A = F.canonical_csr(P.structured_laplacian_2d(16, P.variable_coefficient))
    Q = F.canonical_csr(P.structured_hierarchy_2d(16, 2, "quasi")[0])
    for X, Y in ((sp.csr_matrix(A.T), Q), (sp.csr_matrix(Q.T), A), (A[3:40], Q)):
        got, want = ops.spgemm(X, Y), PS.ScipyOps.spgemm(X, Y)
        got.sort_indices()
        assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
        assert np.array_equal(got.data, want.data)
    got, want = ops.transpose(Q[5:77]), PS.ScipyOps.transpose(Q[5:77])
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
    assert np.array_equal(got.data, want.data)
    assert ops.spgemm(sp.csr_matrix((0, 5)), sp.csr_matrix((5, 3))).shape == (0, 3)
    assert ops.transpose(sp.csr_matrix((4, 0))).shape == (0, 4)


@pytest.mark.parametrize("smoother,transfer,world,n_dist", [("mcgs", "linear", 2, 2), ("mcgs", "quasi", 4, 2),
                                                            ("jacobi", "linear", 3, 1), ("mcgs", "linear", 4, 3)])
def test_strip_hierarchy_is_bit_identical_to_the_single_gpu_cycle(torch_mod, smoother, transfer, world, n_dist):
    from learnmultigrid_b200 import formats as F, partition as PT, problems as P
    from learnmultigrid_b200.distributed import run_virtual_ranks
    from learnmultigrid_b200.distributed_strip import StripHierarchy
    from learnmultigrid_b200.engine import DeviceHierarchy
    N, levels, nu, cycles = 64, 4, 2, 3
    A = F.canonical_csr(P.structured_laplacian_2d(N, P.variable_coefficient))
    Qs = [F.canonical_csr(q) for q in P.structured_hierarchy_2d(N, levels, transfer)]
    b = P.structured_rhs_2d(N)
    rng = np.random.default_rng(3)
    x0 = rng.standard_normal((A.shape[0], 1))
    ns = [A.shape[0]] + [q.shape[1] for q in Qs]
    offs = [PT.block_offsets(n, world) for n in ns]

    h = DeviceHierarchy(A, Qs, smoother=smoother)
    h.set_rhs(b)
    h.set_x(x0)
    params = h.make_params(nu_pre=nu, nu_post=nu, omega=2.0 / 3.0)
    want_x, want_n = [], []
    for _ in range(cycles):
        want_n.append(h.residual_norm())
        h.vcycle(params)
        want_x.append(h.get_x().copy())

    def body(fab):
        r = fab.rank
        hs = StripHierarchy(A[offs[0][r]:offs[0][r + 1]], [q[offs[l][r]:offs[l][r + 1]] for l, q in enumerate(Qs)],
                            offs, fab, n_dist, smoother=smoother, region_bytes=1 << 20, max_sites=256, timeout_s=30.0)
        hs.set_rhs(b[offs[0][r]:offs[0][r + 1]])                      # the owned block is enough
        hs.set_x(x0)
        p = hs.make_params(nu_pre=nu, nu_post=nu, omega=2.0 / 3.0)
        xs, norms = [], []
        for _ in range(cycles):
            norms.append(hs.residual_norm())
            hs.vcycle(p)
            xs.append(hs.get_x().copy())
        hs.check()
        nnz = (list(hs._global_nnzA), list(hs._global_nnzQ))
        hs.close()
        return xs, norms, nnz

    for xs, norms, nnz in run_virtual_ranks(world, body):
        for got, want in zip(xs, want_x):
            assert np.array_equal(got, want)                           # bit for bit
        np.testing.assert_allclose(norms, want_n, rtol=1e-13)         # blockwise sum of the squares
        assert nnz[0] == [lv.nnz_A for lv in h.levels] and nnz[1] == [lv.nnz_Q for lv in h.levels[:-1]]


def test_device_first_fit_colouring_equals_the_host_helper(torch_mod):
    """csrc/color_kernels.cu: the round-based colouring on the device gives the colours of formats.greedy_colors, entry
    for entry (structured operator with identity rows, 7- / 19-point Galerkin levels, unstructured numbering, a random
    unsymmetric pattern); a 1D chain exceeds the round limit and reports UNSUPPORTED"""
    import scipy.sparse as sp
    from learnmultigrid_b200 import _lib, formats as F, problems as P, setup_device as SD
    torch = torch_mod
    S = SD.DeviceSetup(torch, torch.device("cuda", 0))
    N = 128
    A = F.canonical_csr(P.structured_laplacian_2d(N, P.variable_coefficient))
    mats = [A]
    for transfer in ("linear", "quasi"):
        Q = F.canonical_csr(P.structured_hierarchy_2d(N, 2, transfer)[0])
        mats.append(F.canonical_csr(sp.csr_matrix(Q.T @ sp.csc_matrix(A) @ Q)))
    mats.append(F.canonical_csr(P.irregular_p1_2d(64)["A"]))
    mats.append(F.canonical_csr(sp.random(3000, 3000, density=0.003, random_state=1, format="csr")
                                + sp.diags((np.arange(3000) % 5 > 0) * 1.0)))
    for M in mats:
        want, nc = F.greedy_colors(M)
        got = np.asarray(S.first_fit_colors(S.upload(M)))
        assert np.array_equal(got, want)
        assert S.last_color_rounds >= 1
        # the hierarchy setup keeps the colours on the device and hands the colouring the pattern of A^T it already has
        AT = S.transpose(S.upload(M))
        dev_col = S.first_fit_colors(S.upload(M), t_pattern=(AT.indptr, AT.indices), host=False)
        assert dev_col.is_cuda and np.array_equal(dev_col.cpu().numpy(), want)
    chain = F.canonical_csr(sp.diags([np.ones(4999), 2 * np.ones(5000), np.ones(4999)], [-1, 0, 1], format="csr"))
    with pytest.raises(_lib.MgError):
        S.first_fit_colors(S.upload(chain), max_rounds=1000)
    assert np.array_equal(np.asarray(S.first_fit_colors(S.upload(chain), max_rounds=6000)), F.greedy_colors(chain)[0])


def test_implied_columns_are_bit_identical(torch_mod, monkeypatch):
    """implied columns (default on; here with the row floor lowered so that a 66 k-row level uses them): regular slices
    compute their columns from the row; the device-built offset tables equal the host twin, every SELL mode gives the
    same bits as the ordinary kernels, and so does a whole solve (multicolour Gauss-Seidel and Jacobi)"""
    import ctypes
    from learnmultigrid_b200 import _lib, formats as F, problems as P
    from learnmultigrid_b200.engine import DeviceHierarchy
    torch = torch_mod
    lib = _lib.load()
    N, levels = 256, 4
    A = F.canonical_csr(P.structured_laplacian_2d(N, P.variable_coefficient))
    Qs = [F.canonical_csr(q) for q in P.structured_hierarchy_2d(N, levels, "linear")]
    b = P.structured_rhs_2d(N)
    x0 = np.random.default_rng(2).standard_normal((A.shape[0], 1))

    def run(smoother):
        h = DeviceHierarchy(A, Qs, smoother=smoother)
        h.set_rhs(b)
        h.set_x(x0)
        p = h.make_params(nu_pre=2, nu_post=1, omega=2.0 / 3.0)
        out = []
        for _ in range(3):
            out.append(h.residual_norm())
            h.vcycle(p)
            out.append(h.get_x().copy())
        return h, out

    monkeypatch.setenv("MGB_IMPLIED_COLUMNS", "0")
    prev = lib.mg_set_implied_columns(0)
    floor = lib.mg_set_implied_min_rows(1)
    try:
        plain = {s: run(s)[1] for s in ("mcgs", "jacobi")}
        assert run("mcgs")[0].levels[0].A.slice_off is None
        monkeypatch.setenv("MGB_IMPLIED_COLUMNS", "1")
        monkeypatch.setenv("MGB_IMPLIED_MIN_ROWS", "1")
        lib.mg_set_implied_columns(1)
        for s in ("mcgs", "jacobi"):
            h, got = run(s)
            lev = h.levels[0]
            assert lev.A.slice_off is not None and lev.A.struct.d_slice_rec and lev.A.struct.nrec >= 1
            # device-built offsets = host twin on the same (colour-blocked) matrix
            sell = (lev.A.slice_ptr.cpu().numpy(), lev.A.cols.cpu().numpy(), lev.A.vals.cpu().numpy())
            want = F.sell_slice_offsets(sell, lev.A.shape[0], lev.A.uniform_len)
            dev = lev.A.slice_off.cpu().numpy().reshape(-1, 8)[:, :want.shape[1]]      # 8 ints per slice on the device
            reg = want[:, 0] != F.SLICE_IRREGULAR
            assert np.array_equal(dev[:, 0] != F.SLICE_IRREGULAR, reg) and np.array_equal(dev[reg], want[reg])
            assert reg.mean() > 0.7
            for g, w in zip(got, plain[s]):
                assert np.array_equal(np.asarray(g), np.asarray(w))
    finally:
        lib.mg_set_implied_columns(prev)
        lib.mg_set_implied_min_rows(floor)


def test_implied_columns_on_partitioned_levels(torch_mod, monkeypatch):
    """the implied-columns kernels carrying an exchange site, and the partitioned cycle's shortcuts (fused residual of
    the last colour, skipped prolongation rows, norm after the cycle): against the plain single-GPU cycle -- no implied
    columns, one pass per operation -- bit for bit"""
    from learnmultigrid_b200 import _lib, formats as F, problems as P
    from learnmultigrid_b200.distributed import DistributedHierarchy, run_virtual_ranks
    from learnmultigrid_b200.engine import DeviceHierarchy
    lib = _lib.load()
    N, levels, nu, cycles = 256, 4, 1, 3
    A = F.canonical_csr(P.structured_laplacian_2d(N))
    Qs = [F.canonical_csr(q) for q in P.structured_hierarchy_2d(N, levels, "linear")]
    b = P.structured_rhs_2d(N)
    x0 = np.random.default_rng(8).standard_normal((A.shape[0], 1))
    monkeypatch.setenv("MGB_IMPLIED_COLUMNS", "0")
    prev = lib.mg_set_implied_columns(0)
    floor = lib.mg_set_implied_min_rows(1)
    lib.mg_set_cycle_fusion(0)
    try:
        h = DeviceHierarchy(A, Qs, smoother="mcgs")
        h.set_rhs(b)
        h.set_x(x0)
        params = h.make_params(nu_pre=nu, nu_post=nu)
        want, want_norm = [], []
        for _ in range(cycles):
            h.vcycle(params)
            want.append(h.get_x().copy())
            want_norm.append(h.residual_norm())
    finally:
        lib.mg_set_cycle_fusion(1)

    def body(fab):
        hd = DistributedHierarchy(A, Qs, fab, smoother="mcgs", colors=h.colors, n_dist=2, region_bytes=1 << 20,
                                  max_sites=256, timeout_s=30.0)
        assert hd.levels[0].A.slice_off is not None
        assert all(int(lv.flags) == 3 for lv in hd.levels[:-1])
        p = hd.make_params(nu_pre=nu, nu_post=nu)
        out = []
        for mode in ("before", "after"):
            hd.set_rhs(b)
            hd.set_x(x0)
            xs, ns = [], []
            for _ in range(cycles):
                hd.vcycle(p, with_norm=mode == "before", norm_after=mode == "after")
                xs.append(hd.get_x().copy())
                ns.append(hd.last_norm())
            out.append((xs, ns))
        hd.check()
        hd.close()
        return out

    monkeypatch.setenv("MGB_IMPLIED_COLUMNS", "1")
    monkeypatch.setenv("MGB_IMPLIED_MIN_ROWS", "1")
    try:
        lib.mg_set_implied_columns(1)
        res = run_virtual_ranks(2, body)
    finally:
        lib.mg_set_implied_columns(prev)
        lib.mg_set_implied_min_rows(floor)
    for out in res:
        for xs, ns in out:
            for got, w in zip(xs, want):
                assert np.array_equal(got, w)
        np.testing.assert_allclose(out[1][1], want_norm, rtol=1e-12)          # the norm AFTER each cycle
