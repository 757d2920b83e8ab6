"""Host-side logic of the engine (no GPU): formats, orderings, colourings, and the level data handed to the
kernels, checked by emulating the kernels in NumPy (tests/helpers.py) against the CPU oracle."""
import numpy as np
import pytest
import scipy.sparse as sp
from scipy.sparse.linalg import spsolve

from learnmultigrid_b200 import formats as F
from oracle import kernels as K
from oracle.vcycle import OracleMultigrid, geometric_interpolator
from helpers import (bilinear_P, coo_from, emulate_vcycle, load_golden, poisson2d, sell_rowsum)


def test_sell_layout_roundtrip_and_padding():
    rng = np.random.default_rng(0)
    for shape, dens in (((100, 80), 0.1), ((33, 33), 0.3), ((1, 5), 1.0), ((64, 64), 0.02), ((70, 3), 0.5)):
        A = F.canonical_csr(sp.random(*shape, density=dens, random_state=1, format="csr"))
        sell = F.csr_to_sell(A)
        slice_ptr, cols, vals = sell
        assert slice_ptr[0] == 0 and np.all(np.diff(slice_ptr) % 32 == 0)
        assert len(slice_ptr) == (shape[0] + 31) // 32 + 1
        assert cols.min() >= 0 and cols.max() < shape[1]          # padding columns are valid gathers
        x = rng.standard_normal(shape[1])
        assert np.array_equal(sell_rowsum(sell, shape[0], x), A @ x)
        assert np.count_nonzero(vals) == np.count_nonzero(A.data)


def test_empty_rows_and_ragged():
    A = sp.lil_matrix((40, 40))
    A[0, 0] = 2.0
    A[5, :] = 1.0                      # one very long row in the first slice
    A[39, 38] = -1.0
    A = F.canonical_csr(A)
    sell = F.csr_to_sell(A)
    x = np.arange(40, dtype=float)
    assert np.array_equal(sell_rowsum(sell, 40, x), A @ x)


def test_permute_keeps_row_entry_order():
    A = poisson2d(8)
    colors, nc = F.greedy_colors(A)
    assert nc == 2
    perm, cptr = F.color_permutation(colors)
    iperm = F.inverse_permutation(perm)
    Ap = F.permute_csr(A, perm, iperm)
    for newr in (0, 7, 40, 80):
        oldr = perm[newr]
        a0, a1 = A.indptr[oldr], A.indptr[oldr + 1]
        p0, p1 = Ap.indptr[newr], Ap.indptr[newr + 1]
        assert np.array_equal(perm[Ap.indices[p0:p1]], A.indices[a0:a1])     # same entries, same order
        assert np.array_equal(Ap.data[p0:p1], A.data[a0:a1])


def test_greedy_coloring_is_valid_on_nonsymmetric_pattern():
    A = poisson2d(12)                              # row-replaced Dirichlet rows: structurally nonsymmetric
    colors, nc = F.greedy_colors(A)
    assert nc == 2                                 # red-black on the 5-point grid
    S = sp.csr_matrix(A + A.T)
    S.setdiag(0)
    S.eliminate_zeros()
    r, c = S.nonzero()
    assert np.all(colors[r] != colors[c])
    P = bilinear_P(12)
    Ac = F.canonical_csr(sp.csr_matrix(P.T @ A @ P))
    colors, nc = F.greedy_colors(Ac)
    S = sp.csr_matrix(Ac + Ac.T)
    S.setdiag(0)
    S.eliminate_zeros()
    r, c = S.nonzero()
    assert np.all(colors[r] != colors[c]) and nc <= 6


def test_lex_levels_respect_dependencies():
    A = poisson2d(10)
    lp, lr = F.lex_levels(A)
    level = np.empty(A.shape[0], dtype=np.int64)
    for l in range(len(lp) - 1):
        level[lr[lp[l]:lp[l + 1]]] = l
    S = sp.csr_matrix(A + A.T)
    r, c = S.nonzero()
    m = c < r
    assert np.all(level[c[m]] < level[r[m]])
    assert len(lp) - 1 == 2 * 11 - 1 - 2 or len(lp) - 1 <= 2 * 11      # anti-diagonal wavefronts
    # emulate the level-scheduled sweep and compare with the serial kernel, bit for bit
    rng = np.random.default_rng(1)
    x = rng.standard_normal(A.shape[0])
    b = rng.standard_normal(A.shape[0])
    xs = x.copy()
    K.gauss_seidel(A, xs, b, iterations=2)
    xl = x.copy()
    for _ in range(2):
        for l in range(len(lp) - 1):
            K.gauss_seidel_multicolor(A, xl, b, [lr[lp[l]:lp[l + 1]]])
    assert np.array_equal(xs, xl)


def test_geometric_interpolator_csr_equals_reference_dense():
    for n in (3, 4, 5, 6, 9, 17, 33, 257, 1025):
        assert np.array_equal(F.geometric_interpolator_csr(n).toarray(), geometric_interpolator(n))


@pytest.mark.parametrize("smoother,omega", [("jacobi", 2.0 / 3.0), ("jacobi", 1.0), ("mcgs", 1.0)])
@pytest.mark.parametrize("nu", [1, 2, 3])
def test_emulated_device_cycle_matches_oracle_2d(smoother, omega, nu):
    """The level data the engine uploads (permuted SELL operators, transfers, colour blocks), driven by a
    NumPy mirror of cycle.cu, reproduces the oracle V-cycle on a 3-level 2D hierarchy."""
    N = 16
    A = poisson2d(N)
    Qs = [bilinear_P(N), bilinear_P(N // 2)]
    rng = np.random.default_rng(7)
    b = rng.standard_normal(A.shape[0])
    x0 = rng.standard_normal(A.shape[0])
    host = F.build_host_hierarchy(A, Qs, smoother)
    colors = [d["colors"] for d in host]
    o = OracleMultigrid(A, b.reshape(-1, 1), Qs, smoother="mcgs" if smoother == "mcgs" else "jacobi",
                        omega=omega, colors=colors, hoist_setup=True)
    o.build_hierarchy(3)
    want = o.v_cycle(o.matrix, x0.reshape(-1, 1).copy(), b.reshape(-1, 1), nu, 3).ravel()
    perm = host[0]["perm"]
    xin = x0 if perm is None else x0[perm]
    bin_ = b if perm is None else b[perm]
    got = emulate_vcycle(host, smoother, nu, nu, omega, xin, bin_,
                         coarse_solve=lambda Ac, rc: spsolve(sp.csc_matrix(Ac), rc))
    if perm is not None:
        out = np.empty_like(got)
        out[perm] = got
        got = out
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12 * np.linalg.norm(want))


def test_emulated_cycle_matches_oracle_1d_c1():
    c1 = load_golden("c1_1d_1024.npz")
    A = coo_from(c1, "A")
    Q0 = coo_from(c1, "Q_quasi")
    Qs = [Q0, F.geometric_interpolator_csr(513)]
    b = c1["rhs"].ravel()
    host = F.build_host_hierarchy(A, Qs, "jacobi")
    o = OracleMultigrid(A, b.reshape(-1, 1), Qs, smoother="jacobi", omega=2.0 / 3.0, hoist_setup=True)
    o.build_hierarchy(3)
    x = np.zeros(1025)
    xo = np.zeros((1025, 1))
    for _ in range(3):
        x = emulate_vcycle(host, "jacobi", 1, 1, 2.0 / 3.0, x, b,
                           coarse_solve=lambda Ac, rc: spsolve(sp.csc_matrix(Ac), rc))
        xo = o.v_cycle(o.matrix, xo, b.reshape(-1, 1), 1, 3)
        np.testing.assert_allclose(x, xo.ravel(), rtol=0, atol=1e-12 * np.linalg.norm(xo))


def test_galerkin_pattern_matches_scipy_reference_semantics():
    """host hierarchy = csr_matrix(Q.T @ A @ Q) (Multigrid.py:97-98): exact zeros pruned, sorted CSR."""
    A = poisson2d(8)
    P = bilinear_P(8)
    host = F.build_host_hierarchy(A, [P], "jacobi", with_sell=False)
    want = sp.csr_matrix(P.T @ sp.csc_matrix(A) @ P)
    want.sort_indices()
    got = host[1]["A_nat"]
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
    assert np.array_equal(got.data, want.data)
    assert np.all(got.data != 0.0)


def test_structured_colorings_are_valid_on_the_galerkin_hierarchy():
    """(ix+iy)%2 on the 5-point fine operator, (ix+iy)%3 on the 7-point Galerkin operators of linear transfers"""
    import scipy.sparse as sp
    from learnmultigrid_b200 import problems as P
    for coef in (None, P.variable_coefficient):
        N, L = 32, 4
        A = P.structured_laplacian_2d(N, coef)
        Qs = P.structured_hierarchy_2d(N, L, transfer="linear")
        cols = P.structured_colors_2d(N, L)
        assert cols[-1] is None
        for l in range(L - 1):
            assert cols[l].max() == (1 if l == 0 else 2)
            assert P.coloring_is_valid(A, cols[l]), "level %d" % l
            A = sp.csr_matrix(Qs[l].T @ A @ Qs[l])


def test_reverse_post_smoothing_makes_the_cycle_symmetric():
    """V(1,1) with multicolour GS forward before and REVERSED after the coarse correction is a symmetric operator on a
    symmetric problem (what MG-preconditioned CG needs); the reference order is not"""
    from learnmultigrid_b200 import problems as P, formats as F
    from oracle.vcycle import OracleMultigrid
    N, L = 16, 3
    A = P.symmetric_dirichlet(P.structured_laplacian_2d(N, P.variable_coefficient), P.boundary_nodes_2d(N))
    assert abs(A - A.T).max() == 0.0
    Qs = P.structured_hierarchy_2d(N, L, transfer="linear")
    n = A.shape[0]
    rng = np.random.default_rng(0)
    r1, r2 = rng.standard_normal((n, 1)), rng.standard_normal((n, 1))
    import scipy.sparse as sp
    cols, Al = [], A
    for l in range(L - 1):
        cols.append(F.greedy_colors(F.canonical_csr(Al))[0])
        Al = sp.csr_matrix(Qs[l].T @ Al @ Qs[l])
    cols.append(None)
    asym = {}
    for rev in (False, True):
        o = OracleMultigrid(A, r1, Qs, smoother="mcgs", colors=cols, hoist_setup=True, reverse_post=rev)
        o.build_hierarchy(L)
        M = lambda r: o.v_cycle(o.matrix, np.zeros((n, 1)), r, 1, L)
        asym[rev] = abs((M(r1).T @ r2).item() - (r1.T @ M(r2)).item()) / abs((M(r1).T @ r2).item())
    assert asym[True] < 1e-12 < asym[False]


@pytest.mark.parametrize("transfer,expect", [("linear", [2, 3, 3]), ("quasi", None)])
def test_lattice_colorings_are_valid_and_smaller_than_greedy(transfer, expect):
    """colour = (alpha*ix + beta*iy) mod m from the stencil offsets of a small model hierarchy, applied to a larger one"""
    import scipy.sparse as sp
    from learnmultigrid_b200 import problems as P, formats as F
    N, L = 64, 4
    for coef in (None, P.variable_coefficient):
        A = sp.csr_matrix(P.structured_laplacian_2d(N, coef))
        Qs = P.structured_hierarchy_2d(N, L, transfer=transfer)
        cols = P.lattice_colors_2d(N, L, transfer=transfer, coefficient=coef, model_N=32)
        counts = []
        for l in range(L - 1):
            assert P.coloring_is_valid(A, cols[l]), "level %d" % l
            counts.append(int(cols[l].max()) + 1)
            greedy = int(F.greedy_colors(F.canonical_csr(A))[1])
            assert counts[-1] <= greedy
            A = sp.csr_matrix(Qs[l].T @ A @ Qs[l])
        if expect is not None:
            assert counts == expect
        else:
            assert counts[0] == 2 and counts[1] <= 8 and counts[2] <= 14     # 19- and 37-point stencils


def _first_fit_python(A):
    """the definition: symmetrised pattern, rows with off-diagonal entries in index order, then the rows with nothing
    but a diagonal entry; each takes the smallest colour no coloured neighbour has"""
    A = F.canonical_csr(A)
    n = A.shape[0]
    S = sp.csr_matrix((np.ones(A.nnz), A.indices, A.indptr), shape=A.shape)
    S = sp.csr_matrix(S + S.T)
    off = np.array([np.any(A.indices[A.indptr[i]:A.indptr[i + 1]] != i) for i in range(n)], dtype=bool)
    colors = -np.ones(n, dtype=np.int64)
    for i in list(np.flatnonzero(off)) + list(np.flatnonzero(~off)):
        nb = S.indices[S.indptr[i]:S.indptr[i + 1]]
        used = set(colors[nb[nb != i]].tolist())
        c = 0
        while c in used:
            c += 1
        colors[i] = c
    return colors


def test_greedy_coloring_equals_the_python_definition_including_more_than_64_colours():
    rng = np.random.default_rng(11)
    mats = [poisson2d(9), sp.random(150, 150, density=0.05, random_state=2, format="csr") + sp.eye(150)]
    C = sp.lil_matrix((100, 100))
    C[:70, :70] = 1.0                               # a 70-clique: colours 0..69 cross the 64-bit mask boundary
    C[70:, 3] = 1.0                                 # one-directional couplings into the clique
    C.setdiag(1.0)
    C[95, :] = 0.0
    C[95, 95] = 1.0                                 # a diagonal-only row that others do not reference
    mats.append(C.tocsr())
    mats.append(sp.csr_matrix(rng.random((40, 40)) < 0.5) * 1.0 + sp.eye(40))
    for M in mats:
        M = F.canonical_csr(M)
        got, nc = F.greedy_colors(M)
        want = _first_fit_python(M)
        assert np.array_equal(got, want)
        assert nc == want.max() + 1
    assert F.greedy_colors(mats[2])[1] >= 70


def test_greedy_coloring_refuses_more_than_128_colours():
    from learnmultigrid_b200 import _lib
    full = F.canonical_csr(sp.csr_matrix(np.ones((130, 130))))
    with pytest.raises(_lib.MgError):
        F.greedy_colors(full)


def test_interface_helpers_of_the_assembly_and_mesh_mirror():
    """small pieces behind the reference's names: affine interval map, element table, rule tables, invalid orders"""
    from learnmultigrid_b200.assembly.MapReferenceElement import IntervalMap, g_function, inv_g_function
    from learnmultigrid_b200.assembly.Quadrature import Quadrature, Quadrature2D
    from learnmultigrid_b200.assembly.ShapeFunction import Function, GradientTriangle
    from learnmultigrid_b200.assembly.LoadFunction import LoadFunction
    from learnmultigrid_b200.mesh.Mesh1D import Mesh1D, Mesh1DRefinement
    m = IntervalMap(0.25, 0.75)
    assert m.to_physical(0.5) == 0.5 and m.to_reference(0.5) == 0.5
    assert g_function(0.2, 1.0, 3.0) == 1.0 + 0.2 * (3.0 - 1.0) and inv_g_function(1.4, 1.0, 3.0) == (1.4 - 1.0) / (3.0 - 1.0)
    q = Quadrature(3)
    assert np.array_equal(q.get_points(), [0.11270166537926, 0.5, 0.88729833462074])      # 14-digit constants
    assert np.array_equal(q.get_weights(), [0.27777777777778, 0.44444444444444, 0.27777777777778])
    assert Quadrature.order_to_points(2) == "Invalid order" and Quadrature2D.order_to_weights(1) == "Invalid order"
    assert np.array_equal(Quadrature2D(3).get_weights(), [1 / 6] * 3) and Quadrature2D(3).get_points().shape == (3, 2)
    phi = Function(2)
    assert q.compute(phi, (0, 1)) == q.compute(phi, np.array([1, 0]))
    assert abs(q.compute(phi, (0, 0)) - 1 / 3) < 1e-13 and abs(q.compute_single(phi, 1, LoadFunction(lambda x: 1.0)) - 0.5) < 1e-13
    assert GradientTriangle(1).evaluate(None, 2).tolist() == [[0], [1]]
    with pytest.raises(TypeError):
        LoadFunction(3.0)
    mesh = Mesh1D(True, 4)
    mesh.construct()
    els = mesh.elements()
    assert [e.index for e in els] == [0, 1, 2, 3] and els[2].left == 0.5 and els[2].length == 0.25
    fine = Mesh1DRefinement(3, 2)
    fine.construct()
    assert fine.get_ne() == 12 and fine.get_np() == 13 and fine.get_connections().shape == (12, 2)


def test_round_based_colouring_equals_the_serial_first_fit():
    """csrc/color_kernels.cu (the device colouring) run as a serial host emulation with the kernels' per-row code:
    identical colours to mg_host_greedy_color on structured operators with identity rows, Galerkin levels (7-, 19-,
    37-point), unsymmetric random patterns, empty rows, and beyond 64 colours; the number of rounds is the length of
    the longest dependency chain (two grid widths on a row-major grid)"""
    from learnmultigrid_b200 import _lib, problems as P
    N = 24
    A = P.structured_laplacian_2d(N)
    mats = [A]
    for transfer in ("linear", "quasi"):
        cur = sp.csc_matrix(A)
        for Q in P.structured_hierarchy_2d(N, 3, transfer):
            cur = sp.csr_matrix(Q.T @ cur @ Q)
            mats.append(cur)
    mats.append(sp.random(200, 200, density=0.04, random_state=3, format="csr") + sp.diags((np.arange(200) % 4 > 0) * 1.0))
    C = sp.lil_matrix((100, 100))
    C[:70, :70] = 1.0
    C[70:, 3] = 1.0
    C.setdiag(1.0)
    C[95, :] = 0.0
    mats.append(C.tocsr())                                        # 70 colours, one-directional couplings, an empty row
    for M in mats:
        M = F.canonical_csr(M)
        want, nc = F.greedy_colors(M)
        got, nc2, rounds = F.greedy_colors_by_rounds(M)
        assert np.array_equal(got, want) and nc2 == nc
        assert 1 <= rounds <= M.shape[0]
    _, _, rounds = F.greedy_colors_by_rounds(A)
    assert rounds <= 2 * (N + 1) + 2                              # anti-diagonal wavefronts + the identity rows
    with pytest.raises(_lib.MgError):
        F.greedy_colors_by_rounds(F.canonical_csr(sp.csr_matrix(np.ones((130, 130)))))


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_sell_layout_on_ragged_and_empty_inputs(seed):
    """SELL-32 host layout (the cross-check of the device builder) on random ragged matrices: row counts that are not
    multiples of 32, empty rows, one empty matrix, a single row; every row sum equals CSR's in the same order, padding
    has value 0 and a valid column"""
    rng = np.random.default_rng(seed)
    shapes = [(1, 5), (31, 40), (33, 7), (64, 64), (100, 90), (5, 200)]
    for (n, m) in shapes:
        A = sp.random(n, m, density=rng.uniform(0.02, 0.4), random_state=seed * 7 + n, format="lil")
        for r in rng.integers(0, n, size=max(n // 5, 1)):
            A[int(r), :] = 0                                   # empty rows
        A = F.canonical_csr(sp.csr_matrix(A))
        A.eliminate_zeros()
        slice_ptr, cols, vals = F.csr_to_sell(A)
        assert len(slice_ptr) == (n + 31) // 32 + 1 and slice_ptr[0] == 0 and np.all(np.diff(slice_ptr) % 32 == 0)
        assert len(cols) == len(vals) == slice_ptr[-1]
        assert np.all((cols >= 0) & (cols < m)) or slice_ptr[-1] == 0
        assert np.count_nonzero(vals) == np.count_nonzero(A.data)
        x = rng.standard_normal(m)
        want = np.array([np.sum(A.data[A.indptr[i]:A.indptr[i + 1]] * x[A.indices[A.indptr[i]:A.indptr[i + 1]]]
                                * 1.0) if A.indptr[i + 1] > A.indptr[i] else 0.0 for i in range(n)])
        got = sell_rowsum((slice_ptr, cols, vals), n, x)
        np.testing.assert_allclose(got, want, rtol=1e-14, atol=1e-15)
    empty = F.canonical_csr(sp.csr_matrix((0, 4)))
    sp0, c0, v0 = F.csr_to_sell(empty)
    assert len(sp0) == 1 and len(c0) == 0 and len(v0) == 0


def test_slice_relative_columns_of_structured_levels():
    """formats.sell_slice_offsets: on the colour-blocked structured operators nearly every slice is regular (columns =
    row + a per-slice offset table); the columns rebuilt from the offsets equal the stored ones on exactly those slices;
    slices with identity rows, ragged tails or unstructured rows are marked irregular"""
    from learnmultigrid_b200 import problems as P
    N = 128
    A = P.structured_laplacian_2d(N, P.variable_coefficient)
    mats = [F.canonical_csr(A)]
    cur = sp.csc_matrix(A)
    for Q in P.structured_hierarchy_2d(N, 2, "linear"):
        cur = sp.csr_matrix(Q.T @ cur @ Q)
        mats.append(F.canonical_csr(cur))
    for l, M in enumerate(mats):
        colors, nc = F.greedy_colors(M)
        perm, cptr = F.color_permutation(colors)
        Mp = F.permute_csr(M, perm, F.inverse_permutation(perm))
        sell = F.csr_to_sell(Mp)
        slice_ptr, cols, vals = sell
        lens = np.diff(slice_ptr) // 32
        assert lens.min() == lens.max()                               # uniform on structured levels
        L = int(lens.max())
        off = F.sell_slice_offsets(sell, Mp.shape[0], L)
        regular = off[:, 0] != F.SLICE_IRREGULAR
        assert l > 0 or regular.mean() > 0.45          # two slices per grid line hold a boundary node (129 wide here)
        nsl = len(lens)
        c = cols.reshape(nsl, L, 32)
        rows = (np.arange(nsl) * 32)[:, None, None] + np.arange(32)[None, None, :]
        rebuilt = rows + off[:, :, None].astype(np.int64)
        assert np.array_equal(rebuilt[regular], c[regular])
        assert not np.any(np.all((rebuilt == c)[~regular], axis=(1, 2)) & ((np.arange(nsl)[~regular] + 1) * 32 <= Mp.shape[0]))
        # a slice that mixes identity (boundary) rows with stencil rows is never regular; a slice of identity rows
        # only is (padding repeats the row's last column, i.e. the row itself: all offsets 0)
        ident = np.bincount(np.flatnonzero(np.diff(Mp.indptr) == 1) // 32, minlength=nsl)
        full = np.minimum(32, Mp.shape[0] - np.arange(nsl) * 32)
        assert not np.any(regular[(ident > 0) & (ident < full)])
    # at benchmark-like widths almost everything is regular (1 - 2 / (slices per grid line))
    big = F.canonical_csr(P.structured_laplacian_2d(1024))
    colors, _ = F.greedy_colors(big)
    perm, _ = F.color_permutation(colors)
    sell = F.csr_to_sell(F.permute_csr(big, perm, F.inverse_permutation(perm)))
    assert (F.sell_slice_offsets(sell, big.shape[0], 5)[:, 0] != F.SLICE_IRREGULAR).mean() > 0.93
    R = F.canonical_csr(sp.random(96, 96, density=0.1, random_state=1, format="csr") + sp.eye(96))
    sell = F.csr_to_sell(R)
    lens = np.diff(sell[0]) // 32
    if lens.min() == lens.max():
        assert not np.any(F.sell_slice_offsets(sell, 96, int(lens.max()))[:, 0] != F.SLICE_IRREGULAR)
    assert F.sell_slice_offsets((np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.int32), np.zeros(0)), 0, 5).shape == (0, 5)


def _records_case(M):
    """colour-blocked uniform SELL of M, its host-twin offsets padded to 8, a value dictionary if M has <= 256 values"""
    import torch
    colors, nc = F.greedy_colors(M)
    perm, cptr = F.color_permutation(colors)
    Mp = F.permute_csr(M, perm, F.inverse_permutation(perm))
    sell = F.csr_to_sell(Mp)
    slice_ptr, cols, vals = sell
    lens = np.diff(slice_ptr) // 32
    assert lens.min() == lens.max()
    L, nsl = int(lens.max()), len(lens)
    off = np.zeros((nsl, 8), dtype=np.int32)
    off[:, :L] = F.sell_slice_offsets(sell, Mp.shape[0], L)
    irregular = off[:, 0] == F.SLICE_IRREGULAR
    off[irregular, 1:] = 0
    table, idx = np.unique(vals.view(np.int64), return_inverse=True)          # distinct BIT patterns (+0 and -0 differ)
    tt = ti = None
    if len(table) <= 256:
        tt = torch.from_numpy(table.view(np.float64).copy())
        ti = torch.from_numpy(idx.astype(np.uint8))
    return torch, Mp, cols.reshape(nsl, L, 32), vals.reshape(nsl, L, 32), L, torch.from_numpy(off), ti, tt


def test_slice_records_rebuild_the_regular_slices_of_the_matrix():
    """formats.slice_records (the table behind the implied-columns / implied-values kernels), on host tensors: a slice
    with a record id has exactly the columns row + rec_table[id] and -- when the records carry values -- exactly the
    values rec_vals[id] in all of its 32 rows; ids are ordered by frequency; a variable-coefficient operator keeps
    column records only; an operator whose values alternate from row to row falls back to column records"""
    from learnmultigrid_b200 import problems as P
    N = 256                                      # (a 65-wide coarse grid has no regular slice at all)
    A = F.canonical_csr(P.structured_laplacian_2d(N))
    Q = P.structured_hierarchy_2d(N, 2, "linear")[0]
    A1 = F.canonical_csr(sp.csr_matrix(Q.T @ sp.csc_matrix(A) @ Q))
    for M, want_values in ((A, True), (A1, True), (F.canonical_csr(P.structured_laplacian_2d(N, P.variable_coefficient)), False)):
        torch, Mp, c, v, L, off, ti, tt = _records_case(M)
        assert (ti is not None) == want_values
        ids, rec_table, rec_vals, nreg = F.slice_records(torch, off, ti, tt, L)
        ids, rec_table = ids.numpy(), rec_table.numpy()
        assert (rec_vals is not None) == want_values
        nsl = len(ids)
        reg = ids >= 0
        assert nreg == reg.sum() and (not want_values or nreg > 0.4 * nsl)
        rows = (np.arange(nsl) * 32)[:, None, None] + np.arange(32)[None, None, :]
        rebuilt = rows + rec_table[np.maximum(ids, 0)][:, :L, None].astype(np.int64)
        assert np.array_equal(rebuilt[reg], c[reg])
        col_regular = off.numpy()[:, 0] != F.SLICE_IRREGULAR
        if want_values:
            rv = rec_vals.numpy()[np.maximum(ids, 0)][:, :L, None]
            assert np.array_equal(np.broadcast_to(rv, v.shape)[reg].view(np.int64), v[reg].view(np.int64))   # same bits
            # a slice lost its id only because its rows differ in a value
            lost = col_regular & ~reg
            assert np.all(np.any(v[lost] != v[lost][:, :, :1], axis=(1, 2)))
            assert lost.sum() <= 0.1 * col_regular.sum()
        else:
            assert np.array_equal(reg, col_regular)
        counts = np.bincount(ids[reg], minlength=len(rec_table))
        assert np.all(np.diff(counts) <= 0) and counts.min() >= 1            # most frequent record first, none unused
    # values that alternate from row to row: no slice is regular in its values -> the records stay column records
    torch, Mp, c, v, L, off, ti, tt = _records_case(A)
    lane_parity = torch.from_numpy(np.broadcast_to((np.arange(32) % 2).astype(np.uint8), c.shape).copy().reshape(-1))
    ids2, table2, vals2, nreg2 = F.slice_records(torch, off, lane_parity, torch.tensor([1.0, 2.0], dtype=torch.float64), L)
    assert vals2 is None and nreg2 == int((off.numpy()[:, 0] != F.SLICE_IRREGULAR).sum())


def test_color_list_downloads_an_entry_when_it_is_read():
    """setup_device.ColorList: colours computed on the device stay there until the host asks; entries read through the
    list are NumPy arrays from then on, device() hands out what is stored"""
    import torch
    from learnmultigrid_b200.setup_device import ColorList

    class Counting:
        def __init__(self, t):
            self.t, self.downloads = t, 0

        def cpu(self):
            self.downloads += 1
            return self.t

    dev = Counting(torch.tensor([0, 1, 0, 2], dtype=torch.int32))
    host = np.array([1, 0], dtype=np.int32)
    cl = ColorList([dev, host, None])
    assert cl.device(0) is dev and dev.downloads == 0 and len(cl) == 3
    assert cl[2] is None and cl[1] is host and cl[-2] is host
    assert isinstance(cl[0], np.ndarray) and np.array_equal(cl[0], [0, 1, 0, 2]) and dev.downloads == 1
    assert cl[0] is cl[0] and dev.downloads == 1                     # cached
    assert [None if c is None else c.tolist() for c in cl] == [[0, 1, 0, 2], [1, 0], None]
    assert [None if c is None else len(c) for c in cl[0:2]] == [4, 2]
