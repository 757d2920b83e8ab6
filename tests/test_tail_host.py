"""The tail program (csrc/tail.cu: recorder in cycle.cu, dependence analysis that places the grid barriers, per-row
arithmetic) run on HOST arrays through mg_host_tail_vcycle, against the CPU oracle and against the NumPy mirror of the
launch-per-operation cycle (helpers.emulate_vcycle).  The CUDA kernel executes the same __host__ __device__ row code;
tests/test_gpu_tail.py checks it on the GPU against the launch-per-operation path, bit for bit."""
import ctypes

import numpy as np
import pytest
import scipy.sparse as sp

from learnmultigrid_b200 import _lib
from learnmultigrid_b200 import formats as F
from oracle.vcycle import OracleMultigrid
from helpers import bilinear_P, coo_from, emulate_vcycle, load_golden, poisson2d


def vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class HostHierarchy:
    """mg_level array over the host-level dicts of formats.build_host_hierarchy (all pointers are NumPy buffers)"""

    def __init__(self, A, Qs, smoother):
        self.host = F.build_host_hierarchy(A, Qs, smoother)
        L = len(self.host)
        self.keep = []
        self.arr = (_lib.mg_level * L)()
        self.vec = []
        for l, d in enumerate(self.host):
            s = self.arr[l]
            n = d["n"]
            s.n = n
            v = {k: np.zeros(n) for k in ("x", "b", "r", "tmp")}
            self.vec.append(v)
            s.d_x, s.d_b, s.d_r, s.d_tmp = (vp(v[k]) for k in ("x", "b", "r", "tmp"))
            if l < L - 1:
                for name, field, shape in (("A_sell", "A", (n, n)), ("Q_sell", "Q", (n, self.host[l + 1]["n"])),
                                           ("QT_sell", "QT", (self.host[l + 1]["n"], n))):
                    sptr, cols, vals = (np.ascontiguousarray(a) for a in d[name])
                    sptr = sptr.astype(np.int64)
                    cols = cols.astype(np.int32)
                    vals = vals.astype(np.float64)
                    self.keep += [sptr, cols, vals]
                    lens = np.diff(sptr) // 32
                    m = _lib.mg_sell(shape[0], shape[1], len(sptr) - 1, vp(sptr), vp(cols), vp(vals),
                                     int(lens.max()) if len(lens) else 0,
                                     int(lens[0]) if len(lens) and np.all(lens == lens[0]) else 0)
                    setattr(s, field, m)
                dinv = np.ascontiguousarray(d["dinv"], dtype=np.float64)
                self.keep.append(dinv)
                s.d_dinv = vp(dinv)
                if d.get("color_ptr") is not None:
                    cp = (ctypes.c_int64 * len(d["color_ptr"]))(*[int(c) for c in d["color_ptr"]])
                    self.keep.append(cp)
                    s.ncolors = len(d["color_ptr"]) - 1
                    s.h_color_ptr = ctypes.cast(cp, ctypes.POINTER(ctypes.c_int64))
            else:
                inv = np.ascontiguousarray(np.linalg.inv(d["A_nat"].toarray()))
                self.keep.append(inv)
                s.coarse_kind = _lib.MG_COARSE_DENSE
                s.d_coarse_inv = vp(inv)

    def cycle(self, smoother, nu, omega, x, b, shuffle=0, zero_guess_skip=True):
        lib = _lib.load()
        perm = self.host[0]["perm"]
        self.vec[0]["x"][:] = x if perm is None else x[perm]
        self.vec[0]["b"][:] = b if perm is None else b[perm]
        sm = {"jacobi": _lib.MG_SMOOTH_JACOBI, "mcgs": _lib.MG_SMOOTH_MCGS}[smoother]
        params = _lib.mg_cycle_params(sm, nu, nu, omega, 1 if zero_guess_skip else 0, 0)
        _lib.check(lib.mg_host_tail_vcycle(self.arr, len(self.host), ctypes.byref(params), shuffle),
                   "mg_host_tail_vcycle")
        out = self.vec[0]["x"].copy()
        if perm is not None:
            nat = np.empty_like(out)
            nat[perm] = out
            out = nat
        return out

    def stats(self):
        lib = _lib.load()
        o, b, l = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        lib.mg_tail_last_stats(ctypes.byref(o), ctypes.byref(b), ctypes.byref(l))
        return o.value, b.value, l.value


def problem(N=16, levels=3, seed=7):
    A = poisson2d(N)
    Qs = [bilinear_P(N >> k) for k in range(levels - 1)]
    rng = np.random.default_rng(seed)
    return A, Qs, rng.standard_normal(A.shape[0]), rng.standard_normal(A.shape[0])


@pytest.mark.parametrize("smoother,omega", [("jacobi", 2.0 / 3.0), ("jacobi", 1.0), ("mcgs", 1.0)])
@pytest.mark.parametrize("nu", [1, 2, 3])
def test_tail_program_reproduces_the_cycle(smoother, omega, nu):
    A, Qs, b, x0 = problem()
    h = HostHierarchy(A, Qs, smoother)
    got = h.cycle(smoother, nu, omega, x0, b)
    # (1) the NumPy mirror of the launch-per-operation cycle, same coarse solve: bit for bit
    perm = h.host[0]["perm"]
    inv = np.linalg.inv(h.host[-1]["A_nat"].toarray())

    mirror = emulate_vcycle(h.host, smoother, nu, nu, omega, x0 if perm is None else x0[perm],
                            b if perm is None else b[perm], coarse_solve=lambda Ac, rc: _rowwise(inv, rc))
    if perm is not None:
        nat = np.empty_like(mirror)
        nat[perm] = mirror
        mirror = nat
    assert np.array_equal(got, mirror)
    # (2) the oracle V-cycle (SciPy Galerkin, spsolve): 1e-12 relative per cycle
    colors = [d["colors"] for d in h.host]
    o = OracleMultigrid(A, b.reshape(-1, 1), Qs, smoother=smoother, omega=omega, colors=colors, hoist_setup=True)
    o.build_hierarchy(len(Qs) + 1)
    want = o.v_cycle(o.matrix, x0.reshape(-1, 1).copy(), b.reshape(-1, 1), nu, len(Qs) + 1).ravel()
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12 * np.linalg.norm(want))


def _rowwise(inv, rc):
    """x_i = sum_j inv[i,j] * r_j added left to right (what the host mode of the cycle does)"""
    out = np.empty(len(rc))
    for i in range(len(rc)):
        acc = 0.0
        for j in range(len(rc)):
            acc += inv[i, j] * rc[j]
        out[i] = acc
    return out


@pytest.mark.parametrize("smoother,omega,skip", [("jacobi", 2.0 / 3.0, True), ("jacobi", 2.0 / 3.0, False),
                                                 ("mcgs", 1.0, True)])
def test_barrier_placement_survives_shuffled_execution(smoother, omega, skip):
    """inside a barrier-free group of operations the rows may run in ANY order (on the GPU: concurrently); a missing
    barrier shows up as a different result under a shuffled order"""
    A, Qs, b, x0 = problem(N=16, levels=4, seed=3)
    h = HostHierarchy(A, Qs, smoother)
    ref = h.cycle(smoother, 2, omega, x0, b, shuffle=0, zero_guess_skip=skip)
    ops, barriers, launches = h.stats()
    assert launches == 0 and ops > 0                         # host mode launches nothing
    assert barriers < ops - 1                                # at least one pair of operations shares a group
    for seed in (1, 2, 3, 12345):
        assert np.array_equal(h.cycle(smoother, 2, omega, x0, b, shuffle=seed, zero_guess_skip=skip), ref)


def test_shuffled_execution_detects_a_missing_barrier():
    """the check above has teeth: with the barriers removed on purpose, a shuffled order changes the result"""
    lib = _lib.load()
    A, Qs, b, x0 = problem(N=16, levels=3, seed=5)
    h = HostHierarchy(A, Qs, "mcgs")
    ref = h.cycle("mcgs", 1, 1.0, x0, b)
    old = lib.mg_tail_debug_drop_barriers(1)
    try:
        assert np.array_equal(h.cycle("mcgs", 1, 1.0, x0, b, shuffle=0), ref)     # serial order hides the bug
        assert any(not np.array_equal(h.cycle("mcgs", 1, 1.0, x0, b, shuffle=s), ref) for s in (1, 2, 3))
    finally:
        lib.mg_tail_debug_drop_barriers(old)
    assert np.array_equal(h.cycle("mcgs", 1, 1.0, x0, b, shuffle=1), ref)


def test_operation_and_barrier_counts():
    """3 levels, V(1,1), multicolour GS with c_l colours: per non-coarsest level c_l + 1 (residual) + 1 (restriction) on
    the way down, 1 (prolongation) + c_l on the way up, + 1 fill of the zero coarse guess on levels > 0"""
    A, Qs, b, x0 = problem(N=16, levels=3)
    h = HostHierarchy(A, Qs, "mcgs")
    h.cycle("mcgs", 1, 1.0, x0, b)
    ops, barriers, _ = h.stats()
    c = [len(d["color_ptr"]) - 1 for d in h.host[:-1]]
    assert ops == sum(2 * k + 3 for k in c) + 1
    # two stretches (down to the coarsest solve, back up), a barrier between consecutive operations of a stretch --
    # except between the restriction into level 1 and the fill of level 1's iterate, which do not depend on each other
    assert barriers == (ops - 2) - 1
