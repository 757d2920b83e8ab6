"""GPU: the tail program (csrc/tail.cu, one persistent cooperative kernel for the recorded operations of the small
levels) against the launch-per-operation path of the same cycle: iterates BIT FOR BIT, eager and captured in a CUDA
graph, for every threshold from "coarsest smoothed level only" to "the whole cycle"; then against the CPU oracle.
The tail program is off by default (mg_set_tail_max_rows)."""
import ctypes

import numpy as np
import pytest

from helpers import assert_history_close, bilinear_P, poisson2d

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    import torch
    from learnmultigrid_b200 import _lib
    assert torch.cuda.is_available()
    return _lib.load()


@pytest.fixture()
def tail(lib):
    """sets the threshold for one test and restores 'off' afterwards"""
    def set_rows(rows, ctas=2):
        lib.mg_set_tail_max_rows(int(rows))
        lib.mg_set_tail_ctas_per_sm(int(ctas))
    yield set_rows
    lib.mg_set_tail_max_rows(0)
    lib.mg_set_tail_ctas_per_sm(2)


def stats(lib):
    o, b, l = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
    lib.mg_tail_last_stats(ctypes.byref(o), ctypes.byref(b), ctypes.byref(l))
    return o.value, b.value, l.value


def run(h, params, x0, b, cycles, use_graph):
    h.set_rhs(b)
    h.set_x(x0)
    xs = []
    for _ in range(cycles):
        h.vcycle(params, use_graph=use_graph)
        xs.append(h.get_x().copy())
    return xs, h.last_launches


@pytest.mark.parametrize("smoother,omega", [("mcgs", 1.0), ("jacobi", 2.0 / 3.0)])
@pytest.mark.parametrize("nu", [1, 2, 3])
def test_tail_program_is_bit_identical_to_the_launch_per_operation_cycle(lib, tail, smoother, omega, nu):
    from learnmultigrid_b200.engine import DeviceHierarchy
    N = 64
    A = poisson2d(N)
    Qs = [bilinear_P(N >> k) for k in range(4)]          # 4225 / 1089 / 289 / 81 / 25 rows
    rng = np.random.default_rng(11)
    n = A.shape[0]
    b, x0 = rng.standard_normal(n), rng.standard_normal(n)
    h = DeviceHierarchy(A, Qs, smoother=smoother)
    params = h.make_params(nu_pre=nu, nu_post=nu, omega=omega)
    want, launches_plain = run(h, params, x0, b, 3, use_graph=False)
    for rows in (100, 300, 1100, 5000):                   # tail = levels 3.. / 2.. / 1.. / 0.. (the whole cycle)
        tail(rows)
        for use_graph in (False, True):
            got, launches = run(h, params, x0, b, 3, use_graph=use_graph)
            for g, w in zip(got, want):
                assert np.array_equal(g, w), (rows, use_graph)
            assert launches < launches_plain
        if rows == 5000:
            h.vcycle(params, use_graph=False)
            ops, barriers, tl = stats(lib)
            assert 2 <= tl <= -(-ops // 40) + 2           # one stretch down, one up, <= 40 operations per launch
            assert launches_plain - 2 <= ops <= launches_plain and barriers < ops
        tail(0)
    again, launches = run(h, params, x0, b, 3, use_graph=True)
    assert launches == launches_plain and all(np.array_equal(g, w) for g, w in zip(again, want))


def test_tail_program_long_rows_and_many_operations(lib, tail):
    """quasi-L2 transfers: 19-/37-point Galerkin stencils (rows longer than one chunk of 8 entries, non-uniform slice
    lengths), 9+ colours per level, V(3,3): more than 40 operations per stretch, i.e. several cooperative launches"""
    from learnmultigrid_b200 import problems as P
    from learnmultigrid_b200.engine import DeviceHierarchy
    N = 64
    A = P.structured_laplacian_2d(N)
    Qs = P.structured_hierarchy_2d(N, 4, transfer="quasi")
    b = P.structured_rhs_2d(N).ravel()
    rng = np.random.default_rng(5)
    x0 = rng.standard_normal(A.shape[0])
    h = DeviceHierarchy(A, Qs, smoother="mcgs")
    params = h.make_params(nu_pre=3, nu_post=3)
    want, launches_plain = run(h, params, x0, b, 2, use_graph=False)
    for ctas in (1, 2, 4):
        tail(10 ** 9, ctas)
        for use_graph in (False, True):
            got, launches = run(h, params, x0, b, 2, use_graph=use_graph)
            assert all(np.array_equal(g, w) for g, w in zip(got, want)), (ctas, use_graph)
        h.vcycle(params, use_graph=False)
        ops, barriers, tl = stats(lib)
        assert ops > 80 and tl >= 3 and launches < launches_plain / 4


def test_tail_program_through_the_api_matches_the_oracle(lib, tail):
    """SemiGeometricMG.solve with the tail program on: same iteration count and history as the CPU oracle"""
    from learnmultigrid_b200.solvers.Multigrid import SemiGeometricMG
    from oracle.vcycle import OracleMultigrid
    N = 32
    A = poisson2d(N)
    Qs = [bilinear_P(N), bilinear_P(N // 2)]
    rng = np.random.default_rng(2)
    rhs = rng.standard_normal((A.shape[0], 1))
    tail(10 ** 9)
    mg = SemiGeometricMG(A, rhs, Qs)
    mg.solve(levels=3, smoother="GaussSeidel", smooth_steps=2, error=1e-10, max_iterations=40)
    o = OracleMultigrid(A, rhs, Qs, smoother="mcgs", colors=mg.get_hierarchy().colors, hoist_setup=True)
    o.solve(levels=3, smooth_steps=2, error=1e-10, max_iterations=40)
    assert mg.get_iterations() == len(o.track_res) < 40
    assert_history_close(mg.track_res, o.track_res, A, o.solution)


def test_partitioned_cycle_with_a_replicated_tail_program(lib, tail):
    """row-partitioned fine levels (virtual ranks on one GPU), replicated coarse levels run as a tail program:
    bit-identical to the single-GPU cycle"""
    from learnmultigrid_b200.engine import DeviceHierarchy
    from learnmultigrid_b200.distributed import DistributedHierarchy, run_virtual_ranks
    N = 32
    A = poisson2d(N)
    Qs = [bilinear_P(N), bilinear_P(N // 2), bilinear_P(N // 4)]
    rng = np.random.default_rng(3)
    n = A.shape[0]
    b, x0 = rng.standard_normal(n), rng.standard_normal(n)
    h = DeviceHierarchy(A, Qs, smoother="mcgs")
    params = h.make_params(nu_pre=1, nu_post=1)
    want, _ = run(h, params, x0, b, 3, use_graph=False)
    colors = h.colors
    tail(10 ** 9)

    def body(fab):
        d = DistributedHierarchy(A, Qs, fab, smoother="mcgs", colors=colors, n_dist=1, region_bytes=1 << 20,
                                 max_sites=256, timeout_s=10.0)
        d.set_rhs(b)
        d.set_x(x0)
        p = d.make_params(nu_pre=1, nu_post=1)
        xs = []
        for it in range(3):
            d.vcycle(p, use_graph=(it > 0))
            xs.append(d.get_x().copy())
        d.check()
        d.close()
        return xs

    for xs in run_virtual_ranks(2, body):
        assert all(np.array_equal(g, w) for g, w in zip(xs, want))
