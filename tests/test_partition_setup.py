"""Strip-local hierarchy setup (learnmultigrid_b200/partition_setup.py): every rank holds only its row blocks of A_l and
Q_l, fetches the few remote rows it needs and forms its rows of A_{l+1} = Q^T A Q.  Checked against the single-process
product -- SciPy's `csr_matrix(Q.T @ A @ Q)` with A held as CSC, which is what the reference computes
(Multigrid.py:97-98, Solver.py:18) and what the device SpGEMM reproduces -- BIT FOR BIT, values and sparsity patterns,
on virtual ranks (threads) and in a world_size-2 gloo run."""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest
import scipy.sparse as sp

from learnmultigrid_b200 import formats as F
from learnmultigrid_b200 import partition as PT
from learnmultigrid_b200 import partition_setup as PS
from learnmultigrid_b200 import problems as P
from helpers import coo_from, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_ranks(world, fn):
    """fn(fabric_view) on `world` threads (the ThreadFabric of the virtual-rank GPU tests, without a device)"""
    from learnmultigrid_b200.distributed import ThreadFabric
    fab = ThreadFabric(world)
    out, err = [None] * world, [None] * world

    def work(r):
        try:
            out[r] = fn(fab.view(r))
        except BaseException as e:       # noqa: BLE001
            err[r] = e
            fab.abort()
    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for e in err:
        if e is not None and not isinstance(e, threading.BrokenBarrierError):
            raise e
    return out


def global_hierarchy(A, Qs):
    """single-process reference: the reference's expression, level by level"""
    As = [F.canonical_csr(A)]
    cur = sp.csc_matrix(A)
    for Q in Qs:
        Q = F.canonical_csr(Q)
        cur = sp.csr_matrix(Q.T @ cur @ Q)
        cur.sort_indices()
        As.append(F.canonical_csr(cur))
    return As


def same(got, want):
    got, want = sp.csr_matrix(got), sp.csr_matrix(want)
    got.sort_indices()
    want.sort_indices()
    assert got.shape == want.shape
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)   # pattern exact
    assert np.array_equal(got.data, want.data)                                                    # values bit for bit


def strip_run(A, Qs, world):
    A = F.canonical_csr(A)
    Qs = [F.canonical_csr(Q) for Q in Qs]
    ns = [A.shape[0]] + [Q.shape[1] for Q in Qs]
    offs = [PT.block_offsets(n, world) for n in ns]

    def body(fab):
        r = fab.rank
        A_blk = A[offs[0][r]:offs[0][r + 1]]
        Q_blks = [Q[offs[l][r]:offs[l][r + 1]] for l, Q in enumerate(Qs)]
        return PS.build_strip_hierarchy(fab, A_blk, Q_blks, offs)
    return offs, run_ranks(world, body)


CASES = {
    "lap2d_linear": lambda: (P.structured_laplacian_2d(32), P.structured_hierarchy_2d(32, 4, "linear")),
    "varcoef2d_quasi": lambda: (P.structured_laplacian_2d(32, P.variable_coefficient),
                                P.structured_hierarchy_2d(32, 4, "quasi")),
}


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_strip_hierarchy_equals_the_global_product(case, world):
    A, Qs = CASES[case]()
    want = global_hierarchy(A, Qs)
    offs, res = strip_run(A, Qs, world)
    for l in range(len(want)):
        same(sp.vstack([res[r][0][l] for r in range(world)], format="csr"), want[l])
    for l, Q in enumerate(Qs):
        QT = sp.csr_matrix(F.canonical_csr(Q).T)
        QT.sort_indices()
        same(sp.vstack([res[r][1][l] for r in range(world)], format="csr"), QT)


def test_strip_galerkin_1d_c1_three_transfer_types():
    c1 = load_golden("c1_1d_1024.npz")
    A = coo_from(c1, "A")
    for Q in (coo_from(c1, "Q_quasi"), coo_from(c1, "Q_pseudo"), sp.csr_matrix(c1["Q_L2_dense"])):
        want = global_hierarchy(A, [Q])
        offs, res = strip_run(A, [Q], 4)
        same(sp.vstack([res[r][0][1] for r in range(4)], format="csr"), want[1])


def test_fetch_rows_and_more_ranks_than_rows():
    """a rank may own no rows at all (5 rows over 8 ranks) and may need rows from several owners"""
    rng = np.random.default_rng(1)
    M = F.canonical_csr(sp.random(5, 9, density=0.6, random_state=3, format="csr"))
    offs = PT.block_offsets(5, 8)

    def body(fab):
        r = fab.rank
        wanted = np.unique(rng.integers(0, 5, size=4)) if r % 2 == 0 else np.zeros(0, dtype=np.int64)
        wanted = np.array([0, 2, 4]) if r == 3 else wanted
        got = PS.fetch_rows(fab, offs, M[offs[r]:offs[r + 1]], wanted)
        return wanted, got
    for wanted, got in run_ranks(8, body):
        same(got, M[wanted])


def test_only_a_few_grid_lines_are_fetched():
    """the exchange is a halo, not a gather: on a 65^2 grid cut into 4 strips a rank fetches rows of at most two grid
    lines of A and four of Q per neighbour"""
    N = 64
    A = F.canonical_csr(P.structured_laplacian_2d(N))
    Q = F.canonical_csr(P.linear_P_2d(N))
    world = 4
    offs_f, offs_c = PT.block_offsets(A.shape[0], world), PT.block_offsets(Q.shape[1], world)
    W = N + 1
    for r in range(world):
        QT_blk = sp.csr_matrix(Q.T)[offs_c[r]:offs_c[r + 1]]
        F1 = np.unique(QT_blk.indices)
        F2 = np.unique(A[F1].indices)
        for Fset, lines in ((F1, 2), (F2, 4)):
            remote = Fset[(Fset < offs_f[r]) | (Fset >= offs_f[r + 1])]
            assert len(remote) <= 2 * lines * W


WORKER = r"""
import os, sys
import numpy as np, scipy.sparse as sp, torch.distributed as dist
sys.path.insert(0, {root!r})
from learnmultigrid_b200 import formats as F, partition as PT, partition_setup as PS, problems as P
from learnmultigrid_b200.distributed import TorchFabric
dist.init_process_group("gloo")
fab = TorchFabric()
A = F.canonical_csr(P.structured_laplacian_2d(16, P.variable_coefficient))
Qs = [F.canonical_csr(q) for q in P.structured_hierarchy_2d(16, 3, "quasi")]
ns = [A.shape[0]] + [q.shape[1] for q in Qs]
offs = [PT.block_offsets(n, fab.world) for n in ns]
r = fab.rank
A_blks, QT_blks = PS.build_strip_hierarchy(fab, A[offs[0][r]:offs[0][r + 1]],
                                           [q[offs[l][r]:offs[l][r + 1]] for l, q in enumerate(Qs)], offs)
cur = sp.csc_matrix(A)
ok = True
for l, q in enumerate(Qs):
    cur = sp.csr_matrix(q.T @ cur @ q); cur.sort_indices()
    want = F.canonical_csr(cur)[offs[l + 1][r]:offs[l + 1][r + 1]]
    got = sp.csr_matrix(A_blks[l + 1]); got.sort_indices()
    ok = ok and np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices) \
        and np.array_equal(got.data, want.data)
# the composed setup (blocks, first-fit colours coloured in rank order, halo plan, gathered replicated level)
levs, A_rep, Q_rep = PS.strip_local_setup(fab, A[offs[0][r]:offs[0][r + 1]],
                                          [q[offs[l][r]:offs[l][r + 1]] for l, q in enumerate(Qs)], offs, 1, "mcgs")
col, nc = F.greedy_colors(A)
A1 = sp.csr_matrix(Qs[0].T @ sp.csc_matrix(A) @ Qs[0]); A1.sort_indices()
ok = ok and np.array_equal(levs[0].colors, col[offs[0][r]:offs[0][r + 1]]) and levs[0].ncolors == nc \
    and np.array_equal(A_rep.indices, A1.indices) and np.array_equal(A_rep.data, A1.data) \
    and (Q_rep[0] != Qs[1]).nnz == 0
import sys
sys.stdout.write("RANK %d %s\n" % (r, "OK" if ok else "MISMATCH"))      # one write: the two ranks share the pipe
sys.stdout.flush()
dist.destroy_process_group()
"""


def test_two_process_gloo_strip_setup(tmp_path):
    import socket
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    with socket.socket() as sock:                  # a free port: back-to-back runs must not collide in TIME_WAIT
        sock.bind(("127.0.0.1", 0))
        port = str(sock.getsockname()[1])
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=port, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", port, str(script)],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "RANK 0 OK" in out.stdout and "RANK 1 OK" in out.stdout, out.stdout + out.stderr


def test_row_block_generators_equal_slices_of_the_global_operators():
    for N in (8, 16):
        for coef in (None, P.variable_coefficient):
            A = P.structured_laplacian_2d(N, coef)
            n = A.shape[0]
            for r0, r1 in ((0, n), (0, 0), (5, 40), (n - 17, n), (N + 1, 3 * (N + 1))):
                same(P.structured_laplacian_2d(N, coef, rows=(r0, r1)), A[r0:r1])
        Pm = P.linear_P_2d(N)
        for r0, r1 in ((0, Pm.shape[0]), (3, 50), (Pm.shape[0] - 9, Pm.shape[0])):
            same(P.linear_P_2d(N, rows=(r0, r1)), Pm[r0:r1])


def test_hierarchy_from_generated_strips_without_any_global_matrix():
    """every rank generates ITS rows of A_0 and of every P_l, nothing else: the assembled hierarchy equals the global
    one (the precondition for problems larger than one GPU's memory)"""
    N, levels, world = 32, 4, 4
    ns = [(N >> l) + 1 for l in range(levels)]
    offs = [PT.block_offsets(w * w, world) for w in ns]

    def body(fab):
        r = fab.rank
        A_blk = P.structured_laplacian_2d(N, P.variable_coefficient, rows=(offs[0][r], offs[0][r + 1]))
        Q_blks = [P.linear_P_2d(N >> l, rows=(offs[l][r], offs[l][r + 1])) for l in range(levels - 1)]
        return PS.build_strip_hierarchy(fab, A_blk, Q_blks, offs)
    res = run_ranks(world, body)
    want = global_hierarchy(P.structured_laplacian_2d(N, P.variable_coefficient),
                            P.structured_hierarchy_2d(N, levels, "linear"))
    for l in range(levels):
        same(sp.vstack([res[r][0][l] for r in range(world)], format="csr"), want[l])


@pytest.mark.parametrize("world", [2, 3, 5])
def test_halo_plans_from_local_blocks_equal_the_plans_from_global_data(world):
    """partition.RankPlan built from a rank's own row blocks + fetched halo colours = the plan built from the global
    matrices and the global colour array (what DistributedHierarchy uses today), attribute by attribute"""
    N = 16
    A, Qs = P.structured_laplacian_2d(N), P.structured_hierarchy_2d(N, 3, "quasi")
    As = global_hierarchy(A, Qs)
    Qs = [F.canonical_csr(q) for q in Qs]
    QTs = [F.canonical_csr(sp.csr_matrix(q.T)) for q in Qs]
    l = 1                                                                  # a level with a finer and a coarser one
    colors, nc = F.greedy_colors(As[l])
    offs = [PT.block_offsets(a.shape[0], world) for a in As]

    def body(fab):
        r = fab.rank
        A_blk = As[l][offs[l][r]:offs[l][r + 1]]
        QT_blk = QTs[l][offs[l + 1][r]:offs[l + 1][r + 1]]
        Qp_blk = Qs[l - 1][offs[l - 1][r]:offs[l - 1][r + 1]]
        plan = PS.rank_plan_from_blocks(fab, offs[l], offs[l + 1], offs[l - 1], A_blk, QT_blk, Qp_blk,
                                        colors[offs[l][r]:offs[l][r + 1]], nc)
        jac = PS.rank_plan_from_blocks(fab, offs[l], offs[l + 1], offs[l - 1], A_blk, QT_blk, Qp_blk, None, 0)
        return plan, jac
    for r, (got, jac) in enumerate(run_ranks(world, body)):
        ext = PT.level_external_columns(As[l], QTs[l], Qs[l - 1], offs[l], offs[l + 1], offs[l - 1], r)
        want = PT.RankPlan(offs[l], r, ext, colors)
        for name in ("halo_gid", "halo_owner", "halo_color", "perm", "iperm", "color_ptr"):
            assert np.array_equal(getattr(got, name), getattr(want, name)), name
        assert (got.n_own, got.n_halo, got.ncolors, got.neighbours, got.seg, got.seg_color) == \
            (want.n_own, want.n_halo, want.ncolors, want.neighbours, want.seg, want.seg_color)
        wj = PT.RankPlan(offs[l], r, ext, None)
        assert np.array_equal(jac.halo_gid, wj.halo_gid) and jac.seg == wj.seg and jac.perm is None


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_distributed_first_fit_colouring_equals_the_global_one(world):
    """blocks coloured in rank order with the forbidden-colour masks passed along: the colours of
    formats.greedy_colors, for symmetric and unsymmetric patterns, with Dirichlet (diagonal-only) rows, with 19- and
    37-point stencils, and on a random sparse matrix whose couplings cross several blocks"""
    N = 16
    A = P.structured_laplacian_2d(N)
    Qs = P.structured_hierarchy_2d(N, 3, "quasi")
    mats = global_hierarchy(A, Qs)                                     # 5-point + identity rows, 19- and 37-point
    R = sp.random(120, 120, density=0.08, random_state=5, format="csr") + sp.diags((np.arange(120) % 3 > 0) * 1.0)
    mats.append(F.canonical_csr(R))                                    # unsymmetric, long-range, some empty diagonals
    C = sp.lil_matrix((100, 100))
    C[:70, :70] = 1.0                                                  # a 70-clique across the blocks: > 64 colours
    C[70:, 3] = 1.0
    C.setdiag(1.0)
    mats.append(F.canonical_csr(C.tocsr()))
    for M in mats:
        want, nc = F.greedy_colors(M)
        offs = PT.block_offsets(M.shape[0], world)
        res = run_ranks(world, lambda fab: PS.greedy_colors_distributed(fab, offs, M[offs[fab.rank]:offs[fab.rank + 1]]))
        got = np.concatenate([r[0] for r in res])
        assert np.array_equal(got, want)
        assert all(r[1] == nc for r in res)


@pytest.mark.parametrize("smoother", ["mcgs", "jacobi"])
@pytest.mark.parametrize("world", [2, 4])
def test_whole_strip_local_setup_equals_what_the_replicated_setup_derives(world, smoother):
    """strip_local_setup (row blocks in, nothing global held): per partitioned level the operator blocks, first-fit
    colours and halo plans of the replicated setup; below, the gathered first replicated operator and transfers"""
    N, levels, n_dist = 32, 5, 3
    A = P.structured_laplacian_2d(N, P.variable_coefficient)
    Qs = [F.canonical_csr(q) for q in P.structured_hierarchy_2d(N, levels, "linear")]
    As = global_hierarchy(A, Qs)
    QTs = [F.canonical_csr(sp.csr_matrix(q.T)) for q in Qs]
    offs = [PT.block_offsets(a.shape[0], world) for a in As]

    def body(fab):
        r = fab.rank
        A_blk = P.structured_laplacian_2d(N, P.variable_coefficient, rows=(offs[0][r], offs[0][r + 1]))
        Q_blks = [P.linear_P_2d(N >> l, rows=(offs[l][r], offs[l][r + 1])) for l in range(levels - 1)]
        return PS.strip_local_setup(fab, A_blk, Q_blks, offs, n_dist, smoother)
    res = run_ranks(world, body)
    for r, (levs, A_rep, Q_rep) in enumerate(res):
        assert len(levs) == n_dist and len(Q_rep) == levels - 1 - n_dist
        same(A_rep, As[n_dist])
        for k, Q in enumerate(Q_rep):
            same(Q, Qs[n_dist + k])
        for l, lv in enumerate(levs):
            same(lv.A, As[l][offs[l][r]:offs[l][r + 1]])
            same(lv.Q, Qs[l][offs[l][r]:offs[l][r + 1]])
            same(lv.QT, QTs[l][offs[l + 1][r]:offs[l + 1][r + 1]])
            colors = F.greedy_colors(As[l])[0] if smoother == "mcgs" else None
            ext = PT.level_external_columns(As[l], QTs[l], Qs[l - 1] if l else None, offs[l], offs[l + 1],
                                            offs[l - 1] if l else None, r)
            want = PT.RankPlan(offs[l], r, ext, colors)
            for name in ("halo_gid", "halo_owner", "halo_color", "perm", "iperm", "color_ptr"):
                a, b = getattr(lv.plan, name, None), getattr(want, name, None)
                assert (a is None and b is None) or np.array_equal(a, b), (l, name)
            assert (lv.plan.n_own, lv.plan.n_halo, lv.plan.ncolors, lv.plan.seg, lv.plan.seg_color) == \
                (want.n_own, want.n_halo, want.ncolors, want.seg, want.seg_color)
            if smoother == "mcgs":
                assert np.array_equal(lv.colors, colors[offs[l][r]:offs[l][r + 1]])


def test_strip_local_setup_with_row_local_structured_colours():
    """colours given per own row (a structured colouring needs no exchange at all)"""
    N, levels, world = 16, 3, 2
    A = P.structured_laplacian_2d(N)
    Qs = [F.canonical_csr(q) for q in P.structured_hierarchy_2d(N, levels, "linear")]
    As = global_hierarchy(A, Qs)
    offs = [PT.block_offsets(a.shape[0], world) for a in As]
    full = [F.greedy_colors(a)[0] for a in As[:-1]]

    def body(fab):
        r = fab.rank
        own = [full[l][offs[l][r]:offs[l][r + 1]] for l in range(levels - 1)]
        return PS.strip_local_setup(fab, A[offs[0][r]:offs[0][r + 1]],
                                    [q[offs[l][r]:offs[l][r + 1]] for l, q in enumerate(Qs)], offs, 1, "mcgs", own)
    for r, (levs, A_rep, Q_rep) in enumerate(run_ranks(world, body)):
        ext = PT.level_external_columns(As[0], F.canonical_csr(sp.csr_matrix(Qs[0].T)), None, offs[0], offs[1], None, r)
        want = PT.RankPlan(offs[0], r, ext, full[0])
        assert np.array_equal(levs[0].plan.perm, want.perm) and levs[0].plan.seg_color == want.seg_color
        same(A_rep, As[1])
        same(Q_rep[0], Qs[1])


@pytest.mark.parametrize("world,n_dist", [(2, 2), (3, 2), (4, 1), (2, 3)])
def test_partitioned_vcycle_driven_by_the_strip_setup_equals_the_global_oracle(world, n_dist):
    """End to end without any global operator on the ranks: every rank generates its rows, strip_local_setup gives it
    blocks, colours, plans and the replicated tail; the partitioned V-cycle (halo per colour, residual halo, hand-off
    gather, prolongation, halo refresh -- the steps of csrc/cycle.cu, here with the oracle's kernels and send lists
    exchanged the way DistributedHierarchy does) reproduces the single-process oracle cycle BIT FOR BIT.  This is the
    host-side twin of distributed_strip.StripHierarchy."""
    from oracle import kernels as K
    from oracle.vcycle import OracleMultigrid
    N, L, NU = 16, 4, 1
    ns = [((N >> l) + 1) ** 2 for l in range(L)]
    offs = [PT.block_offsets(n, world) for n in ns]
    # the global oracle: verification only
    A0 = P.structured_laplacian_2d(N, P.variable_coefficient)
    Qs = [F.canonical_csr(q) for q in P.structured_hierarchy_2d(N, L, "linear")]
    As = global_hierarchy(A0, Qs)
    cols = [F.greedy_colors(a)[0] for a in As[:-1]] + [None]
    oracle = OracleMultigrid(A0, np.zeros((ns[0], 1)), Qs, smoother="mcgs", colors=cols, hoist_setup=True)
    oracle.build_hierarchy(L)
    rng = np.random.default_rng(1)
    xg, bg = rng.standard_normal(ns[0]), rng.standard_normal(ns[0])
    want, cur = [], xg.reshape(-1, 1).copy()
    for _ in range(2):
        cur = oracle.v_cycle(oracle.matrix, cur, bg.reshape(-1, 1), NU, L)
        want.append(cur.ravel().copy())

    def body(fab):
        r = fab.rank
        A_blk = P.structured_laplacian_2d(N, P.variable_coefficient, rows=(offs[0][r], offs[0][r + 1]))
        Q_blks = [P.linear_P_2d(N >> l, rows=(offs[l][r], offs[l][r + 1])) for l in range(L - 1)]
        strip, A_rep, Q_rep = PS.strip_local_setup(fab, A_blk, Q_blks, offs, n_dist, "mcgs")
        plans = [s.plan for s in strip]
        # send lists, as in DistributedHierarchy: everyone tells the others which rows it reads, in its halo order
        mine = []
        for p in plans:
            d = {}
            for q in p.neighbours:
                s, e = p.seg[q]
                d[q] = (p.halo_gid[s:e], [p.seg_color[(q, c)][0] - s for c in range(p.ncolors)] + [e - s])
            mine.append(d)
        everyone = fab.allgather(mine)
        sends = [{q: (plans[l].local_of_owned(everyone[q][l][r][0]), everyone[q][l][r][1])
                  for q in range(world) if q != r and r in everyone[q][l]} for l in range(n_dist)]

        def positions(l):
            """global id -> position in this rank's level-l vector (partitioned: [own | halo]; replicated: natural)"""
            if l >= n_dist:
                return lambda gid: np.asarray(gid, dtype=np.int64), ns[l]
            g = plans[l].gather_indices()
            order = np.argsort(g)

            def pos(gid):
                k = np.searchsorted(g[order], gid)
                assert np.array_equal(g[order][k], gid)          # every referenced column is owned or in the halo
                return order[k]
            return pos, len(g)

        lev = []
        for l in range(n_dist):
            p, s = plans[l], strip[l]
            pos, nv = positions(l)
            posn, nvn = positions(l + 1)
            own = p.gather_indices()[:p.n_own] - p.o0                       # block rows in colour-blocked order
            if l + 1 < n_dist:
                pn = plans[l + 1]
                own_next = pn.gather_indices()[:pn.n_own] - pn.o0
            else:
                own_next = np.arange(offs[l + 1][r + 1] - offs[l + 1][r])

            def local(M, rows, colpos, ncols):
                B = sp.csr_matrix(M)[rows]
                B.sort_indices()
                return F.raw_csr(B.indptr, colpos(B.indices.astype(np.int64)).astype(np.int32), B.data,
                                 (len(rows), ncols))
            lev.append({"A": local(s.A, own, pos, nv), "Q": local(s.Q, own, posn, nvn),
                        "QT": local(s.QT, own_next, pos, nv), "p": p, "nv": nv})

        def exchange(l, v, color=None):
            p = lev[l]["p"]
            out = {}
            for q, (idx, ptr) in sends[l].items():
                a, b = (0, len(idx)) if color is None else (ptr[color], ptr[color + 1])
                out[q] = v[idx[a:b]].copy()
            got = fab.allgather(out)
            for q in p.neighbours:
                s0, s1 = p.seg[q] if color is None else p.seg_color[(q, color)]
                v[p.n_own + s0:p.n_own + s1] = got[q][r]

        tail_colors = [F.greedy_colors(A_rep)[0]]
        tail = OracleMultigrid(A_rep, np.zeros((ns[n_dist], 1)), Q_rep, smoother="mcgs", colors=tail_colors,
                               hoist_setup=True)
        for a in tail.build_hierarchy(L - n_dist)[1:-1]:
            tail_colors.append(F.greedy_colors(F.canonical_csr(a))[0])

        def smooth(l, x, b):
            d = lev[l]
            p = d["p"]
            Asq = sp.vstack([d["A"], sp.csr_matrix((p.n_halo, d["nv"]))]).tocsr()
            Asq = F.raw_csr(Asq.indptr, Asq.indices, Asq.data, Asq.shape)
            bb = np.concatenate([b, np.zeros(p.n_halo)])
            for _ in range(NU):
                for c in range(p.ncolors):
                    K.gauss_seidel_multicolor(Asq, x, bb, [np.arange(p.color_ptr[c], p.color_ptr[c + 1], dtype=np.int32)])
                    exchange(l, x, c)

        def vcycle(l, x, b):
            d = lev[l]
            p = d["p"]
            smooth(l, x, b)
            res = np.zeros(d["nv"])
            res[:p.n_own] = K.residual(d["A"], x, b)
            exchange(l, res)
            rc = K.spmv(d["QT"], res)
            if l + 1 < n_dist:
                e = np.zeros(lev[l + 1]["nv"])
                vcycle(l + 1, e, rc)
            else:
                bc = np.concatenate(fab.allgather(rc)).reshape(-1, 1)       # hand-off: every rank gets the full rhs
                if L - (l + 1) >= 2:
                    e = tail.v_cycle(tail.matrix, np.zeros_like(bc), bc, NU, L - (l + 1)).ravel()
                else:
                    e = tail._lu.solve(bc.ravel())                          # the tail is the coarsest level alone
            x[:p.n_own] = K.prolong_correct(d["Q"], e, x[:p.n_own])
            exchange(l, x)
            smooth(l, x, b)

        g0 = plans[0].gather_indices()
        x, b = xg[g0].copy(), bg[g0[:plans[0].n_own]].copy()
        out = []
        for _ in range(2):
            vcycle(0, x, b)
            out.append((g0[:plans[0].n_own].copy(), x[:plans[0].n_own].copy()))
        return out

    res = run_ranks(world, body)
    for cyc in range(2):
        got = np.empty(ns[0])
        for rank_out in res:
            gid, val = rank_out[cyc]
            got[gid] = val
        assert np.array_equal(got, want[cyc])


def _rect_mesh(Nx, Ny):
    """Nx x Ny square cells of side 1/Nx with the reference's node numbering and triangle split (Mesh2D.py:63-93)"""
    from learnmultigrid_b200.mesh.Mesh2D import Mesh2D
    W, h = Nx + 1, 1.0 / Nx
    iy, ix = np.divmod(np.arange(W * (Ny + 1)), W)
    k = (np.arange(Ny)[:, None] * W + np.arange(Nx)[None, :]).reshape(-1)
    conn = np.stack([np.stack([k, k + 1, k + W + 1], axis=1), np.stack([k, k + W + 1, k + W], axis=1)], axis=1)
    return Mesh2D(p=np.stack([ix * h, iy * h], axis=1), conn=conn.reshape(-1, 3))


@pytest.mark.parametrize("Nx,Ny", [(4, 8), (8, 4)])
def test_rectangular_strip_generators_match_the_assembly(Nx, Ny):
    """structured_laplacian_2d / structured_rhs_2d / linear_P_2d with Ny != N (the stacked strips of a weak-scaling
    run): the operator and load vector equal the P1 assembly on that mesh, row blocks equal slices, the transfer
    reproduces linear functions, and the strip-local Galerkin hierarchy equals the global one"""
    from learnmultigrid_b200.assembly.StiffnessMatrix import StiffnessMatrix
    from learnmultigrid_b200.assembly.LoadVector import LoadVector
    from learnmultigrid_b200.assembly.LoadFunction import LoadFunction
    from learnmultigrid_b200.assembly.Quadrature import Quadrature2D
    from learnmultigrid_b200.assembly.ShapeFunction import GradientTriangle, FunctionTriangle
    mesh = _rect_mesh(Nx, Ny)
    W = Nx + 1
    n = W * (Ny + 1)
    iy, ix = np.divmod(np.arange(n), W)
    bnd = np.flatnonzero((ix == 0) | (ix == Nx) | (iy == 0) | (iy == Ny))
    for coef in (None, P.variable_coefficient):
        As = sp.lil_matrix(StiffnessMatrix(mesh).compute_stiffness_2d(GradientTriangle(1), Quadrature2D(3), format="csr",
                                                                      coefficient=coef))
        for b in bnd:
            As[b, :] = 0
            As[b, b] = 1.0
        As = sp.csr_matrix(As)
        As.data[abs(As.data) < 1e-13] = 0          # hypotenuse couplings: exact zeros in the generator
        As.eliminate_zeros()
        As.sort_indices()
        G = P.structured_laplacian_2d(Nx, coef, Ny=Ny)
        assert np.array_equal(As.indptr, G.indptr) and np.array_equal(As.indices, G.indices)
        np.testing.assert_allclose(G.data, As.data, rtol=1e-13, atol=1e-14)
        for r0, r1 in ((0, n), (3, 29), (n - 11, n)):
            same(P.structured_laplacian_2d(Nx, coef, rows=(r0, r1), Ny=Ny), G[r0:r1])
    rhs = LoadVector(mesh).compute_rhs_2d(LoadFunction(lambda q: -1.0), FunctionTriangle(1), Quadrature2D(3))
    rhs[bnd] = 0
    np.testing.assert_allclose(P.structured_rhs_2d(Nx, Ny=Ny), rhs, rtol=1e-13, atol=1e-18)
    assert np.array_equal(P.structured_rhs_2d(Nx, rows=(5, 31), Ny=Ny), P.structured_rhs_2d(Nx, Ny=Ny)[5:31])
    Pm = P.linear_P_2d(Nx, Nyf=Ny)
    Wc = Nx // 2 + 1
    cy, cx = np.divmod(np.arange(Wc * (Ny // 2 + 1)), Wc)
    np.testing.assert_allclose(Pm @ (3.0 + 2.0 * cx - 5.0 * cy), 3.0 + ix - 2.5 * iy, rtol=0, atol=1e-13)
    same(P.linear_P_2d(Nx, rows=(7, 33), Nyf=Ny), Pm[7:33])
    Qs = P.structured_hierarchy_2d(Nx, 3, "linear", Ny=Ny)
    want = global_hierarchy(G, Qs)
    offs, res = strip_run(G, Qs, 3)
    for l in range(3):
        same(sp.vstack([res[r][0][l] for r in range(3)], format="csr"), want[l])


@pytest.mark.parametrize("seed,world", [(0, 2), (1, 3), (2, 5), (3, 7)])
def test_strip_hierarchy_on_random_unstructured_operators(seed, world):
    """no grid structure at all: random sparse A (unsymmetric, empty rows) and random transfers with empty rows and
    columns; couplings cross every block boundary, some ranks fetch from all others; two levels of Galerkin products,
    transposes and the composed setup (colours, plans) against the global computation"""
    rng = np.random.default_rng(seed)
    n0, n1, n2 = 230 + 17 * seed, 71 + 5 * seed, 19 + seed
    A = sp.random(n0, n0, density=0.03, random_state=seed, format="csr") + sp.diags((np.arange(n0) % 7 > 0) * 3.0)
    Q0 = sp.random(n0, n1, density=0.05, random_state=seed + 10, format="csr")
    Q1 = sp.random(n1, n2, density=0.15, random_state=seed + 20, format="csr")
    A, Qs = F.canonical_csr(A), [F.canonical_csr(Q0), F.canonical_csr(Q1)]
    want = global_hierarchy(A, Qs)
    offs, res = strip_run(A, Qs, world)
    for l in range(3):
        same(sp.vstack([res[r][0][l] for r in range(world)], format="csr"), want[l])
    for l, Q in enumerate(Qs):
        same(sp.vstack([res[r][1][l] for r in range(world)], format="csr"), F.canonical_csr(sp.csr_matrix(Q.T)))

    def body(fab):
        r = fab.rank
        return PS.strip_local_setup(fab, A[offs[0][r]:offs[0][r + 1]],
                                    [q[offs[l][r]:offs[l][r + 1]] for l, q in enumerate(Qs)], offs, 2, "mcgs")
    QTs = [F.canonical_csr(sp.csr_matrix(q.T)) for q in Qs]
    for r, (levs, A_rep, Q_rep) in enumerate(run_ranks(world, body)):
        same(A_rep, want[2])
        assert Q_rep == []
        for l, lv in enumerate(levs):
            colors = F.greedy_colors(want[l])[0]
            assert np.array_equal(lv.colors, colors[offs[l][r]:offs[l][r + 1]])
            ext = PT.level_external_columns(want[l], QTs[l], Qs[l - 1] if l else None, offs[l], offs[l + 1],
                                            offs[l - 1] if l else None, r)
            ref = PT.RankPlan(offs[l], r, ext, colors)
            assert np.array_equal(lv.plan.halo_gid, ref.halo_gid) and lv.plan.seg_color == ref.seg_color
            assert np.array_equal(lv.plan.perm, ref.perm)
