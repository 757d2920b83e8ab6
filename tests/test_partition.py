"""Host logic of the multi-GPU path: row-block partition and halo plans (NumPy), including a world_size-2 gloo run
in which two processes build their plans independently, exchange halos with torch.distributed and reproduce the
single-process oracle sweep."""
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from learnmultigrid_b200 import formats as F
from learnmultigrid_b200 import partition as PT
from helpers import free_port, poisson2d

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_block_offsets_cover_rows():
    for n, size in ((10, 3), (67125249, 8), (5, 8)):
        o = PT.block_offsets(n, size)
        assert o[0] == 0 and o[-1] == n and np.all(np.diff(o) >= 0) and len(o) == size + 1
    assert list(PT.owner_of(PT.block_offsets(10, 3), np.array([0, 2, 3, 5, 6, 9]))) == [0, 0, 1, 1, 2, 2]


def plans_for(A, size, colors):
    offs = PT.block_offsets(A.shape[0], size)
    plans = []
    for r in range(size):
        ext = PT.external_columns(A.indptr, A.indices, offs[r], offs[r + 1], offs[r], offs[r + 1])
        plans.append(PT.RankPlan(offs, r, ext, colors))
    return offs, plans


def test_halo_plan_is_consistent_between_neighbours():
    A = F.canonical_csr(poisson2d(12))
    colors, nc = F.greedy_colors(A)
    offs, plans = plans_for(A, 3, colors)
    for p in plans:
        assert p.n_own == offs[p.rank + 1] - offs[p.rank]
        assert set(p.neighbours) <= {p.rank - 1, p.rank + 1}
        # halo ordered by owner, then colour, then gid
        key = list(zip(p.halo_owner, p.halo_color, p.halo_gid))
        assert key == sorted(key)
        g = p.gather_indices()
        assert len(g) == p.n_own + p.n_halo and len(set(g)) == len(g)
        # colour blocks of the owned part
        own_colors = colors[g[:p.n_own]]
        assert np.all(np.diff(own_colors) >= 0)
        assert np.array_equal(np.bincount(own_colors, minlength=nc), np.diff(p.color_ptr))
    for s in plans:
        for r in plans:
            if s.rank == r.rank:
                continue
            idx, ptr = s.send_indices(r)
            if s.rank not in r.seg:
                assert len(idx) == 0
                continue
            a, b = r.seg[s.rank]
            # what the sender reads is exactly what the receiver expects, position by position
            sent_gids = s.gather_indices()[idx]
            assert np.array_equal(sent_gids, r.halo_gid[a:b])
            for c in range(nc):
                ca, cb = r.seg_color[(s.rank, c)]
                assert (ptr[c], ptr[c + 1]) == (ca - a, cb - a)
                assert np.all(colors[sent_gids[ptr[c]:ptr[c + 1]]] == c)


def test_partitioned_sweeps_match_global_oracle():
    """emulate 4 ranks in NumPy: local matrices with remapped columns + halo exchange after every colour reproduce the
    global multicolour Gauss-Seidel sweep bit for bit"""
    from oracle import kernels as K
    A = F.canonical_csr(poisson2d(16))
    n = A.shape[0]
    colors, nc = F.greedy_colors(A)
    offs, plans = plans_for(A, 4, colors)
    rng = np.random.default_rng(0)
    x, b = rng.standard_normal(n), rng.standard_normal(n)
    perm, cptr = F.color_permutation(colors)
    want = x.copy()
    K.gauss_seidel_multicolor(A, want, b, [perm[cptr[c]:cptr[c + 1]] for c in range(nc)], iterations=2)
    loc = []
    for p in plans:
        g = p.gather_indices()
        slot = -np.ones(n, dtype=np.int64)
        slot[g] = np.arange(len(g))
        Ab = A[g[:p.n_own]]                                    # owned rows in colour-blocked order
        Al = F.raw_csr(Ab.indptr, slot[Ab.indices], Ab.data, (p.n_own, len(g)))
        loc.append({"A": Al, "x": x[g].copy(), "b": b[g[:p.n_own]].copy(), "g": g})
    for _ in range(2):
        for c in range(nc):
            for p, d in zip(plans, loc):
                rows = np.arange(p.color_ptr[c], p.color_ptr[c + 1], dtype=np.int32)
                # square view of the local operator: pad rows so that the oracle kernel can index halo columns
                Asq = sp.vstack([d["A"], sp.csr_matrix((p.n_halo, d["A"].shape[1]))]).tocsr()
                bb = np.concatenate([d["b"], np.zeros(p.n_halo)])
                K.gauss_seidel_multicolor(F.raw_csr(Asq.indptr, Asq.indices, Asq.data, Asq.shape), d["x"], bb, [rows])
            for r, dr in zip(plans, loc):                     # exchange colour c
                for s, ds in zip(plans, loc):
                    if s.rank in r.seg:
                        idx, ptr = s.send_indices(r)
                        a = r.seg[s.rank][0]
                        dr["x"][r.n_own + a + ptr[c]: r.n_own + a + ptr[c + 1]] = ds["x"][idx[ptr[c]:ptr[c + 1]]]
    got = np.empty(n)
    for p, d in zip(plans, loc):
        got[d["g"][:p.n_own]] = d["x"][:p.n_own]
    assert np.array_equal(got, want)


WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["MGB_ROOT"])
sys.path.insert(0, os.path.join(os.environ["MGB_ROOT"], "tests"))
from learnmultigrid_b200 import formats as F, partition as PT
from oracle import kernels as K
from helpers import poisson2d
import scipy.sparse as sp

dist.init_process_group("gloo")
rank, size = dist.get_rank(), dist.get_world_size()
A = F.canonical_csr(poisson2d(16)); n = A.shape[0]
colors, nc = F.greedy_colors(A)
offs = PT.block_offsets(n, size)
def plan(r):
    return PT.RankPlan(offs, r, PT.external_columns(A.indptr, A.indices, offs[r], offs[r+1], offs[r], offs[r+1]), colors)
me = plan(rank)
nbr = {p: plan(p) for p in me.neighbours}
rng = np.random.default_rng(0)
x, b = rng.standard_normal(n), rng.standard_normal(n)
g = me.gather_indices()
slot = -np.ones(n, dtype=np.int64); slot[g] = np.arange(len(g))
Ab = A[g[:me.n_own]]
Asq = sp.vstack([F.raw_csr(Ab.indptr, slot[Ab.indices], Ab.data, (me.n_own, len(g))), sp.csr_matrix((me.n_halo, len(g)))]).tocsr()
Asq = F.raw_csr(Asq.indptr, Asq.indices, Asq.data, Asq.shape)
xl = x[g].copy(); bl = np.concatenate([b[g[:me.n_own]], np.zeros(me.n_halo)])
for sweep in range(2):
    for c in range(nc):
        rows = np.arange(me.color_ptr[c], me.color_ptr[c+1], dtype=np.int32)
        K.gauss_seidel_multicolor(Asq, xl, bl, [rows])
        reqs, recv = [], {}
        for p in me.neighbours:
            idx, ptr = me.send_indices(nbr[p])
            out = torch.from_numpy(np.ascontiguousarray(xl[idx[ptr[c]:ptr[c+1]]]))
            a0, a1 = me.seg_color[(p, c)]
            recv[p] = torch.empty(a1 - a0, dtype=torch.float64)
            if out.numel(): reqs.append(dist.isend(out, p))
            if recv[p].numel(): reqs.append(dist.irecv(recv[p], p))
        for q in reqs: q.wait()
        for p in me.neighbours:
            a0, a1 = me.seg_color[(p, c)]
            xl[me.n_own + a0: me.n_own + a1] = recv[p].numpy()
# compare the owned part with the global oracle
perm, cptr = F.color_permutation(colors)
want = x.copy()
K.gauss_seidel_multicolor(A, want, b, [perm[cptr[c]:cptr[c+1]] for c in range(nc)], iterations=2)
ok = np.array_equal(xl[:me.n_own], want[g[:me.n_own]])
t = torch.tensor([1 if ok else 0]); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0: print("PARTITION_OK" if int(t.item()) == 1 else "PARTITION_MISMATCH")
dist.destroy_process_group()
'''


def test_two_process_gloo_halo_exchange(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MGB_ROOT=ROOT, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", free_port(), str(script)],
                         env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=240)
    assert "PARTITION_OK" in out.stdout, out.stdout[-3000:]


VCYCLE_WORKER = r'''
import os, sys
import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["MGB_ROOT"])
sys.path.insert(0, os.path.join(os.environ["MGB_ROOT"], "tests"))
from learnmultigrid_b200 import formats as F, partition as PT, problems as P
from oracle import kernels as K
from oracle.vcycle import OracleMultigrid

dist.init_process_group("gloo")
rank, W = dist.get_rank(), dist.get_world_size()
N, L, ND, NU = 16, 4, 2, 1                      # levels 0,1 partitioned; 2,3 replicated
A0 = P.structured_laplacian_2d(N, P.variable_coefficient)
Qs = [F.canonical_csr(q) for q in P.structured_hierarchy_2d(N, L, transfer="linear")]
As = [F.canonical_csr(A0)]
for q in Qs:
    As.append(F.canonical_csr(sp.csr_matrix(q.T @ sp.csc_matrix(As[-1]) @ q)))
QTs = [F.transpose_csr(q) for q in Qs]
cols = [F.greedy_colors(a)[0] for a in As[:-1]] + [None]
offs = [PT.block_offsets(a.shape[0], W) for a in As]

def plan_of(l, r):
    ext = PT.level_external_columns(As[l], QTs[l], Qs[l - 1] if l else None, offs[l], offs[l + 1],
                                    offs[l - 1] if l else None, r)
    return PT.RankPlan(offs[l], r, ext, cols[l])

plans = [plan_of(l, rank) for l in range(ND)]
nbrs = [{p: plan_of(l, p) for p in range(W) if p != rank} for l in range(ND)]

def local(M, rows_g, colmap):
    B = M[rows_g]
    return F.raw_csr(B.indptr, colmap[B.indices].astype(np.int32), B.data, (len(rows_g), int(colmap.max()) + 1))

def colmap_of(l):
    if l < ND:
        g = plans[l].gather_indices()
        m = -np.ones(As[l].shape[0], dtype=np.int64); m[g] = np.arange(len(g)); return m, len(g)
    return np.arange(As[l].shape[0], dtype=np.int64), As[l].shape[0]      # replicated: natural ordering

lev = []
for l in range(ND):
    p = plans[l]
    g = p.gather_indices()
    cm, nv = colmap_of(l)
    cmn, nvn = colmap_of(l + 1)
    own_next = (plans[l + 1].gather_indices()[:plans[l + 1].n_own] if l + 1 < ND
                else np.arange(offs[l + 1][rank], offs[l + 1][rank + 1]))
    Al = local(As[l], g[:p.n_own], cm); Al = F.raw_csr(Al.indptr, Al.indices, Al.data, (p.n_own, nv))
    Ql = local(Qs[l], g[:p.n_own], cmn); Ql = F.raw_csr(Ql.indptr, Ql.indices, Ql.data, (p.n_own, nvn))
    QTl = local(QTs[l], own_next, cm); QTl = F.raw_csr(QTl.indptr, QTl.indices, QTl.data, (len(own_next), nv))
    lev.append({"A": Al, "Q": Ql, "QT": QTl, "p": p, "nv": nv, "own_next": own_next})

def exchange(l, v, color=None):
    p = lev[l]["p"]; reqs, recv = [], {}
    for q in sorted(set(p.neighbours) | {r for r, pl in nbrs[l].items() if rank in pl.seg}):
        pl = nbrs[l][q]
        idx, ptr = p.send_indices(pl)
        a, b = (0, len(idx)) if color is None else (ptr[color], ptr[color + 1])
        out = torch.from_numpy(np.ascontiguousarray(v[idx[a:b]]))
        if q in p.seg:
            s0, s1 = p.seg[q] if color is None else p.seg_color[(q, color)]
        else:
            s0 = s1 = 0
        recv[q] = (s0, torch.empty(s1 - s0, dtype=torch.float64))
        if out.numel(): reqs.append(dist.isend(out, q))
        if s1 > s0: reqs.append(dist.irecv(recv[q][1], q))
    for r_ in reqs: r_.wait()
    for q, (s0, t) in recv.items():
        v[p.n_own + s0: p.n_own + s0 + t.numel()] = t.numpy()

oracle = OracleMultigrid(A0, np.zeros((As[0].shape[0], 1)), Qs, smoother="mcgs", colors=cols, hoist_setup=True)
oracle.build_hierarchy(L)

def smooth(l, x, b):
    d = lev[l]; p = d["p"]
    Asq = sp.vstack([d["A"], sp.csr_matrix((p.n_halo, d["nv"]))]).tocsr()
    Asq = F.raw_csr(Asq.indptr, Asq.indices, Asq.data, Asq.shape)
    bb = np.concatenate([b, np.zeros(p.n_halo)])
    for s in range(NU):
        for c in range(p.ncolors):
            rows = np.arange(p.color_ptr[c], p.color_ptr[c + 1], dtype=np.int32)
            K.gauss_seidel_multicolor(Asq, x, bb, [rows])
            exchange(l, x, c)

def vcycle(l, x, b):
    """x: local vector (own | halo) with a current halo; b: own rows"""
    d = lev[l]; p = d["p"]
    smooth(l, x, b)
    r = np.zeros(d["nv"]); r[:p.n_own] = K.residual(d["A"], x, b)
    exchange(l, r)
    rc = K.spmv(d["QT"], r)
    if l + 1 < ND:
        pc = lev[l + 1]["p"]
        e = np.zeros(lev[l + 1]["nv"])
        vcycle(l + 1, e, rc)
    else:
        sizes = [int(offs[l + 1][q + 1] - offs[l + 1][q]) for q in range(W)]
        mine = torch.zeros(max(sizes), dtype=torch.float64); mine[:len(rc)] = torch.from_numpy(np.ascontiguousarray(rc))
        parts = [torch.empty(max(sizes), dtype=torch.float64) for q in range(W)]      # gloo wants equal sizes
        dist.all_gather(parts, mine)
        bc = np.concatenate([t.numpy()[:k] for t, k in zip(parts, sizes)]).reshape(-1, 1)
        e = oracle.v_cycle(oracle._A_levels[l + 1], np.zeros_like(bc), bc, NU, L - (l + 1), level=l + 1).ravel()
    x[:p.n_own] = K.prolong_correct(d["Q"], e, x[:p.n_own])
    exchange(l, x)
    smooth(l, x, b)

rng = np.random.default_rng(1)
n0 = As[0].shape[0]
xg, bg = rng.standard_normal(n0), rng.standard_normal(n0)
g0 = plans[0].gather_indices()
x = xg[g0].copy(); b = bg[g0[:plans[0].n_own]].copy()
want = xg.reshape(-1, 1).copy()
ok = True
for cyc in range(2):
    vcycle(0, x, b)
    want = oracle.v_cycle(oracle.matrix, want, bg.reshape(-1, 1), NU, L)
    ok = ok and np.array_equal(x[:plans[0].n_own], want.ravel()[g0[:plans[0].n_own]])
t = torch.tensor([1 if ok else 0]); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0: print("VCYCLE_PARTITION_OK" if int(t.item()) == 1 else "VCYCLE_PARTITION_MISMATCH")
dist.destroy_process_group()
'''


def test_two_process_gloo_partitioned_vcycle(tmp_path):
    """the whole partitioned V-cycle of learnmultigrid_b200/distributed.py (halo per colour, residual halo for the
    restriction, all-gather hand-off to the replicated levels, prolongation, halo refresh) as two gloo processes with
    the oracle's kernels: bit-identical to the single-process oracle cycle"""
    script = tmp_path / "vworker.py"
    script.write_text(VCYCLE_WORKER)
    env = dict(os.environ, MGB_ROOT=ROOT, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", free_port(), str(script)],
                         env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert "VCYCLE_PARTITION_OK" in out.stdout, out.stdout[-4000:]


@pytest.mark.parametrize("seed,size", [(0, 2), (1, 3), (2, 5), (3, 8)])
def test_partitioned_sweeps_on_random_unstructured_graphs(seed, size):
    """random symmetric sparsity (no grid structure, long-range couplings: every rank can neighbour every other),
    random world sizes: plans are consistent and the emulated per-colour exchange reproduces the global sweep"""
    from oracle import kernels as K
    rng = np.random.default_rng(seed)
    n = 97 + 13 * seed
    S = sp.random(n, n, density=0.04, random_state=seed, format="csr")
    S = S + S.T + sp.diags(np.full(n, 8.0))
    A = F.canonical_csr(sp.csr_matrix(S))
    colors, nc = F.greedy_colors(A)
    offs, plans = plans_for(A, size, colors)
    x, b = rng.standard_normal(n), rng.standard_normal(n)
    perm, cptr = F.color_permutation(colors)
    want = x.copy()
    K.gauss_seidel_multicolor(A, want, b, [perm[cptr[c]:cptr[c + 1]] for c in range(nc)], iterations=1)
    loc = []
    for p in plans:
        g = p.gather_indices()
        assert len(set(g)) == len(g)
        slot = -np.ones(n, dtype=np.int64)
        slot[g] = np.arange(len(g))
        Ab = A[g[:p.n_own]]
        assert slot[Ab.indices].min() >= 0                      # every referenced column is owned or in the halo
        Asq = sp.vstack([F.raw_csr(Ab.indptr, slot[Ab.indices], Ab.data, (p.n_own, len(g))),
                         sp.csr_matrix((p.n_halo, len(g)))]).tocsr()
        loc.append({"A": F.raw_csr(Asq.indptr, Asq.indices, Asq.data, Asq.shape), "x": x[g].copy(),
                    "b": np.concatenate([b[g[:p.n_own]], np.zeros(p.n_halo)]), "g": g})
    for c in range(nc):
        for p, d in zip(plans, loc):
            rows = np.arange(p.color_ptr[c], p.color_ptr[c + 1], dtype=np.int32)
            if len(rows):
                K.gauss_seidel_multicolor(d["A"], d["x"], d["b"], [rows])
        for r, dr in zip(plans, loc):
            for s, ds in zip(plans, loc):
                if s.rank in r.seg:
                    idx, ptr = s.send_indices(r)
                    a = r.seg[s.rank][0]
                    dr["x"][r.n_own + a + ptr[c]: r.n_own + a + ptr[c + 1]] = ds["x"][idx[ptr[c]:ptr[c + 1]]]
    got = np.empty(n)
    for p, d in zip(plans, loc):
        got[d["g"][:p.n_own]] = d["x"][:p.n_own]
    assert np.array_equal(got, want)
