"""Host logic of the multi-GPU path: row-block partition and halo plans (NumPy), including a world_size-2 gloo run
in which two processes build their plans independently, exchange halos with torch.distributed and reproduce the
single-process oracle sweep."""
import os
import subprocess
import sys

import numpy as np
import scipy.sparse as sp

from learnmultigrid_b200 import formats as F
from learnmultigrid_b200 import partition as PT
from helpers import poisson2d

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
                          "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)],
                         env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=240)
    assert "PARTITION_OK" in out.stdout, out.stdout[-3000:]
