#!/usr/bin/env python
"""Weak-scaling line of the partitioned V-cycle: every GPU holds one N x N square of cells, the squares are stacked in
y (domain [0,1] x [0,W]), so the per-GPU work is fixed as the GPU count W grows and the global problem
((N+1) x (W N + 1) nodes, 537 M DOF at N = 8192, W = 8) is larger than the replicated setup of bench.py could hold.

    torchrun --nnodes=1 --nproc-per-node W --master-addr 127.0.0.1 tools/bench_weak.py --cells 8192 --steps 20

No rank ever forms a global operator: each generates its own rows (problems.structured_laplacian_2d(rows=, Ny=),
linear_P_2d(rows=, Nyf=)) and the hierarchy is built by distributed_strip.StripHierarchy.  Prints one JSON line in the
layout of bench.py (`"scaling": "weak"`).  A step = one V(nu,nu) cycle + the residual norm of its result (one graph, as
in bench.py), CUDA events, max over ranks.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=8192, dest="n", help="cells per side of one GPU's square (not --n: torchrun's parser would claim the abbreviation)")
    ap.add_argument("--levels", type=int, default=6)
    ap.add_argument("--nu", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--coefficient", default="constant", choices=["constant", "variable"])
    ap.add_argument("--min-rows-per-rank", type=int, default=65536)
    ap.add_argument("--colors", default="greedy", choices=["greedy", "structured"],
                    help="greedy: first-fit, coloured block by block in rank order (serial across the ranks); "
                         "structured: (ix+iy) %% 2 / %% 3 evaluated on the own rows, no exchange")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from learnmultigrid_b200 import partition as PT, problems as P
    from learnmultigrid_b200.distributed import TorchFabric
    from learnmultigrid_b200.distributed_strip import StripHierarchy
    fab = TorchFabric()
    r, W = fab.rank, fab.world
    N, Ny, L = a.n, a.n * W, a.levels
    ns = [((N >> l) + 1) * ((Ny >> l) + 1) for l in range(L)]
    offs = [PT.block_offsets(n, W) for n in ns]
    n_dist = 0
    for l in range(L - 1):
        if ns[l] // W >= a.min_rows_per_rank:
            n_dist = l + 1
        else:
            break
    n_dist = max(n_dist, 1)
    coef = P.variable_coefficient if a.coefficient == "variable" else None
    t0 = time.perf_counter()
    A_blk = P.structured_laplacian_2d(N, coef, rows=(offs[0][r], offs[0][r + 1]), Ny=Ny)
    Q_blks = [P.linear_P_2d(N >> l, rows=(offs[l][r], offs[l][r + 1]), Nyf=Ny >> l) for l in range(L - 1)]
    rhs = P.structured_rhs_2d(N, rows=(offs[0][r], offs[0][r + 1]), Ny=Ny)
    colors = None
    if a.colors == "structured":
        colors = []
        for l in range(L - 1):
            Wl = (N >> l) + 1
            lo, hi = (offs[l][r], offs[l][r + 1]) if l < n_dist else (0, ns[l])
            iy, ix = np.divmod(np.arange(lo, hi, dtype=np.int64), Wl)
            colors.append(((ix + iy) % (2 if l == 0 else 3)).astype(np.int32))
        colors.append(None)
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    h = StripHierarchy(A_blk, Q_blks, offs, fab, n_dist, smoother="mcgs", colors=colors, timeout_s=60.0)
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t0
    del A_blk, Q_blks
    params = h.make_params(nu_pre=a.nu, nu_post=a.nu, omega=2.0 / 3.0)
    h.set_rhs(rhs)
    h.zero_x()
    for _ in range(max(a.warmup, 3)):
        h.vcycle(params, norm_after=True)
    torch.cuda.synchronize()
    h.zero_x()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        h.vcycle(params, norm_after=True)
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / a.steps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    res = h.residual_norm()
    h.check()
    cyc = h.cycle_bytes(a.nu, a.nu)
    if r == 0:
        print(json.dumps({
            "metric": "vcycle_fine_grid_dof_per_s", "value": ns[0] / (ms * 1e-3), "unit": "DOF/s", "n_gpus": W,
            "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "dtype": "f64", "data": "synthetic",
            "config": {"workload": "2D structured P1 %s, %d stacked %dx%d squares (%d DOF), %d-level V(%d,%d), multicolour "
                                   "Gauss-Seidel (%s colours), linear transfers; strip-local setup"
                                   % ("Laplacian" if coef is None else "variable-coefficient stiffness", W, N, N, ns[0],
                                      L, a.nu, a.nu, a.colors),
                       "levels_rows": ns, "partitioned_levels": n_dist, "generate_s": round(t_gen, 2),
                       "setup_s": round(t_setup, 2), "residual_after_timed_steps": res,
                       "dof_per_gpu": ns[0] // W},
            "roofline": {"bound": "hbm", "cycle": {"algorithmic_bytes": cyc["total"],
                                                   "achieved_gbs_per_gpu": cyc["total"] / W / (ms * 1e-3) / 1e9}},
            "gpu_launches": int(h.last_launches) * a.steps}))
    h.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
