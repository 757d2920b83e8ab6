#!/usr/bin/env python
"""Multi-process check of the partitioned V-cycle: launch with torchrun, one rank per GPU.  Every rank also builds
the single-GPU hierarchy and compares iterates bit for bit.  Prints DIST_CHECK_OK from rank 0 on success."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--levels", type=int, default=4)
    ap.add_argument("--cycles", type=int, default=4)
    ap.add_argument("--smoother", default="mcgs")
    ap.add_argument("--n-dist", type=int, default=2)
    ap.add_argument("--latency", action="store_true", help="time back-to-back exchange sites instead")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if a.latency:
        return latency(torch, dist)
    from learnmultigrid_b200 import problems as P
    from learnmultigrid_b200.distributed import DistributedHierarchy, TorchFabric
    from learnmultigrid_b200.engine import DeviceHierarchy
    A = P.structured_laplacian_2d(a.size)
    Qs = P.structured_hierarchy_2d(a.size, a.levels, transfer="linear")
    rng = np.random.default_rng(11)
    n = A.shape[0]
    b, x0 = rng.standard_normal(n), rng.standard_normal(n)
    h1 = DeviceHierarchy(A, Qs, smoother=a.smoother)
    fab = TorchFabric()
    hd = DistributedHierarchy(A, Qs, fab, smoother=a.smoother, colors=h1.colors, n_dist=a.n_dist, timeout_s=30.0)
    ok = True
    for h in (h1, hd):
        h.set_rhs(b)
        h.set_x(x0)
    p1 = h1.make_params(nu_pre=1, nu_post=1, omega=2.0 / 3.0)
    pd = hd.make_params(nu_pre=1, nu_post=1, omega=2.0 / 3.0)
    for it in range(a.cycles):
        n1 = h1.residual_norm()
        h1.vcycle(p1)
        hd.vcycle(pd, with_norm=True)
        nd = hd.last_norm()
        x1, xd = h1.get_x(), hd.get_x()
        same = np.array_equal(x1, xd)
        close = abs(n1 - nd) <= 1e-12 * n1
        if fab.rank == 0:
            print("cycle %d: norm single %.15e partitioned %.15e, iterate bit-identical: %s" % (it, n1, nd, same))
        ok = ok and same and close
    hd.check()
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if fab.rank == 0:
        print("DIST_CHECK_OK" if int(t.item()) == 1 else "DIST_CHECK_MISMATCH")
    hd.close()
    dist.destroy_process_group()


def latency(torch, dist):
    """microseconds per exchange site: `sites` ring exchanges of `count` doubles per program, graph-replayed"""
    import ctypes
    from learnmultigrid_b200 import _lib
    from learnmultigrid_b200.distributed import PeerComm, TorchFabric
    lib = _lib.load()
    fab = TorchFabric()
    r, W = fab.rank, fab.world
    comm = PeerComm(fab, torch, region_bytes=8 << 20, max_sites=256, timeout_s=20.0)
    dev = torch.device("cuda", torch.cuda.current_device())
    for count in (0, 1, 1024, 8192, 65536):
        src = torch.zeros(max(count, 1), dtype=torch.float64, device=dev)
        dst = torch.zeros(2 * max(count, 1), dtype=torch.float64, device=dev)
        x = _lib.mg_xfer()
        peers = sorted({(r - 1) % W, (r + 1) % W})
        x.npeers = len(peers)
        for k, q in enumerate(peers):
            x.peer[k] = q
            x.send_cnt[k] = count
            x.recv_off[k] = k * count
            x.recv_cnt[k] = count
        sites = 50
        cap = torch.cuda.Stream()
        cap.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cap):
            h = cap.cuda_stream
            _lib.check(lib.mg_graph_begin(h))
            _lib.check(lib.mg_comm_begin(ctypes.byref(comm.struct)))
            for _ in range(sites):
                _lib.check(lib.mg_comm_exchange(ctypes.byref(comm.struct), ctypes.byref(x), src.data_ptr(), dst.data_ptr(), h))
            _lib.check(lib.mg_comm_end(ctypes.byref(comm.struct), h))
            g = ctypes.c_void_p()
            _lib.check(lib.mg_graph_end(h, ctypes.byref(g)))
        torch.cuda.current_stream().wait_stream(cap)
        st = _lib.stream_handle(torch)
        for _ in range(5):
            _lib.check(lib.mg_graph_launch(g, st))
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 40
        e0.record()
        for _ in range(reps):
            _lib.check(lib.mg_graph_launch(g, st))
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (reps * sites)
        t = torch.tensor([us], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if r == 0:
            print("exchange site, %d peers, %6d doubles per message: %.2f us per site" % (len(peers), count, float(t.item())))
        comm.check()
        lib.mg_graph_destroy(g)
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
