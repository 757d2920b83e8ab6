#!/usr/bin/env python
"""Where the time of one Galerkin product A_c = Q^T A Q goes on the device (setup_device.DeviceSetup.galerkin): the
sub-steps of the level-0 product of the N x N benchmark grid (operator and transfer operator generated in HBM), each
bracketed by a device synchronisation, after one untimed warm-up product.  Prints one JSON line.

    python tools/time_galerkin.py --n 8192
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def csr_bytes(nnz, n):
    return 12 * nnz + 4 * (n + 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--coefficient", default="constant", choices=["constant", "variable"])
    a = ap.parse_args()
    import torch
    from learnmultigrid_b200 import problems_device as PD, setup_device as SD
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    S = SD.DeviceSetup(torch, dev)
    A = PD.structured_laplacian_2d(a.n, None if a.coefficient == "constant" else PD.variable_coefficient)
    Q = PD.structured_hierarchy_2d(a.n, 2)[0]
    QT = S.transpose(Q)
    S.galerkin(A, Q, QT)                      # warm-up: allocator, module loading
    torch.cuda.synchronize()

    steps = {}

    class Timed:
        def __init__(self, name):
            self.name = name

        def __enter__(self):
            torch.cuda.synchronize()
            self.t = time.perf_counter()

        def __exit__(self, *exc):
            torch.cuda.synchronize()
            steps[self.name] = steps.get(self.name, 0.0) + (time.perf_counter() - self.t) * 1e3

    # the same sequence as DeviceSetup.galerkin / spgemm, step by step
    lib = S.lib
    t = torch

    def spgemm(tag, X, Y):
        n = X.shape[0]
        counts = S.empty(n, t.int32)
        avg = Y.nnz / max(Y.shape[0], 1)
        import numpy as np
        est = 2.0 * (X.nnz / max(n, 1)) * avg            # as DeviceSetup.spgemm sizes its first symbolic table
        log_t = min(7, max(5, int(np.ceil(np.log2(max(2.0 * est, 2.0))))))
        with Timed(tag + " symbolic (2^%d slots)" % log_t):
            g = S._group_for(avg, log_t, False)
            S._flag.zero_()
            lib.mg_spgemm_symbolic(n, X.indptr.data_ptr(), X.indices.data_ptr(), Y.indptr.data_ptr(),
                                   Y.indices.data_ptr(), g, log_t, counts.data_ptr(), S._flag.data_ptr(), S.st())
            assert int(S._flag.item()) == 0, "table overflow: DeviceSetup.spgemm would retry with a larger one"
        with Timed(tag + " row maximum + scan"):
            max_row = int(counts.max().item())
            cptr, total = S.scan(counts, n)
        cidx = S.empty(total, t.int32)
        cval = S.empty(total, t.float64)
        log_n = max(4, int(np.ceil(np.log2(max(2 * max_row, 2)))))
        nzc = S.empty(n, t.int32)
        with Timed(tag + " numeric (2^%d slots, group %d)" % (log_n, S._group_for(avg, log_n, True))):
            S._flag.zero_()
            lib.mg_spgemm_numeric(n, *X.ptrs(), *Y.ptrs(), S._group_for(avg, log_n, True), log_n, cptr.data_ptr(),
                                  cidx.data_ptr(), cval.data_ptr(), nzc.data_ptr(), S._flag.data_ptr(), S.st())
            assert int(S._flag.item()) == 0
        with Timed(tag + " non-zero scan"):
            optr, ototal = S.scan(nzc, n)
        if ototal == total:
            return SD.DevCSR((n, Y.shape[1]), cptr, cidx, cval)
        with Timed(tag + " pruning of exact zeros"):
            oidx = S.empty(ototal, t.int32)
            oval = S.empty(ototal, t.float64)
            lib.mg_csr_compact_nonzeros(n, cptr.data_ptr(), cidx.data_ptr(), cval.data_ptr(), optr.data_ptr(),
                                        oidx.data_ptr(), oval.data_ptr(), S.st())
        return SD.DevCSR((n, Y.shape[1]), optr, oidx, oval)

    t0 = time.perf_counter()
    with Timed("transpose A"):
        AT = S.transpose(A)
    T = spgemm("A^T Q:", AT, Q)
    del AT
    C = spgemm("Q^T (A^T Q):", QT, T)
    with Timed("transpose C"):
        Ac = S.transpose(C)
    torch.cuda.synchronize()
    total_ms = (time.perf_counter() - t0) * 1e3
    nbytes = (csr_bytes(A.nnz, A.shape[0]) + csr_bytes(Q.nnz, Q.shape[0]) + csr_bytes(QT.nnz, QT.shape[0])
              + 2 * csr_bytes(T.nnz, A.shape[0]) + csr_bytes(Ac.nnz, Ac.shape[0]))
    print(json.dumps({"what": "level-0 Galerkin product, sub-steps (ms, device synchronised around each)",
                      "grid": "%dx%d" % (a.n + 1, a.n + 1), "rows": A.shape[0], "nnz_A": A.nnz, "nnz_Q": Q.nnz,
                      "nnz_AQ": T.nnz, "coarse_rows": Ac.shape[0], "nnz_Ac": Ac.nnz,
                      "steps_ms": {k: round(v, 3) for k, v in steps.items()}, "total_ms": round(total_ms, 3),
                      "algorithmic_bytes": nbytes, "gb_per_s": round(nbytes / total_ms / 1e6, 1)}))


if __name__ == "__main__":
    main()
