#!/usr/bin/env python
"""BASELINE.json configs[4] (C5): hierarchy-setup sweep (device Galerkin Q^T A Q by two-pass SpGEMM, transposes,
colouring, SELL build, coarsest factorisation) for 1M-64M DOF structured meshes, and an MG-preconditioned CG solve
to ||r||_2 <= 1e-10 on the symmetrically eliminated operator.  Prints one JSON line per size.

SpGEMM bytes per SURVEY.md 8d: read A, Q, Q^T once, write AQ and A_c once:
    S(a,n) + 2 S(q,.) + 2 S(nnz(AQ), n) + S(a_c, n_c),   S(nnz, n) = 12 nnz + 4 (n+1).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def S(nnz, n):
    return 12 * nnz + 4 * (n + 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1024,2048,4096,8192")
    ap.add_argument("--levels", type=int, default=6)
    ap.add_argument("--transfer", default="linear", choices=["linear", "quasi"])
    ap.add_argument("--coefficient", default="variable", choices=["constant", "variable"])
    ap.add_argument("--tol", type=float, default=1e-10)
    a = ap.parse_args()
    import torch
    from learnmultigrid_b200 import problems as P, setup_device as SD, _lib
    from learnmultigrid_b200.solvers.CG import CG
    from learnmultigrid_b200.solvers.Multigrid import SemiGeometricMG
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    for N in [int(v) for v in a.sizes.split(",")]:
        coef = P.variable_coefficient if a.coefficient == "variable" else None
        t0 = time.perf_counter()
        A = P.symmetric_dirichlet(P.structured_laplacian_2d(N, coef), P.boundary_nodes_2d(N))
        rhs = P.structured_rhs_2d(N)
        Qs = P.structured_hierarchy_2d(N, a.levels, transfer=a.transfer)
        t_gen = time.perf_counter() - t0
        n0 = A.shape[0]
        # ---- Galerkin products alone (the reference's per-cycle `csr_matrix(i.T @ A @ i)`, Multigrid.py:97-98)
        Sx = SD.DeviceSetup(torch, dev)
        Ad = Sx.upload(A)
        galerkin = []
        for l, Q in enumerate(Qs):
            Qd = Sx.upload(Q)
            QTd = Sx.transpose(Qd)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            Ac = Sx.galerkin(Ad, Qd, QTd)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            # nnz(AQ) is not kept by galerkin(); bound it from the product of the patterns' row lengths
            AT = Sx.transpose(Ad)
            T = Sx.spgemm(AT, Qd)
            nnz_aq = T.nnz
            del AT, T
            byts = S(Ad.nnz, Ad.shape[0]) + 2 * S(Qd.nnz, Qd.shape[0]) + 2 * S(nnz_aq, Ad.shape[0]) + S(Ac.nnz, Ac.shape[0])
            galerkin.append({"level": l, "rows": Ad.shape[0], "nnz_A": Ad.nnz, "nnz_Q": Qd.nnz, "nnz_AQ": nnz_aq,
                             "nnz_Ac": Ac.nnz, "ms": dt * 1e3, "GBps": byts / dt / 1e9})
            Ad = Ac
            del Qd, QTd
        del Ad, Sx
        torch.cuda.empty_cache()
        # ---- full setup through the API, then PCG
        mg = SemiGeometricMG(A, rhs, Qs)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pre = mg.as_preconditioner(levels=a.levels, smoother="GaussSeidel", smooth_steps=1)
        torch.cuda.synchronize()
        t_setup = time.perf_counter() - t0
        rhs_pinned = torch.from_numpy(np.ascontiguousarray(rhs.reshape(-1))).pin_memory()
        cg = CG(A, rhs_pinned)
        cg.solve(max_iterations=3, error=0.0, preconditioner=pre)          # warm-up: graph capture, kernel loading
        cg = CG(A, rhs_pinned)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cg.solve(max_iterations=200, error=a.tol, preconditioner=pre)
        torch.cuda.synchronize()
        t_solve = time.perf_counter() - t0
        tm = cg.last_timing or {}
        line = {"config": "C5 setup sweep + MG-preconditioned CG", "grid": "%dx%d" % (N + 1, N + 1), "dof": n0,
                "levels": a.levels, "transfer": a.transfer, "coefficient": a.coefficient, "generate_s": round(t_gen, 2),
                "galerkin": galerkin, "galerkin_total_ms": sum(g["ms"] for g in galerkin),
                "setup_total_s": round(t_setup, 3), "pcg_iterations": cg.get_iterations(), "pcg_tol": a.tol,
                "pcg_final_residual": float(cg.track_res[-1, 0]), "pcg_solve_ms": t_solve * 1e3,
                "pcg_split_ms": {"rhs_host_to_device": tm.get("transfer_in_s", 0) * 1e3,
                                 "iterations": tm.get("iterations_s", 0) * 1e3,
                                 "per_iteration": tm.get("iterations_s", 0) * 1e3 / max(cg.get_iterations(), 1),
                                 "solution_device_to_host": tm.get("transfer_out_s", 0) * 1e3},
                "pcg_dof_per_s": n0 / t_solve}
        print(json.dumps(line), flush=True)
        del mg, cg, pre
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
