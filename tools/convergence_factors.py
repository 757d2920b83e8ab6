#!/usr/bin/env python
"""Convergence factors of the V-cycle with the engine's smoothers next to the reference's smoother (BASELINE.json
north star: "multicolour Gauss-Seidel is validated against a SciPy reference using the same colour ordering, and its
convergence factor is reported next to the reference's lexicographic GS").  Runs the CPU oracle (oracle/vcycle.py:
the reference's V-cycle statement by statement; `gs` = PyAMG's index-order sweep = what the reference runs, `mcgs` =
the multicolour sweep the GPU kernels are bit-identical to, `jacobi` = damped Jacobi), so it needs no GPU.

factor = geometric mean of the residual ratios of the last iterations before 1e-11 (first-iteration quirk skipped)."""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from learnmultigrid_b200 import formats as F, problems as P      # noqa: E402
from oracle.vcycle import OracleMultigrid                         # noqa: E402


def colours_for(A, Qs):
    cols, Al = [], sp.csr_matrix(A)
    for Q in Qs:
        cols.append(F.greedy_colors(F.canonical_csr(Al))[0])
        Al = sp.csr_matrix(Q.T @ Al @ Q)
    return cols + [None]


def factor(A, rhs, Qs, smoother, nu, colors=None, omega=2.0 / 3.0):
    o = OracleMultigrid(A, rhs, Qs, smoother=smoother, omega=omega, colors=colors, hoist_setup=True)
    o.solve(levels=len(Qs) + 1, smooth_steps=nu, error=1e-11, max_iterations=60)
    h = o.track_res.ravel()[1:]                 # entry 0 is the sqrt(n) quirk (Multigrid.py:64-66)
    r = h[1:] / h[:-1]
    tail = r[-min(4, len(r)):]
    return float(np.exp(np.mean(np.log(tail)))), len(o.track_res)


def main():
    rows = []
    for name, N, L, transfer, coef in (("2D 257^2 Laplacian, linear transfers, 5 levels", 256, 5, "linear", None),
                                       ("2D 257^2 Laplacian, quasi-L2 transfers, 5 levels", 256, 5, "quasi", None),
                                       ("2D 257^2 variable coefficient, linear transfers, 5 levels", 256, 5, "linear",
                                        P.variable_coefficient),
                                       ("2D 513^2 Laplacian, linear transfers, 6 levels", 512, 6, "linear", None)):
        A = P.structured_laplacian_2d(N, coef)
        rhs = P.structured_rhs_2d(N)
        Qs = P.structured_hierarchy_2d(N, L, transfer=transfer)
        cols = colours_for(A, Qs)
        struct = P.structured_colors_2d(N, L) if transfer == "linear" else None
        for nu in (1, 3):
            line = [name, "V(%d,%d)" % (nu, nu)]
            line.append("%.3f (%d its)" % factor(A, rhs, Qs, "gs", nu))
            line.append("%.3f (%d its)" % factor(A, rhs, Qs, "mcgs", nu, cols))
            line.append("%.3f (%d its)" % factor(A, rhs, Qs, "mcgs", nu, struct) if struct else "-")
            line.append("%.3f (%d its)" % factor(A, rhs, Qs, "jacobi", nu))
            rows.append(line)
            print(" | ".join(line), flush=True)
    pb = P.irregular_p1_2d(128, seed=42)
    print("irregular mesh rows need the NN-built transfers (GPU builder): see tests/test_gpu_neural2d.py")
    return rows


if __name__ == "__main__":
    print("configuration | cycle | lexicographic GS (reference) | multicolour GS, greedy colours | multicolour GS, "
          "structured 2/3 colours | damped Jacobi 2/3")
    main()
