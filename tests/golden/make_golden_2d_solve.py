"""Generate tests/golden/solve_2d.npz: the REFERENCE's own two-level 2D solves of SURVEY.md section 8c
(test/thesis_structured_2d.py:380-414, 457-458), run through oracle/refshim.py:

    mesh = Mesh2D(N*N); A, M, rhs from the reference's 2D assembly with f = -1; Dirichlet rows replaced by identity
    rows and rhs zeroed there; SemiGeometricMG(A, rhs, Q).solve(levels=2, smoother="GaussSeidel", smooth_steps=3,
    error=1e-09, max_iterations=20)

Only Q is restated (the reference reads it from MATLAB files that are not shipped, thesis_compare_2D.py:386):
"quasi" = rownormalise(M_ref P), "linear" = P, with P the linear interpolation from the nested mesh Mesh2D((N/2)^2)
(learnmultigrid_b200.problems.linear_P_2d).  Stored per case: A, Q (COO), rhs, track_res, solution, iterations.
Run in the authoring container:  python tests/golden/make_golden_2d_solve.py
"""
import os
import sys
import warnings

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")

from oracle import refshim  # noqa: E402

refshim.install()
with refshim.quiet():
    from learn_multigrid.mesh.Mesh2D import Mesh2D
    from learn_multigrid.assembly.MassMatrix import MassMatrix
    from learn_multigrid.assembly.StiffnessMatrix import StiffnessMatrix
    from learn_multigrid.assembly.LoadVector import LoadVector
    from learn_multigrid.assembly.LoadFunction import LoadFunction
    from learn_multigrid.assembly.Quadrature import Quadrature2D
    from learn_multigrid.assembly.ShapeFunction import FunctionTriangle, GradientTriangle
    from learn_multigrid.solvers.Multigrid import SemiGeometricMG


def coo(prefix, M):
    c = sp.coo_matrix(sp.csr_matrix(M))
    return {prefix + "_row": c.row.astype(np.int32), prefix + "_col": c.col.astype(np.int32),
            prefix + "_data": c.data.astype(np.float64), prefix + "_shape": np.array(c.shape)}


def main():
    from learnmultigrid_b200.problems import linear_P_2d
    out = {}
    for N, kinds in ((16, ("quasi", "linear")), (32, ("quasi",))):
        with refshim.quiet():
            mesh = Mesh2D(N * N)
            q = Quadrature2D(3)
            A = StiffnessMatrix(mesh).compute_stiffness_2d(GradientTriangle(1), q)
            M = MassMatrix(mesh).compute_mass_2d(FunctionTriangle(1), q)
            rhs = LoadVector(mesh).compute_rhs_2d(LoadFunction(lambda x: -1), FunctionTriangle(1), q)
            p = mesh.p
            border = np.logical_or(np.logical_or(p[:, 0] == 0, p[:, 0] == 1),
                                   np.logical_or(p[:, 1] == 0, p[:, 1] == 1))
            nodes = np.where(border)[0]
            eye = np.eye(len(p))
            A[nodes, :] = eye[nodes, :]                  # thesis_structured_2d.py:407-414
            rhs[nodes] = 0
        P = sp.csr_matrix(linear_P_2d(N))
        for kind in kinds:
            if kind == "quasi":
                B = sp.csr_matrix(sp.csr_matrix(M) @ P)
                Q = sp.csr_matrix(sp.diags(1.0 / np.asarray(B.sum(axis=1)).ravel()) @ B)
            else:
                Q = P.copy()
            Q.sort_indices()
            with refshim.quiet():
                mg = SemiGeometricMG(A.copy(), rhs.copy(), Q)
                mg.solve(levels=2, smoother="GaussSeidel", smooth_steps=3, error=1e-09, max_iterations=20)
            name = "N%d_%s" % (N, kind)
            out.update(coo(name + "_A", A))
            out.update(coo(name + "_Q", Q))
            out[name + "_rhs"] = np.asarray(rhs, dtype=np.float64).reshape(-1, 1)
            out[name + "_track"] = np.asarray(mg.track_res, dtype=np.float64)
            out[name + "_x"] = np.asarray(mg.get_solution(), dtype=np.float64)
            out[name + "_its"] = np.array(mg.get_iterations())
            x = out[name + "_x"].ravel()
            print(name, int(out[name + "_its"]), " ".join("%.10e" % v for v in out[name + "_track"].ravel()))
            print("   ||u|| = %.12e   u(centre) = %.12e   rhs interior = %.6e"
                  % (np.linalg.norm(x), x[(N // 2) * (N + 1) + N // 2], rhs[(N + 1) + 1, 0]))
    np.savez_compressed(os.path.join(HERE, "solve_2d.npz"), **out)
    print(os.path.getsize(os.path.join(HERE, "solve_2d.npz")))


if __name__ == "__main__":
    main()
