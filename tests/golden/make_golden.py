"""Generate tests/golden/*.npz by running the REFERENCE's own code (through oracle/refshim.py) in the
authoring container.  Run:  python tests/golden/make_golden.py
The GPU box has no /root/reference; tests there only read the committed .npz files.

Every array named `ref_*` was produced by unmodified reference modules (learn_multigrid.* imported from
/root/reference) with the shims listed in oracle/refshim.py; nothing in them comes from the product.
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
    from learn_multigrid.mesh.Mesh1D import Mesh1D
    from learn_multigrid.mesh.Mesh2D import Mesh2D
    from learn_multigrid.assembly.StiffnessMatrix import StiffnessMatrix
    from learn_multigrid.assembly.MassMatrix import MassMatrix
    from learn_multigrid.assembly.LoadVector import LoadVector
    from learn_multigrid.assembly.LoadFunction import LoadFunction
    from learn_multigrid.assembly.Quadrature import Quadrature, Quadrature2D
    from learn_multigrid.assembly.ShapeFunction import Function, Gradient, FunctionTriangle, GradientTriangle
    from learn_multigrid.L2_projection.L2Projection import L2Projection
    from learn_multigrid.L2_projection.Intersection import Intersection
    from learn_multigrid.L2_projection.CouplingOperator import CouplingOperator
    from learn_multigrid.solvers.Multigrid import SemiGeometricMG, GeometricMG, NeuralMG_2D
    from learn_multigrid.solvers.Jacobi import Jacobi
    from learn_multigrid.solvers.GaussSeidel import GaussSeidel
    from learn_multigrid.solvers.CG import CG
    from learn_multigrid.solvers.Solver import DirectSolver
    from learn_multigrid.utilities.laplacian import laplacian_1d_fd_bc


def coo(prefix, M):
    M = sp.coo_matrix(M)
    return {prefix + "_row": M.row.astype(np.int32), prefix + "_col": M.col.astype(np.int32),
            prefix + "_data": M.data.astype(np.float64), prefix + "_shape": np.array(M.shape, dtype=np.int64)}


def ones(x):
    return np.ones(shape=np.shape(x))


def problem_1d(ne, regular=True, seed=None):
    if seed is not None:
        np.random.seed(seed)
    m = Mesh1D(regular, ne)
    m.construct()
    A = StiffnessMatrix(m).compute_stiffness_1d(Gradient(2), Quadrature(3))
    M = MassMatrix(m).compute_mass_1d(Function(2), Quadrature(3))
    rhs = LoadVector(m).compute_rhs_1d(ones)
    rhs[0] = 0
    rhs[-1] = 0
    # Dirichlet rows exactly as test/test_NN.py:176-181
    A[1, 0] = 0
    A[-2, -1] = 0
    A[0, :] = 0
    A[-1, :] = 0
    A[0, 0] = 1
    A[-1, -1] = 1
    return m, A, M, rhs


def main():
    out = {}
    with refshim.quiet():
        # ---------------- C1: 1D, ne=1024, coarse 512 (BASELINE.json configs[0]) -----------------------
        m, A, M, rhs = problem_1d(1024)
        mc = Mesh1D(True, 512)
        mc.construct()
        c1 = {}
        c1.update(coo("A", A))
        c1.update(coo("M", M))
        c1["rhs"] = rhs
        c1["x_fine"] = m.get_mesh()
        c1["x_coarse"] = mc.get_mesh()
        Qs = {}
        for typ in ("quasi", "pseudo", "L2"):
            Q, _ = L2Projection(typ, m, mc).compute_transfer_1d()
            Qs[typ] = Q
            if typ == "L2":
                c1["Q_L2_dense"] = Q          # M^-1 B is dense
            else:
                c1.update(coo("Q_" + typ, Q))
        cases = [("quasi", 3, 1, 1e-10, 40), ("quasi", 2, 3, 1e-10, 40), ("quasi", 3, 3, 1e-10, 40),
                 ("pseudo", 3, 3, 1e-10, 40), ("L2", 2, 3, 1e-10, 40), ("quasi", 5, 3, 1e-10, 40)]
        for typ, lev, steps, err, mi in cases:
            mg = SemiGeometricMG(A, rhs, Qs[typ])
            mg.solve(levels=lev, smoother="GaussSeidel", smooth_steps=steps, error=err, max_iterations=mi)
            key = "ref_sgmg_%s_L%d_s%d" % (typ, lev, steps)
            c1[key + "_track"] = mg.track_res
            c1[key + "_x"] = mg.get_solution()
            c1[key + "_its"] = np.array(mg.get_iterations())
        g = GeometricMG(A, rhs)
        g.solve(levels=3, smoother="GaussSeidel", smooth_steps=3, error=1e-10, max_iterations=40)
        c1["ref_gmg_L3_s3_track"] = g.track_res
        c1["ref_gmg_L3_s3_x"] = g.get_solution()
        # testMG.py-style FD matrix (utilities/laplacian.py:22-59), pseudo, levels=3, steps=1, 1e-11
        L, X, rhs_fd = laplacian_1d_fd_bc(m, ones)
        mg = SemiGeometricMG(L, rhs_fd, Qs["pseudo"])
        mg.solve(smoother="GaussSeidel", smooth_steps=1, levels=3, max_iterations=100, error=1e-11)
        c1.update(coo("A_fd", L))
        c1["rhs_fd"] = rhs_fd
        c1["ref_fd_pseudo_L3_s1_track"] = mg.track_res
        c1["ref_fd_pseudo_L3_s1_x"] = mg.get_solution()
        # a second .solve() on the same object keeps counting iterations and skips the sqrt(n) quirk
        mg2 = SemiGeometricMG(A, rhs, Qs["quasi"])
        mg2.solve(levels=2, smoother="GaussSeidel", smooth_steps=1, error=1e-10, max_iterations=2)
        mg2.solve(levels=2, smoother="GaussSeidel", smooth_steps=1, error=1e-10, max_iterations=3)
        c1["ref_twice_track"] = mg2.track_res
        c1["ref_twice_its"] = np.array(mg2.get_iterations())
        # initial guess is mutated in place by the pre-smoother (PyAMG ravel view)
        x0 = np.full((1025, 1), 0.25)
        mg3 = SemiGeometricMG(A, rhs, Qs["quasi"])
        mg3.solve(levels=2, smoother="GaussSeidel", smooth_steps=1, error=1e-10, max_iterations=2,
                  initial_guess=x0)
        c1["ref_guess_track"] = mg3.track_res
        c1["ref_guess_x0_after"] = x0
        c1["ref_guess_x"] = mg3.get_solution()
        np.savez_compressed(os.path.join(HERE, "c1_1d_1024.npz"), **c1)

        # ---------------- small 1D transfer operators (regular + seeded irregular) ---------------------
        t = {}
        for tag, regular, seed in (("reg", True, None), ("irr", False, 42)):
            m, A, M, rhs = problem_1d(16, regular, seed)
            if regular:
                mc = Mesh1D(True, 8)
                mc.construct()
            else:      # nested coarse mesh: every other node of the irregular fine mesh
                mc = Mesh1D(True, 8)
                mc.construct()
                mc.x = m.get_mesh()[::2].copy()
                mc.connection_matrix()
            inter = Intersection(m, mc)
            inter.find_intersections1d()
            ints, coords, union = inter.get_info()
            B = CouplingOperator(inter, m, mc).compute_b_1d(Quadrature(3), Function(2))
            t[tag + "_x_fine"] = m.get_mesh()
            t[tag + "_x_coarse"] = mc.get_mesh()
            t[tag + "_intersections"] = ints
            t[tag + "_int_coord"] = coords
            t[tag + "_B"] = B
            t[tag + "_M"] = M
            t[tag + "_A"] = A
            t[tag + "_rhs"] = rhs
            for typ in ("quasi", "pseudo", "L2"):
                Q, _ = L2Projection(typ, m, mc).compute_transfer_1d()
                t[tag + "_Q_" + typ] = Q
        # non-nested pair as in test/testMG.py:44-49 (Mesh1DRefinement 2*2^3=16 vs 3*2^2=12 elements)
        from learn_multigrid.mesh.Mesh1D import Mesh1DRefinement
        mf = Mesh1DRefinement(coarse_ne=2, n_ref=3)
        mf.construct()
        mcn = Mesh1DRefinement(coarse_ne=3, n_ref=2)
        mcn.construct()
        for typ in ("quasi", "pseudo", "L2"):
            Q, _ = L2Projection(typ, mf, mcn).compute_transfer_1d()
            t["nonnested_Q_" + typ] = Q
        t["nonnested_x_fine"] = mf.get_mesh()
        t["nonnested_x_coarse"] = mcn.get_mesh()
        np.savez_compressed(os.path.join(HERE, "transfer_1d_small.npz"), **t)

        # ---------------- 2D assembly on Mesh2D(16x16 squares) = 289 nodes -----------------------------
        d2 = {}
        for N in (4, 16):
            mesh = Mesh2D(N * N)
            q = Quadrature2D(3)
            A2 = StiffnessMatrix(mesh).compute_stiffness_2d(GradientTriangle(1), q)
            M2 = MassMatrix(mesh).compute_mass_2d(FunctionTriangle(1), q)
            rhs2 = LoadVector(mesh).compute_rhs_2d(LoadFunction(lambda x: -1), FunctionTriangle(1), q)
            d2["N%d_p" % N] = mesh.p
            d2["N%d_conn" % N] = mesh.conn
            d2.update(coo("N%d_A_raw" % N, A2))
            d2.update(coo("N%d_M" % N, M2))
            d2["N%d_rhs_raw" % N] = rhs2.copy()
            # Dirichlet rows as test/thesis_structured_2d.py:407-414
            p = mesh.p
            border = np.logical_or(np.logical_or(p[:, 0] == 0, p[:, 0] == 1),
                                   np.logical_or(p[:, 1] == 0, p[:, 1] == 1))
            nodes = np.where(border)[0]
            eye = np.eye(len(p))
            A2[nodes, :] = eye[nodes, :]
            rhs2[nodes] = 0
            d2.update(coo("N%d_A" % N, A2))
            d2["N%d_rhs" % N] = rhs2
        # rectangular ne (find_balanced_couple, Mesh2D.py:41-61): 12 squares -> 4 x 3
        mesh = Mesh2D(12)
        d2["rect12_p"] = mesh.p
        d2["rect12_conn"] = mesh.conn
        np.savez_compressed(os.path.join(HERE, "assembly_2d.npz"), **d2)

        # ---------------- stationary solvers / CG / direct on the 3x3 matrix (test_solver.py:30-48) ----
        s = {}
        A3 = np.array([[30.0, 1, 15], [28, 60, 3], [100, 19, 150]])
        b3 = np.array([[1.0], [2.0], [3.0]])
        j = Jacobi(A3, b3)
        j.solve(max_iterations=1000, error=1e-12)
        s["jacobi_track"] = j.track_res
        s["jacobi_x"] = j.get_solution()
        gs = GaussSeidel(A3, b3)
        gs.solve(max_iterations=1000, error=1e-12)
        s["gs_track"] = gs.track_res
        s["gs_x"] = gs.get_solution()
        d = DirectSolver(A3, b3)
        d.solve()
        s["direct_x"] = d.get_solution()
        s["direct_res"] = np.array(d.get_residual())
        # CG on the symmetric 1D problem (ne=32)
        m, A, M, rhs = problem_1d(32)
        cg = CG(A, rhs)
        cg.solve(max_iterations=200, error=1e-10)
        s["cg_A"] = A
        s["cg_rhs"] = rhs
        s["cg_track"] = cg.track_res
        s["cg_x"] = cg.get_solution()
        s["cg_its"] = np.array(cg.get_iterations())
        s["A3"] = A3
        s["b3"] = b3
        np.savez_compressed(os.path.join(HERE, "solvers_small.npz"), **s)

        # ---------------- NeuralMG_2D plumbing with a stub predictor ------------------------
        nn = {}
        mesh = Mesh2D(64)      # 8x8 squares, 81 nodes
        q = Quadrature2D(3)
        M2 = MassMatrix(mesh).compute_mass_2d(FunctionTriangle(1), q)
        A2 = StiffnessMatrix(mesh).compute_stiffness_2d(GradientTriangle(1), q)

        class Stub:
            def predict(self, X):
                # deterministic smooth function of the patch so that ordering errors are visible
                X = np.asarray(X, dtype=np.float64)
                w = np.linspace(0.5, 1.5, 31)[None, :]
                return (1.0 + np.tanh(X.sum(axis=1, keepdims=True))) * w
        rhs2 = np.zeros((81, 1))
        nmg = NeuralMG_2D(A2, rhs2, Stub(), M2, np.ones(43), np.zeros(43))
        C, F, Cn, Fn = nmg.coarsening(sp.lil_matrix(M2))
        nn["C"] = np.array(sorted(C))
        nn["C_order"] = np.array(list(C))
        patches, fill = nmg.extract_patches(C, sp.lil_matrix(M2))
        nn["patches"] = patches
        nn["fill"] = fill
        nmg.define_hierarchy(levels=3)
        for k, Qk in enumerate(nmg.l_hierarchy):
            nn["Q%d" % k] = np.asarray(Qk.todense() if sp.issparse(Qk) else Qk, dtype=np.float64)
        nn.update(coo("M", M2))
        np.savez_compressed(os.path.join(HERE, "neural_2d_stub.npz"), **nn)

    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
