"""Generate tests/golden/neural_2d_cases.npz: every intermediate of the REFERENCE's NeuralMG_2D hierarchy builder
(learn_multigrid/solvers/Multigrid.py:401-765), produced by the reference's own methods (oracle/refshim.py) with a
deterministic stub in place of the Keras model (the trained weights are not shipped).  Run in the authoring
container:  python tests/golden/make_golden_neural.py

For each case and level l the file holds:  M (COO of the level's mass matrix as the builder sees it, i.e. after
pre_process), C (coarse nodes in selection order), patches (n,43), fill (n,31), pred (n,31), B (dense), Q (dense),
and dn (the d_neighs table returned by fill_B, -1 padded) -- exactly the statement sequence of define_hierarchy
(:741-765), with the intermediates kept.
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
    from learn_multigrid.assembly.Quadrature import Quadrature2D
    from learn_multigrid.assembly.ShapeFunction import FunctionTriangle, GradientTriangle
    from learn_multigrid.solvers.Multigrid import NeuralMG_2D
from scipy.sparse import lil_matrix  # noqa: E402


class Stub:
    """deterministic smooth function of the (normalised) patch, so that ordering errors are visible"""

    def predict(self, X):
        X = np.asarray(X, dtype=np.float64)
        w = np.linspace(0.5, 1.5, 31)[None, :]
        return (1.0 + np.tanh(X.sum(axis=1, keepdims=True))) * w


def irregular_mesh(ne, seed):
    """the reference's Mesh2D.refine(regular=False) fails under NumPy >= 1.24 (ragged assignment, Mesh2D.py:149-156);
    the product's restated refine produces the refined mesh, the REFERENCE assembles on it"""
    for m in [m for m in sys.modules if m == "learnmultigrid_b200" or m.startswith("learnmultigrid_b200.")]:
        pass
    from learnmultigrid_b200.mesh.Mesh2D import Mesh2D as OurMesh
    np.random.seed(seed)
    m = OurMesh(ne)
    m.refine(regular=False)
    return Mesh2D(p=np.array(m.p), conn=np.array(m.conn))


def flipped_mesh(nsq, flips, seed, jitter=0.15):
    """nsq x nsq squares; the diagonal of the listed squares (ix, iy) runs from the lower-right to the upper-left corner
    instead (so the node that is the lower-right corner of one flipped square and the upper-left corner of another has 8
    neighbours, 7 with one flip); interior nodes are moved by up to `jitter` * h so that no two mass entries are equal
    (the reference ranks a node's neighbours with np.argsort, whose order of equal keys is unspecified)"""
    with refshim.quiet():
        m = Mesh2D(nsq * nsq)
    p, conn = np.array(m.p), np.array(m.conn)
    W = nsq + 1
    for ix, iy in flips:
        k, e = iy * W + ix, 2 * (iy * nsq + ix)
        assert list(conn[e]) == [k, k + 1, k + W + 1] and list(conn[e + 1]) == [k, k + W + 1, k + W]
        conn[e] = [k, k + 1, k + W]
        conn[e + 1] = [k + 1, k + W + 1, k + W]
    rng = np.random.default_rng(seed)
    interior = (p[:, 0] > 0) & (p[:, 0] < 1) & (p[:, 1] > 0) & (p[:, 1] < 1)
    p[interior] += (rng.random((int(interior.sum()), 2)) - 0.5) * 2 * jitter / nsq
    with refshim.quiet():
        return Mesh2D(p=p, conn=conn)


def run_case(mesh, levels, mean, std):
    q = Quadrature2D(3)
    with refshim.quiet():
        M = MassMatrix(mesh).compute_mass_2d(FunctionTriangle(1), q)
        A = StiffnessMatrix(mesh).compute_stiffness_2d(GradientTriangle(1), q)
        n = M.shape[0]
        nmg = NeuralMG_2D(A, np.zeros((n, 1)), Stub(), M, std, mean)
        out = {}
        mass = nmg.M
        d_neighs = {}
        for i in range(levels - 1):                    # statement by statement define_hierarchy (:748-763)
            if d_neighs:
                mass = nmg.pre_process(mass, d_neighs)
            Mc = sp.coo_matrix(sp.csr_matrix(mass))
            out["l%d_M_row" % i], out["l%d_M_col" % i] = Mc.row.astype(np.int32), Mc.col.astype(np.int32)
            out["l%d_M_data" % i], out["l%d_M_shape" % i] = Mc.data.astype(np.float64), np.array(Mc.shape)
            C, F, Cn, Fn = nmg.coarsening(mass)
            mapp = nmg.map_coarse(C)
            patches, idx_fill = nmg.extract_patches(C, mass)
            out["l%d_C" % i] = np.array(list(C), dtype=np.int64)
            out["l%d_patches" % i] = patches.copy()
            out["l%d_fill" % i] = idx_fill.copy()
            pn = (patches - nmg.mean) / nmg.std
            res = nmg.model.predict(pn)
            out["l%d_pred" % i] = res.copy()
            B, d_neighs = nmg.fill_B(res, idx_fill, mass.shape[0], mapp, C)
            out["l%d_B" % i] = B.copy()
            dn = -np.ones((len(C), 6), dtype=np.int64)
            for k, v in d_neighs.items():
                dn[k, :len(v)] = v
            out["l%d_dn" % i] = dn
            row_sums = B.sum(axis=1)
            Q = lil_matrix(B / row_sums[:, np.newaxis])
            out["l%d_Q" % i] = np.asarray(Q.todense(), dtype=np.float64)
            mass = lil_matrix(Q.T @ mass @ Q)
        Mc = sp.coo_matrix(sp.csr_matrix(mass))
        out["final_M_row"], out["final_M_col"] = Mc.row.astype(np.int32), Mc.col.astype(np.int32)
        out["final_M_data"], out["final_M_shape"] = Mc.data.astype(np.float64), np.array(Mc.shape)
        Ac = sp.coo_matrix(sp.csr_matrix(A))
        out["A_row"], out["A_col"], out["A_data"] = Ac.row.astype(np.int32), Ac.col.astype(np.int32), Ac.data
        out["mesh_p"] = np.asarray(mesh.get_points(), dtype=np.float64)          # inputs of the reference's assembly
        out["mesh_conn"] = np.asarray(mesh.get_connections(), dtype=np.int64)
    return out


def main():
    cases = {}
    mean, std = np.zeros(43), np.ones(43)
    rng = np.random.default_rng(7)
    mean2, std2 = rng.standard_normal(43) * 1e-3, 0.5 + rng.random(43)
    with refshim.quiet():
        structured = Mesh2D(64)           # 8 x 8 squares, 81 nodes
        rect = Mesh2D(60)                 # 6 x 10 squares (find_balanced_couple), 77 nodes
        big = Mesh2D(256)                 # 16 x 16, 289 nodes
    cases["s81"] = run_case(structured, 3, mean, std)
    cases["r77"] = run_case(rect, 3, mean2, std2)
    cases["s289"] = run_case(big, 4, mean, std)
    cases["i81"] = run_case(irregular_mesh(16, 42), 3, mean, std)        # 4x4 squares refined irregularly: 81 nodes
    cases["i289"] = run_case(irregular_mesh(64, 7), 3, mean2, std2)     # 8x8 refined: 289 nodes, parents numbered first
    # coarse nodes with 7 and 8 neighbours: extra patch variants (Multigrid.py:631-663)
    cases["f81"] = run_case(flipped_mesh(8, [(1, 2), (2, 1), (3, 4), (5, 2), (6, 1), (6, 5)], 3), 3, mean, std)
    cases["f169"] = run_case(flipped_mesh(12, [(1, 2), (2, 1), (3, 4), (5, 2), (6, 1), (6, 5), (9, 8), (8, 9), (3, 8),
                                               (4, 7), (9, 2), (7, 10)], 11), 3, mean2, std2)
    flat = {"mean2": mean2, "std2": std2}
    for name, d in cases.items():
        for k, v in d.items():
            flat[name + "__" + k] = v
    path = os.path.join(HERE, "neural_2d_cases.npz")
    np.savez_compressed(path, **flat)
    print(path, os.path.getsize(path))
    for name, d in cases.items():
        print(name, [d[k].shape for k in d if k.endswith("_Q")], [d[k].shape[0] for k in d if k.endswith("_patches")])


if __name__ == "__main__":
    main()
