"""Generate tests/golden/mesh_embedding.npz: the REFERENCE's Mesh2D.embedding() (learn_multigrid/mesh/Mesh2D.py:
162-431, two layers of ghost elements around the unit square) run through oracle/refshim.py on structured,
rectangular, regularly refined and irregularly refined meshes.  Run in the authoring container:
    python tests/golden/make_golden_mesh.py

Everything stored is produced by the reference's own Mesh2D methods.  refine() needs shim 4 of oracle/refshim.py
under NumPy >= 1.24 (ragged list assignment, Mesh2D.py:149-156); np.random is seeded before every construction.
Per embedding case: in_p, in_conn (input mesh), p, conn (embedded).  Per refine case `ref_*`: ne, regular, seed,
times, p, conn after `times` calls of refine(regular).
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")

from oracle import refshim  # noqa: E402

refshim.install()
with refshim.quiet():
    Mesh2D = refshim.patch_refine().Mesh2D


def refined(ne, regular, seed, times=1):
    with refshim.quiet():
        np.random.seed(seed)
        m = Mesh2D(ne)
        for _ in range(times):
            m.refine(regular=regular)
    return np.array(m.p), np.array(m.conn)


def main():
    inputs = {}
    with refshim.quiet():
        for ne in (4, 12, 16, 60, 64):
            m = Mesh2D(ne)
            inputs["s%d" % ne] = (np.array(m.p), np.array(m.conn))
    inputs["r4"] = refined(4, True, 0)
    inputs["i16x2"] = refined(16, False, 42, times=2)
    inputs["i12"] = refined(12, False, 3)
    out = {}
    for name, (p, conn) in inputs.items():
        with refshim.quiet():
            e = Mesh2D(p=p.copy(), conn=conn.copy()).embedding()
        out[name + "_in_p"], out[name + "_in_conn"] = p, conn.astype(np.int32)
        out[name + "_p"], out[name + "_conn"] = np.array(e.p), np.array(e.conn).astype(np.int32)
        print(name, len(p), len(conn), "->", e.n_p, e.ne)
    for ne, regular, seed, times in ((4, True, 0, 2), (16, False, 42, 2), (12, False, 3, 1), (60, False, 1, 1)):
        p, conn = refined(ne, regular, seed, times)
        name = "ref_%d_%d_%d_%d" % (ne, int(regular), seed, times)
        out[name + "_p"], out[name + "_conn"] = p, conn.astype(np.int32)
        print(name, p.shape, conn.shape)
    np.savez_compressed(os.path.join(HERE, "mesh_embedding.npz"), **out)


if __name__ == "__main__":
    main()
