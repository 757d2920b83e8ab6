"""time the device coupling operator (L2_projection/coupling2d.py, csrc/assembly_kernels.cu) on a 131 k x 20 k triangle pair"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from learnmultigrid_b200 import problems as P
from learnmultigrid_b200.mesh.Mesh2D import Mesh2D
from learnmultigrid_b200.L2_projection import coupling2d as C
pb = P.irregular_p1_2d(256, seed=1)
coarse = Mesh2D(100 * 100)
pf, tf = np.asarray(pb["mesh"].get_points()), np.asarray(pb["mesh"].get_connections())
pc, tc = np.asarray(coarse.get_points()), np.asarray(coarse.get_connections())
t = time.perf_counter(); f, c = C.candidate_pairs(pf, tf, pc, tc); t_pairs = time.perf_counter() - t
for rep in range(2):
    torch.cuda.synchronize(); t = time.perf_counter()
    B = C.coupling_operator_2d_native(pb["mesh"], coarse)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
print("COUPLING fine tris %d coarse tris %d candidate pairs %d (host binning %.2f s) nnz %d total native call %.3f s" % (len(tf), len(tc), len(f), t_pairs, B.nnz, dt))
