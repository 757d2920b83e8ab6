"""Coupling operator B[f,c] = int phi_f phi_c with the reference's interface
(learn_multigrid/L2_projection/CouplingOperator.py:7-69), vectorised over the intersection segments."""
import numpy as np
import scipy.sparse as sp

from ..assembly.MapReferenceElement import g_function, inv_g_function


class CouplingOperator:

    def __init__(self, intersections, fine_mesh, coarse_mesh):
        self.inter = intersections
        self.coarse_mesh = coarse_mesh
        self.fine_mesh = fine_mesh

    def op_l2g_1d(self):
        intersections, _, _ = self.inter.get_info()
        fine_l2g = np.stack((intersections[:, 0], intersections[:, 0] + 1), axis=1).astype(int)
        coarse_l2g = np.stack((intersections[:, 1], intersections[:, 1] + 1), axis=1).astype(int)
        return fine_l2g, coarse_l2g

    def compute_b_1d(self, q, phi, sparse=False):
        """loc_B[i,j] = (x_b - x_a) * sum_k phi_i(xi_f,k) phi_j(xi_c,k) w_k per intersection segment, 3-point
        Gauss points mapped to the segment and back to each element's reference coordinate (:31-69)."""
        intersections, int_coord, _ = self.inter.get_info()
        conn = self.fine_mesh.get_connections()
        coarse_conn = self.coarse_mesh.get_connections()
        p = q.get_points()
        w = q.get_weights()
        fine_l2g, coarse_l2g = self.op_l2g_1d()
        x_a = int_coord[:, 0]
        x_b = int_coord[:, 1]
        fa, fb = conn[intersections[:, 0], 0], conn[intersections[:, 0], 1]
        ca, cb = coarse_conn[intersections[:, 1], 0], coarse_conn[intersections[:, 1], 1]
        K = len(intersections)
        loc = np.zeros((K, 2, 2))
        fine_ref = [inv_g_function(g_function(p[k], x_a, x_b), fa, fb) for k in range(len(p))]
        coarse_ref = [inv_g_function(g_function(p[k], x_a, x_b), ca, cb) for k in range(len(p))]
        for i in range(2):
            for j in range(2):
                res = 0
                for k in range(len(p)):
                    res = res + phi.evaluate(fine_ref[k], i) * phi.evaluate(coarse_ref[k], j) * w[k]
                loc[:, i, j] = (x_b - x_a) * res
        rows = np.repeat(fine_l2g, 2, axis=1).reshape(-1)
        cols = np.tile(coarse_l2g, (1, 2)).reshape(-1)
        shape = (self.fine_mesh.get_np(), self.coarse_mesh.get_np())
        B = sp.coo_matrix((loc.reshape(-1), (rows, cols)), shape=shape).tocsr()
        return B if sparse else B.toarray()

    def compute_b_2d(self):
        """B[f, c] = int phi_f phi_c on the triangle-triangle intersections (no reference counterpart: the 2D path of
        the reference is a stub; see coupling2d.py).  CSR."""
        from .coupling2d import coupling_operator_2d
        return coupling_operator_2d(self.fine_mesh, self.coarse_mesh)
