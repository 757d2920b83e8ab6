"""Transfer operators Q from the coupling operator B with the reference's interface
(learn_multigrid/L2_projection/L2Projection.py:11-90).

1D as the reference: "L2" Q = M^-1 B, "pseudo" Q = diag(colsum M)^-1 B, "quasi" Q = B / rowsum(B).
2D (the reference's compute_transfer_2d is an unfinished stub, :17-24): for NESTED meshes B_h = M_h P with P the
linear interpolation between the meshes (verified against compute_b_1d in 1D), which gives the
quasi / pseudo operators with one sparse product and a row scaling; "L2" needs a sparse solve and stays on the host.
"""
import time

import numpy as np
import scipy.sparse as sp

from .Intersection import Intersection
from .CouplingOperator import CouplingOperator
from ..assembly.MassMatrix import MassMatrix
from ..assembly.Quadrature import Quadrature, Quadrature2D
from ..assembly.ShapeFunction import Function, FunctionTriangle


class L2Projection:

    def __init__(self, type, fine_mesh, coarse_mesh):
        self.type = type
        self.fine_mesh = fine_mesh
        self.coarse_mesh = coarse_mesh

    def compute_transfer_1d(self, sparse=False):
        """returns (Q, seconds spent in the intersection search) like the reference (:26-57)"""
        fine_mesh = self.fine_mesh
        coarse_mesh = self.coarse_mesh
        inter = Intersection(fine_mesh, coarse_mesh)
        start = time.time()
        inter.find_intersections1d()
        timeL = time.time() - start
        q = Quadrature(3)
        phi = Function(2)
        coup_op = CouplingOperator(inter, fine_mesh, coarse_mesh)
        method = self.type_to_method(self.type)
        if not callable(method):
            raise ValueError("unknown projection type %r" % (self.type,))
        if sparse and self.type in ("quasi", "pseudo"):
            B = coup_op.compute_b_1d(q, phi, sparse=True)
            M = MassMatrix(fine_mesh).compute_mass_1d(phi, q, sparse=True)
            if self.type == "quasi":
                s = np.asarray(B.sum(axis=1)).ravel()
            else:
                s = np.asarray(M.sum(axis=0)).ravel()
            return sp.csr_matrix(sp.diags(1.0 / s) @ B), timeL
        B = coup_op.compute_b_1d(q, phi)
        M = MassMatrix(fine_mesh).compute_mass_1d(phi, q)
        L = method(B, M)
        return L, timeL

    def compute_transfer_2d(self, P=None):
        """Q (CSR) between two P1 triangle meshes of the same domain.
        With P (the linear interpolation coarse -> fine of NESTED meshes, n_f x n_c) the coupling operator is
        B_h = M_h P -- one sparse product.  Without it B_h is integrated on the triangle-triangle
        intersections (coupling2d.coupling_operator_2d), which works for non-nested meshes as well.
        "quasi": Q = B / rowsum(B); "pseudo": Q = diag(colsum M)^-1 B; "L2": Q = M^-1 B (sparse solve, dense result:
        small meshes only), as in 1D (:67-90)."""
        M = MassMatrix(self.fine_mesh).compute_mass_2d(FunctionTriangle(1), Quadrature2D(3), format="csr")
        if P is not None:
            B = sp.csr_matrix(M @ sp.csr_matrix(P))
        else:
            from .coupling2d import coupling_operator_2d
            B = coupling_operator_2d(self.fine_mesh, self.coarse_mesh)
        if self.type == "quasi":
            s = np.asarray(B.sum(axis=1)).ravel()
        elif self.type == "pseudo":
            s = np.asarray(M.sum(axis=0)).ravel()
        elif self.type == "L2":
            from scipy.sparse.linalg import splu
            Q = sp.csr_matrix(splu(sp.csc_matrix(M)).solve(B.toarray()))
            Q.sort_indices()
            return Q
        else:
            raise ValueError("unknown projection type %r" % (self.type,))
        Q = sp.csr_matrix(sp.diags(1.0 / s) @ B)
        Q.sort_indices()
        return Q

    def type_to_method(self, type):
        switcher = {
            "L2": self.compute_l2_1d,
            "pseudo": self.compute_pseudo_1d,
            "quasi": self.compute_quasi_1d,
        }
        return switcher.get(type, "Invalid order")

    @staticmethod
    def compute_l2_1d(B, M):
        inv_M = np.linalg.inv(M)
        return np.dot(inv_M, B)

    @staticmethod
    def compute_pseudo_1d(B, M):
        row_sum = np.sum(M, axis=0)
        diag = np.diag(row_sum)
        return np.linalg.solve(diag, B)

    @staticmethod
    def compute_quasi_1d(B, _):
        row_sums = B.sum(axis=1)
        return B / row_sums[:, np.newaxis]
