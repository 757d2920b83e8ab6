"""P1 stiffness matrix with the reference's interface (learn_multigrid/assembly/StiffnessMatrix.py:5-82),
vectorised over elements.  `coefficient` (optional, new: the reference has no variable coefficient,
test/thesis_variableCoeff_stiff.py is empty) is a callable k(x, y) evaluated at element centroids."""
import numpy as np
import scipy.sparse as sp
from scipy.sparse import lil_matrix

from ._element import (element_jacobians, load_triplets, local_matrix, save_triplets, scatter_elements,
                       triangle_jacobian)
from .MassMatrix import _assemble_1d


class StiffnessMatrix:

    jacobian = staticmethod(triangle_jacobian)

    def __init__(self, mesh):
        self.mesh = mesh
        self.J = self.jacobian
        self.A = lil_matrix([])

    def compute_stiffness_2d(self, d_phi, q, format="lil", coefficient=None):
        """loc_A[i,j] = detJ * sum_k ((J^-T g_i)^T J^-T) g_j * w[i]   (StiffnessMatrix.py:21-36, 53-59 and
        Quadrature2D.compute_grad, Quadrature.py:72-82, same association of the products)."""
        if format == "device":             # CUDA kernels (assembly_device.py); returns a device CSR (setup_device.DevCSR)
            from ..assembly_device import DeviceAssembler
            asm = DeviceAssembler()
            self.A = asm.stiffness(*asm.mesh_to_device(self.mesh), d_phi, q, coefficient)
            return self.A
        n_p = self.mesh.get_np()
        p = self.mesh.get_points()
        conn = np.asarray(self.mesh.get_connections())
        J00, J01, J10, J11, det = element_jacobians(p, conn)
        # inv = transpose(inverse(J)):  J^-1 = [[J11, -J01], [-J10, J00]] / det
        T00 = J11 / det
        T01 = -J10 / det
        T10 = -J01 / det
        T11 = J00 / det
        pts = q.get_points()
        w = q.get_weights()
        g = [np.asarray(d_phi.evaluate(pts[0], i), dtype=float).reshape(2) for i in range(3)]
        ne = len(conn)
        loc = np.zeros((ne, 3, 3))
        for i in range(3):
            a0 = T00 * g[i][0] + T01 * g[i][1]          # a = J^-T g_i
            a1 = T10 * g[i][0] + T11 * g[i][1]
            t0 = a0 * T00 + a1 * T10                    # t = a^T J^-T
            t1 = a0 * T01 + a1 * T11
            for j in range(3):
                s = (t0 * g[j][0] + t1 * g[j][1]) * w[i]
                res = 0
                for _ in range(len(pts)):
                    res = res + s
                loc[:, i, j] = det * res
        if coefficient is not None:
            cx = p[conn, 0].mean(axis=1)
            cy = p[conn, 1].mean(axis=1)
            loc *= np.asarray(coefficient(cx, cy), dtype=float)[:, None, None]
        self.A = scatter_elements(conn, loc, n_p, format)
        return self.A

    def save(self, path="../data/matrices/A"):
        save_triplets(path, self.A)

    def load(self, path):
        self.A = load_triplets(path)
        return self.A

    @staticmethod
    def loc_a_2d(d_J, jac_inv, d_phi, q):
        return local_matrix(3, d_J, lambda i, j: q.compute_grad(d_phi, jac_inv, np.array([i, j])))

    def compute_stiffness_1d(self, dphi, q, sparse=False):
        """loc_A[i,j] = 1/(right-left) * sum_k dphi_i dphi_j w_k  (StiffnessMatrix.py:61-82)"""
        conn = self.mesh.get_connections()
        n_points = self.mesh.get_np()
        h = conn[:, 1] - conn[:, 0]
        c = np.array([[q.compute(dphi, np.array([i, j])) for j in range(2)] for i in range(2)], dtype=float)
        return _assemble_1d((1 / h)[:, None, None] * c[None], n_points, sparse)

    @staticmethod
    def loc_a_1d(dphi, q, left, right):
        return local_matrix(2, 1 / (right - left), lambda i, j: q.compute(dphi, np.array([i, j])))
