"""P1 stiffness matrix with the reference's interface (learn_multigrid/assembly/StiffnessMatrix.py:5-82),
vectorised over elements.  `coefficient` (optional, new: the reference has no variable coefficient,
test/thesis_variableCoeff_stiff.py is empty) is a callable k(x, y) evaluated at element centroids."""
import numpy as np
import scipy.sparse as sp
from scipy.sparse import coo_matrix, lil_matrix

from ._element import element_jacobians, scatter_elements
from .MassMatrix import _assemble_1d


class StiffnessMatrix:

    @staticmethod
    def jacobian(x, y):
        J = np.zeros((2, 2))
        J[0, 0] = x[1] - x[0]
        J[0, 1] = x[2] - x[0]
        J[1, 0] = y[1] - y[0]
        J[1, 1] = y[2] - y[0]
        return J

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
        x_coo = sp.coo_matrix(self.A)
        np.savez(path, row=x_coo.row, col=x_coo.col, data=x_coo.data, shape=x_coo.shape)

    def load(self, path):
        y = np.load(path)
        z = coo_matrix((y['data'], (y['row'], y['col'])), shape=y['shape'])
        z = lil_matrix(z)
        self.A = z
        return z

    @staticmethod
    def loc_a_2d(d_J, jac_inv, d_phi, q):
        loc_A = np.zeros((3, 3))
        for i in range(0, 3):
            for j in range(0, 3):
                loc_A[i, j] = d_J * q.compute_grad(d_phi, jac_inv, np.array([i, j]))
        return loc_A

    def compute_stiffness_1d(self, dphi, q, sparse=False):
        """loc_A[i,j] = 1/(right-left) * sum_k dphi_i dphi_j w_k  (StiffnessMatrix.py:61-82)"""
        conn = self.mesh.get_connections()
        n_points = self.mesh.get_np()
        h = conn[:, 1] - conn[:, 0]
        c = np.array([[q.compute(dphi, np.array([i, j])) for j in range(2)] for i in range(2)], dtype=float)
        return _assemble_1d((1 / h)[:, None, None] * c[None], n_points, sparse)

    @staticmethod
    def loc_a_1d(dphi, q, left, right):
        locA = np.zeros(shape=(2, 2))
        for i in range(0, 2):
            for j in range(0, 2):
                locA[i, j] = 1 / (right - left) * q.compute(dphi, np.array([i, j]))
        return locA
