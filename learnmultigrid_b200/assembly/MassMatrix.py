"""P1 mass matrix with the reference's interface (learn_multigrid/assembly/MassMatrix.py:5-82), vectorised
over elements.  1D returns a dense (n,n) array like the reference (or CSR with sparse=True); 2D returns a
lil_matrix like the reference (or CSR with format="csr")."""
import numpy as np
import scipy.sparse as sp
from scipy.sparse import lil_matrix

from ._element import (element_jacobians, load_triplets, local_matrix, save_triplets, scatter_elements,
                       triangle_jacobian)


class MassMatrix:

    jacobian = staticmethod(triangle_jacobian)

    def __init__(self, mesh):
        self.mesh = mesh
        self.J = self.jacobian
        self.M = lil_matrix([])

    def compute_mass_2d(self, phi, q, format="lil"):
        """loc_M[i,j] = detJ * sum_k phi_i(p_k) phi_j(p_k) w_k  (MassMatrix.py:21-35, 52-59).
        format="device": assembled by CUDA kernels (assembly_device.py), returns a device CSR (setup_device.DevCSR)."""
        if format == "device":
            from ..assembly_device import DeviceAssembler
            asm = DeviceAssembler()
            self.M = asm.mass(*asm.mesh_to_device(self.mesh), phi, q)
            return self.M
        n_p = self.mesh.get_np()
        p = self.mesh.get_points()
        conn = np.asarray(self.mesh.get_connections())
        *_, det = element_jacobians(p, conn)
        c = np.array([[q.compute(phi, np.array([i, j])) for j in range(3)] for i in range(3)], dtype=float)
        loc = det[:, None, None] * c[None, :, :]
        self.M = scatter_elements(conn, loc, n_p, format)
        return self.M

    def save(self, path="../data/matrices/M"):
        save_triplets(path, self.M)

    def load(self, path):
        self.M = load_triplets(path)
        return self.M

    @staticmethod
    def loc_m_2d(d_J, phi, q):
        return local_matrix(3, d_J, lambda i, j: q.compute(phi, np.array([i, j])))

    def compute_mass_1d(self, phi, q, sparse=False):
        """loc_M[i,j] = (right-left) * sum_k phi_i phi_j w_k  (MassMatrix.py:61-82)"""
        conn = self.mesh.get_connections()
        n_points = self.mesh.get_np()
        h = conn[:, 1] - conn[:, 0]
        c = np.array([[q.compute(phi, np.array([i, j])) for j in range(2)] for i in range(2)], dtype=float)
        return _assemble_1d(h[:, None, None] * c[None], n_points, sparse)

    @staticmethod
    def loc_m_1d(phi, q, left, right):
        return local_matrix(2, right - left, lambda i, j: q.compute(phi, np.array([i, j])))


def _assemble_1d(loc, n_points, sparse):
    """tridiagonal assembly of per-element 2x2 blocks in element order (dense like the reference, or CSR)"""
    ne = loc.shape[0]
    diag = np.zeros(n_points)
    diag[:-1] += loc[:, 0, 0]
    diag[1:] += loc[:, 1, 1]
    upper = loc[:, 0, 1].copy()
    lower = loc[:, 1, 0].copy()
    if sparse:
        return sp.diags([lower, diag, upper], [-1, 0, 1], shape=(n_points, n_points), format="csr")
    M = np.zeros(shape=(n_points, n_points))
    i = np.arange(ne)
    M[i, i] = diag[:-1]
    M[n_points - 1, n_points - 1] = diag[-1]
    M[i, i + 1] = upper
    M[i + 1, i] = lower
    return M
