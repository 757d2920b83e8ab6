"""Load vector with the reference's interface (learn_multigrid/assembly/LoadVector.py:4-62), vectorised."""
import numpy as np

from ._element import element_jacobians, triangle_jacobian


class LoadVector:

    jacobian = staticmethod(triangle_jacobian)

    def __init__(self, mesh):
        self.mesh = mesh
        self.J = self.jacobian
        self.rhs = np.array([])

    def compute_rhs_2d(self, fun, phi, q):
        """loc_rhs[i] = detJ * sum_k phi_i(p_k) f(p_k) w_k with f evaluated at the REFERENCE points, exactly as
        the reference does (LoadVector.py:20-51, Quadrature.compute_single :42-48)."""
        n_p = self.mesh.get_np()
        p = self.mesh.get_points()
        conn = np.asarray(self.mesh.get_connections())
        *_, det = element_jacobians(p, conn)
        c = np.array([q.compute_single(phi, i, fun) for i in range(3)], dtype=float)
        rhs = np.zeros(n_p)
        loc = det[:, None] * c[None, :]
        np.add.at(rhs, conn.reshape(-1), loc.reshape(-1))      # element by element, like the reference's loop (:25-33)
        self.rhs = rhs.reshape(n_p, 1)
        return self.rhs

    def save(self, path="../data/matrices/rhs"):
        np.save(path, self.rhs)                  # .npy, as the reference writes it (LoadVector.py:36-43)

    def load(self, path):
        self.rhs = np.load(path)
        return self.rhs

    @staticmethod
    def loc_rhs_2d(d_J, fun, phi, q):
        return np.array([[d_J * q.compute_single(phi, i, fun)] for i in range(3)], dtype=float)

    def compute_rhs_1d(self, f):
        """per-element trapezoid rule (LoadVector.py:53-62)"""
        x = np.asarray(self.mesh.get_mesh(), dtype=float)
        n_points = self.mesh.get_np()
        fx = np.asarray(f(x), dtype=float).reshape(-1) * np.ones(n_points)
        h = x[1:] - x[:-1]
        b = np.zeros(n_points)
        b[:-1] += fx[:-1] * h / 2
        b[1:] += fx[1:] * h / 2
        return b.reshape(n_points, 1)
