"""Load vector with the reference's interface (learn_multigrid/assembly/LoadVector.py:4-62), vectorised."""
import numpy as np

from ._element import element_jacobians


class LoadVector:

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
        for i in range(3):                       # element order inside each local index
            np.add.at(rhs, conn[:, i], loc[:, i])
        self.rhs = rhs.reshape(n_p, 1)
        return self.rhs

    def save(self, path="../data/matrices/rhs"):
        np.save(path, self.rhs)

    def load(self, path):
        self.rhs = np.load(path)
        return self.rhs

    @staticmethod
    def loc_rhs_2d(d_J, fun, phi, q):
        locRHS = np.zeros(shape=(3, 1))
        for i in range(0, 3):
            locRHS[i] = d_J * q.compute_single(phi, i, fun)
        return locRHS

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
