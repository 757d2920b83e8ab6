"""1D finite-difference Laplacians (learn_multigrid/utilities/laplacian.py:9-59)."""
import numpy as np
from scipy.sparse import spdiags

from ..assembly.LoadVector import LoadVector


def laplacian_1d_fd(x, N, f):
    h = (x[-1] - x[0]) / (N - 1)
    X = x[1:N - 1].reshape((N - 2, 1))
    rhs = f(X)
    L = spdiags([[-1] * (N - 2), [2] * (N - 2), [-1] * (N - 2)], [-1, 0, 1], (N - 2), (N - 2))
    return (1 / h ** 2) * L, X, rhs


def laplacian_1d_fd_bc(m, f):
    x = m.get_mesh()
    N = m.get_np()
    h = (x[-1] - x[0]) / (N - 1)
    X = x.reshape((N, 1))
    rhs = LoadVector(m).compute_rhs_1d(f)
    rhs[0] = 0
    rhs[-1] = 0
    L = spdiags([[-1] * N, [2] * N, [-1] * N], [-1, 0, 1], N, N).tolil()
    L = (1 / h ** 2) * L
    L[1, 0] = 0
    L[-2, -1] = 0
    L[0, :] = 0
    L[-1, :] = 0
    L[0, 0] = 1
    L[-1, -1] = 1
    return L, X, rhs
