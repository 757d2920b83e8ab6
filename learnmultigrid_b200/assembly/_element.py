"""Vectorised element geometry shared by the 2D assembly classes."""
import numpy as np
import scipy.sparse as sp


def element_jacobians(p, conn):
    """J per element with the reference's layout (MassMatrix.jacobian, MassMatrix.py:8-14):
    J = [[x1-x0, x2-x0], [y1-y0, y2-y0]]; returns J00, J01, J10, J11, detJ (signed, not abs: :31)."""
    x = p[conn, 0]
    y = p[conn, 1]
    J00 = x[:, 1] - x[:, 0]
    J01 = x[:, 2] - x[:, 0]
    J10 = y[:, 1] - y[:, 0]
    J11 = y[:, 2] - y[:, 0]
    det = J00 * J11 - J01 * J10
    return J00, J01, J10, J11, det


def scatter_elements(conn, loc, n, fmt):
    """global matrix from per-element 3x3 blocks loc[e, i, j]; exact zeros are not stored, exactly like the
    reference's lil_matrix accumulation (M[ix_(l2g,l2g)] += loc)."""
    rows = np.repeat(conn, 3, axis=1).reshape(-1)
    cols = np.tile(conn, (1, 3)).reshape(-1)
    M = sp.coo_matrix((loc.reshape(-1), (rows, cols)), shape=(n, n)).tocsr()
    M.sum_duplicates()
    M.eliminate_zeros()
    M.sort_indices()
    if fmt == "lil":
        return M.tolil()
    if fmt == "csr":
        return M
    raise ValueError("format must be 'lil' or 'csr'")


def triangle_jacobian(x, y):
    """Jacobian of the affine map from the unit triangle to the triangle with vertex coordinates x[0..2], y[0..2]:
    its columns are the edge vectors from vertex 0 (MassMatrix.jacobian, MassMatrix.py:8-14)."""
    return np.array([[x[1] - x[0], x[2] - x[0]],
                     [y[1] - y[0], y[2] - y[0]]], dtype=float)


def local_matrix(n, scale, entry):
    """n x n element matrix  scale * entry(i, j)  (the reference's loc_* helpers)"""
    return np.array([[scale * entry(i, j) for j in range(n)] for i in range(n)], dtype=float)


def save_triplets(path, M):
    """matrix as COO triplets in an .npz file (the layout of the reference's save(), MassMatrix.py:37-43)"""
    coo = sp.coo_matrix(M)
    np.savez(path, row=coo.row, col=coo.col, data=coo.data, shape=coo.shape)


def load_triplets(path):
    """lil_matrix from a file written by save_triplets / by the reference's save()"""
    with np.load(path) as z:
        return sp.coo_matrix((z["data"], (z["row"], z["col"])), shape=tuple(z["shape"])).tolil()

