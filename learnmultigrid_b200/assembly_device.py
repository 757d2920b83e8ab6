"""P1 assembly on the device (csrc/assembly_kernels.cu): mass, stiffness (optionally with a per-element coefficient),
load vector and Dirichlet rows, from a mesh given as points + connectivity.  The reference loops over elements in
Python and adds 3x3 blocks into a lil_matrix (MassMatrix.py:21-35, StiffnessMatrix.py:21-36, LoadVector.py:20-51);
the host classes in learnmultigrid_b200/assembly vectorise that with NumPy; this module is the same arithmetic as
kernels, producing device CSR that DeviceHierarchy / NeuralBuilder consume without a host round trip.

The element constants (shape-function products at the quadrature points, with the reference's 14-digit rounded nodes
and weights) are evaluated once on the host through the same Quadrature / ShapeFunction objects the reference API
takes, so a caller's custom `phi`, `q` or load function keep working.
"""
import ctypes

import numpy as np

from . import _lib
from . import setup_device as SD


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class DeviceAssembler:

    def __init__(self, setup=None):
        torch = _lib.require_cuda()
        self.torch = torch
        self.lib = _lib.load()
        self.S = setup if setup is not None else SD.DeviceSetup(torch, torch.device("cuda", torch.cuda.current_device()))
        self.dev = self.S.dev

    def st(self):
        return _lib.stream_handle(self.torch)

    def mesh_to_device(self, mesh_or_p, conn=None):
        t = self.torch
        if conn is None:
            p, conn = mesh_or_p.get_points(), mesh_or_p.get_connections()
        else:
            p = mesh_or_p
        p = np.ascontiguousarray(np.asarray(p, dtype=np.float64).reshape(-1, 2))
        conn = np.ascontiguousarray(np.asarray(conn).reshape(-1, 3).astype(np.int32))
        return t.from_numpy(p).to(self.dev), t.from_numpy(conn).to(self.dev), len(p), len(conn)

    def _csr_from_contributions(self, rows, cols, vals, n):
        t, S = self.torch, self.S
        m = rows.numel()
        order = S.row_col_order(rows, cols, n, n)
        head, folded = S.empty(m, t.int32), S.empty(m, t.float64)
        _lib.check(self.lib.mg_coo_fold_sum(m, rows.data_ptr(), cols.data_ptr(), vals.data_ptr(), order.data_ptr(),
                                            head.data_ptr(), folded.data_ptr(), self.st()), "mg_coo_fold_sum")
        slot, nnz = S.scan(head, m)
        orow, ocol, oval = S.empty(nnz, t.int32), S.empty(nnz, t.int32), S.empty(nnz, t.float64)
        _lib.check(self.lib.mg_nn_emit(m, rows.data_ptr(), cols.data_ptr(), order.data_ptr(), head.data_ptr(),
                                       slot.data_ptr(), folded.data_ptr(), orow.data_ptr(), ocol.data_ptr(),
                                       oval.data_ptr(), self.st()), "mg_nn_emit")
        indptr = t.searchsorted(orow, t.arange(n + 1, dtype=t.int32, device=self.dev)).to(t.int32)
        return SD.DevCSR((n, n), indptr, ocol, oval)

    def _assemble(self, d_p, d_conn, n_p, ne, kind, const, coef):
        t, S = self.torch, self.S
        rows, cols, vals = S.empty(9 * ne, t.int32), S.empty(9 * ne, t.int32), S.empty(9 * ne, t.float64)
        const = np.ascontiguousarray(const, dtype=np.float64)
        _lib.check(self.lib.mg_assemble_p1_2d(ne, d_p.data_ptr(), d_conn.data_ptr(), kind, _ptr(const), _lib.ptr(coef),
                                              rows.data_ptr(), cols.data_ptr(), vals.data_ptr(), self.st()),
                   "mg_assemble_p1_2d")
        return self._csr_from_contributions(rows, cols, vals, n_p)

    def mass(self, d_p, d_conn, n_p, ne, phi, q):
        """MassMatrix.compute_mass_2d (MassMatrix.py:21-35)"""
        c = np.array([[q.compute(phi, np.array([i, j])) for j in range(3)] for i in range(3)], dtype=float)
        return self._assemble(d_p, d_conn, n_p, ne, 0, c.reshape(-1), None)

    def stiffness(self, d_p, d_conn, n_p, ne, d_phi, q, coefficient=None):
        """StiffnessMatrix.compute_stiffness_2d (StiffnessMatrix.py:21-36); coefficient(x, y) at element centroids"""
        t = self.torch
        pts, w = q.get_points(), q.get_weights()
        g = np.array([np.asarray(d_phi.evaluate(pts[0], i), dtype=float).reshape(2) for i in range(3)])
        const = np.concatenate([g.reshape(-1), np.asarray(w, dtype=float).reshape(3), [float(len(pts))]])
        coef = None
        if coefficient is not None:
            xy = d_p[d_conn.long()]                        # (ne, 3, 2)
            cx, cy = xy[:, :, 0].mean(dim=1), xy[:, :, 1].mean(dim=1)
            k = coefficient(cx.cpu().numpy(), cy.cpu().numpy())
            coef = t.from_numpy(np.ascontiguousarray(np.asarray(k, dtype=np.float64))).to(self.dev)
        return self._assemble(d_p, d_conn, n_p, ne, 1, const, coef)

    def load(self, d_p, d_conn, n_p, ne, fun, phi, q):
        """LoadVector.compute_rhs_2d (LoadVector.py:20-51), contributions added in element order; returns (n_p,)"""
        t, S = self.torch, self.S
        c = np.array([q.compute_single(phi, i, fun) for i in range(3)], dtype=float)
        nodes, vals = S.empty(3 * ne, t.int32), S.empty(3 * ne, t.float64)
        _lib.check(self.lib.mg_assemble_load_p1_2d(ne, d_p.data_ptr(), d_conn.data_ptr(), _ptr(c), nodes.data_ptr(),
                                                   vals.data_ptr(), self.st()), "mg_assemble_load_p1_2d")
        m = 3 * ne
        order = S.row_col_order(nodes, None, n_p, 1)
        head, folded = S.empty(m, t.int32), S.empty(m, t.float64)
        _lib.check(self.lib.mg_coo_fold_sum(m, nodes.data_ptr(), None, vals.data_ptr(), order.data_ptr(),
                                            head.data_ptr(), folded.data_ptr(), self.st()), "mg_coo_fold_sum")
        out = t.zeros(n_p, dtype=t.float64, device=self.dev)
        _lib.check(self.lib.mg_vector_from_runs(m, nodes.data_ptr(), order.data_ptr(), head.data_ptr(),
                                                folded.data_ptr(), out.data_ptr(), self.st()), "mg_vector_from_runs")
        return out

    def dirichlet(self, A, boundary, rhs=None):
        """A[nodes,:] = I[nodes,:]; rhs[nodes] = 0 (thesis_structured_2d.py:407-414)"""
        t, S = self.torch, self.S
        n = A.shape[0]
        flag = t.zeros(n, dtype=t.int32, device=self.dev)
        b = t.from_numpy(np.ascontiguousarray(np.asarray(boundary, dtype=np.int64))).to(self.dev)
        flag[b] = 1
        count = S.empty(n, t.int32)
        _lib.check(self.lib.mg_csr_dirichlet_count(n, A.indptr.data_ptr(), flag.data_ptr(), count.data_ptr(), self.st()),
                   "mg_csr_dirichlet_count")
        optr, total = S.scan(count, n)
        oidx, oval = S.empty(total, t.int32), S.empty(total, t.float64)
        _lib.check(self.lib.mg_csr_dirichlet_fill(n, *A.ptrs(), flag.data_ptr(), optr.data_ptr(), oidx.data_ptr(),
                                                  oval.data_ptr(), self.st()), "mg_csr_dirichlet_fill")
        if rhs is not None:
            rhs[b] = 0.0
        return SD.DevCSR(A.shape, optr, oidx, oval)
