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


def _scaling(kind, B, M):
    """the diagonal a sparse coupling operator is divided by: its row sums ("quasi") or the lumped mass ("pseudo")"""
    if kind == "quasi":
        return np.asarray(B.sum(axis=1)).ravel()
    if kind == "pseudo":
        return np.asarray(M.sum(axis=0)).ravel()
    raise ValueError("unknown projection type %r" % (kind,))


def _row_scaled(B, s):
    Q = sp.csr_matrix(sp.diags(1.0 / s) @ B)
    Q.sort_indices()
    return Q


class L2Projection:
    KINDS = ("L2", "pseudo", "quasi")

    def __init__(self, type, fine_mesh, coarse_mesh):
        self.type = type
        self.fine_mesh = fine_mesh
        self.coarse_mesh = coarse_mesh

    # ---- 1D (dense like the reference, :26-90) -------------------------------------------------------------
    def compute_transfer_1d(self, sparse=False):
        """(Q, seconds spent in the intersection search).  sparse=True builds "quasi" / "pseudo" operators in CSR
        without the dense n x n detour."""
        method = self.type_to_method(self.type)
        if not callable(method):
            raise ValueError("unknown projection type %r" % (self.type,))
        inter = Intersection(self.fine_mesh, self.coarse_mesh)
        t0 = time.time()
        inter.find_intersections1d()
        seconds = time.time() - t0
        rule, hat = Quadrature(3), Function(2)
        coupling = CouplingOperator(inter, self.fine_mesh, self.coarse_mesh)
        mass = MassMatrix(self.fine_mesh)
        if sparse and self.type != "L2":
            B = coupling.compute_b_1d(rule, hat, sparse=True)
            M = mass.compute_mass_1d(hat, rule, sparse=True)
            return _row_scaled(B, _scaling(self.type, B, M)), seconds
        return method(coupling.compute_b_1d(rule, hat), mass.compute_mass_1d(hat, rule)), seconds

    def type_to_method(self, type):
        return getattr(self, "compute_%s_1d" % type.lower(), "Invalid order") if type in self.KINDS else "Invalid order"

    @staticmethod
    def compute_l2_1d(B, M):
        """Q = M^-1 B through the explicit inverse, as the reference forms it (:67-71; same LAPACK calls, same bits)"""
        return np.dot(np.linalg.inv(M), B)

    @staticmethod
    def compute_pseudo_1d(B, M):
        """Q = diag(column sums of M)^-1 B by a dense solve with the diagonal matrix (:73-81)"""
        return np.linalg.solve(np.diag(np.sum(M, axis=0)), B)

    @staticmethod
    def compute_quasi_1d(B, _):
        """Q = B with every row divided by its sum (:83-90)"""
        return B / B.sum(axis=1)[:, np.newaxis]

    # ---- 2D ------------------------------------------------------------------------------------------------
    def compute_transfer_2d(self, P=None, B=None):
        """Q (CSR) between two P1 triangle meshes of the same domain.
        With P (the linear interpolation coarse -> fine of NESTED meshes, n_f x n_c) the coupling operator is
        B_h = M_h P -- one sparse product.  Without it B_h is integrated on the triangle-triangle
        intersections ON THE DEVICE (coupling2d.coupling_operator_2d_native), which works for non-nested meshes as
        well; a coupling operator computed elsewhere can be handed in as B (the CPU tests pass the NumPy model's).
        "quasi": Q = B / rowsum(B); "pseudo": Q = diag(colsum M)^-1 B; "L2": Q = M^-1 B (sparse solve, dense result:
        small meshes only), as in 1D."""
        if self.type not in self.KINDS:
            raise ValueError("unknown projection type %r" % (self.type,))
        M = MassMatrix(self.fine_mesh).compute_mass_2d(FunctionTriangle(1), Quadrature2D(3), format="csr")
        if B is not None:
            B = sp.csr_matrix(B)
        elif P is not None:
            B = sp.csr_matrix(M @ sp.csr_matrix(P))
        else:
            # on the device: candidate pairs by bounding-box binning, clipping / integration per pair, fold
            # (csrc/assembly_kernels.cu); coupling2d.coupling_operator_2d is the NumPy model the tests check it against
            from .coupling2d import coupling_operator_2d_native
            from ..setup_device import DeviceSetup
            from .. import _lib
            Bd = coupling_operator_2d_native(self.fine_mesh, self.coarse_mesh)
            torch = _lib.require_cuda()
            B = DeviceSetup(torch, Bd.values.device).download(Bd)
            B = sp.csr_matrix(B)
            B.sort_indices()
        if self.type == "L2":
            from scipy.sparse.linalg import splu
            Q = sp.csr_matrix(splu(sp.csc_matrix(M)).solve(B.toarray()))
            Q.sort_indices()
            return Q
        return _row_scaled(B, _scaling(self.type, B, M))
