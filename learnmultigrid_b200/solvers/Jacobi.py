"""Jacobi with the reference's interface (learn_multigrid/solvers/Jacobi.py:8-37) on the GPU."""
import numpy as np

from .. import _lib
from .Solver import IterativeSolver


class Jacobi(IterativeSolver):

    def __init__(self, matrix, rhs):
        super().__init__(matrix, rhs)
        self.label = "Jacobi"

    def solve(self, max_iterations=1000, error=1e-12, initial_guess=None, omega=1.0):
        """x <- x + omega * D^-1 (b - A x) until ||b - A x||_2 <= error (absolute), at most max_iterations
        residual evaluations.  omega=1 is the reference (Jacobi.py:35)."""
        d = self._device_csr()
        torch, lib, n = d["torch"], d["lib"], d["n"]
        if initial_guess is None:
            x0 = np.zeros(shape=(self.get_dimension(), 1))
        else:
            x0 = initial_guess
        x = self._upload(x0)
        xo = torch.empty_like(x)
        b = self._upload(self.rhs)
        r = torch.empty_like(x)
        with np.errstate(divide="ignore"):
            dinv = self._upload(1.0 / d["host"].diagonal())
        st = _lib.stream_handle(torch)
        track = []
        for _ in range(0, max_iterations):
            self.iterations += 1
            self._residual(x, b, r)
            self.residual = self._norm(r)
            track.append(self.residual)
            if self.residual <= error:
                break
            _lib.check(lib.mg_jacobi_sweep_csr(n, d["indptr"].data_ptr(), d["indices"].data_ptr(),
                                               d["values"].data_ptr(), dinv.data_ptr(), x.data_ptr(), b.data_ptr(),
                                               xo.data_ptr(), float(omega), st), "mg_jacobi_sweep_csr")
            x, xo = xo, x
        self.solution = x.cpu().numpy().reshape(self.dim, 1)
        self.residual_vector = r.cpu().numpy().reshape(self.dim, 1)
        self.track_res = np.array(track, dtype=float).reshape(-1, 1)
