"""Gauss-Seidel with the reference's interface (learn_multigrid/solvers/GaussSeidel.py:9-39) on the GPU.

The reference forms (D+L)^-1 explicitly and iterates x += (D+L)^-1 r, which is one index-order Gauss-Seidel
sweep per iteration; here the sweep is done exactly (level-scheduled kernel, mg_gs_lex_sweep_csr)."""
import numpy as np

from .. import _lib
from .. import formats as F
from .Solver import IterativeSolver


class GaussSeidel(IterativeSolver):

    def __init__(self, matrix, rhs):
        super().__init__(matrix, rhs)
        self.label = "Gauss-Seidel"

    def solve(self, max_iterations=1000, error=1e-12, initial_guess=None):
        d = self._device_csr()
        torch, lib, n = d["torch"], d["lib"], d["n"]
        if initial_guess is None:
            x0 = np.zeros(shape=(self.get_dimension(), 1))
        else:
            x0 = initial_guess
        x = self._upload(x0)
        b = self._upload(self.rhs)
        r = torch.empty_like(x)
        lp, lr = F.lex_levels(d["host"])
        lp_d = torch.from_numpy(lp).to(d["dev"])
        lr_d = torch.from_numpy(lr).to(d["dev"])
        st = _lib.stream_handle(torch)
        track = []
        for _ in range(0, max_iterations):
            self.iterations += 1
            self._residual(x, b, r)
            self.residual = self._norm(r)
            track.append(self.residual)
            if self.residual <= error:
                break
            _lib.check(lib.mg_gs_lex_sweep_csr(n, d["indptr"].data_ptr(), d["indices"].data_ptr(),
                                               d["values"].data_ptr(), x.data_ptr(), b.data_ptr(), lp_d.data_ptr(),
                                               lr_d.data_ptr(), len(lp) - 1, 1, st), "mg_gs_lex_sweep_csr")
        self.solution = x.cpu().numpy().reshape(self.dim, 1)
        self.residual_vector = r.cpu().numpy().reshape(self.dim, 1)
        self.track_res = np.array(track, dtype=float).reshape(-1, 1)
