"""oracle/vcycle.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

CPU restatement of the reference's multigrid driver and V-cycle, statement by statement:

  * outer loop            learn_multigrid/solvers/Multigrid.py:36-75   (Multigrid.solve)
  * V-cycle               learn_multigrid/solvers/Multigrid.py:77-124  (Multigrid.v_cycle)
  * geometric interpolator learn_multigrid/solvers/Multigrid.py:126-147
  * Solver base           learn_multigrid/solvers/Solver.py:14-21 (matrix stored as csc_matrix)
  * Jacobi update         learn_multigrid/solvers/Jacobi.py:22-35  (omega = 1 there)

Differences from the reference that the north star asks for, all opt-in:
  * `Q_list` gives a transfer operator for EVERY level (the reference only honours the first one and
    uses the dense 1D interpolator below, Multigrid.py:188-197);
  * `smoother` is honoured: "gs" = the committed PyAMG lexicographic Gauss-Seidel path,
    "jacobi" = damped Jacobi (commented-out dispatch at Multigrid.py:85-86,119-120 + Jacobi.py:35),
    "mcgs" = multicolour Gauss-Seidel in a given colour order;
  * `hoist_setup=True` builds the Galerkin operators and the coarse LU once instead of in every
    cycle (Multigrid.py:97-98,106 redo them); arithmetic per cycle is unchanged.
"""
import numpy as np
import scipy.sparse as sp
from scipy.sparse import csc_matrix, csr_matrix
from scipy.sparse.linalg import spsolve, splu

from . import kernels as K


def geometric_interpolator(dimension):
    """Dense 1D linear interpolation of Multigrid.interpolator (Multigrid.py:126-147)."""
    rows = dimension
    cols = int(np.floor((dimension - 1) / 2)) + 1
    mat = np.zeros(shape=(rows, cols))
    i = 1
    for j in range(1, cols - 1):
        mat[i, j] = 1
        i += 1
        mat[i, j] = 2
        i += 1
        mat[i, j] = 1
    mat[0, 0] = 2
    mat[1, 0] = 1
    mat[-1, -1] = 2
    mat[-2, -1] = 1
    return mat / 2


def galerkin(A, Q):
    """A_c = csr_matrix(Q.T @ A @ Q) exactly as Multigrid.py:97-98 (SciPy two-pass SpGEMM, exact zeros
    pruned by the numeric pass, canonical sorted CSR after the final conversion)."""
    Ac = Q.T @ A @ Q
    Ac = csr_matrix(Ac)
    Ac.sort_indices()
    return Ac


def color_rows_from_colors(colors):
    """rows of each colour, colours in increasing index, rows ascending inside a colour."""
    colors = np.asarray(colors)
    return [np.flatnonzero(colors == c).astype(np.int32) for c in range(int(colors.max()) + 1)]


class OracleMultigrid:
    """Restated `Multigrid` / `SemiGeometricMG` (Multigrid.py:26-197) with an explicit hierarchy."""

    def __init__(self, A, rhs, Q_list=None, smoother="gs", omega=1.0, colors=None,
                 hoist_setup=False, geometric_below=False, reverse_post=False):
        self.reverse_post = reverse_post         # multicolour post-smoothing in reverse colour order (engine option)
        # Solver.__init__ (Solver.py:14-21)
        self.dim = rhs.size
        self.matrix = csc_matrix(A)
        self.rhs = np.asarray(rhs, dtype=np.float64).reshape(-1, 1)
        self.solution = np.empty(shape=self.rhs.shape)
        self.residual = 0.0
        self.residual_vector = np.empty(shape=self.rhs.shape)
        self.track_res = np.ndarray(shape=(0, 1), dtype=float)
        self.iterations = 0                      # IterativeSolver.__init__ (Solver.py:72); never reset
        self.Q_list = [csr_matrix(q) for q in (Q_list or [])]   # SemiGeometricMG.__init__ :182
        self.smoother = smoother
        self.omega = float(omega)
        self.colors = colors                     # list (per level) of colour arrays, for "mcgs"
        self.hoist_setup = hoist_setup
        self.geometric_below = geometric_below   # reference behaviour below the first level
        self._A_levels = None
        self._lu = None
        self._color_rows = None

    # -- hierarchy pieces -------------------------------------------------------------------------
    def _interp(self, level, n):
        if level < len(self.Q_list):
            return self.Q_list[level]
        if self.geometric_below:
            return geometric_interpolator(n)     # dense ndarray, as in the reference
        raise ValueError("no transfer operator for level %d" % level)

    def _smooth(self, A, u, rhs, steps, level, post=False):
        """pre/post smoothing (Multigrid.py:88,121); returns the smoothed vector (may alias u)."""
        if self.smoother == "gs":
            K.gauss_seidel(A, u, rhs, iterations=steps)
            return u
        if self.smoother == "jacobi":
            A = csr_matrix(A)
            dinv = 1.0 / A.diagonal()
            return K.jacobi(A, u, rhs, dinv, omega=self.omega, iterations=steps)
        if self.smoother == "mcgs":
            if self._color_rows is None:
                self._color_rows = {}
            if level not in self._color_rows:
                self._color_rows[level] = color_rows_from_colors(self.colors[level])
            rows = self._color_rows[level]
            if post and self.reverse_post:
                rows = rows[::-1]
            K.gauss_seidel_multicolor(A, u, rhs, rows, iterations=steps)
            return u
        raise ValueError("unknown smoother %r" % (self.smoother,))

    def build_hierarchy(self, levels):
        """Galerkin operators for all levels (used when hoist_setup=True and by pattern tests)."""
        A = self.matrix
        out = [A]
        for l in range(levels - 1):
            i = self._interp(l, A.shape[0])
            A = galerkin(A, i) if sp.issparse(i) else csr_matrix(i.T @ A @ i)
            out.append(A)
        self._A_levels = out
        self._lu = splu(csc_matrix(out[-1]))
        return out

    # -- Multigrid.v_cycle (Multigrid.py:77-124) --------------------------------------------------
    def v_cycle(self, A, u0, rhs, smooth_steps, levels, level=0):
        levels -= 1
        u0 = self._smooth(A, u0, rhs, smooth_steps, level)            # :88
        u = u0.copy()                                                 # :89
        res = rhs - A.dot(u)                                          # :90
        i = self._interp(level, A.shape[0])                           # :91
        res_coarse = i.T @ res                                        # :93
        if self.hoist_setup:
            A_coarse = self._A_levels[level + 1]
        else:
            A_coarse = csr_matrix(i.T @ A @ i)                        # :97-98
        if levels != 1:                                               # :102-104
            u_coarse = self.v_cycle(A_coarse, np.zeros(shape=(A_coarse.shape[0], 1)), res_coarse,
                                    smooth_steps, levels, level + 1)
        else:                                                         # :106
            if self.hoist_setup:
                u_coarse = np.reshape(self._lu.solve(np.ravel(res_coarse)), (A_coarse.shape[0], 1))
            else:
                u_coarse = np.reshape(spsolve(A_coarse, res_coarse, use_umfpack=False),
                                      (A_coarse.shape[0], 1))
        u = u + i @ u_coarse                                          # :115
        u = self._smooth(A, u, rhs, smooth_steps, level, post=True)   # :121
        return u

    # -- Multigrid.solve (Multigrid.py:36-75) -----------------------------------------------------
    def solve(self, levels=2, smooth_steps=1, max_iterations=100, error=1e-08, initial_guess=None):
        if initial_guess is None:
            self.solution = np.zeros(shape=(self.dim, 1))
        else:
            self.solution = initial_guess
        if self.hoist_setup and (self._A_levels is None or len(self._A_levels) != levels):
            self.build_hierarchy(levels)
        A = self.matrix
        track_res = np.ndarray(shape=(0, 1), dtype=float)
        for _ in range(0, max_iterations):
            self.iterations += 1
            self.residual_vector = self.rhs - self.matrix.dot(self.solution)
            self.residual = np.linalg.norm(self.residual_vector)
            if self.iterations <= 1:                                  # :64-66 quirk
                self.residual_vector = np.ones(shape=self.solution.shape)
                self.residual = np.linalg.norm(self.residual_vector)
            track_res = np.vstack((track_res, self.residual))
            if self.residual <= error:
                break
            self.solution = self.v_cycle(A, self.solution, self.rhs, smooth_steps, levels)
        self.track_res = track_res
        return self.solution


# -- stationary solvers and CG (Jacobi.py:15-37, GaussSeidel.py:16-39, CG.py:12-50) -----------------
def jacobi_solve(A, rhs, max_iterations=1000, error=1e-12, initial_guess=None, omega=1.0):
    A = csr_matrix(csc_matrix(A))
    rhs = np.asarray(rhs, dtype=np.float64).reshape(-1, 1)
    x = np.zeros_like(rhs) if initial_guess is None else initial_guess
    dinv = (1.0 / A.diagonal()).reshape(-1, 1)
    track = []
    it = 0
    for _ in range(max_iterations):
        it += 1
        r = rhs - A.dot(x)
        res = np.linalg.norm(r)
        track.append(res)
        if res <= error:
            break
        x = x + omega * (dinv * r)
    return x, np.array(track).reshape(-1, 1), it


def gauss_seidel_solve(A, rhs, max_iterations=1000, error=1e-12, initial_guess=None):
    """x += (D+L)^-1 r per iteration (GaussSeidel.py:23-37) done as a forward substitution = one
    lexicographic GS sweep."""
    A = csr_matrix(csc_matrix(A))
    rhs = np.asarray(rhs, dtype=np.float64).reshape(-1, 1)
    x = np.zeros_like(rhs) if initial_guess is None else initial_guess.copy()
    track = []
    it = 0
    for _ in range(max_iterations):
        it += 1
        r = rhs - A.dot(x)
        res = np.linalg.norm(r)
        track.append(res)
        if res <= error:
            break
        K.gauss_seidel(A, x, rhs, iterations=1)
    return x, np.array(track).reshape(-1, 1), it


def cg_solve(A, rhs, max_iterations=1000, error=1e-08, initial_guess=None):
    """Plain CG of CG.py:12-50 (history includes the initial residual)."""
    A = csc_matrix(A)
    rhs = np.asarray(rhs, dtype=np.float64).reshape(-1, 1)
    x = np.zeros_like(rhs) if initial_guess is None else initial_guess
    r = rhs - A.dot(x)
    track = [np.linalg.norm(r)]
    p = r
    it = 0
    for _ in range(max_iterations):
        it += 1
        r2 = (r.T @ r).item()
        Ap = A.dot(p)
        pAp = (p.T @ Ap).item()
        alpha = r2 / pAp
        x = x + alpha * p
        r = r - alpha * A @ p
        res = np.linalg.norm(r)
        track.append(res)
        if res <= error:
            break
        beta = (r.T @ r).item() / r2
        p = r + beta * p
    return x, np.array(track).reshape(-1, 1), it


def pcg_solve(A, rhs, precond, max_iterations=1000, error=1e-08):
    """Preconditioned CG in the statement order of learnmultigrid_b200/solvers/CG.py (the reference's CG.py:12-50 has
    no preconditioner; BASELINE.json configs[4] asks for MG-preconditioned CG).  precond(r) -> z."""
    A = csc_matrix(A)
    rhs = np.asarray(rhs, dtype=np.float64).reshape(-1, 1)
    x = np.zeros_like(rhs)
    r = rhs - A.dot(x)
    track = [np.linalg.norm(r)]
    z = precond(r)
    p = z.copy()
    rz = (r.T @ z).item()
    it = 0
    for _ in range(max_iterations):
        it += 1
        Ap = A.dot(p)
        alpha = rz / (p.T @ Ap).item()
        x = x + alpha * p
        r = r - alpha * Ap
        res = np.linalg.norm(r)
        track.append(res)
        if res <= error:
            break
        z = precond(r)
        rz_new = (r.T @ z).item()
        beta = rz_new / rz
        rz = rz_new
        p = z + beta * p
    return x, np.array(track).reshape(-1, 1), it
