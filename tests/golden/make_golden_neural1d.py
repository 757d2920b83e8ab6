"""Generate tests/golden/neural_1d.npz with the REFERENCE's 1D NeuralMG (learn_multigrid/solvers/Multigrid.py:200-370)
run through oracle/refshim.py with a deterministic stub in place of the Keras model (weights are not shipped):
features, transfer operators of every level, and the residual history / solution of NeuralMG.solve."""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")
from oracle import refshim  # noqa: E402

refshim.install()
with refshim.quiet():
    from learn_multigrid.mesh.Mesh1D import Mesh1D
    from learn_multigrid.assembly.StiffnessMatrix import StiffnessMatrix
    from learn_multigrid.assembly.MassMatrix import MassMatrix
    from learn_multigrid.assembly.LoadVector import LoadVector
    from learn_multigrid.assembly.Quadrature import Quadrature
    from learn_multigrid.assembly.ShapeFunction import Function, Gradient
    from learn_multigrid.solvers.Multigrid import NeuralMG


class Stub:
    """9 outputs per patch; entries 2, 4:7, 8 are the ones construct_B uses.  Computed from the un-normalised features
    so that B = M_h P exactly (the coupling operator on nested meshes), slightly perturbed so that the
    predicted operator is not the trivial one."""

    def __init__(self, mean, std, eps=0.05):
        self.mean, self.std, self.eps = mean, std, eps

    def predict(self, X):
        X = np.asarray(X, dtype=np.float64) * self.std + self.mean
        out = np.zeros((X.shape[0], 9))
        out[:, 2] = 0.5 * X[:, 2]
        out[:, 4] = X[:, 2] + 0.5 * X[:, 1]
        out[:, 5] = X[:, 3] + 0.5 * (X[:, 2] + X[:, 4])
        out[:, 6] = X[:, 4] + 0.5 * X[:, 5]
        out[:, 8] = 0.5 * X[:, 4]
        out *= 1.0 + self.eps * np.sin(np.arange(X.shape[0]))[:, None]
        out[:, [0, 1, 3, 7]] = -7.0            # never used
        return out


def main():
    out = {}
    for name, ne, regular, seed, levels, steps in (("reg64", 64, True, None, 3, 3), ("irr128", 128, False, 5, 4, 2)):
        with refshim.quiet():
            if seed is not None:
                np.random.seed(seed)
            m = Mesh1D(regular, ne)
            m.construct()
            A = StiffnessMatrix(m).compute_stiffness_1d(Gradient(2), Quadrature(3))
            M = MassMatrix(m).compute_mass_1d(Function(2), Quadrature(3))
            rhs = LoadVector(m).compute_rhs_1d(lambda x: np.ones(shape=np.shape(x)))
            A[1, 0] = 0
            A[-2, -1] = 0
            A[0, :] = 0
            A[-1, :] = 0
            A[0, 0] = A[-1, -1] = 1                                  # test/test_NN.py:176-181
            rhs[0] = rhs[-1] = 0
            mean, std = np.zeros(7), np.ones(7) * M.max()
            nmg = NeuralMG(A, rhs, Stub(mean, std), M, std, mean)
            Ml = M
            for l in range(levels - 1):
                out["%s_feat%d" % (name, l)] = nmg.prepare_nn_input(Ml)
                Q = nmg.transfer_op(Ml)
                out["%s_Q%d" % (name, l)] = np.asarray(Q)
                Ml = np.asarray(Q.T @ Ml @ Q)
            nmg.solve(levels=levels, smoother="GaussSeidel", smooth_steps=steps, error=1e-10, max_iterations=40,
                      initial_guess=np.zeros((ne + 1, 1)))
        out[name + "_A"], out[name + "_M"], out[name + "_rhs"] = np.asarray(A), np.asarray(M), np.asarray(rhs)
        out[name + "_std"] = std
        out[name + "_hist"] = nmg.track_res
        out[name + "_sol"] = nmg.solution
        out[name + "_par"] = np.array([levels, steps])
        print(name, nmg.track_res.ravel())
    path = os.path.join(HERE, "neural_1d.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
