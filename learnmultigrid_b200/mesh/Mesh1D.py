"""1D meshes of the unit interval behind the reference's interface (learn_multigrid/mesh/Mesh1D.py:9-96).

    Mesh1D(regular, ne).construct()            nodes in `x`, elements as (left, right) coordinate pairs in `conn`
    Mesh1DRefinement(coarse_ne, n_ref)         the uniform mesh after n_ref bisections of coarse_ne elements

An irregular mesh moves every interior node to the LEFT by a random amount in [h/8, h/4) (Mesh1D.py:30-42).  The
reference draws `np.random.rand()` once per interior node in index order; one vectorised draw of ne - 1 numbers consumes
the legacy global generator in exactly the same way, so a script that seeds NumPy first gets the reference's mesh bit
for bit (tests/test_transfer_1d_golden.py).
"""
import numpy as np

from .Element1D import Element


def _uniform_nodes(ne):
    return np.linspace(0, 1, ne + 1)


def _element_table(x):
    """(ne, 2) array of element end points"""
    x = np.asarray(x, dtype=float)
    return np.column_stack((x[:-1], x[1:]))


class Mesh1D:

    def __init__(self, regular=True, ne=0):
        self.regular = regular
        self._resize(ne)
        self.x = np.array([])

    def _resize(self, ne):
        self.ne, self.np, self.h = ne, ne + 1, 1 / ne
        self.conn = np.ndarray(shape=(ne, 2))

    # ---- construction --------------------------------------------------------------------------------
    def construct(self):
        (self.construct_regular if self.is_regular() else self.construct_irregular)()
        self.connection_matrix()

    def construct_regular(self):
        self.x = _uniform_nodes(self.ne)

    def construct_irregular(self):
        lo, hi = self.h / 8, self.h / 4
        nodes = _uniform_nodes(self.ne)
        shift = (hi - lo) * np.random.rand(max(self.np - 2, 0)) + lo
        nodes[1:self.np - 1] = nodes[1:self.np - 1] - shift
        self.x = nodes

    def connection_matrix(self):
        self.conn = _element_table(self.x)

    # ---- queries -------------------------------------------------------------------------------------
    def is_regular(self):
        return self.regular

    def get_connections(self):
        return self.conn

    def get_ne(self):
        return self.ne

    def get_np(self):
        return self.np

    def get_mesh(self):
        return self.x

    def elements(self):
        """the elements as mesh.Element1D.Element objects (not in the reference)"""
        return [Element(k, a, b) for k, (a, b) in enumerate(self.conn)]

    def plot_mesh(self):
        import matplotlib.pyplot as plt
        plt.plot(self.x, np.zeros(len(self.x)), "ro")
        plt.grid()
        plt.title("Mesh")
        plt.xticks(np.arange(0, 1.1, step=0.1))
        plt.show()


class Mesh1DRefinement(Mesh1D):

    def __init__(self, coarse_ne=2, n_ref=0):
        self.regular = True
        self.n_ref = n_ref
        self._resize(coarse_ne)
        self.x = _uniform_nodes(coarse_ne)

    def construct(self):
        self._resize(self.ne << self.n_ref)
        self.construct_regular()
        self.connection_matrix()
