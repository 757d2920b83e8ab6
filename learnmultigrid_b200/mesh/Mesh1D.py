"""1D meshes with the reference's interface (learn_multigrid/mesh/Mesh1D.py:9-96), vectorised."""
import numpy as np


class Mesh1D:

    def __init__(self, regular=True, ne=0):
        self.regular = regular
        self.ne = ne
        self.np = ne + 1
        self.h = 1 / ne
        self.x = np.array([])
        self.conn = np.ndarray(shape=(self.ne, 2))

    def construct(self):
        if self.is_regular():
            self.construct_regular()
        else:
            self.construct_irregular()
        self.connection_matrix()

    def construct_regular(self):
        self.x = np.linspace(0, 1, self.np)

    def construct_irregular(self):
        """interior nodes shifted left by r in [h/8, h/4) (Mesh1D.py:30-42); draws np.random.rand() once per
        interior node in index order, so a seeded run reproduces the reference's mesh exactly."""
        h = self.h
        tmp = np.linspace(0, 1, self.np)
        b = h / 4
        a = h / 8
        for i in range(1, self.np - 1):
            r = (b - a) * np.random.rand() + a
            tmp[i] = tmp[i] - r
        self.x = tmp

    def is_regular(self):
        return self.regular

    def connection_matrix(self):
        x = self.x
        self.conn = np.stack((x[:-1], x[1:]), axis=1).astype(float)

    def get_connections(self):
        return self.conn

    def get_ne(self):
        return self.ne

    def get_np(self):
        return self.np

    def get_mesh(self):
        return self.x

    def plot_mesh(self):
        import matplotlib.pyplot as plt
        x = self.x
        plt.plot(x, np.zeros(len(x)), 'ro')
        plt.grid()
        plt.title('Mesh')
        plt.show()


class Mesh1DRefinement(Mesh1D):
    """uniform mesh with coarse_ne * 2**n_ref elements (Mesh1D.py:77-96)"""

    def __init__(self, coarse_ne=2, n_ref=0):
        self.ne = coarse_ne
        self.np = coarse_ne + 1
        self.h = 1 / coarse_ne
        self.x = np.linspace(0, 1, self.np)
        self.conn = np.ndarray(shape=(self.ne, 2))
        self.n_ref = n_ref
        self.regular = True

    def construct(self):
        ne = self.ne * (2 ** self.n_ref)
        self.ne = ne
        self.np = ne + 1
        self.h = 1 / ne
        self.x = np.linspace(0, 1, self.np)
        super().connection_matrix()
