"""2D triangular meshes with the reference's interface (learn_multigrid/mesh/Mesh2D.py), vectorised.

Node and element numbering are the reference's: nodes row-major on the unit square (Mesh2D.py:63-74), each
square [k, k+1, k+W+1, k+W] split into triangles [0,1,2] and [0,2,3] (:76-91), i.e. the diagonal runs from the
lower-left to the upper-right corner.  `refine` is the 1 -> 4 red refinement of :95-160 with the same child
order and the same consumption of np.random.rand() (one draw per new edge point).
"""
import math

import numpy as np


def get_prime_factors(number):
    prime_factors = []
    while number % 2 == 0:
        prime_factors.append(2)
        number = number / 2
    for i in range(3, int(math.sqrt(number)) + 1, 2):
        while number % i == 0:
            prime_factors.append(int(i))
            number = number / i
    if number > 2:
        prime_factors.append(int(number))
    return prime_factors


class Mesh2D:

    def __init__(self, ne=0, p=np.array([]), conn=np.array([])):
        if conn.size != 0 and p.size != 0:
            pass
        else:
            self.ne = ne
            [p, conn] = self.construct()
        self.ne = len(conn)
        self.p = p
        self.conn = conn
        self.n_p = len(p)

    @staticmethod
    def find_balanced_couple(ne):
        rad = np.sqrt(ne)
        y = np.mod(rad, 1)
        A = 0
        B = 0
        if y == 0:
            A = B = int(rad)
        else:
            set_A = []
            set_B = []
            queue = get_prime_factors(ne)
            while queue:
                if np.prod(set_B) < np.prod(set_A):
                    set_B.append(queue.pop())
                else:
                    set_A.append(queue.pop())
                A = np.prod(set_A)
                B = np.prod(set_B)
        return int(A), int(B)

    def construct(self):
        h_el, v_el = self.find_balanced_couple(self.ne)
        h_p = np.linspace(0, 1, num=h_el + 1, endpoint=True)
        v_p = np.linspace(0, 1, num=v_el + 1, endpoint=True)
        W = h_el + 1
        p = np.empty((W * (v_el + 1), 2))
        p[:, 0] = np.tile(h_p, v_el + 1)
        p[:, 1] = np.repeat(v_p, W)
        k = (np.arange(v_el)[:, None] * W + np.arange(h_el)[None, :]).reshape(-1)
        conn = np.empty((2 * len(k), 3), dtype=int)
        conn[0::2, 0] = k
        conn[0::2, 1] = k + 1
        conn[0::2, 2] = k + W + 1
        conn[1::2, 0] = k
        conn[1::2, 1] = k + W + 1
        conn[1::2, 2] = k + W
        return p, conn

    def refine(self, regular=True):
        p = [tuple(r) for r in self.p]
        conn = self.conn
        ne = len(conn)
        mid = {}
        new_conn = np.zeros((ne * 4, 3), dtype=int)
        if regular:
            a = 0.5
            b = 0.5
        else:
            a = 0.3
            b = 1 - a
        nm = 0
        for j in range(ne):
            el = conn[j, :]
            t_new = [0, 0, 0]
            for k in range(3):
                left = int(el[k])
                right = int(el[0] if k == 2 else el[k + 1])
                key = (left, right) if left < right else (right, left)
                if key not in mid:
                    r = a + (b - a) * np.random.rand()
                    xn = r * p[left][0] + (1 - r) * p[right][0]
                    yn = r * p[left][1] + (1 - r) * p[right][1]
                    mid[key] = len(p)
                    p.append((xn, yn))
                t_new[k] = mid[key]
            new_conn[nm, :] = [t_new[0], el[1], t_new[1]]
            new_conn[nm + 1, :] = [t_new[1], el[2], t_new[2]]
            new_conn[nm + 2, :] = [t_new[0], t_new[1], t_new[2]]
            new_conn[nm + 3, :] = [el[0], t_new[0], t_new[2]]
            nm += 4
        self.p = np.array(p)
        self.conn = new_conn
        self.n_p = len(p)
        self.ne = len(new_conn)

    def embedding(self):
        raise NotImplementedError("ghost-node embedding (Mesh2D.py:162-431) serves the NN patch extraction at "
                                  "boundaries and is outside the V-cycle hot path (SURVEY.md 8f rank 4)")

    def get_connections(self):
        return self.conn

    def get_points(self):
        return self.p

    def get_ne(self):
        return self.ne

    def get_np(self):
        return self.n_p

    def get_mesh(self):
        return self.p

    def plot_mesh(self):
        import matplotlib.pyplot as plt
        plt.triplot(self.p[:, 0], self.p[:, 1], self.conn)
        plt.show()
