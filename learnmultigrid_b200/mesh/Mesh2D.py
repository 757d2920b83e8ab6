"""2D triangular meshes with the reference's interface (learn_multigrid/mesh/Mesh2D.py), vectorised.

Node and element numbering are the reference's: nodes row-major on the unit square (Mesh2D.py:63-74), each
square [k, k+1, k+W+1, k+W] split into triangles [0,1,2] and [0,2,3] (:76-91), i.e. the diagonal runs from the
lower-left to the upper-right corner.  `refine` is the 1 -> 4 red refinement of :95-160 with the same child
order and the same consumption of np.random.rand() (one draw per new edge point).
"""
import math

import numpy as np

# the reference's Mesh2D module also carries a copy of the 1D refinement mesh (Mesh2D.py:524-546; its parent there
# has no connection_matrix, so only the Mesh1D one can be constructed): same name, the working class
from .Mesh1D import Mesh1DRefinement  # noqa: F401


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
        """1 -> 4 red refinement (Mesh2D.py:95-160), vectorised.  New edge points are numbered in the order the
        reference discovers their edges (element by element, local edges 0-1, 1-2, 2-0), one np.random.rand() draw
        per new point in that order, placed at r * left + (1 - r) * right with left/right as first met."""
        p = np.asarray(self.p, dtype=float)
        conn = np.asarray(self.conn, dtype=int)
        ne, n_p = len(conn), len(p)
        a, b = (0.5, 0.5) if regular else (0.3, 0.7)
        left = conn.reshape(-1)                                   # (element, local edge) row-major = discovery order
        right = conn[:, [1, 2, 0]].reshape(-1)
        key = np.minimum(left, right) * n_p + np.maximum(left, right)
        uniq, first, inverse = np.unique(key, return_index=True, return_inverse=True)
        order = np.argsort(first, kind="stable")                  # unique edges in discovery order
        rank = np.empty_like(order)
        rank[order] = np.arange(len(order))
        first = first[order]
        r = a + (b - a) * np.random.rand(len(first))
        new_p = np.empty((len(first), 2))
        new_p[:, 0] = r * p[left[first], 0] + (1 - r) * p[right[first], 0]
        new_p[:, 1] = r * p[left[first], 1] + (1 - r) * p[right[first], 1]
        t_new = (n_p + rank[inverse.reshape(-1)]).reshape(ne, 3)
        new_conn = np.empty((ne * 4, 3), dtype=int)
        new_conn[0::4] = np.stack([t_new[:, 0], conn[:, 1], t_new[:, 1]], axis=1)
        new_conn[1::4] = np.stack([t_new[:, 1], conn[:, 2], t_new[:, 2]], axis=1)
        new_conn[2::4] = t_new
        new_conn[3::4] = np.stack([conn[:, 0], t_new[:, 0], t_new[:, 2]], axis=1)
        self.p = np.vstack((p, new_p))
        self.conn = new_conn
        self.n_p = len(self.p)
        self.ne = len(new_conn)

    def embedding(self):
        """Two layers of ghost elements around the unit square (Mesh2D.py:162-431), as a new Mesh2D.

        Restated from the reference's behaviour with the same node and element numbering (the original nodes
        keep their ids; per layer: left, right, bottom, top strips in border-edge order, then the four corner
        squares lb, lt, rt, rb), the same exact-equality border tests and the same shift rule
        (`find_h_v_shift`: max over the two opposite sides of the smallest *signed* edge extent), without the
        per-element Python loop and the repeated `np.vstack` of the original.
        """
        pts = [tuple(r) for r in np.asarray(self.p, dtype=float)]
        tris = [np.asarray(self.conn, dtype=int)]
        horizontal = 0
        vertical = 0

        def add(x, y):
            pts.append((x, y))
            return len(pts) - 1

        for _ in range(2):
            p = np.array(pts)
            conn = np.vstack(tris)
            borders = (p[:, 0] == (0 - horizontal), p[:, 0] == (1 + horizontal),
                       p[:, 1] == (0 - vertical), p[:, 1] == (1 + vertical))
            left_edge, right_edge, bottom_edge, top_edge = (
                self.order_edges(self._border_edges(conn, on), p, axis)
                for on, axis in zip(borders, (1, 1, 0, 0)))
            vertical = self.find_h_v_shift(p, [left_edge, right_edge], 1)
            horizontal = self.find_h_v_shift(p, [top_edge, bottom_edge], 0)
            new_tris = []

            def strip(edges, axis, ghost_xy, square_of):
                """one ghost square per border edge; `ghost_xy(node)` places the ghost of a border node,
                `square_of(lo, hi, new_lo, new_hi)` orders the square's corners as the reference does"""
                ghost = {}
                angles = np.zeros((2, 2), dtype=int)
                for i, (a, b) in enumerate(edges):
                    lo, hi = (a, b) if pts[a][axis] < pts[b][axis] else (b, a)
                    for n in (lo, hi):
                        if n not in ghost:
                            ghost[n] = add(*ghost_xy(n))
                    sq, first, last = square_of(lo, hi, ghost[lo], ghost[hi])
                    if i == 0:
                        angles[0, :] = first
                    if i == len(edges) - 1:
                        angles[1, :] = last
                    new_tris.append([sq[0], sq[1], sq[2]])
                    new_tris.append([sq[0], sq[2], sq[3]])
                return angles

            left_angles = strip(left_edge, 1, lambda n: (pts[n][0] - horizontal, pts[n][1]),
                                lambda d, t, nd, nt: ([nd, d, t, nt], [d, nd], [nt, t]))
            right_angles = strip(right_edge, 1, lambda n: (pts[n][0] + horizontal, pts[n][1]),
                                 lambda d, t, nd, nt: ([d, nd, nt, t], [nd, d], [t, nt]))
            bottom_angles = strip(bottom_edge, 0, lambda n: (pts[n][0], pts[n][1] - vertical),
                                  lambda l, r, nl, nr: ([nl, nr, r, l], [nl, l], [r, nr]))
            top_angles = strip(top_edge, 0, lambda n: (pts[n][0], pts[n][1] + vertical),
                               lambda l, r, nl, nr: ([l, r, nr, nl], [l, nl], [nr, r]))

            def extent(pair, axis):
                return abs(pts[pair[1]][axis] - pts[pair[0]][axis])

            # corner squares; the abscissae of both left corners start from left_angles[0, 0] (Mesh2D.py:387)
            n = add(pts[left_angles[0, 0]][0] - extent(left_angles[0], 0),
                    pts[bottom_angles[0, 1]][1] - extent(bottom_angles[0], 1))
            corner = [[n, bottom_angles[0, 0], left_angles[0, 0], left_angles[0, 1]]]
            n = add(pts[left_angles[0, 0]][0] - extent(left_angles[1], 0),
                    pts[top_angles[0, 0]][1] + extent(top_angles[0], 1))
            corner.append([left_angles[1, 0], left_angles[1, 1], top_angles[0, 1], n])
            n = add(pts[right_angles[1, 0]][0] + extent(right_angles[1], 0),
                    pts[top_angles[1, 1]][1] + extent(top_angles[1], 1))
            corner.append([top_angles[1, 1], right_angles[1, 1], n, top_angles[1, 0]])
            n = add(pts[right_angles[0, 1]][0] + extent(right_angles[0], 0),
                    pts[bottom_angles[1, 0]][1] - extent(bottom_angles[1], 1))
            corner.append([bottom_angles[1, 1], n, right_angles[0, 0], right_angles[0, 1]])
            for sq in corner:
                new_tris.append([sq[0], sq[1], sq[2]])
                new_tris.append([sq[0], sq[2], sq[3]])
            tris.append(np.array(new_tris, dtype=int))

        return Mesh2D(p=np.array(pts), conn=np.vstack(tris))

    @staticmethod
    def _border_edges(conn, on_border):
        """edges (pairs of element nodes, in the element's local order) with both ends on the border, in element
        order: the vectorised form of the `find_edges` scan over all elements (Mesh2D.py:177-181, 474-484)"""
        hit = on_border[conn]
        count = hit.sum(axis=1)
        if np.any(count > 2):
            raise ValueError("an element has all three nodes on one border line")
        rows = np.nonzero(count == 2)[0]
        if rows.size == 0:
            return np.ones((1, 2), dtype=int) * -1
        return conn[rows][hit[rows]].reshape(-1, 2)

    @staticmethod
    def find_h_v_shift(p, edges, index):
        lens = []
        for side in edges:
            ext = p[side[:, 1], index] - p[side[:, 0], index]
            lens.append(min(100, ext.min()))
        return max(lens)

    @staticmethod
    def order_edges(edge, p, axis):
        """first row := an edge touching the smallest coordinate, last row := one touching the largest, by
        swapping (the rows in between keep their discovery order), Mesh2D.py:449-471"""
        for row, pick in ((0, np.argmin), (-1, np.argmax)):
            c0 = p[edge[:, 0], axis]
            c1 = p[edge[:, 1], axis]
            pos_1 = pick(c0)
            pos_2 = pick(c1)
            if pick is np.argmin:
                which = pos_1 if c0[pos_1] < c1[pos_2] else pos_2
            else:
                which = pos_1 if c0[pos_1] > c1[pos_2] else pos_2
            edge[[row, which]] = edge[[which, row]]
        return edge

    @staticmethod
    def find_edges(element, border, edge):
        where = np.isin(np.asarray(element), border)
        if where.sum() > 1:
            if edge[0, 0] == -1:
                edge[0, :] = np.asarray(element)[where]
            else:
                edge = np.vstack((edge, np.asarray(element)[where]))
        return edge

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
