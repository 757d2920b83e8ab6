"""Element intersections with the reference's interface (learn_multigrid/L2_projection/Intersection.py).

1D: all overlapping (fine, coarse) element pairs in the reference's discovery order (fine element major,
coarse element minor, :59-77) found by a sorted sweep instead of the O(ne_f * ne_c) double loop; segment end
points are read sequentially from the union of all node coordinates exactly as the reference does (:57,75).
2D: the nested child map of the unfinished stub (:19-34): coarse element i <-> fine elements 4i .. 4i+3.
"""
import numpy as np


class Intersection:

    def __init__(self, fine_mesh, coarse_mesh):
        self.fine_mesh = fine_mesh
        self.coarse_mesh = coarse_mesh
        self.intersections = None
        self.int_coord = None
        self.union = None

    def get_info(self):
        return self.intersections, self.int_coord, self.union

    def get_intersections(self):
        return self.intersections

    def find_intersections2d(self, geometric=False):
        """geometric=False: the reference's stub (nested child map, Intersection.py:19-34).  geometric=True: all
        (fine, coarse) element pairs that overlap with positive area, for arbitrary mesh pairs (coupling2d.py);
        `int_coord` then holds the overlap areas."""
        if geometric:
            from .coupling2d import coupling_operator_2d
            _, pairs, area = coupling_operator_2d(self.fine_mesh, self.coarse_mesh, return_pairs=True)
            order = np.lexsort((pairs[:, 1], pairs[:, 0]))
            self.intersections = pairs[order].astype(int)
            self.int_coord = area[order]
            return
        conn = self.fine_mesh.get_connections()
        c_conn = self.coarse_mesh.get_connections()
        intersected = np.zeros((conn.shape[0], 2), dtype=int)
        fine = np.arange(4 * c_conn.shape[0])
        intersected[fine, 0] = fine
        intersected[fine, 1] = fine // 4
        self.intersections = intersected

    def find_intersections1d(self):
        conn = self.fine_mesh.get_connections()
        coarse_conn = self.coarse_mesh.get_connections()
        union = np.union1d(conn, coarse_conn)
        left, right = conn[:, 0], conn[:, 1]
        c_left, c_right = coarse_conn[:, 0], coarse_conn[:, 1]
        # pair (i, j) intersects unless left_i >= c_right_j or right_i <= c_left_j; both meshes are sorted, so
        # for fine element i the intersecting coarse elements are the contiguous range [lo_i, hi_i)
        lo = np.searchsorted(c_right, left, side="right")
        hi = np.searchsorted(c_left, right, side="left")
        counts = np.maximum(hi - lo, 0)
        fi = np.repeat(np.arange(len(conn)), counts)
        start = np.repeat(lo, counts)
        offs = np.arange(counts.sum()) - np.repeat(np.cumsum(counts) - counts, counts)
        intersections = np.stack((fi, start + offs), axis=1).astype(int).reshape(-1, 2)
        k = np.arange(len(intersections))
        int_coord = np.stack((union[k], union[k + 1]), axis=1).astype(float).reshape(-1, 2)
        self.intersections = intersections
        self.int_coord = int_coord
        self.union = union
