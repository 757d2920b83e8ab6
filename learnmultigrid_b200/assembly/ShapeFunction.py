"""P1 shape functions and their gradients behind the reference's classes
(learn_multigrid/assembly/ShapeFunction.py:6-86): `Function(2)` / `Gradient(2)` on the interval [0, 1] (the reference
calls the two-node element "order 2"), `FunctionTriangle(1)` / `GradientTriangle(1)` on the unit triangle.

`evaluate(points)` stacks the values of all basis functions row by row; `evaluate(points, k)` is basis function k, and
an index outside the basis prints a message and returns -100 as the reference does.  The basis functions are plain
module-level functions collected in one table per class, keyed by the order the constructor takes.
"""
from abc import ABC

import numpy as np

_INVALID = "Invalid order"


def _hat_left(x):
    return 1 - x


def _hat_right(x):
    return x


def _slope_down(_x):
    return -1


def _slope_up(_x):
    return 1


def _tri_origin(p):
    return 1 - p[0] - p[1]


def _tri_x(p):
    return p[0]


def _tri_y(p):
    return p[1]


def _constant_column(gx, gy):
    column = np.array([np.array([gx, gy])]).T          # (2, 1) integer column, the shape the stiffness code multiplies

    def grad(_p):
        return column.copy()
    return grad


_BASES = {
    "Function": {2: (_hat_left, _hat_right)},
    "Gradient": {2: (_slope_down, _slope_up)},
    "FunctionTriangle": {1: (_tri_origin, _tri_x, _tri_y)},
    "GradientTriangle": {1: (_constant_column(-1, -1), _constant_column(1, 0), _constant_column(0, 1))},
}


def _basis(kind, order):
    fns = _BASES[kind].get(order)
    return _INVALID if fns is None else np.array(fns)


class ShapeFunction(ABC):
    _kind = None

    def __init__(self, order):
        self.order = order
        self.phi = None if self._kind is None else _basis(self._kind, order)

    def evaluate(self, points=None, index=None):
        if index is None:
            rows = [np.atleast_1d(f(points)) for f in self.phi]
            return np.vstack([np.empty((0, np.size(points)), dtype=float)] + rows)
        if index in range(0, self.phi.size):
            return self.phi[index](points)
        print("No shape function available")
        return -100

    def get_functions(self):
        return self.phi


class Function(ShapeFunction):
    _kind = "Function"

    @staticmethod
    def order_to_function(order):
        return _basis("Function", order)


class Gradient(ShapeFunction):
    _kind = "Gradient"

    @staticmethod
    def order_to_gradient(order):
        return _basis("Gradient", order)


class FunctionTriangle(ShapeFunction):
    _kind = "FunctionTriangle"

    @staticmethod
    def order_to_function(order):
        return _basis("FunctionTriangle", order)


class GradientTriangle(ShapeFunction):
    _kind = "GradientTriangle"

    @staticmethod
    def order_to_gradient(order):
        return _basis("GradientTriangle", order)
