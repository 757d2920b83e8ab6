"""P1 shape functions and gradients with the reference's interface
(learn_multigrid/assembly/ShapeFunction.py:6-86)."""
from abc import ABC

import numpy as np


class ShapeFunction(ABC):

    def __init__(self, order):
        self.order = order
        self.phi = None

    def evaluate(self, points=None, index=None):
        if index is None:
            result = np.ndarray(shape=(0, np.size(points)), dtype=float)
            for f in self.phi:
                result = np.vstack((result, f(points)))
            return result
        if index in range(0, self.phi.size):
            return self.phi[index](points)
        print("No shape function available")
        return -100

    def get_functions(self):
        return self.phi


class Function(ShapeFunction):
    def __init__(self, order):
        super().__init__(order)
        self.phi = self.order_to_function(order)

    @staticmethod
    def order_to_function(order):
        return {2: np.array([lambda x: 1 - x, lambda x: x])}.get(order, "Invalid order")


class Gradient(ShapeFunction):
    def __init__(self, order):
        super().__init__(order)
        self.phi = self.order_to_gradient(order)

    @staticmethod
    def order_to_gradient(order):
        return {2: np.array([lambda x: -1, lambda x: 1])}.get(order, "Invalid order")


class FunctionTriangle(ShapeFunction):
    def __init__(self, order):
        super().__init__(order)
        self.phi = self.order_to_function(order)

    @staticmethod
    def order_to_function(order):
        return {1: np.array([lambda p: 1 - p[0] - p[1], lambda p: p[0], lambda p: p[1]])}.get(order, "Invalid order")


class GradientTriangle(ShapeFunction):
    def __init__(self, order):
        super().__init__(order)
        self.phi = self.order_to_gradient(order)

    @staticmethod
    def order_to_gradient(order):
        return {1: np.array([lambda p: np.array([np.array([-1, -1])]).T,
                             lambda p: np.array([np.array([1, 0])]).T,
                             lambda p: np.array([np.array([0, 1])]).T])}.get(order, "Invalid order")
