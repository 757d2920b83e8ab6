"""Quadrature rules behind the reference's interface (learn_multigrid/assembly/Quadrature.py:5-98).

What has to be reproduced for bit parity with the reference's assembled matrices:
  * the 3-point Gauss rule on [0, 1] with nodes and weights ROUNDED to 14 digits (:54, :62), and the 3-point triangle
    rule with nodes rounded likewise and weights 1/6 (:87-96);
  * every integral is accumulated from the integer 0, one quadrature point after the other, each term evaluated as
    (f(x_k) * g(x_k)) * w_k;
  * `Quadrature2D.compute_grad` weighs every point with w[i], i the ROW index of the local matrix, not with w[k] (:80)
    -- harmless only because the triangle weights are equal; kept.
The rules live in one table per reference element; `order_to_points` / `order_to_weights` look them up and return
"Invalid order" for anything else, as the reference's switcher does.
"""
import numpy as np

_INVALID = "Invalid order"

# order -> (nodes, weights); integers stay integers (order 1 is the reference's placeholder rule)
_INTERVAL_RULES = {
    1: ((0,), (2,)),
    3: ((0.11270166537926, 0.50000000000000, 0.88729833462074),
        (0.27777777777778, 0.44444444444444, 0.27777777777778)),
}
_A, _B = 0.16666666666667, 0.66666666666667
_TRIANGLE_RULES = {
    3: (((_A, _A), (_A, _B), (_B, _A)), (1 / 6, 1 / 6, 1 / 6)),
}


def _accumulate(terms):
    """sum in the given order starting from the integer 0"""
    total = 0
    for t in terms:
        total += t
    return total


class Quadrature:
    _rules = _INTERVAL_RULES

    def __init__(self, order):
        self.p = self.order_to_points(order)
        self.w = self.order_to_weights(order)

    @classmethod
    def _lookup(cls, order, which):
        rule = cls._rules.get(order)
        return _INVALID if rule is None else np.array(rule[which])

    @classmethod
    def order_to_points(cls, order):
        return cls._lookup(order, 0)

    @classmethod
    def order_to_weights(cls, order):
        return cls._lookup(order, 1)

    def get_points(self):
        return self.p

    def get_weights(self):
        return self.w

    def compute(self, phi, index):
        """int phi_i phi_j over the reference element, index = (i, j)"""
        i, j = index[0], index[1]
        return _accumulate(phi.evaluate(x, i) * phi.evaluate(x, j) * w for x, w in zip(self.p, self.w))

    def compute_inter(self, phi, index, fine_p, coarse_p):
        """int phi_i(fine coordinate) phi_j(coarse coordinate) over an intersection whose quadrature points are given
        in the coordinates of both elements (CouplingOperator.compute_b_1d)"""
        i, j = index[0], index[1]
        return _accumulate(phi.evaluate(xf, i) * phi.evaluate(xc, j) * w for xf, xc, w in zip(fine_p, coarse_p, self.w))

    def compute_single(self, phi, index, fun):
        """int phi_index f over the reference element"""
        return _accumulate(phi.evaluate(x, index) * fun.evaluate(x) * w for x, w in zip(self.p, self.w))


class Quadrature2D(Quadrature):
    _rules = _TRIANGLE_RULES

    def compute_grad(self, d_phi, jac_inv, index):
        """sum over the points of (J^-T grad phi_i)^T (J^-T grad phi_j), times w[i] (see the module docstring)"""
        i, j = index[0], index[1]
        wi = self.get_weights()[i]

        def term(x):
            gi_t = (jac_inv @ d_phi.evaluate(x, i)).T
            return (gi_t @ jac_inv @ d_phi.evaluate(x, j))[0][0] * wi
        return _accumulate(term(x) for x in self.get_points())
