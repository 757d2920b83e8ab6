"""Affine maps between the reference interval [0, 1] and a physical element [x_a, x_b]
(learn_multigrid/assembly/MapReferenceElement.py:10-20).  `IntervalMap` holds an element once and maps many points;
`g_function` / `inv_g_function` are the reference's free functions on top of it.  The arithmetic is the reference's,
operation for operation: x_a + t * (x_b - x_a) and (x - x_a) / (x_b - x_a)."""


class IntervalMap:
    __slots__ = ("left", "length")

    def __init__(self, x_a, x_b):
        self.left = x_a
        self.length = x_b - x_a

    def to_physical(self, t):
        return self.left + t * self.length

    def to_reference(self, x):
        return (x - self.left) / self.length


def g_function(x_ref, x_a, x_b):
    """reference -> physical"""
    return IntervalMap(x_a, x_b).to_physical(x_ref)


def inv_g_function(x, x_a, x_b):
    """physical -> reference"""
    return IntervalMap(x_a, x_b).to_reference(x)
