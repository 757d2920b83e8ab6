"""learn_multigrid/assembly/MapReferenceElement.py:10-20"""


def g_function(x_ref, x_a, x_b):
    return x_a + x_ref * (x_b - x_a)


def inv_g_function(x, x_a, x_b):
    return (x - x_a) / (x_b - x_a)
