"""Right-hand-side function wrapper with the reference's interface (learn_multigrid/assembly/LoadFunction.py:3-14):
`LoadFunction(f).evaluate(x)` is `f(x)`; the object is also callable, which is what the vectorised load assembly uses."""


class LoadFunction:
    __slots__ = ("fun",)

    def __init__(self, fun):
        if not callable(fun):
            raise TypeError("LoadFunction needs a callable f(x)")
        self.fun = fun

    def __call__(self, points=None):
        return self.fun(points)

    evaluate = __call__

    def get_functions(self):
        return self.fun
