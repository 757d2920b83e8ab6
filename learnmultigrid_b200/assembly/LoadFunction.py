class LoadFunction:
    """learn_multigrid/assembly/LoadFunction.py:3-14"""

    def __init__(self, fun):
        self.fun = fun

    def evaluate(self, points=None):
        return self.fun(points)

    def get_functions(self):
        return self.fun
