"""`Element` exists in the reference as an empty placeholder (learn_multigrid/mesh/Element1D.py); scripts only import
it.  Here it can optionally carry an interval, which Mesh1D.elements() uses."""


class Element:

    def __init__(self, index=None, left=None, right=None):
        self.index, self.left, self.right = index, left, right

    @property
    def length(self):
        return None if self.left is None or self.right is None else self.right - self.left

    def __repr__(self):
        return "Element(%r, %r, %r)" % (self.index, self.left, self.right)
