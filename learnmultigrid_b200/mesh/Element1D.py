class Element:
    """placeholder kept for import compatibility (learn_multigrid/mesh/Element1D.py)"""

    def __init__(self):
        pass
