"""plot_intersections(fine_mesh, coarse_mesh, union) with the reference's signature
(learn_multigrid/utilities/plots.py:5-33): three stacked axes with the fine nodes, the coarse nodes and the union of
both (the segment end points of Intersection.find_intersections1d).  matplotlib is imported when the function is
called, so that the package does not depend on it (plotting is outside the solve path)."""
import numpy as np


def plot_intersections(fine_mesh, coarse_mesh, union):
    import matplotlib.pyplot as plt
    panels = (("Fine Mesh", np.asarray(fine_mesh.get_mesh())), ("Coarse Mesh", np.asarray(coarse_mesh.get_mesh())),
              ("Intersections", np.asarray(union)))
    for k, (title, x) in enumerate(panels):
        plt.subplot(3, 1, k + 1)
        plt.plot(x, np.zeros(len(x)), "ro")
        plt.grid()
        plt.xticks(np.arange(0, 1.1, 0.1))
        plt.title(title)
    plt.show()
