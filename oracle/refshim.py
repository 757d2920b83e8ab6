"""oracle/refshim.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Makes the reference's OWN modules importable in the authoring container so that the restatement
can be diffed against them and golden vectors generated (tests/golden/make_golden.py).  /root/reference
does not exist on the GPU box: nothing run there imports this module.

Shims:
  1. no-op `matplotlib`, `matplotlib.pyplot` (Solver.py:5, Mesh1D.py:4, Mesh2D.py:4-5 import them);
  2. `pyamg.relaxation.relaxation.gauss_seidel` = oracle.kernels.gauss_seidel, the restated PyAMG
     kernel (PyAMG is not installed and cannot be: no network);
  3. `np.asscalar` for CG.py:30,32,47; `np.int` for Mesh2D.refine;
  4. (opt-in, `patch_refine()`) Mesh2D.refine builds its per-element scratch as `np.zeros((3, 1))`, whose
     elements NumPy >= 1.24 no longer accepts inside the list assignments of Mesh2D.py:149-156; the module's `np`
     is replaced by a proxy that hands out `np.zeros(3)` for exactly that shape.  Nothing else changes, the
     reference's statements run as they stand.
"""
import contextlib
import io
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("LEARNMG_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "learn_multigrid"))


def install():
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    from . import kernels as K

    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")

        def _noop(*a, **k):
            return None
        for name in ("plot", "yscale", "legend", "title", "ylabel", "xlabel", "show", "grid", "xticks",
                     "triplot", "figure", "savefig", "scatter", "text", "annotate", "axis", "subplots"):
            setattr(plt, name, _noop)
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if "pyamg" not in sys.modules:
        pyamg = types.ModuleType("pyamg")
        relax = types.ModuleType("pyamg.relaxation")
        relax2 = types.ModuleType("pyamg.relaxation.relaxation")
        relax2.gauss_seidel = K.gauss_seidel
        relax.relaxation = relax2
        pyamg.relaxation = relax
        sys.modules["pyamg"] = pyamg
        sys.modules["pyamg.relaxation"] = relax
        sys.modules["pyamg.relaxation.relaxation"] = relax2
    if not hasattr(np, "asscalar"):
        np.asscalar = lambda a: a.item()
    if not hasattr(np, "int"):
        np.int = int
    # the product also ships a drop-in package called learn_multigrid; make sure the REFERENCE wins here
    for m in [m for m in sys.modules if m == "learn_multigrid" or m.startswith("learn_multigrid.")]:
        del sys.modules[m]
    if REFERENCE_ROOT in sys.path:
        sys.path.remove(REFERENCE_ROOT)
    sys.path.insert(0, REFERENCE_ROOT)


@contextlib.contextmanager
def quiet():
    """The reference prints on every constructor and iteration (Multigrid.py:32,61,68)."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def patch_refine():
    """shim 4: let the reference's own Mesh2D.refine run under NumPy >= 1.24 (call after install())"""
    import learn_multigrid.mesh.Mesh2D as ref_mesh

    class _NumpyProxy:
        def __getattr__(self, name):
            return getattr(np, name)

        @staticmethod
        def zeros(shape, *args, **kwargs):
            if shape == (3, 1):
                shape = (3,)
            return np.zeros(shape, *args, **kwargs)

    ref_mesh.np = _NumpyProxy()
    return ref_mesh
