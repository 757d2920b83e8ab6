"""The C-ABI library loads and exports exactly what include/mgb200.h declares (no compute calls: no GPU here)."""
import ctypes
import os
import re

from learnmultigrid_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(testing=False):
    """product prototypes of mgb200.h, or (testing=True) those inside its `#ifdef MGB_TESTING` blocks"""
    text = open(os.path.join(ROOT, "include", "mgb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    blocks = re.findall(r"#ifdef MGB_TESTING(.*?)#endif", text, flags=re.S)
    if testing:
        text = "\n".join(blocks)
    else:
        text = re.sub(r"#ifdef MGB_TESTING.*?#endif", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    names = declared_functions()
    assert len(names) >= 30
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "declared in mgb200.h but not exported: " + n
        assert n in _lib._SIGNATURES, "declared in mgb200.h but not bound in _lib.py: " + n
    for n in _lib._SIGNATURES:
        assert n in names, "bound in _lib.py but not declared in mgb200.h: " + n


def test_host_emulations_live_in_the_testing_library_only():
    """the serial host twins of device code (mg_host_nn_*, mg_host_coupling_pairs_p1_2d, mg_host_color_rounds) are test
    infrastructure: exported by libmgb200_testing.so, absent from the product library"""
    twins = declared_functions(testing=True)
    assert len(twins) == 5 and sorted(twins) == sorted(_lib._TESTING_SIGNATURES)
    product = ctypes.CDLL(_lib.LIB_PATH)
    testing = _lib.load_testing()
    for n in twins:
        assert not hasattr(product, n), "test-only symbol in the product library: " + n
        assert hasattr(testing, n)


def test_library_loads_and_reports_version():
    lib = _lib.load()
    assert lib.mg_version() == 200
    assert lib.mg_last_error() is not None


def test_struct_layouts_match_header_sizes():
    # mg_sell: 3 x int64 + 3 pointers + 2 x int64 + 2 pointers + 2 x int32 + 4 pointers; mg_cycle_params: 3 x int32 (+pad) + double + 3 x int32 (+pad)
    assert ctypes.sizeof(_lib.mg_sell) == 136
    assert ctypes.sizeof(_lib.mg_cycle_params) == 40
    assert ctypes.sizeof(_lib.mg_level) % 8 == 0
    lib = _lib.load()
    for which, st in enumerate((_lib.mg_sell, _lib.mg_level, _lib.mg_cycle_params, _lib.mg_bcr, _lib.mg_comm,
                                _lib.mg_xfer, _lib.mg_dist_level, _lib.mg_bcr_dist, _lib.mg_dist_norm)):
        assert lib.mg_struct_size(which) == ctypes.sizeof(st), st.__name__


def test_host_helpers_run_without_gpu():
    import numpy as np
    import scipy.sparse as sp
    from learnmultigrid_b200 import formats as F
    A = sp.csr_matrix(sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(9, 9)))
    colors, nc = F.greedy_colors(A)
    assert nc == 2 and list(colors) == [0, 1] * 4 + [0]
    lp, lr = F.lex_levels(A)
    assert len(lp) == 10 and list(lr) == list(range(9))


def test_no_cpu_fallback_without_cuda():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import numpy as np
    import scipy.sparse as sp
    from learnmultigrid_b200.solvers.Multigrid import SemiGeometricMG
    A = sp.csr_matrix(sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(9, 9)))
    mg = SemiGeometricMG(A, np.ones((9, 1)), sp.csr_matrix(np.ones((9, 5))))
    with pytest.raises(_lib.MgError):
        mg.solve(levels=2, smoother="Jacobi")
