"""oracle/kernels.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

ctypes bindings of oracle/relax_kernels.c (plain C restatement of PyAMG's gauss_seidel and
SciPy's csr_matvec as reached from learn_multigrid/solvers/Multigrid.py:62,88,90,93,115,121).
"""
import ctypes
import os
import subprocess

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle_relax.so")
_lib = None


def build(force=False):
    """Compile relax_kernels.c with gcc (oracle/Makefile)."""
    if force or not os.path.exists(_LIB_PATH) or (
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "relax_kernels.c"))):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        i32p = ctypes.POINTER(ctypes.c_int32)
        f64p = ctypes.POINTER(ctypes.c_double)
        i64 = ctypes.c_int64
        _lib.oracle_gauss_seidel.argtypes = [i32p, i32p, f64p, f64p, f64p, i64, ctypes.c_int]
        _lib.oracle_gauss_seidel_rows.argtypes = [i32p, i32p, f64p, f64p, f64p, i32p, i64]
        _lib.oracle_residual.argtypes = [i32p, i32p, f64p, f64p, f64p, f64p, i64]
        _lib.oracle_spmv.argtypes = [i32p, i32p, f64p, f64p, f64p, i64]
        _lib.oracle_jacobi.argtypes = [i32p, i32p, f64p, f64p, f64p, f64p, f64p, ctypes.c_double, i64]
        _lib.oracle_prolong_correct.argtypes = [i32p, i32p, f64p, f64p, f64p, i64]
        for f in ("oracle_gauss_seidel", "oracle_gauss_seidel_rows", "oracle_residual", "oracle_spmv",
                  "oracle_jacobi", "oracle_prolong_correct"):
            getattr(_lib, f).restype = None
    return _lib


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def _f64(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _csr(A):
    """CSR view (what PyAMG does first: `A = csr_matrix(A)` if not CSR); entries keep their storage order."""
    return sp.csr_matrix(A)


def gauss_seidel(A, x, b, iterations=1, sweep="forward"):
    """PyAMG `gauss_seidel(A, x, b, iterations, sweep='forward')` restated.

    In place on `x` through a ravel() view exactly like PyAMG (x must be contiguous fp64,
    shape (n,) or (n,1)); call sites Multigrid.py:88,121."""
    if sweep != "forward":
        raise ValueError("only the forward sweep is on the reference path")
    A = _csr(A)
    xv = np.ravel(x)
    if not np.shares_memory(xv, x):
        raise ValueError("x must be contiguous (PyAMG raises here too)")
    bv = np.ascontiguousarray(np.ravel(b), dtype=np.float64)
    if xv.dtype != np.float64 or A.dtype != np.float64:
        raise TypeError("arguments A, x, b must have the same dtype (float64)")
    ip, ipp = _i32(A.indptr)
    ij, ijp = _i32(A.indices)
    lib().oracle_gauss_seidel(ipp, ijp, _f64(A.data), _f64(xv), _f64(bv), A.shape[0], int(iterations))


def gauss_seidel_multicolor(A, x, b, color_rows, iterations=1):
    """Multicolour GS: for each sweep, for each colour in order, update that colour's rows in place.
    `color_rows` is a list of int arrays (rows of each colour)."""
    A = _csr(A)
    xv = np.ravel(x)
    assert np.shares_memory(xv, x)
    bv = np.ascontiguousarray(np.ravel(b), dtype=np.float64)
    ip, ipp = _i32(A.indptr)
    ij, ijp = _i32(A.indices)
    rows = [_i32(r) for r in color_rows]
    for _ in range(int(iterations)):
        for r, rp in rows:
            lib().oracle_gauss_seidel_rows(ipp, ijp, _f64(A.data), _f64(xv), _f64(bv), rp, r.size)


def residual(A, x, b):
    """r = b - A x (SciPy csr_matvec order); returns the shape of b."""
    A = _csr(A)
    xv = np.ascontiguousarray(np.ravel(x), dtype=np.float64)
    bv = np.ascontiguousarray(np.ravel(b), dtype=np.float64)
    r = np.empty_like(bv)
    ip, ipp = _i32(A.indptr)
    ij, ijp = _i32(A.indices)
    lib().oracle_residual(ipp, ijp, _f64(A.data), _f64(xv), _f64(bv), _f64(r), A.shape[0])
    return r.reshape(np.shape(b))


def spmv(A, x):
    A = _csr(A)
    xv = np.ascontiguousarray(np.ravel(x), dtype=np.float64)
    y = np.empty(A.shape[0], dtype=np.float64)
    ip, ipp = _i32(A.indptr)
    ij, ijp = _i32(A.indices)
    lib().oracle_spmv(ipp, ijp, _f64(A.data), _f64(xv), _f64(y), A.shape[0])
    return y.reshape((A.shape[0],) + np.shape(x)[1:])


def jacobi(A, x, b, dinv, omega=1.0, iterations=1):
    """`iterations` damped-Jacobi sweeps x <- x + omega*(dinv*(b - A x)); returns the new x
    (same shape as x).  omega=1 is learn_multigrid/solvers/Jacobi.py:35."""
    A = _csr(A)
    cur = np.array(np.ravel(x), dtype=np.float64, copy=True)
    nxt = np.empty_like(cur)
    bv = np.ascontiguousarray(np.ravel(b), dtype=np.float64)
    dv = np.ascontiguousarray(np.ravel(dinv), dtype=np.float64)
    ip, ipp = _i32(A.indptr)
    ij, ijp = _i32(A.indices)
    for _ in range(int(iterations)):
        lib().oracle_jacobi(ipp, ijp, _f64(A.data), _f64(dv), _f64(cur), _f64(bv), _f64(nxt),
                            float(omega), A.shape[0])
        cur, nxt = nxt, cur
    return cur.reshape(np.shape(x))


def prolong_correct(Q, e, u):
    """u + Q e (Multigrid.py:115); returns a new array shaped like u."""
    Q = _csr(Q)
    ev = np.ascontiguousarray(np.ravel(e), dtype=np.float64)
    out = np.array(np.ravel(u), dtype=np.float64, copy=True)
    ip, ipp = _i32(Q.indptr)
    ij, ijp = _i32(Q.indices)
    lib().oracle_prolong_correct(ipp, ijp, _f64(Q.data), _f64(ev), _f64(out), Q.shape[0])
    return out.reshape(np.shape(u))
