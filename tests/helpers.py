"""Shared test helpers: golden loading, small problem builders, NumPy emulation of the device kernels."""
import os

import numpy as np
import scipy.sparse as sp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def coo_from(d, prefix):
    return sp.coo_matrix((d[prefix + "_data"], (d[prefix + "_row"], d[prefix + "_col"])),
                         shape=tuple(d[prefix + "_shape"])).tocsr()


def history_tolerance(A, x, rel=1e-12):
    """Absolute tolerance for residual-norm histories: two fp64 evaluations of the same V-cycle agree to
    `rel` relative to ||x|| (the north star's per-V-cycle bar), i.e. their residual norms ||b - A x|| agree to
    rel * ||A||_inf * ||x||_2."""
    A = sp.csr_matrix(A)
    return rel * abs(A).sum(axis=1).max() * float(np.linalg.norm(x))


def assert_history_close(got, want, A, x, rel=1e-12):
    got = np.asarray(got).ravel()
    want = np.asarray(want).ravel()
    assert len(got) == len(want), "iteration counts differ: %d vs %d" % (len(got), len(want))
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=history_tolerance(A, x, rel))


def poisson2d(N, dirichlet_rows=True):
    """5-point P1 stiffness on the (N+1)^2 structured grid with row-replaced Dirichlet rows
    (the reference's structured 2D operator)."""
    W = N + 1
    n = W * W
    idx = np.arange(n).reshape(W, W)
    rows, cols, vals = [], [], []
    for dy, dx in ((0, 0), (0, 1), (0, -1), (1, 0), (-1, 0)):
        src = idx[max(0, -dy):W - max(0, dy), max(0, -dx):W - max(0, dx)]
        dst = idx[max(0, dy):W - max(0, -dy), max(0, dx):W - max(0, -dx)]
        rows.append(src.ravel())
        cols.append(dst.ravel())
        vals.append(np.full(src.size, 4.0 if (dy == 0 and dx == 0) else -1.0))
    A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n)).tolil()
    if dirichlet_rows:
        b = np.zeros((W, W), dtype=bool)
        b[0, :] = b[-1, :] = b[:, 0] = b[:, -1] = True
        for i in np.flatnonzero(b.ravel()):
            A.rows[i] = [i]
            A.data[i] = [1.0]
    return sp.csr_matrix(A)


def bilinear_P(Nf):
    """linear interpolation on the P1 right-triangle mesh (diagonal lower-left -> upper-right), (Nf+1)^2 x (Nf/2+1)^2"""
    Wf, Wc = Nf + 1, Nf // 2 + 1
    rows, cols, vals = [], [], []
    for iy in range(Wf):
        for ix in range(Wf):
            r = iy * Wf + ix
            cx, cy = ix // 2, iy // 2
            if ix % 2 == 0 and iy % 2 == 0:
                ent = [(cx, cy, 1.0)]
            elif ix % 2 == 1 and iy % 2 == 0:
                ent = [(cx, cy, 0.5), (cx + 1, cy, 0.5)]
            elif ix % 2 == 0 and iy % 2 == 1:
                ent = [(cx, cy, 0.5), (cx, cy + 1, 0.5)]
            else:
                ent = [(cx, cy, 0.5), (cx + 1, cy + 1, 0.5)]
            for x, y, v in ent:
                rows.append(r)
                cols.append(y * Wc + x)
                vals.append(v)
    return sp.csr_matrix((vals, (rows, cols)), shape=(Wf * Wf, Wc * Wc))


# ---------------------------------------------------------------------------------------------------------
# NumPy emulation of the SELL kernels and of cycle.cu's orchestration (host-logic tests only)
def sell_rowsum(sell, n, x, rows=None, skip_diag=False):
    slice_ptr, cols, vals = sell
    rows = np.arange(n, dtype=np.int64) if rows is None else np.asarray(rows, dtype=np.int64)
    s = rows >> 5
    lane = rows & 31
    base = slice_ptr[s]
    ln = (slice_ptr[s + 1] - base) // 32
    acc = np.zeros(len(rows))
    diag = np.zeros(len(rows))
    for k in range(int(ln.max()) if len(rows) else 0):
        m = ln > k
        idx = base[m] + k * 32 + lane[m]
        c = cols[idx]
        v = vals[idx]
        prod = v * x[c]
        if skip_diag:
            isd = c == rows[m]
            dm = diag[m]
            dm[isd & (v != 0.0)] = v[isd & (v != 0.0)]
            diag[m] = dm
            am = acc[m]
            am[~isd] = am[~isd] + prod[~isd]
            acc[m] = am
        else:
            acc[m] = acc[m] + prod
    return (acc, diag) if skip_diag else acc


def emulate_vcycle(levels, smoother, nu_pre, nu_post, omega, x0, b0, zero_guess_skip=True, coarse_solve=None):
    """Mirror of vcycle_rec in learnmultigrid_b200/csrc/cycle.cu on the host-level dicts of
    formats.build_host_hierarchy (vectors in the levels' own orderings)."""
    L = len(levels)
    xs = [None] * L
    bs = [None] * L
    xs[0] = x0.copy()
    bs[0] = b0.copy()

    def smooth(l, x, b, steps, zero_guess):
        d = levels[l]
        n = d["n"]
        if zero_guess:
            x = np.zeros(n)
        for s in range(steps):
            if smoother == "jacobi":
                if s == 0 and zero_guess and zero_guess_skip:
                    x = 0.0 + omega * (d["dinv"] * b)
                else:
                    r = b - sell_rowsum(d["A_sell"], n, x)
                    x = x + omega * (d["dinv"] * r)
            elif smoother == "mcgs":
                cp = d["color_ptr"]
                for c in range(len(cp) - 1):
                    rows = np.arange(cp[c], cp[c + 1])
                    acc, diag = sell_rowsum(d["A_sell"], n, x, rows, skip_diag=True)
                    ok = diag != 0.0
                    x = x.copy()
                    x[rows[ok]] = (b[rows[ok]] - acc[ok]) / diag[ok]
            else:
                raise ValueError(smoother)
        return x

    def rec(l):
        d = levels[l]
        if l == L - 1:
            xs[l] = coarse_solve(d["A_nat"], bs[l])
            return
        n = d["n"]
        x = smooth(l, xs[l], bs[l], nu_pre, l > 0)
        r = bs[l] - sell_rowsum(d["A_sell"], n, x)
        bs[l + 1] = sell_rowsum(d["QT_sell"], levels[l + 1]["n"], r)
        rec(l + 1)
        x = x + sell_rowsum(d["Q_sell"], n, xs[l + 1])
        xs[l] = smooth(l, x, bs[l], nu_post, False)

    rec(0)
    return xs[0]


def free_port():
    """a TCP port that is free right now (torchrun rendezvous: back-to-back runs must not collide in TIME_WAIT)"""
    import socket
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        return str(sock.getsockname()[1])

