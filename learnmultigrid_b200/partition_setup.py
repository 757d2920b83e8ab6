"""Strip-local Galerkin products: the hierarchy setup of a row-partitioned solve WITHOUT any rank holding a global
operator ("AQ needs remote rows of Q referenced by halo columns (one exchange of CSR row slices), Q^T(AQ)
needs a sparse transpose exchange across strip boundaries").  Host logic + exchange plan, independent of where the
local sparse products run (`ops`: SciPy here for the CPU tests; the device SpGEMM of setup_device.py has the same
contract), so that DistributedHierarchy can stop forming the global hierarchy on every GPU (DESIGN 12, weak scaling).

What one rank owns on level l:  rows [f0, f1) of A_l and of Q_l (global column ids), rows [c0, c1) of Q_l^T.
What it must produce:           rows [c0, c1) of A_{l+1} = Q^T A Q  with the bits of the single-process product.

The single-process evaluation is  T = A^T Q,  C = Q^T T,  A_c = C^T  (setup_device.DeviceSetup.galerkin, which is how
SciPy evaluates `csr_matrix(i.T @ A @ i)` with A held as CSC, learn_multigrid/solvers/Multigrid.py:97-98), every
entry accumulated over the inner index in ascending order, exact zeros dropped after each product.  For the owned
coarse rows i:

    A_c[i, j] = C[j, i] = sum_f Q[f, j] * T[f, i],      T[f, i] = sum_g A[g, f] * Q[g, i]

so the rank needs   F1 = {g : Q[g, i] != 0 for an owned i}      = the columns of its Q^T rows   -> rows F1 of A,
                    F2 = {f : A[g, f] != 0 for a g in F1}       = the columns of those A rows   -> rows F2 of Q,
and nothing else.  Rows are fetched from their owners (`fetch_rows`: two all-gathers through the fabric -- setup-time
traffic, a few grid lines per neighbour); the three local products then run on index-compressed blocks.  Compression
is monotone, so every sum keeps its order and its bits: the assembled row blocks equal the global product exactly,
values and sparsity pattern (tests/test_partition_setup.py, 2-8 ranks, 1D / 2D, linear and quasi-L2 transfers).
"""
import numpy as np
import scipy.sparse as sp

from . import formats as F
from . import partition as PT


class ScipyOps:
    """local sparse kernels with the contract of setup_device.DeviceSetup (CSR in, CSR out, sorted columns,
    csr_matmat accumulation order, exact zeros pruned)"""

    @staticmethod
    def spgemm(A, B):
        C = sp.csr_matrix(sp.csr_matrix(A) @ sp.csr_matrix(B))
        C.eliminate_zeros()            # csr_matmat already skips exact-zero sums; explicit for clarity
        C.sort_indices()
        return C

    @staticmethod
    def transpose(A):
        T = sp.csr_matrix(sp.csr_matrix(A).T)
        T.sort_indices()
        return T


def _canon(M):
    M = sp.csr_matrix(M)
    M.sort_indices()
    return M


def fetch_rows(fabric, offsets, block, wanted):
    """Rows `wanted` (sorted unique global row ids) of a row-partitioned CSR matrix whose rows [offsets[r], offsets[r+1])
    live on rank r as `block` (global column ids).  Returns a CSR matrix with len(wanted) rows in the order of `wanted`.
    Collective: every rank must call it."""
    rank, world = fabric.rank, fabric.world
    wanted = np.asarray(wanted, dtype=np.int64)
    block = _canon(block)
    o0 = int(offsets[rank])
    owner = PT.owner_of(offsets, wanted) if len(wanted) else np.zeros(0, dtype=np.int64)
    requests = {int(q): wanted[owner == q] for q in np.unique(owner) if int(q) != rank}
    everyone = fabric.allgather(requests)                      # everyone[p][q] = rows p wants from q
    replies = {}
    for p in range(world):
        ids = everyone[p].get(rank)
        if ids is None or p == rank:
            continue
        sub = block[np.asarray(ids, dtype=np.int64) - o0]
        replies[p] = (sub.indptr.astype(np.int64), sub.indices.astype(np.int64), sub.data.astype(np.float64))
    answers = fabric.allgather(replies)                        # answers[q][p] = what q sends to p
    parts = []
    ncols = block.shape[1]
    for q in sorted(set(int(v) for v in np.unique(owner))):
        ids = wanted[owner == q]
        if q == rank:
            sub = block[ids - o0]
        else:
            ip, ix, va = answers[q][rank]
            sub = sp.csr_matrix((va, ix, ip), shape=(len(ids), ncols))
        parts.append(sub)
    if not parts:
        return sp.csr_matrix((0, ncols))
    out = sp.vstack(parts, format="csr")                       # owners ascending = global row order (blocks contiguous)
    out.sort_indices()
    return out


def fetch_values(fabric, offsets, own_values, wanted):
    """entries `wanted` (sorted unique global ids) of a vector whose block [offsets[r], offsets[r+1]) lives on rank r
    (colours of the halo nodes, for instance).  Collective."""
    rank, world = fabric.rank, fabric.world
    wanted = np.asarray(wanted, dtype=np.int64)
    own_values = np.asarray(own_values)
    o0 = int(offsets[rank])
    owner = PT.owner_of(offsets, wanted) if len(wanted) else np.zeros(0, dtype=np.int64)
    everyone = fabric.allgather({int(q): wanted[owner == q] for q in np.unique(owner) if int(q) != rank})
    answers = fabric.allgather({p: own_values[np.asarray(everyone[p][rank], dtype=np.int64) - o0]
                                for p in range(world) if p != rank and rank in everyone[p]})
    out = np.empty(len(wanted), dtype=own_values.dtype)
    for q in np.unique(owner):
        sel = owner == q
        out[sel] = own_values[wanted[sel] - o0] if int(q) == rank else answers[int(q)][rank]
    return out


def greedy_colors_distributed(fabric, offsets, A_blk):
    """The first-fit colouring of formats.greedy_colors (index order on the symmetrised pattern, rows with only a
    diagonal entry last) of a row-partitioned matrix: returns (colours of the own rows, number of colours of the whole
    matrix), identical to the single-process result.  The colouring is sequential by nature, so the blocks are coloured
    in rank order (world rounds, one all-gather each): a rank receives the colours of the lower ranks' rows it refers to
    and the colours those ranks' rows forbid for its own rows, colours its rows with off-diagonal entries, and publishes
    the same for the others; the rows with only a diagonal entry are coloured locally at the end.  Collective."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    rank, world = fabric.rank, fabric.world
    A_blk = _canon(A_blk)
    o0, o1 = int(offsets[rank]), int(offsets[rank + 1])
    n_own = o1 - o0
    cols = np.asarray(A_blk.indices, dtype=np.int64)
    ext = np.unique(cols[(cols < o0) | (cols >= o1)])
    n_ext = len(ext)
    local = np.where((cols >= o0) & (cols < o1), cols - o0, n_own + np.searchsorted(ext, cols)).astype(np.int32)
    indptr = np.ascontiguousarray(A_blk.indptr, dtype=np.int32)
    ext_color = -np.ones(max(n_ext, 1), dtype=np.int32)
    lo = np.zeros(n_own + n_ext + 1, dtype=np.uint64)
    hi = np.zeros(n_own + n_ext + 1, dtype=np.uint64)
    colors = -np.ones(max(n_own, 1), dtype=np.int32)
    ext_owner = PT.owner_of(offsets, ext) if n_ext else np.zeros(0, dtype=np.int64)
    wanted_by = fabric.allgather(ext)                          # wanted_by[p] = global ids rank p refers to

    def vp(a):
        return a.ctypes.data_as(ctypes.c_void_p)

    def run(phase):
        rc = lib.mg_host_greedy_color_block(n_own, n_ext, vp(indptr), vp(local), vp(ext_color), vp(lo), vp(hi),
                                            vp(colors), phase)
        if rc < 0:
            _lib.check(rc, "mg_host_greedy_color_block")

    for turn in range(world):
        payload = None
        if turn == rank:
            run(0)
            # for every other rank: the colours of my rows it refers to, and what my rows forbid for its rows
            payload = {}
            for p in range(world):
                if p == rank:
                    continue
                ids = wanted_by[p]
                mine = ids[(ids >= o0) & (ids < o1)]
                sel = np.flatnonzero(ext_owner == p)
                push = sel[(lo[n_own + sel] != 0) | (hi[n_own + sel] != 0)]
                payload[p] = (mine, colors[mine - o0].copy(), ext[push], lo[n_own + push].copy(), hi[n_own + push].copy())
        got = fabric.allgather(payload)[turn]
        if turn != rank and got is not None and rank in got:
            ids, cs, pushed, plo, phi = got[rank]
            if len(ids):
                ext_color[np.searchsorted(ext, ids)] = cs          # -1 for rows the owner colours in its last phase
            if len(pushed):
                lo[pushed - o0] |= plo
                hi[pushed - o0] |= phi
    run(1)
    ncolors = max(fabric.allgather(int(colors[:n_own].max()) + 1 if n_own else 0))
    return colors[:n_own].copy(), int(ncolors)


def rank_plan_from_blocks(fabric, offs, offs_next, offs_prev, A_blk, QT_blk, Q_prev_blk, own_colors, ncolors):
    """partition.RankPlan of this rank on one level from its own row blocks only: the external columns of its rows of
    A_l, of Q_l^T (coarse rows it owns) and of Q_{l-1} (finer rows it owns), and the colours of those halo nodes fetched
    from their owners.  own_colors None = no colouring (Jacobi).  Collective."""
    rank = fabric.rank
    o0, o1 = int(offs[rank]), int(offs[rank + 1])
    parts = []
    for blk in (A_blk, QT_blk, Q_prev_blk):
        if blk is None:
            continue
        cols = np.asarray(_canon(blk).indices, dtype=np.int64)
        parts.append(np.unique(cols[(cols < o0) | (cols >= o1)]))
    ext = np.unique(np.concatenate(parts)) if parts else np.zeros(0, dtype=np.int64)
    if own_colors is None:
        fabric.allgather(None)                       # keep the collective sequence identical on every rank
        fabric.allgather(None)
        return PT.RankPlan(offs, rank, ext, None)
    ext_colors = fetch_values(fabric, offs, np.asarray(own_colors), ext)
    return PT.RankPlan(offs, rank, ext, (np.asarray(own_colors), ext_colors, int(ncolors)))


def _compress_columns(M, keep):
    """columns `keep` (sorted unique) of M renumbered 0..len(keep)-1; all of M's columns must be in `keep`"""
    M = _canon(M)
    pos = np.searchsorted(keep, M.indices)
    if len(M.indices) and (pos.max() >= len(keep) or np.any(keep[pos] != M.indices)):
        raise ValueError("a column outside the compression set")
    return sp.csr_matrix((M.data, pos.astype(np.int64), M.indptr), shape=(M.shape[0], len(keep)))


def galerkin_row_block(fabric, offs_f, offs_c, A_blk, Q_blk, QT_blk, ops=ScipyOps):
    """Rows [offs_c[rank], offs_c[rank+1]) of A_c = Q^T A Q (global column ids) from this rank's row blocks of A and Q
    (fine rows [offs_f[rank], offs_f[rank+1])) and of Q^T (coarse rows).  Collective.  Bit-identical to the rows of the
    single-process product T = A^T Q, C = Q^T T, A_c = C^T."""
    rank = fabric.rank
    QT_blk = _canon(QT_blk)
    n_own = QT_blk.shape[0]
    n_c = int(offs_c[-1])
    F1 = np.unique(QT_blk.indices).astype(np.int64)            # fine rows g with Q[g, i] != 0 for an owned i
    A_F1 = fetch_rows(fabric, offs_f, A_blk, F1)               # rows F1 of A
    F2 = np.unique(A_F1.indices).astype(np.int64)              # fine rows f reached through A[g, f]
    Q_F2 = fetch_rows(fabric, offs_f, Q_blk, F2)               # rows F2 of Q
    J = np.unique(Q_F2.indices).astype(np.int64)               # coarse columns j of the owned rows of A_c
    if n_own == 0:
        return sp.csr_matrix((0, n_c))
    # T[F2, own] = (A[F1, F2])^T  Q[F1, own]      (sum over g in F1, ascending)
    Q_F1_own = ops.transpose(_compress_columns(QT_blk, F1))    # |F1| x n_own  = Q[F1, own]
    AT_loc = ops.transpose(_compress_columns(A_F1, F2))        # |F2| x |F1|   = A^T[F2, F1]
    T_loc = ops.spgemm(AT_loc, Q_F1_own)                       # |F2| x n_own
    # C[J, own] = (Q[F2, J])^T  T[F2, own]        (sum over f in F2, ascending)
    QT_loc = ops.transpose(_compress_columns(Q_F2, J))         # |J| x |F2|
    C_loc = ops.spgemm(QT_loc, T_loc)                          # |J| x n_own
    Ac_loc = _canon(ops.transpose(C_loc))                      # n_own x |J|
    return sp.csr_matrix((Ac_loc.data, J[Ac_loc.indices], Ac_loc.indptr), shape=(n_own, n_c))


def transpose_row_block(fabric, offs_rows, offs_cols, M_blk):
    """Rows [offs_cols[rank], offs_cols[rank+1]) of M^T from the row blocks of M (rows partitioned by offs_rows):
    every rank cuts its block by the column ranges of the others and the pieces are exchanged (the "sparse transpose
    exchange" of SURVEY 8e).  Entries of a row of M^T come out in ascending column order = ascending row of M, the order
    the restriction kernel adds in."""
    rank, world = fabric.rank, fabric.world
    M_blk = sp.csc_matrix(_canon(M_blk))
    r0 = int(offs_rows[rank])
    pieces = {}
    for q in range(world):
        c0, c1 = int(offs_cols[q]), int(offs_cols[q + 1])
        sub = sp.csr_matrix(M_blk[:, c0:c1].T)                 # (c1-c0) x n_own_rows
        sub.sort_indices()
        pieces[q] = (sub.indptr.astype(np.int64), sub.indices.astype(np.int64) + r0, sub.data.astype(np.float64))
    everyone = fabric.allgather(pieces)
    n_rows_total = int(offs_rows[-1])
    n_own = int(offs_cols[rank + 1] - offs_cols[rank])
    acc = sp.csr_matrix((n_own, n_rows_total))
    for p in range(world):                                     # blocks have disjoint column ranges: a plain sum
        ip, ix, va = everyone[p][rank]
        acc = acc + sp.csr_matrix((va, ix, ip), shape=(n_own, n_rows_total))
    acc = sp.csr_matrix(acc)
    acc.sort_indices()
    return acc


def build_strip_hierarchy(fabric, A_blk, Q_blks, offsets, ops=ScipyOps):
    """Row blocks of every level operator of the Galerkin hierarchy from this rank's row block of A_0 and of every
    Q_l.  offsets[l] = block offsets of level l (len world+1).  Returns ([A_0 block, A_1 block, ...],
    [Q_0^T block, ...]) -- the per-rank inputs DistributedHierarchy cuts out of the global matrices today."""
    A_blks, QT_blks = [_canon(A_blk)], []
    for l, Q_blk in enumerate(Q_blks):
        QT = transpose_row_block(fabric, offsets[l], offsets[l + 1], Q_blk)
        QT_blks.append(QT)
        A_blks.append(galerkin_row_block(fabric, offsets[l], offsets[l + 1], A_blks[l], Q_blk, QT, ops))
    return A_blks, QT_blks


def gather_row_blocks(fabric, block, n_cols=None):
    """The row blocks of all ranks stacked in rank order = the global matrix, on every rank (the first replicated
    level of a partitioned hierarchy and the transfer operators below it are small: DistributedHierarchy keeps them in
    full on every GPU).  Collective."""
    block = _canon(block)
    parts = fabric.allgather((block.indptr.astype(np.int64), block.indices.astype(np.int64),
                              block.data.astype(np.float64), block.shape))
    n_cols = block.shape[1] if n_cols is None else int(n_cols)
    mats = [sp.csr_matrix((va, ix, ip), shape=(shape[0], n_cols)) for ip, ix, va, shape in parts]
    out = sp.vstack(mats, format="csr")
    out.sort_indices()
    return out


class StripLevel:
    """what one rank holds of a partitioned level after strip_local_setup: its rows of A_l and Q_l (fine rows it owns),
    its rows of Q_l^T (coarse rows it owns) -- global column ids --, the colours of its rows and its halo plan"""

    def __init__(self, A, Q, QT, colors, ncolors, plan):
        self.A, self.Q, self.QT, self.colors, self.ncolors, self.plan = A, Q, QT, colors, ncolors, plan


def strip_local_setup(fabric, A_blk, Q_blks, offsets, n_dist, smoother="mcgs", colors=None, ops=ScipyOps):
    """The whole host side of a partitioned hierarchy setup from this rank's row blocks only.

    A_blk    : rows [offsets[0][rank], offsets[0][rank+1]) of A_0 (global column ids)
    Q_blks   : rows [offsets[l][rank], offsets[l][rank+1]) of every Q_l
    offsets  : block offsets per level (partition.block_offsets), len = number of levels
    n_dist   : levels 0..n_dist-1 stay partitioned, level n_dist and below are replicated
    colors   : optional per-level colours of the OWN rows (e.g. a structured colouring evaluated on the own rows);
               default for "mcgs": the first-fit colouring, coloured block by block in rank order
    Returns (levels, A_rep, Q_rep): `levels[l]` a StripLevel for l < n_dist; `A_rep` = A_{n_dist} in full and `Q_rep` =
    [Q_l in full for l >= n_dist], gathered once -- the inputs of the replicated tail, whose remaining Galerkin
    products are small and formed redundantly on every rank as before.  Everything equals what DistributedHierarchy
    derives from the global operators today (tests/test_partition_setup.py).  Collective."""
    L = len(Q_blks) + 1
    if len(offsets) != L:
        raise ValueError("one offsets array per level expected")
    if not 1 <= n_dist <= L - 1:
        raise ValueError("n_dist must be between 1 and levels-1")
    if smoother not in ("jacobi", "mcgs"):
        raise ValueError("partitioned levels support 'jacobi' and 'mcgs'")
    Q_blks = [_canon(Q) for Q in Q_blks]
    A_blks, QT_blks = build_strip_hierarchy(fabric, A_blk, Q_blks[:n_dist], offsets, ops)
    levels = []
    for l in range(n_dist):
        own_col, nc = None, 0
        if smoother == "mcgs":
            if colors is not None and colors[l] is not None:
                own_col = np.ascontiguousarray(colors[l], dtype=np.int32)
                if len(own_col) != A_blks[l].shape[0]:
                    raise ValueError("level %d: one colour per own row expected" % l)
                nc = max(fabric.allgather(int(own_col.max()) + 1 if len(own_col) else 0))
            else:
                own_col, nc = greedy_colors_distributed(fabric, offsets[l], A_blks[l])
        plan = rank_plan_from_blocks(fabric, offsets[l], offsets[l + 1], offsets[l - 1] if l else None, A_blks[l],
                                     QT_blks[l], Q_blks[l - 1] if l else None, own_col, nc)
        levels.append(StripLevel(A_blks[l], Q_blks[l], QT_blks[l], own_col, nc, plan))
    A_rep = gather_row_blocks(fabric, A_blks[n_dist], int(offsets[n_dist][-1]))
    Q_rep = [gather_row_blocks(fabric, Q_blks[l], int(offsets[l + 1][-1])) for l in range(n_dist, L - 1)]
    return levels, A_rep, Q_rep
