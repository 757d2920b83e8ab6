"""Row-block partition of a multigrid level and its halo plan (pure NumPy host logic, no device work).

A level with n rows is split into `size` contiguous row blocks (strips of the row-major structured grid).  A rank's
level vectors are laid out as [owned rows, colour-blocked | halo], the halo ordered by (owning rank, colour, global
id) so that the entries one neighbour sends for one colour form one contiguous range.  The plan is a pure function
of (offsets, rank, external column set, colours): every rank can compute its neighbours' plans itself, which is how
send lists are obtained without communication (the hierarchy setup is replicated, see distributed.py).
"""
import numpy as np


def block_offsets(n, size):
    """offsets[r] .. offsets[r+1] = rows of rank r (equal contiguous blocks)"""
    return np.array([(n * r) // size for r in range(size + 1)], dtype=np.int64)


def owner_of(offsets, gid):
    return np.searchsorted(offsets, gid, side="right") - 1


class RankPlan:
    """halo plan of one rank on one level"""

    def __init__(self, offsets, rank, ext_cols, colors=None):
        """colors: None, the GLOBAL colour array of the level, or -- for a setup in which no rank holds global data
        (partition_setup.py) -- a tuple (colours of the owned rows, colours of the sorted unique external columns,
        number of colours of the level)."""
        self.offsets = np.asarray(offsets, dtype=np.int64)
        self.rank = int(rank)
        self.o0, self.o1 = int(self.offsets[rank]), int(self.offsets[rank + 1])
        self.n_own = self.o1 - self.o0
        ext = np.unique(np.asarray(ext_cols, dtype=np.int64))
        if len(ext) and (np.any((ext >= self.o0) & (ext < self.o1))):
            raise ValueError("external column set contains owned columns")
        owner = owner_of(self.offsets, ext) if len(ext) else np.zeros(0, dtype=np.int64)
        if isinstance(colors, tuple):
            own_col, hcol_in, ncolors = colors
            own_col, hcol_in = np.asarray(own_col), np.asarray(hcol_in)
            if len(own_col) != self.n_own or len(hcol_in) != len(ext):
                raise ValueError("local colour arrays do not match the row block / the external column set")
            colors = None
            local_colors = True
        else:
            local_colors = False
        if local_colors or colors is not None:
            if not local_colors:
                colors = np.asarray(colors)
                own_col = colors[self.o0:self.o1]
                ncolors = int(colors.max()) + 1 if len(colors) else 0
                hcol_in = colors[ext] if len(ext) else np.zeros(0, dtype=np.int64)
            self.ncolors = int(ncolors)
            self.perm = np.argsort(own_col, kind="stable").astype(np.int32)          # new -> old (block-local)
            self.color_ptr = np.zeros(self.ncolors + 1, dtype=np.int64)
            np.cumsum(np.bincount(own_col, minlength=self.ncolors), out=self.color_ptr[1:])
            hcol = np.asarray(hcol_in, dtype=np.int64) if len(ext) else np.zeros(0, dtype=np.int64)
        else:
            self.ncolors = 0
            self.perm = None
            self.color_ptr = None
            hcol = np.zeros(len(ext), dtype=np.int64)
        if self.perm is not None:
            self.iperm = np.empty(self.n_own, dtype=np.int32)
            self.iperm[self.perm] = np.arange(self.n_own, dtype=np.int32)
        else:
            self.iperm = None
        order = np.lexsort((ext, hcol, owner))                   # by owner, then colour, then global id
        self.halo_gid = ext[order]
        self.halo_owner = owner[order]
        self.halo_color = hcol[order]
        self.n_halo = len(ext)
        self.neighbours = [int(p) for p in np.unique(self.halo_owner)]
        self.seg = {}            # owner -> (start, end) in the halo
        self.seg_color = {}      # (owner, colour) -> (start, end)
        nc = max(self.ncolors, 1)
        for p in self.neighbours:
            idx = np.flatnonzero(self.halo_owner == p)
            s0 = int(idx[0])
            self.seg[p] = (s0, int(idx[-1]) + 1)
            cnt = np.bincount(self.halo_color[idx], minlength=nc)
            start = s0
            for c in range(nc):
                self.seg_color[(p, c)] = (start, start + int(cnt[c]))
                start += int(cnt[c])

    def local_of_owned(self, gid):
        """local (colour-blocked) index of owned global rows"""
        loc = np.asarray(gid, dtype=np.int64) - self.o0
        if np.any((loc < 0) | (loc >= self.n_own)):
            raise ValueError("row not owned")
        return loc.astype(np.int32) if self.iperm is None else self.iperm[loc]

    def send_indices(self, receiver_plan):
        """local indices (into this rank's owned part) of the values `receiver_plan.rank` needs from this rank, in
        the receiver's halo order; plus per-colour offsets into that list"""
        r = receiver_plan
        if self.rank not in r.seg:
            return np.zeros(0, dtype=np.int32), np.zeros(max(r.ncolors, 1) + 1, dtype=np.int64)
        s, e = r.seg[self.rank]
        idx = self.local_of_owned(r.halo_gid[s:e])
        nc = max(r.ncolors, 1)
        ptr = np.zeros(nc + 1, dtype=np.int64)
        for c in range(nc):
            ptr[c] = r.seg_color[(self.rank, c)][0] - s
        ptr[nc] = e - s
        return idx, ptr

    def gather_indices(self):
        """global ids of the local vector layout [owned (colour-blocked) | halo]"""
        own = np.arange(self.o0, self.o1, dtype=np.int64)
        if self.perm is not None:
            own = own[self.perm]
        return np.concatenate([own, self.halo_gid])


def external_columns(indptr, indices, r0, r1, c0, c1):
    """sorted unique columns outside [c0,c1) referenced by rows [r0,r1) of a CSR matrix"""
    cols = np.asarray(indices[indptr[r0]:indptr[r1]], dtype=np.int64)
    ext = cols[(cols < c0) | (cols >= c1)]
    return np.unique(ext)


def level_external_columns(A, QT, Q_prev, offs, offs_next, offs_prev, rank):
    """The halo of a rank's level vector: the columns outside its row block that its rows of A_l (smoothing, residual),
    its rows of Q_l^T (restriction: rows = the coarse rows it owns, columns = this level) and its rows of Q_{l-1}
    (prolongation from this level: rows = the finer rows it owns) refer to.  CSR inputs in natural ordering; QT /
    Q_prev may be None.  distributed.DistributedHierarchy computes the same set on the device."""
    o0, o1 = int(offs[rank]), int(offs[rank + 1])
    parts = [external_columns(A.indptr, A.indices, o0, o1, o0, o1)]
    if QT is not None:
        c0, c1 = int(offs_next[rank]), int(offs_next[rank + 1])
        parts.append(external_columns(QT.indptr, QT.indices, c0, c1, o0, o1))
    if Q_prev is not None:
        f0, f1 = int(offs_prev[rank]), int(offs_prev[rank + 1])
        parts.append(external_columns(Q_prev.indptr, Q_prev.indices, f0, f1, o0, o1))
    return np.unique(np.concatenate(parts))
