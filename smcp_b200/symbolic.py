"""Host-side symbolic analysis of chordal sparsity patterns (NumPy, PyTorch-free).

This is the once-per-solve setup that precedes the GPU hot path (SURVEY.md §3.4,
reference call sites ``src/python/solvers.py:301-319`` / ``1542-1560``).  The reference
delegates it to ``chompack.symbolic`` / ``maxcardsearch`` / ``peo`` and
``cvxopt.amd.order`` (none of which are vendored); what is restated here is their
published behaviour:

* ``maxcardsearch``  – maximum-cardinality search; the *reverse* visiting order is a
  perfect elimination ordering iff the pattern is chordal (Tarjan & Yannakakis 1984).
* ``peo``            – zero-fill test of an ordering.
* ``embed``          – symbolic Cholesky (elimination tree + column structures) that
  produces the chordal embedding (filled pattern) under an ordering.
* ``min_degree``     – a minimum-degree ordering (stand-in for SuiteSparse AMD, whose
  tie-breaking cannot be observed here; both solvers accept an explicit ``p``).
* ``Symbolic``       – clique tree with *maximal* supernodes (Pothen & Sun: a parent joins
  the supernode of its first qualifying child in post-order, no amalgamation — the
  reference never passes a ``merge_function``), supernodal post-order, relative indices
  and the flat ``blkval`` storage layout used by every device kernel.

All index arrays are int64 on the host (``int_t = Py_ssize_t`` in ``src/C/cvxopt.h:46``);
the device copies are int32.
"""
from __future__ import annotations

import heapq
import os

import numpy as np

__all__ = ["maxcardsearch", "peo", "embed", "min_degree", "Symbolic", "lower_pattern",
           "task_partition"]


# --------------------------------------------------------------------------------------
# pattern helpers
# --------------------------------------------------------------------------------------
def lower_pattern(n, I, J):
    """Sorted, de-duplicated lower-triangular CCS pattern (with full diagonal) from
    coordinate lists.  Returns (colptr, rowind)."""
    I = np.asarray(I, dtype=np.int64)
    J = np.asarray(J, dtype=np.int64)
    r = np.maximum(I, J)
    c = np.minimum(I, J)
    d = np.arange(n, dtype=np.int64)
    r = np.concatenate([r, d])
    c = np.concatenate([c, d])
    key = np.unique(c * n + r)
    c = key // n
    r = key % n
    colptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(colptr, c + 1, 1)
    colptr = np.cumsum(colptr)
    return colptr, r


def _adjacency(n, colptr, rowind):
    """Symmetric adjacency (CSR-like: ptr, idx) without the diagonal."""
    cols = np.repeat(np.arange(n, dtype=np.int64), np.diff(colptr))
    rows = np.asarray(rowind, dtype=np.int64)
    off = rows != cols
    a = np.concatenate([rows[off], cols[off]])
    b = np.concatenate([cols[off], rows[off]])
    order = np.lexsort((b, a))
    a = a[order]
    b = b[order]
    ptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(ptr, a + 1, 1)
    ptr = np.cumsum(ptr)
    return ptr, b


def _py_maxcardsearch(n, colptr, rowind):
    """Maximum cardinality search on a lower-triangular pattern.

    Returns an ordering ``p`` (``p[k]`` = vertex eliminated k-th) that is a perfect
    elimination ordering whenever the pattern is chordal: the reverse of the MCS visiting
    order.  Ties are broken towards the largest vertex index so that a pattern already in
    PEO (band, arrow) keeps the identity ordering.
    (Replaces ``chompack.maxcardsearch``, call sites ``solvers.py:301, 1542``.)
    """
    ptr, idx = _adjacency(n, colptr, rowind)
    weight = np.zeros(n, dtype=np.int64)
    visited = np.zeros(n, dtype=bool)
    # bucket queue keyed by weight, max-heaps on vertex index inside a bucket
    heap = [(0, -v) for v in range(n)]
    heapq.heapify(heap)
    order = np.empty(n, dtype=np.int64)
    k = n - 1
    while heap:
        w, negv = heapq.heappop(heap)
        v = -negv
        if visited[v] or -w != weight[v]:
            continue
        visited[v] = True
        order[k] = v
        k -= 1
        for u in idx[ptr[v]:ptr[v + 1]]:
            if not visited[u]:
                weight[u] += 1
                heapq.heappush(heap, (-weight[u], -u))
    return order


def _perm_lower(n, colptr, rowind, p):
    """Lower-triangular pattern of P A Pᵀ, i.e. entry (i,j) of the result is A[p[i],p[j]]."""
    ip = np.empty(n, dtype=np.int64)
    ip[np.asarray(p, dtype=np.int64)] = np.arange(n, dtype=np.int64)
    cols = np.repeat(np.arange(n, dtype=np.int64), np.diff(colptr))
    rows = np.asarray(rowind, dtype=np.int64)
    return lower_pattern(n, ip[rows], ip[cols])


def _py_embed(n, colptr, rowind, p=None):
    """Symbolic Cholesky of the pattern under ordering ``p``.

    Returns ``(fcolptr, frowind, parent)``: the filled (chordal) lower pattern in the
    *permuted* index space, CCS with sorted rows and the diagonal first in each column,
    and the elimination tree.  (The embedding step of ``chompack.symbolic(Va, p)``,
    ``solvers.py:305-308``.)
    """
    if p is not None:
        colptr, rowind = _perm_lower(n, colptr, rowind, p)
    parent = np.full(n, -1, dtype=np.int64)
    children = [[] for _ in range(n)]
    struct = [None] * n
    for j in range(n):
        s = rowind[colptr[j]:colptr[j + 1]]
        s = s[s > j]
        ch = children[j]
        if ch:
            parts = [s]
            for c in ch:
                parts.append(struct[c][1:])  # struct[c][0] == j
            s = np.unique(np.concatenate(parts))
        struct[j] = s
        if len(s):
            parent[j] = s[0]
            children[s[0]].append(j)
    counts = np.array([len(s) + 1 for s in struct], dtype=np.int64)
    fcolptr = np.zeros(n + 1, dtype=np.int64)
    fcolptr[1:] = np.cumsum(counts)
    frowind = np.empty(fcolptr[-1], dtype=np.int64)
    for j in range(n):
        frowind[fcolptr[j]] = j
        frowind[fcolptr[j] + 1:fcolptr[j + 1]] = struct[j]
    return fcolptr, frowind, parent


def peo(n, colptr, rowind, p):
    """True iff ``p`` is a perfect elimination ordering of the pattern (zero fill).
    (Replaces ``chompack.peo``, ``solvers.py:302, 1543``.)"""
    pc, pr = _perm_lower(n, colptr, rowind, p)
    fc, fr, _ = embed(n, pc, pr)
    return int(fc[-1]) == int(pc[-1])



def _py_min_degree(n, colptr, rowind):
    """Minimum (external) degree ordering on the elimination graph with lazy heap updates.

    Stand-in for ``cvxopt.amd.order`` (``solvers.py:192-198, 278-279``).  Exact-degree
    elimination-graph algorithm; ties broken by vertex index, so it is deterministic.
    """
    ptr, idx = _adjacency(n, colptr, rowind)
    adj = [set(idx[ptr[v]:ptr[v + 1]].tolist()) for v in range(n)]
    deg = [len(a) for a in adj]
    heap = [(deg[v], v) for v in range(n)]
    heapq.heapify(heap)
    done = [False] * n
    order = []
    while heap:
        d, v = heapq.heappop(heap)
        if done[v] or d != deg[v]:
            continue
        done[v] = True
        order.append(v)
        nb = adj[v]
        for u in nb:
            au = adj[u]
            au.discard(v)
            au |= nb
            au.discard(u)
        for u in nb:
            du = len(adj[u])
            if du != deg[u]:
                deg[u] = du
            heapq.heappush(heap, (deg[u], u))
        adj[v] = set()
    return np.asarray(order, dtype=np.int64)


# --------------------------------------------------------------------------------------
# native twins (smcp_b200/csrc/host_symbolic.cpp): same tie-breaking, bit-identical results.
# The Python functions above are the specification; the drivers go through these dispatchers.
# --------------------------------------------------------------------------------------
def _native():
    """The shared library if it is built (host symbolic code needs no GPU), else None.
    ``SMCP_B200_NO_NATIVE_HOST=1`` forces the pure-NumPy specification (the CPU baseline of bench.py
    uses it so that no library of this repository is mapped in the reference arm)."""
    if os.environ.get("SMCP_B200_NO_NATIVE_HOST"):
        return None
    try:
        from . import device
        return device.load_library()
    except (RuntimeError, OSError, AttributeError):
        return None


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def min_degree(n, colptr, rowind):
    """Minimum-degree ordering (native ``smcp_host_min_degree``; see ``_py_min_degree``)."""
    lib = _native()
    if lib is None:
        return _py_min_degree(n, colptr, rowind)
    perm = np.empty(n, dtype=np.int64)
    if lib.smcp_host_min_degree(int(n), _i64(colptr), _i64(rowind), perm) != 0:
        raise RuntimeError("smcp_host_min_degree failed")
    return perm


def maxcardsearch(n, colptr, rowind):
    """Reverse MCS order (native ``smcp_host_maxcardsearch``; see ``_py_maxcardsearch``)."""
    lib = _native()
    if lib is None:
        return _py_maxcardsearch(n, colptr, rowind)
    order = np.empty(n, dtype=np.int64)
    if lib.smcp_host_maxcardsearch(int(n), _i64(colptr), _i64(rowind), order) != 0:
        raise RuntimeError("smcp_host_maxcardsearch failed")
    return order


def embed(n, colptr, rowind, p=None):
    """Chordal embedding under ordering ``p`` (native ``smcp_host_embed``; see ``_py_embed``)."""
    lib = _native()
    if lib is None:
        return _py_embed(n, colptr, rowind, p)
    if p is not None:
        colptr, rowind = _perm_lower(n, colptr, rowind, p)
    colptr, rowind = _i64(colptr), _i64(rowind)
    fcolptr = np.empty(n + 1, dtype=np.int64)
    parent = np.empty(n, dtype=np.int64)
    if lib.smcp_host_embed(int(n), colptr, rowind, fcolptr, None, None) != 0:
        raise RuntimeError("smcp_host_embed failed")
    frowind = np.empty(int(fcolptr[-1]), dtype=np.int64)
    if lib.smcp_host_embed(int(n), colptr, rowind, fcolptr, frowind.ctypes.data, parent.ctypes.data) != 0:
        raise RuntimeError("smcp_host_embed failed")
    return fcolptr, frowind, parent


# --------------------------------------------------------------------------------------
# clique tree / supernodal layout
# --------------------------------------------------------------------------------------
def _postorder(parent):
    """Post-order of a forest given by ``parent`` (-1 = root); children visited in
    increasing index order."""
    n = len(parent)
    head = [[] for _ in range(n)]
    roots = []
    for v in range(n):
        pv = parent[v]
        if pv < 0:
            roots.append(v)
        else:
            head[pv].append(v)
    post = np.empty(n, dtype=np.int64)
    k = 0
    for r in roots:
        stack = [(r, 0)]
        while stack:
            v, i = stack.pop()
            if i < len(head[v]):
                stack.append((v, i + 1))
                stack.append((head[v][i], 0))
            else:
                post[k] = v
                k += 1
    return post


def _py_aaidx(nsn, snpar, nn, na, nj, relptr, relidx, blkptr, updptr, nupd):
    """Specification of ``smcp_host_aaidx``: blkval offsets of the alpha x alpha entries of every
    supernode, column-major per supernode (top-down: parents have larger index)."""
    aaidx = np.empty(nupd, dtype=np.int64)
    for k in range(nsn - 1, -1, -1):
        a = int(na[k])
        if a == 0:
            continue
        pk = snpar[k]
        rel = relidx[relptr[k]:relptr[k + 1]]
        ri = np.repeat(rel[None, :], a, axis=0).T   # ri[i,j] = rel[i]
        rj = ri.T                                   # rj[i,j] = rel[j]
        lo_r = np.maximum(ri, rj)
        lo_c = np.minimum(ri, rj)
        out = np.empty((a, a), dtype=np.int64)
        nnp = int(nn[pk])
        njp = int(nj[pk])
        own = lo_c < nnp
        out[own] = blkptr[pk] + lo_c[own] * njp + lo_r[own]
        if not own.all():
            nap = int(na[pk])
            par = aaidx[updptr[pk]:updptr[pk + 1]].reshape(nap, nap)  # [col, row] storage
            rr = lo_r[~own] - nnp
            cc = lo_c[~own] - nnp
            out[~own] = par[cc, rr]
        # store column-major: element (i,j) at j*a+i
        aaidx[updptr[k]:updptr[k + 1]] = out.T.reshape(-1)
    return aaidx


def amalgamate(n, colptr, rowind, tol):
    """Relaxed supernodes (an OPT-IN performance knob: ``solvers.options['amalgamation']``):
    greedy post-order merging of a supernode into its parent whenever the explicit zeros this stores,
    ``nn_child * (nj_parent - na_child)``, are at most ``tol`` times the entries of the merged block.  Input
    and output are filled patterns in the SAME perfect elimination ordering; the output contains the input.
    A merged clique is ``nu_child + nu_parent + alpha_parent``, so on the band pattern (4995 blocks of 6 x 1)
    ``tol = 0.3`` gives ~1000 blocks of ~10 x 5.  The enlarged pattern changes the cone the barrier lives on,
    hence the iterates: the default (0) keeps the reference's pattern bit for bit."""
    if tol <= 0.0:
        return colptr, rowind
    symb = Symbolic(n, colptr, rowind)
    nsn = symb.nsn
    perm = symb.perm
    members = [list(perm[symb.snptr[k]:symb.snptr[k + 1]]) for k in range(nsn)]       # Vp indices
    alpha = [perm[symb.rowidx[symb.rowptr[k] + symb.nn[k]:symb.rowptr[k + 1]]] for k in range(nsn)]
    nnc = symb.nn.astype(np.int64).copy()
    njc = symb.nj.astype(np.int64).copy()
    zacc = np.zeros(nsn, dtype=np.int64)         # explicit zeros already stored in a (merged) supernode
    for k in range(nsn):                         # post-order: children first; a child may have absorbed its own children
        pk = int(symb.snpar[k])
        if pk < 0:
            continue
        na_k = njc[k] - nnc[k]
        zeros = zacc[k] + zacc[pk] + nnc[k] * (njc[pk] - na_k)
        if zeros <= tol * (nnc[k] + nnc[pk]) * (njc[pk] + nnc[k]):
            members[pk] = members[k] + members[pk]
            members[k] = None
            nnc[pk] += nnc[k]
            njc[pk] += nnc[k]
            zacc[pk] = zeros
    newcols = [None] * n
    for k in range(nsn):
        if members[k] is None:
            continue
        mem = np.sort(np.asarray(members[k], dtype=np.int64))
        al = np.sort(np.asarray(alpha[k], dtype=np.int64))
        for t, j in enumerate(mem):
            newcols[int(j)] = np.concatenate([mem[t:], al])
    ncolptr = np.zeros(n + 1, dtype=np.int64)
    ncolptr[1:] = np.cumsum([len(c) for c in newcols])
    return ncolptr, np.concatenate(newcols).astype(np.int64)


def _py_supernodes(n, colptr, rowind, parent):
    """NumPy specification of ``smcp_host_supernodes`` (csrc/host_symbolic.cpp): maximal supernodes with the
    Pothen-Sun first-qualifying-child rule, supernodes in post-order with contiguous columns, row lists
    (own columns, then the separator ascending) and the positions of every separator in its parent's row list."""
    colcount = np.diff(colptr)
    post = _postorder(parent)
    rep = np.full(n, -1, dtype=np.int64)      # representative (first vertex) of v's supernode
    for j in post:
        if rep[j] < 0:
            rep[j] = j
        pj = parent[j]
        if pj >= 0 and rep[pj] < 0 and colcount[j] - 1 == colcount[pj]:
            rep[pj] = rep[j]
    # chains: vertices of a supernode in path order (child -> parent); post-order visits
    # a path's vertices in that order
    members = {}
    for j in post:
        members.setdefault(int(rep[j]), []).append(int(j))
    reps = sorted(members.keys())
    sn_of_rep = {r: k for k, r in enumerate(reps)}
    nsn = len(reps)
    snpar0 = np.full(nsn, -1, dtype=np.int64)
    for k, r in enumerate(reps):
        top = members[r][-1]
        if parent[top] >= 0:
            snpar0[k] = sn_of_rep[int(rep[parent[top]])]
    snpost = _postorder(snpar0)
    renum = np.empty(nsn, dtype=np.int64)
    renum[snpost] = np.arange(nsn, dtype=np.int64)
    snptr = np.zeros(nsn + 1, dtype=np.int64)
    perm = np.empty(n, dtype=np.int64)        # internal index -> Vp index
    k0 = 0
    for knew, kold in enumerate(snpost):
        mem = members[reps[kold]]
        perm[k0:k0 + len(mem)] = mem
        k0 += len(mem)
        snptr[knew + 1] = k0
    iperm = np.empty(n, dtype=np.int64)
    iperm[perm] = np.arange(n, dtype=np.int64)
    snpar = np.full(nsn, -1, dtype=np.int64)
    snpar[renum] = np.where(snpar0 >= 0, renum[np.maximum(snpar0, 0)], -1)
    nn = np.diff(snptr)
    # separator = structure of the last (top) vertex of the chain
    top = perm[snptr[1:] - 1] if nsn else np.zeros(0, dtype=np.int64)
    na = colcount[top] - 1
    nj = nn + na
    rowptr = np.zeros(nsn + 1, dtype=np.int64)
    rowptr[1:] = np.cumsum(nj)
    rowidx = np.empty(rowptr[-1], dtype=np.int64)
    for k in range(nsn):
        r0 = rowptr[k]
        rowidx[r0:r0 + nn[k]] = np.arange(snptr[k], snptr[k + 1])
        t = top[k]
        a = iperm[rowind[colptr[t] + 1:colptr[t + 1]]]
        a.sort()
        rowidx[r0 + nn[k]:rowptr[k + 1]] = a
    # relative indices: alpha_k inside gamma_parent
    relptr = np.zeros(nsn + 1, dtype=np.int64)
    relptr[1:] = np.cumsum(na)
    relidx = np.empty(relptr[-1], dtype=np.int64)
    for k in range(nsn):
        pk = snpar[k]
        if pk < 0:
            continue
        a = rowidx[rowptr[k] + nn[k]:rowptr[k + 1]]
        g = rowidx[rowptr[pk]:rowptr[pk + 1]]
        pos = np.searchsorted(g, a)
        if np.any(pos >= len(g)) or np.any(g[np.minimum(pos, len(g) - 1)] != a):  # pragma: no cover
            raise ValueError("pattern is not chordal in the given ordering")
        relidx[relptr[k]:relptr[k + 1]] = pos
    return nsn, perm, snptr, snpar, rowptr, rowidx, relptr, relidx


class Symbolic:
    """Supernodal clique tree of a *chordal* lower-triangular pattern given in a perfect
    elimination ordering (the ``symb = symbolic(Vp)`` object of ``solvers.py:314, 1555``).

    Input: ``(n, colptr, rowind)`` of the filled pattern ``Vp`` (CCS, rows sorted,
    diagonal present).  The object fixes

    * the supernode partition (maximal supernodes, Pothen–Sun first-qualifying-child rule),
    * an internal post-ordered relabelling ``iperm``/``perm`` under which every supernode
      is a contiguous column range and supernodes are numbered in post-order,
    * the flat value layout ``blkval``: supernode k owns a dense column-major block of
      shape ``(nn_k+na_k) x nn_k`` at ``blkptr[k]`` whose rows are ``rowidx[rowptr[k]:
      rowptr[k+1]]`` (own columns first, then the separator ``alpha_k`` ascending),
    * ``relidx``: positions of ``alpha_k`` inside the parent's row list,
    * ``aaidx``: for every supernode the ``blkval`` offsets of the ``alpha x alpha``
      entries (lower triangle, column-major ``na x na``; strictly-upper slots mirror the
      lower ones) — lets any kernel gather ``X_{alpha alpha}`` without walking the tree,
    * ``updptr``: offsets of the per-supernode ``na x na`` update-matrix workspace,
    * ``vec2blk``: ``blkval`` offset of the q-th non-zero of ``Vp`` in CCS order (the
      "vector space" of ``Av`` rows, ``solvers.py:318-319``), and ``wdot`` the trace
      inner-product weights (2 strict lower, 1 diagonal, 0 padding).
    """

    def __init__(self, n, colptr, rowind):
        n = int(n)
        colptr = np.asarray(colptr, dtype=np.int64)
        rowind = np.asarray(rowind, dtype=np.int64)
        self.n = n
        self.colptr = colptr
        self.rowind = rowind
        self.nvp = int(colptr[-1])
        colcount = np.diff(colptr)
        # elimination tree of a filled pattern: parent = first sub-diagonal row
        parent = np.full(n, -1, dtype=np.int64)
        has = colcount > 1
        parent[has] = rowind[colptr[:-1][has] + 1]
        self.parent = parent

        # supernode partition, post-ordered relabelling, row lists, relative indices: native when the
        # library is built (csrc/host_symbolic.cpp: smcp_host_supernodes), else the NumPy specification
        lib = _native()
        if lib is not None:
            import ctypes
            nsn_c = ctypes.c_int64(0)
            perm = np.empty(n, dtype=np.int64)
            snptr = np.empty(n + 1, dtype=np.int64)
            snpar = np.empty(max(n, 1), dtype=np.int64)
            rowptr = np.empty(n + 1, dtype=np.int64)
            rowidx = np.empty(n + self.nvp, dtype=np.int64)
            relptr = np.empty(n + 1, dtype=np.int64)
            relidx = np.empty(max(self.nvp, 1), dtype=np.int64)
            rc = lib.smcp_host_supernodes(n, _i64(colptr), _i64(rowind), ctypes.byref(nsn_c), perm, snptr, snpar, rowptr,
                                          rowidx, relptr, relidx)
            if rc == -2:
                raise ValueError("pattern is not chordal in the given ordering")
            if rc != 0:
                raise RuntimeError("smcp_host_supernodes failed")
            nsn = int(nsn_c.value)
            snptr, snpar, rowptr, relptr = snptr[:nsn + 1].copy(), snpar[:nsn].copy(), rowptr[:nsn + 1].copy(), relptr[:nsn + 1].copy()
            rowidx, relidx = rowidx[:rowptr[-1]].copy(), relidx[:relptr[-1]].copy()
        else:
            nsn, perm, snptr, snpar, rowptr, rowidx, relptr, relidx = _py_supernodes(n, colptr, rowind, parent)
        self.nsn = nsn
        iperm = np.empty(n, dtype=np.int64)
        iperm[perm] = np.arange(n, dtype=np.int64)
        self.perm, self.iperm, self.snptr, self.snpar = perm, iperm, snptr, snpar
        nn = np.diff(snptr)
        nj = np.diff(rowptr)
        na = nj - nn
        self.nn, self.na, self.nj = nn, na, nj
        self.rowptr, self.rowidx, self.relptr, self.relidx = rowptr, rowidx, relptr, relidx
        blkptr = np.zeros(nsn + 1, dtype=np.int64)
        blkptr[1:] = np.cumsum(nj * nn)
        self.blkptr = blkptr
        self.nblk = int(blkptr[-1])
        updptr = np.zeros(nsn + 1, dtype=np.int64)
        updptr[1:] = np.cumsum(na * na)
        self.updptr = updptr
        self.nupd = int(updptr[-1])

        # children lists (ascending = post-order among siblings)
        chptr = np.zeros(nsn + 1, dtype=np.int64)
        haspar = snpar >= 0
        np.add.at(chptr, snpar[haspar] + 1, 1)
        chptr = np.cumsum(chptr)
        kids = np.nonzero(haspar)[0]
        chidx = kids[np.argsort(snpar[kids], kind="stable")].astype(np.int64)
        self.chptr = chptr
        self.chidx = chidx

        # height (leaves 0) and depth (roots 0) levels
        height = np.zeros(nsn, dtype=np.int64)
        for k in range(nsn):
            pk = snpar[k]
            if pk >= 0 and height[pk] < height[k] + 1:
                height[pk] = height[k] + 1
        depth = np.zeros(nsn, dtype=np.int64)
        for k in range(nsn - 1, -1, -1):
            pk = snpar[k]
            if pk >= 0:
                depth[k] = depth[pk] + 1
        self.height = height
        self.depth = depth

        # alpha x alpha gather map (top-down: parents have larger index)
        lib = _native()
        if lib is not None:
            aaidx = np.empty(self.nupd, dtype=np.int64)
            if lib.smcp_host_aaidx(nsn, _i64(snpar), _i64(nn), _i64(nj), _i64(relptr), _i64(relidx), _i64(blkptr),
                                   _i64(updptr), aaidx) != 0:
                raise RuntimeError("smcp_host_aaidx failed")
        else:
            aaidx = _py_aaidx(nsn, snpar, nn, na, nj, relptr, relidx, blkptr, updptr, self.nupd)
        self.aaidx = aaidx

        # vector space <-> blkval
        cols = np.repeat(np.arange(n, dtype=np.int64), colcount)
        ir = iperm[rowind]
        ic = iperm[cols]
        r = np.maximum(ir, ic)
        c = np.minimum(ir, ic)
        sn_of_col = np.repeat(np.arange(nsn, dtype=np.int64), nn)
        kc = sn_of_col[c]
        lc = c - snptr[kc]
        # row position of r in supernode kc's row list
        # own rows: r - snptr[kc] if r < snptr[kc+1] else search in alpha
        pos = np.empty(self.nvp, dtype=np.int64)
        own = r < snptr[kc + 1]
        pos[own] = r[own] - snptr[kc[own]]
        if not own.all():
            idx = np.nonzero(~own)[0]
            # search per supernode (vectorised through a global key)
            key_rows = rowidx + np.repeat(np.arange(nsn, dtype=np.int64), nj) * n
            q = np.searchsorted(key_rows, r[idx] + kc[idx] * n)
            if np.any(key_rows[q] != r[idx] + kc[idx] * n):  # pragma: no cover
                raise ValueError("pattern entry outside the clique tree (not chordal?)")
            pos[idx] = q - rowptr[kc[idx]]
        self.vec2blk = blkptr[kc] + lc * nj[kc] + pos
        self.Ip = rowind.copy()
        self.Jp = cols
        self.diag_vec = np.nonzero(rowind == cols)[0]          # "Id" of solvers.py:349
        wdot = np.zeros(self.nblk, dtype=np.float64)
        wdot[self.vec2blk] = 2.0
        wdot[self.vec2blk[self.diag_vec]] = 1.0
        self.wdot = wdot
        # blkval offset of the diagonal entry of internal column i
        sn_i = sn_of_col
        li = np.arange(n, dtype=np.int64) - snptr[sn_i]
        self.diag_blk = blkptr[sn_i] + li * nj[sn_i] + li
        if self.nvp != int(np.count_nonzero(wdot)):
            raise ValueError("pattern is not chordal in the given ordering")  # pragma: no cover
        if int(np.sum(nn * (nn + 1) // 2 + na * nn)) != self.nvp:
            raise ValueError("pattern is not chordal in the given ordering (fill needed)")

    # ------------------------------------------------------------------
    def tasks(self, max_tasks_work=None):
        """Partition of the clique tree into sequential *tasks* for the persistent
        dependency-driven kernels: a task is a connected piece of the tree given as a
        list of supernodes in post-order; its dependencies are the tasks holding the
        children of its supernodes that are outside the piece.

        Greedy rule: walk supernodes in post-order, a supernode is merged into the task
        of its children when it has exactly one child task to wait for ... (see
        ``task_partition`` for the parametrised version used by the device layer).
        """
        return task_partition(self, max_tasks_work)

    def summary(self):
        return dict(n=self.n, nvp=self.nvp, nsn=self.nsn, nblk=self.nblk, nupd=self.nupd,
                    max_nn=int(self.nn.max()), max_na=int(self.na.max()),
                    max_nj=int(self.nj.max()), height=int(self.height.max()))


def task_partition(symb, small_work=None):
    """Group supernodes into tasks (see ``Symbolic.tasks``).

    A supernode joins the task of its *last* child (the one immediately preceding it in
    post-order) when (a) it has a single child, or (b) the accumulated work of the whole
    subtree below it is at most ``small_work`` flops (so small subtrees are walked by one
    CTA without any inter-CTA hand-off).  Returns ``(task_ptr, task_sn, dep_ptr, dep_idx,
    task_of_sn)`` with tasks numbered in a topological (post-order) sequence.
    """
    nsn = symb.nsn
    nn, na = symb.nn, symb.na
    work = (nn ** 3) // 3 + na * nn * nn + na * na * nn + 1
    if small_work is None:
        small_work = 0
    sub = work.astype(np.int64).copy()
    nchild = np.diff(symb.chptr)
    for k in range(nsn):
        pk = symb.snpar[k]
        if pk >= 0:
            sub[pk] += sub[k]
    task_of = np.full(nsn, -1, dtype=np.int64)
    # a subtree that is "small" becomes one task rooted at the highest small ancestor
    small = sub <= small_work
    ntask = 0
    tasks = []
    for k in range(nsn):
        ch = symb.chidx[symb.chptr[k]:symb.chptr[k + 1]]
        if len(ch) and small[k]:
            # merge all child tasks into one (children subtrees are contiguous in post-order)
            t0 = task_of[ch[0]]
            for c in ch[1:]:
                tc = task_of[c]
                if tc != t0:
                    tasks[t0].extend(tasks[tc])
                    for s in tasks[tc]:
                        task_of[s] = t0
                    tasks[tc] = None
            tasks[t0].append(k)
            task_of[k] = t0
        elif len(ch) == 1:
            t0 = task_of[ch[0]]
            tasks[t0].append(k)
            task_of[k] = t0
        else:
            tasks.append([k])
            task_of[k] = len(tasks) - 1
    # renumber surviving tasks by their last supernode (topological)
    alive = [t for t in range(len(tasks)) if tasks[t] is not None]
    alive.sort(key=lambda t: tasks[t][-1])
    renum = {t: i for i, t in enumerate(alive)}
    T = len(alive)
    task_ptr = np.zeros(T + 1, dtype=np.int64)
    task_sn = np.empty(nsn, dtype=np.int64)
    for i, t in enumerate(alive):
        lst = sorted(tasks[t])
        task_ptr[i + 1] = task_ptr[i] + len(lst)
        task_sn[task_ptr[i]:task_ptr[i + 1]] = lst
    task_of = np.array([renum[int(t)] for t in task_of], dtype=np.int64)
    deps = [set() for _ in range(T)]
    for k in range(nsn):
        pk = symb.snpar[k]
        if pk >= 0 and task_of[pk] != task_of[k]:
            deps[task_of[pk]].add(int(task_of[k]))
    dep_ptr = np.zeros(T + 1, dtype=np.int64)
    dep_idx = []
    for i in range(T):
        d = sorted(deps[i])
        dep_idx.extend(d)
        dep_ptr[i + 1] = len(dep_idx)
    return task_ptr, task_sn, dep_ptr, np.asarray(dep_idx, dtype=np.int64), task_of
