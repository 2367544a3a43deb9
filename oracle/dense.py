"""Dense NumPy ground truth for the chordal kernels (TEST INFRASTRUCTURE; see
``oracle/__init__.py``).  Everything is formed as dense n x n matrices in the internal
(post-ordered) index space of a ``Symbolic`` and projected back on the pattern, so it is
only usable for n up to a few hundred.  It shares *no* arithmetic with
``oracle.supernodal`` or the CUDA kernels — only the storage maps.
"""
from __future__ import annotations

import numpy as np


def _block_coords(symb):
    """(rows, cols, offsets) of every lower-triangular pattern entry in blkval."""
    rows, cols, offs = [], [], []
    for k in range(symb.nsn):
        nn, nj = int(symb.nn[k]), int(symb.nj[k])
        r = symb.rowidx[symb.rowptr[k]:symb.rowptr[k + 1]]
        for j in range(nn):
            rr = r[j:]
            rows.append(rr)
            cols.append(np.full(len(rr), symb.snptr[k] + j))
            offs.append(symb.blkptr[k] + j * nj + np.arange(j, nj))
    return np.concatenate(rows), np.concatenate(cols), np.concatenate(offs)


def to_dense(symb, x, symmetric=True):
    """Dense matrix (internal order) of a chordal matrix / factor stored in blkval."""
    r, c, o = _block_coords(symb)
    M = np.zeros((symb.n, symb.n))
    M[r, c] = np.ravel(x)[o]
    if symmetric:
        M = M + np.tril(M, -1).T
    return M


def project(symb, M):
    """blkval array holding the pattern entries (lower triangle) of dense M."""
    r, c, o = _block_coords(symb)
    x = np.zeros(symb.nblk)
    x[o] = M[r, c]
    return x


def cholesky(symb, x):
    S = to_dense(symb, x)
    try:
        L = np.linalg.cholesky(S)
    except np.linalg.LinAlgError:
        raise ArithmeticError("not positive definite")
    return project(symb, L)


def llt(symb, l):
    L = to_dense(symb, l, symmetric=False)
    return project(symb, L @ L.T)


def projected_inverse(symb, l):
    L = to_dense(symb, l, symmetric=False)
    return project(symb, np.linalg.inv(L @ L.T))


def completion_residual(symb, l, x):
    """max |P((L L^T)^{-1}) - X| — the defining property of ``completion``."""
    return float(np.max(np.abs(projected_inverse(symb, l) - project(symb, to_dense(symb, x)))))


def hessian(symb, l, u):
    """P(S^{-1} U S^{-1}), S = L L^T."""
    L = to_dense(symb, l, symmetric=False)
    Sinv = np.linalg.inv(L @ L.T)
    return project(symb, Sinv @ to_dense(symb, u) @ Sinv)


def dot(symb, x, y):
    return float(np.sum(to_dense(symb, x) * to_dense(symb, y)))


def schur(symb, l, Us):
    """H_ij = tr(U_i S^{-1} U_j S^{-1}) for a list of chordal matrices U."""
    L = to_dense(symb, l, symmetric=False)
    Sinv = np.linalg.inv(L @ L.T)
    D = [to_dense(symb, u) for u in Us]
    W = [Sinv @ d @ Sinv for d in D]
    m = len(D)
    H = np.empty((m, m))
    for i in range(m):
        for j in range(m):
            H[i, j] = np.sum(D[i] * W[j])
    return H
