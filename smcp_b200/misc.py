"""Host index helpers and problem builders (NumPy restatement of the hot members of the
reference's C extension ``smcp.misc``, ``src/C/misc.c:1057-1102``).

Sparse matrices are ``scipy.sparse.csc_matrix`` with int64 indices — the CCS layout of a
cvxopt ``spmatrix`` (``src/C/cvxopt.h:48-69``).  The per-iteration members of ``misc``
(``Av_to_spmatrix``, ``scal_diag``, ``SCMcolumn2``) have no host counterpart here: they are
replaced by device kernels behind ``include/smcp_b200.h``.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def ind2sub(n, ind):
    """(I, J) with I = ind % n, J = ind // n (``misc.c:387-408``)."""
    ind = np.asarray(ind, dtype=np.int64)
    return ind % n, ind // n


def sub2ind(siz, I, J):
    """Linear index I + m*J for a matrix of size (m, n) (``misc.c:428-445``)."""
    m = int(siz[0])
    return np.asarray(I, dtype=np.int64) + m * np.asarray(J, dtype=np.int64)


def as_csc(A):
    """Canonical CCS (sorted row indices, no duplicates, int64)."""
    A = sp.csc_matrix(A)
    A.sum_duplicates()
    A.sort_indices()
    A.indptr = A.indptr.astype(np.int64)
    A.indices = A.indices.astype(np.int64)
    return A


def nzcolumns(A):
    """Number of non-zero columns (= rows) touched by each A_i, i = 1..m
    (``misc.c:682-730``): distinct values of {r % n} U {r // n} over the stored entries of
    column i of A."""
    A = as_csc(A)
    n = int(round(np.sqrt(A.shape[0])))
    m = A.shape[1] - 1
    Nz = np.zeros(m, dtype=np.int64)
    colptr, rows = A.indptr, A.indices
    col_of = np.repeat(np.arange(m + 1, dtype=np.int64), np.diff(colptr))
    sel = col_of >= 1
    if not np.any(sel):
        return Nz
    c = col_of[sel] - 1
    r = rows[sel]
    keys = np.concatenate([c * n + r % n, c * n + r // n])
    keys = np.unique(keys)
    np.add.at(Nz, keys // n, 1)
    return Nz


def matperm(nzc, Nmax):
    """Constraint permutation (``misc.c:750-773``): constraints with more than ``Nmax``
    non-zero columns first (original order), the ``Ns`` others last in reverse order."""
    nzc = np.asarray(nzc, dtype=np.int64)
    m = len(nzc)
    Nmax = max(int(Nmax), 0)
    dense = np.nonzero(nzc > Nmax)[0]
    sparse_ = np.nonzero(nzc <= Nmax)[0]
    pm = np.empty(m, dtype=np.int64)
    pm[:len(dense)] = dense
    pm[len(dense):] = sparse_[::-1]
    return pm, int(len(sparse_))


def phase1_sdp(A, u):
    """Phase-I problem data of order n+2 with m+1 constraints (``misc.c:1004-1054``):
    objective e_n e_n^T; A_i' = A_i (re-strided) - u_i e_n e_n^T; last constraint
    I_n (+) 0 (+) 1."""
    A = as_csc(A)
    n = int(round(np.sqrt(A.shape[0])))
    m = A.shape[1] - 1
    u = np.asarray(u, dtype=np.float64).ravel()
    n2 = n + 2
    rows, cols, vals = [np.array([n * n2 + n])], [np.array([0])], [np.array([1.0])]
    colptr = A.indptr
    for i in range(1, m + 1):
        r = A.indices[colptr[i]:colptr[i + 1]]
        v = A.data[colptr[i]:colptr[i + 1]]
        rows.append(np.concatenate([r + 2 * (r // n), [n * n2 + n]]))
        vals.append(np.concatenate([v, [-u[i - 1]]]))
        cols.append(np.full(len(r) + 1, i))
    d = np.arange(n, dtype=np.int64)
    rows.append(np.concatenate([d * n2 + d, [n2 * n2 - 1]]))
    vals.append(np.ones(n + 1))
    cols.append(np.full(n + 1, m + 1))
    return as_csc(sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                                shape=(n2 * n2, m + 2)))
