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


# --------------------------------------------------------------------------------------
# sparse SDPA ("dat-s") files: the format of SDPLIB and of the reference's benchmark problems
# (``misc.sdpa_read`` / ``sdpa_readhead`` / ``sdpa_write``, ``src/C/misc.c:139-352``)
# --------------------------------------------------------------------------------------
def _sdpa_tokens(text):
    """Numbers of an SDPA body: everything that is not part of a number is a separator
    (the format allows ``{ } ( ) ,`` around the block structure and the vector b)."""
    import re
    return re.findall(r"[+-]?(?:\d+\.?\d*(?:[eEdD][+-]?\d+)?|\.\d+(?:[eEdD][+-]?\d+)?)", text)


def sdpa_read(fname, neg=False):
    """``A, b, blockstruct = sdpa_read(fname[, neg=False])`` (``misc.c:139-245``).

    Reads a sparse SDPA file into the CVXOPT-facing layout of SMCP: ``A`` is CCS of shape
    n^2 x (m+1), column k = vec of the lower triangle of F_k (entry (i, j), i <= j, of block
    ``blk`` goes to row ``(i-1+off)*n + (j-1+off)``), ``b`` the m-vector of the file and
    ``blockstruct`` the signed block sizes (negative = diagonal block).  ``neg=True`` negates
    all data, which turns the SDPA primal into SMCP's standard form (``base.py:177-195``).
    Unlike the reference the entries need not be sorted by matrix number."""
    with open(fname, "r") as fh:
        lines = fh.readlines()
    k = 0
    while k < len(lines) and (lines[k].lstrip()[:1] in ("*", '"') or not lines[k].strip()):
        k += 1
    m = int(_sdpa_tokens(lines[k])[0])
    nblocks = int(_sdpa_tokens(lines[k + 1])[0])
    tok = _sdpa_tokens("".join(lines[k + 2:]).replace("D", "e").replace("d", "e"))
    blockstruct = np.array([int(float(t)) for t in tok[:nblocks]], dtype=np.int64)
    boff = np.concatenate([[0], np.cumsum(np.abs(blockstruct))])
    n = int(boff[-1])
    b = np.array([float(t) for t in tok[nblocks:nblocks + m]], dtype=np.float64)
    body = np.array([float(t) for t in tok[nblocks + m:]], dtype=np.float64)
    body = body[:5 * (len(body) // 5)].reshape(-1, 5)
    mno = body[:, 0].astype(np.int64)
    bno = body[:, 1].astype(np.int64)
    ii = body[:, 2].astype(np.int64) + boff[bno - 1]
    jj = body[:, 3].astype(np.int64) + boff[bno - 1]
    v = body[:, 4]
    keep = v != 0
    rows = (ii[keep] - 1) * n + (jj[keep] - 1)
    sgn = -1.0 if neg else 1.0
    A = sp.csc_matrix((sgn * v[keep], (rows, mno[keep])), shape=(n * n, m + 1))
    return as_csc(A), sgn * b, blockstruct


def sdpa_readhead(fname):
    """``n, m, blockstruct = sdpa_readhead(fname)``: the header of a sparse SDPA file."""
    with open(fname, "r") as fh:
        lines = fh.readlines()
    k = 0
    while k < len(lines) and (lines[k].lstrip()[:1] in ("*", '"') or not lines[k].strip()):
        k += 1
    m = int(_sdpa_tokens(lines[k])[0])
    nblocks = int(_sdpa_tokens(lines[k + 1])[0])
    tok = _sdpa_tokens("".join(lines[k + 2:k + 4]))
    blockstruct = np.array([int(float(t)) for t in tok[:nblocks]], dtype=np.int64)
    return int(np.abs(blockstruct).sum()), m, blockstruct


def sdpa_write(fname, A, b, blockstruct, neg=False):
    """Writes ``(A, b, blockstruct)`` as a sparse SDPA file (``misc.c:281-352``): one line
    ``<matno> <blkno> <i> <j> <value>`` per stored lower-triangular entry, written as the upper
    triangular entry (j, i) the format asks for; 12 significant digits like the reference."""
    A = as_csc(A)
    b = np.asarray(b, dtype=np.float64).ravel()
    blockstruct = np.asarray(blockstruct, dtype=np.int64).ravel()
    n = int(np.abs(blockstruct).sum())
    boff = np.concatenate([[0], np.cumsum(np.abs(blockstruct))])
    sgn = -1.0 if neg else 1.0
    with open(fname, "w") as fh:
        fh.write("* sparse SDPA data file (created by smcp_b200)\n")
        fh.write("%i = m\n" % len(b))
        fh.write("%i = nBlocks\n" % len(blockstruct))
        fh.write(" ".join("%i" % s for s in blockstruct) + "\n")
        fh.write(" ".join("%.12g" % (sgn * x) for x in b) + "\n")
        for k in range(A.shape[1]):
            r = A.indices[A.indptr[k]:A.indptr[k + 1]]
            v = A.data[A.indptr[k]:A.indptr[k + 1]]
            Il, Jl = r % n, r // n                      # row >= column (lower triangle)
            if np.any(Jl > Il):
                raise ValueError("strictly upper triangular element in A")
            blk = np.searchsorted(boff, Jl, side="right")        # 1-based block of the column
            if np.any(Il >= boff[blk]):
                raise ValueError("matrix contains elements outside the blocks")
            for q in range(len(r)):
                if v[q] != 0.0:
                    fh.write("%i %i %i %i %.12g\n" % (k, blk[q], Jl[q] - boff[blk[q] - 1] + 1,
                                                      Il[q] - boff[blk[q] - 1] + 1, sgn * v[q]))
