"""Supernodal CPU restatement of the chordal-matrix kernels (TEST INFRASTRUCTURE, see
``oracle/__init__.py``; parity unpinned — chompack is not vendored).

Each routine restates the multifrontal recursion SMCP obtains from ``chompack`` at the
call sites listed below (SURVEY.md Appendix A gives the block formulas; they follow
Andersen/Dahl/Vandenberghe 2010, the paper cited at ``doc/source/index.rst:12-16``) with
NumPy/SciPy dense BLAS/LAPACK on per-supernode blocks — the same kind of call sequence
chompack issues, which makes it the timed CPU baseline as well.

All routines work on value arrays of shape ``(B, nblk)`` (a batch of B chordal matrices on
one pattern, ``blkval`` layout of ``smcp_b200.symbolic.Symbolic``) and modify them in
place, mirroring chompack's in-place semantics (``solvers.py:873-874``: callers copy first).

Notation per supernode k: nu = own columns (nn), alpha = separator rows (na),
block = [X_nunu ; X_alphanu] (nj x nn), Lt = L_alphanu L_nunu^{-1}.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sl


def _as2d(x):
    x = np.asarray(x)
    return x.reshape(1, -1) if x.ndim == 1 else x


def _blk(symb, X, k):
    """(B, nj, nn) view of supernode k's block (column-major nj x nn in blkval)."""
    nn, nj = int(symb.nn[k]), int(symb.nj[k])
    b0 = int(symb.blkptr[k])
    return X[:, b0:b0 + nn * nj].reshape(X.shape[0], nn, nj).transpose(0, 2, 1)


def _sym(a):
    """Full symmetric matrix from the lower triangle of the trailing two dims."""
    lo = np.tril(a)
    return lo + np.tril(a, -1).swapaxes(-1, -2)


def _rel(symb, k):
    return symb.relidx[symb.relptr[k]:symb.relptr[k + 1]]


def _children(symb, k):
    return symb.chidx[symb.chptr[k]:symb.chptr[k + 1]]


def _frontal(symb, X, k, upd):
    """Full symmetric frontal matrix [X_nn X_an^T; X_an 0] + extend-add of the children's
    update matrices (App. A.1 / A.4 pass 1)."""
    nn, nj = int(symb.nn[k]), int(symb.nj[k])
    B = X.shape[0]
    blk = _blk(symb, X, k)
    F = np.zeros((B, nj, nj))
    F[:, :nn, :nn] = _sym(blk[:, :nn, :])
    F[:, nn:, :nn] = blk[:, nn:, :]
    F[:, :nn, nn:] = blk[:, nn:, :].swapaxes(1, 2)
    for c in _children(symb, k):
        r = _rel(symb, c)
        F[:, r[:, None], r[None, :]] += upd[c]
        upd[c] = None
    return F


def _gather_aa(symb, X, k):
    """X_{alpha alpha} (B, na, na), full symmetric, gathered through ``aaidx``."""
    na = int(symb.na[k])
    idx = symb.aaidx[symb.updptr[k]:symb.updptr[k + 1]]
    return X[:, idx].reshape(X.shape[0], na, na).swapaxes(1, 2)


def _chol(a):
    """Batched lower Cholesky with dpotrf's failure rule (pivot <= 0 or NaN)."""
    try:
        L = np.linalg.cholesky(a)
    except np.linalg.LinAlgError:
        raise ArithmeticError("matrix is not positive definite")
    if not np.all(np.isfinite(L)):
        raise ArithmeticError("matrix is not positive definite")
    return L


def _write_lower(blk, nn, Lnn, Lan):
    blk[:, :nn, :] = np.tril(Lnn)
    blk[:, nn:, :] = Lan


# --------------------------------------------------------------------------------------
def cholesky(symb, X):
    """X <- L with L L^T = X (App. A.1; chompack.cholesky, e.g. ``solvers.py:884, 2354``).
    Raises ArithmeticError if X is not positive definite."""
    X = _as2d(X)
    upd = [None] * symb.nsn
    for k in range(symb.nsn):
        nn = int(symb.nn[k])
        F = _frontal(symb, X, k, upd)
        Lnn = _chol(F[:, :nn, :nn])
        if symb.na[k]:
            # L_an = F_an L_nn^{-T}
            Lan = sl.solve_triangular(Lnn, F[:, nn:, :nn].swapaxes(1, 2), lower=True).swapaxes(1, 2)
            upd[k] = F[:, nn:, nn:] - Lan @ Lan.swapaxes(1, 2)
        else:
            Lan = F[:, nn:, :nn]
        _write_lower(_blk(symb, X, k), nn, Lnn, Lan)


def llt(symb, L):
    """L <- P_V(L L^T) (App. A.6; chompack.llt, ``solvers.py:904, 1721``)."""
    L = _as2d(L)
    upd = [None] * symb.nsn
    for k in range(symb.nsn):
        nn = int(symb.nn[k])
        blk = _blk(symb, L, k)
        Lf = blk.copy()
        Lf[:, :nn, :] = np.tril(Lf[:, :nn, :])
        P = Lf @ Lf.swapaxes(1, 2)                     # (nj, nj) outer product
        for c in _children(symb, k):
            r = _rel(symb, c)
            P[:, r[:, None], r[None, :]] += upd[c]
            upd[c] = None
        if symb.na[k]:
            upd[k] = P[:, nn:, nn:]
        _write_lower(blk, nn, P[:, :nn, :nn], P[:, nn:, :nn])


def projected_inverse(symb, L):
    """L <- Y = P_V((L L^T)^{-1}) (App. A.2; chompack.projected_inverse,
    ``solvers.py:891, 2361``).  Root to leaves; Y_aa is gathered from ancestors' output."""
    L = _as2d(L)
    B = L.shape[0]
    for k in range(symb.nsn - 1, -1, -1):
        nn, na = int(symb.nn[k]), int(symb.na[k])
        blk = _blk(symb, L, k)
        Lnn = np.tril(blk[:, :nn, :])
        eye = np.broadcast_to(np.eye(nn), (B, nn, nn))
        Linv = sl.solve_triangular(Lnn, eye, lower=True)
        Dinv = Linv.swapaxes(1, 2) @ Linv
        if na:
            Lt = sl.solve_triangular(Lnn, blk[:, nn:, :].swapaxes(1, 2), lower=True, trans='T').swapaxes(1, 2)
            # Lt = L_an L_nn^{-1}
            Yaa = _gather_aa(symb, L, k)
            Yan = -Yaa @ Lt
            Ynn = Dinv - Lt.swapaxes(1, 2) @ Yan
        else:
            Yan = blk[:, nn:, :]
            Ynn = Dinv
        _write_lower(blk, nn, 0.5 * (Ynn + Ynn.swapaxes(1, 2)), Yan)


def completion(symb, X):
    """X <- L with P_V((L L^T)^{-1}) = X (App. A.3; chompack.completion, e.g.
    ``solvers.py:874, 2344``).  Raises ArithmeticError if X has no positive definite
    completion.  Every supernode only needs original entries of X, so the result is
    assembled out of place."""
    X = _as2d(X)
    out = np.zeros_like(X)
    for k in range(symb.nsn):
        nn, na = int(symb.nn[k]), int(symb.na[k])
        blk = _blk(symb, X, k)
        Xnn = _sym(blk[:, :nn, :])
        if na:
            Xaa = _gather_aa(symb, X, k)
            R = _chol(Xaa)
            Xan = blk[:, nn:, :]
            Z = sl.solve_triangular(R, Xan, lower=True)                  # R^{-1} X_an
            Delta = Xnn - Z.swapaxes(1, 2) @ Z
            W = sl.solve_triangular(R, Z, lower=True, trans='T')         # X_aa^{-1} X_an
        else:
            Delta = Xnn
        # L_nn lower with L_nn L_nn^T = Delta^{-1}:  Delta = U U^T (U upper), L_nn = U^{-T}
        Rf = _chol(Delta[:, ::-1, ::-1])
        U = Rf[:, ::-1, ::-1]                                            # upper, U U^T = Delta
        eye = np.broadcast_to(np.eye(nn), Delta.shape)
        Lnn = sl.solve_triangular(U, eye, lower=False).swapaxes(1, 2)    # U^{-T}, lower
        Lan = -W @ Lnn if na else blk[:, nn:, :]
        _write_lower(_blk(symb, out, k), nn, Lnn, Lan)
    X[...] = out


# --------------------------------------------------------------------------------------
class HessianFactor:
    """Per-(L, Y) data shared by all Hessian evaluations at one scaling point:
    L_nn, Lt = L_an L_nn^{-1}, Y_aa and (lazily) its Cholesky factor."""

    def __init__(self, symb, L, Y):
        L = _as2d(L)
        Y = _as2d(Y)
        assert L.shape[0] == 1 and Y.shape[0] == 1
        self.symb = symb
        self.Lnn, self.Lt, self.Yaa, self.Raa = [], [], [], [None] * symb.nsn
        for k in range(symb.nsn):
            nn, na = int(symb.nn[k]), int(symb.na[k])
            blk = _blk(symb, L, k)[0]
            Lnn = np.tril(blk[:nn, :])
            self.Lnn.append(Lnn)
            if na:
                self.Lt.append(sl.solve_triangular(Lnn, blk[nn:, :].T, lower=True, trans='T').T)
                self.Yaa.append(_gather_aa(symb, Y, k)[0].copy())
            else:
                self.Lt.append(np.zeros((0, nn)))
                self.Yaa.append(np.zeros((0, 0)))

    def chol_Yaa(self, k):
        if self.Raa[k] is None:
            self.Raa[k] = _chol(self.Yaa[k])
        return self.Raa[k]


def _dsolve2(Lnn, K):
    """D^{-1} K D^{-1} with D = Lnn Lnn^T (K batched (B, nn, nn))."""
    T = sl.cho_solve((Lnn, True), K)
    return sl.cho_solve((Lnn, True), T.swapaxes(1, 2)).swapaxes(1, 2)


def hessian(hf, U):
    """U <- P_V(S^{-1} U S^{-1}), S = L L^T (App. A.4; ``hessian(L,Y,U,inv=False,adj=None)``
    e.g. ``solvers.py:483, 524, 531, 1913, 1952, 1959``)."""
    symb = hf.symb
    U = _as2d(U)
    upd = [None] * symb.nsn
    # pass 1 (post-order) fused with the per-supernode scaling
    for k in range(symb.nsn):
        nn, na = int(symb.nn[k]), int(symb.na[k])
        F = _frontal(symb, U, k, upd)
        Lnn, Lt = hf.Lnn[k], hf.Lt[k]
        Knn = F[:, :nn, :nn]
        if na:
            Kan = F[:, nn:, :nn] - Lt @ Knn
            upd[k] = F[:, nn:, nn:] - Lt @ F[:, :nn, nn:] - Kan @ Lt.T
            D = Lnn @ Lnn.T
            Man = hf.Yaa[k] @ sl.cho_solve((Lnn, True), Kan.swapaxes(1, 2)).swapaxes(1, 2)
        else:
            Man = F[:, nn:, :nn]
        Mnn = _dsolve2(Lnn, Knn)
        _write_lower(_blk(symb, U, k), nn, 0.5 * (Mnn + Mnn.swapaxes(1, 2)), Man)
    # pass 3 (reverse post-order)
    for k in range(symb.nsn - 1, -1, -1):
        nn, na = int(symb.nn[k]), int(symb.na[k])
        if not na:
            continue
        blk = _blk(symb, U, k)
        Lt = hf.Lt[k]
        Mnn = _sym(blk[:, :nn, :])
        Man = blk[:, nn:, :]
        Zaa = _gather_aa(symb, U, k)
        Zan = Man - Zaa @ Lt
        Znn = Mnn - Lt.T @ Man - Zan.swapaxes(1, 2) @ Lt
        _write_lower(blk, nn, 0.5 * (Znn + Znn.swapaxes(1, 2)), Zan)


def hessian_inv(hf, Z):
    """Z <- H^{-1}(Z): the inverse of ``hessian`` (App. A.5; ``hessian(...,inv=True,
    adj=None)`` e.g. ``solvers.py:405, 1104, 1735, 2021``).  One post-order sweep: each
    supernode reads its own block and the still-untouched alpha x alpha entries of its
    ancestors, then extend-adds its children's updates."""
    symb = hf.symb
    Z = _as2d(Z)
    upd = [None] * symb.nsn
    for k in range(symb.nsn):
        nn, na = int(symb.nn[k]), int(symb.na[k])
        blk = _blk(symb, Z, k)
        Lnn, Lt = hf.Lnn[k], hf.Lt[k]
        D = Lnn @ Lnn.T
        Znn = _sym(blk[:, :nn, :])
        if na:
            Zan = blk[:, nn:, :]
            Zaa = _gather_aa(symb, Z, k)
            Man = Zan + Zaa @ Lt
            Mnn = Znn + Lt.T @ Zan + Man.swapaxes(1, 2) @ Lt
            Knn = D @ Mnn @ D
            Kan = sl.cho_solve((hf.chol_Yaa(k), True), Man) @ D
            Fan = Kan + Lt @ Knn
            Faa = Lt @ Kan.swapaxes(1, 2) + Fan @ Lt.T
        else:
            Knn = D @ Znn @ D
            Fan = blk[:, nn:, :]
            Faa = None
        Fnn = Knn.copy()
        for c in _children(symb, k):
            r = _rel(symb, c)
            own = r < nn
            ro, ra = r[own], r[~own] - nn
            Uc = upd[c]
            Fnn[:, ro[:, None], ro[None, :]] += Uc[:, own][:, :, own]
            if na:
                Fan[:, ra[:, None], ro[None, :]] += Uc[:, ~own][:, :, own]
                Faa[:, ra[:, None], ra[None, :]] += Uc[:, ~own][:, :, ~own]
            upd[c] = None
        if na:
            upd[k] = Faa
        _write_lower(blk, nn, 0.5 * (Fnn + Fnn.swapaxes(1, 2)), Fan)


def hessian_half(hf, U, adj, inv):
    """The half factors of the Hessian (``chompack.hessian(L, Y, U, adj=False/True, inv=...)``;
    ``solvers.py:917, 978, 1121, 1126``).  With Y_aa = R R^T per supernode and pass 1 / pass 3 the two
    congruence sweeps of ``hessian`` (App. A.4):
        G        (adj=False, inv=False) = half scaling o pass 1   : (K_nn, K_an) -> (L^-1 K_nn L^-T, R^T K_an L^-T)
        G^adj    (adj=True,  inv=False) = pass 3 o half scaling^adj: (V_nn, V_an) -> (L^-T V_nn L^-1, R V_an L^-1)
        G^-1     (adj=False, inv=True)  and  G^-adj (adj=True, inv=True) are their inverses,
    so that hessian = G^adj o G, hessian_inv = G^-1 o G^-adj and ||G(U)||^2 = U . hessian(U)."""
    symb = hf.symb
    U = _as2d(U)
    T = lambda a: a.swapaxes(1, 2)
    if not inv and not adj:
        upd = [None] * symb.nsn
        for k in range(symb.nsn):
            nn, na = int(symb.nn[k]), int(symb.na[k])
            F = _frontal(symb, U, k, upd)
            Lnn, Lt = hf.Lnn[k], hf.Lt[k]
            Knn = F[:, :nn, :nn]
            X1 = sl.solve_triangular(Lnn, Knn, lower=True)                            # L^-1 K
            Gnn = T(sl.solve_triangular(Lnn, T(X1), lower=True))                      # L^-1 K L^-T
            if na:
                Kan = F[:, nn:, :nn] - Lt @ Knn
                upd[k] = F[:, nn:, nn:] - Lt @ F[:, :nn, nn:] - Kan @ Lt.T
                Gan = hf.chol_Yaa(k).T @ T(sl.solve_triangular(Lnn, T(Kan), lower=True))   # R^T K_an L^-T
            else:
                Gan = F[:, nn:, :nn]
            _write_lower(_blk(symb, U, k), nn, 0.5 * (Gnn + T(Gnn)), Gan)
    elif not inv and adj:
        for k in range(symb.nsn - 1, -1, -1):
            nn, na = int(symb.nn[k]), int(symb.na[k])
            blk = _blk(symb, U, k)
            Lnn, Lt = hf.Lnn[k], hf.Lt[k]
            Vnn = _sym(blk[:, :nn, :])
            X1 = sl.solve_triangular(Lnn, Vnn, lower=True, trans='T')                  # L^-T V
            Mnn = T(sl.solve_triangular(Lnn, T(X1), lower=True, trans='T'))           # L^-T V L^-1
            if na:
                Man = hf.chol_Yaa(k) @ T(sl.solve_triangular(Lnn, T(blk[:, nn:, :]), lower=True, trans='T'))   # R V_an L^-1
                Zaa = _gather_aa(symb, U, k)
                Zan = Man - Zaa @ Lt
                Znn = Mnn - Lt.T @ Man - T(Zan) @ Lt
            else:
                Zan, Znn = blk[:, nn:, :], Mnn
            _write_lower(blk, nn, 0.5 * (Znn + T(Znn)), Zan)
    elif inv and adj:
        for k in range(symb.nsn):
            nn, na = int(symb.nn[k]), int(symb.na[k])
            blk = _blk(symb, U, k)
            Lnn, Lt = hf.Lnn[k], hf.Lt[k]
            Znn = _sym(blk[:, :nn, :])
            if na:
                Zan = blk[:, nn:, :]
                Zaa = _gather_aa(symb, U, k)
                Man = Zan + Zaa @ Lt
                Mnn = Znn + Lt.T @ Zan + T(Man) @ Lt
                Van = sl.solve_triangular(hf.chol_Yaa(k), Man, lower=True) @ Lnn        # R^-1 M_an L
            else:
                Mnn, Van = Znn, blk[:, nn:, :]
            Vnn = Lnn.T @ Mnn @ Lnn
            _write_lower(blk, nn, 0.5 * (Vnn + T(Vnn)), Van)
    else:
        upd = [None] * symb.nsn
        for k in range(symb.nsn):
            nn, na = int(symb.nn[k]), int(symb.na[k])
            blk = _blk(symb, U, k)
            Lnn, Lt = hf.Lnn[k], hf.Lt[k]
            Knn = Lnn @ _sym(blk[:, :nn, :]) @ Lnn.T
            if na:
                Kan = sl.solve_triangular(hf.chol_Yaa(k), blk[:, nn:, :], lower=True, trans='T') @ Lnn.T   # R^-T V_an L^T
                Fan = Kan + Lt @ Knn
                Faa = Lt @ T(Kan) + Fan @ Lt.T
            else:
                Fan, Faa = blk[:, nn:, :], None
            Fnn = Knn.copy()
            for c in _children(symb, k):
                r = _rel(symb, c)
                own = r < nn
                ro, ra = r[own], r[~own] - nn
                Uc = upd[c]
                Fnn[:, ro[:, None], ro[None, :]] += Uc[:, own][:, :, own]
                if na:
                    Fan[:, ra[:, None], ro[None, :]] += Uc[:, ~own][:, :, own]
                    Faa[:, ra[:, None], ra[None, :]] += Uc[:, ~own][:, :, ~own]
                upd[c] = None
            if na:
                upd[k] = Faa
            _write_lower(blk, nn, 0.5 * (Fnn + T(Fnn)), Fan)


def trsm(symb, L, Bm, trans='N'):
    """Dense right-hand sides: Bm <- L^{-1} Bm ('N') or L^{-T} Bm ('T'), Bm is n x k with
    rows in the *internal* order of ``symb`` (App. A.8; ``chompack.trsm`` at
    ``solvers.py:491-492, 1921-1922``)."""
    L = _as2d(L)
    rng = range(symb.nsn) if trans == 'N' else range(symb.nsn - 1, -1, -1)
    for k in rng:
        nn = int(symb.nn[k])
        blk = _blk(symb, L, k)[0]
        Lnn = np.tril(blk[:nn, :])
        rows = symb.rowidx[symb.rowptr[k]:symb.rowptr[k + 1]]
        rn, ra = rows[:nn], rows[nn:]
        if trans == 'N':
            Bm[rn] = sl.solve_triangular(Lnn, Bm[rn], lower=True)
            if len(ra):
                Bm[ra] -= blk[nn:, :] @ Bm[rn]
        else:
            if len(ra):
                Bm[rn] -= blk[nn:, :].T @ Bm[ra]
            Bm[rn] = sl.solve_triangular(Lnn, Bm[rn], lower=True, trans='T')


def dot(symb, X, Y):
    """Trace inner product of two chordal matrices (App. A.7; ``chompack.dot``)."""
    return float(np.dot(np.ravel(X) * symb.wdot, np.ravel(Y)))


def sumlogdiag(symb, L):
    """sum(log(diag(L))) (``sum(log(L.diag()))`` at ``solvers.py:395, 925``)."""
    return float(np.sum(np.log(np.ravel(L)[symb.diag_blk])))
