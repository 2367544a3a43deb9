"""CPU backend for the SMCP drivers: the reference's Newton-system path restated on NumPy /
SciPy (TEST INFRASTRUCTURE and timed CPU baseline; see ``oracle/__init__.py``).

It implements ``smcp_b200.chordal.BackendProtocol`` so that ``smcp_b200.solvers`` — which
follows ``src/python/solvers.py`` — runs unchanged on it.  ``schur_factor`` restates
``kkt_chol`` (``solvers.py:477-504`` == ``1906-1934``) column by column exactly as the
reference does, including the sparse-constraint technique built on
``misc.SCMcolumn2`` (``src/C/misc.c:620-663``).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sl
import scipy.sparse as sp

from . import supernodal as sn


class OracleBackend:
    name = "oracle"

    def __init__(self, symb, batch_columns=1):
        self.symb = symb
        self.batch_columns = int(batch_columns)   # 1 = per-column loop like the reference
        self.Av = None
        self.H = None
        self.stats = {"hessian_cols": 0, "schur_calls": 0}

    # -- storage ----------------------------------------------------------------------
    def new(self):
        return np.zeros(self.symb.nblk)

    def clone(self, buf):
        return buf.copy()

    def release(self, buf):
        pass

    def from_vec(self, v):
        x = np.zeros(self.symb.nblk)
        x[self.symb.vec2blk] = v
        return x

    def to_vec(self, buf):
        return buf[self.symb.vec2blk].copy()

    # -- level 1 ----------------------------------------------------------------------
    def axpy(self, a, x, y):
        if a == 1.0:
            y += x
        elif a == -1.0:
            y -= x
        else:
            y += a * x

    def scal(self, a, x):
        x *= a

    def dot(self, x, y):
        return sn.dot(self.symb, x, y)

    def sumlogdiag(self, buf):
        return sn.sumlogdiag(self.symb, buf)

    # -- factorizations ---------------------------------------------------------------
    def cholesky(self, buf):
        sn.cholesky(self.symb, buf)

    def completion(self, buf):
        sn.completion(self.symb, buf)

    def projected_inverse(self, buf):
        sn.projected_inverse(self.symb, buf)

    def trsm(self, Lbuf, B, trans="N"):
        """chompack.trsm: returns L^{-1} B / L^{-T} B (rows of B in the internal order)."""
        out = np.array(B, dtype=np.float64)
        sn.trsm(self.symb, Lbuf, out, trans)
        return out

    def llt(self, buf):
        sn.llt(self.symb, buf)

    # -- hessian ----------------------------------------------------------------------
    def hessian_factor(self, Lbuf, Ybuf):
        hf = sn.HessianFactor(self.symb, Lbuf, Ybuf)
        hf.Lbuf = Lbuf.copy()
        return hf

    def hessian_apply(self, hf, bufs, inv, adj=None):
        for b in bufs:
            if adj is not None:
                sn.hessian_half(hf, b, bool(adj), bool(inv))
            elif inv:
                sn.hessian_inv(hf, b)
            else:
                sn.hessian(hf, b)

    # -- operator ---------------------------------------------------------------------
    def set_operator(self, Av, Ns):
        symb = self.symb
        self.Av = sp.csc_matrix(Av)
        self.m = self.Av.shape[1]
        self.Ns = int(Ns)
        self.AvT = self.Av.T.tocsr()
        self.halfdiag = np.ones(symb.nvp)
        self.halfdiag[symb.diag_vec] = 0.5
        self.H = np.zeros((self.m, self.m), order="F")

    def Amap(self, buf):
        return 2.0 * (self.AvT @ (self.to_vec(buf) * self.halfdiag))

    def Amap_col(self, buf, i):
        col = self.Av[:, [i]]
        return float(2.0 * (col.T @ (self.to_vec(buf) * self.halfdiag))[0])

    def Aadj(self, y):
        return self.from_vec(self.Av @ np.asarray(y, dtype=np.float64).ravel())

    # -- Schur complement -------------------------------------------------------------
    def schur_assemble(self, hf, columns=None):
        """Lower triangle of H, H_ij = A_i . Hess(A_j) (``solvers.py:479-497``).  ``columns``: list of
        (j0, j1) column ranges to assemble (the blocks a rank owns); the other columns stay zero."""
        symb, Av, m, Ns = self.symb, self.Av, self.m, self.Ns
        H = self.H
        H[...] = 0.0
        md = m - Ns
        B = max(1, self.batch_columns)
        own = np.ones(m, dtype=bool)
        if columns is not None:
            own[:] = False
            for c0, c1 in columns:
                own[c0:c1] = True
        # technique 1: one Hessian evaluation per "dense" constraint
        dense_cols = [j for j in range(md) if own[j]]
        for b0 in range(0, len(dense_cols), B):
            batch = dense_cols[b0:b0 + B]
            U = np.zeros((len(batch), symb.nblk))
            for q, j in enumerate(batch):
                c0, c1 = Av.indptr[j], Av.indptr[j + 1]
                U[q, symb.vec2blk[Av.indices[c0:c1]]] = Av.data[c0:c1]
            sn.hessian(hf, U)
            self.stats["hessian_cols"] += len(batch)
            for q, j in enumerate(batch):
                at = U[q, symb.vec2blk] * self.halfdiag
                H[j:, j] = 2.0 * (self.AvT[j:, :] @ at)
        # technique 2: sparse constraints through columns of S^{-1}
        if Ns:
            Ip, Jp = symb.Ip, symb.Jp
            for j in range(Ns):
                jj = md + j
                if not own[jj]:
                    continue
                c0, c1 = Av.indptr[jj], Av.indptr[jj + 1]
                rows_j = Av.indices[c0:c1]
                K = np.unique(np.concatenate([Ip[rows_j], Jp[rows_j]]))
                V = np.zeros((symb.n, len(K)))
                V[symb.iperm[K], np.arange(len(K))] = 1.0
                sn.trsm(symb, hf.Lbuf, V, 'N')
                sn.trsm(symb, hf.Lbuf, V, 'T')
                kkl = np.zeros(symb.n, dtype=np.int64)
                kkl[K] = np.arange(len(K))
                _scm_column2(H, Av, V, symb.iperm, Ip, Jp, kkl, jj)
        self.stats["schur_calls"] += 1
        return H

    def schur_factor(self, hf):
        H = self.schur_assemble(hf)
        try:
            self.Hf = sl.cholesky(H, lower=True, check_finite=False)
        except sl.LinAlgError:
            raise ArithmeticError("Schur complement is not positive definite")
        if not np.all(np.isfinite(self.Hf)):
            raise ArithmeticError("Schur complement is not positive definite")

    # -- kktsolver='qr' (solvers.py:413-475): Z = [G(A_1) ... G(A_m)], Householder QR like the reference
    def schur_factor_qr(self, hf):
        symb, Av, m = self.symb, self.Av, self.m
        Z = np.zeros((m, symb.nblk))
        for j in range(m):
            c0, c1 = Av.indptr[j], Av.indptr[j + 1]
            Z[j, symb.vec2blk[Av.indices[c0:c1]]] = Av.data[c0:c1]
        sn.hessian_half(hf, Z, False, False)
        self.Zw = (Z * np.sqrt(symb.wdot)[None, :]).T.copy()          # nblk x m, weighted: Z^T Z = trace inner products
        R = np.linalg.qr(self.Zw, mode="r")
        d = np.abs(np.diag(R))
        if not np.all(np.isfinite(R)) or d.min() <= 1e-14 * d.max():
            raise ArithmeticError("Z is rank deficient")
        Ru = R * np.sign(np.diag(R))[:, None]
        self.Hf = np.ascontiguousarray(Ru.T)                            # lower factor of Z^T Z
        self.H = None

    def gram_factor(self):
        G = 2.0 * (self.AvT @ sp.diags(self.halfdiag) @ self.Av).toarray()      # <A_i, A_j>, trace inner product
        try:
            self.Hf = sl.cholesky(G, lower=True, check_finite=False)
        except sl.LinAlgError:
            raise ArithmeticError("the constraint matrices are linearly dependent")

    def z_tmul(self, buf):
        return self.Zw.T @ (buf * np.sqrt(self.symb.wdot))

    def z_mul(self, y):
        w = np.sqrt(self.symb.wdot)
        out = self.Zw @ np.asarray(y, dtype=np.float64).ravel()
        return np.divide(out, w, out=np.zeros_like(out), where=w > 0)

    def schur_solve(self, y):
        return sl.cho_solve((self.Hf, True), np.asarray(y, dtype=np.float64).ravel(),
                            check_finite=False)


def _scm_column2(H, Av, V, iperm, Ip, Jp, kkl, j):
    """Column j (rows >= j) of H for a sparse constraint — ``misc.SCMcolumn2``
    (``src/C/misc.c:620-663``): sum over the entries (alpha; r, c) of A_j and (beta; r1, c1)
    of A_i of alpha*beta*[V(r1,r)V(c1,c) + (r1 != c1) V(c1,r)V(r1,c)], alpha doubled off the
    diagonal.  V rows are in the internal order of the symbolic object."""
    m = Av.shape[1]
    c0, c1 = Av.indptr[j], Av.indptr[j + 1]
    kj = Av.indices[c0:c1]
    alpha = Av.data[c0:c1] * np.where(Ip[kj] != Jp[kj], 2.0, 1.0)
    kr, kc = kkl[Ip[kj]], kkl[Jp[kj]]
    e0, e1 = Av.indptr[j], Av.indptr[m]
    ke = Av.indices[e0:e1]
    beta = Av.data[e0:e1]
    col_e = np.repeat(np.arange(j, m), np.diff(Av.indptr[j:m + 1]))
    r1, c1i = iperm[Ip[ke]], iperm[Jp[ke]]
    off = (Ip[ke] != Jp[ke]).astype(np.float64)
    acc = np.zeros(len(ke))
    for p in range(len(kj)):
        t = V[r1, kr[p]] * V[c1i, kc[p]] + off * V[c1i, kr[p]] * V[r1, kc[p]]
        acc += alpha[p] * beta * t
    H[j:, j] = np.bincount(col_e - j, weights=acc, minlength=m - j)
