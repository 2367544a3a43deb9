"""Host mirror of the chompack operator API the SMCP drivers are written against
(``cspmatrix``, ``cholesky``, ``completion``, ``projected_inverse``,
``llt``, ``hessian``, ``dot``; imported by the reference at ``src/python/solvers.py:82-97,
1364-1379``).  Same names, same in-place semantics, same ``ArithmeticError`` convention.

A ``cspmatrix`` is a thin handle: the values live wherever the *backend* keeps them — in
HBM behind the C-ABI for the product backend (``smcp_b200.device.DeviceBackend``).  The
backend protocol (duck-typed, see ``BackendProtocol``) is the seam through which the parity
tests run the very same driver code on the CPU oracle.
"""
from __future__ import annotations

import numpy as np

__all__ = ["cspmatrix", "cholesky", "completion", "projected_inverse", "llt", "hessian",
           "dot", "BackendProtocol"]


class BackendProtocol:
    """Documentation of the primitive set a backend provides.  ``buf`` is an opaque value
    buffer holding one chordal matrix in ``blkval`` layout.

    storage      new() -> buf (zeros); clone(buf) -> buf; release(buf)
                 from_vec(v) -> buf   (v: |Vp| values, lower-triangular CCS order of Vp)
                 to_vec(buf) -> ndarray
    level-1      axpy(a, x, y): y += a*x;  scal(a, x);  dot(x, y) -> float
                 sumlogdiag(buf) -> float  (= sum(log(diag)))
    factor       cholesky(buf); completion(buf)  -> raise ArithmeticError
                 projected_inverse(buf); llt(buf)
    hessian      hessian_factor(Lbuf, Ybuf) -> token
                 hessian_apply(token, [bufs], inv: bool)
    operator     set_operator(Av, Ns): Av = |Vp| x m CCS in vector-space row order, the
                 last Ns columns are the "sparse" constraints (solvers.py:246-268)
                 Amap(buf) -> ndarray(m);  Amap_col(buf, i) -> float;  Aadj(y) -> buf
    schur        schur_factor(token): assemble H (lower) and factor it in place,
                 raise ArithmeticError if H is not positive definite
                 schur_solve(y) -> ndarray(m)    (potrs)
    """


class cspmatrix:
    """Chordal sparse symmetric matrix (or Cholesky factor) on a fixed supernodal pattern
    — the handle type behind ``chompack.cspmatrix`` at every reference call site."""

    __slots__ = ("ops", "buf", "version", "__weakref__")

    def __init__(self, ops, buf=None):
        self.ops = ops
        self.buf = ops.new() if buf is None else buf
        self.version = 0

    def __del__(self):
        try:
            self.ops.release(self.buf)
        except Exception:
            pass

    # -- construction / extraction ------------------------------------------------
    @classmethod
    def from_vec(cls, ops, v):
        """``cspmatrix(symb) + spmatrix``: project values given in vector-space order."""
        return cls(ops, ops.from_vec(np.ascontiguousarray(v, dtype=np.float64)))

    @classmethod
    def identity(cls, ops, alpha=1.0):
        symb = ops.symb
        v = np.zeros(symb.nvp)
        v[symb.diag_vec] = alpha
        return cls.from_vec(ops, v)

    def copy(self):
        return cspmatrix(self.ops, self.ops.clone(self.buf))

    def to_vec(self):
        """Values of ``X.spmatrix(reordered=False, symmetric=False)`` (lower CCS of Vp)."""
        return self.ops.to_vec(self.buf)

    def sumlogdiag(self):
        """``sum(log(X.diag()))`` (``solvers.py:395, 925``)."""
        return self.ops.sumlogdiag(self.buf)

    def scale(self, a):
        """``blas.scal(a, X.blkval)``."""
        self.ops.scal(float(a), self.buf)
        self.version += 1
        return self

    # -- arithmetic (new objects, like chompack) -----------------------------------
    def __add__(self, other):
        r = self.copy()
        r.ops.axpy(1.0, other.buf, r.buf)
        return r

    def __sub__(self, other):
        r = self.copy()
        r.ops.axpy(-1.0, other.buf, r.buf)
        return r

    def __iadd__(self, other):
        self.ops.axpy(1.0, other.buf, self.buf)
        self.version += 1
        return self

    def __isub__(self, other):
        self.ops.axpy(-1.0, other.buf, self.buf)
        self.version += 1
        return self

    def __mul__(self, a):
        r = self.copy()
        r.ops.scal(float(a), r.buf)
        return r

    __rmul__ = __mul__

    def __neg__(self):
        return self * -1.0


def _inplace(X):
    X.version += 1
    return X


def cholesky(X):
    """In place X -> L, L L^T = X; ``ArithmeticError`` if X is not positive definite."""
    X.ops.cholesky(_inplace(X).buf)


def completion(X):
    """In place X -> L with P_V((L L^T)^{-1}) = X; ``ArithmeticError`` if X has no
    positive definite completion."""
    X.ops.completion(_inplace(X).buf)


def projected_inverse(L):
    """In place L -> P_V((L L^T)^{-1})."""
    L.ops.projected_inverse(_inplace(L).buf)


def llt(L):
    """In place L -> P_V(L L^T)."""
    L.ops.llt(_inplace(L).buf)


def dot(X, Y):
    """Trace inner product of two chordal matrices."""
    return X.ops.dot(X.buf, Y.buf)


_hf_cache = {}


def _factor_token(L, Y):
    ops = L.ops
    key = id(ops)
    ent = _hf_cache.get(key)
    sig = (id(L), L.version, id(Y), Y.version)
    if ent is None or ent[0] != sig:
        tok = ops.hessian_factor(L.buf, Y.buf)
        _hf_cache.clear()           # one live scaling point at a time, like the reference
        _hf_cache[key] = (sig, tok, L, Y)
        return tok
    return ent[1]


def hessian(L, Y, U, inv=False, adj=None):
    """``chompack.hessian(L, Y, U, adj=None, inv=...)``: U <- P_V(S^{-1} U S^{-1}) with
    S = L L^T (``inv=False``) or its inverse map (``inv=True``).  U is a cspmatrix or a list
    of them (evaluated as one device batch).  ``adj=False`` / ``adj=True`` apply the half factors
    G / G^adj (``inv=True``: G^-1 / G^-adj) with ``hessian = G^adj o G``, as chompack does
    (``solvers.py:917, 978, 1121, 1126``).  The drivers only need ``||G(u)||`` and take the cheaper
    ``hessian_norm``; the half factors themselves run through the generic supernodal sweeps."""
    Us = U if isinstance(U, (list, tuple)) else [U]
    tok = _factor_token(L, Y)
    if adj is None:
        L.ops.hessian_apply(tok, [_inplace(u).buf for u in Us], bool(inv))
    else:
        L.ops.hessian_apply(tok, [_inplace(u).buf for u in Us], bool(inv), bool(adj))


def hessian_norm(L, Y, u, inv):
    """``sqrt(dot(du,du))`` after ``hessian(L,Y,[du],adj=True,inv=inv)`` (or the ``adj=False``
    half when ``inv=False``), evaluated as ``sqrt(u . H^{+-1}(u))`` — identical because
    H = G^adj G (``solvers.py:916-918, 977-979, 1119-1128``)."""
    w = u.copy()
    hessian(L, Y, w, inv=inv)
    val = dot(u, w)
    return float(np.sqrt(val)) if val > 0.0 else 0.0


def schur_token(L, Y):
    return _factor_token(L, Y)
