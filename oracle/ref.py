"""NumPy-facing wrapper around the reference's OWN C extension ``smcp.misc`` compiled from
``/root/reference/src/C/misc.c`` into ``oracle/_ref/misc.so`` (recipe: ``oracle/Makefile``;
ABI shim for the absent ``cvxopt.base``: ``oracle/refshim/cvxopt_base_shim.c``).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  This is the one place where parity
is pinned on code of the reference itself: the members of ``smcp.misc`` that sit on the
Newton-system path (``misc.c:1057-1102``) are called unmodified and compared with

* ``smcp_b200/misc.py``          nzcolumns / matperm / phase1_sdp / ind2sub / sub2ind
* ``oracle/backend.py``          the sparse-constraint Schur column (``_scm_column2``)
* the CUDA kernel ``scm_sparse_kernel`` (through ``smcp_kkt_assemble``) — via the golden
  vectors this module generated (``tests/golden/misc_ref_*.npz``, ``scripts/make_golden.py``)

``available()`` is False when the shared objects were not built (no ``/root/reference`` and
no prebuilt ``oracle/_ref``); callers skip in that case.
"""
from __future__ import annotations

import os
import sys

import numpy as np

_REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_mod = None
_base = None


def available():
    return os.path.exists(os.path.join(_REF, "misc.so")) and os.path.exists(os.path.join(_REF, "cvxopt", "base.so"))


def _load():
    global _mod, _base
    if _mod is None:
        if not available():
            raise RuntimeError("oracle/_ref is not built (run `make -C oracle`; needs /root/reference)")
        if "cvxopt" in sys.modules and not getattr(sys.modules["cvxopt"], "__file__", "").startswith(_REF):
            raise RuntimeError("a real cvxopt is already imported; refusing to shadow it with the ABI shim")
        sys.path.insert(0, _REF)
        try:
            import cvxopt.base as base      # the shim
            import misc as refmisc          # the reference's misc.c, unmodified
        finally:
            sys.path.remove(_REF)
        _mod, _base = refmisc, base
    return _mod, _base


# ---- conversions ----------------------------------------------------------------------
def imatrix(a):
    _, base = _load()
    a = np.ascontiguousarray(np.asarray(a, dtype=np.int64).ravel())
    return base.matrix_from(a.tobytes(), len(a), 1, 0)


def dmatrix(a):
    """dense float64 matrix (2-D arrays are stored column-major like cvxopt)."""
    _, base = _load()
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    return base.matrix_from(np.asfortranarray(a).tobytes(order="F"), a.shape[0], a.shape[1], 1)


def spmatrix(A):
    """scipy CSC (sorted indices) -> shim spmatrix"""
    import scipy.sparse as sp
    _, base = _load()
    A = sp.csc_matrix(A)
    A.sort_indices()
    return base.spmatrix_from(np.ascontiguousarray(A.data, dtype=np.float64).tobytes(),
                              np.ascontiguousarray(A.indptr, dtype=np.int64).tobytes(),
                              np.ascontiguousarray(A.indices, dtype=np.int64).tobytes(),
                              A.shape[0], A.shape[1])


def to_numpy(M):
    nr, nc = M.size
    dt = np.int64 if M.id == 0 else np.float64
    return np.frombuffer(M.tobytes(), dtype=dt).reshape((nr, nc), order="F").copy()


def to_scipy(S):
    import scipy.sparse as sp
    nr, nc = S.size
    return sp.csc_matrix((np.frombuffer(S.values_bytes(), dtype=np.float64).copy(),
                          np.frombuffer(S.rowind_bytes(), dtype=np.int64).copy(),
                          np.frombuffer(S.colptr_bytes(), dtype=np.int64).copy()), shape=(nr, nc))


# ---- the reference's functions ----------------------------------------------------------
def nzcolumns(A):
    """misc.c:682-730"""
    m, _ = _load()
    return to_numpy(m.nzcolumns(spmatrix(A))).ravel()


def matperm(nzc, Nmax):
    """misc.c:750-773"""
    m, _ = _load()
    pm, Ns = m.matperm(imatrix(nzc), int(Nmax))
    return to_numpy(pm).ravel(), int(Ns)


def ind2sub(n, ind):
    """misc.c:387-408"""
    m, _ = _load()
    I, J = m.ind2sub(int(n), imatrix(ind))
    return to_numpy(I).ravel(), to_numpy(J).ravel()


def sub2ind(siz, I, J):
    """misc.c:427-447"""
    m, _ = _load()
    return to_numpy(m.sub2ind(tuple(int(s) for s in siz), imatrix(I), imatrix(J))).ravel()


def phase1_sdp(A, u):
    """misc.c:1004-1054"""
    m, _ = _load()
    return to_scipy(m.phase1_sdp(spmatrix(A), dmatrix(u)))


def Av_to_spmatrix(Av, Ip, Jp, j, n, scale=False):
    """misc.c:475-521"""
    m, _ = _load()
    return to_scipy(m.Av_to_spmatrix(spmatrix(Av), imatrix(Ip), imatrix(Jp), int(j), int(n), bool(scale)))


def scal_diag(values_of_Vp, colptr, rowind, shape, Id, t=0.5):
    """misc.c:542-557 on a CCS matrix given by its arrays; returns the scaled values."""
    import scipy.sparse as sp
    m, _ = _load()
    S = spmatrix(sp.csc_matrix((values_of_Vp, rowind, colptr), shape=shape))
    m.scal_diag(S, imatrix(Id), float(t))
    return np.frombuffer(S.values_bytes(), dtype=np.float64).copy()


def SCMcolumn2(H, Av, V, Ip, Jp, Kl, j):
    """misc.c:620-663: column j (rows >= j) of the Schur complement for a sparse constraint.
    H: m x m array (updated copy is returned); V: n x k dense (columns of S^-1), Kl: vertex ->
    column of V."""
    m, _ = _load()
    Hm = dmatrix(H)
    m.SCMcolumn2(Hm, spmatrix(Av), dmatrix(V), imatrix(Ip), imatrix(Jp), imatrix(Kl), int(j))
    return to_numpy(Hm)


def sdpa_read(fname, neg=False):
    """The reference's own ``misc.sdpa_read`` (``misc.c:139-245``) -> (scipy csc, b, blockstruct)."""
    mod, _ = _load()
    A, b, bs = mod.sdpa_read(fname, neg=bool(neg))
    return to_scipy(A), to_numpy(b).ravel(), to_numpy(bs).ravel().astype(np.int64)
