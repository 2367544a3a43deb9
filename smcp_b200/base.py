"""Problem containers and random problem builders (API surface of ``smcp.base``,
reference ``src/python/base.py``).

``SDP`` keeps the reference's data layout — ``A`` CCS ``n^2 x (m+1)`` with column 0 = vec(C)
and lower-triangular entries only, ``b`` dense — and its three solve entry points on the
hot path: ``solve_feas`` (``base.py:346-368``), ``solve_esd`` (``316-344``) and
``solve_phase1`` (``370-470``), plus SDPA (dat-s) / pickle I/O with the reference's signatures.
``solve_cvxopt`` and the robust-LS converters are outside the path and not provided.

The generators restate ``band_SDP`` (``base.py:598-636``), ``mtxnorm_SDP`` (``707-778``),
``rand_SDP`` (``879-949``) and ``mk_rand`` (``514-560``) with NumPy's ``default_rng``:
cvxopt's RNG stream (``setseed``/``normal``) cannot be reproduced, so instances differ from
the reference's for the same seed but follow the same construction.  ``maxcut_SDP`` builds
the max-cut relaxation the reference reads from SDPLIB files.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import misc, solvers
from .symbolic import Symbolic, embed, maxcardsearch, min_degree, lower_pattern

__all__ = ["SDP", "band_SDP", "mtxnorm_SDP", "rand_SDP", "maxcut_SDP", "mk_rand", "completion"]


class SDP(object):
    """SDP object: ``n``, ``m``, ``A`` (CCS n^2 x (m+1)), ``b``, aggregate sparsity ``I``."""

    def __init__(self, A=None, b=None, name=None, filename=None):
        """``SDP(A, b)`` from problem data, or ``SDP(filename)`` / ``SDP(filename='x.dat-s')`` from a sparse
        SDPA (dat-s) or pickle file (``base.py:62-85, 177-195``: the file's primal is negated into SMCP's
        standard form)."""
        if filename is not None:
            if A is not None and not isinstance(A, str):
                raise TypeError("give either problem data or a file name")
            A = filename
        self._I = None
        self._ischordal = None
        self._blockstruct = None
        self._X0 = self._y0 = self._S0 = None
        if isinstance(A, str):
            import os
            fp, ext = os.path.splitext(A)
            if ext == ".bz2" and fp.endswith(".pkl"):
                self._load(A)
                return
            if ext == ".bz2":
                # compressed SDPLIB files (``base.py:180-190``): decompress next to a temporary name
                import bz2, tempfile
                with open(A, "rb") as fc, tempfile.NamedTemporaryFile("wb", suffix=".dat-s", delete=False) as fo:
                    fo.write(bz2.decompress(fc.read()))
                    tmp = fo.name
                try:
                    self._A, self._b, self._blockstruct = misc.sdpa_read(tmp, neg=True)
                finally:
                    os.remove(tmp)
                fp = os.path.splitext(fp)[0]
            elif ext == ".dat-s":
                self._A, self._b, self._blockstruct = misc.sdpa_read(A, neg=True)
            elif ext == ".pkl":
                self._load(A)
                return
            else:
                raise NameError("Unknown file extension")
            self._pname = name or os.path.basename(fp)
            return
        self._A = misc.as_csc(A) if A is not None else None
        self._b = np.asarray(b, dtype=np.float64).ravel() if b is not None else None
        self._pname = name

    def _load(self, fname):
        """Load problem data saved by ``save`` (``base.py:219-238``)."""
        import bz2, pickle
        raw = open(fname, "rb").read()
        D = pickle.loads(bz2.decompress(raw) if fname.endswith(".bz2") else raw)
        self._A = misc.as_csc(D["A"])
        self._b = np.asarray(D["b"], dtype=np.float64).ravel()
        self._X0, self._y0, self._S0, self._pname = D["X0"], D["y0"], D["S0"], D["pname"]

    def save(self, fname=None, compress=False):
        """Save problem data (and the generator's strictly feasible point, if any) to a pickle file
        ``<fname>.pkl`` or ``<fname>.pkl.bz2`` (``base.py:240-270``); refuses to overwrite."""
        import bz2, os, pickle
        fname = (fname or self._pname) + (".pkl.bz2" if compress else ".pkl")
        if os.path.isfile(fname):
            raise IOError("file %s already exists" % fname)
        D = {"A": self._need(), "b": self._b, "X0": self._X0, "y0": self._y0, "S0": self._S0, "pname": self._pname}
        raw = pickle.dumps(D)
        with open(fname, "wb") as f:
            f.write(bz2.compress(raw) if compress else raw)
        return fname

    def write_sdpa(self, fname=None, compress=False):
        """Writes the problem to the sparse SDPA file ``<fname>.dat-s`` (``<fname>`` defaults to the problem
        name; ``compress=True``: ``<fname>.dat-s.bz2``), refusing to overwrite an existing file
        (``base.py:197-217``; one block unless the object was read from a file with a block structure).
        Returns the name of the file written."""
        import bz2, os
        bs = self._blockstruct if self._blockstruct is not None else np.array([self.n], dtype=np.int64)
        if fname is None:
            fname = self._pname
        if fname is None:
            raise ValueError("the problem has no name: give a file name")
        fname += ".dat-s"
        if os.path.isfile(fname):
            raise IOError("file %s already exists" % fname)
        if compress and os.path.isfile(fname + ".bz2"):
            raise IOError("file %s already exists" % (fname + ".bz2"))
        misc.sdpa_write(fname, self._need(), self.b, bs, neg=True)
        if compress:
            with open(fname, "rb") as fi, open(fname + ".bz2", "wb") as fo:
                fo.write(bz2.compress(fi.read()))
            os.remove(fname)
            fname += ".bz2"
        return fname

    def __str__(self):
        return "<SDP: n=%i, m=%i, nnz=%i> %s" % (self.n, self.m, self.nnz, self._pname)

    def _need(self):
        if self._A is None:
            raise AttributeError("SDP object has not been initialized")
        return self._A

    @property
    def n(self):
        return int(np.sqrt(self._need().shape[0]))

    @property
    def m(self):
        return self._need().shape[1] - 1

    @property
    def A(self):
        return self._need()

    def get_A(self, i=None):
        """A if i is None, otherwise A_i as an n x n lower-triangular sparse matrix."""
        A = self._need()
        if i is None:
            return A
        if not 0 <= i <= self.m:
            raise ValueError("index is out of range")
        r = A.indices[A.indptr[i]:A.indptr[i + 1]]
        v = A.data[A.indptr[i]:A.indptr[i + 1]]
        Il, Jl = misc.ind2sub(self.n, r)
        return sp.csc_matrix((v, (Il, Jl)), shape=(self.n, self.n))

    @property
    def b(self):
        if self._b is None:
            raise AttributeError("SDP object has not been initialized")
        return self._b

    @property
    def I(self):
        """Aggregate sparsity pattern as absolute (linear) indices."""
        if self._I is None:
            self._I = np.unique(self._need().indices)
        return self._I

    @property
    def V(self):
        I, J = misc.ind2sub(self.n, self.I)
        return sp.csc_matrix((np.zeros(len(I)), (I, J)), shape=(self.n, self.n))

    @property
    def nnz(self):
        return len(self.I)

    @property
    def issparse(self):
        return len(self.I) <= 0.5 * (self.n * (self.n + 1) / 2)

    def get_nnz(self, i=None):
        """Number of non-zeros in the lower triangle of A_0 .. A_m, or of A_i (``base.py:279-293``)."""
        cnt = np.diff(self._need().indptr)
        if i is None:
            return cnt
        if not 0 <= i <= self.m:
            raise ValueError("index out of range")
        return int(cnt[i])

    nnzs = property(get_nnz, doc="Vector with number of nonzeros in lower triangle of A0,A1,...,Am")

    def get_nzcols(self, i=None):
        """Number of non-zero columns of A_1 .. A_m, or of A_i, i >= 1 (``base.py:299-310``)."""
        nzc = misc.nzcolumns(self._need())
        if i is None:
            return nzc
        if not 0 < i <= self.m:
            raise ValueError("index out of range")
        return int(nzc[i - 1])

    nzcols = property(get_nzcols, doc="Vector with number of nonzero columns in A1,..,Am")

    @property
    def ischordal(self):
        if self._ischordal is None:
            I, J = misc.ind2sub(self.n, self.I)
            cp, ri = lower_pattern(self.n, I, J)
            p = maxcardsearch(self.n, cp, ri)
            fc, _, _ = embed(self.n, cp, ri, p)
            self._ischordal = bool(fc[-1] == cp[-1])
        return self._ischordal

    # -- solvers ------------------------------------------------------------------
    def solve_esd(self, kktsolver="chol", scaling="primal", primalstart=None, dualstart=None, p=None):
        return solvers.chordalsolver_esd(self.A, self.b, kktsolver=kktsolver, scaling=scaling,
                                         primalstart=primalstart, dualstart=dualstart, p=p)

    def solve_feas(self, kktsolver="chol", scaling="primal", primalstart=None, dualstart=None):
        return solvers.chordalsolver_feas(self.A, self.b, kktsolver=kktsolver, scaling=scaling,
                                          primalstart=primalstart, dualstart=dualstart)

    def solve_phase1(self, kktsolver="chol", MM=1e5):
        """Primal Phase I with the feasible-start solver; returns ``(X0, sol)`` with a primal
        feasible ``X0`` (``base.py:370-470``).  The least-norm start (``base.py:383-396``: sparse
        ``syrk`` + CHOLMOD in the reference) runs on the backend: Gram matrix of the constraints by one
        triangular product, dense Cholesky, ``A^adj``."""
        from .chordal import cspmatrix, completion
        n, m = self.n, self.m
        k = 1e-3
        A = self._need()
        Id = np.arange(n, dtype=np.int64) * (n + 1)
        # least-norm start X0 = sum_j u_j A_j with G u = b, G_ij = <A_i, A_j> (base.py:383-396: the reference
        # forms As^T As with syrk and solves with CHOLMOD).  On the backend: G as ONE triangular product of
        # the constraint vectors (DMMA / TMA on the device), the dense Cholesky of the Schur complement, and
        # A^adj for the combination -- on the same embedding the completability test below needs.
        opt = solvers._read_options(n, True)
        prob = solvers._Problem(A, self.b, opt, "qr", None)          # 'qr': no constraint reordering (Ns = 0)
        ops, symb, pm = prob.ops, prob.symb, prob.p
        dense_bytes = 8.0 * symb.nblk * m
        if dense_bytes <= (16 << 30):
            ops.gram_factor()
            u = ops.schur_solve(prob.b)
            Xc = prob.Aadj(u)
            X0 = sp.tril(prob.to_original(Xc), format="csc")
            X0.eliminate_zeros()
        else:
            # constraint vectors too large for the dense product: sparse normal equations on the host
            As = A[:, 1:].tocsc().copy()
            scale = np.ones(A.shape[0])
            scale[Id] = 1.0 / np.sqrt(2.0)
            As = (sp.diags(scale) @ As).tocsc()
            u = spla.spsolve((As.T @ As).tocsc(), self.b)
            x = 0.5 * (A[:, 1:] @ u)
            I, J = misc.ind2sub(n, self.I)
            X0 = sp.csc_matrix((x[self.I], (I, J)), shape=(n, n))
            Xc = prob.from_original(X0)

        def completable(Z):
            L = Z.copy()
            try:
                completion(L)
                return True
            except ArithmeticError:
                return False

        if completable(Xc):
            return X0, None

        trA = np.zeros(m + 1)
        trA[:m] = np.asarray(A[Id, 1:].sum(axis=0)).ravel()
        trA[-1] = MM
        P1 = SDP(misc.phase1_sdp(A, trA), np.concatenate([self.b - k * trA[:m], [MM]]))

        tMIN, tMAX = 0.0, 1.0
        while True:
            t = (tMIN + tMAX) / 2.0
            if completable(Xc.copy() + cspmatrix.identity(ops, t)):
                tMAX = t
                if tMAX - tMIN < 1e-1:
                    break
            else:
                tMAX *= 2.0
                tMIN = t
        tt = t + 1.0
        U = (X0 + tt * sp.identity(n, format="csc")).tocsc()
        trU = U.diagonal().sum()
        Z0 = sp.block_diag([U, sp.diags([tt + k, MM - trU])], format="csc")
        sol = P1.solve_feas(primalstart={"x": Z0}, kktsolver=kktsolver)
        s = sol["x"][n, n] - k
        if s > 0:
            return None, P1
        sol.pop("y")
        sol.pop("s")
        Xz = sol.pop("x")
        X0 = (Xz[:n, :n] - s * sp.identity(n, format="csc")).tocsc()
        return X0, sol


# --------------------------------------------------------------------------------------
# generators
# --------------------------------------------------------------------------------------
def _pattern_cols(n, colptr):
    return np.repeat(np.arange(n, dtype=np.int64), np.diff(colptr))


def mk_rand(n, colptr, rowind, cone="posdef", seed=0):
    """Random matrix with the given lower pattern that is positive definite ('posdef') or
    has a positive definite completion ('completable'): U = P_V(sum_i u_i u_i^T) with
    u_i ~ N(0, I/n), plus the first shift t*I, t in {0.1, 0.2, 0.4, ...}, that puts it in
    the cone (``base.py:514-560``).  The cone tests run on the solver backend (the chordal
    Cholesky / completion kernels), as in the reference.  Returns values in pattern order."""
    from .chordal import cspmatrix, cholesky, completion
    if cone not in ("posdef", "completable"):
        raise ValueError("cone must be 'posdef' (default) or 'completable' ")
    rng = np.random.default_rng(seed)
    colptr = np.asarray(colptr, dtype=np.int64)
    rowind = np.asarray(rowind, dtype=np.int64)
    cols = _pattern_cols(n, colptr)
    U = np.zeros(len(rowind))
    chunk = max(1, min(n, (1 << 24) // max(n, 1)))
    for i0 in range(0, n, chunk):
        G = rng.standard_normal((n, min(chunk, n - i0))) / np.sqrt(n)
        U += np.einsum("ij,ij->i", G[rowind], G[cols])
    # embedding + symbolic for the cone test
    pm = maxcardsearch(n, colptr, rowind)
    fc, fr, _ = embed(n, colptr, rowind, pm)
    if fc[-1] != colptr[-1]:
        pm = min_degree(n, colptr, rowind)
        fc, fr, _ = embed(n, colptr, rowind, pm)
    symb = Symbolic(n, fc, fr)
    ops = solvers._make_backend(symb)
    ip = np.empty(n, dtype=np.int64)
    ip[pm] = np.arange(n, dtype=np.int64)
    pr, pc = ip[rowind], ip[cols]
    key = np.minimum(pr, pc) * n + np.maximum(pr, pc)
    vkey = symb.Jp * n + symb.Ip
    pos = np.searchsorted(vkey, key)            # Vp is CCS-sorted: keys ascending
    diag = rowind == cols
    test = cholesky if cone == "posdef" else completion
    t = 0.1
    Ut = U.copy()
    while True:
        v = np.zeros(symb.nvp)
        v[pos] = Ut
        Z = cspmatrix.from_vec(ops, v)
        try:
            test(Z)
            return Ut
        except ArithmeticError:
            Ut = U.copy()
            Ut[diag] += t
            t *= 2.0


def _band_pattern(n, bw):
    I = np.concatenate([np.arange(j, min(j + bw + 1, n)) for j in range(n)])
    J = np.concatenate([np.full(min(j + bw + 1, n) - j, j) for j in range(n)])
    return I.astype(np.int64), J.astype(np.int64)


class band_SDP(SDP):
    """Random SDP with band structure, ``band_SDP(n, m, bw, seed=0)`` (``base.py:563-636``):
    pattern {(i,j): j <= i <= j+bw}; every A_i dense on the band with N(0, 1/|V|^2) entries;
    A_0 = S0 + sum y0_i A_i and b_i = A_i . X0 with S0 > 0 and X0 completable, so a strictly
    feasible primal-dual pair is known."""

    def __init__(self, n, m, bw, seed=0):
        SDP.__init__(self)
        if type(seed) is not int:
            raise ValueError("seed must be an integer")
        self._bw = bw
        rng = np.random.default_rng(seed)
        I1, J1 = _band_pattern(n, bw)
        N = len(I1)
        Il = misc.sub2ind((n, n), I1, J1)
        diag = I1 == J1
        cp, ri = lower_pattern(n, I1, J1)
        y0 = rng.standard_normal(m)
        y0 /= np.linalg.norm(y0)
        S0 = mk_rand(n, cp, ri, "posdef", seed)
        X0 = mk_rand(n, cp, ri, "completable", seed)
        A_ = rng.standard_normal((N, m + 1)) * (1.0 / N)
        A_[:, 0] = S0 + A_[:, 1:] @ y0
        x = X0.copy()
        x[diag] *= 0.5
        self._b = 2.0 * (A_[:, 1:].T @ x)
        rows = np.tile(Il, m + 1)
        cols = np.repeat(np.arange(m + 1, dtype=np.int64), N)
        self._A = misc.as_csc(sp.csc_matrix((A_.T.reshape(-1), (rows, cols)), shape=(n * n, m + 1)))
        self._X0 = sp.csc_matrix((X0, (I1, J1)), shape=(n, n))
        self._S0 = sp.csc_matrix((S0, (I1, J1)), shape=(n, n))
        self._y0 = y0
        self._pname = "band_n%i_m%i_bw%i" % (n, m, bw)

    @property
    def bw(self):
        return self._bw


class mtxnorm_SDP(SDP):
    """Matrix-norm minimisation  min ||F(x) + G||_2  as an SDP of order n = p+q
    (``base.py:639-778``): the p x q block (rows q..n-1, columns 0..q-1) holds G (column 0,
    dense N(0,1)) and F_i (columns 1..r, ``nz = round(density*p*q)`` random positions,
    N(0,1)); the last column is -I and b = (0,...,0,-1), so m = r+1."""

    def __init__(self, p, q, r, density=1.0, seed=0):
        SDP.__init__(self)
        if type(seed) is not int:
            raise ValueError("seed must be an integer")
        # density: one float for all F_i or a list with one float per F_i (base.py:718-760)
        if type(density) is float and 0.0 < density <= 1.0:
            dens = [density] * r
        elif isinstance(density, (list, tuple)) and len(density) == r and all(0.0 < float(v) <= 1.0 for v in density):
            dens = [float(v) for v in density]
        else:
            raise TypeError("density must be a float between 0 and 1 or a list of r such floats")
        rng = np.random.default_rng(seed)
        n = p + q
        self._p, self._q = p, q
        I1 = np.tile(np.arange(q, n, dtype=np.int64), q)
        J1 = np.repeat(np.arange(q, dtype=np.int64), p)
        Il = misc.sub2ind((n, n), I1, J1)
        rows = [Il]
        cols = [np.zeros(p * q, dtype=np.int64)]
        vals = [rng.standard_normal(p * q)]
        for j in range(1, r + 1):
            nz = min(max(1, int(round(dens[j - 1] * p * q))), p * q)
            sel = Il if nz == p * q else rng.choice(Il, size=nz, replace=False)
            rows.append(sel)
            cols.append(np.full(nz, j, dtype=np.int64))
            vals.append(rng.standard_normal(nz))
        rows.append(np.arange(0, n * n, n + 1, dtype=np.int64))
        cols.append(np.full(n, r + 1, dtype=np.int64))
        vals.append(-np.ones(n))
        self._A = misc.as_csc(sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                                            shape=(n * n, r + 2)))
        self._b = np.zeros(r + 1)
        self._b[-1] = -1.0
        self._pname = "mtxnorm_p%i_q%i_r%i_d%s" % (p, q, r, density if type(density) is float else "list")


class rand_SDP(SDP):
    """Random SDP on an arbitrary pattern, ``rand_SDP(V, m, density=1.0, seed=0)``
    (``base.py:864-949``): ``V`` is an n x n sparse matrix (its lower triangle is the
    pattern); each A_i has ``nz = round(density*|V|)`` entries at random pattern positions;
    A_0 = S0 + sum y0_k A_k, b_k = A_k . X0 — strictly feasible by construction."""

    def __init__(self, V, m, density=1.0, seed=0):
        SDP.__init__(self)
        if type(seed) is not int:
            raise ValueError("seed must be an integer")
        rng = np.random.default_rng(seed)
        V = sp.tril(sp.csc_matrix(V), format="coo")
        n = V.shape[0]
        cp, ri = lower_pattern(n, V.row, V.col)
        J1 = _pattern_cols(n, cp)
        I1 = ri
        N = len(I1)
        Il = misc.sub2ind((n, n), I1, J1)
        diag = I1 == J1
        y0 = rng.standard_normal(m)
        y0 /= np.linalg.norm(y0)
        S0 = mk_rand(n, cp, ri, "posdef", seed)
        X0 = mk_rand(n, cp, ri, "completable", seed)
        # density: one float for all constraints or a list with one float per constraint
        # (``base.py:903-931``)
        if isinstance(density, float):
            nzs = [min(max(1, int(round(density * N))), N)] * m
        elif isinstance(density, (list, tuple)) and len(density) == m:
            nzs = [min(max(1, int(round(float(v) * N))), N) for v in density]
        else:
            raise TypeError("density must be a float or a list of m floats")
        pos = [np.arange(N) if nz == N else rng.choice(N, size=nz, replace=False) for nz in nzs]
        vals = [rng.standard_normal(nz) for nz in nzs]
        x = X0.copy()
        x[diag] *= 0.5
        a0 = S0.copy()
        self._b = np.empty(m)
        for j in range(m):
            np.add.at(a0, pos[j], vals[j] * y0[j])
            self._b[j] = 2.0 * np.dot(vals[j], x[pos[j]])
        pos_all = np.concatenate(pos) if m else np.zeros(0, dtype=np.int64)
        rows = np.concatenate([Il, Il[pos_all]])
        cols = np.concatenate([np.zeros(N, dtype=np.int64), np.repeat(np.arange(1, m + 1, dtype=np.int64), nzs)])
        data = np.concatenate([a0] + vals)
        self._A = misc.as_csc(sp.csc_matrix((data, (rows, cols)), shape=(n * n, m + 1)))
        self._X0 = sp.csc_matrix((X0, (I1, J1)), shape=(n, n))
        self._S0 = sp.csc_matrix((S0, (I1, J1)), shape=(n, n))
        self._y0 = y0
        self._pname = "rand_n%i_m%i" % (n, m)


class maxcut_SDP(SDP):
    """Max-cut relaxation  min C.X  s.t. X_ii = 1, X >= 0  with C = -(1/4) Laplacian of a
    graph given by its edge list (the SDPLIB ``maxG*``/``mcp*`` family the reference
    benchmarks read from ``dat-s`` files; ``doc/source/benchmarks/index.rst``)."""

    def __init__(self, n, edges, weights=None):
        SDP.__init__(self)
        edges = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
        e = edges[edges[:, 0] != edges[:, 1]]
        i = np.maximum(e[:, 0], e[:, 1])
        j = np.minimum(e[:, 0], e[:, 1])
        w = np.ones(len(e)) if weights is None else np.asarray(weights, dtype=np.float64)
        Lap = sp.coo_matrix((np.concatenate([-w, w, w]),
                             (np.concatenate([i, i, j]), np.concatenate([j, i, j]))), shape=(n, n)).tocsc()
        Lap.sum_duplicates()
        C = (-0.25 * sp.tril(Lap)).tocoo()
        d = np.arange(n, dtype=np.int64)
        rows = np.concatenate([misc.sub2ind((n, n), C.row, C.col), d * (n + 1)])
        cols = np.concatenate([np.zeros(len(C.row), dtype=np.int64), d + 1])
        vals = np.concatenate([C.data, np.ones(n)])
        self._A = misc.as_csc(sp.csc_matrix((vals, (rows, cols)), shape=(n * n, n + 1)))
        self._b = np.ones(n)
        self._pname = "maxcut_n%i" % n


def completion(X):
    """Maximum-determinant positive definite completion of the sparse symmetric matrix ``X``
    (lower triangle used) as a dense n x n array; raises ``ArithmeticError`` if none exists
    (``base.py:952-973``: ``symbolic`` under a fill-reducing ordering, ``chompack.completion``,
    then two ``chompack.trsm`` on the identity: X_hat = L^{-T} L^{-1})."""
    from .chordal import cspmatrix
    from . import chordal
    X = sp.tril(sp.csc_matrix(X), format="coo")
    n = X.shape[0]
    cp, ri = lower_pattern(n, X.row, X.col)
    p = maxcardsearch(n, cp, ri)
    fc, fr, _ = embed(n, cp, ri, p)
    if fc[-1] != cp[-1]:
        p = min_degree(n, cp, ri)
        fc, fr, _ = embed(n, cp, ri, p)
    symb = Symbolic(n, fc, fr)
    ops = solvers._make_backend(symb)
    lo = sp.tril(X, format="csc")
    full = (lo + sp.tril(lo, -1).T).tocsr()
    L = cspmatrix.from_vec(ops, np.asarray(full[p[symb.Ip], p[symb.Jp]]).ravel())
    chordal.completion(L)                       # ArithmeticError if X has no PD completion
    Z = ops.trsm(L.buf, np.eye(n), "N")
    Z = ops.trsm(L.buf, Z, "T")                 # internal order: vertex p[perm[i]] at position i
    order = p[symb.perm]
    out = np.empty((n, n))
    out[np.ix_(order, order)] = Z
    return out
