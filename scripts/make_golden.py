"""Generate the golden fixtures under tests/golden/ (run in the build container, where
/root/reference exists; the fixtures are committed because the reference cannot travel to
the GPU box).

    make -C oracle            # builds oracle/_ref from /root/reference/src/C/misc.c
    python scripts/make_golden.py

Three families:

1. ``misc_ref_*.npz``  — inputs and outputs of the REFERENCE'S OWN compiled code
   (oracle/_ref/misc.so = /root/reference/src/C/misc.c, unmodified, called through
   oracle/ref.py): nzcolumns, matperm, ind2sub/sub2ind, phase1_sdp, Av_to_spmatrix,
   scal_diag and SCMcolumn2 (the one arithmetic kernel of the hot path that lives in the
   reference tree, misc.c:620-663).  The SCMcolumn2 fixtures carry a complete small
   problem (pattern, scaling point S, Av, Ns) plus the Schur columns the reference code
   produced from V = columns of S^{-1} (dense numpy inverse), so that the CUDA path can be
   checked end to end against reference output.
2. ``dense_*.npz``     — dense-NumPy ground truth (oracle/dense.py) of every chordal
   kernel on fixed patterns: cholesky, projected inverse, completion input/output pair,
   llt, Hessian, inverse Hessian, Schur complement.
3. ``driver_*.json``   — per-iteration traces of the restated drivers on the CPU oracle
   backend for the reference's own test problem (tests/test_basic.py:9-19) and a small
   band SDP; the GPU driver parity tests compare iteration counts and objectives with them.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "tests", "golden")


def rand_A(n, m, rng, kmin=2, kmax=8, dense_cols=()):
    rows, cols, vals = [], [], []
    for c in range(m + 1):
        k = n * n if c in dense_cols else int(rng.integers(kmin, kmax))
        I, J = rng.integers(0, n, k), rng.integers(0, n, k)
        r = np.unique(np.maximum(I, J) + n * np.minimum(I, J))
        rows += list(r)
        cols += [c] * len(r)
        vals += list(rng.standard_normal(len(r)))
    from smcp_b200 import misc
    return misc.as_csc(sp.csc_matrix((vals, (rows, cols)), shape=(n * n, m + 1)))


def gen_misc(ref):
    rng = np.random.default_rng(20261017)
    out = {}
    for t, (n, m) in enumerate([(7, 5), (12, 9), (30, 17)]):
        A = rand_A(n, m, rng, dense_cols=(2,) if t else ())
        nz = ref.nzcolumns(A)
        pm, Ns = ref.matperm(nz, int(0.3 * n))
        u = rng.standard_normal(m)
        P1 = ref.phase1_sdp(A, u)
        ind = rng.integers(0, n * n, 25)
        I, J = ref.ind2sub(n, ind)
        out.update({
            "A%d_data" % t: A.data, "A%d_indices" % t: A.indices, "A%d_indptr" % t: A.indptr,
            "A%d_shape" % t: np.array(A.shape), "nz%d" % t: nz, "Nmax%d" % t: np.array(int(0.3 * n)),
            "pm%d" % t: pm, "Ns%d" % t: np.array(Ns), "u%d" % t: u,
            "P%d_data" % t: P1.data, "P%d_indices" % t: P1.indices, "P%d_indptr" % t: P1.indptr,
            "P%d_shape" % t: np.array(P1.shape), "ind%d" % t: ind, "I%d" % t: I, "J%d" % t: J,
            "lin%d" % t: ref.sub2ind((n, n), I, J),
        })
    out["ncases"] = np.array(3)
    np.savez_compressed(os.path.join(OUT, "misc_ref_index.npz"), **out)


def gen_scm(ref):
    """Complete small problems for the sparse-constraint technique (solvers.py:489-497)."""
    import smcp_b200 as S
    from smcp_b200 import solvers
    from smcp_b200.solvers import _Problem, _read_options
    from oracle.backend import OracleBackend
    from oracle import dense as dn
    solvers.options["show_progress"] = False
    solvers.set_backend_factory(lambda symb: OracleBackend(symb))
    rng = np.random.default_rng(7)
    cases = []
    n = 40
    e = rng.integers(0, n, size=(70, 2))
    cases.append(("maxcut", S.maxcut_SDP(n, e)))
    n = 60
    e = rng.integers(0, n, size=(50, 2))
    V = sp.coo_matrix((np.ones(50 + n), (np.concatenate([e[:, 0], np.arange(n)]),
                                         np.concatenate([e[:, 1], np.arange(n)]))), shape=(n, n))
    cases.append(("randsparse", S.rand_SDP(V, 25, density=0.03, seed=4)))
    for name, P in cases:
        opt = _read_options(P.n, False)
        pr = _Problem(P.A, P.b, opt, "chol", None)
        symb, Av, m, Ns = pr.symb, pr.Av, pr.m, pr.Ns
        assert Ns > 0
        s = np.zeros(symb.nvp)
        s[symb.diag_vec] = 2.0
        s += 0.05 * rng.standard_normal(symb.nvp)
        # dense S in the Vp index space (rows/cols = Ip/Jp) and its inverse
        Sd = np.zeros((symb.n, symb.n))
        Sd[symb.Ip, symb.Jp] = s
        Sd = Sd + np.tril(Sd, -1).T
        Sinv = np.linalg.inv(Sd)
        H = np.zeros((m, m), order="F")
        md = m - Ns
        # technique 2 columns with the REFERENCE'S SCMcolumn2 (V = S^{-1}[:, K], Kl: vertex -> column)
        for j in range(md, m):
            c0, c1 = Av.indptr[j], Av.indptr[j + 1]
            rows_j = Av.indices[c0:c1]
            K = np.unique(np.concatenate([symb.Ip[rows_j], symb.Jp[rows_j]]))
            Vm = np.asfortranarray(Sinv[:, K])
            Kl = np.zeros(symb.n, dtype=np.int64)
            Kl[K] = np.arange(len(K))
            H = ref.SCMcolumn2(H, Av, Vm, symb.Ip, symb.Jp, Kl, j)
        # technique 1 columns (if any) from dense algebra: H_ij = A_i . S^-1 A_j S^-1
        if md:
            def mat(j):
                c0, c1 = Av.indptr[j], Av.indptr[j + 1]
                M = np.zeros((symb.n, symb.n))
                M[symb.Ip[Av.indices[c0:c1]], symb.Jp[Av.indices[c0:c1]]] = Av.data[c0:c1]
                return M + np.tril(M, -1).T
            for j in range(md):
                W = Sinv @ mat(j) @ Sinv
                for i in range(j, m):
                    H[i, j] = np.sum(mat(i) * W)
        np.savez_compressed(os.path.join(OUT, "misc_ref_scm_%s.npz" % name),
                            n=np.array(symb.n), vp_colptr=symb.colptr, vp_rowind=symb.rowind,
                            s_vec=s, Av_data=Av.data, Av_indices=Av.indices, Av_indptr=Av.indptr,
                            Av_shape=np.array(Av.shape), Ns=np.array(Ns), H_lower=np.tril(H))
        # Av_to_spmatrix + scal_diag on the same problem (misc.c:475-557)
        j = md
        Aj = ref.Av_to_spmatrix(Av, symb.Ip, symb.Jp, j, symb.n)
        Id = symb.diag_vec
        vals = np.arange(1.0, symb.nvp + 1.0)
        sc = ref.scal_diag(vals, symb.colptr, symb.rowind, (symb.n, symb.n), Id, 0.5)
        np.savez_compressed(os.path.join(OUT, "misc_ref_av_%s.npz" % name), j=np.array(j),
                            Aj_data=Aj.data, Aj_indices=Aj.indices, Aj_indptr=Aj.indptr, scal_in=vals,
                            scal_out=sc, Id=Id)
    solvers.set_backend_factory(None)


def gen_dense():
    from conftest import make_symbolic, random_pd
    from oracle import dense as dn
    cases = {"mixed": (40, 30, 1, 0), "band5": (60, 0, 5, 7), "chain": (36, -3, 6, 12), "tree": (50, 0, -3, 10)}
    for name, spec in cases.items():
        symb = make_symbolic(*spec)
        rng = np.random.default_rng(99)
        x = random_pd(symb, 5)
        l = dn.cholesky(symb, x)
        y = dn.projected_inverse(symb, l)
        u = rng.standard_normal((3, symb.nblk)) * (symb.wdot > 0)
        z = np.stack([dn.hessian(symb, l, ui) for ui in u])
        Us = rng.standard_normal((4, symb.nblk)) * (symb.wdot > 0)
        np.savez_compressed(os.path.join(OUT, "dense_%s.npz" % name), spec=np.array(spec),
                            vp_colptr=symb.colptr, vp_rowind=symb.rowind, x=x, chol=l, projinv=y,
                            llt=dn.llt(symb, l), u=u, hess=z, schur_U=Us, schur_H=dn.schur(symb, l, Us),
                            dot_xy=np.array(dn.dot(symb, x, y)))


def gen_drivers():
    import smcp_b200 as S
    from smcp_b200 import solvers
    from oracle.backend import OracleBackend
    solvers.options["show_progress"] = False
    solvers.set_backend_factory(lambda symb: OracleBackend(symb, batch_columns=32))
    traces = {}
    c = np.array([-6., -4., -5.])
    G = np.array([[16., 7., 24., -8., 8., -1., 0., -1., 0., 0., 7., -5., 1., -5., 1., -7., 1., -7., -4.],
                  [-14., 2., 7., -13., -18., 3., 0., 0., -1., 0., 3., 13., -6., 13., 12., -10., -6., -10., -28.],
                  [5., 0., -15., 12., -6., 17., 0., 0., 0., -1., 9., 6., -6., 6., -7., -7., -6., -7., -11.]]).T
    h = np.array([-3., 5., 12., -2., -14., -13., 10., 0., 0., 0., 68., -30., -19., -30., 99., 23., -19., 23., 10.])
    sol = solvers.conelp(c, G, h, {'l': 2, 'q': [4, 4], 's': [3]})
    traces["conelp_test_basic"] = {"status": sol["status"], "iterations": sol["iterations"],
                                   "x": [float(v) for v in np.asarray(sol["x"]).ravel()],
                                   "primal objective": sol["primal objective"], "dual objective": sol["dual objective"]}
    for name, P, kw, method in [
            ("band_feas_primal", S.band_SDP(60, 20, 3, seed=7), {"scaling": "primal"}, "feas"),
            ("band_feas_dual", S.band_SDP(60, 20, 3, seed=7), {"scaling": "dual"}, "feas"),
            ("band_esd", S.band_SDP(30, 10, 2, seed=1), {}, "esd"),
            ("mtxnorm_esd", S.mtxnorm_SDP(12, 4, 9, density=0.6, seed=1), {}, "esd")]:
        sol = getattr(P, "solve_" + method)(kktsolver="chol", **kw)
        traces[name] = {"status": sol["status"], "iterations": sol["iterations"],
                        "primal objective": sol["primal objective"], "dual objective": sol["dual objective"],
                        "gap": sol["gap"], "y": [float(v) for v in np.asarray(sol["y"]).ravel()],
                        # per-iteration statistics: the self-dual embedding is compared iterate by iterate down
                        # to the rounding floor of its Newton solves (see tests/test_golden.py)
                        "trace": [[float(r[k]) for k in ("pcost", "dcost", "gap", "pres", "dres")] for r in sol["trace"]
                                  if all(r.get(k) is not None for k in ("pcost", "dcost", "gap", "pres", "dres"))]}
    solvers.set_backend_factory(None)
    with open(os.path.join(OUT, "driver_traces.json"), "w") as f:
        json.dump(traces, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    from oracle import ref
    if not ref.available():
        raise SystemExit("oracle/_ref is not built: run `make -C oracle` (needs /root/reference)")
    if sys.argv[1:] != ["drivers"]:        # `make_golden.py drivers` regenerates driver_traces.json only
        gen_misc(ref)
        gen_scm(ref)
        gen_dense()
    gen_drivers()
    print("golden fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print("  %-40s %8d bytes" % (f, os.path.getsize(os.path.join(OUT, f))))
