"""Golden-vector tests (fixtures: tests/golden/, generator: scripts/make_golden.py).

CPU (not gpu):  host restatements and the supernodal oracle against (a) output of the
                reference's own compiled C code (misc_ref_*), (b) dense ground truth (dense_*);
                when oracle/_ref is built, also live against the reference code on fresh inputs.
GPU          :  the CUDA path through the C ABI against the same fixtures.
"""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(GOLD, name))


def _csc(z, key):
    return sp.csc_matrix((z[key + "_data"], z[key + "_indices"], z[key + "_indptr"]), shape=tuple(z[key + "_shape"]))


# ---------------------------------------------------------------------------------------
# smcp.misc: our NumPy restatement == the reference's compiled code (bit-exact, integers/bytes)
# ---------------------------------------------------------------------------------------
def test_misc_index_helpers_match_reference_golden():
    from smcp_b200 import misc
    z = _load("misc_ref_index.npz")
    for t in range(int(z["ncases"])):
        A = _csc(z, "A%d" % t)
        n = int(round(np.sqrt(A.shape[0])))
        nz = misc.nzcolumns(A)
        assert np.array_equal(nz, z["nz%d" % t])
        pm, Ns = misc.matperm(nz, int(z["Nmax%d" % t]))
        assert np.array_equal(pm, z["pm%d" % t]) and Ns == int(z["Ns%d" % t])
        P = misc.phase1_sdp(A, z["u%d" % t])
        Pr = _csc(z, "P%d" % t)
        assert P.shape == Pr.shape
        assert np.array_equal(P.indptr, Pr.indptr) and np.array_equal(P.indices, Pr.indices)
        assert np.array_equal(P.data, Pr.data)
        I, J = misc.ind2sub(n, z["ind%d" % t])
        assert np.array_equal(I, z["I%d" % t]) and np.array_equal(J, z["J%d" % t])
        assert np.array_equal(misc.sub2ind((n, n), I, J), z["lin%d" % t])


def test_misc_live_against_reference_build():
    """Fresh random inputs through oracle/_ref/misc.so (the reference's misc.c, unmodified)."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built (no /root/reference)")
    from smcp_b200 import misc
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLD), "..", "scripts"))
    from make_golden import rand_A
    rng = np.random.default_rng(1234)
    for n, m in [(3, 1), (9, 6), (25, 31)]:
        A = rand_A(n, m, rng, dense_cols=(1,))
        nz = ref.nzcolumns(A)
        assert np.array_equal(misc.nzcolumns(A), nz)
        for nmax in (0, 2, n // 2, n + 1, -3):
            pr, Nr = ref.matperm(nz, nmax)
            pm, Ns = misc.matperm(nz, nmax)
            assert np.array_equal(pm, pr) and Ns == Nr
        u = rng.standard_normal(m)
        Pr, P = ref.phase1_sdp(A, u), misc.phase1_sdp(A, u)
        assert np.array_equal(P.indptr, Pr.indptr) and np.array_equal(P.indices, Pr.indices)
        assert np.array_equal(P.data, Pr.data)


@pytest.mark.parametrize("name", ["maxcut", "randsparse"])
def test_oracle_schur_sparse_technique_matches_reference_scmcolumn2(name):
    """The oracle's Schur complement (technique 2, oracle/backend.py) against the columns the
    reference's SCMcolumn2 produced (misc.c:620-663)."""
    from smcp_b200.symbolic import Symbolic
    from oracle.backend import OracleBackend
    z = _load("misc_ref_scm_%s.npz" % name)
    symb = Symbolic(int(z["n"]), z["vp_colptr"], z["vp_rowind"])
    ob = OracleBackend(symb, batch_columns=8)
    Av = _csc(z, "Av")
    ob.set_operator(Av, int(z["Ns"]))
    L = ob.from_vec(z["s_vec"])
    ob.cholesky(L)
    Y = ob.clone(L)
    ob.projected_inverse(Y)
    H = np.tril(ob.schur_assemble(ob.hessian_factor(L, Y)))
    Href = z["H_lower"]
    assert np.linalg.norm(H - Href) <= 1e-12 * np.linalg.norm(Href)


@pytest.mark.parametrize("name", ["maxcut", "randsparse"])
def test_av_to_spmatrix_and_scal_diag_semantics(name):
    """misc.Av_to_spmatrix / scal_diag (misc.c:475-557) are replaced on the device by a scatter
    of column j of Av into blkval and by the trace weights; check the scatter map against the
    reference's output: entry q of Av[:, j] lands at (Ip[q], Jp[q])."""
    from smcp_b200.symbolic import Symbolic
    z = _load("misc_ref_scm_%s.npz" % name)
    g = _load("misc_ref_av_%s.npz" % name)
    symb = Symbolic(int(z["n"]), z["vp_colptr"], z["vp_rowind"])
    Av = _csc(z, "Av")
    j = int(g["j"])
    c0, c1 = Av.indptr[j], Av.indptr[j + 1]
    M = sp.csc_matrix((Av.data[c0:c1], (symb.Ip[Av.indices[c0:c1]], symb.Jp[Av.indices[c0:c1]])),
                      shape=(symb.n, symb.n))
    M.sort_indices()
    assert np.array_equal(M.indptr, g["Aj_indptr"]) and np.array_equal(M.indices, g["Aj_indices"])
    assert np.array_equal(M.data, g["Aj_data"])
    v = g["scal_in"].copy()
    v[symb.diag_vec] *= 0.5
    assert np.array_equal(v, g["scal_out"])
    assert np.array_equal(symb.diag_vec, g["Id"])


# ---------------------------------------------------------------------------------------
# chordal kernels: supernodal oracle == dense ground truth fixtures
# ---------------------------------------------------------------------------------------
DENSE = ["mixed", "band5", "chain", "tree"]


def _dense_case(name):
    from smcp_b200.symbolic import Symbolic
    z = _load("dense_%s.npz" % name)
    symb = Symbolic(int(len(z["vp_colptr"]) - 1), z["vp_colptr"], z["vp_rowind"])
    assert symb.nblk == len(z["x"])
    return symb, z


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("name", DENSE)
def test_supernodal_oracle_matches_dense_golden(name):
    from oracle import supernodal as sn
    symb, z = _dense_case(name)
    w = symb.wdot > 0
    L = z["x"].copy()[None, :]
    sn.cholesky(symb, L)
    assert _rel(L[0] * w, z["chol"] * w) < 1e-12
    Y = L.copy()
    sn.projected_inverse(symb, Y)
    assert _rel(Y[0] * w, z["projinv"] * w) < 1e-11
    C = Y.copy()
    sn.completion(symb, C)                      # completion(P(S^-1)) = chol(S)
    assert _rel(C[0] * w, z["chol"] * w) < 1e-9
    T = L.copy()
    sn.llt(symb, T)
    assert _rel(T[0] * w, z["llt"] * w) < 1e-12
    hf = sn.HessianFactor(symb, L[0], Y[0])
    U = z["u"].copy()
    sn.hessian(hf, U)
    assert _rel(U * w, z["hess"] * w) < 1e-10
    sn.hessian_inv(hf, U)
    assert _rel(U * w, z["u"] * w) < 1e-8
    assert abs(sn.dot(symb, z["x"], z["projinv"]) - float(z["dot_xy"])) < 1e-9 * abs(float(z["dot_xy"]))


def test_driver_traces_oracle():
    """The restated drivers on the oracle backend reproduce the committed traces (guards the
    host control flow: iteration counts are exact, objectives to 1e-9)."""
    import smcp_b200 as S
    from smcp_b200 import solvers
    from oracle.backend import OracleBackend
    with open(os.path.join(GOLD, "driver_traces.json")) as f:
        tr = json.load(f)
    solvers.options["show_progress"] = False
    solvers.set_backend_factory(lambda symb: OracleBackend(symb, batch_columns=32))
    sol = S.band_SDP(60, 20, 3, seed=7).solve_feas(kktsolver="chol", scaling="primal")
    g = tr["band_feas_primal"]
    assert sol["status"] == g["status"] and sol["iterations"] == g["iterations"]
    assert abs(sol["primal objective"] - g["primal objective"]) <= 1e-9 * max(1, abs(g["primal objective"]))


# ---------------------------------------------------------------------------------------
# GPU: the CUDA path against the same fixtures
# ---------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", DENSE)
def test_cuda_kernels_match_dense_golden(name):
    from smcp_b200.device import DeviceBackend
    symb, z = _dense_case(name)
    dev = DeviceBackend(symb, small_work=2000)
    w = symb.wdot > 0
    L = dev.set_blk(z["x"])
    dev.cholesky(L)
    assert _rel(dev.get_blk(L) * w, z["chol"] * w) < 1e-12
    Y = dev.clone(L)
    dev.projected_inverse(Y)
    assert _rel(dev.get_blk(Y) * w, z["projinv"] * w) < 1e-11
    C = dev.clone(Y)
    dev.completion(C)
    assert _rel(dev.get_blk(C) * w, z["chol"] * w) < 1e-9
    T = dev.clone(L)
    dev.llt(T)
    assert _rel(dev.get_blk(T) * w, z["llt"] * w) < 1e-12
    tok = dev.hessian_factor(L, Y)
    for k in range(z["u"].shape[0]):
        U = dev.set_blk(z["u"][k])
        dev.hessian_apply(tok, [U], False)
        assert _rel(dev.get_blk(U) * w, z["hess"][k] * w) < 1e-10
        dev.hessian_apply(tok, [U], True)
        assert _rel(dev.get_blk(U) * w, z["u"][k] * w) < 1e-8
    X = dev.set_blk(z["x"])
    assert abs(dev.dot(X, Y) - float(z["dot_xy"])) < 1e-9 * abs(float(z["dot_xy"]))
    # batch of 64 copies (exercises the batched kernels of the Schur assembly)
    B = 64
    buf = dev.alloc_batch(B)
    dev.set_batch(buf, np.repeat(z["u"][:1], B, axis=0))
    dev.hessian_batch(tok, buf, B, False)
    out = dev.get_batch(buf, B)
    assert _rel(out * w, np.repeat(z["hess"][:1], B, axis=0) * w) < 1e-10
    dev.free_batch(buf)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["maxcut", "randsparse"])
def test_cuda_schur_matches_reference_scmcolumn2(name):
    """smcp_kkt_assemble (sparse-constraint technique on the device) against the Schur columns
    computed by the reference's SCMcolumn2 (misc.c:620-663)."""
    from smcp_b200.symbolic import Symbolic
    from smcp_b200.device import DeviceBackend
    z = _load("misc_ref_scm_%s.npz" % name)
    symb = Symbolic(int(z["n"]), z["vp_colptr"], z["vp_rowind"])
    dev = DeviceBackend(symb, small_work=2000)
    dev.set_operator(_csc(z, "Av"), int(z["Ns"]))
    L = dev.from_vec(z["s_vec"])
    dev.cholesky(L)
    Y = dev.clone(L)
    dev.projected_inverse(Y)
    dev.schur_assemble(dev.hessian_factor(L, Y))
    H = np.tril(dev.get_H())
    Href = z["H_lower"]
    assert np.linalg.norm(H - Href) <= 1e-12 * np.linalg.norm(Href)


@pytest.mark.gpu
def test_cuda_driver_traces():
    import smcp_b200 as S
    from smcp_b200 import solvers
    with open(os.path.join(GOLD, "driver_traces.json")) as f:
        tr = json.load(f)
    solvers.options["show_progress"] = False
    solvers.set_backend_factory(None)
    runs = {
        "band_feas_primal": lambda: S.band_SDP(60, 20, 3, seed=7).solve_feas(kktsolver="chol", scaling="primal"),
        "band_feas_dual": lambda: S.band_SDP(60, 20, 3, seed=7).solve_feas(kktsolver="chol", scaling="dual"),
        "band_esd": lambda: S.band_SDP(30, 10, 2, seed=1).solve_esd(kktsolver="chol"),
        "mtxnorm_esd": lambda: S.mtxnorm_SDP(12, 4, 9, density=0.6, seed=1).solve_esd(kktsolver="chol"),
    }
    for name, fn in runs.items():
        sol, g = fn(), tr[name]
        if "feas" in name:
            # feasible-start solver: the north star's bar -- same status, iteration count +-1, objectives 1e-8
            assert sol["status"] == g["status"], name
            assert abs(sol["iterations"] - g["iterations"]) <= 1, (name, sol["iterations"], g["iterations"])
            for key in ("primal objective", "dual objective"):
                assert abs(sol[key] - g[key]) <= 1e-8 * max(1.0, abs(g[key])), (name, key, sol[key], g[key])
            y = np.asarray(sol["y"]).ravel()
            assert np.linalg.norm(y - np.array(g["y"])) <= 1e-6 * max(1.0, np.linalg.norm(g["y"])), name
            continue
        # Self-dual embedding.  Root cause of its irreproducible exit (profiles/r02_esd_rootcause.md): from
        # the iteration at which pres / dres pass ~1e-7 the Newton equations are only solved to an ABSOLUTE
        # residual of ~1e-8 (options['debug'] prints it; the loss is in the Schur complement itself, an
        # extended-precision solve of the same H does not help), i.e. at the level of feastol = 1e-8, so
        # the last iterations are a random walk around the tolerance: on the CPU oracle 1e-15 of noise moves
        # the exit of band_esd from iteration 28 to 22 or beyond 100.  What IS reproducible, and pinned here:
        # every iterate down to that floor, and the optimum itself.
        gt = np.array(g["trace"])
        floor = next((k for k in range(len(gt)) if max(gt[k, 3], gt[k, 4]) < 3e-7), len(gt))
        st = np.array([[r[k] for k in ("pcost", "dcost", "gap", "pres", "dres")] for r in sol["trace"]
                       if all(r.get(k) is not None for k in ("pcost", "dcost", "gap", "pres", "dres"))])
        assert len(st) >= floor and floor >= 10, (name, len(st), floor)
        for k in range(floor):
            for c in (0, 1):
                assert abs(st[k, c] - gt[k, c]) <= 1e-6 * max(1.0, abs(gt[k, c])), (name, k, c, st[k, c], gt[k, c])
            for c in (2, 3, 4):
                assert abs(st[k, c] - gt[k, c]) <= 1e-3 * gt[k, c] + 1e-10, (name, k, c, st[k, c], gt[k, c])
        if sol["status"] == "optimal":
            # two exits inside the tolerance box (abstol = reltol = 1e-6) agree to the box, not better
            for key in ("primal objective", "dual objective"):
                assert abs(sol[key] - g[key]) <= 2e-6 * max(1.0, abs(g[key])), (name, key, sol[key], g[key])
            y = np.asarray(sol["y"]).ravel()
            assert np.linalg.norm(y - np.array(g["y"])) <= 1e-5 * max(1.0, np.linalg.norm(g["y"])), name
        else:
            # did not land inside the tolerance box: it must at least have reached the floor at the optimum
            best = min(range(len(st)), key=lambda k: max(st[k, 3], st[k, 4]))
            assert max(st[best, 3], st[best, 4]) <= 1e-7, (name, st[best])
            assert abs(st[best, 0] - g["primal objective"]) <= 1e-6 * max(1.0, abs(g["primal objective"])), name
