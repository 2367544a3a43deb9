"""Segment-parallel chain kernels (csrc/chordal_chain.cuh) on long band patterns: several
segments per sweep, single matrix and batch, well- and ill-conditioned scaling points, against
the CPU oracle and against the sequential warp sweeps of the same library."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _band_symb(n, bw):
    from smcp_b200.symbolic import Symbolic, lower_pattern
    I = np.concatenate([np.arange(j, min(j + bw + 1, n)) for j in range(n)])
    J = np.concatenate([np.full(min(j + bw + 1, n) - j, j) for j in range(n)])
    cp, ri = lower_pattern(n, I, J)
    return Symbolic(n, cp, ri)


def _scaling_point(symb, n, bw, eps, seed):
    """banded PSD sum of window rank-ones + eps*I: cond ~ 1/eps"""
    rng = np.random.default_rng(seed)
    S = np.zeros((n, n))
    for k in range(n - bw):
        if rng.random() < 0.5:
            v = rng.standard_normal(bw + 1)
            S[k:k + bw + 1, k:k + bw + 1] += 0.2 * np.outer(v, v)
    S += eps * np.eye(n)
    return S[symb.Ip, symb.Jp]       # vector-space order (band patterns keep the natural order)


def _rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("n,bw,eps,tol", [(700, 5, 1e-1, 1e-12), (1500, 5, 1e-6, 1e-9), (400, 1, 1e-2, 1e-12),
                                          (333, 7, 1e-3, 1e-11), (2100, 3, 1e-9, 1e-7), (40, 5, 1e-1, 1e-12)])
def test_chain_hessian_matches_oracle_and_sequential(n, bw, eps, tol):
    from smcp_b200.device import DeviceBackend
    from oracle import supernodal as sn
    symb = _band_symb(n, bw)
    assert np.array_equal(symb.perm, np.arange(n))
    svec = _scaling_point(symb, n, bw, eps, n + bw)
    dev = DeviceBackend(symb)
    os.environ["SMCP_B200_NO_CHAIN"] = "1"
    try:
        seq = DeviceBackend(symb)
    finally:
        os.environ.pop("SMCP_B200_NO_CHAIN", None)
    rng = np.random.default_rng(1)
    U = rng.standard_normal((3, symb.nblk)) * (symb.wdot > 0)
    outs = {}
    for name, be in (("chain", dev), ("seq", seq)):
        L = be.from_vec(svec)
        be.cholesky(L)
        Y = be.clone(L)
        be.projected_inverse(Y)
        tok = be.hessian_factor(L, Y)
        res = []
        for k in range(U.shape[0]):
            u = be.set_blk(U[k])
            be.hessian_apply(tok, [u], False)
            fwd = be.get_blk(u)
            v = be.set_blk(U[k])
            be.hessian_apply(tok, [v], True)            # inverse map on a fresh input
            res.append((fwd, be.get_blk(v)))
        B = 40
        buf = be.alloc_batch(B)
        be.set_batch(buf, np.repeat(U[:1], B, axis=0))
        be.hessian_batch(tok, buf, B, False)
        bat = be.get_batch(buf, B)
        be.free_batch(buf)
        T = be.clone(L)
        be.llt(T)
        outs[name] = (res, bat, be.get_blk(T), be.get_blk(L), be.get_blk(Y))
    w = symb.wdot > 0
    # chain == sequential sweeps of the same library
    for k in range(U.shape[0]):
        assert _rel(outs["chain"][0][k][0] * w, outs["seq"][0][k][0] * w) < tol
        assert _rel(outs["chain"][0][k][1] * w, outs["seq"][0][k][1] * w) < 1e-12
    assert _rel(outs["chain"][1] * w, np.repeat(outs["chain"][0][0][0][None, :], 40, axis=0) * w) < 1e-13
    assert _rel(outs["chain"][2] * w, outs["seq"][2] * w) < 1e-13
    # chain == CPU oracle (supernodal restatement)
    Lo, Yo = outs["chain"][3].copy(), outs["chain"][4].copy()
    hf = sn.HessianFactor(symb, Lo, Yo)
    Uo = U.copy()
    sn.hessian(hf, Uo)
    for k in range(U.shape[0]):
        assert _rel(outs["chain"][0][k][0] * w, Uo[k] * w) < tol
    To = Lo.copy()[None, :]
    sn.llt(symb, To)
    assert _rel(outs["chain"][2] * w, To[0] * w) < 1e-12


def test_full_solve_converges_with_accuracy_guard():
    """Solve a band SDP to SMCP's default tolerances.  Late iterates have cond(S) > 1e10: the
    segment propagators grow and the segment-parallel sweeps lose digits (the solver used to stall
    at a primal residual of ~1e-5 on the benchmark problem); the guard in chain_prepare switches
    those scaling points to the sequential sweeps.  The result must match a run that uses the
    sequential sweeps throughout: same status, same iteration count, same objectives."""
    import smcp_b200 as S
    from smcp_b200 import solvers
    from smcp_b200.device import DeviceBackend
    solvers.options["show_progress"] = False
    solvers.options["maxiters"] = 100
    solvers.set_backend_factory(lambda symb: DeviceBackend(symb))
    sols = {}
    for mode in ("guarded", "sequential"):
        if mode == "sequential":
            os.environ["SMCP_B200_NO_CHAIN"] = "1"
        try:
            P = S.band_SDP(1500, 120, 5, seed=0)
            sols[mode] = P.solve_feas(kktsolver="chol", primalstart={"x": P._X0})
        finally:
            os.environ.pop("SMCP_B200_NO_CHAIN", None)
    solvers.set_backend_factory(None)
    a, b = sols["guarded"], sols["sequential"]
    assert a["status"] == "optimal" and b["status"] == "optimal", (a["status"], b["status"])
    assert abs(a["iterations"] - b["iterations"]) <= 1, (a["iterations"], b["iterations"])
    assert a["primal infeasibility"] <= 1e-8
    for key in ("primal objective", "dual objective"):
        assert abs(a[key] - b[key]) <= 1e-8 * max(1.0, abs(b[key])), (key, a[key], b[key])
