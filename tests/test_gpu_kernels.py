"""Parity of every CUDA kernel family with the CPU oracle, through the C ABI
(``smcp_b200.device.DeviceBackend`` -> ``libsmcp_b200.so``).  FP64 throughout; tolerance
1e-11 relative (Frobenius) unless stated — far inside the north star's 1e-8."""
import numpy as np
import pytest

from conftest import PATTERNS, make_symbolic, random_pd

pytestmark = pytest.mark.gpu

TOL = 1e-11


def relerr(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def _cases():
    out = []
    for p in PATTERNS:
        out.append(p + ("auto",))
        if p[3] >= 7 or p[:3] in ((30, 0, 3), (12, 0, 0)):
            # tiny-clique patterns also run through the CTA-per-supernode kernels, and with a
            # task partition that cuts the tree into many tasks (cross-warp hand-off)
            out.append(p + ("cta",))
            out.append(p + ("split",))
        if p[3] in (0, 1, 3, 6, 14):
            # general trees: also with the dense multi-CTA path (bigfront.cu) forced on every
            # supernode with >= 3 rows (and all their ancestors), separators included ...
            out.append(p + ("bigtop3",))
        if p[3] in (1, 5):
            # the independent top-set operations issued on three concurrent lanes (streams)
            out.append(p + ("bigtop3lanes",))
        if p[3] in (0, 1, 6, 14):
            # ... and with a threshold that splits the tree between the two code paths
            out.append(p + ("bigtop6",))
        if p[1] == 0 and p[2] >= 1:
            # band patterns: chain clique trees take the segment-parallel kernels by default;
            # keep the warp-per-chain sweeps covered on them as well
            out.append(p + ("nochain",))
    return out


@pytest.fixture(scope="module", params=_cases(), ids=lambda p: "n%d_e%d_bw%d_%s" % (p[0], p[1], p[2], p[4]))
def setup(request):
    import os
    from smcp_b200.device import DeviceBackend
    from oracle import supernodal as sn
    n, ne, bw, seed, mode = request.param
    symb = make_symbolic(n, ne, bw, seed)
    if mode == "cta":
        os.environ["SMCP_B200_NO_SMALL"] = "1"
    if mode == "nochain":
        os.environ["SMCP_B200_NO_CHAIN"] = "1"
    if mode.startswith("bigtop"):
        os.environ["SMCP_B200_BIG_NJ"] = mode[6:].replace("lanes", "")
        os.environ["SMCP_B200_LANES"] = "3" if mode.endswith("lanes") else "1"

    try:
        dev = DeviceBackend(symb, small_work=0 if mode == "split" else 2000)
    finally:
        os.environ.pop("SMCP_B200_NO_SMALL", None)
        os.environ.pop("SMCP_B200_NO_CHAIN", None)
        os.environ.pop("SMCP_B200_BIG_NJ", None)
        os.environ.pop("SMCP_B200_LANES", None)
    s = random_pd(symb, seed)
    l = s.copy()
    sn.cholesky(symb, l)
    y = l.copy()
    sn.projected_inverse(symb, y)
    return symb, dev, s, l, y


def test_roundtrip_vec(setup):
    symb, dev, s, l, y = setup
    v = np.random.default_rng(0).standard_normal(symb.nvp)
    b = dev.from_vec(v)
    assert np.array_equal(dev.to_vec(b), v)
    blk = dev.get_blk(b)
    assert np.array_equal(blk[symb.vec2blk], v)
    assert np.count_nonzero(blk) <= symb.nvp


def test_cholesky(setup):
    from oracle import supernodal as sn
    symb, dev, s, l, y = setup
    b = dev.set_blk(s)
    dev.cholesky(b)
    assert relerr(dev.get_blk(b), l) < TOL
    assert abs(dev.sumlogdiag(b) - sn.sumlogdiag(symb, l)) < 1e-10 * max(1, symb.n)


def test_cholesky_not_pd(setup):
    symb, dev, s, l, y = setup
    bad = s.copy()
    bad[symb.diag_blk[symb.n // 2]] = -1.0
    b = dev.set_blk(bad)
    with pytest.raises(ArithmeticError):
        dev.cholesky(b)
    nanm = s.copy()
    nanm[symb.diag_blk[0]] = np.nan
    with pytest.raises(ArithmeticError):
        dev.cholesky(dev.set_blk(nanm))


def test_llt(setup):
    symb, dev, s, l, y = setup
    b = dev.set_blk(l)
    dev.llt(b)
    assert relerr(dev.get_blk(b), s) < TOL


def test_projected_inverse(setup):
    symb, dev, s, l, y = setup
    b = dev.set_blk(l)
    dev.projected_inverse(b)
    assert relerr(dev.get_blk(b), y) < TOL


def test_trsm(setup):
    """chompack.trsm (solvers.py:491-492, 1921-1922): dense right-hand sides against the supernodal
    factor, both directions, against the oracle and against a dense triangular solve."""
    from oracle import supernodal as sn
    symb, dev, s, l, y = setup
    rng = np.random.default_rng(3)
    n = symb.n
    B = rng.standard_normal((n, 5))
    Lb = dev.set_blk(l)
    for trans in ("N", "T"):
        ref = B.copy()
        sn.trsm(symb, l, ref, trans)
        got = dev.trsm(Lb, B, trans)
        assert relerr(got, ref) < 1e-10


def test_trsm_many_rhs(setup):
    """The dense inverse of the sparse-constraint technique solves with n right-hand sides: warp-per-column sweep over
    the small supernodes (trsm_warp_kernel), runs of thin top-set supernodes in one launch (thin_trsm_run_kernel)."""
    from oracle import supernodal as sn
    symb, dev, s, l, y = setup
    rng = np.random.default_rng(4)
    n = symb.n
    B = rng.standard_normal((n, 70))
    Lb = dev.set_blk(l)
    for trans in ("N", "T"):
        ref = B.copy()
        sn.trsm(symb, l, ref, trans)
        got = dev.trsm(Lb, B, trans)
        assert relerr(got, ref) < 1e-10


def test_completion(setup):
    from oracle import supernodal as sn
    symb, dev, s, l, y = setup
    b = dev.set_blk(y)
    dev.completion(b)
    lc = y.copy()
    sn.completion(symb, lc)
    assert relerr(dev.get_blk(b), lc) < 1e-9          # conditioning of the completion map
    assert relerr(dev.get_blk(b), l) < 1e-9           # completion(P(S^-1)) recovers chol(S)
    bad = y.copy()
    bad[symb.diag_blk[symb.n // 2]] = -1.0
    with pytest.raises(ArithmeticError):
        dev.completion(dev.set_blk(bad))


def test_dot_axpy_scal(setup):
    from oracle import supernodal as sn
    symb, dev, s, l, y = setup
    bs, by = dev.set_blk(s), dev.set_blk(y)
    assert abs(dev.dot(bs, by) - sn.dot(symb, s, y)) <= 1e-12 * abs(sn.dot(symb, s, y)) + 1e-300
    dev.axpy(0.37, bs, by)
    assert np.array_equal(dev.get_blk(by), y + 0.37 * s)     # mul then add, no FMA: bit-exact
    dev.scal(-1.7, by)
    assert np.array_equal(dev.get_blk(by), (y + 0.37 * s) * -1.7)


@pytest.mark.parametrize("batch", [1, 5])
def test_hessian_forward_inverse(setup, batch):
    from oracle import supernodal as sn
    symb, dev, s, l, y = setup
    rng = np.random.default_rng(11)
    hf = sn.HessianFactor(symb, l, y)
    tok = dev.hessian_factor(dev.set_blk(l), dev.set_blk(y))
    U = rng.standard_normal((batch, symb.nblk)) * (symb.wdot > 0)
    W = U.copy()
    sn.hessian(hf, W)
    bufs = [dev.set_blk(u) for u in U]
    dev.hessian_apply(tok, bufs, False)
    got = np.array([dev.get_blk(b) for b in bufs])
    assert relerr(got, W) < TOL
    dev.hessian_apply(tok, bufs, True)
    back = np.array([dev.get_blk(b) for b in bufs])
    Wi = W.copy()
    sn.hessian_inv(hf, Wi)
    assert relerr(back, Wi) < 1e-9
    assert relerr(back, U) < 1e-8


@pytest.mark.parametrize("inv,adj", [(False, False), (False, True), (True, True), (True, False)])
def test_hessian_half_factors(setup, inv, adj):
    """chompack.hessian(L, Y, U, adj=False/True, inv=...) (solvers.py:917, 978, 1121, 1126): the half
    factors G, G^adj, G^-adj, G^-1 against the oracle, the identity hessian = G^adj o G, and the Newton
    decrement ||G(u)||^2 = u . hessian(u) that the drivers evaluate through hessian_norm()."""
    from oracle import supernodal as sn
    symb, dev, s, l, y = setup
    rng = np.random.default_rng(13)
    hf = sn.HessianFactor(symb, l, y)
    tok = dev.hessian_factor(dev.set_blk(l), dev.set_blk(y))
    U = rng.standard_normal((2, symb.nblk)) * (symb.wdot > 0)
    W = U.copy()
    sn.hessian_half(hf, W, adj, inv)
    bufs = [dev.set_blk(u) for u in U]
    dev.hessian_apply(tok, bufs, inv, adj)
    got = np.array([dev.get_blk(b) for b in bufs])
    assert relerr(got * (symb.wdot > 0), W * (symb.wdot > 0)) < 1e-10
    # second half: the composition is the full (inverse) Hessian
    dev.hessian_apply(tok, bufs, inv, not adj)
    full = np.array([dev.get_blk(b) for b in bufs])
    F = U.copy()
    (sn.hessian_inv if inv else sn.hessian)(hf, F)
    if adj == inv:      # forward: G then G^adj; inverse: G^-adj then G^-1
        assert relerr(full * (symb.wdot > 0), F * (symb.wdot > 0)) < 1e-9
        for k in range(2):
            n2 = sn.dot(symb, got[k], got[k])
            assert abs(n2 - sn.dot(symb, U[k], F[k])) <= 1e-9 * abs(n2)


def test_probe_batch(setup):
    symb, dev, s, l, y = setup
    rng = np.random.default_rng(5)
    d = rng.standard_normal(symb.nblk) * (symb.wdot > 0)
    gam = np.array([0.0, 1e-3, 0.1, 0.5, 1.0, 4.0, 50.0])
    ok, sld = dev.probe("cholesky", dev.set_blk(s), dev.set_blk(d), gam)
    from oracle import supernodal as sn
    for g, o, v in zip(gam, ok, sld):
        t = s + g * d
        try:
            sn.cholesky(symb, t)
            assert o and abs(v - sn.sumlogdiag(symb, t)) < 1e-9 * max(1, symb.n)
        except ArithmeticError:
            assert not o
    ok2, _ = dev.probe("completion", dev.set_blk(y), dev.set_blk(d), gam * 1e-2)
    for g, o in zip(gam * 1e-2, ok2):
        t = y + g * d
        try:
            sn.completion(symb, t)
            assert o
        except ArithmeticError:
            assert not o


@pytest.mark.parametrize("batch", [3, 40])
def test_contiguous_batch(setup, batch):
    """cholesky / llt / projected_inverse / completion / Hessian on `batch` matrices stored
    contiguously (the layout of the Schur-complement assembly and of the probes)."""
    from oracle import supernodal as sn
    symb, dev, s, l, y = setup
    rng = np.random.default_rng(3)
    shifts = 1.0 + rng.random(batch)
    S = np.array([s + sh * (symb.wdot == 1.0) for sh in shifts])      # s + sh*I, still PD
    Ls = S.copy()
    for r in Ls:
        sn.cholesky(symb, r)
    buf = dev.alloc_batch(batch)
    dev.set_batch(buf, S)
    info = dev.cholesky_batch(buf, batch)
    assert not info.any()
    assert relerr(dev.get_batch(buf, batch), Ls) < TOL
    dev.llt_batch(buf, batch)
    assert relerr(dev.get_batch(buf, batch), S) < TOL
    dev.set_batch(buf, Ls)
    dev.projected_inverse_batch(buf, batch)
    Ys = Ls.copy()
    for r in Ys:
        sn.projected_inverse(symb, r)
    assert relerr(dev.get_batch(buf, batch), Ys) < TOL
    info = dev.completion_batch(buf, batch)
    assert not info.any()
    assert relerr(dev.get_batch(buf, batch), Ls) < 1e-9
    # Hessian at the scaling point (l, y) applied to the whole batch in one call
    hf = sn.HessianFactor(symb, l, y)
    tok = dev.hessian_factor(dev.set_blk(l), dev.set_blk(y))
    U = rng.standard_normal((batch, symb.nblk)) * (symb.wdot > 0)
    W = U.copy()
    sn.hessian(hf, W)
    dev.set_batch(buf, U)
    dev.hessian_batch(tok, buf, batch, False)
    assert relerr(dev.get_batch(buf, batch), W) < TOL
    dev.hessian_batch(tok, buf, batch, True)
    assert relerr(dev.get_batch(buf, batch), U) < 1e-8
    # one indefinite matrix in the middle of the batch is reported, the others are not
    bad = S.copy()
    bad[batch // 2, symb.diag_blk[symb.n // 2]] = -1.0
    dev.set_batch(buf, bad)
    info = dev.cholesky_batch(buf, batch)
    assert info[batch // 2] != 0 and np.count_nonzero(info) == 1
    dev.free_batch(buf)
