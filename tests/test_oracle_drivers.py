"""CPU tests: the drivers on the oracle backend (host logic + oracle numerics)."""
import numpy as np
import pytest

c = np.array([-6., -4., -5.])
G = np.array([[16., 7., 24., -8., 8., -1., 0., -1., 0., 0., 7., -5., 1., -5., 1., -7., 1., -7., -4.],
              [-14., 2., 7., -13., -18., 3., 0., 0., -1., 0., 3., 13., -6., 13., 12., -10., -6., -10., -28.],
              [5., 0., -15., 12., -6., 17., 0., 0., 0., -1., 9., 6., -6., 6., -7., -7., -6., -7., -11.]]).T
h = np.array([-3., 5., 12., -2., -14., -13., 10., 0., 0., 0., 68., -30., -19., -30., 99., 23., -19., 23., 10.])
CONELP = (c, G, h, {'l': 2, 'q': [4, 4], 's': [3]})


@pytest.fixture
def oracle_backend():
    from smcp_b200 import solvers
    from oracle.backend import OracleBackend
    solvers.options["show_progress"] = False
    solvers.set_backend_factory(lambda symb: OracleBackend(symb, batch_columns=32))
    yield


def test_conelp_known_answer(oracle_backend):
    """Data of the reference's own test (tests/test_basic.py:9-19); the optimum is the one
    printed in the CVXOPT user guide for this cone LP."""
    from smcp_b200 import solvers
    sol = solvers.conelp(*CONELP)
    assert sol["status"] == "optimal"
    assert np.allclose(sol["x"], [-1.22091527, 0.09663315, 3.57750167], atol=2e-6)
    assert abs(sol["primal objective"] - sol["dual objective"]) < 1e-5


def test_structural_doc_pins(oracle_backend):
    """The structural known answers of the docs (doc/source/documentation/index.rst:593-596,
    605-608): <SDP: n=100, m=100, nnz=297> and mtxnorm(200,10,200) -> n=210, m=201, nnz=2210."""
    import smcp_b200 as S
    P = S.band_SDP(100, 100, 2, seed=10)
    assert (P.n, P.m, P.nnz) == (100, 100, 297)
    Q = S.mtxnorm_SDP(200, 10, 200, seed=0)
    assert (Q.n, Q.m, Q.nnz) == (210, 201, 2210)


@pytest.mark.parametrize("scaling", ["primal", "dual"])
def test_feas_band(oracle_backend, scaling):
    import smcp_b200 as S
    P = S.band_SDP(30, 10, 2, seed=1)
    sol = P.solve_feas(scaling=scaling)
    assert sol["status"] == "optimal"
    assert sol["primal infeasibility"] < 1e-8 and sol["dual infeasibility"] < 1e-8
    assert abs(sol["primal objective"] - sol["dual objective"]) < 1e-5
    # known strictly feasible pair => optimal value between the generator's bounds
    X, Sm, y = sol["x"], sol["s"], sol["y"]
    n = P.n
    A = P.A
    Cm = P.get_A(0)
    Cfull = Cm + sp_tril_t(Cm)
    assert abs((Cfull.multiply(X)).sum() - sol["primal objective"]) < 1e-8
    # dual feasibility: C - sum y_i A_i = S
    R = Cfull.copy()
    for i in range(P.m):
        Ai = P.get_A(i + 1)
        R = R - y[i] * (Ai + sp_tril_t(Ai))
    assert abs(R - Sm).max() < 1e-7
    assert np.linalg.eigvalsh(Sm.toarray()).min() > -1e-9


def sp_tril_t(M):
    import scipy.sparse as sp
    return sp.tril(M, -1).T


def test_esd_band(oracle_backend):
    import smcp_b200 as S
    P = S.band_SDP(30, 10, 2, seed=1)
    sol = P.solve_esd()
    assert sol["status"] == "optimal"
    assert abs(sol["primal objective"] - 3.1762985) < 1e-5


def test_kktsolver_qr_matches_chol(oracle_backend):
    """kktsolver='qr' (solvers.py:413-475; here in SYRK form on the half factors, Householder QR in the
    oracle): the same Newton systems as 'chol', hence the same iterates to solver accuracy."""
    import smcp_b200 as S
    P = S.band_SDP(40, 12, 3, seed=2)
    a = P.solve_feas(kktsolver="chol", primalstart={"x": P._X0})
    b = P.solve_feas(kktsolver="qr", primalstart={"x": P._X0})
    assert a["status"] == b["status"] == "optimal"
    assert abs(a["iterations"] - b["iterations"]) <= 1
    for key in ("primal objective", "dual objective"):
        assert abs(a[key] - b[key]) <= 1e-8 * max(1.0, abs(a[key]))
    # the self-dual embedding, whose exit is fragile with 'chol' (see tests/test_golden.py), converges with 'qr'
    c = S.band_SDP(30, 10, 2, seed=1).solve_esd(kktsolver="qr")
    assert c["status"] == "optimal" and c["iterations"] <= 25
    assert abs(c["primal objective"] - 3.17629856) <= 1e-6


def test_phase1_then_feas(oracle_backend):
    """example.py:22-35 flow: mtxnorm problem, phase 1 for a primal start, then solve_feas."""
    import smcp_b200 as S
    P = S.mtxnorm_SDP(12, 3, 8, seed=0)
    X0, info = P.solve_phase1()
    assert X0 is not None
    sol = P.solve_feas(primalstart={"x": X0})
    assert sol["status"] == "optimal"
    assert sol["primal objective"] < 0


def test_options_validation(oracle_backend):
    import smcp_b200 as S
    from smcp_b200 import solvers
    P = S.band_SDP(10, 3, 1, seed=0)
    solvers.options["maxiters"] = 0
    with pytest.raises(ValueError):
        P.solve_feas()
    solvers.options["maxiters"] = 1.5
    with pytest.raises(TypeError):
        P.solve_esd()
    solvers.options["maxiters"] = 100
    with pytest.raises(ValueError):
        P.solve_feas(scaling="both")
    with pytest.raises(ValueError):
        P.solve_feas(kktsolver="lu")


def test_product_backend_fails_loudly_without_cuda():
    """No CPU fallback: with no CUDA device the default backend must raise."""
    import smcp_b200 as S
    from smcp_b200 import solvers
    from smcp_b200.device import Context
    solvers.set_backend_factory(None)
    try:
        Context.get(0)
        pytest.skip("a CUDA device is present")
    except RuntimeError:
        pass
    with pytest.raises(RuntimeError):
        S.band_SDP(10, 3, 1, seed=0)


def test_lp_socp_sdp_front_ends(oracle_backend):
    """solvers.lp / socp / sdp (solvers.py:2602-2699) are thin wrappers over conelp: the CVXOPT
    user-guide cone LP split into its parts must give the same optimum through each of them, and a
    small LP has a known vertex solution."""
    from smcp_b200 import solvers
    c_, G_, h_, dims = CONELP
    # LP: min -4x - 5y  s.t. 2x + y <= 3, x + 2y <= 3, x, y >= 0  ->  x = y = 1
    sol = solvers.lp(np.array([-4.0, -5.0]), np.array([[2.0, 1.0], [1.0, 2.0], [-1.0, 0.0], [0.0, -1.0]]),
                     np.array([3.0, 3.0, 0.0, 0.0]))
    assert sol["status"] == "optimal" and np.allclose(sol["x"], [1.0, 1.0], atol=1e-6)
    # the 's' part of the user-guide problem alone, through sdp(), against conelp() on the same data
    Gs, hs = G_[10:19, :], h_[10:19].reshape(3, 3, order="F")
    a = solvers.conelp(c_, Gs, h_[10:19], {"l": 0, "q": [], "s": [3]})
    b = solvers.sdp(c_, Gs=[Gs], hs=[hs])
    assert a["status"] == b["status"]
    if a["status"] == "optimal":
        assert np.allclose(a["x"], b["x"], atol=1e-7)
        assert b["zs"][0].shape == (3, 3) and b["zl"] is None
    # 'l' + 'q' parts through socp()
    a = solvers.conelp(c_, G_[:10, :], h_[:10], {"l": 2, "q": [4, 4], "s": []})
    b = solvers.socp(c_, Gl=G_[:2, :], hl=h_[:2], Gq=[G_[2:6, :], G_[6:10, :]], hq=[h_[2:6], h_[6:10]])
    assert a["status"] == b["status"]
    if a["status"] == "optimal":
        assert np.allclose(a["x"], b["x"], atol=1e-7)
        assert len(b["zq"]) == 2 and len(b["sq"][1]) == 4 and len(b["zl"]) == 2


def test_base_completion_is_max_det(oracle_backend):
    """base.completion(X) (base.py:952-973): the result agrees with X on the pattern and its inverse
    is zero off the pattern (the defining property of the maximum-determinant completion); a matrix
    without a positive definite completion raises ArithmeticError."""
    import scipy.sparse as sp
    import smcp_b200 as S
    rng = np.random.default_rng(0)
    n = 12
    G = rng.standard_normal((n, n))
    W = G @ G.T + n * np.eye(n)
    mask = np.eye(n, dtype=bool)
    for (i, j) in [(1, 0), (2, 1), (3, 2), (5, 2), (7, 3), (8, 7), (9, 0), (10, 9), (11, 4), (6, 5), (4, 3)]:
        mask[i, j] = mask[j, i] = True
    X = sp.csc_matrix(np.tril(W * mask))
    Xh = S.completion(X)
    assert np.allclose(Xh, Xh.T, atol=1e-12)
    assert np.allclose(Xh[mask], W[mask], rtol=1e-10)
    # inverse vanishes outside the (chordal embedding of the) pattern; the pattern above is a forest + diagonal = chordal
    Inv = np.linalg.inv(Xh)
    assert np.abs(Inv[~mask]).max() < 1e-9 * np.abs(Inv).max()
    bad = W * mask
    bad[0, 0] = -1.0
    with pytest.raises(ArithmeticError):
        S.completion(sp.csc_matrix(np.tril(bad)))


def test_c_trsm_matches_numpy():
    """oracle/csn.c (the compiled chordal trsm of bench.py's CPU baseline) against its specification
    oracle/supernodal.py:trsm, both directions, on a pattern with fill and mixed supernode sizes."""
    import scipy.sparse as sp
    from oracle import csn, supernodal as sn
    from smcp_b200 import symbolic as sy
    if not csn.available():
        pytest.skip("oracle/_ref/csn.so not built (make -C oracle)")
    n = 300
    rng = np.random.default_rng(5)
    e = rng.integers(0, n, size=(4 * n, 2))
    I = np.concatenate([np.maximum(e[:, 0], e[:, 1]), np.arange(n)])
    J = np.concatenate([np.minimum(e[:, 0], e[:, 1]), np.arange(n)])
    Lp = sp.tril(sp.coo_matrix((np.ones(len(I)), (I, J)), shape=(n, n))).tocsc()
    Lp.sort_indices()
    cp, ri = Lp.indptr.astype(np.int64), Lp.indices.astype(np.int64)
    out = sy.embed(n, cp, ri, sy.min_degree(n, cp, ri))
    symb = sy.Symbolic(n, out[0], out[1])
    L = rng.standard_normal(symb.nblk) * 0.05
    L[symb.diag_blk] = 1.0 + rng.random(len(symb.diag_blk))
    for k in (1, 7, 40):
        B = rng.standard_normal((n, k))
        for tr in ("N", "T"):
            ref = B.copy()
            sn.trsm(symb, L, ref, tr)
            got = np.ascontiguousarray(B.copy())
            csn.trsm(symb, L, got, tr)
            assert np.linalg.norm(got - ref) <= 1e-12 * np.linalg.norm(ref)
