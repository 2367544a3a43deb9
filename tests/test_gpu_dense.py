"""Dense kernels behind the Schur complement and the large frontal matrices, through the C ABI on
host buffers (smcp_dense_potrf / smcp_dense_trsm / smcp_dense_gemm), against NumPy/SciPy LAPACK:
the reference's calls are cvxopt.lapack.potrf / potrs (src/python/solvers.py:501, 526) and
chompack's per-supernode BLAS.  Tolerances are relative Frobenius errors, written per test."""
import ctypes as C

import numpy as np
import pytest
import scipy.linalg as sl

pytestmark = pytest.mark.gpu


def _ctx():
    from smcp_b200.device import Context
    return Context.get()


def _spd(m, seed, cond_shift=None):
    rng = np.random.default_rng(seed)
    G = rng.standard_normal((m, min(m, 300)))
    return G @ G.T / G.shape[1] + (cond_shift if cond_shift is not None else 1.0) * np.eye(m)


def _potrf(A, ncols, lda=None):
    from smcp_b200.device import _ck
    ctx = _ctx()
    m = A.shape[0]
    lda = lda or m
    Af = np.zeros((lda, m), order="F")
    Af[:m, :] = np.tril(A)
    Af[:m, :] += np.triu(np.full((m, m), 7.5), 1)          # the strict upper triangle must not be referenced
    flat = Af.reshape(-1, order="F").copy()
    info = np.zeros(1, dtype=np.int32)
    ms = C.c_double()
    _ck(ctx.lib, ctx.lib.smcp_dense_potrf(ctx.h, flat, lda, m, ncols, info, C.byref(ms)))
    return flat.reshape((lda, m), order="F")[:m, :], int(info[0]), ms.value


# m <= 2560: one cooperative launch (potrf_tile_kernel); above: column blocks of 512 + DMMA updates
@pytest.mark.parametrize("m", [1, 2, 63, 64, 65, 127, 128, 129, 200, 333, 640, 1000, 1186, 2049, 2560, 2561, 3100])
def test_potrf_full(m):
    A = _spd(m, m)
    out, info, _ = _potrf(A, m)
    assert info == 0
    L = np.tril(out)
    Lref = np.linalg.cholesky(A)
    assert np.linalg.norm(L - Lref) <= 1e-12 * np.linalg.norm(Lref)
    assert np.all(out[np.triu_indices(m, 1)] == 7.5)


@pytest.mark.parametrize("m,ncols", [(70, 1), (70, 9), (300, 64), (300, 100), (1131, 1), (1140, 9), (1186, 600), (700, 699),
                                     (2000, 1300), (3000, 700), (3000, 1024), (3300, 1100)])
def test_potrf_partial(m, ncols):
    """Leading ncols columns factored, Schur complement in the trailing block: a frontal matrix of
    the supernodal Cholesky."""
    A = _spd(m, 3 * m + ncols)
    out, info, _ = _potrf(A, ncols)
    assert info == 0
    L11 = np.linalg.cholesky(A[:ncols, :ncols])
    L21 = sl.solve_triangular(L11, A[:ncols, ncols:], lower=True).T
    Sc = A[ncols:, ncols:] - L21 @ L21.T
    got = np.tril(out)
    assert np.linalg.norm(got[:ncols, :ncols] - L11) <= 1e-12 * np.linalg.norm(L11)
    assert np.linalg.norm(got[ncols:, :ncols] - L21) <= 1e-12 * max(1.0, np.linalg.norm(L21))
    assert np.linalg.norm(got[ncols:, ncols:] - np.tril(Sc)) <= 1e-12 * np.linalg.norm(Sc)


def test_potrf_leading_dimension():
    m, lda = 500, 517
    A = _spd(m, 5)
    out, info, _ = _potrf(A, m, lda=lda)
    assert info == 0
    Lref = np.linalg.cholesky(A)
    assert np.linalg.norm(np.tril(out) - Lref) <= 1e-12 * np.linalg.norm(Lref)


@pytest.mark.parametrize("m,bad", [(40, 7), (200, 150), (300, 1), (500, 333), (1000, 1000), (2700, 2000)])
def test_potrf_info(m, bad):
    A = _spd(m, m + 1)
    L = np.linalg.cholesky(A)
    A2 = A.copy()
    A2[bad - 1, bad - 1] -= L[bad - 1, bad - 1] ** 2 + 1.0
    _, info, _ = _potrf(A2, m)
    assert info == bad


@pytest.mark.parametrize("m", [12, 70, 900])
def test_potrf_ill_conditioned(m):
    """cond ~ 1e12: the factor still reproduces A to working precision (backward stability)."""
    rng = np.random.default_rng(1)
    Q, _ = np.linalg.qr(rng.standard_normal((m, m)))
    A = (Q * np.logspace(0, -12, m)) @ Q.T
    A = 0.5 * (A + A.T)
    out, info, _ = _potrf(A, m)
    assert info == 0
    L = np.tril(out)
    assert np.linalg.norm(L @ L.T - A) <= 1e-13 * np.linalg.norm(A)


@pytest.mark.parametrize("trans", [0, 1])
@pytest.mark.parametrize("n,nrhs", [(1, 1), (5, 3), (63, 8), (64, 9), (65, 1), (130, 17), (200, 200), (1186, 1), (1186, 1186),
                                    (1131, 9), (2816, 40), (3000, 50),
                                    # one or two right-hand sides on a large factor: cluster kernel on inverted diagonal blocks
                                    (300, 1), (1131, 2), (2000, 2), (5000, 1)])
def test_trsm(n, nrhs, trans):
    """n <= 2816: slab kernel (one launch); above: the blocked launch chain of front.cu."""
    from smcp_b200.device import _ck
    ctx = _ctx()
    rng = np.random.default_rng(n + nrhs)
    L = np.tril(rng.standard_normal((n, n))) / np.sqrt(n) + 2.0 * np.eye(n)
    B = rng.standard_normal((n, nrhs))
    ldl, ldb = n + 3, n + 1
    Lf = np.full((ldl, n), 9.25, order="F")
    Lf[:n, :] = L + np.triu(np.full((n, n), 9.25), 1)
    Bf = np.full((ldb, nrhs), -3.0, order="F")
    Bf[:n, :] = B
    bflat = Bf.reshape(-1, order="F").copy()
    _ck(ctx.lib, ctx.lib.smcp_dense_trsm(ctx.h, trans, Lf.reshape(-1, order="F").copy(), ldl, n, bflat, ldb, nrhs, None))
    out = bflat.reshape((ldb, nrhs), order="F")
    ref = sl.solve_triangular(L, B, lower=True, trans="T" if trans else "N")
    assert np.linalg.norm(out[:n] - ref) <= 1e-12 * np.linalg.norm(ref)
    assert np.all(out[n:] == -3.0)


@pytest.mark.parametrize("ta,tb", [(1, 1), (0, 0), (0, 1), (1, 0)])
@pytest.mark.parametrize("M,N,K,tri", [(300, 300, 4000, 1), (257, 130, 64, 0), (1000, 1000, 512, 1), (128, 128, 8, 1), (90, 700, 33, 0),
                                       # N <= 16 with K > 16: matrix times a few vectors (gemm_thin_kernel); K <= 16: rank-k kernel
                                       (1131, 1, 1131, 0), (700, 9, 300, 0), (5, 16, 1000, 0), (40, 3, 17, 0), (1131, 1131, 1, 1),
                                       (333, 200, 16, 0),
                                       # K-major operands (ta = tb = 1) with even leading dimensions: TMA-fed kernel, with split-K
                                       (1000, 1000, 30000, 1), (77, 513, 2050, 0), (130, 257, 1001, 0), (129, 129, 256, 1)])
def test_gemm(ta, tb, M, N, K, tri):
    from smcp_b200.device import _ck
    ctx = _ctx()
    rng = np.random.default_rng(M + N + K)
    A = rng.standard_normal((M, K))
    B = rng.standard_normal((N, K))
    C0 = rng.standard_normal((M, N))
    Af = np.asfortranarray(A.T if ta else A)
    Bf = np.asfortranarray(B.T if tb else B)
    cflat = np.asfortranarray(C0).reshape(-1, order="F").copy()
    _ck(ctx.lib, ctx.lib.smcp_dense_gemm(ctx.h, ta, tb, Af.reshape(-1, order="F").copy(), Af.shape[0],
                                         Bf.reshape(-1, order="F").copy(), Bf.shape[0], cflat, M, M, N, K, -1.0, 1, tri, None))
    out = cflat.reshape((M, N), order="F")
    ref = C0 - A @ B.T
    if tri:
        mask = np.tril(np.ones((M, N), dtype=bool))
        assert np.linalg.norm((out - ref)[mask]) <= 1e-12 * np.linalg.norm(ref)
        assert np.array_equal(out[~mask], C0[~mask])
    else:
        assert np.linalg.norm(out - ref) <= 1e-12 * np.linalg.norm(ref)
