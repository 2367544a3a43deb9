"""Sparse SDPA (dat-s) reader / writer, pinned on the reference's own
``misc.sdpa_read`` compiled unmodified from ``src/C/misc.c`` (oracle/_ref)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from smcp_b200 import misc


SDPA_TEXT = """* a small block-diagonal problem (two blocks, the second one diagonal)
"second comment line"
3 = mDIM
2 = nBLOCK
{2, -3}
{1.5, -2, 0.25}
0 1 1 1 1.0
0 1 1 2 -0.5
0 1 2 2 2.0
0 2 1 1 3.0
0 2 3 3 1.0
1 1 1 1 1.0
1 2 2 2 -1.0
2 1 1 2 0.75
2 2 1 1 0
3 1 2 2 1e-3
3 2 3 3 2.5D+0
"""


@pytest.fixture()
def sdpa_file(tmp_path):
    p = tmp_path / "small.dat-s"
    p.write_text(SDPA_TEXT)
    return str(p)


def test_sdpa_read_known_answer(sdpa_file):
    A, b, bs = misc.sdpa_read(sdpa_file)
    n = 5
    assert A.shape == (n * n, 4) and np.array_equal(bs, [2, -3])
    assert np.allclose(b, [1.5, -2.0, 0.25])
    D = A.toarray()
    # entry (i, j), i <= j, of block with offset o -> row (i-1+o)*n + (j-1+o)
    assert D[0 * n + 0, 0] == 1.0 and D[0 * n + 1, 0] == -0.5 and D[1 * n + 1, 0] == 2.0
    assert D[2 * n + 2, 0] == 3.0 and D[4 * n + 4, 0] == 1.0
    assert D[0, 1] == 1.0 and D[3 * n + 3, 1] == -1.0
    assert D[0 * n + 1, 2] == 0.75 and np.count_nonzero(D[:, 2]) == 1          # explicit zero dropped
    assert D[1 * n + 1, 3] == 1e-3 and D[4 * n + 4, 3] == 2.5                   # Fortran exponent
    An, bn, _ = misc.sdpa_read(sdpa_file, neg=True)
    assert np.array_equal(An.toarray(), -D) and np.array_equal(bn, -b)
    assert misc.sdpa_readhead(sdpa_file)[:2] == (5, 3)


def test_sdpa_read_matches_reference_build(sdpa_file):
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    for neg in (False, True):
        A, b, bs = misc.sdpa_read(sdpa_file, neg=neg)
        Ar, br, bsr = ref.sdpa_read(sdpa_file, neg=neg)
        assert np.array_equal(bs, bsr) and np.array_equal(b, br)
        assert A.shape == Ar.shape and (sp.csc_matrix(A) != sp.csc_matrix(Ar)).nnz == 0


def test_sdpa_roundtrip_and_constructor(tmp_path):
    import smcp_b200 as S
    rng = np.random.default_rng(0)
    n, m = 6, 4
    rows, cols, vals = [], [], []
    for k in range(m + 1):
        for _ in range(5):
            i, j = sorted(rng.integers(0, n, size=2))
            rows.append(j + n * i)            # lower triangle: row j >= column i
            cols.append(k)
            vals.append(float(rng.standard_normal()))
    A = misc.as_csc(sp.csc_matrix((vals, (rows, cols)), shape=(n * n, m + 1)))
    b = rng.standard_normal(m)
    P = S.SDP(A, b)
    # the reference's signature (base.py:197-217): write_sdpa(fname=None, compress=False) appends '.dat-s'
    # (and '.bz2'), defaults to the problem name and refuses to overwrite
    f = P.write_sdpa(str(tmp_path / "rt"))
    assert f == str(tmp_path / "rt.dat-s")
    for Q in (S.SDP(f), S.SDP(filename=f)):
        assert Q.n == n and Q.m == m
        assert np.allclose(Q.b, b, rtol=1e-11) and abs(Q.A - A).max() <= 1e-11 * abs(A).max()
    with pytest.raises(IOError):
        P.write_sdpa(str(tmp_path / "rt"))
    fz = P.write_sdpa(str(tmp_path / "rtz"), compress=True)
    assert fz.endswith(".dat-s.bz2") and not os.path.exists(fz[:-4])
    Qz = S.SDP(fz)
    assert abs(Qz.A - A).max() <= 1e-11 * abs(A).max()
    with pytest.raises(IOError):
        P.write_sdpa(str(tmp_path / "rtz"), compress=True)
    with pytest.raises(ValueError):
        P.write_sdpa()                       # unnamed problem, no file name
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        assert S.SDP(A, b, name="named").write_sdpa() == "named.dat-s"
    finally:
        os.chdir(cwd)


def test_sdp_from_bz2_and_counts(sdpa_file, tmp_path):
    import bz2
    import smcp_b200 as S
    z = str(tmp_path / "small.dat-s.bz2")
    with open(sdpa_file, "rb") as fi, open(z, "wb") as fo:
        fo.write(bz2.compress(fi.read()))
    P, Q = S.SDP(sdpa_file), S.SDP(z)
    assert (P.A != Q.A).nnz == 0 and np.array_equal(P.b, Q.b) and Q._pname == "small"
    assert np.array_equal(P.get_nnz(), [5, 2, 1, 2]) and P.get_nnz(1) == 2
    assert np.array_equal(P.get_nzcols(), P.nzcols) and P.get_nzcols(1) == 2
    with pytest.raises(ValueError):
        P.get_nzcols(0)
    with pytest.raises(NameError):
        S.SDP(str(tmp_path / "x.txt"))


@pytest.mark.parametrize("compress", [False, True])
def test_save_and_load_pickle(tmp_path, compress):
    import smcp_b200 as S
    from smcp_b200 import solvers
    from oracle.backend import OracleBackend
    solvers.set_backend_factory(lambda symb: OracleBackend(symb))      # generators probe the cone
    try:
        P = S.band_SDP(12, 4, 2, seed=3)
    finally:
        solvers.set_backend_factory(None)
    f = P.save(str(tmp_path / "prob"), compress=compress)
    Q = S.SDP(f)
    assert (P.A != Q.A).nnz == 0 and np.array_equal(P.b, Q.b) and Q._pname == P._pname
    assert (P._X0 != Q._X0).nnz == 0
    with pytest.raises(IOError):
        P.save(str(tmp_path / "prob"), compress=compress)
