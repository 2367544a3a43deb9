"""Native host symbolic analysis (csrc/host_symbolic.cpp) against its Python specification
(smcp_b200/symbolic.py): orderings, filled patterns and elimination trees must be bit-identical
(BASELINE north star: "chordal pattern, elimination tree and supernode partition must match
bit-exactly")."""
import numpy as np
import pytest

from smcp_b200 import symbolic as sy


def _random_pattern(n, nedges, seed, band=0):
    rng = np.random.default_rng(seed)
    e = rng.integers(0, n, size=(nedges, 2))
    I, J = [e[:, 0]], [e[:, 1]]
    for b in range(1, band + 1):
        I.append(np.arange(b, n))
        J.append(np.arange(0, n - b))
    return sy.lower_pattern(n, np.concatenate(I), np.concatenate(J))


def test_native_library_is_used():
    assert sy._native() is not None, "libsmcp_b200.so must be built (python __graft_entry__.py)"


@pytest.mark.parametrize("n,ne,band,seed", [(1, 0, 0, 0), (2, 1, 0, 1), (30, 40, 0, 2), (200, 300, 1, 3),
                                           (400, 450, 0, 4), (300, 2000, 0, 5), (150, 0, 3, 6), (64, 0, 0, 7)])
def test_native_matches_python(n, ne, band, seed):
    cp, ri = _random_pattern(n, ne, seed, band)
    p_nat, p_py = sy.min_degree(n, cp, ri), sy._py_min_degree(n, cp, ri)
    assert np.array_equal(p_nat, p_py)
    m_nat, m_py = sy.maxcardsearch(n, cp, ri), sy._py_maxcardsearch(n, cp, ri)
    assert np.array_equal(m_nat, m_py)
    for p in (None, p_py, m_py):
        a, b = sy.embed(n, cp, ri, p), sy._py_embed(n, cp, ri, p)
        for x, y in zip(a, b):
            assert x.dtype == np.int64 and np.array_equal(x, y)


def test_chordal_pattern_zero_fill():
    # band pattern is chordal: MCS order is a PEO, embedding adds nothing
    n = 120
    cp, ri = _random_pattern(n, 0, 0, band=4)
    p = sy.maxcardsearch(n, cp, ri)
    assert np.array_equal(p, np.arange(n))
    fc, fr, par = sy.embed(n, cp, ri, p)
    assert int(fc[-1]) == int(cp[-1]) and np.array_equal(fr, ri)
    assert np.array_equal(par[:-1], np.arange(1, n)) and par[-1] == -1
    assert sy.peo(n, cp, ri, p)


def test_min_degree_dense_tail_shortcut():
    # a clique plus pendant vertices: exercises the "rest is a clique" shortcut
    n = 40
    I, J = np.meshgrid(np.arange(25), np.arange(25))
    I = np.concatenate([I.ravel(), np.arange(25, n)])
    J = np.concatenate([J.ravel(), np.arange(0, n - 25)])
    cp, ri = sy.lower_pattern(n, I, J)
    assert np.array_equal(sy.min_degree(n, cp, ri), sy._py_min_degree(n, cp, ri))


@pytest.mark.parametrize("n,ne,band,seed", [(30, 40, 0, 2), (200, 300, 1, 3), (300, 2000, 0, 5), (150, 0, 3, 6), (1, 0, 0, 0)])
def test_native_aaidx_matches_python(n, ne, band, seed):
    cp, ri = _random_pattern(n, ne, seed, band)
    p = sy.min_degree(n, cp, ri)
    fc, fr, _ = sy.embed(n, cp, ri, p)
    symb = sy.Symbolic(n, fc, fr)                # built with the native map
    ref = sy._py_aaidx(symb.nsn, symb.snpar, symb.nn, symb.na, symb.nj, symb.relptr, symb.relidx, symb.blkptr,
                       symb.updptr, symb.nupd)
    assert symb.aaidx.dtype == np.int64 and np.array_equal(symb.aaidx, ref)


@pytest.mark.parametrize("n,ne,band,seed", [(1, 0, 0, 0), (30, 40, 0, 2), (200, 300, 1, 3), (300, 2000, 0, 5), (150, 0, 3, 6), (64, 0, 0, 7)])
def test_native_supernodes_match_python(n, ne, band, seed, monkeypatch):
    """smcp_host_supernodes (partition, post-ordered relabelling, row lists, relative indices) against the NumPy
    specification: every array of the Symbolic object bit-identical."""
    cp, ri = _random_pattern(n, ne, seed, band)
    p = sy.min_degree(n, cp, ri)
    fc, fr, _ = sy.embed(n, cp, ri, p)
    a = sy.Symbolic(n, fc, fr)
    monkeypatch.setenv("SMCP_B200_NO_NATIVE_HOST", "1")
    b = sy.Symbolic(n, fc, fr)
    for k in ("nsn", "perm", "iperm", "snptr", "snpar", "rowptr", "rowidx", "relptr", "relidx", "chptr", "chidx", "blkptr",
              "updptr", "aaidx", "vec2blk", "height", "depth", "diag_blk"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k


def test_amalgamation_is_a_chordal_superset():
    """Opt-in relaxed supernodes: the merged pattern contains the original, is still a filled pattern in the same
    ordering (Symbolic accepts it), and has fewer, larger supernodes; tol = 0 is the identity."""
    n = 400
    cp, ri = _random_pattern(n, 0, 0, band=5)
    assert sy.amalgamate(n, cp, ri, 0.0)[0] is cp
    s0 = sy.Symbolic(n, cp, ri)
    cp2, ri2 = sy.amalgamate(n, cp, ri, 0.3)
    s1 = sy.Symbolic(n, cp2, ri2)
    assert s1.nsn < s0.nsn // 3 and s1.nn.max() >= 3
    old = set(zip(np.repeat(np.arange(n), np.diff(cp)).tolist(), ri.tolist()))
    new = set(zip(np.repeat(np.arange(n), np.diff(cp2)).tolist(), ri2.tolist()))
    assert old <= new and len(new) < 2 * len(old)
    cp3, ri3 = _random_pattern(120, 150, 3, band=1)
    p = sy.min_degree(120, cp3, ri3)
    fc, fr, _ = sy.embed(120, cp3, ri3, p)
    fc2, fr2 = sy.amalgamate(120, fc, fr, 0.2)
    s2 = sy.Symbolic(120, fc2, fr2)                 # raises if the merged pattern were not chordal in this order
    assert s2.nsn <= sy.Symbolic(120, fc, fr).nsn
