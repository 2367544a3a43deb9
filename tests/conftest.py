import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_available():
    try:
        from smcp_b200.device import Context
        Context.get()
        return True
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # without a CUDA device the gpu tests are skipped (the product path itself still fails loudly: see
    # test_oracle_drivers.test_product_backend_fails_loudly_without_cuda); `-m gpu` on the box runs them all
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if gpu_items and not _cuda_available():
        skip = pytest.mark.skip(reason="no CUDA device")
        for it in gpu_items:
            it.add_marker(skip)


@pytest.fixture(autouse=True)
def _reset_solver_state():
    from smcp_b200 import solvers
    saved = dict(solvers.options)
    yield
    solvers.options.clear()
    solvers.options.update(saved)
    solvers.set_backend_factory(None)


def tree_pattern(n, depth, seed):
    """Chordal lower pattern with tiny cliques: vertex j is joined to `depth` consecutive
    ancestors in a random recursive tree (cliques of at most depth+1 vertices)."""
    from smcp_b200.symbolic import lower_pattern
    rng = np.random.default_rng(seed)
    par = np.full(n, -1, dtype=np.int64)
    I, J = [], []
    for j in range(1, n):
        par[j] = rng.integers(max(0, j - 6), j)
        a = j
        for _ in range(depth):
            a = par[a]
            if a < 0:
                break
            I.append(j)
            J.append(int(a))
    return lower_pattern(n, I, J)


def random_pattern(n, nextra, bw, seed):
    """Lower pattern: band of width bw plus nextra random off-diagonal pairs (bw < 0: a
    tree-structured pattern with cliques of at most |bw|+1 vertices)."""
    from smcp_b200.symbolic import lower_pattern
    if bw < 0:
        return tree_pattern(n, -bw, seed)
    if nextra < 0:
        # chain of overlapping cliques {t*stride, ..., t*stride + bw - 1}: supernodes with
        # several columns AND a separator (nn = stride, na = bw - stride)
        stride = -nextra
        I, J = [], []
        for t0 in range(0, n, stride):
            c = list(range(t0, min(n, t0 + bw)))
            for a in c:
                for b in c:
                    if a >= b:
                        I.append(a)
                        J.append(b)
        return lower_pattern(n, I, J)
    rng = np.random.default_rng(seed)
    I, J = [], []
    for j in range(n):
        for i in range(j, min(n, j + bw + 1)):
            I.append(i)
            J.append(j)
    e = rng.integers(0, n, size=(nextra, 2))
    I += list(e[:, 0])
    J += list(e[:, 1])
    return lower_pattern(n, I, J)


def make_symbolic(n, nextra, bw, seed):
    from smcp_b200.symbolic import Symbolic, embed, min_degree, maxcardsearch
    cp, ri = random_pattern(n, nextra, bw, seed)
    p = maxcardsearch(n, cp, ri)
    fc, fr, _ = embed(n, cp, ri, p)
    if fc[-1] != cp[-1]:
        p = min_degree(n, cp, ri)
        fc, fr, _ = embed(n, cp, ri, p)
    return Symbolic(n, fc, fr)


def random_pd(symb, seed, shift=0.5):
    """blkval array of a random positive definite chordal matrix."""
    from oracle import dense as dn
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(symb.nblk) * (symb.wdot > 0)
    M = dn.to_dense(symb, x)
    M += np.diag(np.abs(M).sum(1) + shift)
    return dn.project(symb, M)


PATTERNS = [
    # (n, extra edges, bandwidth, seed)
    (40, 30, 1, 0),      # sparse random, mixed supernodes
    (60, 40, 2, 1),
    (30, 0, 3, 2),       # band: chain of tiny cliques
    (25, 200, 1, 3),     # nearly dense: one large supernode
    (1, 0, 0, 4),        # 1 x 1
    (12, 0, 0, 5),       # diagonal: n independent 1 x 1 supernodes (forest)
    (150, 120, 2, 6),
    # tiny cliques (every clique <= 8 vertices): warp-per-chain kernels of chordal_small.cu
    (300, 0, 5, 7),      # band, bandwidth 5: the benchmark's chain of 6 x 1 blocks
    (90, 0, 7, 8),       # band, bandwidth 7: 8 x 8 fronts (the largest the small path takes)
    (80, 0, 1, 9),       # tridiagonal
    (400, 0, -3, 10),    # random tree, cliques of 4: many tasks, cross-task dependencies
    (250, 0, -6, 11),    # random tree, cliques of 7
    (120, -3, 6, 12),    # clique chain: 3-column supernodes with 3-row separators
    (100, -2, 8, 13),    # clique chain: 2-column supernodes with 6-row separators
    (180, 2500, 1, 14),  # heavy fill: a dense root supernode of ~170 columns (dense root path, front.cu)
]
