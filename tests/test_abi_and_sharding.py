"""CPU tests of the drop-in boundary and the multi-GPU host logic.

* every function declared in include/smcp_b200.h is exported by libsmcp_b200.so and bound by
  the ctypes layer (no compute call is made: there is no GPU here);
* the 1-D block-cyclic column ownership used to shard the Schur complement covers every
  column exactly once; a world_size-2 gloo run assembles disjoint column blocks of H with the
  oracle on each rank and the exchange reproduces the single-rank H.
"""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "smcp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(smcp_[A-Za-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from smcp_b200 import device
    names = _declared()
    assert len(names) >= 45
    lib = device.load_library()
    for nm in names:
        assert hasattr(lib, nm), "declared in include/smcp_b200.h but not exported: " + nm
    assert sorted(device.API) == names, (set(names) ^ set(device.API))
    out = subprocess.run(["nm", "-D", "--defined-only", device._LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (smcp_[A-Za-z0-9_]+)", out))
    assert set(names) <= exported


def test_library_has_sm100a_code():
    from smcp_b200 import device
    out = subprocess.run(["cuobjdump", "-lelf", device._LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out[:300]


def test_no_cpu_fallback_in_ctx_create():
    """Without a device smcp_ctx_create must fail and say so."""
    import ctypes as C
    from smcp_b200 import device
    lib = device.load_library()
    h = C.c_void_p()
    rc = lib.smcp_ctx_create(0, C.byref(h))
    if rc == 0:
        lib.smcp_ctx_destroy(h)
        pytest.skip("a CUDA device is present")
    assert b"no CPU fallback" in lib.smcp_last_error()


@pytest.mark.parametrize("m,nranks,block", [(1000, 8, 64), (10, 4, 64), (513, 2, 64), (64, 3, 16), (1, 2, 8)])
def test_block_cyclic_ownership(m, nranks, block):
    from smcp_b200.device import owned_column_blocks
    seen = np.zeros(m, dtype=int)
    for r in range(nranks):
        for c0, c1 in owned_column_blocks(m, r, nranks, block):
            assert 0 <= c0 < c1 <= m and c0 % block == 0
            assert (c0 // block) % nranks == r
            seen[c0:c1] += 1
    assert np.all(seen == 1)


_WORKER = r"""
# One process per rank (gloo): the host-side logic of the N > 1 path on the CPU oracle.
#  1. every rank assembles ONLY the column blocks of H it owns (1-D block-cyclic, like
#     smcp_kkt_assemble_cyclic): dense-technique and sparse-technique columns;
#  2. the blocks are factored with the schedule of d_potrf (csrc/dense.cu): the owner of block q
#     factors its panel and broadcasts it, every rank applies it to the blocks it owns;
#  3. every rank must end with the factor of the H a single process assembles -- and the exchange
#     of the UNFACTORED blocks (smcp_kkt_allgather) must reproduce that H bit for bit.
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np
import scipy.linalg as sl
import scipy.sparse as sp
import torch, torch.distributed as dist
from smcp_b200.device import owned_column_blocks
from smcp_b200 import solvers
from smcp_b200.solvers import _Problem, _read_options
from oracle.backend import OracleBackend
import smcp_b200 as S
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
solvers.options["show_progress"] = False
solvers.set_backend_factory(lambda symb: OracleBackend(symb, batch_columns=8))
rng = np.random.default_rng(7)
e = rng.integers(0, 40, size=(50, 2))
V = sp.coo_matrix((np.ones(50 + 40), (np.concatenate([e[:, 0], np.arange(40)]), np.concatenate([e[:, 1], np.arange(40)]))), shape=(40, 40))
cases = [("band-dense", S.band_SDP(30, 21, 2, seed=3), 4), ("rand-sparse", S.rand_SDP(V, 23, density=0.04, seed=4), 3)]
for name, P, block in cases:
    pr = _Problem(P.A, P.b, _read_options(P.n, False), "chol", None)
    ob, symb, m = pr.ops, pr.symb, pr.m
    s = np.zeros(symb.nvp); s[symb.diag_vec] = 2.0
    s += 0.05 * np.random.default_rng(0).standard_normal(symb.nvp)
    L = ob.from_vec(s); ob.cholesky(L); Y = ob.clone(L); ob.projected_inverse(Y)
    hf = ob.hessian_factor(L, Y)
    mine = owned_column_blocks(m, rank, world, block)
    cols0 = ob.stats["hessian_cols"]
    Hloc = np.tril(ob.schur_assemble(hf, columns=mine)).copy()          # this rank's blocks only
    done = ob.stats["hessian_cols"] - cols0
    assert done == sum(min(c1, m - pr.Ns) - min(c0, m - pr.Ns) for c0, c1 in mine), (name, "assembled foreign columns")
    for c0, c1 in owned_column_blocks(m, (rank + 1) %% world, world, block):
        assert not Hloc[:, c0:c1].any(), (name, "foreign block is not empty")
    Href = np.tril(ob.schur_assemble(hf)).copy()                          # single-process reference
    # (3) exchange of the unfactored blocks
    Hx = Hloc.copy()
    for q, c0 in enumerate(range(0, m, block)):
        c1 = min(m, c0 + block)
        t = torch.from_numpy(np.ascontiguousarray(Hx[:, c0:c1]))
        dist.broadcast(t, src=q %% world)
        Hx[:, c0:c1] = t.numpy()
    assert np.array_equal(Hx, Href), (name, "sharded assembly + exchange does not reproduce H")
    # (2) block-cyclic right-looking Cholesky with panel broadcasts
    Hd = Hloc.copy()
    nb = (m + block - 1) // block
    for q in range(nb):
        c0, c1 = q * block, min(m, (q + 1) * block)
        if q %% world == rank:
            Hd[c0:c1, c0:c1] = np.linalg.cholesky(Hd[c0:c1, c0:c1] + np.tril(Hd[c0:c1, c0:c1], -1).T)
            if c1 < m:
                Hd[c1:, c0:c1] = sl.solve_triangular(Hd[c0:c1, c0:c1], Hd[c1:, c0:c1].T, lower=True).T
        t = torch.from_numpy(np.ascontiguousarray(Hd[:, c0:c1]))
        dist.broadcast(t, src=q %% world)
        Hd[:, c0:c1] = t.numpy()
        for q2 in range(q + 1, nb):
            if q2 %% world != rank:
                continue
            d0, d1 = q2 * block, min(m, (q2 + 1) * block)
            Hd[d0:, d0:d1] -= Hd[d0:, c0:c1] @ Hd[d0:d1, c0:c1].T
    Lref = np.linalg.cholesky(Href + np.tril(Href, -1).T)
    assert np.linalg.norm(np.tril(Hd) - Lref) <= 1e-12 * np.linalg.norm(Lref), (name, "distributed factor differs")
    # every rank holds the same factor bit for bit
    t = torch.from_numpy(np.ascontiguousarray(np.tril(Hd)))
    t0 = t.clone()
    dist.broadcast(t0, src=0)
    assert torch.equal(t, t0), (name, "ranks disagree on the factor")
dist.barrier()
if rank == 0:
    print("OK")
"""


def test_sharded_schur_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert "OK" in out.stdout


def test_bench_arms_print_the_same_config():
    """bench.py: the b200 arm and the reference arm describe the workload with identical `config` objects
    (the driver compares them), for every N, and BASELINE.json's headline configuration is the default."""
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import bench
    assert bench.DEFAULT == "rand_n2000_m10000"
    for world in (1, 2, 8):
        a, b = bench.config_of(bench.DEFAULT, world), bench.config_of(bench.DEFAULT, world)
        assert a == b and a["workload"] == bench.DEFAULT and a["n"] == 2000 and a["m"] == 10000
        assert "l2" in a and ("x%d " % world) in a["parallelism"]
    assert "model" not in bench.config_of(bench.DEFAULT, 1)
