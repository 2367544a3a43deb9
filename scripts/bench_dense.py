"""Dense Schur-complement kernels on their own: lapack.potrf / potrs replacements at several m
(one GPU, or block-cyclic under torchrun).  Prints TFLOP/s of the factorisation (m^3/3 flops)
and the time of one solve.

    python scripts/bench_dense.py 1000 4000 10000
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 \
        scripts/bench_dense.py 10000
"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smcp_b200 as S
from smcp_b200 import device, solvers
from smcp_b200.device import DeviceBackend, Context, _ck
from smcp_b200.symbolic import Symbolic, lower_pattern

rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
ctx = Context.get(local)
if world > 1:
    import torch
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idt = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        idt = torch.frombuffer(bytearray(device.comm_unique_id()), dtype=torch.uint8).clone()
    dist.broadcast(idt, src=0)
    device.init_comm(rank, world, bytes(idt.numpy().tobytes()), block=int(os.environ.get("SMCP_BLOCK", "256")), device=local)

n = 8000
I = np.concatenate([np.arange(n), np.arange(1, n), np.arange(2, n), np.arange(3, n)])
J = np.concatenate([np.arange(n), np.arange(0, n - 1), np.arange(0, n - 2), np.arange(0, n - 3)])
cp, ri = lower_pattern(n, I, J)
symb = Symbolic(n, cp, ri)
for m in [int(a) for a in sys.argv[1:]] or [1000, 4000]:
    ops = DeviceBackend(symb)
    ops.set_operator(sp.random(symb.nvp, m, density=2.0 / symb.nvp, random_state=1, format="csc"), 0)
    rng = np.random.default_rng(0)
    H = rng.uniform(-1.0, 1.0, size=(m, m))
    H = np.tril(H, -1)
    H[np.arange(m), np.arange(m)] = m                  # diagonally dominant -> positive definite
    Hf = np.asfortranarray(H).reshape(-1, order="F")
    info = np.zeros(1, dtype=np.int32)
    best = 1e9
    for rep in range(4):
        _ck(ops.lib, ops.lib.smcp_kkt_set_H(ops._op, Hf))
        if world > 1:
            dist.barrier()
        ctx.sync()
        t0 = time.perf_counter()
        _ck(ops.lib, ops.lib.smcp_kkt_factor_block(ops._op, int(os.environ.get("SMCP_BLOCK", "256")), rank, world, info))
        best = min(best, time.perf_counter() - t0)
        assert info[0] == 0
    rhs = rng.standard_normal(m)
    ts = 1e9
    for rep in range(5):
        ctx.sync()
        t0 = time.perf_counter()
        z = ops.schur_solve(rhs)
        ts = min(ts, time.perf_counter() - t0)
    Hs = H + np.tril(H, -1).T
    res = np.linalg.norm(Hs @ z - rhs) / np.linalg.norm(rhs)
    if m <= 4096:
        Lref = np.linalg.cholesky(Hs)
        errL = np.abs(np.tril(ops.get_H()) - Lref).max() / np.abs(Lref).max()
    else:
        errL = float("nan")
    if world > 1:
        t = torch.tensor([best, ts], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best, ts = float(t[0]), float(t[1])
    if rank == 0:
        print("m=%6d gpus=%d potrf %9.3f ms  %6.2f TFLOP/s   potrs %8.3f ms   |Hz-b|/|b| %.1e  |L-Lref|max rel %.1e"
              % (m, world, best * 1e3, m ** 3 / 3.0 / best / 1e12, ts * 1e3, res, errL), flush=True)
    del ops
