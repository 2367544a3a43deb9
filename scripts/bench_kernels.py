"""Device time of the dense kernels on their own (C ABI on host buffers, CUDA events around the
kernel only): lapack.potrf replacement, triangular solves, DMMA GEMM.

    python scripts/bench_kernels.py [potrf|trsm|gemm ...]
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smcp_b200.device import Context, _ck

ctx = Context.get()
what = sys.argv[1:] or ["potrf", "trsm", "gemm"]
rng = np.random.default_rng(0)


def best(fn, reps=4):
    t = 1e30
    for _ in range(reps):
        t = min(t, fn())
    return t


if "potrf" in what:
    for m, nc in [(1000, 1000), (1186, 1186), (1131, 1), (1140, 9), (2000, 2000), (2560, 2560), (4096, 4096), (10000, 10000)]:
        H = np.tril(rng.uniform(-1.0, 1.0, size=(m, m)), -1)
        H[np.arange(m), np.arange(m)] = m
        flat0 = np.asfortranarray(H).reshape(-1, order="F")
        info = np.zeros(1, dtype=np.int32)
        ms = C.c_double()

        def run():
            f = flat0.copy()
            _ck(ctx.lib, ctx.lib.smcp_dense_potrf(ctx.h, f, m, m, nc, info, C.byref(ms)))
            assert info[0] == 0
            return ms.value
        t = best(run)
        fl = nc ** 3 / 3.0 + (m - nc) * nc * m
        print("potrf m=%6d ncols=%6d  %9.3f ms  %7.2f TFLOP/s" % (m, nc, t, fl / t / 1e9), flush=True)

if "trsm" in what:
    for n, nrhs in [(1186, 1186), (1186, 9), (1131, 1), (1131, 1131), (2500, 2500), (200, 200), (679, 679)]:
        L = np.tril(rng.standard_normal((n, n))) / np.sqrt(n) + 2.0 * np.eye(n)
        Lf = np.asfortranarray(L).reshape(-1, order="F")
        B0 = np.asfortranarray(rng.standard_normal((n, nrhs))).reshape(-1, order="F")
        ms = C.c_double()
        for trans in (0, 1):
            def run():
                b = B0.copy()
                _ck(ctx.lib, ctx.lib.smcp_dense_trsm(ctx.h, trans, Lf, n, n, b, n, nrhs, C.byref(ms)))
                return ms.value
            t = best(run)
            print("trsm n=%5d nrhs=%5d trans=%d  %9.3f ms  %7.2f TFLOP/s" % (n, nrhs, trans, t, n * n * nrhs / t / 1e9), flush=True)

if "gemm" in what:
    for (ta, tb, M, N, K, tri) in [(1, 1, 1000, 1000, 30000, 1), (0, 0, 9488, 9488, 512, 1), (0, 0, 9744, 9744, 256, 1),
                                   (0, 0, 4096, 4096, 4096, 0), (0, 1, 1186, 1186, 1186, 0), (0, 0, 1131, 1131, 64, 1)]:
        A = np.asfortranarray(rng.standard_normal((K, M) if ta else (M, K)))
        B = np.asfortranarray(rng.standard_normal((K, N) if tb else (N, K)))
        Cm = np.zeros(M * N)
        ms = C.c_double()

        def run():
            _ck(ctx.lib, ctx.lib.smcp_dense_gemm(ctx.h, ta, tb, A.reshape(-1, order="F"), A.shape[0], B.reshape(-1, order="F"), B.shape[0],
                                                 Cm, M, M, N, K, 1.0, 0, tri, C.byref(ms)))
            return ms.value
        t = best(run, 3)
        fl = 2.0 * K * (M * (N + 1) / 2.0 if tri else M * N)
        print("gemm ta=%d tb=%d M=%5d N=%5d K=%6d tri=%d  %9.3f ms  %7.2f TFLOP/s (algorithmic)" % (ta, tb, M, N, K, tri, t, fl / t / 1e9), flush=True)
