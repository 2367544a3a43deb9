"""Run one BASELINE.json configuration through the public API on the CUDA backend for a few
iterations and report seconds per iteration (device families included).

    python scripts/run_config.py C4 [iters]      # mtxnorm p=q=200, r=500 (solve_esd)
    python scripts/run_config.py C3 [iters] [n] [m]   # rand_SDP, sparse aggregate pattern
    python scripts/run_config.py C5 [iters] [n]  # max-cut on a random sparse graph (solve_esd)
    python scripts/run_config.py C2 [iters]      # band n=5000 m=1000 (solve_feas)
"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smcp_b200 as S
from smcp_b200 import solvers
from smcp_b200.device import Context

cfg = sys.argv[1]
full = len(sys.argv) > 2 and sys.argv[2] == "full"      # solve to the default tolerances, report time-to-solve
iters = 100 if full else (int(sys.argv[2]) if len(sys.argv) > 2 else 4)
t0 = time.time()
kw = {}
if cfg == "C2":
    P = S.band_SDP(5000, 1000, 5, seed=0)
    method, kw = "feas", {"primalstart": {"x": P._X0}}
elif cfg == "C4":
    P = S.mtxnorm_SDP(200, 200, 500, density=1.0, seed=0)
    method = "esd"
    if os.environ.get("RUNCFG_METHOD") == "feas":
        # the reference's benchmark flow ("M1"): primal phase 1, then the feasible-start solver
        t1 = time.time()
        X0, sol1 = P.solve_phase1(kktsolver="chol")
        print("phase 1: %.2f s, %s iterations" % (time.time() - t1, sol1["iterations"] if sol1 else 0), flush=True)
        method, kw = "feas", {"primalstart": {"x": X0}}
elif cfg == "C3":
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
    m = int(sys.argv[4]) if len(sys.argv) > 4 else 10000
    rng = np.random.default_rng(0)
    ne = 6 * n                 # |V| (lower triangle) ~ 8 n must exceed m: the A_i live on V
    e = rng.integers(0, n, size=(ne, 2))
    I = np.concatenate([np.maximum(e[:, 0], e[:, 1]), np.arange(n), np.arange(1, n)])
    J = np.concatenate([np.minimum(e[:, 0], e[:, 1]), np.arange(n), np.arange(0, n - 1)])
    V = sp.coo_matrix((np.ones(len(I)), (I, J)), shape=(n, n))
    P = S.rand_SDP(V, m, density=0.005, seed=0)
    method, kw = "feas", ({"primalstart": {"x": P._X0}} if getattr(P, "_X0", None) is not None else {})
elif cfg == "C5":
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
    rng = np.random.default_rng(0)
    e = rng.integers(0, n, size=(3 * n // 2, 2))
    P = S.maxcut_SDP(n, e)
    method = "esd"
    if os.environ.get("RUNCFG_METHOD") == "feas":
        # X = I is strictly feasible for the max-cut relaxation (diag(X) = 1)
        method, kw = "feas", {"primalstart": {"x": sp.identity(n, format="csc")}}
else:
    raise SystemExit("unknown config")
print("%s: %s generated in %.1f s" % (cfg, P, time.time() - t0), flush=True)
ctx = Context.get()
solvers.options["maxiters"] = iters
solvers.options["show_progress"] = False
stamps = {}


per_iter = {}
_prev = {}


def hook(name, it):
    ctx.sync()
    stamps[it] = time.perf_counter()
    if os.environ.get("RUNCFG_PER_ITER"):
        # API regions of the iteration that just ended (device ms, calls)
        cur = {nm: ctx.region_get(nm) for nm in ctx.region_names()}
        per_iter[it - 1] = {nm: (v[0] - _prev.get(nm, (0.0, 0))[0], v[1] - _prev.get(nm, (0.0, 0))[1]) for nm, v in cur.items()}
        _prev.clear()
        _prev.update(cur)


solvers._iteration_hook = hook
# pass 1: wall clock per iteration, no per-launch timing (launches overlap with the host)
t0 = time.time()
sol = getattr(P, "solve_" + method)(kktsolver=os.environ.get("RUNCFG_KKT", "chol"), **kw)
print("solve_%s: status %s, %d iterations, %.2f s total (incl. setup)" % (method, sol["status"], sol["iterations"], time.time() - t0))
its = sorted(stamps)
if full:
    t_it = stamps[its[-1]] - stamps[its[0]] if len(its) > 1 else float("nan")
    print("  time-to-solve: %.3f s in the iterations (%d timed), %.2f ms/iteration; pobj %.10e dobj %.10e gap %.2e pres %.1e dres %.1e"
          % (t_it, len(its) - 1, 1e3 * t_it / max(1, len(its) - 1), sol["primal objective"], sol["dual objective"], sol["gap"],
             sol["primal infeasibility"] or 0.0, sol["dual infeasibility"] or 0.0))
    raise SystemExit(0)
for a, b in zip(its[:-1], its[1:]):
    print("  iteration %d: %.2f ms" % (a, 1e3 * (stamps[b] - stamps[a])))
for itn in sorted(per_iter)[1:]:
    row = sorted(per_iter[itn].items(), key=lambda kv: -kv[1][0])
    print("  regions of iteration %d: " % itn + ", ".join("%s %.1f/%d" % (nm.replace("op_", "").replace("kkt_", "k_"), ms, c) for nm, (ms, c) in row if c))
print("  API regions over the whole solve (device ms between event pairs, no synchronisation; calls):")
for nm in sorted(ctx.region_names(), key=lambda k: -ctx.region_get(k)[0]):
    ms, calls = ctx.region_get(nm)
    if calls:
        print("    %-24s %10.2f ms %6d calls %9.3f ms/call" % (nm, ms, calls, ms / calls))
if os.environ.get("RUNCFG_NOPROF"):
    raise SystemExit(0)
# pass 2: the same iterations with per-launch CUDA-event timing (serialises the launches)
stamps.clear()
ctx.prof_reset()
ctx.prof_enable(True)
sol = getattr(P, "solve_" + method)(kktsolver=os.environ.get("RUNCFG_KKT", "chol"), **kw)
ctx.prof_enable(False)
rows = []
for nm in ctx.prof_names():
    ms, cnt = ctx.prof_get(nm)
    if cnt:
        rows.append((ms, nm, cnt))
tot = sum(r[0] for r in rows)
nit = max(1, sol["iterations"])
print("  kernel families over %d iterations (per-launch event timing on):" % nit)
for ms, nm, cnt in sorted(rows, reverse=True)[:14]:
    print("  %-28s %10.3f ms %7d launches %5.1f%%" % (nm, ms, cnt, 100 * ms / tot))
print("  device total %.1f ms" % tot)
