"""Per-kernel-family device time of a few IPM iterations (CUDA events around each launch)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import smcp_b200 as S
from smcp_b200 import solvers
from smcp_b200.device import Context

n, m, bw, iters = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
t0 = time.time()
P = S.band_SDP(n, m, bw, seed=0)
print("generated", P, "in %.1fs" % (time.time() - t0), flush=True)
ctx = Context.get()
solvers.options["maxiters"] = iters
solvers.options["show_progress"] = True
t0 = time.time()
sol = P.solve_feas(primalstart={"x": P._X0})
print("plain run: %.3f s total, %.4f s/iter (incl. setup)" % (time.time() - t0, sol["time"] / max(1, sol["iterations"])))
ctx.prof_enable(True)
ctx.prof_reset()
solvers.options["show_progress"] = False
t0 = time.time()
sol = P.solve_feas(primalstart={"x": P._X0})
ctx.prof_enable(False)
names = ctx.prof_names()
tot = 0.0
rows = []
for nm in names:
    ms, cnt = ctx.prof_get(nm)
    tot += ms
    rows.append((ms, nm, cnt))
for ms, nm, cnt in sorted(rows, reverse=True):
    if cnt:
        print("%-20s %10.3f ms %8d launches  %8.1f us/launch" % (nm, ms, cnt, 1e3 * ms / cnt))
print("device total %.3f ms over %d iterations -> %.3f ms/iter" % (tot, sol["iterations"], tot / max(1, sol["iterations"])))
