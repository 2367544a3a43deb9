"""Host wall time per C-ABI call over a few iterations of the bench workload (debug aid).
    SMCP_B200_HOSTPROF=1 python scripts/host_prof.py [n m bw iters]"""
import os
import sys
import time
os.environ["SMCP_B200_HOSTPROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smcp_b200 as S
from smcp_b200 import solvers, device
from smcp_b200.device import Context

n, m, bw, iters = (int(a) for a in (sys.argv[1:5] if len(sys.argv) > 4 else (5000, 1000, 5, 8)))
P = S.band_SDP(n, m, bw, seed=0)
ctx = Context.get()
solvers.options["show_progress"] = False
solvers.options["maxiters"] = iters
stamps = {}
snaps = {}


def hook(name, it):
    ctx.sync()
    stamps[it] = time.perf_counter()
    snaps[it] = {k: tuple(v) for k, v in device.HOSTPROF.items()}
    if it == 3:
        device.HOSTPROF.clear()
        snaps[it] = {}


solvers._iteration_hook = hook
sol = P.solve_feas(kktsolver="chol", primalstart={"x": P._X0})
its = sorted(stamps)
print("iterations (ms):", " ".join("%.2f" % (1e3 * (stamps[b] - stamps[a])) for a, b in zip(its[:-1], its[1:])))
for a, b in zip(its[:-1], its[1:]):
    if a < 3:
        continue
    d = {k: (v[0] - snaps[a].get(k, (0, 0.0))[0], v[1] - snaps[a].get(k, (0, 0.0))[1]) for k, v in snaps[b].items()}
    top = sorted(d.items(), key=lambda kv: -kv[1][1])[:5]
    print("  iteration %d: %.2f ms wall; " % (a, 1e3 * (stamps[b] - stamps[a])) +
          ", ".join("%s x%d %.2f ms" % (k.replace("smcp_", ""), c, 1e3 * t) for k, (c, t) in top))
nit = max(1, its[-1] - 3)
tot = sum(v[1] for v in device.HOSTPROF.values())
print("host time inside the C ABI per iteration: %.2f ms" % (1e3 * tot / nit))
for name, (cnt, sec) in sorted(device.HOSTPROF.items(), key=lambda kv: -kv[1][1])[:14]:
    print("  %-28s %7.1f calls/iter %8.3f ms/iter" % (name, cnt / nit, 1e3 * sec / nit))
