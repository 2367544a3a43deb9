"""cProfile of IPM iterations W..W+K on the CUDA backend (host-side overhead hunting)."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smcp_b200 as S
from smcp_b200 import solvers
from smcp_b200.device import Context
n, m, bw, W, K = (int(a) for a in sys.argv[1:6])
P = S.band_SDP(n, m, bw, seed=0)
ctx = Context.get()
solvers.options["maxiters"] = W + K
solvers.options["show_progress"] = False
pr = cProfile.Profile()
stamps = {}
def hook(name, it):
    ctx.sync()
    stamps[it] = time.perf_counter()
    if it == W:
        pr.enable()
    if it == W + K:
        pr.disable()
solvers._iteration_hook = hook
sol = P.solve_feas(primalstart={"x": P._X0})
pr.disable()
its = sorted(stamps)
print("iterations", sol["iterations"], "status", sol["status"])
for a, b in zip(its[:-1], its[1:]):
    print("iteration %d: %.2f ms" % (a, 1e3 * (stamps[b] - stamps[a])))
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
