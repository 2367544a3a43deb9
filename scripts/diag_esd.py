"""Convergence diagnostic for the self-dual embedding on mtxnorm / max-cut problems.
   python scripts/diag_esd.py mtx p q r | maxcut n"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import smcp_b200 as S
from smcp_b200 import solvers
kind = sys.argv[1]
if kind == "mtx":
    p, q, r = (int(a) for a in sys.argv[2:5])
    P = S.mtxnorm_SDP(p, q, r, density=1.0, seed=0)
else:
    n = int(sys.argv[2])
    rng = np.random.default_rng(0)
    P = S.maxcut_SDP(n, rng.integers(0, n, size=(3 * n // 2, 2)))
solvers.options["show_progress"] = False
solvers.options["maxiters"] = 70
t0 = time.time()
sol = P.solve_esd(kktsolver="chol")
print("%s %s BIG_FLOPS=%s: status %s, %d iterations, %.2f s; pobj %.10e dobj %.10e gap %.2e pres %.1e dres %.1e" % (
    kind, sys.argv[2:], os.environ.get("SMCP_B200_BIG_FLOPS"), sol["status"], sol["iterations"], time.time() - t0,
    sol["primal objective"], sol["dual objective"], sol["gap"], sol["primal infeasibility"] or 0, sol["dual infeasibility"] or 0))
for r_ in sol["trace"]:
    if r_["iter"] % 6 == 0 or r_["iter"] >= sol["iterations"] - 2:
        f = lambda k: float(r_.get(k) or 0.0)
        print("  it %3d gap %.2e pres %.1e dres %.1e step %.2e" % (r_["iter"], f("gap"), f("pres"), f("dres"), f("step")))
