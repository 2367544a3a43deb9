"""Per-kernel-family device time of ONE chordal operation on a BASELINE pattern (CUDA events per launch).

    python scripts/op_profile.py C3 completion cholesky hessian hessian_inv hessian_prep_inv
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from smcp_b200 import solvers
from smcp_b200.solvers import _Problem, _read_options
from smcp_b200.device import Context

cfg = sys.argv[1]
ops_wanted = sys.argv[2:] or ["completion", "cholesky", "hessian", "hessian_inv"]
P = bench.build_problem({"C3": "rand_n2000_m10000", "C2": "band_n5000_m1000_bw5"}[cfg])
solvers.options["show_progress"] = False
pr = _Problem(P.A, P.b, _read_options(P.n, True), "chol", None)
dev, symb = pr.ops, pr.symb
ctx = Context.get()
X = pr.from_original(P._X0).buf
S = pr.from_original(P._S0).buf
L = dev.clone(X)
dev.completion(L)
tok = dev.hessian_factor(L, X)
U = dev.clone(S)
dev.hessian_apply(tok, [U], True)      # warm-up incl. the lazily built chol(Y_aa)


def run(name):
    if name == "completion":
        T = dev.clone(X); return lambda: dev.completion(T), T
    if name == "cholesky":
        T = dev.clone(S); return lambda: dev.cholesky(T), T
    if name == "hessian":
        T = dev.clone(S); return lambda: dev.hessian_apply(tok, [T], False), T
    if name == "hessian_inv":
        T = dev.clone(S); return lambda: dev.hessian_apply(tok, [T], True), T
    if name == "kkt_assemble":
        return (lambda: dev.schur_assemble(tok)), None
    if name == "kkt_factor":
        return (lambda: dev.schur_factor(tok)), None
    if name == "hessian_prep_inv":
        return (lambda: (dev.hessian_factor(L, X), dev.hessian_apply(dev._tok, [dev.clone(S)], True))), None
    raise SystemExit("unknown op " + name)


for name in ops_wanted:
    fn, _ = run(name)
    fn()
    ctx.sync()
    fn, _ = run(name)
    ctx.prof_reset()
    ctx.prof_enable(True)
    fn()
    ctx.prof_enable(False)
    rows = []
    for nm in ctx.prof_names():
        ms, cnt = ctx.prof_get(nm)
        if cnt:
            rows.append((ms, nm, cnt))
    tot = sum(r[0] for r in rows)
    print("%s on %s: %.3f ms in %d launches (per-launch event timing)" % (name, cfg, tot, sum(r[2] for r in rows)))
    for ms, nm, cnt in sorted(rows, reverse=True)[:10]:
        print("    %-22s %9.3f ms %6d launches %9.1f us/launch" % (nm, ms, cnt, 1e3 * ms / cnt))
