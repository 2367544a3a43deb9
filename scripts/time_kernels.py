"""Device time of every chordal kernel family on a band pattern, single matrix and batch
(CUDA events around each launch; prints ms per launch and the effective GB/s)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from smcp_b200.symbolic import Symbolic, lower_pattern
from smcp_b200.device import DeviceBackend, Context

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
bw = int(sys.argv[2]) if len(sys.argv) > 2 else 5
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
reps = 3
I = np.concatenate([np.arange(j, min(j + bw + 1, n)) for j in range(n)])
J = np.concatenate([np.full(min(j + bw + 1, n) - j, j) for j in range(n)])
cp, ri = lower_pattern(n, I, J)
symb = Symbolic(n, cp, ri)
dev = DeviceBackend(symb)
ctx = Context.get()
rng = np.random.default_rng(0)
v = 0.05 * rng.standard_normal(symb.nvp)
v[symb.diag_vec] = 2.0
S = dev.from_vec(v)
L = dev.clone(S)
dev.cholesky(L)
Y = dev.clone(L)
dev.projected_inverse(Y)
tok = dev.hessian_factor(L, Y)
U1 = dev.from_vec(rng.standard_normal(symb.nvp))
Ub = dev.alloc_batch(batch)
host = np.zeros(symb.nblk)
host[symb.vec2blk] = rng.standard_normal(symb.nvp)
dev.set_batch(Ub, np.repeat(host[None, :], min(batch, 8), axis=0))
Sb = dev.alloc_batch(batch)
sblk = dev.get_blk(S)
for k in range(batch):
    dev.lib.smcp_csp_copy(dev.sym, dev._slot(Sb, k), S, 1)
ctx.prof_enable(True)
for it in range(reps + 1):
    if it == 1:
        ctx.prof_reset()
    x = dev.clone(S); dev.cholesky(x); dev.llt(x)
    y = dev.clone(L); dev.projected_inverse(y); dev.completion(y)
    u = dev.clone(U1); dev.hessian_apply(tok, [u], False); dev.hessian_apply(tok, [u], True)
    dev.hessian_batch(tok, Ub, batch, False)
    dev.cholesky_batch(Sb, batch); dev.llt_batch(Sb, batch)
    for b in (x, y, u):
        dev.release(b)
ctx.prof_enable(False)
names = sorted(ctx.prof_names())
print("pattern: band n=%d bw=%d, nsn=%d, nblk=%d, batch=%d" % (n, bw, symb.nsn, symb.nblk, batch))
for nm in names:
    ms, cnt = ctx.prof_get(nm)
    if cnt:
        w = ctx.prof_get_work(nm) / cnt
        per = ms / cnt
        print("%-26s %9.3f ms/launch  %5d launches  %7.0f matrices/launch  %8.1f GB/s (16 B/entry)" % (
            nm, per, cnt, w, 16.0 * symb.nvp * w / (per * 1e-3) / 1e9))
