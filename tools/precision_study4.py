"""Does flushing fp16 subnormal operands (lo parts) explain the GPU's primal error?"""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from oracle.net import Net, init_params, rel_l2
from oracle import cosmology as oc
torch.set_num_threads(8)
P = init_params(42)
g = np.load('tests/golden/cand_noncubic_6.npz')
x = np.random.default_rng(int(g['seed'])).standard_normal((1, 3, 104, 112, 120), dtype=np.float32)
z, Om = 1.0, 0.25
Dz = float(np.float32(oc.growth_factor(z, Om))); vf = float(np.float32(oc.vel_norm(z, Om)))
h = torch.float16
TINY = 2.0 ** -14
def r1(t): return t.to(h).to(t.dtype)
def r2(t):
    a = t.to(h).to(t.dtype); return a + (t - a).to(h).to(t.dtype)
def r2_ftz(t):
    a = t.to(h).to(t.dtype); l = (t - a).to(h).to(t.dtype)
    l = torch.where(l.abs() < TINY, torch.zeros_like(l), l)
    return a + l
def r2_scaled(S):
    def f(t):
        ts = t * S; a = ts.to(h).to(t.dtype); l = (ts - a).to(h).to(t.dtype); return (a + l) / S
    return f
cfgs = {
 'split (gradual underflow)': dict(xp=r2, wp=r2, xt=r1, wt=r1, dw=r1, dx=r1),
 'split, FTZ on x-lo and W-lo': dict(xp=r2_ftz, wp=r2_ftz, xt=r1, wt=r1, dw=r1, dx=r1),
 'split, FTZ on x-lo only': dict(xp=r2_ftz, wp=r2, xt=r1, wt=r1, dw=r1, dx=r1),
 'split, W scaled x256 (no subnormal W-lo)': dict(xp=r2, wp=r2_scaled(256.0), xt=r1, wt=r1, dw=r1, dx=r1),
 'split, W x256 and x x16': dict(xp=r2_scaled(16.0), wp=r2_scaled(256.0), xt=r1, wt=r1, dw=r1, dx=r1),
}
for name, ops in cfgs.items():
    t = time.time()
    out = Net(True, True, torch.float32, ops=ops).forward(P, x, np.float32(Om), Dz, vf)
    print('%-42s disp %.3e vel %.3e (%.0fs)' % (name, rel_l2(out[0].numpy(), g['disp']), rel_l2(out[1].numpy(), g['vel']), time.time()-t), flush=True)
