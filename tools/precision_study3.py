"""Numerics study 3 on the z=1, Om=0.25 non-cubic case: what limits the velocity?"""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from oracle.net import Net, init_params, rel_l2
from oracle import cosmology as oc
torch.set_num_threads(8)
P = init_params(42)
x = np.random.default_rng(5).standard_normal((1, 3, 104, 112, 120), dtype=np.float32)
g = np.load('tests/golden/noncubic.npz')
z, Om = 1.0, 0.25
Dz = float(np.float32(oc.growth_factor(z, Om))); vf = float(np.float32(oc.vel_norm(z, Om)))
h = torch.float16
def r1(t): return t.to(h).to(t.dtype)
def r2(t):
    a = t.to(h).to(t.dtype); return a + (t - a).to(h).to(t.dtype)
def noisy(eps):
    gen = torch.Generator().manual_seed(0)
    def f(t): return r2(t) * (1 + eps * torch.randn(t.shape, generator=gen, dtype=t.dtype))
    return f
cfgs = {
 'A split primal, fp16 tangent': dict(xp=r2, wp=r2, xt=r1, wt=r1, dw=r1, dx=r1),
 'B A + primal noise 1e-5': dict(xp=noisy(1e-5), wp=r2, xt=r1, wt=r1, dw=r1, dx=r1),
 'C split primal, split dx only': dict(xp=r2, wp=r2, xt=r1, wt=r1, dw=r1, dx=r2),
 'D split primal, tangent all split': dict(xp=r2, wp=r2, xt=r2, wt=r2, dw=r2, dx=r2),
 'E D + primal noise 1e-5': dict(xp=noisy(1e-5), wp=r2, xt=r2, wt=r2, dw=r2, dx=r2),
}
for name, ops in cfgs.items():
    t = time.time()
    out = Net(True, True, torch.float32, ops=ops).forward(P, x, np.float32(Om), Dz, vf)
    print('%-36s disp %.3e vel %.3e (%.0fs)' % (name, rel_l2(out[0].numpy(), g['disp']), rel_l2(out[1].numpy(), g['vel']), time.time()-t), flush=True)
