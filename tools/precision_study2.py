"""Numerics study 2: split-precision primal.  python tools/precision_study2.py [N]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from oracle.net import Net, init_params, rel_l2
from oracle import cosmology as cosmo
N = int(sys.argv[1]) if len(sys.argv) > 1 else 104
torch.set_num_threads(8)
params = init_params(42)
x = np.random.default_rng(1234).standard_normal((1, 3, N, N, N), dtype=np.float32)
z, Om = 0.5, 0.3
Dz = float(cosmo.growth_factor(z, Om)); vf = float(cosmo.vel_norm(z, Om))
ref = [r.numpy() for r in Net(True, True, torch.float64).forward(params, x, Om, Dz, vf)]
def r1(dt): return lambda t: t.to(dt).to(t.dtype)
def r2(dt):
    def f(t):
        h = t.to(dt).to(t.dtype); return h + (t - h).to(dt).to(t.dtype)
    return f
h, b = torch.float16, torch.bfloat16
cfgs = {
 'fp16 1-pass': dict(xp=r1(h), wp=r1(h), xt=r1(h), wt=r1(h), dw=r1(h), dx=r1(h)),
 'fp16 split x,w (3-prod) + fp16 tangent': dict(xp=r2(h), wp=r2(h), xt=r1(h), wt=r1(h), dw=r1(h), dx=r1(h)),
 'fp16 split x only': dict(xp=r2(h), wp=r1(h), xt=r1(h), wt=r1(h), dw=r1(h), dx=r1(h)),
 'bf16 split x,w + bf16 tangent': dict(xp=r2(b), wp=r2(b), xt=r1(b), wt=r1(b), dw=r1(b), dx=r1(b)),
 'fp16 split primal + split tangent ops': dict(xp=r2(h), wp=r2(h), xt=r2(h), wt=r2(h), dw=r2(h), dx=r2(h)),
}
for name, ops in cfgs.items():
    t = time.time()
    out = Net(True, True, torch.float32, ops=ops).forward(params, x, Om, Dz, vf)
    print('%-42s disp %.3e vel %.3e (%.0fs)' % (name, rel_l2(out[0].numpy(), ref[0]), rel_l2(out[1].numpy(), ref[1]), time.time()-t), flush=True)
