"""Numerics study: emulate fp16 / bf16 storage of activations and packed weights (fp32
accumulation) in the CPU oracle and compare with the fp64 oracle.  Decides the operand dtype
of the tcgen05 path.  Usage: python tools/precision_study.py [N]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from oracle.net import Net, init_params, rel_l2
from oracle import cosmology as cosmo

N = int(sys.argv[1]) if len(sys.argv) > 1 else 104
torch.set_num_threads(8)
params = init_params(42)
rng = np.random.default_rng(1234)
x = rng.standard_normal((1, 3, N, N, N), dtype=np.float32)
z, Om = 0.5, 0.3
Dz = float(cosmo.growth_factor(z, Om)); vf = float(cosmo.vel_norm(z, Om))
print('Dz', Dz, 'vel_fac', vf)
t = time.time()
ref = Net(True, True, torch.float64).forward(params, x, Om, Dz, vf)
print('fp64 %.1fs' % (time.time() - t))
ref = [r.numpy() for r in ref]
def q(dt):
    return lambda t: t.to(dt).to(t.dtype)
for name, qa, qw in [('fp32', None, None), ('fp16', q(torch.float16), q(torch.float16)),
                     ('bf16', q(torch.bfloat16), q(torch.bfloat16)),
                     ('fp16act+fp32w', q(torch.float16), None), ('fp32act+fp16w', None, q(torch.float16))]:
    t = time.time()
    out = Net(True, True, torch.float32, q_act=qa, q_w=qw).forward(params, x, Om, Dz, vf)
    print('%-14s disp %.3e vel %.3e  (%.1fs)' % (name, rel_l2(out[0].numpy(), ref[0]), rel_l2(out[1].numpy(), ref[1]), time.time() - t), flush=True)
