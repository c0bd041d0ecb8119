"""Run a few single-subbox forwards (for ncu / quick per-launch timing).
python tools/run_one.py [N=224] [reps=3] [precision=split] [vel|novel]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import jax_nbody_emulator_with_dj_b200 as nb
from jax_nbody_emulator_with_dj_b200._engine import Engine
N = int(sys.argv[1]) if len(sys.argv) > 1 else 224
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
prec = sys.argv[3] if len(sys.argv) > 3 else 'split'
vel = (sys.argv[4] if len(sys.argv) > 4 else 'vel') == 'vel'
params = nb.init_params(42)
x = torch.randn((1, 3, N, N, N), device='cuda', generator=torch.Generator('cuda').manual_seed(1))
model = nb.StyleNBodyEmulatorVelCore() if vel else nb.StyleNBodyEmulatorCore()
model.precision = prec
Dz, vf = nb.growth_factor(0.5, 0.3), nb.vel_norm(0.5, 0.3)
args = (0.3, Dz, vf) if vel else (0.3, Dz)
print('first apply...', file=sys.stderr, flush=True)
model.apply(params, x, *args)
torch.cuda.synchronize()
print('first apply done', file=sys.stderr, flush=True)
eng = Engine.get()
eng.set_profiling(True)
t = time.time()
for _ in range(reps):
    model.apply(params, x, *args)
torch.cuda.synchronize()
print('%.1f ms per forward' % ((time.time() - t) / reps * 1e3))
tot_t = tot_f = 0
for n, ms, f in eng.get_profile():
    tot_t += ms; tot_f += f
    print('%-26s %8.3f ms %8.1f TFLOP/s' % (n, ms, f / (ms * 1e-3) / 1e12 if ms > 0 else 0))
print('total %.2f ms, %.1f TFLOP/s algorithmic' % (tot_t, tot_f / (tot_t * 1e-3) / 1e12))
