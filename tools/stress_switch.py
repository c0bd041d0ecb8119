"""Stress the model-switch path (vel <-> displacement-only, Style <-> premodulated) without launch blocking."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import jax_nbody_emulator_with_dj_b200 as nb
P = nb.init_params(42)
x128 = np.random.default_rng(4).standard_normal((1, 3, 128, 128, 128), dtype=np.float32)
x104 = np.random.default_rng(3).standard_normal((1, 3, 104, 112, 120), dtype=np.float32)
Dz, vf = nb.growth_factor(1.0, 0.3), nb.vel_norm(1.0, 0.3)
pmv = nb.modulate_emulator_parameters_vel(P, 1.0, 0.3)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
ref = {}
for it in range(n):
    for tag, f in (('s', lambda: nb.StyleNBodyEmulatorVelCore().apply(P, x104, 0.3, Dz, vf)),
                   ('p', lambda: nb.NBodyEmulatorVelCore().apply(pmv, x128, Dz, vf)),
                   ('n', lambda: nb.StyleNBodyEmulatorCore().apply(P, x128, 0.3, Dz)),
                   ('S', lambda: nb.StyleNBodyEmulatorVelCore().apply(P, x128, 0.3, Dz, vf))):
        try:
            out = f()
            torch.cuda.synchronize()
        except Exception as e:
            print('FAIL iteration', it, 'step', tag, str(e)[:200], flush=True)
            sys.exit(1)
        out = out if isinstance(out, tuple) else (out,)
        if tag in ref:
            assert all(np.array_equal(a, b) for a, b in zip(out, ref[tag])), ('not reproducible', it, tag)
        ref[tag] = out
print('ok', n, 'iterations')
