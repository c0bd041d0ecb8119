"""Bit-level regression check between two builds of libnbe_b200.so (A/B runs with NBE_LIB):
python tools/bitcheck.py out.npz            -> run a few forwards, save the outputs
python tools/bitcheck.py out.npz ref.npz    -> ... and compare with a previous run, bit for bit"""
import sys
import numpy as np
sys.path.insert(0, '.')
import jax_nbody_emulator_with_dj_b200 as nb

P = nb.init_params(42)
out = {}
x = np.random.default_rng(3).standard_normal((1, 3, 104, 112, 120), dtype=np.float32)
for tag, (z, Om) in {'a': (0.5, 0.3), 'b': (2.0, 0.2)}.items():
    print('step', tag, flush=True)
    d, v = nb.StyleNBodyEmulatorVelCore().apply(P, x, Om, nb.growth_factor(z, Om), nb.vel_norm(z, Om))
    out['d' + tag], out['v' + tag] = d, v
x = np.random.default_rng(4).standard_normal((1, 3, 128, 128, 128), dtype=np.float32)
print('step pmv', flush=True)
pmv = nb.modulate_emulator_parameters_vel(P, 1.0, 0.3)
print('step p', flush=True)
d, v = nb.NBodyEmulatorVelCore().apply(pmv, x, nb.growth_factor(1.0, 0.3), nb.vel_norm(1.0, 0.3))
out['dp'], out['vp'] = d, v
print('step n', flush=True)
out['dn'] = nb.StyleNBodyEmulatorCore().apply(P, x, 0.3, nb.growth_factor(1.0, 0.3))
np.savez(sys.argv[1], **out)
if len(sys.argv) > 2:
    ref = np.load(sys.argv[2])
    bad = [k for k in out if not np.array_equal(out[k], ref[k])]
    for k in bad:
        print(k, 'differs: rel', float(np.linalg.norm(out[k] - ref[k]) / np.linalg.norm(ref[k])))
    print('BIT-IDENTICAL' if not bad else 'DIFFERENT: %s' % bad)
