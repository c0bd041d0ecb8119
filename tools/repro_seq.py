import sys, numpy as np
sys.path.insert(0, '.')
import jax_nbody_emulator_with_dj_b200 as nb
P = nb.init_params(42)
seq = sys.argv[1]
x128 = np.random.default_rng(4).standard_normal((1, 3, 128, 128, 128), dtype=np.float32)
x104 = np.random.default_rng(3).standard_normal((1, 3, 104, 112, 120), dtype=np.float32)
Dz, vf = nb.growth_factor(1.0, 0.3), nb.vel_norm(1.0, 0.3)
pmv = None
from jax_nbody_emulator_with_dj_b200._engine import Engine
for c in seq:
    if c == 's': nb.StyleNBodyEmulatorVelCore().apply(P, x104, 0.3, Dz, vf)
    if c == 'b': nb.StyleNBodyEmulatorVelCore().apply(P, x104, 0.2, nb.growth_factor(2.0, 0.2), nb.vel_norm(2.0, 0.2))
    if c == 'S': nb.StyleNBodyEmulatorVelCore().apply(P, x128, 0.3, Dz, vf)
    if c == 'P': pmv = nb.modulate_emulator_parameters_vel(P, 1.0, 0.3)
    if c == 'p': nb.NBodyEmulatorVelCore().apply(pmv or nb.modulate_emulator_parameters_vel(P, 1.0, 0.3), x128, Dz, vf)
    if c == 'n': nb.StyleNBodyEmulatorCore().apply(P, x128, 0.3, Dz)
    if c == 'm': nb.StyleNBodyEmulatorCore().apply(P, x104, 0.3, Dz)
    e = Engine.get()
    print('step', c, 'fold', e.lib.nbe_fold_active(e.h), flush=True)
print('ok', seq)
