"""GPU: evaluate every golden fixture with the product; prints rel-L2 errors next to the
fixture's own conditioning numbers (used to choose well-conditioned test inputs)."""
import glob, os, sys
import numpy as np
sys.path.insert(0, '.')
import jax_nbody_emulator_with_dj_b200 as nb
from oracle.net import rel_l2
P = nb.init_params(42)
for f in sorted(glob.glob('tests/golden/*.npz')):
    g = np.load(f)
    if 'size' in g.files:
        continue
    z, Om = np.atleast_1d(g['z']).astype(np.float32), np.atleast_1d(g['Om']).astype(np.float32)
    shape = tuple(int(s) for s in g['shape']) if 'shape' in g.files else (int(g['N']),) * 3
    B = g['disp'].shape[0]
    x = np.random.default_rng(int(g['seed'])).standard_normal((B, 3) + shape, dtype=np.float32)
    Dz, vf = nb.growth_factor(z, Om), nb.vel_norm(z, Om)
    d, v = nb.StyleNBodyEmulatorVelCore().apply(P, x, Om, Dz, vf)
    if 'stride' in g.files:
        st = int(g['stride']); d = d[:, :, ::st, ::st, ::st]; v = v[:, :, ::st, ::st, ::st]
    extra = ' cond_vel %.1e' % g['cond_vel'] if 'cond_vel' in g.files else ''
    extra += ' emu_vel %.1e' % g['emu_vel'] if 'emu_vel' in g.files else ''
    print('%-34s disp %.2e vel %.2e%s' % (os.path.basename(f), rel_l2(d, g['disp']), rel_l2(v, g['vel']), extra), flush=True)
