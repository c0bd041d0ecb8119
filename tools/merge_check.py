import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import jax_nbody_emulator_with_dj_b200 as nb
P = nb.init_params(42)
S = 512
box = np.random.default_rng(7).standard_normal((3, S, S, S), dtype=np.float32)
proc = nb.SubboxProcessor(nb.StyleNBodyEmulatorVelCore(), P, nb.SubboxConfig(size=(S, S, S), ndiv=(4, 4, 4)))
d0, v0 = proc.process_box(box, 0.5, 0.3, show_progress=False)
for m in ((2, 2, 1), (2, 2, 2)):
    try:
        d, v = proc.process_box(box, 0.5, 0.3, show_progress=False, merge=m)
        t = time.perf_counter()
        d, v = proc.process_box(box, 0.5, 0.3, show_progress=False, merge=m)
        t = time.perf_counter() - t
        print('merge', m, 'bit-identical', bool(np.array_equal(d, d0) and np.array_equal(v, v0)), '%.3f s  %.1f M particles/s e2e' % (t, S ** 3 / t / 1e6), flush=True)
    except Exception as e:
        print('merge', m, 'FAILED', str(e)[:300], flush=True)
print('hbm used GB', (torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 1e9)
