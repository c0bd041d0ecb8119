"""Small driver for compute-sanitizer (one tool per gpurun call):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
Exercises the host path of nbe_process_box on tiny periodic boxes -- windowed incremental upload on
its own stream, plan / arena re-allocation inside a call sequence (the case the withdrawn round-1
split upload faulted on), sharded ranges, block-layout outputs -- and every conv kernel instance of
the Style+vel and displacement-only nets at 104^3."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jax_nbody_emulator_with_dj_b200 as nb

P = nb.init_params(42)
f = lambda shape, seed: np.random.default_rng(seed).standard_normal(shape, dtype=np.float32)
res = []
for size, ndiv in (((8, 8, 16), (1, 1, 2)), ((8, 16, 24), (1, 1, 1)), ((8, 8, 16), (1, 1, 2)), ((16, 8, 8), (2, 1, 1))):
    cfg = nb.SubboxConfig(size=size, ndiv=ndiv)
    proc = nb.SubboxProcessor(nb.StyleNBodyEmulatorVelCore(), P, cfg)
    box = f((3,) + size, 3)
    d, v = proc.process_box(box, 0.5, 0.3, show_progress=False)
    if int(cfg.n_subboxes) > 1:
        acc = np.zeros_like(d)
        for r in range(2):
            acc += proc.process_box(box, 0.5, 0.3, show_progress=False, shard=(r, 2), gather="none")[0]
        assert np.array_equal(acc, d)
    res.append((size, float(np.abs(d).mean()), float(np.abs(v).mean())))
    proc.close()
d1 = nb.SubboxProcessor(nb.StyleNBodyEmulatorCore(), P, nb.SubboxConfig(size=(8, 8, 8), ndiv=(1, 1, 1))).process_box(
    f((3, 8, 8, 8), 4), 0.5, 0.3, show_progress=False)
torch.cuda.synchronize()
print("sanitize_small ok", res, float(np.abs(d1).mean()))
