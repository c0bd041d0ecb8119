import sys, numpy as np, torch
sys.path.insert(0, '.')
import jax_nbody_emulator_with_dj_b200 as nb
S = 512
cfg = nb.SubboxConfig(size=(S, S, S), ndiv=(4, 4, 4))
proc = nb.SubboxProcessor(nb.StyleNBodyEmulatorVelCore(), nb.init_params(42), cfg)
host_t = torch.empty((3, S, S, S), dtype=torch.float32, pin_memory=True)
host_t.normal_()
host = host_t.numpy()
for i in range(2):
    r = proc.process_box(host, 0.5, 0.3, show_progress=False, shard=(0, 8), gather="none"); del r
