"""Halo-amortising tiling on the 512^3 Style+vel box: merged tiles (2,2,1) and (2,2,2) against the
64-subbox run — bit-identity, device-resident and end-to-end particles/s (three repetitions each)."""
import json, sys, time, numpy as np, torch
sys.path.insert(0, '.')
import jax_nbody_emulator_with_dj_b200 as nb
from jax_nbody_emulator_with_dj_b200._engine import Engine
P = nb.init_params(42)
S = 512
host_t = torch.empty((3, S, S, S), dtype=torch.float32, pin_memory=True)
box = host_t.numpy()
np.random.default_rng(7).standard_normal(box.shape, dtype=np.float32, out=box)
cfg = nb.SubboxConfig(size=(S, S, S), ndiv=(4, 4, 4))
proc = nb.SubboxProcessor(nb.StyleNBodyEmulatorVelCore(), P, cfg)
d0, v0 = proc.process_box(box, 0.5, 0.3, show_progress=False)
d0, v0 = d0.copy(), v0.copy()
res = {}
for m in ((1, 1, 1), (2, 2, 1), (2, 2, 2)):
    kw = {} if m == (1, 1, 1) else dict(merge=m)
    try:
        ts = []
        for _ in range(4):
            t = time.perf_counter()
            d, v = proc.process_box(box, 0.5, 0.3, show_progress=False, **kw)
            ts.append(time.perf_counter() - t)
        ok = bool(np.array_equal(d, d0) and np.array_equal(v, v0))
        res[str(m)] = dict(bit_identical=ok, e2e_s=[round(t, 3) for t in ts], e2e_Mps=round(S ** 3 / min(ts[1:]) / 1e6, 1))
        print('merge', m, res[str(m)], flush=True)
    except Exception as e:
        print('merge', m, 'FAILED', str(e)[:300], flush=True)
# device-resident, as bench.py's `value` is measured
eng = Engine.get(0)
Dz, vf = np.float32(nb.growth_factor(0.5, 0.3)), np.float32(nb.vel_norm(0.5, 0.3))
box_dev = host_t.cuda()
dd = torch.zeros((3, S, S, S), dtype=torch.float32, device='cuda'); vd = torch.zeros_like(dd)
for m in ((2, 2, 1), (2, 2, 2)):
    try:
        mcfg, (mc, ma, mp) = proc.merged_config(m)
        n = int(mcfg.n_subboxes)
        run = lambda: eng.process_box_dev(box_dev, mcfg.size, mcfg.crop_size, mp, mc, ma, 0, n, Dz, vf, dd, vd)
        run(); torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e-3)
        ok = bool(np.array_equal(dd.cpu().numpy(), d0))
        res['dev' + str(m)] = dict(bit_identical=ok, s=[round(t, 3) for t in ts], Mps=round(S ** 3 / min(ts) / 1e6, 1), tiles=n)
        print('device-resident merge', m, res['dev' + str(m)], flush=True)
    except Exception as e:
        print('device merge', m, 'FAILED', str(e)[:300], flush=True)
res['hbm_used_GB'] = round((torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 1e9, 1)
print(json.dumps(res))
open('gpurun_out/merge_check.json', 'w').write(json.dumps(res, indent=1))
