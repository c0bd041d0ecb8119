"""Time the pre-/post-emulator kernels on one B200 (CUDA events, warm-up 3, 10 reps, inputs > L2).
python tools/bench_density.py [n=512]"""
import sys, json
import torch
sys.path.insert(0, '.')
import jax_nbody_emulator_with_dj_b200 as nb
from jax_nbody_emulator_with_dj_b200._engine import Engine
import ctypes as C

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
eng = Engine.get()
g = torch.Generator('cuda').manual_seed(1)
psi = torch.randn((3, n, n, n), device='cuda', generator=g) * 1.5
delta = torch.empty((n, n, n), device='cuda')
st = torch.cuda.current_stream().cuda_stream
dims = (C.c_int32 * 3)(n, n, n)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


out = {"n": n}
np_ = n ** 3
for order, name in ((2, "CIC"), (3, "TSC"), (4, "PCS")):
    ms = timed(lambda: eng._ck(eng.lib.nbe_density_from_psi(eng.h, psi.data_ptr(), dims, float(n), n, order, delta.data_ptr(), st)))
    # algorithmic bytes: read 3 fp32 per particle, zero + read-modify-write the mesh once, p^3 fp32 reductions per particle
    alg = np_ * 12 + 3 * np_ * 4
    out[f"paint_{name}"] = {"ms": ms, "particles_per_s": np_ / ms * 1e3, "reductions_per_s": np_ * order ** 3 / ms * 1e3,
                            "alg_GBps": alg / ms / 1e6}
dk = torch.fft.rfftn(delta).contiguous()
bins = torch.empty((3, int(3 ** 0.5 * (n // 2)) + 1), device='cuda', dtype=torch.float64)
ms = timed(lambda: eng._ck(eng.lib.nbe_pk_bins(eng.h, dk.data_ptr(), n, 2, bins.shape[1], bins.data_ptr(), st)))
out["pk_bins"] = {"ms": ms, "alg_GBps": dk.numel() * 8 / ms / 1e6}
ms = timed(lambda: eng._ck(eng.lib.nbe_mas_deconvolve(eng.h, dk.data_ptr(), n, 2, st)))
out["mas_deconvolve"] = {"ms": ms, "alg_GBps": 2 * dk.numel() * 8 / ms / 1e6}
pk = torch.empty((3,) + tuple(dk.shape), device='cuda', dtype=torch.complex64)
ms = timed(lambda: eng._ck(eng.lib.nbe_za_psi_k(eng.h, dk.data_ptr(), n, float(n), pk.data_ptr(), st)))
out["za_psi_k"] = {"ms": ms, "alg_GBps": 4 * dk.numel() * 8 / ms / 1e6}
ms = timed(lambda: torch.fft.rfftn(delta))
out["cufft_rfftn"] = {"ms": ms}
ms = timed(lambda: nb.power_spectrum(delta, float(n), MAS="CIC"), reps=3)
out["power_spectrum_total"] = {"ms": ms}
print(json.dumps(out))
