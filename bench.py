#!/usr/bin/env python
"""Headline benchmark: output particles/s of the 512^3 Style+velocity box (BASELINE.json
config 3: create_emulator(compute_vel=True), ndiv=(4,4,4) = 64 subboxes of 224^3 -> 128^3),
sharded over N GPUs (one process per GPU, no data-path collective).

  python bench.py --gpus 1 --steps K --warmup W            # our arm
  python bench.py --impl reference --gpus N ...            # CPU arm (oracle port, see below)

One "step" = one pass over the rank's share of the box.  `value` is measured with the input
box and outputs resident in HBM (CUDA events on the launching stream, max over ranks); `e2e`
is the DEFAULT public call -- ``SubboxProcessor.process_box(host_box, z, Om)`` -- with HOST numpy
buffers, H2D and D2H inside the timed region, ending with the whole (3, S, S, S) displacement and
velocity boxes in ONE pair of host arrays.  At N > 1 that call is made by rank 0 alone, which drives
all N GPUs from one process (one host thread per GPU, nbe_process_box_multi) while the other ranks
wait on a host-side (gloo) barrier.  Prints ONE JSON line on rank 0.

Other BASELINE.json configs (not driver-run; lines committed under profiles/):
  --config 2   StyleNBodyEmulatorVelCore, batch 8 x 128^3, per-sample (z, Om), 1 GPU
  --config 4   premodulated NBodyEmulatorVelCore, 1024^3 box, ndiv 8 (512 subboxes), --gpus N in one process
  --config 5   Style+vel 2048^3 box, ndiv 16 (4096 subboxes), streamed as D-slabs through pinned staging
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

A100_README_PARTICLES_PER_S = 512 ** 3 / 44.9      # reference README.md:252 (A100-40GB, fp32, 64 subboxes)
FLOP_PER_SUBBOX_224 = 35.8597e12                   # SURVEY.md Appendix A (Style+vel, N=224)
FLOP_PER_SUBBOX_128 = 4.31668e12


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = sorted(sm)[len(sm) // 2] if sm else None
        return {"sm_mhz": busy, "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_oracle_time(n, threads, repeats=1):
    """Time the CPU oracle (fp32 torch, all host threads) on one Style+vel n^3 subbox."""
    import torch
    from oracle.net import Net, init_params
    from oracle import cosmology as oc
    torch.set_num_threads(threads)
    params = init_params(42)
    x = np.random.default_rng(1234).standard_normal((1, 3, n, n, n), dtype=np.float32)
    Dz, vf = float(oc.growth_factor(0.5, 0.3)), float(oc.vel_norm(0.5, 0.3))
    net = Net(True, True, torch.float32)
    best = None
    for _ in range(repeats):
        t = time.perf_counter()
        with torch.no_grad():
            net.forward(params, x, 0.3, Dz, vf)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    return best


def reference_arm(args):
    """The reference (JAX/Flax) cannot be imported in this image (no jax / flax wheels, no
    network) and has no compilable native code, so the CPU arm is the oracle port: the same
    math in torch-CPU fp32 on all host threads.  Each step = one bounded sample (one subbox)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = 224 if cores >= 64 else 128
    flop = FLOP_PER_SUBBOX_224 if n == 224 else FLOP_PER_SUBBOX_128
    for _ in range(min(args.warmup, 1)):
        cpu_oracle_time(n, cores)
    times = [cpu_oracle_time(n, cores) for _ in range(max(1, min(args.steps, 3)))]
    t = float(np.mean(times))
    # particles/s of the 512^3 / ndiv 4 workload: one 224^3 subbox yields 128^3 particles; a
    # 128^3 sample is converted by its FLOP share (same layers, smaller volume)
    sub_time_224 = t * (FLOP_PER_SUBBOX_224 / flop)
    value = 128 ** 3 / sub_time_224
    sample = (f"one {n}^3 Style+vel subbox per step, fp32 torch-CPU oracle, {cores} threads"
              + ("" if n == 224 else "; scaled to a 224^3 subbox by algorithmic FLOPs (x%.2f)" % (FLOP_PER_SUBBOX_224 / flop)))
    line = {
        "impl": "reference", "metric": "output particles/s, 512^3 Style+vel box (ndiv 4x4x4)", "value": value,
        "unit": "particles/s", "n_gpus": args.gpus, "steps": len(times), "warmup": min(args.warmup, 1),
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "512^3 Style+vel box, ndiv=(4,4,4), 64 subboxes 224^3->128^3 (CPU: bounded sample)",
                   "weights": "fixed-seed random init (pretrained blob absent)"},
        "cpu_baseline": {"value": value, "unit": "particles/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "particles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def e2e_h2d_bytes(cfg, crop_idx, plen, n_sub, ngpu):
    """Bytes the windowed upload moves per box: per GPU the (D planes) x (H rows) x S2 window its share reads."""
    from jax_nbody_emulator_with_dj_b200.subbox import shard_range
    per = sum(plen)
    tabs = crop_idx.reshape(n_sub, per)
    tot = 0
    for g in range(ngpu):
        lo, hi = shard_range(n_sub, g, ngpu)
        if hi > lo:
            nd = np.unique(tabs[lo:hi, :plen[0]]).size
            nh = np.unique(tabs[lo:hi, plen[0]:plen[0] + plen[1]]).size
            tot += 3 * nd * nh * int(cfg.size[2]) * np.dtype(cfg.dtype).itemsize
    return tot


def fill_gaussian(host_t, seed, device=0, chunk=32):
    """Synthetic N(0,1) box straight into a pinned host tensor: generated on the GPU plane-chunk by
    plane-chunk (numpy would need minutes for 1024^3), deterministic in (seed, plane)."""
    import torch
    C_, S0 = host_t.shape[0], host_t.shape[1]
    for d0 in range(0, S0, chunk):
        d1 = min(S0, d0 + chunk)
        g = torch.Generator(device=f"cuda:{device}")
        g.manual_seed(seed * 1000003 + d0)
        t = torch.randn((C_, d1 - d0) + tuple(host_t.shape[2:]), generator=g, device=f"cuda:{device}", dtype=torch.float32)
        for ch in range(C_):              # contiguous per channel on both sides: a plain DMA into pinned memory
            host_t[ch, d0:d1].copy_(t[ch].to(host_t.dtype), non_blocking=True)
    torch.cuda.synchronize(device)


def roofline_from_profile(prof, peak_key="bf16_tflops_sustained"):
    pk, pk_src = peaks()
    conv = [(n, t, f) for (n, t, f) in prof if f > 0 and t > 0]
    if not conv:
        return None
    dom = max(conv, key=lambda r: r[1])
    ach = dom[2] / (dom[1] * 1e-3) / 1e12
    peak = float(pk.get(peak_key, pk.get("bf16_tflops")))
    tot_t = sum(t for _, t, _ in prof)
    tot_f = sum(f for _, _, f in prof)
    return {"bound": "tensor", "kernel": dom[0], "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "traffic": None, "peak_source": pk_src + f" ({peak_key}; fp16 operands run at the bf16 rate)",
            "kernel_share_of_step": dom[1] / tot_t if tot_t else None,
            "net_achieved": tot_f / (tot_t * 1e-3) / 1e12 if tot_t else None,
            "net_frac": tot_f / (tot_t * 1e-3) / 1e12 / peak if tot_t else None,
            "note": "algorithmic FLOPs (SURVEY 8d: primal + two tangent products per layer) / mean launch duration; split precision with the folded tangent (DESIGN 4.2) executes 4/3 of them, 5/3 without"}


def config2(args):
    """BASELINE config 2: StyleNBodyEmulatorVelCore, one batch of 8 x (3,128^3) -> 8 x (3,32^3), (z, Om)
    different per sample, one GPU.  A step = the per-sample weight modulation (8 weight sets) + 8 forward
    passes.  `value`: input resident in HBM (nbe_modulate + nbe_forward); `e2e`: model.apply(params, x, Om,
    Dz, vel_fac) with numpy in / numpy out."""
    import torch
    import jax_nbody_emulator_with_dj_b200 as nb
    from jax_nbody_emulator_with_dj_b200._engine import Engine
    torch.cuda.set_device(0)
    B, N = 8, 128
    zs, Oms = np.linspace(0, 2, B).astype(np.float32), np.linspace(0.1, 0.5, B).astype(np.float32)
    Dz, vf = nb.growth_factor(zs, Oms).astype(np.float32), nb.vel_norm(zs, Oms).astype(np.float32)
    x = np.random.default_rng(1234).standard_normal((B, 3, N, N, N), dtype=np.float32)
    params = nb.init_params(42)
    model = nb.StyleNBodyEmulatorVelCore()
    model.precision = args.precision
    eng = Engine.get(0)
    eng.set_precision(args.precision)
    eng.set_params(params, False, True, model.eps)
    xd = torch.from_numpy(x).cuda()
    particles = B * (N - 96) ** 3

    def step():
        eng._mod_key = None                       # the reference modulates on every call: so does a step
        eng.modulate(Oms, Dz)
        return eng.forward(xd, Dz, vf, True)
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    clocks = ClockSampler(0); clocks.start()
    eng.set_profiling(True); eng.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    launches = eng.launch_count(reset=True)
    prof = eng.get_profile(); eng.set_profiling(False)
    clk = clocks.stop()
    roof = roofline_from_profile(prof, "bf16_tflops")
    if roof:
        roof["note"] += "; per-launch means over the 8 samples of a step; burst peak (each step is ~65 ms)"
    # modulation kernel: HBM GB/s (algorithmic bytes: fp32 params read once + fp32 W, dW and fp16 operand tensors written per sample)
    torch.cuda.synchronize()
    m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    m0.record()
    for _ in range(20):
        eng._mod_key = None
        eng.modulate(Oms, Dz)
    m1.record(); torch.cuda.synchronize()
    mod_ms = m0.elapsed_time(m1) / 20
    e2e = None
    if not args.no_e2e:
        model.apply(params, x, Oms, Dz, vf)
        t0 = time.perf_counter()
        for _ in range(max(1, args.steps)):
            eng._mod_key = None
            d, v = model.apply(params, x, Oms, Dz, vf)
        t = (time.perf_counter() - t0) / max(1, args.steps)
        e2e = {"value": particles / t, "unit": "particles/s", "h2d_bytes_per_step": int(x.nbytes),
               "d2h_bytes_per_step": int(d.nbytes + v.nbytes), "ms_per_step": t * 1e3,
               "api": "StyleNBodyEmulatorVelCore.apply(params, x[8,3,128^3] numpy, Om[8], Dz[8], vel_fac[8]) -> numpy"}
    cpu = None
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        t = cpu_oracle_time(128, cores)
        cpu = {"value": 32 ** 3 / t, "unit": "particles/s", "cores": cores, "kind": "port",
               "sample": "one of the 8 samples (128^3 -> 32^3 Style+vel), fp32 torch-CPU oracle, %.1f s" % t}
    line = {"metric": "output particles/s, StyleNBodyEmulatorVelCore batch 8 x 128^3 -> 32^3, per-sample (z, Om)",
            "value": particles / (ms * 1e-3), "unit": "particles/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16x2-split (fp16 hi+lo operands, fp32 accumulate)" if args.precision == "split" else "f16",
            "data": "synthetic",
            "config": {"workload": "BASELINE config 2: StyleNBodyEmulatorVelCore, batch 8, 128^3 -> 32^3, z = linspace(0,2,8), "
                                   "Om = linspace(0.1,0.5,8); a step = 8 weight modulations + 8 forward passes",
                       "weights": "fixed-seed random init", "l2": "activations of one sample (2.1 GB) exceed L2",
                       "flop_per_step": 8 * FLOP_PER_SUBBOX_128},
            "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
            "modulation_kernel": {"ms": mod_ms, "samples": B,
                                  "algorithmic_GB": (13.4e6 + B * (2 * 13.4e6 + 20.1e6)) / 1e9,
                                  "GBps": (13.4e6 + B * (2 * 13.4e6 + 20.1e6)) / 1e9 / (mod_ms * 1e-3)}}
    print(json.dumps(line), flush=True)


def config_big(args):
    """BASELINE configs 4 / 5, one process driving --gpus N GPUs through the public API; host buffers only
    (there is no HBM-resident variant of a box that is streamed by definition), so `value` is the same
    measurement as `e2e`."""
    import resource
    import torch
    import jax_nbody_emulator_with_dj_b200 as nb
    from jax_nbody_emulator_with_dj_b200._engine import Engine
    ng = min(args.gpus, torch.cuda.device_count())
    devs = list(range(ng))
    z, Om = 0.5, 0.3
    params = nb.init_params(42)
    if args.config == 4:
        S, nd = 1024, 8
        model = nb.NBodyEmulatorVelCore()
        params = nb.modulate_emulator_parameters_vel(params, z, Om)
        name = "premodulated NBodyEmulatorVelCore"
    else:
        S, nd = 2048, 16
        model = nb.StyleNBodyEmulatorVelCore()
        name = "StyleNBodyEmulatorVelCore (Style+vel)"
    model.precision = args.precision
    cfg = nb.SubboxConfig(size=(S, S, S), ndiv=(nd, nd, nd))
    proc = nb.SubboxProcessor(model, params, cfg)
    n_sub = int(cfg.n_subboxes)
    c = int(cfg.crop_size[0])
    clocks = ClockSampler(0)
    free0 = [torch.cuda.mem_get_info(g)[0] for g in devs]
    extra = {}
    if args.config == 4:
        host_t = torch.empty((3, S, S, S), dtype=torch.float32, pin_memory=True)
        fill_gaussian(host_t, 1234)
        host = host_t.numpy()
        crop_idx, add0, plen = cfg.flat_tables()
        for _ in range(max(1, args.warmup)):
            r = proc.process_box(host, z, Om, show_progress=False, devices=devs)
            del r
        for e in [Engine.get(g) for g in devs]:
            e.launch_count(reset=True)
        clocks.start()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            d, v = proc.process_box(host, z, Om, show_progress=False, devices=devs)
            chk = [float(d[:, ::97, ::101, ::103].astype(np.float64).sum()), float(v[:, ::97, ::101, ::103].astype(np.float64).sum())]
            del d, v
        t = (time.perf_counter() - t0) / args.steps
        particles = S ** 3
        h2d, d2h = e2e_h2d_bytes(cfg, crop_idx, plen, n_sub, ng), 2 * 3 * S ** 3 * 4
        api = "SubboxProcessor.process_box(host numpy (3,1024^3), z, Om) [defaults]"
        extra["checksum"] = chk
    else:
        n_slabs = nd if args.slabs is None else min(nd, args.slabs)
        acc = {"sum_d": 0.0, "sum_v": 0.0, "planes": 0, "fill_s": 0.0, "sink_s": 0.0}

        gen = torch.Generator(device="cuda:0")

        def source(planes, out):          # synthetic periodic box: plane p is N(0,1) seeded by p (same values whenever re-read)
            t1 = time.perf_counter()
            ot = torch.from_numpy(out)
            for j0 in range(0, len(planes), 16):
                ps = planes[j0:j0 + 16]
                buf = torch.empty((3, len(ps), S, S), device="cuda:0", dtype=torch.float32)
                for j, pl in enumerate(ps):
                    gen.manual_seed(1234 * 1000003 + int(pl))
                    buf[:, j] = torch.randn((3, S, S), generator=gen, device="cuda:0", dtype=torch.float32)
                for ch in range(3):       # per channel: contiguous on both sides => one DMA into the pinned staging buffer
                    ot[ch, j0:j0 + len(ps)].copy_(buf[ch], non_blocking=True)
            torch.cuda.synchronize(0)
            acc["fill_s"] += time.perf_counter() - t1

        def sink(k, d0, ds, vs):          # the consumer: a strided checksum (stands in for the fp16 / memmap writer)
            t1 = time.perf_counter()
            acc["sum_d"] += float(ds[:, ::8, ::64, ::64].astype(np.float64).sum())
            acc["sum_v"] += float(vs[:, ::8, ::64, ::64].astype(np.float64).sum())
            acc["planes"] += ds.shape[1]
            acc["sink_s"] += time.perf_counter() - t1
        proc.process_box_streamed(source, sink, z, Om, devices=devs, max_slabs=1)       # warm-up: contexts, arenas, staging
        for k in acc:
            acc[k] = 0
        for e in [Engine.get(g) for g in devs]:
            e.launch_count(reset=True)
        clocks.start()
        t0 = time.perf_counter()
        done = proc.process_box_streamed(source, sink, z, Om, devices=devs, max_slabs=n_slabs)
        t = time.perf_counter() - t0
        particles = done * c * S * S
        h2d = done * 3 * (c + 96) * S * S * 4          # whole staged planes are within reach of the GPUs; windows are a subset
        d2h = done * 2 * 3 * c * S * S * 4
        api = "SubboxProcessor.process_box_streamed(plane_source, slab_sink, z, Om): %d of %d D-slabs" % (done, nd)
        extra.update({"slabs_done": done, "slabs_total": nd, "checksum": [acc["sum_d"], acc["sum_v"]],
                      "host_fill_s": acc["fill_s"], "host_sink_s": acc["sink_s"],
                      "staging_pinned_gb": 2 * (3 * (c + 96) + 2 * 3 * c) * S * S * 4 / 1e9})
    clk = clocks.stop()
    launches = sum(Engine.get(g).launch_count() for g in devs)
    used = [(free0[i] - torch.cuda.mem_get_info(g)[0]) / 1e9 for i, g in enumerate(devs)]
    val = particles / t
    line = {"metric": f"output particles/s, {S}^3 box, {name}, ndiv {nd}", "value": val, "unit": "particles/s", "n_gpus": ng,
            "steps": args.steps if args.config == 4 else 1, "warmup": max(1, args.warmup) if args.config == 4 else 1,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f16x2-split (fp16 hi+lo operands, fp32 accumulate)" if args.precision == "split" else "f16",
            "data": "synthetic",
            "config": {"workload": f"BASELINE config {args.config}: {name}, {S}^3 box, ndiv=({nd},{nd},{nd}) = {n_sub} subboxes "
                                   f"{c + 96}^3->{c}^3, {ng} GPU(s) driven by one process",
                       "weights": "fixed-seed random init", "l2": "inputs larger than L2",
                       "note": "value == e2e: host buffers, H2D and D2H inside the timed region"},
            "clocks": clk,
            "e2e": {"value": val, "unit": "particles/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "api": api},
            "gpu_launches": int(launches), "roofline": None, "cpu_baseline": None,
            "memory": {"host_peak_rss_gb": resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e6, "hbm_used_gb_per_gpu": used},
            **extra}
    line["roofline"] = {"bound": "tensor", "achieved": val * 17.10e6 / 1e12 / ng, "unit": "TFLOP/s per GPU (algorithmic, whole job incl. host path)",
                        "peak": peaks()[0].get("bf16_tflops_sustained"), "frac": val * 17.10e6 / 1e12 / ng / peaks()[0].get("bf16_tflops_sustained"),
                        "traffic": None, "note": "17.10 MFLOP per output particle (SURVEY 8d, crop 128 Style/premod+vel)"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--box", type=int, default=512)
    ap.add_argument("--ndiv", type=int, default=4)
    ap.add_argument("--precision", default="split", choices=["split", "fp16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-alt", action="store_true")
    ap.add_argument("--kernels", action="store_true", help="print the per-launch table to stderr")
    ap.add_argument("--config", type=int, default=3, choices=[2, 3, 4, 5], help="BASELINE.json config (1-based)")
    ap.add_argument("--slabs", type=int, default=None, help="config 5: process only the first SLABS D-slabs")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    if args.config == 2:
        return config2(args)
    if args.config in (4, 5):
        return config_big(args)

    import torch
    import jax_nbody_emulator_with_dj_b200 as nb
    from jax_nbody_emulator_with_dj_b200._engine import Engine
    from jax_nbody_emulator_with_dj_b200.subbox import shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner on stdout when the first communicator comes up; stdout must
        # carry exactly one JSON line, so point fd 1 at stderr until the communicator exists
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    if args.gpus != world and rank == 0:
        print(f"[bench] note: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)

    S, nd = args.box, args.ndiv
    z, Om = 0.5, 0.3
    cfg = nb.SubboxConfig(size=(S, S, S), ndiv=(nd, nd, nd))
    model = nb.StyleNBodyEmulatorVelCore()
    model.precision = args.precision
    params = nb.init_params(42)
    proc = nb.SubboxProcessor(model, params, cfg)
    n_sub = int(cfg.n_subboxes)
    lo, hi = shard_range(n_sub, rank, world)
    particles_total = int(np.prod(cfg.crop_size)) * n_sub

    # synthetic Gaussian box in pinned host memory
    host_t = torch.empty((3, S, S, S), dtype=torch.float32, pin_memory=True)
    host = host_t.numpy()
    np.random.default_rng(1234).standard_normal(host.shape, dtype=np.float32, out=host)

    eng = Engine.get(local)
    eng.set_precision(args.precision)
    eng.set_params(params, False, True, model.eps)
    Dz = np.float32(nb.growth_factor(z, Om))
    vf = np.float32(nb.vel_norm(z, Om))
    eng.modulate(np.float32(Om), Dz)
    crop_idx, add0, plen = cfg.flat_tables()
    box_dev = host_t.cuda()
    disp_dev = torch.zeros((3, S, S, S), dtype=torch.float32, device="cuda")
    vel_dev = torch.zeros((3, S, S, S), dtype=torch.float32, device="cuda")

    def step_dev():
        eng.process_box_dev(box_dev, cfg.size, cfg.crop_size, plen, crop_idx, add0, lo, hi - lo, Dz, vf,
                            disp_dev, vel_dev)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_dev()
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    eng.set_profiling(True)
    eng.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_dev()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    launches = eng.launch_count(reset=True)
    prof = eng.get_profile()
    eng.set_profiling(False)
    fold_on = int(eng.lib.nbe_fold_active(eng.h))
    clk = clocks.stop()
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    ms_per_step = ms_total / args.steps
    value = particles_total / (ms_per_step * 1e-3)

    # roofline of the dominant kernel (largest share of the step), live CUDA-event timings
    pk, pk_src = peaks()
    conv = [(n, t, f) for (n, t, f) in prof if f > 0 and t > 0]
    roof = None
    if conv:
        dom = max(conv, key=lambda r: r[1])
        ach = dom[2] / (dom[1] * 1e-3) / 1e12
        peak = float(pk.get("bf16_tflops_sustained", pk.get("bf16_tflops")))
        tot_t = sum(t for _, t, _ in prof)
        tot_f = sum(f for _, _, f in prof)
        roof = {"bound": "tensor", "kernel": dom[0], "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak, "traffic": None, "peak_source": pk_src + " (bf16 sustained; fp16 operands run at the same rate)",
                "kernel_share_of_step": dom[1] / tot_t if tot_t else None,
                "net_achieved": tot_f / (tot_t * 1e-3) / 1e12 if tot_t else None,
                "net_frac": tot_f / (tot_t * 1e-3) / 1e12 / peak if tot_t else None,
                "note": "algorithmic FLOPs (SURVEY 8d: primal + two tangent products per layer) / mean launch duration; split precision with the folded tangent (DESIGN 4.2) executes 4/3 of them, 5/3 without"}
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                with open(tp) as f:
                    roof["traffic"] = json.load(f).get(dom[0])
            except Exception:
                pass
    if args.kernels and rank == 0:
        for n, t, f in prof:
            print(f"[kernel] {n:26s} {t:8.3f} ms  {f / (t * 1e-3) / 1e12 if t > 0 else 0:8.1f} TFLOP/s", file=sys.stderr)

    # supplementary: the single-product fp16 mode (misses the velocity gate; reported for context)
    alt = None
    if args.precision == "split" and not args.no_alt:
        eng.set_precision("fp16")
        eng.modulate(np.float32(Om), Dz)
        step_dev(); barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(); step_dev(); a1.record(); barrier()
        t_alt = torch.tensor([a0.elapsed_time(a1)], device="cuda")
        if dist is not None:
            dist.all_reduce(t_alt, op=dist.ReduceOp.MAX)
        alt = {"precision": "fp16 single product", "value": particles_total / (float(t_alt.item()) * 1e-3),
               "unit": "particles/s", "note": "rel-L2 disp 5e-4, vel ~1e-2: does not meet the 1e-3 velocity gate"}
        eng.set_precision(args.precision)
        eng.modulate(np.float32(Om), Dz)

    # supplementary: halo-amortising tiling (bit-identical output, fewer halo FLOPs), SURVEY 8(f1)
    amort = None
    if (S, nd) == (512, 4) and not args.no_alt:
        def timed_merge(merge):
            mcfg, (mc, ma, mp) = proc.merged_config(merge)
            mlo, mhi = shard_range(int(mcfg.n_subboxes), rank, world)
            def step_m():
                eng.process_box_dev(box_dev, mcfg.size, mcfg.crop_size, mp, mc, ma, mlo, mhi - mlo, Dz, vf, disp_dev, vel_dev)
            step_m(); barrier()
            m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            m0.record(); step_m(); m1.record(); barrier()
            t_m = torch.tensor([m0.elapsed_time(m1)], device="cuda")
            if dist is not None:
                dist.all_reduce(t_m, op=dist.ReduceOp.MAX)
            return int(mcfg.n_subboxes), particles_total / (float(t_m.item()) * 1e-3)
        try:
            n221, v221 = timed_merge((2, 2, 1))
            amort = {"merge": [2, 2, 1], "subboxes": n221, "value": v221, "unit": "particles/s",
                     "note": "same box and output bits; 16 tiles 352x352x224 -> 256x256x128 instead of 64 x 224^3 -> 128^3"}
            try:                        # 8 tiles 352^3 -> 256^3 (about 120 GB of HBM with the other arenas of this run)
                n222, v222 = timed_merge((2, 2, 2))
                amort = {"merge": [2, 2, 2], "subboxes": n222, "value": v222, "unit": "particles/s",
                         "note": "same box and output bits; 8 tiles 352^3 -> 256^3 instead of 64 x 224^3 -> 128^3",
                         "merge_221": {"subboxes": n221, "value": v221}}
            except Exception as e:
                amort["merge_222_error"] = str(e)[:200]
        except Exception as e:          # e.g. not enough HBM for the larger arena
            amort = {"error": str(e)[:200]}

    # end-to-end through the DEFAULT public call: host numpy in, whole box in one pair of host arrays out.
    # N > 1: rank 0 alone makes the call and drives all N GPUs (one host thread per GPU inside the
    # library); the other ranks release what they can and wait on a HOST barrier (an NCCL barrier would
    # park a spinning kernel on the very GPUs rank 0 is using).
    e2e = None
    if not args.no_e2e:
        hostbar = None
        if dist is not None:
            hostbar = dist.new_group(backend="gloo")
            del box_dev, disp_dev, vel_dev
            torch.cuda.empty_cache()
            if rank != 0:
                eng.release_workspace()       # rank 0 is about to run its own context on this GPU
            dist.barrier(group=hostbar)
        if rank == 0:
            devs = list(range(world))
            d_, v_ = proc.process_box(host, z, Om, show_progress=False, shard=(0, 1), devices=devs)   # warm-up (pins, contexts, arenas)
            assert d_.shape == (3, S, S, S) and v_.shape == (3, S, S, S)
            chk = float(d_[:, ::61, ::67, ::71].astype(np.float64).sum())
            del d_, v_
            n_e2e = max(1, args.steps)
            for g in devs:
                torch.cuda.synchronize(g)
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                d_, v_ = proc.process_box(host, z, Om, show_progress=False, shard=(0, 1), devices=devs)
                del d_, v_
            t_e2e = (time.perf_counter() - t0) / n_e2e
            import resource
            e2e = {"value": particles_total / t_e2e, "unit": "particles/s",
                   "h2d_bytes_per_step": int(e2e_h2d_bytes(cfg, crop_idx, plen, n_sub, world)),
                   "d2h_bytes_per_step": int(2 * 3 * particles_total * 4), "ms_per_step": t_e2e * 1e3,
                   "api": "SubboxProcessor.process_box(host numpy, z, Om) [defaults] -> nbe_process_box_multi, "
                          f"{world} GPU(s) driven by one process; returns the assembled (3,S,S,S) disp + vel host arrays",
                   "host_rss_gb": resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e6, "checksum": chk}
        if dist is not None:
            dist.barrier(group=hostbar)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        t = cpu_oracle_time(128, cores)
        sub224 = t * FLOP_PER_SUBBOX_224 / FLOP_PER_SUBBOX_128
        cpu = {"value": 128 ** 3 / sub224, "unit": "particles/s", "cores": cores, "kind": "port",
               "sample": "one 128^3->32^3 Style+vel subbox, fp32 torch-CPU oracle (%.1f s); scaled to the 224^3 "
                         "subbox of this workload by algorithmic FLOPs (x8.31)" % t}

    if rank == 0:
        line = {
            "metric": "output particles/s, 512^3 Style+vel box (ndiv 4x4x4)" if (S, nd) == (512, 4)
                      else f"output particles/s, {S}^3 Style+vel box (ndiv {nd})",
            "value": value, "unit": "particles/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": value / A100_README_PARTICLES_PER_S if (S, nd) == (512, 4) else None,
            "dtype": "f16x2-split (fp16 hi+lo operands, fp32 accumulate)" if args.precision == "split" else "f16",
            "data": "synthetic",
            "config": {"workload": f"create_emulator(compute_vel=True) {S}^3 box, ndiv=({nd},{nd},{nd}), "
                                   f"{n_sub} subboxes {plen[0]}^3->{cfg.crop_size[0]}^3, Style+vel",
                       "weights": "fixed-seed random init (pretrained blob absent from the reference checkout)",
                       "sharding": f"{world} rank(s), contiguous subbox ranges, no data-path collective",
                       "l2": "inputs larger than L2 (1.6 GB box, >10 GB activations per subbox)",
                       "precision": args.precision,
                       "tangent": "folded: (dx + a.x)*W + beta.(x*W), 4 tensor-core products per layer" if fold_on
                                  else "x*dW + dx*W, 5 tensor-core products per layer",
                       "baseline_note": "vs_baseline = value / (512^3 / 44.9 s), the reference README's A100-40GB fp32 figure"},
            "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "alt_precision": alt, "halo_amortised": amort,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
