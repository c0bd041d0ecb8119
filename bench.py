#!/usr/bin/env python
"""Headline benchmark: output particles/s of the 512^3 Style+velocity box (BASELINE.json
config 3: create_emulator(compute_vel=True), ndiv=(4,4,4) = 64 subboxes of 224^3 -> 128^3),
sharded over N GPUs (one process per GPU, no data-path collective).

  python bench.py --gpus 1 --steps K --warmup W            # our arm
  python bench.py --impl reference --gpus N ...            # CPU arm (oracle port, see below)

One "step" = one pass over the rank's share of the box.  `value` is measured with the input
box and outputs resident in HBM (CUDA events on the launching stream, max over ranks); `e2e`
goes through SubboxProcessor.process_box with HOST (pinned) numpy buffers, H2D and D2H
inside the timed region.  Prints ONE JSON line on rank 0.
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
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

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
                "note": "algorithmic FLOPs (SURVEY 8d) / mean launch duration; split precision executes 5/3 of them"}
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
        try:
            mcfg, (mc, ma, mp) = proc.merged_config((2, 2, 1))
            mlo, mhi = shard_range(int(mcfg.n_subboxes), rank, world)
            def step_m():
                eng.process_box_dev(box_dev, mcfg.size, mcfg.crop_size, mp, mc, ma, mlo, mhi - mlo, Dz, vf, disp_dev, vel_dev)
            step_m(); barrier()
            m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            m0.record(); step_m(); m1.record(); barrier()
            t_m = torch.tensor([m0.elapsed_time(m1)], device="cuda")
            if dist is not None:
                dist.all_reduce(t_m, op=dist.ReduceOp.MAX)
            amort = {"merge": [2, 2, 1], "subboxes": int(mcfg.n_subboxes), "value": particles_total / (float(t_m.item()) * 1e-3),
                     "unit": "particles/s", "note": "same box and output bits; 16 tiles 352x352x224 -> 256x256x128 instead of 64 x 224^3 -> 128^3"}
        except Exception as e:          # e.g. not enough HBM for the larger arena
            amort = {"error": str(e)[:200]}

    # end-to-end through the public API: host numpy in, host numpy out
    e2e = None
    if not args.no_e2e:
        proc.process_box(host, z, Om, show_progress=False, gather="none", copy=False)     # warm-up (pins, allocs)
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, args.steps)
        for _ in range(n_e2e):
            proc.process_box(host, z, Om, show_progress=False, gather="none", copy=False)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        t_e2e = float(dt.item()) / n_e2e
        own = (hi - lo) * int(np.prod(cfg.crop_size))
        per = sum(plen)
        tabs = crop_idx.reshape(n_sub, per)[lo:hi, :plen[0]]
        planes = int(np.unique(tabs).size)          # D-planes this rank uploads (slab + halo)
        e2e = {"value": particles_total / t_e2e, "unit": "particles/s", "h2d_bytes_per_step": int(3 * planes * S * S * 4),
               "d2h_bytes_per_step": int(2 * 3 * own * 4), "ms_per_step": t_e2e * 1e3,
               "api": "SubboxProcessor.process_box(host numpy) -> nbe_process_box"}

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
                       "baseline_note": "vs_baseline = value / (512^3 / 44.9 s), the reference README's A100-40GB fp32 figure"},
            "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "alt_precision": alt, "halo_amortised": amort,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
