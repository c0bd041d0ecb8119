#!/usr/bin/env python
"""Batch CLI with the on-disk contract of the reference's examples/run_jax_emulator.py
(:186-355): glob patterns of cosmology files ([Om, Ob, h, ns, s8, z]), displacement files
((3, N, N, N) .npy) and output directories; writes emu_dis.npy (and emu_vel.npy).

    python examples/run_b200_emulator.py --cosmo_param_files 'sims/*/params.npy' \\
        --displacement_files 'sims/*/dis.npy' --output_dirs 'sims/*/' --ndiv 4 --vel

Weights: --params FILE (.npz holding the pickled {'block': {'layer': {...}}} tree, as the
reference ships) or --random-weights SEED; by default the packaged blob is loaded (absent from
this checkout, in which case the run stops with the loader's error).
"""
from __future__ import annotations

import argparse
import os
import sys
import time
from glob import glob
from pathlib import Path

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def triple(text: str):
    vals = [int(v) for v in text.strip("()").split(",")]
    if len(vals) == 1:
        return (vals[0],) * 3
    if len(vals) == 3:
        return tuple(vals)
    raise argparse.ArgumentTypeError(f"Expected 1 or 3 values, got {len(vals)}")


def precision(text: str):
    if text in ("f16", "f32"):
        return np.float16 if text == "f16" else np.float32
    raise argparse.ArgumentTypeError(f"precision must be 'f32' or 'f16', got '{text}'")


def files(pattern: str):
    found = sorted(Path(p) for p in glob(pattern))
    if not found:
        raise argparse.ArgumentTypeError(f"No files match pattern: {pattern}")
    for p in found:
        if not p.is_file() or not os.access(p, os.R_OK):
            raise argparse.ArgumentTypeError(f"Input file path is not a readable file: {p}")
    return found


def dirs(pattern: str):
    found = sorted(Path(p) for p in glob(pattern))
    if not found:
        raise argparse.ArgumentTypeError(f"No directories match pattern: {pattern}")
    for p in found:
        if not p.is_dir() or not os.access(p, os.W_OK):
            raise argparse.ArgumentTypeError(f"Output directory is not writable: {p}")
    return found


def read_cosmology(path: Path):
    data = np.load(path)
    Om, z = float(data[0]), float(data[-1])
    if not 0.1 <= Om <= 0.5:
        sys.exit(f"in file {path}: Om={Om:.4f} out of valid range [0.1, 0.5]")
    if not 0.0 <= z <= 3.0:
        sys.exit(f"in file {path}: z={z:.4f} out of valid range [0.0, 3.0]")
    return Om, z


def box_shape(paths):
    shape = None
    for p in paths:
        s = np.load(p, mmap_mode="r").shape
        if len(s) != 4:
            sys.exit(f"in file {p}: input array ndim {len(s)} is not 4")
        if s[0] != 3:
            sys.exit(f"in file {p}: first dimension {s[0]} is not 3 (expected 3 displacement components)")
        if shape is not None and s != shape:
            sys.exit(f"in file {p}: input array shape {s} differs from first file shape {shape}")
        shape = s
    return shape


def parser():
    ap = argparse.ArgumentParser(description="Batch process displacement fields with the B200 N-body emulator.")
    ap.add_argument("--cosmo_param_files", type=files, required=True)
    ap.add_argument("--displacement_files", type=files, required=True)
    ap.add_argument("--output_dirs", type=dirs, required=True)
    ap.add_argument("--ndiv", type=triple, required=True)
    ap.add_argument("--vel", action=argparse.BooleanOptionalAction, default=True)
    ap.add_argument("--style", action=argparse.BooleanOptionalAction, default=True)
    ap.add_argument("--precision", type=precision, default=np.float32)
    ap.add_argument("--output-precision", type=precision, default=np.float16, dest="output_precision")
    ap.add_argument("--quiet", "-q", action="store_true")
    ap.add_argument("--params", type=Path, default=None, help=".npz with the pickled parameter tree under 'params'")
    ap.add_argument("--random-weights", type=int, default=None, metavar="SEED")
    ap.add_argument("--boxsize", type=float, default=None,
                    help="box side in the units of the displacement: also write emu_delta.npy (density contrast of "
                         "q + emu_dis) and emu_pk.txt (k, P(k), Nmodes; MAS-corrected), the step scripts/core.py:446-458 "
                         "runs after the emulator")
    ap.add_argument("--mas-worder", type=int, default=2, choices=(2, 3, 4), help="mass assignment: 2 CIC, 3 TSC, 4 PCS")
    return ap


def load_tree(args):
    import jax_nbody_emulator_with_dj_b200 as nb
    if args.random_weights is not None:
        return nb.init_params(args.random_weights)
    if args.params is not None:
        with np.load(args.params, allow_pickle=True) as f:
            return {"params": f["params"].item()}
    return nb.load_default_parameters()


def main(argv=None):
    args = parser().parse_args(argv)
    n = len(args.cosmo_param_files)
    if not (n == len(args.displacement_files) == len(args.output_dirs)):
        sys.exit("Number of files must match:\n"
                 f"  cosmo_param_files: {n}\n  displacement_files: {len(args.displacement_files)}\n"
                 f"  output_dirs: {len(args.output_dirs)}")
    shape = box_shape(args.displacement_files)
    cosmologies = [read_cosmology(p) for p in args.cosmo_param_files]
    import jax_nbody_emulator_with_dj_b200 as nb
    cfg = nb.SubboxConfig(size=shape[1:], ndiv=args.ndiv, dtype=args.precision, output_dtype=args.output_precision)
    tree = load_tree(args)
    print(f"Processing {n} simulation(s): box {shape[1:]}, ndiv {args.ndiv}, vel {args.vel}, style {args.style}")
    emu = None
    if args.style:
        emu = nb.create_emulator(premodulate=False, compute_vel=args.vel, load_params=False, processor_config=cfg)
        emu.params = emu.processor.params = tree
    for i, (dis, (Om, z), out) in enumerate(zip(args.displacement_files, cosmologies, args.output_dirs)):
        if not args.style:
            emu = nb.create_emulator(premodulate=True, compute_vel=args.vel, load_params=False, processor_config=cfg)
            mod = nb.modulate_emulator_parameters_vel if args.vel else nb.modulate_emulator_parameters
            emu.params = emu.processor.params = mod(tree, z, Om)
        box = np.load(dis)
        t0 = time.time()
        res = emu.process_box(box, z=z, Om=Om, show_progress=not args.quiet)
        dt = time.time() - t0
        if args.vel:
            np.save(out / "emu_dis.npy", res[0])
            np.save(out / "emu_vel.npy", res[1])
        else:
            np.save(out / "emu_dis.npy", res)
        if args.boxsize is not None:
            if len(set(shape[1:])) != 1:
                sys.exit(f"--boxsize needs a cubic box, got {shape[1:]}")
            psi = np.asarray(res[0] if args.vel else res, dtype=np.float32)
            delta = nb.get_delta_from_psi(psi, args.boxsize, worder=args.mas_worder)
            pk = nb.power_spectrum(delta, args.boxsize, MAS=nb.mas_name_from_worder(args.mas_worder))
            np.save(out / "emu_delta.npy", delta)
            np.savetxt(out / "emu_pk.txt", np.column_stack(pk), header="k  P(k)  Nmodes")
        print(f"[{i + 1}/{n}] z={z:.4f}, Om={Om:.4f}: {dt:.2f}s -> {out}")
    print("\nDone!")


if __name__ == "__main__":
    main()
