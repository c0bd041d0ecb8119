"""CLI contract of examples/run_b200_emulator.py (same flags / files as the reference's
examples/run_jax_emulator.py:186-355).  Argument handling on CPU, one end-to-end run on GPU."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("run_b200_emulator", os.path.join(ROOT, "examples", "run_b200_emulator.py"))
cli = importlib.util.module_from_spec(spec)
spec.loader.exec_module(cli)


def _sim(tmp_path, size=(8, 8, 16), Om=0.3, z=0.5, seed=31):
    d = tmp_path / "sim0"
    d.mkdir()
    np.save(d / "params.npy", np.array([Om, 0.049, 0.67, 0.96, 0.8, z]))
    np.save(d / "dis.npy", np.random.default_rng(seed).standard_normal((3,) + size, dtype=np.float32))
    return d


def test_argument_parsing(tmp_path):
    d = _sim(tmp_path)
    ap = cli.parser()
    a = ap.parse_args(["--cosmo_param_files", str(d / "params.npy"), "--displacement_files", str(d / "dis.npy"),
                       "--output_dirs", str(d), "--ndiv", "1,1,2", "--no-style", "--precision", "f16", "-q"])
    assert a.ndiv == (1, 1, 2) and a.vel is True and a.style is False and a.precision == np.float16
    assert a.output_precision == np.float16 and a.quiet
    assert cli.triple("4") == (4, 4, 4) and cli.triple("(2, 4, 4)") == (2, 4, 4)
    with pytest.raises(SystemExit):
        ap.parse_args(["--cosmo_param_files", str(d / "nope*.npy"), "--displacement_files", str(d / "dis.npy"),
                       "--output_dirs", str(d), "--ndiv", "2"])
    assert cli.read_cosmology(d / "params.npy") == (0.3, 0.5)
    np.save(d / "bad.npy", np.array([0.9, 0, 0, 0, 0, 0.5]))
    with pytest.raises(SystemExit, match="out of valid range"):
        cli.read_cosmology(d / "bad.npy")
    np.save(d / "bad4.npy", np.zeros((2, 4, 4, 4), np.float32))
    with pytest.raises(SystemExit, match="is not 3"):
        cli.box_shape([d / "bad4.npy"])
    assert cli.box_shape([d / "dis.npy"]) == (3, 8, 8, 16)


@pytest.mark.gpu
def test_end_to_end_files(tmp_path, golden_dir):
    from oracle.net import rel_l2
    g = np.load(os.path.join(golden_dir, "box.npz"))
    d = _sim(tmp_path, seed=int(g["seed"]))
    cli.main(["--cosmo_param_files", str(d / "params.npy"), "--displacement_files", str(d / "dis.npy"),
              "--output_dirs", str(d), "--ndiv", "1,1,2", "--random-weights", "42", "--output-precision", "f32", "-q"])
    dis, vel = np.load(d / "emu_dis.npy"), np.load(d / "emu_vel.npy")
    assert dis.dtype == np.float32 and dis.shape == (3, 8, 8, 16)
    assert rel_l2(dis, g["disp"]) <= 1e-3 and rel_l2(vel, g["vel"]) <= 1e-3
    cli.main(["--cosmo_param_files", str(d / "params.npy"), "--displacement_files", str(d / "dis.npy"),
              "--output_dirs", str(d), "--ndiv", "1,1,2", "--random-weights", "42", "--no-vel", "--no-style", "-q"])
    dis16 = np.load(d / "emu_dis.npy")
    assert dis16.dtype == np.float16 and rel_l2(dis16, g["disp"]) <= 2e-3


@pytest.mark.gpu
def test_density_and_power_spectrum_outputs(tmp_path):
    """--boxsize: emu_delta.npy / emu_pk.txt, the step the reference driver runs after the emulator
    (scripts/core.py:446-458), checked against the numpy oracle on the CLI's own displacement."""
    from oracle import density as D
    d = _sim(tmp_path, size=(16, 16, 16), seed=5)
    cli.main(["--cosmo_param_files", str(d / "params.npy"), "--displacement_files", str(d / "dis.npy"),
              "--output_dirs", str(d), "--ndiv", "1,1,1", "--random-weights", "42", "--output-precision", "f32",
              "--no-vel", "--boxsize", "16", "--mas-worder", "3", "-q"])
    dis, delta, pk = np.load(d / "emu_dis.npy"), np.load(d / "emu_delta.npy"), np.loadtxt(d / "emu_pk.txt")
    ref = D.delta_from_psi(dis.astype(np.float64), 16.0, worder=3)
    assert np.abs(delta - ref).max() / np.sqrt(np.mean((1 + ref) ** 2)) < 2e-5
    k, P, N = D.power_spectrum(ref, 16.0, MAS="TSC")
    assert np.array_equal(pk[:, 2], N) and np.allclose(pk[:, 1], P, rtol=1e-3) and np.allclose(pk[:, 0], k, rtol=1e-6)
    with pytest.raises(SystemExit):
        d2 = _sim(tmp_path / "x", size=(8, 8, 16)) if (tmp_path / "x").mkdir() is None else None
        cli.main(["--cosmo_param_files", str(d2 / "params.npy"), "--displacement_files", str(d2 / "dis.npy"),
                  "--output_dirs", str(d2), "--ndiv", "1,1,2", "--random-weights", "42", "--boxsize", "8", "-q"])
