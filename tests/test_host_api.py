"""CPU tests of the host side: C-ABI library loads and exports every symbol include/nbe.h
declares (no compute calls), cosmology helpers vs the oracle, parameter-tree validation,
drop-in error behaviour, loud failure without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import jax_nbody_emulator_with_dj_b200 as nb
from jax_nbody_emulator_with_dj_b200 import _lib
from jax_nbody_emulator_with_dj_b200._engine import LAYERS, flatten_params, dtype_code
from oracle import cosmology as oc
from oracle.net import init_params as oracle_init, layer_table

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAS_GPU = torch.cuda.is_available()


def test_library_builds_and_exports_header_symbols():
    path = _lib.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    hdr = open(os.path.join(ROOT, "include", "nbe.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(nbe_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/nbe.h but not exported"
    assert set(names) == set(_lib.EXPORTS)
    lib.nbe_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.nbe_version()


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    out = os.popen(f"cuobjdump -sass {_lib.LIB_PATH} 2>/dev/null").read()
    if not out:
        pytest.skip("cuobjdump unavailable")
    assert "UTCHMMA" in out and "UTMALDG" in out and "LDTM" in out
    assert "HMMA.16" not in out          # no legacy mma.sync path


@pytest.mark.skipif(HAS_GPU, reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu():
    x = np.zeros((1, 3, 104, 104, 104), np.float32)
    with pytest.raises(nb.NBEError, match="no CPU fallback"):
        nb.StyleNBodyEmulatorVelCore().apply(nb.init_params(), x, 0.3, 0.8, 50.0)
    cfg = nb.SubboxConfig(size=(8, 8, 8), ndiv=(1, 1, 1))
    emu = nb.create_emulator(load_params=False, processor_config=cfg)
    emu.processor.params = nb.init_params()
    with pytest.raises(nb.NBEError):
        emu.process_box(np.zeros((3, 8, 8, 8), np.float32), 0.5, 0.3, show_progress=False)


def test_cosmology_matches_oracle_and_readme():
    zs = np.array([0.0, 0.5, 1.0, 2.0, 3.0])
    for Om in (0.1, 0.3, 0.5, 0.9):
        for f in ("growth_factor", "hubble_rate", "growth_rate", "dlogH_dloga", "vel_norm", "acc_norm"):
            a = getattr(nb, f)(zs, Om)
            b = getattr(oc, f)(zs, np.full_like(zs, Om))
            assert a.dtype == np.float32 and a.shape == zs.shape
            assert np.allclose(a, b, rtol=2e-6), (f, Om)
    assert abs(float(nb.growth_factor(0.0, 0.3)) - 1.0) < 1e-6          # tests/test_cosmology.py:18-38
    assert abs(float(nb.hubble_rate(0.0, 0.3)) - 100.0) < 1e-4
    assert abs(float(nb.growth_factor(0.5, 0.3)) - 0.77318) < 1e-5      # README.md:178-180
    assert abs(float(nb.vel_norm(0.5, 0.3)) - 50.538) < 1e-3
    assert np.ndim(nb.growth_factor(0.5, 0.3)) == 0                      # shape preserved for scalars


def test_layer_table_and_init_match_oracle():
    assert [tuple(r) for r in LAYERS] == [tuple(r) for r in layer_table()]
    a, b = oracle_init(42), nb.init_params(42)
    for blk in a["params"]:
        for lay in a["params"][blk]:
            for k, v in a["params"][blk][lay].items():
                assert np.array_equal(v, b["params"][blk][lay][k])
    pm = nb.init_params(1, premodulated=True, compute_vel=True)
    assert set(pm["params"]["conv_c"]["conv_0"]) == {"weight", "bias", "dweight"}
    assert nb.NBodyEmulatorCore().init(7)["params"]["conv_c"]["skip"].keys() == {"weight", "bias"}


def test_flatten_params_validation():
    p = nb.init_params(3)
    arr, keep = flatten_params(p, False, True)
    assert len(arr) == 33 and arr[0].block == b"conv_l00" and arr[0].layer == b"skip" and arr[0].cin == 3
    with pytest.raises(ValueError, match="style model needs"):
        flatten_params(nb.init_params(3, premodulated=True), False, True)
    with pytest.raises(ValueError, match="dweight"):
        flatten_params(nb.init_params(3, premodulated=True, compute_vel=False), True, True)
    with pytest.raises(ValueError, match="un-modulated"):
        flatten_params(p, True, True)
    bad = nb.init_params(3)
    bad["params"]["conv_c"]["conv_0"]["weight"] = np.zeros((32, 64, 3, 3, 3), np.float32)
    with pytest.raises(ValueError, match="expected"):
        flatten_params(bad, False, True)
    bad = nb.init_params(3)
    del bad["params"]["up_r1"]
    with pytest.raises(ValueError, match="no layer"):
        flatten_params(bad, False, True)
    with pytest.raises(ValueError):
        flatten_params(None, False, True)


def test_create_emulator_contract():        # nbody_emulator.py:268-384, tests/test_nbody_emulator.py
    emu = nb.create_emulator(load_params=False)
    assert isinstance(emu.model, nb.StyleNBodyEmulatorVelCore) and emu.params is None and emu.processor is None
    assert emu.premodulate is False and emu.compute_vel is True and emu.dtype == np.float32
    assert isinstance(nb.create_emulator(premodulate=True, load_params=False).model, nb.NBodyEmulatorVelCore)
    assert isinstance(nb.create_emulator(premodulate=True, compute_vel=False, load_params=False).model, nb.NBodyEmulatorCore)
    assert isinstance(nb.create_emulator(compute_vel=False, load_params=False).model, nb.StyleNBodyEmulatorCore)
    m = emu.model
    assert (m.style_size, m.in_chan, m.out_chan, m.mid_chan, m.eps) == (2, 3, 3, 64, 1e-8)
    with pytest.raises(ValueError, match="No parameters loaded"):
        emu.apply(np.zeros((1, 3, 128, 128, 128), np.float32), 0.0, 0.3)
    with pytest.raises(ValueError, match="No processor created"):
        emu.process_box(np.zeros((3, 64, 64, 64), np.float32), 0.0, 0.3)
    # processor_config.dtype overrides the dtype kwarg (tests/test_nbody_emulator.py:463-475)
    cfg = nb.SubboxConfig(size=(64, 64, 64), ndiv=(2, 2, 2), dtype=np.float16)
    e2 = nb.create_emulator(load_params=False, processor_config=cfg, dtype=np.float32)
    assert e2.dtype == np.float16 and isinstance(e2.processor, nb.SubboxProcessor)
    assert e2.processor.premodulate is False and e2.processor.compute_vel is True
    assert nb.create_emulator(load_params=False, dtype=np.float16).dtype == np.float16
    # premodulate=True with load_params=False does not raise
    nb.create_emulator(premodulate=True, load_params=False)
    # the pretrained blob is absent from the reference checkout: loading must fail, not fall back
    with pytest.raises((FileNotFoundError, OSError)):
        nb.create_emulator(load_params=True)
    with pytest.raises((FileNotFoundError, OSError, ValueError)):
        nb.create_emulator(premodulate=True, load_params=True)


def test_model_rejects_unsupported_shapes():
    with pytest.raises(NotImplementedError):
        nb.StyleNBodyEmulatorVelCore(mid_chan=32)._check()
    from jax_nbody_emulator_with_dj_b200.models import _prep_x
    if not HAS_GPU:
        return
    with pytest.raises(ValueError, match="multiples of 8"):
        _prep_x(np.zeros((1, 3, 100, 104, 104), np.float32))
    with pytest.raises(ValueError, match=r"\(B, 3, D, H, W\)"):
        _prep_x(np.zeros((3, 104, 104, 104), np.float32))


def test_dtype_codes():
    assert dtype_code(np.float32) == 0 and dtype_code(np.float16) == 1 and dtype_code("bfloat16") == 2
    assert dtype_code(torch.float16) == 1 and dtype_code(torch.bfloat16) == 2
    with pytest.raises(ValueError):
        dtype_code(np.float64)


def test_public_names():                    # __init__.py:73-95
    for n in ["create_emulator", "NBodyEmulator", "SubboxConfig", "SubboxProcessor", "load_default_parameters",
              "modulate_emulator_parameters", "modulate_emulator_parameters_vel", "growth_factor", "hubble_rate",
              "growth_rate", "dlogH_dloga", "vel_norm", "acc_norm", "StyleNBodyEmulatorCore",
              "StyleNBodyEmulatorVelCore", "NBodyEmulatorCore", "NBodyEmulatorVelCore"]:
        assert hasattr(nb, n) and n in nb.__all__
