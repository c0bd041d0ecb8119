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


def test_load_default_parameters_reads_the_reference_file_format(tmp_path, monkeypatch):
    """nbody_emulator.py:115-129: the weights ship as ONE pickled object array 'params' holding
    {block: {layer: {weight, bias, style_weight, style_bias}}}.  The blob is absent from the reference
    checkout, so the round trip is exercised on a synthetic file written in exactly that format."""
    tree = nb.init_params(7)["params"]
    f = tmp_path / "nbody_emulator_params.npz"
    np.savez(f, params=np.array(tree, dtype=object))
    got = nb.load_default_parameters(f)
    assert set(got) == {"params"} and list(got["params"]) == list(tree)
    for b, l, co, ci, k in LAYERS:
        for key, shape in (("weight", (co, ci, k, k, k)), ("bias", (co,)), ("style_weight", (ci, 2)), ("style_bias", (ci,))):
            assert got["params"][b][l][key].shape == shape
            assert np.array_equal(got["params"][b][l][key], tree[b][l][key])
    flatten_params(got, False, True)                      # the tree the C ABI accepts
    # NBE_PARAMS environment override, and create_emulator(load_params=True) through it
    monkeypatch.setenv("NBE_PARAMS", str(f))
    assert np.array_equal(nb.load_default_parameters()["params"]["conv_c"]["skip"]["bias"], tree["conv_c"]["skip"]["bias"])
    emu = nb.create_emulator(load_params=True)
    assert emu.params is not None and "conv_r01" in emu.params["params"]
    emu2 = nb.create_emulator(load_params=True, params_path=f, processor_config=nb.SubboxConfig(size=(8, 8, 8), ndiv=(1, 1, 1)))
    assert emu2.processor.params is emu2.params
    with pytest.raises(ValueError, match="premodulate_z and premodulate_Om"):
        nb.create_emulator(premodulate=True, load_params=True)
    monkeypatch.delenv("NBE_PARAMS")
    with pytest.raises(FileNotFoundError, match="NBE_PARAMS"):
        nb.load_default_parameters(tmp_path / "missing.npz")


def test_partially_styled_tree_is_rejected_not_passed_through():
    P = nb.init_params(3)
    del P["params"]["conv_l1"]["skip"]["style_weight"]
    with pytest.raises(ValueError, match="partially styled"):
        nb.modulate_emulator_parameters(P, 0.5, 0.3)
    # a fully premodulated tree is passed through layer by layer like the reference does (:184, :263)
    pm = nb.init_params(3, premodulated=True)
    out = nb.modulate_emulator_parameters_vel(pm, 0.5, 0.3)
    assert out["params"]["conv_l00"]["conv_0"] is pm["params"]["conv_l00"]["conv_0"]


def test_params_fingerprint_sees_replaced_and_overwritten_leaves():
    from jax_nbody_emulator_with_dj_b200._engine import params_fingerprint
    P = nb.init_params(5)
    f0 = params_fingerprint(P)
    assert f0 == params_fingerprint(P)
    P["params"]["conv_l2"]["conv_1"]["weight"] = P["params"]["conv_l2"]["conv_1"]["weight"].copy()
    f1 = params_fingerprint(P)
    assert f1 != f0                                    # new leaf object
    P["params"]["down_l1"]["conv_0"]["weight"][...] *= 2.0
    assert params_fingerprint(P) != f1                 # same object, new contents
    assert params_fingerprint({"params": {}}) is None


def test_output_pool_never_rewrites_an_array_the_caller_still_holds(monkeypatch):
    """Returned boxes are the page-locked memory the GPUs wrote into; a buffer is recycled only after
    the caller dropped the array (and every view of it)."""
    from jax_nbody_emulator_with_dj_b200 import subbox as sb
    made = []

    def fake_pinned(shape, np_dtype):
        t = torch.zeros(shape, dtype=torch.float32)
        made.append(t)
        return t, t.numpy()
    monkeypatch.setattr(sb, "_pinned_zeros", fake_pinned)
    pool = sb._OutputPool()
    key, shape = ("d", (3, 4, 4, 4), "float32"), (3, 4, 4, 4)
    a = pool.take(key, shape, np.float32, (0, 8))
    a[...] = 1.0
    b = pool.take(key, shape, np.float32, (0, 8))        # `a` is alive: a second buffer
    assert len(made) == 2 and not np.shares_memory(a, b) and np.all(a == 1.0)
    view = a[0]
    del a
    c = pool.take(key, shape, np.float32, (0, 8))        # a view keeps the first buffer alive
    assert len(made) == 3
    del view, b
    d = pool.take(key, shape, np.float32, (0, 8))        # now one of them is free: recycled, not re-zeroed
    assert len(made) == 3
    d[...] = 5.0
    del d
    e = pool.take(key, shape, np.float32, (0, 4))        # another owned range: stale voxels are cleared
    assert len(made) == 3 and np.all(e == 0)
    assert c is not None
