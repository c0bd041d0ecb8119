"""GPU tests (-m gpu) of round 2's additions: per-layer activation parity (every stored activation of
the net against the oracle's intermediates, so a ResNet / resample / transpose layer fails on its own
and not only through the net output), the host path (windowed upload, stream ordering, table
validation, output ownership) and the one-call multi-GPU entry points (self-skip below 2 GPUs)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import jax_nbody_emulator_with_dj_b200 as nb
from jax_nbody_emulator_with_dj_b200._engine import Engine
from oracle.net import Net, rel_l2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = nb.init_params(42)

# internal activation id (csrc/nbe_api.cu: ActId) -> name of the oracle intermediate (oracle/net.py: Net.cap)
ACTS = ['in', 'conv_l00.conv_0', 'conv_l00', 'conv_l01.conv_0', 'conv_l01', 'down_l0', 'conv_l1.conv_0', 'conv_l1',
        'down_l1', 'conv_l2.conv_0', 'conv_l2', 'down_l2', 'conv_c.conv_0', 'conv_c', 'up_r2', 'conv_r2.conv_0',
        'conv_r2', 'up_r1', 'conv_r1.conv_0', 'conv_r1', 'up_r0', 'conv_r00.conv_0', 'conv_r00', 'conv_r01.conv_0']


def field(shape, seed):
    return np.random.default_rng(seed).standard_normal(shape, dtype=np.float32)


def _read_act(eng, aid, which):
    shp = (C.c_int32 * 4)()
    eng.lib.nbe_debug_read_act(eng.h, aid, 0, None, 0, shp)
    buf = np.empty(tuple(shp), dtype=np.float16)
    r = eng.lib.nbe_debug_read_act(eng.h, aid, which, C.c_void_p(buf.ctypes.data), buf.nbytes, shp)
    assert r == buf.nbytes, r
    return buf.astype(np.float32)


def test_every_stored_activation_matches_the_oracle():
    """N = 104 Style+vel: primal (hi + lo) to 1e-5 and Dz-tangent to 2e-3 per activation tensor.  The
    tangent bound is the fp16 storage of dx (2^-12 per element) plus what the layers before it
    accumulated; a wrong tap, crop, parity or skip term is an O(1) error in exactly one row below."""
    x = field((1, 3, 104, 104, 104), 1234)
    z, Om = 0.5, 0.3
    Dz, vf = float(nb.growth_factor(z, Om)), float(nb.vel_norm(z, Om))
    d, v = nb.StyleNBodyEmulatorVelCore().apply(P, x, Om, Dz, vf)
    net = Net(True, True, torch.float32)
    net.cap = {}
    with torch.no_grad():
        rd, rv = net.forward(P, x, Om, Dz, vf)
    eng = Engine.get()
    worst = {}
    folded = 0
    for aid, name in enumerate(ACTS):
        rx, rdx = net.cap[name]
        rx = rx[0].permute(1, 2, 3, 0).numpy()
        hi = _read_act(eng, aid, 0)
        if name == 'in':                       # 16-channel record [xh | xl | xh | 0]
            assert rel_l2(hi[..., 0:3] + hi[..., 3:6], rx) < 1e-6, name
            assert np.array_equal(hi[..., 0:3], hi[..., 6:9]) and not hi[..., 9:].any()
            continue
        assert hi.shape == rx.shape, (name, hi.shape, rx.shape)
        ex = rel_l2(hi + _read_act(eng, aid, 1), rx)
        # tangent folding: a tensor read by a FOLD launch stores dx' = dx + a (.) x (a of that launch's 3^3 conv)
        a = np.zeros(rx.shape[-1], dtype=np.float32)
        assert eng.lib.nbe_debug_act_fold(eng.h, aid, 0, a.ctypes.data_as(C.POINTER(C.c_float)), a.size) == a.size
        folded += bool(a.any())
        edx = rel_l2(_read_act(eng, aid, 2), rdx[0].permute(1, 2, 3, 0).numpy() + a * rx)
        worst[name] = (ex, edx)
        assert ex < 1e-5 and edx < 2e-3, (name, ex, edx)
    assert len(worst) == 23
    assert folded == (20 if eng.lib.nbe_fold_active(eng.h) else 0)
    assert rel_l2(d, rd.numpy()) < 1e-5 and rel_l2(v, rv.numpy()) < 1e-3


def test_modulation_on_a_busy_stream_is_ordered_before_process_box():
    """nbe_modulate runs on the caller's (torch) stream, nbe_process_box on the context's own
    non-blocking streams: an event orders them.  Queue ~100 ms of work on the torch stream, change Om
    (so the weights are re-modulated behind it) and call process_box at once."""
    size, ndiv = (8, 8, 16), (1, 1, 2)
    box = field((3,) + size, 31)
    proc = nb.SubboxProcessor(nb.StyleNBodyEmulatorVelCore(), P, nb.SubboxConfig(size=size, ndiv=ndiv))
    ref = {}
    for Om in (0.25, 0.35):
        torch.cuda.synchronize()
        ref[Om] = [a.copy() for a in proc.process_box(box, 0.5, Om, show_progress=False)]
        torch.cuda.synchronize()
    a = torch.randn(8192, 8192, device="cuda")
    for Om in (0.25, 0.35, 0.25):
        for _ in range(40):
            a = torch.tanh(a @ a) * 0.01           # keeps the torch stream busy while the modulation is queued
        d, v = proc.process_box(box, 0.5, Om, show_progress=False)
        assert np.array_equal(d, ref[Om][0]) and np.array_equal(v, ref[Om][1])
    torch.cuda.synchronize()


def test_c_abi_rejects_out_of_range_tables():
    size, ndiv = (8, 8, 16), (1, 1, 2)
    cfg = nb.SubboxConfig(size=size, ndiv=ndiv)
    proc = nb.SubboxProcessor(nb.StyleNBodyEmulatorVelCore(), P, cfg)
    box = field((3,) + size, 1)
    d, v = proc.process_box(box, 0.5, 0.3, show_progress=False)        # known-good state
    eng = Engine.get()
    crop_idx, add0, plen = cfg.flat_tables()
    bad = crop_idx.copy(); bad[5] = 8                                    # D index == size
    with pytest.raises(ValueError, match="crop_idx"):
        eng.process_box(box, 0, size, cfg.crop_size, plen, bad, add0, 0, 2, 0.77, 50.0, d, v, 0)
    bad = crop_idx.copy(); bad[-1] = -1
    with pytest.raises(ValueError, match="crop_idx"):
        eng.process_box(box, 0, size, cfg.crop_size, plen, bad, add0, 0, 2, 0.77, 50.0, d, v, 0)
    bada = add0.copy(); bada[5] = 9                                      # W anchor 9 + crop 8 > 16
    with pytest.raises(ValueError, match="add_idx0"):
        eng.process_box(box, 0, size, cfg.crop_size, plen, crop_idx, bada, 0, 2, 0.77, 50.0, d, v, 0)
    d2, v2 = proc.process_box(box, 0.5, 0.3, show_progress=False)      # the context is still usable
    assert np.array_equal(d2, d) and np.array_equal(v2, v)


def test_returned_boxes_belong_to_the_caller():
    """Zero-copy outputs: the arrays are the pinned memory the GPU wrote, and a later call never
    rewrites an array that is still referenced; once dropped, the buffer is recycled."""
    size, ndiv = (8, 8, 16), (1, 1, 2)
    proc = nb.SubboxProcessor(nb.StyleNBodyEmulatorVelCore(), P, nb.SubboxConfig(size=size, ndiv=ndiv))
    b1, b2 = field((3,) + size, 1), field((3,) + size, 2)
    d1, v1 = proc.process_box(b1, 0.5, 0.3, show_progress=False)
    keep = d1.copy()
    d2, v2 = proc.process_box(b2, 0.5, 0.3, show_progress=False)
    assert not np.shares_memory(d1, d2) and np.array_equal(d1, keep) and not np.array_equal(d1, d2)
    addr = d2.ctypes.data
    del d2, v2
    d3, v3 = proc.process_box(b1, 0.5, 0.3, show_progress=False)
    assert d3.ctypes.data == addr and np.array_equal(d3, keep)          # recycled buffer, same answer
    dc, vc = proc.process_box(b1, 0.5, 0.3, show_progress=False, copy=True)
    assert np.array_equal(dc, keep) and dc.flags["OWNDATA"]


def test_windowed_upload_four_shards_union_equals_unsharded():
    """224^3 subboxes on a 256^3 box cut into 4 shards of 2 subboxes: each shard's device window is
    224 of 256 D-planes x 224 of 256 H-rows, wrapped around the periodic boundary (two D-runs, two
    H-runs = four rectangles per channel), uploaded incrementally."""
    size, ndiv = (256, 256, 256), (2, 2, 2)
    box = field((3,) + size, 99)
    proc = nb.SubboxProcessor(nb.StyleNBodyEmulatorVelCore(), P, nb.SubboxConfig(size=size, ndiv=ndiv))
    d, v = [a.copy() for a in proc.process_box(box, 0.5, 0.3, show_progress=False)]
    ds, vs = np.zeros_like(d), np.zeros_like(v)
    for r in range(4):
        dr, vr = proc.process_box(box, 0.5, 0.3, show_progress=False, shard=(r, 4), gather="none")
        ds += dr
        vs += vr
    assert np.array_equal(ds, d) and np.array_equal(vs, v)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_one_call_multi_gpu_is_bit_identical_to_one_gpu():
    """SubboxProcessor.process_box drives every visible GPU by default (nbe_process_box_multi: one host
    thread per GPU, one pinned input, one output)."""
    size, ndiv = (256, 256, 256), (2, 2, 2)
    box = field((3,) + size, 5)
    proc = nb.SubboxProcessor(nb.StyleNBodyEmulatorVelCore(), P, nb.SubboxConfig(size=size, ndiv=ndiv))
    d1, v1 = proc.process_box(box, 0.5, 0.3, show_progress=False, devices=[0])
    cur = torch.cuda.current_device()
    dn, vn = proc.process_box(box, 0.5, 0.3, show_progress=False)                 # default: all GPUs
    assert torch.cuda.current_device() == cur                                    # the caller's device is restored
    assert np.array_equal(dn, d1) and np.array_equal(vn, v1)
    for g in range(torch.cuda.device_count()):
        assert Engine.get(g).launch_count() > 0
    d2, v2 = proc.process_box(box, 0.5, 0.3, show_progress=False, devices=[1, 0])
    assert np.array_equal(d2, d1) and np.array_equal(v2, v1)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_nccl_block_gather_two_ranks(tmp_path):
    """One process per GPU under torchrun: gather='all' all-gathers the (subbox, 3, c, c, c) records on
    the devices; every rank must end with the single-process result, bit for bit."""
    script = os.path.join(ROOT, "tools", "dist_gather_check.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29631", script, str(tmp_path / "out.json")],
                       capture_output=True, text=True, env=env, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    import json
    res = json.load(open(tmp_path / "out.json"))
    assert res["bit_identical_all_ranks"] is True and res["gather"]["GBps"] > 1.0


def test_streamed_slabs_are_bit_identical_to_process_box():
    """BASELINE config 5 plumbing at test size: the 256^3 box walked as two D-slabs through
    page-locked staging buffers (plane source / slab sink callbacks) gives exactly process_box's bits;
    only (crop+96) input planes and crop output planes are ever resident on the host side."""
    size, ndiv = (256, 256, 256), (2, 2, 2)
    box = field((3,) + size, 17)
    proc = nb.SubboxProcessor(nb.StyleNBodyEmulatorVelCore(), P, nb.SubboxConfig(size=size, ndiv=ndiv))
    d, v = proc.process_box(box, 0.5, 0.3, show_progress=False)
    asked, got_d, got_v = [], np.zeros_like(d), np.zeros_like(v)

    def source(planes, out):
        asked.append(planes.copy())
        assert out.shape == (3, 224, 256, 256)
        out[...] = box[:, planes]

    def sink(k, d0, ds, vs):
        got_d[:, d0:d0 + 128] = ds
        got_v[:, d0:d0 + 128] = vs

    assert proc.process_box_streamed(source, sink, 0.5, 0.3) == 2
    assert np.array_equal(asked[0], np.arange(-48, 176) % 256) and np.array_equal(asked[1], np.arange(80, 304) % 256)
    assert np.array_equal(got_d, d) and np.array_equal(got_v, v)
