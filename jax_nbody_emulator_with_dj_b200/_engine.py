"""Per-GPU engine: owns one ``nbe_ctx`` (include/nbe.h) and feeds it from numpy / torch.

PyTorch is used only for device memory, streams and pinned host buffers; all compute goes
through libnbe_b200.so.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import NBEError

# (block, layer, cout, cin, k): the 33 conv layers in tree order
# (style_nbody_emulator_vel_core.py:46-103; style_blocks_vel.py:112-135).
BLOCKS = (
    ("conv_l00", "res", 3, 64), ("conv_l01", "res", 64, 64), ("down_l0", "resample", 64, 64),
    ("conv_l1", "res", 64, 64), ("down_l1", "resample", 64, 64), ("conv_l2", "res", 64, 64),
    ("down_l2", "resample", 64, 64), ("conv_c", "res", 64, 64), ("up_r2", "resample", 64, 64),
    ("conv_r2", "res", 128, 64), ("up_r1", "resample", 64, 64), ("conv_r1", "res", 128, 64),
    ("up_r0", "resample", 64, 64), ("conv_r00", "res", 128, 64), ("conv_r01", "res", 64, 3),
)


def layer_table():
    rows = []
    for name, kind, cin, cout in BLOCKS:
        if kind == "res":
            mid = max(cin, cout)
            rows += [(name, "skip", cout, cin, 1), (name, "conv_0", mid, cin, 3), (name, "conv_1", cout, mid, 3)]
        else:
            rows.append((name, "conv_0", cout, cin, 2))
    return rows


LAYERS = layer_table()


def _torch():
    import torch
    return torch


def dtype_code(dt):
    """numpy / torch / string dtype -> NBE_* code."""
    torch = _torch()
    if isinstance(dt, torch.dtype):
        m = {torch.float32: _lib.NBE_F32, torch.float16: _lib.NBE_F16, torch.bfloat16: _lib.NBE_BF16}
        if dt not in m:
            raise ValueError(f"unsupported dtype {dt}")
        return m[dt]
    name = dt if isinstance(dt, str) and dt == "bfloat16" else np.dtype(dt).name
    m = {"float32": _lib.NBE_F32, "float16": _lib.NBE_F16, "bfloat16": _lib.NBE_BF16}
    if name not in m:
        raise ValueError(f"unsupported dtype {dt}")
    return m[name]


def torch_dtype(code):
    torch = _torch()
    return {_lib.NBE_F32: torch.float32, _lib.NBE_F16: torch.float16, _lib.NBE_BF16: torch.bfloat16}[code]


def _f32(a):
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(a), dtype=np.float32)


def flatten_params(params, premod, vel):
    """Param tree -> (LayerParams array, keepalive list).  Shapes are checked against the
    reference tree (tests/test_style_nbody_emulator_vel_core.py:391-446)."""
    if params is None or "params" not in params:
        raise ValueError("params must be a dict {'params': {block: {layer: {...}}}}")
    P = params["params"]
    arr = (_lib.LayerParams * len(LAYERS))()
    keep = []
    for n, (b, l, co, ci, k) in enumerate(LAYERS):
        try:
            lp = P[b][l]
        except KeyError:
            raise ValueError(f"parameter tree has no layer {b}/{l}") from None
        w = _f32(lp["weight"])
        if w.shape != (co, ci, k, k, k):
            raise ValueError(f"{b}/{l}/weight has shape {w.shape}, expected {(co, ci, k, k, k)} "
                             "(only mid_chan=64, in_chan=out_chan=3 is supported)")
        bias = _f32(lp["bias"])
        if bias.shape != (co,):
            raise ValueError(f"{b}/{l}/bias has shape {bias.shape}, expected {(co,)}")
        keep += [w, bias]
        e = arr[n]
        e.block, e.layer = b.encode(), l.encode()
        e.weight, e.bias = w.ctypes.data, bias.ctypes.data
        e.cout, e.cin, e.k = co, ci, k
        if premod:
            if "style_weight" in lp:
                raise ValueError(f"{b}/{l}: premodulated model got un-modulated (style) parameters")
            if vel:
                if "dweight" not in lp:
                    raise ValueError(f"{b}/{l}: velocity model needs 'dweight' (use modulate_emulator_parameters_vel)")
                dw = _f32(lp["dweight"])
                if dw.shape != w.shape:
                    raise ValueError(f"{b}/{l}/dweight has shape {dw.shape}")
                keep.append(dw)
                e.dweight = dw.ctypes.data
        else:
            if "style_weight" not in lp or "style_bias" not in lp:
                raise ValueError(f"{b}/{l}: style model needs 'style_weight' and 'style_bias'")
            sw, sb = _f32(lp["style_weight"]), _f32(lp["style_bias"])
            if sw.shape != (ci, 2) or sb.shape != (ci,):
                raise ValueError(f"{b}/{l}: style_weight {sw.shape} / style_bias {sb.shape} (style_size must be 2)")
            keep += [sw, sb]
            e.style_weight, e.style_bias = sw.ctypes.data, sb.ctypes.data
    return arr, keep


def params_fingerprint(params):
    """Cheap change detector for a parameter tree: per leaf (id, address, shape, sum of a 61-strided
    sample).  O(n_leaves) Python plus ~1 % of the weights read."""
    fp = []
    try:
        P = params["params"]
        for b, l, *_ in LAYERS:
            for name, a in sorted(P[b][l].items()):
                if hasattr(a, "detach"):
                    fp.append((name, id(a), a.data_ptr(), tuple(a.shape), float(a.detach().reshape(-1)[::61].sum())))
                else:
                    arr = np.asarray(a)
                    fp.append((name, id(a), arr.__array_interface__["data"][0], arr.shape, float(arr.reshape(-1)[::61].sum())))
    except (KeyError, TypeError, AttributeError):
        return None                    # malformed tree: flatten_params raises the proper error
    return tuple(fp)


class Engine:
    """One per (process, GPU)."""

    _instances = {}

    @classmethod
    def get(cls, device=None):
        torch = _torch()
        if not torch.cuda.is_available():
            raise NBEError("no CUDA device visible: the B200 kernels are the only compute path (no CPU fallback)")
        if device is None:
            device = torch.cuda.current_device()
        device = int(device)
        if device not in cls._instances:
            cls._instances[device] = cls(device)
        return cls._instances[device]

    def __init__(self, device):
        self.lib = _lib.load()
        self.device = device
        h = C.c_void_p()
        rc = self.lib.nbe_create(C.byref(h), device)
        if rc != 0:
            raise NBEError(f"nbe_create(device={device}) failed with {rc} (needs an sm_100 GPU)")
        self.h = h
        self._params_key = None
        self._params_ref = None
        self._mod_key = None
        self.precision = _lib.NBE_PREC_SPLIT

    def _ck(self, rc):
        if rc != 0:
            msg = self.lib.nbe_last_error(self.h).decode(errors="replace")
            if rc == -1:
                raise ValueError(msg)
            raise NBEError(f"libnbe error {rc}: {msg}")

    # ---- configuration
    def set_precision(self, precision):
        if isinstance(precision, str):
            precision = {"split": _lib.NBE_PREC_SPLIT, "fp16": _lib.NBE_PREC_FP16}[precision]
        if precision != self.precision:
            self._ck(self.lib.nbe_set_precision(self.h, precision))
            self.precision = precision
            self._mod_key = None

    def set_params(self, params, premod, vel, eps=1e-8, fingerprint=None):
        """Upload the tree unless the SAME tree with the same leaves was uploaded last.  The cache key
        holds the identity, address and shape of every leaf plus a strided content sample, so replacing
        a leaf (``params['params'][b][l]['weight'] = w2``) or overwriting one in place is seen; call
        ``invalidate()`` after any other kind of in-place edit."""
        if fingerprint is None:
            fingerprint = params_fingerprint(params)
        key = (id(params), bool(premod), bool(vel), float(eps), fingerprint)
        if key == self._params_key and self._params_ref is params:
            return
        arr, keep = flatten_params(params, premod, vel)
        self._ck(self.lib.nbe_set_params(self.h, arr, len(LAYERS), int(premod), int(vel), float(eps)))
        del keep
        self._params_key, self._params_ref = key, params
        self._mod_key = None

    def invalidate(self):
        self._params_key = self._params_ref = self._mod_key = None

    def modulate(self, Om, Dz):
        torch = _torch()
        Dz = np.ascontiguousarray(np.atleast_1d(np.asarray(Dz, dtype=np.float32)))
        if Om is None:
            key = ("premod",)
            if key == self._mod_key:
                return
            one = np.ones(1, dtype=np.float32)
            self._ck(self.lib.nbe_modulate(self.h, None, one.ctypes.data_as(C.POINTER(C.c_float)), 1,
                                           C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
            self._mod_key = key
            return
        Om = np.ascontiguousarray(np.atleast_1d(np.asarray(Om, dtype=np.float32)))
        key = (Om.tobytes(), Dz.tobytes())
        if key == self._mod_key:
            return
        self._ck(self.lib.nbe_modulate(self.h, Om.ctypes.data_as(C.POINTER(C.c_float)),
                                       Dz.ctypes.data_as(C.POINTER(C.c_float)), int(Dz.shape[0]),
                                       C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        self._mod_key = key

    def get_modulated(self, layer_index, sample=0, want_dw=True):
        _, _, co, ci, k = LAYERS[layer_index]
        w = np.empty((co, ci, k, k, k), dtype=np.float32)
        dw = np.empty_like(w) if want_dw else None
        self._ck(self.lib.nbe_get_modulated(self.h, layer_index, sample, w.ctypes.data,
                                            dw.ctypes.data if want_dw else None))
        return w, dw

    # ---- compute
    def forward(self, x, Dz, vel_fac, vel):
        """x: torch CUDA tensor (B,3,n0,n1,n2) contiguous.  Returns disp[, vel] on the device."""
        torch = _torch()
        B, ch, n0, n1, n2 = x.shape
        code = dtype_code(x.dtype)
        out_shape = (B, 3, n0 - 96, n1 - 96, n2 - 96)
        disp = torch.empty(out_shape, dtype=x.dtype, device=x.device)
        velo = torch.empty(out_shape, dtype=x.dtype, device=x.device) if vel else None
        dims = (C.c_int32 * 3)(n0, n1, n2)
        Dz = np.ascontiguousarray(np.broadcast_to(np.atleast_1d(np.asarray(Dz, dtype=np.float32)), (B,)))
        fp = C.POINTER(C.c_float)
        vf = None
        if vel:
            vf = np.ascontiguousarray(np.broadcast_to(np.atleast_1d(np.asarray(vel_fac, dtype=np.float32)), (B,)))
        self._ck(self.lib.nbe_forward(self.h, C.c_void_p(x.data_ptr()), code, B, dims, Dz.ctypes.data_as(fp),
                                      vf.ctypes.data_as(fp) if vel else None, C.c_void_p(disp.data_ptr()),
                                      C.c_void_p(velo.data_ptr()) if vel else None, code,
                                      C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return (disp, velo) if vel else disp

    def process_box(self, in_host, in_code, size, crop, plen, crop_idx, add_idx0, first, count, Dz, vel_fac,
                    disp_host, vel_host, out_code):
        ip = C.POINTER(C.c_int32)
        a3 = lambda t: (C.c_int32 * 3)(*[int(v) for v in t])
        self._ck(self.lib.nbe_process_box(
            self.h, C.c_void_p(in_host.ctypes.data), in_code, a3(size), a3(crop), a3(plen),
            crop_idx.ctypes.data_as(ip), add_idx0.ctypes.data_as(ip), int(first), int(count), float(Dz),
            float(vel_fac), C.c_void_p(disp_host.ctypes.data),
            C.c_void_p(vel_host.ctypes.data) if vel_host is not None else None, out_code))

    def process_box_blocks(self, in_host, in_code, size, crop, plen, crop_idx, add_idx0, first, count, Dz, vel_fac,
                           disp_blk, vel_blk, out_code):
        """nbe_process_box with the outputs left on the device as (count, 3, c0, c1, c2) torch tensors."""
        ip = C.POINTER(C.c_int32)
        a3 = lambda t: (C.c_int32 * 3)(*[int(v) for v in t])
        self._ck(self.lib.nbe_process_box_blocks(
            self.h, C.c_void_p(in_host.ctypes.data), in_code, a3(size), a3(crop), a3(plen),
            crop_idx.ctypes.data_as(ip), add_idx0.ctypes.data_as(ip), int(first), int(count), float(Dz),
            float(vel_fac), C.c_void_p(disp_blk.data_ptr()),
            C.c_void_p(vel_blk.data_ptr()) if vel_blk is not None else None, out_code))

    @staticmethod
    def process_box_multi(engines, in_host, in_code, size, crop, plen, crop_idx, add_idx0, first, count, Dz, vel_fac,
                          disp_host, vel_host, out_code, out_size0=0):
        """One call over several GPUs (nbe_process_box_multi): one host thread per engine inside the
        library, one shared pinned input, one shared output."""
        e0 = engines[0]
        ip = C.POINTER(C.c_int32)
        a3 = lambda t: (C.c_int32 * 3)(*[int(v) for v in t])
        hs = (C.c_void_p * len(engines))(*[e.h.value for e in engines])
        e0._ck(e0.lib.nbe_process_box_multi(
            hs, len(engines), C.c_void_p(in_host.ctypes.data), in_code, a3(size), a3(crop), a3(plen),
            crop_idx.ctypes.data_as(ip), add_idx0.ctypes.data_as(ip), int(first), int(count), float(Dz),
            float(vel_fac), C.c_void_p(disp_host.ctypes.data),
            C.c_void_p(vel_host.ctypes.data) if vel_host is not None else None, out_code, int(out_size0)))

    def process_box_dev(self, box_dev, size, crop, plen, crop_idx, add_idx0, first, count, Dz, vel_fac,
                        disp_dev, vel_dev):
        """Device-resident variant: torch CUDA tensors (3,size) in / out, async on the current stream."""
        torch = _torch()
        ip = C.POINTER(C.c_int32)
        a3 = lambda t: (C.c_int32 * 3)(*[int(v) for v in t])
        self._ck(self.lib.nbe_process_box_dev(
            self.h, C.c_void_p(box_dev.data_ptr()), dtype_code(box_dev.dtype), a3(size), a3(crop), a3(plen),
            crop_idx.ctypes.data_as(ip), add_idx0.ctypes.data_as(ip), int(first), int(count), float(Dz),
            float(vel_fac), C.c_void_p(disp_dev.data_ptr()),
            C.c_void_p(vel_dev.data_ptr()) if vel_dev is not None else None, dtype_code(disp_dev.dtype),
            C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))

    # ---- instrumentation
    def launch_count(self, reset=False):
        return int(self.lib.nbe_launch_count(self.h, int(reset)))

    def set_profiling(self, on):
        self._ck(self.lib.nbe_set_profiling(self.h, int(on)))

    def get_profile(self):
        cap = 64
        names = (C.c_char_p * cap)()
        ms = (C.c_float * cap)()
        fl = (C.c_double * cap)()
        n = self.lib.nbe_get_profile(self.h, cap, names, ms, fl)
        return [(names[i].decode(), float(ms[i]), float(fl[i])) for i in range(min(n, cap))]

    def release_workspace(self):
        """Free the activation arena / plans / staging buffers (re-created on demand)."""
        self._ck(self.lib.nbe_release_workspace(self.h))

    def workspace_bytes(self, dims):
        return int(self.lib.nbe_workspace_bytes(self.h, (C.c_int32 * 3)(*dims)))
