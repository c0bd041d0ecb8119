"""ctypes binding of libnbe_b200.so (C ABI in include/nbe.h) and its in-tree build.

The product path has no CPU fallback: if the shared library is missing or no sm_100 GPU is
visible, calls raise ``NBEError`` loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NBE_LIB") or os.path.join(HERE, "libnbe_b200.so")
CSRC = os.path.join(HERE, "csrc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "static"]

NBE_F32, NBE_F16, NBE_BF16 = 0, 1, 2
NBE_PREC_SPLIT, NBE_PREC_FP16 = 0, 1

EXPORTS = (
    "nbe_create", "nbe_destroy", "nbe_last_error", "nbe_version", "nbe_set_params", "nbe_set_precision",
    "nbe_modulate", "nbe_get_modulated", "nbe_forward", "nbe_process_box", "nbe_process_box_dev",
    "nbe_process_box_multi", "nbe_process_box_blocks",
    "nbe_workspace_bytes", "nbe_release_workspace", "nbe_host_register", "nbe_host_unregister",
    "nbe_launch_count", "nbe_set_profiling", "nbe_get_profile", "nbe_debug_read_act",
    "nbe_fold_active", "nbe_debug_act_fold",
    "nbe_density_from_psi", "nbe_mas_deconvolve", "nbe_pk_bins", "nbe_za_psi_k",
)


class NBEError(RuntimeError):
    pass


class LayerParams(C.Structure):
    _fields_ = [("block", C.c_char_p), ("layer", C.c_char_p), ("weight", C.c_void_p), ("dweight", C.c_void_p),
                ("bias", C.c_void_p), ("style_weight", C.c_void_p), ("style_bias", C.c_void_p),
                ("cout", C.c_int32), ("cin", C.c_int32), ("k", C.c_int32)]


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))]


def build(force=False, verbose=False):
    """Compile csrc/ for sm_100a into libnbe_b200.so (in-tree, so it travels with the repo)."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not force and os.path.exists(LIB_PATH):
        newest = max(os.path.getmtime(s) for s in sources() + [os.path.join(HERE, "..", "include", "nbe.h")])
        if os.path.getmtime(LIB_PATH) >= newest:
            return LIB_PATH
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB_PATH, os.path.join(CSRC, "nbe_api.cu")]
    if verbose:
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise NBEError("nvcc failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


_lib = None
_lock = threading.Lock()


def load():
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise NBEError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        vp, i32p, f32p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_float)
        lib.nbe_create.argtypes = [C.POINTER(vp), C.c_int]
        lib.nbe_destroy.argtypes = [vp]
        lib.nbe_destroy.restype = None
        lib.nbe_last_error.argtypes = [vp]
        lib.nbe_last_error.restype = C.c_char_p
        lib.nbe_version.restype = C.c_char_p
        lib.nbe_set_params.argtypes = [vp, C.POINTER(LayerParams), C.c_int, C.c_int, C.c_int, C.c_float]
        lib.nbe_set_precision.argtypes = [vp, C.c_int]
        lib.nbe_modulate.argtypes = [vp, f32p, f32p, C.c_int, vp]
        lib.nbe_get_modulated.argtypes = [vp, C.c_int, C.c_int, vp, vp]
        lib.nbe_forward.argtypes = [vp, vp, C.c_int, C.c_int, i32p, f32p, f32p, vp, vp, C.c_int, vp]
        lib.nbe_process_box.argtypes = [vp, vp, C.c_int, i32p, i32p, i32p, i32p, i32p, C.c_int, C.c_int,
                                        C.c_float, C.c_float, vp, vp, C.c_int]
        lib.nbe_process_box_multi.argtypes = [C.POINTER(vp), C.c_int, vp, C.c_int, i32p, i32p, i32p, i32p, i32p, C.c_int,
                                              C.c_int, C.c_float, C.c_float, vp, vp, C.c_int, C.c_int32]
        lib.nbe_process_box_blocks.argtypes = lib.nbe_process_box.argtypes
        lib.nbe_process_box_dev.argtypes = [vp, vp, C.c_int, i32p, i32p, i32p, i32p, i32p, C.c_int, C.c_int,
                                            C.c_float, C.c_float, vp, vp, C.c_int, vp]
        lib.nbe_host_register.argtypes = [vp, vp, C.c_size_t]
        lib.nbe_host_unregister.argtypes = [vp, vp]
        lib.nbe_density_from_psi.argtypes = [vp, vp, i32p, C.c_float, C.c_int32, C.c_int32, vp, vp]
        lib.nbe_mas_deconvolve.argtypes = [vp, vp, C.c_int32, C.c_int32, vp]
        lib.nbe_pk_bins.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_int32, vp, vp]
        lib.nbe_za_psi_k.argtypes = [vp, vp, C.c_int32, C.c_float, vp, vp]
        lib.nbe_release_workspace.argtypes = [vp]
        lib.nbe_workspace_bytes.argtypes = [vp, i32p]
        lib.nbe_workspace_bytes.restype = C.c_size_t
        lib.nbe_launch_count.argtypes = [vp, C.c_int]
        lib.nbe_launch_count.restype = C.c_int64
        lib.nbe_set_profiling.argtypes = [vp, C.c_int]
        lib.nbe_get_profile.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p), f32p, C.POINTER(C.c_double)]
        lib.nbe_debug_read_act.argtypes = [vp, C.c_int, C.c_int, vp, C.c_size_t, i32p]
        lib.nbe_debug_read_act.restype = C.c_longlong
        lib.nbe_fold_active.argtypes = [vp]
        lib.nbe_debug_act_fold.argtypes = [vp, C.c_int, C.c_int, f32p, C.c_int]
        for name in EXPORTS:
            getattr(lib, name)
        _lib = lib
        return lib
