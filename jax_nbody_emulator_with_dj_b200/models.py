"""The four core models with the reference's names, attributes and ``apply`` signatures.

  StyleNBodyEmulatorVelCore.apply(params, x, Om, Dz, vel_fac)   style_nbody_emulator_vel_core.py:105
  StyleNBodyEmulatorCore.apply(params, x, Om, Dz)               style_nbody_emulator_core.py:100
  NBodyEmulatorVelCore.apply(params, x, Dz, vel_fac)            nbody_emulator_vel_core.py:103
  NBodyEmulatorCore.apply(params, x, Dz)                        nbody_emulator_core.py

``x`` is (B, 3, D, H, W) with D, H, W multiples of 8 and >= 104; outputs are
(B, 3, D-96, H-96, W-96) in the dtype of ``x``.  numpy in -> numpy out; torch CUDA tensors
in -> torch CUDA tensors out (no host round trip).  All arithmetic runs in the sm_100a
kernels of libnbe_b200; there is no CPU implementation in this package.

Flax's ``model.init(key, x, ...)`` is offered as a seeded numpy initialiser with the same
tree (the JAX PRNG stream itself is not reproducible without JAX).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from ._engine import LAYERS, Engine, _torch


def init_params(seed=42, premodulated=False, compute_vel=True, dtype=np.float32):
    """Seeded parameter tree (PCG64).  Style tree: weight ~ N(0,1)/sqrt(fan_in),
    style_weight ~ N(0,1)/sqrt(cin), style_bias = 1 + 0.1 N(0,1), bias = 0.1 N(0,1)
    (bias / style_bias perturbed so that they are exercised; Flax's own init is
    lecun_normal / ones / zeros, style_layers_vel.py:55-75)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    P = {}
    for b, l, co, ci, k in LAYERS:
        fan_in = ci * k ** 3
        ent = {
            "weight": (rng.standard_normal((co, ci, k, k, k)) / np.sqrt(fan_in)).astype(dtype),
            "bias": (0.1 * rng.standard_normal(co)).astype(dtype),
            "style_weight": (rng.standard_normal((ci, 2)) / np.sqrt(ci)).astype(dtype),
            "style_bias": (1.0 + 0.1 * rng.standard_normal(ci)).astype(dtype),
        }
        if premodulated:
            sw = ent.pop("style_weight")
            ent.pop("style_bias")
            if compute_vel:
                ent["dweight"] = (ent["weight"] * (sw[:, 1] * 0.5)[None, :, None, None, None]).astype(dtype)
        P.setdefault(b, {})[l] = ent
    return {"params": P}


def _seed_of(key):
    if key is None:
        return 42
    try:
        return int(np.asarray(key).ravel()[-1])
    except Exception:
        return 42


def _prep_x(x):
    """-> (torch CUDA tensor, was_numpy, unbatched)."""
    torch = _torch()
    was_numpy = not isinstance(x, torch.Tensor)
    if was_numpy:
        xa = np.asarray(x)
        if xa.dtype not in (np.float32, np.float16):
            xa = xa.astype(np.float32)
        x = torch.from_numpy(np.ascontiguousarray(xa))
    if x.ndim != 5 or x.shape[1] != 3:
        raise ValueError(f"x must have shape (B, 3, D, H, W), got {tuple(x.shape)}")
    for n in x.shape[2:]:
        if n % 8 != 0 or n < 104:
            raise ValueError(f"spatial sizes must be multiples of 8 and >= 104 (out = in - 96), got {tuple(x.shape[2:])}")
    if not x.is_cuda:
        x = x.cuda(non_blocking=True)
    return x.contiguous(), was_numpy


def _ret(t, was_numpy):
    if isinstance(t, tuple):
        return tuple(_ret(v, was_numpy) for v in t)
    return t.cpu().numpy() if was_numpy else t


@dataclass
class _CoreBase:
    in_chan: int = 3
    out_chan: int = 3
    mid_chan: int = 64
    eps: float = 1e-8

    _premod = False
    _vel = False
    precision = "split"        # "split" (default, meets 1e-3 on velocity) or "fp16"

    def _check(self):
        if (self.in_chan, self.out_chan, self.mid_chan) != (3, 3, 64):
            raise NotImplementedError("the B200 kernels are specialised for in_chan=out_chan=3, mid_chan=64")
        if getattr(self, "style_size", 2) != 2:
            raise NotImplementedError("style_size must be 2 (Om, Dz)")

    def _run(self, params, x, Om, Dz, vel_fac):
        self._check()
        Engine.get()                      # fails loudly when no B200 / library is present
        x, was_numpy = _prep_x(x)
        B = x.shape[0]
        eng = Engine.get(x.device.index)
        eng.set_precision(self.precision)
        eng.set_params(params, self._premod, self._vel, self.eps)
        Dz = np.broadcast_to(np.atleast_1d(np.asarray(_host(Dz), dtype=np.float32)), (B,))
        if self._premod:
            eng.modulate(None, Dz)
        else:
            Om = np.broadcast_to(np.atleast_1d(np.asarray(_host(Om), dtype=np.float32)), (B,))
            eng.modulate(Om, Dz)
        vf = None
        if self._vel:
            vf = np.broadcast_to(np.atleast_1d(np.asarray(_host(vel_fac), dtype=np.float32)), (B,))
        return _ret(eng.forward(x, Dz, vf, self._vel), was_numpy)

    def init(self, key=None, *args, **kwargs):
        return init_params(_seed_of(key), premodulated=self._premod, compute_vel=self._vel)


def _host(v):
    if hasattr(v, "detach"):
        return v.detach().cpu().numpy()
    return v


@dataclass
class StyleNBodyEmulatorVelCore(_CoreBase):
    style_size: int = 2
    _premod = False
    _vel = True

    def apply(self, params, x, Om, Dz, vel_fac):
        return self._run(params, x, Om, Dz, vel_fac)


@dataclass
class StyleNBodyEmulatorCore(_CoreBase):
    style_size: int = 2
    _premod = False
    _vel = False

    def apply(self, params, x, Om, Dz):
        return self._run(params, x, Om, Dz, None)


@dataclass
class NBodyEmulatorVelCore(_CoreBase):
    _premod = True
    _vel = True

    def apply(self, params, x, Dz, vel_fac):
        return self._run(params, x, None, Dz, vel_fac)


@dataclass
class NBodyEmulatorCore(_CoreBase):
    _premod = True
    _vel = False

    def apply(self, params, x, Dz):
        return self._run(params, x, None, Dz, None)
