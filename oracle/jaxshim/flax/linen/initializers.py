"""``flax.linen.initializers`` stand-in (same call signatures; the draws are numpy's, not threefry)."""
import numpy as _np

from jax import random as _random
from jax._core import asarray as _asarray


def zeros(key, shape, dtype=_np.float32):
    return _asarray(_np.zeros(shape), dtype=dtype)


def ones(key, shape, dtype=_np.float32):
    return _asarray(_np.ones(shape), dtype=dtype)


def zeros_init():
    return zeros


def ones_init():
    return ones


def constant(value):
    return lambda key, shape, dtype=_np.float32: _asarray(_np.full(shape, value), dtype=dtype)


def normal(stddev=1e-2):
    return lambda key, shape, dtype=_np.float32: _asarray(_np.asarray(_random.normal(key, shape)) * stddev, dtype=dtype)


def variance_scaling(scale, mode, distribution, in_axis=-2, out_axis=-1):
    def init(key, shape, dtype=_np.float32):
        shape = tuple(shape)
        if len(shape) < 2:
            fan_in = fan_out = shape[0] if shape else 1
        else:
            rec = int(_np.prod(shape)) // (shape[in_axis] * shape[out_axis])
            fan_in, fan_out = shape[in_axis] * rec, shape[out_axis] * rec
        n = {"fan_in": fan_in, "fan_out": fan_out, "fan_avg": (fan_in + fan_out) / 2}[mode]
        return _asarray(_np.asarray(_random.normal(key, shape)) * _np.sqrt(scale / n), dtype=dtype)
    return init


def lecun_normal(in_axis=-2, out_axis=-1):
    return variance_scaling(1.0, "fan_in", "truncated_normal", in_axis, out_axis)


def he_normal(in_axis=-2, out_axis=-1):
    return variance_scaling(2.0, "fan_in", "truncated_normal", in_axis, out_axis)
