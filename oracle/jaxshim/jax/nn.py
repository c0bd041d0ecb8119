"""``jax.nn`` stand-in (see ../README.md)."""
import numpy as _np

from ._core import canon as _canon


def leaky_relu(x, negative_slope=0.01):
    """jax.nn.leaky_relu: where(x >= 0, x, negative_slope * x)."""
    a = _np.asarray(x)
    return _canon(_np.where(a >= 0, a, _np.asarray(negative_slope, dtype=a.dtype) * a))


def relu(x):
    a = _np.asarray(x)
    return _canon(_np.maximum(a, 0))
