"""``jax.scipy.special`` stand-in: scipy's fp64 hyp2f1, result in JAX's default width."""
import numpy as _np
import scipy.special as _sp

from .._core import canon as _canon


def hyp2f1(a, b, c, x):
    return _canon(_sp.hyp2f1(*[_np.asarray(v, dtype=_np.float64) for v in (a, b, c, x)]))
