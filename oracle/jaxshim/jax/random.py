"""``jax.random`` stand-in: numpy Generators keyed by the PRNGKey integers (NOT threefry; the
reference never depends on the random stream, its tests only need reproducible draws)."""
import numpy as _np

from ._core import asarray as _asarray, canon_dtype as _canon_dtype


def PRNGKey(seed):
    return _asarray(_np.array([0, int(seed) & 0xFFFFFFFF], dtype=_np.uint32))


key = PRNGKey


def _gen(k):
    k = _np.asarray(k).astype(_np.uint64).ravel()
    return _np.random.default_rng([int(v) for v in k])


def split(k, num=2):
    g = _gen(k)
    return _asarray(g.integers(0, 2 ** 32, size=(num, 2), dtype=_np.uint64).astype(_np.uint32))


def normal(k, shape=(), dtype=_np.float32):
    return _asarray(_gen(k).standard_normal(shape), dtype=_canon_dtype(dtype))


def uniform(k, shape=(), dtype=_np.float32, minval=0.0, maxval=1.0):
    return _asarray(_gen(k).uniform(minval, maxval, shape), dtype=_canon_dtype(dtype))


def randint(k, shape, minval, maxval, dtype=_np.int32):
    return _asarray(_gen(k).integers(minval, maxval, size=shape), dtype=_canon_dtype(dtype))


def fold_in(k, data):
    g = _np.random.default_rng([int(v) for v in _np.asarray(k).astype(_np.uint64).ravel()] + [int(data) & 0xFFFFFFFF])
    return _asarray(g.integers(0, 2 ** 32, size=(2,), dtype=_np.uint64).astype(_np.uint32))
