"""Stand-in for the ``jax`` package: exactly the subset the reference sources call, on numpy and
torch-CPU (test infrastructure, see ../README.md).  Not the JAX runtime."""
import functools

import numpy as _np

from . import _core
from ._core import Array

__version__ = "0.0-shim"


class _Config:
    def update(self, name, value):
        if name == "jax_enable_x64":
            _core._X64[0] = bool(value)

    @property
    def jax_enable_x64(self):
        return _core._X64[0]

    def read(self, name):
        return _core._X64[0] if name == "jax_enable_x64" else None


config = _Config()


def jit(fun=None, **_kw):
    """Identity (also as ``@jax.jit`` / ``@partial(jax.jit, static_argnums=...)``)."""
    if fun is None:
        return lambda f: f
    return fun


def device_put(x, device=None):
    return _core.asarray(x) if not isinstance(x, (dict, list, tuple)) else tree_util.tree_map(_core.asarray, x)


def devices(backend=None):
    return ["cpu:0 (numpy/torch stand-in)"]


def default_backend():
    return "cpu"


def vmap(fun, in_axes=0, out_axes=0):
    """Loop over the mapped axis and stack (semantics of jax.vmap for the cases the reference uses:
    in_axes an int, None, or a tuple of those; out_axes 0)."""
    assert out_axes == 0

    @functools.wraps(fun)
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = {_np.shape(a)[ax] for a, ax in zip(args, axes) if ax is not None}
        assert len(n) == 1, f"vmap: inconsistent mapped sizes {n}"
        outs = []
        for i in range(n.pop()):
            outs.append(fun(*[a if ax is None else _np.take(_core.asarray(a), i, axis=ax).view(Array)
                              for a, ax in zip(args, axes)]))
        if isinstance(outs[0], (tuple, list)):
            return type(outs[0])(numpy.stack([o[j] for o in outs]) for j in range(len(outs[0])))
        return numpy.stack(outs)
    return mapped


def jvp(fun, primals, tangents):
    """Forward derivative by a 4th-order central difference in fp64.  The reference only
    differentiates scalar cosmology functions (cosmology.py:60-75); the network tangent is
    hand-written there, not traced."""
    assert len(primals) == 1 and len(tangents) == 1
    was = _core._X64[0]
    out0 = fun(primals[0])
    _core._X64[0] = True
    try:
        x = _np.asarray(primals[0], dtype=_np.float64)
        t = _np.asarray(tangents[0], dtype=_np.float64)
        h = 1e-3
        f = lambda e: _np.asarray(fun((x + e * t).view(Array)), dtype=_np.float64)
        d = (8.0 * (f(h) - f(-h)) - (f(2 * h) - f(-2 * h))) / (12.0 * h)
    finally:
        _core._X64[0] = was
    return out0, _core.asarray(d, dtype=_np.asarray(out0).dtype if was else _core.canon_dtype(_np.asarray(out0).dtype))


def grad(*a, **k):
    raise NotImplementedError("jax.grad is not part of the stand-in (only the reference's tests use it)")


value_and_grad = grad


class custom_jvp:
    def __init__(self, fun, nondiff_argnums=()):
        self.fun = fun
        functools.update_wrapper(self, fun)

    def defjvp(self, f):
        return f

    def __call__(self, *a, **k):
        return self.fun(*a, **k)


from . import numpy, lax, nn, random, tree_util, scipy  # noqa: E402,F401
from .tree_util import tree_map  # noqa: E402,F401
