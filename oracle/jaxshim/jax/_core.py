"""Array type and dtype canonicalisation of the JAX stand-in (test infrastructure, see ../README.md).

JAX semantics that the reference relies on and plain numpy does not have:
  * default 32-bit types: python scalars, float64 and int64 inputs become float32 / int32
    unless ``jax_enable_x64`` is set;
  * ``x.at[idx].set(v)`` functional updates;
  * mixed int/float arithmetic stays in float32.
"""
import numpy as np

_X64 = [False]

_DOWN = {np.dtype(np.float64): np.dtype(np.float32), np.dtype(np.int64): np.dtype(np.int32),
         np.dtype(np.uint64): np.dtype(np.uint32), np.dtype(np.complex128): np.dtype(np.complex64)}


def x64_enabled():
    return _X64[0]


def canon_dtype(dt):
    if dt is None:
        return None
    dt = np.dtype(dt)
    return dt if _X64[0] else _DOWN.get(dt, dt)


class _AtIndex:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def _upd(self, fn):
        out = np.array(self.arr, copy=True).view(Array)
        fn(out.view(np.ndarray))
        return out

    def set(self, v):
        return self._upd(lambda o: o.__setitem__(self.idx, v))

    def add(self, v):
        return self._upd(lambda o: np.add.at(o, self.idx, v))

    def multiply(self, v):
        return self._upd(lambda o: np.multiply.at(o, self.idx, v))

    def get(self):
        return canon(self.arr.view(np.ndarray)[self.idx])


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIndex(self.arr, idx)


class Array(np.ndarray):
    """numpy array with JAX's ``.at`` and 32-bit result types."""
    __array_priority__ = 100

    @property
    def at(self):
        return _At(self)

    def block_until_ready(self):
        return self

    def astype(self, dtype, *a, **k):
        return np.ndarray.astype(self.view(np.ndarray), canon_dtype(dtype), *a, **k).view(Array)

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        ins = tuple(i.view(np.ndarray) if isinstance(i, Array) else i for i in inputs)
        if out is not None:
            kwargs["out"] = tuple(o.view(np.ndarray) if isinstance(o, Array) else o for o in out)
        res = getattr(ufunc, method)(*ins, **kwargs)
        if out is not None:
            return out[0] if len(out) == 1 else out
        # JAX's weak typing: bf16 array (op) python scalar stays bf16.  numpy does this itself for its own float16
        # (NEP 50) but promotes ml_dtypes' bfloat16 to float32.
        arrs = [i for i in ins if isinstance(i, np.ndarray)]
        if (arrs and all(a.dtype.name == "bfloat16" for a in arrs) and isinstance(res, np.ndarray) and res.dtype == np.float32
                and all(isinstance(i, (np.ndarray, int, float)) for i in ins)):
            res = res.astype(arrs[0].dtype)
        return canon(res)

    def __setitem__(self, k, v):
        raise TypeError("JAX arrays are immutable; use x.at[idx].set(v)")

    def __hash__(self):
        raise TypeError("unhashable type: Array")


def canon(out):
    """numpy result -> Array with JAX's default widths (recurses into tuples / lists)."""
    if isinstance(out, (tuple, list)):
        return type(out)(canon(o) for o in out)
    if isinstance(out, (np.ndarray, np.generic)):
        a = np.asarray(out)
        t = canon_dtype(a.dtype)
        if t != a.dtype:
            a = a.astype(t)
        return a.view(Array)
    return out


def asarray(x, dtype=None):
    if hasattr(x, "detach") and hasattr(x, "numpy"):     # torch tensor
        x = x.detach().cpu().numpy()
    if dtype is not None:
        return np.asarray(x, dtype=canon_dtype(dtype)).view(Array)
    return canon(np.asarray(x))
