"""``jax.numpy`` stand-in: numpy functions with JAX's 32-bit defaults (see ../README.md)."""
import numpy as _np

from ._core import Array, asarray as _asarray, canon as _canon, canon_dtype as _canon_dtype

ndarray = Array
float16, float32, float64 = _np.float16, _np.float32, _np.float64
int8, int16, int32, int64 = _np.int8, _np.int16, _np.int32, _np.int64
uint8, uint32 = _np.uint8, _np.uint32
bool_ = _np.bool_
try:                                    # numpy has no bf16 of its own; ml_dtypes provides the scalar type
    import ml_dtypes as _ml
    bfloat16 = _ml.bfloat16
except ImportError:                     # pragma: no cover
    pass
complex64 = _np.complex64
pi, inf, nan, newaxis, e = _np.pi, _np.inf, _np.nan, None, _np.e
dtype = _np.dtype
finfo, iinfo = _np.finfo, _np.iinfo


def _unwrap(a):
    if isinstance(a, Array):
        return a.view(_np.ndarray)
    if isinstance(a, (list, tuple)):
        return type(a)(_unwrap(x) for x in a)
    return a


def _wrap(f):
    def g(*args, **kw):
        if kw.get("dtype", None) is not None:
            kw["dtype"] = _canon_dtype(kw["dtype"])
        return _canon(f(*[_unwrap(a) for a in args], **{k: _unwrap(v) for k, v in kw.items()}))
    g.__name__ = getattr(f, "__name__", "f")
    return g


def array(x, dtype=None, copy=True, ndmin=0):
    a = _asarray(x, dtype)
    if ndmin:
        a = _np.array(a, ndmin=ndmin).view(Array)
    return _np.array(a, copy=True).view(Array)


def asarray(x, dtype=None):
    return _asarray(x, dtype)


def astype(x, dtype):
    return _asarray(x).astype(dtype)


_NAMES = """abs absolute add all allclose any arange arccos arcsin arctan arctan2 argmax argmin argsort
atleast_1d atleast_2d atleast_3d broadcast_to ceil clip concatenate corrcoef cos cosh cumsum diff divide dot
einsum equal exp expand_dims eye flip floor full full_like greater isclose isfinite isinf isnan less linspace
log log10 log1p log2 logspace matmul max maximum mean meshgrid min minimum moveaxis multiply negative ones
ones_like outer pad power prod ravel real repeat reshape roll round sign sin sinh sqrt square squeeze stack
std subtract sum swapaxes take tan tanh tensordot tile transpose tril triu var where zeros zeros_like
array_equal median linalg fft""".split()
for _n in _NAMES:
    _f = getattr(_np, _n)
    globals()[_n] = _wrap(_f) if callable(_f) else _f
del _n, _f


def rsqrt(x):
    return 1.0 / sqrt(x)  # noqa: F821
