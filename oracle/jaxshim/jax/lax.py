"""``jax.lax`` stand-in: conv_general_dilated by its definition (see ../README.md)."""
import numpy as _np
import torch
import torch.nn.functional as F

from ._core import Array, asarray as _asarray, canon as _canon


def conv_general_dilated(lhs, rhs, window_strides, padding, lhs_dilation=None, rhs_dilation=None,
                         dimension_numbers=None, feature_group_count=1, batch_group_count=1,
                         precision=None, preferred_element_type=None):
    """out[n,o,p] = sum_{i,k} lhs_dilated_padded[n,i,p*stride + k*rhs_dil] * rhs[o,i,k]
    (cross-correlation, no kernel flip, as XLA defines it).  Layouts: NCDHW / OIDHW only."""
    if dimension_numbers is not None:
        assert tuple(dimension_numbers) == ("NCDHW", "OIDHW", "NCDHW"), dimension_numbers
    assert feature_group_count == 1 and batch_group_count == 1
    lhs, rhs = _np.asarray(lhs), _np.asarray(rhs)
    assert lhs.dtype == rhs.dtype, (lhs.dtype, rhs.dtype)  # jax.lax requires equal operand dtypes
    narrow = lhs.dtype if lhs.dtype.name in ("float16", "bfloat16") else None
    if narrow is not None:                                  # torch-CPU has no 16-bit conv3d: fp32 math, rounded result
        lhs, rhs = lhs.astype(_np.float32), rhs.astype(_np.float32)
    x = torch.from_numpy(_np.ascontiguousarray(lhs))
    w = torch.from_numpy(_np.ascontiguousarray(rhs))
    nd = x.dim() - 2
    assert nd == 3
    if lhs_dilation is not None and tuple(lhs_dilation) != (1,) * nd:
        ld = tuple(lhs_dilation)
        shape = list(x.shape[:2]) + [(n - 1) * d + 1 for n, d in zip(x.shape[2:], ld)]
        z = x.new_zeros(shape)
        z[(slice(None), slice(None)) + tuple(slice(None, None, d) for d in ld)] = x
        x = z
    if isinstance(padding, str):
        assert padding.upper() == "VALID", padding
    else:
        pads = [tuple(p) for p in padding]
        flat = []
        for lo, hi in reversed(pads):                       # F.pad takes the last dim first
            flat += [lo, hi]
        x = F.pad(x, flat)
    rd = tuple(rhs_dilation) if rhs_dilation is not None else 1
    y = F.conv3d(x, w, None, stride=tuple(window_strides), dilation=rd).numpy()
    return _canon(y.astype(narrow) if narrow is not None else y)


def rsqrt(x):
    return _canon(1.0 / _np.sqrt(_np.asarray(x)))


def stop_gradient(x):
    return x
