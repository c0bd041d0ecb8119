"""Displacement -> density contrast -> P(k): the step the reference runs right after the emulator.

The reference's driver does this through two third-party packages (absent here):
``dj.get_delta_from_psi(psi, method="pm", res=..., worder=..., deconvolve=...)`` of DISCO-DJ
(``scripts/core.py:396-409, 446-458``) and ``PKL.Pk(delta, boxsize, axis=0, MAS=...)`` of Pylians3
(``scripts/utils.py:1083-1090``).  Here the mass assignment, the MAS deconvolution and the shell
binning are CUDA kernels of ``libnbe_b200.so`` (``csrc/density.cuh``); only the FFT is a library call
(cuFFT through ``torch.fft``).  There is no CPU fallback.
"""
import ctypes as C
from collections import namedtuple

import numpy as np

from ._engine import Engine, _torch

MAS_ORDER = {"NGP": 1, "CIC": 2, "TSC": 3, "PCS": 4, "None": 0, None: 0}

PkResult = namedtuple("PkResult", ["k3D", "Pk", "Nmodes3D"])


def mas_name_from_worder(worder):
    """DISCO-DJ mass-assignment order -> Pylians MAS name (``scripts/utils.py:119-124``)."""
    mapper = {2: "CIC", 3: "TSC", 4: "PCS"}
    if worder not in mapper:
        raise ValueError(f"Unsupported mass-assignment order: {worder}")
    return mapper[worder]


def _dev(a, torch, dtype):
    if isinstance(a, torch.Tensor):
        return a.to(device="cuda", dtype=dtype).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to("cuda").to(dtype)


def get_delta_from_psi(psi, boxsize, res=None, worder=2, deconvolve=False, as_numpy=None):
    """Density contrast of unit-mass particles at ``q + psi`` (q: regular lattice) on a ``res``^3 mesh.

    ``psi``: (3, n0, n1, n2) — the emulator's output layout — or mesh layout (n0, n1, n2, 3) as
    DISCO-DJ takes it; units of ``boxsize``.  numpy in -> numpy out, torch in -> torch (cuda) out.
    """
    torch = _torch()
    eng = Engine.get()
    if as_numpy is None:
        as_numpy = not isinstance(psi, torch.Tensor)
    if psi.ndim != 4:
        raise ValueError(f"`psi` must be (3, n0, n1, n2) or (n0, n1, n2, 3), got shape={tuple(psi.shape)}")
    if psi.shape[0] != 3 and psi.shape[-1] == 3:
        psi = psi.permute(3, 0, 1, 2) if isinstance(psi, torch.Tensor) else np.moveaxis(psi, -1, 0)
    if psi.shape[0] != 3:
        raise ValueError(f"`psi` must have 3 components, got shape={tuple(psi.shape)}")
    if worder not in (1, 2, 3, 4):
        raise ValueError(f"Unsupported mass-assignment order: {worder}")
    x = _dev(psi, torch, torch.float32)
    n = (C.c_int32 * 3)(*x.shape[1:])
    res = int(res or x.shape[1])
    if res < 1:
        raise ValueError(f"`res` must be >= 1, got {res}.")
    delta = torch.empty((res, res, res), device="cuda", dtype=torch.float32)
    st = torch.cuda.current_stream().cuda_stream
    eng._ck(eng.lib.nbe_density_from_psi(eng.h, x.data_ptr(), n, float(boxsize), res, int(worder), delta.data_ptr(), st))
    if deconvolve:
        delta = deconvolve_mas_kernel(delta, worder, as_numpy=False)
    return delta.cpu().numpy() if as_numpy else delta


def deconvolve_mas_kernel(delta, worder, as_numpy=None):
    """Divide a gridded field by the MAS window of order ``worder`` (``scripts/utils.py:136-150``)."""
    torch = _torch()
    eng = Engine.get()
    if as_numpy is None:
        as_numpy = not isinstance(delta, torch.Tensor)
    d = _dev(delta, torch, torch.float32)
    if d.ndim != 3 or len(set(d.shape)) != 1:
        raise ValueError(f"`delta` must be cubic 3D, got shape={tuple(d.shape)}")
    res = d.shape[0]
    dk = torch.fft.rfftn(d).contiguous()
    st = torch.cuda.current_stream().cuda_stream
    eng._ck(eng.lib.nbe_mas_deconvolve(eng.h, dk.data_ptr(), res, int(worder), st))
    out = torch.fft.irfftn(dk, s=d.shape)
    return out.cpu().numpy() if as_numpy else out


def power_spectrum(delta, boxsize, MAS=None):
    """3D monopole with Pylians' conventions; returns ``PkResult(k3D, Pk, Nmodes3D)`` (numpy float64)."""
    torch = _torch()
    eng = Engine.get()
    if MAS not in MAS_ORDER:
        raise ValueError(f"Unsupported MAS: {MAS}")
    d = _dev(delta, torch, torch.float32)
    if d.ndim != 3 or len(set(d.shape)) != 1:
        raise ValueError(f"`delta` must be cubic 3D, got shape={tuple(d.shape)}")
    res = d.shape[0]
    nb = int(np.sqrt(3.0) * (res // 2)) + 1
    dk = torch.fft.rfftn(d).contiguous()
    out = torch.empty((3, nb), device="cuda", dtype=torch.float64)
    st = torch.cuda.current_stream().cuda_stream
    eng._ck(eng.lib.nbe_pk_bins(eng.h, dk.data_ptr(), res, MAS_ORDER[MAS], nb, out.data_ptr(), st))
    p, k, n = out.cpu().numpy()
    sel = slice(1, nb)
    with np.errstate(invalid="ignore", divide="ignore"):
        k3d = k[sel] / n[sel] * (2 * np.pi / boxsize)
        pk = p[sel] / n[sel] * (boxsize / res ** 2) ** 3
    return PkResult(k3d, pk, n[sel])


def za_displacement_from_delta(delta_linear, boxsize, as_numpy=None):
    """First-order LPT (Zel'dovich) displacement of a linear density field: div(psi) = -delta.

    The step before the emulator (``scripts/core.py:396-397``: ``dj.with_lpt(n_order=1)`` +
    ``evaluate_lpt_psi_at_a``; the growth factor is already in ``delta_linear`` when it is given at
    the target epoch).  Returns psi (3, n, n, n) fp32 in the units of ``boxsize`` — the emulator's
    input layout."""
    torch = _torch()
    eng = Engine.get()
    if as_numpy is None:
        as_numpy = not isinstance(delta_linear, torch.Tensor)
    d = _dev(delta_linear, torch, torch.float32)
    if d.ndim != 3 or len(set(d.shape)) != 1:
        raise ValueError(f"`delta_linear` must be cubic 3D, got shape={tuple(d.shape)}")
    res = d.shape[0]
    dk = torch.fft.rfftn(d).contiguous()
    pk = torch.empty((3,) + tuple(dk.shape), device="cuda", dtype=torch.complex64)
    st = torch.cuda.current_stream().cuda_stream
    eng._ck(eng.lib.nbe_za_psi_k(eng.h, dk.data_ptr(), res, float(boxsize), pk.data_ptr(), st))
    psi = torch.fft.irfftn(pk, s=d.shape, dim=(1, 2, 3)).contiguous()
    return psi.cpu().numpy() if as_numpy else psi
