"""Flat-LCDM scalars for the emulator (host side; three floats per call).

Same definitions and names as the reference ``cosmology.py`` (growth_factor :33-40,
hubble_rate :43-46, growth_rate :100-113, dlogH_dloga :116-127, vel_norm :130-141,
acc_norm :144-155): D = a g(-OL a^3/Om) / g(-OL/Om) with g = 2F1(1, 1/3; 11/6; x).  The
reference differentiates log D with ``jax.jvp``; here the derivative of the closed form is
written out (d/dx 2F1(a,b;c;x) = ab/c 2F1(a+1,b+1;c+1;x)).  Evaluated with scipy in float64
and returned as float32 (the dtype the reference's JAX functions produce), shape-preserving.
"""
from __future__ import annotations

import numpy as np
from scipy.special import hyp2f1

_A, _B, _C = 1.0, 1.0 / 3.0, 11.0 / 6.0


def _g(x):
    return hyp2f1(_A, _B, _C, x)


def _dg(x):
    return (_A * _B / _C) * hyp2f1(_A + 1.0, _B + 1.0, _C + 1.0, x)


def _prep(z, Om):
    z = np.asarray(z, dtype=np.float64)
    Om = np.asarray(Om, dtype=np.float64)
    return z, Om


def _out(v):
    v = np.asarray(v, dtype=np.float32)
    return v if v.ndim else np.float32(v)


def _growth64(z, Om):
    a = 1.0 / (1.0 + z)
    OL = 1.0 - Om
    return a * _g(-OL * a ** 3 / Om) / _g(-OL / Om)


def _rate64(z, Om):
    a = 1.0 / (1.0 + z)
    x = -(1.0 - Om) * a ** 3 / Om
    return 1.0 + 3.0 * x * _dg(x) / _g(x)


def _hubble64(z, Om):
    return 100.0 * np.sqrt(Om * (1.0 + z) ** 3 + (1.0 - Om))


def _dlogH64(z, Om):
    m = Om * (1.0 + z) ** 3
    return -1.5 * m / (m + 1.0 - Om)


def growth_factor(z, Om):
    """Linear growth function for flat LambdaCDM, normalised to 1 at z = 0."""
    z, Om = _prep(z, Om)
    return _out(_growth64(z, Om))


def hubble_rate(z, Om):
    """Hubble parameter in [h km/s/Mpc]."""
    z, Om = _prep(z, Om)
    return _out(_hubble64(z, Om))


def growth_rate(z, Om):
    """f = dlogD/dloga."""
    z, Om = _prep(z, Om)
    return _out(_rate64(z, Om))


def dlogH_dloga(z, Om):
    z, Om = _prep(z, Om)
    return _out(_dlogH64(z, Om))


def vel_norm(z, Om):
    """Velocity normalisation D f H / (1+z) [km/s]."""
    z, Om = _prep(z, Om)
    return _out(_growth64(z, Om) * _rate64(z, Om) * _hubble64(z, Om) / (1.0 + z))


def acc_norm(z, Om):
    z, Om = _prep(z, Om)
    return _out(_growth64(z, Om) * _rate64(z, Om) * _hubble64(z, Om) ** 2 * _dlogH64(z, Om) / (1.0 + z))
