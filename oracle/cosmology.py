"""CPU oracle of the flat-LCDM scalars (test infrastructure, see oracle/__init__.py).

Follows the *definitions* in /root/reference/src/jax_nbody_emulator/cosmology.py
(growth_factor :33-40, hubble_rate :43-46, growth_rate :100-113, vel_norm :130-141) but
evaluates them by an independent route -- the growth integral

    D(a) ∝ H(a) ∫_0^a da' / (a' H(a'))^3          (exact for flat LCDM)

with scipy quadrature in float64 -- instead of the 2F1 closed form the reference (and the
product package) use, so that agreement is a real check.  Pinned against the values quoted
in the reference README.md:178-180 and the limits asserted in tests/test_cosmology.py.
"""
from __future__ import annotations

import numpy as np
from scipy.integrate import quad


def _E(a, Om):
    return np.sqrt(Om / a ** 3 + (1.0 - Om))


def _growth_unnorm(a, Om):
    integ, _ = quad(lambda x: 1.0 / (x * _E(x, Om)) ** 3, 0.0, a, epsabs=0, epsrel=1e-12, limit=200)
    return _E(a, Om) * integ


def growth_factor(z, Om):
    z, Om = np.broadcast_arrays(np.asarray(z, dtype=np.float64), np.asarray(Om, dtype=np.float64))
    out = np.empty(z.shape, dtype=np.float64)
    for i in np.ndindex(z.shape):
        out[i] = _growth_unnorm(1.0 / (1.0 + z[i]), Om[i]) / _growth_unnorm(1.0, Om[i])
    return out


def hubble_rate(z, Om):
    z = np.asarray(z, dtype=np.float64)
    Om = np.asarray(Om, dtype=np.float64)
    return 100.0 * np.sqrt(Om * (1.0 + z) ** 3 + (1.0 - Om))


def growth_rate(z, Om):
    """f = dlnD/dlna, analytic derivative of the growth integral:
    f = dlnE/dlna + 1 / (a^2 E^3 I(a))."""
    z, Om = np.broadcast_arrays(np.asarray(z, dtype=np.float64), np.asarray(Om, dtype=np.float64))
    out = np.empty(z.shape, dtype=np.float64)
    for i in np.ndindex(z.shape):
        a = 1.0 / (1.0 + z[i])
        om = Om[i]
        E = _E(a, om)
        integ, _ = quad(lambda x: 1.0 / (x * _E(x, om)) ** 3, 0.0, a, epsabs=0, epsrel=1e-12, limit=200)
        dlnE = -1.5 * om / a ** 3 / E ** 2
        out[i] = dlnE + 1.0 / (a ** 2 * E ** 3 * integ)
    return out


def dlogH_dloga(z, Om):
    z = np.asarray(z, dtype=np.float64)
    Om = np.asarray(Om, dtype=np.float64)
    a = 1.0 / (1.0 + z)
    return -1.5 * Om / a ** 3 / (Om / a ** 3 + 1.0 - Om)


def vel_norm(z, Om):
    z = np.asarray(z, dtype=np.float64)
    return growth_factor(z, Om) * growth_rate(z, Om) * hubble_rate(z, Om) / (1.0 + z)


def acc_norm(z, Om):
    z = np.asarray(z, dtype=np.float64)
    return (growth_factor(z, Om) * growth_rate(z, Om) * hubble_rate(z, Om) ** 2
            * dlogH_dloga(z, Om) / (1.0 + z))
