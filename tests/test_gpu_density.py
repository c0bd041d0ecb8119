"""GPU parity of the post-emulator step (displacement -> density -> P(k)) against
oracle/density.py, through the C ABI (nbe_density_from_psi / nbe_mas_deconvolve / nbe_pk_bins)."""
import numpy as np
import pytest
import torch

import jax_nbody_emulator_with_dj_b200 as nb
from oracle import density as D

pytestmark = pytest.mark.gpu

TOL_DELTA = 2e-5      # fp32 atomics in arbitrary order vs the fp64 oracle, relative to rms(1 + delta)
TOL_PK = 1e-4         # fp32 FFT (cuFFT) vs fp64 numpy FFT, per shell


def psi_field(n, seed, amp):
    return (np.random.default_rng(seed).standard_normal((3,) + n) * amp).astype(np.float32)


@pytest.mark.parametrize("order", [1, 2, 3, 4])
@pytest.mark.parametrize("n,res", [((16, 16, 16), 16), ((12, 16, 20), 24), ((16, 16, 16), 9)])
def test_density_matches_oracle(order, n, res):
    psi = psi_field(n, 7 + order, 3.0)                     # several cells: exercises the periodic wrap
    got = nb.get_delta_from_psi(psi, 20.0, res=res, worder=order)
    ref = D.delta_from_psi(psi.astype(np.float64), 20.0, res=res, worder=order)
    assert got.shape == (res, res, res) and got.dtype == np.float32
    scale = np.sqrt(np.mean((1 + ref) ** 2))
    if order == 1:            # NGP: a particle within fp32 rounding of a cell edge may land next door
        assert np.mean(np.abs(got - ref) > 1e-4) < 2e-3
    else:
        assert np.abs(got - ref).max() / scale < TOL_DELTA
    assert abs(float(got.astype(np.float64).mean())) < 1e-5


def test_unperturbed_lattice_and_mesh_layout():
    z = np.zeros((3, 32, 32, 32), np.float32)
    assert np.abs(nb.get_delta_from_psi(z, 1.0, worder=2)).max() == 0.0
    psi = psi_field((8, 8, 8), 1, 0.5)
    a = nb.get_delta_from_psi(psi, 8.0, worder=3)
    b = nb.get_delta_from_psi(np.moveaxis(psi, 0, -1), 8.0, worder=3)        # DISCO-DJ's (n, n, n, 3)
    assert np.array_equal(a, b) or np.abs(a - b).max() < 1e-5
    t = nb.get_delta_from_psi(torch.from_numpy(psi).cuda(), 8.0, worder=3)   # torch in -> torch out, no host trip
    assert isinstance(t, torch.Tensor) and t.is_cuda and np.abs(t.cpu().numpy() - a).max() < 1e-5


@pytest.mark.parametrize("res", [15, 16, 32])
@pytest.mark.parametrize("mas", [None, "CIC", "TSC", "PCS"])
def test_power_spectrum_matches_oracle(res, mas):
    d = np.random.default_rng(res).standard_normal((res, res, res)).astype(np.float32)
    got = nb.power_spectrum(d, 250.0, MAS=mas)
    k, P, N = D.power_spectrum(d.astype(np.float64), 250.0, MAS=mas)
    assert np.array_equal(got.Nmodes3D, N)                 # integer work: bit-exact
    ok = N > 0
    assert np.allclose(got.k3D[ok], k[ok], rtol=1e-12)
    assert np.allclose(got.Pk[ok], P[ok], rtol=TOL_PK)


def test_deconvolve_and_end_to_end_pk():
    psi = psi_field((32, 32, 32), 11, 0.8)
    for order in (2, 3, 4):
        got = nb.get_delta_from_psi(psi, 32.0, worder=order, deconvolve=True)
        ref = D.delta_from_psi(psi.astype(np.float64), 32.0, worder=order, deconvolve=True)
        assert np.abs(got - ref).max() / np.sqrt(np.mean((1 + ref) ** 2)) < 2e-4
        # painted field + MAS-corrected estimator == deconvolved field + plain estimator
        a = nb.power_spectrum(nb.get_delta_from_psi(psi, 32.0, worder=order), 32.0, MAS=nb.mas_name_from_worder(order))
        b = nb.power_spectrum(got, 32.0, MAS=None)
        assert np.allclose(a.Pk, b.Pk, rtol=2e-3)


def test_argument_errors():
    with pytest.raises(ValueError):
        nb.get_delta_from_psi(np.zeros((3, 4, 4, 4), np.float32), 1.0, worder=5)
    with pytest.raises(ValueError):
        nb.get_delta_from_psi(np.zeros((2, 4, 4, 4), np.float32), 1.0)
    with pytest.raises(ValueError):
        nb.power_spectrum(np.zeros((4, 4, 5), np.float32), 1.0)
    with pytest.raises(ValueError):
        nb.mas_name_from_worder(1)


def test_large_mesh_mass_conservation_and_shift():
    """BASELINE-size property check (256^3 particles): mass conservation and shift covariance."""
    g = torch.Generator("cuda").manual_seed(3)
    psi = torch.randn((3, 256, 256, 256), device="cuda", generator=g) * 1.7
    a = nb.get_delta_from_psi(psi, 256.0, worder=2)
    assert abs(float(a.double().mean())) < 1e-6
    b = nb.get_delta_from_psi(psi + 2.0, 256.0, worder=2)          # two cells along every axis
    assert float((torch.roll(a, (2, 2, 2), (0, 1, 2)) - b).abs().max()) < 2e-3


@pytest.mark.parametrize("res", [15, 16, 48])
def test_zeldovich_displacement_matches_oracle(res):
    d = np.random.default_rng(res).standard_normal((res, res, res)).astype(np.float32)
    got = nb.za_displacement_from_delta(d, 75.0)
    ref = D.za_displacement(d.astype(np.float64), 75.0)
    assert got.shape == (3, res, res, res) and got.dtype == np.float32
    assert np.sqrt(np.mean((got - ref) ** 2) / np.mean(ref ** 2)) < 1e-5      # fp32 cuFFT vs fp64 numpy
    # the emulator takes it as is: (3, n, n, n), units of the box
    t = nb.za_displacement_from_delta(torch.from_numpy(d).cuda(), 75.0)
    assert t.is_cuda and t.is_contiguous() and np.abs(t.cpu().numpy() - got).max() < 1e-5
