"""CPU checks of the density / P(k) oracle (oracle/density.py) against closed-form answers.

The reference takes both steps from DISCO-DJ and Pylians3 (absent here: parity unpinned against
them), so the oracle is pinned on what those algorithms must satisfy exactly."""
import numpy as np
import pytest

from oracle import density as D


@pytest.mark.parametrize("order", [1, 2, 3, 4])
def test_unperturbed_lattice_is_uniform(order):
    assert np.abs(D.delta_from_psi(np.zeros((3, 8, 8, 8)), 50.0, worder=order)).max() < 1e-12


@pytest.mark.parametrize("order", [2, 3, 4])
@pytest.mark.parametrize("res", [12, 16])
def test_mass_is_conserved_and_weights_are_a_partition_of_unity(order, res):
    rng = np.random.default_rng(order)
    psi = rng.standard_normal((3, 8, 8, 8)) * 7.0                 # several cells, wraps around the box
    d = D.delta_from_psi(psi, 40.0, res=res, worder=order)
    assert abs(d.mean()) < 1e-12
    assert d.min() >= -1 - 1e-12


def test_cic_single_particle_weights():
    psi = np.zeros((3, 1, 1, 1)); psi[:, 0, 0, 0] = [0.25, 0.5, 0.0]          # box 4, res 4 -> mesh units
    rho = (D.delta_from_psi(psi, 4.0, res=4, worder=2) + 1) / 64
    assert np.isclose(rho[0, 0, 0], 0.75 * 0.5) and np.isclose(rho[1, 0, 0], 0.25 * 0.5)
    assert np.isclose(rho[0, 1, 0], 0.75 * 0.5) and np.isclose(rho[1, 1, 0], 0.25 * 0.5)
    assert np.isclose(rho.sum(), 1.0)


def test_uniform_shift_by_one_cell_changes_nothing():
    rng = np.random.default_rng(3)
    psi = rng.standard_normal((3, 8, 8, 8))
    a = D.delta_from_psi(psi, 8.0, worder=3)
    b = D.delta_from_psi(psi + 1.0, 8.0, worder=3)                # one cell in every direction
    assert np.allclose(np.roll(a, (1, 1, 1), (0, 1, 2)), b, atol=1e-12)


@pytest.mark.parametrize("res", [15, 16])
def test_mode_count_and_white_noise_level(res):
    g = np.random.default_rng(res).standard_normal((res, res, res))
    k, P, N = D.power_spectrum(g, 100.0)
    # independent modes of a real field: (res^3 - n_self) / 2 + n_self - 1 (k = 0 dropped)
    n_self = 8 if res % 2 == 0 else 1
    assert N.sum() == (res ** 3 - n_self) // 2 + n_self - 1
    assert abs(np.average(P, weights=N) / (100.0 ** 3 / res ** 3) - 1) < 0.08
    kf = 2 * np.pi / 100.0
    assert np.all(k >= np.arange(1, len(k) + 1) * kf - 1e-12) and np.all(k < np.arange(2, len(k) + 2) * kf)


def test_single_plane_wave_lands_in_its_shell():
    res, L = 16, 10.0
    x = np.arange(res) / res
    d = 0.1 * np.cos(2 * np.pi * 3 * x)[:, None, None] * np.ones((1, res, res))
    k, P, N = D.power_spectrum(d, L)
    # |delta_k|^2 = (A/2)^2 at k = (+-3, 0, 0); one independent mode in shell 3
    expect = (0.05 ** 2) * L ** 3 / N[2]
    assert np.isclose(P[2], expect, rtol=1e-10)
    assert np.all(P[np.arange(len(P)) != 2] < 1e-20)


def test_deconvolution_inverts_the_window():
    rng = np.random.default_rng(5)
    d = rng.standard_normal((12, 12, 12))
    w = D._mas_window(12, 2)
    smooth = np.fft.irfftn(np.fft.rfftn(d) * w, s=d.shape, axes=(0, 1, 2))
    assert np.allclose(D.deconvolve_mas(smooth, 2), d, atol=1e-10)
    k, P0, _ = D.power_spectrum(d, 30.0)
    _, P1, _ = D.power_spectrum(smooth, 30.0, MAS="CIC")
    assert np.allclose(P0, P1, rtol=1e-9)


def test_zeldovich_plane_wave_and_divergence():
    res, L = 16, 40.0
    x = np.arange(res) * (L / res)
    kk = 2 * np.pi * 2 / L
    d = 0.2 * np.cos(kk * x)[None, :, None] * np.ones((res, 1, res))
    psi = D.za_displacement(d, L)
    assert np.allclose(psi[1], -0.2 / kk * np.sin(kk * x)[None, :, None] * np.ones((res, 1, res)), atol=1e-12)
    assert np.abs(psi[0]).max() < 1e-12 and np.abs(psi[2]).max() < 1e-12
    g = np.random.default_rng(0).standard_normal((12, 12, 12)); g -= g.mean()
    psi = D.za_displacement(g, 12.0)
    # spectral divergence of psi gives back -delta except on the zeroed Nyquist planes (none for these modes if we low-pass)
    gk = np.fft.rfftn(g); k1 = np.fft.fftfreq(12, 1 / 12.0); kz = np.arange(7.0)
    KX, KY, KZ = np.meshgrid(k1, k1, kz, indexing="ij")
    low = (np.abs(KX) < 6) & (np.abs(KY) < 6) & (KZ < 6)
    div = sum(1j * (2 * np.pi / 12.0) * K * np.fft.rfftn(psi[i]) for i, K in enumerate((KX, KY, KZ)))
    assert np.allclose(div[low], -gk[low], atol=1e-9)
