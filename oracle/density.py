"""CPU oracle for the step right after the emulator: displacement -> density
contrast -> P(k).  TEST INFRASTRUCTURE ONLY: imported by tests/ and nothing else.

The reference delegates both steps to third-party packages that are absent from
/root/reference and from this image:

  * density:  `dj.get_delta_from_psi(psi, method="pm", res, worder, deconvolve)` of DISCO-DJ
    (scripts/core.py:396-409, 446-458; MAS orders 2/3/4 = CIC/TSC/PCS, scripts/utils.py:119-124);
  * spectrum: `PKL.Pk(delta, boxsize, axis=0, MAS, ...)` of Pylians3 (scripts/utils.py:1083-1090),
    read back through `.k3D`, `.Pk[:, 0]`, `.Nmodes3D`.

PARITY UNPINNED against those packages (they cannot be run here).  What is restated is their
published algorithm:

  * B-spline mass assignment of order p (p = 2 CIC, 3 TSC, 4 PCS) of unit-mass particles at
    x = q + psi on a periodic mesh, q the regular lattice; delta = rho / mean(rho) - 1;
  * optional deconvolution: delta_k /= prod_i sinc(pi k_i / res)^p;
  * Pylians' estimator: P(k) = |delta_k|^2 * L^3 / res^6 averaged over independent modes of the
    real FFT, bins of width k_F = 2 pi / L indexed by floor(|k| / k_F), k of a bin = mean |k| of
    its modes, the k = 0 mode excluded, MAS correction applied per mode before squaring.
"""
import numpy as np

MAS_ORDER = {"NGP": 1, "CIC": 2, "TSC": 3, "PCS": 4, "None": 0, None: 0}


def _weights(x, order):
    """Grid offsets and B-spline weights of order `order` for positions x (mesh units)."""
    if order == 1:
        i0 = np.floor(x + 0.5).astype(np.int64)
        return i0[None], np.ones((1,) + x.shape, x.dtype)
    if order == 2:
        i0 = np.floor(x).astype(np.int64)
        d = x - i0
        return np.stack([i0, i0 + 1]), np.stack([1 - d, d])
    if order == 3:
        ic = np.floor(x + 0.5).astype(np.int64)
        d = x - ic                                       # in [-0.5, 0.5)
        return (np.stack([ic - 1, ic, ic + 1]),
                np.stack([0.5 * (0.5 - d) ** 2, 0.75 - d * d, 0.5 * (0.5 + d) ** 2]))
    if order == 4:
        i0 = np.floor(x).astype(np.int64)
        d = x - i0                                       # in [0, 1)
        w0 = (1 - d) ** 3 / 6
        w1 = (4 - 6 * d ** 2 + 3 * d ** 3) / 6
        w2 = (4 - 6 * (1 - d) ** 2 + 3 * (1 - d) ** 3) / 6
        w3 = d ** 3 / 6
        return np.stack([i0 - 1, i0, i0 + 1, i0 + 2]), np.stack([w0, w1, w2, w3])
    raise ValueError(f"Unsupported mass-assignment order: {order}")


def delta_from_psi(psi, boxsize, res=None, worder=2, deconvolve=False, dtype=np.float64):
    """psi: (3, n0, n1, n2) displacement in the units of `boxsize`.  Returns delta (res, res, res)."""
    psi = np.asarray(psi, dtype)
    n = psi.shape[1:]
    res = int(res or n[0])
    rho = np.zeros((res, res, res), dtype)
    idx, wts = [], []
    for a in range(3):
        q = np.arange(n[a], dtype=dtype) * (res / n[a])
        shape = [1, 1, 1]; shape[a] = n[a]
        x = q.reshape(shape) + psi[a] * (res / boxsize)
        i, w = _weights(x.reshape(-1), worder)
        idx.append(np.mod(i, res)); wts.append(w)
    p = idx[0].shape[0]
    for a in range(p):
        for b in range(p):
            for c in range(p):
                np.add.at(rho, (idx[0][a], idx[1][b], idx[2][c]), wts[0][a] * wts[1][b] * wts[2][c])
    delta = rho * (res ** 3 / float(np.prod(n))) - 1
    if deconvolve:
        delta = deconvolve_mas(delta, worder)
    return delta


def _mas_window(res, order):
    k = np.fft.fftfreq(res, 1.0 / res)
    kz = np.arange(res // 2 + 1)
    w = lambda kk: np.sinc(kk / res) ** order
    return w(k)[:, None, None] * w(k)[None, :, None] * w(kz)[None, None, :]


def deconvolve_mas(delta, order):
    dk = np.fft.rfftn(delta)
    dk /= _mas_window(delta.shape[0], order)
    return np.fft.irfftn(dk, s=delta.shape, axes=(0, 1, 2))


def power_spectrum(delta, boxsize, MAS=None):
    """Pylians-style 3D monopole.  Returns (k3D, Pk, Nmodes3D) for bins 1 .. floor(sqrt(3) res / 2)."""
    delta = np.asarray(delta, np.float64)
    res = delta.shape[0]
    order = MAS_ORDER[MAS]
    dk = np.fft.rfftn(delta)
    if order:
        dk = dk / _mas_window(res, order)
    mid = res // 2
    kx = np.fft.fftfreq(res, 1.0 / res)
    if res % 2 == 0:
        kx[kx == -mid] = mid                            # Pylians keeps the Nyquist row at +res/2
    kz = np.arange(res // 2 + 1, dtype=np.float64)
    KX, KY, KZ = np.meshgrid(kx, kx, kz, indexing="ij")
    kmod = np.sqrt(KX ** 2 + KY ** 2 + KZ ** 2)
    # independent modes of the half-complex cube (the kz = 0 and kz = Nyquist planes hold each mode twice)
    plane = (KZ == 0) | ((KZ == mid) & (res % 2 == 0))
    dup = plane & ((KX < 0) | (((KX == 0) | ((KX == mid) & (res % 2 == 0))) & (KY < 0)))
    keep = ~dup & (kmod > 0)
    nb = int(np.sqrt(3.0) * mid) + 1
    kidx = kmod.astype(np.int64)
    kidx[~keep] = nb                                    # trash bin
    p = (dk.real ** 2 + dk.imag ** 2)
    pk = np.bincount(kidx.ravel(), weights=p.ravel(), minlength=nb + 1)[:nb]
    ks = np.bincount(kidx.ravel(), weights=kmod.ravel(), minlength=nb + 1)[:nb]
    nm = np.bincount(kidx.ravel(), minlength=nb + 1)[:nb].astype(np.float64)
    sel = slice(1, nb)
    kf = 2 * np.pi / boxsize
    with np.errstate(invalid="ignore", divide="ignore"):
        k3d = ks[sel] / nm[sel] * kf
        Pk = pk[sel] / nm[sel] * (boxsize / res ** 2) ** 3
    return k3d, Pk, nm[sel]


def za_displacement(delta, boxsize):
    """psi (3, n, n, n) with psi_k = i k / k^2 delta_k (scripts/core.py:396-397, DISCO-DJ 1LPT);
    k = 0 and the Nyquist component of each derivative zeroed."""
    delta = np.asarray(delta, np.float64)
    res = delta.shape[0]
    mid = res // 2
    kf = 2 * np.pi / boxsize
    k1 = np.fft.fftfreq(res, 1.0 / res)
    kz = np.arange(res // 2 + 1, dtype=np.float64)
    KX, KY, KZ = np.meshgrid(k1, k1, kz, indexing="ij")
    k2 = KX ** 2 + KY ** 2 + KZ ** 2
    dk = np.fft.rfftn(delta)
    with np.errstate(divide="ignore", invalid="ignore"):
        s = np.where(k2 > 0, 1.0 / (kf * k2), 0.0)
    out = []
    for K in (KX, KY, KZ):
        Kd = np.where((np.abs(K) == mid) & (res % 2 == 0), 0.0, K)
        out.append(np.fft.irfftn(1j * s * Kd * dk, s=delta.shape, axes=(0, 1, 2)))
    return np.stack(out)
