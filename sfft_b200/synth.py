"""
Seeded synthetic image pairs for parity tests and benchmarks.

A star field (density `density` px^-2, log-uniform fluxes) rendered with a Gaussian PSF
sigma_r on the reference and a spatially varying sigma_s(x, y) on the science image, a
flux scale, a polynomial differential background and Gaussian noise.  A fraction of the
stars is flagged "variable" and zeroed (stamp-wise) in the masked images mREF / mSCI.
Everything is float64; callers cast to the run precision.
"""
import numpy as np

__all__ = ["make_pair", "CONFIG_SEEDS"]

CONFIG_SEEDS = {1: 20261018, 2: 20261019, 3: 20261020, 4: 20261021, 5: 20261022}


def _render(shape, x, y, flux, sigma, half=7):
    N0, N1 = shape
    img = np.zeros(shape, dtype=np.float64)
    d = np.arange(-half, half + 1)
    ix, iy = np.rint(x).astype(np.int64), np.rint(y).astype(np.int64)
    fx, fy = x - ix, y - iy
    gx = np.exp(-0.5 * ((d[None, :] - fx[:, None]) / sigma[:, None]) ** 2)     # (n, 2h+1)
    gy = np.exp(-0.5 * ((d[None, :] - fy[:, None]) / sigma[:, None]) ** 2)
    norm = flux / (2.0 * np.pi * sigma ** 2)
    stamp = norm[:, None, None] * gx[:, :, None] * gy[:, None, :]              # (n, 2h+1, 2h+1)
    rr = (ix[:, None] + d[None, :]) % N0                                       # circular field: matches
    cc = (iy[:, None] + d[None, :]) % N1                                       # the circular model exactly
    np.add.at(img, (rr[:, :, None], cc[:, None, :]), stamp)
    return img, (ix, iy)


def make_pair(N0, N1, seed, varying_psf=True, density=2e-3, var_frac=0.01,
              flux_scale=1.3, bkg_coeffs=(5.0, 3.0, -2.0, 1.0), noise=(1.0, 1.5), chunk=8192):
    """Return dict(REF, SCI, mREF, mSCI) of float64 (N0, N1) arrays.

    bkg = c0 + c1*x + c2*y + c3*x*y with x, y in (0, 1]  (constant when varying_psf is False).
    """
    rng = np.random.default_rng(seed)
    nstar = max(8, int(density * N0 * N1))
    x = rng.uniform(0, N0, nstar)
    y = rng.uniform(0, N1, nstar)
    flux = 10.0 ** rng.uniform(2.0, 5.0, nstar)
    sig_r = np.full(nstar, 1.2)
    if varying_psf:
        sig_s = 1.6 + 0.4 * x / N0 + 0.2 * y / N1
    else:
        sig_s = np.full(nstar, 1.8)
    REF = np.zeros((N0, N1))
    SCI = np.zeros((N0, N1))
    for s in range(0, nstar, chunk):
        sl = slice(s, s + chunk)
        REF += _render((N0, N1), x[sl], y[sl], flux[sl], sig_r[sl])[0]
        SCI += _render((N0, N1), x[sl], y[sl], flux_scale * flux[sl], sig_s[sl])[0]
    cx = ((np.arange(N0) + 1.0) / N0)[:, None]
    cy = ((np.arange(N1) + 1.0) / N1)[None, :]
    if varying_psf:
        c0, c1, c2, c3 = bkg_coeffs
        bkg = c0 + c1 * cx + c2 * cy + c3 * cx * cy
    else:
        bkg = bkg_coeffs[0] + 0.0 * cx * cy
    REF += rng.normal(0.0, noise[0], (N0, N1))
    SCI += bkg + rng.normal(0.0, noise[1], (N0, N1))
    # variables: zero their stamps in the masked images
    nvar = max(1, int(var_frac * nstar))
    vidx = rng.choice(nstar, nvar, replace=False)
    mask = np.zeros((N0, N1), dtype=bool)
    d = np.arange(-7, 8)
    rr = (np.rint(x[vidx]).astype(np.int64)[:, None] + d[None, :]) % N0
    cc = (np.rint(y[vidx]).astype(np.int64)[:, None] + d[None, :]) % N1
    mask[rr[:, :, None], cc[:, None, :]] = True
    mREF, mSCI = REF.copy(), SCI.copy()
    mREF[mask] = 0.0
    mSCI[mask] = 0.0
    return dict(REF=REF, SCI=SCI, mREF=mREF, mSCI=mSCI)
