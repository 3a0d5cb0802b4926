"""Seeded synthetic DEMs: fractal relief + regional tilt + one
diffused scarp + white noise, stored as float32 and cast to float64 — the same path a
GDAL Float32 raster takes through ``DEMGrid.load`` (dem.py:317 of the reference)."""
import numpy as np
from scipy.special import erf


def synthetic_dem(n, seed, de=1.0, relief=30.0, nx=None):
    ny = int(n)
    nx = ny if nx is None else int(nx)
    rng = np.random.default_rng(seed)
    ky = np.fft.fftfreq(ny)[:, None]
    kx = np.fft.rfftfreq(nx)[None, :]
    k = np.sqrt(kx ** 2 + ky ** 2)
    k[0, 0] = 1
    shape = k.shape
    spec = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) * k ** -2.5
    spec[0, 0] = 0
    z = np.fft.irfft2(spec, s=(ny, nx))
    z = relief * z / z.std()
    y, x = np.mgrid[0:ny, 0:nx] * de
    xr = (x - x.mean()) * np.cos(0.3) + (y - y.mean()) * np.sin(0.3)
    z = z + 650 + 0.02 * x + 1.5 * erf(xr / (2 * np.sqrt(10.)))
    z = z + 0.03 * rng.standard_normal((ny, nx))
    return z.astype(np.float32).astype(np.float64)
