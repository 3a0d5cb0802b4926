"""float64 NumPy model of the GPU schedule (padded pow2 domain, r2c half spectra,
support-box template, offset bookkeeping) checked against the oracle."""
import sys; sys.path.insert(0, '.')
import numpy as np
from oracle import scarplet_oracle as O

def nextpow2(n):
    p = 1
    while p < n: p *= 2
    return p

def model_match_template(z, dx, dy, kind, scale, age, angle, force_pad=False, dtype=np.float64):
    ny, nx = z.shape
    x, y = O.axis_vectors(nx, ny, dx)
    alpha = -angle
    ca, sa = np.cos(alpha), np.sin(alpha)
    c, d = O.template_constants(kind, scale, age, nx)
    # support box (conservative), offsets relative to a0 = ny//2, b0 = nx//2
    ceff = c
    if kind == O.RICKER:
        ceff = min(c, np.sqrt(745.2) / (np.pi * age) + dx)
    ex = ceff * abs(ca) + d * abs(sa); ey = ceff * abs(sa) + d * abs(ca)
    a0, b0 = ny // 2, nx // 2
    sx_lo = max(-int(ex / dx) - 2, -b0); sx_hi = min(int(ex / dx) + 2, nx - 1 - b0)
    sy_lo = max(-int(ey / dx) - 2, -a0); sy_hi = min(int(ey / dx) + 2, ny - 1 - a0)
    # template on the box, float64, exactly like the reference
    bb = np.arange(sx_lo, sx_hi + 1); aa = np.arange(sy_lo, sy_hi + 1)
    X, Y = np.meshgrid(x[b0 + bb], y[a0 + aa])
    xr = X * ca + Y * sa; yr = -X * sa + Y * ca
    mask = (abs(xr) < c) & (abs(yr) < d)
    if kind == O.SCARP:
        W = (-xr / (2. * age ** (3 / 2.) * np.sqrt(np.pi))) * np.exp(-xr ** 2. / (4. * age)) * mask
    else:
        W = (1. - 2. * (np.pi * age * xr) ** 2.) * np.exp(-(np.pi * age * xr) ** 2.) * mask
    M = W != 0
    n = M.sum() + O.EPS; ts = (W ** 2).sum()
    # geometry per axis
    def geom(N, lo, hi):
        delta = -(N & 1)
        if (N & (N - 1)) == 0 and not force_pad:
            return N, delta, N          # periodic, P == N
        P = nextpow2(N + hi - lo + 1)
        split = N - delta - lo + 1      # q < split -> s = q, else s = q - P
        return P, delta, split
    Py, dly, spy = geom(ny, sy_lo, sy_hi); Px, dlx, spx = geom(nx, sx_lo, sx_hi)
    qy = np.arange(Py); sy = np.where(qy < spy, qy, qy - Py); gy = np.mod(sy, ny)
    qx = np.arange(Px); sx = np.where(qx < spx, qx, qx - Px); gx = np.mod(sx, nx)
    curv = O.directional_laplacian(z, dx, dy, angle)
    Cp = curv[np.ix_(gy, gx)].astype(dtype)
    C2p = (curv ** 2)[np.ix_(gy, gx)].astype(dtype)
    Tp = np.zeros((Py, Px), dtype); Mp = np.zeros((Py, Px), dtype)
    Tp[np.ix_(aa % Py, bb % Px)] = W; Mp[np.ix_(aa % Py, bb % Px)] = M
    # half spectra along x (r2c), full along y
    FC = np.fft.rfft2(Cp); FC2 = np.fft.rfft2(C2p); FT = np.fft.rfft2(Tp); FM = np.fft.rfft2(Mp)
    Gt = np.fft.ifft(FT * FC, axis=0); Gm = np.fft.ifft(FM * FC2, axis=0)   # inverse column pass
    # inverse row pass: one packed complex FFT of X = Gt + i Gm with Hermitian extension along kx
    Xf = np.zeros((Py, Px), complex)
    h = Px // 2
    Xf[:, :h + 1] = Gt + 1j * Gm
    k = np.arange(1, h)
    Xf[:, Px - k] = np.conj(Gt[:, k]) + 1j * np.conj(Gm[:, k])
    out = np.fft.ifft(Xf, axis=1)
    xc_p = out.real; T3_p = out.imag
    my = (np.arange(ny) - dly) % Py; mx = (np.arange(nx) - dlx) % Px
    xc = xc_p[np.ix_(my, mx)]; T3 = T3_p[np.ix_(my, mx)]
    amp = xc / ts; T1 = ts * amp ** 2
    err = (T1 - 2 * amp * xc + T3) / n + O.EPS
    snr = np.abs(T1 / err)
    wl = O.window_limits(kind, nx, ny, dx, alpha, c, d)
    amp[wl] = 0; snr[wl] = 0
    return amp, snr, (Py, Px)

rng = np.random.default_rng(0)
for (ny, nx, kind, scale, age, angle, fp) in [
        (64, 64, O.SCARP, 8, 2.0, 0.3, False), (64, 64, O.SCARP, 8, 2.0, 0.3, True), (61, 75, O.SCARP, 8, 2.0, -1.1, False),
        (60, 77, O.SCARP, 6, 4.0, np.pi / 2, False), (50, 50, O.SCARP, 40, 10.0, 0.2, False),
        (64, 128, O.RICKER, 5, 0.2, 0.7, False), (63, 90, O.RICKER, 5, 0.2, -0.4, False), (63, 90, O.RICKER, 5, 0.02, -0.4, False)]:
    z = np.cumsum(np.cumsum(rng.standard_normal((ny, nx)), 0), 1) * 0.01
    a0, _, _, s0 = O.match_template(z, 1.0, 1.0, kind, scale, age, angle)
    a1, s1, P = model_match_template(z, 1.0, 1.0, kind, scale, age, angle, fp)
    print(ny, nx, kind, scale, age, round(angle, 2), 'P', P, 'valid', (s0 > 0).sum(), 'amp err', np.abs(a1 - a0).max() / max(np.abs(a0).max(), 1e-300),
          'snr relerr', (np.abs(s1 - s0) / np.maximum(s0, 1e-300))[s0 > 0].max() if (s0 > 0).any() else None, 'masks', ((s0 > 0) == (s1 > 0)).all())
