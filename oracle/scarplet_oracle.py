"""CPU oracle for the scarplet template-matching hot path.

TEST INFRASTRUCTURE ONLY.  This module is a float64 NumPy restatement of the
algorithm in the reference package (stgl/scarplet 0.1.4).  It exists so that the
CUDA path in ``scarplet_b200`` can be checked against it.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it.  The product package never does.

Parity is PINNED: ``tests/test_oracle_golden.py`` checks every function here
against the reference's own golden files (``scarplet/tests/results/*.npy``,
re-packed under ``tests/golden/`` by ``tests/golden/make_golden.py``) and
against outputs of the unmodified reference imported in the build container.

Third-party arithmetic the reference reaches through wheels that are not under
``/root/reference`` (all unpinned in ``requirements.txt:1-7``): FFTW3 via
``pyfftw`` (here: ``numpy.fft``, pocketfft), ``numexpr`` (here: plain NumPy),
``scipy.special.erfinv`` (same).

Every function cites the reference ``file:line`` it follows (paths relative to
the reference checkout).
"""

import multiprocessing as mp
from functools import partial

import numpy as np
from scipy.special import erfinv

EPS = np.spacing(1)  # core.py:339

SCARP = "scarp"
RICKER = "ricker"  # Channel is an alias of Ricker, WindowedTemplate.py:523-525
RIGHT_UPPER = "right_upper_break"
LEFT_UPPER = "left_upper_break"


class Grid(object):
    """Minimal stand-in for ``DEMGrid`` as the hot path consumes it
    (dem.py:203-218, 351-372): ``_griddata`` plus ``_georef_info.dx/dy``."""

    class _Geo(object):
        def __init__(self, dx, dy):
            self.dx = dx
            self.dy = dy

    def __init__(self, z, dx=1.0, dy=None):
        self._griddata = np.array(z, dtype=np.float64)
        self._georef_info = Grid._Geo(dx, dx if dy is None else dy)


# ---------------------------------------------------------------------------
# dem.py
# ---------------------------------------------------------------------------

def directional_laplacian(z, dx, dy, alpha):
    """Directional second derivative of the grid (dem.py:68-107).

    Does NOT mutate ``z`` (the reference zero-fills NaNs in place, dem.py:85-86;
    the values it computes are the same).
    """
    z = np.asarray(z, dtype=np.float64)
    nan_idx = np.isnan(z)
    if nan_idx.any():
        z = np.where(nan_idx, 0.0, z)
    ny, nx = z.shape

    # mixed derivative: first differences along x divided by dx, then first
    # differences of those along y divided by dx AGAIN (dem.py:88-89), zero
    # first column and first row (dem.py:90-93)
    gx = (z[:, 1:] - z[:, :-1]) / dx
    dxy = np.zeros((ny, nx))
    dxy[1:, 1:] = (gx[1:, :] - gx[:-1, :]) / dx

    # np.diff(z, 2, axis) is a difference of first differences
    dxx = np.zeros((ny, nx))
    fx = z[:, 1:] - z[:, :-1]
    dxx[:, 1:-1] = (fx[:, 1:] - fx[:, :-1]) / dx ** 2      # dem.py:95-97

    dyy = np.zeros((ny, nx))
    fy = z[1:, :] - z[:-1, :]
    dyy[1:-1, :] = (fy[1:, :] - fy[:-1, :]) / dy ** 2      # dem.py:99-101

    ca = np.cos(alpha)
    sa = np.sin(alpha)
    out = dxx * ca ** 2 - 2 * dxy * sa * ca + dyy * sa ** 2  # dem.py:103-104
    out[nan_idx] = np.nan                                    # dem.py:105
    return out


def laplacian(z, dx, dy):
    """dem.py:62-66"""
    return directional_laplacian(z, dx, dy, 0)


# ---------------------------------------------------------------------------
# WindowedTemplate.py
# ---------------------------------------------------------------------------

def axis_vectors(nx, ny, de):
    """Centred pixel coordinates (WindowedTemplate.py:50-53)."""
    x = de * np.linspace(1, nx, num=nx)
    y = de * np.linspace(1, ny, num=ny)
    x = x - np.mean(x)
    y = y - np.mean(y)
    return x, y


def rotated_coordinates(nx, ny, de, alpha):
    """WindowedTemplate.py:49-59; ``alpha`` is the template's (negated) angle."""
    x, y = axis_vectors(nx, ny, de)
    X, Y = np.meshgrid(x, y)
    xr = X * np.cos(alpha) + Y * np.sin(alpha)
    yr = -X * np.sin(alpha) + Y * np.cos(alpha)
    return xr, yr


def scarp_halfwidth(kt):
    """WindowedTemplate.py:156-157"""
    return abs(2 * np.sqrt(kt) * erfinv(0.9))


def template_constants(kind, scale, age, nx):
    """(c, d) of the curvature window for a template kind
    (Scarp: WindowedTemplate.py:150-157; Ricker: :485-492)."""
    if kind in (SCARP, RIGHT_UPPER, LEFT_UPPER):
        return scarp_halfwidth(age), scale
    if kind == RICKER:
        return nx, scale
    raise ValueError(kind)


def window_mask(nx, ny, de, alpha, c, d):
    """WindowedTemplate.py:61-64"""
    xr, yr = rotated_coordinates(nx, ny, de, alpha)
    return (abs(xr) < c) & (abs(yr) < d)


def window_limits(kind, nx, ny, de, alpha, c, d):
    """Edge mask (WindowedTemplate.py:66-84; Ricker override :494-495)."""
    if kind == RICKER:
        return np.zeros((ny, nx), dtype=bool)
    x4 = d * np.cos(alpha - np.pi / 2)
    y4 = d * np.sin(alpha - np.pi / 2)
    x1 = d * np.cos(alpha)
    y1 = d * np.sin(alpha)
    an_y = abs((x4 - x1) + 2 * c * np.cos(alpha - np.pi / 2))
    an_x = abs((y1 - y4) + 2 * c * np.sin(alpha - np.pi / 2))
    x, y = axis_vectors(nx, ny, de)
    X, Y = np.meshgrid(x, y)
    return ((X < (min(x) + an_x)) | (X > (max(x) - an_x))
            | (Y < (min(y) + an_y)) | (Y > (max(y) - an_y)))


def template_array(kind, scale, age, angle, nx, ny, de):
    """Windowed template raster for one (scale, age, angle).

    Scarp: WindowedTemplate.py:159-183.  Ricker/Channel: :497-520.
    RightFacingUpperBreakScarp: :246-255 (negated ``template_numexpr``; with a
    modern numexpr ``kt ** (3/2)`` is ``kt ** 1.5``, :209).
    LeftFacingUpperBreakScarp inherits ``Scarp.template``.
    """
    alpha = -angle                                   # :151, :489
    c, d = template_constants(kind, scale, age, nx)
    xr, yr = rotated_coordinates(nx, ny, de, alpha)
    mask = (abs(xr) < c) & (abs(yr) < d)
    if kind in (SCARP, LEFT_UPPER):
        kt = age
        W = (-xr / (2. * kt ** (3 / 2.) * np.sqrt(np.pi))) \
            * np.exp(-xr ** 2. / (4. * kt))
        return W * mask
    if kind == RIGHT_UPPER:
        kt = age
        W = (-xr / (2 * kt ** (3 / 2) * np.sqrt(np.pi))) \
            * np.exp(-xr ** 2 / (4 * kt))
        return -(W * mask)
    if kind == RICKER:
        f = age
        W = (1. - 2. * (np.pi * f * xr) ** 2.) * np.exp(-(np.pi * f * xr) ** 2.)
        return W * mask
    raise ValueError(kind)


def err_mask(kind, angle, nx, ny, de):
    """Optional SNR mask of the upper-break templates
    (WindowedTemplate.py:257-267, :294-304); None for the others."""
    if kind == RIGHT_UPPER:
        xr, _ = rotated_coordinates(nx, ny, de, -angle)
        return xr <= 0
    if kind == LEFT_UPPER:
        xr, _ = rotated_coordinates(nx, ny, de, -angle)
        return xr >= 0
    return None


# ---------------------------------------------------------------------------
# core.py
# ---------------------------------------------------------------------------

def match_template(z, dx, dy, kind, scale, age, angle):
    """One (scale, age, angle) fit over the whole raster (core.py:297-377).

    Returns ``(amp[ny,nx], age, angle, snr[ny,nx])``.
    """
    fft2, ifft2, fftshift = np.fft.fft2, np.fft.ifft2, np.fft.fftshift
    curv = directional_laplacian(z, dx, dy, angle)           # :340
    ny, nx = curv.shape
    de = dx                                                  # :343
    t = template_array(kind, scale, age, angle, nx, ny, de)  # :345-346

    M = t != 0                                               # :348
    fm2 = fft2(M)
    n = np.sum(M) + EPS                                      # :350
    fc = fft2(curv)
    ft = fft2(t)
    fc2 = fft2(curv ** 2)
    tsum = np.sum(t ** 2)                                    # :356

    xcorr = np.real(fftshift(ifft2(ft * fc)))                # :359
    amp = xcorr / tsum                                       # :360
    T1 = tsum * (amp ** 2)                                   # :362
    T3 = fftshift(ifft2(fc2 * fm2))                          # :363
    with np.errstate(divide='ignore', invalid='ignore'):
        error = (1 / n) * np.real(T1 - 2 * amp * xcorr + T3) + EPS  # :366
        snr = np.abs(T1 / error)                             # :367

    em = err_mask(kind, angle, nx, ny, de)                   # :369-371
    if em is not None:
        snr[em] = 0

    c, d = template_constants(kind, scale, age, nx)
    wl = window_limits(kind, nx, ny, de, -angle, c, d)       # :373-375
    amp[wl] = 0
    snr[wl] = 0
    return amp, age, angle, snr


def match_template_plugin(z, grid_dx, grid_dy, Template, scale, age, angle, **kwargs):
    """core.py:339-375 for ANY template class -- the plugin surface of the path: a class
    callable as ``Template(scale, age, angle, nx, ny, de)`` (core.py:345) with
    ``template()`` (:346), ``get_window_limits()`` (:373) and optionally ``get_err_mask()``
    (:369-371).  Same arithmetic as ``match_template`` above."""
    fft2, ifft2, fftshift = np.fft.fft2, np.fft.ifft2, np.fft.fftshift
    curv = directional_laplacian(z, grid_dx, grid_dy, angle)   # :340
    ny, nx = curv.shape
    template_obj = Template(scale, age, angle, nx, ny, grid_dx, **kwargs)   # :343-345 (kwargs may hold dx, dy)
    t = template_obj.template()                              # :346
    M = t != 0                                               # :348
    fm2 = fft2(M)
    n = np.sum(M) + EPS                                      # :350
    fc = fft2(curv)
    ft = fft2(t)
    fc2 = fft2(curv ** 2)
    tsum = np.sum(t ** 2)                                    # :356
    xcorr = np.real(fftshift(ifft2(ft * fc)))                # :359
    amp = xcorr / tsum                                       # :360
    T1 = tsum * (amp ** 2)                                   # :362
    T3 = fftshift(ifft2(fc2 * fm2))                          # :363
    with np.errstate(divide='ignore', invalid='ignore'):
        error = (1 / n) * np.real(T1 - 2 * amp * xcorr + T3) + EPS  # :366
        snr = np.abs(T1 / error)                             # :367
    if hasattr(template_obj, 'get_err_mask'):                # :369-371
        snr[template_obj.get_err_mask()] = 0
    mask = template_obj.get_window_limits()                  # :373-375
    amp[mask] = 0
    snr[mask] = 0
    return amp, age, angle, snr


def calculate_best_fit_parameters_plugin(z, dx, dy, Template, scale, age,
                                         ang_max=np.pi / 2, ang_min=-np.pi / 2):
    """core.py:139-195 with a plugin class (serial: the Pool only changes who computes)."""
    ny, nx = z.shape
    results = (match_template_plugin(z, dx, dy, Template, scale, age, a)
               for a in search_angles(ang_min, ang_max))
    return np.stack(compare(results, ny, nx))                # :186-193


def compare(results, ny, nx):
    """Running per-pixel best-SNR select (core.py:198-243).

    Strict ``>`` / ``<``: the first maximum wins and an exact tie zeroes the
    pixel in all four planes.
    """
    best_amp = np.zeros((ny, nx))
    best_age = np.zeros((ny, nx))
    best_angle = np.zeros((ny, nx))
    best_snr = np.zeros((ny, nx))
    with np.errstate(invalid='ignore'):
        for this_amp, this_age, this_angle, this_snr in results:
            keep = best_snr > this_snr
            take = best_snr < this_snr
            best_amp = keep * best_amp + take * this_amp
            best_age = keep * best_age + take * this_age
            best_angle = keep * best_angle + take * this_angle
            best_snr = keep * best_snr + take * this_snr
    return best_amp, best_age, best_angle, best_snr


def search_angles(ang_min=-np.pi / 2, ang_max=np.pi / 2):
    """1-degree orientation grid (core.py:173-175)."""
    ang_stepsize = 1
    num_angles = int((180 / np.pi) * (ang_max - ang_min) / ang_stepsize + 1)
    return np.linspace(ang_min, ang_max, num_angles)


def default_ages():
    """core.py:107, :286"""
    return 10 ** np.arange(0, 3.5, 0.1)


def _fit_angle(z, dx, dy, kind, scale, age, angle):
    return match_template(z, dx, dy, kind, scale, age, angle)


def calculate_best_fit_parameters(z, dx, dy, kind, scale, age,
                                  ang_max=np.pi / 2, ang_min=-np.pi / 2,
                                  processes=1):
    """Angle sweep at one age, reduced with ``compare`` (core.py:139-195).

    ``processes`` > 1 uses a ``multiprocessing.Pool`` with ordered ``imap`` and
    ``chunksize=1`` exactly as the reference does (core.py:180-183); the
    reference always uses ``mp.cpu_count()`` workers.
    Returns an ndarray (4, ny, nx): [amp, age, angle, snr] (core.py:190-193).
    """
    z = np.asarray(z, dtype=np.float64)
    ny, nx = z.shape
    angles = search_angles(ang_min, ang_max)
    if processes and processes > 1:
        with mp.Pool(processes=processes) as pool:
            work = partial(_fit_angle, z, dx, dy, kind, scale, age)
            best = compare(pool.imap(work, angles, chunksize=1), ny, nx)
    else:
        best = compare((match_template(z, dx, dy, kind, scale, age, a)
                        for a in angles), ny, nx)
    return np.stack(best)


def match(z, dx, dy, kind, processes=1, **kwargs):
    """core.py:266-294: one angle sweep when ``age`` is given, otherwise the
    35-age sweep reduced hierarchically (returns a tuple of four planes)."""
    if 'age' in kwargs:
        return calculate_best_fit_parameters(z, dx, dy, kind,
                                             processes=processes, **kwargs)
    ny, nx = np.asarray(z).shape
    stacks = [calculate_best_fit_parameters(z, dx, dy, kind, age=age,
                                            processes=processes, **kwargs)
              for age in default_ages()]
    return compare(stacks, ny, nx)


def calculate_best_fit_parameters_serial(z, dx, dy, kind, scale,
                                         ang_max=np.pi / 2,
                                         ang_min=-np.pi / 2,
                                         ages=None):
    """Flat angle-outer / age-inner sweep (core.py:65-136)."""
    z = np.asarray(z, dtype=np.float64)
    ny, nx = z.shape
    ages = default_ages() if ages is None else ages
    angles = search_angles(ang_min, ang_max)
    return compare((match_template(z, dx, dy, kind, scale, age, angle)
                    for angle in angles for age in ages), ny, nx)


def calculate_best_fit_parameters_serial_plugin(z, grid_dx, grid_dy, Template, scale, ang_max=np.pi / 2,
                                                ang_min=-np.pi / 2, ages=None, **kwargs):
    """core.py:65-136 with a plugin class: flat angle-outer / age-inner sweep, keyword
    arguments forwarded to the template's constructor (core.py:116-121)."""
    ny, nx = z.shape
    ages = default_ages() if ages is None else ages
    return compare((match_template_plugin(z, grid_dx, grid_dy, Template, scale, age, angle, **kwargs)
                    for angle in search_angles(ang_min, ang_max) for age in ages), ny, nx)


# ---------------------------------------------------------------------------
# dem.py, either side of the match path
# ---------------------------------------------------------------------------

def estimate_curvature_noiselevel(z, dx, dy, sigma=100):
    """``CalculationMixin._estimate_curvature_noiselevel`` (dem.py:152-179): for 180
    directions, mean and standard deviation of the directional Laplacian minus its Gaussian
    low-pass.  ``sigma`` is 100 in the reference (dem.py:172); a parameter here so that small
    test rasters exercise the reflecting boundary the same way."""
    from scipy import ndimage                                # dem.py:164
    angles = np.linspace(0, np.pi, num=180)                  # :166
    mean, sd = [], []
    z = np.array(z, dtype=np.float64)
    for alpha in angles:                                     # :171-175
        del2z = directional_laplacian(z, dx, dy, alpha)
        # the reference's Laplacian zero-fills NaN cells of the grid IN PLACE (dem.py:85-86), so
        # only the first direction sees them (as NaN, :105); the other 179 see zeros
        z[np.isnan(z)] = 0
        lowpass = ndimage.gaussian_filter(del2z, sigma)
        highpass = del2z - lowpass
        with np.errstate(all='ignore'):
            mean.append(np.nanmean(highpass))
            sd.append(np.nanstd(highpass))
    return angles, mean, sd


def fill_nodata_pass(z, max_search_distance):
    """One pass of the nodata fill that stands in for ``rasterio.fill.fillnodata``
    (dem.py:406-408; rasterio / GDAL are not installable here, so this restates the
    REPLACEMENT the CUDA path defines, not GDAL's algorithm -- parity with GDAL unpinned): a
    NaN cell becomes the inverse-distance-weighted mean of the nearest valid cell along each
    of the eight row / column / diagonal rays within ``max_search_distance`` cells."""
    ny, nx = z.shape
    out = z.copy()
    rays = [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1)]
    for r, c in np.argwhere(np.isnan(z)):
        wsum = vsum = 0.0
        for dr, dc in rays:
            step = 1.4142135623730951 if dr and dc else 1.0
            s = 1
            while s * step <= max_search_distance:
                rr, cc = r + s * dr, c + s * dc
                if rr < 0 or rr >= ny or cc < 0 or cc >= nx:
                    break
                v = z[rr, cc]
                if v == v:
                    w = 1.0 / (s * step)
                    wsum += w
                    vsum += w * v
                    break
                s += 1
        if wsum > 0:
            out[r, c] = vsum / wsum
    return out


def fill_nodata(z):
    """``DEMGrid._fill_nodata`` (dem.py:388-414): passes with the reference's search distance
    (half of the longest per-row / per-column nodata count, :403-405) until no NaN is left
    (:402)."""
    z = np.array(z, dtype=np.float64)
    num, prev = int(np.isnan(z).sum()), None
    while num > 0 and num != prev:
        mask = np.isnan(z)
        dist = max(mask.sum(axis=1).max(), mask.sum(axis=0).max()) / 2
        z = fill_nodata_pass(z, max(dist, 1.0))
        prev, num = num, int(np.isnan(z).sum())
    return z
