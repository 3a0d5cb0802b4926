"""Host-side mirror of scarplet/core.py for the template-matching path.

Same function names, positional order, keyword arguments and return layouts as the
reference; the work is done by the CUDA library behind ``engine.Plan``:

* the Pool fan-out over orientations (core.py:180-183) and the age list
  comprehension (core.py:288-291) become one batched sweep on the device;
* ``match_template``'s six FFTs and numexpr passes (core.py:340-375) become the
  kernels in ``csrc/sb_kernels.cuh``;
* the serial ``compare`` fold (core.py:227-241) runs in registers inside the last
  kernel (first maximum wins; see DESIGN.md for the one documented difference on
  exact float64 ties).

Plans (device workspace, twiddle tables, page-locked result buffers) are cached per raster
geometry, so that a second ``match`` on a raster of the same shape pays for nothing but the
DEM upload, the search and the result download; ``release()`` frees them.
"""
import threading

import numpy as np

from . import engine
from . import params as P
from .engine import Plan
from .templates import device_spec


def _grid_fields(data):
    z = data._griddata
    return z, data._georef_info.dx, data._georef_info.dy


# ---------------------------------------------------------------------------
# plan cache
# ---------------------------------------------------------------------------
_CACHE_SLOTS = 2
_cache = []                      # [(key, plan)], most recently used last
_cache_lock = threading.Lock()


class _Lease(object):
    """A cached plan taken out for one call (context manager: goes back on exit)."""

    def __init__(self, key, plan):
        self.key, self.plan = key, plan

    def __enter__(self):
        return self.plan

    def __exit__(self, exc_type, exc, tb):
        if exc_type is not None:        # a failed call leaves no half-used plan behind
            self.plan.close()
            return False
        with _cache_lock:
            _cache.append((self.key, self.plan))
            while len(_cache) > _CACHE_SLOTS:
                _cache.pop(0)[1].close()
        return False


def _plan_for(data, states=1):
    """A plan for ``data``'s geometry with its DEM uploaded, out of the cache when one is
    free (a plan is used by one call at a time: concurrent callers get their own)."""
    z, dx, dy = _grid_fields(data)
    ny, nx = z.shape
    d = engine.DEFAULTS
    key = (ny, nx, float(dx), float(dy), d["precision"], d["workspace_mb"], d["max_fft"])
    plan = None
    with _cache_lock:
        for i, (k, p) in enumerate(_cache):
            if k == key:
                plan = _cache.pop(i)[1]
                break
    for attempt in (0, 1):
        try:
            if plan is None:
                plan = Plan(ny, nx, dx, dy)
            if plan.states != states:
                plan.set_states(states)
            plan.set_dem(z)
            break
        except engine.SbError:
            # most likely device memory held by cached plans of other shapes: drop them, try once more
            if plan is not None:
                plan.close()
                plan = None
            if attempt == 1:
                raise
            release()
    return _Lease(key, plan)


def release():
    """Free the cached plans (device workspace and page-locked result buffers)."""
    with _cache_lock:
        while _cache:
            _cache.pop()[1].close()


# ---------------------------------------------------------------------------
# plugin templates (classes without an on-device generator)
# ---------------------------------------------------------------------------
def _plugin_match_template(plan, data, Template, scale, age, angle, **kwargs):
    """core.py:339-375 for a template class without an on-device generator -- the plugin
    surface: the class is instantiated and asked for its raster and masks exactly as the
    reference does; correlation and fit run on the device (``sb_match_template_raster``).
    One host-rendered raster per call: slow by design, any ``WindowedTemplate`` works."""
    z, dx, _ = _grid_fields(data)
    ny, nx = z.shape
    tobj = Template(scale, age, angle, nx, ny, dx, **kwargs)        # core.py:345
    amp, snr = plan.match_template_raster(tobj.template(), angle)   # core.py:346-367
    if hasattr(tobj, 'get_err_mask'):                               # core.py:369-371
        snr[np.asarray(tobj.get_err_mask(), dtype=bool)] = 0
    mask = np.asarray(tobj.get_window_limits(), dtype=bool)         # core.py:373-375
    amp[mask] = 0
    snr[mask] = 0
    return amp, snr


def _plugin_sweep(data, Template, scale, ages, ang_max, ang_min, order, **kwargs):
    """The reference's own loops (core.py:100-134, 180-188, 288-291) around
    ``_plugin_match_template``, folded with ``compare``'s exact semantics."""
    z, _, _ = _grid_fields(data)
    ny, nx = z.shape
    angles = P.search_angles(ang_min, ang_max)
    with _plan_for(data) as plan:
        def one(age, angle):
            amp, snr = _plugin_match_template(plan, data, Template, scale, age, angle, **kwargs)
            return amp, age, angle, snr
        if order == "angle_major":                                  # core.py:100-134
            best = np.zeros((4, ny, nx), dtype=np.float64)
            for angle in angles:
                for age in ages:
                    plan.compare_fold(best, *one(age, angle))
            return best
        outer = np.zeros((4, ny, nx), dtype=np.float64)
        for age in ages:                                            # core.py:288-291
            best = np.zeros((4, ny, nx), dtype=np.float64)
            for angle in angles:                                    # core.py:180-188
                plan.compare_fold(best, *one(age, angle))
            if len(ages) == 1:
                return best
            plan.compare_fold(outer, best[0], best[1], best[2], best[3])
        return outer


# ---------------------------------------------------------------------------
# the reference API
# ---------------------------------------------------------------------------
def match_template(data, Template, scale, age, angle, **kwargs):
    """Fit one (scale, age, angle) template to the directional curvature
    (core.py:297-377).  Returns ``(amp, age, angle, snr)`` with float64 planes.  Extra
    keyword arguments go to the template's constructor like the reference's (core.py:345);
    a class that takes any is served through its own ``template()``."""
    spec = None if kwargs else device_spec(Template)
    with _plan_for(data) as plan:
        if spec is None:
            amp, snr = _plugin_match_template(plan, data, Template, scale, age, angle, **kwargs)
        else:
            amp, snr = plan.match_template(spec, scale, age, angle)
    return amp, age, angle, snr


def _sweep(data, Template, scales, ages, ang_max, ang_min, order, **kwargs):
    """One search per entry of ``scales`` in a single device sweep (they share each
    orientation's curvature spectra); returns the list of (4, ny, nx) stacks."""
    spec = None if kwargs else device_spec(Template)
    if spec is None:
        return [_plugin_sweep(data, Template, s, ages, ang_max, ang_min, order, **kwargs) for s in scales]
    angles = P.search_angles(ang_min, ang_max)
    with _plan_for(data, states=len(scales)) as plan:
        a_rec, t_rec, age_of, angle_of = plan.build_sweep(spec, list(scales), ages, angles, order)
        plan.reset()
        plan.sweep(a_rec, t_rec)
        return [plan.finalize(age_of, angle_of, state=k) for k in range(len(scales))]


def calculate_best_fit_parameters(dem, Template, scale, age, ang_max=np.pi / 2,
                                  ang_min=-np.pi / 2, **kwargs):
    """Best amplitude / orientation / SNR at one age over a 1-degree orientation
    search (core.py:139-195).  Returns ndarray (4, ny, nx): [amp, age, angle, snr].
    Like the reference (core.py:182) further keyword arguments are accepted and not
    forwarded to the template."""
    return _sweep(dem, Template, [scale], [age], ang_max, ang_min, "age_major")[0]


def calculate_best_fit_parameters_serial(dem, Template, scale, ang_max=np.pi / 2,
                                         ang_min=-np.pi / 2, **kwargs):
    """Flat search over orientations x the 35 default ages (core.py:65-136); keyword
    arguments are forwarded to the template's constructor (core.py:116-121).
    Returns ``(best_amp, best_age, best_angle, best_snr)``."""
    out = _sweep(dem, Template, [scale], P.default_ages(), ang_max, ang_min, "angle_major", **kwargs)[0]
    return out[0], out[1], out[2], out[3]


def _match_args(kwargs):
    ang_max = kwargs.get('ang_max', np.pi / 2)
    ang_min = kwargs.get('ang_min', -np.pi / 2)
    if 'age' in kwargs:
        return [kwargs['age']], ang_max, ang_min, True
    # extension: ``ages=`` replaces the hard-coded 35-age grid of core.py:286
    ages = kwargs.get('ages', None)
    ages = P.default_ages() if ages is None else np.asarray(ages, dtype=np.float64)
    return ages, ang_max, ang_min, False


def match(data, Template, **kwargs):
    """core.py:266-294.  With ``age=`` one orientation search, returned as a stacked
    ndarray (4, ny, nx); without it the 35-age search, returned like
    ``compare`` as a tuple of four planes."""
    if 'age' in kwargs:
        return calculate_best_fit_parameters(data, Template, **kwargs)
    ages, ang_max, ang_min, _ = _match_args(kwargs)
    out = _sweep(data, Template, [kwargs['scale']], ages, ang_max, ang_min, "age_major")[0]
    return out[0], out[1], out[2], out[3]


def match_scales(data, Template, scales, **kwargs):
    """One ``match`` result per template scale -- the multi-scale product the reference
    publishes as one 4-band raster per scale (CHANGELOG.md:20-24), which its users obtain by
    calling ``match`` in a loop (core.py:266-294 per scale).  Here all scales run in ONE
    device sweep: an orientation's curvature spectra are built once and every scale folds
    into its own best state.  Returns ``{scale: result}`` with ``result`` exactly what
    ``match(..., scale=scale)`` returns."""
    scales = list(scales)
    ages, ang_max, ang_min, single = _match_args(kwargs)
    res = _sweep(data, Template, scales, ages, ang_max, ang_min, "age_major")
    return {s: (r if single else (r[0], r[1], r[2], r[3])) for s, r in zip(scales, res)}


def compare(results, ny, nx):
    """Per-pixel best-SNR select over an iterable of ``(amp, age, angle, snr)``
    (core.py:198-243), with the reference's exact semantics (strict compares, an
    exact tie zeroes the pixel), folded on the device in float64."""
    best = np.zeros((4, ny, nx), dtype=np.float64)
    with Plan(ny, nx, 1.0, 1.0) as plan:
        for this_amp, this_age, this_angle, this_snr in results:
            plan.compare_fold(best, this_amp, this_age, this_angle, this_snr)
    return best[0], best[1], best[2], best[3]
