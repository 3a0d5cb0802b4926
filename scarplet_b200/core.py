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
"""
import numpy as np

from . import params as P
from .engine import Plan
from .templates import device_spec


def _grid_fields(data):
    z = data._griddata
    return z, data._georef_info.dx, data._georef_info.dy


def _plan_for(data, **plan_kwargs):
    z, dx, dy = _grid_fields(data)
    ny, nx = z.shape
    plan = Plan(ny, nx, dx, dy, **plan_kwargs)
    plan.set_dem(z)
    return plan


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


def match_template(data, Template, scale, age, angle, **kwargs):
    """Fit one (scale, age, angle) template to the directional curvature
    (core.py:297-377).  Returns ``(amp, age, angle, snr)`` with float64 planes."""
    spec = device_spec(Template)
    with _plan_for(data) as plan:
        if spec is None:
            amp, snr = _plugin_match_template(plan, data, Template, scale, age, angle, **kwargs)
        else:
            amp, snr = plan.match_template(spec, scale, age, angle)
    return amp, age, angle, snr


def _sweep(data, Template, scale, ages, ang_max, ang_min, order, plan=None):
    spec = device_spec(Template)
    if spec is None:
        return _plugin_sweep(data, Template, scale, ages, ang_max, ang_min, order)
    angles = P.search_angles(ang_min, ang_max)
    own = plan is None
    if own:
        plan = _plan_for(data)
    try:
        a_rec, t_rec, age_of, angle_of = plan.build_sweep(spec, scale, ages, angles, order)
        plan.reset()
        plan.sweep(a_rec, t_rec)
        return plan.finalize(age_of, angle_of)
    finally:
        if own:
            plan.close()


def calculate_best_fit_parameters(dem, Template, scale, age, ang_max=np.pi / 2,
                                  ang_min=-np.pi / 2, **kwargs):
    """Best amplitude / orientation / SNR at one age over a 1-degree orientation
    search (core.py:139-195).  Returns ndarray (4, ny, nx): [amp, age, angle, snr]."""
    return _sweep(dem, Template, scale, [age], ang_max, ang_min, "age_major")


def calculate_best_fit_parameters_serial(dem, Template, scale, ang_max=np.pi / 2,
                                         ang_min=-np.pi / 2, **kwargs):
    """Flat search over orientations x the 35 default ages (core.py:65-136).
    Returns ``(best_amp, best_age, best_angle, best_snr)``."""
    out = _sweep(dem, Template, scale, P.default_ages(), ang_max, ang_min, "angle_major")
    return out[0], out[1], out[2], out[3]


def match(data, Template, **kwargs):
    """core.py:266-294.  With ``age=`` one orientation search, returned as a stacked
    ndarray (4, ny, nx); without it the 35-age search, returned like
    ``compare`` as a tuple of four planes."""
    if 'age' in kwargs:
        return calculate_best_fit_parameters(data, Template, **kwargs)
    scale = kwargs['scale']
    ang_max = kwargs.get('ang_max', np.pi / 2)
    ang_min = kwargs.get('ang_min', -np.pi / 2)
    # extension: ``ages=`` replaces the hard-coded 35-age grid of core.py:286
    ages = kwargs.get('ages', None)
    ages = P.default_ages() if ages is None else np.asarray(ages, dtype=np.float64)
    out = _sweep(data, Template, scale, ages, ang_max, ang_min, "age_major")
    return out[0], out[1], out[2], out[3]


def match_scales(data, Template, scales, **kwargs):
    """One ``match`` result per template scale -- the multi-scale product the reference
    publishes as one 4-band raster per scale (CHANGELOG.md:20-24), which its users obtain by
    calling ``match`` in a loop (core.py:266-294 per scale).  The DEM is uploaded and its
    second differences are built once; every scale is its own search with its own best state.
    Returns ``{scale: result}`` with ``result`` exactly what ``match(..., scale=scale)`` returns."""
    spec = device_spec(Template)
    out = {}
    if spec is None:
        for scale in scales:
            out[scale] = match(data, Template, scale=scale, **kwargs)
        return out
    ang_max = kwargs.get('ang_max', np.pi / 2)
    ang_min = kwargs.get('ang_min', -np.pi / 2)
    with _plan_for(data) as plan:
        for scale in scales:
            if 'age' in kwargs:
                out[scale] = _sweep(data, Template, scale, [kwargs['age']], ang_max, ang_min,
                                    "age_major", plan=plan)
            else:
                ages = kwargs.get('ages', None)
                ages = P.default_ages() if ages is None else np.asarray(ages, dtype=np.float64)
                res = _sweep(data, Template, scale, ages, ang_max, ang_min, "age_major", plan=plan)
                out[scale] = (res[0], res[1], res[2], res[3])
    return out


def compare(results, ny, nx):
    """Per-pixel best-SNR select over an iterable of ``(amp, age, angle, snr)``
    (core.py:198-243), with the reference's exact semantics (strict compares, an
    exact tie zeroes the pixel), folded on the device in float64."""
    best = np.zeros((4, ny, nx), dtype=np.float64)
    with Plan(ny, nx, 1.0, 1.0) as plan:
        for this_amp, this_age, this_angle, this_snr in results:
            plan.compare_fold(best, this_amp, this_age, this_angle, this_snr)
    return best[0], best[1], best[2], best[3]
