"""Template plugin surface (host-side mirror of scarplet/WindowedTemplate.py).

Same constructors, attributes and method names as the reference classes, so code
written against ``scarplet.WindowedTemplate`` keeps working::

    Template(scale, age, angle, nx, ny, de)        # core.py:345
    .template()            -> (ny, nx) float64     # evaluated by the CUDA generator
    .get_window_limits()   -> (ny, nx) bool edge mask
    .get_err_mask()        -> (ny, nx) bool        # upper-break templates only
    .get_coordinates() / .get_mask()

Inside a sweep none of these methods run: each built-in class carries a
``_sb_spec`` (``params.DeviceSpec``) that selects the on-device generator, and the
kernels evaluate template, window and masks per (scale, age, angle) in registers.
"""
import numpy as np

from . import params as P


class WindowedTemplate(object):
    """Base of the plugin surface (WindowedTemplate.py:12-84)."""

    _sb_spec = None

    def _axes(self):
        return P.axis_vectors(self.nx, self.ny, self.de)

    def get_coordinates(self):
        """Rotated coordinates of every pixel (WindowedTemplate.py:49-59)."""
        xs, ys = self._axes()
        gx, gy = np.meshgrid(xs, ys)
        ca, sa = np.cos(self.alpha), np.sin(self.alpha)
        return gx * ca + gy * sa, -gx * sa + gy * ca

    def get_mask(self):
        """Curvature window (WindowedTemplate.py:61-64)."""
        xr, yr = self.get_coordinates()
        return (abs(xr) < self.c) & (abs(yr) < self.d)

    def get_window_limits(self):
        """Edge mask (WindowedTemplate.py:66-84), built from the same index rectangle
        the kernels use."""
        xs, ys = self._axes()
        i_lo, i_hi, j_lo, j_hi = P.window_rectangle(xs, ys, self.alpha, self.c, self.d)
        mask = np.ones((self.ny, self.nx), dtype=bool)
        if i_hi >= i_lo and j_hi >= j_lo:
            mask[i_lo:i_hi + 1, j_lo:j_hi + 1] = False
        return mask

    def _age_parameter(self):
        raise NotImplementedError

    def template(self):
        """Windowed template raster, rendered on the GPU in float64."""
        from .engine import Plan
        with Plan(self.ny, self.nx, self.de, self.de) as plan:
            return plan.render_template(self._sb_spec, self.d, self._age_parameter(),
                                        -self.alpha)


class Scarp(WindowedTemplate):
    """Diffusion-degraded vertical scarp (WindowedTemplate.py:87-183)."""

    _sb_spec = P.DeviceSpec(P.KIND_SCARP)

    def __init__(self, d, kt, alpha, nx, ny, de):
        self.d = d
        self.kt = kt
        self.alpha = -alpha                      # WindowedTemplate.py:151
        self.nx = nx
        self.ny = ny
        self.de = de
        self.c = P.scarp_halfwidth(kt)           # WindowedTemplate.py:156-157

    def _age_parameter(self):
        return self.kt

    def template_numexpr(self):
        """The reference's numexpr twin (WindowedTemplate.py:185-215) computes the same
        raster with a modern numexpr."""
        return self.template()


class RightFacingUpperBreakScarp(Scarp):
    """Negated scarp template + lower-half SNR mask (WindowedTemplate.py:218-267)."""

    _sb_spec = P.DeviceSpec(P.KIND_SCARP, sign=-1.0, errmode=P.ERRMASK_XR_LE0)

    def get_err_mask(self):
        xr, _ = self.get_coordinates()
        return xr <= 0


class LeftFacingUpperBreakScarp(Scarp):
    """WindowedTemplate.py:270-304"""

    _sb_spec = P.DeviceSpec(P.KIND_SCARP, errmode=P.ERRMASK_XR_GE0)

    def get_err_mask(self):
        xr, _ = self.get_coordinates()
        return xr >= 0


class Ricker(WindowedTemplate):
    """Ricker wavelet across the profile (WindowedTemplate.py:434-520)."""

    _sb_spec = P.DeviceSpec(P.KIND_RICKER, edge_mask=False)

    def __init__(self, d, f, alpha, nx, ny, de):
        self.d = d
        self.f = f
        self.alpha = -alpha                      # WindowedTemplate.py:489
        self.nx = nx
        self.ny = ny
        self.c = nx                              # WindowedTemplate.py:491
        self.de = de

    def _age_parameter(self):
        return self.f

    def get_window_limits(self):
        return np.zeros((self.ny, self.nx), dtype=bool)   # WindowedTemplate.py:494-495


class Channel(Ricker):
    """Alias used for fluvial channels (WindowedTemplate.py:523-525)."""
    pass


_BUILTIN_BY_NAME = {cls.__name__: cls for cls in
                    (Scarp, RightFacingUpperBreakScarp, LeftFacingUpperBreakScarp, Ricker, Channel)}


# the plugin surface a subclass may override (core.py:345-346, 369-375)
_SURFACE = ("__init__", "template", "template_numexpr", "get_coordinates", "get_mask",
            "get_window_limits", "get_err_mask")


def device_spec(Template):
    """Device generator for a template class, or ``None``.

    Only a class that IS one of the built-ins keeps the on-device generator: this package's
    classes, the reference's own built-ins (matched by name when they come from
    ``scarplet.WindowedTemplate``, so ``sl.match(data, scarplet.WindowedTemplate.Scarp)`` style
    call sites keep working), and subclasses of this package's built-ins that leave the whole
    plugin surface (constructor, ``template()``, masks, coordinates) untouched or declare their
    own ``_sb_spec``.  A subclass that overrides any of it -- the reference's
    ``ShiftedTemplateMixin`` pattern (WindowedTemplate.py:307-431) rebuilt on these classes, a
    rescaled ``template()`` ... -- is a plugin class: ``None`` is returned and it is served
    through its own methods (core._plugin_*), never silently replaced by its base."""
    if not isinstance(Template, type):
        return None
    if Template in _BUILTIN_BY_NAME.values():
        return Template._sb_spec
    mod = getattr(Template, "__module__", "")
    name = getattr(Template, "__name__", "")
    if mod.startswith("scarplet.") and name in _BUILTIN_BY_NAME:
        return _BUILTIN_BY_NAME[name]._sb_spec
    for cls in Template.__mro__:
        if cls in _BUILTIN_BY_NAME.values():
            return cls._sb_spec               # nothing overridden on the way down to a built-in
        if "_sb_spec" in cls.__dict__ and cls.__dict__["_sb_spec"] is not None:
            return cls.__dict__["_sb_spec"]   # the subclass names its generator itself
        if any(attr in cls.__dict__ for attr in _SURFACE):
            return None
    return None        # a plugin class: core.py renders it on the host (generic path)
