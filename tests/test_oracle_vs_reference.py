"""Validates the NumPy oracle against the UNMODIFIED reference imported from
/root/reference (build container only; skipped on the GPU box)."""
import numpy as np
import pytest

from oracle import ref_import, scarplet_oracle as O

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref():
    sl = ref_import.import_reference()
    from scarplet import WindowedTemplate as WT
    return sl, WT


def _dem(ny, nx, seed):
    rng = np.random.default_rng(seed)
    z = np.cumsum(np.cumsum(rng.standard_normal((ny, nx)), 0), 1) * 0.01
    return z.astype(np.float32).astype(np.float64)


@pytest.mark.parametrize("cls,kind,scale,age", [
    ("Scarp", O.SCARP, 9, 2.5), ("Channel", O.RICKER, 5, 0.2), ("Ricker", O.RICKER, 4, 0.05),
    ("RightFacingUpperBreakScarp", O.RIGHT_UPPER, 9, 2.5), ("LeftFacingUpperBreakScarp", O.LEFT_UPPER, 7, 6.0)])
@pytest.mark.parametrize("shape,de", [((60, 70), 1.0), ((57, 81), 2.0)])
def test_match_template_identical(ref, cls, kind, scale, age, shape, de):
    sl, WT = ref
    z = _dem(shape[0], shape[1], 3)
    for angle in (-np.pi / 2, -0.7, 0.0, 0.4, np.pi / 2):
        grid = ref_import.make_grid(z, de, de)
        amp, _, _, snr = sl.match_template(grid, getattr(WT, cls), scale * de, age, angle)
        oamp, _, _, osnr = O.match_template(z, de, de, kind, scale * de, age, angle)
        assert np.array_equal(amp, oamp) and np.array_equal(snr, osnr)


def test_templates_and_masks_identical(ref):
    sl, WT = ref
    for nx, ny, de in ((64, 50, 1.0), (51, 77, 0.5)):
        for angle in (-1.3, 0.0, 0.9):
            t = WT.Scarp(12, 4.0, angle, nx, ny, de)
            assert np.array_equal(t.template(), O.template_array(O.SCARP, 12, 4.0, angle, nx, ny, de))
            assert np.array_equal(t.get_window_limits(), O.window_limits(O.SCARP, nx, ny, de, -angle, t.c, 12))
            r = WT.Ricker(6, 0.1, angle, nx, ny, de)
            assert np.array_equal(r.template(), O.template_array(O.RICKER, 6, 0.1, angle, nx, ny, de))
            e = WT.LeftFacingUpperBreakScarp(12, 4.0, angle, nx, ny, de)
            assert np.array_equal(e.get_err_mask(), O.err_mask(O.LEFT_UPPER, angle, nx, ny, de))


def test_search_and_age_sweep_identical(ref):
    sl, WT = ref
    z = _dem(48, 56, 5)
    grid = ref_import.make_grid(z, 1.0, 1.0)
    res = sl.calculate_best_fit_parameters(grid, WT.Scarp, 8, 3.0)
    assert np.array_equal(res, O.calculate_best_fit_parameters(z, 1.0, 1.0, O.SCARP, 8, 3.0, processes=2))
    ser = sl.calculate_best_fit_parameters_serial(grid, WT.Scarp, 8, ang_max=0.05, ang_min=-0.05)
    oser = O.calculate_best_fit_parameters_serial(z, 1.0, 1.0, O.SCARP, 8, ang_max=0.05, ang_min=-0.05)
    for a, b in zip(ser, oser):
        assert np.array_equal(a, b)


def test_plugin_template_identical(ref):
    """A user-defined template class through the unmodified reference and through the oracle's
    restatement of the plugin surface (core.py:345-375)."""
    sl, WT = ref
    from plugin_templates import Ridge
    z = _dem(64, 72, 9)
    grid = ref_import.make_grid(z, 1.0, 1.0)
    for angle in (-1.1, 0.0, 0.6):
        amp, _, _, snr = sl.match_template(grid, Ridge, 7, 1.5, angle)
        oamp, _, _, osnr = O.match_template_plugin(z, 1.0, 1.0, Ridge, 7, 1.5, angle)
        assert np.array_equal(amp, oamp) and np.array_equal(snr, osnr)
    res = sl.calculate_best_fit_parameters(grid, Ridge, 7, 1.5, ang_max=0.1, ang_min=-0.1)
    ores = O.calculate_best_fit_parameters_plugin(z, 1.0, 1.0, Ridge, 7, 1.5, ang_max=0.1, ang_min=-0.1)
    assert np.array_equal(res, ores)


def test_serial_sweep_forwards_kwargs_identical(ref):
    """core.py:116-121: keyword arguments reach the template's constructor in the serial sweep."""
    sl, WT = ref
    z = _dem(40, 48, 2)
    grid = ref_import.make_grid(z, 1.0, 1.0)
    ser = sl.calculate_best_fit_parameters_serial(grid, WT.ShiftedLeftFacingUpperBreakScarp, 6,
                                                  ang_max=0.01, ang_min=-0.01, dx=3, dy=2)
    oser = O.calculate_best_fit_parameters_serial_plugin(z, 1.0, 1.0, WT.ShiftedLeftFacingUpperBreakScarp, 6,
                                                         ang_max=0.01, ang_min=-0.01, dx=3, dy=2)
    for a, b in zip(ser, oser):
        assert np.array_equal(a, b)


def test_curvature_noiselevel_identical(ref):
    """dem.py:152-179 of the unmodified reference (sigma 100 is hard-coded there, so the raster
    is small and the filter runs on its reflecting boundary throughout)."""
    z = _dem(30, 36, 4)
    z[7, 9] = np.nan
    grid = ref_import.make_grid(z.copy(), 2.0, 2.0)
    angles, mean, sd = grid._estimate_curvature_noiselevel()
    oangles, omean, osd = O.estimate_curvature_noiselevel(z, 2.0, 2.0, sigma=100)
    assert np.array_equal(angles, oangles)
    assert np.array_equal(mean, omean, equal_nan=True) and np.array_equal(sd, osd, equal_nan=True)
