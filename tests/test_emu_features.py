"""Round-2 features through the CPU emulator build of the kernels (tests/emu) and on the
host: row-slab plans (spatial sharding), mixed FFT tile lengths, get_err_mask on the fast fit
kernel, subclass / keyword handling of the plugin surface, the curvature noise level, nodata
filling and the GeoTIFF writer.  Index logic only -- GPU parity is tests/test_gpu_parity.py."""
import os

import numpy as np
import pytest

from oracle import scarplet_oracle as O
from tests.parity import assert_parity, stack_report


def _search(plan, z, spec, scale, ages, angles, order="age_major"):
    plan.set_dem(z)
    a, t, age_of, angle_of = plan.build_sweep(spec, scale, ages, angles, order)
    plan.reset()
    plan.sweep(a, t)
    return plan.finalize(age_of, angle_of)


@pytest.mark.parametrize("shape,bands,tmpl", [
    ((150, 170), 3, "Scarp"),       # padded domain, bands shorter than a tile
    ((128, 128), 2, "Scarp"),       # power-of-two raster: the whole-raster plan is periodic, the bands are not
    ((96, 100), 2, "Channel"),      # no edge mask: the periodic wrap-around of the raster is live output
])
def test_row_slabs_equal_whole_raster(emu_lib, shape, bands, tmpl):
    """A raster cut into row bands with a halo (one band per rank, BASELINE config 5) gives
    the whole-raster result: template centring, edge masks and the circular wrap-around are
    evaluated in full-raster coordinates, a band holds only its own DEM rows + halo."""
    from scarplet_b200 import params as P, distributed as D, templates as T
    from scarplet_b200.engine import Plan
    from scarplet_b200.synth import synthetic_dem
    ny, nx = shape
    z = synthetic_dem(ny, seed=11, nx=nx)
    cls = getattr(T, tmpl)
    spec = cls._sb_spec
    scale, ages = (10, [2.0, 20.0]) if tmpl == "Scarp" else (5, [0.2])
    angles = P.search_angles(-np.pi / 2, np.pi / 2)[2::25]
    with Plan(ny, nx, 1.0, 1.0) as plan:
        whole = _search(plan, z, spec, scale, ages, angles)
        s_all, n_all = plan.curv_stats()
    halo = D.slab_halo(spec, scale, ages, angles, nx, ny, 1.0)
    parts, stats = [], []
    for r in range(bands):
        lo, hi = D.shard_bounds(ny, bands, r)
        with Plan(ny, nx, 1.0, 1.0, slab=(lo, hi, halo)) as plan:
            r0, nrows = plan.dem_rows()
            assert nrows <= hi - lo + 2 * halo
            plan.set_dem(z)                               # the band is cut out of the whole raster
            stats.append(plan.curv_stats())
            plan.set_curv_stats(s_all, n_all)             # what share_dem_stats distributes
            a, t, age_of, angle_of = plan.build_sweep(spec, scale, ages, angles)
            plan.reset()
            plan.sweep(a, t)
            out = plan.finalize(age_of, angle_of)
            assert out.shape == (4, hi - lo, nx)
            assert plan.device_bytes > 0
            parts.append(out)
    got = np.concatenate(parts, axis=1)
    assert np.isclose(sum(s for s, _ in stats), s_all, rtol=1e-12) and sum(n for _, n in stats) == n_all
    ref = np.stack(O.compare((O.match_template(z, 1.0, 1.0, O.SCARP if tmpl == "Scarp" else O.RICKER, scale, age, ang)
                              for age in ages for ang in angles), ny, nx))
    for res in (whole, got):
        rep = stack_report(res, ref, odd_template=(tmpl == "Scarp"))
        assert rep["mask_mismatch_unexplained"] == 0 and rep["index_agreement"] >= 0.999, rep
        assert rep["snr_rel_max_strong"] <= 1e-4, rep
    # the bands agree with the whole-raster plan to rounding (different FFT domains)
    rep = stack_report(got, whole, odd_template=(tmpl == "Scarp"))
    assert rep["mask_equal"] and rep["index_agreement"] >= 0.999 and rep["snr_rel_max_strong"] < 2e-5, rep


def test_slab_halo_too_small_is_an_error(emu_lib):
    from scarplet_b200 import params as P, templates as T
    from scarplet_b200._lib import SbError
    from scarplet_b200.engine import Plan
    from scarplet_b200.synth import synthetic_dem
    z = synthetic_dem(150, seed=1, nx=128)
    angles = P.search_angles(-0.1, 0.1)
    with Plan(150, 128, 1.0, 1.0, slab=(40, 90, 3)) as plan:
        plan.set_dem(z)
        a, t, _, _ = plan.build_sweep(T.Scarp._sb_spec, 10, [2.0], angles)
        plan.reset()
        with pytest.raises(SbError, match="halo too small"):
            plan.sweep(a, t)


def test_nan_reaches_every_band(emu_lib):
    """dem.py:105 + the reference's full-raster fft2: a NaN anywhere poisons every un-masked
    pixel; a band that does not hold the cell learns about it through the shared statistics."""
    from scarplet_b200 import params as P, distributed as D, templates as T
    from scarplet_b200.engine import Plan
    from scarplet_b200.synth import synthetic_dem
    ny, nx = 120, 96
    z = synthetic_dem(ny, seed=5, nx=nx)
    z[10, 40] = np.nan
    spec, angles = T.Scarp._sb_spec, P.search_angles(-np.pi / 2, np.pi / 2)[::30]
    ref = O.compare((O.match_template(z, 1.0, 1.0, O.SCARP, 8, 3.0, a) for a in angles), ny, nx)
    halo = D.slab_halo(spec, 8, [3.0], angles, nx, ny, 1.0)
    stats, plans = [], []
    for r in range(2):
        lo, hi = D.shard_bounds(ny, 2, r)
        plan = Plan(ny, nx, 1.0, 1.0, slab=(lo, hi, halo))
        plan.set_dem(z)
        stats.append(plan.curv_stats())
        plans.append(plan)
    assert not np.isfinite(stats[0][0]) and np.isfinite(stats[1][0])
    parts = []
    for plan in plans:
        plan.set_curv_stats(sum(s for s, _ in stats), sum(n for _, n in stats))
        a, t, age_of, angle_of = plan.build_sweep(spec, 8, [3.0], angles)
        plan.reset()
        plan.sweep(a, t)
        parts.append(plan.finalize(age_of, angle_of))
        plan.close()
    got = np.concatenate(parts, axis=1)
    for k in (0, 3):
        assert np.array_equal(np.isnan(got[k]), np.isnan(ref[k]))
        assert np.array_equal(got[k] == 0, ref[k] == 0)


def test_mixed_tile_lengths(emu_lib):
    """Tiles of different FFT lengths along an axis (least summed length) give the result of
    equal tiles and of the oracle."""
    from scarplet_b200 import params as P, templates as T
    from scarplet_b200.engine import Plan
    from scarplet_b200.synth import synthetic_dem
    ny, nx = 300, 460
    z = synthetic_dem(ny, seed=3, nx=nx)
    angles = P.search_angles(-np.pi / 2, np.pi / 2)[5::40]
    outs, areas = {}, {}
    for mixed in (1, 0):
        with Plan(ny, nx, 1.0, 1.0, max_fft=256) as plan:
            plan.set_option("mixed_tiles", mixed)
            outs[mixed] = _search(plan, z, T.Scarp._sb_spec, 12, [2.0, 30.0], angles)
            areas[mixed] = plan.fft_area
            geo = plan.last_geometry()
            assert geo["tiles_x"] >= 2 and geo["tiles_y"] >= 2
    assert areas[1] < areas[0]                            # e.g. 256 + 256 + 128 columns instead of 3 x 256
    ref = np.stack(O.compare((O.match_template(z, 1.0, 1.0, O.SCARP, 12, age, ang)
                              for age in (2.0, 30.0) for ang in angles), ny, nx))
    for mixed in (1, 0):
        rep = stack_report(outs[mixed], ref)
        assert rep["mask_mismatch_unexplained"] == 0 and rep["index_agreement"] >= 0.999, rep
        assert rep["snr_rel_max_strong"] <= 1e-4 and rep["amp_rel_max_strong"] <= 1e-4, rep


@pytest.mark.parametrize("cls,kind", [("LeftFacingUpperBreakScarp", O.LEFT_UPPER),
                                      ("RightFacingUpperBreakScarp", O.RIGHT_UPPER)])
@pytest.mark.parametrize("shape", [(128, 128), (90, 150)])
def test_err_mask_templates_on_the_fast_path(emu_lib, cls, kind, shape):
    """get_err_mask (core.py:369-371) inside a search narrows each row's candidate columns: the
    pipelined fit kernel takes the float64-exact column ranges from k_err_cross and must give
    the result of the simple kernels (which evaluate xr per pixel) and of the oracle."""
    from scarplet_b200 import params as P, templates as T
    from scarplet_b200.engine import Plan
    from scarplet_b200.synth import synthetic_dem
    ny, nx = shape
    z = synthetic_dem(ny, seed=17, nx=nx)
    angles = P.search_angles(-np.pi / 2, np.pi / 2)[::15]
    spec = getattr(T, cls)._sb_spec
    outs = {}
    for fast in (1, 0):
        with Plan(ny, nx, 1.0, 1.0) as plan:
            plan.set_option("fast", fast)
            outs[fast] = _search(plan, z, spec, 9, [3.0, 12.0], angles)
    ref = np.stack(O.compare((O.match_template(z, 1.0, 1.0, kind, 9, age, ang)
                              for age in (3.0, 12.0) for ang in angles), ny, nx))
    assert np.array_equal(outs[1][3] > 0, outs[0][3] > 0)
    for fast in (1, 0):
        rep = stack_report(outs[fast], ref)
        assert rep["mask_mismatch_unexplained"] == 0 and rep["index_agreement"] >= 0.999, rep
        assert rep["snr_rel_max_strong"] <= 1e-4 and rep["amp_rel_max_strong"] <= 1e-4, rep


def test_overriding_subclass_is_a_plugin(emu_lib):
    """A subclass of a built-in that overrides part of the plugin surface must be served through
    its own methods, never silently through its base's device generator (the reference's
    ShiftedTemplateMixin pattern, WindowedTemplate.py:307-431)."""
    import scarplet_b200 as sl
    from scarplet_b200 import templates as T
    from scarplet_b200.synth import synthetic_dem

    class Half(T.Scarp):
        def template(self):
            return 0.5 * super().template()

    class Renamed(T.Scarp):
        pass

    class OwnSpec(T.Scarp):
        _sb_spec = T.RightFacingUpperBreakScarp._sb_spec

        def get_err_mask(self):
            xr, _ = self.get_coordinates()
            return xr <= 0

    assert T.device_spec(Half) is None
    assert T.device_spec(Renamed) is T.Scarp._sb_spec
    assert T.device_spec(OwnSpec) is T.RightFacingUpperBreakScarp._sb_spec
    assert T.device_spec(T.Channel) is T.Ricker._sb_spec
    z = synthetic_dem(64, seed=3, nx=80)
    grid = sl.DEMGrid(z, 1.0)
    amp_h, _, _, snr_h = sl.match_template(grid, Half, 8, 3.0, 0.4)
    amp, _, _, snr = sl.match_template(grid, T.Scarp, 8, 3.0, 0.4)
    v = snr > 0
    assert v.sum() > 500 and np.array_equal(snr_h > 0, v)
    # half the template => twice the amplitude, same SNR (core.py:360-367)
    assert np.allclose(amp_h[v], 2 * amp[v], rtol=2e-5, atol=4e-6 * np.abs(amp[v]).max())
    res_h = sl.calculate_best_fit_parameters(grid, Half, 8, 3.0, ang_max=0.1, ang_min=-0.1)
    res = sl.calculate_best_fit_parameters(grid, T.Scarp, 8, 3.0, ang_max=0.1, ang_min=-0.1)
    w = res[3] > 0
    assert np.allclose(res_h[0][w], 2 * res[0][w], rtol=1e-4, atol=1e-6 * np.abs(res[0][w]).max())
    sl.release()


def test_kwargs_reach_the_template_constructor(emu_lib):
    """core.py:116-121, 345: calculate_best_fit_parameters_serial and match_template forward
    keyword arguments to the template; calculate_best_fit_parameters accepts and drops them
    (core.py:182)."""
    import scarplet_b200 as sl
    from scarplet_b200 import templates as T
    from scarplet_b200.synth import synthetic_dem
    seen = []

    class Shifted(T.Scarp):
        def __init__(self, d, kt, alpha, nx, ny, de, dx=0, dy=0):
            T.Scarp.__init__(self, d, kt, alpha, nx, ny, de)
            seen.append((dx, dy))
            self.shift = (int(dy), int(dx))

        def template(self):
            return np.roll(T.Scarp.template(self), self.shift, axis=(0, 1))

    z = synthetic_dem(48, seed=3, nx=64)
    grid = sl.DEMGrid(z, 1.0)
    sl.match_template(grid, Shifted, 6, 3.0, 0.2, dx=2, dy=1)
    assert seen[-1] == (2, 1)
    n0 = len(seen)
    out = sl.calculate_best_fit_parameters_serial(grid, Shifted, 6, ang_max=0.01, ang_min=-0.01, dx=3, dy=0)
    assert len(out) == 4 and len(seen) - n0 == 2 * 35 and set(seen[n0:]) == {(3, 0)}    # 2 orientations x 35 ages
    ref = O.calculate_best_fit_parameters_serial_plugin(z, 1.0, 1.0, Shifted, 6, ang_max=0.01, ang_min=-0.01,
                                                        dx=3, dy=0)
    rep = stack_report(np.stack(out), np.stack(ref))
    assert rep["mask_mismatch_unexplained"] == 0 and rep["index_agreement"] >= 0.999, rep
    with pytest.raises(TypeError):                      # the reference's Scarp takes no keywords either
        sl.match_template(grid, T.Scarp, 6, 3.0, 0.2, dx=2)
    res = sl.calculate_best_fit_parameters(grid, T.Scarp, 6, 3.0, ang_max=0.01, ang_min=-0.01, dx=5)
    assert res.shape == (4, 48, 64)
    sl.release()


def test_curvature_noise_level(emu_lib):
    """dem.py:152-179 against the oracle's restatement (scipy.ndimage.gaussian_filter per
    direction), with and without nodata cells."""
    import scarplet_b200 as sl
    from scarplet_b200.synth import synthetic_dem
    z = synthetic_dem(70, seed=9, nx=90)
    for with_nan in (False, True):
        if with_nan:
            z = z.copy()
            z[20, 30] = np.nan
            z[55:57, 80] = np.nan
        keep = z.copy()
        angles, mean, sd = sl.DEMGrid(z, 2.0, 2.0)._estimate_curvature_noiselevel(sigma=6)
        r_angles, r_mean, r_sd = O.estimate_curvature_noiselevel(z, 2.0, 2.0, sigma=6)
        assert np.array_equal(z, keep, equal_nan=True)
        assert len(mean) == 180 and isinstance(mean, list) and np.array_equal(angles, r_angles)
        scale = np.max(np.abs(r_sd))
        assert np.allclose(sd, r_sd, rtol=1e-9, atol=0)
        assert np.allclose(mean, r_mean, rtol=0, atol=1e-11 * scale)


def test_fill_nodata(emu_lib):
    """dem.py:388-414: in place, loops until nothing is left, flags the grid as interpolated;
    the device pass equals the oracle's restatement of the same ray search."""
    import scarplet_b200 as sl
    from scarplet_b200.synth import synthetic_dem
    z = synthetic_dem(60, seed=4, nx=75)
    holes = z.copy()
    holes[10:14, 20:29] = np.nan
    holes[40, :] = np.nan
    holes[0, 0] = np.nan
    grid = sl.DEMGrid(holes, 1.0)
    grid._fill_nodata()
    assert grid.is_interpolated and not np.isnan(grid._griddata).any()
    assert np.array_equal(grid.nodata_mask, np.isnan(holes))
    assert np.array_equal(grid._griddata[~np.isnan(holes)], holes[~np.isnan(holes)])
    ref = O.fill_nodata(holes)
    assert np.allclose(grid._griddata, ref, rtol=1e-13, atol=0)
    # interpolated values stay near the surface they replace
    assert np.abs(grid._griddata - z)[np.isnan(holes)].max() < 10.0          # relief: 30 m
    g2 = sl.DEMGrid(np.where(np.isnan(holes), -9999.0, holes), 1.0)
    g2.nodata_value = -9999.0
    g2._fill_nodata()
    assert np.allclose(g2._griddata, ref, rtol=1e-13, atol=0)


def test_geotiff_writer(tmp_path):
    """dem.py:291-306 and the 4-band result product (CHANGELOG.md:20) without GDAL: the files
    parse back (own reader and PIL), band order amplitude, age, orientation, SNR."""
    from PIL import Image
    import scarplet_b200 as sl
    from scarplet_b200 import geotiff
    rng = np.random.default_rng(0)
    res = rng.standard_normal((4, 37, 53))
    georef = sl.dem.GeorefInfo(2.0, -2.0, 53, 37)
    georef.geo_transform = (500000.0, 2.0, 0.0, 4100000.0, 0.0, -2.0)
    path = str(tmp_path / "results.tif")
    n = sl.save_results(path, res, georef, epsg=32610)
    assert n == os.path.getsize(path)
    bands, tags = geotiff.read_geotiff(path)
    assert bands.shape == (4, 37, 53) and np.array_equal(bands, res.astype(np.float32))
    assert geotiff.BAND_ORDER == ("amplitude", "age", "orientation", "snr")
    assert tags[277] == (4,) and tags[339] == (3, 3, 3, 3) and tags[258] == (32, 32, 32, 32)
    assert tags[33550] == (2.0, 2.0, 0.0) and tags[33922] == (0.0, 0.0, 0.0, 500000.0, 4100000.0, 0.0)
    keys = tags[34735]
    assert keys[:4] == (1, 1, 0, 3) and (3072, 0, 1, 32610) in [keys[4 + 4 * i: 8 + 4 * i] for i in range(3)]
    # single band (DEMGrid.save): PIL reads it as mode F with the same values and geo tags
    grid = sl.DEMGrid(rng.standard_normal((20, 31)) * 100, 2.0, -2.0)
    grid._georef_info.geo_transform = georef.geo_transform
    p1 = str(tmp_path / "dem.tif")
    grid.save(p1)
    im = Image.open(p1)
    assert im.mode == "F" and im.size == (31, 20)
    assert np.array_equal(np.asarray(im), grid._griddata.astype(np.float32))
    assert tuple(im.tag_v2[33550]) == (2.0, 2.0, 0.0)
    # many strips and BigTIFF layout
    big = rng.standard_normal((2, 300, 40))
    p2 = str(tmp_path / "strips.tif")
    geotiff.write_geotiff(p2, big, rows_per_strip=7, nodata=-9999.0)
    b2, t2 = geotiff.read_geotiff(p2)
    assert np.array_equal(b2, big.astype(np.float32)) and len(t2[273]) == 43
    assert b"".join(t2[42113]).rstrip(b"\0") == b"-9999.0"
