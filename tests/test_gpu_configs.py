"""Parity at the BASELINE.json configurations and for the entry points
the single-shape tests do not reach, with the north_star tolerances enforced as BOUNDS
(tests/parity.py): max relative error of amplitude / SNR <= 1e-4 on every pixel at or above
the median reference SNR, <= 5e-5 of all valid pixels above 1e-4, orientation / age identical
on >= 99.9 % of the valid pixels with EVERY disagreement explained, masks exact.

Seeded, non-mirrored DEMs (scarplet_b200.synth.synthetic_dem = SURVEY 8d's generator).  Where
the raster is too large for the oracle, the oracle runs on a crop: the template support is
compact, so a crop with the raster's parity reproduces its interior.
Every report is kept as JSON (gpurun_out/parity/, copied to profiles/ per round).
"""
import numpy as np
import pytest

from tests.parity import assert_parity, save_report, stack_report

pytestmark = pytest.mark.gpu

CORES = 16


def _scarp_margin(scale, kt_max):
    """Distance from a crop edge beyond which the crop's own edge mask (WindowedTemplate.py:66-84:
    up to d*sqrt(2) + 2c) and the template support cannot reach."""
    from scarplet_b200 import params as P
    return int(1.4143 * scale + 2 * P.scarp_halfwidth(kt_max) + 16)


def _oracle_sweep(crop, kind, scale, ages):
    """match()'s hierarchical reduce (core.py:285-292) over ``ages``: per-age stacks + result."""
    from oracle import scarplet_oracle as O
    ny, nx = crop.shape
    stacks = [O.calculate_best_fit_parameters(crop, 1.0, 1.0, kind, scale, float(a), processes=CORES)
              for a in ages]
    if len(ages) == 1:
        return stacks[0], np.stack(stacks)
    return np.stack(O.compare(stacks, ny, nx)), np.stack(stacks)


def test_c1_full_raster(cuda_lib):
    """C1: sl.match Scarp scale=100 age=10, +-90 deg at 1 deg, 1024 x 1024 (seed 0): the whole
    raster against the oracle."""
    import scarplet_b200 as sl
    from scarplet_b200.WindowedTemplate import Scarp
    from scarplet_b200.synth import synthetic_dem
    from oracle import scarplet_oracle as O
    z = synthetic_dem(1024, seed=0)
    res = sl.match(sl.DEMGrid(z, 1.0, -1.0), Scarp, scale=100, age=10, ang_min=-np.pi / 2, ang_max=np.pi / 2)
    ref = O.calculate_best_fit_parameters(z, 1.0, -1.0, O.SCARP, 100, 10.0, processes=CORES)
    rep = stack_report(res, ref)
    save_report("c1_full_raster", rep, {"config": "C1 1024^2 Scarp scale 100 age 10, 181 angles, full raster"})
    assert rep["valid"] > 900000
    assert_parity(rep)


def _wrap_composite(z, h):
    """First h and last h + 1 rows / columns of an odd raster put together: its circular
    neighbourhood of the origin (the reference convolves circularly, core.py:359), with the
    raster's parity."""
    idx_y = np.r_[0:h, z.shape[0] - (h + 1):z.shape[0]]
    idx_x = np.r_[0:h, z.shape[1] - (h + 1):z.shape[1]]
    return z[np.ix_(idx_y, idx_x)], idx_y, idx_x


def test_c2_channel_odd_raster(cuda_lib):
    """C2: sl.match Channel scale=10 age=0.1 on 3601 x 3601 (odd: padded 4096 domain, pixel
    units, seed 1, SRTM-like relief): an odd interior crop, and the border strip where the
    wrap-around of the reference's circular convolution is live output (no edge mask,
    WindowedTemplate.py:494-495)."""
    import scarplet_b200 as sl
    from scarplet_b200.WindowedTemplate import Channel
    from scarplet_b200.synth import synthetic_dem
    from oracle import scarplet_oracle as O
    n = 3601
    z = synthetic_dem(n, seed=1, relief=300.0)
    res = sl.match(sl.DEMGrid(z, 1.0, -1.0), Channel, scale=10, age=0.1, ang_min=-np.pi / 2, ang_max=np.pi / 2)
    assert res.shape == (4, n, n) and np.isfinite(res).all()
    # Ricker support along xr ends where exp(-u^2) underflows: |xr| < 87 px at f = 0.1
    m = 110
    c0, size = 1500, 701
    ref, _ = _oracle_sweep(z[c0:c0 + size, c0:c0 + size], O.RICKER, 10, [0.1])
    rep = stack_report(res[:, c0 + m:c0 + size - m, c0 + m:c0 + size - m], ref[:, m:size - m, m:size - m],
                       odd_template=False)
    save_report("c2_interior_crop", rep, {"config": "C2 3601^2 Channel scale 10 f 0.1, 701^2 crop at 1500"})
    assert rep["valid"] > 200000
    assert_parity(rep)
    # wrap-around: rows / columns 0..h-m and n-h+m..n-1 around the origin of the periodic raster
    h = 350
    comp, iy, ix = _wrap_composite(z, h)
    ref, _ = _oracle_sweep(comp, O.RICKER, 10, [0.1])
    keep = np.r_[0:h - m, h + m:2 * h + 1]                 # away from the composite's artificial seam
    sub = res[np.ix_(np.arange(4), iy[keep], ix[keep])]
    rep = stack_report(sub, ref[np.ix_(np.arange(4), keep, keep)], odd_template=False)
    save_report("c2_wraparound_border", rep, {"config": "C2 3601^2 Channel: border strip through the periodic seam"})
    assert rep["valid"] > 200000
    assert_parity(rep)


def test_c3_multi_age_crop(cuda_lib):
    """C3: Scarp scale=100, ages spanning kt = 1 ... 3162, 181 angles on the 4096 x 4096 seed-2
    DEM (not mirrored): crop against the oracle's match() reduce, every orientation / age
    disagreement explained by the reference's exact-tie reset at +-90 degrees or by a top-two
    gap below the tolerance."""
    import scarplet_b200 as sl
    from scarplet_b200.WindowedTemplate import Scarp
    from scarplet_b200.synth import synthetic_dem
    from oracle import scarplet_oracle as O
    n = 4096
    z = synthetic_dem(n, seed=2)
    ages = np.logspace(0, 3.5, 30)[[0, 6, 12, 18, 24, 29]]
    res = np.stack(sl.match(sl.DEMGrid(z, 1.0), Scarp, scale=100, ages=ages, ang_min=-np.pi / 2, ang_max=np.pi / 2))
    m = _scarp_margin(100, ages.max())
    size, c0 = 2 * m + 260, 1700
    ref, stacks = _oracle_sweep(z[c0:c0 + size, c0:c0 + size], O.SCARP, 100, ages)
    inner = (slice(None), slice(m, size - m), slice(m, size - m))
    rep = stack_report(res[:, c0 + m:c0 + size - m, c0 + m:c0 + size - m], ref[inner],
                       ref_age_stacks=stacks[(slice(None),) + inner], ages=ages)
    save_report("c3_six_ages_crop", rep, {"config": "C3 4096^2 Scarp scale 100, 6 ages kt 1..3162 x 181 angles, "
                                                    "%d^2 crop at %d, margin %d" % (size, c0, m)})
    assert rep["valid"] > 50000
    assert_parity(rep)
    assert set(np.unique(res[1][res[3] > 0])) <= set(ages)


def test_c4_multi_scale_crops(cuda_lib):
    """C4: match_scales 25/50/100/200 at age 10 on the 8192 x 8192 seed-3 DEM, all four scales in
    one device sweep (shared curvature spectra, one best state per scale): one crop per scale
    against the oracle."""
    import scarplet_b200 as sl
    from scarplet_b200.WindowedTemplate import Scarp
    from scarplet_b200.synth import synthetic_dem
    from oracle import scarplet_oracle as O
    n = 8192
    z = synthetic_dem(n, seed=3)
    scales = (25, 50, 100, 200)
    multi = sl.match_scales(sl.DEMGrid(z, 1.0), Scarp, scales, age=10.0)
    for k, scale in enumerate(scales):
        m = _scarp_margin(scale, 10.0)
        size, c0 = 2 * m + 300, 3000 + 400 * k
        ref, _ = _oracle_sweep(z[c0:c0 + size, c0:c0 + size], O.SCARP, scale, [10.0])
        rep = stack_report(multi[scale][:, c0 + m:c0 + size - m, c0 + m:c0 + size - m], ref[:, m:size - m, m:size - m])
        save_report("c4_scale_%d_crop" % scale, rep, {"config": "C4 8192^2 Scarp scale %d age 10 (one sweep for 4 "
                                                                "scales), %d^2 crop at %d" % (scale, size, c0)})
        assert rep["valid"] > 50000
        assert_parity(rep)
    del multi
    sl.release()


def test_serial_sweep_angle_major(cuda_lib):
    """calculate_best_fit_parameters_serial (core.py:65-136): the flat angle-outer / age-inner
    sweep over the 35 default ages, on the device, against the oracle's."""
    import scarplet_b200 as sl
    from scarplet_b200.WindowedTemplate import Scarp
    from scarplet_b200.synth import synthetic_dem
    from oracle import scarplet_oracle as O
    z = synthetic_dem(220, seed=31, nx=260)
    out = sl.calculate_best_fit_parameters_serial(sl.DEMGrid(z, 1.0), Scarp, 14, ang_max=np.pi / 6, ang_min=-np.pi / 6)
    assert isinstance(out, tuple) and len(out) == 4
    ref = O.calculate_best_fit_parameters_serial(z, 1.0, 1.0, O.SCARP, 14, ang_max=np.pi / 6, ang_min=-np.pi / 6)
    rep = stack_report(np.stack(out), np.stack(ref))
    save_report("serial_sweep", rep, {"config": "serial sweep 220x260 Scarp scale 14, 35 ages x 61 angles"})
    assert rep["valid"] > 10000
    assert_parity(rep)
    assert set(np.unique(out[1][out[3] > 0])) <= set(O.default_ages())


def test_tiled_multi_age_vs_oracle(cuda_lib):
    """Halo-padded tiles of mixed FFT lengths (forced with a small max_fft), two ages: against the
    ORACLE, not against the single-domain run."""
    from scarplet_b200 import params as P
    from scarplet_b200.engine import Plan
    from scarplet_b200.synth import synthetic_dem
    from scarplet_b200.templates import Scarp
    from oracle import scarplet_oracle as O
    ny, nx = 700, 900
    z = synthetic_dem(ny, seed=9, nx=nx)
    angles = P.search_angles(-np.pi / 2, np.pi / 2)[1::12]     # without the degenerate -90 / +90 degree pair
    ages = [5.0, 40.0]
    with Plan(ny, nx, 1.0, 1.0, max_fft=256) as plan:
        plan.set_dem(z)
        a, t, age_of, angle_of = plan.build_sweep(Scarp._sb_spec, 20, ages, angles)
        plan.reset()
        plan.sweep(a, t)
        res = plan.finalize(age_of, angle_of)
        geo = plan.last_geometry()
    assert geo["tiles_y"] > 2 and geo["tiles_x"] > 2
    import multiprocessing as mp
    from functools import partial
    stacks = []
    with mp.Pool(CORES) as pool:
        for age in ages:
            work = partial(O.match_template, z, 1.0, 1.0, O.SCARP, 20, age)
            stacks.append(np.stack(O.compare(pool.imap(work, angles, chunksize=1), ny, nx)))
    ref = np.stack(O.compare(stacks, ny, nx))
    rep = stack_report(res, ref, ref_age_stacks=np.stack(stacks), ages=ages)
    save_report("tiled_two_ages", rep, {"config": "700x900, max_fft 256 (%d x %d tiles), Scarp scale 20, 2 ages x 15 angles"
                                                  % (geo["tiles_y"], geo["tiles_x"])})
    assert_parity(rep)


@pytest.mark.parametrize("cls,kind", [("LeftFacingUpperBreakScarp", "left_upper_break"),
                                      ("RightFacingUpperBreakScarp", "right_upper_break")])
def test_err_mask_search_fast_path(cuda_lib, cls, kind):
    """get_err_mask templates (core.py:369-371) on the pipelined kernels (float64-exact column
    ranges per raster row) against the oracle, padded and periodic domains."""
    import scarplet_b200 as sl
    from scarplet_b200 import WindowedTemplate as WT
    from scarplet_b200.synth import synthetic_dem
    from oracle import scarplet_oracle as O
    for shape in ((512, 512), (300, 410)):
        z = synthetic_dem(shape[0], seed=41, nx=shape[1])
        res = sl.calculate_best_fit_parameters(sl.DEMGrid(z, 1.0), getattr(WT, cls), 18, 6.0)
        ref = O.calculate_best_fit_parameters(z, 1.0, 1.0, kind, 18, 6.0, processes=CORES)
        rep = stack_report(res, ref)
        save_report("errmask_%s_%dx%d" % (cls, shape[0], shape[1]), rep)
        assert_parity(rep)
    sl.release()


def test_spatial_bands_on_one_device(cuda_lib):
    """Row-slab plans (what each rank of a spatially sharded raster runs, BASELINE config 5) put
    together on one device: against the oracle and the whole-raster plan."""
    from scarplet_b200 import params as P, distributed as D
    from scarplet_b200.engine import Plan
    from scarplet_b200.synth import synthetic_dem
    from scarplet_b200.templates import Scarp
    from oracle import scarplet_oracle as O
    ny, nx = 1200, 640
    z = synthetic_dem(ny, seed=13, nx=nx)
    angles = P.search_angles(-np.pi / 2, np.pi / 2)[1::6]      # without the degenerate -90 / +90 degree pair
    ages = [3.0, 30.0]
    spec = Scarp._sb_spec
    halo = D.slab_halo(spec, 25, ages, angles, nx, ny, 1.0)
    parts, stats, plans = [], [], []
    for r in range(3):
        lo, hi = D.shard_bounds(ny, 3, r)
        plan = Plan(ny, nx, 1.0, 1.0, slab=(lo, hi, halo))
        plan.set_dem(z)
        stats.append(plan.curv_stats())
        plans.append(plan)
    for plan in plans:
        plan.set_curv_stats(sum(s for s, _ in stats), sum(c for _, c in stats))
        a, t, age_of, angle_of = plan.build_sweep(spec, 25, ages, angles)
        plan.reset()
        plan.sweep(a, t)
        parts.append(plan.finalize(age_of, angle_of))
        plan.close()
    got = np.concatenate(parts, axis=1)
    import multiprocessing as mp
    from functools import partial
    stacks = []
    with mp.Pool(CORES) as pool:
        for age in ages:
            work = partial(O.match_template, z, 1.0, 1.0, O.SCARP, 25, age)
            stacks.append(np.stack(O.compare(pool.imap(work, angles, chunksize=1), ny, nx)))
    ref = np.stack(O.compare(stacks, ny, nx))
    rep = stack_report(got, ref, ref_age_stacks=np.stack(stacks), ages=ages)
    save_report("spatial_bands_one_device", rep, {"config": "1200x640 in 3 row bands, halo %d, Scarp scale 25, 2 ages x 30 angles" % halo})
    assert_parity(rep)


def test_noise_level_and_nodata_fill(cuda_lib):
    """SURVEY 8f-3 / 8f-4 on the device: dem.py:152-179 and dem.py:388-414 against the oracle."""
    import scarplet_b200 as sl
    from scarplet_b200.synth import synthetic_dem
    from oracle import scarplet_oracle as O
    z = synthetic_dem(300, seed=9, nx=420)
    z[100:103, 200] = np.nan
    angles, mean, sd = sl.DEMGrid(z, 2.0, 2.0)._estimate_curvature_noiselevel(sigma=25)
    _, r_mean, r_sd = O.estimate_curvature_noiselevel(z, 2.0, 2.0, sigma=25)
    assert np.allclose(sd, r_sd, rtol=1e-9, atol=0)
    assert np.allclose(mean, r_mean, rtol=0, atol=1e-11 * np.max(r_sd))
    holes = z.copy()
    holes[40:60, 70:90] = np.nan
    holes[250, :] = np.nan
    grid = sl.DEMGrid(holes, 1.0)
    grid._fill_nodata()
    assert not np.isnan(grid._griddata).any()
    assert np.allclose(grid._griddata, O.fill_nodata(holes), rtol=1e-12, atol=0)


def test_template_shares_accumulate_to_the_single_search(cuda_lib):
    """What the ranks of an orientation-sharded search compute, one share after the other on one
    device into the same best state: bit-identical to the single sweep.  An orientation's
    curvature spectra, a template's correlation and fit must not depend on where in a batch they
    are computed (the merged multi-GPU result is then bit-identical too: scratch/mgpu_check.py)."""
    from scarplet_b200 import params as P
    from scarplet_b200.engine import Plan
    from scarplet_b200.synth import synthetic_dem
    from scarplet_b200.templates import Channel, Scarp
    angles = P.search_angles(-np.pi / 2, np.pi / 2)
    for spec, n, nx, scale, ages in ((Scarp._sb_spec, 1024, 1024, 50, [3.0, 30.0, 300.0]),
                                     (Channel._sb_spec, 701, 701, 10, [0.1]),
                                     (Scarp._sb_spec, 4096, 384, 60, [2.0, 9.0, 40.0, 180.0, 800.0])):
        z = synthetic_dem(n, seed=11, nx=nx)
        with Plan(n, nx, 1.0, 1.0) as plan:
            plan.set_dem(z)
            a, t, age_of, angle_of = plan.build_sweep(spec, scale, ages, angles)
            plan.reset()
            plan.sweep(a, t)
            single = plan.finalize(age_of, angle_of)
            for world in (8, 3):
                plan.reset()
                for rank in range(world):
                    a, t, _, _ = plan.build_sweep(spec, scale, ages, angles, template_share=(rank, world))
                    plan.sweep(a, t)
                shares = plan.finalize(age_of, angle_of)
                assert np.array_equal(shares, single), (n, nx, world)
