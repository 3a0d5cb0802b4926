"""Host-side logic (no GPU): float64 parameter precomputation, the C ABI surface, the
host mirror of the reference API."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import scarplet_oracle as O
from scarplet_b200 import _lib, params as P
from scarplet_b200 import templates as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_axis_vectors_and_angles_match_oracle():
    for nx, ny, de in ((200, 200, 1.0), (465, 870, 2.0), (3601, 77, 0.5)):
        x, y = P.axis_vectors(nx, ny, de)
        ox, oy = O.axis_vectors(nx, ny, de)
        assert np.array_equal(x, ox) and np.array_equal(y, oy)
    assert np.array_equal(P.search_angles(-np.pi / 2, np.pi / 2), O.search_angles())
    assert len(P.search_angles(-np.pi / 2, np.pi / 2)) == 181
    assert np.array_equal(P.default_ages(), O.default_ages()) and len(P.default_ages()) == 35


@pytest.mark.parametrize("nx,ny,de", [(200, 200, 1.0), (150, 91, 2.0), (64, 301, 0.5)])
def test_window_rectangle_equals_reference_mask(nx, ny, de):
    """get_window_limits (WindowedTemplate.py:66-84) as an index rectangle: exact."""
    x, y = P.axis_vectors(nx, ny, de)
    rng = np.random.default_rng(nx)
    for angle in list(O.search_angles()[::15]) + list(rng.uniform(-1.6, 1.6, 10)):
        for kt, d in ((10., 20. * de), (300., 8. * de), (1., 60. * de)):
            c = P.scarp_halfwidth(kt)
            i_lo, i_hi, j_lo, j_hi = P.window_rectangle(x, y, -angle, c, d)
            mask = np.ones((ny, nx), dtype=bool)
            if i_hi >= i_lo and j_hi >= j_lo:
                mask[i_lo:i_hi + 1, j_lo:j_hi + 1] = False
            assert np.array_equal(mask, O.window_limits(O.SCARP, nx, ny, de, -angle, c, d))


@pytest.mark.parametrize("nx,ny,de", [(120, 90, 1.0), (77, 65, 2.0)])
def test_support_box_covers_template(nx, ny, de):
    x, y = P.axis_vectors(nx, ny, de)
    rng = np.random.default_rng(ny)
    for _ in range(25):
        angle = rng.uniform(-np.pi / 2, np.pi / 2)
        for spec, kind, scale, age in ((T.Scarp._sb_spec, O.SCARP, 15 * de, 10 ** rng.uniform(0, 3)),
                                       (T.Ricker._sb_spec, O.RICKER, 6 * de, rng.uniform(0.005, 0.4))):
            rec = P.template_record(spec, scale, age, angle, nx, ny, de, x, y, 0, 0)
            t = O.template_array(kind, scale, age, angle, nx, ny, de)
            ii, jj = np.nonzero(t)
            if ii.size == 0:
                continue
            a0, b0 = ny // 2, nx // 2
            assert rec.sy_lo <= ii.min() - a0 and ii.max() - a0 <= rec.sy_hi
            assert rec.sx_lo <= jj.min() - b0 and jj.max() - b0 <= rec.sx_hi
            assert rec.sy_lo >= -a0 and rec.sy_hi <= ny - 1 - a0
            s = rec.tscale
            assert s > 0 and np.log2(s) == np.round(np.log2(s))      # exact power of two


def test_template_record_matches_reference_constants():
    x, y = P.axis_vectors(100, 100, 1)
    rec = P.template_record(T.Scarp._sb_spec, 100, 10, 0.3, 100, 100, 1, x, y, 3, 7)
    assert rec.c == O.scarp_halfwidth(10) and rec.d == 100
    assert rec.cos_t == np.cos(-0.3) and rec.sin_t == np.sin(-0.3)
    assert rec.k0 == 2. * 10 ** (3 / 2.) * np.sqrt(np.pi) and rec.k1 == 40.
    assert (rec.angle_id, rec.idx, rec.kind, rec.errmode) == (3, 7, 0, 0)
    rec = P.template_record(T.Channel._sb_spec, 10, 0.1, 0.3, 100, 100, 1, x, y, 0, 0)
    assert rec.c == 100 and rec.k0 == np.pi * 0.1 and rec.kind == 1
    assert (rec.i_lo, rec.i_hi, rec.j_lo, rec.j_hi) == (0, 99, 0, 99)     # Ricker: no edge mask
    rec = P.template_record(T.RightFacingUpperBreakScarp._sb_spec, 10, 5., 0.3, 100, 100, 1, x, y, 0, 0)
    assert rec.sign == -1.0 and rec.errmode == P.ERRMASK_XR_LE0


def test_plugin_surface_host_methods_match_oracle():
    for cls, kind in ((T.Scarp, O.SCARP), (T.LeftFacingUpperBreakScarp, O.LEFT_UPPER),
                      (T.RightFacingUpperBreakScarp, O.RIGHT_UPPER)):
        obj = cls(12, 4.0, 0.7, 64, 50, 1.0)
        assert obj.alpha == -0.7 and obj.c == O.scarp_halfwidth(4.0)
        assert np.array_equal(obj.get_window_limits(), O.window_limits(kind, 64, 50, 1.0, -0.7, obj.c, 12))
        assert np.array_equal(obj.get_mask(), O.window_mask(64, 50, 1.0, -0.7, obj.c, 12))
        em = O.err_mask(kind, 0.7, 64, 50, 1.0)
        if em is not None:
            assert np.array_equal(obj.get_err_mask(), em)
    r = T.Channel(6, 0.1, -0.2, 40, 30, 2.0)
    assert r.c == 40 and not r.get_window_limits().any()
    assert T.device_spec(T.Channel) is T.Ricker._sb_spec
    # a class without an on-device generator is served through its own methods (plugin path)
    assert T.device_spec(dict) is None


def test_sweep_index_orders():
    """Flat indices reproduce the reference's reduction orders (core.py:285-292, 116-134)."""
    class FakePlan(object):
        nx, ny, dx = 64, 64, 1.0
        x, y = P.axis_vectors(64, 64, 1.0)
    from scarplet_b200.engine import Plan
    angles = P.search_angles(-0.05, 0.05)
    ages = [1.0, 10.0, 100.0]
    (a, na), (t, nt, _w), age_of, angle_of = Plan.build_sweep(FakePlan, T.Scarp._sb_spec, 8, ages, angles, "age_major")
    assert na == len(angles) and nt == len(angles) * 3
    for k in range(nt):
        assert age_of[t[k].idx] == ages[k % 3] and angle_of[t[k].idx] == angles[t[k].angle_id]
        assert t[k].idx == (k % 3) * len(angles) + t[k].angle_id
    (a, na), (t, nt, _w), age_of, angle_of = Plan.build_sweep(FakePlan, T.Scarp._sb_spec, 8, ages, angles, "angle_major",
                                                          angle_slice=(2, 5))
    assert na == 3 and nt == 9 and [t[k].idx for k in range(nt)] == list(range(6, 15))


def test_header_symbols_exported_by_library():
    """Every function declared in include/scarplet_b200.h is exported by the built
    library and bound by the ctypes layer (no compute calls here: no GPU)."""
    header = open(os.path.join(ROOT, "include", "scarplet_b200.h")).read()
    declared = set(re.findall(r"\b(sb_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.EXPORTED), declared ^ set(_lib.EXPORTED)
    import __graft_entry__
    __graft_entry__.build()
    lib = ctypes.CDLL(_lib.library_path())
    for name in declared:
        assert hasattr(lib, name), name
    assert ctypes.sizeof(_lib.SbTemplate) == 8 * 8 + 14 * 4
    assert ctypes.sizeof(_lib.SbAngle) == 32


def test_product_has_no_cpu_path():
    """Without a CUDA device the library must refuse to create a plan, loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    import __graft_entry__
    __graft_entry__.build()
    from scarplet_b200.engine import Plan
    _lib._use_library(None)
    with pytest.raises(_lib.SbError, match="no CUDA device|CPU path"):
        Plan(64, 64, 1.0, 1.0)
    # nothing under scarplet_b200/ may import the oracle
    pkg = os.path.join(ROOT, "scarplet_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(dirpath, fn)).read().lower().replace("no cpu", ""), fn


def test_shard_bounds():
    from scarplet_b200.distributed import shard_bounds
    for n, w in ((181, 8), (181, 2), (5, 8), (30, 4)):
        cuts = [shard_bounds(n, w, r) for r in range(w)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in cuts]
        assert max(sizes) - min(sizes) <= 1
    assert [shard_bounds(181, 8, r)[1] - shard_bounds(181, 8, r)[0] for r in range(8)] == [23] * 5 + [22] * 3


@pytest.mark.parametrize("ny,nx,de,cls,scale,ages", [
    (200, 200, 1.0, "Scarp", 100.0, [10.0]),
    (465, 870, 2.0, "Scarp", 30.0, None),
    (300, 301, 1.0, "Channel", 10.0, [0.1, 0.02]),
    (150, 170, 1.0, "RightFacingUpperBreakScarp", 10.0, [2.0, 20.0]),
    (64, 96, -1.0, "Scarp", 8.0, [2.0]),               # falling axes: per-record fallback
])
def test_bulk_template_records_equal_single(ny, nx, de, cls, scale, ages):
    """The vectorised record builder (one call per sweep) against the per-template one that
    mirrors the reference expression by expression."""
    from scarplet_b200 import params as P
    from scarplet_b200 import templates as T
    from scarplet_b200._lib import SbTemplate
    spec = getattr(T, cls)._sb_spec
    ages = P.default_ages() if ages is None else np.asarray(ages, dtype=np.float64)
    angles = P.search_angles(-np.pi / 2, np.pi / 2)
    x, y = P.axis_vectors(nx, ny, de)
    A, G = len(angles), len(ages)
    idx = np.arange(G)[None, :] * A + np.arange(A)[:, None]
    recs = P.template_records(spec, scale, ages, angles, nx, ny, de, x, y, np.arange(A), idx)
    assert recs.shape == (A, G) and recs.dtype.itemsize == ctypes.sizeof(SbTemplate)
    for a in range(0, A, 3):
        for g in range(G):
            one = P.template_record(spec, scale, ages[g], angles[a], nx, ny, de, x, y, a, idx[a, g])
            assert tuple(getattr(one, n) for n, _ in SbTemplate._fields_) == tuple(recs[a, g].tolist())
    arr = P.records_to_ctypes(recs)
    assert arr[G + 1].idx == idx[1, 1 % G] if G > 1 else arr[1].idx == idx[1, 0]


def test_template_shares_partition_the_search(monkeypatch):
    """``build_sweep(template_share=(rank, world))``: the ranks' records are a partition of the
    whole search (orientation-major, contiguous), indices unchanged; shares of equal estimated
    device time by default (a few per cent apart in count), of equal count (+- 1 template) with
    ``SB_SHARE_BALANCE=count``."""
    from scarplet_b200.engine import Plan

    class FakePlan(object):
        nx, ny, dx = 64, 64, 1.0
        x, y = P.axis_vectors(64, 64, 1.0)
    angles = P.search_angles(-np.pi / 2, np.pi / 2)
    ages = np.logspace(0, 3.5, 30)
    _, (full, n_full, _w), age_of, angle_of = Plan.build_sweep(FakePlan, T.Scarp._sb_spec, [8, 12], ages, angles, "age_major")
    assert n_full == 181 * 30 * 2
    for world, mode in ((1, "cost"), (2, "cost"), (8, "cost"), (7, "cost"), (8, "count"), (7, "count")):
        monkeypatch.setenv("SB_SHARE_BALANCE", mode)
        FakePlan._sweep_memo = []
        seen, sizes = [], []
        for rank in range(world):
            (a, na), (t, nt, _w), _, _ = Plan.build_sweep(FakePlan, T.Scarp._sb_spec, [8, 12], ages, angles, "age_major",
                                                      template_share=(rank, world))
            sizes.append(nt)
            for k in range(nt):
                assert 0 <= t[k].angle_id < na
                seen.append((t[k].idx, t[k].state, round(angle_of[t[k].idx], 12)))
                # the angle record the template points at is its own orientation
                assert np.isclose(a[t[k].angle_id].cos_a, np.cos(angle_of[t[k].idx]))
        assert sum(sizes) == n_full and min(sizes) > 0
        if mode == "count":
            assert max(sizes) - min(sizes) <= 1
        assert seen == [(full[k].idx, full[k].state, round(angle_of[full[k].idx], 12)) for k in range(n_full)]


def test_share_cuts_edge_cases(monkeypatch):
    """``Plan._share_cuts``: monotone cut points from 0 to the number of templates for any world
    size -- more ranks than templates leaves some shares empty, never a template unassigned."""
    from scarplet_b200.engine import Plan

    class FakePlan(object):
        nx, ny, dx = 200, 160, 1.0
        x, y = P.axis_vectors(200, 160, 1.0)
    angles = P.search_angles(-np.pi / 2, np.pi / 2)
    for mode in ("cost", "count"):
        monkeypatch.setenv("SB_SHARE_BALANCE", mode)
        for n_angles, ages, world in ((181, [10.0], 8), (181, [1.0, 100.0, 3000.0], 5), (3, [10.0], 8), (1, [10.0], 2),
                                      (181, np.logspace(0, 3.5, 30), 1)):
            cuts = Plan._share_cuts(FakePlan, T.Scarp._sb_spec, [20.0], np.asarray(ages), angles[:n_angles], world)
            n = n_angles * len(ages)
            assert len(cuts) == world + 1 and cuts[0] == 0 and cuts[-1] == n and np.all(np.diff(cuts) >= 0), (mode, cuts)
            if n >= 2 * world:
                assert np.all(np.diff(cuts) > 0)
