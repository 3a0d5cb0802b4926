"""BASELINE.json configs 1, 2, 4, 5 (and the 16384^2 north-star case) on one GPU:
device time of the search, throughput, and parity against the oracle on a crop
(compact template support => an oracle run on a crop reproduces the interior).

    python scratch/run_configs.py [c1 c2 c4 ns c5] > gpurun_out/configs.json

Test/measurement infrastructure only (imports oracle/).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import __graft_entry__  # noqa: E402
from oracle import scarplet_oracle as O  # noqa: E402
from parity import stack_report  # noqa: E402
from scarplet_b200 import params as P  # noqa: E402
from scarplet_b200.engine import Plan  # noqa: E402
from scarplet_b200.synth import synthetic_dem  # noqa: E402
from scarplet_b200.templates import Channel, Scarp  # noqa: E402

KEEP = ("valid", "index_agreement", "mask_mismatch_unexplained", "tie_reset_pixels", "snr_rel_p50",
        "snr_rel_max", "amp_rel_max", "frac_snr_over_tol", "frac_amp_over_tol", "disagree_snr_gap_max")


def big_dem(n, seed):
    """Large rasters: a 4096^2 fractal tile repeated with per-tile mirroring, plus tilt,
    one scarp and fresh white noise (cheap to build; parity is judged on crops)."""
    if n <= 8192:
        return synthetic_dem(n, seed)
    base = synthetic_dem(4096, seed) - 650.0
    reps = n // 4096
    rows = []
    for i in range(reps):
        row = [base[::(-1 if i % 2 else 1), ::(-1 if j % 2 else 1)] for j in range(reps)]
        rows.append(np.concatenate(row, axis=1))
    z = np.concatenate(rows, axis=0)
    rng = np.random.default_rng(seed)
    z = z + 650.0 + 0.03 * rng.standard_normal(z.shape, dtype=np.float32)
    return z.astype(np.float32).astype(np.float64)


def scarp_margin(scale, kt_max):
    """Distance from a crop edge beyond which the crop's own edge mask
    (WindowedTemplate.py:66-84: up to d*sqrt(2) + 2c) and template support cannot reach."""
    return int(1.4143 * scale + 2 * P.scarp_halfwidth(kt_max) + 16)


def timed_search(z, spec, scale, ages, angles, repeats=2):
    ny, nx = z.shape
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        plan = Plan(ny, nx, 1.0, 1.0, device=0, stream=stream.cuda_stream)
        t0 = time.perf_counter()
        plan.set_dem(z)
        a, t, age_of, angle_of = plan.build_sweep(spec, scale, ages, angles, "age_major")
        host_s = time.perf_counter() - t0
        ms = []
        for _ in range(repeats):
            plan.reset()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            plan.sweep(a, t)
            e1.record(stream)
            stream.synchronize()
            ms.append(e0.elapsed_time(e1))
        out = plan.finalize(age_of, angle_of)
        geo = plan.last_geometry()
        plan.close()
    evals = ny * nx * len(ages) * len(angles)
    return out, {"ms": min(ms), "px_evals": evals, "Mpx_evals_per_s": evals / (min(ms) * 1e-3) / 1e6,
                 "setup_host_s": host_s, "geometry": geo}


def crop_parity(z, res, kind, scale, ages, c0, size, margin, odd=True):
    crop = z[c0:c0 + size, c0:c0 + size]
    t0 = time.perf_counter()
    if len(ages) == 1:
        ref = O.calculate_best_fit_parameters(crop, 1.0, 1.0, kind, scale, float(ages[0]), processes=16)
    else:
        # match()'s hierarchical reduce (core.py:285-292) over the given age list
        stacks = [O.calculate_best_fit_parameters(crop, 1.0, 1.0, kind, scale, float(a), processes=16)
                  for a in ages]
        ref = np.stack(O.compare(stacks, size, size))
    dt = time.perf_counter() - t0
    sub = res[:, c0 + margin:c0 + size - margin, c0 + margin:c0 + size - margin]
    rsub = ref[:, margin:size - margin, margin:size - margin]
    rep = stack_report(sub, rsub, odd_template=odd)
    out = {k: rep[k] for k in KEEP}
    out["oracle_s"] = dt
    out["crop"] = [c0, size, margin]
    return out


def main():
    which = sys.argv[1:] or ["c1", "c2", "c4", "ns"]
    __graft_entry__.build()
    angles = P.search_angles(-np.pi / 2, np.pi / 2)
    results = {}

    if "c1" in which:
        z = synthetic_dem(1024, 0)
        res, info = timed_search(z, Scarp._sb_spec, 100, [10.0], angles, repeats=3)
        ref = O.calculate_best_fit_parameters(z, 1.0, 1.0, O.SCARP, 100, 10.0, processes=16)
        rep = stack_report(res, ref)
        info["parity_full_raster"] = {k: rep[k] for k in KEEP}
        results["C1 Scarp scale=100 age=10, 1024^2"] = info
        print("c1 done", file=sys.stderr)

    if "c2" in which:
        z = synthetic_dem(3601, 1, relief=300.0)
        res, info = timed_search(z, Channel._sb_spec, 10, [0.1], angles, repeats=3)
        info["finite"] = bool(np.isfinite(res).all())
        # Ricker support along xr: exp underflow at u^2 > 745 => |xr| < 87 px; no edge mask
        # the crop must have the raster's parity (odd): the template's pixel-centre offsets are
        # half-integer on an even axis (WindowedTemplate.py:50-53)
        info["parity_crop"] = crop_parity(z, res, O.RICKER, 10, [0.1], 1500, 701, 110, odd=False)
        results["C2 Channel scale=10 age=0.1, 3601^2 (dx=1)"] = info
        print("c2 done", file=sys.stderr)

    if "c4" in which:
        z = synthetic_dem(8192, 3)
        per_scale = {}
        tot_ms, tot_evals = 0.0, 0
        for scale in (25, 50, 100, 200):
            res, info = timed_search(z, Scarp._sb_spec, scale, [10.0], angles, repeats=2)
            margin = scarp_margin(scale, 10.0)
            size = 2 * margin + 300
            info["parity_crop"] = crop_parity(z, res, O.SCARP, scale, [10.0], 3000, size, margin)
            per_scale[str(scale)] = info
            tot_ms += info["ms"]
            tot_evals += info["px_evals"]
        results["C4 multi-scale Scarp 25/50/100/200 age=10, 8192^2"] = {
            "per_scale": per_scale, "ms": tot_ms, "px_evals": tot_evals,
            "Mpx_evals_per_s": tot_evals / (tot_ms * 1e-3) / 1e6}
        print("c4 done", file=sys.stderr)

    if "ns" in which:
        n = 16384
        z = big_dem(n, 4)
        ages = np.logspace(0, 3.5, 30)
        res, info = timed_search(z, Scarp._sb_spec, 100, ages, angles, repeats=1)
        info["finite"] = bool(np.isfinite(res).all())
        info["snr_nonneg"] = bool((res[3] >= 0).all())
        hit = res[3] > 0
        info["ages_in_grid"] = bool(np.isin(res[1][hit], ages).all())
        info["angles_in_grid"] = bool(np.isin(res[2][hit], angles).all())
        info["valid_fraction"] = float(hit.mean())
        # crop parity on a 6-age subset of the same search, across a tile seam (8192 boundary region)
        sub_ages = ages[::5]
        res6, info6 = timed_search(z, Scarp._sb_spec, 100, sub_ages, angles, repeats=1)
        margin = scarp_margin(100, sub_ages.max())
        info["parity_crop_6_ages_at_tile_seam"] = crop_parity(z, res6, O.SCARP, 100, sub_ages,
                                                              7800, 2 * margin + 160, margin)
        info["six_age_run"] = {k: info6[k] for k in ("ms", "Mpx_evals_per_s")}
        results["north-star: Scarp scale=100, 30 ages x 181 angles, 16384^2, ONE GPU"] = info
        print("ns done", file=sys.stderr)

    if "c5" in which:
        n = 32768
        z = big_dem(n, 4)
        res, info = timed_search(z, Scarp._sb_spec, 100, [10.0], angles, repeats=1)
        info["finite"] = bool(np.isfinite(res).all())
        info["parity_crop"] = crop_parity(z, res, O.SCARP, 100, [10.0], 16200, 2 * scarp_margin(100, 10.0) + 200,
                                          scarp_margin(100, 10.0))
        results["C5 (one age of 30): Scarp scale=100 age=10 x 181 angles, 32768^2, ONE GPU"] = info
        print("c5 done", file=sys.stderr)

    print(json.dumps(results, indent=1))


if __name__ == "__main__":
    main()
