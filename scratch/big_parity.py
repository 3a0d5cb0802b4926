"""Parity of the two large BASELINE workloads against the oracle, on ONE GPU (the multi-GPU runs
are bit-identical to / within rounding of the single-device ones: scratch/mgpu_check.py):

* north-star: the FULL 16384^2 x 30 ages x 181 angles Scarp search on the seeded, non-mirrored
  device-generated DEM (bench.device_dem_rows); a crop across a seam of the mixed-length FFT
  tiles is compared with the oracle's match() reduce over all 30 ages;
* C5: band 1 of 8 of the 32768^2 raster as a row-slab plan (what rank 1 of the 8-GPU run
  holds: its rows + halo only), 6 ages spanning kt = 1 ... 3162; a crop in the band's first
  rows -- the seam with band 0 -- against the oracle.

    python scratch/big_parity.py [ns] [c5] > gpurun_out/big_parity.json
Test / measurement infrastructure only (imports oracle/).
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
from bench import device_dem_rows  # noqa: E402
from oracle import scarplet_oracle as O  # noqa: E402
from parity import assert_parity, stack_report  # noqa: E402
from scarplet_b200 import distributed as D  # noqa: E402
from scarplet_b200 import params as P  # noqa: E402
from scarplet_b200.engine import Plan  # noqa: E402
from scarplet_b200.templates import Scarp  # noqa: E402

CORES = os.cpu_count() or 16
KEEP = ("valid", "index_agreement", "disagree", "disagree_tie_reset", "disagree_top_two_within_tol",
        "disagree_unexplained", "mask_mismatch_unexplained", "tie_reset_pixels", "snr_rel_p50", "snr_rel_max_strong",
        "amp_rel_max_strong", "frac_snr_over_tol", "frac_amp_over_tol", "snr_median_ref")


def oracle_reduce(crop, ages):
    t0 = time.perf_counter()
    stacks = [O.calculate_best_fit_parameters(crop, 1.0, 1.0, O.SCARP, 100, float(a), processes=CORES) for a in ages]
    ref = np.stack(O.compare(stacks, crop.shape[0], crop.shape[1]))
    return ref, np.stack(stacks), time.perf_counter() - t0


def margin(kt_max):
    return int(1.4143 * 100 + 2 * P.scarp_halfwidth(kt_max) + 16)


def crop_report(res_rows, row_base, z_crop, c0y, c0x, size, m, ages):
    """res_rows: (4, rows, nx) stack whose first row is raster row row_base."""
    ref, stacks, dt = oracle_reduce(z_crop, ages)
    sub = res_rows[:, c0y + m - row_base:c0y + size - m - row_base, c0x + m:c0x + size - m]
    inner = (slice(None), slice(m, size - m), slice(m, size - m))
    rep = stack_report(sub, ref[inner], ref_age_stacks=stacks[(slice(None),) + inner], ages=ages)
    out = {k: rep[k] for k in KEEP if k in rep}
    out["oracle_s"] = dt
    out["crop"] = {"row0": c0y, "col0": c0x, "size": size, "margin": m}
    try:
        assert_parity(rep)
        out["within_north_star_tolerances"] = True
    except AssertionError:
        out["within_north_star_tolerances"] = False
    return out


def main():
    which = sys.argv[1:] or ["ns", "c5"]
    __graft_entry__.build()
    device = torch.device("cuda", 0)
    angles = P.search_angles(-np.pi / 2, np.pi / 2)
    spec = Scarp._sb_spec
    results = {}
    stream = torch.cuda.Stream(device=device)
    with torch.cuda.stream(stream):
        if "ns" in which:
            n, ages = 16384, np.logspace(0, 3.5, 30)
            z = device_dem_rows(n, 4, 0, n, device)
            stream.synchronize()
            with Plan(n, n, 1.0, 1.0, device=0, stream=stream.cuda_stream) as plan:
                plan.set_dem_device(z.data_ptr())
                a, t, age_of, angle_of = plan.build_sweep(spec, 100, ages, angles)
                plan.reset()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                plan.sweep(a, t)
                e1.record(stream)
                stream.synchronize()
                ms = e0.elapsed_time(e1)
                geo = plan.last_geometry()
                m = margin(ages.max())
                size = 2 * m + 200
                seam = -(-n * 3783 // 16867)             # first output row of the second FFT tile (4 x 4096 + 2048)
                c0 = seam - size // 2
                rows = (c0, c0 + size)
                res = plan.finalize(age_of, angle_of, rows=rows)
            zc = z[c0:c0 + size, c0:c0 + size].cpu().numpy()
            rep = crop_report(res, c0, zc, c0, c0, size, m, ages)
            rep.update({"search_ms_one_gpu": ms, "px_evals": n * n * 30 * 181, "geometry": geo,
                        "Mpx_evals_per_s_one_gpu": n * n * 30 * 181 / (ms * 1e-3) / 1e6,
                        "dem": "bench.device_dem_rows(16384, seed 4): seeded, not mirrored"})
            results["north-star 16384^2, 30 ages x 181 angles (full search), crop across an FFT-tile seam"] = rep
            del z
            torch.cuda.empty_cache()
            print("ns done", file=sys.stderr)
        if "c5" in which:
            n = 32768
            ages = np.logspace(0, 3.5, 30)[[0, 6, 12, 18, 24, 29]]
            lo, hi = D.shard_bounds(n, 8, 1)                           # band 1 of 8
            halo = D.slab_halo(spec, 100, ages, angles, n, n, 1.0)
            with Plan(n, n, 1.0, 1.0, device=0, stream=stream.cuda_stream, slab=(lo, hi, halo)) as plan:
                r0, nrows = plan.dem_rows()
                z = device_dem_rows(n, 4, r0, nrows, device)
                stream.synchronize()
                plan.set_dem_device(z.data_ptr())
                a, t, age_of, angle_of = plan.build_sweep(spec, 100, ages, angles)
                plan.reset()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                plan.sweep(a, t)
                e1.record(stream)
                stream.synchronize()
                ms = e0.elapsed_time(e1)
                geo = plan.last_geometry()
                m = margin(ages.max())
                size = 2 * m + 200
                c0y = lo - m                                              # interior = the band's first 200 rows
                c0x = -(-n // 9) - size // 2                              # across the first x-tile seam (9 x 4096)
                res = plan.finalize(age_of, angle_of, rows=(lo, lo + size))
                mem = plan.device_bytes
            zc = device_dem_rows(n, 4, c0y, size, device)[:, c0x:c0x + size].cpu().numpy()
            rep = crop_report(res, lo, zc, c0y, c0x, size, m, ages)
            rep.update({"band": [lo, hi], "halo_rows": halo, "dem_rows_held": nrows, "search_ms": ms, "geometry": geo,
                        "plan_device_GB": mem / 1e9,
                        "Mpx_evals_per_s_one_gpu": (hi - lo) * n * len(ages) * 181 / (ms * 1e-3) / 1e6})
            results["C5 32768^2: band 1 of 8 as a row-slab plan, 6 ages x 181 angles, crop at the seam with band 0"] = rep
            print("c5 done", file=sys.stderr)
    print(json.dumps(results, indent=1))


if __name__ == "__main__":
    main()
