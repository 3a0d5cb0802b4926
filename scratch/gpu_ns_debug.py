"""Why did the 16384^2 six-age run agree with the oracle on only 95.8 % of (age, angle)?"""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "scratch"))
from oracle import scarplet_oracle as O
from parity import stack_report
from scarplet_b200 import params as P
from scarplet_b200.engine import Plan
from scarplet_b200.templates import Scarp
from run_configs import big_dem, scarp_margin

angles = P.search_angles(-np.pi / 2, np.pi / 2)
ages = np.logspace(0, 3.5, 30)[::5]
KEEP = ("valid", "index_agreement", "snr_rel_max", "frac_snr_over_tol", "disagree_snr_gap_max", "mask_mismatch_unexplained")


def gpu(z, max_fft=None):
    ny, nx = z.shape
    with Plan(ny, nx, 1.0, 1.0, max_fft=max_fft) as plan:
        plan.set_dem(z)
        a, t, age_of, angle_of = plan.build_sweep(Scarp._sb_spec, 100, ages, angles, "age_major")
        plan.reset(); plan.sweep(a, t)
        return plan.finalize(age_of, angle_of), plan.last_geometry()


n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
z = big_dem(n, 4)
m = scarp_margin(100, ages.max())
size = 2 * m + 160
c0 = 7800 if n == 16384 else n // 2 - size // 2
crop = z[c0:c0 + size, c0:c0 + size]
stacks = [O.calculate_best_fit_parameters(crop, 1.0, 1.0, O.SCARP, 100, float(a), processes=16) for a in ages]
ref = np.stack(O.compare(stacks, size, size))
rsub = ref[:, m:-m, m:-m]
small, g0 = gpu(crop)
rep = stack_report(small[:, m:-m, m:-m], rsub)
print("crop as its own raster on the GPU (P=%d):" % g0["Py"], json.dumps({k: rep[k] for k in KEEP}), flush=True)
for max_fft in (None, 4096):
    res, geo = gpu(z, max_fft)
    sub = res[:, c0 + m:c0 + size - m, c0 + m:c0 + size - m]
    rep = stack_report(sub, rsub)
    print("full raster %d, P=%d tiles=%d tile_out~%d:" % (n, geo["Py"], geo["tiles_y"], -(-n // geo["tiles_y"])),
          json.dumps({k: rep[k] for k in KEEP}), flush=True)
    bad = (rsub[3] > 0) & ~(np.isclose(sub[1], rsub[1], rtol=1e-12) & np.isclose(sub[2], rsub[2], atol=1e-12))
    ii, jj = np.nonzero(bad)
    if len(ii):
        gap = np.abs(sub[3] - rsub[3])[bad] / rsub[3][bad]
        print("  disagree: n=%d rows %d..%d cols %d..%d (raster coords), gap p50 %.2e max %.2e, frac gap>1e-4 %.3f"
              % (len(ii), ii.min() + c0 + m, ii.max() + c0 + m, jj.min() + c0 + m, jj.max() + c0 + m,
                 np.median(gap), gap.max(), (gap > 1e-4).mean()))
        print("  row histogram (16 bins):", np.histogram(ii, bins=16, range=(0, 160))[0].tolist())
        print("  col histogram (16 bins):", np.histogram(jj, bins=16, range=(0, 160))[0].tolist())
        k = min(8, len(ii))
        for a in range(k):
            i, j = ii[a], jj[a]
            print("   px (%d,%d): gpu age %.1f ang %.4f snr %.6g amp %.5g | ref age %.1f ang %.4f snr %.6g amp %.5g"
                  % (i + c0 + m, j + c0 + m, sub[1, i, j], sub[2, i, j], sub[3, i, j], sub[0, i, j],
                     rsub[1, i, j], rsub[2, i, j], rsub[3, i, j], rsub[0, i, j]))
