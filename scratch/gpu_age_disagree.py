"""Where do multi-age searches disagree with the oracle on the best (age, angle)?"""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__
from oracle import scarplet_oracle as O
from parity import stack_report
from scarplet_b200 import params as P
from scarplet_b200.engine import Plan
from scarplet_b200.synth import synthetic_dem
from scarplet_b200.templates import Scarp, Channel

__graft_entry__.build()
angles = P.search_angles(-np.pi / 2, np.pi / 2)
all_ages = np.logspace(0, 3.5, 30)


def gpu(z, spec, scale, ages, precision=32, max_fft=None):
    ny, nx = z.shape
    with Plan(ny, nx, 1.0, 1.0, precision=precision, max_fft=max_fft) as plan:
        plan.set_dem(z)
        a, t, age_of, angle_of = plan.build_sweep(spec, scale, ages, angles, "age_major")
        plan.reset(); plan.sweep(a, t)
        return plan.finalize(age_of, angle_of), plan.last_geometry()


def oracle(crop, kind, scale, ages):
    stacks = [O.calculate_best_fit_parameters(crop, 1.0, 1.0, kind, scale, float(a), processes=16) for a in ages]
    if len(stacks) == 1:
        return stacks[0], stacks
    return np.stack(O.compare(stacks, *crop.shape)), stacks


def analyse(name, res, ref, stacks, ages, m, odd=True):
    sub = res[:, m:-m, m:-m] if res.shape == ref.shape else res
    rsub = ref[:, m:-m, m:-m]
    rep = stack_report(sub, rsub, odd_template=odd)
    # per-pixel second-best reference SNR over ages (best angle per age): how flat is the optimum?
    snr_by_age = np.stack([s[3][m:-m, m:-m] for s in stacks])
    srt = np.sort(snr_by_age, axis=0)
    valid = rsub[3] > 0
    out = {k: rep[k] for k in ("valid", "index_agreement", "snr_rel_max", "frac_snr_over_tol", "disagree_snr_gap_max",
                               "mask_mismatch_unexplained", "tie_reset_pixels")}
    bad = valid & ~((np.isclose(sub[1], rsub[1], rtol=1e-12) & np.isclose(sub[2], rsub[2], atol=1e-12)))
    gap = np.abs(sub[3] - rsub[3]) / np.where(valid, rsub[3], 1)
    out["n_disagree"] = int(bad.sum())
    if bad.any():
        out["disagree_gap_p50"] = float(np.median(gap[bad]))
        out["disagree_gap_p99"] = float(np.quantile(gap[bad], 0.99))
        out["disagree_frac_gap_over_1e-4"] = float((gap[bad] > 1e-4).mean())
        out["disagree_same_age_frac"] = float(np.isclose(sub[1], rsub[1], rtol=1e-12)[bad].mean())
        out["disagree_angle_step_deg_p50"] = float(np.median(np.abs(sub[2] - rsub[2])[bad]) * 180 / np.pi)
        if len(stacks) > 1:
            out["top2_age_gap_p50_at_disagree"] = float(np.median(((srt[-1] - srt[-2]) / srt[-1])[bad]))
    print(name, json.dumps(out), flush=True)


n, c0, size = 2048, 600, 774
z = synthetic_dem(n, seed=7)
crop = z[c0:c0 + size, c0:c0 + size]
for label, ages in (("age=2512 only", all_ages[-1:]), ("age=10 only", np.array([10.0])), ("6 ages", all_ages[::5])):
    m = int(1.4143 * 100 + 2 * P.scarp_halfwidth(ages.max()) + 16)
    ref, stacks = oracle(crop, O.SCARP, 100, ages)
    for prec in (32, 64):
        res, geo = gpu(z, Scarp._sb_spec, 100, ages, precision=prec)
        analyse("%s, complex%d, %s" % (label, prec * 4, "P=%d tiles=%d" % (geo["Py"], geo["tiles_y"])),
                res[:, c0:c0 + size, c0:c0 + size], ref, stacks, ages, m)

# C2 analogue with matching parity: odd raster, odd crop, tiles
z = synthetic_dem(3601, 1, relief=300.0)
res, geo = gpu(z, Channel._sb_spec, 10, [0.1])
for size in (701, 700):
    crop = z[1500:1500 + size, 1500:1500 + size]
    ref, stacks = oracle(crop, O.RICKER, 10, [0.1])
    analyse("C2 Channel crop %d (raster 3601), P=%d tiles=%d" % (size, geo["Py"], geo["tiles_y"]),
            res[:, 1500:1500 + size, 1500:1500 + size], ref, stacks, [0.1], 110, odd=False)
