"""Tile seams and mirrored terrain: which one lowers the (age, angle) agreement?"""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import scarplet_oracle as O
from parity import stack_report
from scarplet_b200 import params as P
from scarplet_b200.engine import Plan
from scarplet_b200.synth import synthetic_dem
from scarplet_b200.templates import Scarp

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


def oracle(crop):
    stacks = [O.calculate_best_fit_parameters(crop, 1.0, 1.0, O.SCARP, 100, float(a), processes=16) for a in ages]
    return np.stack(O.compare(stacks, *crop.shape))


m = int(1.4143 * 100 + 2 * P.scarp_halfwidth(ages.max()) + 16)
size = 2 * m + 160
base = synthetic_dem(2048, seed=4)
plain = synthetic_dem(4096, seed=5)
mirrored = np.block([[base, base[:, ::-1]], [base[::-1, :], base[::-1, ::-1]]])
rng = np.random.default_rng(4)
mirrored = (mirrored + 0.03 * rng.standard_normal(mirrored.shape)).astype(np.float32).astype(np.float64)
for name, z in (("plain", plain), ("mirrored", mirrored)):
    single, g1 = gpu(z)
    tiled, g2 = gpu(z, max_fft=2048)
    rep = stack_report(tiled, single)
    print(name, "tiled(P=%d x%d) vs single(P=%d):" % (g2["Py"], g2["tiles_y"], g1["Py"]), json.dumps({k: rep[k] for k in KEEP}), flush=True)
    seam = g2["tiles_y"] and (4096 // g2["tiles_y"])
    for label, c0 in (("crop at mirror line/centre", 2048 - size // 2), ("crop at tile seam", seam - size // 2), ("crop elsewhere", 700)):
        ref = oracle(z[c0:c0 + size, c0:c0 + size])
        for nm, res in (("single", single), ("tiled", tiled)):
            rep = stack_report(res[:, c0 + m:c0 + size - m, c0 + m:c0 + size - m], ref[:, m:-m, m:-m])
            print(name, label, c0, nm, json.dumps({k: rep[k] for k in KEEP}), flush=True)
