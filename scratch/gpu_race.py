"""Small age sweeps through every pipelined kernel, for compute-sanitizer --tool racecheck /
memcheck: radix-16 persistent column kernel (Py = 1024), radix-64 column kernel (Py = 4096: pairs
of two-warp groups behind named barriers, halves swapped through the exchange buffers), padded
tiles, fit sub-streams, two best states, a row slab, get_err_mask templates."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scarplet_b200 import params as P
from scarplet_b200.engine import Plan
from scarplet_b200.synth import synthetic_dem
from scarplet_b200.templates import LeftFacingUpperBreakScarp, Scarp

ANGLES = P.search_angles(-np.pi / 2, np.pi / 2)
for shape, scale, step, kw in (((1024, 256), 12, 16, {}), ((300, 260), 12, 16, {}), ((4096, 160), 12, 45, {}),
                               ((4096, 160), [8, 200], 60, {"states": 2}), ((600, 256), 12, 30, {"slab": (150, 400, 40)})):
    z = synthetic_dem(shape[0], seed=3, nx=shape[1])
    angles = ANGLES[::step]
    for spec in (Scarp._sb_spec, LeftFacingUpperBreakScarp._sb_spec):
        with Plan(shape[0], shape[1], 1.0, 1.0, **kw) as plan:
            plan.set_dem(z)
            a, t, age_of, angle_of = plan.build_sweep(spec, scale, [2.0, 5.0, 10.0, 20.0, 40.0], angles)
            plan.reset()
            plan.sweep(a, t)
            out = plan.finalize(age_of, angle_of)
            print(shape, kw, plan.last_geometry(), float(np.nanmax(out[3])), flush=True)
