"""Small age sweep through every pipelined kernel (persistent column kernel included), for
compute-sanitizer --tool racecheck / memcheck."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scarplet_b200 import params as P
from scarplet_b200.engine import Plan
from scarplet_b200.synth import synthetic_dem
from scarplet_b200.templates import Scarp

for shape in ((1024, 256), (300, 260)):
    z = synthetic_dem(shape[0], seed=3, nx=shape[1])
    angles = P.search_angles(-np.pi / 2, np.pi / 2)[::16]
    with Plan(shape[0], shape[1], 1.0, 1.0) as plan:
        plan.set_dem(z)
        a, t, age_of, angle_of = plan.build_sweep(Scarp._sb_spec, 12, [2.0, 5.0, 10.0, 20.0, 40.0], angles)
        plan.reset()
        plan.sweep(a, t)
        out = plan.finalize(age_of, angle_of)
        print(shape, plan.last_geometry(), float(np.nanmax(out[3])))
