import sys, os; sys.path.insert(0,'.')
import numpy as np
import scarplet_b200 as sl
from scarplet_b200.synth import synthetic_dem
from scarplet_b200.WindowedTemplate import Channel
from scarplet_b200.engine import Plan
from scarplet_b200 import params as P
z = synthetic_dem(257, seed=255, nx=255)
nang = int(sys.argv[1]) if len(sys.argv) > 1 else 181
angles = P.search_angles(-np.pi/2, np.pi/2)[:nang]
outs = []
with Plan(257, 255, 1.0, 1.0) as plan:
    plan.set_dem(z)
    a, t, age_of, angle_of = plan.build_sweep(Channel._sb_spec, 8, [0.15], angles)
    for it in range(3):
        plan.reset(); plan.sweep(a, t); outs.append(plan.finalize(age_of, angle_of))
print('deterministic:', np.array_equal(outs[0], outs[1]), np.array_equal(outs[1], outs[2]), 'diff px', (outs[0][3] != outs[1][3]).sum())
