import sys; sys.path.insert(0,'.')
import numpy as np
import scarplet_b200 as sl
from scarplet_b200.synth import synthetic_dem
from scarplet_b200.WindowedTemplate import Channel
from scarplet_b200.engine import Plan
from scarplet_b200 import params as P
from oracle import scarplet_oracle as O
z = synthetic_dem(257, seed=255, nx=255)
angles = O.search_angles()
spec = Channel._sb_spec
def sweep(plan, angs):
    a, t, age_of, angle_of = plan.build_sweep(spec, 8, [0.15], angs)
    plan.reset(); plan.sweep(a, t); return plan.finalize(age_of, angle_of)
for ws in (8192, 64):
    with Plan(257, 255, 1.0, 1.0, workspace_mb=ws) as plan:
        plan.set_dem(z)
        # single-template sweeps vs raw
        for ai in (0, 37, 90):
            out = sweep(plan, angles[ai:ai+1])
            amp, snr = plan.match_template(spec, 8, 0.15, angles[ai])
            print('ws', ws, 'angle', ai, 'sweep-vs-raw snr rel max', (np.abs(out[3]-snr)/np.maximum(snr,1e-30)).max(), 'amp', np.abs(out[0]-amp).max(), plan.last_geometry())
        # multi-template sweep: check each pixel's snr equals the raw snr of the chosen angle
        angs = angles[30:46]
        out = sweep(plan, angs)
        raws = np.stack([plan.match_template(spec, 8, 0.15, a)[1] for a in angs])
        best = raws.max(0)
        print('ws', ws, 'multi sweep vs max of raws: rel max', (np.abs(out[3]-best)/best).max(), 'argmax agree', (np.argmax(raws,0) == np.round((out[2]-angs[0])/(np.pi/180)).astype(int)).mean(), plan.last_geometry())
        for k, a in enumerate(angs[:16]):
            sel = np.isclose(out[2], a)
            if sel.any(): print('   angle', 30+k, 'chosen px', sel.sum(), 'rel err vs its raw', (np.abs(out[3]-raws[k])/raws[k])[sel].max())
