import sys; sys.path.insert(0,'.')
import numpy as np
import scarplet_b200 as sl
from scarplet_b200.synth import synthetic_dem
from scarplet_b200.WindowedTemplate import Channel
from scarplet_b200.engine import Plan
from oracle import scarplet_oracle as O
z = synthetic_dem(257, seed=255, nx=255)
angles = O.search_angles()
spec = Channel._sb_spec
res = sl.calculate_best_fit_parameters(sl.DEMGrid(z, 1.0), Channel, 8, 0.15)
with Plan(257, 255, 1.0, 1.0) as plan:
    plan.set_dem(z)
    raws = np.stack([plan.match_template(spec, 8, 0.15, a)[1] for a in angles])
    a, t, age_of, angle_of = plan.build_sweep(spec, 8, [0.15], angles)
    plan.reset(); plan.sweep(a, t); res2 = plan.finalize(age_of, angle_of); print(plan.last_geometry())
print('api vs plan sweep equal', np.array_equal(res, res2))
idx = np.round((res[2] + np.pi/2)/(np.pi/180)).astype(int)
chosen_raw = np.take_along_axis(raws, idx[None], 0)[0]
rel = np.abs(res[3] - chosen_raw)/chosen_raw
print('sweep snr vs raw snr of chosen angle: p50 %.2e p99 %.2e max %.2e frac>1e-4 %.4f' % (np.median(rel), np.quantile(rel,.99), rel.max(), (rel>1e-4).mean()))
bad = rel > 1e-4
print('bad by angle idx:', np.bincount(idx[bad], minlength=181).nonzero()[0][:40], np.bincount(idx[bad], minlength=181).max())
ys, xs = np.nonzero(bad); print('bad rows hist', np.histogram(ys, bins=8, range=(0,257))[0], 'cols', np.histogram(xs, bins=8, range=(0,255))[0])
best = raws.max(0); print('sweep best vs max raws rel max', (np.abs(res[3]-best)/best).max(), 'argmax agree', (raws.argmax(0)==idx).mean())
