import sys; sys.path.insert(0,'.')
import numpy as np
import scarplet_b200 as sl
from scarplet_b200.synth import synthetic_dem
from scarplet_b200.WindowedTemplate import Channel
from scarplet_b200.engine import Plan
from oracle import scarplet_oracle as O
z = synthetic_dem(257, seed=255, nx=255)
angles = O.search_angles()
bad = 0
with Plan(257, 255, 1.0, 1.0) as plan:
    plan.set_dem(z)
    for ai in range(0, 181, 9):
        a = angles[ai]
        t = plan.render_template(Channel._sb_spec, 8, 0.15, a)
        ref = O.template_array(O.RICKER, 8, 0.15, a, 255, 257, 1.0)
        nm = ((t != 0) != (ref != 0)).sum()
        amp, snr = plan.match_template(Channel._sb_spec, 8, 0.15, a)
        ramp, _, _, rsnr = O.match_template(z, 1., 1., O.RICKER, 8, 0.15, a)
        rel = np.abs(snr - rsnr) / rsnr
        print(ai, 'M mismatch', nm, 'n', (ref != 0).sum(), 'val err', np.abs(t - ref).max(), 'snr rel p50/p99/max', np.median(rel), np.quantile(rel, .99), rel.max(), 'amp err', np.abs(amp - ramp).max() / np.abs(ramp).max())
