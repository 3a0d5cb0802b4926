import sys, time, os
sys.path.insert(0, '.')
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
from scarplet_b200.engine import Plan
from scarplet_b200 import params as P
from scarplet_b200.templates import Scarp
from scarplet_b200.synth import synthetic_dem
print(torch.cuda.get_device_name(0))
for n, na, ages in ((1024, 181, [10.0]), (4096, 16, list(np.logspace(0, 3.5, 4))), (4096, 16, [10.0])):
    z = synthetic_dem(n, seed=2)
    st = torch.cuda.Stream()
    with Plan(n, n, 1.0, 1.0, stream=st.cuda_stream) as plan:
        plan.set_dem(z)
        angles = P.search_angles(-np.pi/2, np.pi/2)[:na]
        a, t, age_of, angle_of = plan.build_sweep(Scarp._sb_spec, 100, ages, angles)
        for it in range(3):
            plan.reset()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(st); plan.sweep(a, t); e1.record(st); st.synchronize()
            ms = e0.elapsed_time(e1)
            evals = n * n * len(angles) * len(ages)
            print(n, na, len(ages), 'ms %.2f' % ms, 'Gpx-evals/s %.2f' % (evals / ms / 1e6), plan.last_geometry(), 'launches', plan.launches)
        out = plan.finalize(age_of, angle_of)
        print('valid', (out[3] > 0).sum(), 'snr max', out[3].max())
