import sys, os; sys.path.insert(0,'.')
import numpy as np
import scarplet_b200 as sl
from scarplet_b200.synth import synthetic_dem
from scarplet_b200.WindowedTemplate import Channel
from oracle import scarplet_oracle as O
from tests.parity import stack_report
z = synthetic_dem(257, seed=255, nx=255)
res = sl.calculate_best_fit_parameters(sl.DEMGrid(z, 1.0), Channel, 8, 0.15)
ref = O.calculate_best_fit_parameters(z, 1.0, 1.0, O.RICKER, 8, 0.15, processes=8)
ref1 = O.calculate_best_fit_parameters(z, 1.0, 1.0, O.RICKER, 8, 0.15, processes=1)
print('oracle pool vs serial equal', np.array_equal(ref, ref1))
print(stack_report(res, ref, odd_template=False))
print(stack_report(res, ref1, odd_template=False))
os.makedirs('gpurun_out', exist_ok=True)
np.savez_compressed('gpurun_out/channel_debug.npz', z=z, res=res, ref=ref, ref1=ref1)
