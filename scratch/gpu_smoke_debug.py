import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import scarplet_b200 as sl
from scarplet_b200.WindowedTemplate import Scarp
from scarplet_b200.synth import synthetic_dem
from oracle import scarplet_oracle as O
from parity import stack_report
z = synthetic_dem(256, seed=5)
grid = sl.DEMGrid(z, 1.0)
res = sl.match(grid, Scarp, scale=20, age=10., ang_min=-np.pi / 2, ang_max=np.pi / 2)
ref = O.match(z, 1.0, 1.0, O.SCARP, scale=20, age=10., ang_min=-np.pi / 2, ang_max=np.pi / 2)
rep = stack_report(res, ref)
print(json.dumps({k: rep[k] for k in ("valid", "tie_reset_pixels", "mask_mismatch_unexplained", "mask_mismatch", "index_agreement", "snr_rel_max")}))
d = (res[3] > 0) != (ref[3] > 0)
ii, jj = np.nonzero(d)
print("mismatch", d.sum(), list(zip(ii[:6].tolist(), jj[:6].tolist())), res[3][d][:6], ref[3][d][:6], res[2][d][:6])
