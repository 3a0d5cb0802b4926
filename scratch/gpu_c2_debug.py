import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import scarplet_b200 as sl
from scarplet_b200.WindowedTemplate import Channel
from scarplet_b200.synth import synthetic_dem
from oracle import scarplet_oracle as O
from parity import stack_report
KEEP = ("valid", "index_agreement", "snr_rel_max", "frac_snr_over_tol", "mask_mismatch_unexplained", "snr_rel_p50")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1201
relief = float(sys.argv[2]) if len(sys.argv) > 2 else 300.0
z = synthetic_dem(n, 1, relief=relief)
res = sl.calculate_best_fit_parameters(sl.DEMGrid(z, 1.0), Channel, 10, 0.1)
ref = O.calculate_best_fit_parameters(z, 1.0, 1.0, O.RICKER, 10, 0.1, processes=16)
rep = stack_report(res, ref, odd_template=False)
print("gpu vs full oracle:", json.dumps({k: rep[k] for k in KEEP}))
c0, size, m = n // 3, 500, 110
crop = z[c0:c0 + size, c0:c0 + size]
rc = O.calculate_best_fit_parameters(crop, 1.0, 1.0, O.RICKER, 10, 0.1, processes=16)
rep2 = stack_report(ref[:, c0 + m:c0 + size - m, c0 + m:c0 + size - m], rc[:, m:size - m, m:size - m], odd_template=False)
print("full oracle vs crop oracle (interior):", json.dumps({k: rep2[k] for k in KEEP}))
rep3 = stack_report(res[:, c0 + m:c0 + size - m, c0 + m:c0 + size - m], rc[:, m:size - m, m:size - m], odd_template=False)
print("gpu vs crop oracle (interior):", json.dumps({k: rep3[k] for k in KEEP}))
