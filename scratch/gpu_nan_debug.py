import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scarplet_b200 as sl
from scarplet_b200.WindowedTemplate import Scarp
from scarplet_b200.synth import synthetic_dem
from oracle import scarplet_oracle as O
shape = (200, 333)
z = synthetic_dem(shape[0], seed=7, nx=shape[1])
z[shape[0] // 3, shape[1] // 2] = np.nan
res = sl.calculate_best_fit_parameters(sl.DEMGrid(z, 1.0), Scarp, 16, 5.0)
ref = O.calculate_best_fit_parameters(z, 1.0, 1.0, O.SCARP, 16, 5.0, processes=8)
for p in (0, 3):
    a, b = np.isnan(res[p]), np.isnan(ref[p])
    print("plane", p, "nan res", a.sum(), "nan ref", b.sum(), "mismatch", (a != b).sum(), "zero res", (res[p] == 0).sum(), "zero ref", (ref[p] == 0).sum())
    ii, jj = np.nonzero(a != b)
    if len(ii):
        print("  rows", ii.min(), ii.max(), "cols", jj.min(), jj.max(), "first", list(zip(ii[:5], jj[:5])), "res", res[p][ii[:5], jj[:5]], "ref", ref[p][ii[:5], jj[:5]])
print("age eq", np.array_equal(res[1], ref[1]), "ang eq", np.array_equal(res[2], ref[2]), np.unique(res[1]), np.unique(res[2])[:5])
