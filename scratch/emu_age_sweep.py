import sys; sys.path.insert(0,'.')
import numpy as np, time
from tests.emu.build_emu import build
from scarplet_b200 import _lib
lib = _lib.open_library(build()); _lib._use_library(lib)
import scarplet_b200 as sb
from scarplet_b200.WindowedTemplate import Scarp
from tests.parity import stack_report
prec = int(sys.argv[1])
sb.configure(precision=prec)
z = np.load('tests/golden/synthetic_dem_f32.npy').astype(np.float64)
gold = np.load('tests/golden/reference_goldens.npz')['synthetic_match1']
t0=time.time(); res = sb.match(sb.DEMGrid(z,1.0), Scarp, scale=100, ang_max=np.pi/2, ang_min=-np.pi/2); print('precision',prec,'%.0fs'%(time.time()-t0))
print(stack_report(np.stack(res), gold))
print('allclose', [bool(np.allclose(res[i],gold[i])) for i in range(4)]); np.save('/tmp/res_age_%d.npy'%prec, np.stack(res))
