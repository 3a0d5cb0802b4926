import sys; sys.path.insert(0,'.')
import numpy as np
import scarplet_b200 as sl
from scarplet_b200.synth import synthetic_dem
from scarplet_b200.WindowedTemplate import Channel
from oracle import scarplet_oracle as O
from tests.parity import stack_report
z = synthetic_dem(257, seed=255, nx=255)
res = sl.calculate_best_fit_parameters(sl.DEMGrid(z, 1.0), Channel, 8, 0.15)
ref = O.calculate_best_fit_parameters(z, 1.0, 1.0, O.RICKER, 8, 0.15, processes=8)
rep = stack_report(res, ref, odd_template=False); print(rep)
angles = O.search_angles()
S = np.stack([O.match_template(z,1.,1.,O.RICKER,8,0.15,a)[3] for a in angles])
srt = np.sort(S,axis=0); gap = (srt[-1]-srt[-2])/srt[-1]
dis = (ref[3]>0) & ~np.isclose(res[2], ref[2])
if dis.any():
    print('disagree', dis.sum(), 'gap pct', np.percentile(gap[dis],[0,50,90,99,100]))
    i_ref = S.argmax(0)
    ang_idx = np.round((res[2]+np.pi/2)/(np.pi/180)).astype(int)
    d = np.abs(ang_idx - i_ref)[dis]
    print('angle idx diff hist', np.bincount(d)[:6], (d>=179).sum())
    print('snr at disagree', np.percentile(ref[3][dis],[0,50,100]), 'all', np.percentile(ref[3][ref[3]>0],[0,50,100]))
    print('our snr vs ref snr at disagree rel', np.percentile((np.abs(res[3]-ref[3])/ref[3])[dis],[0,50,100]))
    ys,xs = np.nonzero(dis); print('rows', np.percentile(ys,[0,50,100]), 'cols', np.percentile(xs,[0,50,100]))
