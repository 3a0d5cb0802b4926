import sys, os; sys.path.insert(0,'.')
import numpy as np
import scarplet_b200 as sl
from scarplet_b200.WindowedTemplate import Channel
from tests.parity import stack_report
d = np.load('scratch/channel_debug_ref.npz')
res = sl.calculate_best_fit_parameters(sl.DEMGrid(d['z'], 1.0), Channel, 8, 0.15)
rep = stack_report(res, d['ref'], odd_template=False)
print(sys.argv[1:], 'idx agree', rep['index_agreement'], 'snr_rel_max', rep['snr_rel_max'], 'frac>tol', rep['frac_snr_over_tol'])
