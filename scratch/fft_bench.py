"""Developer micro-benchmark: the batched length-4096 FFT kernel alone, radix-16 core (3 stages,
2 exchanges, 256 threads per transform) against the radix-64 core (2 passes, 1 exchange, 64
threads per transform).  Reads + writes 2 * 8 B per point from HBM."""
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__
__graft_entry__.build()
from scarplet_b200.engine import Plan
out = {}
with Plan(16, 16, 1.0, 1.0) as plan:
    for rows in (4096, 16384):
        for r64 in (False, True):
            ms = plan.debug_fft_bench(4096, rows, reps=20, radix64=r64)
            pts = rows * 4096
            out["rows%d_%s" % (rows, "r64" if r64 else "r16")] = {"ms": ms, "Gpts_per_s": pts / ms / 1e6, "GBps": pts * 16 / ms / 1e6}
print(json.dumps(out, indent=1))
