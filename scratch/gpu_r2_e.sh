#!/bin/bash
# Round 2, GPU call E (1 GPU): parity of the two large workloads against the oracle (north-star
# full 30-age search on a non-mirrored 16384^2 DEM; C5 band seam), bench with the new fit kernel.
cd "$(dirname "$0")/.."
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/e_bench_c3.json 2> gpurun_out/e_bench_c3.err
timeout 1500 python scratch/big_parity.py ns c5 > gpurun_out/e_big_parity.json 2> gpurun_out/e_big_parity.err
python -c "
import json
d=json.loads(open('gpurun_out/e_bench_c3.json').read().strip().splitlines()[-1]); print(' value %.0f ms %.2f' % (d['value'], d['ms_per_step'])); print({k: round(v['ms_per_step'],1) for k,v in d['roofline']['kernels'].items()})
"
cat gpurun_out/e_big_parity.json; tail -5 gpurun_out/e_big_parity.err
