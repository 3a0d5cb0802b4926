#!/bin/bash
# Round 2, GPU call H (1 GPU): linear-combination set-up on hardware: tests + all single-GPU configs.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/parity
export SB_PARITY_DIR=gpurun_out/parity
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/h_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/h_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/h_bench_c3.json 2> gpurun_out/h_bench_c3.err
for w in c1 c2 c4; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/h_bench_$w.json 2> gpurun_out/h_bench_$w.err
done
tail -6 gpurun_out/h_tests.log
for f in gpurun_out/h_bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(' value %.0f ms %.2f e2e %s dropin %s' % (d['value'], d['ms_per_step'], (d.get('e2e') or {}).get('value'), (d.get('e2e_dropin') or {}).get('warm_value'))); print({k: round(v['ms_per_step'],2) for k,v in d['roofline']['kernels'].items()})
except Exception as e: print(' ??', e); print(open('$f'.replace('.json','.err')).read()[-800:])
"; done
