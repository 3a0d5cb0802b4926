#!/bin/bash
# Round 2, GPU call M (1 GPU): gbuf store-path change -- tests, bench C3/C2/C4 + a tiled case, then sanitizers.
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/m_tests.log 2>&1; tail -3 gpurun_out/m_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/m_bench_c3.json 2> gpurun_out/m_bench_c3.err
for w in c1 c2 c4; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/m_bench_$w.json 2> gpurun_out/m_bench_$w.err
done
timeout 600 python bench.py --workload ns --ages 3 --steps 1 --warmup 1 --no-cpu-baseline --no-dropin > gpurun_out/m_bench_ns3.json 2> gpurun_out/m_bench_ns3.err
for f in gpurun_out/m_bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(' value %.0f ms %.2f e2e %s dropin %s' % (d['value'], d['ms_per_step'], (d.get('e2e') or {}).get('value'), (d.get('e2e_dropin') or {}).get('warm_value'))); print(' ', d.get('kernel_ms_per_step'))
except Exception as e: print(' ??', e); print(open('$f'.replace('.json','.err')).read()[-800:])
"; done
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python scratch/gpu_race.py > gpurun_out/l_racecheck.log 2>&1
tail -5 gpurun_out/l_racecheck.log
timeout 420 compute-sanitizer --tool memcheck python scratch/gpu_race.py > gpurun_out/l_memcheck.log 2>&1
tail -4 gpurun_out/l_memcheck.log
timeout 420 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -k "search_vs_oracle or nan_in_dem or plugin_template or err_mask or noise_level or spatial or serial" > gpurun_out/l_memcheck_tests.log 2>&1
tail -4 gpurun_out/l_memcheck_tests.log
