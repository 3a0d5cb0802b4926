#!/bin/bash
# Round 2, GPU call Q (1 GPU): the kept measurements of the final build -- bench lines (ours and the
# reference arm), launch list, ncu --set full of the dominant kernels at the real C3 batch.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/parity; export SB_PARITY_DIR=gpurun_out/parity
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/q_tests.log 2>&1; tail -3 gpurun_out/q_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/q_bench_c3.json 2> gpurun_out/q_bench_c3.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/q_bench_c3_reference.json 2> gpurun_out/q_bench_c3_reference.err
for w in c1 c2 c4; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/q_bench_$w.json 2> gpurun_out/q_bench_$w.err
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/q_launches.csv \
  python bench.py --steps 1 --warmup 1 --e2e-steps 1 --profile-steps 0 --no-cpu-baseline --no-dropin > gpurun_out/q_launch_run.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_conv_cols_r|k_fit_rows_g' -s 20 -c 2 \
  -o /tmp/q_prof -f python bench.py --steps 1 --warmup 1 --e2e-steps 1 --profile-steps 0 --no-cpu-baseline --no-dropin > gpurun_out/q_ncu.log 2>&1
ncu -i /tmp/q_prof.ncu-rep --page raw --csv > gpurun_out/q_prof_raw.csv 2>/dev/null
ncu -i /tmp/q_prof.ncu-rep --page source --csv --print-source sass > gpurun_out/q_prof_source.csv 2>/dev/null
gzip -f gpurun_out/q_prof_source.csv
timeout 600 ncu --set full --clock-control none -k regex:'k_combine_spectra|k_diff_rows_f|k_tmpl_rows|k_curv_cols' -s 2 -c 5 \
  -o /tmp/q_prof2 -f python bench.py --steps 1 --warmup 1 --e2e-steps 1 --profile-steps 0 --no-cpu-baseline --no-dropin > gpurun_out/q_ncu2.log 2>&1
ncu -i /tmp/q_prof2.ncu-rep --page raw --csv > gpurun_out/q_prof2_raw.csv 2>/dev/null
du -sh gpurun_out
for f in gpurun_out/q_bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(' value %.0f ms %.2f e2e %s dropin %s cpu %s' % (d['value'], d['ms_per_step'], (d.get('e2e') or {}).get('value'), (d.get('e2e_dropin') or {}).get('warm_value'), (d.get('cpu_baseline') or {}).get('value')))
except Exception as e: print(' ??', e); print(open('$f'.replace('.json','.err')).read()[-800:])
"; done
