#!/bin/bash
# usage (on the GPU box): scratch/run_variants.sh v1 v2 ...   -- benches each scratch/variants/<v>.so
# a variant may be given as name:ENV=VAL,ENV2=VAL2
for spec in "$@"; do
  v=${spec%%:*}; envs=""
  if [[ "$spec" == *:* ]]; then envs=$(echo "${spec#*:}" | tr ',' ' '); fi
  cp scratch/variants/$v.so scarplet_b200/libscarplet_b200.so
  env $envs python bench.py --steps 2 --warmup 2 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['kernel_ms_per_step']
print('$spec value %.0f ms %.1f conv %.1f fit %.1f'%(d['value'],d['ms_per_step'],k['k_conv_cols'],k['k_fit_rows']))"
done
