#!/bin/bash
# like run_variants.sh but prints every kernel's time
for spec in "$@"; do
  v=${spec%%:*}; envs=""
  if [[ "$spec" == *:* ]]; then envs=$(echo "${spec#*:}" | tr ',' ' '); fi
  cp scratch/variants/$v.so scarplet_b200/libscarplet_b200.so
  env $envs python bench.py --steps 2 --warmup 2 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['kernel_ms_per_step']
print('$spec', {a: round(b,1) for a,b in k.items()})"
done
