#!/bin/bash
# developer ablation: time the per-template kernels with parts of their memory traffic switched off
for d in "$@"; do
  SB_DBG=$d timeout 200 python bench.py --ages 4 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_ms_per_step']
print('SB_DBG=$d conv %.1f fit %.1f total %.1f'%(k['k_conv_cols'],k['k_fit_rows'],d['ms_per_step']))"
done
