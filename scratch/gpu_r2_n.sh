#!/bin/bash
# Round 2, GPU call N (2 GPUs): row kernel with early input requests (A/B by SB_FIT_EARLY), the
# NCCL consistency check incl. the shared DEM upload, the 2-GPU bench line.
cd "$(dirname "$0")/.."
for e in 0 1; do
  CUDA_VISIBLE_DEVICES=0 SB_FIT_EARLY=$e timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/n_bench_c3_early$e.json 2> gpurun_out/n_bench_c3_early$e.err
done
for e in 0 1; do
  CUDA_VISIBLE_DEVICES=0 SB_FIT_EARLY=$e timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/n_bench_c2_early$e.json 2> gpurun_out/n_bench_c2_early$e.err
  CUDA_VISIBLE_DEVICES=0 SB_FIT_EARLY=$e timeout 600 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/n_bench_c4_early$e.json 2> gpurun_out/n_bench_c4_early$e.err
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 scratch/mgpu_check.py > gpurun_out/n_mgpu_check.log 2>&1
echo "mgpu_check rc=$?" >> gpurun_out/n_mgpu_check.log
timeout 600 $TR --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/n_bench_c3_g2.json 2> gpurun_out/n_bench_c3_g2.err
grep -v "^$" gpurun_out/n_mgpu_check.log | tail -8 | cut -c1-260
for f in gpurun_out/n_bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); r=d['roofline']
    print(' value %.0f ms %.2f e2e %s merge %s' % (d['value'], d['ms_per_step'], (d.get('e2e') or {}).get('value'), d.get('merge_ms_per_step')))
    print(' ', {k:round(v['ms_per_step'],2) for k,v in r.get('kernels',{}).items()})
except Exception as e: print(' ??', e); print(open('$f'.replace('.json','.err')).read()[-800:])
"; done
