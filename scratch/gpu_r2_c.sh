#!/bin/bash
# Round 2, GPU call C (2 GPUs): NCCL paths on real devices -- banded / replicated merge equality,
# spatial sharding equality, then the 2-GPU bench lines.
cd "$(dirname "$0")/.."
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 scratch/mgpu_check.py > gpurun_out/c_mgpu_check.log 2>&1
echo "mgpu_check rc=$?" >> gpurun_out/c_mgpu_check.log
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/c_bench_c3_g$N.json 2> gpurun_out/c_bench_c3_g$N.err
timeout 600 $TR --master-port 29513 bench.py --gpus $N --workload ns --ages 3 --steps 1 --warmup 1 > gpurun_out/c_bench_ns3_g$N.json 2> gpurun_out/c_bench_ns3_g$N.err
timeout 600 $TR --master-port 29514 bench.py --gpus $N --workload c5 --size 16384 --ages 2 --steps 1 --warmup 1 > gpurun_out/c_bench_c5small_g$N.json 2> gpurun_out/c_bench_c5small_g$N.err
timeout 600 $TR --master-port 29515 bench.py --gpus $N --workload c4 --steps 2 --warmup 2 > gpurun_out/c_bench_c4_g$N.json 2> gpurun_out/c_bench_c4_g$N.err
tail -12 gpurun_out/c_mgpu_check.log
for f in gpurun_out/c_bench_*_g$N.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(' value %.0f ms %.2f wall %.2f merge %.2f mem %.1f GB e2e %s tiles %s' % (d['value'], d['ms_per_step'], d['host_wall_ms_per_step'], d['merge_ms_per_step'], d['device_gb_per_rank_max'], (d.get('e2e') or {}).get('value'), d['config']['tiles']))
except Exception as e: print(' ??', e); print(open('$f'.replace('.json','.err')).read()[-1500:])
"; done
