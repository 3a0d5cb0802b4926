#!/bin/bash
# Round 2, GPU call O (final build) (8 GPUs): consistency over NCCL at 8 ranks, then the multi-GPU workloads:
# C3 (driver's line), north-star 16384^2 (orientations sharded), C5 32768^2 (rows sharded), C4.
cd "$(dirname "$0")/.."
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 scratch/mgpu_check.py > gpurun_out/o_mgpu_check.log 2>&1
echo "mgpu_check rc=$?" >> gpurun_out/o_mgpu_check.log
timeout 300 $TR --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/o_bench_c3_g$N.json 2> gpurun_out/o_bench_c3_g$N.err
timeout 600 $TR --master-port 29513 bench.py --gpus $N --workload ns --steps 2 --warmup 3 --profile-steps 1 > gpurun_out/o_bench_ns_g$N.json 2> gpurun_out/o_bench_ns_g$N.err
timeout 900 $TR --master-port 29514 bench.py --gpus $N --workload c5 --steps 1 --warmup 3 --profile-steps 0 --e2e-steps 1 > gpurun_out/o_bench_c5_g$N.json 2> gpurun_out/o_bench_c5_g$N.err
timeout 300 $TR --master-port 29515 bench.py --gpus $N --workload c4 --steps 3 --warmup 3 > gpurun_out/o_bench_c4_g$N.json 2> gpurun_out/o_bench_c4_g$N.err
tail -4 gpurun_out/o_mgpu_check.log
for f in gpurun_out/o_bench_*_g$N.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(' value %.0f ms %.2f wall %.2f merge %s mem %.1f GB e2e %s tiles %s' % (d['value'], d['ms_per_step'], d['host_wall_ms_per_step'], d['merge_ms_per_step'], d['device_gb_per_rank_max'], (d.get('e2e') or {}).get('value'), d['config']['tiles']))
except Exception as e: print(' ??', e); print(open('$f'.replace('.json','.err')).read()[-1500:])
"; done
