#!/bin/bash
# Round 2, GPU call I (2 GPUs): GPU tests on one device, then the NCCL consistency check and the
# 2-GPU bench line with template-balanced shares.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/parity
export SB_PARITY_DIR=gpurun_out/parity
( time CUDA_VISIBLE_DEVICES=0 timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/i_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/i_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 scratch/mgpu_check.py > gpurun_out/i_mgpu_check.log 2>&1
echo "mgpu_check rc=$?" >> gpurun_out/i_mgpu_check.log
timeout 600 $TR --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/i_bench_c3_g2.json 2> gpurun_out/i_bench_c3_g2.err
tail -6 gpurun_out/i_tests.log; grep -v "^$" gpurun_out/i_mgpu_check.log | tail -8 | cut -c1-220
python -c "
import json
d=json.loads(open('gpurun_out/i_bench_c3_g2.json').read().strip().splitlines()[-1]); print(' value %.0f ms %.2f merge %s e2e %s' % (d['value'], d['ms_per_step'], d['merge_ms_per_step'], (d.get('e2e') or {}).get('value')))
"
