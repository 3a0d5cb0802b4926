#!/bin/bash
# Round 2, GPU call P (2 GPUs): e2e of the orientation-sharded search with the whole-raster
# upload per rank against the shared upload (SB_BENCH_SHARED_UPLOAD), final build.
cd "$(dirname "$0")/.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for u in 0 1 0 1; do
  SB_BENCH_SHARED_UPLOAD=$u timeout 300 $TR --master-port 2954$u bench.py --gpus 2 --steps 2 --warmup 3 --e2e-steps 4 --profile-steps 0 > gpurun_out/p_bench_c3_g2_u$u.json 2> gpurun_out/p_bench_c3_g2_u$u.err
  python -c "
import json
d=json.loads(open('gpurun_out/p_bench_c3_g2_u$u.json').read().strip().splitlines()[-1]); print('upload=$u value %.0f ms %.2f e2e %.0f ms %.2f merge %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['merge_ms_per_step']))"
done
