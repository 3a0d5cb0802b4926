#!/bin/bash
# Round 2, GPU call S (N GPUs): C3 bench line of the final build with the e2e phase diagnostic,
# template shares of equal count against shares of equal estimated device time.
cd "$(dirname "$0")/.."
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for m in count cost; do
SB_SHARE_BALANCE=$m timeout 300 $TR --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --e2e-phases --profile-steps 0 > gpurun_out/s_bench_c3_g${N}_$m.json 2> gpurun_out/s_bench_c3_g${N}_$m.err
python -c "
import json
d=json.loads(open('gpurun_out/s_bench_c3_g${N}_$m.json').read().strip().splitlines()[-1]); print('$m: value %.0f ms %.2f merge %s e2e %.0f ms %.2f' % (d['value'], d['ms_per_step'], d['merge_ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])); print(d['e2e']['phases_ms_min_max_over_ranks'])" || tail -20 gpurun_out/s_bench_c3_g${N}_$m.err
done
