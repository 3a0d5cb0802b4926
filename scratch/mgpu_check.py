"""Multi-GPU consistency (run under torchrun, one rank per GPU): the orientation-sharded
search merged over NCCL must equal the single-GPU search bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 scratch/mgpu_check.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from scarplet_b200 import distributed as D  # noqa: E402
from scarplet_b200 import params as P  # noqa: E402
from scarplet_b200.engine import Plan  # noqa: E402
from scarplet_b200.synth import synthetic_dem  # noqa: E402
from scarplet_b200.templates import Channel, Scarp  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    angles = P.search_angles(-np.pi / 2, np.pi / 2)
    ok = True
    for name, spec, n, scale, ages in (("Scarp", Scarp._sb_spec, 1024, 50, [3.0, 30.0, 300.0]),
                                       ("Channel", Channel._sb_spec, 700, 10, [0.1])):
        z = synthetic_dem(n, seed=11)
        stream = torch.cuda.Stream(device=device)
        with torch.cuda.stream(stream):
            with Plan(n, n, 1.0, 1.0, device=local, stream=stream.cuda_stream) as plan:
                plan.set_dem(z)
                merged = D.sharded_search(plan, spec, scale, ages, angles, device=device)
                a, t, age_of, angle_of = plan.build_sweep(spec, scale, ages, angles)
                plan.reset()
                plan.sweep(a, t)
                single = plan.finalize(age_of, angle_of)
        same = np.array_equal(merged, single)
        ok &= same
        print("rank %d %s %dx%d: merged == single-GPU: %s (valid px %d)"
              % (rank, name, n, n, same, int((single[3] > 0).sum())), flush=True)
    flag = torch.tensor([1 if ok else 0], device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_CHECK", "PASS" if flag.item() == 1 else "FAIL", flush=True)
    return 0 if flag.item() == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
