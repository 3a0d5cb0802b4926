"""Timing of the two ways a rank of an orientation-sharded search gets the raster onto its
device (under torchrun): Plan.set_dem(host) -- the whole raster over PCIe -- against
distributed.set_dem_sharded -- 1 / world of the rows over PCIe, all-gather for the rest --
with the parts of the latter timed separately.  Host wall clock around device synchronisation."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from scarplet_b200 import distributed as D  # noqa: E402
from scarplet_b200.engine import Plan  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    n = 4096
    z = torch.from_numpy(np.random.default_rng(1).normal(size=(n, n))).pin_memory()
    stream = torch.cuda.Stream(device=device)

    def timed(fn, reps=5):
        out = []
        for _ in range(reps + 2):
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            out.append((time.perf_counter() - t0) * 1e3)
        return float(np.median(out[2:]))

    with torch.cuda.stream(stream):
        with Plan(n, n, 1.0, 1.0, device=local, stream=stream.cuda_stream) as plan:
            res = {"set_dem(host, whole raster)": timed(lambda: plan.set_dem(z.numpy())),
                   "set_dem_sharded": timed(lambda: D.set_dem_sharded(plan, z, device))}
            per = -(-n // world)
            buf = torch.zeros((world * per, n), dtype=torch.float64, device=device)
            lo, hi = rank * per, min((rank + 1) * per, n)
            res["  H2D of my rows"] = timed(lambda: buf[lo:hi].copy_(z[lo:hi], non_blocking=True))
            res["  all_gather_into_tensor (in place)"] = timed(
                lambda: dist.all_gather_into_tensor(buf.view(-1), buf[rank * per:(rank + 1) * per].view(-1)))
            tmp = torch.empty_like(buf)
            res["  all_gather_into_tensor (separate output)"] = timed(
                lambda: dist.all_gather_into_tensor(tmp.view(-1), buf[rank * per:(rank + 1) * per].view(-1)))
            res["  set_dem_device"] = timed(lambda: plan.set_dem_device(buf.data_ptr()))
            res["  H2D whole raster (torch copy_)"] = timed(lambda: buf[:n].copy_(z, non_blocking=True))
    print("rank %d of %d: %s" % (rank, world, {k: round(v, 2) for k, v in res.items()}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
