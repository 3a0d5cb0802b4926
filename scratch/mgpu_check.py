"""Multi-GPU consistency (run under torchrun, one rank per GPU), on real devices over NCCL:

* orientation-sharded search, replicated merge (two all-reduces) and banded merge (all-to-all +
  device fold): both must equal the single-GPU search bit for bit -- with the plan on its own
  stream (fenced hand-overs) and on torch's current stream;
* row-sharded search (spatial sharding with halos, no data-path collective): the bands put
  together must equal the whole-raster search to rounding, DEM uploaded from the host and
  generated on the device.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 scratch/mgpu_check.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from parity import stack_report  # noqa: E402
from scarplet_b200 import distributed as D  # noqa: E402
from scarplet_b200 import params as P  # noqa: E402
from scarplet_b200.engine import Plan  # noqa: E402
from scarplet_b200.synth import synthetic_dem  # noqa: E402
from scarplet_b200.templates import Channel, Scarp  # noqa: E402


def gather_bands(band, lo, hi, ny, nx, device, rank, world):
    """All bands on rank 0 as one (4, ny, nx) array (test helper: through the device)."""
    full = torch.zeros((4, ny, nx), dtype=torch.float64, device=device)
    full[:, lo:hi] = torch.from_numpy(band).to(device)
    dist.all_reduce(full, op=dist.ReduceOp.SUM)
    return full.cpu().numpy()


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    angles = P.search_angles(-np.pi / 2, np.pi / 2)
    ok = True
    report = {}
    for name, spec, n, scale, ages in (("Scarp", Scarp._sb_spec, 1024, 50, [3.0, 30.0, 300.0]),
                                       ("Channel", Channel._sb_spec, 701, 10, [0.1])):
        z = synthetic_dem(n, seed=11)
        for own_stream in (False, True):
            stream = torch.cuda.Stream(device=device)
            with torch.cuda.stream(stream):
                kw = {} if own_stream else {"stream": stream.cuda_stream}
                with Plan(n, n, 1.0, 1.0, device=local, **kw) as plan:
                    plan.set_dem(z)
                    merged = D.sharded_search(plan, spec, scale, ages, angles, device=device)
                    lo, hi, band = D.sharded_search(plan, spec, scale, ages, angles, device=device, merge="bands")
                    a, t, age_of, angle_of = plan.build_sweep(spec, scale, ages, angles)
                    plan.reset()
                    plan.sweep(a, t)
                    single = plan.finalize(age_of, angle_of)
                    # DEM uploaded 1 / world per rank + all-gather (twice: the device buffer is reused)
                    for _ in range(2):
                        D.set_dem_sharded(plan, z, device)
                        lo2, hi2, band2 = D.sharded_search(plan, spec, scale, ages, angles, device=device, merge="bands")
                banded = gather_bands(band, lo, hi, n, n, device, rank, world)
                banded2 = gather_bands(band2, lo2, hi2, n, n, device, rank, world)
            same = np.array_equal(merged, single)
            same_b = np.array_equal(banded, single)
            same_u = np.array_equal(banded2, single)
            ok &= same and same_b and same_u
            print("rank %d %s %dx%d own_stream=%s: replicated == single: %s, banded == single: %s, "
                  "banded after the shared upload == single: %s (valid px %d)"
                  % (rank, name, n, n, own_stream, same, same_b, same_u, int((single[3] > 0).sum())), flush=True)
        # rows sharded: host DEM
        plan, (lo, hi) = D.spatial_plan(n, n, 1.0, 1.0, spec, scale, ages, angles, device=local)
        with plan:
            plan.set_dem(z)
            D.share_dem_stats(plan, device=device)
            lo, hi, band = D.spatial_search(plan, spec, scale, ages, angles)
            mem = plan.device_bytes
        spatial = gather_bands(band, lo, hi, n, n, device, rank, world)
        rep = stack_report(spatial, single, odd_template=(name == "Scarp"))
        good = rep["mask_equal"] and rep["index_agreement"] >= 0.9995 and rep["snr_rel_max_strong"] < 5e-5
        ok &= bool(good)
        report[name] = {k: rep[k] for k in ("valid", "mask_equal", "index_agreement", "snr_rel_max_strong", "amp_rel_max_strong")}
        print("rank %d %s rows %d..%d of %d: spatial vs whole: %s %s, plan bytes %.1f MB"
              % (rank, name, lo, hi, n, good, report[name], mem / 1e6), flush=True)
    # device-generated DEM (bench.py's generator for the large rasters): bands vs whole raster
    sys.path.insert(0, ROOT)
    from bench import device_dem_rows
    n, ages, scale = 2048, [3.0, 300.0], 50
    spec = Scarp._sb_spec
    plan, (lo, hi) = D.spatial_plan(n, n, 1.0, 1.0, spec, scale, ages, angles[::4], device=local)
    with plan:
        r0, nrows = plan.dem_rows()
        zd = device_dem_rows(n, 4, r0, nrows, device)
        torch.cuda.synchronize()
        plan.set_dem_device(zd.data_ptr())
        D.share_dem_stats(plan, device=device)
        lo, hi, band = D.spatial_search(plan, spec, scale, ages, angles[::4])
    spatial = gather_bands(band, lo, hi, n, n, device, rank, world)
    zfull = device_dem_rows(n, 4, 0, n, device)
    # the band every rank generated is the same rows of the whole raster
    rows = (r0 + torch.arange(nrows, device=device)) % n
    same_dem = bool(torch.equal(zd, zfull[rows]))
    with Plan(n, n, 1.0, 1.0, device=local) as plan:
        plan.set_dem(zfull.cpu().numpy())
        a, t, age_of, angle_of = plan.build_sweep(spec, scale, ages, angles[::4])
        plan.reset()
        plan.sweep(a, t)
        single = plan.finalize(age_of, angle_of)
    rep = stack_report(spatial, single)
    good = same_dem and rep["mask_equal"] and rep["index_agreement"] >= 0.9995 and rep["snr_rel_max_strong"] < 5e-5
    ok &= bool(good)
    print("rank %d device DEM 2048: band == rows of the whole raster: %s; spatial vs whole: %s agreement %.6f strong max %.2e"
          % (rank, same_dem, good, rep["index_agreement"], rep["snr_rel_max_strong"]), flush=True)
    flag = torch.tensor([1 if ok else 0], device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "all_ok": bool(flag.item()), "spatial": report}), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
