"""world_size-2 checks of the multi-GPU paths on CPU: gloo backend, the emulator build of
the kernels standing in for the two devices.

* orientations sharded across ranks, best states merged (a) replicated: all-reduce(MAX) on
  packed keys + all-reduce(SUM) on amplitudes, (b) banded: all-to-all of row bands + device
  fold -- both must equal the single-rank search bit for bit;
* rows sharded across ranks (spatial sharding, BASELINE config 5): no data-path collective,
  the bands put together must equal the whole-raster search (to rounding: other FFT domains).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SHAPE = (96, 128)
SCALE, AGES = 8, [2.0, 9.0]


def _inputs():
    from scarplet_b200 import params as P
    from scarplet_b200.synth import synthetic_dem
    z = synthetic_dem(SHAPE[0], seed=7, nx=SHAPE[1], relief=3.0)
    return z, P.search_angles(-np.pi / 2, np.pi / 2)[::9]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), SB_EMU_WORKERS="2")
    import torch
    import torch.distributed as dist
    from tests.emu.build_emu import build
    from scarplet_b200 import _lib
    _lib._use_library(_lib.open_library(build()))
    from scarplet_b200 import distributed as D
    from scarplet_b200.engine import Plan
    from scarplet_b200.templates import Scarp
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cpu = torch.device("cpu")
    z, angles = _inputs()
    ny, nx = SHAPE
    with Plan(ny, nx, 1.0, 1.0) as plan:
        plan.set_dem(z)
        out = D.sharded_search(plan, Scarp._sb_spec, SCALE, AGES, angles, "age_major", device=cpu)
        np.save(os.path.join(out_dir, "replicated%d.npy" % rank), out)
        lo, hi, band = D.sharded_search(plan, Scarp._sb_spec, SCALE, AGES, angles, "age_major", device=cpu,
                                        merge="bands")
        assert (lo, hi) == D.shard_bounds(ny, world, rank) and band.shape == (4, hi - lo, nx)
        np.save(os.path.join(out_dir, "bands%d.npy" % rank), band)
        # every rank uploads half of the rows, all-gather for the rest (twice: the buffer is reused)
        for _ in range(2):
            D.set_dem_sharded(plan, z, cpu)
            lo, hi, band = D.sharded_search(plan, Scarp._sb_spec, SCALE, AGES, angles, "age_major", device=cpu,
                                            merge="bands")
        np.save(os.path.join(out_dir, "bands_shared_upload%d.npy" % rank), band)
    # a row count the ranks do not divide: the last chunk of the all-gather is padded
    with Plan(ny - 1, nx, 1.0, 1.0) as plan:
        plan.set_dem(z[:-1])
        a = plan.directional_laplacian(0.3)
        D.set_dem_sharded(plan, z[:-1], cpu)
        assert np.array_equal(plan.directional_laplacian(0.3), a, equal_nan=True)
    plan, (lo, hi) = D.spatial_plan(ny, nx, 1.0, 1.0, Scarp._sb_spec, SCALE, AGES, angles)
    with plan:
        plan.set_dem(z)
        D.share_dem_stats(plan, device=cpu)
        lo2, hi2, band = D.spatial_search(plan, Scarp._sb_spec, SCALE, AGES, angles)
        assert (lo2, hi2) == (lo, hi)
        np.save(os.path.join(out_dir, "spatial%d.npy" % rank), band)
    dist.destroy_process_group()


def test_two_rank_searches_equal_single_rank(tmp_path, emu_lib):
    import torch.multiprocessing as mp
    from tests.parity import stack_report
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    from scarplet_b200.engine import Plan
    from scarplet_b200.templates import Scarp
    z, angles = _inputs()
    with Plan(SHAPE[0], SHAPE[1], 1.0, 1.0) as plan:
        plan.set_dem(z)
        a, t, age_of, angle_of = plan.build_sweep(Scarp._sb_spec, SCALE, AGES, angles)
        plan.reset()
        plan.sweep(a, t)
        single = plan.finalize(age_of, angle_of)
    assert (single[3] > 0).sum() > 1000
    r0 = np.load(tmp_path / "replicated0.npy")
    r1 = np.load(tmp_path / "replicated1.npy")
    assert np.array_equal(r0, r1) and np.array_equal(r0, single)
    bands = np.concatenate([np.load(tmp_path / ("bands%d.npy" % r)) for r in range(2)], axis=1)
    assert np.array_equal(bands, single)
    bands = np.concatenate([np.load(tmp_path / ("bands_shared_upload%d.npy" % r)) for r in range(2)], axis=1)
    assert np.array_equal(bands, single)
    spatial = np.concatenate([np.load(tmp_path / ("spatial%d.npy" % r)) for r in range(2)], axis=1)
    rep = stack_report(spatial, single)
    assert rep["mask_equal"] and rep["index_agreement"] >= 0.999 and rep["snr_rel_max_strong"] < 2e-5, rep
