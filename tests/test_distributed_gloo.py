"""world_size-2 check of the multi-GPU path on CPU: gloo backend, the emulator build of
the kernels standing in for the two devices.  Orientations are sharded across ranks,
best states merged with all-reduce(MAX) on packed keys + all-reduce(SUM) on amplitudes;
the result must equal the single-rank search bit for bit."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), SB_EMU_WORKERS="2")
    import torch
    import torch.distributed as dist
    from tests.emu.build_emu import build
    from scarplet_b200 import _lib
    _lib._use_library(_lib.open_library(build()))
    from scarplet_b200 import params as P, distributed as D
    from scarplet_b200.engine import Plan
    from scarplet_b200.synth import synthetic_dem
    from scarplet_b200.templates import Scarp
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z = synthetic_dem(96, seed=7, nx=128, relief=3.0)
    angles = P.search_angles(-np.pi / 2, np.pi / 2)[::9]
    with Plan(96, 128, 1.0, 1.0) as plan:
        plan.set_dem(z)
        out = D.sharded_search(plan, Scarp._sb_spec, 8, [2.0, 9.0], angles, "age_major",
                               device=torch.device("cpu"))
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), out)
    dist.destroy_process_group()


def test_two_rank_merge_equals_single_rank(tmp_path, emu_lib):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / "rank0.npy")
    r1 = np.load(tmp_path / "rank1.npy")
    assert np.array_equal(r0, r1)
    from scarplet_b200 import params as P
    from scarplet_b200.engine import Plan
    from scarplet_b200.synth import synthetic_dem
    from scarplet_b200.templates import Scarp
    z = synthetic_dem(96, seed=7, nx=128, relief=3.0)
    angles = P.search_angles(-np.pi / 2, np.pi / 2)[::9]
    with Plan(96, 128, 1.0, 1.0) as plan:
        plan.set_dem(z)
        a, t, age_of, angle_of = plan.build_sweep(Scarp._sb_spec, 8, [2.0, 9.0], angles)
        plan.reset()
        plan.sweep(a, t)
        single = plan.finalize(age_of, angle_of)
    assert np.array_equal(r0, single)
    assert (single[3] > 0).sum() > 1000
