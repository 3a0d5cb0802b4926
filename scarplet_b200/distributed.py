"""Multi-GPU orientation search: one process per GPU (``torch.distributed``, NCCL over
NVLink), the orientation list sharded contiguously across ranks — the reference's only
parallel axis (``Pool.imap`` over angles, core.py:180-183).  Every rank holds the whole
DEM and its own best state; the single exchange step is the per-pixel best-SNR merge:
one all-reduce(MAX) on packed 64-bit keys and one all-reduce(SUM) on the winners'
amplitudes (12 bytes per pixel)."""
import numpy as np


def shard_bounds(n_items, world_size, rank):
    """Contiguous shard [lo, hi) of ``n_items`` for ``rank`` (sizes differ by <= 1)."""
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def merge_best_state(plan, device, group=None):
    """All ranks end with the same merged best state in ``plan``."""
    import torch
    import torch.distributed as dist
    n = plan.ny * plan.nx
    keys = torch.empty(n, dtype=torch.int64, device=device)
    amp = torch.empty(n, dtype=torch.float32, device=device)
    plan.best_pack(keys.data_ptr())
    dist.all_reduce(keys, op=dist.ReduceOp.MAX, group=group)
    plan.best_select(keys.data_ptr(), amp.data_ptr())
    dist.all_reduce(amp, op=dist.ReduceOp.SUM, group=group)
    plan.best_unpack(keys.data_ptr(), amp.data_ptr())
    return keys, amp


def sharded_search(plan, spec, scale, ages, angles, order="age_major", device=None,
                   group=None, finalize=True):
    """Run this rank's shard of the orientation search on ``plan`` (DEM already set),
    merge across ranks and decode.  Returns the (4, ny, nx) stack on every rank (or
    ``None`` with ``finalize=False``, leaving the merged state in the plan)."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(len(angles), world, rank)
    a_rec, t_rec, age_of, angle_of = plan.build_sweep(spec, scale, ages, angles, order,
                                                     angle_slice=(lo, hi))
    plan.reset()
    plan.sweep(a_rec, t_rec)
    if world > 1:
        merge_best_state(plan, device, group)
    if not finalize:
        return None
    return plan.finalize(age_of, angle_of)
