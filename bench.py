#!/usr/bin/env python
"""Benchmark of the template-matching hot path (contract in the task statement).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU algorithm

Workload (BASELINE.json config 3, the 1-GPU case of the north-star search): Scarp,
scale 100, 30 log-spaced ages (10^0..10^3.5) x 181 orientations (+-90 deg at 1 deg) on
a seeded synthetic 4096 x 4096 DEM.  One "step" = the whole search = 9.11e10
template-pixel evaluations.  With N > 1 ranks the orientation list is sharded across
GPUs (strong scaling, total work fixed) and merged with two NCCL all-reduces.

Metric: template-pixel evaluations per second (Mpx-evals/s).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "template_pixel_evals_per_sec"
UNIT = "Mpx-evals/s"


def workload(args):
    ages = np.logspace(0, 3.5, args.ages) if args.ages > 1 else np.array([10.0])
    return {"n": args.size, "seed": 2, "scale": 100.0, "ages": ages,
            "ang_min": -np.pi / 2, "ang_max": np.pi / 2}


def config_dict(args, wl, n_angles, extra=None):
    cfg = {"workload": "C3: calculate_best_fit_parameters Scarp scale=100, %d log-spaced ages "
                       "(10^0-10^3.5) x %d angles on a synthetic %dx%d DEM (seed %d)"
                       % (len(wl["ages"]), n_angles, wl["n"], wl["n"], wl["seed"]),
           "size": wl["n"], "n_ages": int(len(wl["ages"])), "n_angles": int(n_angles),
           "template": "Scarp", "scale": wl["scale"],
           "px_evals_per_step": int(wl["n"]) ** 2 * int(len(wl["ages"])) * int(n_angles),
           "l2": "working set per step (>= 16 GB of spectra and intermediates) exceeds the 126 MB L2; no flush needed"}
    if extra:
        cfg.update(extra)
    return cfg


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.samples:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------
# reference CPU algorithm (oracle port of scarplet/core.py) on the host cores
# ---------------------------------------------------------------------------
def _cpu_worker_count(n):
    cores = os.cpu_count() or 1
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 32 << 30
    per_worker = 200 * n * n          # ~155 B/px live + slack
    return max(1, min(cores, int(avail * 0.6 // per_worker)))


def cpu_reference_sample(z, scale, ages, angles, workers):
    """The reference's own flow for the sample: Pool fan-out over orientations at each
    age, ordered imap, parent-side compare (core.py:139-195, 266-294), NumPy/pocketfft in
    place of numexpr/pyfftw.  Returns (px_evals, seconds)."""
    from oracle import scarplet_oracle as O
    import multiprocessing as mp
    from functools import partial
    ny, nx = z.shape
    t0 = time.perf_counter()
    stacks = []
    with mp.Pool(processes=workers) as pool:
        for age in ages:
            work = partial(O.match_template, z, 1.0, 1.0, O.SCARP, scale, age)
            best = O.compare(pool.imap(work, angles, chunksize=1), ny, nx)
            stacks.append(np.stack(best))
    O.compare(stacks, ny, nx)
    dt = time.perf_counter() - t0
    return ny * nx * len(ages) * len(angles), dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from scarplet_b200.synth import synthetic_dem
    from scarplet_b200 import params as P
    wl = workload(args)
    angles_all = P.search_angles(wl["ang_min"], wl["ang_max"])
    n = wl["n"]
    z = synthetic_dem(n, wl["seed"])
    workers = _cpu_worker_count(n)
    # bounded sample of the same workload: one orientation per worker at one age
    sel = np.linspace(0, len(angles_all) - 1, min(workers, len(angles_all))).astype(int)
    angles = angles_all[sel]
    ages = wl["ages"][len(wl["ages"]) // 2: len(wl["ages"]) // 2 + 1]
    times = []
    evals = 0
    for it in range(args.warmup + args.steps):
        evals, dt = cpu_reference_sample(z, wl["scale"], ages, angles, workers)
        if it >= args.warmup:
            times.append(dt)
    sec = float(np.mean(times))
    value = evals / sec / 1e6
    sample = "%d orientations x %d age of the %dx%d search per step (extrapolates linearly)" % (
        len(angles), len(ages), n, n)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, wl, len(angles_all), {"sample": sample}),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port",
                             "sample": sample,
                             "note": "oracle port of scarplet/core.py (NumPy/pocketfft standing in for "
                                     "numexpr/pyfftw, which are not installable here), mp.Pool like the reference"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------
# this repo's CUDA path
# ---------------------------------------------------------------------------
ALGO_BYTES = {
    # algorithmic bytes per template-pixel evaluation of each per-template kernel
    # (complex64 half spectra; DESIGN.md "byte model")
    "k_conv_cols": 16.0,   # read F[curv], F[curv^2] (8) + write both inverse-column planes (8)
    "k_fit_rows": 8.0,     # read both planes (8); best state amortised over the batch
}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__
    from scarplet_b200.synth import synthetic_dem
    from scarplet_b200 import params as P
    from scarplet_b200.engine import Plan
    from scarplet_b200.templates import Scarp
    from scarplet_b200 import distributed as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.barrier()

    wl = workload(args)
    n = wl["n"]
    angles = P.search_angles(wl["ang_min"], wl["ang_max"])
    ages = wl["ages"]
    z = synthetic_dem(n, wl["seed"])
    z_pinned = torch.from_numpy(z).pin_memory()
    evals_per_step = n * n * len(ages) * len(angles)

    stream = torch.cuda.Stream(device=device)
    sampler = ClockSampler(local)
    with torch.cuda.stream(stream):
        plan = Plan(n, n, 1.0, 1.0, device=local, stream=stream.cuda_stream)
        if args.fast is not None:
            plan.set_option("fast", args.fast)
        plan.set_dem(z_pinned.numpy())                       # inputs resident in HBM
        lo, hi = D.shard_bounds(len(angles), world, rank)
        a_rec, t_rec, age_of, angle_of = plan.build_sweep(Scarp._sb_spec, wl["scale"], ages, angles,
                                                         "age_major", angle_slice=(lo, hi))

        def step():
            plan.reset()
            plan.sweep(a_rec, t_rec)
            if world > 1:
                D.merge_best_state(plan, device)

        def fence():
            stream.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        for _ in range(args.warmup):
            step()
        fence()
        plan.set_option("profile", 1)
        plan.profile(reset=True)
        launches0 = plan.launches
        if rank == 0:
            sampler.start()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        fence()
        clocks = sampler.stop() if rank == 0 else None
        ms_total = e0.elapsed_time(e1)
        prof = plan.profile(reset=True)
        plan.set_option("profile", 0)
        launches = plan.launches - launches0
        geo = plan.last_geometry()

        t_ms = torch.tensor([ms_total], dtype=torch.float64, device=device)
        n_launch = torch.tensor([launches], dtype=torch.int64, device=device)
        if world > 1:
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(n_launch, op=dist.ReduceOp.SUM)
        ms_step = float(t_ms.item()) / args.steps
        value = evals_per_step / (ms_step * 1e-3) / 1e6

        # ---- end to end through the public API with host buffers ---------------
        def e2e_step():
            plan.set_dem(z_pinned.numpy())                   # H2D inside the timed region
            out = D.sharded_search(plan, Scarp._sb_spec, wl["scale"], ages, angles, "age_major",
                                   device=device, finalize=(rank == 0))
            return out

        for _ in range(3):          # untimed: the plan's page-locked result pool fills (engine._result_array)
            out = e2e_step()
        fence()
        t0 = time.perf_counter()
        out = None
        for _ in range(args.e2e_steps):
            out = e2e_step()
        fence()
        e2e_s = (time.perf_counter() - t0) / args.e2e_steps
        t_e2e = torch.tensor([e2e_s], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        e2e_value = evals_per_step / float(t_e2e.item()) / 1e6
        h2d = int(z.nbytes) * world + sum(int(ctypes_sizeof(x)) for x in (a_rec[0], t_rec[0]))
        d2h = int(out.nbytes) if out is not None else 0
        plan.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel -----------------------------------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
    per_template = {k: v for k, v in prof.items() if k in ALGO_BYTES and v[1] > 0}
    dom = max(per_template, key=lambda k: per_template[k][0]) if per_template else "k_conv_cols"
    dom_ms, dom_launches = prof.get(dom, (0.0, 0))
    my_evals = n * n * len(ages) * (hi - lo) * args.steps       # rank 0's share
    achieved = ALGO_BYTES[dom] * my_evals / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    # DRAM bytes of the same kernel from the committed `ncu --set full` capture
    # (profiles/traffic.json, bytes per px-eval) scaled to this run's average launch
    traffic = None
    traffic_src = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f).get(dom)
        if tj and dom_launches:
            traffic = tj["dram_bytes_per_px_eval"] * my_evals / dom_launches
            traffic_src = tj.get("source")
    except Exception:
        pass
    total_kernel_ms = sum(v[0] for v in prof.values())
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak if peak else None, "traffic": traffic, "traffic_source": traffic_src,
                "achieved_bytes_per_launch": (ALGO_BYTES[dom] * my_evals / dom_launches) if dom_launches else None,
                "peak_source": peak_src,
                "algorithmic_bytes_per_px_eval": ALGO_BYTES[dom],
                "avg_launch_ms": dom_ms / dom_launches if dom_launches else None,
                "launches": dom_launches,
                "share_of_step": dom_ms / total_kernel_ms if total_kernel_ms else None,
                "kernel_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()},
                "whole_step_hbm_frac": (24.0 * my_evals / (total_kernel_ms * 1e-3) / 1e9 / peak)
                if total_kernel_ms else None}

    # ---- CPU baseline: oracle port on this host's cores, bounded sample --------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        workers = _cpu_worker_count(n)
        sel = np.linspace(0, len(angles) - 1, min(workers, len(angles))).astype(int)
        c_ages = ages[len(ages) // 2: len(ages) // 2 + 1]
        evals, dt = cpu_reference_sample(z, wl["scale"], c_ages, angles[sel], workers)
        cpu = {"value": evals / dt / 1e6, "unit": UNIT, "cores": workers, "kind": "port",
               "sample": "%d orientations x 1 age of the %dx%d search, %.1f s" % (len(sel), n, n, dt)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, wl, len(angles), {"parallelism": "orientations sharded over %d GPU(s)" % world,
                                                         "fft_domain": [geo["Py"], geo["Px"]],
                                                         "tiles": [geo["tiles_y"], geo["tiles_x"]],
                                                         "batches": [geo["angle_batch"], geo["template_batch"]]}),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": float(t_e2e.item()) * 1e3, "steps": args.e2e_steps,
                    "api": "Plan.set_dem(host) + distributed.sharded_search(...) -> (4,ny,nx) float64 on host"},
            "gpu_launches": int(n_launch.item()),
            "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def ctypes_sizeof(arr):
    import ctypes
    return ctypes.sizeof(arr)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=4096)
    ap.add_argument("--ages", type=int, default=30)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fast", type=int, default=None, help="developer switch: 0 = simple kernels")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
