#!/usr/bin/env python
"""Benchmark of the template-matching hot path (contract in the task statement).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU algorithm

Default workload = BASELINE.json config 3 (the 1-GPU case of the north-star search): Scarp,
scale 100, 30 log-spaced ages (10^0..10^3.5) x 181 orientations (+-90 deg at 1 deg) on a
seeded synthetic 4096 x 4096 DEM.  One "step" = the whole search = 9.11e10 template-pixel
evaluations.  With N > 1 ranks the orientation list is sharded across GPUs (strong scaling,
total work fixed) and the best states are merged with one all-to-all of row bands.

Other workloads (kept measurements, not the driver's line): ``--workload c1|c2|c4`` (the other
single-GPU BASELINE configs), ``--workload ns`` (north-star: 16384^2, 30 ages x 181 angles,
orientations sharded), ``--workload c5`` (32768^2, the same search, the raster sharded into row
bands with halos: no data-path collective).  ``ns`` and ``c5`` generate their DEM on the device.

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

WORKLOADS = {
    "c1": dict(n=1024, seed=0, template="Scarp", scales=[100.0], ages=[10.0], relief=30.0, shard="orientations",
               label="C1: sl.match Scarp scale=100 age=10"),
    "c2": dict(n=3601, seed=1, template="Channel", scales=[10.0], ages=[0.1], relief=300.0, shard="orientations",
               label="C2: sl.match Channel scale=10 age=0.1 (pixel units, SRTM-like relief)"),
    "c3": dict(n=4096, seed=2, template="Scarp", scales=[100.0], ages=30, relief=30.0, shard="orientations",
               label="C3: calculate_best_fit_parameters Scarp scale=100, 30 log-spaced ages (10^0-10^3.5)"),
    "c4": dict(n=8192, seed=3, template="Scarp", scales=[25.0, 50.0, 100.0, 200.0], ages=[10.0], relief=30.0,
               shard="orientations", label="C4: multi-scale Scarp sweep (scale 25/50/100/200) age=10, one sweep"),
    "ns": dict(n=16384, seed=4, template="Scarp", scales=[100.0], ages=30, relief=30.0, shard="orientations",
               device_dem=True, label="north-star: Scarp scale=100, 30 log-spaced ages"),
    "c5": dict(n=32768, seed=4, template="Scarp", scales=[100.0], ages=30, relief=30.0, shard="rows",
               device_dem=True, label="C5: regional DEM in row bands with halos, Scarp scale=100, 30 log-spaced ages"),
}


def workload(args):
    wl = dict(WORKLOADS[args.workload])
    if args.size:
        wl["n"] = args.size
    if args.ages:
        wl["ages"] = args.ages
    if args.shard:
        wl["shard"] = args.shard
    a = wl["ages"]
    wl["ages"] = (np.logspace(0, 3.5, a) if a > 1 else np.array([10.0])) if isinstance(a, int) else np.asarray(a, float)
    wl["ang_min"], wl["ang_max"] = -np.pi / 2, np.pi / 2
    wl["name"] = args.workload
    return wl


def px_evals(wl, n_angles):
    return int(wl["n"]) ** 2 * len(wl["ages"]) * int(n_angles) * len(wl["scales"])


def config_dict(wl, n_angles, extra=None):
    cfg = {"workload": "%s x %d angles on a synthetic %dx%d DEM (seed %d)"
                       % (wl["label"], n_angles, wl["n"], wl["n"], wl["seed"]),
           "size": wl["n"], "n_ages": int(len(wl["ages"])), "n_angles": int(n_angles), "scales": list(wl["scales"]),
           "template": wl["template"], "px_evals_per_step": px_evals(wl, n_angles),
           "l2": "working set per step (>= 16 GB of spectra and intermediates at C3) exceeds the 126 MB L2; "
                 "no flush needed"}
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


def cpu_reference_sample(z, kind, scale, ages, angles, workers):
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
            work = partial(O.match_template, z, 1.0, 1.0, kind, scale, age)
            best = O.compare(pool.imap(work, angles, chunksize=1), ny, nx)
            stacks.append(np.stack(best))
    O.compare(stacks, ny, nx)
    dt = time.perf_counter() - t0
    return ny * nx * len(ages) * len(angles), dt


def _oracle_kind(wl):
    from oracle import scarplet_oracle as O
    return O.SCARP if wl["template"] == "Scarp" else O.RICKER


def _cpu_sample(wl, angles_all, z):
    """Bounded sample of the workload: one orientation per worker at the middle age / first scale."""
    n = wl["n"]
    workers = _cpu_worker_count(n)
    sel = np.linspace(0, len(angles_all) - 1, min(workers, len(angles_all))).astype(int)
    ages = wl["ages"][len(wl["ages"]) // 2: len(wl["ages"]) // 2 + 1]
    sample = "%d orientations x %d age of the %dx%d search per step (extrapolates linearly)" % (
        len(sel), len(ages), n, n)
    return workers, angles_all[sel], ages, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from scarplet_b200.synth import synthetic_dem
    from scarplet_b200 import params as P
    wl = workload(args)
    if wl.get("device_dem"):
        wl["n"] = 4096          # the CPU arm cannot hold the large rasters: same search on a 4096^2 sample
    angles_all = P.search_angles(wl["ang_min"], wl["ang_max"])
    n = wl["n"]
    z = synthetic_dem(n, wl["seed"], relief=wl["relief"])
    workers, angles, ages, sample = _cpu_sample(wl, angles_all, z)
    times = []
    evals = 0
    for it in range(args.warmup + args.steps):
        evals, dt = cpu_reference_sample(z, _oracle_kind(wl), wl["scales"][0], ages, angles, workers)
        if it >= args.warmup:
            times.append(dt)
    sec = float(np.mean(times))
    value = evals / sec / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(wl, len(angles_all), {"sample": sample}),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port",
                             "sample": sample,
                             "note": "oracle port of scarplet/core.py (NumPy/pocketfft standing in for "
                                     "numexpr/pyfftw, which are not installable here), mp.Pool like the reference"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------
# device-side DEM for the rasters the host generator cannot hold
# ---------------------------------------------------------------------------
def device_dem_rows(n, seed, row0, nrows, device, relief=30.0, tile=4096):
    """Rows ``(row0 + arange(nrows)) mod n`` of a seeded n x n DEM, float64 on ``device``.

    SURVEY 8d's recipe at scale: a ``tile``^2 spectral-synthesis fractal (amplitude ~ k^-2.5,
    zero DC, sigma = relief) repeated over the raster and modulated by a smooth amplitude field
    that does not share its period (no two repeats are alike, nothing is mirrored), + 650 m +
    0.02 x regional tilt + one diffused scarp 1.5 erf(x_rot / (2 sqrt(10))) at 0.3 rad + white
    noise sigma = 0.03 m drawn per 256-row block from its own seeded generator (any band of rows
    is reproducible on any rank), rounded to float32 like a GDAL Float32 raster (dem.py:317)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    t = min(tile, n)
    ky = torch.fft.fftfreq(t, device=device)[:, None]
    kx = torch.fft.rfftfreq(t, device=device)[None, :]
    k = torch.sqrt(kx ** 2 + ky ** 2)
    k[0, 0] = 1
    spec = torch.complex(torch.randn(k.shape, generator=g, device=device), torch.randn(k.shape, generator=g, device=device))
    spec = spec * k ** -2.5
    spec[0, 0] = 0
    base = torch.fft.irfft2(spec, s=(t, t))
    base = (relief * base / base.std()).to(torch.float32)
    del spec, k
    rows = (row0 + torch.arange(nrows, device=device)) % n
    cols = torch.arange(n, device=device)
    z = base[rows % t][:, cols % t].to(torch.float64)
    yy = rows.to(torch.float64)[:, None]
    xx = cols.to(torch.float64)[None, :]
    amp = 1.0 + 0.25 * torch.sin(2 * np.pi * yy / (0.37 * n) + 1.3) * torch.cos(2 * np.pi * xx / (0.41 * n) + 0.7)
    z *= amp
    del amp
    c = (n - 1) / 2.0
    xr = (xx - c) * np.cos(0.3) + (yy - c) * np.sin(0.3)
    z += 650.0 + 0.02 * xx + 1.5 * torch.erf(xr / (2 * np.sqrt(10.0)))
    del xr
    blk = 256
    for b in sorted(set((rows // blk).tolist())):
        gb = torch.Generator(device=device)
        gb.manual_seed(int(seed) * 1000003 + int(b))
        noise = torch.randn((blk, n), generator=gb, device=device, dtype=torch.float32)
        sel = (rows // blk) == b
        z[sel] += 0.03 * noise[(rows[sel] % blk)].to(torch.float64)
    return z.to(torch.float32).to(torch.float64).contiguous()


# ---------------------------------------------------------------------------
# roofline bookkeeping
# ---------------------------------------------------------------------------
def algorithmic_bytes(kernel, tpa, batch, n_angles=181, angle_batch=64):
    """HBM bytes one template-pixel evaluation needs from each kernel of the complex64 schedule
    (``tpa`` = templates per orientation, ``batch`` = templates per fold
    launch).  Half spectra: a plane of Py x (Px/2+1) complex values is 4 B per pixel and field."""
    if kernel == "k_curv_rows":
        # once per FFT tile: five packed plane pairs through a row pass (read 16, write 8) and a
        # column pass (read 8, write 8) -- k_diff_rows_f + k_curv_cols
        return 5 * (16.0 + 8.0 + 8.0 + 8.0) / (tpa * n_angles)
    if kernel == "k_curv_cols":
        # k_combine_spectra: nine planes read once per batch of orientations (36 B/px), both spectra
        # of every orientation written (8 B/px)
        return (36.0 / max(angle_batch, 1) + 8.0) / tpa
    if kernel == "k_tmpl_rows":
        return 0.6                       # the template's row spectra on its support rows only
    if kernel == "k_conv_cols":
        # write both inverse-column planes (8); the curvature-spectrum columns are staged in shared
        # memory once per run of same-angle templates by the persistent kernels (8 / tpa), else per template
        return 8.0 + (8.0 / tpa if tpa >= 4 else 8.0) + 0.6
    if kernel == "k_fit_rows":
        return 8.0 + 12.0 / max(batch, 1)    # read both planes (8) + best-state read-modify-write per launch
    return 0.0


def load_profile_counters():
    """On-chip counters and DRAM bytes of the dominant kernels from the committed
    ``ncu --set full`` capture of the benchmarked batch shape (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


# ---------------------------------------------------------------------------
# this repo's CUDA path
# ---------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__
    from scarplet_b200.synth import synthetic_dem
    from scarplet_b200 import params as P
    from scarplet_b200.engine import Plan
    from scarplet_b200 import templates as T
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
    spec = getattr(T, wl["template"])._sb_spec
    angles = P.search_angles(wl["ang_min"], wl["ang_max"])
    ages, scales = wl["ages"], wl["scales"]
    scale_arg = scales if len(scales) > 1 else scales[0]
    evals_per_step = px_evals(wl, len(angles))
    rows_mode = wl["shard"] == "rows"
    stream = torch.cuda.Stream(device=device)
    sampler = ClockSampler(local)

    with torch.cuda.stream(stream):
        # ---- plan and DEM ----------------------------------------------------------------
        if rows_mode:
            plan, (row_lo, row_hi) = D.spatial_plan(n, n, 1.0, 1.0, spec, scale_arg, ages, angles,
                                                    device=local, stream=stream.cuda_stream, states=len(scales))
        else:
            plan = Plan(n, n, 1.0, 1.0, device=local, stream=stream.cuda_stream, states=len(scales))
            row_lo, row_hi = 0, n
        if args.fast is not None:
            plan.set_option("fast", args.fast)
        r0, nrows = plan.dem_rows()
        z = z_pinned = z_dev = None
        if wl.get("device_dem"):
            z_dev = device_dem_rows(n, wl["seed"], r0, nrows, device, relief=wl["relief"])
            stream.synchronize()
            plan.set_dem_device(z_dev.data_ptr())
        else:
            z = synthetic_dem(n, wl["seed"], relief=wl["relief"])
            z_pinned = torch.from_numpy(np.take(z, (r0 + np.arange(nrows)) % n, axis=0) if nrows != n else z).pin_memory()
            plan.set_dem(z_pinned.numpy())                       # inputs resident in HBM
        if rows_mode:
            D.share_dem_stats(plan, device=device)
            a_rec, t_rec, age_of, angle_of = plan.build_sweep(spec, scale_arg, ages, angles, "age_major")
        else:
            a_rec, t_rec, age_of, angle_of = plan.build_sweep(spec, scale_arg, ages, angles, "age_major",
                                                             template_share=(rank, world))
        my_templates = t_rec[1]
        merge_ms = []

        def step(timed=False):
            plan.reset()
            plan.sweep(a_rec, t_rec)
            if world > 1 and not rows_mode:
                m0 = torch.cuda.Event(enable_timing=True)
                m1 = torch.cuda.Event(enable_timing=True)
                m0.record(stream)
                for s in range(len(scales)):
                    D.merge_best_bands(plan, device, state=s)
                m1.record(stream)
                if timed:
                    merge_ms.append((m0, m1))

        def fence():
            stream.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        for _ in range(args.warmup):
            step()
        fence()
        # ---- timed region: K steps, CUDA events on the launching stream, no per-kernel events ----
        launches0 = plan.launches
        if rank == 0:
            sampler.start()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        t_wall0 = time.perf_counter()
        e0.record(stream)
        for _ in range(args.steps):
            step(timed=True)
        e1.record(stream)
        fence()
        wall_ms = (time.perf_counter() - t_wall0) * 1e3 / args.steps
        clocks = sampler.stop() if rank == 0 else None
        ms_total = e0.elapsed_time(e1)
        launches = plan.launches - launches0
        geo = plan.last_geometry()
        fft_area = plan.fft_area
        # ---- the same steps again with an event pair around every launch: per-kernel times --------
        plan.set_option("profile", 1)
        plan.profile(reset=True)
        prof_steps = args.steps if args.profile_steps is None else args.profile_steps
        for _ in range(prof_steps):
            step()
        fence()
        prof = plan.profile(reset=True)
        plan.set_option("profile", 0)
        if prof_steps:
            prof = {k: (v[0] * args.steps / prof_steps, v[1] * args.steps / prof_steps) for k, v in prof.items()}
        merge = float(np.mean([a.elapsed_time(b) for a, b in merge_ms])) if merge_ms else 0.0

        stats = torch.tensor([ms_total, merge, float(plan.device_bytes + torch.cuda.max_memory_allocated(device)), -merge],
                             dtype=torch.float64, device=device)
        n_launch = torch.tensor([launches], dtype=torch.int64, device=device)
        if world > 1:
            dist.all_reduce(stats, op=dist.ReduceOp.MAX)
            dist.all_reduce(n_launch, op=dist.ReduceOp.SUM)
        ms_step = float(stats[0].item()) / args.steps
        value = evals_per_step / (ms_step * 1e-3) / 1e6

        # ---- end to end through the public multi-GPU API with host buffers -------------------
        # every rank: H2D of its DEM (whole raster, or its band), search, merge, decode and D2H of
        # the row band it owns afterwards -- the distributed form of the (4, ny, nx) result
        e2e = None
        if not wl.get("device_dem"):
            shared_upload = world > 1 and not rows_mode and os.environ.get("SB_BENCH_SHARED_UPLOAD", "1") != "0"

            def e2e_step():
                if shared_upload:      # H2D of 1 / world of the rows per rank + all-gather over NVLink
                    D.set_dem_sharded(plan, z_pinned, device)
                else:
                    plan.set_dem(z_pinned.numpy())               # H2D inside the timed region
                if rows_mode:
                    D.share_dem_stats(plan, device=device)
                    return D.spatial_search(plan, spec, scale_arg, ages, angles)
                return D.sharded_search(plan, spec, scale_arg, ages, angles, "age_major", device=device, merge="bands")

            for _ in range(5 if world > 1 else 3):   # untimed: the plan's page-locked result pool fills
                out = e2e_step()                     # (engine._result_array), NCCL connects its channels
            phases = None
            if args.e2e_phases and not rows_mode:
                # diagnostic (untimed, not the e2e number): where one e2e step spends its time on this
                # rank -- host clock after a stream synchronise at the end of each phase
                acc = np.zeros(5)
                for _ in range(4):
                    if world > 1:
                        dist.barrier()
                    fence()
                    t = [time.perf_counter()]
                    if shared_upload:
                        D.set_dem_sharded(plan, z_pinned, device)
                    else:
                        plan.set_dem(z_pinned.numpy())
                    stream.synchronize(); t.append(time.perf_counter())
                    a_s, t_s, age_s, angle_s = plan.build_sweep(spec, scale_arg, ages, angles, "age_major",
                                                               template_share=(rank, world))
                    plan.reset()
                    plan.sweep(a_s, t_s)
                    stream.synchronize(); t.append(time.perf_counter())
                    band = D.merge_best_bands(plan, device) if world > 1 else (0, n)
                    stream.synchronize(); t.append(time.perf_counter())
                    out = plan.finalize(age_s, angle_s, rows=band)
                    stream.synchronize(); t.append(time.perf_counter())
                    if world > 1:
                        dist.barrier()
                    t.append(time.perf_counter())
                    acc += np.diff(t) * 1e3
                acc /= 4
                ph = torch.tensor(acc, dtype=torch.float64, device=device)
                ph_max = ph.clone()
                if world > 1:
                    dist.all_reduce(ph_max, op=dist.ReduceOp.MAX)
                    dist.all_reduce(ph, op=dist.ReduceOp.MIN)
                names = ("upload", "sweep", "merge", "decode_download", "wait_for_ranks")
                phases = {k: [round(float(a), 2), round(float(b), 2)] for k, a, b in zip(names, ph.tolist(), ph_max.tolist())}
                by_rank = [torch.zeros(1, dtype=torch.float64, device=device) for _ in range(world)]
                mine = torch.tensor([acc[1]], dtype=torch.float64, device=device)
                if world > 1:
                    dist.all_gather(by_rank, mine)
                else:
                    by_rank = [mine]
                phases["sweep_by_rank"] = [round(float(v.item()), 2) for v in by_rank]
                phases["templates_by_rank"] = "SB_SHARE_BALANCE=%s" % os.environ.get("SB_SHARE_BALANCE", "cost")
            # enough steps for ~1.3 s of timed work (short multi-GPU steps are sensitive to single
            # host hiccups), at least 2, at most 10
            e2e_steps = args.e2e_steps or int(min(10, max(2, -(-1300.0 // ms_step))))
            fence()
            t0 = time.perf_counter()
            out = None
            marks = []
            for _ in range(e2e_steps):
                out = e2e_step()
                marks.append(time.perf_counter())
            fence()
            e2e_s = (time.perf_counter() - t0) / e2e_steps
            step_ms = [round((b - a) * 1e3, 2) for a, b in zip([t0] + marks[:-1], marks)]
            t_e2e = torch.tensor([e2e_s], dtype=torch.float64, device=device)
            h2d_dem = z_pinned.numel() * 8
            if shared_upload:
                h2d_dem = max(0, min((rank + 1) * -(-n // world), n) - rank * -(-n // world)) * z_pinned.shape[1] * 8
            io = torch.tensor([int(h2d_dem) + ctypes_sizeof(a_rec[0]) + ctypes_sizeof(t_rec[0]),
                               int(out[2].nbytes)], dtype=torch.int64, device=device)
            if world > 1:
                dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
                dist.all_reduce(io, op=dist.ReduceOp.SUM)
            e2e = {"value": evals_per_step / float(t_e2e.item()) / 1e6, "unit": UNIT,
                   "h2d_bytes_per_step": int(io[0].item()), "d2h_bytes_per_step": int(io[1].item()),
                   "ms_per_step": float(t_e2e.item()) * 1e3, "steps": e2e_steps, "rank0_step_ms": step_ms,
                   "phases_ms_min_max_over_ranks": phases,
                   "api": "per rank: %s + distributed.%s -> (row_lo, row_hi, (4, rows, nx) float64 on host); "
                          "the ranks' row bands together are the (4, ny, nx) result"
                          % ("distributed.set_dem_sharded(host: 1 / world of the rows over PCIe, all-gather)" if shared_upload
                             else "Plan.set_dem(host)",
                             "spatial_search(...)" if rows_mode else "sharded_search(..., merge='bands')")}
            del out
        plan.close()

        # ---- the drop-in entry itself: sl.match(DEMGrid, ...) cold and warm (1 GPU) ---------
        dropin = None
        if world == 1 and not wl.get("device_dem") and not args.no_dropin:
            import scarplet_b200 as sl
            grid = sl.DEMGrid(z, 1.0)
            cls = getattr(sl.WindowedTemplate, wl["template"])

            def call():
                if len(scales) > 1:
                    return sl.match_scales(grid, cls, scales, age=float(ages[0]))
                if len(ages) == 1:
                    return sl.match(grid, cls, scale=scales[0], age=float(ages[0]))
                return sl.match(grid, cls, scale=scales[0], ages=ages)
            times = []
            for _ in range(4):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                res = call()
                times.append(time.perf_counter() - t0)
                del res
            sl.release()
            dropin = {"api": "sl.match(DEMGrid, %s, ...) -> float64 stack(s) on host (plans cached per raster shape)" % wl["template"],
                      "cold_ms": times[0] * 1e3, "warm_ms": float(np.mean(times[2:])) * 1e3,
                      "warm_value": evals_per_step / float(np.mean(times[2:])) / 1e6, "unit": UNIT}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline: every kernel, three ways -------------------------------------------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
    my_evals = int(n) * int(row_hi - row_lo) * my_templates * args.steps   # rank 0's share
    tpa = len(ages) * len(scales)
    batch = max(1, geo["template_batch"])
    counters = load_profile_counters()
    per_kernel = {}
    total_ms = sum(v[0] for v in prof.values())
    total_bytes = 0.0
    for k, (ms, cnt) in prof.items():
        if cnt == 0:
            continue
        ab_k = algorithmic_bytes(k, tpa, batch, len(angles), geo["angle_batch"])
        b = ab_k * my_evals
        total_bytes += b
        entry = {"ms_per_step": ms / args.steps, "launches_per_step": cnt / args.steps,
                 "algorithmic_bytes_per_px_eval": ab_k,
                 "algorithmic_GBps": b / (ms * 1e-3) / 1e9 if ms > 0 else None,
                 "algorithmic_frac": b / (ms * 1e-3) / 1e9 / peak if ms > 0 else None,
                 "share_of_step": ms / total_ms if total_ms else None}
        c = counters.get(k)
        if c and ms > 0:
            dram = c["dram_bytes_per_px_eval"] * my_evals
            entry.update({"dram_bytes_per_px_eval": c["dram_bytes_per_px_eval"],
                          "dram_GBps": dram / (ms * 1e-3) / 1e9, "dram_frac": dram / (ms * 1e-3) / 1e9 / peak,
                          "lsu_wavefront_pct": c.get("lsu_wavefront_pct"), "fp32_pipe_pct": c.get("fma_pipe_pct"),
                          "issue_active_pct": c.get("issue_active_pct"), "counters_source": c.get("source")})
        per_kernel[k] = entry
    if not per_kernel:
        per_kernel = {"k_conv_cols": {"ms_per_step": 0.0, "algorithmic_GBps": None, "algorithmic_frac": None, "share_of_step": None}}
        prof = {"k_conv_cols": (0.0, 0)}
    heavy = {k: v for k, v in per_kernel.items() if k in ("k_conv_cols", "k_fit_rows")}
    dom = max(heavy, key=lambda k: heavy[k]["ms_per_step"]) if heavy else max(per_kernel, key=lambda k: per_kernel[k]["ms_per_step"])
    d = per_kernel[dom]
    dom_ms, dom_launches = prof[dom]
    ab = algorithmic_bytes(dom, tpa, batch) * my_evals
    roofline = {"bound": "hbm", "kernel": dom, "achieved": d["algorithmic_GBps"], "peak": peak, "unit": "GB/s",
                "frac": d["algorithmic_frac"],
                "traffic": (d["dram_bytes_per_px_eval"] * my_evals / dom_launches) if "dram_bytes_per_px_eval" in d else None,
                "traffic_source": d.get("counters_source"),
                "dram_frac": d.get("dram_frac"), "fp32_pipe_pct": d.get("fp32_pipe_pct"),
                "lsu_wavefront_pct": d.get("lsu_wavefront_pct"),
                "achieved_bytes_per_launch": ab / dom_launches if dom_launches else None,
                "peak_source": peak_src, "algorithmic_bytes_per_px_eval": algorithmic_bytes(dom, tpa, batch),
                "avg_launch_ms": dom_ms / dom_launches if dom_launches else None, "launches": dom_launches,
                "share_of_step": d["share_of_step"],
                "whole_step": {"algorithmic_bytes_per_px_eval": total_bytes / my_evals if my_evals else None,
                               "algorithmic_frac": total_bytes / (total_ms * 1e-3) / 1e9 / peak if total_ms else None,
                               "kernel_ms_per_step": total_ms / args.steps},
                "kernels": per_kernel,
                "note": "frac = algorithmic bytes / CUDA-event time / measured copy peak; dram_frac = DRAM bytes of the "
                        "committed ncu capture scaled to this run / the same time; the binding resource of both heavy "
                        "kernels is on-chip (FP32 pipe and shared-memory wavefronts), see fp32_pipe_pct / lsu_wavefront_pct"}

    # ---- CPU baseline: oracle port on this host's cores, bounded sample --------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline and z is not None:
        workers, c_angles, c_ages, sample = _cpu_sample(wl, angles, z)
        evals, dt = cpu_reference_sample(z, _oracle_kind(wl), scales[0], c_ages, c_angles, workers)
        cpu = {"value": evals / dt / 1e6, "unit": UNIT, "cores": workers, "kind": "port",
               "sample": "%s, %.1f s" % (sample, dt)}

    area = float(row_hi - row_lo) * n
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(wl, len(angles), {
                "parallelism": ("raster sharded into %d row band(s) with halos" if rows_mode
                                else "orientations sharded over %d GPU(s)") % world,
                "fft_domain_max": [geo["Py"], geo["Px"]], "tiles": [geo["tiles_y"], geo["tiles_x"]],
                "fft_area_over_raster_area": fft_area / area if area else None,
                "batches": [geo["angle_batch"], geo["template_batch"]]}),
            "clocks": clocks, "e2e": e2e, "e2e_dropin": dropin,
            "gpu_launches": int(n_launch.item()),
            # on the stream between the end of a rank's sweep and the end of its merge: the minimum over
            # ranks is the exchange itself (the last rank to arrive waits for nobody), the maximum adds the
            # wait of the first rank for the last (load imbalance)
            "merge_ms_per_step": {"min_over_ranks": -float(stats[3].item()), "max_over_ranks": float(stats[1].item())},
            "device_gb_per_rank_max": float(stats[2].item()) / 1e9,
            "host_wall_ms_per_step": wall_ms,
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
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--shard", default=None, choices=["orientations", "rows"])
    ap.add_argument("--size", type=int, default=None)
    ap.add_argument("--ages", type=int, default=None)
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--profile-steps", type=int, default=None, help="steps of the per-kernel timing pass (default: --steps; 0: skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dropin", action="store_true")
    ap.add_argument("--e2e-phases", action="store_true", help="diagnostic: untimed per-phase host clock of an e2e step")
    ap.add_argument("--fast", type=int, default=None, help="developer switch: 0 = simple kernels")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
