#!/usr/bin/env python
"""bench.py - grid-cell-years/s of the two HDP hot paths (thresholds + heatwave metrics) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--cells C]

One "step" = thresholds for every measure of the workload, then the full percentile x definition metric
sweep for every measure (BASELINE.json configs[1]: CMIP6-like 180x360, 30 y + 86 y, 3 measures, 10 x 6).
Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for how every field is obtained.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "grid-cell-years/s (threshold+metric)"
UNIT = "grid-cell-years/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cmip6_1deg")
    ap.add_argument("--cells", type=int, default=0, help="override the number of cells per GPU (debugging only)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=8.0, help="target CPU time per path for the cpu_baseline sample")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(label: str, cells: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel(s) in `label`, from the committed ncu
    `--set full` capture of this round (profiles/traffic.json, written by tools/ncu_summary.py), scaled to `cells`."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        k = json.load(f)["kernels"]
    names = [n for n in label.split("+") if n in k]             # (the hand-over pass of k_thr_seg behind k_thr_cand is not captured: ~0 bytes)
    if not names:
        return None
    return sum(k[n]["dram_bytes"] * cells / k[n]["cells"] for n in names)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms.  It is started before the warm-up (nvidia-smi needs a few
    hundred ms to deliver its first line) and `stop(t0, t1)` keeps the samples whose timestamps fall inside the timed region
    [t0, t1] (time.time()); a region shorter than the sampling period falls back to the samples within 0.5 s around it, which
    were taken under the same load (warm-up steps before, end-to-end steps after)."""
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_id: str):
        self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", gpu_id, f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=self.file, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self, t0: float, t1: float):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)                                   # let the sample that covers the end of the region arrive
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.file.flush()
        self.file.seek(0)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for line in self.file.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(parts[1]), float(parts[2]), [n for n, v in zip(names, parts[4:8]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        self.file.close()
        os.unlink(self.file.name)
        inside = [r for r in rows if t0 <= r[0] <= t1]
        where = "timed region"
        if not inside:
            inside = [r for r in rows if t0 - 0.5 <= r[0] <= t1 + 0.5]
            where = "within 0.5 s of the timed region (region shorter than the sampling period)"
        if inside:
            out.update(sm_mhz=float(np.median([r[1] for r in inside])), sm_max_mhz=float(max(r[2] for r in inside)),
                       reasons=sorted({n for r in inside for n in r[3]}), samples=len(inside), sampled=where)
        return out


# --------------------------------------------------------------------------------------------------
# CPU legs (oracle port): cpu_baseline of our arm and the whole of --impl reference
# --------------------------------------------------------------------------------------------------
def host_cores() -> int:
    """Cores this process may run on.  torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU legs are meant to use every
    host core, so they ask the oracle for this many threads explicitly instead of relying on the OpenMP default."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def bind_to_gpu_numa_node(gpu_id: str):
    """Multi-GPU boxes have several NUMA nodes: run this rank on the cores next to its GPU, so that the pinned host buffers of
    the end-to-end leg are allocated (first touch) in the memory its PCIe root is attached to.  Returns the previous affinity
    (the CPU baseline leg restores it: it is meant to use every host core)."""
    try:
        import pynvml
        previous = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByUUID(gpu_id) if gpu_id.startswith("GPU-") else pynvml.nvmlDeviceGetHandleByIndex(int(gpu_id))
        n_words = (max(previous) + 64) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, n_words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (word >> b) & 1} & previous
        if cpus:
            os.sched_setaffinity(0, cpus)
        return previous
    except Exception:                                      # no NVML / no permission: stay where we are
        return None


def cpu_tables(wl):
    from hdp_b200 import _tables as tb
    wt = wl.window_tables()
    if not wl.run_years:
        return wt, None, None
    return wt, tb.doy_map(wl.run_axis().dayofyr), wl.seasons()


def cpu_pass(wl, base, run, is_south, threads=None):
    """One pass of the reference algorithm (oracle port, OpenMP over cells) on [T, n] host arrays.
    Returns (seconds thresholds, seconds metrics, thresholds, metrics)."""
    import oracle
    threads = threads or host_cores()
    wt, dm, st = cpu_tables(wl)
    win = wt.window_samples()
    t0 = time.perf_counter()
    thr = oracle.thresholds_batch(base, win, wl.percentiles, threads)
    t1 = time.perf_counter()
    met = None
    if run is not None:
        met = oracle.metrics_batch(run, thr, dm, wl.defs, st.north, st.south, is_south, threads)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1, thr, met


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port of its Numba kernels; the Python/Numba
    reference itself cannot travel to the GPU box) on all host cores, on a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from hdp_b200 import workloads, synth
    wl = workloads.get(args.workload)
    cores = host_cores()
    lat, _ = synth.grid_latitudes(wl.n_lat, wl.n_lon)
    n = args.cells or max(8 * cores, 64)
    sel = np.linspace(0, wl.cells - 1, n).astype(np.int64)
    lat_s = lat[sel]
    base = synth.gridded_field(lat_s, wl.base_axis().dayofyr, seed=1234, device="cpu").numpy()
    run = synth.gridded_field(lat_s, wl.run_axis().dayofyr, seed=2234, trend=4.0, device="cpu").numpy() if wl.run_years else None
    south = (lat_s < 0).astype(np.uint8)
    for _ in range(max(args.warmup, 1)):
        cpu_pass(wl, base[:, : min(n, cores)], None if run is None else run[:, : min(n, cores)], south[: min(n, cores)])
    t_thr = t_met = 0.0
    for _ in range(args.steps):
        a, b, _, _ = cpu_pass(wl, base, run, south)
        t_thr += a; t_met += b
    total = t_thr + t_met
    cy = n * (wl.base_years + wl.run_years) * args.steps          # one measure per sampled cell
    value = cy / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
        "config": {"workload": wl.name, "description": wl.description, "percentiles": len(wl.percentiles),
                   "definitions": len(wl.defs), "sample_cells": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n} cells of {wl.name} (one measure) per step, thresholds + metric sweep, oracle/hdp_oracle.c with OpenMP",
                         "thresholds_cell_years_per_s": n * wl.base_years * args.steps / max(t_thr, 1e-9),
                         "metrics_cell_years_per_s": (n * wl.run_years * args.steps / t_met) if t_met > 0 else None},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, warnings) was diverted to stderr."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    args = parse_args()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)            # libraries that print to fd 1 (NCCL's version banner) must not pollute the JSON line
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from hdp_b200 import _core, _tables as tb, synth, workloads

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: hdp_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    gpu_id = str(getattr(torch.cuda.get_device_properties(dev), "uuid", local_rank))
    if not gpu_id.startswith("GPU-") and len(gpu_id) > 8:
        gpu_id = "GPU-" + gpu_id
    full_affinity = bind_to_gpu_numa_node(gpu_id) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    wl = workloads.get(args.workload)
    wt = wl.window_tables()
    st = wl.seasons() if wl.run_years else None                       # thresholds-only workloads (era5_025deg) have no run
    dm = tb.doy_map(wl.run_axis().dayofyr) if wl.run_years else None
    lat, _ = synth.grid_latitudes(wl.n_lat, wl.n_lon)
    if args.cells:
        lat = lat[np.linspace(0, wl.cells - 1, args.cells).astype(np.int64)]
    C = lat.size
    south = torch.as_tensor((lat < 0).astype(np.uint8), device=dev)
    q, defs = wl.percentiles, wl.defs
    P, D, Y, n_doy = len(q), len(defs), (st.n_years if st else 0), wt.n_doy
    offsets = [5.0, -5.0, 0.0][: wl.measures]                       # tmax / tmin / tavg

    # weak scaling: every rank owns a full grid (its own ensemble member, seeded by rank), no collective on the path
    base, run = [], []
    for m, off in enumerate(offsets):
        seed = 1234 + 10 * m + 1000 * rank
        base.append(synth.gridded_field(lat, wl.base_axis().dayofyr, seed=seed, offset=off, device=dev))
        run.append(synth.gridded_field(lat, wl.run_axis().dayofyr, seed=seed + 1, offset=off, trend=4.0, device=dev)
                   if wl.run_years else None)
    thr = [torch.empty((C, n_doy, P), dtype=torch.float64, device=dev) for _ in offsets]
    out = [torch.empty((4, P, D, Y, C), dtype=torch.uint16, device=dev) if wl.run_years else None for _ in offsets]

    def step():
        for m in range(len(offsets)):
            _core.thresholds_array(base[m], wt, q, out=thr[m])
            if run[m] is not None:
                _core.metrics_array(run[m], thr[m], dm, defs, st.north, st.south, south, out=out[m])

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(gpu_id) if rank == 0 else None
    for _ in range(args.warmup):
        step()
    sync_all()

    _core.timing_enable(True)
    _core.timing_read()
    launches0 = _core.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_region0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    sync_all()
    t_region1 = time.time()
    elapsed_ms = e0.elapsed_time(e1)
    launches = _core.launch_count() - launches0
    kern = _core.timing_read()
    _core.timing_enable(False)
    clocks = sampler.stop(t_region0, t_region1) if sampler else None
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())

    cy_per_step = wl.cell_years(C) * world
    value = cy_per_step * args.steps / (elapsed_ms / 1e3)

    # ---- per-kernel times (CUDA events on the launching stream, inside the timed region) -> roofline
    by_kernel = {}
    for name, ms in kern:
        by_kernel.setdefault(name, []).append(ms)
    # one "launch" below = everything a kernel does for ONE measure (k_thr_cell runs as several cell-chunk launches per measure)
    passes = args.steps * len(offsets)
    kernel_ms = {k: float(np.sum(v)) / passes for k, v in by_kernel.items()}
    kernel_launches = {k: len(v) / passes for k, v in by_kernel.items()}
    kernel_share = {k: float(np.sum(v)) / max(sum(np.sum(x) for x in by_kernel.values()), 1e-9) for k, v in by_kernel.items()}
    thr_kernels = ("k_thr_generic", "k_thr_cand", "k_thr_seg", "k_thr_ranked", "normalize")
    alg_bytes = {k: wl.bytes_thresholds(C) for k in thr_kernels}
    alg_bytes.update({"k_hot_words": wl.bytes_metrics(C), "k_scan": wl.bytes_metrics(C)})
    peak, peak_src = peaks()
    dominant = max(kernel_share, key=kernel_share.get) if kernel_share else None
    roofline = None
    if dominant:
        # thresholds: the whole path-1 bytes are charged to its dominant kernel; metrics: path-2 bytes are charged to
        # the PAIR k_hot_words + k_scan (they are one pass split in two launches), so use their summed duration.
        if dominant in ("k_hot_words", "k_scan"):
            dur = kernel_ms.get("k_hot_words", 0.0) + kernel_ms.get("k_scan", 0.0)
            label = "k_hot_words+k_scan"
        else:
            dur = sum(kernel_ms.get(k, 0.0) for k in thr_kernels)         # the transposing copy is part of path 1
            label = "+".join(k for k in thr_kernels if k in kernel_ms)
        achieved = alg_bytes[dominant] / (dur / 1e3) / 1e9
        roofline = {"bound": "hbm", "kernel": label, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": measured_traffic(label, C), "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes[dominant],
                    "kernel_ms": kernel_ms, "kernel_share": kernel_share, "kernel_launches_per_measure": kernel_launches,
                    "unit_of_launch": "all launches of the kernel for one measure (64 800 cells)"}
    # whole-step roofline: all algorithmic bytes of the step over the step time
    step_bytes = wl.measures * (wl.bytes_thresholds(C) + (wl.bytes_metrics(C) if wl.run_years else 0))
    step_gbs = step_bytes / (elapsed_ms / args.steps / 1e3) / 1e9

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32+f64", "data": "synthetic",
        "config": {"workload": wl.name, "description": wl.description, "cells_per_gpu": int(C), "base_days": len(wl.base_axis()),
                   "run_days": len(wl.run_axis()) if wl.run_years else 0, "measures": wl.measures, "percentiles": P, "definitions": D,
                   "window_radius": wl.radius, "sharding": f"cells x{world} (one full grid per GPU, no collective)",
                   "l2": "inputs per step far larger than L2 (no flush needed)",
                   "host_numa_binding": bool(full_affinity)},
        "gpu_launches": int(launches),
        "step_roofline": {"algorithmic_bytes_per_step": int(step_bytes), "achieved_gbs": step_gbs, "frac": step_gbs / peak},
        "roofline": roofline,
        "clocks": clocks,
    }

    # ---- e2e: the same step through the host-buffer C ABI (hdp_b200_*_host), H2D and D2H inside the timed region
    if not args.no_e2e and wl.run_years:
        h_base = torch.empty(base[0].shape, dtype=torch.float32, pin_memory=True)
        h_run = torch.empty(run[0].shape, dtype=torch.float32, pin_memory=True)
        h_thr = torch.empty(thr[0].shape, dtype=torch.float64, pin_memory=True)
        h_out = torch.empty(out[0].shape, dtype=torch.uint16, pin_memory=True)
        h_base.copy_(base[0]); h_run.copy_(run[0])
        torch.cuda.synchronize()
        south_h = (lat < 0).astype(np.uint8)

        def e2e_step():
            # host measures are reused for the M measures (same bytes moved as M distinct measures)
            for _ in range(wl.measures):
                _core.thresholds_host(h_base.numpy(), wt, q, out=h_thr.numpy())
                _core.metrics_host(h_run.numpy(), h_thr.numpy(), dm, defs, st.north, st.south, south_h, out=h_out.numpy())

        # release the resident arrays of the device-only leg so the host pipeline has room
        e2e_step()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        sync_all()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        thr_bytes = C * n_doy * P * 8
        h2d = wl.measures * (h_base.numel() * 4 + h_run.numel() * 4 + thr_bytes + C)
        d2h = wl.measures * (thr_bytes + h_out.numel() * 2)
        line["e2e"] = {"value": cy_per_step * args.e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                       "d2h_bytes_per_step": int(d2h), "steps": args.e2e_steps, "ms_per_step": 1e3 * dt / args.e2e_steps,
                       "api": "hdp_b200_thresholds_host + hdp_b200_metrics_host (pinned host buffers)"}
        # parity of the two arms on this run's data: host pipeline vs device-resident results of measure 0
        line["e2e"]["matches_device_path"] = bool(torch.equal(h_out.view(torch.int16).to(dev), out[0].view(torch.int16)) and
                                                  torch.equal(h_thr.to(dev).view(torch.int64), thr[0].view(torch.int64)))

    # ---- cpu_baseline: oracle port on a bounded sample of measure 0, all host cores, rank 0 only; also a parity check
    if rank == 0 and not args.no_cpu:
        if full_affinity:
            os.sched_setaffinity(0, full_affinity)         # the CPU baseline uses every host core
        cores = host_cores()
        n_probe = min(C, max(cores, 8))
        sel = np.linspace(0, C - 1, n_probe).astype(np.int64)
        sel_t = torch.as_tensor(sel, device=dev)

        def fetch(n_sel):
            b = base[0][:, n_sel].cpu().numpy()
            r = run[0][:, n_sel].cpu().numpy() if wl.run_years else None
            return b, r

        b_h, r_h = fetch(sel_t)
        a, b, _, _ = cpu_pass(wl, b_h, r_h, (lat[sel] < 0).astype(np.uint8))          # calibration (also warms the caches)
        per_cell = (a + b) / n_probe
        n = int(min(C, max(n_probe, 2 * args.cpu_seconds / max(per_cell, 1e-6))))
        sel = np.linspace(0, C - 1, n).astype(np.int64)
        sel_t = torch.as_tensor(sel, device=dev)
        b_h, r_h = fetch(sel_t)
        t_thr, t_met, thr_cpu, met_cpu = cpu_pass(wl, b_h, r_h, (lat[sel] < 0).astype(np.uint8))
        cpu_cy = n * (wl.base_years + wl.run_years)
        line["cpu_baseline"] = {
            "value": cpu_cy / (t_thr + t_met), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} cells of measure 0 of {wl.name} (thresholds + full metric sweep), oracle/hdp_oracle.c with OpenMP",
            "thresholds_cell_years_per_s": n * wl.base_years / max(t_thr, 1e-9),
            "metrics_cell_years_per_s": (n * wl.run_years / t_met) if t_met > 0 else None,
            "seconds": t_thr + t_met,
        }
        thr_gpu = thr[0][sel_t].cpu().numpy()
        parity = {"thresholds_bit_exact": bool(np.array_equal((thr_gpu + 0.0).view(np.uint64), (thr_cpu + 0.0).view(np.uint64)))}
        if met_cpu is not None:
            got = out[0].view(torch.int16)[..., sel_t].cpu().numpy().view(np.uint16).astype(np.int64).transpose(1, 2, 4, 0, 3)
            parity["metrics_bit_exact"] = bool(np.array_equal(got, met_cpu))
            hot = float(np.mean(r_h > thr_cpu[np.arange(n)[None, :], dm[:, None], 0]))
            line["config"]["hot_day_fraction_lowest_percentile"] = hot
        line["parity_on_sample"] = parity

    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
