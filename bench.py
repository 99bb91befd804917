#!/usr/bin/env python
"""Benchmark of the CWT hot path (BASELINE.json metric: CWT scale-points/s and HBM GB/s
vs roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload: Morlet(6) CWT of a synthetic linear chirp, N = 2^20,
256 scales (s0=2, dj=1/16, J=255), fp64.  One "step" = one full transform of one signal
(forward FFT + every scale).  Under torchrun every rank transforms its own signal on its
own GPU (weak scaling, no data-path collective); `value` = all ranks' scale-points / max
time over ranks.

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
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

N0 = 2 ** 20
DT, S0, DJ, J = 1.0, 2.0, 1.0 / 16, 255
F0 = 6.0
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernels, from the
# `ncu --set full` captures summarised under profiles/ (filled in per round; None = not captured)
# r1: PassBBody<double,1>, 16-row launch: 268.5 MB read + 215.5 MB written = 30.25 MB per row
# (algorithmic 16.78 MB per row: the Z intermediate of the two-kernel scales is read back from
# DRAM); scaled to the 32-row launches of the bench step.
# DRAM bytes (read + write) per scale ROW of the W-writing kernels, from the `ncu --set full`
# captures summarised in profiles/r1/ncu_r1_*.txt (16-row launches): PassB<1024> 484.84 MB,
# PassB<512> 478.56 MB, Single<1024> 214.78 MB, Direct<8> 217.69 MB per launch.
TRAFFIC_PER_ROW = {"PassBBody<double, 1, 1024>": 484.84e6 / 16, "PassBBody<double, 1, 512>": 478.56e6 / 16,
                   "SingleBody": 214.78e6 / 16, "DirectBody": 217.69e6 / 16}


def traffic_per_row(kernel_name):
    return TRAFFIC_PER_ROW.get(kernel_name, TRAFFIC_PER_ROW.get(kernel_name.split("<")[0]))


METRIC = "cwt_scale_points_per_sec"
UNIT = "scale-points/s"


def chirp(n, phase=0.0):
    t = np.arange(n) / n
    return np.sin(2 * np.pi * (50 * t + (n / 8) * t ** 2) + phase)


def scales():
    return S0 * 2 ** (np.arange(0, J + 1) * DJ)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    return rank, world, local, dist


def dist_barrier(dist, local):
    if dist is not None:
        import torch
        dist.barrier(device_ids=[local])
        torch.cuda.synchronize()


def dist_max(dist, local, value):
    if dist is None:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device="cuda:%d" % local)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def cpu_baseline(sample_scales, workers):
    """Oracle (numpy/scipy port of the reference algorithm) on a bounded sample of the
    same workload: `sample_scales` of the 256 scales, evenly spread, full N."""
    from oracle import cwt_oracle as orc
    x = chirp(N0)
    sj = scales()
    idx = np.linspace(0, len(sj) - 1, sample_scales).round().astype(int)
    lam = orc.Morlet(F0).flambda()
    fr = 1.0 / (lam * sj[idx])
    t0 = time.perf_counter()
    W = orc.cwt(x, DT, wavelet=orc.Morlet(F0), freqs=fr, workers=workers)[0]
    dt = time.perf_counter() - t0
    assert W.shape == (sample_scales, N0)
    return sample_scales * N0 / dt, dt


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm for this path (the oracle port;
    the reference is pure Python and cannot travel to the GPU box), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = 16
    for _ in range(args.warmup and 1):
        cpu_baseline(4, cores)
    tot_t, tot_pts = 0.0, 0
    for _ in range(args.steps):
        v, dt = cpu_baseline(sample, cores)
        tot_t += dt
        tot_pts += sample * N0
    value = tot_pts / tot_t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d of 256 scales (evenly spread), full N=2^20, per step; "
                                   "scipy.fft with workers=%d" % (sample, cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config():
    return {"workload": "config2: Morlet(6) CWT, synthetic chirp N=2^20, 256 scales "
                        "(s0=2, dj=1/16, J=255), fp64, one signal per GPU",
            "n": N0, "scales": J + 1, "wavelet": "morlet(6)",
            "l2": "no explicit flush: each step writes 4.29 GB of coefficients (>> 126 MB L2) "
                  "and re-reads only the 16 MiB spectrum it just produced"}


def run_ours(args):
    rank, world, local, dist = dist_setup(args.gpus)
    import pycwt_b200 as pycwt
    from pycwt_b200 import _engine
    eng = _engine.Engine(local)
    sj = scales()
    S = len(sj)
    x = chirp(N0, phase=0.1 * rank)
    pts = S * N0

    # ---- value: inputs resident in HBM, kernels only (CUDA events in the engine) ----
    dsig = eng.dev_alloc(x.nbytes)
    eng.h2d(dsig, x)
    eng.cwt_dev(dsig, 0, N0, DT, sj, _engine.MORLET, F0, _engine.F64)   # plans + first run
    if args.warmup > 0:
        eng.bench_last(args.warmup)
    launches_per_step = eng.last_launch_count()
    sampler = ClockSampler(local)
    dist_barrier(dist, local)
    eng.sync()
    sampler.start()
    ms = eng.bench_last(args.steps)          # mean device ms per step, events on the engine stream
    eng.sync()
    dist_barrier(dist, local)
    clocks = sampler.stop()
    ms_max = dist_max(dist, local, ms)
    value = world * pts / (ms_max * 1e-3)

    prof = eng.profile_last()    # per-kernel-type event times of one more (untimed) step
    if args.kernels_only:   # for ncu: no e2e leg, no CPU baseline
        if rank == 0:
            print(json.dumps({"kernels_only": True, "ms_per_step": ms_max, "value": value,
                              "launches_per_step": launches_per_step,
                              "kernels": {k["name"]: [k["launches"], round(k["ms"], 4), k["rows"]]
                                          for k in prof}}))
        return

    # ---- e2e: public API, host buffers, H2D + D2H inside the timed region ----
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(2):   # warm-up: default engine, device buffers, the two pinned result buffers
        W, *_ = pycwt.cwt(x, DT, DJ, S0, J, pycwt.Morlet(F0))
    dist_barrier(dist, local)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        W, *_ = pycwt.cwt(x, DT, DJ, S0, J, pycwt.Morlet(F0))
    t_e2e = (time.perf_counter() - t0) / e2e_steps
    assert W.shape == (S, N0)
    t_e2e = dist_max(dist, local, t_e2e)
    e2e_val = world * pts / t_e2e

    line = None
    if rank == 0:
        peak, peak_src = measured_peaks()
        alg_bytes = pts * 16 + N0 * 8
        achieved = alg_bytes / (ms_max * 1e-3) / 1e9
        sample = 32
        cores = 1
        # the CPU leg is timed at N = 1 only (rank 0 of a multi-rank job reports null)
        cpu_v, cpu_t = cpu_baseline(sample, cores) if world == 1 else (None, None)
        # dominant kernel = the type with the largest share of the step; the kernels that
        # write W (Single/Direct/PassB) carry 16 B of algorithmic bytes per scale-point,
        # PassA/Band launches are intermediate work of the same scales (0 algorithmic bytes).
        writers = [k for k in prof if k["name"].split("<")[0] in ("SingleBody", "DirectBody", "PassBBody")
                   and not k["name"].endswith("-1>")]
        dom = max(prof, key=lambda k: k["ms"])
        domw = max(writers, key=lambda k: k["ms"])
        kern_ms = sum(k["ms"] for k in prof)
        dom_bytes = domw["rows"] * N0 * 16
        dom_rf = {"kernel": domw["name"], "launches_per_step": domw["launches"],
                  "ms_per_step": domw["ms"], "share_of_step": domw["ms"] / kern_ms,
                  "rows": domw["rows"], "algorithmic_bytes": dom_bytes,
                  "achieved": dom_bytes / (domw["ms"] * 1e-3) / 1e9}
        # traffic and algorithmic bytes are both per (average) launch of the dominant kernel
        tpr = traffic_per_row(domw["name"])
        dom_traffic = None if tpr is None else tpr * domw["rows"] / domw["launches"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(x.nbytes),
                    "d2h_bytes_per_step": int(pts * 16), "ms_per_step": 1e3 * t_e2e,
                    "steps": e2e_steps},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": {"bound": "hbm", "achieved": dom_rf["achieved"], "peak": peak, "unit": "GB/s",
                         "frac": dom_rf["achieved"] / peak, "traffic": dom_traffic,
                         "peak_source": peak_src, "kernel": dom_rf["kernel"],
                         "launches_per_step": dom_rf["launches_per_step"],
                         "ms_per_step": dom_rf["ms_per_step"], "share_of_step": dom_rf["share_of_step"],
                         "algorithmic_bytes_per_step": dom_bytes,
                         "algorithmic_bytes_per_launch": dom_bytes / domw["launches"],
                         "largest_kernel_any": dom["name"],
                         "step": {"achieved": achieved, "frac": achieved / peak,
                                  "algorithmic_bytes": alg_bytes,
                                  # secondary figure of SURVEY 8d: 5 N log2 N flops per (un-pruned)
                                  # inverse transform of the reference algorithm, per second
                                  "nominal_fp64_tflops": S * 5.0 * N0 * np.log2(N0) / (ms_max * 1e-3) / 1e12,
                                  "note": "whole step: forward FFT + all per-scale inverse transforms"},
                         "kernels": {k["name"]: {"launches": k["launches"], "ms": round(k["ms"], 4),
                                                 "rows": k["rows"]} for k in prof}},
            "cpu_baseline": {
                "value": cpu_v, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": ("%d of 256 scales (evenly spread), full N=2^20; %.1f s" % (sample, cpu_t))
                if cpu_v is not None else "timed at N=1 only"},
            "clocks": clocks,
        }
    if dist is not None:
        dist.barrier(device_ids=[local])
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--kernels-only", action="store_true",
                    help="profiling aid: time the resident-input kernels only")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
