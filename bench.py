#!/usr/bin/env python3
"""bench.py — headline benchmark of the hot path (BASELINE.json metric).

Metric   field-muls/s on a 2^24-coefficient forward NTT over the 64-bit prime 2^64-2^32+1
         (algorithmic count (n/2)·log2 n = 201 326 592 per transform, independent of the radix used).
Step     one in-place forward NTT of one 2^24-coefficient polynomial per GPU, operands resident in
         HBM (`value`), or pinned host → HBM → NTT → pinned host through the C ABI (`e2e`).
N > 1    one process per GPU (torchrun), independent transforms per rank, no data-path collective
         (SURVEY §8e "batched NTTs"): weak scaling, value = N·muls / max-over-ranks time.
--impl reference   the reference's own algorithm (oracle/ronk_oracle.c: faithful recursive
         fft_recursive, src/polynomial/mod.rs:295-323) on the box's host cores, one independent
         2^24 transform per thread.

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GL = 0xFFFFFFFF00000001
LOG_N = 24
N = 1 << LOG_N
MULS_PER_NTT = (N // 2) * LOG_N          # 201 326 592
ALG_BYTES_PER_NTT = 16 * N               # one read + one write of the data
METRIC = "field-muls/s on 2^24-coeff 64-bit-prime NTT"
UNIT = "field-muls/s"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
            }
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.02)
        except Exception as e:  # NVML missing: report, never fake
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def run_reference(args):
    """--impl reference: the reference algorithm on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, args.ref_threads or cores))
    for _ in range(args.warmup):
        oracle.bench_fft_threads(GL, N, threads)
    t = 0.0
    for _ in range(args.steps):
        secs, _ = oracle.bench_fft_threads(GL, N, threads)
        t += secs
    ms = 1e3 * t / args.steps
    value = threads * MULS_PER_NTT / (t / args.steps)
    sample = f"{threads} independent 2^24-point transforms per step, one per host thread (faithful recursive fft_recursive)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic (splitmix64 mod p, seed 42+thread)",
        "config": {"workload": "2^24-coeff forward NTT, Goldilocks p=2^64-2^32+1, g=7", "parallelism": f"{threads} host threads"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from ronkathon_b200 import Context, ops
    ctx = Context(local_rank, torch.cuda.current_stream().cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # synthetic operand, generated on the device (splitmix64 mod p, seed 42 + rank)
    data = ops.splitmix_fill(ctx, N, 42 + rank, GL, dev)
    ctx.sync()
    for _ in range(args.warmup):
        ops.ntt_(ctx, data, LOG_N)
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.05)

    # ---- timed region 1: whole-job throughput, operands resident in HBM -------------------------
    launches0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        ops.ntt_(ctx, data, LOG_N)
    e1.record()
    barrier()
    launches = ctx.launches - launches0
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = world * MULS_PER_NTT / (ms_per_step * 1e-3)

    # ---- timed region 2: per-kernel durations (CUDA events around every launch, same stream) ----
    ctx.prof_enable(True)
    barrier()
    for _ in range(args.steps):
        ops.ntt_(ctx, data, LOG_N)
    recs = ctx.prof_fetch()
    ctx.prof_enable(False)
    per = {}
    for name, ms in recs:
        per.setdefault(name, []).append(ms)
    kern = {k: sum(v) / len(v) for k, v in per.items()}
    dom = max(kern, key=kern.get)
    stages = {"ntt_pass1": 12, "ntt_pass2": 12}
    peak, peak_src = measured_peaks()
    # algorithmic bytes of one launch = 16·n · (butterfly stages this launch does / 24)
    alg_bytes = ALG_BYTES_PER_NTT * stages.get(dom, 24) / 24
    achieved = alg_bytes / (kern[dom] * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(dom)
    whole_ms = sum(kern.values())
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "peak_source": peak_src,
        "alg_bytes_per_launch": alg_bytes, "kernel_ms": kern,
        "whole_ntt": {"alg_bytes": ALG_BYTES_PER_NTT, "ms": whole_ms,
                      "achieved": ALG_BYTES_PER_NTT / (whole_ms * 1e-3) / 1e9,
                      "frac": ALG_BYTES_PER_NTT / (whole_ms * 1e-3) / 1e9 / peak},
    }

    # ---- timed region 3: end to end through the C ABI with HOST buffers -------------------------
    # Every step uploads its own pinned host buffer, transforms it and downloads the result
    # (ronk_ntt_u64_host_submit / _wait).  Three steps are in flight (three device slots): the upload
    # of step i+1, the kernels of step i and the download of step i-1 overlap on the full-duplex link.
    SLOTS = 3
    hosts = [torch.empty(N, dtype=torch.int64).pin_memory() for _ in range(SLOTS)]
    for h in hosts:
        h.copy_(data.cpu())
    e2e_steps = max(4, min(args.steps, 12))

    def e2e_run(steps):
        for i in range(steps):
            ctx.call("ronk_ntt_u64_host_submit", GL, 7, hosts[i % SLOTS].data_ptr(), LOG_N, 1, 0, i % SLOTS)
            if i >= SLOTS - 1:
                ctx.call("ronk_ntt_u64_host_wait", (i - (SLOTS - 1)) % SLOTS)
        for s in range(SLOTS):
            ctx.call("ronk_ntt_u64_host_wait", s)

    e2e_run(4)
    barrier()
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * MULS_PER_NTT / float(te.item())
    # single-step latency (no overlap) for reference
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        ctx.call("ronk_ntt_u64_host", GL, 7, hosts[0].data_ptr(), LOG_N, 1, 0)
    e2e_latency_ms = 1e3 * (time.perf_counter() - t0) / 3

    sampler.stop_flag = True
    sampler.join(timeout=2)
    clocks = sampler.result()

    # ---- CPU baseline beside it (rank 0, N = 1 only) ---------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle
        secs, _ = oracle.bench_fft_threads(GL, N, 1)
        cpu_baseline = {"value": MULS_PER_NTT / secs, "unit": UNIT, "cores": 1, "kind": "port",
                        "sample": "one full 2^24-point faithful recursive fft (polynomial/mod.rs:295-323 restated in C), "
                                  f"single thread as in the reference, {secs:.2f} s; host has {os.cpu_count()} cores"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic (splitmix64 mod p generated on device, seed 42+rank)",
            "config": {"workload": "2^24-coeff forward NTT, Goldilocks p=2^64-2^32+1, g=7, in place, natural order in/out",
                       "per_gpu": "1 transform per step", "parallelism": f"independent transforms × {world} GPU(s), no collective",
                       "l2": "working set 256 MiB (data + workspace) > 126 MB L2; no flush needed"},
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * N, "d2h_bytes_per_step": 8 * N,
                    "ms_per_step": 1e3 * float(te.item()), "single_step_latency_ms": e2e_latency_ms,
                    "api": "ronk_ntt_u64_host_submit/_wait, pinned host buffers, 3 steps in flight"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
