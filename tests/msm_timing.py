#!/usr/bin/env python3
"""kzg::commit timing on one GPU (BASELINE config 4: 2^20 (point, scalar) pairs, SURVEY §8d inputs).
Not collected by pytest; run as `python tests/msm_timing.py`.  Checks the result against the oracle's
bucket commit, then reports the whole call (device-resident inputs, 4-byte result back on the host)
and the two kernels separately (the library's own CUDA-event profiling).  One JSON line."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from gpu_util import msm_inputs  # noqa: E402
from ronkathon_b200 import Context, ops  # noqa: E402


def main():
    torch.cuda.set_device(0)
    ctx = Context(0, torch.cuda.current_stream().cuda_stream)
    out = {}
    for log_n in (10, 16, 20, 24):
        n = 1 << log_n
        pts, sc = msm_inputs(n)
        P, S = torch.from_numpy(pts).cuda(), torch.from_numpy(sc).cuda()
        got = ops.msm(ctx, P, S)
        ok = got == oracle.commit(sc, pts, fast=True)
        for _ in range(3):
            ops.msm(ctx, P, S)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        a.record()
        for _ in range(iters):
            ops.msm(ctx, P, S)
        b.record()
        torch.cuda.synchronize()
        call_ms = a.elapsed_time(b) / iters
        ctx.prof_enable(True)
        for _ in range(5):
            ops.msm(ctx, P, S)
        ctx.sync()
        recs = ctx.prof_fetch()
        ctx.prof_enable(False)
        k = {}
        for name, ms in recs:
            k.setdefault(name, []).append(ms)
        out[f"2^{log_n}"] = {"bit_exact": bool(ok), "call_ms": round(call_ms, 4),
                             "kernel_ms": {name: round(float(np.median(v)), 4) for name, v in k.items()},
                             "point_adds_per_s": n / (call_ms * 1e-3)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
