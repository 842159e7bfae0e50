#!/usr/bin/env python3
"""Timing of the BASELINE configs other than the headline metric, on one GPU (SURVEY §8d inputs):
config 2 (2^20 forward NTT), config 3 (2^23 × 2^23-coefficient product through 2^24-point transforms),
config 4 (kzg::commit, 2^20 pairs; also 2^10 … 2^24) and one GPU's share of config 5 (512 × 2^16).
Not collected by pytest; run as `python tests/config_timing.py`.  Every result is first checked against the
oracle (bit-exact, or by the size-independent property named in the code), then timed with CUDA events,
device-resident operands; per-kernel times come from the library's own event profiling.  One JSON line."""
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


GL = oracle.GOLDILOCKS


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def kernels(ctx, fn):
    ctx.prof_enable(True)
    for _ in range(5):
        fn()
    ctx.sync()
    recs = ctx.prof_fetch()
    ctx.prof_enable(False)
    k = {}
    for name, ms in recs:
        k.setdefault(name, []).append(ms)
    return {name: round(float(np.median(v)) * len(v) / 5, 4) for name, v in k.items()}   # ms per call


def main():
    torch.cuda.set_device(0)
    ctx = Context(0, torch.cuda.current_stream().cuda_stream)
    out = {}
    only_msm = "--only" in sys.argv and sys.argv[sys.argv.index("--only") + 1] == "msm"
    if not only_msm:
        ntt_configs(ctx, out)
    msm_configs(ctx, out)
    print(json.dumps(out))


def ntt_configs(ctx, out):
    # config 2: one 2^20 transform
    n = 1 << 20
    a = oracle.splitmix(GL, 42, n)
    d = torch.from_numpy(a.view(np.int64)).cuda()
    ops.ntt_(ctx, d, 20)
    ctx.sync()
    ok = np.array_equal(d.cpu().numpy().view(np.uint64), oracle.ntt_fast(GL, a))
    ms = timed(lambda: ops.ntt_(ctx, d, 20))
    out["config2_ntt_2^20"] = {"bit_exact": bool(ok), "ms": round(ms, 4), "field_muls_per_s": (n // 2) * 20 / (ms * 1e-3),
                               "kernel_ms": kernels(ctx, lambda: ops.ntt_(ctx, d, 20))}
    # config 3: a, b with 2^23 coefficients each, product of 2^24 - 1 coefficients
    m = 1 << 23
    A, B = ops.splitmix_fill(ctx, m, 42), ops.splitmix_fill(ctx, m, 43)
    c = ops.poly_mul(ctx, A, B)
    ctx.sync()
    ah, bh, ch = (t.cpu().numpy().view(np.uint64) for t in (A, B, c))
    x = 0x123456789ABCDEF % GL   # c(x) == a(x)·b(x) at a random point, c[0] and the top coefficient directly
    ok = (oracle.poly_eval_horner(GL, ch, x) == oracle.mul(GL, oracle.poly_eval_horner(GL, ah, x), oracle.poly_eval_horner(GL, bh, x))
          and int(ch[0]) == oracle.mul(GL, int(ah[0]), int(bh[0])) and int(ch[-1]) == oracle.mul(GL, int(ah[-1]), int(bh[-1])))
    ms = timed(lambda: ops.poly_mul(ctx, A, B), iters=10)
    out["config3_poly_mul_2^23x2^23"] = {"identity_checks": bool(ok), "ms": round(ms, 4),
                                         "field_muls_per_s": 637534208 / (ms * 1e-3),
                                         "kernel_ms": kernels(ctx, lambda: ops.poly_mul(ctx, A, B))}
    del A, B, c
    # one GPU's share of config 5: 512 transforms of 2^16 points
    bt = 512
    a = oracle.splitmix(GL, 42, bt << 16)
    d = torch.from_numpy(a.view(np.int64)).cuda()
    ops.ntt_(ctx, d, 16, batch=bt)
    ctx.sync()
    h = d.cpu().numpy().view(np.uint64)
    ok = all(np.array_equal(h[i << 16:(i + 1) << 16], oracle.ntt_fast(GL, a[i << 16:(i + 1) << 16])) for i in (0, 255, 511))
    ms = timed(lambda: ops.ntt_(ctx, d, 16, batch=bt))
    out["config5_share_512x2^16"] = {"bit_exact_sampled": bool(ok), "ms": round(ms, 4),
                                     "field_muls_per_s": bt * (1 << 15) * 16 / (ms * 1e-3),
                                     "kernel_ms": kernels(ctx, lambda: ops.ntt_(ctx, d, 16, batch=bt))}
    del d
    # single transforms of every two-pass size (ms), to see the mid-size regime
    sweep = {}
    for lg in range(14, 25):
        dd = ops.splitmix_fill(ctx, 1 << lg, 42)
        sweep[f"2^{lg}"] = round(timed(lambda: ops.ntt_(ctx, dd, lg)), 4)
        del dd
    out["single_transform_ms"] = sweep


def msm_configs(ctx, out):
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
        out[f"config4_msm_2^{log_n}"] = {"bit_exact": bool(ok), "call_ms": round(call_ms, 4),
                             "kernel_ms": {name: round(float(np.median(v)), 4) for name, v in k.items()},
                             "point_adds_per_s": n / (call_ms * 1e-3)}


if __name__ == "__main__":
    main()
