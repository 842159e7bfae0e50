#!/usr/bin/env python3
"""Device-resident timing of the 2^23 x 2^23-coefficient product (config 3) and of fused-multiply transforms, ms per call."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ronkathon_b200 import Context, ops
GL = 0xFFFFFFFF00000001
torch.cuda.set_device(0)
ctx = Context(0, torch.cuda.current_stream().cuda_stream)
def timed(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return round(a.elapsed_time(b) / iters, 4)
A, B = ops.splitmix_fill(ctx, 1 << 23, 42, GL, "cuda"), ops.splitmix_fill(ctx, 1 << 23, 43, GL, "cuda")
x, m = ops.splitmix_fill(ctx, 1 << 24, 1, GL, "cuda"), ops.splitmix_fill(ctx, 1 << 24, 2, GL, "cuda")
print({"poly_mul_2^23x2^23": timed(lambda: ops.poly_mul(ctx, A, B)), "ntt_mul_2^24": timed(lambda: ops.ntt_mul_(ctx, x, m, 24)),
       "ntt_2^24": timed(lambda: ops.ntt_(ctx, x, 24))})
