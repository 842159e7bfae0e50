#!/usr/bin/env python3
"""Device-resident timing of batched transforms (CUDA events, ms per call): python tools/time_sizes.py 20:1 20:4 20:16 16:512"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ronkathon_b200 import Context, ops  # noqa: E402

GL = 0xFFFFFFFF00000001
torch.cuda.set_device(0)
ctx = Context(0, torch.cuda.current_stream().cuda_stream)
out = {}
for spec in sys.argv[1:]:
    lg, batch = (int(v) for v in spec.split(":"))
    d = ops.splitmix_fill(ctx, batch << lg, 3, GL, "cuda")
    for _ in range(3):
        ops.ntt_(ctx, d, lg, batch)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 30
    a.record()
    for _ in range(iters):
        ops.ntt_(ctx, d, lg, batch)
    b.record()
    torch.cuda.synchronize()
    out[spec] = round(a.elapsed_time(b) / iters, 4)
    del d
print(out)
