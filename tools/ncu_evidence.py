#!/usr/bin/env python3
"""Workload for the ncu captures behind profiles/r02_*: one call of each kernel family the round-1 evidence lacked
 — inverse and fused-multiply 2^24 transforms, the 2^16 shapes of BASELINE config 5, the
element-wise field kernels, the div_linear scan of kzg::open and kzg::commit at 2^20 terms (histogram kernels).
Run under ncu with `-k regex:<pattern>`; without ncu it just executes once (and checks nothing: parity lives in
tests/)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import msm_terms  # noqa: E402
from ronkathon_b200 import Context, ops, _lib  # noqa: E402

GL = 0xFFFFFFFF00000001
torch.cuda.set_device(0)
ctx = Context(0, torch.cuda.current_stream().cuda_stream)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = 3
if which in ("all", "ntt24"):
    a = ops.splitmix_fill(ctx, 1 << 24, 1, GL, "cuda")
    m = ops.splitmix_fill(ctx, 1 << 24, 2, GL, "cuda")
    for _ in range(reps):
        ops.ntt_(ctx, a, 24)
        ops.ntt_mul_(ctx, a, m, 24)
        ops.ntt_(ctx, a, 24, inverse=True)
if which in ("all", "ntt16"):
    b = ops.splitmix_fill(ctx, 512 << 16, 3, GL, "cuda")
    for _ in range(reps):
        ops.ntt_(ctx, b, 16, 512)
        ops.ntt_(ctx, b, 16, 512, inverse=True)
if which in ("all", "ntt20"):   # BASELINE config 2: passes A1 / A2 of the tile kernel + the radix-16 pass C
    c2 = ops.splitmix_fill(ctx, 1 << 20, 6, GL, "cuda")
    for _ in range(reps):
        ops.ntt_(ctx, c2, 20)
        ops.ntt_(ctx, c2, 20, inverse=True)
if which in ("all", "ntt16c"):  # one 2^16-point transform: the 16-CTA cluster kernel
    c3 = ops.splitmix_fill(ctx, 1 << 16, 7, GL, "cuda")
    for _ in range(reps):
        ops.ntt_(ctx, c3, 16)
        ops.ntt_(ctx, c3, 16, inverse=True)
if which in ("all", "field"):
    x = ops.splitmix_fill(ctx, 1 << 24, 4, GL, "cuda")
    y = ops.splitmix_fill(ctx, 1 << 24, 5, GL, "cuda")
    out = torch.empty_like(x)
    for name in ("ronk_field_add_u64", "ronk_field_sub_u64", "ronk_field_mul_u64"):
        for _ in range(reps):
            ctx.call(name, GL, _lib._ptr(x), _lib._ptr(y), _lib._ptr(out), x.numel())
    q = torch.empty(1 << 24, dtype=torch.int64, device="cuda")
    rem = torch.empty(1, dtype=torch.int64, device="cuda")
    for _ in range(reps):   # Polynomial::div by (x - z): the kzg::open quotient
        ctx.call("ronk_poly_div_linear_u64", GL, _lib._ptr(x), x.numel(), GL - 5, 1, _lib._ptr(q), _lib._ptr(rem))
if which in ("all", "msm"):
    pts, sc = msm_terms(1 << 20)
    P, S = torch.from_numpy(pts).cuda(), torch.from_numpy(sc).cuda()
    for _ in range(reps):
        ops.msm(ctx, P, S)
if which == "msm24":
    pts, sc = msm_terms(1 << 24)
    P, S = torch.from_numpy(pts).cuda(), torch.from_numpy(sc).cuda()
    for _ in range(reps):
        ops.msm(ctx, P, S)
ctx.sync()
print("ok")
