#!/usr/bin/env python3
"""Multi-GPU validation + timing (run under torchrun, one rank per GPU, NCCL over NVLink):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29511 tools/multi_gpu_check.py

Checks, bit-exactly against the oracle: (1) one 2^20-point transform spread over the N GPUs
(local NTT → twiddle column → ONE all-to-all → cross-rank butterflies), (2) BASELINE config 5's
batched transforms sharded by contiguous batch ranges (no collective), (3) kzg::commit with
index-range shards + 68-byte bucket all-gather.  Then times config 5 (4096 × 2^16 over N GPUs) and
the distributed 2^24 transform with CUDA events, max over ranks.  Prints one JSON line on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from ronkathon_b200 import Context, ops  # noqa: E402
from ronkathon_b200 import dist as rd  # noqa: E402

GL = oracle.GOLDILOCKS


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = Context(local, torch.cuda.current_stream().cuda_stream)
    lo = rd.LocalOps(ctx)
    res = {"n_gpus": world}

    # (1) one large transform across the group
    lg = 20
    a = oracle.splitmix(GL, 42, 1 << lg)
    out = rd.ntt_distributed(lo, ops.to_device(a[rank::world].copy(), dev), lg)
    full = ops.to_host(rd.gather_distributed_output(out, lg))
    res["dist_ntt_2_20_bit_exact"] = bool(np.array_equal(full, oracle.ntt_fast(GL, a)))

    # (1b) the same transform with the exchange fused into the final kernel (P2P loads over NVLink)
    fused = rd.FusedDistributedNTT(ctx, lg)
    out_f = fused.run(ops.to_device(a[rank::world].copy(), dev))
    full_f = ops.to_host(rd.gather_distributed_output(out_f, lg))
    res["fused_dist_ntt_2_20_bit_exact"] = bool(np.array_equal(full_f, oracle.ntt_fast(GL, a)))
    fused.close()

    # (2) config-5 shape, reduced batch for the oracle check: 64 × 2^16
    batch, lgb = 64, 16
    data = oracle.splitmix(GL, 7, batch << lgb)
    b0, b1 = rd.shard_range(batch, rank, world)
    shard = ops.to_device(data[b0 << lgb:b1 << lgb].copy(), dev)
    rd.ntt_batch_sharded(lo, shard, lgb)
    ctx.sync()
    got = ops.to_host(shard)
    ok = True
    for b in {b0, (b0 + b1) // 2, b1 - 1}:
        ok &= bool(np.array_equal(got[(b - b0) << lgb:(b - b0 + 1) << lgb], oracle.ntt_fast(GL, data[b << lgb:(b + 1) << lgb])))
    res["batched_sharded_bit_exact"] = ok

    # (3) MSM 2^20 terms, index-range shards
    from gpu_util import msm_inputs
    n = 1 << 20
    pts, sc = msm_inputs(n)
    i0, i1 = rd.shard_range(n, rank, world)
    got = rd.msm_distributed(lo, torch.from_numpy(pts[i0:i1].copy()).to(dev), torch.from_numpy(sc[i0:i1].copy()).to(dev))
    if rank == 0:
        res["msm_2_20_bit_exact"] = got == oracle.commit(sc, pts, fast=True)

    def timed(fn, iters=10, warm=3):
        for _ in range(warm):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # config 5 at full size: 4096 × 2^16, strong scaling (total fixed)
    total_batch = 4096
    b0, b1 = rd.shard_range(total_batch, rank, world)
    buf = ops.splitmix_fill(ctx, (b1 - b0) << 16, 100 + rank, GL, dev)
    ms = timed(lambda: rd.ntt_batch_sharded(lo, buf, 16))
    res["config5_ms"] = ms
    res["config5_field_muls_per_s"] = total_batch * (1 << 15) * 16 / (ms * 1e-3)

    # one 2^24 transform across the group (capacity mode)
    lg = 24
    locbuf = ops.splitmix_fill(ctx, (1 << lg) // world, 5 + rank, GL, dev)
    scratch = locbuf.clone()

    def one():
        scratch.copy_(locbuf)
        rd.ntt_distributed(lo, scratch, lg)
    res["dist_ntt_2_24_ms"] = timed(one, iters=5)
    fused24 = rd.FusedDistributedNTT(ctx, lg)
    outbuf = torch.empty_like(locbuf)
    res["fused_dist_ntt_2_24_ms"] = timed(lambda: fused24.run(locbuf, outbuf), iters=5)
    fused24.close()

    # MSM 2^20 timing
    P, S = torch.from_numpy(pts[i0:i1].copy()).to(dev), torch.from_numpy(sc[i0:i1].copy()).to(dev)
    res["msm_2_20_ms"] = timed(lambda: rd.msm_distributed(lo, P, S), iters=5)
    if rank == 0:
        print(json.dumps(res), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
