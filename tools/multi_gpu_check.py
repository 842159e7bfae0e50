#!/usr/bin/env python3
"""Multi-GPU parity checks through the C ABI's ronk_dist_* entry points (run under torchrun, one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29511 tools/multi_gpu_check.py

Bit-exact against the oracle: the distributed transform (both exchange flavours) over a sweep of sizes and batch
counts — including local transforms that are single-tile, two-pass and 4096-point-per-tile — with Goldilocks and
with a generic-modulus (Montgomery) field; contiguous batch shards with an uneven split; kzg::commit over
index-range shards (incl. empty shards and a rejected term seen by every rank); two contexts alternating in one
process (device binding).  bench.py --gpus N carries the full-size timings (`multi` block).
Prints one JSON line on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from bench import msm_terms  # noqa: E402
from ronkathon_b200 import Context, RonkPanic, ops  # noqa: E402
from ronkathon_b200 import dist as rd  # noqa: E402

GL = oracle.GOLDILOCKS


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = Context(local, torch.cuda.current_stream().cuda_stream)
    dctx = rd.DistContext(ctx)
    res = {"n_gpus": world}
    lg_w = world.bit_length() - 1

    def gather_host(t):
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t.contiguous())
        return [ops.to_host(p) for p in parts]

    def check_dist(p, g, log_n, batch, flavour):
        n, m = 1 << log_n, (1 << log_n) // world
        blk = m // world
        full = [oracle.splitmix(p, 900 + 31 * b + log_n, n) for b in range(batch)]   # same on every rank
        loc = ops.to_device(np.concatenate([a[rank::world] for a in full]), dev)
        dctx.ntt_dist(loc, log_n, batch, flavour, p=p, g=g)
        ctx.sync()
        outs = gather_host(loc)
        ok = True
        for b in range(batch):
            X = np.empty(n, dtype=np.uint64)
            for s in range(world):
                o = outs[s][b * m:(b + 1) * m].reshape(world, blk)
                for q in range(world):
                    X[s * blk + m * q: s * blk + m * q + blk] = o[q]
            ok = ok and bool(np.array_equal(X, oracle.ntt_fast(p, full[b], g=g)))
        return ok

    sweep = {}
    for flavour, fname in ((rd.DIST_NCCL, "nccl"), (rd.DIST_FUSED, "fused")):
        for log_n in sorted({max(2 * lg_w, 4), 8, 12, 15, 16, 20, 12 + lg_w + 12 if 24 + lg_w <= 26 else 24}):
            if log_n < 2 * lg_w:
                continue
            for batch in (1, 3):
                if log_n >= 24 and batch > 1:
                    continue
                sweep[f"{fname}_2^{log_n}_x{batch}"] = check_dist(GL, 7, log_n, batch, flavour)
        # generic modulus through the Montgomery policy: p = 2^32·k + 1 style prime with 2-adicity 32 → use 0xFFFFFFFF00000001
        # with a non-default generator (routes to MontField), and a 31-bit NTT prime
        sweep[f"{fname}_mont_gl_g5_2^12"] = check_dist(GL, pow(7, 5, GL), 12, 2, flavour)
        sweep[f"{fname}_mont_p2013265921_2^10"] = check_dist(2013265921, 31, 10, 2, flavour)   # 15·2^27 + 1
    res["dist_ntt"] = sweep
    res["dist_ntt_all_bit_exact"] = all(sweep.values())

    # contiguous batch shards, uneven split (world does not divide 13 for world ≥ 2 except …)
    total, lgb = 13, 14
    data = oracle.splitmix(GL, 7, total << lgb)
    lo, hi = dctx.shard_range(total)
    shard = ops.to_device(data[lo << lgb:hi << lgb].copy() if hi > lo else np.zeros(0, dtype=np.uint64), dev)
    got_lo, got_hi = dctx.ntt_batch_sharded(shard if hi > lo else torch.zeros(1, dtype=torch.int64, device=dev), lgb, total)
    ctx.sync()
    ok = (got_lo, got_hi) == (lo, hi)
    for b in range(lo, hi):
        ok = ok and bool(np.array_equal(ops.to_host(shard)[(b - lo) << lgb:(b - lo + 1) << lgb],
                                        oracle.ntt_fast(GL, data[b << lgb:(b + 1) << lgb])))
    flags = [None] * world
    dist.all_gather_object(flags, bool(ok))
    res["batch_sharded_uneven_bit_exact"] = all(flags)

    # kzg::commit over index-range shards: 2^16 + 5 terms, a 3-term input (empty shards), a rejected term
    commits = {}
    for n in ((1 << 16) + 5, 3):
        pts, sc = msm_terms(n)
        i0, i1 = dctx.shard_range(n)
        P = torch.from_numpy(pts[i0:i1].copy()).to(dev) if i1 > i0 else torch.zeros((1, 4), dtype=torch.uint8, device=dev)
        S = torch.from_numpy(sc[i0:i1].copy()).to(dev) if i1 > i0 else torch.zeros(0, dtype=torch.uint8, device=dev)
        got = dctx.msm(P, S)
        commits[str(n)] = bool(got == oracle.commit(sc, pts, fast=True))
    pts, sc = msm_terms(1 << 12)
    sc[len(sc) - 1] = 17            # not an F17 residue, lands in the last rank's shard
    i0, i1 = dctx.shard_range(len(sc))
    try:
        dctx.msm(torch.from_numpy(pts[i0:i1].copy()).to(dev), torch.from_numpy(sc[i0:i1].copy()).to(dev))
        rejected = False
    except RonkPanic:
        rejected = True
    flags = [None] * world
    dist.all_gather_object(flags, rejected)
    commits["rejected_on_every_rank"] = all(flags)
    res["commit_sharded"] = commits

    # two contexts for two devices alternating in ONE process (rank 0 only; needs ≥ 2 visible GPUs)
    if rank == 0 and torch.cuda.device_count() >= 2:
        other = Context(1, 0)
        a = oracle.splitmix(GL, 3, 1 << 16)
        d0 = ops.to_device(a, dev)
        d1 = ops.to_device(a, torch.device("cuda", 1))
        for _ in range(3):
            ops.ntt_(ctx, d0, 16)
            ops.ntt_(other, d1, 16)
            ops.ntt_(ctx, d0, 16, inverse=True)
            ops.ntt_(other, d1, 16, inverse=True)
        ctx.sync(); other.sync()
        res["two_contexts_one_process"] = bool(np.array_equal(ops.to_host(d0), a) and np.array_equal(ops.to_host(d1), a)
                                               and torch.cuda.current_device() == local)
        other.close()
    dist.barrier()
    dctx.close()
    if rank == 0:
        print(json.dumps(res), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
