"""world_size-2 (and 4) `gloo` tests of the multi-GPU host logic on CPU: sharding, the single
all-to-all of the distributed transform, block-cyclic output layout, MSM bucket all-gather.
The per-rank compute is a CPU stand-in built on the oracle (test infrastructure); on a GPU box the
same code runs with ronkathon_b200.dist.LocalOps over NCCL."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle

GL = oracle.GOLDILOCKS


class OracleOps:
    """CPU stand-in for dist.LocalOps (same interface), backed by the oracle."""

    def __init__(self, p=GL, g=7):
        self.p, self.g = p, g

    def root_of_unity(self, n):
        return oracle.root_of_unity(self.p, n, self.g)

    def _np(self, x):
        return x.numpy().view(np.uint64)

    def ntt(self, x, log_n, batch=1, inverse=False):
        a = self._np(x)
        n = 1 << log_n
        for b in range(batch):
            a[b * n:(b + 1) * n] = oracle.ntt_fast(self.p, a[b * n:(b + 1) * n], inverse=inverse, g=self.g)
        return x

    def mul_powers(self, x, base, scale=1):
        a = self._np(x)
        cur = scale
        tab = np.empty(len(a), dtype=np.uint64)
        for i in range(len(a)):
            tab[i] = cur
            cur = oracle.mul(self.p, cur, base)
        a[:] = oracle.vec_mul(self.p, a, tab)
        return x

    def cross_dft(self, x, log_g, stride, count, inverse=False):
        a = self._np(x)
        G = 1 << log_g
        for k in range(count):
            col = a[k:k + G * stride:stride].copy()
            a[k:k + G * stride:stride] = oracle.ntt_fast(self.p, col, inverse=inverse, g=self.g) if G > 1 else col
        return x

    def msm_buckets(self, points, scalars):
        pts, sc = points.numpy().reshape(-1, 4), scalars.numpy()
        out = b""
        for s in range(17):
            acc = oracle.INF
            for i in np.nonzero(sc == s)[0]:
                acc = oracle.point_add(acc, bytes(pts[i]))
            out += oracle.INF if s == 0 else acc
        return out

    def msm_combine(self, sets):
        n = len(sets) // 68
        B = [oracle.INF] * 17
        for k in range(n):
            for s in range(17):
                B[s] = oracle.point_add(B[s], sets[68 * k + 4 * s:68 * k + 4 * s + 4])
        run, tot = oracle.INF, oracle.INF
        for s in range(16, 0, -1):
            run = oracle.point_add(run, B[s])
            tot = oracle.point_add(tot, run)
        return tot

    def sync(self):
        pass


def _worker(rank, world, port, log_n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ronkathon_b200 import dist as rd
        ops = OracleOps()
        n = 1 << log_n
        a = oracle.splitmix(GL, 42, n)
        # --- one large transform across the group ---
        local = torch.from_numpy(a[rank::world].copy().view(np.int64))
        out = rd.ntt_distributed(ops, local, log_n)
        full = rd.gather_distributed_output(out, log_n).numpy().view(np.uint64)
        ok_ntt = bool(np.array_equal(full, oracle.ntt_fast(GL, a)))
        # --- batched transforms, contiguous shards, no collective ---
        batch, lg = 10, 8
        data = oracle.splitmix(GL, 7, batch << lg)
        lo, hi = rd.shard_range(batch, rank, world)
        shard = torch.from_numpy(data[lo << lg:hi << lg].copy().view(np.int64))
        rd.ntt_batch_sharded(ops, shard, lg)
        exp = np.concatenate([oracle.ntt_fast(GL, data[b << lg:(b + 1) << lg]) for b in range(lo, hi)]) if hi > lo else np.empty(0, np.uint64)
        ok_batch = bool(np.array_equal(shard.numpy().view(np.uint64), exp))
        # --- MSM: index-range shards + bucket all-gather ---
        rng = np.random.default_rng(3)
        table = [oracle.point_smul(bytes([1, 0, 2, 0]), k) for k in range(17)]
        pts = np.frombuffer(b"".join(table[int(k)] for k in rng.integers(0, 17, 200)), dtype=np.uint8).copy().reshape(-1, 4)
        sc = rng.integers(0, 17, 200).astype(np.uint8)
        lo, hi = rd.shard_range(200, rank, world)
        got = rd.msm_distributed(ops, torch.from_numpy(pts[lo:hi].copy()), torch.from_numpy(sc[lo:hi].copy()))
        ok_msm = got == oracle.commit(sc, pts, fast=True)
        q.put((rank, ok_ntt, ok_batch, ok_msm))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,log_n", [(2, 10), (4, 12)])
def test_distributed_logic_gloo(world, log_n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + world * 7 + log_n
    procs = [ctx.Process(target=_worker, args=(r, world, port, log_n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == list(range(world))
    for rank, ok_ntt, ok_batch, ok_msm in res:
        assert ok_ntt and ok_batch and ok_msm, (rank, ok_ntt, ok_batch, ok_msm)


def test_shard_range_covers_everything():
    from ronkathon_b200.dist import shard_range
    for total in (0, 1, 7, 8, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
