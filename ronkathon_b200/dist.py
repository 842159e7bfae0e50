"""Multi-GPU layer: one process per GPU, `torch.distributed` for the plumbing.

Three shardings of the hot path:

* batched transforms (BASELINE config 5): independent units, contiguous batch ranges per rank,
  NO collective on the data path (`shard_range`, `ntt_batch_sharded`);
* one large transform across G = world_size GPUs: rank r owns the decimated slice a[r::G]; a local
  n/G-point NTT, the twiddle column ω_n^(r·k'), ONE all-to-all, then G-point butterflies across the
  received blocks (`ntt_distributed`).  Output is block-cyclic: rank s ends with
  X[s·m/G + k'' + m·q] at position q·(m/G) + k'' (m = n/G);
* kzg::commit: index-range shards → 17 bucket sums per rank → one all-gather of 68 bytes → local
  combine (`msm_distributed`).

The per-rank compute goes through `LocalOps`; the default implementation calls the CUDA kernels in
libronk_b200.so.  Tests substitute a CPU stand-in so the sharding / exchange logic runs under the
`gloo` backend without a GPU (the stand-in lives in tests/, not here: no CPU path in the product).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib
from ._lib import GOLDILOCKS


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [lo, hi) range of `total` units owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class LocalOps:
    """Per-rank compute used by the distributed algorithms, on libronk_b200.so."""

    def __init__(self, ctx, p: int = GOLDILOCKS, g: int = 7):
        self.ctx, self.p, self.g = ctx, p, g
        # torch's collectives and caching allocator are ordered on torch's CURRENT stream, the kernels on the
        # context's: the algorithms below interleave the two without events, which is only correct when they are
        # the same stream.  Refuse anything else instead of racing silently.
        if torch.cuda.is_available() and hasattr(ctx, "stream"):
            cur = torch.cuda.current_stream().cuda_stream
            if int(ctx.stream) != int(cur):
                raise ValueError("LocalOps needs a Context created on torch's current stream "
                                 f"(context stream {ctx.stream:#x}, torch current stream {cur:#x})")

    def root_of_unity(self, n: int) -> int:
        import ctypes as C
        out = C.c_uint64()
        if _lib.lib().ronk_root_of_unity(self.p, self.g, n, C.byref(out)) != 0:
            raise _lib.RonkPanic(1, "n must divide p - 1")
        return out.value

    def ntt(self, x, log_n: int, batch: int = 1, inverse: bool = False):
        self.ctx.call("ronk_ntt_u64", self.p, self.g, _lib._ptr(x), log_n, batch, int(inverse))
        return x

    def mul_powers(self, x, base: int, scale: int = 1):
        """x[i] *= scale·base^i"""
        tab = torch.empty_like(x)
        self.ctx.call("ronk_field_powers_u64", self.p, base, scale, _lib._ptr(tab), x.numel())
        self.ctx.call("ronk_field_mul_u64", self.p, _lib._ptr(x), _lib._ptr(tab), _lib._ptr(x), x.numel())
        return x

    def cross_dft(self, x, log_g: int, stride: int, count: int, inverse: bool = False):
        self.ctx.call("ronk_ntt_strided_small_u64", self.p, self.g, _lib._ptr(x), log_g, stride, count, int(inverse))
        return x

    def msm_buckets(self, points, scalars) -> bytes:
        from . import ops
        return ops.msm_buckets(self.ctx, points, scalars)

    def msm_combine(self, sets: bytes) -> bytes:
        from . import ops
        return ops.msm_combine(self.ctx, sets)

    def sync(self):
        self.ctx.sync()


def ntt_batch_sharded(ops: LocalOps, shard, log_n: int, inverse: bool = False):
    """Config-5 path: this rank's contiguous slice of the batch, transformed in place. No collective."""
    n = 1 << log_n
    assert shard.numel() % n == 0
    return ops.ntt(shard, log_n, shard.numel() // n, inverse)


def ntt_distributed(ops: LocalOps, local, log_n: int, group=None):
    """Forward transform of ONE 2^log_n-point polynomial spread cyclically over the group.

    `local` holds a[rank::G] (length m = n/G, int64 view of uint64 residues).  Returns a tensor of
    length m laid out [q][k''] with value X[(rank·m/G + k'') + m·q]."""
    G = dist.get_world_size(group)
    r = dist.get_rank(group)
    log_g = G.bit_length() - 1
    assert 1 << log_g == G, "world size must be a power of two"
    n = 1 << log_n
    m = n // G
    assert local.numel() == m and m % G == 0
    ops.ntt(local, log_n - log_g)                       # Y_r[k'] = Σ_j a[r + Gj] ω_m^(jk')
    if r:
        ops.mul_powers(local, pow(ops.root_of_unity(n), r, ops.p))  # Z_r[k'] = ω_n^(r k') Y_r[k']
    ops.sync()
    recv = torch.empty_like(local)
    dist.all_to_all_single(recv, local, group=group)    # recv[r'][k''] = Z_r'[rank·m/G + k'']
    ops.cross_dft(recv, log_g, m // G, m // G)          # X[k' + m q] = Σ_r' ω_G^(r' q) Z_r'[k']
    ops.sync()
    return recv


def gather_distributed_output(out, log_n: int, group=None):
    """Collect ntt_distributed's block-cyclic outputs into natural order on every rank (tests / small n)."""
    G = dist.get_world_size(group)
    n = 1 << log_n
    m = n // G
    parts = [torch.empty_like(out) for _ in range(G)]
    dist.all_gather(parts, out, group=group)
    full = torch.empty(n, dtype=out.dtype, device=out.device)
    for s, part in enumerate(parts):
        blk = part.view(G, m // G)
        for q in range(G):
            lo = s * (m // G) + m * q
            full[lo:lo + m // G] = blk[q]
    return full


def msm_distributed(ops: LocalOps, points_shard, scalars_shard, group=None) -> bytes:
    """kzg::commit over index-range shards: all-gather of the 17 bucket sums, local combine."""
    mine = ops.msm_buckets(points_shard, scalars_shard)
    G = dist.get_world_size(group)
    send = torch.tensor(list(mine), dtype=torch.uint8)
    backend = dist.get_backend(group)
    if backend == "nccl":
        send = send.cuda()
    parts = [torch.empty_like(send) for _ in range(G)]
    dist.all_gather(parts, send, group=group)
    return ops.msm_combine(b"".join(bytes(p.cpu().tolist()) for p in parts))


class FusedDistributedNTT:
    """One 2^log_n-point transform across the group with the exchange FUSED into the final kernel.

    Each rank keeps its local n/G-point transform in an IPC-exported device buffer; after a host
    barrier every rank launches `ronk_ntt_cross_rank_fused_u64`, whose threads read the peers'
    blocks directly over NVLink (P2P loads), apply the twiddle column and do the G-point
    butterflies — no NCCL all-to-all, no staging buffer.  Same output layout as `ntt_distributed`."""

    def __init__(self, ctx, log_n: int, group=None, p: int = GOLDILOCKS, g: int = 7):
        import ctypes as C
        self.ctx, self.log_n, self.group, self.p, self.g = ctx, log_n, group, p, g
        self.G = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.log_g = self.G.bit_length() - 1
        assert 1 << self.log_g == self.G and self.G >= 2
        self.m = (1 << log_n) // self.G
        self.buf = _lib.vp()
        ctx.call("ronk_dev_alloc", C.byref(self.buf), self.m * 8)
        handle = (C.c_uint8 * 64)()
        ctx.call("ronk_ipc_export", self.buf, C.cast(handle, _lib.vp))
        handles = [None] * self.G
        dist.all_gather_object(handles, bytes(handle), group=group)
        self.peers = (C.c_void_p * self.G)()
        self._opened = []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.peers[r] = self.buf.value
            else:
                ptr = _lib.vp()
                hb = (C.c_uint8 * 64).from_buffer_copy(h)
                ctx.call("ronk_ipc_open", C.cast(hb, _lib.vp), C.byref(ptr))
                self.peers[r] = ptr.value
                self._opened.append(ptr)

    def run(self, local, out=None):
        """`local`: this rank's a[rank::G] (CUDA int64 tensor, m words). Returns the block-cyclic output tensor."""
        import ctypes as C
        assert local.numel() == self.m
        if out is None:
            out = torch.empty_like(local)
        self.ctx.call("ronk_memcpy_d2d", self.buf, _lib._ptr(local), self.m * 8)
        self.ctx.call("ronk_ntt_u64", self.p, self.g, self.buf, self.log_n - self.log_g, 1, 0)
        self.ctx.sync()
        dist.barrier(group=self.group)          # every rank's Y_r is complete and visible
        self.ctx.call("ronk_ntt_cross_rank_fused_u64", self.p, self.g, C.cast(self.peers, _lib.vp), self.log_g,
                      self.rank, self.log_n, _lib._ptr(out))
        self.ctx.sync()
        dist.barrier(group=self.group)          # peers are done reading before anyone overwrites its buffer
        return out

    def close(self):
        for ptr in self._opened:
            self.ctx.call("ronk_ipc_close", ptr)
        self._opened = []
        if self.buf:
            self.ctx.call("ronk_dev_free", self.buf)
            self.buf = None


DIST_NCCL, DIST_FUSED = 0, 1


class DistContext:
    """The multi-GPU entry points of the C ABI (`ronk_dist_*`, `ronk_ntt_u64_dist`, `ronk_ntt_u64_batch_sharded`,
    `ronk_msm_pluto_ext_dist`): the library owns the NCCL communicator, the IPC-exported exchange buffers and
    the kernels; this class only carries the 128-byte ncclUniqueId from rank 0 to the other ranks over whatever
    `torch.distributed` group the host already has (gloo or nccl) — the job a Rust host would give to MPI or a
    TCP store."""

    def __init__(self, ctx, group=None):
        import ctypes as C
        self.ctx, self.group = ctx, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        uid = (C.c_uint8 * 128)()
        if self.rank == 0:
            rc = _lib.lib().ronk_dist_unique_id(C.cast(uid, _lib.vp))
            if rc != 0:
                raise _lib.RonkError(rc, "ronk_dist_unique_id: libnccl.so.2 not available")
        box = [bytes(uid)]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        uid = (C.c_uint8 * 128).from_buffer_copy(box[0])
        ctx.call("ronk_dist_init", C.cast(uid, _lib.vp), self.rank, self.world)

    def shard_range(self, total: int):
        return shard_range(total, self.rank, self.world)

    def ntt_batch_sharded(self, shard, log_n: int, total_batch: int, inverse: bool = False, p: int = GOLDILOCKS, g: int = 7):
        """This rank's contiguous range of `total_batch` independent transforms, in place; no collective."""
        import ctypes as C
        lo, hi = C.c_uint64(), C.c_uint64()
        self.ctx.call("ronk_ntt_u64_batch_sharded", p, g, _lib._ptr(shard), log_n, total_batch, int(inverse),
                      C.byref(lo), C.byref(hi))
        return lo.value, hi.value

    def ntt_dist(self, local, log_n: int, batch: int = 1, flavour: int = DIST_FUSED, p: int = GOLDILOCKS, g: int = 7):
        """`local` = [batch][n/G] with local[b][j] = a_b[rank + G·j]; in place → [batch][G][n/G²] block-cyclic."""
        assert local.numel() == batch * ((1 << log_n) // self.world)
        self.ctx.call("ronk_ntt_u64_dist", p, g, _lib._ptr(local), log_n, batch, flavour)
        return local

    def msm(self, points_shard, scalars_shard) -> bytes:
        """kzg::commit over index-range shards (device tensors); the full commitment on every rank."""
        import numpy as np
        out = np.empty(4, dtype=np.uint8)
        n = scalars_shard.numel()
        self.ctx.call("ronk_msm_pluto_ext_dist", _lib._ptr(points_shard), n, _lib._ptr(scalars_shard), n, _lib._ptr(out))
        return out.tobytes()

    def barrier(self):
        self.ctx.call("ronk_dist_barrier")

    def close(self):
        if self.ctx is not None:
            self.ctx.call("ronk_dist_finalize")
            self.ctx = None

