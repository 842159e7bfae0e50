"""Device-resident operator API (torch tensors hold the HBM buffers; libronk_b200.so does the work).

This is the path bench.py times: buffers stay on the GPU, calls are asynchronous on the
context's stream.  Tensors are int64 views of uint64 canonical residues (torch has no uint64
arithmetic; nothing here does arithmetic in torch).
"""
from __future__ import annotations

import numpy as np

from . import _lib
from ._lib import GOLDILOCKS, Context


def _check_u64(t):
    import torch
    assert t.is_cuda and t.dtype == torch.int64 and t.is_contiguous(), "need a contiguous CUDA int64 (uint64 view) tensor"


def to_device(a: np.ndarray, device="cuda"):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint64).view(np.int64)).to(device)


def to_host(t) -> np.ndarray:
    return t.cpu().numpy().view(np.uint64)


def ntt_(ctx: Context, data, log_n: int, batch: int = 1, inverse: bool = False, p: int = GOLDILOCKS, g: int = 7):
    """In-place Polynomial::fft / ifft over `batch` contiguous transforms."""
    _check_u64(data)
    assert data.numel() == (batch << log_n)
    ctx.call("ronk_ntt_u64", p, g, _lib._ptr(data), log_n, batch, int(inverse))
    return data


def ntt_mul_(ctx: Context, data, mul, log_n: int, batch: int = 1, p: int = GOLDILOCKS, g: int = 7):
    """data ← NTT(data) ⊙ mul, the point-wise product fused into the last stage."""
    _check_u64(data); _check_u64(mul)
    ctx.call("ronk_ntt_mul_u64", p, g, _lib._ptr(data), _lib._ptr(mul), log_n, batch)
    return data


def poly_mul(ctx: Context, a, b, p: int = GOLDILOCKS, g: int = 7):
    """Polynomial::mul — returns a new tensor with len(a)+len(b)-1 coefficients."""
    import torch
    _check_u64(a); _check_u64(b)
    out = torch.empty(a.numel() + b.numel() - 1, dtype=torch.int64, device=a.device)
    ctx.call("ronk_poly_mul_u64", p, g, _lib._ptr(a), a.numel(), _lib._ptr(b), b.numel(), _lib._ptr(out))
    return out


def poly_eval(ctx: Context, coeffs, xs, p: int = GOLDILOCKS):
    import torch
    _check_u64(coeffs); _check_u64(xs)
    out = torch.empty_like(xs)
    ctx.call("ronk_poly_eval_u64", p, _lib._ptr(coeffs), coeffs.numel(), _lib._ptr(xs), xs.numel(), _lib._ptr(out))
    return out


def field_binop(ctx: Context, op: str, a, b, p: int = GOLDILOCKS):
    import torch
    _check_u64(a); _check_u64(b)
    out = torch.empty_like(a)
    ctx.call({"add": "ronk_field_add_u64", "sub": "ronk_field_sub_u64", "mul": "ronk_field_mul_u64",
              "div": "ronk_field_div_u64"}[op], p, _lib._ptr(a), _lib._ptr(b), _lib._ptr(out), a.numel())
    return out


def splitmix_fill(ctx: Context, n: int, seed: int, p: int = GOLDILOCKS, device="cuda"):
    import torch
    out = torch.empty(n, dtype=torch.int64, device=device)
    ctx.call("ronk_splitmix_fill_u64", p, seed, _lib._ptr(out), n)
    return out


def msm(ctx: Context, points, scalars) -> bytes:
    """kzg::commit on device-resident packed points (uint8 [n,4]) and scalars (uint8 [n])."""
    import torch
    assert points.is_cuda and points.dtype == torch.uint8 and scalars.dtype == torch.uint8
    out = np.empty(4, dtype=np.uint8)
    ctx.call("ronk_msm_pluto_ext", _lib._ptr(points), points.numel() // 4, _lib._ptr(scalars), scalars.numel(),
             _lib._ptr(out))
    return out.tobytes()


def msm_buckets(ctx: Context, points, scalars) -> bytes:
    out = np.empty(68, dtype=np.uint8)
    ctx.call("ronk_msm_pluto_ext_buckets", _lib._ptr(points), points.numel() // 4, _lib._ptr(scalars),
             scalars.numel(), _lib._ptr(out))
    return out.tobytes()


def msm_combine(ctx: Context, bucket_sets: bytes) -> bytes:
    arr = np.frombuffer(bucket_sets, dtype=np.uint8).copy()
    out = np.empty(4, dtype=np.uint8)
    ctx.call("ronk_msm_combine_buckets_host", _lib._ptr(arr), len(arr) // 68, _lib._ptr(out))
    return out.tobytes()
