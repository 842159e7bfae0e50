"""Shared helpers for the -m gpu tests: one Context on torch's current stream, seeded inputs."""
import numpy as np
import pytest

import oracle

GL = oracle.GOLDILOCKS
_ctx = None


def ctx():
    global _ctx
    import torch
    from ronkathon_b200 import Context, set_default_context
    if _ctx is None:
        assert torch.cuda.is_available(), "GPU tests need a B200"
        torch.cuda.set_device(0)
        _ctx = Context(0, torch.cuda.current_stream().cuda_stream)
        set_default_context(_ctx)
    return _ctx


def dev(a):
    from ronkathon_b200 import ops
    return ops.to_device(np.ascontiguousarray(a, dtype=np.uint64))


def host(t):
    from ronkathon_b200 import ops
    ctx().sync()
    return ops.to_host(t)


def summary(x):
    idx = np.arange(1, len(x) + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return {"first": int(x[0]), "second": int(x[1]), "last": int(x[-1]),
                "sum_mod_2_64": int(np.sum(x, dtype=np.uint64)),
                "weighted_sum_mod_2_64": int(np.sum(x * idx, dtype=np.uint64)),
                "xor": int(np.bitwise_xor.reduce(x))}


def msm_inputs(n, seed_pts=44, seed_sc=45):
    """points k·G1 + l·G2 with (k,l) from splitmix(seed 44) mod 17, scalars seed 45."""
    G1, G2 = bytes([1, 0, 2, 0]), bytes([36, 0, 0, 31])
    table = {}
    for k in range(17):
        for l in range(17):
            table[(k, l)] = oracle.point_add(oracle.point_smul(G1, k), oracle.point_smul(G2, l))
    kl = oracle.splitmix(17, seed_pts, 2 * n).astype(np.int64)
    lut = np.zeros((17, 17, 4), dtype=np.uint8)
    for (k, l), v in table.items():
        lut[k, l] = np.frombuffer(v, dtype=np.uint8)
    pts = lut[kl[0::2], kl[1::2]]
    sc = oracle.splitmix(17, seed_sc, n).astype(np.uint8)
    return np.ascontiguousarray(pts), sc
