"""GPU parity: AffinePoint arithmetic and kzg::commit (Pippenger MSM) through the C ABI vs the
reference KATs (src/curve/pluto_curve.rs tests, src/kzg/tests.rs) and the oracle."""
import numpy as np
import pytest

import oracle
from conftest import pt
from gpu_util import ctx, msm_inputs

pytestmark = pytest.mark.gpu


def test_curve_kats(kats):
    from ronkathon_b200 import AffinePoint, G1_GENERATOR, G2_GENERATOR, PlutoScalarField, RonkPanic
    ctx()
    c = kats["curve"]
    g = G1_GENERATOR
    for k, v in c["multiples_of_G1"].items():
        assert g * PlutoScalarField(int(k)) == AffinePoint(bytes(v))
    assert g.double() == AffinePoint(bytes(c["multiples_of_G1"]["2"]))
    assert g + g.double() == AffinePoint(bytes(c["multiples_of_G1"]["3"]))
    for a, b in c["negatives"]:
        assert -AffinePoint(bytes(a)) == AffinePoint(bytes(b))
    assert G2_GENERATOR.double() == AffinePoint(bytes(c["two_G2"]))
    inf = AffinePoint.infinity()
    assert g + inf == g and inf + g == g and g + (-g) == inf
    assert AffinePoint.new(36, (0, 31)) == G2_GENERATOR
    with pytest.raises(RonkPanic):          # pluto_curve.rs:204-213 false_point()
        AffinePoint.new(36, (0, 81))
    for G in (G1_GENERATOR, G2_GENERATOR):  # order 17
        acc = G
        for _ in range(16):
            acc = acc + G
        assert acc == inf
    assert g * PlutoScalarField(0) == inf


def test_point_add_exhaustive_subgroup_vs_oracle():
    """All 290×290 sums inside the 17-torsion (incl. Infinity, P+P, P+(-P)) in one kernel launch."""
    c = ctx()
    pts, _ = msm_inputs(1)
    G1, G2 = bytes([1, 0, 2, 0]), bytes([36, 0, 0, 31])
    elems = [oracle.INF] + [oracle.point_add(oracle.point_smul(G1, k), oracle.point_smul(G2, l))
                            for k in range(17) for l in range(17)]
    elems = list(dict.fromkeys(elems))
    A = np.frombuffer(b"".join(x for x in elems for _ in elems), dtype=np.uint8).copy()
    B = np.frombuffer(b"".join(y for _ in elems for y in elems), dtype=np.uint8).copy()
    out = np.empty_like(A)
    c.call("ronk_point_add_pluto_ext_host", A.ctypes.data, B.ctypes.data, out.ctypes.data, len(A) // 4)
    exp = b"".join(oracle.point_add(x, y) for x in elems for y in elems)
    assert out.tobytes() == exp


def test_kzg_kats(kats):
    from ronkathon_b200 import AffinePoint, RonkPanic, kzg
    ctx()
    k = kats["kzg"]
    g1, g2 = kzg.setup()
    assert [p.raw for p in g1] == [bytes(v) for v in k["g1srs"]]
    assert [p.raw for p in g2] == [bytes(v) for v in k["g2srs"]]
    for c in k["commit"]:
        assert kzg.commit(c["coeffs"], g1).raw == pt(c["out"])
    for o in k["open"]:
        assert kzg.open_(o["coeffs"], o["z"], g1).raw == bytes(o["out"])
    with pytest.raises(RonkPanic):           # kzg/setup.rs:53
        kzg.commit([1] * 8, g1)
    assert kzg.commit([], g1) == AffinePoint.infinity()   # empty sum → Infinity (curve/mod.rs:219-223)
    assert kzg.commit([0, 0, 0], g1) == AffinePoint.infinity()


@pytest.mark.parametrize("n", [1, 2, 17, 1000, 4097, 1 << 15])
def test_msm_vs_literal_commit(n):
    """Pippenger buckets == the reference's literal Σ (repeated-addition scalar mul) loop."""
    import torch
    from ronkathon_b200 import ops
    c = ctx()
    pts, sc = msm_inputs(n, 44 + n, 45 + n)
    got = ops.msm(c, torch.from_numpy(pts).cuda(), torch.from_numpy(sc).cuda())
    assert got == oracle.commit(sc, pts, fast=True)
    if n <= 4097:
        assert got == oracle.commit(sc, pts)


def test_config4_msm_2_20_and_bucket_combine():
    """BASELINE config 4 at full size, plus the multi-GPU combine path: per-shard buckets folded
    by ronk_msm_combine_buckets_host must equal the single-device commit."""
    import torch
    from ronkathon_b200 import ops
    c = ctx()
    n = 1 << 20
    pts, sc = msm_inputs(n)
    P, S = torch.from_numpy(pts).cuda(), torch.from_numpy(sc).cuda()
    got = ops.msm(c, P, S)
    assert got == oracle.commit(sc, pts, fast=True)
    shards = 8
    sets = b"".join(ops.msm_buckets(c, P[i * n // shards:(i + 1) * n // shards].contiguous(),
                                    S[i * n // shards:(i + 1) * n // shards].contiguous()) for i in range(shards))
    assert ops.msm_combine(c, sets) == got
    # general curve points outside the 17-torsion: 3·(every on-curve point found by scanning x)
    allpts = []
    for x0 in range(101):
        for y0 in range(101):
            for y1 in (0, 1):
                b = bytes([x0, 0, y0, y1])
                if oracle.on_curve(b):
                    allpts.append(b)
    arr = np.frombuffer(b"".join(allpts), dtype=np.uint8).copy().reshape(-1, 4)
    scs = (np.arange(len(arr)) % 17).astype(np.uint8)
    assert ops.msm(c, torch.from_numpy(arr).cuda(), torch.from_numpy(scs).cuda()) == oracle.commit(scs, arr, fast=True)


def test_msm_rejects_bad_input():
    import torch
    from ronkathon_b200 import RonkPanic, ops
    c = ctx()
    pts, sc = msm_inputs(64)
    bad = pts.copy(); bad[10] = [36, 0, 0, 81]           # off-curve
    with pytest.raises(RonkPanic):
        ops.msm(c, torch.from_numpy(bad).cuda(), torch.from_numpy(sc).cuda())
    bad_sc = sc.copy(); bad_sc[5] = 17                    # not a canonical F17 residue
    with pytest.raises(RonkPanic):
        ops.msm(c, torch.from_numpy(pts).cuda(), torch.from_numpy(bad_sc).cuda())
    bad = pts.copy(); bad[3] = [101, 0, 2, 0]             # non-canonical coordinate
    with pytest.raises(RonkPanic):
        ops.msm(c, torch.from_numpy(bad).cuda(), torch.from_numpy(sc).cuda())


def _full_group_points():
    """The affine points of E(F_101²) with x1 ∈ {0, 1} and x0 ≡ 0 (mod 4), found by scanning — far outside the
    17-torsion."""
    out = []
    for x0 in range(0, 101, 4):
        for x1 in (0, 1):
            for y0 in range(101):
                for y1 in range(101):
                    b = bytes([x0, x1, y0, y1])
                    if oracle.on_curve(b):
                        out.append(b)
    rng = np.random.default_rng(8)                       # widen: sums of random pairs (the oracle's addition law)
    for _ in range(3000):
        i, j = rng.integers(0, len(out), 2)
        q = oracle.point_add(out[i], out[j])
        if q != b"\xff" * 4:
            out.append(q)
    return np.frombuffer(b"".join(out), dtype=np.uint8).copy().reshape(-1, 4)


def test_commit_paths_agree_on_the_full_group_and_unaligned_input():
    """kzg::commit three ways — group coordinates (default), point histogram (RONK_MSM_COORD=0), Pippenger buckets
    (RONK_MSM_COORD=0 RONK_MSM_HIST=0) — on points of the whole curve group with Infinity and zero scalars mixed in,
    against the oracle; sizes around the 4-term vector width and views that break the 16-byte alignment."""
    import os
    import torch
    from ronkathon_b200 import Context, ops
    base = _full_group_points()
    rng = np.random.default_rng(21)
    ctxs = {"coord": ctx()}
    for name, env in (("hist", {"RONK_MSM_COORD": "0"}), ("buckets", {"RONK_MSM_COORD": "0", "RONK_MSM_HIST": "0"})):
        os.environ.update(env)
        try:
            ctxs[name] = Context(0, torch.cuda.current_stream().cuda_stream)
        finally:
            for k in env:
                os.environ.pop(k)
    for n in (1, 3, 4, 5, 63, 1021, 65537, (1 << 18) + 7):
        pts = base[rng.integers(0, len(base), n)].copy()
        pts[rng.integers(0, n, max(1, n // 50))] = 0xFF            # Infinity terms
        sc = rng.integers(0, 17, n).astype(np.uint8)
        want = oracle.commit(sc, pts, fast=True)
        P, S = torch.from_numpy(pts).cuda(), torch.from_numpy(sc).cuda()
        for name, c in ctxs.items():
            assert ops.msm(c, P, S) == want, (name, n)
        if n > 8:   # drop 1 / 3 leading terms: the point pointer is no longer 16-byte aligned, the scalar pointer odd
            for k in (1, 3):
                assert ops.msm(ctxs["coord"], P[k:], S[k:]) == oracle.commit(sc[k:], pts[k:], fast=True), (n, k)
    # repeated calls on one context: the kernel's global accumulators clean themselves
    pts, sc = msm_inputs(5000)
    P, S = torch.from_numpy(pts).cuda(), torch.from_numpy(sc).cuda()
    want = oracle.commit(sc, pts, fast=True)
    for _ in range(5):
        assert ops.msm(ctxs["coord"], P, S) == want
    # a rejected call leaves nothing behind either
    bad = pts.copy(); bad[77] = [36, 0, 0, 81]
    from ronkathon_b200 import RonkPanic
    with pytest.raises(RonkPanic):
        ops.msm(ctxs["coord"], torch.from_numpy(bad).cuda(), S)
    assert ops.msm(ctxs["coord"], P, S) == want
