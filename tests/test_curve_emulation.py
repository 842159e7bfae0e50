"""CPU-tier check of the curve arithmetic the MSM kernels use: ronkathon_b200/csrc/msm_curve.cuh compiled for the
host (tests/emu/curve_emu.cpp — a test fixture, never part of the product) against the oracle:
GF(101²) mul / inverse on every element, point addition on all 290² pairs of the order-289 subgroup + Infinity
with both inverse flavours, the on-curve test, and the scan-based Σ s·B_s of the finishing kernel."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
PU8 = C.POINTER(C.c_uint8)


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(HERE, "emu", "curve_emu.cpp")
    so = os.path.join(HERE, "emu", "libcurve_emu.so")
    hdrs = [os.path.join(HERE, "..", "ronkathon_b200", "csrc", h) for h in ("msm_curve.cuh", "field.cuh")]
    newest = max(os.path.getmtime(x) for x in [src] + hdrs)
    if not os.path.exists(so) or os.path.getmtime(so) < newest:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", so, src])
    lib = C.CDLL(so)
    lib.emu_point_add.argtypes = [PU8, PU8, PU8, C.c_uint64, C.c_int]
    lib.emu_point_valid.argtypes = [PU8, PU8, C.c_uint64]
    lib.emu_gf_op.argtypes = [C.c_int, PU8, PU8, PU8, C.c_uint64]
    lib.emu_bucket_combine.argtypes = [PU8, PU8]
    PU32 = C.POINTER(C.c_uint32)
    lib.emu_group_tables.argtypes = [PU32, PU32]
    lib.emu_group_tables.restype = C.c_int
    lib.emu_coord_commit.argtypes = [PU32, PU32, PU8, PU8, C.c_uint64, PU8]
    lib.emu_coord_commit.restype = C.c_int
    return lib


def _p(a):
    return a.ctypes.data_as(PU8)


def subgroup():
    G1, G2 = bytes([1, 0, 2, 0]), bytes([36, 0, 0, 31])
    pts = [oracle.point_add(oracle.point_smul(G1, k), oracle.point_smul(G2, l)) for k in range(17) for l in range(17)]
    assert len(set(pts)) == 289
    return pts


def test_gf101_2_mul_and_inverse_all_elements(emu):
    elems = np.array([(a, b) for a in range(101) for b in range(101)], dtype=np.uint8)
    other = np.roll(elems, 4099, axis=0).copy()
    out = np.empty_like(elems)
    emu.emu_gf_op(0, _p(elems), _p(other), _p(out), len(elems))
    for i in range(0, len(elems), 37):
        assert tuple(out[i]) == tuple(oracle.gf_mul(tuple(elems[i]), tuple(other[i])))
    emu.emu_gf_op(1, _p(elems), _p(elems), _p(out), len(elems))
    for i in range(1, len(elems)):                       # x · x^-1 == 1 for every non-zero element
        assert tuple(oracle.gf_mul(tuple(elems[i]), tuple(out[i]))) == (1, 0)


@pytest.mark.parametrize("use_table", [0, 1])
def test_point_add_all_subgroup_pairs(emu, use_table):
    pts = subgroup() + [b"\xff" * 4]
    a = np.frombuffer(b"".join(p for p in pts for _ in pts), dtype=np.uint8).copy()
    b = np.frombuffer(b"".join(q for _ in pts for q in pts), dtype=np.uint8).copy()
    out = np.empty_like(a)
    emu.emu_point_add(_p(a), _p(b), _p(out), len(pts) ** 2, use_table)
    k = 0
    for p in pts:
        for q in pts:
            assert bytes(out[4 * k:4 * k + 4]) == oracle.point_add(p, q), (p, q)
            k += 1


def test_on_curve_predicate(emu):
    pts = subgroup()
    rng = np.random.default_rng(3)
    cand = rng.integers(0, 101, size=(4000, 4), dtype=np.uint8)
    cand[:289] = np.frombuffer(b"".join(pts), dtype=np.uint8).reshape(-1, 4)
    cand[300] = (36, 0, 0, 81)        # off-curve point of src/curve/pluto_curve.rs:225-234
    cand[301] = (101, 0, 2, 0)        # non-canonical coordinate
    ok = np.empty(len(cand), dtype=np.uint8)
    emu.emu_point_valid(_p(np.ascontiguousarray(cand)), _p(ok), len(cand))
    for i, c in enumerate(cand):
        expect = bytes(c) == b"\xff" * 4 or (all(v < 101 for v in c) and oracle.on_curve(bytes(c)))  # 0xFF×4 = Infinity
        assert bool(ok[i]) == expect, (i, c)
    assert ok[:289].all() and not ok[300] and not ok[301]   # the subgroup (incl. Infinity) is valid input


def test_scan_based_bucket_combination(emu):
    """msm_finish_kernel's last warp: suffix scan + tree sum over the 16 buckets == Σ s·B_s (the reference's
    scalar multiplication by repeated addition, summed)."""
    pts = subgroup() + [b"\xff" * 4]
    rng = np.random.default_rng(9)
    for _ in range(200):
        B = [pts[i] for i in rng.integers(0, len(pts), 16)]
        expect = b"\xff" * 4
        for s, P in enumerate(B, start=1):
            expect = oracle.point_add(expect, oracle.point_smul(P, s % 17) if s < 17 else P)
        out = np.empty(4, dtype=np.uint8)
        buf = np.frombuffer(b"".join(B), dtype=np.uint8).copy()
        emu.emu_bucket_combine(_p(buf), _p(out))
        assert bytes(out) == expect


# ---- kzg::commit in group coordinates (msm_coord_kernel): the host-built tables and the per-term step ----------
MSM_BINS, EXP = 20402, 102


@pytest.fixture(scope="module")
def group_tables(emu):
    bintab = np.empty(MSM_BINS + 2, dtype=np.uint32)
    pttab = np.empty(EXP * EXP, dtype=np.uint32)
    PU32 = C.POINTER(C.c_uint32)
    assert emu.emu_group_tables(bintab.ctypes.data_as(PU32), pttab.ctypes.data_as(PU32)) == 1
    return bintab, pttab


def _unpack(w):
    w = int(w)
    return bytes([w & 0xFF, (w >> 8) & 0xFF, (w >> 16) & 0xFF, w >> 24])


def test_group_tables_are_a_bijection_onto_the_curve(group_tables):
    """pttab enumerates Infinity + 10 403 distinct curve points, bintab is its inverse and carries the y of every bin's
    point (the kernel's is_on_curve), and every bin without a curve point is marked empty."""
    bintab, pttab = group_tables
    assert int(pttab[0]) == 0xFFFFFFFF
    pts = [_unpack(w) for w in pttab[1:]]
    assert len(set(pts)) == EXP * EXP - 1 and b"\xff" * 4 not in pts
    assert all(oracle.on_curve(P) for P in pts)
    filled = 0
    for k, w in enumerate(pttab):
        if k == 0:
            continue
        P = _unpack(w)
        ybit = (P[2] > 50) if P[2] else (P[3] > 50)
        e = int(bintab[2 * (P[0] + 101 * P[1]) + int(ybit)])
        assert e & 0xFFFF == P[2] | (P[3] << 8) and (e >> 16) & 0xFF == k // EXP and e >> 24 == k % EXP
        filled += 1
    assert filled == EXP * EXP - 1
    assert int((bintab[:MSM_BINS] != 0xFFFFFFFF).sum()) == EXP * EXP - 1   # all other bins are empty


def test_group_tables_are_a_homomorphism(group_tables):
    """pttab[a][b] + pttab[c][d] == pttab[a+c][b+d] under the reference's addition law (oracle), incl. doubling,
    inverse pairs and Infinity: the property the two dot products mod 102 rest on."""
    _, pttab = group_tables
    rng = np.random.default_rng(11)
    idx = rng.integers(0, EXP, size=(3000, 4))
    idx[:EXP, 2:] = idx[:EXP, :2]                       # doubling
    idx[EXP:2 * EXP, 2:] = (-idx[EXP:2 * EXP, :2]) % EXP  # P + (-P)
    idx[2 * EXP:2 * EXP + 50, :2] = 0                   # Infinity + Q
    for a, b, c, d in idx:
        got = oracle.point_add(_unpack(pttab[EXP * a + b]), _unpack(pttab[EXP * c + d]))
        assert got == _unpack(pttab[EXP * ((a + c) % EXP) + (b + d) % EXP]), (a, b, c, d)
    # exponent 102: the basis points have order exactly 102
    for k in (EXP, 1):
        P, acc = _unpack(pttab[k]), b"\xff" * 4
        for m in range(1, EXP + 1):
            acc = oracle.point_add(acc, P)
            assert (acc == b"\xff" * 4) == (m == EXP)


def test_coord_commit_matches_oracle_commit(emu, group_tables):
    """Σ s_i·P_i through the tables == the oracle's kzg::commit (kzg/setup.rs:48-60) on arbitrary curve points with
    Infinity and zero scalars mixed in; off-curve / non-canonical / scalar ≥ 17 terms are rejected."""
    bintab, pttab = group_tables
    PU32 = C.POINTER(C.c_uint32)
    tb, tp = bintab.ctypes.data_as(PU32), pttab.ctypes.data_as(PU32)
    rng = np.random.default_rng(5)
    for n in (1, 2, 5, 17, 200, 1500):
        words = pttab[rng.integers(0, EXP * EXP, n)]
        pts = np.frombuffer(b"".join(_unpack(w) for w in words), dtype=np.uint8).reshape(n, 4).copy()
        sc = rng.integers(0, 17, n).astype(np.uint8)
        out = np.empty(4, dtype=np.uint8)
        assert emu.emu_coord_commit(tb, tp, _p(pts), _p(sc), n, _p(out)) == 0
        assert bytes(out) == oracle.commit(sc, pts, fast=True)
    pts = np.array([[36, 0, 0, 81]], dtype=np.uint8)      # off the curve (pluto_curve.rs:225-234)
    out = np.empty(4, dtype=np.uint8)
    assert emu.emu_coord_commit(tb, tp, _p(pts), _p(np.array([3], dtype=np.uint8)), 1, _p(out)) == 1
    pts = np.array([[101, 0, 2, 0]], dtype=np.uint8)      # non-canonical coordinate
    assert emu.emu_coord_commit(tb, tp, _p(pts), _p(np.array([3], dtype=np.uint8)), 1, _p(out)) == 1
    pts = np.array([[1, 0, 2, 0]], dtype=np.uint8)        # scalar not an F17 residue
    assert emu.emu_coord_commit(tb, tp, _p(pts), _p(np.array([17], dtype=np.uint8)), 1, _p(out)) == 1
