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
