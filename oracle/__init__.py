"""ctypes front-end for the CPU oracle (oracle/ronk_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package `ronkathon_b200` never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libronk_oracle.so")

GOLDILOCKS = 0xFFFFFFFF00000001
INF = bytes([0xFF] * 4)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ronk_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B" if force else "all"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        u64, p64, i32, pu8 = C.c_uint64, C.POINTER(C.c_uint64), C.c_int, C.POINTER(C.c_uint8)
        sig = {
            "orc_new": (u64, [u64, u64]), "orc_add": (u64, [u64, u64, u64]), "orc_sub": (u64, [u64, u64, u64]),
            "orc_mul": (u64, [u64, u64, u64]), "orc_neg": (u64, [u64, u64]), "orc_pow": (u64, [u64, u64, u64]),
            "orc_pow_literal": (u64, [u64, u64, u64]),
            "orc_inverse": (i32, [u64, u64, p64]), "orc_div": (i32, [u64, u64, u64, p64]),
            "orc_rem": (i32, [u64, u64, u64, p64]), "orc_is_prime": (i32, [u64]),
            "orc_find_primitive_element": (u64, [u64]),
            "orc_primitive_root_of_unity": (i32, [u64, u64, u64, p64]),
            "orc_poly_eval": (u64, [u64, p64, u64, u64]), "orc_poly_degree": (u64, [p64, u64]),
            "orc_poly_leading": (u64, [p64, u64]),
            "orc_poly_pow_mult": (None, [u64, p64, u64, u64, u64, p64]),
            "orc_poly_add": (None, [u64, p64, u64, p64, u64, p64]),
            "orc_poly_sub": (None, [u64, p64, u64, p64, u64, p64]),
            "orc_poly_neg": (None, [u64, p64, u64, p64]),
            "orc_poly_mul": (None, [u64, p64, u64, p64, u64, p64]),
            "orc_poly_divrem": (i32, [u64, p64, u64, p64, u64, p64, p64]),
            "orc_dft": (i32, [u64, u64, p64, u64, p64]), "orc_fft": (i32, [u64, u64, p64, u64]),
            "orc_ifft": (i32, [u64, u64, p64, u64]),
            "orc_lagrange_eval": (i32, [u64, u64, p64, u64, u64, p64]),
            "orc_ntt_fast": (i32, [u64, u64, p64, u64, i32]),
            "orc_splitmix_fill": (None, [u64, u64, p64, u64]),
            "orc_vec_mul": (None, [u64, p64, p64, p64, u64]),
            "orc_poly_eval_horner": (u64, [u64, p64, u64, u64]),
            "orc_gf_add": (None, [pu8, pu8, pu8]), "orc_gf_sub": (None, [pu8, pu8, pu8]),
            "orc_gf_neg": (None, [pu8, pu8]), "orc_gf_mul": (None, [pu8, pu8, pu8]),
            "orc_gf_inv": (i32, [pu8, pu8]),
            "orc_point_on_curve": (i32, [pu8]), "orc_point_add": (i32, [pu8, pu8, pu8]),
            "orc_point_neg": (None, [pu8, pu8]), "orc_point_double": (i32, [pu8, pu8]),
            "orc_point_smul": (i32, [pu8, u64, pu8]),
            "orc_commit": (i32, [pu8, u64, pu8, u64, pu8]),
            "orc_commit_fast": (i32, [pu8, u64, pu8, u64, pu8]),
            "orc_setup": (i32, [pu8, pu8]), "orc_open": (i32, [pu8, u64, C.c_uint8, pu8, u64, pu8]),
            "orc_rs_encode": (i32, [u64, u64, p64, u64, u64, p64, p64]),
            "orc_rs_decode": (i32, [u64, p64, p64, u64, p64]),
            "orc_bench_fft_threads": (C.c_double, [u64, u64, u64, i32, p64]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(_lib, name)
            fn.restype, fn.argtypes = res, args
    return _lib


class OraclePanic(Exception):
    """Raised where the reference would panic / assert / unwrap a None."""


def _p64(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def _pu8(a: np.ndarray):
    assert a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _arr(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint64))


def _chk(rc):
    if rc != 0:
        raise OraclePanic("reference would panic here")


# ---- field ---------------------------------------------------------------------------------
def generator(p: int) -> int:
    g = lib().orc_find_primitive_element(p)
    if g == 0:
        raise OraclePanic("generator not found")
    return g


def add(p, a, b): return lib().orc_add(p, a % p, b % p)
def sub(p, a, b): return lib().orc_sub(p, a % p, b % p)
def mul(p, a, b): return lib().orc_mul(p, a % p, b % p)
def neg(p, a): return lib().orc_neg(p, a % p)
def pow_(p, a, e): return lib().orc_pow(p, a % p, e)
def pow_literal(p, a, e): return lib().orc_pow_literal(p, a % p, e)


def inverse(p, a):
    out = C.c_uint64()
    _chk(lib().orc_inverse(p, a % p, C.byref(out)))
    return out.value


def div(p, a, b):
    out = C.c_uint64()
    _chk(lib().orc_div(p, a % p, b % p, C.byref(out)))
    return out.value


def root_of_unity(p, n, g=None):
    out = C.c_uint64()
    _chk(lib().orc_primitive_root_of_unity(p, generator(p) if g is None else g, n, C.byref(out)))
    return out.value


def splitmix(p, seed, n) -> np.ndarray:
    out = np.empty(n, dtype=np.uint64)
    lib().orc_splitmix_fill(p, seed, _p64(out), n)
    return out


def vec_mul(p, a, b):
    a, b = _arr(a), _arr(b); out = np.empty(len(a), np.uint64)
    lib().orc_vec_mul(p, _p64(a), _p64(b), _p64(out), len(a)); return out


def poly_eval_horner(p, c, x):
    c = _arr(c)
    return lib().orc_poly_eval_horner(p, _p64(c), len(c), x % p)


# ---- polynomial -----------------------------------------------------------------------------
def poly_eval(p, c, x):
    c = _arr(c)
    return lib().orc_poly_eval(p, _p64(c), len(c), x % p)


def poly_add(p, a, b):
    a, b = _arr(a), _arr(b); out = np.empty(len(a), np.uint64)
    lib().orc_poly_add(p, _p64(a), len(a), _p64(b), len(b), _p64(out)); return out


def poly_sub(p, a, b):
    a, b = _arr(a), _arr(b); out = np.empty(len(a), np.uint64)
    lib().orc_poly_sub(p, _p64(a), len(a), _p64(b), len(b), _p64(out)); return out


def poly_neg(p, a):
    a = _arr(a); out = np.empty(len(a), np.uint64)
    lib().orc_poly_neg(p, _p64(a), len(a), _p64(out)); return out


def poly_mul(p, a, b):
    a, b = _arr(a), _arr(b); out = np.empty(len(a) + len(b) - 1, np.uint64)
    lib().orc_poly_mul(p, _p64(a), len(a), _p64(b), len(b), _p64(out)); return out


def poly_divrem(p, a, b):
    a, b = _arr(a), _arr(b)
    q, r = np.empty(len(a), np.uint64), np.empty(len(a), np.uint64)
    _chk(lib().orc_poly_divrem(p, _p64(a), len(a), _p64(b), len(b), _p64(q), _p64(r)))
    return q, r


def poly_pow_mult(p, a, d2, coeff):
    a = _arr(a); out = np.empty(len(a) + d2, np.uint64)
    lib().orc_poly_pow_mult(p, _p64(a), len(a), d2, coeff % p, _p64(out)); return out


def poly_degree(a): a = _arr(a); return lib().orc_poly_degree(_p64(a), len(a))
def poly_leading(a): a = _arr(a); return lib().orc_poly_leading(_p64(a), len(a))


def dft(p, a, g=None):
    a = _arr(a); out = np.empty(len(a), np.uint64)
    _chk(lib().orc_dft(p, generator(p) if g is None else g, _p64(a), len(a), _p64(out))); return out


def fft(p, a, g=None):
    """Faithful recursive radix-2 fft (polynomial/mod.rs:273-323)."""
    v = _arr(a).copy()
    _chk(lib().orc_fft(p, generator(p) if g is None else g, _p64(v), len(v))); return v


def ifft(p, a, g=None):
    v = _arr(a).copy()
    _chk(lib().orc_ifft(p, generator(p) if g is None else g, _p64(v), len(v))); return v


def ntt_fast(p, a, inverse=False, g=None):
    v = _arr(a).copy()
    _chk(lib().orc_ntt_fast(p, generator(p) if g is None else g, _p64(v), len(v), int(inverse))); return v


def lagrange_eval(p, c, x, g=None):
    c = _arr(c); out = C.c_uint64()
    _chk(lib().orc_lagrange_eval(p, generator(p) if g is None else g, _p64(c), len(c), x % p, C.byref(out)))
    return out.value


def rs_encode(p, msg, n, g=None):
    msg = _arr(msg); xs, ys = np.empty(n, np.uint64), np.empty(n, np.uint64)
    _chk(lib().orc_rs_encode(p, generator(p) if g is None else g, _p64(msg), len(msg), n, _p64(xs), _p64(ys)))
    return xs, ys


def rs_decode(p, xs, ys, k):
    """Message::decode (codes/reed_solomon.rs:55-107) on the first k coordinates."""
    xs, ys = _arr(xs)[:k].copy(), _arr(ys)[:k].copy(); out = np.empty(k, np.uint64)
    _chk(lib().orc_rs_decode(p, _p64(xs), _p64(ys), k, _p64(out)))
    return out


# ---- GF(101^2), curve, kzg --------------------------------------------------------------------
def _b(x, n):
    a = np.ascontiguousarray(np.asarray(list(x), dtype=np.uint8)); assert a.size == n; return a


def gf_add(a, b): o = np.empty(2, np.uint8); lib().orc_gf_add(_pu8(_b(a, 2)), _pu8(_b(b, 2)), _pu8(o)); return tuple(int(v) for v in o)
def gf_sub(a, b): o = np.empty(2, np.uint8); lib().orc_gf_sub(_pu8(_b(a, 2)), _pu8(_b(b, 2)), _pu8(o)); return tuple(int(v) for v in o)
def gf_neg(a): o = np.empty(2, np.uint8); lib().orc_gf_neg(_pu8(_b(a, 2)), _pu8(o)); return tuple(int(v) for v in o)
def gf_mul(a, b): o = np.empty(2, np.uint8); lib().orc_gf_mul(_pu8(_b(a, 2)), _pu8(_b(b, 2)), _pu8(o)); return tuple(int(v) for v in o)


def gf_inv(a):
    o = np.empty(2, np.uint8); _chk(lib().orc_gf_inv(_pu8(_b(a, 2)), _pu8(o))); return tuple(int(v) for v in o)


def point(x0, x1=0, y0=0, y1=0) -> bytes:
    """Wire format of AffinePoint<PlutoExtendedCurve>: x0,x1,y0,y1 (x = x0 + x1 t)."""
    return bytes([x0, x1, y0, y1])


def on_curve(P) -> bool: return bool(lib().orc_point_on_curve(_pu8(_b(P, 4))))


def point_add(P, Q):
    o = np.empty(4, np.uint8); _chk(lib().orc_point_add(_pu8(_b(P, 4)), _pu8(_b(Q, 4)), _pu8(o))); return bytes(o)


def point_neg(P): o = np.empty(4, np.uint8); lib().orc_point_neg(_pu8(_b(P, 4)), _pu8(o)); return bytes(o)


def point_double(P):
    o = np.empty(4, np.uint8); _chk(lib().orc_point_double(_pu8(_b(P, 4)), _pu8(o))); return bytes(o)


def point_smul(P, s):
    o = np.empty(4, np.uint8); _chk(lib().orc_point_smul(_pu8(_b(P, 4)), s, _pu8(o))); return bytes(o)


def _pts(points) -> np.ndarray:
    if isinstance(points, np.ndarray):
        a = np.ascontiguousarray(points, dtype=np.uint8).reshape(-1)
    else:
        a = np.frombuffer(b"".join(bytes(p) for p in points), dtype=np.uint8).copy()
    assert a.size % 4 == 0
    return a


def commit(scalars, points, fast=False):
    pts = _pts(points); sc = np.ascontiguousarray(np.asarray(scalars, dtype=np.uint8))
    o = np.empty(4, np.uint8)
    fn = lib().orc_commit_fast if fast else lib().orc_commit
    _chk(fn(_pu8(pts), pts.size // 4, _pu8(sc), sc.size, _pu8(o)))
    return bytes(o)


def setup():
    g1, g2 = np.empty(28, np.uint8), np.empty(8, np.uint8)
    _chk(lib().orc_setup(_pu8(g1), _pu8(g2)))
    return [bytes(g1[4 * i:4 * i + 4]) for i in range(7)], [bytes(g2[4 * i:4 * i + 4]) for i in range(2)]


def open_(coeffs, z, points):
    pts = _pts(points); c = np.ascontiguousarray(np.asarray(coeffs, dtype=np.uint8)); o = np.empty(4, np.uint8)
    _chk(lib().orc_open(_pu8(c), c.size, z, _pu8(pts), pts.size // 4, _pu8(o)))
    return bytes(o)


def bench_fft_threads(p, n, threads, g=None):
    cs = C.c_uint64()
    secs = lib().orc_bench_fft_threads(p, generator(p) if g is None else g, n, threads, C.byref(cs))
    if secs < 0:
        raise OraclePanic("fft failed")
    return secs, cs.value
