"""§8f "next" rows, one thin layer above the hot path (all arithmetic in libronk_b200.so):

* Reed–Solomon `Message::encode::<N>` (src/codes/reed_solomon.rs:42-52): the codeword is the
  message polynomial evaluated at the N-th roots of unity ω_N^i — i.e. `Polynomial::dft` of the
  message zero-padded to N coefficients (a plain NTT when N is a power of two).
* Reed–Solomon `Message::decode` (reed_solomon.rs:55-107): Lagrange interpolation through the first
  K coordinates (`ronk_poly_interpolate_u64_host`).
* Shamir `split` (src/shamir/mod.rs:53-58): `Polynomial::evaluate` at x = 1..n, one batched kernel.
"""
from __future__ import annotations

import numpy as np

from .polynomial import Polynomial


def rs_encode(message, n: int, field):
    """Returns [(x_i, y_i)] with x_i = ω_n^i, y_i = m(x_i)  (reed_solomon.rs:42-52).
    Panics (RonkPanic) if n ∤ p-1, like primitive_root_of_unity."""
    k = len(message)
    assert n >= k, "codeword must be at least as long as the message"
    w = field.primitive_root_of_unity(n)
    padded = Polynomial(list(message) + [0] * (n - k), field)
    ys = padded.fft() if n & (n - 1) == 0 and n > 1 else padded.dft()
    xs = Polynomial([1] + [0] * (n - 1), field)  # x_i = ω^i: evaluate X at the roots, i.e. dft of the monomial X
    xs.coefficients = np.roll(xs.coefficients, 1) if n > 1 else xs.coefficients
    xvals = xs.dft().coefficients if n > 1 else np.array([1 % field.ORDER], dtype=np.uint64)
    assert int(xvals[1 % n]) == w.value or n == 1
    return [(field(int(x)), field(int(y))) for x, y in zip(xvals, ys.coefficients)]


def rs_decode(codeword, k: int, field):
    """Message::decode (reed_solomon.rs:55-107): the message is the interpolant through the first
    k coordinates of the codeword, one O(k²) device interpolation.  codeword: [(x_i, y_i)]."""
    from . import _lib
    assert len(codeword) >= k, "codeword must be at least as long as the message"
    xs = np.array([int(getattr(x, "value", x)) for x, _ in codeword[:k]], dtype=np.uint64)
    ys = np.array([int(getattr(y, "value", y)) for _, y in codeword[:k]], dtype=np.uint64)
    out = np.empty(k, dtype=np.uint64)
    _lib.default_context().call("ronk_poly_interpolate_u64_host", field.ORDER, _lib._ptr(xs), _lib._ptr(ys), k,
                                _lib._ptr(out))
    return [field(int(v)) for v in out]


def shamir_shares(coefficients, n: int, field):
    """Evaluations of the sharing polynomial at x = 1..n (shamir/mod.rs:53-58), one kernel launch."""
    poly = Polynomial(coefficients, field)
    return list(zip(range(1, n + 1), poly.evaluate_many(range(1, n + 1))))
