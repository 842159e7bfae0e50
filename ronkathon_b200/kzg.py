"""Host-side mirror of ronkathon's kzg module (src/kzg/setup.rs): setup / commit / open.
`commit` is the Pippenger bucket-MSM kernel in libronk_b200.so; `check` (pairing) is out of scope."""
from __future__ import annotations

import numpy as np

from . import _lib
from .curve import G1_GENERATOR, G2_GENERATOR, AffinePoint
from .field import PlutoScalarField
from .polynomial import Polynomial


def setup():
    """kzg/setup.rs:10-43: tau = 2; 7 G1 powers and 2 G2 powers."""
    tau = PlutoScalarField(2)
    g1 = [G1_GENERATOR * tau.pow(i) for i in range(7)]
    g2 = [G2_GENERATOR * tau.pow(i) for i in range(2)]
    return g1, g2


def _pack(points) -> np.ndarray:
    if isinstance(points, np.ndarray):
        return np.ascontiguousarray(points, dtype=np.uint8).reshape(-1)
    return np.frombuffer(b"".join(p.raw for p in points), dtype=np.uint8).copy()


def commit(coeffs, g1_srs) -> AffinePoint:
    """kzg/setup.rs:48-60: Σ g1_srs[i]·coeffs[i]; asserts g1_srs.len() >= coeffs.len()."""
    pts = _pack(g1_srs)
    # coefficients are PlutoScalarField values; raw ints are reduced as PlutoScalarField::new does
    # (prime/mod.rs:48-51: value % P) — never wrapped by the uint8 cast
    sc = np.array([int(getattr(c, "value", c)) % 17 for c in coeffs], dtype=np.uint8)
    out = np.empty(4, dtype=np.uint8)
    n_pts = len(pts) // 4
    # the zip stops at the shorter sequence, so only the first len(coeffs) points are shipped
    _lib.default_context().call("ronk_msm_pluto_ext_host", _lib._ptr(pts), n_pts, _lib._ptr(sc), len(sc),
                                _lib._ptr(out))
    return AffinePoint(out.tobytes())


def open_(coeffs, eval_point, g1_srs) -> AffinePoint:
    """kzg/setup.rs:63-78: poly / (x - z) by Polynomial::div, then commit the quotient."""
    poly = Polynomial(coeffs, PlutoScalarField)
    z = PlutoScalarField(getattr(eval_point, "value", eval_point))
    divisor = Polynomial([(-z).value, 1], PlutoScalarField)
    q = poly / divisor
    return commit([int(v) for v in q.coefficients], g1_srs)


def commit_lagrange(evaluations, g1_srs) -> AffinePoint:
    """Commit to a polynomial given in the LAGRANGE basis over the 2^k-th roots of unity of F17 — the form in
    which the PLONK compiler emits its selector / permutation polynomials (compiler/program.rs:118-226,
    `Polynomial<Lagrange<PlutoScalarField>, PlutoScalarField, GROUP_ORDER>`).  The prover step the reference
    never wrote: Lagrange → monomial by the inverse transform (polynomial/mod.rs:430-453, on the GPU), then
    kzg::commit (kzg/setup.rs:48-60, the MSM kernel)."""
    from .polynomial import Lagrange
    poly = evaluations if isinstance(evaluations, Polynomial) else Polynomial(evaluations, PlutoScalarField, Lagrange)
    if poly.basis is not Lagrange:
        raise _lib.RonkPanic(1, "commit_lagrange expects a Lagrange-basis polynomial")
    mono = poly.ifft()
    return commit([int(v) for v in mono.coefficients], g1_srs)


def commit_preprocessed(polys: dict, g1_srs) -> dict:
    """Commitments to a `CommonPreprocessedInput` (compiler/program.rs:59-63: ql, qr, qm, qo, qc, s1, s2, s3),
    each given as GROUP_ORDER evaluations."""
    return {name: commit_lagrange(ev, g1_srs) for name, ev in polys.items()}

