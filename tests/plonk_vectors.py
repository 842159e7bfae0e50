"""Selector / permutation polynomials of the reference's PLONK compiler tests, as evaluation vectors over F17
(compiler/program.rs:350-420).  GROUP_ORDER = 4 values are the ones the reference asserts; the 8- and
16-point vectors extend the same circuit with the compiler's padding rule (unused rows: selectors 0,
permutation cells of the `None` variable chained among themselves) so the transform is exercised at every
GROUP_ORDER the compiler instantiates (4 and 8) plus the largest radix-2 domain F17 has (16)."""
import numpy as np

P = 17
G = 14  # PlutoScalarField::PRIMITIVE_ELEMENT (prime/mod.rs:87-123 finds 14 for p = 17)

# compiler/program.rs:360-377 (s_polys) and :389-419 (selector_polys): asserted coefficient vectors
REFERENCE_N4 = {
    "s1": [4, 3, 1, 15], "s2": [9, 13, 16, 14], "s3": [2, 5, 8, 12],
    "ql": [1, 0, 0, 0], "qr": [0, 0, 0, 2], "qm": [0, 0, 16, 1], "qo": [0, 1, 1, 1], "qc": [0, 8, 12, 0],
}


def padded(n: int) -> dict:
    """The same 4-gate circuit in a GROUP_ORDER = n domain: selectors zero-padded; for the permutation
    polynomials the cells keep their labels' structure K·ω^row (K = 1, 2, 3 per column,
    compiler/utils.rs `label`) — rows ≥ 4 hold the identity permutation."""
    w = pow(G, (P - 1) // n, P)
    out = {k: v + [0] * (n - 4) for k, v in REFERENCE_N4.items() if k.startswith("q")}
    for col, name in enumerate(("s1", "s2", "s3")):
        out[name] = [((col + 1) * pow(w, r, P)) % P for r in range(n)]
    return out


def srs(oracle, n: int):
    """[τ^i]G1 for i < n with the reference's τ = 2 (kzg/setup.rs:13); the first 7 equal kzg::setup()."""
    g1 = bytes([1, 0, 2, 0])
    return np.frombuffer(b"".join(oracle.point_smul(g1, pow(2, i, 17)) for i in range(n)), dtype=np.uint8).reshape(n, 4).copy()
