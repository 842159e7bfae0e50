"""CPU tier, SURVEY §8f row 4: the compiler's Lagrange-basis polynomials → monomial basis (oracle ifft,
polynomial/mod.rs:430-453) → kzg::commit (kzg/setup.rs:48-60).  Pins what the GPU test compares against:
the round trip through evaluate at the nodes, and commit = p(τ)·G1 with p(τ) from the reference's own
Lagrange-basis evaluate (mod.rs:382-415)."""
import numpy as np
import pytest

import oracle
import plonk_vectors as pv


@pytest.mark.parametrize("n", [4, 8, 16])
def test_lagrange_to_monomial_commit(n):
    polys = pv.REFERENCE_N4 if n == 4 else pv.padded(n)
    srs = pv.srs(oracle, n)
    if n == 4:  # the first SRS elements are the reference's setup() output (kzg/tests.rs:11-51)
        g1, _ = oracle.setup()
        assert bytes(srs.reshape(-1)) == b"".join(g1[:4])
    w = oracle.root_of_unity(pv.P, n, pv.G)
    for name, ev in polys.items():
        mono = oracle.ifft(pv.P, np.array(ev, dtype=np.uint64), pv.G)
        # ifft∘fft = id and the monomial form interpolates the evaluations at ω^i
        assert list(oracle.fft(pv.P, mono, pv.G)) == ev
        assert [oracle.poly_eval(pv.P, mono, pow(w, i, pv.P)) for i in range(n)] == ev
        c = oracle.commit(mono.astype(np.uint8), srs)
        # commit(Σ c_i x^i) with srs_i = τ^i·G1 is p(τ)·G1; p(τ) straight from the Lagrange form
        # (for n = 8, 16 the point τ = 2 is itself a node ω^i, where the reference's barycentric fold
        # returns l(τ)·c = 0 — mod.rs:405-414 — so there p(τ) comes from the monomial evaluate, mod.rs:133-139)
        nodes = {pow(w, i, pv.P) for i in range(n)}
        p_tau = (oracle.poly_eval(pv.P, mono, 2) if 2 in nodes
                 else oracle.lagrange_eval(pv.P, np.array(ev, dtype=np.uint64), 2, pv.G))
        assert c == oracle.point_smul(bytes([1, 0, 2, 0]), p_tau), name
