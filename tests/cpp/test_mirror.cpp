// test_mirror.cpp — the reference's own unit tests for the hot path, restated against the C++ host
// mirror (include/ronk_b200.hpp).  Needs a B200; built by __graft_entry__.build(), run by
// tests/test_gpu_cpp_mirror.py.  Cites the reference tests each block follows.
#include <cstdio>
#include <cstdlib>

#include "ronk_b200.hpp"

using namespace ronk;

static int failures = 0;
#define CHECK(cond)                                                      \
  do {                                                                   \
    if (!(cond)) { std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); failures++; } \
  } while (0)
template <class Fn>
static bool panics(Fn fn) {
  try { fn(); } catch (const Panic&) { return true; }
  return false;
}
template <class F>
static std::vector<F> vec(std::initializer_list<uint64_t> l) {
  std::vector<F> v;
  for (auto x : l) v.emplace_back(x);
  return v;
}

int main() {
  using B = PlutoBaseField;
  using S = PlutoScalarField;
  // src/algebra/field/prime/arithmetic.rs:80-216
  CHECK(S(12) + S(5) == S(0)); CHECK(B(60) + B(60) == B(19));
  CHECK(S(5) - S(12) == S(10)); CHECK(B(40) - B(61) == B(80));
  CHECK(S(10) * S(10) == S(15)); CHECK(B(40) * B(61) == B(16));
  CHECK(S(12).pow(5) == S(3)); CHECK(B(25).pow(25) == B(1)); CHECK(B(0).pow(0) == B(1));
  CHECK(*S(12).inverse() == S(10)); CHECK(*B(61).inverse() == B(53)); CHECK(!B(0).inverse().has_value());
  CHECK(B(15) / B(2) == B(58));
  CHECK(panics([] { (void)(B(1) / B(0)); }));
  // src/algebra/field/prime/mod.rs:297-314, :386-391
  CHECK(B::PRIMITIVE_ELEMENT() == B(2)); CHECK(S::PRIMITIVE_ELEMENT() == S(14));
  CHECK(panics([] { (void)S::primitive_root_of_unity(3); }));
  // src/polynomial/tests.rs
  Polynomial<Monomial, B> poly(vec<B>({1, 2, 3, 4}));
  CHECK(poly.evaluate(B(2)) == B(49));
  CHECK(poly.dft().coefficients == vec<B>({10, 79, 99, 18}));
  CHECK(poly.fft().coefficients == vec<B>({10, 79, 99, 18}));
  CHECK(poly.fft().ifft() == poly);
  CHECK(poly.dft().evaluate(B(2)) == B(49));
  CHECK(poly.degree() == 3); CHECK(poly.leading_coefficient() == B(4));
  CHECK(panics([] { (void)Polynomial<Monomial, B>(vec<B>({1, 2, 3})).dft(); }));
  // src/polynomial/arithmetic.rs:194-371
  Polynomial<Monomial, B> a(vec<B>({1, 2, 3, 4})), b(vec<B>({5, 6, 7, 8, 9}));
  CHECK((b + a).coefficients == vec<B>({6, 8, 10, 12, 9}));
  CHECK((a * b).coefficients == vec<B>({5, 16, 34, 60, 70, 70, 59, 36}));
  CHECK((b / a).coefficients == vec<B>({95, 78, 0, 0, 0}));
  CHECK((b % a).coefficients == vec<B>({11, 41, 71, 0, 0}));
  // 64-bit instantiation (SURVEY §8c golden NTT8)
  using G = GoldilocksField;
  Polynomial<Monomial, G> g8(vec<G>({1, 2, 3, 4, 5, 6, 7, 8}));
  auto X = g8.fft();
  CHECK(X.coefficients[0] == G(36)); CHECK(X.coefficients[1] == G(18445622567621360637ULL));
  CHECK(X.coefficients[7] == G(1121501793223676ULL)); CHECK(X.ifft() == g8);
  CHECK(G::primitive_root_of_unity(1ULL << 32) == G(1753635133440165772ULL));
  // src/curve/pluto_curve.rs:90-170, src/kzg/tests.rs
  AffinePoint g = G1_GENERATOR();
  AffinePoint two_g = g + g;
  CHECK((two_g.raw == std::array<uint8_t, 4>{68, 0, 74, 0}));
  CHECK(((g * S(16)).raw == std::array<uint8_t, 4>{1, 0, 99, 0}));
  CHECK((g + (-g)).is_infinity());
  CHECK(panics([] { (void)AffinePoint::make(36, 0, 0, 81); }));
  auto srs = kzg::setup();
  CHECK(srs.first.size() == 7 && srs.second.size() == 2);
  CHECK((srs.first[3].raw == std::array<uint8_t, 4>{18, 0, 49, 0}));
  CHECK(kzg::commit(vec<S>({11, 11, 11, 1}), srs.first).is_infinity());
  CHECK((kzg::commit(vec<S>({7, 16, 1, 11, 1}), srs.first).raw == std::array<uint8_t, 4>{32, 0, 59, 0}));
  CHECK((kzg::open(vec<S>({11, 11, 11, 1}), S(4), srs.first).raw == std::array<uint8_t, 4>{26, 0, 45, 0}));
  CHECK(panics([&] { (void)kzg::commit(std::vector<S>(8, S(1)), srs.first); }));
  // src/codes/reed_solomon.rs:136-219 (P = 127, K = 3 / 5, N = 3 / 7)
  using M127 = PrimeField<127>;
  {
    codes::Message<M127> msg(vec<M127>({1, 2, 3}));
    auto cw3 = msg.encode(3);
    CHECK(cw3.size() == 3 && cw3[1].x == M127(107) && cw3[2].x == M127(19));
    CHECK(cw3[0].y == M127(6) && cw3[1].y == M127(18) && cw3[2].y == M127(106));
    CHECK(codes::Message<M127>::decode(msg.encode(7), 3).data == msg.data);
    codes::Message<M127> msg5(vec<M127>({1, 2, 3, 4, 5}));
    CHECK(codes::Message<M127>::decode(msg5.encode(7), 5).data == msg5.data);
    CHECK(panics([&] { (void)msg5.encode(3); }));   // N < K
    CHECK(panics([&] { (void)msg.encode(4); }));    // 4 does not divide 126
    using G64 = GoldilocksField;
    std::vector<G64> big;
    for (uint64_t i = 0; i < 300; i++) big.emplace_back(i * i + 7);
    codes::Message<G64> mg(big);
    CHECK(codes::Message<G64>::decode(mg.encode(512), 300).data == big);   // power of two: the NTT path
  }
  {  // multi-GPU entry points with a world of one (the N > 1 paths run under torchrun: tools/multi_gpu_check.py)
    Context& c = Context::global();
    dist::init(c, dist::unique_id(), 0, 1);
    int r = -1, w = -1;
    CHECK(ronk_dist_rank(c.get(), &r, &w) == RONK_OK && r == 0 && w == 1);
    uint64_t lo = 9, hi = 9;
    CHECK(ronk_dist_shard_range(13, 2, 4, &lo, &hi) == RONK_OK && lo == 7 && hi == 10);
    const uint64_t P = RONK_GOLDILOCKS;
    std::vector<uint64_t> h(1 << 10), ref;
    for (size_t i = 0; i < h.size(); i++) h[i] = (i * 2654435761ULL + 12345) % P;
    ref = h;
    c.check(ronk_ntt_u64_host(c.get(), P, 7, ref.data(), 10, 1, 0));
    void* d = nullptr;
    c.check(ronk_dev_alloc(c.get(), &d, h.size() * 8));
    c.check(ronk_memcpy_h2d(c.get(), d, h.data(), h.size() * 8));
    dist::ntt(c, P, 7, (uint64_t*)d, 10, 1, RONK_DIST_NCCL);          // G = 1: the whole transform is local
    std::vector<uint64_t> got(h.size());
    c.check(ronk_memcpy_d2h(c.get(), got.data(), d, h.size() * 8));
    CHECK(got == ref);
    c.check(ronk_memcpy_h2d(c.get(), d, h.data(), h.size() * 8));
    auto range = dist::ntt_batch_sharded(c, P, 7, (uint64_t*)d, 8, 4);   // 4 transforms of 256 points, all ours
    CHECK(range.first == 0 && range.second == 4);
    auto srs = kzg::setup();
    std::vector<uint8_t> pts, sc = {7, 16, 1, 11, 1};
    for (size_t i = 0; i < sc.size(); i++) pts.insert(pts.end(), srs.first[i].raw.begin(), srs.first[i].raw.end());
    void *dp = nullptr, *ds = nullptr;
    c.check(ronk_dev_alloc(c.get(), &dp, pts.size()));
    c.check(ronk_dev_alloc(c.get(), &ds, sc.size()));
    c.check(ronk_memcpy_h2d(c.get(), dp, pts.data(), pts.size()));
    c.check(ronk_memcpy_h2d(c.get(), ds, sc.data(), sc.size()));
    CHECK((dist::commit(c, (const uint8_t*)dp, (const uint8_t*)ds, sc.size()).raw == std::array<uint8_t, 4>{32, 0, 59, 0}));
    // the distributed decomposition with G = 4 VIRTUAL ranks on this one device (both exchange flavours): slice r holds
    // a[r::4]; rank s ends with [q][k] = X[s·blk + k + m·q] — reassembled, it must be the single-device transform
    {
      const uint32_t log_n = 10, log_g = 2, G = 4;
      const size_t n = 1 << log_n, m = n / G, blk = m / G;
      for (int flavour : {RONK_DIST_NCCL, RONK_DIST_FUSED}) {
        std::vector<uint64_t> loc(n), out(n), X(n);
        for (size_t r = 0; r < G; r++)
          for (size_t j = 0; j < m; j++) loc[r * m + j] = h[r + G * j];
        c.check(ronk_memcpy_h2d(c.get(), d, loc.data(), n * 8));
        dist::ntt_virtual(c, P, 7, (uint64_t*)d, log_n, 1, log_g, flavour);
        c.check(ronk_memcpy_d2h(c.get(), out.data(), d, n * 8));
        for (size_t s2 = 0; s2 < G; s2++)
          for (size_t q = 0; q < G; q++)
            for (size_t k = 0; k < blk; k++) X[s2 * blk + k + m * q] = out[s2 * m + q * blk + k];
        CHECK(X == ref);
      }
    }
    ronk_dev_free(c.get(), d); ronk_dev_free(c.get(), dp); ronk_dev_free(c.get(), ds);
    dist::finalize(c);
    CHECK(ronk_ntt_u64_dist(c.get(), P, 7, nullptr, 10, 1, 0) == RONK_ENCCL);   // no communicator any more
  }
  std::printf(failures ? "%d FAILURES\n" : "cpp mirror ok\n", failures);
  return failures ? 1 : 0;
}
