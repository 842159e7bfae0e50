// msm_curve.cuh — GF(101²) and AffinePoint<PlutoExtendedCurve> arithmetic shared by the MSM kernels and the
// element-wise point kernels of msm.cu.  Host-compilable (RONK_DEV), so tests/emu can run it on the CPU tier.
// Mirrors src/algebra/field/extension/gf_101_2.rs (inverse :35-47, Mul :86-100) and src/curve/mod.rs
// (Add :178-213, Neg :225-235, is_on_curve :130-139); curve y² = x³ + 3 (src/curve/pluto_curve.rs:40-51).
#pragma once
#include "field.cuh"

namespace ronk {

constexpr u32 Q101 = 101;
constexpr u32 PT_INF = 0xFFFFFFFFu;

struct Gf { u32 c0, c1; };
struct Pt { Gf x, y; bool inf; };

RONK_DEV u32 fq_mul(u32 a, u32 b) { return (a * b) % Q101; }
RONK_DEV u32 fq_add(u32 a, u32 b) { u32 s = a + b; return s >= Q101 ? s - Q101 : s; }
RONK_DEV u32 fq_sub(u32 a, u32 b) { return a >= b ? a - b : a + Q101 - b; }
RONK_DEV u32 fq_neg(u32 a) { return a ? Q101 - a : 0; }
// a^99 = a^-1 (Fermat; prime/mod.rs:62-72).  99 = 0b1100011.
RONK_DEV u32 fq_inv(u32 a) {
  const u32 a2 = fq_mul(a, a), a3 = fq_mul(a2, a);
  const u32 a6 = fq_mul(a3, a3), a12 = fq_mul(a6, a6), a24 = fq_mul(a12, a12);
  const u32 a48 = fq_mul(a24, a24), a96 = fq_mul(a48, a48);
  return fq_mul(a96, a3);
}

RONK_DEV Gf gf_add(Gf a, Gf b) { return {fq_add(a.c0, b.c0), fq_add(a.c1, b.c1)}; }
RONK_DEV Gf gf_sub(Gf a, Gf b) { return {fq_sub(a.c0, b.c0), fq_sub(a.c1, b.c1)}; }
RONK_DEV Gf gf_neg(Gf a) { return {fq_neg(a.c0), fq_neg(a.c1)}; }
RONK_DEV bool gf_eq(Gf a, Gf b) { return a.c0 == b.c0 && a.c1 == b.c1; }
// (a0 + a1 t)(b0 + b1 t) mod (t² + 2) = (a0b0 - 2a1b1) + (a0b1 + a1b0) t
RONK_DEV Gf gf_mul(Gf a, Gf b) {
  return {(a.c0 * b.c0 + 99u * (a.c1 * b.c1 % Q101)) % Q101, (a.c0 * b.c1 + a.c1 * b.c0) % Q101};
}
// conj / norm, norm = a0² + 2a1²  (gf_101_2.rs:35-47); caller guarantees a != 0
RONK_DEV Gf gf_inv(Gf a) {
  const u32 s = fq_inv((a.c0 * a.c0 + 2u * a.c1 * a.c1) % Q101);
  return {fq_mul(a.c0, s), fq_mul(fq_neg(a.c1), s)};
}

// norm inverse from a table tab[a] = a^-1 mod 101 (a in 1..100)
RONK_DEV Gf gf_inv_tab(Gf a, const uint8_t* tab) {
  const u32 s = tab[(a.c0 * a.c0 + 2u * a.c1 * a.c1) % Q101];
  return {fq_mul(a.c0, s), fq_mul(fq_neg(a.c1), s)};
}

RONK_DEV Pt pt_unpack(u32 w) {
  Pt p;
  p.inf = (w == PT_INF);
  p.x = {w & 0xFF, (w >> 8) & 0xFF};
  p.y = {(w >> 16) & 0xFF, w >> 24};
  return p;
}
RONK_DEV u32 pt_pack(const Pt& p) {
  return p.inf ? PT_INF : (p.x.c0 | (p.x.c1 << 8) | (p.y.c0 << 16) | (p.y.c1 << 24));
}
// Well-formed (canonical coordinates) and on y² = x³ + 3  (curve/mod.rs:130-139).
RONK_DEV bool pt_valid(u32 w) {
  if (w == PT_INF) return true;
  const Pt p = pt_unpack(w);
  if (p.x.c0 >= Q101 || p.x.c1 >= Q101 || p.y.c0 >= Q101 || p.y.c1 >= Q101) return false;
  const Gf lhs = gf_mul(p.y, p.y);
  const Gf rhs = gf_add(gf_mul(gf_mul(p.x, p.x), p.x), Gf{3, 0});
  return gf_eq(lhs, rhs);
}
// AffinePoint + AffinePoint  (curve/mod.rs:178-213), same case order as the reference.
RONK_DEV Pt pt_add(const Pt& a, const Pt& b) {
  if (a.inf) return b;
  if (b.inf) return a;
  const bool same_x = gf_eq(a.x, b.x);
  if (same_x && gf_eq(a.y, gf_neg(b.y))) { Pt r; r.inf = true; r.x = {0, 0}; r.y = {0, 0}; return r; }
  Gf num, den;
  if (same_x && gf_eq(a.y, b.y)) {  // tangent: 3x² / 2y   (a = 0)
    num = gf_mul(Gf{3, 0}, gf_mul(a.x, a.x));
    den = gf_add(a.y, a.y);
  } else {                          // chord: (y2 - y1) / (x2 - x1)
    num = gf_sub(b.y, a.y);
    den = gf_sub(b.x, a.x);
  }
  const Gf lam = gf_mul(num, gf_inv(den));
  Pt r;
  r.inf = false;
  r.x = gf_sub(gf_sub(gf_mul(lam, lam), a.x), b.x);
  r.y = gf_sub(gf_mul(lam, gf_sub(a.x, r.x)), a.y);
  return r;
}
RONK_DEV u32 pt_add_w(u32 a, u32 b) { return pt_pack(pt_add(pt_unpack(a), pt_unpack(b))); }

// Same addition law with the table inverse (MSM kernels).
RONK_DEV u32 pt_add_t(u32 wa, u32 wb, const uint8_t* tab) {
  if (wa == PT_INF) return wb;
  if (wb == PT_INF) return wa;
  const Pt a = pt_unpack(wa), b = pt_unpack(wb);
  const bool same_x = gf_eq(a.x, b.x);
  if (same_x && gf_eq(a.y, gf_neg(b.y))) return PT_INF;
  Gf num, den;
  if (same_x && gf_eq(a.y, b.y)) {
    num = gf_mul(Gf{3, 0}, gf_mul(a.x, a.x));
    den = gf_add(a.y, a.y);
  } else {
    num = gf_sub(b.y, a.y);
    den = gf_sub(b.x, a.x);
  }
  const Gf lam = gf_mul(num, gf_inv_tab(den, tab));
  Pt r;
  r.inf = false;
  r.x = gf_sub(gf_sub(gf_mul(lam, lam), a.x), b.x);
  r.y = gf_sub(gf_mul(lam, gf_sub(a.x, r.x)), a.y);
  return pt_pack(r);
}
RONK_DEV void build_inv_table(uint8_t* tab, u32 tid, u32 nthr) {
  for (u32 a = tid; a < Q101; a += nthr) tab[a] = (uint8_t)(a ? fq_inv(a) : 0);
}

}  // namespace ronk
