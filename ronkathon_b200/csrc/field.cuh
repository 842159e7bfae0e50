// field.cuh — device-side PrimeField arithmetic (sm_100a, integer pipes only).
//
// Mirrors src/algebra/field/prime/arithmetic.rs (Add :6, Sub :22-27, Mul :37, Neg :64) on
// canonical residues.  Two field policies share every kernel in this library:
//   GoldilocksField  p = 2^64 - 2^32 + 1 baked in: reduction by 2^64 ≡ 2^32-1, 2^96 ≡ -1 and
//                    multiplication by the 16th roots of unity as shifts (they are powers of two);
//   MontField        any odd modulus < 2^64 given at run time (Montgomery REDC) — the path the
//                    reference's own moduli p = 101 / 17 / 127 run through as the bit-exact cross-check.
// "Twiddle form" is what mul_tw's second operand must be in: plain residues for Goldilocks,
// Montgomery form (w·2^64 mod p) for MontField, so data never leaves normal form.
#pragma once
#include <cstdint>

namespace ronk {

typedef uint64_t u64;
typedef unsigned int u32;

// Every arithmetic routine is __host__ __device__ so that tests/emu can compile the very same
// source for the CPU and check the kernel logic without a GPU (test infrastructure only — the
// product library never runs these on the host).
#if defined(__CUDACC__)
#define RONK_DEV __host__ __device__ __forceinline__
#define RONK_HD __host__ __device__ __forceinline__
#else
#define RONK_DEV inline
#define RONK_HD inline
#endif

RONK_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(a, b);
#else
  return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

constexpr u64 GL_P = 0xFFFFFFFF00000001ULL;
constexpr u64 GL_EPS = 0xFFFFFFFFULL;  // 2^64 mod p

// ------------------------------------------------------------------------------------------
// Goldilocks
// ------------------------------------------------------------------------------------------
struct GoldilocksField {
  RONK_HD u64 modulus() const { return GL_P; }

  // (a + b) mod p, canonical in → canonical out.
  RONK_DEV u64 add(u64 a, u64 b) const {
    u64 s = a + b;
    // s wrapped (carry) or s >= p  ⇔  a + b >= p  (a, b < p): subtract p == add EPS mod 2^64
    u64 t = s + GL_EPS;
    return (s < a || t < s) ? t : s;
  }
  // (a - b) mod p: borrow → add p back (== subtract EPS mod 2^64)
  RONK_DEV u64 sub(u64 a, u64 b) const {
    u64 d = a - b;
    return (a < b) ? d - GL_EPS : d;
  }
  RONK_DEV u64 neg(u64 a) const { return a ? GL_P - a : 0; }

  // 128-bit (hi:lo) → canonical residue.  x = lo + hi_lo·2^64 + hi_hi·2^96 ≡ lo + hi_lo·EPS - hi_hi.
  RONK_DEV u64 reduce128(u64 lo, u64 hi) const {
    u64 hh = hi >> 32, hl = hi & 0xFFFFFFFFULL;
    u64 t0 = lo - hh;
    if (lo < hh) t0 -= GL_EPS;          // wrapped: +p
    u64 t1 = (hl << 32) - hl;           // hl·EPS < 2^64
    u64 r = t0 + t1;
    if (r < t1) r += GL_EPS;            // wrapped: 2^64 ≡ EPS, cannot wrap twice
    return (r >= GL_P) ? r - GL_P : r;
  }
  RONK_DEV u64 mul(u64 a, u64 b) const { return reduce128(a * b, mulhi64(a, b)); }
  RONK_DEV u64 mul_tw(u64 a, u64 w) const { return mul(a, w); }
  RONK_DEV u64 to_tw(u64 w) const { return w; }

  // a · 2^S mod p for a compile-time S in [0, 192)  (2 has order 192: 2^96 ≡ -1).
  template <int S>
  RONK_DEV u64 mul_pow2(u64 a) const {
    static_assert(S >= 0 && S < 192, "shift out of range");
    if constexpr (S == 0) {
      return a;
    } else if constexpr (S >= 96) {
      return neg(mul_pow2<S - 96>(a));
    } else if constexpr (S < 64) {
      return reduce128(a << S, a >> (64 - S));
    } else {
      return mul_pow2<S - 32>(mul_pow2<32>(a));
    }
  }
  // a · ω16^E (forward) or a · ω16^-E (INV), E in [0,8).  ω16 = g^((p-1)/16) = 2^156 for g = 7.
  template <int E, bool INV>
  RONK_DEV u64 w16(u64 a) const {
    constexpr int fwd = (156 * E) % 192;
    constexpr int sh = INV ? (192 - fwd) % 192 : fwd;
    return mul_pow2<sh>(a);
  }
};

// ------------------------------------------------------------------------------------------
// Generic odd modulus < 2^64, Montgomery multiplication with R = 2^64.
// ------------------------------------------------------------------------------------------
struct MontField {
  u64 p;        // modulus (odd)
  u64 pinv;     // p^-1 mod 2^64
  u64 r2;       // 2^128 mod p
  u64 w16t[8];  // ω16^e (direction already applied) in Montgomery form; unused entries = R mod p

  RONK_HD u64 modulus() const { return p; }

  RONK_DEV u64 add(u64 a, u64 b) const {
    u64 s = a + b;
    return (s < a || s >= p) ? s - p : s;
  }
  RONK_DEV u64 sub(u64 a, u64 b) const {
    u64 d = a - b;
    return (a < b) ? d + p : d;
  }
  RONK_DEV u64 neg(u64 a) const { return a ? p - a : 0; }
  // REDC(a·b) = a·b·2^-64 mod p, subtractive form (no 129-bit intermediate for p > 2^63).
  RONK_DEV u64 redc_mul(u64 a, u64 b) const {
    u64 lo = a * b, hi = mulhi64(a, b);
    u64 m = lo * pinv;
    u64 mp = mulhi64(m, p);
    u64 t = hi - mp;
    return (hi < mp) ? t + p : t;
  }
  RONK_DEV u64 mul_tw(u64 a, u64 w_mont) const { return redc_mul(a, w_mont); }
  RONK_DEV u64 mul(u64 a, u64 b) const { return redc_mul(redc_mul(a, b), r2); }
  RONK_DEV u64 to_tw(u64 w) const { return redc_mul(w, r2); }
  template <int E, bool INV>
  RONK_DEV u64 w16(u64 a) const {
    if constexpr (E == 0) return a;
    return redc_mul(a, w16t[E]);
  }
};

// a^e by square-and-multiply (value of Field::pow, prime/mod.rs:74-84).
template <class F>
RONK_DEV u64 field_pow(const F& f, u64 a, u64 e) {
  u64 r = 1 % f.modulus();
  u64 base = a;
  while (e) {
    if (e & 1) r = f.mul(r, base);
    base = f.mul(base, base);
    e >>= 1;
  }
  return r;
}

// ------------------------------------------------------------------------------------------
// Host-side scalar helpers (plan building only: roots of unity, Montgomery constants).
// ------------------------------------------------------------------------------------------
inline u64 h_mulmod(u64 a, u64 b, u64 p) { return (u64)(((unsigned __int128)a * b) % p); }
inline u64 h_powmod(u64 a, u64 e, u64 p) {
  u64 r = 1 % p;
  a %= p;
  while (e) {
    if (e & 1) r = h_mulmod(r, a, p);
    a = h_mulmod(a, a, p);
    e >>= 1;
  }
  return r;
}
inline u64 h_inv64(u64 p) {  // p^-1 mod 2^64 (p odd), Newton
  u64 x = p;
  for (int i = 0; i < 6; i++) x *= 2 - p * x;
  return x;
}

}  // namespace ronk
