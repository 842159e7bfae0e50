// ntt12_kernel.cuh — the tile kernel specialised for 4096-point transforms per tile (log_m = 12: both passes
// of the 2^24-point transform the headline metric is quoted on, and any pass whose transform length is 4096).
//
// Same algorithm, tile maps, round schedule and twiddle tables as ntt_tile_kernel (ntt_kernel.cuh); what
// changes is that every shape parameter is a compile-time constant and shared memory uses an ADDITIVE padded
// layout instead of the XOR swizzle, so that no phase spends integer-pipe instructions on addresses:
//
//   word(e) = e + (e >> (LC+4) << LC) + (e >> (2·LC+8) << LC)        (LC = log2 of the tile's column count)
//
// i.e. 2^LC pad words after every 2^(LC+4) elements and again after every 2^(2·LC+8): ≈ +6.4 % shared memory.
// word() is additive over disjoint bit fields of e, and every access pattern of the kernel enumerates
// e = e(thread) | e(step) with disjoint fields (the tile index of a loaded / stored element is a bit permutation
// of the thread-and-step counter), hence word(e) = word(e(thread)) + word(e(step)) where the second term is a
// compile-time constant: each of the 64–96 shared-memory accesses a thread makes per tile is `LDS/STS [R + imm]`.
// The pads make every pattern conflict-free (64-bit accesses are served per half-warp against 16 eight-byte
// banks = word mod 16):
//   tile load, round 0, round 1 : the 16 lanes differ in e bits [0,4)                     → 16 distinct banks
//   round 2 (window [LC, LC+4)) : lanes differ in bits [0,LC) ∪ [LC+4, 8)       — first pad term
//   un-bit-reversing store      : lanes differ in bits [0,LC) ∪ [2·LC+8, LC+12) — second pad term
// (tools/smem_layout_audit.py enumerates all of them; tests/emu runs these very functions on the CPU tier.)
// Round 1 measured 40–47 % of the executed instructions of both passes as address arithmetic (swizzle XORs,
// bit-reversal, Gray-code walks, per-element index math) on the already saturated ALU pipe
// (profiles/r02c_ntt_metrics.txt, per-line counts in DESIGN.md §3); this kernel has none of it.
#pragma once
#include "ntt_kernel.cuh"

namespace ronk {

template <int LC>
struct N12 {
  static constexpr int LM = 12, TL = LM + LC, KK = TL - 5;
  static constexpr u32 T = 1u << TL, NTHR = T / 32, C = 1u << LC;
  static constexpr u32 word(u32 e) { return e + ((e >> (LC + 4)) << LC) + ((e >> (2 * LC + 8)) << LC); }
  static constexpr u32 TILE_WORDS = (word(T - 1) + 2u) & ~1u;  // even: the twiddle table behind it is a 16-byte TMA target
  static constexpr u32 TW_OFF1 = TW_ROW << 8;                  // ntt_tw2d_layout(12): round 0 at 0, round 1 behind it
  static constexpr u32 TW_WORDS = (TW_OFF1 + (TW_ROW << 4) + 1u) & ~1u;
};
RONK_HD constexpr u32 bitrev12c(u32 v) {
  u32 r = 0;
  for (int i = 0; i < 12; i++) r |= ((v >> i) & 1u) << (11 - i);
  return r;
}

// ---------------- load: HBM → shared ----------------
template <int MODE, int LC>
RONK_DEV void n12_load(u64* smem, const NttTileArgs& A, u32 tile, u32 tid) {
  using L = N12<LC>;
  const u32 b = tile / A.tiles_per_batch, sub = tile - b * A.tiles_per_batch;
  const u64* src;
  u64 stride;  // global words between element tid + j·NTHR and tid + (j+1)·NTHR
  if (MODE == MODE_PASS1) {
    src = A.src + ((u64)b << A.log_n) + ((u64)sub << LC) + ((u64)(tid >> LC) << A.log_n2) + (tid & (L::C - 1u));
    stride = (u64)(L::NTHR >> LC) << A.log_n2;
  } else {
    src = A.src + ((u64)b << A.log_n) + ((u64)sub << L::TL) + tid;
    stride = L::NTHR;
  }
  u64* const s = smem + L::word(tid);
  constexpr int LB = (MODE == MODE_PASS1) ? 8 : 16;  // loads in flight per thread (cf. RONK_LD_BATCH)
#pragma unroll
  for (int j0 = 0; j0 < 32; j0 += LB) {
    u64 v[LB];
#pragma unroll
    for (int i = 0; i < LB; i++) v[i] = src[(u64)(j0 + i) * stride];
#pragma unroll
    for (int i = 0; i < LB; i++) s[L::word((u32)(j0 + i) << L::KK)] = v[i];
  }
}

// ---------------- one radix-16 round (R = 0, 1, 2: window of the transform index from the top down) ----------------
template <class F, bool INV, int LC, int R>
RONK_DEV void n12_round(const F& f, u64* smem, const u64* tw, u32 tid) {
  using L = N12<LC>;
  constexpr int WB = LC + 8 - 4 * R;   // lowest tile-index bit of the window
  constexpr int LCUR = 12 - 4 * R;     // log2 of the sub-transform length
#pragma unroll
  for (int g = 0; g < 2; g++) {
    const u32 t = tid + (u32)g * L::NTHR;
    const u32 e0 = ((t >> WB) << (WB + 4)) | (t & ((1u << WB) - 1u));
    u64* const s = smem + L::word(e0);
    u64 x[16];
#pragma unroll
    for (int q = 0; q < 16; q++) x[q] = s[L::word((u32)q << WB)];
    radix_network<4, INV>(f, x);
    if constexpr (LCUR > 4) {
      const u32 i2 = (e0 >> LC) & ((1u << (LCUR - 4)) - 1u);
      const u64* row = tw + (R == 0 ? 0u : L::TW_OFF1) + i2 * TW_ROW;
#pragma unroll
      for (int j = 1; j < 16; j++) {
        const int k1 = ((j & 1) << 3) | ((j & 2) << 1) | ((j & 4) >> 1) | ((j & 8) >> 3);
        x[j] = f.mul_tw(x[j], row[k1]);
      }
    }
#pragma unroll
    for (int q = 0; q < 16; q++) s[L::word((u32)q << WB)] = x[q];
  }
}

// ---------------- store: shared → HBM, un-bit-reversing on the fly ----------------
// PASS2: g = tid + j·NTHR enumerates the tile in output order: k1_in = g mod C, k2 = g / C; the value sits at
//        tile index (bitrev12(k2) << LC) | k1_in and goes to X[k1 + N1·k2].
template <class F, int LC, bool FMUL>
RONK_DEV void n12_store_pass2(const F& f, const u64* smem, const NttTileArgs& A, u32 tile, u32 tid) {
  using L = N12<LC>;
  const u32 b = tile / A.tiles_per_batch, sub = tile - b * A.tiles_per_batch;
  const u32 k1_in = tid & (L::C - 1u), k2_t = tid >> LC;
  const u32 e_t = (bitrev(k2_t, 12) << LC) | k1_in;
  const u64* const s = smem + L::word(e_t);
  const u64 off_t = ((u64)b << A.log_n) + ((u64)sub << LC) + k1_in + ((u64)k2_t << A.log_n1);
  const u64 stride = (u64)(L::NTHR >> LC) << A.log_n1;
  u64* const dst = A.dst + off_t;
  if (FMUL) {
    constexpr int SB = 8;
#pragma unroll
    for (int j0 = 0; j0 < 32; j0 += SB) {
      u64 m[SB];
#pragma unroll
      for (int i = 0; i < SB; i++) m[i] = A.mul_src[(off_t + (u64)(j0 + i) * stride) & A.mul_mask];
#pragma unroll
      for (int i = 0; i < SB; i++) {
        const u32 ej = bitrev12c(((u32)(j0 + i) << L::KK) >> LC) << LC;  // compile-time after unrolling
        dst[(u64)(j0 + i) * stride] = f.mul(s[L::word(ej)], m[i]);
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; j++) {
      const u32 ej = bitrev12c(((u32)j << L::KK) >> LC) << LC;
      dst[(u64)j * stride] = s[L::word(ej)];
    }
  }
}

// PASS1: g = tid + j·NTHR in workspace order W[k1 / C2][j2][k1 % C2]: rem = g mod (C·C2), k1 = (g / (C·C2))·C2 +
//        rem mod C2, column c = rem / C2; the value sits at tile index (bitrev12(k1) << LC) | c and is multiplied
//        by the inter-pass twiddle ω_n^(±j2·k1) — from the n-word table laid out like the workspace (one coalesced
//        load at the store's own offset) or, without the table, stepped along the thread's k1 progression.
template <class F, bool INV, int LC, int LC2>
RONK_DEV void n12_store_pass1(const F& f, const u64* smem, const NttTileArgs& A, u32 tile, u32 tid) {
  using L = N12<LC>;
  constexpr int CL = LC + LC2;
  static_assert(L::KK >= CL, "a thread's step must not touch the (column, k1 % C2) bits");
  const u32 b = tile / A.tiles_per_batch, sub = tile - b * A.tiles_per_batch;
  const u32 rem = tid & ((1u << CL) - 1u);
  const u32 c = rem >> LC2, k1_in = rem & ((1u << LC2) - 1u);
  const u32 k1_t = ((tid >> CL) << LC2) | k1_in;
  const u32 e_t = (bitrev(k1_t, 12) << LC) | c;
  const u64* const s = smem + L::word(e_t);
  const u32 j2 = (sub << LC) | c;
  const u64 base = (u64)b << A.log_n;
  const u64 off_t = ((u64)(tid >> CL) << (A.log_n2 + LC2)) + ((u64)j2 << LC2) + k1_in;
  const u64 stride = (u64)(L::NTHR >> CL) << (A.log_n2 + LC2);
  u64* const dst = A.dst + base + off_t;
  if (A.tw_full) {
    const u64* const twp = A.tw_full + off_t;
    constexpr int TB = 8;
#pragma unroll
    for (int j0 = 0; j0 < 32; j0 += TB) {
      u64 w[TB];
#pragma unroll
      for (int i = 0; i < TB; i++) w[i] = ld_tw(twp + (u64)(j0 + i) * stride);
#pragma unroll
      for (int i = 0; i < TB; i++) {
        // step j adds j·NTHR to g: only the k1 block index moves, by j·(NTHR >> CL) — bits [LC2, 12) of k1
        const u32 ej = bitrev12c((((u32)(j0 + i) << L::KK) >> CL) << LC2) << LC;
        dst[(u64)(j0 + i) * stride] = f.mul_tw(s[L::word(ej)], w[i]);
      }
    }
    return;
  }
  const u32 nmask = (A.log_n >= 32) ? 0xFFFFFFFFu : ((1u << A.log_n) - 1u);
  const u32 lomask = (1u << A.log_lo) - 1u;
  const u32 dk1 = (L::NTHR >> CL) << LC2;
  u32 ex0 = j2 * k1_t, exd = j2 * dk1;
  if (INV) { ex0 = (0u - ex0) & nmask; exd = (0u - exd) & nmask; }
  u64 w = f.mul_tw(ld_tw(A.tw_lo + (ex0 & lomask)), ld_tw(A.tw_hi + (ex0 >> A.log_lo)));
  const u64 rho = f.mul_tw(ld_tw(A.tw_lo + (exd & lomask)), ld_tw(A.tw_hi_plain + (exd >> A.log_lo)));
#pragma unroll
  for (int j = 0; j < 32; j++) {
    const u32 ej = bitrev12c((((u32)j << L::KK) >> CL) << LC2) << LC;
    dst[(u64)j * stride] = f.mul_tw(s[L::word(ej)], w);
    w = f.mul_tw(w, rho);
  }
}

#if defined(__CUDACC__)
// Shared memory: [ tile: TILE_WORDS·8 B | twiddles: TW_WORDS·8 B | mbarrier: 8 B ]
template <class F, int MODE, bool INV, int LC, int LC2, bool FMUL>
__global__ void __launch_bounds__(N12<LC>::NTHR, N12<LC>::NTHR >= 512 ? 1 : 2)
    ntt12_kernel(const F f, const NttTileArgs A) {
  using L = N12<LC>;
  extern __shared__ __align__(128) u64 smem[];
  const u32 tid = threadIdx.x, tile = blockIdx.x;
  u64* tw = smem + L::TILE_WORDS;
  u64* bar = tw + L::TW_WORDS;
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar, L::TW_WORDS * 8u);
    tma_bulk_g2s(tw, A.tw_tile, L::TW_WORDS * 8u, bar);  // lands while the tile itself is being loaded
  }
  n12_load<MODE, LC>(smem, A, tile, tid);
  if (MODE == MODE_PASS1 && A.prefetch_dist && tile + A.prefetch_dist < gridDim.x)
    ntt_prefetch_pass1(A, tile + A.prefetch_dist, tid, L::NTHR);
  __syncthreads();
  mbar_wait(bar, 0);
  n12_round<F, INV, LC, 0>(f, smem, tw, tid);
  __syncthreads();
  n12_round<F, INV, LC, 1>(f, smem, tw, tid);
  __syncthreads();
  n12_round<F, INV, LC, 2>(f, smem, tw, tid);
  __syncthreads();
  if (MODE == MODE_PASS1) n12_store_pass1<F, INV, LC, LC2>(f, smem, A, tile, tid);
  else n12_store_pass2<F, LC, FMUL>(f, smem, A, tile, tid);
}
#endif  // __CUDACC__

// host: can this launch take the specialised kernel?
inline bool ntt12_applicable(const NttTileArgs& A, int mode) {
  if (A.log_m != 12 || A.src_len != NTT_UNBOUNDED || A.dst_len != NTT_UNBOUNDED) return false;
  if (A.tw_words != N12<1>::TW_WORDS) return false;
  if (mode == MODE_PASS1) return A.log_c == 2 && (A.log_c2 == 1 || A.log_c2 == 2);
  if (mode == MODE_PASS2) return A.log_c == 1 || A.log_c == 2;
  return false;
}

}  // namespace ronk
