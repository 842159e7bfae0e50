// ntt12_kernel.cuh — the tile kernel specialised for 4096-point transforms per tile (log_m = 12: both passes
// of the 2^24-point transform the headline metric is quoted on, and any pass whose transform length is 4096).
//
// Same algorithm, tile maps, round schedule and twiddle tables as ntt_tile_kernel (ntt_kernel.cuh).  What changes:
//  * every shape parameter is a compile-time constant;
//  * the tile is held as PAIRS of field elements (the two lowest columns of the tile, 16 bytes): a thread's
//    32 elements per round are 16 pairs, so every shared-memory and most global accesses are 128-bit
//    (LDS.128 / STS.128 / LDG.128 / STG.128) — half the load/store instructions of the 64-bit kernel, and the
//    general twiddle of a pair is shared (the two halves differ only in the column, not in the transform index);
//  * shared memory uses an ADDITIVE padded layout instead of the XOR swizzle.  In units of pairs
//        word(e) = e + (e >> (LP+4) << LP) + (e >> S2 << LP)      LP = log2(columns) - 1,
//        S2 = 2·LP + 9 (pass 2) or 2·LP + 8 (pass 1)
//    i.e. 2^LP pad slots after every 2^(LP+4) pairs and after every 2^S2 (≈ +6.4 % shared memory).  word() is
//    additive over disjoint bit fields of e, and every access pattern enumerates e = e(thread) | e(step) with
//    disjoint fields (the tile index of a loaded / stored element is a bit permutation of the thread-and-step
//    counter), so word(e) = word(e(thread)) + constant: every access is `[R + imm]`, no address arithmetic.
//    The pads make all patterns conflict-free (128-bit accesses: 8 lanes against eight 16-byte banks =
//    word mod 8; the 64-bit reads of the pass-1 store: 16 lanes against sixteen 8-byte banks):
//      tile load, rounds 0 and 1 : lanes differ in e bits [0,3)
//      round 2 (window [LP, LP+4)): lanes differ in bits [0,LP) ∪ [LP+4, 7)          — first pad term
//      un-bit-reversing stores   : lanes differ in bits [0,LP) ∪ the top 3-LP (pass 2) / 2 (pass 1) bits — second
//    (tests/emu runs these very functions on the CPU tier and audits the banks: emu_layout12_worst_conflict).
// Round 1 spent 40–47 % of the executed instructions of both passes on address arithmetic (swizzle XORs,
// bit reversal, Gray-code walks, per-element index math) on the saturated ALU pipe (profiles/r02c_ntt_*); the
// first version of this kernel removed that but then stalled on the load/store queues (mio_throttle 1.7, lg_throttle
// 0.6 per issued instruction, profiles/r02d_ntt12_metrics.txt) — hence the 128-bit accesses.
#pragma once
#include "ntt_kernel.cuh"

namespace ronk {

struct alignas(16) u64x2 {
  u64 a, b;
};

template <int LC, int MODE>
struct N12 {
  static_assert(LC >= 1 && LC <= 2, "the pair is the lowest column bit");
  static constexpr int LP = LC - 1;        // log2 of the pairs per tile row
  static constexpr int TLP = 12 + LP;      // log2 of the pairs per tile
  static constexpr int KK = TLP - 4;       // log2 of the thread count: 16 pairs per thread
  static constexpr int S2 = (MODE == MODE_PASS2) ? 2 * LP + 9 : 2 * LP + 8;
  static constexpr u32 PAIRS = 1u << TLP, NTHR = 1u << KK, C = 1u << LC;
  static constexpr u32 word(u32 e) { return e + ((e >> (LP + 4)) << LP) + ((e >> S2) << LP); }
  static constexpr u32 TILE_SLOTS = word(PAIRS - 1) + 1u;      // 16-byte slots
  static constexpr u32 TW_OFF1 = TW_ROW << 8;                  // ntt_tw2d_layout(12): round 0 at 0, round 1 behind it
  static constexpr u32 TW_WORDS = (TW_OFF1 + (TW_ROW << 4) + 1u) & ~1u;
};
RONK_HD constexpr u32 bitrev12c(u32 v) {
  u32 r = 0;
  for (int i = 0; i < 12; i++) r |= ((v >> i) & 1u) << (11 - i);
  return r;
}
RONK_HD constexpr u32 br4(int j) { return (u32)(((j & 1) << 3) | ((j & 2) << 1) | ((j & 4) >> 1) | ((j & 8) >> 3)); }
RONK_DEV u64x2 ld_pair(const u64* p) { return *reinterpret_cast<const u64x2*>(p); }
RONK_DEV void st_pair(u64* p, u64x2 v) { *reinterpret_cast<u64x2*>(p) = v; }

// ---------------- round 0, fed straight from HBM ----------------
// A thread's round-0 group is 16 pairs that differ only in the top four bits of the transform index: it loads
// exactly those from global memory into registers (for every q the warp's 32 loads are contiguous — 512 B in pass 2,
// 16 rows × 32 B in pass 1), transforms them and only then writes the tile: no separate load phase, one
// shared-memory round trip and one CTA barrier less per tile.
template <class F, int MODE, bool INV, int LC>
RONK_DEV void n12_round0_load(const F& f, u64x2* smem, const u64* tw, const NttTileArgs& A, u32 tile, u32 tid,
                              u64* tw_barrier = nullptr) {
  using L = N12<LC, MODE>;
  constexpr int WB = L::LP + 8;
  const u32 b = tile / A.tiles_per_batch, sub = tile - b * A.tiles_per_batch;
  const u64* src;
  u64 stride;  // global words between the group's pairs q and q + 1
  if (MODE == MODE_PASS1) {  // pair = (row j1, column pair c'): tile index (j1 << LP) | c'
    src = A.src + ((u64)b << A.log_n) + ((u64)sub << LC) + ((u64)(tid >> L::LP) << A.log_n2) + 2u * (tid & ((1u << L::LP) - 1u));
    stride = (u64)256 << A.log_n2;
  } else {
    src = A.src + ((u64)b << A.log_n) + ((u64)sub << (L::TLP + 1)) + 2u * tid;
    stride = 2u << WB;
  }
  u64 xa[16], xb[16];
#pragma unroll
  for (int q = 0; q < 16; q++) {
    const u64x2 v = ld_pair(src + (u64)q * stride);
    xa[q] = v.a;
    xb[q] = v.b;
  }
  radix_network<4, INV>(f, xa);
  radix_network<4, INV>(f, xb);
#if defined(__CUDA_ARCH__)
  if (tw_barrier) mbar_wait(tw_barrier, 0);  // the twiddle table (TMA bulk copy) landed while the tile was being fetched
#endif
  const u32 i2 = (tid >> L::LP) & 255u;
  const u64* row = tw + i2 * TW_ROW;
#pragma unroll
  for (int j = 1; j < 16; j++) {
    const int k1 = ((j & 1) << 3) | ((j & 2) << 1) | ((j & 4) >> 1) | ((j & 8) >> 3);
    const u64 w = row[k1];
    xa[j] = f.mul_tw(xa[j], w);
    xb[j] = f.mul_tw(xb[j], w);
  }
  u64x2* const s = smem + L::word(tid);
#pragma unroll
  for (int q = 0; q < 16; q++) s[L::word((u32)q << WB)] = u64x2{xa[q], xb[q]};
}

// ---------------- one radix-16 round (R = 0, 1, 2: window of the transform index from the top down) ----------------
template <class F, int MODE, bool INV, int LC, int R>
RONK_DEV void n12_round(const F& f, u64x2* smem, const u64* tw, u32 tid) {
  using L = N12<LC, MODE>;
  constexpr int WB = L::LP + 8 - 4 * R;  // lowest pair-index bit of the window
  constexpr int LCUR = 12 - 4 * R;       // log2 of the sub-transform length
  const u32 e0 = ((tid >> WB) << (WB + 4)) | (tid & ((1u << WB) - 1u));
  u64x2* const s = smem + L::word(e0);
  u64 xa[16], xb[16];
#pragma unroll
  for (int q = 0; q < 16; q++) {
    const u64x2 v = s[L::word((u32)q << WB)];
    xa[q] = v.a;
    xb[q] = v.b;
  }
  radix_network<4, INV>(f, xa);
  radix_network<4, INV>(f, xb);
  if constexpr (LCUR > 4) {
    const u32 i2 = (e0 >> L::LP) & ((1u << (LCUR - 4)) - 1u);
    const u64* row = tw + (R == 0 ? 0u : L::TW_OFF1) + i2 * TW_ROW;
#pragma unroll
    for (int j = 1; j < 16; j++) {
      const int k1 = ((j & 1) << 3) | ((j & 2) << 1) | ((j & 4) >> 1) | ((j & 8) >> 3);
      const u64 w = row[k1];  // the two halves of a pair sit in the same row of the transform: one twiddle
      xa[j] = f.mul_tw(xa[j], w);
      xb[j] = f.mul_tw(xb[j], w);
    }
  }
#pragma unroll
  for (int q = 0; q < 16; q++) s[L::word((u32)q << WB)] = u64x2{xa[q], xb[q]};
}

// ---------------- stores ----------------
// PASS2, round 2 with the store fused: after the last round a thread holds the 16 pairs (i_hi, i_mid fixed, i_lo = q)
// of its group; pair q goes to X[k1 + N1·k2] with k2 = bitrev12(i) = bitrev4(q)·256 + bitrev4(i_mid)·16 + bitrev4(i_hi).
// Every store of pass 2 is an isolated 16-byte segment anyway (two adjacent k1, rows N1 words apart), so going
// through shared memory once more to un-bit-reverse bought nothing: the registers are stored directly.
template <class F, bool INV, int LC, bool FMUL>
RONK_DEV void n12_round2_store_pass2(const F& f, const u64x2* smem, const NttTileArgs& A, u32 tile, u32 tid) {
  using L = N12<LC, MODE_PASS2>;
  constexpr int WB = L::LP;
  const u32 b = tile / A.tiles_per_batch, sub = tile - b * A.tiles_per_batch;
  const u32 e0 = ((tid >> WB) << (WB + 4)) | (tid & ((1u << WB) - 1u));
  const u64x2* const s = smem + L::word(e0);
  u64 xa[16], xb[16];
#pragma unroll
  for (int q = 0; q < 16; q++) {
    const u64x2 v = s[L::word((u32)q << WB)];
    xa[q] = v.a;
    xb[q] = v.b;
  }
  radix_network<4, INV>(f, xa);
  radix_network<4, INV>(f, xb);
  // register j is tile position i = i_up·16 + j, which holds X[k2 = bitrev12(i)] = X[bitrev4(j)·256 + bitrev8(i_up)]
  const u32 i_up = e0 >> (L::LP + 4), k1p = e0 & ((1u << L::LP) - 1u);
  const u32 k2_t = bitrev(i_up, 8);
  const u64 off_t = ((u64)b << A.log_n) + ((u64)sub << LC) + 2u * k1p + ((u64)k2_t << A.log_n1);
  const u64 stride = (u64)256 << A.log_n1;
  u64* const dst = A.dst + off_t;
  if (FMUL) {
    constexpr int SB = 4;
#pragma unroll
    for (int j0 = 0; j0 < 16; j0 += SB) {
      u64x2 m[SB];
#pragma unroll
      for (int i = 0; i < SB; i++) m[i] = ld_pair(A.mul_src + ((off_t + (u64)br4(j0 + i) * stride) & A.mul_mask));
#pragma unroll
      for (int i = 0; i < SB; i++)
        st_pair(dst + (u64)br4(j0 + i) * stride, u64x2{f.mul(xa[j0 + i], m[i].a), f.mul(xb[j0 + i], m[i].b)});
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; j++) st_pair(dst + (u64)br4(j) * stride, u64x2{xa[j], xb[j]});
  }
}

// PASS1 (C2 = 2 columns in the pass-2 tile): the workspace is W[k1 / 2][j2][k1 % 2].  g = tid + j·NTHR enumerates
//        (k1 block, column): column c = g mod C, block = g / C; its two values k1 = 2·block + h are contiguous in
//        W (16 bytes) and sit in the tile at element index (bitrev12(k1) << LC) | c — half (c & 1) of pair
//        (bitrev12(k1) << LP) | (c >> 1), where bitrev12(2·block + h) = bitrev11(block) | h << 11.  Each is
//        multiplied by the inter-pass twiddle ω_n^(±j2·k1) from the n-word table laid out like W (A.tw_full: one
//        128-bit load at the store's own offset).
template <class F, int LC>
RONK_DEV void n12_store_pass1(const F& f, const u64x2* smem, const NttTileArgs& A, u32 tile, u32 tid) {
  using L = N12<LC, MODE_PASS1>;
  static_assert(L::KK >= LC, "a thread's step must not touch the column bits");
  const u32 b = tile / A.tiles_per_batch, sub = tile - b * A.tiles_per_batch;
  const u32 c = tid & (L::C - 1u), blk_t = tid >> LC;
  const u32 e_t = (bitrev(blk_t, 11) << L::LP) | (c >> 1);
  const u64* const s = reinterpret_cast<const u64*>(smem + L::word(e_t)) + (c & 1u);
  const u32 j2 = (sub << LC) | c;
  const u64 off_t = ((u64)blk_t << (A.log_n2 + 1)) + ((u64)j2 << 1);
  const u64 stride = (u64)(L::NTHR >> LC) << (A.log_n2 + 1);
  u64* const dst = A.dst + ((u64)b << A.log_n) + off_t;
  const u64* const twp = A.tw_full + off_t;
  constexpr u32 EH1 = (1u << 11) << L::LP;  // h = 1: the top bit of the transform index
  constexpr int TB = 8;                     // table loads in flight (128-bit each)
#pragma unroll
  for (int j0 = 0; j0 < 16; j0 += TB) {
    u64x2 w[TB];
#pragma unroll
    for (int i = 0; i < TB; i++) w[i] = ld_pair(twp + (u64)(j0 + i) * stride);
#pragma unroll
    for (int i = 0; i < TB; i++) {
      // step j adds j·NTHR to g: the block index moves by j·(NTHR >> LC); bitrev12(2·x) = bitrev11(x)
      const u32 ej = bitrev12c((((u32)(j0 + i) << L::KK) >> LC) << 1) << L::LP;
      const u64 v0 = s[2u * L::word(ej)], v1 = s[2u * L::word(ej | EH1)];
      st_pair(dst + (u64)(j0 + i) * stride, u64x2{f.mul_tw(v0, w[i].a), f.mul_tw(v1, w[i].b)});
    }
  }
}

#if defined(__CUDACC__)
// Shared memory: [ tile: TILE_SLOTS·16 B | twiddles: TW_WORDS·8 B | mbarrier: 8 B ]
template <class F, int MODE, bool INV, int LC, bool FMUL>
__global__ void __launch_bounds__(N12<LC, MODE>::NTHR, N12<LC, MODE>::NTHR >= 512 ? 1 : 2)
    ntt12_kernel(const F f, const NttTileArgs A) {
  using L = N12<LC, MODE>;
  extern __shared__ __align__(128) u64 smem_raw[];
  u64x2* smem = reinterpret_cast<u64x2*>(smem_raw);
  const u32 tid = threadIdx.x, tile = blockIdx.x;
  u64* tw = smem_raw + 2u * L::TILE_SLOTS;
  u64* bar = tw + L::TW_WORDS;
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar, L::TW_WORDS * 8u);
    tma_bulk_g2s(tw, A.tw_tile, L::TW_WORDS * 8u, bar);  // lands while the tile itself is being loaded
  }
  __syncthreads();  // the mbarrier is initialised before anyone waits on it
  if (MODE == MODE_PASS1 && A.prefetch_dist && tile + A.prefetch_dist < gridDim.x)
    ntt_prefetch_pass1(A, tile + A.prefetch_dist, tid, L::NTHR);
  if (MODE == MODE_PASS2 && A.prefetch_dist && tile + A.prefetch_dist < gridDim.x && (tid & 7u) == 0) {
    // round 0 consumes its loads at once, so their latency is exposed: ask L2 for the tile of the CTA that will
    // follow this one on the SM (one 128-byte line per 8 threads and step)
    const u32 nt = tile + A.prefetch_dist, nb = nt / A.tiles_per_batch, nsub = nt - nb * A.tiles_per_batch;
    const u64* p = A.src + ((u64)nb << A.log_n) + ((u64)nsub << (L::TLP + 1)) + 2u * tid;
#pragma unroll
    for (int q = 0; q < 16; q++) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + ((u64)q << (L::LP + 9))));
  }
  n12_round0_load<F, MODE, INV, LC>(f, smem, tw, A, tile, tid, bar);
  __syncthreads();
  n12_round<F, MODE, INV, LC, 1>(f, smem, tw, tid);
  // rounds 1 and 2 of one i_hi digit are done by the same 2^(LP+4) consecutive threads — half a warp or a warp —
  // so the warps run free from here: no CTA barrier between the rounds
  __syncwarp();
  if constexpr (MODE == MODE_PASS1) {
    n12_round<F, MODE, INV, LC, 2>(f, smem, tw, tid);
    __syncthreads();
    n12_store_pass1<F, LC>(f, smem, A, tile, tid);
  } else {
    n12_round2_store_pass2<F, INV, LC, FMUL>(f, smem, A, tile, tid);
  }
}
#endif  // __CUDACC__

// host: can this launch take the specialised kernel?
inline bool ntt12_applicable(const NttTileArgs& A, int mode) {
  if (A.log_m != 12 || A.src_len != NTT_UNBOUNDED || A.dst_len != NTT_UNBOUNDED) return false;
  if (A.tw_words != N12<1, MODE_PASS2>::TW_WORDS) return false;
  if (((uintptr_t)A.src & 15u) || ((uintptr_t)A.dst & 15u)) return false;  // 128-bit global accesses
  if (mode == MODE_PASS1)  // the stepped form of the inter-pass twiddle lives in the generic kernel only
    return A.log_c == 2 && A.log_c2 == 1 && A.tw_full && !((uintptr_t)A.tw_full & 15u);
  if (mode == MODE_PASS2) return A.log_c == 1 && !((A.flags & NTT_FLAG_MUL) && ((uintptr_t)A.mul_src & 15u));
  return false;
}

}  // namespace ronk
