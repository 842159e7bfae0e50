// curve_emu.cpp — TEST INFRASTRUCTURE: compiles ronkathon_b200/csrc/msm_curve.cuh for the host so the CPU test
// tier can check the GF(101²) / curve arithmetic the MSM kernels use (both the Fermat-inverse and the
// table-inverse addition) against the oracle without a GPU.  Never linked into libronk_b200.so.
#include <cstdint>

#include "../../ronkathon_b200/csrc/msm_curve.cuh"

using namespace ronk;

static u32 pack4(const uint8_t* p) { return (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24); }
static void unpack4(u32 w, uint8_t* p) {
  p[0] = (uint8_t)w; p[1] = (uint8_t)(w >> 8); p[2] = (uint8_t)(w >> 16); p[3] = (uint8_t)(w >> 24);
}

extern "C" {

// element-wise over n packed points (4 bytes each, 0xFF×4 = Infinity); use_table selects pt_add_t
void emu_point_add(const uint8_t* a, const uint8_t* b, uint8_t* out, uint64_t n, int use_table) {
  uint8_t tab[104];
  build_inv_table(tab, 0, 1);
  for (uint64_t i = 0; i < n; i++) {
    const u32 wa = pack4(a + 4 * i), wb = pack4(b + 4 * i);
    unpack4(use_table ? pt_add_t(wa, wb, tab) : pt_add_w(wa, wb), out + 4 * i);
  }
}
void emu_point_valid(const uint8_t* a, uint8_t* ok, uint64_t n) {
  for (uint64_t i = 0; i < n; i++) ok[i] = pt_valid(pack4(a + 4 * i)) ? 1 : 0;
}
// GF(101²): op 0 mul, 1 inverse of a (b ignored); elements as (c0, c1) byte pairs
void emu_gf_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out, uint64_t n) {
  for (uint64_t i = 0; i < n; i++) {
    const Gf x = {a[2 * i], a[2 * i + 1]}, y = {b[2 * i], b[2 * i + 1]};
    const Gf r = op == 0 ? gf_mul(x, y) : gf_inv(x);
    out[2 * i] = (uint8_t)r.c0; out[2 * i + 1] = (uint8_t)r.c1;
  }
}
// the bucket combination of msm_finish_kernel's last warp, lane-parallel on 16 values:
// suffix scan (run_k = B_{k+1} + … + B_16) then tree sum — must equal Σ s·B_s
void emu_bucket_combine(const uint8_t* buckets16, uint8_t* out) {
  uint8_t tab[104];
  build_inv_table(tab, 0, 1);
  u32 v[32];
  for (int k = 0; k < 32; k++) v[k] = k < 16 ? pack4(buckets16 + 4 * k) : PT_INF;
  for (int off = 1; off < 16; off <<= 1) {
    u32 nv[32];
    for (int k = 0; k < 32; k++) {
      const u32 other = k + off < 32 ? v[k + off] : v[k];  // __shfl_down semantics
      nv[k] = (k + off < 16) ? pt_add_t(v[k], other, tab) : v[k];
    }
    for (int k = 0; k < 32; k++) v[k] = nv[k];
  }
  for (int off = 8; off > 0; off >>= 1) {
    u32 nv[32];
    for (int k = 0; k < 32; k++) {
      const u32 other = k + off < 32 ? v[k + off] : v[k];
      nv[k] = (k < off) ? pt_add_t(v[k], other, tab) : v[k];
    }
    for (int k = 0; k < 32; k++) v[k] = nv[k];
  }
  unpack4(v[0], out);
}

// group-coordinate tables of msm_coord_kernel (build_group_tables, host plan code): bintab[MSM_BINS], pttab[102²];
// returns 1 when a basis was found and (a, b) → a·G1 + b·G2 is injective
int emu_group_tables(uint32_t* bintab, uint32_t* pttab) { return build_group_tables(bintab, pttab) ? 1 : 0; }
// the per-term step of msm_coord_kernel and its final lookup, serially: Σ s_i·P_i through the tables.
// returns 0 ok, 1 rejected term (the kernel's error flag)
int emu_coord_commit(const uint32_t* bintab, const uint32_t* pttab, const uint8_t* points, const uint8_t* scalars, uint64_t n,
                     uint8_t* out) {
  u32 acc_a = 0, acc_b = 0;
  for (uint64_t i = 0; i < n; i++) {
    const u32 w = pack4(points + 4 * i), s = scalars[i];
    if (s >= 17u) return 1;
    if (w == PT_INF) continue;
    if ((w & 0xFF) >= Q101 || ((w >> 8) & 0xFF) >= Q101 || ((w >> 16) & 0xFF) >= Q101 || (w >> 24) >= Q101) return 1;
    const u32 e = bintab[pt_bin(w)];
    if ((e & 0xFFFFu) != (w >> 16)) return 1;
    acc_a = (acc_a + s * ((e >> 16) & 0xFFu)) % MSM_EXP;
    acc_b = (acc_b + s * (e >> 24)) % MSM_EXP;
  }
  unpack4(pttab[MSM_EXP * acc_a + acc_b], out);
  return 0;
}

}  // extern "C"
