// ronk_internal.h — context object and launch helpers shared by the translation units of
// libronk_b200.so.  Not part of the public ABI (include/ronk_b200.h is).
#pragma once
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <unordered_set>
#include <vector>

#include "../../include/ronk_b200.h"
#include "field.cuh"

namespace ronk {

struct ProfRec {
  char name[32];
  cudaEvent_t start, stop;
};

struct NttPlan {
  u64 p = 0, g = 0;
  u32 log_n = 0;
  bool two_pass = false;
  u32 log_n1 = 0, log_n2 = 0;
  u64* tw1 = nullptr;        // ω_{N1}^e (single pass: ω_n^e), twiddle form
  u64* tw2 = nullptr;        // ω_{N2}^e == ω_n^(e·N1)
  u64* tw1_2d[2] = {nullptr, nullptr};  // per-round 2-D tables of pass 1 / single (forward, inverse)
  u64* tw2_2d[2] = {nullptr, nullptr};  // per-round 2-D tables of pass 2
  u64* tw_lo = nullptr;      // ω_n^x, x < N1
  u64* tw_hi_inv = nullptr;  // ω_n^(y·N1) · n^-1
  u64 scale_inv = 0;         // n^-1, twiddle form (single-pass inverse)
  // full inter-pass twiddle tables ω_n^(±j2·k1) [· n^-1] in the pass-1 workspace layout, per direction and
  // per log2(C2) of that layout (built on first use; n words each)
  std::map<u32, u64*> tw_full[2];
  // three-pass 2^24 transform (ntt3_kernel.cuh): ω_256^x and the 64 Ki-entry pass-2 table, per direction
  u64* tw256[2] = {nullptr, nullptr};
  u64* t2[2] = {nullptr, nullptr};
  u64* t1[2] = {nullptr, nullptr};  // optional n-word pass-1 twiddle table (RONK_NTT3_T1=1)
};

}  // namespace ronk

// Tuning switches, read from the environment ONCE at ronk_ctx_create (never on the launch path).
struct ronk_tune {
  int ntt3_min_batch16 = 1; // RONK_NTT3_MIN_BATCH16: smallest batch of 2^16-point transforms that takes the 256-point-tile kernels
  int ntt16_cluster_max_batch = 2;  // RONK_NTT16_CLUSTER_MAX_BATCH: up to this many 2^16-point transforms go through ntt16c_kernel
                                    // (one launch, 16-CTA cluster per transform, DSMEM exchange, in place, no workspace); 0 = never.
                                    // Measured (profiles/r02m_ab.txt): 10.5 µs per transform either way at batch 1–2, the two-launch
                                    // path wins from batch 4 on (8 transforms: 10.9 vs 15.6 µs)
  int ntt3_ng1_tiles = 6;   // RONK_NTT3_NG1_TILES: grids below this many tiles per SM run one group per thread (256 threads per tile);
                            // measured (profiles/r02r_switches.txt): 6 vs 3 — two 2^20-point transforms 0.0453 vs 0.0488 ms, rest equal
  int ntt3_split = 1;       // RONK_NTT3_SPLIT: 2^17 … 2^19 and 2^25 / 2^26 as a radix-2/4/8 register pass + interleaved 2^16- / 2^24-point tile transforms
  int ntt3_split_min16 = 1; // RONK_NTT3_SPLIT_MIN16: … for 2^17 … 2^19 only from this many 2^16-point sub-transforms on.  Measured
                            // (profiles/r03e_ab.txt): even ONE transform gains (2^18: 14.6 vs 16.6 µs, 2^19: 16.5 vs 22.7 µs), so 1
  int ntt3_mid = 1;         // RONK_NTT3_MID: 2^21 … 2^23-point transforms through the 256-point-tile kernels (first pass of 32 / 64 / 128 points)
  int ntt3_20 = 1;          // RONK_NTT3_20: 2^20-point transforms as 16 interleaved 2^16-point tile transforms + one radix-16 pass
  int ntt3_t1 = 1;          // RONK_NTT3_T1: pass-1 twiddles ω_n^(k1·m) of the 2^24-point transform from a 128 MiB table per direction
                            // (built at first use; no memory → stepped) instead of stepping.  Measured: 1 % at 6 CTAs per SM (off
                            // then), 3.4 % at 5 (0.2473 vs 0.2559 ms, profiles/r02r_switches.txt) — on
  int ntt3_pdl = 1;         // RONK_NTT3_PDL: programmatic dependent launch between the three passes
  int ntt3 = 1;             // RONK_NTT3: 2^24-point transforms as three passes of 256-point tiles (ntt3_kernel.cuh)
  int pdl = 1;              // RONK_PDL: programmatic dependent launch of pass 2 behind pass 1
  int pf_dist2 = 1;         // RONK_PF_DIST2: the same for pass 2 of the specialised kernel (round 0 fed from HBM)
  int pf_dist = 1;          // RONK_PF_DIST: pass-1 L2 prefetch distance in waves of co-resident CTAs (0 = off)
  int single_tile_log = 12; // RONK_SINGLE_TILE_LOG: preferred tile size when several small transforms share a tile
  int tile1 = 14, tile2 = 13, tile_adapt = 1;  // RONK_TILE1 / RONK_TILE2 / RONK_TILE_ADAPT
  int fast12 = 0;           // RONK_FAST12: the specialised 4096-point-per-tile kernel (ntt12_kernel.cuh) where it applies — opt-in:
                            // 20 % fewer instructions but slower on B200 (0.392 vs 0.379 ms); needs RONK_TW_TABLE=1 for pass 1
  int msm_split = 0;        // RONK_MSM_SPLIT: ≥ 2^22 terms: every other term to an L2-resident histogram (global RED)
  int msm_coord = 1;        // RONK_MSM_COORD: kzg::commit in group coordinates (two dot products mod 102 + one lookup); 0 = the paths below
  int msm_hist = 1;         // RONK_MSM_HIST: kzg::commit through the point-indexed histogram (1) or the bucket kernels (0)
  int tw_table = 0;         // RONK_TW_TABLE: inter-pass twiddles from an n-word table (1) or stepped w ← w·ρ (0)
};

struct ronk_ctx {
  int device = 0;
  ronk_tune tune;
  std::unordered_set<const void*> smem_attr_done;  // kernels whose >48 KiB dynamic-smem attribute is set on `device`
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  std::string err;
  uint64_t launches = 0;
  bool prof = false;
  std::vector<ronk::ProfRec> prof_log;
  std::map<std::tuple<uint64_t, uint64_t, uint32_t>, ronk::NttPlan> plans;
  void* ws = nullptr;  // workspace (two-pass intermediate, poly_mul operands, msm partials)
  size_t ws_bytes = 0;
  void* ws2 = nullptr;  // second scratch buffer (poly_mul)
  size_t ws2_bytes = 0;
  // two-slot host pipeline (ronk_ntt_u64_host_submit / _wait)
  cudaStream_t copy_in = nullptr, copy_out = nullptr;
  static constexpr int kSlots = 3;
  void* slot_buf[kSlots] = {};
  size_t slot_bytes[kSlots] = {};
  cudaEvent_t ev_h2d[kSlots] = {}, ev_compute[kSlots] = {}, ev_d2h[kSlots] = {};
  bool slot_pending[kSlots] = {};
  int cluster16_state = 0;   // ntt16c_kernel: 0 = not probed, 1 = usable, -1 = the device refuses 16-CTA clusters of its footprint
  void* dist = nullptr;      // ronk::DistState (dist.cu): communicator, peer mappings, staging — null until ronk_dist_init
  void* msm_ytab = nullptr;  // uint16_t[20402]: y of the curve point in each histogram bin (msm.cu), built on first use
  void* msm_done = nullptr;  // u32 completion counter of msm_hist_finish_kernel
  void* msm_coord = nullptr; // msm_coord_kernel: bintab[20404] | pttab[10404] | counter, Σa, Σb (msm.cu), built on first use
  int* d_flag = nullptr;  // device error flag
  int* h_flag = nullptr;  // pinned host mirror: h_flag[0] = error flag, h_flag[1..31] = small results (msm.cu)
};

namespace ronk {

inline int set_err(ronk_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}

#define RONK_CUDA(ctx, expr)                                                                  \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return ronk::set_err(ctx, _e == cudaErrorMemoryAllocation ? RONK_ENOMEM : RONK_ECUDA,   \
                           std::string(#expr ": ") + cudaGetErrorString(_e));                 \
  } while (0)

#define RONK_TRY(expr)          \
  do {                          \
    int _rc = (expr);           \
    if (_rc != RONK_OK) return _rc; \
  } while (0)

// Binds the calling thread to the context's device for the duration of one C-ABI call and restores the
// caller's device on exit (contexts for several GPUs may coexist in one process; the caller — torch, a
// Rust host — keeps its own current device).  Every extern "C" entry point that takes a ctx opens with it,
// so allocations, attribute calls and launches inside the call all land on ctx->device.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(const ronk_ctx* ctx) {
    if (!ctx) return;
    if (cudaGetDevice(&prev) == cudaSuccess && prev != ctx->device) switched = cudaSetDevice(ctx->device) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// Opt a kernel into > 48 KiB of dynamic shared memory, once per (context, kernel).
template <class K>
inline int ensure_smem_attr(ronk_ctx* ctx, K kernel, int bytes);

// Brackets a kernel launch with the launch counter and (optionally) profiling events.
struct LaunchScope {
  ronk_ctx* ctx;
  bool on;
  cudaEvent_t start = nullptr, stop = nullptr;
  const char* name;
  LaunchScope(ronk_ctx* c, const char* n) : ctx(c), on(c->prof), name(n) {
    ctx->launches++;  // the device is already bound by the entry point's DeviceGuard
    if (on) {
      cudaEventCreate(&start);
      cudaEventCreate(&stop);
      cudaEventRecord(start, ctx->stream);
    }
  }
  ~LaunchScope() {
    if (on) {
      cudaEventRecord(stop, ctx->stream);
      ProfRec r;
      std::memset(&r, 0, sizeof(r));
      std::snprintf(r.name, sizeof(r.name), "%s", name);
      r.start = start;
      r.stop = stop;
      ctx->prof_log.push_back(r);
    }
  }
};

template <class K>
inline int ensure_smem_attr(ronk_ctx* ctx, K kernel, int bytes) {
  const void* key = reinterpret_cast<const void*>(kernel);
  if (ctx->smem_attr_done.count(key)) return RONK_OK;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return set_err(ctx, RONK_ECUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
  ctx->smem_attr_done.insert(key);
  return RONK_OK;
}

inline int check_launch(ronk_ctx* ctx, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err(ctx, RONK_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
  return RONK_OK;
}

inline int ensure_ws(ronk_ctx* ctx, void** buf, size_t* cap, size_t bytes) {
  if (*cap >= bytes) return RONK_OK;
  if (*buf) {
    cudaStreamSynchronize(ctx->stream);
    cudaFree(*buf);
    *buf = nullptr;
    *cap = 0;
  }
  cudaError_t e = cudaMalloc(buf, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_err(ctx, RONK_ENOMEM, "workspace allocation failed");
  }
  *cap = bytes;
  return RONK_OK;
}

// Field policy construction (host).
inline bool is_goldilocks_fast(u64 p, u64 g) { return p == GL_P && g == 7; }
int make_mont_field(ronk_ctx* ctx, u64 p, u64 g, bool inverse, MontField* out);  // ntt.cu
int validate_modulus(ronk_ctx* ctx, u64 p);                                       // field_ops.cu

// Internal device-pointer entry points used across translation units.
int ntt_device(ronk_ctx* ctx, u64 p, u64 g, u64* data, const u64* mul, u32 log_n, u32 batch, int inverse);
int ntt_device_shared_mul(ronk_ctx* ctx, u64 p, u64 g, const u64* src, u64* dst, const u64* mul, u32 log_n, u32 batch);
int ntt_device_bounded(ronk_ctx* ctx, u64 p, u64 g, const u64* src, u64 src_len, u64* dst, u64 dst_len, const u64* mul,
                       u32 log_n, int inverse);

}  // namespace ronk
