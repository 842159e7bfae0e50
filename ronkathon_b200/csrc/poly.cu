// poly.cu — Polynomial<Monomial|Lagrange, F, D> operations other than the transforms.
// Mirrors src/polynomial/arithmetic.rs (Add :16-35, Sub :49-68, Mul :97-119, Div/Rem :121-146) and
// src/polynomial/mod.rs (evaluate :133-139, quotient_and_remainder :170-225, dft :240-258,
// Lagrange evaluate :382-415).
#include "ntt_kernel.cuh"
#include "ronk_internal.h"

namespace ronk {

// out[i] = a[i] ± (i < db ? b[i] : 0), i < da   (arithmetic.rs:23-34 / :56-67)
template <class F, bool SUB>
__global__ void poly_addsub_kernel(const F f, const u64* a, size_t da, const u64* b, size_t db, u64* out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < da; i += stride) {
    const u64 y = (i < db) ? b[i] : 0ULL;
    out[i] = SUB ? f.sub(a[i], y) : f.add(a[i], y);
  }
}

// Schoolbook product (arithmetic.rs:110-118): one thread per output coefficient.
template <class F>
__global__ void poly_mul_schoolbook_kernel(const F f, const u64* a, size_t da, const u64* b, size_t db, u64* c) {
  const size_t L = da + db - 1;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < L; k += stride) {
    const size_t lo = (k >= db) ? k - db + 1 : 0, hi = (k < da) ? k : da - 1;
    u64 acc = 0;
    for (size_t i = lo; i <= hi; i++) acc = f.add(acc, f.mul(a[i], b[k - i]));
    c[k] = acc;
  }
}

// evaluate (mod.rs:133-139): out[pt] = Σ_j c_j x^j.  One CTA per point; thread t owns the
// coefficients j ≡ t (mod 256) (coalesced), Horner in x^256, then a shared-memory tree sum.
template <class F>
__global__ void poly_eval_kernel(const F f, const u64* c, size_t d, const u64* xs, u64* out) {
  __shared__ u64 red[256];
  const u32 t = threadIdx.x;
  const u64 x = xs[blockIdx.x];
  const u64 y = field_pow(f, x, 256);
  u64 acc = 0;
  if (t < d) {
    const size_t kmax = (d - 1 - t) / 256;
    for (size_t k = kmax + 1; k-- > 0;) acc = f.add(f.mul(acc, y), c[t + 256 * k]);
    acc = f.mul(acc, field_pow(f, x, (u64)t));
  }
  red[t] = acc;
  __syncthreads();
  for (u32 s = 128; s > 0; s >>= 1) {
    if (t < s) red[t] = f.add(red[t], red[t + s]);
    __syncthreads();
  }
  if (t == 0) out[blockIdx.x] = red[0];
}

// Lagrange-basis evaluate (mod.rs:382-415), literal fold semantics: the closure's early
// `return c` replaces the accumulator, and l(x) multiplies the fold afterwards.
template <class F>
__global__ void lagrange_eval_kernel(const F f, const u64* c, const u64* nodes, u32 n, u64 x, u64* out, int* flag) {
  __shared__ u64 red[256];
  __shared__ u64 lred[256];
  __shared__ u32 hit;  // index j* with nodes[j*] == x, or n
  const u32 t = threadIdx.x;
  if (t == 0) hit = n;
  __syncthreads();
  for (u32 j = t; j < n; j += blockDim.x)
    if (nodes[j] == x) hit = j;
  __syncthreads();
  const u32 jstar = hit;
  u64 acc = 0, lx = 1 % f.modulus();
  for (u32 j = t; j < n; j += blockDim.x) {
    lx = f.mul(lx, f.sub(x, nodes[j]));
    // The reference builds EVERY weight first (mod.rs:386-393), so a repeated node panics in F::ONE.div
    // whatever x is — also for the terms the fold later discards.
    u64 wden = 1 % f.modulus();  // w_j^-1 = Π_{m≠j} (x_j - x_m)
    for (u32 m = 0; m < n; m++)
      if (m != j) wden = f.mul(wden, f.sub(nodes[j], nodes[m]));
    if (wden == 0) { atomicExch(flag, 1); continue; }
    if (jstar < n && j <= jstar) {
      if (j == jstar) acc = f.add(acc, c[j]);
      continue;
    }
    const u64 den = f.mul(wden, f.sub(x, nodes[j]));
    if (den == 0) { atomicExch(flag, 1); continue; }
    acc = f.add(acc, f.mul(c[j], field_pow(f, den, f.modulus() - 2)));
  }
  red[t] = acc;
  lred[t] = lx;
  __syncthreads();
  for (u32 s = blockDim.x / 2; s > 0; s >>= 1) {
    if (t < s) {
      red[t] = f.add(red[t], red[t + s]);
      lred[t] = f.mul(lred[t], lred[t + s]);
    }
    __syncthreads();
  }
  if (t == 0) out[0] = f.mul(lred[0], red[0]);
}

// quotient_and_remainder (mod.rs:170-225).  Single CTA; q and r have da terms.
// flag: 1 = all-zero divisor / non-invertible, 2 = index out of range (the reference panics).
template <class F>
__global__ void poly_divrem_kernel(const F f, const u64* a, u32 da, const u64* b, u32 db, u64* q, u64* r, int* flag) {
  __shared__ u32 s_deg, s_ok;
  __shared__ u64 s_s;
  const u32 t = threadIdx.x, nt = blockDim.x;
  for (u32 i = t; i < da; i += nt) { q[i] = 0; r[i] = a[i]; }
  __shared__ u32 s_rdeg, s_rfound;
  if (t == 0) { s_rfound = 0; s_rdeg = 0; }
  __syncthreads();
  for (u32 i = t; i < db; i += nt)
    if (b[i] != 0) { atomicMax(&s_rdeg, i); s_rfound = 1; }
  __syncthreads();
  const u32 rhs_degree = s_rdeg;
  const bool rhs_nonzero = s_rfound != 0;
  u64 cinv = 0;
  if (rhs_nonzero) cinv = field_pow(f, b[rhs_degree], f.modulus() - 2);
  u32 plen = da;  // p_coeffs.len() after trim_zeros
  while (true) {
    // find degree of the (trimmed) dividend: highest nonzero below plen
    if (t == 0) { s_deg = 0; s_ok = 0; }
    __syncthreads();
    for (u32 i = t; i < plen; i += nt)
      if (r[i] != 0) { atomicMax(&s_deg, i); s_ok = 1; }
    __syncthreads();
    const bool any = s_ok != 0;
    const u32 p_degree = s_deg;
    if (!(any && plen >= db)) break;                       // :183-184
    if (!rhs_nonzero) { if (t == 0) atomicExch(flag, 1); break; }  // rposition().unwrap()
    if (p_degree < rhs_degree) break;                      // :190-192
    const u32 diff = p_degree - rhs_degree;
    if (diff + db > plen) { if (t == 0) atomicExch(flag, 2); break; }  // p_coeffs[diff + i] out of range
    __syncthreads();
    if (t == 0) { s_s = f.mul(r[p_degree], cinv); q[diff] = s_s; }
    __syncthreads();
    const u64 s = s_s;
    for (u32 i = t; i < db; i += nt) r[diff + i] = f.sub(r[diff + i], f.mul(b[i], s));
    __syncthreads();
    plen = p_degree;  // coefficient p_degree is now 0; trim_zeros continues below via the next scan
    // trim_zeros pops every trailing zero: the next iteration's scan finds the new top, and
    // plen must equal (new top + 1) for the `plen >= db` test.
    if (t == 0) { s_deg = 0; s_ok = 0; }
    __syncthreads();
    for (u32 i = t; i < plen; i += nt)
      if (r[i] != 0) { atomicMax(&s_deg, i); s_ok = 1; }
    __syncthreads();
    plen = s_ok ? s_deg + 1 : 0;
    __syncthreads();
  }
}


// ---- division by a linear factor (b0 + b1·x), §8f row 1 ------------------------------------------
// quotient_and_remainder (mod.rs:170-225) specialised to the divisor kzg::open builds
// (kzg/setup.rs:72-75: [-z, 1]).  With z = -b0/b1 the long division is the suffix recurrence
//   h_j = a_j + z·h_{j+1}  (h_d = 0),   q_j = h_{j+1} / b1,   r = h_0 = a(z),
// a first-order linear recurrence, i.e. a scan: (1) every chunk of 4096 coefficients is folded to
// its Horner value S_c = Σ a_{c0+i} z^i, (2) one CTA turns the S_c into the carry-in of every chunk,
// (3) every chunk redoes its local scan seeded with the carry and writes q.  2 reads + 1 write per
// coefficient; the general single-CTA kernel above needs O(D) sequential steps for the same result.
constexpr int DL_THR = 256, DL_PER = 16, DL_CHUNK = DL_THR * DL_PER;
RONK_DEV u32 dl_pad(u32 i) { return i + (i >> 4); }  // 16-word rows padded to 17: conflict-free both ways

template <class F>
RONK_DEV u64 pow2k_tw(const F& f, u64 x_tw, int k) {  // x^(2^k), twiddle form in and out
  for (int i = 0; i < k; i++) x_tw = f.mul_tw(x_tw, x_tw);
  return x_tw;
}

template <class F, bool APPLY>
__global__ void __launch_bounds__(DL_THR)
div_linear_kernel(const F f, const u64* __restrict__ a, u64 d, u64 z, const u64* __restrict__ carry, u64* __restrict__ S,
                  u64 scale, u64* __restrict__ q, u64* __restrict__ rem) {
  __shared__ u64 tile[DL_CHUNK + DL_CHUNK / 16];
  __shared__ u64 v[DL_THR + 1];
  const u32 tid = threadIdx.x;
  const u64 c0 = (u64)blockIdx.x * DL_CHUNK;
  const u64 z_tw = f.to_tw(z);
  for (u32 i = tid; i < (u32)DL_CHUNK; i += DL_THR) tile[dl_pad(i)] = (c0 + i < d) ? a[c0 + i] : 0ULL;  // coalesced
  __syncthreads();
  u64 x[DL_PER];
#pragma unroll
  for (int k = 0; k < DL_PER; k++) x[k] = tile[dl_pad(tid * DL_PER + k)];
  u64 t = 0;  // Σ_k x[k] z^k
#pragma unroll
  for (int k = DL_PER - 1; k >= 0; k--) t = f.add(x[k], f.mul_tw(t, z_tw));
  v[tid] = t;
  if (tid == 0) v[DL_THR] = APPLY ? carry[blockIdx.x] : 0ULL;  // element 256 = carry into this chunk
  u64 M = pow2k_tw(f, z_tw, 4);                                  // z^16
  for (u32 off = 1; off <= (u32)DL_THR; off <<= 1) {             // suffix scan, uniform multiplier per level
    __syncthreads();
    const bool on = tid + off <= (u32)DL_THR;
    const u64 other = on ? v[tid + off] : 0ULL;
    __syncthreads();
    if (on) v[tid] = f.add(v[tid], f.mul_tw(other, M));
    M = f.mul_tw(M, M);
  }
  __syncthreads();
  if (!APPLY) {
    if (tid == 0) S[blockIdx.x] = v[0];
    return;
  }
  const u64 s_tw = f.to_tw(scale);
  u64 run = v[tid + 1];  // h at the first index above this thread's range
#pragma unroll
  for (int k = DL_PER - 1; k >= 0; k--) {
    run = f.add(x[k], f.mul_tw(run, z_tw));  // h_{base+k}
    x[k] = run;
  }
  if (blockIdx.x == 0 && tid == 0) rem[0] = x[0];  // r = h_0 (not scaled: a = q·(b1·x + b0) + r)
#pragma unroll
  for (int k = 0; k < DL_PER; k++) tile[dl_pad(tid * DL_PER + k)] = f.mul_tw(x[k], s_tw);
  __syncthreads();
  // q_j = h_{j+1}/b1: h at chunk position i goes to q[c0 + i - 1]; the top coefficient q_{d-1} is 0
  for (u32 i = tid; i < (u32)DL_CHUNK; i += DL_THR) {
    const u64 j = c0 + i;
    if (j >= 1 && j < d) q[j - 1] = tile[dl_pad(i)];
  }
  if (c0 + DL_CHUNK >= d && c0 < d && tid == 0) q[d - 1] = 0ULL;
}

// carry[c] = Σ_{c' > c} S_{c'} Z^{c'-c-1}, Z = z^4096.  One CTA: every thread owns a contiguous block of
// chunks (local Horner), thread 0 chains the ≤ 1024 block values, then every thread replays its block.
template <class F>
__global__ void __launch_bounds__(1024)
div_linear_carry_kernel(const F f, const u64* __restrict__ S, u32 nchunks, u64 z, u64* __restrict__ carry) {
  __shared__ u64 L[1024];
  __shared__ u64 G[1025];
  const u32 t = threadIdx.x;
  const u64 Z = pow2k_tw(f, f.to_tw(z), 12);
  const u32 B = (nchunks + 1023) / 1024;
  const u32 lo = (u32)min((u64)t * B, (u64)nchunks), hi = (u32)min((u64)lo + B, (u64)nchunks);
  u64 acc = 0;
  for (u32 c = hi; c-- > lo;) acc = f.add(S[c], f.mul_tw(acc, Z));
  L[t] = acc;
  __syncthreads();
  if (t == 0) {
    u64 ZB = f.to_tw(1 % f.modulus());  // Z^B
    {
      u64 base = Z;
      for (u32 e = B; e; e >>= 1) {
        if (e & 1) ZB = f.mul_tw(ZB, base);
        base = f.mul_tw(base, base);
      }
    }
    u64 g = 0;
    G[1024] = 0;
    for (u32 i = 1024; i-- > 0;) {
      g = f.add(L[i], f.mul_tw(g, ZB));
      G[i] = g;  // Horner value of all chunks from block i upwards
    }
  }
  __syncthreads();
  u64 run = G[t + 1];
  for (u32 c = hi; c-- > lo;) {
    carry[c] = run;
    run = f.add(S[c], f.mul_tw(run, Z));
  }
}

template <class F>
static int div_linear_with_field(ronk_ctx* ctx, const F& f, const u64* a, size_t d, u64 z, u64 scale, u64* q, u64* rem) {
  const size_t nchunks = (d + DL_CHUNK - 1) / DL_CHUNK;
  if (nchunks > 0x7FFFFFFFULL) return set_err(ctx, RONK_EUNSUPPORTED, "polynomial too long");
  RONK_TRY(ensure_ws(ctx, &ctx->ws2, &ctx->ws2_bytes, 2 * nchunks * sizeof(u64)));
  u64* S = (u64*)ctx->ws2;
  u64* carry = S + nchunks;
  {
    LaunchScope ls(ctx, "div_linear_fold");
    div_linear_kernel<F, false><<<(u32)nchunks, DL_THR, 0, ctx->stream>>>(f, a, d, z, nullptr, S, scale, nullptr, nullptr);
  }
  RONK_TRY(check_launch(ctx, "div_linear_kernel<fold>"));
  {
    LaunchScope ls(ctx, "div_linear_carry");
    div_linear_carry_kernel<F><<<1, 1024, 0, ctx->stream>>>(f, S, (u32)nchunks, z, carry);
  }
  RONK_TRY(check_launch(ctx, "div_linear_carry_kernel"));
  {
    LaunchScope ls(ctx, "div_linear_apply");
    div_linear_kernel<F, true><<<(u32)nchunks, DL_THR, 0, ctx->stream>>>(f, a, d, z, carry, nullptr, scale, q, rem);
  }
  return check_launch(ctx, "div_linear_kernel<apply>");
}

// a / (b0 + b1·x): q (d terms, top one 0) and the scalar remainder, all device pointers; q may not alias a.
static int div_linear_device(ronk_ctx* ctx, u64 p, const u64* a, size_t d, u64 b0, u64 b1, u64* q, u64* rem) {
  if (!ctx || !a || !q || !rem) return set_err(ctx, RONK_EINVAL, "null argument");
  if (q == a) return set_err(ctx, RONK_EINVAL, "the quotient may not alias the dividend");
  RONK_TRY(validate_modulus(ctx, p));
  if (d == 0) return set_err(ctx, RONK_EINVAL, "empty dividend");
  if (b0 >= p || b1 >= p) return set_err(ctx, RONK_EINVAL, "non-canonical divisor coefficient");
  if (b1 == 0) return set_err(ctx, RONK_EINVAL, "divisor is not linear (leading coefficient 0)");
  const u64 b1inv = h_powmod(b1, p - 2, p);
  const u64 z = h_mulmod(b0 ? p - b0 : 0, b1inv, p);
  if (p == GL_P) {
    GoldilocksField f;
    return div_linear_with_field(ctx, f, a, d, z, b1inv, q, rem);
  }
  MontField f;
  RONK_TRY(make_mont_field(ctx, p, 0, false, &f));
  return div_linear_with_field(ctx, f, a, d, z, b1inv, q, rem);
}

// ---- Lagrange interpolation (Reed–Solomon decode, §8f row 2) ------------------------------------
// Message::decode (codes/reed_solomon.rs:55-107) interpolates the first K coordinates:
//   data[i] = Σ_j y_j · (-1)^i e_{K-1-i}(x \ x_j) / Π_{k≠j}(x_k - x_j)
// which is coefficient i of Σ_j y_j · M(X)/(X - x_j) / M'(x_j), M = Π (X - x_k).  The reference
// enumerates combinations (exponential in K); the value is the unique interpolant, computed here as
// (1) M by K in-place products, (2) per node j a synthetic division M/(X - x_j) giving q_j and
// d_j = q_j(x_j) = Π_{k≠j}(x_j - x_k), c_j = y_j/d_j, (3) out = Σ_j c_j q_j with a warp-shuffle
// reduction per coefficient.  A repeated node gives d_j = 0: the reference's `/` panics → flag.
template <class F>
__global__ void __launch_bounds__(1024)
interp_master_kernel(const F f, const u64* __restrict__ xs, u32 k, u64* __restrict__ m0, u64* __restrict__ m1) {
  // m0/m1: k+1 words each, ping-pong; the result ends in (k odd ? m1 : m0)
  const u32 t = threadIdx.x;
  for (u32 i = t; i <= k; i += blockDim.x) m0[i] = (i == 0) ? 1 % f.modulus() : 0ULL;
  __syncthreads();
  u64* cur = m0;
  u64* nxt = m1;
  for (u32 s = 0; s < k; s++) {  // cur has degree s; nxt = cur · (X - x_s)
    const u64 x = xs[s];
    for (u32 i = t; i <= s + 1; i += blockDim.x) {
      const u64 lo = (i >= 1) ? cur[i - 1] : 0ULL;
      const u64 hi = (i <= s) ? f.mul(cur[i], x) : 0ULL;
      nxt[i] = f.sub(lo, hi);
    }
    __syncthreads();
    u64* tmp = cur; cur = nxt; nxt = tmp;
  }
}

template <class F>
__global__ void __launch_bounds__(256)
interp_nodes_kernel(const F f, const u64* __restrict__ M, const u64* __restrict__ xs, const u64* __restrict__ ys, u32 k,
                    u64* __restrict__ partial /* [ceil(k/32)][k] */, int* flag) {
  const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = j < k;
  const u64 x = live ? xs[j] : 0ULL;
  // pass 1: d_j = q_j(x_j), with q_j[i-1] = M[i] + x_j·q_j[i], q_j[k-1] = M[k] = 1, Horner from the top
  u64 qv = 0, d = 0;
  for (u32 i = k; i >= 1; i--) {
    qv = f.add(M[i], f.mul(qv, x));  // q_j[i-1]
    d = f.add(f.mul(d, x), qv);
  }
  u64 c = 0;
  if (live) {
    if (d == 0) atomicExch(flag, 1);
    else c = f.mul(ys[j], field_pow(f, d, f.modulus() - 2));
  }
  // pass 2: Σ over the warp's nodes of c_j · q_j[i-1]
  const u32 warp = j >> 5, lane = threadIdx.x & 31;
  qv = 0;
  for (u32 i = k; i >= 1; i--) {
    qv = f.add(M[i], f.mul(qv, x));
    u64 term = f.mul(c, qv);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) term = f.add(term, __shfl_down_sync(0xFFFFFFFFu, term, off));
    if (lane == 0) partial[(size_t)warp * k + (i - 1)] = term;
  }
}

template <class F>
__global__ void interp_sum_kernel(const F f, const u64* __restrict__ partial, u32 k, u32 nwarps, u64* __restrict__ out) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k) return;
  u64 acc = 0;
  for (u32 w = 0; w < nwarps; w++) acc = f.add(acc, partial[(size_t)w * k + i]);
  out[i] = acc;
}

template <class F>
static int interp_with_field(ronk_ctx* ctx, const F& f, const u64* xs, const u64* ys, u32 k, u64* out) {
  const u32 blocks = (k + 255) / 256, nwarps = blocks * 8;
  const size_t words = 2 * (size_t)(k + 1) + (size_t)nwarps * k;
  RONK_TRY(ensure_ws(ctx, &ctx->ws2, &ctx->ws2_bytes, words * sizeof(u64)));
  u64* m0 = (u64*)ctx->ws2;
  u64* m1 = m0 + (k + 1);
  u64* partial = m1 + (k + 1);
  RONK_CUDA(ctx, cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), ctx->stream));
  {
    LaunchScope ls(ctx, "interp_master");
    interp_master_kernel<F><<<1, 1024, 0, ctx->stream>>>(f, xs, k, m0, m1);
  }
  RONK_TRY(check_launch(ctx, "interp_master_kernel"));
  const u64* M = (k & 1) ? m1 : m0;
  {
    LaunchScope ls(ctx, "interp_nodes");
    interp_nodes_kernel<F><<<blocks, 256, 0, ctx->stream>>>(f, M, xs, ys, k, partial, ctx->d_flag);
  }
  RONK_TRY(check_launch(ctx, "interp_nodes_kernel"));
  {
    LaunchScope ls(ctx, "interp_sum");
    interp_sum_kernel<F><<<(k + 255) / 256, 256, 0, ctx->stream>>>(f, partial, k, nwarps, out);
  }
  return check_launch(ctx, "interp_sum_kernel");
}

static int grid_for(ronk_ctx* ctx, size_t n, int threads) {
  size_t blocks = (n + threads - 1) / threads;
  size_t cap = (size_t)ctx->sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks == 0) blocks = 1;
  return (int)blocks;
}

template <class F>
static int poly_mul_with_field(ronk_ctx* ctx, const F& f, u64 p, u64 g, const u64* a, size_t da, const u64* b,
                               size_t db, u64* c) {
  const size_t L = da + db - 1;
  u32 log_n = 0;
  while (((size_t)1 << log_n) < L) log_n++;
  const bool ntt_ok = g != 0 && log_n >= 1 && log_n <= 26 && (p - 1) % ((u64)1 << log_n) == 0;
  // NTT cost ~ 3·n·log n / 2 multiplies vs da·db for schoolbook
  const double school = (double)da * (double)db;
  const double viantt = 1.5 * (double)((size_t)1 << log_n) * (double)log_n + 4096.0;
  if (!ntt_ok || school <= viantt) {
    LaunchScope ls(ctx, "poly_mul_schoolbook");
    poly_mul_schoolbook_kernel<F><<<grid_for(ctx, L, 128), 128, 0, ctx->stream>>>(f, a, da, b, db, c);
    return check_launch(ctx, "poly_mul_schoolbook_kernel");
  }
  const size_t n = (size_t)1 << log_n;
  RONK_TRY(ensure_ws(ctx, &ctx->ws2, &ctx->ws2_bytes, 2 * n * sizeof(u64)));
  u64* A = (u64*)ctx->ws2;
  u64* B = A + n;
  // the zero padding of a and b and the clipping of the product to L coefficients happen inside the
  // transforms' load / store phases (no pad-copy kernels, no final device copy)
  RONK_TRY(ntt_device_bounded(ctx, p, g, a, da, A, n, nullptr, log_n, 0));  // Â
  RONK_TRY(ntt_device_bounded(ctx, p, g, b, db, B, n, A, log_n, 0));        // B̂ ⊙ Â fused into the last stage
  return ntt_device_bounded(ctx, p, g, B, n, c, L, nullptr, log_n, 1);      // back to coefficients, L of them
}

static int poly_mul_device(ronk_ctx* ctx, u64 p, u64 g, const u64* a, size_t da, const u64* b, size_t db, u64* c) {
  if (!ctx || !a || !b || !c) return set_err(ctx, RONK_EINVAL, "null argument");
  if (da == 0 || db == 0) return set_err(ctx, RONK_EINVAL, "empty polynomial (D + D2 - 1 underflows)");
  RONK_TRY(validate_modulus(ctx, p));
  if (is_goldilocks_fast(p, g) || (p == GL_P && g == 0)) {
    GoldilocksField f;
    return poly_mul_with_field(ctx, f, p, g, a, da, b, db, c);
  }
  MontField f;
  RONK_TRY(make_mont_field(ctx, p, 0, false, &f));
  return poly_mul_with_field(ctx, f, p, g, a, da, b, db, c);
}

template <bool SUB>
static int poly_addsub(ronk_ctx* ctx, u64 p, const u64* a, size_t da, const u64* b, size_t db, u64* out) {
  if (!ctx || (da && (!a || !out)) || (db && !b)) return set_err(ctx, RONK_EINVAL, "null argument");
  RONK_TRY(validate_modulus(ctx, p));
  if (da == 0) return RONK_OK;
  const char* name = SUB ? "poly_sub" : "poly_add";
  if (p == GL_P) {
    GoldilocksField f;
    LaunchScope ls(ctx, name);
    poly_addsub_kernel<GoldilocksField, SUB><<<grid_for(ctx, da, 256), 256, 0, ctx->stream>>>(f, a, da, b, db, out);
  } else {
    MontField f;
    RONK_TRY(make_mont_field(ctx, p, 0, false, &f));
    LaunchScope ls(ctx, name);
    poly_addsub_kernel<MontField, SUB><<<grid_for(ctx, da, 256), 256, 0, ctx->stream>>>(f, a, da, b, db, out);
  }
  return check_launch(ctx, name);
}

static int poly_eval_device(ronk_ctx* ctx, u64 p, const u64* c, size_t d, const u64* xs, size_t m, u64* out) {
  if (!ctx || (m && (!xs || !out)) || (d && !c)) return set_err(ctx, RONK_EINVAL, "null argument");
  RONK_TRY(validate_modulus(ctx, p));
  if (m == 0) return RONK_OK;
  if (m > 0x7FFFFFFFULL) return set_err(ctx, RONK_EUNSUPPORTED, "too many points");
  if (p == GL_P) {
    GoldilocksField f;
    LaunchScope ls(ctx, "poly_eval");
    poly_eval_kernel<GoldilocksField><<<(u32)m, 256, 0, ctx->stream>>>(f, c, d, xs, out);
  } else {
    MontField f;
    RONK_TRY(make_mont_field(ctx, p, 0, false, &f));
    LaunchScope ls(ctx, "poly_eval");
    poly_eval_kernel<MontField><<<(u32)m, 256, 0, ctx->stream>>>(f, c, d, xs, out);
  }
  return check_launch(ctx, "poly_eval_kernel");
}

// nodes[i] = ω_n^i (plain residues)
static int roots_table(ronk_ctx* ctx, u64 p, u64 g, u64 n, u64* nodes) {
  u64 w;
  if (ronk_root_of_unity(p, g, n, (uint64_t*)&w) != RONK_OK)
    return set_err(ctx, RONK_EINVAL, "n must divide p - 1 (no primitive n-th root of unity)");
  if (n > 0x7FFFFFFFULL) return set_err(ctx, RONK_EUNSUPPORTED, "n too large");
  // plain residues: use the Goldilocks policy's identity to_tw for GL, and for Mont build then
  // convert back would be wasteful — a dedicated kernel keeps it simple.
  if (p == GL_P) {
    GoldilocksField f;
    LaunchScope ls(ctx, "pow_table");
    pow_table_kernel<GoldilocksField><<<((u32)n + 255) / 256, 256, 0, ctx->stream>>>(f, w, 1, nodes, (u32)n);
  } else {
    MontField f;
    RONK_TRY(make_mont_field(ctx, p, 0, false, &f));
    // to_tw(x) = x·R; passing s = R^-1 yields plain residues
    const u64 r1 = (u64)((((unsigned __int128)1) << 64) % p);
    const u64 rinv = h_powmod(r1, p - 2, p);
    LaunchScope ls(ctx, "pow_table");
    pow_table_kernel<MontField><<<((u32)n + 255) / 256, 256, 0, ctx->stream>>>(f, w, rinv, nodes, (u32)n);
  }
  return check_launch(ctx, "pow_table_kernel");
}

static int dft_device(ronk_ctx* ctx, u64 p, u64 g, const u64* in, u64 n, u64* out) {
  if (!ctx || !in || !out) return set_err(ctx, RONK_EINVAL, "null argument");
  RONK_TRY(validate_modulus(ctx, p));
  if (g == 0 || g >= p) return set_err(ctx, RONK_EINVAL, "generator out of range");
  if (n == 0 || (p - 1) % n != 0)
    return set_err(ctx, RONK_EINVAL, "n must divide p - 1 (no primitive n-th root of unity)");
  RONK_TRY(ensure_ws(ctx, &ctx->ws, &ctx->ws_bytes, n * sizeof(u64)));
  RONK_TRY(roots_table(ctx, p, g, n, (u64*)ctx->ws));
  return poly_eval_device(ctx, p, in, n, (const u64*)ctx->ws, n, out);  // X[i] = a(ω^i)
}

}  // namespace ronk

using namespace ronk;

extern "C" {

int ronk_poly_mul_u64(ronk_ctx* ctx, uint64_t p, uint64_t g, const uint64_t* a, size_t da, const uint64_t* b,
                      size_t db, uint64_t* c) {
  ronk::DeviceGuard _dg(ctx);
  return poly_mul_device(ctx, p, g, (const u64*)a, da, (const u64*)b, db, (u64*)c);
}

int ronk_poly_add_u64(ronk_ctx* ctx, uint64_t p, const uint64_t* a, size_t da, const uint64_t* b, size_t db,
                      uint64_t* out) {
  ronk::DeviceGuard _dg(ctx);
  return poly_addsub<false>(ctx, p, (const u64*)a, da, (const u64*)b, db, (u64*)out);
}
int ronk_poly_sub_u64(ronk_ctx* ctx, uint64_t p, const uint64_t* a, size_t da, const uint64_t* b, size_t db,
                      uint64_t* out) {
  ronk::DeviceGuard _dg(ctx);
  return poly_addsub<true>(ctx, p, (const u64*)a, da, (const u64*)b, db, (u64*)out);
}

int ronk_poly_eval_u64(ronk_ctx* ctx, uint64_t p, const uint64_t* coeffs, size_t d, const uint64_t* xs, size_t m,
                       uint64_t* out) {
  ronk::DeviceGuard _dg(ctx);
  return poly_eval_device(ctx, p, (const u64*)coeffs, d, (const u64*)xs, m, (u64*)out);
}

int ronk_dft_u64(ronk_ctx* ctx, uint64_t p, uint64_t g, const uint64_t* in, uint64_t n, uint64_t* out) {
  ronk::DeviceGuard _dg(ctx);
  return dft_device(ctx, p, g, (const u64*)in, n, (u64*)out);
}

int ronk_poly_div_linear_u64(ronk_ctx* ctx, uint64_t p, const uint64_t* a, size_t d, uint64_t b0, uint64_t b1,
                             uint64_t* q, uint64_t* rem) {
  ronk::DeviceGuard _dg(ctx);
  return div_linear_device(ctx, p, (const u64*)a, d, b0, b1, (u64*)q, (u64*)rem);
}

// ---- host-pointer variants ---------------------------------------------------------------------
static int up(ronk_ctx* ctx, u64** d, const void* h, size_t n) {
  RONK_CUDA(ctx, cudaMalloc((void**)d, (n ? n : 1) * sizeof(u64)));
  if (n) RONK_CUDA(ctx, cudaMemcpyAsync(*d, h, n * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
  return RONK_OK;
}
static int down(ronk_ctx* ctx, void* h, const u64* d, size_t n) {
  if (n) RONK_CUDA(ctx, cudaMemcpyAsync(h, d, n * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
  RONK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return RONK_OK;
}
struct DevBuf {  // frees on scope exit
  u64* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
};

int ronk_poly_mul_u64_host(ronk_ctx* ctx, uint64_t p, uint64_t g, const uint64_t* a, size_t da, const uint64_t* b,
                           size_t db, uint64_t* c) {
  ronk::DeviceGuard _dg(ctx);
  if (!ctx || !a || !b || !c) return set_err(ctx, RONK_EINVAL, "null argument");
  if (da == 0 || db == 0) return set_err(ctx, RONK_EINVAL, "empty polynomial (D + D2 - 1 underflows)");
  DevBuf A, B, C;
  RONK_TRY(up(ctx, &A.p, a, da));
  RONK_TRY(up(ctx, &B.p, b, db));
  RONK_CUDA(ctx, cudaMalloc((void**)&C.p, (da + db - 1) * sizeof(u64)));
  RONK_TRY(poly_mul_device(ctx, p, g, A.p, da, B.p, db, C.p));
  return down(ctx, c, C.p, da + db - 1);
}

int ronk_poly_eval_u64_host(ronk_ctx* ctx, uint64_t p, const uint64_t* coeffs, size_t d, const uint64_t* xs, size_t m,
                            uint64_t* out) {
  ronk::DeviceGuard _dg(ctx);
  if (!ctx || (m && (!xs || !out)) || (d && !coeffs)) return set_err(ctx, RONK_EINVAL, "null argument");
  DevBuf C, X, O;
  RONK_TRY(up(ctx, &C.p, coeffs, d));
  RONK_TRY(up(ctx, &X.p, xs, m));
  RONK_CUDA(ctx, cudaMalloc((void**)&O.p, (m ? m : 1) * sizeof(u64)));
  RONK_TRY(poly_eval_device(ctx, p, C.p, d, X.p, m, O.p));
  return down(ctx, out, O.p, m);
}

int ronk_dft_u64_host(ronk_ctx* ctx, uint64_t p, uint64_t g, const uint64_t* in, uint64_t n, uint64_t* out) {
  ronk::DeviceGuard _dg(ctx);
  if (!ctx || !in || !out) return set_err(ctx, RONK_EINVAL, "null argument");
  DevBuf I, O;
  RONK_TRY(up(ctx, &I.p, in, n));
  RONK_CUDA(ctx, cudaMalloc((void**)&O.p, (n ? n : 1) * sizeof(u64)));
  RONK_TRY(dft_device(ctx, p, g, I.p, n, O.p));
  return down(ctx, out, O.p, n);
}

int ronk_poly_lagrange_eval_u64_host(ronk_ctx* ctx, uint64_t p, uint64_t g, const uint64_t* coeffs, size_t n,
                                     uint64_t x, uint64_t* out) {
  ronk::DeviceGuard _dg(ctx);
  if (!ctx || !coeffs || !out) return set_err(ctx, RONK_EINVAL, "null argument");
  RONK_TRY(validate_modulus(ctx, p));
  if (g == 0 || g >= p || x >= p) return set_err(ctx, RONK_EINVAL, "argument out of range");
  if (n == 0 || (p - 1) % n != 0)
    return set_err(ctx, RONK_EINVAL, "n must divide p - 1 (Lagrange::new asserts)");  // mod.rs:361
  if (n > (1u << 20)) return set_err(ctx, RONK_EUNSUPPORTED, "n too large for the O(n²) barycentric form");
  DevBuf C, N, O;
  RONK_TRY(up(ctx, &C.p, coeffs, n));
  RONK_CUDA(ctx, cudaMalloc((void**)&N.p, n * sizeof(u64)));
  RONK_CUDA(ctx, cudaMalloc((void**)&O.p, sizeof(u64)));
  RONK_TRY(roots_table(ctx, p, g, n, N.p));
  RONK_CUDA(ctx, cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), ctx->stream));
  if (p == GL_P) {
    GoldilocksField f;
    LaunchScope ls(ctx, "lagrange_eval");
    lagrange_eval_kernel<GoldilocksField><<<1, 256, 0, ctx->stream>>>(f, C.p, N.p, (u32)n, x, O.p, ctx->d_flag);
  } else {
    MontField f;
    RONK_TRY(make_mont_field(ctx, p, 0, false, &f));
    LaunchScope ls(ctx, "lagrange_eval");
    lagrange_eval_kernel<MontField><<<1, 256, 0, ctx->stream>>>(f, C.p, N.p, (u32)n, x, O.p, ctx->d_flag);
  }
  RONK_TRY(check_launch(ctx, "lagrange_eval_kernel"));
  RONK_CUDA(ctx, cudaMemcpyAsync(ctx->h_flag, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  RONK_TRY(down(ctx, out, O.p, 1));  // synchronises the stream
  if (*ctx->h_flag)  // mod.rs:386-393: F::ONE.div(x_j - x_m) panics when two nodes coincide (g not of order n)
    return set_err(ctx, RONK_EINVAL, "Lagrange evaluate: repeated node (the reference divides by zero)");
  return RONK_OK;
}

int ronk_poly_interpolate_u64_host(ronk_ctx* ctx, uint64_t p, const uint64_t* xs, const uint64_t* ys, size_t k,
                                   uint64_t* out) {
  ronk::DeviceGuard _dg(ctx);
  if (!ctx || (k && (!xs || !ys || !out))) return set_err(ctx, RONK_EINVAL, "null argument");
  RONK_TRY(validate_modulus(ctx, p));
  if (k == 0) return RONK_OK;
  if (k > 8192) return set_err(ctx, RONK_EUNSUPPORTED, "more than 8192 nodes (O(K²) interpolation)");
  for (size_t i = 0; i < k; i++)
    if (xs[i] >= p || ys[i] >= p) return set_err(ctx, RONK_EINVAL, "non-canonical residue");
  DevBuf X, Y, O;
  RONK_TRY(up(ctx, &X.p, xs, k));
  RONK_TRY(up(ctx, &Y.p, ys, k));
  RONK_CUDA(ctx, cudaMalloc((void**)&O.p, k * sizeof(u64)));
  if (p == GL_P) {
    GoldilocksField f;
    RONK_TRY(interp_with_field(ctx, f, X.p, Y.p, (u32)k, O.p));
  } else {
    MontField f;
    RONK_TRY(make_mont_field(ctx, p, 0, false, &f));
    RONK_TRY(interp_with_field(ctx, f, X.p, Y.p, (u32)k, O.p));
  }
  RONK_CUDA(ctx, cudaMemcpyAsync(ctx->h_flag, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  RONK_TRY(down(ctx, out, O.p, k));
  if (*ctx->h_flag) return set_err(ctx, RONK_EINVAL, "interpolation: repeated x coordinate (the reference divides by zero)");
  return RONK_OK;
}

int ronk_poly_divrem_u64_host(ronk_ctx* ctx, uint64_t p, const uint64_t* a, size_t da, const uint64_t* b, size_t db,
                              uint64_t* q, uint64_t* r) {
  ronk::DeviceGuard _dg(ctx);
  if (!ctx || (da && (!a || !q || !r)) || (db && !b)) return set_err(ctx, RONK_EINVAL, "null argument");
  RONK_TRY(validate_modulus(ctx, p));
  if (da > 0x7FFFFFF0ULL || db > 0x7FFFFFF0ULL) return set_err(ctx, RONK_EUNSUPPORTED, "polynomial too long");
  if (da == 0) return RONK_OK;
  DevBuf A, B, Qd, Rd;
  RONK_TRY(up(ctx, &A.p, a, da));
  RONK_CUDA(ctx, cudaMalloc((void**)&Qd.p, da * sizeof(u64)));
  RONK_CUDA(ctx, cudaMalloc((void**)&Rd.p, da * sizeof(u64)));
  if (db == 2 && b[1] != 0 && b[1] < p && b[0] < p) {
    // linear divisor (the kzg::open case): device-wide scan instead of D sequential steps; the
    // remainder is the constant a(z), zero-padded to da terms like the reference's array
    RONK_CUDA(ctx, cudaMemsetAsync(Rd.p, 0, da * sizeof(u64), ctx->stream));
    RONK_TRY(div_linear_device(ctx, p, A.p, da, b[0], b[1], Qd.p, Rd.p));
    RONK_CUDA(ctx, cudaMemcpyAsync(q, Qd.p, da * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    return down(ctx, r, Rd.p, da);
  }
  RONK_TRY(up(ctx, &B.p, b, db));
  RONK_CUDA(ctx, cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), ctx->stream));
  if (p == GL_P) {
    GoldilocksField f;
    LaunchScope ls(ctx, "poly_divrem");
    poly_divrem_kernel<GoldilocksField><<<1, 256, 0, ctx->stream>>>(f, A.p, (u32)da, B.p, (u32)db, Qd.p, Rd.p, ctx->d_flag);
  } else {
    MontField f;
    RONK_TRY(make_mont_field(ctx, p, 0, false, &f));
    LaunchScope ls(ctx, "poly_divrem");
    poly_divrem_kernel<MontField><<<1, 256, 0, ctx->stream>>>(f, A.p, (u32)da, B.p, (u32)db, Qd.p, Rd.p, ctx->d_flag);
  }
  RONK_TRY(check_launch(ctx, "poly_divrem_kernel"));
  RONK_CUDA(ctx, cudaMemcpyAsync(ctx->h_flag, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  RONK_CUDA(ctx, cudaMemcpyAsync(q, Qd.p, da * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
  RONK_TRY(down(ctx, r, Rd.p, da));
  if (*ctx->h_flag) return set_err(ctx, RONK_EINVAL, "polynomial division: the reference would panic on this divisor");
  return RONK_OK;
}

}  // extern "C"
