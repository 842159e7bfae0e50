// Micro-kernels behind DESIGN.md §3.2's per-operation SASS budget (compile only; no GPU needed):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -cubin -o ops.cubin ops.cu
//   cuobjdump -sass ops.cubin > ops.sass && python hist.py ops.sass
#include "../../ronkathon_b200/csrc/ntt_kernel.cuh"
using namespace ronk;
extern "C" __global__ void k_net(u64* d) {  // one radix-16 network on 16 elements
  GoldilocksField f; u64 x[16];
  for (int i = 0; i < 16; i++) x[i] = d[threadIdx.x * 16 + i];
  radix_network<4, false>(f, x);
  for (int i = 0; i < 16; i++) d[threadIdx.x * 16 + i] = x[i];
}
extern "C" __global__ void k_mul16(u64* d, const u64* w) {  // 16 general multiplies
  GoldilocksField f; u64 x[16];
  for (int i = 0; i < 16; i++) x[i] = d[threadIdx.x * 16 + i];
  for (int i = 0; i < 16; i++) x[i] = f.mul(x[i], w[threadIdx.x * 16 + i]);
  for (int i = 0; i < 16; i++) d[threadIdx.x * 16 + i] = x[i];
}
extern "C" __global__ void k_addsub8(u64* d) {  // 8 butterflies (add + sub)
  GoldilocksField f; u64 x[16];
  for (int i = 0; i < 16; i++) x[i] = d[threadIdx.x * 16 + i];
  for (int i = 0; i < 8; i++) { u64 a = x[i], b = x[i + 8]; x[i] = f.add(a, b); x[i + 8] = f.sub(a, b); }
  for (int i = 0; i < 16; i++) d[threadIdx.x * 16 + i] = x[i];
}
extern "C" __global__ void k_mul1(u64* d, const u64* w) { GoldilocksField f; d[threadIdx.x] = f.mul(d[threadIdx.x], w[threadIdx.x]); }
template <int S> __global__ void k_sh(u64* d) { GoldilocksField f; d[threadIdx.x] = f.mul_pow2<S>(f.sub(d[threadIdx.x], d[threadIdx.x + 64])); }
template __global__ void k_sh<12>(u64*);
template __global__ void k_sh<48>(u64*);
template __global__ void k_sh<84>(u64*);
template __global__ void k_sh<32>(u64*);
template __global__ void k_sh<64>(u64*);
