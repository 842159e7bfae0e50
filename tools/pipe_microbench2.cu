// pipe_microbench2.cu — second, SASS-verified issue-rate probe for the B200 integer pipes (round 2).
// Every loop body is inline PTX whose SASS opcode is pinned by construction (carry-out forces IADD3,
// a run-time multiplier forces IMAD, lop3 / shf exist on the ALU pipe only); `cuobjdump -sass` of
// this file was checked for each MODE before the numbers were used in DESIGN.md §3.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_microbench2 pipe_microbench2.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
// 8 independent accumulators a0..a7 (32-bit), w0..w3 (64-bit)
#define A4(op) op(0) op(1) op(2) op(3)
#define A8(op) op(0) op(1) op(2) op(3) op(4) op(5) op(6) op(7)
#define S(x) #x
#define IADDC(i) "add.cc.u32 %" S(i) ",%" S(i) ",%12;"
#define LOP(i) "lop3.b32 %" S(i) ",%" S(i) ",%12,%13,0x96;"
#define SHFO(i) "shf.l.wrap.b32 %" S(i) ",%" S(i) ",%12,%13;"
#define IMAD(i) "mad.lo.u32 %" S(i) ",%" S(i) ",%12,%13;"
#define IHI(i) "mad.hi.u32 %" S(i) ",%" S(i) ",%12,%13;"
template <int MODE>
__global__ void __launch_bounds__(256) k(unsigned long long* out, unsigned long long* cyc, unsigned m, unsigned m2) {
  unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  unsigned long long w0 = a0, w1 = a1, w2 = a2, w3 = a3;
  unsigned long long t0 = clock64();
#define OPS : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7), "+l"(w0), "+l"(w1), "+l"(w2), "+l"(w3) : "r"(m), "r"(m2)
#define WIDE(i) "{.reg .u32 l,h; mov.b64 {l,h},%" S(i) "; mul.wide.u32 %" S(i) ",l,%12;}"
#define CHAIN "add.cc.u32 %0,%0,%12; addc.cc.u32 %1,%1,%13; addc.cc.u32 %2,%2,%12; addc.cc.u32 %3,%3,%13; addc.cc.u32 %4,%4,%12; addc.cc.u32 %5,%5,%13; addc.cc.u32 %6,%6,%12; addc.u32 %7,%7,%13;"
#pragma unroll 1
  for (int i = 0; i < ITERS; i++) {
    if (MODE == 0) asm volatile(CHAIN OPS);                                            // IADD3 + 7 IADD3.X carry chain
    if (MODE == 1) asm volatile(A8(LOP) OPS);                                          // 8 LOP3
    if (MODE == 2) asm volatile(A8(SHFO) OPS);                                         // 8 SHF
    if (MODE == 3) asm volatile(A8(IMAD) OPS);                                         // 8 IMAD
    if (MODE == 4) asm volatile(WIDE(8) WIDE(9) WIDE(10) WIDE(11) OPS);  // 4 IMAD.WIDE
    if (MODE == 5) asm volatile(LOP(0) IMAD(4) LOP(1) IMAD(5) LOP(2) IMAD(6) LOP(3) IMAD(7) OPS);  // 4 ALU : 4 IMAD
    if (MODE == 6) asm volatile(LOP(0) IMAD(4) IMAD(5) LOP(1) IMAD(6) IMAD(7) LOP(2) IMAD(4) IMAD(5) LOP(3) IMAD(6) IMAD(7) OPS);  // 4 ALU : 8 IMAD
    if (MODE == 7) asm volatile(LOP(0) LOP(1) IMAD(4) LOP(2) LOP(3) IMAD(5) LOP(0) LOP(1) IMAD(6) LOP(2) LOP(3) IMAD(7) OPS);  // 8 ALU : 4 IMAD
    if (MODE == 8) asm volatile(LOP(0) WIDE(8) LOP(1) WIDE(9) LOP(2) WIDE(10) LOP(3) WIDE(11) OPS);  // 4 ALU : 4 WIDE
    if (MODE == 9) asm volatile(LOP(0) LOP(1) LOP(2) WIDE(8) LOP(3) LOP(4) LOP(5) WIDE(9) OPS);  // 6 ALU : 2 WIDE
    if (MODE == 10) asm volatile(IMAD(0) IMAD(1) IMAD(2) WIDE(8) IMAD(3) IMAD(4) IMAD(5) WIDE(9) OPS);       // 6 IMAD : 2 WIDE
    if (MODE == 11) asm volatile(LOP(0) LOP(1) IMAD(4) IMAD(5) WIDE(8) LOP(2) LOP(3) IMAD(6) IMAD(7) WIDE(9) OPS);  // 4 ALU : 4 IMAD : 2 WIDE
    if (MODE == 12) asm volatile(LOP(0) IMAD(4) SHFO(1) IMAD(5) LOP(2) IMAD(6) SHFO(3) IMAD(7) OPS);         // 4 ALU(lop/shf) : 4 IMAD
    if (MODE == 14) asm volatile(A8(IHI) OPS);                                         // 8 IMAD.HI
    if (MODE == 15) asm volatile(LOP(0) IHI(4) LOP(1) IHI(5) LOP(2) IHI(6) LOP(3) IHI(7) OPS);         // 4 LOP3 : 4 IMAD.HI
    if (MODE == 16) asm volatile(IMAD(0) IHI(4) IMAD(1) IHI(5) IMAD(2) IHI(6) IMAD(3) IHI(7) OPS);     // 4 IMAD : 4 IMAD.HI
    if (MODE == 13) asm volatile(LOP(0) IMAD(4) IMAD(5) IMAD(6) LOP(1) IMAD(7) IMAD(4) IMAD(5) OPS);     // 2 ALU : 6 IMAD
  }
  unsigned long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + w0 + w1 + w2 + w3;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
static const int NINSTR[] = {8, 8, 8, 8, 8, 8, 12, 12, 8, 8, 8, 10, 8, 8, 8, 8, 8};
template <int MODE>
void run(const char* name, int sms, int blocks_per_sm, unsigned long long* out, unsigned long long* cyc) {
  const int threads = 256, blocks = sms * blocks_per_sm;
  k<MODE><<<blocks, threads>>>(out, cyc, 0x9e3779b9u, 12345u);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int r = 0; r < 4; r++) k<MODE><<<blocks, threads>>>(out, cyc, 0x9e3779b9u, 12345u);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 4;
  unsigned long long* h = new unsigned long long[blocks];
  cudaMemcpy(h, cyc, blocks * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
  const double warp_instr_per_smsp = (double)blocks_per_sm * threads / 32 / 4 * ITERS * NINSTR[MODE];
  const double wps = blocks_per_sm * threads / 128.0;
  printf("%-34s warps/SMSP=%2.0f  SMSP-cycles per loop iteration per warp: clock64 %.2f  wall@1965MHz %.2f  (eff.clock %.0f MHz, nominal IPC %.3f)\n", name,
         wps, avg / ITERS / wps, ms * 1.965e6 / ITERS / wps, avg / ms / 1e3, warp_instr_per_smsp / avg);
  delete[] h;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  printf("device %s, %d SMs\n", p.name, sms);
  unsigned long long *out, *cyc;
  cudaMalloc(&out, (size_t)sms * 8 * 256 * 8); cudaMalloc(&cyc, sms * 8 * 8);
  for (int bps : {2, 8}) {
    run<0>("carry chain IADD3+7 IADD3.X", sms, bps, out, cyc);
    run<1>("8 LOP3", sms, bps, out, cyc);
    run<2>("8 SHF", sms, bps, out, cyc);
    run<3>("8 IMAD", sms, bps, out, cyc);
    run<4>("8 IMAD.WIDE", sms, bps, out, cyc);
    run<5>("4 LOP3 : 4 IMAD", sms, bps, out, cyc);
    run<6>("4 LOP3 : 8 IMAD", sms, bps, out, cyc);
    run<13>("2 LOP3 : 6 IMAD", sms, bps, out, cyc);
    run<7>("8 LOP3 : 4 IMAD", sms, bps, out, cyc);
    run<8>("4 LOP3 : 4 WIDE", sms, bps, out, cyc);
    run<9>("6 LOP3 : 2 WIDE", sms, bps, out, cyc);
    run<10>("6 IMAD : 2 WIDE", sms, bps, out, cyc);
    run<11>("4 LOP3 : 4 IMAD : 2 WIDE", sms, bps, out, cyc);
    run<12>("2 LOP3 + 2 SHF : 4 IMAD", sms, bps, out, cyc);
    run<14>("8 IMAD.HI", sms, bps, out, cyc);
    run<15>("4 LOP3 : 4 IMAD.HI", sms, bps, out, cyc);
    run<16>("4 IMAD : 4 IMAD.HI", sms, bps, out, cyc);
  }
  return 0;
}
