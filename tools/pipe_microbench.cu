// pipe_microbench.cu — measures per-SM issue throughput of the integer / FP64 pipes on B200 to
// guide the NTT butterfly design (which pipe mix a 64-bit modular butterfly should aim for).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_microbench pipe_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048
#define REP8(x) x x x x x x x x

template <int MODE>
__global__ void k(unsigned long long* out, unsigned long long* cyc) {
  unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  unsigned long long w0 = a0, w1 = a1, w2 = a2, w3 = a3, w4 = a4, w5 = a5, w6 = a6, w7 = a7;
  double d0 = a0, d1 = a1, d2 = a2, d3 = a3, d4 = a4, d5 = a5, d6 = a6, d7 = a7;
  unsigned m = blockIdx.x * 2654435761u + 12345u;
  double dm = 1.0000001;
  unsigned long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < ITERS; i++) {
    if (MODE == 0) {  // 32-bit add (ALU pipe: IADD3)
      asm volatile("add.u32 %0,%0,%8; add.u32 %1,%1,%8; add.u32 %2,%2,%8; add.u32 %3,%3,%8; add.u32 %4,%4,%8; add.u32 %5,%5,%8; add.u32 %6,%6,%8; add.u32 %7,%7,%8;"
                   : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7) : "r"(m));
    } else if (MODE == 1) {  // mad.wide.u32 (FMA pipe: IMAD.WIDE)
      asm volatile("{.reg .u32 l0,h0; mov.b64 {l0,h0},%0; mad.wide.u32 %0,l0,%8,%0;} {.reg .u32 l1,h1; mov.b64 {l1,h1},%1; mad.wide.u32 %1,l1,%8,%1;} {.reg .u32 l2,h2; mov.b64 {l2,h2},%2; mad.wide.u32 %2,l2,%8,%2;} {.reg .u32 l3,h3; mov.b64 {l3,h3},%3; mad.wide.u32 %3,l3,%8,%3;} {.reg .u32 l4,h4; mov.b64 {l4,h4},%4; mad.wide.u32 %4,l4,%8,%4;} {.reg .u32 l5,h5; mov.b64 {l5,h5},%5; mad.wide.u32 %5,l5,%8,%5;} {.reg .u32 l6,h6; mov.b64 {l6,h6},%6; mad.wide.u32 %6,l6,%8,%6;} {.reg .u32 l7,h7; mov.b64 {l7,h7},%7; mad.wide.u32 %7,l7,%8,%7;}"
                   : "+l"(w0), "+l"(w1), "+l"(w2), "+l"(w3), "+l"(w4), "+l"(w5), "+l"(w6), "+l"(w7) : "r"(m));
    } else if (MODE == 2) {  // fma.rn.f64 (FP64 pipe: DFMA)
      asm volatile("fma.rn.f64 %0,%0,%8,%8; fma.rn.f64 %1,%1,%8,%8; fma.rn.f64 %2,%2,%8,%8; fma.rn.f64 %3,%3,%8,%8; fma.rn.f64 %4,%4,%8,%8; fma.rn.f64 %5,%5,%8,%8; fma.rn.f64 %6,%6,%8,%8; fma.rn.f64 %7,%7,%8,%8;"
                   : "+d"(d0), "+d"(d1), "+d"(d2), "+d"(d3), "+d"(d4), "+d"(d5), "+d"(d6), "+d"(d7) : "d"(dm));
    } else if (MODE == 3) {  // 4 IADD + 4 IMAD.WIDE interleaved
      asm volatile("add.u32 %0,%0,%8; {.reg .u32 l4,h4; mov.b64 {l4,h4},%4; mad.wide.u32 %4,l4,%8,%4;} add.u32 %1,%1,%8; {.reg .u32 l5,h5; mov.b64 {l5,h5},%5; mad.wide.u32 %5,l5,%8,%5;} add.u32 %2,%2,%8; {.reg .u32 l6,h6; mov.b64 {l6,h6},%6; mad.wide.u32 %6,l6,%8,%6;} add.u32 %3,%3,%8; {.reg .u32 l7,h7; mov.b64 {l7,h7},%7; mad.wide.u32 %7,l7,%8,%7;}"
                   : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+l"(w4), "+l"(w5), "+l"(w6), "+l"(w7) : "r"(m));
    } else if (MODE == 4) {  // 4 IADD + 4 DFMA interleaved
      asm volatile("add.u32 %0,%0,%8; fma.rn.f64 %4,%4,%9,%9; add.u32 %1,%1,%8; fma.rn.f64 %5,%5,%9,%9; add.u32 %2,%2,%8; fma.rn.f64 %6,%6,%9,%9; add.u32 %3,%3,%8; fma.rn.f64 %7,%7,%9,%9;"
                   : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+d"(d4), "+d"(d5), "+d"(d6), "+d"(d7) : "r"(m), "d"(dm));
    } else if (MODE == 5) {  // 64-bit add with carry chain: add.cc + addc (2 ALU instrs per 64-bit add)
      asm volatile("add.cc.u32 %0,%0,%8; addc.u32 %1,%1,%8; add.cc.u32 %2,%2,%8; addc.u32 %3,%3,%8; add.cc.u32 %4,%4,%8; addc.u32 %5,%5,%8; add.cc.u32 %6,%6,%8; addc.u32 %7,%7,%8;"
                   : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7) : "r"(m));
    } else if (MODE == 6) {  // mad.lo.u32 (FMA pipe: IMAD 32-bit)
      asm volatile("mad.lo.u32 %0,%0,%8,%8; mad.lo.u32 %1,%1,%8,%8; mad.lo.u32 %2,%2,%8,%8; mad.lo.u32 %3,%3,%8,%8; mad.lo.u32 %4,%4,%8,%8; mad.lo.u32 %5,%5,%8,%8; mad.lo.u32 %6,%6,%8,%8; mad.lo.u32 %7,%7,%8,%8;"
                   : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7) : "r"(m));
    } else if (MODE == 7) {  // 3-way: 3 IADD + 3 IMAD.WIDE + 2 DFMA
      asm volatile("add.u32 %0,%0,%8; {.reg .u32 l3,h3; mov.b64 {l3,h3},%3; mad.wide.u32 %3,l3,%8,%3;} fma.rn.f64 %6,%6,%9,%9; add.u32 %1,%1,%8; {.reg .u32 l4,h4; mov.b64 {l4,h4},%4; mad.wide.u32 %4,l4,%8,%4;} fma.rn.f64 %7,%7,%9,%9; add.u32 %2,%2,%8; {.reg .u32 l5,h5; mov.b64 {l5,h5},%5; mad.wide.u32 %5,l5,%8,%5;}"
                   : "+r"(a0), "+r"(a1), "+r"(a2), "+l"(w3), "+l"(w4), "+l"(w5), "+d"(d6), "+d"(d7) : "r"(m), "d"(dm));
    } else if (MODE == 8) {  // lop3 / shf (ALU pipe logic + funnel shift)
      asm volatile("shf.l.wrap.b32 %0,%0,%1,%8; shf.l.wrap.b32 %1,%1,%2,%8; shf.l.wrap.b32 %2,%2,%3,%8; shf.l.wrap.b32 %3,%3,%4,%8; shf.l.wrap.b32 %4,%4,%5,%8; shf.l.wrap.b32 %5,%5,%6,%8; shf.l.wrap.b32 %6,%6,%7,%8; shf.l.wrap.b32 %7,%7,%0,%8;"
                   : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7) : "r"(m & 31));
    }
  }
  unsigned long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] =
      a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + w0 + w1 + w2 + w3 + w4 + w5 + w6 + w7 +
      (unsigned long long)(d0 + d1 + d2 + d3 + d4 + d5 + d6 + d7);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int sms, int blocks_per_sm, int threads, unsigned long long* out, unsigned long long* cyc) {
  int blocks = sms * blocks_per_sm;
  k<MODE><<<blocks, threads>>>(out, cyc);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(out, cyc);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  unsigned long long* h = new unsigned long long[blocks];
  cudaMemcpy(h, cyc, blocks * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
  double ops_per_sm = (double)blocks_per_sm * threads * ITERS * 8;
  printf("%-28s blocks/SM=%d thr=%d  cycles=%.0f  thread-instr/clk/SM=%.1f  ms=%.3f  Ginstr/s=%.0f\n", name, blocks_per_sm,
         threads, avg, ops_per_sm / avg, ms, (double)blocks * threads * ITERS * 8 / ms / 1e6);
  delete[] h;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  printf("device %s, %d SMs\n", p.name, sms);
  unsigned long long *out, *cyc;
  cudaMalloc(&out, (size_t)sms * 8 * 1024 * 8); cudaMalloc(&cyc, sms * 8 * 8);
  for (int thr : {256, 1024}) {
    int bps = 2048 / thr;
    run<0>("iadd (ALU)", sms, bps, thr, out, cyc);
    run<5>("add.cc/addc (ALU carry)", sms, bps, thr, out, cyc);
    run<8>("shf (ALU)", sms, bps, thr, out, cyc);
    run<6>("mad.lo.u32 (IMAD)", sms, bps, thr, out, cyc);
    run<1>("mad.wide.u32 (IMAD.WIDE)", sms, bps, thr, out, cyc);
    run<2>("fma.f64 (DFMA)", sms, bps, thr, out, cyc);
    run<3>("iadd+imad.wide 1:1", sms, bps, thr, out, cyc);
    run<4>("iadd+dfma 1:1", sms, bps, thr, out, cyc);
    run<7>("iadd+imad.wide+dfma 3:3:2", sms, bps, thr, out, cyc);
  }
  return 0;
}
