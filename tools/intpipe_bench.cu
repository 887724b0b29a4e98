// Integer-pipe probes for sm_100a: issue throughput of the multiply-add forms a 256-bit field multiplication is built from.
// Every multiply has a loop-variant operand (the previous result), so the compiler cannot hoist it out of the loop -- the
// first version of this probe fed loop-invariant operands to mad.wide and measured the additions that were left.
// Results: profiles/r01_intpipe_peak.jsonl; they are the "integer roofline" quoted in DESIGN.md section 4.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

template <int MODE>
__global__ void k(uint32_t *out, uint32_t seed) {
  uint32_t a[ILP], b[ILP];
  uint64_t w[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) { a[i] = threadIdx.x * (i + 3) + seed; b[i] = (seed * 2654435761u + i + threadIdx.x * 7919u) | 1; w[i] = ((uint64_t)a[i] << 32) | b[i]; }
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      if (MODE == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(a[(i + 1) % ILP]));           // IMAD
      if (MODE == 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(a[(i + 1) % ILP]));           // IMAD.HI
      if (MODE == 2) { uint32_t lo = (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32); asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(lo), "r"(b[i])); }  // IMAD.WIDE, 64-bit addend
      if (MODE == 3) { uint32_t lo = (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32); asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(lo), "r"(b[i])); }       // IMAD.WIDE, no addend
      if (MODE == 4) asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(a[(i + 1) % ILP]));
    }
    if (MODE == 5) {  // fused lo/hi pairs with carry predicates (IMAD.WIDE.U32.X): two independent chains of four lanes, variant multiplier
      uint32_t y = a[0] | 1;
      asm volatile("mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\tmadc.lo.cc.u32 %2, %8, %9, %2;\n\tmadc.hi.cc.u32 %3, %8, %9, %3;\n\t"
                   "madc.lo.cc.u32 %4, %8, %9, %4;\n\tmadc.hi.cc.u32 %5, %8, %9, %5;\n\tmadc.lo.cc.u32 %6, %8, %9, %6;\n\tmadc.hi.u32 %7, %8, %9, %7;"
                   : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]) : "r"(b[it & 7]), "r"(y));
      uint32_t z = b[1] | 1;
      asm volatile("mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\tmadc.lo.cc.u32 %2, %8, %9, %2;\n\tmadc.hi.cc.u32 %3, %8, %9, %3;\n\t"
                   "madc.lo.cc.u32 %4, %8, %9, %4;\n\tmadc.hi.cc.u32 %5, %8, %9, %5;\n\tmadc.lo.cc.u32 %6, %8, %9, %6;\n\tmadc.hi.u32 %7, %8, %9, %7;"
                   : "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]) : "r"(a[it & 7]), "r"(z));
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) r ^= a[i] ^ b[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char *name, double opsPerIter) {
  int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  int blocks = nsm * 8, threads = 256;
  uint32_t *out; cudaMalloc(&out, blocks * threads * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; i++) k<MODE><<<blocks, threads>>>(out, 12345);
  cudaEventRecord(e0);
  const int reps = 10;
  for (int i = 0; i < reps; i++) k<MODE><<<blocks, threads>>>(out, 12345 + i);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double ops = (double)blocks * threads * ITERS * opsPerIter * reps;
  double per_smsp = ops / 32.0 / (ms * 1e-3 * khz * 1e3) / (nsm * 4.0);
  printf("{\"op\": \"%s\", \"Tops_per_s\": %.3f, \"issue_cycles_per_warp_instruction_per_SM_subpartition\": %.2f, \"ms\": %.3f, \"sms\": %d}\n", name, ops / (ms * 1e-3) / 1e12,
         1.0 / per_smsp, ms / reps, nsm);
  cudaFree(out);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, p.multiProcessorCount, p.clockRate);
  run<0>("mad.lo.u32 (IMAD)", ILP);
  run<1>("mad.hi.u32 (IMAD.HI)", ILP);
  run<2>("mad.wide.u32 with 64-bit addend (IMAD.WIDE)", ILP);
  run<3>("mul.wide.u32 (IMAD.WIDE, RZ addend)", ILP);
  run<4>("add.u32 x2 (IADD3)", 2 * ILP);
  run<5>("mad.lo.cc / madc.hi.cc pairs (IMAD.WIDE.U32.X carry chains)", 8);
  return 0;
}
