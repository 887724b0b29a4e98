// Instruction-throughput probes for the integer forms a 256-bit field multiplication can be built from (sm_100a).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
template <int MODE>
__global__ void k(uint32_t *out, const uint32_t *in) {
  uint32_t a[8], b[8], c[8];
  uint64_t w[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { a[i] = in[threadIdx.x + 32 * i]; b[i] = in[threadIdx.x + 32 * (i + 8)]; c[i] = in[threadIdx.x + 32 * (i + 16)]; w[i] = ((uint64_t)a[i] << 32) | b[i]; }
  for (int it = 0; it < ITERS; it++) {
    if (MODE == 0) {  // mad.wide, all operands distinct
#pragma unroll
      for (int i = 0; i < 8; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b[i]));
    }
    if (MODE == 1) {  // mad.wide, shared multiplier
#pragma unroll
      for (int i = 0; i < 8; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b[0]));
    }
    if (MODE == 2) {  // 32-bit mad.lo, distinct
#pragma unroll
      for (int i = 0; i < 8; i++) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(c[i]) : "r"(a[i]), "r"(b[i]));
    }
    if (MODE == 3) {  // fused lo/hi carry chains: two independent chains of 4 lanes
      asm volatile("mad.lo.cc.u32 %0, %8, %12, %0;\n\tmadc.hi.cc.u32 %1, %8, %12, %1;\n\tmadc.lo.cc.u32 %2, %9, %12, %2;\n\tmadc.hi.cc.u32 %3, %9, %12, %3;\n\t"
                   "madc.lo.cc.u32 %4, %10, %12, %4;\n\tmadc.hi.cc.u32 %5, %10, %12, %5;\n\tmadc.lo.cc.u32 %6, %11, %12, %6;\n\tmadc.hi.u32 %7, %11, %12, %7;"
                   : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]), "+r"(c[4]), "+r"(c[5]), "+r"(c[6]), "+r"(c[7]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[it & 7]));
      asm volatile("mad.lo.cc.u32 %0, %8, %12, %0;\n\tmadc.hi.cc.u32 %1, %8, %12, %1;\n\tmadc.lo.cc.u32 %2, %9, %12, %2;\n\tmadc.hi.cc.u32 %3, %9, %12, %3;\n\t"
                   "madc.lo.cc.u32 %4, %10, %12, %4;\n\tmadc.hi.cc.u32 %5, %10, %12, %5;\n\tmadc.lo.cc.u32 %6, %11, %12, %6;\n\tmadc.hi.u32 %7, %11, %12, %7;"
                   : "+r"(a[0 + 4]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]));
    }
    if (MODE == 4) {  // mul.wide (no addend) + xor consumer on the alu pipe
#pragma unroll
      for (int i = 0; i < 8; i++) { uint64_t d; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(d) : "r"(a[i]), "r"(b[i])); c[i] ^= (uint32_t)d ^ (uint32_t)(d >> 32); a[i] += 1; }
    }
    if (MODE == 5) {  // 3-input adds
#pragma unroll
      for (int i = 0; i < 8; i++) c[i] = c[i] + a[i] + b[i];
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] ^= c[(i + 1) & 7];
    }
    if (MODE == 6) {  // mad.hi distinct
#pragma unroll
      for (int i = 0; i < 8; i++) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(c[i]) : "r"(a[i]), "r"(b[i]));
    }
    if (MODE == 7) {  // mad.wide row style: shared a, distinct b, distinct acc
#pragma unroll
      for (int i = 0; i < 8; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[it & 7]), "r"(b[i]));
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) r ^= a[i] ^ b[i] ^ c[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE>
void run(const char *name, double ops_per_iter, uint32_t *in) {
  int blocks = 148 * 8, threads = 256;
  uint32_t *out; cudaMalloc(&out, blocks * threads * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(out, in);
  cudaEventRecord(e0);
  for (int i = 0; i < 5; i++) k<MODE><<<blocks, threads>>>(out, in);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)blocks * threads * ITERS * ops_per_iter * 5;
  double per_smsp_cycle = ops / 32.0 / (ms * 1e-3 * 1.965e9) / (148 * 4);
  printf("{\"probe\": \"%s\", \"Tops_per_s\": %.3f, \"warp_instr_per_clk_per_smsp\": %.3f}\n", name, ops / (ms * 1e-3) / 1e12, per_smsp_cycle);
  cudaFree(out);
}
int main() {
  uint32_t *in; cudaMalloc(&in, 4 * 32 * 24); cudaMemset(in, 0x5a, 4 * 32 * 24);
  run<0>("mad.wide.u32 distinct a,b,acc", 8, in);
  run<1>("mad.wide.u32 shared b", 8, in);
  run<2>("mad.lo.u32 distinct", 8, in);
  run<3>("fused lo/hi carry chain (wide.X), per wide", 8, in);
  run<4>("mul.wide + lop3 consumer, per wide", 8, in);
  run<5>("iadd3 (+lop3), per iadd3", 8, in);
  run<6>("mad.hi.u32 distinct", 8, in);
  run<7>("mad.wide.u32 shared a (row style)", 8, in);
  return 0;
}
