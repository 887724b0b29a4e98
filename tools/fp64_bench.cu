// FP64-pipe probes for sm_100a: issue rate of DFMA / DADD, and whether they overlap with the IMAD.WIDE carry chains of the
// integer field multiplication (same warp, and different warps of one SM sub-partition).  Decides whether a second field
// multiplier on the FP64 pipe can add throughput.  Results: profiles/r02_fp64_pipe_probe.jsonl
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

// MODE 0: fma.rn.f64   1: fma.rz.f64   2: add.f64   3: IMAD.WIDE carry chains only
//      4: same warp: 8 DFMA + 16 IMAD.WIDE per iteration      5: warps alternate (even: DFMA loop, odd: IMAD.WIDE loop)
//      6: same warp: 8 DFMA + 16 64-bit integer adds          7: 64-bit integer adds only
template <int MODE>
__global__ void k(uint32_t *out, uint32_t seed) {
  double x[ILP], y[ILP];
  uint32_t a[8], b[8];
  uint64_t w[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) { x[i] = 1.0 + (threadIdx.x + i + seed) * 1e-9; y[i] = 1.0 - (threadIdx.x * 3 + i) * 1e-10; w[i] = threadIdx.x * 77 + i + seed; }
#pragma unroll
  for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * (i + 3) + seed; b[i] = (seed * 2654435761u + i + threadIdx.x * 7919u) | 1; }
  const int warp = threadIdx.x >> 5;
  const bool do_f = MODE == 0 || MODE == 1 || MODE == 2 || MODE == 4 || MODE == 6 || (MODE == 5 && (warp & 1) == 0);
  const bool do_i = MODE == 3 || MODE == 4 || (MODE == 5 && (warp & 1) == 1);
  const bool do_a = MODE == 6 || MODE == 7;
  for (int it = 0; it < ITERS; it++) {
    if (do_f) {
#pragma unroll
      for (int i = 0; i < ILP; i++) {
        if (MODE == 1) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(y[i]), "d"(y[(i + 1) % ILP]));
        else if (MODE == 2) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x[i]) : "d"(y[i]));
        else asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(y[i]), "d"(y[(i + 1) % ILP]));
      }
    }
    if (do_i) {
      uint32_t yy = a[0] | 1;
      asm volatile("mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\tmadc.lo.cc.u32 %2, %8, %9, %2;\n\tmadc.hi.cc.u32 %3, %8, %9, %3;\n\t"
                   "madc.lo.cc.u32 %4, %8, %9, %4;\n\tmadc.hi.cc.u32 %5, %8, %9, %5;\n\tmadc.lo.cc.u32 %6, %8, %9, %6;\n\tmadc.hi.u32 %7, %8, %9, %7;"
                   : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]) : "r"(b[it & 7]), "r"(yy));
      uint32_t z = b[1] | 1;
      asm volatile("mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\tmadc.lo.cc.u32 %2, %8, %9, %2;\n\tmadc.hi.cc.u32 %3, %8, %9, %3;\n\t"
                   "madc.lo.cc.u32 %4, %8, %9, %4;\n\tmadc.hi.cc.u32 %5, %8, %9, %5;\n\tmadc.lo.cc.u32 %6, %8, %9, %6;\n\tmadc.hi.u32 %7, %8, %9, %7;"
                   : "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]) : "r"(a[it & 7]), "r"(z));
    }
    if (do_a) {
#pragma unroll
      for (int i = 0; i < ILP; i++) { w[i] += w[(i + 1) % ILP] ^ (uint64_t)it; w[i] += (uint64_t)__double_as_longlong(x[i]); }
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) r ^= (uint32_t)__double_as_longlong(x[i]) ^ (uint32_t)(__double_as_longlong(x[i]) >> 32) ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
#pragma unroll
  for (int i = 0; i < 8; i++) r ^= a[i] ^ b[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char *name, double fOps, double iOps, double aOps) {
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
  const double frac = MODE == 5 ? 0.5 : 1.0;  // alternating warps: each class runs on half the warps
  const double warp_iters = (double)blocks * threads / 32.0 * ITERS * reps * frac;
  const double cyc = ms * 1e-3 * khz * 1e3 * nsm * 4.0;  // SM sub-partition cycles available
  printf("{\"mode\": \"%s\", \"ms\": %.3f, \"dfma_per_smsp_cycle\": %.4f, \"imad_wide_per_smsp_cycle\": %.4f, \"iadd64_per_smsp_cycle\": %.4f, "
         "\"Tdfma_per_s\": %.3f, \"Timad_wide_per_s\": %.3f}\n",
         name, ms / reps, warp_iters * fOps / cyc, warp_iters * iOps / cyc, warp_iters * aOps / cyc, warp_iters * fOps * 32 / (ms * 1e-3) / 1e12,
         warp_iters * iOps * 32 / (ms * 1e-3) / 1e12);
  cudaFree(out);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, p.multiProcessorCount, p.clockRate);
  run<0>("fma.rn.f64 only", ILP, 0, 0);
  run<1>("fma.rz.f64 only", ILP, 0, 0);
  run<2>("add.rn.f64 only", ILP, 0, 0);
  run<3>("IMAD.WIDE carry chains only", 0, 16, 0);
  run<4>("same warp: 8 DFMA + 16 IMAD.WIDE", ILP, 16, 0);
  run<5>("alternating warps: DFMA | IMAD.WIDE", ILP, 16, 0);
  run<6>("same warp: 8 DFMA + 16 64-bit adds", ILP, 0, 16);
  run<7>("64-bit integer adds only", 0, 0, 16);
  return 0;
}
