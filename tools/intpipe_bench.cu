// Integer-pipe peak microbenchmark for sm_100a: measures issue throughput of the
// 32-bit multiply-add forms a big-integer kernel is built from.  The result is
// the "integer roofline" denominator quoted in DESIGN.md / bench.py.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

template <int MODE>
__global__ void k(uint32_t *out, uint32_t seed) {
  uint32_t a[ILP], b = seed | 1, c = seed * 2654435761u + threadIdx.x;
  uint64_t w[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) { a[i] = threadIdx.x * (i + 3) + seed; w[i] = a[i]; }
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      if (MODE == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
      if (MODE == 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
      if (MODE == 2) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b));
      if (MODE == 3) asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
      if (MODE == 4) {  // carry-chained lo/hi pair (one 32x32->64 product accumulated)
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;"
                     : "+r"(a[i]), "+r"(a[(i + 1) % ILP]) : "r"(b), "r"(c));
      }
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) r ^= a[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
double run(const char *name, int opsPerIter) {
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
  double ops = (double)blocks * threads * ITERS * ILP * opsPerIter * reps;
  double tops = ops / (ms * 1e-3) / 1e12;
  printf("{\"op\": \"%s\", \"Tops_per_s\": %.3f, \"ms\": %.3f, \"sms\": %d}\n", name, tops, ms / reps, nsm);
  cudaFree(out);
  return tops;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, p.multiProcessorCount, p.clockRate);
  run<0>("mad.lo.u32", 1);
  run<1>("mad.hi.u32", 1);
  run<2>("mad.wide.u32", 1);
  run<3>("add.u32 x2", 2);
  run<4>("mad.lo.cc+madc.hi", 2);
  return 0;
}
