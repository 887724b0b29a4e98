// Microbenchmark of the GF(2^255-19) multiplication forms on sm_100a (developer tool; results in profiles/).
#include <cstdio>
#include <cuda_runtime.h>
#include "../bulletproofs_r1cs_gadgets_b200/csrc/ge25519.h"

// variant B: same products, floor carries with masks (no rounding bias)
__device__ __forceinline__ void carry_floor(fe &out, int64_t h[10]) {
  int64_t c;
#pragma unroll
  for (int r = 0; r < 2; r++) {
#pragma unroll
    for (int i = 0; i < 10; i++) {
      int sh = (i & 1) ? 25 : 26;
      c = h[i] >> sh; h[i] &= ((1LL << sh) - 1);
      if (i < 9) h[i + 1] += c; else h[0] += 19 * c;
    }
    if (r == 0) { /* second pass only needs first two limbs in practice */ }
  }
#pragma unroll
  for (int i = 0; i < 10; i++) out.v[i] = (int32_t)h[i];
}
__device__ __forceinline__ void carry_floor1(fe &out, int64_t h[10]) {
  int64_t c;
#pragma unroll
  for (int i = 0; i < 10; i++) {
    int sh = (i & 1) ? 25 : 26;
    c = h[i] >> sh; h[i] &= ((1LL << sh) - 1);
    if (i < 9) h[i + 1] += c; else h[0] += 19 * c;
  }
  c = h[0] >> 26; h[0] &= ((1LL << 26) - 1); h[1] += c;
#pragma unroll
  for (int i = 0; i < 10; i++) out.v[i] = (int32_t)h[i];
}
__device__ __forceinline__ void mul_wide(int64_t h[10], const fe &f, const fe &g) {
  int32_t g19[10], f2[10];
#pragma unroll
  for (int i = 0; i < 10; i++) { g19[i] = 19 * g.v[i]; f2[i] = 2 * f.v[i]; }
#pragma unroll
  for (int k = 0; k < 10; k++) h[k] = 0;
#pragma unroll
  for (int i = 0; i < 10; i++)
#pragma unroll
    for (int j = 0; j < 10; j++) {
      int32_t a = ((i & 1) && (j & 1)) ? f2[i] : f.v[i];
      int32_t b = (i + j >= 10) ? g19[j] : g.v[j];
      h[(i + j) % 10] += (int64_t)a * (int64_t)b;
    }
}
__device__ __noinline__ fe mulB_fn(fe f, fe g) { int64_t h[10]; mul_wide(h, f, g); fe o; carry_floor1(o, h); return o; }
__device__ __noinline__ fe mulNoCarry_fn(fe f, fe g) { int64_t h[10]; mul_wide(h, f, g); fe o; for (int i = 0; i < 10; i++) o.v[i] = (int32_t)(h[i] >> 20); return o; }

template <int MODE>
__global__ void kern(fe *io, int iters) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  fe a = io[2 * t], b = io[2 * t + 1];
  for (int i = 0; i < iters; i++) {
    if (MODE == 0) { fe_mul_inl(a, a, b); fe_mul_inl(b, b, a); }                       // inlined, dependent
    if (MODE == 1) { a = fe_mul_fn(a, b); b = fe_mul_fn(b, a); }                        // call, dependent
    if (MODE == 2) { a = mulB_fn(a, b); b = mulB_fn(b, a); }                            // floor carry
    if (MODE == 3) { a = mulNoCarry_fn(a, b); b = mulNoCarry_fn(b, a); }                // products only
    if (MODE == 4) { fe c, d; fe_mul_inl(c, a, b); fe_mul_inl(d, b, b); fe_add(a, c, d); fe_sub(b, c, d); fe_carry(b);}  // 2 independent inlined
    if (MODE == 5) { a = fe_sq_fn(a); b = fe_sq_fn(b); }
  }
  io[2 * t] = a; io[2 * t + 1] = b;
}
template <int MODE>
void run(const char *name, int blocks_per_sm, int threads, fe *d, int muls_per_iter) {
  int iters = 2000;
  int blocks = 148 * blocks_per_sm;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<MODE><<<blocks, threads>>>(d, 10);
  cudaEventRecord(e0);
  kern<MODE><<<blocks, threads>>>(d, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double muls = (double)blocks * threads * iters * muls_per_iter;
  printf("{\"variant\": \"%s\", \"warps_per_sm\": %d, \"Gmul_per_s\": %.1f, \"ms\": %.2f}\n", name, blocks_per_sm * threads / 32, muls / ms / 1e6, ms);
}
int main() {
  fe *d; cudaMalloc(&d, sizeof(fe) * 2 * 148 * 16 * 256);
  cudaMemset(d, 1, sizeof(fe) * 2 * 148 * 16 * 256);
  for (int bps : {2, 4, 8, 16}) {
    run<0>("inline_dep", bps, 128, d, 2);
    run<1>("call_dep", bps, 128, d, 2);
    run<2>("call_floorcarry", bps, 128, d, 2);
    run<3>("call_products_only", bps, 128, d, 2);
    run<4>("inline_2indep", bps, 128, d, 2);
    run<5>("call_sq", bps, 128, d, 2);
  }
  return 0;
}
