// Shared body of the field/point microbenchmark; compiled twice (legacy 10-limb headers, current headers).
#include <cstdio>
#include <cuda_runtime.h>
#include "ge25519.h"
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(IMPL, name)

// in: 64 uniform bytes per thread -> point; table: 256 niels points from multiples of the first point
__global__ void FN(_k_setup)(const uint8_t *uni, ge_p3 *pts, ge_niels *tbl, int n) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  ge_p3 p; ristretto_from_uniform(p, uni + 64 * (t & 255));
  pts[t] = p;
  if (t < 256) { ge_niels nl; ge_to_niels(nl, p); tbl[t] = nl; }
}
__global__ void FN(_k_mul)(ge_p3 *pts, int n, int iters) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  fe a = pts[t].X, b = pts[t].Y;
  for (int i = 0; i < iters; i++) { fe_mul(a, a, b); fe_mul(b, b, a); }
  pts[t].X = a; pts[t].Y = b;
}
__global__ void FN(_k_sq)(ge_p3 *pts, int n, int iters) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  fe a = pts[t].X, b = pts[t].Y;
  for (int i = 0; i < iters; i++) { fe_sq(a, a); fe_sq(b, b); }
  pts[t].X = a; pts[t].Y = b;
}
__global__ void FN(_k_madd)(ge_p3 *pts, const ge_niels *tbl, int n, int iters) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  ge_p3 acc = pts[t];
  unsigned idx = t * 2654435761u;
  for (int i = 0; i < iters; i++) {
    idx = idx * 1664525u + 1013904223u;
    ge_niels q; load_struct(q, tbl + ((idx >> 8) & 255));
    ge_madd(acc, acc, q, (idx >> 7) & 1);
  }
  pts[t] = acc;
}
__global__ void FN(_k_dbl)(ge_p3 *pts, int n, int iters) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  ge_p3 acc = pts[t];
  for (int i = 0; i < iters; i++) { ge_dbl_p2(acc, acc); }
  ge_dbl(acc, acc);
  pts[t] = acc;
}
// canonical outputs for cross-implementation comparison: fe bytes of X, Y (as field elements) and the ristretto encoding
__global__ void FN(_k_dump)(const ge_p3 *pts, uint8_t *out, int n, int mode) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  if (mode == 0) { fe_tobytes(out + 64 * t, pts[t].X); fe_tobytes(out + 64 * t + 32, pts[t].Y); }
  else { ristretto_encode(out + 64 * t, pts[t]); for (int i = 0; i < 32; i++) out[64 * t + 32 + i] = 0; }
}

extern "C" void FN(_run)(const uint8_t *d_uni, int n, int warps_per_sm_unused, float *ms_out, uint8_t *h_out /* 4 * n * 64 */) {
  ge_p3 *pts; ge_niels *tbl; uint8_t *d_out;
  cudaMalloc(&pts, sizeof(ge_p3) * n); cudaMalloc(&tbl, sizeof(ge_niels) * 256); cudaMalloc(&d_out, 64 * n);
  const int T = 128, Bk = (n + T - 1) / T;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int it_mul = 1000, it_madd = 256, it_dbl = 512;
  for (int test = 0; test < 4; test++) {
    FN(_k_setup)<<<Bk, T>>>(d_uni, pts, tbl, n);
    // warm-up (also i-cache)
    if (test == 0) FN(_k_mul)<<<Bk, T>>>(pts, n, 4);
    if (test == 1) FN(_k_sq)<<<Bk, T>>>(pts, n, 4);
    if (test == 2) FN(_k_madd)<<<Bk, T>>>(pts, tbl, n, 4);
    if (test == 3) FN(_k_dbl)<<<Bk, T>>>(pts, n, 4);
    FN(_k_setup)<<<Bk, T>>>(d_uni, pts, tbl, n);
    cudaEventRecord(e0);
    if (test == 0) FN(_k_mul)<<<Bk, T>>>(pts, n, it_mul);
    if (test == 1) FN(_k_sq)<<<Bk, T>>>(pts, n, it_mul);
    if (test == 2) FN(_k_madd)<<<Bk, T>>>(pts, tbl, n, it_madd);
    if (test == 3) FN(_k_dbl)<<<Bk, T>>>(pts, n, it_dbl);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms_out[test], e0, e1);
    FN(_k_dump)<<<Bk, T>>>(pts, d_out, n, test < 2 ? 0 : 1);
    cudaMemcpy(h_out + (size_t)test * 64 * n, d_out, 64 * n, cudaMemcpyDeviceToHost);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
  cudaFree(pts); cudaFree(tbl); cudaFree(d_out);
}
