// FP64-pipe field (csrc/fed25519.h) against the integer field, same operations as kern.cuh, outputs in canonical form
#include <cstdio>
#include <cuda_runtime.h>
#ifndef FPN
#define FPN(x) fp_##x
#endif
#include "ge25519.h"
#include "fed25519.h"

static __device__ void to_half_niels(ged_niels &h, const ge_niels &n) {
  fe inv2; for (int i = 0; i < 8; i++) inv2.v[i] = 0xffffffffu; inv2.v[0] = 0xfffffff7u; inv2.v[7] = 0x3fffffffu;  // (p + 1) / 2
  fe a, b, c; fe_mul(a, n.ypx, inv2); fe_mul(b, n.ymx, inv2); fe_mul(c, n.xy2d, inv2);
  fed da, db, dc; fed_from_fe(da, a); fed_from_fe(db, b); fed_from_fe(dc, c);
  for (int i = 0; i < 5; i++) { h.v[i] = da.v[i]; h.v[5 + i] = db.v[i]; h.v[10 + i] = dc.v[i]; }
  h.v[15] = 0;
}
__global__ void FPN(k_setup)(const uint8_t *uni, ge_p3 *pts, ged_niels *tbl, int n) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  ge_p3 p; ristretto_from_uniform(p, uni + 64 * (t & 255));
  pts[t] = p;
  if (t < 256) { ge_niels nl; ge_to_niels(nl, p); ged_niels h; to_half_niels(h, nl); tbl[t] = h; }
}
__global__ void FPN(k_mul)(ge_p3 *pts, int n, int iters) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  fei a, b; fei_from_fe(a, pts[t].X); fei_from_fe(b, pts[t].Y);
  for (int i = 0; i < iters; i++) {
    fed da, db; fed_from_fei(da, a); fed_from_fei(db, b);
    fed_mul(a, da, db);
    fed_from_fei(da, a);
    fed_mul(b, db, da);
  }
  fe_from_fei(pts[t].X, a); fe_from_fei(pts[t].Y, b);
}
__global__ void FPN(k_sq)(ge_p3 *pts, int n, int iters) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  fei a, b; fei_from_fe(a, pts[t].X); fei_from_fe(b, pts[t].Y);
  { fed d; fei one; fei_1(one); fed o; fed_from_fei(o, one); fed_from_fei(d, a); fed_mul(a, d, o); fed_from_fei(d, b); fed_mul(b, d, o); }  // balance the limbs
  for (int i = 0; i < iters; i++) {
    fed da, db; fed_from_fei(da, a); fed_from_fei(db, b);
    fed_sq(a, da); fed_sq(b, db);
  }
  fe_from_fei(pts[t].X, a); fe_from_fei(pts[t].Y, b);
}
__global__ void FPN(k_madd)(ge_p3 *pts, const ged_niels *tbl, int n, int iters) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  gei_p3 acc;
  { ge_p3 p = pts[t]; fei_from_fe_balanced(acc.X, p.X); fei_from_fe_balanced(acc.Y, p.Y); fei_from_fe_balanced(acc.Z, p.Z); fei_from_fe_balanced(acc.T, p.T); }
  unsigned idx = t * 2654435761u;
  for (int i = 0; i < iters; i++) {
    idx = idx * 1664525u + 1013904223u;
    ged_niels q; load_struct(q, tbl + ((idx >> 8) & 255));
    gei_madd(acc, acc, q, (idx >> 7) & 1);
  }
  ge_p3 r; fe_from_fei(r.X, acc.X); fe_from_fei(r.Y, acc.Y); fe_from_fei(r.Z, acc.Z); fe_from_fei(r.T, acc.T);
  pts[t] = r;
}
__global__ void FPN(k_dump)(const ge_p3 *pts, uint8_t *out, int n, int mode) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  if (mode == 0) { fe_tobytes(out + 64 * t, pts[t].X); fe_tobytes(out + 64 * t + 32, pts[t].Y); }
  else { ristretto_encode(out + 64 * t, pts[t]); for (int i = 0; i < 32; i++) out[64 * t + 32 + i] = 0; }
}

extern "C" void FPN(run)(const uint8_t *d_uni, int n, int, float *ms_out, uint8_t *h_out /* 4 * n * 64 */) {
  ge_p3 *pts; ged_niels *tbl; uint8_t *d_out;
  cudaMalloc(&pts, sizeof(ge_p3) * n); cudaMalloc(&tbl, sizeof(ged_niels) * 256); cudaMalloc(&d_out, 64 * n);
  const int T = 128, Bk = (n + T - 1) / T;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int it_mul = 1000, it_madd = 256;
  for (int test = 0; test < 3; test++) {
    FPN(k_setup)<<<Bk, T>>>(d_uni, pts, tbl, n);
    if (test == 0) FPN(k_mul)<<<Bk, T>>>(pts, n, 4);
    if (test == 1) FPN(k_sq)<<<Bk, T>>>(pts, n, 4);
    if (test == 2) FPN(k_madd)<<<Bk, T>>>(pts, tbl, n, 4);
    FPN(k_setup)<<<Bk, T>>>(d_uni, pts, tbl, n);
    cudaEventRecord(e0);
    if (test == 0) FPN(k_mul)<<<Bk, T>>>(pts, n, it_mul);
    if (test == 1) FPN(k_sq)<<<Bk, T>>>(pts, n, it_mul);
    if (test == 2) FPN(k_madd)<<<Bk, T>>>(pts, tbl, n, it_madd);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms_out[test], e0, e1);
    FPN(k_dump)<<<Bk, T>>>(pts, d_out, n, test < 2 ? 0 : 1);
    cudaMemcpy(h_out + (size_t)test * 64 * n, d_out, 64 * n, cudaMemcpyDeviceToHost);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
  cudaFree(pts); cudaFree(tbl); cudaFree(d_out);
}
