// GF(2^255-19) on the FP64 pipe: the second field multiplier of the point-addition kernels.
//
// An IMAD.WIDE (32x32 -> 64 bits) occupies the multiplier for ~4.5 issue cycles per warp on sm_100a, a DFMA for 3
// (tools/fp64_bench.cu, profiles/r02_fp64_pipe_probe.jsonl), and a DFMA multiplies 52-bit operands: per cycle the FP64
// form moves 1.5x the operand bits.  Representation: five signed limbs of radix 2^51 ("balanced": a product leaves limbs
// in [-2^50 - 2^13, 2^50 + 2^13]), held as int64 between operations (fei: additions are integer additions) and as
// integer-valued doubles when they are multiplication operands (fed).  One limb product a*b, |a*b| < 2^103, is split
// exactly into h = floor(a*b / 2^52) and l = a*b - 2^52 h in [0, 2^52) by two fused multiply-adds in round-to-zero mode
// (the technique of Emmart, Zheng, Weems, "Faster modular exponentiation using double precision floating point
// arithmetic on the GPU", ARITH 2018):
//     hi = fma_rz(a, b, C1)           C1 = 1.5 * 2^104: the sum stays in [2^104, 2^105) where one ulp is 2^52,
//                                     so hi = C1 + 2^52 h and the mantissa field of hi moves by h
//     lo = fma_rz(a, b, C2 - hi)      C2 = C1 + 2^52: lo = 2^52 + l exactly, mantissa field = l
// and the BIT PATTERNS of hi and lo are summed per column as 64-bit integers (the biases are taken out once per column).
// Everything is exact integer arithmetic: results do not depend on the pipe they were computed on, and the host twin below
// states the same function with __int128.  Replaces (for the hot kernels) the 8x32-bit saturated form of fe25519.h, which
// stays the storage format of every buffer; reference: curve25519-dalek's FieldElement (Cargo.toml:8, un-vendored).
#pragma once
#include "fe25519.h"

struct fed { double v[5]; };   // multiplication operand: integer-valued, |v_i| <= 2^51 + 2^15
struct fei { int64_t v[5]; };  // balanced radix-2^51 limbs

#define FED_BITS_C1 0x4678000000000000ll  // bit pattern of 1.5 * 2^104
#define FED_BITS_P52 0x4330000000000000ll // bit pattern of 2^52
#define FED_MASK51 ((1ll << 51) - 1)

HD void fei_0(fei &h) {
#pragma unroll
  for (int i = 0; i < 5; i++) h.v[i] = 0;
}
HD void fei_1(fei &h) { fei_0(h); h.v[0] = 1; }
HD void fei_add(fei &h, const fei &f, const fei &g) {
#pragma unroll
  for (int i = 0; i < 5; i++) h.v[i] = f.v[i] + g.v[i];
}
HD void fei_sub(fei &h, const fei &f, const fei &g) {
#pragma unroll
  for (int i = 0; i < 5; i++) h.v[i] = f.v[i] - g.v[i];
}
// operand form of a limb vector (exact: |v| < 2^53)
HD void fed_from_fei(fed &d, const fei &f) {
#pragma unroll
  for (int i = 0; i < 5; i++) d.v[i] = (double)f.v[i];
}
HD void fed_neg_if(fed &d, int neg) {
#if defined(__CUDA_ARCH__)
  const long long m = (long long)(neg != 0) << 63;
#pragma unroll
  for (int i = 0; i < 5; i++) d.v[i] = __longlong_as_double(__double_as_longlong(d.v[i]) ^ m);
#else
  if (neg) for (int i = 0; i < 5; i++) d.v[i] = -d.v[i];
#endif
}
HD void fed_select(fed &h, const fed &f, const fed &g, int b) {
#pragma unroll
  for (int i = 0; i < 5; i++) h.v[i] = b ? g.v[i] : f.v[i];
}

// H += bit pattern of (C1 + 2^52 floor(a b / 2^52)),  L += bit pattern of (2^52 + (a b mod 2^52))
HD void fed_prod(int64_t &H, int64_t &L, double a, double b) {
#if defined(__CUDA_ARCH__)
  const double c1 = __longlong_as_double(FED_BITS_C1);
  const double c2 = __longlong_as_double(FED_BITS_C1 + 1);  // C1 + one ulp = C1 + 2^52
  const double hi = __fma_rz(a, b, c1);
  const double lo = __fma_rz(a, b, __dsub_rn(c2, hi));
  H += __double_as_longlong(hi);
  L += __double_as_longlong(lo);
#else
  const __int128 p = (__int128)(int64_t)a * (__int128)(int64_t)b;
  const int64_t h = (int64_t)(p >> 52);  // arithmetic shift: floor
  H += FED_BITS_C1 + h;
  L += FED_BITS_P52 + (int64_t)(p - ((__int128)h << 52));
#endif
}

// columns H[k], L[k] (biases removed) of a 5x5 product -> balanced limbs.  With c_k = 2^52 H[k] + L[k] at weight 2^(51 k):
// value = sum_k (L[k] + 2 H[k-1]) 2^(51 k), k = 0..9; the upper five columns fold with 2^255 = 19.
HD void fed_reduce(fei &out, const int64_t H[9], const int64_t L[9]) {
  int64_t T[10];
  T[0] = L[0];
#pragma unroll
  for (int k = 1; k < 9; k++) T[k] = L[k] + 2 * H[k - 1];
  T[9] = 2 * H[8];
  int64_t c[5], r[5];
#pragma unroll
  for (int k = 0; k < 5; k++) {
    const int64_t U = T[k] + 19 * T[k + 5];       // |U| < 2^59.3
    const int64_t V = U + (1ll << 50);
    c[k] = V >> 51;                               // |c| < 2^9
    r[k] = (V & FED_MASK51) - (1ll << 50);        // in [-2^50, 2^50)
  }
  out.v[0] = r[0] + 19 * c[4];
#pragma unroll
  for (int k = 1; k < 5; k++) out.v[k] = r[k] + c[k - 1];
}

HD void fed_mul_inl(fei &out, const fed &f, const fed &g) {
  int64_t H[9], L[9];
#pragma unroll
  for (int k = 0; k < 9; k++) {
    const int cnt = k < 5 ? k + 1 : 9 - k;        // products in column k
    H[k] = -(int64_t)cnt * FED_BITS_C1;           // wraps mod 2^64; the true column sums are small
    L[k] = -(int64_t)cnt * FED_BITS_P52;
  }
#pragma unroll
  for (int i = 0; i < 5; i++)
#pragma unroll
    for (int j = 0; j < 5; j++) fed_prod(H[i + j], L[i + j], f.v[i], g.v[j]);
  fed_reduce(out, H, L);
}
// squaring: 15 limb products, the cross products against the doubled operand
HD void fed_sq_inl(fei &out, const fed &f) {
  int64_t H[9], L[9];
#pragma unroll
  for (int k = 0; k < 9; k++) {
    int cnt = 0;
    for (int i = 0; i < 5; i++) { const int j = k - i; if (j >= i && j < 5) cnt++; }
    H[k] = -(int64_t)cnt * FED_BITS_C1;
    L[k] = -(int64_t)cnt * FED_BITS_P52;
  }
  double f2[5];
#pragma unroll
  for (int i = 0; i < 5; i++) f2[i] = f.v[i] + f.v[i];
#pragma unroll
  for (int i = 0; i < 5; i++)
#pragma unroll
    for (int j = i; j < 5; j++) fed_prod(H[i + j], L[i + j], f.v[i], i == j ? f.v[j] : f2[j]);
  fed_reduce(out, H, L);
}

#if defined(__CUDACC__) && BP_FE_CALL
static __device__ __noinline__ fei fed_mul_fn(fed f, fed g) { fei h; fed_mul_inl(h, f, g); return h; }
static __device__ __noinline__ fei fed_sq_fn(fed f) { fei h; fed_sq_inl(h, f); return h; }
#endif
HD void fed_mul(fei &out, const fed &f, const fed &g) {
#if defined(__CUDA_ARCH__) && BP_FE_CALL
  out = fed_mul_fn(f, g);
#else
  fed_mul_inl(out, f, g);
#endif
}
HD void fed_sq(fei &out, const fed &f) {
#if defined(__CUDA_ARCH__) && BP_FE_CALL
  out = fed_sq_fn(f);
#else
  fed_sq_inl(out, f);
#endif
}

// ---- conversions between the storage form (fe, 8 x 32 bits, any 256-bit representative) and the limb form ----
// non-negative limbs, l_0 < 2^51 + 19, l_1..l_4 < 2^51
HD void fei_from_fe(fei &h, const fe &f) {
  uint64_t w[4];
#pragma unroll
  for (int i = 0; i < 4; i++) w[i] = (uint64_t)f.v[2 * i] | ((uint64_t)f.v[2 * i + 1] << 32);
  h.v[0] = (int64_t)(w[0] & FED_MASK51) + 19 * (int64_t)(w[3] >> 63);
  h.v[1] = (int64_t)(((w[0] >> 51) | (w[1] << 13)) & FED_MASK51);
  h.v[2] = (int64_t)(((w[1] >> 38) | (w[2] << 26)) & FED_MASK51);
  h.v[3] = (int64_t)(((w[2] >> 25) | (w[3] << 39)) & FED_MASK51);
  h.v[4] = (int64_t)((w[3] >> 12) & FED_MASK51);
}
HD void fed_from_fe(fed &d, const fe &f) { fei t; fei_from_fe(t, f); fed_from_fei(d, t); }
// same value, limbs in [-2^50 - 2^5, 2^50 + 2^5]: the form sums of two elements may be multiplied in (input |v_i| < 2^62)
HD void fei_balance(fei &h, const fei &f) {
  int64_t c[5], r[5];
#pragma unroll
  for (int k = 0; k < 5; k++) {
    const int64_t V = f.v[k] + (1ll << 50);
    c[k] = V >> 51;
    r[k] = (V & FED_MASK51) - (1ll << 50);
  }
  h.v[0] = r[0] + 19 * c[4];
#pragma unroll
  for (int k = 1; k < 5; k++) h.v[k] = r[k] + c[k - 1];
}
HD void fei_from_fe_balanced(fei &h, const fe &f) { fei t; fei_from_fe(t, f); fei_balance(h, t); }
// any limb vector with |v_i| < 2^52 -> a 256-bit representative
HD void fe_from_fei(fe &f, const fei &h) {
  // + 4p keeps every limb positive: 4p = (2^53 - 76, 2^53 - 4, 2^53 - 4, 2^53 - 4, 2^53 - 4)
  uint64_t l[5];
  l[0] = (uint64_t)(h.v[0] + ((1ll << 53) - 76));
#pragma unroll
  for (int i = 1; i < 5; i++) l[i] = (uint64_t)(h.v[i] + ((1ll << 53) - 4));
  // one sequential carry pass: l_1..l_4 < 2^51, l_0 < 2^51 + 19 * 8
#pragma unroll
  for (int i = 0; i < 4; i++) { l[i + 1] += l[i] >> 51; l[i] &= (uint64_t)FED_MASK51; }
  l[0] += 19 * (l[4] >> 51); l[4] &= (uint64_t)FED_MASK51;
  // pack: value < 2^255 + 2^59
  uint64_t w[4], c;
  w[0] = l[0] + (l[1] << 51); c = w[0] < l[0];
  uint64_t t = (l[1] >> 13) + c;             // < 2^38 + 1
  w[1] = t + (l[2] << 38); c = w[1] < t;
  t = (l[2] >> 26) + c;
  w[2] = t + (l[3] << 25); c = w[2] < t;
  t = (l[3] >> 39) + c;
  w[3] = t + (l[4] << 12);
#pragma unroll
  for (int i = 0; i < 4; i++) { f.v[2 * i] = (uint32_t)w[i]; f.v[2 * i + 1] = (uint32_t)(w[i] >> 32); }
}

// ---- points: accumulator in limb form, table entries as "half Niels" operands ----
// Half Niels: ((y + x)/2, (y - x)/2, d x y) -- every product of the mixed addition is then HALF the textbook one and so are
// E, F, G, H (with D = Z instead of 2 Z): the result is the same projective point with all four coordinates divided by four,
// and no operand exceeds 2^51 + 2^14 in magnitude.  128-byte entries (one line): 3 x 5 doubles + padding.
struct alignas(16) ged_niels { double v[16]; };  // v[0..5) = (y+x)/2, v[5..10) = (y-x)/2, v[10..15) = d x y
struct gei_p3 { fei X, Y, Z, T; };

HD void gei_identity(gei_p3 &p) { fei_0(p.X); fei_1(p.Y); fei_1(p.Z); fei_0(p.T); }

// r = p + (neg ? -q : q)
HD void gei_madd(gei_p3 &r, const gei_p3 &p, const ged_niels &q, int neg) {
  fei a, b, A, B, C;
  fei_add(a, p.Y, p.X); fei_sub(b, p.Y, p.X);
  fed da, db, dt, qp, qm, qt;
#pragma unroll
  for (int i = 0; i < 5; i++) { qp.v[i] = neg ? q.v[5 + i] : q.v[i]; qm.v[i] = neg ? q.v[i] : q.v[5 + i]; qt.v[i] = q.v[10 + i]; }
  fed_from_fei(da, a); fed_from_fei(db, b); fed_from_fei(dt, p.T);
  fed_neg_if(dt, neg);
  fed_mul(A, da, qp); fed_mul(B, db, qm); fed_mul(C, dt, qt);
  fei E, F, G, Hh;
  fei_sub(E, A, B); fei_add(Hh, A, B); fei_sub(F, p.Z, C); fei_add(G, p.Z, C);
  fed dE, dF, dG, dH;
  fed_from_fei(dE, E); fed_from_fei(dF, F); fed_from_fei(dG, G); fed_from_fei(dH, Hh);
  fed_mul(r.X, dE, dF); fed_mul(r.Y, dG, dH); fed_mul(r.Z, dF, dG); fed_mul(r.T, dE, dH);
}
