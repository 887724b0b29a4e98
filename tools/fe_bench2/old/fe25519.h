// GF(2^255-19) arithmetic, 10 signed limbs in radix 2^25.5 (limb i holds ceil(25.5*i) bits),
// 64-bit column accumulators.  On sm_100a every limb product is one IMAD.WIDE on the fma pipe;
// carries run on the alu pipe.  Replaces curve25519-dalek's FieldElement (reference
// Cargo.toml:8, an un-vendored dependency); representation and schedule are our own.
//
// Bounds (standard for this radix): fe_mul/fe_sq accept limbs up to ~1.65*2^26 (even) /
// 1.65*2^25 (odd) in magnitude, i.e. a sum or difference of up to three carried elements,
// and return carried limbs (|h_even| <= 1.01*2^25, |h_odd| <= 1.01*2^24).
#pragma once
#include "hd.h"
#include "constants.h"

struct fe { int32_t v[10]; };

HD void fe_0(fe &h) {
#pragma unroll
  for (int i = 0; i < 10; i++) h.v[i] = 0;
}
HD void fe_1(fe &h) { fe_0(h); h.v[0] = 1; }
HD void fe_add(fe &h, const fe &f, const fe &g) {
#pragma unroll
  for (int i = 0; i < 10; i++) h.v[i] = f.v[i] + g.v[i];
}
HD void fe_sub(fe &h, const fe &f, const fe &g) {
#pragma unroll
  for (int i = 0; i < 10; i++) h.v[i] = f.v[i] - g.v[i];
}
HD void fe_neg(fe &h, const fe &f) {
#pragma unroll
  for (int i = 0; i < 10; i++) h.v[i] = -f.v[i];
}
// h = b ? g : f   (branch-free)
HD void fe_select(fe &h, const fe &f, const fe &g, int b) {
  int32_t m = -(int32_t)(b != 0);
#pragma unroll
  for (int i = 0; i < 10; i++) h.v[i] = f.v[i] ^ (m & (f.v[i] ^ g.v[i]));
}
HD void fe_cswap(fe &f, fe &g, int b) {
  int32_t m = -(int32_t)(b != 0);
#pragma unroll
  for (int i = 0; i < 10; i++) { int32_t x = m & (f.v[i] ^ g.v[i]); f.v[i] ^= x; g.v[i] ^= x; }
}
// h = b ? -f : f
HD void fe_cneg(fe &h, const fe &f, int b) {
  int32_t m = -(int32_t)(b != 0);
#pragma unroll
  for (int i = 0; i < 10; i++) h.v[i] = (f.v[i] ^ m) - m;
}

// carry 10 wide columns into limbs
HD void fe_carry_wide(fe &out, int64_t h[10]) {
  int64_t c;
  c = (h[0] + (1LL << 25)) >> 26; h[1] += c; h[0] -= c << 26;
  c = (h[4] + (1LL << 25)) >> 26; h[5] += c; h[4] -= c << 26;
  c = (h[1] + (1LL << 24)) >> 25; h[2] += c; h[1] -= c << 25;
  c = (h[5] + (1LL << 24)) >> 25; h[6] += c; h[5] -= c << 25;
  c = (h[2] + (1LL << 25)) >> 26; h[3] += c; h[2] -= c << 26;
  c = (h[6] + (1LL << 25)) >> 26; h[7] += c; h[6] -= c << 26;
  c = (h[3] + (1LL << 24)) >> 25; h[4] += c; h[3] -= c << 25;
  c = (h[7] + (1LL << 24)) >> 25; h[8] += c; h[7] -= c << 25;
  c = (h[4] + (1LL << 25)) >> 26; h[5] += c; h[4] -= c << 26;
  c = (h[8] + (1LL << 25)) >> 26; h[9] += c; h[8] -= c << 26;
  c = (h[9] + (1LL << 24)) >> 25; h[0] += c * 19; h[9] -= c << 25;
  c = (h[0] + (1LL << 25)) >> 26; h[1] += c; h[0] -= c << 26;
#pragma unroll
  for (int i = 0; i < 10; i++) out.v[i] = (int32_t)h[i];
}

HD void fe_mul_inl(fe &out, const fe &f, const fe &g) {
  int32_t g19[10], f2[10];
#pragma unroll
  for (int i = 0; i < 10; i++) { g19[i] = 19 * g.v[i]; f2[i] = 2 * f.v[i]; }
  int64_t h[10];
#pragma unroll
  for (int k = 0; k < 10; k++) h[k] = 0;
#pragma unroll
  for (int i = 0; i < 10; i++) {
#pragma unroll
    for (int j = 0; j < 10; j++) {
      // both odd -> the 2^25.5 radix needs a doubling; wrap past limb 9 -> times 19
      int32_t a = ((i & 1) && (j & 1)) ? f2[i] : f.v[i];
      int32_t b = (i + j >= 10) ? g19[j] : g.v[j];
      h[(i + j) % 10] += (int64_t)a * (int64_t)b;
    }
  }
  fe_carry_wide(out, h);
}

HD void fe_sq_wide(int64_t h[10], const fe &f) {
  int32_t f2[10], fw[10];  // fw[j] = f[j] * (19 if wrapped) * (2 if odd and partner odd) chosen per pair below
#pragma unroll
  for (int i = 0; i < 10; i++) f2[i] = 2 * f.v[i];
  (void)fw;
#pragma unroll
  for (int k = 0; k < 10; k++) h[k] = 0;
#pragma unroll
  for (int i = 0; i < 10; i++) {
#pragma unroll
    for (int j = i; j < 10; j++) {
      int32_t a = (i != j) ? f2[i] : f.v[i];
      int32_t b = f.v[j];
      if ((i & 1) && (j & 1)) b *= 2;
      if (i + j >= 10) b *= 19;
      h[(i + j) % 10] += (int64_t)a * (int64_t)b;
    }
  }
}
HD void fe_sq_inl(fe &out, const fe &f) {
  int64_t h[10];
  fe_sq_wide(h, f);
  fe_carry_wide(out, h);
}
// out = 2 f^2
HD void fe_sq2_inl(fe &out, const fe &f) {
  int64_t h[10];
  fe_sq_wide(h, f);
#pragma unroll
  for (int k = 0; k < 10; k++) h[k] += h[k];
  fe_carry_wide(out, h);
}
// On the device the three multiplication primitives are real functions (operands and result travel in registers):
// a point addition is then ~10 calls instead of ~2000 inlined instructions per multiplication site, which keeps the
// hot loops inside the instruction cache (ncu showed 20 % "no_instructions" stalls with everything inlined).
#ifndef BP_FE_CALL
#define BP_FE_CALL 1
#endif
#if defined(__CUDACC__) && BP_FE_CALL
static __device__ __noinline__ fe fe_mul_fn(fe f, fe g) { fe h; fe_mul_inl(h, f, g); return h; }
static __device__ __noinline__ fe fe_sq_fn(fe f) { fe h; fe_sq_inl(h, f); return h; }
static __device__ __noinline__ fe fe_sq2_fn(fe f) { fe h; fe_sq2_inl(h, f); return h; }
#endif
HD void fe_mul(fe &out, const fe &f, const fe &g) {
#if defined(__CUDA_ARCH__) && BP_FE_CALL
  out = fe_mul_fn(f, g);
#else
  fe_mul_inl(out, f, g);
#endif
}
HD void fe_sq(fe &out, const fe &f) {
#if defined(__CUDA_ARCH__) && BP_FE_CALL
  out = fe_sq_fn(f);
#else
  fe_sq_inl(out, f);
#endif
}
HD void fe_sq2(fe &out, const fe &f) {
#if defined(__CUDA_ARCH__) && BP_FE_CALL
  out = fe_sq2_fn(f);
#else
  fe_sq2_inl(out, f);
#endif
}
HD void fe_sqn(fe &out, const fe &f, int n) {
  fe_sq(out, f);
  for (int i = 1; i < n; i++) fe_sq(out, out);
}

HD void fe_frombytes(fe &out, const uint8_t *s) {
  auto ld3 = [&](int o) { return (int64_t)((uint64_t)s[o] | ((uint64_t)s[o + 1] << 8) | ((uint64_t)s[o + 2] << 16)); };
  auto ld4 = [&](int o) { return ld3(o) | (int64_t)((uint64_t)s[o + 3] << 24); };
  int64_t h[10];
  h[0] = ld4(0);
  h[1] = ld3(4) << 6;
  h[2] = ld3(7) << 5;
  h[3] = ld3(10) << 3;
  h[4] = ld3(13) << 2;
  h[5] = ld4(16);
  h[6] = ld3(20) << 7;
  h[7] = ld3(23) << 5;
  h[8] = ld3(26) << 4;
  h[9] = (ld3(29) & 8388607) << 2;
  int64_t c;
  c = (h[9] + (1LL << 24)) >> 25; h[0] += c * 19; h[9] -= c << 25;
  c = (h[1] + (1LL << 24)) >> 25; h[2] += c; h[1] -= c << 25;
  c = (h[3] + (1LL << 24)) >> 25; h[4] += c; h[3] -= c << 25;
  c = (h[5] + (1LL << 24)) >> 25; h[6] += c; h[5] -= c << 25;
  c = (h[7] + (1LL << 24)) >> 25; h[8] += c; h[7] -= c << 25;
  c = (h[0] + (1LL << 25)) >> 26; h[1] += c; h[0] -= c << 26;
  c = (h[2] + (1LL << 25)) >> 26; h[3] += c; h[2] -= c << 26;
  c = (h[4] + (1LL << 25)) >> 26; h[5] += c; h[4] -= c << 26;
  c = (h[6] + (1LL << 25)) >> 26; h[7] += c; h[6] -= c << 26;
  c = (h[8] + (1LL << 25)) >> 26; h[9] += c; h[8] -= c << 26;
#pragma unroll
  for (int i = 0; i < 10; i++) out.v[i] = (int32_t)h[i];
}

// canonical little-endian encoding (fully reduced mod p)
HD void fe_tobytes(uint8_t *s, const fe &f) {
  int32_t h[10];
#pragma unroll
  for (int i = 0; i < 10; i++) h[i] = f.v[i];
  int32_t q = (19 * h[9] + (1 << 24)) >> 25;
#pragma unroll
  for (int i = 0; i < 10; i++) q = (h[i] + q) >> ((i & 1) ? 25 : 26);
  h[0] += 19 * q;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    int sh = (i & 1) ? 25 : 26;
    int32_t c = h[i] >> sh; h[i + 1] += c; h[i] -= c << sh;
  }
  { int32_t c = h[9] >> 25; h[9] -= c << 25; }
  uint32_t w[8];
  uint64_t acc = 0; int bits = 0, wi = 0;
#pragma unroll
  for (int i = 0; i < 10; i++) {
    int sh = (i & 1) ? 25 : 26;
    acc |= (uint64_t)(uint32_t)h[i] << bits; bits += sh;
    if (bits >= 32) { w[wi++] = (uint32_t)acc; acc >>= 32; bits -= 32; }
  }
  if (wi < 8) w[wi++] = (uint32_t)acc;
#pragma unroll
  for (int i = 0; i < 8; i++) { s[4 * i] = (uint8_t)w[i]; s[4 * i + 1] = (uint8_t)(w[i] >> 8); s[4 * i + 2] = (uint8_t)(w[i] >> 16); s[4 * i + 3] = (uint8_t)(w[i] >> 24); }
}

HD int fe_isnegative(const fe &f) { uint8_t s[32]; fe_tobytes(s, f); return s[0] & 1; }
HD int fe_iszero(const fe &f) {
  uint8_t s[32]; fe_tobytes(s, f);
  uint8_t r = 0;
#pragma unroll
  for (int i = 0; i < 32; i++) r |= s[i];
  return r == 0;
}
HD int fe_equal(const fe &a, const fe &b) { fe d; fe_sub(d, a, b); return fe_iszero(d); }
HD void fe_abs(fe &h, const fe &f) { fe_cneg(h, f, fe_isnegative(f)); }

// z^(2^252-3)
HD void fe_pow22523(fe &out, const fe &z) {
  fe t0, t1, t2;
  fe_sq(t0, z); fe_sqn(t1, t0, 2); fe_mul(t1, z, t1); fe_mul(t0, t0, t1);
  fe_sq(t0, t0); fe_mul(t0, t1, t0);
  fe_sqn(t1, t0, 5); fe_mul(t0, t1, t0);
  fe_sqn(t1, t0, 10); fe_mul(t1, t1, t0);
  fe_sqn(t2, t1, 20); fe_mul(t1, t2, t1);
  fe_sqn(t1, t1, 10); fe_mul(t0, t1, t0);
  fe_sqn(t1, t0, 50); fe_mul(t1, t1, t0);
  fe_sqn(t2, t1, 100); fe_mul(t1, t2, t1);
  fe_sqn(t1, t1, 50); fe_mul(t0, t1, t0);
  fe_sqn(t0, t0, 2); fe_mul(out, t0, z);
}
// z^(p-2)
HD void fe_invert(fe &out, const fe &z) {
  fe t0, t1, t2, t3;
  fe_sq(t0, z); fe_sqn(t1, t0, 2); fe_mul(t1, z, t1); fe_mul(t0, t0, t1);
  fe_sq(t2, t0); fe_mul(t1, t1, t2);
  fe_sqn(t2, t1, 5); fe_mul(t1, t2, t1);
  fe_sqn(t2, t1, 10); fe_mul(t2, t2, t1);
  fe_sqn(t3, t2, 20); fe_mul(t2, t3, t2);
  fe_sqn(t2, t2, 10); fe_mul(t1, t2, t1);
  fe_sqn(t2, t1, 50); fe_mul(t2, t2, t1);
  fe_sqn(t3, t2, 100); fe_mul(t2, t3, t2);
  fe_sqn(t2, t2, 50); fe_mul(t1, t2, t1);
  fe_sqn(t1, t1, 5); fe_mul(out, t1, t0);
}

#define FE_CONST(name) HD void name(fe &h) { const int32_t c[10] = name##_LIMBS; _Pragma("unroll") for (int i = 0; i < 10; i++) h.v[i] = c[i]; }
FE_CONST(FE_D)
FE_CONST(FE_2D)
FE_CONST(FE_SQRTM1)
FE_CONST(FE_ONE_MINUS_D_SQ)
FE_CONST(FE_D_MINUS_ONE_SQ)
FE_CONST(FE_SQRT_AD_MINUS_ONE)
FE_CONST(FE_INVSQRT_A_MINUS_D)

// RFC 9496 section 4.2 SQRT_RATIO_M1: returns was_square; r = non-negative sqrt(u/v) or sqrt(i*u/v)
HD int fe_sqrt_ratio_m1(fe &r, const fe &u, const fe &v) {
  fe v3, v7, t, check, neg_u, neg_u_i, i;
  FE_SQRTM1(i);
  fe_sq(v3, v); fe_mul(v3, v3, v);
  fe_sq(v7, v3); fe_mul(v7, v7, v);
  fe_mul(t, u, v7); fe_pow22523(t, t);
  fe_mul(r, u, v3); fe_mul(r, r, t);
  fe_sq(check, r); fe_mul(check, check, v);
  fe_neg(neg_u, u); fe_mul(neg_u_i, neg_u, i);
  int correct = fe_equal(check, u), flipped = fe_equal(check, neg_u), flipped_i = fe_equal(check, neg_u_i);
  fe ri; fe_mul(ri, r, i);
  fe_select(r, r, ri, flipped | flipped_i);
  fe_abs(r, r);
  return correct | flipped;
}
