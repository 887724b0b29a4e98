// Scalar field F_l, l = 2^252 + 27742317777372353535851937790883648493, in Montgomery form
// (R = 2^256) on 4 x 64-bit limbs.  Replaces curve25519-dalek's Scalar (reference
// Cargo.toml:8; call sites e.g. src/gadget_poseidon.rs:123,162, src/scalar_utils.rs:236).
// All device-resident scalars are kept in Montgomery form; bytes crossing the C-ABI are
// canonical 32-byte little-endian.
#pragma once
#include "hd.h"
#include "constants.h"

struct alignas(16) scm { uint64_t v[4]; };  // value * 2^256 mod l, always < l

HD void sc_const_l(uint64_t l[4]) { const uint64_t c[4] = SC_L_LIMBS; l[0] = c[0]; l[1] = c[1]; l[2] = c[2]; l[3] = c[3]; }
HD scm sc_zero() { scm r; r.v[0] = r.v[1] = r.v[2] = r.v[3] = 0; return r; }
HD scm sc_one() { const uint64_t c[4] = SC_R_LIMBS; scm r; r.v[0] = c[0]; r.v[1] = c[1]; r.v[2] = c[2]; r.v[3] = c[3]; return r; }
HD scm sc_r2() { const uint64_t c[4] = SC_R2_LIMBS; scm r; r.v[0] = c[0]; r.v[1] = c[1]; r.v[2] = c[2]; r.v[3] = c[3]; return r; }
HD scm sc_r3() { const uint64_t c[4] = SC_R3_LIMBS; scm r; r.v[0] = c[0]; r.v[1] = c[1]; r.v[2] = c[2]; r.v[3] = c[3]; return r; }

HD uint64_t addc64(uint64_t a, uint64_t b, uint64_t &carry) {
  uint64_t s = a + carry; uint64_t c1 = s < carry; uint64_t r = s + b; carry = c1 + (r < b); return r;
}
HD uint64_t subb64(uint64_t a, uint64_t b, uint64_t &borrow) {
  uint64_t d = a - b; uint64_t b1 = a < b; uint64_t r = d - borrow; borrow = b1 + (d < borrow); return r;
}
// r = a - l if a >= l else a   (a < 2l)
HD void sc_cond_sub_l(uint64_t a[4], uint64_t top) {
  uint64_t l[4]; sc_const_l(l);
  uint64_t t[4], br = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) t[i] = subb64(a[i], l[i], br);
  // if top is set or no borrow, take t
  uint64_t take = (top != 0) | (br == 0);
  uint64_t m = (uint64_t)0 - take;
#pragma unroll
  for (int i = 0; i < 4; i++) a[i] = (a[i] & ~m) | (t[i] & m);
}
HD scm sc_add(const scm &a, const scm &b) {
  scm r; uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) r.v[i] = addc64(a.v[i], b.v[i], c);
  sc_cond_sub_l(r.v, c);
  return r;
}
HD scm sc_sub(const scm &a, const scm &b) {
  scm r; uint64_t br = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) r.v[i] = subb64(a.v[i], b.v[i], br);
  uint64_t l[4]; sc_const_l(l);
  uint64_t m = (uint64_t)0 - (br != 0), c = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) r.v[i] = addc64(r.v[i], l[i] & m, c);
  return r;
}
HD scm sc_neg(const scm &a) { return sc_sub(sc_zero(), a); }
HD int sc_is_zero(const scm &a) { return (a.v[0] | a.v[1] | a.v[2] | a.v[3]) == 0; }
HD int sc_equal(const scm &a, const scm &b) { return ((a.v[0] ^ b.v[0]) | (a.v[1] ^ b.v[1]) | (a.v[2] ^ b.v[2]) | (a.v[3] ^ b.v[3])) == 0; }

// Montgomery product a*b/2^256 mod l (CIOS); a may be any 256-bit value, b < l  ->  result < l.
// Two forms: 4 x 64-bit words with 128-bit products for host code, and 8 x 32-bit words for the device,
// where one 32x32+64 multiply-add is a single IMAD.WIDE (the emulation build uses the device form too,
// so the CPU-side kernel-body tests cover it).
#if defined(__CUDA_ARCH__) || defined(BP_HOST_EMUL)
HD scm sc_montmul(const scm &a, const scm &b) {
  const uint64_t l64[4] = SC_L_LIMBS;
  uint32_t l[8], aw[8], bw[8];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    l[2 * i] = (uint32_t)l64[i]; l[2 * i + 1] = (uint32_t)(l64[i] >> 32);
    aw[2 * i] = (uint32_t)a.v[i]; aw[2 * i + 1] = (uint32_t)(a.v[i] >> 32);
    bw[2 * i] = (uint32_t)b.v[i]; bw[2 * i + 1] = (uint32_t)(b.v[i] >> 32);
  }
  const uint32_t ninv = (uint32_t)SC_NINV;
  uint32_t t[10];
#pragma unroll
  for (int i = 0; i < 10; i++) t[i] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint64_t acc, carry = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) { acc = (uint64_t)aw[j] * bw[i] + t[j] + carry; t[j] = (uint32_t)acc; carry = acc >> 32; }
    acc = (uint64_t)t[8] + carry; t[8] = (uint32_t)acc; t[9] = (uint32_t)(acc >> 32);
    uint32_t m = t[0] * ninv;
    acc = (uint64_t)m * l[0] + t[0]; carry = acc >> 32;
#pragma unroll
    for (int j = 1; j < 8; j++) { acc = (uint64_t)m * l[j] + t[j] + carry; t[j - 1] = (uint32_t)acc; carry = acc >> 32; }
    acc = (uint64_t)t[8] + carry; t[7] = (uint32_t)acc; t[8] = t[9] + (uint32_t)(acc >> 32);
  }
  scm r;
#pragma unroll
  for (int i = 0; i < 4; i++) r.v[i] = (uint64_t)t[2 * i] | ((uint64_t)t[2 * i + 1] << 32);
  sc_cond_sub_l(r.v, t[8]);
  sc_cond_sub_l(r.v, 0);
  return r;
}
#else
HD scm sc_montmul(const scm &a, const scm &b) {
  uint64_t l[4]; sc_const_l(l);
  uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint64_t hi, lo, c, k;
    uint64_t bi = b.v[i];
    // t += a * b[i]
    mul64wide(a.v[0], bi, hi, lo); k = 0; t0 = addc64(t0, lo, k); c = hi + k;
    mul64wide(a.v[1], bi, hi, lo); k = 0; t1 = addc64(t1, lo, k); hi += k; k = 0; t1 = addc64(t1, c, k); c = hi + k;
    mul64wide(a.v[2], bi, hi, lo); k = 0; t2 = addc64(t2, lo, k); hi += k; k = 0; t2 = addc64(t2, c, k); c = hi + k;
    mul64wide(a.v[3], bi, hi, lo); k = 0; t3 = addc64(t3, lo, k); hi += k; k = 0; t3 = addc64(t3, c, k); c = hi + k;
    k = 0; t4 = addc64(t4, c, k); t5 = k;
    // t += m * l ; t >>= 64
    uint64_t m = t0 * SC_NINV;
    mul64wide(m, l[0], hi, lo); k = 0; (void)addc64(t0, lo, k); c = hi + k;
    mul64wide(m, l[1], hi, lo); k = 0; t0 = addc64(t1, lo, k); hi += k; k = 0; t0 = addc64(t0, c, k); c = hi + k;
    // l[2] == 0
    k = 0; t1 = addc64(t2, c, k); c = k;
    mul64wide(m, l[3], hi, lo); k = 0; t2 = addc64(t3, lo, k); hi += k; k = 0; t2 = addc64(t2, c, k); c = hi + k;
    k = 0; t3 = addc64(t4, c, k); t4 = t5 + k;
  }
  scm r; r.v[0] = t0; r.v[1] = t1; r.v[2] = t2; r.v[3] = t3;
  sc_cond_sub_l(r.v, t4);
  sc_cond_sub_l(r.v, 0);
  return r;
}
#endif
HD scm sc_mul(const scm &a, const scm &b) { return sc_montmul(a, b); }  // both Montgomery -> Montgomery
HD scm sc_sqr(const scm &a) { return sc_montmul(a, a); }
HD scm sc_muladd(const scm &a, const scm &b, const scm &c) { return sc_add(sc_montmul(a, b), c); }

HD void load_le64x4(uint64_t w[4], const uint8_t *s) {
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint64_t x = 0;
#pragma unroll
    for (int j = 7; j >= 0; j--) x = (x << 8) | s[8 * i + j];
    w[i] = x;
  }
}
// any 32 bytes (reduced mod l) -> Montgomery
HD scm sc_from_bytes_mod_order(const uint8_t *s) { scm x; load_le64x4(x.v, s); return sc_montmul(x, sc_r2()); }
// 64 bytes, wide reduction -> Montgomery:  lo*R + hi*2^256*R = montmul(lo,R^2) + montmul(hi,R^3)
HD scm sc_from_bytes_wide(const uint8_t *s) {
  scm lo, hi; load_le64x4(lo.v, s); load_le64x4(hi.v, s + 32);
  return sc_add(sc_montmul(lo, sc_r2()), sc_montmul(hi, sc_r3()));
}
HD scm sc_from_words_wide(const uint64_t w[8]) {
  scm lo, hi;
  lo.v[0] = w[0]; lo.v[1] = w[1]; lo.v[2] = w[2]; lo.v[3] = w[3]; hi.v[0] = w[4]; hi.v[1] = w[5]; hi.v[2] = w[6]; hi.v[3] = w[7];
  return sc_add(sc_montmul(lo, sc_r2()), sc_montmul(hi, sc_r3()));
}
// Montgomery -> canonical integer limbs
HD void sc_to_canonical(uint64_t w[4], const scm &a) {
  scm one; one.v[0] = 1; one.v[1] = one.v[2] = one.v[3] = 0;
  scm r = sc_montmul(a, one);
  w[0] = r.v[0]; w[1] = r.v[1]; w[2] = r.v[2]; w[3] = r.v[3];
}
HD void sc_tobytes(uint8_t *s, const scm &a) {
  uint64_t w[4]; sc_to_canonical(w, a);
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) s[8 * i + j] = (uint8_t)(w[i] >> (8 * j));
}
// canonical check: returns 1 and sets r when the 32 bytes encode an integer < l
HD int sc_from_canonical_bytes(scm &r, const uint8_t *s) {
  scm x; load_le64x4(x.v, s);
  uint64_t l[4]; sc_const_l(l);
  uint64_t br = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) (void)subb64(x.v[i], l[i], br);
  r = sc_montmul(x, sc_r2());
  return br != 0;
}
HD scm sc_from_u64(uint64_t x) { scm a; a.v[0] = x; a.v[1] = a.v[2] = a.v[3] = 0; return sc_montmul(a, sc_r2()); }

// a^(l-2) (Fermat); invert(0) == 0 like Scalar::invert on zero in the reference's dependency
// (reference src/scalar_utils.rs:305-307 relies on it).  Kept as the cross-check of sc_invert.
HD scm sc_invert_fermat(const scm &a) {
  const uint64_t e[4] = SC_LM2_LIMBS;
  scm acc = a;  // bit 252
  for (int i = 251; i >= 0; i--) {
    acc = sc_sqr(acc);
    if ((e[i >> 6] >> (i & 63)) & 1) acc = sc_montmul(acc, a);
  }
  return acc;
}
// Inversion by the binary extended Euclidean algorithm (variable time; the witness is the prover's own secret and
// the reference's inversion is the only constant-time step we replace here -- see DESIGN.md).  About 380 iterations of
// 256-bit add/shift instead of 312 dependent Montgomery products: ~6x shorter dependency chain, which is what the
// one-thread-per-proof witness kernel is bound by.  Invariants: xA * a = A, xB * a = B (mod l).
HD scm sc_invert(const scm &am) {
  if (sc_is_zero(am)) return sc_zero();
  uint64_t l[4]; sc_const_l(l);
  uint64_t A[4] = {am.v[0], am.v[1], am.v[2], am.v[3]}, Bv[4] = {l[0], l[1], l[2], l[3]};
  uint64_t xA[4] = {1, 0, 0, 0}, xB[4] = {0, 0, 0, 0};
  for (int it = 0; it < 1024; it++) {
    const bool a1 = (A[0] == 1) & ((A[1] | A[2] | A[3]) == 0), b1 = (Bv[0] == 1) & ((Bv[1] | Bv[2] | Bv[3]) == 0);
    if (a1 | b1) break;
    const uint64_t aodd = A[0] & 1, bodd = Bv[0] & 1;
    // swap so that: both odd -> A >= B ; otherwise A is the even one
    uint64_t br = 0, t[4];
#pragma unroll
    for (int i = 0; i < 4; i++) t[i] = subb64(A[i], Bv[i], br);  // t = A - B, br = (A < B)
    const uint64_t both = aodd & bodd;
    const uint64_t sw = both ? br : aodd;  // both odd: swap when A < B; one even: swap when A is the odd one
    const uint64_t m = (uint64_t)0 - sw;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      uint64_t d = (A[i] ^ Bv[i]) & m; A[i] ^= d; Bv[i] ^= d;
      uint64_t e = (xA[i] ^ xB[i]) & m; xA[i] ^= e; xB[i] ^= e;
    }
    if (both) {
      // A -= B (A >= B now), xA -= xB (mod l)
      uint64_t b2 = 0;
#pragma unroll
      for (int i = 0; i < 4; i++) A[i] = subb64(A[i], Bv[i], b2);
      uint64_t b3 = 0, c3 = 0;
#pragma unroll
      for (int i = 0; i < 4; i++) xA[i] = subb64(xA[i], xB[i], b3);
      const uint64_t mm = (uint64_t)0 - (b3 != 0);
#pragma unroll
      for (int i = 0; i < 4; i++) xA[i] = addc64(xA[i], l[i] & mm, c3);
    }
    // A is even: halve A, halve xA modulo l
    A[0] = (A[0] >> 1) | (A[1] << 63); A[1] = (A[1] >> 1) | (A[2] << 63); A[2] = (A[2] >> 1) | (A[3] << 63); A[3] >>= 1;
    const uint64_t mo = (uint64_t)0 - (xA[0] & 1);
    uint64_t c4 = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) xA[i] = addc64(xA[i], l[i] & mo, c4);
    xA[0] = (xA[0] >> 1) | (xA[1] << 63); xA[1] = (xA[1] >> 1) | (xA[2] << 63); xA[2] = (xA[2] >> 1) | (xA[3] << 63); xA[3] = (xA[3] >> 1) | (c4 << 63);
  }
  const bool a1 = (A[0] == 1) & ((A[1] | A[2] | A[3]) == 0);
  scm r;
#pragma unroll
  for (int i = 0; i < 4; i++) r.v[i] = a1 ? xA[i] : xB[i];
  // r = (a R)^-1 as a plain integer; the Montgomery form of a^-1 is r * R^2 mod l = montmul(r, R^3)
  return sc_montmul(r, sc_r3());
}
HD scm sc_pow_u32(const scm &a, uint32_t e) {
  scm acc = sc_one(), base = a;
  while (e) { if (e & 1) acc = sc_montmul(acc, base); base = sc_sqr(base); e >>= 1; }
  return acc;
}
