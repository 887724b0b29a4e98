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
HD scm sc_montmul_any(const scm &a, const scm &b) {
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
HD scm sc_montmul_any(const scm &a, const scm &b) {
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
#if defined(__CUDA_ARCH__)
// Device form for REDUCED operands (a, b < l): word-serial Montgomery multiplication on two carry-save accumulators
// of 64-bit lanes at even / odd word positions, so that every row a[.] * b_i and every reduction row l[.] * m_i is a
// carry chain of fused IMAD.WIDE.U32 (mad.lo.cc / madc.hi.cc pairs) over non-overlapping lanes; the division by 2^32
// after each row is a swap of the two accumulators' roles.  l has zero words 4..6, so a reduction row is 5 products.
// 104 wide multiplies and ~100 adds instead of ~400 instructions of 64-bit emulation; the algorithm was checked against
// big-integer arithmetic in a Python model of the carry flag before being written here (tools/mont_sim.py).
#define SC_REDC_STEP(E, O)                                                                                                   \
  {                                                                                                                          \
    const uint32_t mi = E[0] * (uint32_t)SC_NINV;                                                                            \
    asm("mad.lo.cc.u32 %0, %8, %11, %0;\n\tmadc.hi.cc.u32 %1, %8, %11, %1;\n\t"                                              \
        "madc.lo.cc.u32 %2, %9, %11, %2;\n\tmadc.hi.cc.u32 %3, %9, %11, %3;\n\t"                                             \
        "addc.cc.u32 %4, %4, 0;\n\taddc.cc.u32 %5, %5, 0;\n\t"                                                               \
        "madc.lo.cc.u32 %6, %10, %11, %6;\n\tmadc.hi.u32 %7, %10, %11, %7;"                                                  \
        : "+r"(O[0]), "+r"(O[1]), "+r"(O[2]), "+r"(O[3]), "+r"(O[4]), "+r"(O[5]), "+r"(O[6]), "+r"(O[7])                     \
        : "r"(l1), "r"(l3), "r"(l7), "r"(mi));                                                                               \
    asm("mad.lo.cc.u32 %0, %9, %11, %0;\n\tmadc.hi.cc.u32 %1, %9, %11, %1;\n\t"                                              \
        "madc.lo.cc.u32 %2, %10, %11, %2;\n\tmadc.hi.cc.u32 %3, %10, %11, %3;\n\t"                                           \
        "addc.cc.u32 %4, %4, 0;\n\taddc.cc.u32 %5, %5, 0;\n\taddc.cc.u32 %6, %6, 0;\n\taddc.cc.u32 %7, %7, 0;\n\t"           \
        "addc.u32 %8, %8, 0;"                                                                                                \
        : "+r"(E[0]), "+r"(E[1]), "+r"(E[2]), "+r"(E[3]), "+r"(E[4]), "+r"(E[5]), "+r"(E[6]), "+r"(E[7]), "+r"(O[7])         \
        : "r"(l0), "r"(l2), "r"(mi));                                                                                        \
  }
// next row: E (lanes at words 0,2,4,6 after the shift) += a_even * bi; O is rebuilt two words lower and += a_odd * bi
#define SC_ROW_STEP(E, O, bi)                                                                                                \
  {                                                                                                                          \
    asm("add.cc.u32 %8, %8, %1;\n\t"                                                                                         \
        "madc.lo.cc.u32 %0, %9, %13, %2;\n\tmadc.hi.cc.u32 %1, %9, %13, %3;\n\t"                                             \
        "madc.lo.cc.u32 %2, %10, %13, %4;\n\tmadc.hi.cc.u32 %3, %10, %13, %5;\n\t"                                           \
        "madc.lo.cc.u32 %4, %11, %13, %6;\n\tmadc.hi.cc.u32 %5, %11, %13, %7;\n\t"                                           \
        "madc.lo.cc.u32 %6, %12, %13, 0;\n\tmadc.hi.u32 %7, %12, %13, 0;"                                                    \
        : "+r"(O[0]), "+r"(O[1]), "+r"(O[2]), "+r"(O[3]), "+r"(O[4]), "+r"(O[5]), "+r"(O[6]), "+r"(O[7]), "+r"(E[0])         \
        : "r"(aw[1]), "r"(aw[3]), "r"(aw[5]), "r"(aw[7]), "r"(bi));                                                          \
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\tmadc.hi.cc.u32 %1, %9, %13, %1;\n\t"                                              \
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\tmadc.hi.cc.u32 %3, %10, %13, %3;\n\t"                                           \
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\tmadc.hi.cc.u32 %5, %11, %13, %5;\n\t"                                           \
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\tmadc.hi.cc.u32 %7, %12, %13, %7;\n\t"                                           \
        "addc.u32 %8, %8, 0;"                                                                                                \
        : "+r"(E[0]), "+r"(E[1]), "+r"(E[2]), "+r"(E[3]), "+r"(E[4]), "+r"(E[5]), "+r"(E[6]), "+r"(E[7]), "+r"(O[7])         \
        : "r"(aw[0]), "r"(aw[2]), "r"(aw[4]), "r"(aw[6]), "r"(bi));                                                          \
  }
HD scm sc_montmul(const scm &a, const scm &b) {
  const uint64_t l64[4] = SC_L_LIMBS;
  const uint32_t l0 = (uint32_t)l64[0], l1 = (uint32_t)(l64[0] >> 32), l2 = (uint32_t)l64[1], l3 = (uint32_t)(l64[1] >> 32),
                 l7 = (uint32_t)(l64[3] >> 32);
  uint32_t aw[8], bw[8];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    aw[2 * i] = (uint32_t)a.v[i]; aw[2 * i + 1] = (uint32_t)(a.v[i] >> 32);
    bw[2 * i] = (uint32_t)b.v[i]; bw[2 * i + 1] = (uint32_t)(b.v[i] >> 32);
  }
  uint32_t X[8], Y[8];  // the two accumulators; their even/odd roles alternate from row to row
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const uint64_t pe = (uint64_t)aw[2 * k] * bw[0], po = (uint64_t)aw[2 * k + 1] * bw[0];
    X[2 * k] = (uint32_t)pe; X[2 * k + 1] = (uint32_t)(pe >> 32);
    Y[2 * k] = (uint32_t)po; Y[2 * k + 1] = (uint32_t)(po >> 32);
  }
  SC_REDC_STEP(X, Y)
  SC_ROW_STEP(Y, X, bw[1]) SC_REDC_STEP(Y, X)
  SC_ROW_STEP(X, Y, bw[2]) SC_REDC_STEP(X, Y)
  SC_ROW_STEP(Y, X, bw[3]) SC_REDC_STEP(Y, X)
  SC_ROW_STEP(X, Y, bw[4]) SC_REDC_STEP(X, Y)
  SC_ROW_STEP(Y, X, bw[5]) SC_REDC_STEP(Y, X)
  SC_ROW_STEP(X, Y, bw[6]) SC_REDC_STEP(X, Y)
  SC_ROW_STEP(Y, X, bw[7]) SC_REDC_STEP(Y, X)
  // even accumulator is Y (Y[0] == 0), odd is X: result words 1..8 = X[j] + Y[j + 1]
  uint32_t r[8];
  asm("add.cc.u32 %0, %8, %16;\n\taddc.cc.u32 %1, %9, %17;\n\taddc.cc.u32 %2, %10, %18;\n\taddc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\taddc.cc.u32 %5, %13, %21;\n\taddc.cc.u32 %6, %14, %22;\n\taddc.u32 %7, %15, 0;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(X[0]), "r"(X[1]), "r"(X[2]), "r"(X[3]), "r"(X[4]), "r"(X[5]), "r"(X[6]), "r"(X[7]),
        "r"(Y[1]), "r"(Y[2]), "r"(Y[3]), "r"(Y[4]), "r"(Y[5]), "r"(Y[6]), "r"(Y[7]));
  scm out;
#pragma unroll
  for (int i = 0; i < 4; i++) out.v[i] = (uint64_t)r[2 * i] | ((uint64_t)r[2 * i + 1] << 32);
  sc_cond_sub_l(out.v, 0);
  return out;
}
#else
HD scm sc_montmul(const scm &a, const scm &b) { return sc_montmul_any(a, b); }
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
HD scm sc_from_bytes_mod_order(const uint8_t *s) { scm x; load_le64x4(x.v, s); return sc_montmul_any(x, sc_r2()); }
// 64 bytes, wide reduction -> Montgomery:  lo*R + hi*2^256*R = montmul(lo,R^2) + montmul(hi,R^3)
HD scm sc_from_bytes_wide(const uint8_t *s) {
  scm lo, hi; load_le64x4(lo.v, s); load_le64x4(hi.v, s + 32);
  return sc_add(sc_montmul_any(lo, sc_r2()), sc_montmul_any(hi, sc_r3()));
}
HD scm sc_from_words_wide(const uint64_t w[8]) {
  scm lo, hi;
  lo.v[0] = w[0]; lo.v[1] = w[1]; lo.v[2] = w[2]; lo.v[3] = w[3]; hi.v[0] = w[4]; hi.v[1] = w[5]; hi.v[2] = w[6]; hi.v[3] = w[7];
  return sc_add(sc_montmul_any(lo, sc_r2()), sc_montmul_any(hi, sc_r3()));
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
  r = sc_montmul_any(x, sc_r2());
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
// Inversion by batched division steps (Bernstein-Yang "safegcd", variable-time form): 30 division steps at a time are
// decided on the low 32 bits of f, g alone and recorded as a 2x2 integer matrix, which is then applied to the full
// 9 x 30-bit signed-limb values (f, g) and, modulo l, to the Bezout coefficients (d, e).  About 19 batches of ~400
// instructions instead of ~380 iterations of a 256-bit add/shift binary GCD: the one-thread-per-proof witness kernel is
// bound by the length of exactly this dependency chain (188 sequential inversions per Poseidon permutation).
// Variable time: the witness is the prover's own secret and the reference's constant-time inversion is the step we
// replace -- see DESIGN.md.  Invariants: d * a = f, e * a = g (mod l); ends with g = 0, f = +-1.
struct sc_s30 { int32_t v[9]; };
#define SC_M30 0x3fffffff
HD int sc_ctz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __ffs((int)x) - 1;
#else
  return __builtin_ctz(x);
#endif
}
// 30 division steps on the low words; returns the new eta and the transition matrix t = (u v; q r) scaled by 2^30
HD int32_t sc_divsteps30(int32_t eta, uint32_t f0, uint32_t g0, int32_t t[4]) {
  uint32_t u = 1, v = 0, q = 0, r = 1, f = f0, g = g0;
  int i = 30;
  for (;;) {
    const int zeros = sc_ctz32(g | (0xffffffffu << i));
    g >>= zeros; u <<= zeros; v <<= zeros; eta -= zeros; i -= zeros;
    if (i == 0) break;
    if (eta < 0) {
      uint32_t tmp;
      eta = -eta;
      tmp = f; f = g; g = 0u - tmp;
      tmp = u; u = q; q = 0u - tmp;
      tmp = v; v = r; r = 0u - tmp;
    }
    // cancel up to min(eta + 1, i, 6) low bits of g with a multiple of f: w = -g / f mod 2^bits
    int limit = (eta + 1) > i ? i : (eta + 1);
    if (limit > 6) limit = 6;
    const uint32_t m = (1u << limit) - 1u;
    const uint32_t finv = f * (2u - f * f);  // f * f = 1 mod 8, one Newton step: inverse of f mod 64
    const uint32_t w = ((0u - g) * finv) & m;
    g += f * w; q += u * w; r += v * w;
  }
  t[0] = (int32_t)u; t[1] = (int32_t)v; t[2] = (int32_t)q; t[3] = (int32_t)r;
  return eta;
}
// (f, g) <- t (f, g) / 2^30   (exact)
HD void sc_update_fg30(sc_s30 &f, sc_s30 &g, const int32_t t[4]) {
  const int64_t u = t[0], v = t[1], q = t[2], r = t[3];
  int64_t cf = u * f.v[0] + v * g.v[0], cg = q * f.v[0] + r * g.v[0];
  cf >>= 30; cg >>= 30;
#pragma unroll
  for (int i = 1; i < 9; i++) {
    const int64_t fi = f.v[i], gi = g.v[i];
    cf += u * fi + v * gi; cg += q * fi + r * gi;
    f.v[i - 1] = (int32_t)cf & SC_M30; cf >>= 30;
    g.v[i - 1] = (int32_t)cg & SC_M30; cg >>= 30;
  }
  f.v[8] = (int32_t)cf; g.v[8] = (int32_t)cg;
}
// (d, e) <- t (d, e) / 2^30 mod l; d, e stay in (-2l, l)
HD void sc_update_de30(sc_s30 &d, sc_s30 &e, const int32_t t[4]) {
  const int32_t L30[9] = SC_L30_LIMBS;
  const int64_t u = t[0], v = t[1], q = t[2], r = t[3];
  const int32_t sd = d.v[8] >> 31, se = e.v[8] >> 31;
  int32_t md = (t[0] & sd) + (t[1] & se), me = (t[2] & sd) + (t[3] & se);
  int64_t cd = u * d.v[0] + v * e.v[0], ce = q * d.v[0] + r * e.v[0];
  md -= (int32_t)(((uint32_t)SC_LINV30 * (uint32_t)cd + (uint32_t)md) & SC_M30);
  me -= (int32_t)(((uint32_t)SC_LINV30 * (uint32_t)ce + (uint32_t)me) & SC_M30);
  cd += (int64_t)L30[0] * md; ce += (int64_t)L30[0] * me;
  cd >>= 30; ce >>= 30;
#pragma unroll
  for (int i = 1; i < 9; i++) {
    const int64_t di = d.v[i], ei = e.v[i];
    cd += u * di + v * ei; ce += q * di + r * ei;
    cd += (int64_t)L30[i] * md; ce += (int64_t)L30[i] * me;
    d.v[i - 1] = (int32_t)cd & SC_M30; cd >>= 30;
    e.v[i - 1] = (int32_t)ce & SC_M30; ce >>= 30;
  }
  d.v[8] = (int32_t)cd; e.v[8] = (int32_t)ce;
}
HD scm sc_invert(const scm &am) {
  if (sc_is_zero(am)) return sc_zero();
  const int32_t L30[9] = SC_L30_LIMBS;
  sc_s30 f, g, d, e;
#pragma unroll
  for (int i = 0; i < 9; i++) { f.v[i] = L30[i]; d.v[i] = 0; e.v[i] = 0; }
  e.v[0] = 1;
  // 4 x 64 -> 9 x 30
#pragma unroll
  for (int i = 0; i < 9; i++) {
    const int bit = 30 * i, w = bit >> 6, off = bit & 63;
    uint64_t x = am.v[w] >> off;
    if (off > 34 && w + 1 < 4) x |= am.v[w + 1] << (64 - off);
    g.v[i] = (int32_t)(x & SC_M30);
  }
  int32_t eta = -1;
  for (int it = 0; it < 40; it++) {
    int32_t t[4];
    eta = sc_divsteps30(eta, (uint32_t)f.v[0], (uint32_t)g.v[0], t);
    sc_update_de30(d, e, t);
    sc_update_fg30(f, g, t);
    int32_t nz = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) nz |= g.v[i];
    if (nz == 0) break;
  }
  // f = +-1: the inverse is sign(f) * d; normalise from (-2l, l) into [0, l)
  const int32_t sneg = f.v[8] >> 31;
  int32_t add = d.v[8] >> 31;
#pragma unroll
  for (int i = 0; i < 9; i++) d.v[i] = ((d.v[i] + (L30[i] & add)) ^ sneg) - sneg;
#pragma unroll
  for (int i = 0; i < 8; i++) { d.v[i + 1] += d.v[i] >> 30; d.v[i] &= SC_M30; }
  add = d.v[8] >> 31;
#pragma unroll
  for (int i = 0; i < 9; i++) d.v[i] += L30[i] & add;
#pragma unroll
  for (int i = 0; i < 8; i++) { d.v[i + 1] += d.v[i] >> 30; d.v[i] &= SC_M30; }
  // 9 x 30 -> 4 x 64
  scm rr;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    uint64_t acc = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) {
      const int lo = 30 * i - 64 * k;  // bit position of limb i relative to word k
      if (lo >= 64 || lo + 30 <= 0) continue;
      const uint64_t x = (uint64_t)(uint32_t)d.v[i];
      acc |= lo >= 0 ? (x << lo) : (x >> (-lo));
    }
    rr.v[k] = acc;
  }
  // rr = (a R)^-1 as a plain integer; the Montgomery form of a^-1 is rr * R^2 mod l = montmul(rr, R^3)
  return sc_montmul(rr, sc_r3());
}
HD scm sc_pow_u32(const scm &a, uint32_t e) {
  scm acc = sc_one(), base = a;
  while (e) { if (e & 1) acc = sc_montmul(acc, base); base = sc_sqr(base); e >>= 1; }
  return acc;
}
