// GF(2^255-19) arithmetic, 8 saturated 32-bit limbs: an element is ANY 256-bit integer congruent to the
// value mod p (2^256 = 38 mod p), so there are no limb bounds to track -- every function accepts and
// returns arbitrary 256-bit representatives, and only fe_tobytes reduces fully.
//
// On sm_100a a limb product is one IMAD.WIDE.U32 with carry-in/carry-out predicates: ptxas fuses each
// mad.lo.cc / madc.hi.cc pair below into IMAD.WIDE.U32[.X], so the 8x8 schoolbook product is 64 of them
// plus 8 for the 2^256 = 38 fold (72 fma-pipe instructions, ~125 in total; the radix-2^25.5 form this
// replaces needed 100 IMAD.WIDE with a 64-bit addend plus ~100 carry/shift instructions).  Products of
// equal column parity go to one of two accumulators (even-/odd-aligned 64-bit lanes) so that a row of four
// products is one carry chain with no overlap between neighbouring lanes.
//
// Replaces curve25519-dalek's FieldElement (reference Cargo.toml:8, an un-vendored dependency);
// representation and schedule are our own.  The host (emulation-test) build uses plain 64-bit C.
#pragma once
#include "hd.h"
#include "constants.h"

struct fe { uint32_t v[8]; };

HD void fe_0(fe &h) {
#pragma unroll
  for (int i = 0; i < 8; i++) h.v[i] = 0;
}
HD void fe_1(fe &h) { fe_0(h); h.v[0] = 1; }

// h = f + g
HD void fe_add(fe &h, const fe &f, const fe &g) {
#if defined(__CUDA_ARCH__)
  uint32_t r0, r1, r2, r3, r4, r5, r6, r7, c;
  asm("add.cc.u32 %0, %9, %17;\n\taddc.cc.u32 %1, %10, %18;\n\taddc.cc.u32 %2, %11, %19;\n\taddc.cc.u32 %3, %12, %20;\n\t"
      "addc.cc.u32 %4, %13, %21;\n\taddc.cc.u32 %5, %14, %22;\n\taddc.cc.u32 %6, %15, %23;\n\taddc.cc.u32 %7, %16, %24;\n\t"
      "addc.u32 %8, 0, 0;"
      : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7), "=r"(c)
      : "r"(f.v[0]), "r"(f.v[1]), "r"(f.v[2]), "r"(f.v[3]), "r"(f.v[4]), "r"(f.v[5]), "r"(f.v[6]), "r"(f.v[7]),
        "r"(g.v[0]), "r"(g.v[1]), "r"(g.v[2]), "r"(g.v[3]), "r"(g.v[4]), "r"(g.v[5]), "r"(g.v[6]), "r"(g.v[7]));
  uint32_t t = c * 38u;
  asm("add.cc.u32 %0, %0, %9;\n\taddc.cc.u32 %1, %1, 0;\n\taddc.cc.u32 %2, %2, 0;\n\taddc.cc.u32 %3, %3, 0;\n\t"
      "addc.cc.u32 %4, %4, 0;\n\taddc.cc.u32 %5, %5, 0;\n\taddc.cc.u32 %6, %6, 0;\n\taddc.cc.u32 %7, %7, 0;\n\t"
      "addc.u32 %8, 0, 0;"
      : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7), "=r"(c)
      : "r"(t));
  h.v[0] = r0 + c * 38u;  // a second wrap leaves a value below 38: no further carry
  h.v[1] = r1; h.v[2] = r2; h.v[3] = r3; h.v[4] = r4; h.v[5] = r5; h.v[6] = r6; h.v[7] = r7;
#else
  uint32_t r[8]; uint64_t c = 0;
  for (int i = 0; i < 8; i++) { c += (uint64_t)f.v[i] + g.v[i]; r[i] = (uint32_t)c; c >>= 32; }
  c *= 38;
  for (int i = 0; i < 8; i++) { c += r[i]; r[i] = (uint32_t)c; c >>= 32; }
  r[0] += (uint32_t)c * 38u;
  for (int i = 0; i < 8; i++) h.v[i] = r[i];
#endif
}
// h = f - g
HD void fe_sub(fe &h, const fe &f, const fe &g) {
#if defined(__CUDA_ARCH__)
  uint32_t r0, r1, r2, r3, r4, r5, r6, r7, b;
  asm("sub.cc.u32 %0, %9, %17;\n\tsubc.cc.u32 %1, %10, %18;\n\tsubc.cc.u32 %2, %11, %19;\n\tsubc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\tsubc.cc.u32 %5, %14, %22;\n\tsubc.cc.u32 %6, %15, %23;\n\tsubc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;"
      : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7), "=r"(b)
      : "r"(f.v[0]), "r"(f.v[1]), "r"(f.v[2]), "r"(f.v[3]), "r"(f.v[4]), "r"(f.v[5]), "r"(f.v[6]), "r"(f.v[7]),
        "r"(g.v[0]), "r"(g.v[1]), "r"(g.v[2]), "r"(g.v[3]), "r"(g.v[4]), "r"(g.v[5]), "r"(g.v[6]), "r"(g.v[7]));
  uint32_t t = b & 38u;  // b = 0xffffffff on borrow: the wrapped value is 2^256 too large, i.e. 38 too large mod p
  asm("sub.cc.u32 %0, %0, %9;\n\tsubc.cc.u32 %1, %1, 0;\n\tsubc.cc.u32 %2, %2, 0;\n\tsubc.cc.u32 %3, %3, 0;\n\t"
      "subc.cc.u32 %4, %4, 0;\n\tsubc.cc.u32 %5, %5, 0;\n\tsubc.cc.u32 %6, %6, 0;\n\tsubc.cc.u32 %7, %7, 0;\n\t"
      "subc.u32 %8, 0, 0;"
      : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7), "=r"(b)
      : "r"(t));
  h.v[0] = r0 - (b & 38u);  // a second wrap leaves a value within 38 of 2^256: no further borrow
  h.v[1] = r1; h.v[2] = r2; h.v[3] = r3; h.v[4] = r4; h.v[5] = r5; h.v[6] = r6; h.v[7] = r7;
#else
  uint32_t r[8]; int64_t c = 0;
  for (int i = 0; i < 8; i++) { c += (int64_t)f.v[i] - (int64_t)g.v[i]; r[i] = (uint32_t)c; c >>= 32; }
  int64_t t = c ? 38 : 0; c = 0;
  for (int i = 0; i < 8; i++) { c += (int64_t)r[i] - (i == 0 ? t : 0); r[i] = (uint32_t)c; c >>= 32; }
  if (c) r[0] -= 38u;
  for (int i = 0; i < 8; i++) h.v[i] = r[i];
#endif
}
HD void fe_neg(fe &h, const fe &f) { fe z; fe_0(z); fe_sub(h, z, f); }
// h = b ? g : f   (branch-free)
HD void fe_select(fe &h, const fe &f, const fe &g, int b) {
  uint32_t m = (uint32_t)(-(int32_t)(b != 0));
#pragma unroll
  for (int i = 0; i < 8; i++) h.v[i] = f.v[i] ^ (m & (f.v[i] ^ g.v[i]));
}
HD void fe_cswap(fe &f, fe &g, int b) {
  uint32_t m = (uint32_t)(-(int32_t)(b != 0));
#pragma unroll
  for (int i = 0; i < 8; i++) { uint32_t x = m & (f.v[i] ^ g.v[i]); f.v[i] ^= x; g.v[i] ^= x; }
}
// h = b ? -f : f
HD void fe_cneg(fe &h, const fe &f, int b) { fe n; fe_neg(n, f); fe_select(h, f, n, b); }

#if defined(__CUDA_ARCH__)
// one carry chain of four 32x32 products into four neighbouring 64-bit lanes, carry-out into the word above
#define FE_ROW4(A, s, x0, x1, x2, x3, y)                                                                                     \
  asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\tmadc.hi.cc.u32 %1, %9, %13, %1;\n\t"                                                \
      "madc.lo.cc.u32 %2, %10, %13, %2;\n\tmadc.hi.cc.u32 %3, %10, %13, %3;\n\t"                                             \
      "madc.lo.cc.u32 %4, %11, %13, %4;\n\tmadc.hi.cc.u32 %5, %11, %13, %5;\n\t"                                             \
      "madc.lo.cc.u32 %6, %12, %13, %6;\n\tmadc.hi.cc.u32 %7, %12, %13, %7;\n\t"                                             \
      "addc.u32 %8, %8, 0;"                                                                                                  \
      : "+r"(A[s]), "+r"(A[s + 1]), "+r"(A[s + 2]), "+r"(A[s + 3]), "+r"(A[s + 4]), "+r"(A[s + 5]), "+r"(A[s + 6]),          \
        "+r"(A[s + 7]), "+r"(A[s + 8])                                                                                       \
      : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y))

// 2^256 = 38: R (8 words) = W[0..7] + 38 * W[8..15], fully folded back into 256 bits
HD void fe_fold512(fe &out, const uint32_t W[16]) {
  uint32_t R[9];
#pragma unroll
  for (int i = 0; i < 8; i++) R[i] = W[i];
  R[8] = 0;
  const uint32_t c38 = 38;
  FE_ROW4(R, 0, W[8], W[10], W[12], W[14], c38);
  asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\tmadc.hi.cc.u32 %1, %8, %12, %1;\n\t"
      "madc.lo.cc.u32 %2, %9, %12, %2;\n\tmadc.hi.cc.u32 %3, %9, %12, %3;\n\t"
      "madc.lo.cc.u32 %4, %10, %12, %4;\n\tmadc.hi.cc.u32 %5, %10, %12, %5;\n\t"
      "madc.lo.cc.u32 %6, %11, %12, %6;\n\tmadc.hi.u32 %7, %11, %12, %7;"
      : "+r"(R[1]), "+r"(R[2]), "+r"(R[3]), "+r"(R[4]), "+r"(R[5]), "+r"(R[6]), "+r"(R[7]), "+r"(R[8])
      : "r"(W[9]), "r"(W[11]), "r"(W[13]), "r"(W[15]), "r"(c38));
  uint32_t t = R[8] * 38u, c;  // R[8] <= 39
  asm("add.cc.u32 %0, %0, %9;\n\taddc.cc.u32 %1, %1, 0;\n\taddc.cc.u32 %2, %2, 0;\n\taddc.cc.u32 %3, %3, 0;\n\t"
      "addc.cc.u32 %4, %4, 0;\n\taddc.cc.u32 %5, %5, 0;\n\taddc.cc.u32 %6, %6, 0;\n\taddc.cc.u32 %7, %7, 0;\n\t"
      "addc.u32 %8, 0, 0;"
      : "+r"(R[0]), "+r"(R[1]), "+r"(R[2]), "+r"(R[3]), "+r"(R[4]), "+r"(R[5]), "+r"(R[6]), "+r"(R[7]), "=r"(c)
      : "r"(t));
  R[0] += c * 38u;
#pragma unroll
  for (int i = 0; i < 8; i++) out.v[i] = R[i];
}
#endif

#if defined(__CUDA_ARCH__)
// shorter carry chains (1-3 lanes) for the triangular cross products of a squaring
#define FE_ROW3(A, s, x0, x1, x2, y)                                                                                         \
  asm("mad.lo.cc.u32 %0, %7, %10, %0;\n\tmadc.hi.cc.u32 %1, %7, %10, %1;\n\t"                                                \
      "madc.lo.cc.u32 %2, %8, %10, %2;\n\tmadc.hi.cc.u32 %3, %8, %10, %3;\n\t"                                               \
      "madc.lo.cc.u32 %4, %9, %10, %4;\n\tmadc.hi.cc.u32 %5, %9, %10, %5;\n\t"                                               \
      "addc.u32 %6, %6, 0;"                                                                                                  \
      : "+r"(A[s]), "+r"(A[s + 1]), "+r"(A[s + 2]), "+r"(A[s + 3]), "+r"(A[s + 4]), "+r"(A[s + 5]), "+r"(A[s + 6])           \
      : "r"(x0), "r"(x1), "r"(x2), "r"(y))
#define FE_ROW2(A, s, x0, x1, y)                                                                                             \
  asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\tmadc.hi.cc.u32 %1, %5, %7, %1;\n\t"                                                  \
      "madc.lo.cc.u32 %2, %6, %7, %2;\n\tmadc.hi.cc.u32 %3, %6, %7, %3;\n\t"                                                 \
      "addc.u32 %4, %4, 0;"                                                                                                  \
      : "+r"(A[s]), "+r"(A[s + 1]), "+r"(A[s + 2]), "+r"(A[s + 3]), "+r"(A[s + 4])                                           \
      : "r"(x0), "r"(x1), "r"(y))
#define FE_ROW1(A, s, x0, y)                                                                                                 \
  asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"                                \
      : "+r"(A[s]), "+r"(A[s + 1]), "+r"(A[s + 2])                                                                           \
      : "r"(x0), "r"(y))
#endif

#ifndef BP_FE_KARATSUBA
#define BP_FE_KARATSUBA 0
#endif
#if defined(__CUDA_ARCH__)
// 4x4 words -> 8 words, same even/odd lane scheme as the 8x8 product (16 wide multiplies)
HD void fe_mul4(uint32_t r[8], const uint32_t a[4], const uint32_t b[4]) {
  uint32_t E[9], O[9];
#pragma unroll
  for (int i = 0; i < 9; i++) { E[i] = 0; O[i] = 0; }
  FE_ROW2(E, 0, a[0], a[2], b[0]); FE_ROW2(O, 1, a[1], a[3], b[0]);
  FE_ROW2(E, 2, a[1], a[3], b[1]); FE_ROW2(O, 1, a[0], a[2], b[1]);
  FE_ROW2(E, 2, a[0], a[2], b[2]); FE_ROW2(O, 3, a[1], a[3], b[2]);
  FE_ROW2(E, 4, a[1], a[3], b[3]); FE_ROW2(O, 3, a[0], a[2], b[3]);
  r[0] = E[0];
  asm("add.cc.u32 %0, %7, %14;\n\taddc.cc.u32 %1, %8, %15;\n\taddc.cc.u32 %2, %9, %16;\n\taddc.cc.u32 %3, %10, %17;\n\t"
      "addc.cc.u32 %4, %11, %18;\n\taddc.cc.u32 %5, %12, %19;\n\taddc.u32 %6, %13, %20;"
      : "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]),
        "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]));
}
// d = |x - y| (4 words), m = all-ones when x < y
HD void fe_absdiff4(uint32_t d[4], uint32_t &m, const uint32_t x[4], const uint32_t y[4]) {
  uint32_t t0, t1, t2, t3;
  asm("sub.cc.u32 %0, %5, %9;\n\tsubc.cc.u32 %1, %6, %10;\n\tsubc.cc.u32 %2, %7, %11;\n\tsubc.cc.u32 %3, %8, %12;\n\tsubc.u32 %4, 0, 0;"
      : "=r"(t0), "=r"(t1), "=r"(t2), "=r"(t3), "=r"(m)
      : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]));
  t0 ^= m; t1 ^= m; t2 ^= m; t3 ^= m;
  const uint32_t one = m & 1u;
  asm("add.cc.u32 %0, %4, %8;\n\taddc.cc.u32 %1, %5, 0;\n\taddc.cc.u32 %2, %6, 0;\n\taddc.u32 %3, %7, 0;"
      : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]) : "r"(t0), "r"(t1), "r"(t2), "r"(t3), "r"(one));
}
// one level of (subtractive) Karatsuba: 3 x 16 wide multiplies instead of 64.  With z0 = f_lo g_lo, z2 = f_hi g_hi and
// zm = |f_lo - f_hi| |g_lo - g_hi|, the middle term f_lo g_hi + f_hi g_lo is z0 + z2 -/+ zm (sign from the two differences).
HD void fe_mul_kara(fe &out, const fe &f, const fe &g) {
  uint32_t z0[8], z2[8], zm[8], da[4], db[4], sa, sb;
  fe_mul4(z0, f.v, g.v);
  fe_mul4(z2, f.v + 4, g.v + 4);
  fe_absdiff4(da, sa, f.v, f.v + 4);
  fe_absdiff4(db, sb, g.v, g.v + 4);
  fe_mul4(zm, da, db);
  uint32_t t[9];
  asm("add.cc.u32 %0, %9, %17;\n\taddc.cc.u32 %1, %10, %18;\n\taddc.cc.u32 %2, %11, %19;\n\taddc.cc.u32 %3, %12, %20;\n\t"
      "addc.cc.u32 %4, %13, %21;\n\taddc.cc.u32 %5, %14, %22;\n\taddc.cc.u32 %6, %15, %23;\n\taddc.cc.u32 %7, %16, %24;\n\t"
      "addc.u32 %8, 0, 0;"
      : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(t[8])
      : "r"(z0[0]), "r"(z0[1]), "r"(z0[2]), "r"(z0[3]), "r"(z0[4]), "r"(z0[5]), "r"(z0[6]), "r"(z0[7]),
        "r"(z2[0]), "r"(z2[1]), "r"(z2[2]), "r"(z2[3]), "r"(z2[4]), "r"(z2[5]), "r"(z2[6]), "r"(z2[7]));
  // mid = t - zm when the differences have equal signs (M = ~0: add the complement plus one), t + zm otherwise (M = 0)
  const uint32_t M = ~(sa ^ sb);
  uint32_t x[8], mid[9], dummy = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = zm[i] ^ M;
  asm("add.cc.u32 %9, %27, 1;\n\t"
      "addc.cc.u32 %0, %10, %19;\n\taddc.cc.u32 %1, %11, %20;\n\taddc.cc.u32 %2, %12, %21;\n\taddc.cc.u32 %3, %13, %22;\n\t"
      "addc.cc.u32 %4, %14, %23;\n\taddc.cc.u32 %5, %15, %24;\n\taddc.cc.u32 %6, %16, %25;\n\taddc.cc.u32 %7, %17, %26;\n\t"
      "addc.u32 %8, %18, %27;"
      : "=r"(mid[0]), "=r"(mid[1]), "=r"(mid[2]), "=r"(mid[3]), "=r"(mid[4]), "=r"(mid[5]), "=r"(mid[6]), "=r"(mid[7]), "=r"(mid[8]), "=r"(dummy)
      : "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]), "r"(t[8]),
        "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]), "r"(M));
  (void)dummy;
  uint32_t W[16];
#pragma unroll
  for (int i = 0; i < 8; i++) { W[i] = z0[i]; W[8 + i] = z2[i]; }
  asm("add.cc.u32 %0, %0, %12;\n\taddc.cc.u32 %1, %1, %13;\n\taddc.cc.u32 %2, %2, %14;\n\taddc.cc.u32 %3, %3, %15;\n\t"
      "addc.cc.u32 %4, %4, %16;\n\taddc.cc.u32 %5, %5, %17;\n\taddc.cc.u32 %6, %6, %18;\n\taddc.cc.u32 %7, %7, %19;\n\t"
      "addc.cc.u32 %8, %8, %20;\n\taddc.cc.u32 %9, %9, 0;\n\taddc.cc.u32 %10, %10, 0;\n\taddc.u32 %11, %11, 0;"
      : "+r"(W[4]), "+r"(W[5]), "+r"(W[6]), "+r"(W[7]), "+r"(W[8]), "+r"(W[9]), "+r"(W[10]), "+r"(W[11]), "+r"(W[12]), "+r"(W[13]),
        "+r"(W[14]), "+r"(W[15])
      : "r"(mid[0]), "r"(mid[1]), "r"(mid[2]), "r"(mid[3]), "r"(mid[4]), "r"(mid[5]), "r"(mid[6]), "r"(mid[7]), "r"(mid[8]));
  fe_fold512(out, W);
}
#endif

HD void fe_mul_inl(fe &out, const fe &f, const fe &g) {
#if defined(__CUDA_ARCH__)
#if BP_FE_KARATSUBA
  fe_mul_kara(out, f, g);
  return;
#endif
  uint32_t E[17], O[17];  // word w of the even-/odd-aligned accumulator
#pragma unroll
  for (int i = 0; i < 17; i++) { E[i] = 0; O[i] = 0; }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int p = i & 1, q = 1 - p;
    FE_ROW4(E, i + p, f.v[p], f.v[p + 2], f.v[p + 4], f.v[p + 6], g.v[i]);  // j = p, p+2, ..: i + j even
    FE_ROW4(O, i + q, f.v[q], f.v[q + 2], f.v[q + 4], f.v[q + 6], g.v[i]);  // i + j odd
  }
  uint32_t W[16];
  W[0] = E[0];
  uint32_t cm, dummy = 0;  // merge in two carry chains (asm operand limit); cm carries between them
  asm("add.cc.u32 %0, %8, %15;\n\taddc.cc.u32 %1, %9, %16;\n\taddc.cc.u32 %2, %10, %17;\n\taddc.cc.u32 %3, %11, %18;\n\t"
      "addc.cc.u32 %4, %12, %19;\n\taddc.cc.u32 %5, %13, %20;\n\taddc.cc.u32 %6, %14, %21;\n\taddc.u32 %7, 0, 0;"
      : "=r"(W[1]), "=r"(W[2]), "=r"(W[3]), "=r"(W[4]), "=r"(W[5]), "=r"(W[6]), "=r"(W[7]), "=r"(cm)
      : "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]),
        "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]));
  asm("add.cc.u32 %8, %25, 0xffffffff;\n\t"
      "addc.cc.u32 %0, %9, %17;\n\taddc.cc.u32 %1, %10, %18;\n\taddc.cc.u32 %2, %11, %19;\n\taddc.cc.u32 %3, %12, %20;\n\t"
      "addc.cc.u32 %4, %13, %21;\n\taddc.cc.u32 %5, %14, %22;\n\taddc.cc.u32 %6, %15, %23;\n\taddc.u32 %7, %16, %24;"
      : "=r"(W[8]), "=r"(W[9]), "=r"(W[10]), "=r"(W[11]), "=r"(W[12]), "=r"(W[13]), "=r"(W[14]), "=r"(W[15]), "=r"(dummy)
      : "r"(E[8]), "r"(E[9]), "r"(E[10]), "r"(E[11]), "r"(E[12]), "r"(E[13]), "r"(E[14]), "r"(E[15]),
        "r"(O[8]), "r"(O[9]), "r"(O[10]), "r"(O[11]), "r"(O[12]), "r"(O[13]), "r"(O[14]), "r"(O[15]), "r"(cm));
  (void)dummy;
  fe_fold512(out, W);
#else
  uint32_t W[16];
  uint64_t acc = 0, hi = 0;  // 96-bit column accumulator (hi counts 2^64)
  for (int k = 0; k < 15; k++) {
    for (int i = 0; i < 8; i++) {
      int j = k - i;
      if (j < 0 || j > 7) continue;
      uint64_t pr = (uint64_t)f.v[i] * g.v[j];
      acc += pr; if (acc < pr) hi++;
    }
    W[k] = (uint32_t)acc;
    acc = (acc >> 32) | (hi << 32); hi = 0;
  }
  W[15] = (uint32_t)acc;
  // R = lo + 38 * hi
  uint32_t R[8]; uint64_t c = 0;
  for (int i = 0; i < 8; i++) { c += (uint64_t)W[i] + 38ull * W[8 + i]; R[i] = (uint32_t)c; c >>= 32; }
  c *= 38;
  for (int i = 0; i < 8; i++) { c += R[i]; R[i] = (uint32_t)c; c >>= 32; }
  R[0] += (uint32_t)c * 38u;
  for (int i = 0; i < 8; i++) out.v[i] = R[i];
#endif
}


// squaring: 28 cross products (doubled by a 1-bit shift of the merged accumulators) + 8 squares = 36 wide multiplies
// instead of 64; rows are processed in increasing i so that a chain's carry-out always lands above every lane used so far
HD void fe_sq_inl(fe &out, const fe &f) {
#if defined(__CUDA_ARCH__)
  uint32_t E[17], O[17];
#pragma unroll
  for (int i = 0; i < 17; i++) { E[i] = 0; O[i] = 0; }
  const uint32_t *a = f.v;
  FE_ROW4(O, 1, a[1], a[3], a[5], a[7], a[0]);  FE_ROW3(E, 2, a[2], a[4], a[6], a[0]);
  FE_ROW3(O, 3, a[2], a[4], a[6], a[1]);        FE_ROW3(E, 4, a[3], a[5], a[7], a[1]);
  FE_ROW3(O, 5, a[3], a[5], a[7], a[2]);        FE_ROW2(E, 6, a[4], a[6], a[2]);
  FE_ROW2(O, 7, a[4], a[6], a[3]);              FE_ROW2(E, 8, a[5], a[7], a[3]);
  FE_ROW2(O, 9, a[5], a[7], a[4]);              FE_ROW1(E, 10, a[6], a[4]);
  FE_ROW1(O, 11, a[6], a[5]);                   FE_ROW1(E, 12, a[7], a[5]);
  FE_ROW1(O, 13, a[7], a[6]);
  // C = E + O (words 1..15; word 0 of the cross sum is empty), then W = 2C + sum a_i^2 2^(64 i)
  uint32_t C[16];
  C[0] = 0;
  uint32_t cm, dummy = 0;
  asm("add.cc.u32 %0, %8, %15;\n\taddc.cc.u32 %1, %9, %16;\n\taddc.cc.u32 %2, %10, %17;\n\taddc.cc.u32 %3, %11, %18;\n\t"
      "addc.cc.u32 %4, %12, %19;\n\taddc.cc.u32 %5, %13, %20;\n\taddc.cc.u32 %6, %14, %21;\n\taddc.u32 %7, 0, 0;"
      : "=r"(C[1]), "=r"(C[2]), "=r"(C[3]), "=r"(C[4]), "=r"(C[5]), "=r"(C[6]), "=r"(C[7]), "=r"(cm)
      : "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]),
        "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]));
  asm("add.cc.u32 %8, %25, 0xffffffff;\n\t"
      "addc.cc.u32 %0, %9, %17;\n\taddc.cc.u32 %1, %10, %18;\n\taddc.cc.u32 %2, %11, %19;\n\taddc.cc.u32 %3, %12, %20;\n\t"
      "addc.cc.u32 %4, %13, %21;\n\taddc.cc.u32 %5, %14, %22;\n\taddc.cc.u32 %6, %15, %23;\n\taddc.u32 %7, %16, %24;"
      : "=r"(C[8]), "=r"(C[9]), "=r"(C[10]), "=r"(C[11]), "=r"(C[12]), "=r"(C[13]), "=r"(C[14]), "=r"(C[15]), "=r"(dummy)
      : "r"(E[8]), "r"(E[9]), "r"(E[10]), "r"(E[11]), "r"(E[12]), "r"(E[13]), "r"(E[14]), "r"(E[15]),
        "r"(O[8]), "r"(O[9]), "r"(O[10]), "r"(O[11]), "r"(O[12]), "r"(O[13]), "r"(O[14]), "r"(O[15]), "r"(cm));
  (void)dummy;
  uint32_t W[16];
#pragma unroll
  for (int i = 15; i >= 1; i--) W[i] = __funnelshift_l(C[i - 1], C[i], 1);
  W[0] = 0;
  asm("mad.lo.cc.u32 %0, %16, %16, %0;\n\tmadc.hi.cc.u32 %1, %16, %16, %1;\n\t"
      "madc.lo.cc.u32 %2, %17, %17, %2;\n\tmadc.hi.cc.u32 %3, %17, %17, %3;\n\t"
      "madc.lo.cc.u32 %4, %18, %18, %4;\n\tmadc.hi.cc.u32 %5, %18, %18, %5;\n\t"
      "madc.lo.cc.u32 %6, %19, %19, %6;\n\tmadc.hi.cc.u32 %7, %19, %19, %7;\n\t"
      "madc.lo.cc.u32 %8, %20, %20, %8;\n\tmadc.hi.cc.u32 %9, %20, %20, %9;\n\t"
      "madc.lo.cc.u32 %10, %21, %21, %10;\n\tmadc.hi.cc.u32 %11, %21, %21, %11;\n\t"
      "madc.lo.cc.u32 %12, %22, %22, %12;\n\tmadc.hi.cc.u32 %13, %22, %22, %13;\n\t"
      "madc.lo.cc.u32 %14, %23, %23, %14;\n\tmadc.hi.u32 %15, %23, %23, %15;"
      : "+r"(W[0]), "+r"(W[1]), "+r"(W[2]), "+r"(W[3]), "+r"(W[4]), "+r"(W[5]), "+r"(W[6]), "+r"(W[7]), "+r"(W[8]), "+r"(W[9]),
        "+r"(W[10]), "+r"(W[11]), "+r"(W[12]), "+r"(W[13]), "+r"(W[14]), "+r"(W[15])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]));
  fe_fold512(out, W);
#else
  fe_mul_inl(out, f, f);
#endif
}
// out = 2 f^2
HD void fe_sq2_inl(fe &out, const fe &f) { fe t; fe_sq_inl(t, f); fe_add(out, t, t); }

// On the device the multiplication primitives are real functions (operands and result travel in registers): a point
// addition is then ~10 calls instead of ~1000 inlined instructions, which keeps the hot loops inside the instruction cache.
#ifndef BP_FE_CALL
#define BP_FE_CALL 1
#endif
#if defined(__CUDACC__) && BP_FE_CALL
static __device__ __noinline__ fe fe_mul_fn(fe f, fe g) { fe h; fe_mul_inl(h, f, g); return h; }
static __device__ __noinline__ fe fe_sq_fn(fe f) { fe h; fe_sq_inl(h, f); return h; }
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
HD void fe_sq2(fe &out, const fe &f) { fe t; fe_sq(t, f); fe_add(out, t, t); }
// Call-or-inline choice per use.  A call moves its 16 + 8 operand registers with IMAD.MOV -- which issues on the same
// multiplier pipe as the IMAD.WIDE of the product itself -- so the point operations of the hot loops expand the field
// multiplications in place (INL = true: +16 % mixed additions per second, profiles/r02_field_call_vs_inline.jsonl) while
// everything with many multiplication sites (inversions, encodings, set-up) keeps the calls and a small instruction footprint.
#ifndef BP_GE_INLINE
#define BP_GE_INLINE 1  // 0: every multiplication is a call (the round-1 build), for A/B runs
#endif
template <bool INL> HD void fe_mul_x(fe &out, const fe &f, const fe &g) {
  if constexpr (INL && BP_GE_INLINE) fe_mul_inl(out, f, g); else fe_mul(out, f, g);
}
template <bool INL> HD void fe_sq_x(fe &out, const fe &f) {
  if constexpr (INL && BP_GE_INLINE) fe_sq_inl(out, f); else fe_sq(out, f);
}
template <bool INL> HD void fe_sq2_x(fe &out, const fe &f) { fe t; fe_sq_x<INL>(t, f); fe_add(out, t, t); }
HD void fe_sqn(fe &out, const fe &f, int n) {
  fe_sq(out, f);
  for (int i = 1; i < n; i++) fe_sq(out, out);
}

// little-endian bytes -> element; bit 255 is ignored (RFC 7748/8032 convention, as the 10-limb form did)
HD void fe_frombytes(fe &out, const uint8_t *s) {
#pragma unroll
  for (int i = 0; i < 8; i++)
    out.v[i] = (uint32_t)s[4 * i] | ((uint32_t)s[4 * i + 1] << 8) | ((uint32_t)s[4 * i + 2] << 16) | ((uint32_t)s[4 * i + 3] << 24);
  out.v[7] &= 0x7fffffffu;
}

// fully reduced words (value in [0, p))
HD void fe_reduce_words(uint32_t r[8], const fe &f) {
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = f.v[i];
  // twice: fold bit 255 (2^255 = 19); afterwards r < 2^255
#pragma unroll
  for (int round = 0; round < 2; round++) {
    uint64_t c = 19ull * (r[7] >> 31);
    r[7] &= 0x7fffffffu;
#pragma unroll
    for (int i = 0; i < 8; i++) { c += r[i]; r[i] = (uint32_t)c; c >>= 32; }
  }
  // r >= p  <=>  r + 19 >= 2^255
  uint32_t t[8]; uint64_t c = 19;
#pragma unroll
  for (int i = 0; i < 8; i++) { c += r[i]; t[i] = (uint32_t)c; c >>= 32; }
  uint32_t m = (uint32_t)(-(int32_t)(t[7] >> 31));
  t[7] &= 0x7fffffffu;
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = r[i] ^ (m & (r[i] ^ t[i]));
}
// canonical little-endian encoding (fully reduced mod p)
HD void fe_tobytes(uint8_t *s, const fe &f) {
  uint32_t w[8];
  fe_reduce_words(w, f);
#pragma unroll
  for (int i = 0; i < 8; i++) { s[4 * i] = (uint8_t)w[i]; s[4 * i + 1] = (uint8_t)(w[i] >> 8); s[4 * i + 2] = (uint8_t)(w[i] >> 16); s[4 * i + 3] = (uint8_t)(w[i] >> 24); }
}

HD int fe_isnegative(const fe &f) { uint32_t w[8]; fe_reduce_words(w, f); return (int)(w[0] & 1); }
HD int fe_iszero(const fe &f) {
  uint32_t w[8]; fe_reduce_words(w, f);
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) r |= w[i];
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

#define FE_CONST(name) HD void name(fe &h) { const uint32_t c[8] = name##_LIMBS; _Pragma("unroll") for (int i = 0; i < 8; i++) h.v[i] = c[i]; }
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
