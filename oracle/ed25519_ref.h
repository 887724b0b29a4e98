/* TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C) of the arithmetic the reference
 * gets from curve25519-dalek 2.x (reference Cargo.toml:8): scalar field mod l,
 * GF(2^255-19), Edwards points in extended coordinates and the ristretto255 encoding
 * (RFC 9496).  dalek is not on disk (un-vendored dependency), so this follows the
 * published algorithms and is pinned against the RFC 9496 vectors and oracle/bp_pyref.py.
 * Used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg only.
 */
#ifndef ED25519_REF_H
#define ED25519_REF_H
#include <stdint.h>
#include <string.h>
#include <stdlib.h>

typedef uint64_t u64;
typedef uint32_t u32;
typedef uint8_t u8;
typedef unsigned __int128 u128;

/* ------------------------------------------------------------------ fe: 5 x 51 bits */
typedef struct { u64 v[5]; } fe;
#define MASK51 ((1ULL << 51) - 1)

static void fe_0(fe *h) { memset(h, 0, sizeof *h); }
static void fe_1(fe *h) { fe_0(h); h->v[0] = 1; }

static void fe_carry(fe *h) {
  u64 c;
  c = h->v[0] >> 51; h->v[0] &= MASK51; h->v[1] += c;
  c = h->v[1] >> 51; h->v[1] &= MASK51; h->v[2] += c;
  c = h->v[2] >> 51; h->v[2] &= MASK51; h->v[3] += c;
  c = h->v[3] >> 51; h->v[3] &= MASK51; h->v[4] += c;
  c = h->v[4] >> 51; h->v[4] &= MASK51; h->v[0] += 19 * c;
}
static void fe_add(fe *h, const fe *f, const fe *g) {
  for (int i = 0; i < 5; i++) h->v[i] = f->v[i] + g->v[i];
  fe_carry(h);
}
static void fe_sub(fe *h, const fe *f, const fe *g) {
  /* f + 8p - g ; inputs have limbs < 2^52 */
  h->v[0] = f->v[0] + 0x3FFFFFFFFFFF68ULL - g->v[0];
  for (int i = 1; i < 5; i++) h->v[i] = f->v[i] + 0x3FFFFFFFFFFFF8ULL - g->v[i];
  fe_carry(h);
}
static void fe_neg(fe *h, const fe *f) { fe z; fe_0(&z); fe_sub(h, &z, f); }

static void fe_mul(fe *h, const fe *f, const fe *g) {
  u64 f0 = f->v[0], f1 = f->v[1], f2 = f->v[2], f3 = f->v[3], f4 = f->v[4];
  u64 g0 = g->v[0], g1 = g->v[1], g2 = g->v[2], g3 = g->v[3], g4 = g->v[4];
  u64 g1_19 = 19 * g1, g2_19 = 19 * g2, g3_19 = 19 * g3, g4_19 = 19 * g4;
  u128 r0 = (u128)f0 * g0 + (u128)f1 * g4_19 + (u128)f2 * g3_19 + (u128)f3 * g2_19 + (u128)f4 * g1_19;
  u128 r1 = (u128)f0 * g1 + (u128)f1 * g0 + (u128)f2 * g4_19 + (u128)f3 * g3_19 + (u128)f4 * g2_19;
  u128 r2 = (u128)f0 * g2 + (u128)f1 * g1 + (u128)f2 * g0 + (u128)f3 * g4_19 + (u128)f4 * g3_19;
  u128 r3 = (u128)f0 * g3 + (u128)f1 * g2 + (u128)f2 * g1 + (u128)f3 * g0 + (u128)f4 * g4_19;
  u128 r4 = (u128)f0 * g4 + (u128)f1 * g3 + (u128)f2 * g2 + (u128)f3 * g1 + (u128)f4 * g0;
  u64 c;
  r1 += (u64)(r0 >> 51); u64 h0 = (u64)r0 & MASK51;
  r2 += (u64)(r1 >> 51); u64 h1 = (u64)r1 & MASK51;
  r3 += (u64)(r2 >> 51); u64 h2 = (u64)r2 & MASK51;
  r4 += (u64)(r3 >> 51); u64 h3 = (u64)r3 & MASK51;
  c = (u64)(r4 >> 51);   u64 h4 = (u64)r4 & MASK51;
  h0 += c * 19; c = h0 >> 51; h0 &= MASK51; h1 += c;
  h->v[0] = h0; h->v[1] = h1; h->v[2] = h2; h->v[3] = h3; h->v[4] = h4;
}
static void fe_sq(fe *h, const fe *f) { fe_mul(h, f, f); }
static void fe_sqn(fe *h, const fe *f, int n) { fe_sq(h, f); for (int i = 1; i < n; i++) fe_sq(h, h); }

static void fe_frombytes(fe *h, const u8 s[32]) {
  u64 w[4];
  memcpy(w, s, 32);
  h->v[0] = w[0] & MASK51;
  h->v[1] = ((w[0] >> 51) | (w[1] << 13)) & MASK51;
  h->v[2] = ((w[1] >> 38) | (w[2] << 26)) & MASK51;
  h->v[3] = ((w[2] >> 25) | (w[3] << 39)) & MASK51;
  h->v[4] = (w[3] >> 12) & MASK51; /* drops bit 255 */
}
static void fe_tobytes(u8 s[32], const fe *f) {
  fe t = *f;
  fe_carry(&t); fe_carry(&t);
  /* now t < 2^255 + small; compute t - p if t >= p */
  u64 q = (t.v[0] + 19) >> 51;
  q = (t.v[1] + q) >> 51; q = (t.v[2] + q) >> 51; q = (t.v[3] + q) >> 51; q = (t.v[4] + q) >> 51;
  t.v[0] += 19 * q;
  u64 c;
  c = t.v[0] >> 51; t.v[0] &= MASK51; t.v[1] += c;
  c = t.v[1] >> 51; t.v[1] &= MASK51; t.v[2] += c;
  c = t.v[2] >> 51; t.v[2] &= MASK51; t.v[3] += c;
  c = t.v[3] >> 51; t.v[3] &= MASK51; t.v[4] += c;
  t.v[4] &= MASK51;
  u64 w[4];
  w[0] = t.v[0] | (t.v[1] << 51);
  w[1] = (t.v[1] >> 13) | (t.v[2] << 38);
  w[2] = (t.v[2] >> 26) | (t.v[3] << 25);
  w[3] = (t.v[3] >> 39) | (t.v[4] << 12);
  memcpy(s, w, 32);
}
static int fe_isneg(const fe *f) { u8 s[32]; fe_tobytes(s, f); return s[0] & 1; }
static int fe_iszero(const fe *f) { u8 s[32]; fe_tobytes(s, f); u8 r = 0; for (int i = 0; i < 32; i++) r |= s[i]; return r == 0; }
static int fe_eq(const fe *a, const fe *b) { u8 x[32], y[32]; fe_tobytes(x, a); fe_tobytes(y, b); return memcmp(x, y, 32) == 0; }
static void fe_abs(fe *h, const fe *f) { if (fe_isneg(f)) fe_neg(h, f); else *h = *f; }

/* z^(2^252-3) */
static void fe_pow22523(fe *out, const fe *z) {
  fe t0, t1, t2;
  fe_sq(&t0, z); fe_sqn(&t1, &t0, 2); fe_mul(&t1, z, &t1); fe_mul(&t0, &t0, &t1);
  fe_sq(&t0, &t0); fe_mul(&t0, &t1, &t0);
  fe_sqn(&t1, &t0, 5); fe_mul(&t0, &t1, &t0);
  fe_sqn(&t1, &t0, 10); fe_mul(&t1, &t1, &t0);
  fe_sqn(&t2, &t1, 20); fe_mul(&t1, &t2, &t1);
  fe_sqn(&t1, &t1, 10); fe_mul(&t0, &t1, &t0);
  fe_sqn(&t1, &t0, 50); fe_mul(&t1, &t1, &t0);
  fe_sqn(&t2, &t1, 100); fe_mul(&t1, &t2, &t1);
  fe_sqn(&t1, &t1, 50); fe_mul(&t0, &t1, &t0);
  fe_sqn(&t0, &t0, 2); fe_mul(out, &t0, z);
}
/* z^(p-2) */
static void fe_invert(fe *out, const fe *z) {
  fe t0, t1, t2, t3;
  fe_sq(&t0, z); fe_sqn(&t1, &t0, 2); fe_mul(&t1, z, &t1); fe_mul(&t0, &t0, &t1);
  fe_sq(&t2, &t0); fe_mul(&t1, &t1, &t2);
  fe_sqn(&t2, &t1, 5); fe_mul(&t1, &t2, &t1);
  fe_sqn(&t2, &t1, 10); fe_mul(&t2, &t2, &t1);
  fe_sqn(&t3, &t2, 20); fe_mul(&t2, &t3, &t2);
  fe_sqn(&t2, &t2, 10); fe_mul(&t1, &t2, &t1);
  fe_sqn(&t2, &t1, 50); fe_mul(&t2, &t2, &t1);
  fe_sqn(&t3, &t2, 100); fe_mul(&t2, &t3, &t2);
  fe_sqn(&t2, &t2, 50); fe_mul(&t1, &t2, &t1);
  fe_sqn(&t1, &t1, 5); fe_mul(out, &t1, &t0);
}

static fe FE_D, FE_2D, FE_SQRTM1, FE_ONE_MINUS_D_SQ, FE_D_MINUS_ONE_SQ, FE_SQRT_AD_MINUS_ONE, FE_INVSQRT_A_MINUS_D;

static void fe_from_u64(fe *h, u64 x) { fe_0(h); h->v[0] = x & MASK51; h->v[1] = x >> 51; }

/* RFC 9496 SQRT_RATIO_M1: returns was_square, r = sqrt(u/v) or sqrt(i*u/v), non-negative */
static int fe_sqrt_ratio_m1(fe *r, const fe *u, const fe *v) {
  fe v3, v7, t, check, neg_u, neg_u_i;
  fe_sq(&v3, v); fe_mul(&v3, &v3, v);
  fe_sq(&v7, &v3); fe_mul(&v7, &v7, v);
  fe_mul(&t, u, &v7); fe_pow22523(&t, &t);
  fe_mul(r, u, &v3); fe_mul(r, r, &t);
  fe_sq(&check, r); fe_mul(&check, &check, v);
  fe_neg(&neg_u, u); fe_mul(&neg_u_i, &neg_u, &FE_SQRTM1);
  int correct = fe_eq(&check, u), flipped = fe_eq(&check, &neg_u), flipped_i = fe_eq(&check, &neg_u_i);
  if (flipped || flipped_i) fe_mul(r, r, &FE_SQRTM1);
  fe_abs(r, r);
  return correct || flipped;
}

/* ------------------------------------------------------------------ sc: scalars mod l, 4 x 64, canonical */
typedef struct { u64 v[4]; } sc;
static const sc SC_L = {{0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL, 0, 0x1000000000000000ULL}};
static u64 SC_NINV; /* -l^{-1} mod 2^64 */
static sc SC_R, SC_R2, SC_R3; /* 2^256, 2^512, 2^768 mod l */

static int sc_geq(const sc *a, const sc *b) {
  for (int i = 3; i >= 0; i--) { if (a->v[i] > b->v[i]) return 1; if (a->v[i] < b->v[i]) return 0; }
  return 1;
}
static u64 sc_sub_raw(sc *r, const sc *a, const sc *b) {
  u64 borrow = 0;
  for (int i = 0; i < 4; i++) { u128 t = (u128)a->v[i] - b->v[i] - borrow; r->v[i] = (u64)t; borrow = (u64)(t >> 64) & 1; }
  return borrow;
}
static u64 sc_add_raw(sc *r, const sc *a, const sc *b) {
  u64 carry = 0;
  for (int i = 0; i < 4; i++) { u128 t = (u128)a->v[i] + b->v[i] + carry; r->v[i] = (u64)t; carry = (u64)(t >> 64); }
  return carry;
}
static void sc_add(sc *r, const sc *a, const sc *b) { sc_add_raw(r, a, b); if (sc_geq(r, &SC_L)) sc_sub_raw(r, r, &SC_L); }
static void sc_sub(sc *r, const sc *a, const sc *b) { if (sc_sub_raw(r, a, b)) sc_add_raw(r, r, &SC_L); }
static void sc_neg(sc *r, const sc *a) { sc z = {{0, 0, 0, 0}}; sc_sub(r, &z, a); }
static int sc_iszero(const sc *a) { return (a->v[0] | a->v[1] | a->v[2] | a->v[3]) == 0; }
static int sc_eq(const sc *a, const sc *b) { return memcmp(a, b, sizeof(sc)) == 0; }
static void sc_from_u64(sc *r, u64 x) { r->v[0] = x; r->v[1] = r->v[2] = r->v[3] = 0; }

/* Montgomery product a*b/2^256 mod l; requires b < l (a may be any 256-bit value) */
static void sc_montmul(sc *r, const sc *a, const sc *b) {
  u64 t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u64 c = 0;
    for (int j = 0; j < 4; j++) { u128 p = (u128)a->v[j] * b->v[i] + t[j] + c; t[j] = (u64)p; c = (u64)(p >> 64); }
    u128 s = (u128)t[4] + c; t[4] = (u64)s; t[5] = (u64)(s >> 64);
    u64 m = t[0] * SC_NINV;
    u128 p = (u128)m * SC_L.v[0] + t[0]; c = (u64)(p >> 64);
    for (int j = 1; j < 4; j++) { p = (u128)m * SC_L.v[j] + t[j] + c; t[j - 1] = (u64)p; c = (u64)(p >> 64); }
    s = (u128)t[4] + c; t[3] = (u64)s; t[4] = t[5] + (u64)(s >> 64);
  }
  sc x = {{t[0], t[1], t[2], t[3]}};
  if (t[4] || sc_geq(&x, &SC_L)) sc_sub_raw(&x, &x, &SC_L);
  if (sc_geq(&x, &SC_L)) sc_sub_raw(&x, &x, &SC_L);
  *r = x;
}
static void sc_mul(sc *r, const sc *a, const sc *b) { sc t; sc_montmul(&t, a, b); sc_montmul(r, &t, &SC_R2); }
static void sc_muladd(sc *r, const sc *a, const sc *b, const sc *c) { sc t; sc_mul(&t, a, b); sc_add(r, &t, c); }
static void sc_from_bytes_mod_order(sc *r, const u8 s[32]) { sc x; memcpy(&x, s, 32); sc_montmul(r, &x, &SC_R); }
static void sc_from_bytes_wide(sc *r, const u8 s[64]) {
  sc lo, hi, a, b; memcpy(&lo, s, 32); memcpy(&hi, s + 32, 32);
  sc_montmul(&a, &lo, &SC_R); sc_montmul(&b, &hi, &SC_R2); sc_add(r, &a, &b);
}
static void sc_tobytes(u8 s[32], const sc *a) { memcpy(s, a, 32); }
static int sc_from_canonical(sc *r, const u8 s[32]) { memcpy(r, s, 32); return !sc_geq(r, &SC_L); }
/* a^(l-2); invert(0) == 0 as in dalek */
static void sc_invert(sc *r, const sc *a) {
  sc e = SC_L; e.v[0] -= 2;
  sc am, acc; sc_montmul(&am, a, &SC_R2); acc = SC_R;
  for (int i = 252; i >= 0; i--) {
    sc_montmul(&acc, &acc, &acc);
    if ((e.v[i >> 6] >> (i & 63)) & 1) sc_montmul(&acc, &acc, &am);
  }
  sc one = {{1, 0, 0, 0}}; sc_montmul(r, &acc, &one);
}

/* ------------------------------------------------------------------ ge: extended coordinates, a = -1 */
typedef struct { fe X, Y, Z, T; } ge;
static void ge_identity(ge *p) { fe_0(&p->X); fe_1(&p->Y); fe_1(&p->Z); fe_0(&p->T); }
static void ge_add(ge *r, const ge *p, const ge *q) {
  fe A, B, C, D, E, F, G, H, t0, t1;
  fe_sub(&t0, &p->Y, &p->X); fe_sub(&t1, &q->Y, &q->X); fe_mul(&A, &t0, &t1);
  fe_add(&t0, &p->Y, &p->X); fe_add(&t1, &q->Y, &q->X); fe_mul(&B, &t0, &t1);
  fe_mul(&C, &p->T, &q->T); fe_mul(&C, &C, &FE_2D);
  fe_mul(&D, &p->Z, &q->Z); fe_add(&D, &D, &D);
  fe_sub(&E, &B, &A); fe_sub(&F, &D, &C); fe_add(&G, &D, &C); fe_add(&H, &B, &A);
  fe_mul(&r->X, &E, &F); fe_mul(&r->Y, &G, &H); fe_mul(&r->Z, &F, &G); fe_mul(&r->T, &E, &H);
}
static void ge_dbl(ge *r, const ge *p) {
  fe A, B, C, E, F, G, H, t0;
  fe_sq(&A, &p->X); fe_sq(&B, &p->Y); fe_sq(&C, &p->Z); fe_add(&C, &C, &C);
  fe_add(&t0, &p->X, &p->Y); fe_sq(&t0, &t0);
  fe_add(&H, &A, &B);             /* H' = A + B */
  fe_sub(&E, &H, &t0);            /* E' = A + B - (X+Y)^2 = -E */
  fe_sub(&G, &A, &B);             /* G' = A - B = -G */
  fe_add(&F, &C, &G);             /* F' = C + A - B = -(G - C) = -F */
  /* X3 = E*F = E'*F', Y3 = G*H where H = D - B = -(A+B) => G*H = G'*H', T3 = E*H = E'*H', Z3 = F*G = F'*G' */
  fe_mul(&r->X, &E, &F); fe_mul(&r->Y, &G, &H); fe_mul(&r->Z, &F, &G); fe_mul(&r->T, &E, &H);
}
static void ge_neg(ge *r, const ge *p) { fe_neg(&r->X, &p->X); r->Y = p->Y; r->Z = p->Z; fe_neg(&r->T, &p->T); }
static int ge_is_identity_ristretto(const ge *p) {
  /* ristretto equality with (0,1): X*1 == Y*0 or Y*1 == X*0  ->  X == 0 or Y == 0 */
  return fe_iszero(&p->X) || fe_iszero(&p->Y);
}

static int ristretto_decode(ge *p, const u8 s[32]) {
  fe sf, ss, u1, u2, u2s, v, t, invsqrt, den_x, den_y, one;
  u8 chk[32];
  fe_frombytes(&sf, s); fe_tobytes(chk, &sf);
  if (memcmp(chk, s, 32) != 0 || (s[0] & 1)) return 0;
  fe_1(&one);
  fe_sq(&ss, &sf); fe_sub(&u1, &one, &ss); fe_add(&u2, &one, &ss); fe_sq(&u2s, &u2);
  fe_sq(&t, &u1); fe_mul(&t, &t, &FE_D); fe_neg(&t, &t); fe_sub(&v, &t, &u2s);
  fe_mul(&t, &v, &u2s);
  int ok = fe_sqrt_ratio_m1(&invsqrt, &one, &t);
  fe_mul(&den_x, &invsqrt, &u2);
  fe_mul(&den_y, &invsqrt, &den_x); fe_mul(&den_y, &den_y, &v);
  fe_add(&t, &sf, &sf); fe_mul(&t, &t, &den_x); fe_abs(&p->X, &t);
  fe_mul(&p->Y, &u1, &den_y);
  fe_1(&p->Z);
  fe_mul(&p->T, &p->X, &p->Y);
  if (!ok || fe_isneg(&p->T) || fe_iszero(&p->Y)) return 0;
  return 1;
}
static void ristretto_encode(u8 s[32], const ge *p) {
  fe u1, u2, t, invsqrt, den1, den2, z_inv, ix0, iy0, ench, x, y, den_inv, one;
  fe_1(&one);
  fe_add(&u1, &p->Z, &p->Y); fe_sub(&t, &p->Z, &p->Y); fe_mul(&u1, &u1, &t);
  fe_mul(&u2, &p->X, &p->Y);
  fe_sq(&t, &u2); fe_mul(&t, &t, &u1);
  fe_sqrt_ratio_m1(&invsqrt, &one, &t);
  fe_mul(&den1, &invsqrt, &u1); fe_mul(&den2, &invsqrt, &u2);
  fe_mul(&z_inv, &den1, &den2); fe_mul(&z_inv, &z_inv, &p->T);
  fe_mul(&ix0, &p->X, &FE_SQRTM1); fe_mul(&iy0, &p->Y, &FE_SQRTM1);
  fe_mul(&ench, &den1, &FE_INVSQRT_A_MINUS_D);
  fe_mul(&t, &p->T, &z_inv);
  if (fe_isneg(&t)) { x = iy0; y = ix0; den_inv = ench; } else { x = p->X; y = p->Y; den_inv = den2; }
  fe_mul(&t, &x, &z_inv);
  if (fe_isneg(&t)) fe_neg(&y, &y);
  fe_sub(&t, &p->Z, &y); fe_mul(&t, &t, &den_inv); fe_abs(&t, &t);
  fe_tobytes(s, &t);
}
static void ristretto_elligator(ge *p, const fe *t0) {
  fe r, u, v, s, sp, c, N, w0, w1, w2, w3, one, t, rpd;
  fe_1(&one);
  fe_sq(&r, t0); fe_mul(&r, &r, &FE_SQRTM1);
  fe_add(&u, &r, &one); fe_mul(&u, &u, &FE_ONE_MINUS_D_SQ);
  fe_mul(&t, &r, &FE_D); fe_add(&t, &t, &one); fe_neg(&t, &t); /* -1 - r*d */
  fe_add(&rpd, &r, &FE_D); fe_mul(&v, &t, &rpd);
  int sq = fe_sqrt_ratio_m1(&s, &u, &v);
  fe_mul(&sp, &s, t0); fe_abs(&sp, &sp); fe_neg(&sp, &sp);
  if (!sq) { s = sp; c = r; } else { fe_neg(&c, &one); }
  fe_sub(&t, &r, &one); fe_mul(&N, &c, &t); fe_mul(&N, &N, &FE_D_MINUS_ONE_SQ); fe_sub(&N, &N, &v);
  fe_add(&w0, &s, &s); fe_mul(&w0, &w0, &v);
  fe_mul(&w1, &N, &FE_SQRT_AD_MINUS_ONE);
  fe_sq(&t, &s); fe_sub(&w2, &one, &t); fe_add(&w3, &one, &t);
  fe_mul(&p->X, &w0, &w3); fe_mul(&p->Y, &w2, &w1); fe_mul(&p->Z, &w1, &w3); fe_mul(&p->T, &w0, &w2);
}
static void ristretto_from_uniform(ge *p, const u8 b[64]) {
  fe r0, r1; ge p0, p1;
  fe_frombytes(&r0, b); fe_frombytes(&r1, b + 32);
  ristretto_elligator(&p0, &r0); ristretto_elligator(&p1, &r1);
  ge_add(p, &p0, &p1);
}

/* --- multiscalar multiplication (variable time): Straus w-NAF(5) for small n, Pippenger above,
 * the same split dalek's vartime_multiscalar_mul makes (n < 190 -> Straus).  Outputs are
 * canonical encodings, so algorithm choice cannot change proof bytes. */
static void sc_naf5(int8_t naf[256], const sc *s) {
  u64 x[5] = {s->v[0], s->v[1], s->v[2], s->v[3], 0};
  memset(naf, 0, 256);
  int pos = 0; u64 carry = 0;
  while (pos < 256) {
    int idx = pos >> 6, bit = pos & 63;
    u64 buf = bit < 59 ? (x[idx] >> bit) : ((x[idx] >> bit) | (x[idx + 1] << (64 - bit)));
    u64 window = carry + (buf & 31);
    if ((window & 1) == 0) { pos += 1; continue; }
    if (window < 16) { carry = 0; naf[pos] = (int8_t)window; } else { carry = 1; naf[pos] = (int8_t)((int)window - 32); }
    pos += 5;
  }
}
static void ge_msm_straus(ge *out, int n, const sc *scalars, const ge *points) {
  int8_t (*nafs)[256] = malloc((size_t)n * 256);
  ge (*tab)[8] = malloc((size_t)n * sizeof(ge[8]));
  for (int i = 0; i < n; i++) {
    sc_naf5(nafs[i], &scalars[i]);
    ge p2; ge_dbl(&p2, &points[i]);
    tab[i][0] = points[i];
    for (int j = 1; j < 8; j++) ge_add(&tab[i][j], &tab[i][j - 1], &p2);
  }
  ge r; ge_identity(&r);
  int top = 255;
  for (; top >= 0; top--) { int any = 0; for (int i = 0; i < n; i++) if (nafs[i][top]) { any = 1; break; } if (any) break; }
  for (int b = top; b >= 0; b--) {
    ge_dbl(&r, &r);
    for (int i = 0; i < n; i++) {
      int d = nafs[i][b];
      if (d > 0) ge_add(&r, &r, &tab[i][d >> 1]);
      else if (d < 0) { ge t; ge_neg(&t, &tab[i][(-d) >> 1]); ge_add(&r, &r, &t); }
    }
  }
  *out = r; free(nafs); free(tab);
}
static void ge_msm_pippenger(ge *out, int n, const sc *scalars, const ge *points) {
  int c = n < 500 ? 6 : n < 800 ? 7 : n < 3000 ? 8 : n < 12000 ? 10 : n < 50000 ? 12 : 13;
  int nb = 1 << c, nwin = (253 + c - 1) / c;
  ge *buckets = malloc(sizeof(ge) * nb);
  u8 *used = malloc(nb);
  ge res; ge_identity(&res);
  for (int w = nwin - 1; w >= 0; w--) {
    for (int i = 0; i < c; i++) ge_dbl(&res, &res);
    memset(used, 0, nb);
    int sh = w * c;
    for (int i = 0; i < n; i++) {
      int idx = sh >> 6, bit = sh & 63;
      u64 d = scalars[i].v[idx] >> bit;
      if (bit + c > 64 && idx < 3) d |= scalars[i].v[idx + 1] << (64 - bit);
      d &= (u64)(nb - 1);
      if (!d) continue;
      if (used[d]) ge_add(&buckets[d], &buckets[d], &points[i]); else { buckets[d] = points[i]; used[d] = 1; }
    }
    ge run, tot; int hr = 0, ht = 0;
    for (int d = nb - 1; d > 0; d--) {
      if (used[d]) { if (hr) ge_add(&run, &run, &buckets[d]); else { run = buckets[d]; hr = 1; } }
      if (hr) { if (ht) ge_add(&tot, &tot, &run); else { tot = run; ht = 1; } }
    }
    if (ht) ge_add(&res, &res, &tot);
  }
  *out = res; free(buckets); free(used);
}
static void ge_msm(ge *out, int n, const sc *scalars, const ge *points) {
  if (n == 0) { ge_identity(out); return; }
  if (n < 190) ge_msm_straus(out, n, scalars, points); else ge_msm_pippenger(out, n, scalars, points);
}
static void ge_scalarmult(ge *out, const sc *s, const ge *p) { ge_msm_straus(out, 1, s, p); }

static ge GE_BASEPOINT;
static const u8 RISTRETTO_BASEPOINT_COMPRESSED[32] = {
    0xe2, 0xf2, 0xae, 0x0a, 0x6a, 0xbc, 0x4e, 0x71, 0xa8, 0x84, 0xa9, 0x61, 0xc5, 0x00, 0x51, 0x5f,
    0x58, 0xe3, 0x0b, 0x6a, 0xa5, 0x82, 0xdd, 0x8d, 0xb6, 0xa6, 0x59, 0x45, 0xe0, 0x8d, 0x2d, 0x76};

static int ed_ref_inited = 0;
static void ed_ref_init(void) {
  if (ed_ref_inited) return;
  /* d = -121665/121666 */
  fe a, b, one; fe_1(&one);
  fe_from_u64(&a, 121665); fe_from_u64(&b, 121666);
  fe_invert(&b, &b); fe_mul(&FE_D, &a, &b); fe_neg(&FE_D, &FE_D);
  fe_add(&FE_2D, &FE_D, &FE_D);
  /* sqrt(-1) = 2^((p-1)/4): (p-1)/4 = 2^253 - 5 ; use 2^((p-1)/4) = 2 * 2^(2^253-6) ... simpler: sqrt_ratio of -1 */
  {
    /* compute via pow22523: for u = -1, v = 1 the candidate r = u * (u)^((p-5)/8); iterate instead with exponent */
    fe two; fe_from_u64(&two, 2);
    /* 2^((p-1)/4) = (2^((p-5)/8))^2 * 2 */
    fe t; fe_pow22523(&t, &two); fe_sq(&t, &t); fe_mul(&FE_SQRTM1, &t, &two);
    /* fix sign so that it matches RFC 9496 constant (which is the non-negative... check): */
    fe chk; fe_sq(&chk, &FE_SQRTM1); fe m1; fe_neg(&m1, &one);
    if (!fe_eq(&chk, &m1)) abort();
    /* RFC 9496 SQRT_M1 = 19681161376707505956807079304988542015446066515923890162744021073123829784752 (even) */
    if (fe_isneg(&FE_SQRTM1)) fe_neg(&FE_SQRTM1, &FE_SQRTM1);
  }
  fe t;
  fe_sq(&t, &FE_D); fe_sub(&FE_ONE_MINUS_D_SQ, &one, &t);
  fe_sub(&t, &FE_D, &one); fe_sq(&FE_D_MINUS_ONE_SQ, &t);
  /* sqrt(a*d - 1) with a = -1: RFC constant is odd ("negative") */
  fe adm1; fe_neg(&adm1, &FE_D); fe_sub(&adm1, &adm1, &one);
  fe_sqrt_ratio_m1(&FE_SQRT_AD_MINUS_ONE, &adm1, &one);
  {
    /* RFC 9496: SQRT_AD_MINUS_ONE = 2506306895338462347411141415870215270124453150249265646007921048261043075 0235 (odd) */
    if (!fe_isneg(&FE_SQRT_AD_MINUS_ONE)) fe_neg(&FE_SQRT_AD_MINUS_ONE, &FE_SQRT_AD_MINUS_ONE);
  }
  fe amd; fe_neg(&amd, &one); fe_sub(&amd, &amd, &FE_D);
  fe_sqrt_ratio_m1(&FE_INVSQRT_A_MINUS_D, &one, &amd);
  /* scalar constants */
  u64 inv = 1; for (int i = 0; i < 6; i++) inv *= 2 - SC_L.v[0] * inv; SC_NINV = (u64)0 - inv;
  sc r = {{1, 0, 0, 0}};
  for (int i = 0; i < 256; i++) sc_add(&r, &r, &r);
  SC_R = r;
  for (int i = 0; i < 256; i++) sc_add(&r, &r, &r);
  SC_R2 = r;
  for (int i = 0; i < 256; i++) sc_add(&r, &r, &r);
  SC_R3 = r;
  if (!ristretto_decode(&GE_BASEPOINT, RISTRETTO_BASEPOINT_COMPRESSED)) abort();
  ed_ref_inited = 1;
}
#endif
