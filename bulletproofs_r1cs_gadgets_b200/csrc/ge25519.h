// Edwards25519 points (a = -1) in extended coordinates and the ristretto255 encoding
// (RFC 9496).  Replaces curve25519-dalek's RistrettoPoint / CompressedRistretto (reference
// Cargo.toml:8; call sites e.g. src/gadget_poseidon.rs:584-587).  All outputs that reach proof
// bytes go through ristretto_encode, which is canonical, so the choice of addition formulas
// and MSM algorithm cannot change a single byte of a proof.
#pragma once
#include "fe25519.h"

struct alignas(32) ge_p3 { fe X, Y, Z, T; };          // 128 B: x = X/Z, y = Y/Z, xy = T/Z            (four 256-bit loads, hd.h)
struct alignas(32) ge_niels { fe ypx, ymx, xy2d; };   // affine: y+x, y-x, 2d*x*y   (96 B = three 32-byte sectors = three 256-bit loads)
struct ge_cached { fe YpX, YmX, Z, T2d; };
struct ge_p1p1 { fe X, Y, Z, T; };

HD void ge_identity(ge_p3 &p) { fe_0(p.X); fe_1(p.Y); fe_1(p.Z); fe_0(p.T); }
HD void ge_niels_identity(ge_niels &n) { fe_1(n.ypx); fe_1(n.ymx); fe_0(n.xy2d); }

// INL: expand the field multiplications in place (hot loops) instead of calling them, see fe_mul_x
template <bool INL = false> HD void ge_p1p1_to_p3(ge_p3 &r, const ge_p1p1 &p) {
  fe_mul_x<INL>(r.X, p.X, p.T); fe_mul_x<INL>(r.Y, p.Y, p.Z); fe_mul_x<INL>(r.Z, p.Z, p.T); fe_mul_x<INL>(r.T, p.X, p.Y);
}
// projective result only (T not produced): valid input for a following doubling
template <bool INL = false> HD void ge_p1p1_to_p2(ge_p3 &r, const ge_p1p1 &p) {
  fe_mul_x<INL>(r.X, p.X, p.T); fe_mul_x<INL>(r.Y, p.Y, p.Z); fe_mul_x<INL>(r.Z, p.Z, p.T);
}
template <bool INL = false> HD void ge_to_cached(ge_cached &c, const ge_p3 &p) {
  fe d2; FE_2D(d2);
  fe_add(c.YpX, p.Y, p.X); fe_sub(c.YmX, p.Y, p.X); c.Z = p.Z; fe_mul_x<INL>(c.T2d, p.T, d2);
}
// kept for the call sites written for the lazy 10-limb form: saturated limbs carry nothing over
HD void fe_carry(fe &) {}

// r = p + (neg ? -q : q), q affine niels.  7 multiplications after completion.
template <bool INL = false> HD void ge_madd_p1p1(ge_p1p1 &r, const ge_p3 &p, const ge_niels &q, int neg) {
  fe a, b, t0, qp, qm, qt;
  fe_select(qp, q.ypx, q.ymx, neg);
  fe_select(qm, q.ymx, q.ypx, neg);
  fe_cneg(qt, q.xy2d, neg);
  fe_add(a, p.Y, p.X); fe_sub(b, p.Y, p.X);
  fe_mul_x<INL>(r.Z, a, qp);      // A
  fe_mul_x<INL>(r.Y, b, qm);      // B
  fe_mul_x<INL>(r.T, qt, p.T);    // C
  fe_add(t0, p.Z, p.Z);           // D
  fe_sub(r.X, r.Z, r.Y); fe_add(r.Y, r.Z, r.Y);
  fe_add(r.Z, t0, r.T); fe_sub(r.T, t0, r.T);
}
template <bool INL = false> HD void ge_madd(ge_p3 &r, const ge_p3 &p, const ge_niels &q, int neg) {
  ge_p1p1 t; ge_madd_p1p1<INL>(t, p, q, neg); ge_p1p1_to_p3<INL>(r, t);
}
// r = p + (neg ? -q : q), q cached
template <bool INL = false> HD void ge_add_cached_p1p1(ge_p1p1 &r, const ge_p3 &p, const ge_cached &q, int neg) {
  fe a, b, t0, qp, qm, qt;
  fe_select(qp, q.YpX, q.YmX, neg);
  fe_select(qm, q.YmX, q.YpX, neg);
  fe_cneg(qt, q.T2d, neg);
  fe_add(a, p.Y, p.X); fe_sub(b, p.Y, p.X);
  fe_mul_x<INL>(r.Z, a, qp);
  fe_mul_x<INL>(r.Y, b, qm);
  fe_mul_x<INL>(r.T, qt, p.T);
  fe_mul_x<INL>(r.X, p.Z, q.Z);
  fe_add(t0, r.X, r.X);
  fe_sub(r.X, r.Z, r.Y); fe_add(r.Y, r.Z, r.Y);
  fe_add(r.Z, t0, r.T); fe_sub(r.T, t0, r.T);
}
template <bool INL = false> HD void ge_add_cached(ge_p3 &r, const ge_p3 &p, const ge_cached &q, int neg) {
  ge_p1p1 t; ge_add_cached_p1p1<INL>(t, p, q, neg); ge_p1p1_to_p3<INL>(r, t);
}
template <bool INL = false> HD void ge_add(ge_p3 &r, const ge_p3 &p, const ge_p3 &q) {
  ge_cached c; ge_to_cached<INL>(c, q); ge_add_cached<INL>(r, p, c, 0);
}
template <bool INL = false> HD void ge_sub(ge_p3 &r, const ge_p3 &p, const ge_p3 &q) {
  ge_cached c; ge_to_cached<INL>(c, q); ge_add_cached<INL>(r, p, c, 1);
}
// doubling; only X, Y, Z of p are read
template <bool INL = false> HD void ge_dbl_p1p1(ge_p1p1 &r, const ge_p3 &p) {
  fe t0;
  fe_sq_x<INL>(r.X, p.X); fe_sq_x<INL>(r.Z, p.Y); fe_sq2_x<INL>(r.T, p.Z);
  fe_add(r.Y, p.X, p.Y); fe_sq_x<INL>(t0, r.Y);
  fe_add(r.Y, r.Z, r.X); fe_sub(r.Z, r.Z, r.X);
  fe_sub(r.X, t0, r.Y); fe_sub(r.T, r.T, r.Z);
}
template <bool INL = false> HD void ge_dbl(ge_p3 &r, const ge_p3 &p) { ge_p1p1 t; ge_dbl_p1p1<INL>(t, p); ge_p1p1_to_p3<INL>(r, t); }
template <bool INL = false> HD void ge_dbl_p2(ge_p3 &r, const ge_p3 &p) { ge_p1p1 t; ge_dbl_p1p1<INL>(t, p); ge_p1p1_to_p2<INL>(r, t); }
// Point-level device functions for kernels with SEVERAL addition / doubling sites (generator fold, per-proof bucket method,
// bucket reduction): one copy of the expanded field arithmetic per kernel, operands and result in registers.
#if defined(__CUDACC__)
static __device__ __noinline__ ge_p3 ge_add_cached_fn(ge_p3 p, ge_cached q, int neg) { ge_p3 r; ge_add_cached<true>(r, p, q, neg); return r; }
static __device__ __noinline__ ge_p3 ge_dbl_fn(ge_p3 p, int need_t) {
  ge_p1p1 t; ge_dbl_p1p1<true>(t, p);
  ge_p3 r; ge_p1p1_to_p2<true>(r, t);
  if (need_t) fe_mul_x<true>(r.T, t.X, t.Y); else r.T = p.T;
  return r;
}
#endif
HD void ge_add_cached_f(ge_p3 &r, const ge_p3 &p, const ge_cached &q, int neg) {
#if defined(__CUDA_ARCH__) && BP_GE_INLINE
  r = ge_add_cached_fn(p, q, neg);
#else
  ge_add_cached(r, p, q, neg);
#endif
}
HD void ge_add_f(ge_p3 &r, const ge_p3 &p, const ge_p3 &q) { ge_cached c; ge_to_cached<true>(c, q); ge_add_cached_f(r, p, c, 0); }
HD void ge_dbl_f(ge_p3 &r, const ge_p3 &p) {
#if defined(__CUDA_ARCH__) && BP_GE_INLINE
  r = ge_dbl_fn(p, 1);
#else
  ge_dbl(r, p);
#endif
}
HD void ge_dbl_p2_f(ge_p3 &r, const ge_p3 &p) {
#if defined(__CUDA_ARCH__) && BP_GE_INLINE
  r = ge_dbl_fn(p, 0);
#else
  ge_dbl_p2(r, p);
#endif
}
HD void ge_neg(ge_p3 &r, const ge_p3 &p) { fe_neg(r.X, p.X); r.Y = p.Y; r.Z = p.Z; fe_neg(r.T, p.T); }

// affine normalisation -> niels form (one inversion)
HD void ge_to_niels(ge_niels &n, const ge_p3 &p) {
  fe zi, x, y, d2; FE_2D(d2);
  fe_invert(zi, p.Z); fe_mul(x, p.X, zi); fe_mul(y, p.Y, zi);
  fe_add(n.ypx, y, x); fe_carry(n.ypx);
  fe_sub(n.ymx, y, x); fe_carry(n.ymx);
  fe_mul(n.xy2d, x, y); fe_mul(n.xy2d, n.xy2d, d2);
}
HD void ge_normalize(ge_p3 &r, const ge_p3 &p) {
  fe zi; fe_invert(zi, p.Z);
  fe_mul(r.X, p.X, zi); fe_mul(r.Y, p.Y, zi); fe_1(r.Z); fe_mul(r.T, r.X, r.Y);
}
// ristretto equality with the identity coset: X == 0 or Y == 0
HD int ge_is_identity_ristretto(const ge_p3 &p) { return fe_iszero(p.X) | fe_iszero(p.Y); }

// RFC 9496 section 4.3.1 Decode; returns 1 on success
HD int ristretto_decode(ge_p3 &p, const uint8_t *s) {
  fe sf, ss, u1, u2, u2s, v, t, invsqrt, den_x, den_y, one, d;
  uint8_t chk[32];
  fe_frombytes(sf, s); fe_tobytes(chk, sf);
  int canonical = 1;
#pragma unroll
  for (int i = 0; i < 32; i++) canonical &= (chk[i] == s[i]);
  if (!canonical || (s[0] & 1)) return 0;
  fe_1(one); FE_D(d);
  fe_sq(ss, sf); fe_sub(u1, one, ss); fe_add(u2, one, ss); fe_sq(u2s, u2);
  fe_sq(t, u1); fe_mul(t, t, d); fe_neg(t, t); fe_sub(v, t, u2s);
  fe_mul(t, v, u2s);
  int ok = fe_sqrt_ratio_m1(invsqrt, one, t);
  fe_mul(den_x, invsqrt, u2);
  fe_mul(den_y, invsqrt, den_x); fe_mul(den_y, den_y, v);
  fe_add(t, sf, sf); fe_mul(t, t, den_x); fe_abs(p.X, t);
  fe_mul(p.Y, u1, den_y);
  fe_1(p.Z);
  fe_mul(p.T, p.X, p.Y);
  if (!ok || fe_isnegative(p.T) || fe_iszero(p.Y)) return 0;
  return 1;
}
// RFC 9496 section 4.3.2 Encode
HD void ristretto_encode(uint8_t *s, const ge_p3 &p) {
  fe u1, u2, t, invsqrt, den1, den2, z_inv, ix0, iy0, ench, x, y, den_inv, one, i, isad;
  fe_1(one); FE_SQRTM1(i); FE_INVSQRT_A_MINUS_D(isad);
  fe_add(u1, p.Z, p.Y); fe_sub(t, p.Z, p.Y); fe_mul(u1, u1, t);
  fe_mul(u2, p.X, p.Y);
  fe_sq(t, u2); fe_mul(t, t, u1);
  fe_sqrt_ratio_m1(invsqrt, one, t);
  fe_mul(den1, invsqrt, u1); fe_mul(den2, invsqrt, u2);
  fe_mul(z_inv, den1, den2); fe_mul(z_inv, z_inv, p.T);
  fe_mul(ix0, p.X, i); fe_mul(iy0, p.Y, i);
  fe_mul(ench, den1, isad);
  fe_mul(t, p.T, z_inv);
  int rotate = fe_isnegative(t);
  fe_select(x, p.X, iy0, rotate); fe_select(y, p.Y, ix0, rotate); fe_select(den_inv, den2, ench, rotate);
  fe_mul(t, x, z_inv);
  fe_cneg(y, y, fe_isnegative(t));
  fe_sub(t, p.Z, y); fe_mul(t, t, den_inv); fe_abs(t, t);
  fe_tobytes(s, t);
}
// RFC 9496 section 4.3.4 MAP (Elligator 2 variant)
HD void ristretto_elligator(ge_p3 &p, const fe &t0) {
  fe r, u, v, s, sp, c, N, w0, w1, w2, w3, one, t, rpd, i, d, omds, dmos, sadm;
  fe_1(one); FE_SQRTM1(i); FE_D(d); FE_ONE_MINUS_D_SQ(omds); FE_D_MINUS_ONE_SQ(dmos); FE_SQRT_AD_MINUS_ONE(sadm);
  fe_sq(r, t0); fe_mul(r, r, i);
  fe_add(u, r, one); fe_mul(u, u, omds);
  fe_mul(t, r, d); fe_add(t, t, one); fe_neg(t, t);
  fe_add(rpd, r, d); fe_mul(v, t, rpd);
  int sq = fe_sqrt_ratio_m1(s, u, v);
  fe_mul(sp, s, t0); fe_abs(sp, sp); fe_neg(sp, sp);
  fe mone; fe_neg(mone, one);
  fe_select(s, sp, s, sq); fe_select(c, r, mone, sq);
  fe_sub(t, r, one); fe_mul(N, c, t); fe_mul(N, N, dmos); fe_sub(N, N, v);
  fe_add(w0, s, s); fe_mul(w0, w0, v);
  fe_mul(w1, N, sadm);
  fe_sq(t, s); fe_sub(w2, one, t); fe_add(w3, one, t);
  fe_mul(p.X, w0, w3); fe_mul(p.Y, w2, w1); fe_mul(p.Z, w1, w3); fe_mul(p.T, w0, w2);
}
HD void ristretto_from_uniform(ge_p3 &p, const uint8_t *b) {
  fe r0, r1; ge_p3 p0, p1;
  fe_frombytes(r0, b); fe_frombytes(r1, b + 32);
  ristretto_elligator(p0, r0); ristretto_elligator(p1, r1);
  ge_add(p, p0, p1);
}
