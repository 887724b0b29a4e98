/* TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C) of the reference's proving path.
 *
 * The reference (lovesh/bulletproofs-r1cs-gadgets) gets Prover::prove / Verifier::verify,
 * the inner-product argument and the generator chains from the un-vendored `bulletproofs`
 * fork (reference Cargo.toml:22-26; call sites e.g. src/gadget_vsmt_2.rs:289-347,355-395),
 * so this file follows the published protocol and is pinned by
 *   - RFC 9496 ristretto255 vectors, the Merlin "test protocol" vector,
 *   - byte-equality with the independent big-int restatement oracle/bp_pyref.py,
 *   - the reference's constants file and gadget shapes (tests/test_oracle.py).
 * PARITY UNPINNED against dalek's bytes: the reference holds no golden proofs.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  It is the checker, never the product path.
 *
 * Native witness helpers restate src/gadget_poseidon.rs:189-280,428-443 (permutation,
 * hash_2), src/gadget_vsmt_2.rs:171-209 (multiplier order of the membership circuit) and
 * src/gadget_mimc.rs:19-39,55-79.
 */
#include "merlin_ref.h"
#include <pthread.h>
#include <stdio.h>

enum { BPO_OK = 0, BPO_ERR_GENS = 1, BPO_ERR_FORMAT = 2, BPO_ERR_VERIFY = 3, BPO_ERR_MISSING = 4, BPO_ERR_ARG = 5 };
enum { K_COMMITTED = 0, K_LEFT = 1, K_RIGHT = 2, K_OUT = 3, K_ONE = 4 };

/* ------------------------------------------------------------------ generators */
static ge PC_B, PC_BB;
static u8 PC_B_C[32], PC_BB_C[32];
static ge *GENS_G = NULL, *GENS_H = NULL;
static u32 GENS_N = 0;
static pthread_mutex_t GENS_MU = PTHREAD_MUTEX_INITIALIZER;
static int bpo_inited = 0;

int bpo_init(void) {
  pthread_mutex_lock(&GENS_MU);
  if (!bpo_inited) {
    ed_ref_init();
    PC_B = GE_BASEPOINT;
    memcpy(PC_B_C, RISTRETTO_BASEPOINT_COMPRESSED, 32);
    u8 h[64];
    sha3_512(h, PC_B_C, 32);
    ristretto_from_uniform(&PC_BB, h);
    ristretto_encode(PC_BB_C, &PC_BB);
    bpo_inited = 1;
  }
  pthread_mutex_unlock(&GENS_MU);
  return 0;
}

static void gens_chain(ge *out, char which, u32 n) {
  u8 seed[15 + 5];
  memcpy(seed, "GeneratorsChain", 15);
  seed[15] = (u8)which; seed[16] = seed[17] = seed[18] = seed[19] = 0; /* party 0, u32 LE */
  keccak_xof k; shake256_init(&k, seed, sizeof seed);
  for (u32 i = 0; i < n; i++) { u8 b[64]; keccak_squeeze(&k, b, 64); ristretto_from_uniform(&out[i], b); }
}

int bpo_ensure_gens(u32 N) {
  bpo_init();
  pthread_mutex_lock(&GENS_MU);
  if (GENS_N < N) {
    free(GENS_G); free(GENS_H);
    GENS_G = malloc(sizeof(ge) * N); GENS_H = malloc(sizeof(ge) * N);
    gens_chain(GENS_G, 'G', N); gens_chain(GENS_H, 'H', N);
    GENS_N = N;
  }
  pthread_mutex_unlock(&GENS_MU);
  return 0;
}

void bpo_pedersen(u8 B[32], u8 Bb[32]) { bpo_init(); memcpy(B, PC_B_C, 32); memcpy(Bb, PC_BB_C, 32); }

int bpo_gens_compressed(int which, u32 n, u8 *out) {
  bpo_ensure_gens(n);
  const ge *src = which ? GENS_H : GENS_G;
  for (u32 i = 0; i < n; i++) ristretto_encode(out + 32 * i, &src[i]);
  return 0;
}

static void pc_commit(ge *out, const sc *v, const sc *r) {
  sc s[2] = {*v, *r}; ge p[2] = {PC_B, PC_BB};
  ge_msm(out, 2, s, p);
}
void bpo_commit(const u8 v[32], const u8 r[32], u8 out[32]) {
  bpo_init();
  sc a, b; sc_from_bytes_mod_order(&a, v); sc_from_bytes_mod_order(&b, r);
  ge p; pc_commit(&p, &a, &b); ristretto_encode(out, &p);
}

/* KAT helpers */
int bpo_ristretto_roundtrip(const u8 in[32], u8 out[32]) { bpo_init(); ge p; if (!ristretto_decode(&p, in)) return 1; ristretto_encode(out, &p); return 0; }
void bpo_from_uniform(const u8 in[64], u8 out[32]) { bpo_init(); ge p; ristretto_from_uniform(&p, in); ristretto_encode(out, &p); }
int bpo_scalarmult(const u8 s[32], const u8 p[32], u8 out[32]) {
  bpo_init(); ge P, R; sc k; if (!ristretto_decode(&P, p)) return 1;
  sc_from_bytes_mod_order(&k, s); ge_scalarmult(&R, &k, &P); ristretto_encode(out, &R); return 0;
}
int bpo_msm(u32 n, const u8 *scalars, const u8 *points, u8 out[32]) {
  bpo_init();
  sc *s = malloc(sizeof(sc) * (n ? n : 1)); ge *p = malloc(sizeof(ge) * (n ? n : 1));
  int rc = 0;
  for (u32 i = 0; i < n; i++) { sc_from_bytes_mod_order(&s[i], scalars + 32 * i); if (!ristretto_decode(&p[i], points + 32 * i)) rc = 1; }
  if (!rc) { ge r; ge_msm(&r, (int)n, s, p); ristretto_encode(out, &r); }
  free(s); free(p); return rc;
}
void bpo_merlin_kat(const u8 *label, u32 ll, const u8 *ml, u32 mll, const u8 *msg, u32 msgl, const u8 *cl, u32 cll, u8 *out, u32 outl) {
  /* Transcript::new(label); append_message(ml,msg); challenge_bytes(cl,out) with arbitrary byte labels */
  transcript t; ts_init(&t, label, ll);
  u8 l4[4]; u32le(l4, msgl);
  strobe_meta_ad(&t.s, ml, mll, 0); strobe_meta_ad(&t.s, l4, 4, 1); strobe_ad(&t.s, msg, msgl, 0);
  u32le(l4, outl);
  strobe_meta_ad(&t.s, cl, cll, 0); strobe_meta_ad(&t.s, l4, 4, 1); strobe_prf(&t.s, out, outl);
}
void bpo_sc_invert(const u8 in[32], u8 out[32]) { bpo_init(); sc a, r; sc_from_bytes_mod_order(&a, in); sc_invert(&r, &a); sc_tobytes(out, &r); }
void bpo_sc_mul(const u8 a[32], const u8 b[32], u8 out[32]) { bpo_init(); sc x, y, r; sc_from_bytes_mod_order(&x, a); sc_from_bytes_mod_order(&y, b); sc_mul(&r, &x, &y); sc_tobytes(out, &r); }
void bpo_sc_from_wide(const u8 in[64], u8 out[32]) { bpo_init(); sc r; sc_from_bytes_wide(&r, in); sc_tobytes(out, &r); }

/* ------------------------------------------------------------------ circuit view */
typedef struct {
  u32 n, m, q;
  const u32 *cons_ptr; /* q+1 */
  const u8 *kind;      /* nnz */
  const u32 *idx;      /* nnz */
  const u8 *coeff;     /* nnz x 32, canonical LE */
} circuit;

static void flatten(const circuit *c, const sc *z, sc *wL, sc *wR, sc *wO, sc *wV, sc *wc) {
  sc zero = {{0, 0, 0, 0}};
  for (u32 i = 0; i < c->n; i++) wL[i] = wR[i] = wO[i] = zero;
  for (u32 i = 0; i < c->m; i++) wV[i] = zero;
  *wc = zero;
  sc ez = *z;
  for (u32 k = 0; k < c->q; k++) {
    for (u32 t = c->cons_ptr[k]; t < c->cons_ptr[k + 1]; t++) {
      sc co, p; memcpy(&co, c->coeff + 32 * (size_t)t, 32);
      sc_mul(&p, &ez, &co);
      u32 i = c->idx[t];
      switch (c->kind[t]) {
        case K_LEFT: sc_add(&wL[i], &wL[i], &p); break;
        case K_RIGHT: sc_add(&wR[i], &wR[i], &p); break;
        case K_OUT: sc_add(&wO[i], &wO[i], &p); break;
        case K_COMMITTED: sc_sub(&wV[i], &wV[i], &p); break;
        default: sc_sub(wc, wc, &p); break;
      }
    }
    sc_mul(&ez, &ez, z);
  }
}

static void sc_ip(sc *r, const sc *a, const sc *b, u32 n) {
  sc acc = {{0, 0, 0, 0}}, t;
  for (u32 i = 0; i < n; i++) { sc_mul(&t, &a[i], &b[i]); sc_add(&acc, &acc, &t); }
  *r = acc;
}

static u32 next_pow2(u32 n) { u32 N = 1; while (N < n) N <<= 1; return N; }
static u32 ilog2(u32 N) { u32 k = 0; while ((1u << k) < N) k++; return k; }

/* InnerProductProof::create.  G,H are consumed (folded in place). */
static void ipa_create(transcript *t, const ge *Q, const sc *Gf, const sc *Hf, ge *G, ge *H, sc *a, sc *b, u32 N, u8 *out /* 2k points then a,b */) {
  ts_append(t, "dom-sep", "ipp v1", 6);
  ts_append_u64(t, "n", N);
  u32 n = N; int first = 1; u32 round = 0;
  sc *s = malloc(sizeof(sc) * (N + 1)); ge *p = malloc(sizeof(ge) * (N + 1));
  while (n != 1) {
    n >>= 1;
    sc *aL = a, *aR = a + n, *bL = b, *bR = b + n;
    ge *GL = G, *GR = G + n, *HL = H, *HR = H + n;
    sc cL, cR; sc_ip(&cL, aL, bR, n); sc_ip(&cR, aR, bL, n);
    ge Lp, Rp;
    for (u32 i = 0; i < n; i++) {
      if (first) { sc_mul(&s[i], &aL[i], &Gf[n + i]); sc_mul(&s[n + i], &bR[i], &Hf[i]); }
      else { s[i] = aL[i]; s[n + i] = bR[i]; }
      p[i] = GR[i]; p[n + i] = HL[i];
    }
    s[2 * n] = cL; p[2 * n] = *Q;
    ge_msm(&Lp, (int)(2 * n + 1), s, p);
    for (u32 i = 0; i < n; i++) {
      if (first) { sc_mul(&s[i], &aR[i], &Gf[i]); sc_mul(&s[n + i], &bL[i], &Hf[n + i]); }
      else { s[i] = aR[i]; s[n + i] = bL[i]; }
      p[i] = GL[i]; p[n + i] = HR[i];
    }
    s[2 * n] = cR; p[2 * n] = *Q;
    ge_msm(&Rp, (int)(2 * n + 1), s, p);
    u8 *Lc = out + 64 * round, *Rc = Lc + 32;
    ristretto_encode(Lc, &Lp); ristretto_encode(Rc, &Rp);
    ts_append_point(t, "L", Lc); ts_append_point(t, "R", Rc);
    sc u, ui; ts_challenge_scalar(t, "u", &u); sc_invert(&ui, &u);
    for (u32 i = 0; i < n; i++) {
      sc x, y;
      sc_mul(&x, &aL[i], &u); sc_mul(&y, &ui, &aR[i]); sc_add(&a[i], &x, &y);
      sc_mul(&x, &bL[i], &ui); sc_mul(&y, &u, &bR[i]); sc_add(&b[i], &x, &y);
      sc gs[2], hs[2]; ge gp[2] = {GL[i], GR[i]}, hp[2] = {HL[i], HR[i]};
      if (first) { sc_mul(&gs[0], &ui, &Gf[i]); sc_mul(&gs[1], &u, &Gf[n + i]); sc_mul(&hs[0], &u, &Hf[i]); sc_mul(&hs[1], &ui, &Hf[n + i]); }
      else { gs[0] = ui; gs[1] = u; hs[0] = u; hs[1] = ui; }
      ge_msm_straus(&G[i], 2, gs, gp); ge_msm_straus(&H[i], 2, hs, hp);
    }
    first = 0; round++;
  }
  sc_tobytes(out + 64 * round, &a[0]); sc_tobytes(out + 64 * round + 32, &b[0]);
  free(s); free(p);
}

size_t bpo_proof_len(u32 n) { u32 N = next_pow2(n ? n : 1); return 32 * (14 + 2 * ilog2(N) + 2); }

/* Prover::prove, single phase.  Witness scalars canonical LE.
 * Output: V[m][32] and the untagged field tuple (A_I1,A_O1,S1,A_I2,A_O2,S2,T_1,T_3,T_4,T_5,T_6,
 * t_x,t_x_blinding,e_blinding,L_0,R_0,...,a,b). */
static int prove_one(const circuit *c, const u8 *aLb, const u8 *aRb, const u8 *aOb, const u8 *vb_, const u8 *vblb,
                     const u8 *label, u32 label_len, const u8 entropy[32], u32 gens_capacity, u8 *V_out, u8 *proof) {
  u32 n = c->n, m = c->m;
  u32 N = next_pow2(n ? n : 1), k = ilog2(N);
  if (gens_capacity < n || gens_capacity < N) return BPO_ERR_GENS;
  bpo_ensure_gens(N);
  transcript t; ts_init(&t, label, label_len);
  ts_append(&t, "dom-sep", "r1cs v1", 7);
  sc *v = malloc(sizeof(sc) * (m + 1)), *vbl = malloc(sizeof(sc) * (m + 1));
  for (u32 i = 0; i < m; i++) {
    sc_from_bytes_mod_order(&v[i], vb_ + 32 * i); sc_from_bytes_mod_order(&vbl[i], vblb + 32 * i);
    ge P; pc_commit(&P, &v[i], &vbl[i]); ristretto_encode(V_out + 32 * i, &P);
    ts_append_point(&t, "V", V_out + 32 * i);
  }
  ts_append_u64(&t, "m", m);
  transcript_rng rng; trng_begin(&rng, &t);
  for (u32 i = 0; i < m; i++) { u8 b[32]; sc_tobytes(b, &vbl[i]); trng_rekey(&rng, "v_blinding", b, 32); }
  trng_finalize(&rng, entropy);
  sc *aL = malloc(sizeof(sc) * N * 12);
  sc *aR = aL + N, *aO = aR + N, *sL = aO + N, *sR = sL + N, *wL = sR + N, *wR = wL + N, *wO = wR + N,
     *lv = wO + N, *rv = lv + N, *eyi = rv + N, *tmp = eyi + N;
  for (u32 i = 0; i < n; i++) {
    sc_from_bytes_mod_order(&aL[i], aLb + 32 * i); sc_from_bytes_mod_order(&aR[i], aRb + 32 * i); sc_from_bytes_mod_order(&aO[i], aOb + 32 * i);
  }
  sc i_b, o_b, s_b;
  trng_scalar(&rng, &i_b); trng_scalar(&rng, &o_b); trng_scalar(&rng, &s_b);
  for (u32 i = 0; i < n; i++) trng_scalar(&rng, &sL[i]);
  for (u32 i = 0; i < n; i++) trng_scalar(&rng, &sR[i]);
  /* A_I1, A_O1, S1 */
  sc *ms = malloc(sizeof(sc) * (2 * N + 2)); ge *mp = malloc(sizeof(ge) * (2 * N + 2));
  u8 *pf = proof;
  {
    ge P;
    ms[0] = i_b; mp[0] = PC_BB;
    for (u32 i = 0; i < n; i++) { ms[1 + i] = aL[i]; mp[1 + i] = GENS_G[i]; ms[1 + n + i] = aR[i]; mp[1 + n + i] = GENS_H[i]; }
    ge_msm(&P, (int)(2 * n + 1), ms, mp); ristretto_encode(pf, &P);
    ms[0] = o_b;
    for (u32 i = 0; i < n; i++) ms[1 + i] = aO[i];
    ge_msm(&P, (int)(n + 1), ms, mp); ristretto_encode(pf + 32, &P);
    ms[0] = s_b;
    for (u32 i = 0; i < n; i++) { ms[1 + i] = sL[i]; ms[1 + n + i] = sR[i]; }
    ge_msm(&P, (int)(2 * n + 1), ms, mp); ristretto_encode(pf + 64, &P);
  }
  ts_append_point(&t, "A_I1", pf); ts_append_point(&t, "A_O1", pf + 32); ts_append_point(&t, "S1", pf + 64);
  ts_append(&t, "dom-sep", "r1cs-1phase", 11);
  memset(pf + 96, 0, 96);
  ts_append_point(&t, "A_I2", pf + 96); ts_append_point(&t, "A_O2", pf + 128); ts_append_point(&t, "S2", pf + 160);
  sc y, z; ts_challenge_scalar(&t, "y", &y); ts_challenge_scalar(&t, "z", &z);
  sc *wV = malloc(sizeof(sc) * (m + 1)), wc;
  flatten(c, &z, wL, wR, wO, wV, &wc);
  sc y_inv; sc_invert(&y_inv, &y);
  sc one = {{1, 0, 0, 0}}, zero = {{0, 0, 0, 0}};
  eyi[0] = one; for (u32 i = 1; i < N; i++) sc_mul(&eyi[i], &eyi[i - 1], &y_inv);
  /* l1,l2,l3,r0,r1,r3 held in scratch */
  sc *l1 = malloc(sizeof(sc) * (n + 1) * 6), *l2 = l1 + n, *l3 = l2 + n, *r0 = l3 + n, *r1 = r0 + n, *r3 = r1 + n;
  sc ey = one;
  for (u32 i = 0; i < n; i++) {
    sc x;
    sc_mul(&x, &eyi[i], &wR[i]); sc_add(&l1[i], &aL[i], &x);
    l2[i] = aO[i]; l3[i] = sL[i];
    sc_sub(&r0[i], &wO[i], &ey);
    sc_mul(&x, &ey, &aR[i]); sc_add(&r1[i], &x, &wL[i]);
    sc_mul(&r3[i], &ey, &sR[i]);
    sc_mul(&ey, &ey, &y);
  }
  sc tp[7], a_, b_;
  sc_ip(&tp[1], l1, r0, n);
  sc_ip(&a_, l1, r1, n); sc_ip(&b_, l2, r0, n); sc_add(&tp[2], &a_, &b_);
  sc_ip(&a_, l2, r1, n); sc_ip(&b_, l3, r0, n); sc_add(&tp[3], &a_, &b_);
  sc_ip(&a_, l1, r3, n); sc_ip(&b_, l3, r1, n); sc_add(&tp[4], &a_, &b_);
  sc_ip(&tp[5], l2, r3, n);
  sc_ip(&tp[6], l3, r3, n);
  sc tb[7];
  static const int TJ[5] = {1, 3, 4, 5, 6};
  for (int j = 0; j < 5; j++) trng_scalar(&rng, &tb[TJ[j]]);
  for (int j = 0; j < 5; j++) {
    ge P; pc_commit(&P, &tp[TJ[j]], &tb[TJ[j]]); ristretto_encode(pf + 192 + 32 * j, &P);
  }
  static const char *TL[5] = {"T_1", "T_3", "T_4", "T_5", "T_6"};
  for (int j = 0; j < 5; j++) ts_append_point(&t, TL[j], pf + 192 + 32 * j);
  sc u, x; ts_challenge_scalar(&t, "u", &u); ts_challenge_scalar(&t, "x", &x);
  sc_ip(&tb[2], wV, vbl, m);
  sc xp[7]; xp[0] = one; for (int j = 1; j < 7; j++) sc_mul(&xp[j], &xp[j - 1], &x);
  sc t_x = zero, t_xb = zero;
  for (int j = 1; j < 7; j++) { sc q; sc_mul(&q, &tp[j], &xp[j]); sc_add(&t_x, &t_x, &q); sc_mul(&q, &tb[j], &xp[j]); sc_add(&t_xb, &t_xb, &q); }
  for (u32 i = 0; i < n; i++) {
    sc q, acc;
    sc_mul(&acc, &l1[i], &xp[1]); sc_mul(&q, &l2[i], &xp[2]); sc_add(&acc, &acc, &q); sc_mul(&q, &l3[i], &xp[3]); sc_add(&lv[i], &acc, &q);
    sc_mul(&q, &r1[i], &xp[1]); sc_add(&acc, &r0[i], &q); sc_mul(&q, &r3[i], &xp[3]); sc_add(&rv[i], &acc, &q);
  }
  for (u32 i = n; i < N; i++) { lv[i] = zero; sc_neg(&rv[i], &ey); sc_mul(&ey, &ey, &y); }
  sc e_b;
  { sc q; sc_mul(&q, &x, &s_b); sc_add(&q, &q, &o_b); sc_mul(&q, &q, &x); sc_add(&q, &q, &i_b); sc_mul(&e_b, &q, &x); }
  sc_tobytes(pf + 352, &t_x); sc_tobytes(pf + 384, &t_xb); sc_tobytes(pf + 416, &e_b);
  ts_append_scalar(&t, "t_x", &t_x); ts_append_scalar(&t, "t_x_blinding", &t_xb); ts_append_scalar(&t, "e_blinding", &e_b);
  sc w; ts_challenge_scalar(&t, "w", &w);
  ge Q; ge_scalarmult(&Q, &w, &PC_B);
  sc *Gf = tmp, *Hf = malloc(sizeof(sc) * N);
  for (u32 i = 0; i < N; i++) { Gf[i] = i < n ? one : u; sc_mul(&Hf[i], &eyi[i], &Gf[i]); }
  ge *G = malloc(sizeof(ge) * N), *H = malloc(sizeof(ge) * N);
  memcpy(G, GENS_G, sizeof(ge) * N); memcpy(H, GENS_H, sizeof(ge) * N);
  ipa_create(&t, &Q, Gf, Hf, G, H, lv, rv, N, pf + 448);
  (void)k;
  free(G); free(H); free(Hf); free(l1); free(wV); free(ms); free(mp); free(aL); free(v); free(vbl);
  return BPO_OK;
}

int bpo_prove(u32 n, u32 m, u32 q, const u32 *cons_ptr, const u8 *kind, const u32 *idx, const u8 *coeff,
              const u8 *aL, const u8 *aR, const u8 *aO, const u8 *v, const u8 *vbl,
              const u8 *label, u32 label_len, const u8 *entropy, u32 gens_capacity, u8 *V_out, u8 *proof) {
  circuit c = {n, m, q, cons_ptr, kind, idx, coeff};
  bpo_init();
  return prove_one(&c, aL, aR, aO, v, vbl, label, label_len, entropy, gens_capacity, V_out, proof);
}

/* Verifier::verify */
int bpo_verify(u32 n, u32 m, u32 q, const u32 *cons_ptr, const u8 *kind, const u32 *idx, const u8 *coeff,
               const u8 *V, const u8 *proof, size_t proof_len, const u8 *label, u32 label_len, const u8 *entropy, u32 gens_capacity) {
  bpo_init();
  circuit c = {n, m, q, cons_ptr, kind, idx, coeff};
  u32 N = next_pow2(n ? n : 1), k = ilog2(N);
  if (proof_len % 32 || proof_len < 32 * 16) return BPO_ERR_FORMAT;
  if (gens_capacity < N) return BPO_ERR_GENS;
  if (proof_len != 32 * (size_t)(14 + 2 * k + 2)) return BPO_ERR_VERIFY;
  bpo_ensure_gens(N);
  const u8 *pf = proof;
  sc t_x, t_xb, e_b, pa, pb;
  if (!sc_from_canonical(&t_x, pf + 352) || !sc_from_canonical(&t_xb, pf + 384) || !sc_from_canonical(&e_b, pf + 416)) return BPO_ERR_FORMAT;
  const u8 *ipp = pf + 448;
  if (!sc_from_canonical(&pa, ipp + 64 * k) || !sc_from_canonical(&pb, ipp + 64 * k + 32)) return BPO_ERR_FORMAT;
  transcript t; ts_init(&t, label, label_len);
  ts_append(&t, "dom-sep", "r1cs v1", 7);
  for (u32 i = 0; i < m; i++) ts_append_point(&t, "V", V + 32 * i);
  ts_append_u64(&t, "m", m);
  if (!ts_validate_and_append_point(&t, "A_I1", pf) || !ts_validate_and_append_point(&t, "A_O1", pf + 32) ||
      !ts_validate_and_append_point(&t, "S1", pf + 64)) return BPO_ERR_VERIFY;
  ts_append(&t, "dom-sep", "r1cs-1phase", 11);
  ts_append_point(&t, "A_I2", pf + 96); ts_append_point(&t, "A_O2", pf + 128); ts_append_point(&t, "S2", pf + 160);
  sc y, z; ts_challenge_scalar(&t, "y", &y); ts_challenge_scalar(&t, "z", &z);
  static const char *TL[5] = {"T_1", "T_3", "T_4", "T_5", "T_6"};
  for (int j = 0; j < 5; j++) if (!ts_validate_and_append_point(&t, TL[j], pf + 192 + 32 * j)) return BPO_ERR_VERIFY;
  sc u, x; ts_challenge_scalar(&t, "u", &u); ts_challenge_scalar(&t, "x", &x);
  ts_append_scalar(&t, "t_x", &t_x); ts_append_scalar(&t, "t_x_blinding", &t_xb); ts_append_scalar(&t, "e_blinding", &e_b);
  sc w; ts_challenge_scalar(&t, "w", &w);
  sc *wL = malloc(sizeof(sc) * (N + 1) * 6), *wR = wL + N, *wO = wR + N, *s = wO + N, *yinv = s + N, *ynwR = yinv + N;
  sc *wV = malloc(sizeof(sc) * (m + 1)), wc;
  sc zero = {{0, 0, 0, 0}}, one = {{1, 0, 0, 0}};
  flatten(&c, &z, wL, wR, wO, wV, &wc);
  for (u32 i = n; i < N; i++) wL[i] = wR[i] = wO[i] = zero;
  ts_append(&t, "dom-sep", "ipp v1", 6);
  ts_append_u64(&t, "n", N);
  sc ch[32], chi[32], chsq[32], chisq[32];
  int rc = BPO_OK;
  for (u32 j = 0; j < k; j++) {
    if (!ts_validate_and_append_point(&t, "L", ipp + 64 * j) || !ts_validate_and_append_point(&t, "R", ipp + 64 * j + 32)) { rc = BPO_ERR_VERIFY; goto done; }
    ts_challenge_scalar(&t, "u", &ch[j]);
    sc_invert(&chi[j], &ch[j]); sc_mul(&chsq[j], &ch[j], &ch[j]); sc_mul(&chisq[j], &chi[j], &chi[j]);
  }
  {
    sc allinv = one;
    for (u32 j = 0; j < k; j++) sc_mul(&allinv, &allinv, &chi[j]);
    s[0] = allinv;
    for (u32 i = 1; i < N; i++) { u32 lg = 31 - (u32)__builtin_clz(i); sc_mul(&s[i], &s[i - (1u << lg)], &chsq[(k - 1) - lg]); }
    sc y_inv; sc_invert(&y_inv, &y);
    yinv[0] = one; for (u32 i = 1; i < N; i++) sc_mul(&yinv[i], &yinv[i - 1], &y_inv);
    sc delta = zero;
    for (u32 i = 0; i < N; i++) { if (i < n) { sc_mul(&ynwR[i], &wR[i], &yinv[i]); sc q; sc_mul(&q, &ynwR[i], &wL[i]); sc_add(&delta, &delta, &q); } else ynwR[i] = zero; }
    transcript_rng rng; trng_begin(&rng, &t); trng_finalize(&rng, entropy);
    sc r; trng_scalar(&rng, &r);
    u32 total = 6 + m + 5 + 2 + 2 * N + 2 * k;
    sc *ms = malloc(sizeof(sc) * total); ge *mp = malloc(sizeof(ge) * total);
    sc xx, xxx, rxx; sc_mul(&xx, &x, &x); sc_mul(&xxx, &xx, &x); sc_mul(&rxx, &r, &xx);
    u32 o = 0;
    ms[0] = x; ms[1] = xx; ms[2] = xxx; sc_mul(&ms[3], &u, &x); sc_mul(&ms[4], &u, &xx); sc_mul(&ms[5], &u, &xxx);
    for (int j = 0; j < 6; j++) { if (!ristretto_decode(&mp[j], pf + 32 * j)) rc = BPO_ERR_VERIFY; }
    o = 6;
    for (u32 i = 0; i < m; i++) { sc_mul(&ms[o], &wV[i], &rxx); if (!ristretto_decode(&mp[o], V + 32 * i)) rc = BPO_ERR_VERIFY; o++; }
    sc rx; sc_mul(&rx, &r, &x);
    ms[o] = rx; sc_mul(&ms[o + 1], &rxx, &x); sc_mul(&ms[o + 2], &rxx, &xx); sc_mul(&ms[o + 3], &rxx, &xxx); sc_mul(&ms[o + 4], &ms[o + 2], &xx);
    for (int j = 0; j < 5; j++) { if (!ristretto_decode(&mp[o + j], pf + 192 + 32 * j)) rc = BPO_ERR_VERIFY; }
    o += 5;
    { /* B: w*(t_x - a*b) + r*(x^2*(wc+delta) - t_x) */
      sc ab, q1, q2; sc_mul(&ab, &pa, &pb); sc_sub(&q1, &t_x, &ab); sc_mul(&q1, &q1, &w);
      sc_add(&q2, &wc, &delta); sc_mul(&q2, &q2, &xx); sc_sub(&q2, &q2, &t_x); sc_mul(&q2, &q2, &r);
      sc_add(&ms[o], &q1, &q2); mp[o] = PC_B; o++;
      sc_mul(&q1, &r, &t_xb); sc_add(&q1, &q1, &e_b); sc_neg(&ms[o], &q1); mp[o] = PC_BB; o++;
    }
    for (u32 i = 0; i < N; i++) { /* g */
      sc uf = i < n ? one : u, q1, q2;
      sc_mul(&q1, &x, &ynwR[i]); sc_mul(&q2, &pa, &s[i]); sc_sub(&q1, &q1, &q2); sc_mul(&ms[o], &uf, &q1); mp[o] = GENS_G[i]; o++;
    }
    for (u32 i = 0; i < N; i++) { /* h */
      sc uf = i < n ? one : u, q1, q2;
      sc_mul(&q1, &x, &wL[i]); sc_add(&q1, &q1, &wO[i]); sc_mul(&q2, &pb, &s[N - 1 - i]); sc_sub(&q1, &q1, &q2);
      sc_mul(&q1, &q1, &yinv[i]); sc_sub(&q1, &q1, &one); sc_mul(&ms[o], &uf, &q1); mp[o] = GENS_H[i]; o++;
    }
    for (u32 j = 0; j < k; j++) { ms[o] = chsq[j]; if (!ristretto_decode(&mp[o], ipp + 64 * j)) rc = BPO_ERR_VERIFY; o++; }
    for (u32 j = 0; j < k; j++) { ms[o] = chisq[j]; if (!ristretto_decode(&mp[o], ipp + 64 * j + 32)) rc = BPO_ERR_VERIFY; o++; }
    if (rc == BPO_OK) {
      /* identity points (A_I2 etc. are 32 zero bytes = identity encoding) decode fine */
      ge chk; ge_msm(&chk, (int)o, ms, mp);
      if (!ge_is_identity_ristretto(&chk)) rc = BPO_ERR_VERIFY;
    }
    free(ms); free(mp);
  }
done:
  free(wL); free(wV);
  return rc;
}

/* ------------------------------------------------------------------ native witnesses */
typedef struct { u32 width, fb, pr, fe; sc mds[36]; sc rk[960]; } poseidon_params;
static poseidon_params PP;

int bpo_poseidon_set_params(const u8 *consts /* 36 MDS then round keys, 32B LE canonical */, u32 nconst, u32 width, u32 fb, u32 fe, u32 pr) {
  bpo_init();
  if (width != 6 || nconst < 36 + (fb + pr + fe) * width || (fb + pr + fe) * width > 960) return BPO_ERR_ARG;
  PP.width = width; PP.fb = fb; PP.pr = pr; PP.fe = fe;
  for (u32 i = 0; i < 36; i++) memcpy(&PP.mds[i], consts + 32 * i, 32);
  for (u32 i = 0; i < (fb + pr + fe) * width; i++) memcpy(&PP.rk[i], consts + 32 * (36 + i), 32);
  return 0;
}
static void sbox_apply(sc *r, const sc *x, int inverse) {
  if (inverse) { sc_invert(r, x); } else { sc t; sc_mul(&t, x, x); sc_mul(r, &t, x); }
}
/* gadget_poseidon.rs:189-280; optional trace receives every S-box input in circuit order */
static void poseidon_perm(sc st[6], int inverse, sc *trace_in, u32 *ntrace) {
  u32 off = 0, total = PP.fb + PP.pr + PP.fe, tr = 0;
  for (u32 rnd = 0; rnd < total; rnd++) {
    int full = rnd < PP.fb || rnd >= PP.fb + PP.pr;
    for (u32 i = 0; i < 6; i++) {
      sc_add(&st[i], &st[i], &PP.rk[off++]);
      if (full || i == 5) { if (trace_in) trace_in[tr++] = st[i]; sbox_apply(&st[i], &st[i], inverse); }
    }
    sc nx[6];
    for (u32 i = 0; i < 6; i++) {
      sc acc = {{0, 0, 0, 0}}, t;
      for (u32 j = 0; j < 6; j++) { sc_mul(&t, &st[j], &PP.mds[i * 6 + j]); sc_add(&acc, &acc, &t); }
      nx[i] = acc;
    }
    memcpy(st, nx, sizeof nx);
  }
  if (ntrace) *ntrace = tr;
}
void bpo_poseidon_perm(const u8 *in, int inverse, u8 *out) {
  sc st[6]; for (int i = 0; i < 6; i++) sc_from_bytes_mod_order(&st[i], in + 32 * i);
  poseidon_perm(st, inverse, NULL, NULL);
  for (int i = 0; i < 6; i++) sc_tobytes(out + 32 * i, &st[i]);
}
static void poseidon_hash2(sc *out, const sc *xl, const sc *xr, int inverse, sc *trace_in, u32 *ntrace) {
  sc st[6]; memset(st, 0, sizeof st);
  st[1] = *xl; st[2] = *xr; sc_from_u64(&st[3], 101);
  poseidon_perm(st, inverse, trace_in, ntrace);
  *out = st[1];
}
void bpo_poseidon_hash2(const u8 xl[32], const u8 xr[32], int inverse, u8 out[32]) {
  sc a, b, r; sc_from_bytes_mod_order(&a, xl); sc_from_bytes_mod_order(&b, xr);
  poseidon_hash2(&r, &a, &b, inverse, NULL, NULL); sc_tobytes(out, &r);
}
/* multipliers emitted by one S-box, in circuit order (gadget_poseidon.rs:141-185, gadget_zero_nonzero.rs:46-66) */
static u32 sbox_multipliers(const sc *x, int inverse, u8 *aL, u8 *aR, u8 *aO, u32 at) {
  sc zero = {{0, 0, 0, 0}};
  if (inverse) {
    sc inv, o; sc_invert(&inv, x); sc_mul(&o, x, &inv);
    sc_tobytes(aL + 32 * at, x); sc_tobytes(aR + 32 * at, &inv); sc_tobytes(aO + 32 * at, &o); at++;
    sc_tobytes(aL + 32 * at, x); sc_tobytes(aR + 32 * at, &zero); sc_tobytes(aO + 32 * at, &zero); at++;
    sc_tobytes(aL + 32 * at, x); sc_tobytes(aR + 32 * at, &inv); sc_tobytes(aO + 32 * at, &o); at++;
  } else {
    sc sq, cu; sc_mul(&sq, x, x); sc_mul(&cu, &sq, x);
    sc_tobytes(aL + 32 * at, x); sc_tobytes(aR + 32 * at, x); sc_tobytes(aO + 32 * at, &sq); at++;
    sc_tobytes(aL + 32 * at, &sq); sc_tobytes(aR + 32 * at, x); sc_tobytes(aO + 32 * at, &cu); at++;
  }
  return at;
}
/* Poseidon 2:1 preimage circuit witness (gadget_poseidon.rs:691-785): returns multipliers written */
u32 bpo_poseidon_hash2_witness(const u8 xl[32], const u8 xr[32], int inverse, u8 *aL, u8 *aR, u8 *aO, u8 out_hash[32]) {
  sc a, b, h; sc_from_bytes_mod_order(&a, xl); sc_from_bytes_mod_order(&b, xr);
  sc tr[188]; u32 nt = 0, at = 0;
  poseidon_hash2(&h, &a, &b, inverse, tr, &nt);
  for (u32 s = 0; s < nt; s++) at = sbox_multipliers(&tr[s], inverse, aL, aR, aO, at);
  sc_tobytes(out_hash, &h);
  return at;
}
/* VSMT-2 membership circuit witness (gadget_vsmt_2.rs:171-209), inverse S-box; returns multipliers written */
u32 bpo_vsmt2_witness(u32 depth, const u8 leaf[32], const u8 *bits, const u8 *sibs, u8 *aL, u8 *aR, u8 *aO, u8 root[32]) {
  sc cur; sc_from_bytes_mod_order(&cur, leaf);
  sc one = {{1, 0, 0, 0}};
  u32 at = 0;
  for (u32 lv = 0; lv < depth; lv++) {
    sc b, omb, sib, l1, l2, r1, r2, left, right;
    sc_from_u64(&b, bits[lv]); sc_sub(&omb, &one, &b);
    sc_from_bytes_mod_order(&sib, sibs + 32 * lv);
    sc_mul(&l1, &omb, &cur); sc_mul(&l2, &b, &sib); sc_mul(&r1, &b, &cur); sc_mul(&r2, &omb, &sib);
    sc_tobytes(aL + 32 * at, &omb); sc_tobytes(aR + 32 * at, &cur); sc_tobytes(aO + 32 * at, &l1); at++;
    sc_tobytes(aL + 32 * at, &b); sc_tobytes(aR + 32 * at, &sib); sc_tobytes(aO + 32 * at, &l2); at++;
    sc_tobytes(aL + 32 * at, &b); sc_tobytes(aR + 32 * at, &cur); sc_tobytes(aO + 32 * at, &r1); at++;
    sc_tobytes(aL + 32 * at, &omb); sc_tobytes(aR + 32 * at, &sib); sc_tobytes(aO + 32 * at, &r2); at++;
    sc_add(&left, &l1, &l2); sc_add(&right, &r1, &r2);
    sc tr[188]; u32 nt = 0;
    poseidon_hash2(&cur, &left, &right, 1, tr, &nt);
    for (u32 s = 0; s < nt; s++) at = sbox_multipliers(&tr[s], 1, aL, aR, aO, at);
  }
  sc_tobytes(root, &cur);
  return at;
}
/* MiMC (gadget_mimc.rs:19-39 native, :55-79 circuit order) */
u32 bpo_mimc_witness(const u8 xl_[32], const u8 xr_[32], u32 rounds, const u8 *constants, u8 *aL, u8 *aR, u8 *aO, u8 image[32]) {
  sc xl, xr; sc_from_bytes_mod_order(&xl, xl_); sc_from_bytes_mod_order(&xr, xr_);
  u32 at = 0;
  for (u32 j = 0; j < rounds; j++) {
    sc c, t, sq, cu, nl;
    sc_from_bytes_mod_order(&c, constants + 32 * j);
    sc_add(&t, &xl, &c); sc_mul(&sq, &t, &t); sc_mul(&cu, &sq, &t);
    if (aL) {
      sc_tobytes(aL + 32 * at, &t); sc_tobytes(aR + 32 * at, &t); sc_tobytes(aO + 32 * at, &sq); at++;
      sc_tobytes(aL + 32 * at, &sq); sc_tobytes(aR + 32 * at, &t); sc_tobytes(aO + 32 * at, &cu); at++;
    }
    sc_add(&nl, &cu, &xr); xr = xl; xl = nl;
  }
  sc_tobytes(image, &xl);
  return at;
}

/* ------------------------------------------------------------------ batched CPU baseline (threads over proofs) */
typedef struct {
  const circuit *c; u32 B, depth, nthreads, tid, gens_capacity;
  const u8 *aL, *aR, *aO;           /* [B][n][32] or NULL when witness_kind != 0 */
  const u8 *v, *vbl, *entropy;      /* [B][m][32], [B][m][32], [B][32] */
  const u8 *label; u32 label_len;
  int witness_kind;                 /* 0 = supplied, 1 = vsmt2(depth): v = leaf, bits.., sibs.., statics */
  u8 *V_out, *proofs; size_t proof_stride; int *status;
} batch_job;

static void *batch_worker(void *arg) {
  batch_job *j = arg;
  u32 n = j->c->n, m = j->c->m;
  u8 *wa = NULL;
  if (j->witness_kind) wa = malloc((size_t)n * 32 * 3);
  for (u32 p = j->tid; p < j->B; p += j->nthreads) {
    const u8 *aL, *aR, *aO;
    const u8 *v = j->v + (size_t)p * m * 32;
    if (j->witness_kind == 1) {
      u8 bits[256], root[32];
      for (u32 d = 0; d < j->depth; d++) bits[d] = v[32 * (1 + d)];
      bpo_vsmt2_witness(j->depth, v, bits, v + 32 * (1 + j->depth), wa, wa + (size_t)n * 32, wa + (size_t)n * 64, root);
      aL = wa; aR = wa + (size_t)n * 32; aO = wa + (size_t)n * 64;
    } else {
      aL = j->aL + (size_t)p * n * 32; aR = j->aR + (size_t)p * n * 32; aO = j->aO + (size_t)p * n * 32;
    }
    j->status[p] = prove_one(j->c, aL, aR, aO, v, j->vbl + (size_t)p * m * 32, j->label, j->label_len,
                             j->entropy + 32 * (size_t)p, j->gens_capacity, j->V_out + (size_t)p * m * 32, j->proofs + (size_t)p * j->proof_stride);
  }
  free(wa);
  return NULL;
}

int bpo_prove_batch(u32 n, u32 m, u32 q, const u32 *cons_ptr, const u8 *kind, const u32 *idx, const u8 *coeff,
                    u32 B, int witness_kind, u32 depth, const u8 *aL, const u8 *aR, const u8 *aO, const u8 *v, const u8 *vbl,
                    const u8 *label, u32 label_len, const u8 *entropy, u32 gens_capacity, u32 nthreads,
                    u8 *V_out, u8 *proofs, size_t proof_stride, int *status) {
  bpo_init();
  circuit c = {n, m, q, cons_ptr, kind, idx, coeff};
  bpo_ensure_gens(next_pow2(n ? n : 1));
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256]; batch_job jobs[256];
  for (u32 t = 0; t < nthreads; t++) {
    jobs[t] = (batch_job){&c, B, depth, nthreads, t, gens_capacity, aL, aR, aO, v, vbl, entropy, label, label_len, witness_kind, V_out, proofs, proof_stride, status};
    pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
  }
  for (u32 t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  return 0;
}
