/* Compiled ABI check: a plain-C translation unit that includes include/bp_b200.h and calls the library the
 * way a foreign-language binding would -- header and library cannot drift without this failing to compile, link or run.
 * The circuit is the reference's smallest one, `factors` (src/factors.rs:12-21): commit p, q, r; (_, _, o) = multiply(p, q);
 * constrain(o - r).  Then: prove, verify, reject a wrong product, and the same statement through the batched circuit API
 * (witness program + bp_prove_batch / bp_verify_batch / bp_verify_batch_combined).
 * Built by tests/test_capi.py with  gcc -std=c11 -Wall -Wextra -Werror  against the emulation build on the CPU and against
 * libbp_b200.so on the GPU box. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bp_b200.h"

#define CHECK(x) do { int32_t rc_ = (x); if (rc_ != BP_OK) { fprintf(stderr, "capi_smoke: %s -> %d (line %d)\n", #x, (int)rc_, __LINE__); return 1; } } while (0)
#define EXPECT(x) do { if (!(x)) { fprintf(stderr, "capi_smoke: expectation failed: %s (line %d)\n", #x, __LINE__); return 1; } } while (0)

static void scalar_u64(uint8_t out[32], uint64_t x) { memset(out, 0, 32); for (int i = 0; i < 8; i++) out[i] = (uint8_t)(x >> (8 * i)); }
static bp_term term(bp_var v, int64_t c) {  /* small signed coefficient mod l: -1 is l - 1 */
  static const uint8_t L_MINUS_1[32] = {0xec, 0xd3, 0xf5, 0x5c, 0x1a, 0x63, 0x12, 0x58, 0xd6, 0x9c, 0xf7, 0xa2, 0xde, 0xf9, 0xde, 0x14,
                                        0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0x10};
  bp_term t; t.var = v;
  if (c >= 0) scalar_u64(t.coeff, (uint64_t)c); else memcpy(t.coeff, L_MINUS_1, 32);
  return t;
}

/* factors gadget (reference src/factors.rs:12-21) against any constraint system */
static int32_t factors_gadget(bp_cs *cs, bp_var p, bp_var q, bp_var r) {
  bp_term lp = term(p, 1), lq = term(q, 1);
  bp_var out[3];
  int32_t rc = bp_cs_multiply(cs, &lp, 1, &lq, 1, out);
  if (rc) return rc;
  bp_term lc[2] = {term(out[2], 1), term(r, -1)};
  return bp_cs_constrain(cs, lc, 2);
}

int main(void) {
  static const uint8_t label[] = "Factors";
  const size_t ll = sizeof label - 1;
  bp_gens *g = NULL;
  CHECK(bp_gens_new(16, &g));
  EXPECT(bp_gens_capacity(g) == 16 && bp_version() > 0);
  uint8_t vals[3][32], blind[3][32], V[3][32], entropy[32];
  scalar_u64(vals[0], 1000003); scalar_u64(vals[1], 998244353); scalar_u64(vals[2], 1000003ull * 998244353ull);
  for (int i = 0; i < 3; i++) scalar_u64(blind[i], 0x1234567 + 77 * (uint64_t)i);
  memset(entropy, 0x5a, 32);

  /* prover (reference src/factors.rs:56-75) */
  bp_cs *pr = NULL; bp_var pv[3];
  CHECK(bp_prover_new(g, label, ll, &pr));
  for (int i = 0; i < 3; i++) CHECK(bp_prover_commit(pr, vals[i], blind[i], V[i], &pv[i]));
  CHECK(factors_gadget(pr, pv[0], pv[1], pv[2]));
  EXPECT(bp_cs_num_multipliers(pr) == 1 && bp_cs_num_constraints(pr) == 3 && bp_cs_num_commitments(pr) == 3);
  uint8_t proof[2048]; size_t plen = sizeof proof;
  CHECK(bp_prover_prove(pr, entropy, proof, &plen));
  EXPECT(plen == bp_cs_proof_len(pr) && plen == 32 * 16);
  bp_cs_free(pr);
  uint8_t chk[32];
  CHECK(bp_pc_commit(g, 1, vals[0], blind[0], chk));
  EXPECT(memcmp(chk, V[0], 32) == 0);

  /* verifier (reference src/factors.rs:78-100): accepts; a different r commitment is rejected */
  for (int wrong = 0; wrong < 2; wrong++) {
    bp_cs *vf = NULL; bp_var vv[3];
    CHECK(bp_verifier_new(g, label, ll, &vf));
    for (int i = 0; i < 3; i++) CHECK(bp_verifier_commit(vf, (wrong && i == 2) ? V[0] : V[i], &vv[i]));
    CHECK(factors_gadget(vf, vv[0], vv[1], vv[2]));
    int32_t rc = bp_verifier_verify(vf, proof, plen, entropy);
    EXPECT(rc == (wrong ? BP_ERR_VERIFICATION : BP_OK));
    bp_cs_free(vf);
  }

  /* the same circuit compiled once and proved as a batch of 3 (witness program recorded by multiply) */
  bp_cs *rec = NULL; bp_var rv[3]; uint8_t zero[32] = {0};
  CHECK(bp_verifier_new(g, label, ll, &rec));
  for (int i = 0; i < 3; i++) CHECK(bp_verifier_commit(rec, zero, &rv[i]));
  CHECK(factors_gadget(rec, rv[0], rv[1], rv[2]));
  bp_circuit *c = NULL;
  CHECK(bp_circuit_compile(rec, &c));
  bp_cs_free(rec);
  EXPECT(bp_circuit_num_multipliers(c) == 1 && bp_circuit_num_commitments(c) == 3 && bp_circuit_proof_len(c) == plen && bp_circuit_has_witness_program(c));
  enum { B = 3 };
  uint8_t bv[B][3][32], bb[B][3][32], be[B][32], bV[B][3][32], bproofs[B][32 * 16];
  int32_t status[B], combined = -1;
  for (int p = 0; p < B; p++) {
    scalar_u64(bv[p][0], 3 + (uint64_t)p); scalar_u64(bv[p][1], 1000 + 7 * (uint64_t)p); scalar_u64(bv[p][2], (3 + (uint64_t)p) * (1000 + 7 * (uint64_t)p));
    for (int i = 0; i < 3; i++) scalar_u64(bb[p][i], 99 + 3 * (uint64_t)p + (uint64_t)i);
    memset(be[p], 0x11 * (p + 1), 32);
  }
  CHECK(bp_prove_batch(g, c, B, label, ll, &bv[0][0][0], &bb[0][0][0], &be[0][0], NULL, NULL, NULL, NULL, NULL, &bV[0][0][0], &bproofs[0][0], status));
  EXPECT(status[0] == 0 && status[1] == 0 && status[2] == 0);
  CHECK(bp_verify_batch(g, c, B, label, ll, &bV[0][0][0], &bproofs[0][0], &be[0][0], NULL, status));
  EXPECT(status[0] == 0 && status[1] == 0 && status[2] == 0);
  CHECK(bp_verify_batch_combined(g, c, B, label, ll, &bV[0][0][0], &bproofs[0][0], &be[0][0], NULL, status, &combined));
  EXPECT(combined == BP_OK);
  bproofs[1][352] ^= 1;  /* t_x of proof 1 */
  CHECK(bp_verify_batch(g, c, B, label, ll, &bV[0][0][0], &bproofs[0][0], &be[0][0], NULL, status));
  EXPECT(status[0] == 0 && status[1] == BP_ERR_VERIFICATION && status[2] == 0);
  CHECK(bp_verify_batch_combined(g, c, B, label, ll, &bV[0][0][0], &bproofs[0][0], &be[0][0], NULL, status, &combined));
  EXPECT(combined == BP_ERR_VERIFICATION);
  /* wire format round trip */
  uint8_t wire[32 * 16 + 1], back[32 * 16 + 96];
  int64_t wl = bp_proof_to_wire(bproofs[0], plen, wire, sizeof wire);
  EXPECT(wl == (int64_t)plen + 1 - 96);
  EXPECT(bp_proof_from_wire(wire, (size_t)wl, back, sizeof back) == (int64_t)plen && memcmp(back, bproofs[0], plen) == 0);
  bp_circuit_free(c);
  bp_gens_free(g);
  EXPECT(bp_launch_count() > 0);
  printf("capi_smoke ok\n");
  return 0;
}
