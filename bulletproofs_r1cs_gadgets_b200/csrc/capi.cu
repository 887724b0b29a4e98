// C-ABI (include/bp_b200.h) over the recorder (recorder.h) and the device engine (engine.h).
#include "recorder.h"
#include <algorithm>
#include <memory>

// device staging of one stream slot of the host-buffer streaming entry points (bp_prove_stream_*_host)
struct HostStage {
  uint8_t *v = nullptr, *vb = nullptr, *e = nullptr, *aux = nullptr, *pub = nullptr, *V = nullptr, *P = nullptr; int32_t *S = nullptr;
  uint32_t cap = 0, B = 0; bool active = false;
  void release() { void *ps[] = {v, vb, e, aux, pub, V, P, S}; for (void *p : ps) dev_free(p); *this = HostStage(); }
};
struct bp_circuit { BpCircuit *c; HostStage stage[2]; };

static scm load_scalar(const uint8_t *b) { return sc_from_bytes_mod_order(b); }
static LC lc_from_terms(const bp_term *t, size_t n) {
  LC lc; lc.terms.reserve(n);
  for (size_t i = 0; i < n; i++) lc.terms.push_back(Term{t[i].var, load_scalar(t[i].coeff)});
  return lc;
}
static bool var_ok(const bp_cs *cs, bp_var v) {
  switch (v.kind) {
    case BP_VAR_COMMITTED: return v.index < cs->V.size();
    case BP_VAR_MULT_LEFT: case BP_VAR_MULT_RIGHT: case BP_VAR_MULT_OUT: return v.index < cs->num_mult;
    case BP_VAR_ONE: return true;
    case BP_VAR_PUBLIC: return v.index < cs->pub.size();
    default: return false;
  }
}
static bool lc_ok(const bp_cs *cs, const bp_term *t, size_t n) {
  if (n && !t) return false;
  for (size_t i = 0; i < n; i++) if (!var_ok(cs, t[i].var)) return false;
  return true;
}

extern "C" {

int32_t bp_version(void) { return 1; }
int64_t bp_launch_count(void) { return engine_launch_count(); }
void bp_profile_enable(int32_t on) { engine_profile_enable(on); }
int32_t bp_profile_report(char *buf, size_t cap) { return engine_profile_report(buf, cap); }

// ------------------------------------------------------------------------------------------------ generators
int32_t bp_gens_new(uint32_t capacity, bp_gens **out) {
  if (!out) return BP_ERR_INVALID_ARGUMENT;
  BpGens *g = nullptr;
  int rc = gens_create(capacity, &g);
  if (rc) return rc;
  *out = new bp_gens{g};
  return BP_OK;
}
void bp_gens_free(bp_gens *g) { if (g) { gens_free(g->g); delete g; } }
uint32_t bp_gens_capacity(const bp_gens *g) { return g ? g->g->capacity : 0; }
int32_t bp_gens_pedersen(const bp_gens *g, uint8_t B[32], uint8_t Bb[32]) {
  if (!g) return BP_ERR_INVALID_ARGUMENT;
  memcpy(B, g->g->pc_c, 32); memcpy(Bb, g->g->pc_c + 32, 32);
  return BP_OK;
}
int32_t bp_gens_export(const bp_gens *g, int32_t which, uint32_t count, uint8_t *out) {
  if (!g || !out) return BP_ERR_INVALID_ARGUMENT;
  return gens_export(g->g, which, count, out);
}
int32_t bp_pc_commit(const bp_gens *g, uint32_t count, const uint8_t *v, const uint8_t *r, uint8_t *out) {
  if (!g || !v || !r || !out) return BP_ERR_INVALID_ARGUMENT;
  return engine_commit(g->g, (int)count, v, r, out);
}

// ------------------------------------------------------------------------------------------------ tier 1
static int32_t cs_new(const bp_gens *g, const uint8_t *label, size_t n, bool prover, bp_cs **out) {
  if (!g || !out || (n && !label)) return BP_ERR_INVALID_ARGUMENT;
  bp_cs *cs = new bp_cs();
  cs->gens = g; cs->is_prover = prover; cs->label.assign(label, label + n);
  *out = cs;
  return BP_OK;
}
int32_t bp_prover_new(const bp_gens *g, const uint8_t *label, size_t n, bp_cs **out) { return cs_new(g, label, n, true, out); }
int32_t bp_verifier_new(const bp_gens *g, const uint8_t *label, size_t n, bp_cs **out) { return cs_new(g, label, n, false, out); }
void bp_cs_free(bp_cs *cs) { delete cs; }

int32_t bp_prover_commit(bp_cs *cs, const uint8_t v[32], const uint8_t vb[32], uint8_t V_out[32], bp_var *var) {
  if (!cs || !cs->is_prover || !v || !vb || !V_out || !var) return BP_ERR_INVALID_ARGUMENT;
  std::array<uint8_t, 32> V;
  int rc = engine_commit(cs->gens->g, 1, v, vb, V.data());
  if (rc) return rc;
  cs->v.push_back(load_scalar(v)); cs->vbl.push_back(load_scalar(vb)); cs->V.push_back(V);
  memcpy(V_out, V.data(), 32);
  *var = bp_var{BP_VAR_COMMITTED, (uint32_t)cs->V.size() - 1};
  return BP_OK;
}
int32_t bp_verifier_commit(bp_cs *cs, const uint8_t V[32], bp_var *var) {
  if (!cs || cs->is_prover || !V || !var) return BP_ERR_INVALID_ARGUMENT;
  std::array<uint8_t, 32> a; memcpy(a.data(), V, 32);
  cs->V.push_back(a);
  *var = bp_var{BP_VAR_COMMITTED, (uint32_t)cs->V.size() - 1};
  return BP_OK;
}
int32_t bp_cs_multiply(bp_cs *cs, const bp_term *l, size_t nl, const bp_term *r, size_t nr, bp_var out[3]) {
  if (!cs || !out || !lc_ok(cs, l, nl) || !lc_ok(cs, r, nr)) return BP_ERR_INVALID_ARGUMENT;
  cs->multiply(lc_from_terms(l, nl), lc_from_terms(r, nr), out);
  return BP_OK;
}
int32_t bp_cs_allocate_multiplier(bp_cs *cs, const uint8_t *l, const uint8_t *r, bp_var out[3]) {
  if (!cs || !out) return BP_ERR_INVALID_ARGUMENT;
  if (l && r) { scm a = load_scalar(l), b = load_scalar(r); return cs->allocate_multiplier(&a, &b, out); }
  return cs->allocate_multiplier(nullptr, nullptr, out);
}
int32_t bp_cs_allocate_single(bp_cs *cs, const uint8_t *value, bp_var *var, bp_var *out_var, int32_t *has_out) {
  if (!cs || !var) return BP_ERR_INVALID_ARGUMENT;
  int has = 0; int rc;
  if (value) { scm a = load_scalar(value); rc = cs->allocate_single(0, &a, nullptr, var, out_var, &has); }
  else rc = cs->allocate_single(0, nullptr, nullptr, var, out_var, &has);
  if (has_out) *has_out = has;
  return rc;
}
int32_t bp_cs_evaluate_lc(bp_cs *cs, const bp_term *lc, size_t n, uint8_t out[32]) {
  if (!cs || !out || !lc_ok(cs, lc, n)) return BP_ERR_INVALID_ARGUMENT;
  scm v;
  if (!cs->eval(lc_from_terms(lc, n), v)) return BP_ERR_MISSING_ASSIGNMENT;
  sc_tobytes(out, v);
  return BP_OK;
}
int32_t bp_cs_public_input(bp_cs *cs, const uint8_t *value, bp_var *var) {
  if (!cs || !var) return BP_ERR_INVALID_ARGUMENT;
  cs->pub.push_back(value ? load_scalar(value) : sc_zero());
  *var = bp_var{BP_VAR_PUBLIC, (uint32_t)cs->pub.size() - 1};
  return BP_OK;
}
int32_t bp_cs_constrain(bp_cs *cs, const bp_term *lc, size_t n) {
  if (!cs || !lc_ok(cs, lc, n)) return BP_ERR_INVALID_ARGUMENT;
  cs->constrain(lc_from_terms(lc, n));
  return BP_OK;
}
uint64_t bp_cs_num_constraints(const bp_cs *cs) { return cs ? cs->num_constraints() : 0; }
uint64_t bp_cs_num_multipliers(const bp_cs *cs) { return cs ? cs->num_mult : 0; }
uint64_t bp_cs_num_commitments(const bp_cs *cs) { return cs ? cs->V.size() : 0; }
static uint32_t np2(uint32_t n) { uint32_t N = 1; while (N < n) N <<= 1; return N; }
static uint32_t lg2(uint32_t N) { uint32_t k = 0; while ((1u << k) < N) k++; return k; }
size_t bp_cs_proof_len(const bp_cs *cs) { return cs ? 32 * (size_t)(14 + 2 * lg2(np2(cs->num_mult ? cs->num_mult : 1)) + 2) : 0; }

static int32_t compile_cs(const bp_cs *cs, bool with_tape, BpCircuit **out) {
  size_t nnz = cs->terms.size();
  std::vector<uint8_t> kind(nnz ? nnz : 1); std::vector<uint32_t> idx(nnz ? nnz : 1); std::vector<scm> co(nnz ? nnz : 1);
  for (size_t t = 0; t < nnz; t++) { kind[t] = (uint8_t)cs->terms[t].var.kind; idx[t] = cs->terms[t].var.index; co[t] = cs->terms[t].coeff; }
  if (!with_tape || cs->pending >= 0) {
    int rc = circuit_create(cs->num_mult, (uint32_t)cs->V.size(), (uint32_t)cs->pub.size(), (uint32_t)cs->num_constraints(), cs->cons_ptr.data(), kind.data(), idx.data(),
                            co.data(), nullptr, 0, 0, nullptr, nullptr, nullptr, nullptr, nullptr, out);
    if (!rc && !cs->fixed_idx.empty()) {
      rc = circuit_set_fixed_commitments(*out, (uint32_t)cs->fixed_idx.size(), cs->fixed_idx.data(), cs->fixed_V[0].data());
      if (rc) { circuit_free(*out); *out = nullptr; }
    }
    return rc;
  }
  size_t wn = cs->wlc_terms.size();
  std::vector<uint8_t> wkind(wn ? wn : 1); std::vector<uint32_t> widx(wn ? wn : 1); std::vector<scm> wco(wn ? wn : 1);
  for (size_t t = 0; t < wn; t++) { wkind[t] = (uint8_t)cs->wlc_terms[t].var.kind; widx[t] = cs->wlc_terms[t].var.index; wco[t] = cs->wlc_terms[t].coeff; }
  HostPoseidonTape pt{}; std::vector<scm> mds;
  if (!cs->pblocks.empty()) {
    const bp_poseidon_params *pp = cs->pparams.get();
    for (auto &row : pp->mds) mds.insert(mds.end(), row.begin(), row.end());
    pt = HostPoseidonTape{cs->pblocks.data(), (uint32_t)cs->pblocks.size(), pp->round_keys.data(), (uint32_t)pp->round_keys.size(), mds.data(),
                          pp->full_rounds_beginning, pp->partial_rounds, pp->full_rounds_end};
  }
  int rc = circuit_create(cs->num_mult, (uint32_t)cs->V.size(), (uint32_t)cs->pub.size(), (uint32_t)cs->num_constraints(), cs->cons_ptr.data(), kind.data(), idx.data(),
                          co.data(), cs->tape.data(), cs->naux, (uint32_t)cs->wlc_ptr.size() - 1, cs->wlc_ptr.data(), wkind.data(), widx.data(),
                          wco.data(), cs->pblocks.empty() ? nullptr : &pt, out);
  if (!rc && !cs->fixed_idx.empty()) {
    rc = circuit_set_fixed_commitments(*out, (uint32_t)cs->fixed_idx.size(), cs->fixed_idx.data(), cs->fixed_V[0].data());
    if (rc) { circuit_free(*out); *out = nullptr; }
  }
  return rc;
}

struct DevBuf {
  uint8_t *p = nullptr;
  int alloc(size_t n) { return dev_malloc((void **)&p, n); }
  ~DevBuf() { dev_free(p); }
};
static void scalars_to_bytes(std::vector<uint8_t> &out, const std::vector<scm> &v) {
  out.resize(v.size() * 32 + 32);
  for (size_t i = 0; i < v.size(); i++) sc_tobytes(out.data() + 32 * i, v[i]);
}

int32_t bp_prover_prove(bp_cs *cs, const uint8_t entropy[32], uint8_t *proof, size_t *proof_len) {
  if (!cs || !cs->is_prover || !entropy || !proof || !proof_len) return BP_ERR_INVALID_ARGUMENT;
  if (cs->pending >= 0) return BP_ERR_MISSING_ASSIGNMENT;
  size_t plen = bp_cs_proof_len(cs);
  if (*proof_len < plen) { *proof_len = plen; return BP_ERR_INVALID_ARGUMENT; }
  const BpGens *g = cs->gens->g;
  uint32_t n = cs->num_mult, m = (uint32_t)cs->V.size();
  if (g->capacity < n || g->capacity < np2(n ? n : 1)) return BP_ERR_INVALID_GENERATORS_LENGTH;
  BpCircuit *c = nullptr;
  int rc = compile_cs(cs, false, &c);
  if (rc) return rc;
  std::vector<uint8_t> hv, hvb, haL, haR, haO;
  scalars_to_bytes(hv, cs->v); scalars_to_bytes(hvb, cs->vbl); scalars_to_bytes(haL, cs->aL); scalars_to_bytes(haR, cs->aR); scalars_to_bytes(haO, cs->aO);
  DevBuf dv, dvb, de, daL, daR, daO, dV, dP, dS;
  if (dv.alloc(m * 32 + 32) || dvb.alloc(m * 32 + 32) || de.alloc(32) || daL.alloc(n * 32 + 32) || daR.alloc(n * 32 + 32) || daO.alloc(n * 32 + 32) ||
      dV.alloc(m * 32 + 32) || dP.alloc(plen) || dS.alloc(sizeof(int))) { circuit_free(c); return BP_ERR_OOM; }
  dev_stream s = 0;
  dev_h2d(dv.p, hv.data(), m * 32, s); dev_h2d(dvb.p, hvb.data(), m * 32, s); dev_h2d(de.p, entropy, 32, s);
  dev_h2d(daL.p, haL.data(), n * 32, s); dev_h2d(daR.p, haR.data(), n * 32, s); dev_h2d(daO.p, haO.data(), n * 32, s);
  ProveArgs a{}; a.B = 1; a.label = cs->label.data(); a.label_len = (int)cs->label.size();
  a.v = dv.p; a.vbl = dvb.p; a.entropy = de.p; a.aL = daL.p; a.aR = daR.p; a.aO = daO.p; a.aux = nullptr;
  a.V_out = dV.p; a.proofs = dP.p; a.status = (int *)dS.p;
  rc = engine_prove(g, c, a, 1, s);
  int st = 0;
  if (!rc) { dev_d2h(proof, dP.p, plen, s); dev_d2h(&st, dS.p, sizeof st, s); if (dev_sync(s)) rc = BP_ERR_CUDA; }
  circuit_free(c);
  if (rc) return rc;
  *proof_len = plen;
  return st;
}

int32_t bp_verifier_verify(bp_cs *cs, const uint8_t *proof, size_t proof_len, const uint8_t entropy[32]) {
  if (!cs || cs->is_prover || !proof || !entropy) return BP_ERR_INVALID_ARGUMENT;
  if (proof_len % 32 || proof_len < 32 * 16) return BP_ERR_FORMAT;
  const BpGens *g = cs->gens->g;
  uint32_t n = cs->num_mult, m = (uint32_t)cs->V.size();
  if (g->capacity < np2(n ? n : 1)) return BP_ERR_INVALID_GENERATORS_LENGTH;
  if (proof_len != bp_cs_proof_len(cs)) return BP_ERR_VERIFICATION;
  BpCircuit *c = nullptr;
  int rc = compile_cs(cs, false, &c);
  if (rc) return rc;
  DevBuf dV, dP, de, dS, dpub;
  std::vector<uint8_t> hpub; scalars_to_bytes(hpub, cs->pub);
  if (dV.alloc(m * 32 + 32) || dP.alloc(proof_len) || de.alloc(32) || dS.alloc(sizeof(int)) || dpub.alloc(hpub.size())) { circuit_free(c); return BP_ERR_OOM; }
  dev_stream s = 0;
  dev_h2d(dpub.p, hpub.data(), hpub.size(), s);
  for (uint32_t i = 0; i < m; i++) dev_h2d(dV.p + 32 * i, cs->V[i].data(), 32, s);
  dev_h2d(dP.p, proof, proof_len, s); dev_h2d(de.p, entropy, 32, s);
  VerifyArgs a{}; a.B = 1; a.label = cs->label.data(); a.label_len = (int)cs->label.size(); a.V = dV.p; a.proofs = dP.p; a.entropy = de.p; a.pub = dpub.p; a.status = (int *)dS.p;
  rc = engine_verify(g, c, a, s);
  int st = 0;
  if (!rc) { dev_d2h(&st, dS.p, sizeof st, s); if (dev_sync(s)) rc = BP_ERR_CUDA; }
  circuit_free(c);
  return rc ? rc : st;
}

// ------------------------------------------------------------------------------------------------ gadgets
int32_t bp_poseidon_params_new(const uint8_t *constants, size_t nconst, uint32_t width, uint32_t fb, uint32_t fe, uint32_t pr, bp_poseidon_params **out) {
  if (!constants || !out || width < 3) return BP_ERR_INVALID_ARGUMENT;
  size_t total = (size_t)(fb + fe + pr) * width;
  if (nconst < (size_t)width * width + total) return BP_ERR_INVALID_ARGUMENT;
  bp_poseidon_params *p = new bp_poseidon_params();
  p->width = width; p->full_rounds_beginning = fb; p->full_rounds_end = fe; p->partial_rounds = pr;
  p->mds.assign(width, std::vector<scm>(width));
  for (uint32_t i = 0; i < width; i++) for (uint32_t j = 0; j < width; j++) p->mds[i][j] = load_scalar(constants + 32 * ((size_t)i * width + j));
  p->round_keys.resize(total);
  for (size_t i = 0; i < total; i++) p->round_keys[i] = load_scalar(constants + 32 * ((size_t)width * width + i));
  *out = p;
  return BP_OK;
}
void bp_poseidon_params_free(bp_poseidon_params *p) { delete p; }
int32_t bp_poseidon_hash_2(const bp_poseidon_params *p, const uint8_t xl[32], const uint8_t xr[32], int32_t sbox, uint8_t out[32]) {
  if (!p || !xl || !xr || !out || p->width != 6) return BP_ERR_INVALID_ARGUMENT;
  sc_tobytes(out, poseidon_hash_2(*p, load_scalar(xl), load_scalar(xr), sbox));
  return BP_OK;
}
int32_t bp_gadget_allocate_statics(bp_cs *cs, uint32_t num_statics, bp_var *out_vars) {  // gadget_poseidon.rs:554-608
  if (!cs || !out_vars || num_statics < 2) return BP_ERR_INVALID_ARGUMENT;
  uint8_t zero[32] = {0}, pad[32] = {0}; pad[0] = 101;
  for (uint32_t i = 0; i < num_statics; i++) {
    const uint8_t *val = i == 1 ? pad : zero;
    int rc;
    uint8_t V[32];
    if (cs->is_prover) rc = bp_prover_commit(cs, val, zero, V, &out_vars[i]);
    else { rc = engine_commit(cs->gens->g, 1, val, zero, V); if (!rc) rc = bp_verifier_commit(cs, V, &out_vars[i]); }
    if (rc) return rc;
    // a compiled circuit remembers these slots: its batch verifiers reject a proof whose caller-supplied V differs here
    cs->fixed_idx.push_back(out_vars[i].index);
    std::array<uint8_t, 32> fv; memcpy(fv.data(), V, 32); cs->fixed_V.push_back(fv);
  }
  return BP_OK;
}
int32_t bp_gadget_poseidon_hash_2(bp_cs *cs, const bp_poseidon_params *p, bp_var xl, bp_var xr, const bp_var *statics, uint32_t ns,
                                  int32_t sbox, const uint8_t expected[32]) {  // gadget_poseidon.rs:470-486
  if (!cs || !p || !statics || !expected || !var_ok(cs, xl) || !var_ok(cs, xr)) return BP_ERR_INVALID_ARGUMENT;
  std::vector<LC> st; for (uint32_t i = 0; i < ns; i++) { if (!var_ok(cs, statics[i])) return BP_ERR_INVALID_ARGUMENT; st.push_back(LC(statics[i])); }
  LC h; int rc = poseidon_hash_2_constraints(*cs, *p, LC(xl), LC(xr), st, sbox, h);
  if (rc) return rc;
  cs->constrain(h - LC::constant(load_scalar(expected)));
  return BP_OK;
}
int32_t bp_gadget_poseidon_hash_2_public(bp_cs *cs, const bp_poseidon_params *p, bp_var xl, bp_var xr, const bp_var *statics, uint32_t ns,
                                         int32_t sbox, bp_var expected) {
  if (!cs || !p || !statics || !var_ok(cs, expected) || !var_ok(cs, xl) || !var_ok(cs, xr)) return BP_ERR_INVALID_ARGUMENT;
  std::vector<LC> st; for (uint32_t i = 0; i < ns; i++) { if (!var_ok(cs, statics[i])) return BP_ERR_INVALID_ARGUMENT; st.push_back(LC(statics[i])); }
  LC h; int rc = poseidon_hash_2_constraints(*cs, *p, LC(xl), LC(xr), st, sbox, h);
  if (rc) return rc;
  cs->constrain(h - LC(expected));
  return BP_OK;
}
int32_t bp_gadget_vsmt2_verif(bp_cs *cs, const bp_poseidon_params *p, uint32_t depth, const uint8_t root[32], bp_var leaf,
                              const bp_var *bits, const bp_var *nodes, const bp_var *statics, uint32_t ns) {
  if (!cs || !p || !root || !bits || !nodes || !statics || !var_ok(cs, leaf)) return BP_ERR_INVALID_ARGUMENT;
  for (uint32_t i = 0; i < depth; i++) if (!var_ok(cs, bits[i]) || !var_ok(cs, nodes[i])) return BP_ERR_INVALID_ARGUMENT;
  for (uint32_t i = 0; i < ns; i++) if (!var_ok(cs, statics[i])) return BP_ERR_INVALID_ARGUMENT;
  return vsmt2_verif_gadget(*cs, *p, depth, LC::constant(load_scalar(root)), leaf, bits, nodes, statics, ns);
}
int32_t bp_gadget_vsmt2_verif_public(bp_cs *cs, const bp_poseidon_params *p, uint32_t depth, bp_var root, bp_var leaf, const bp_var *bits,
                                     const bp_var *nodes, const bp_var *statics, uint32_t ns) {
  if (!cs || !p || !bits || !nodes || !statics || !var_ok(cs, leaf) || !var_ok(cs, root)) return BP_ERR_INVALID_ARGUMENT;
  for (uint32_t i = 0; i < depth; i++) if (!var_ok(cs, bits[i]) || !var_ok(cs, nodes[i])) return BP_ERR_INVALID_ARGUMENT;
  for (uint32_t i = 0; i < ns; i++) if (!var_ok(cs, statics[i])) return BP_ERR_INVALID_ARGUMENT;
  return vsmt2_verif_gadget(*cs, *p, depth, LC(root), leaf, bits, nodes, statics, ns);
}
int32_t bp_poseidon_hash_4(const bp_poseidon_params *p, const uint8_t in[4][32], int32_t sbox, uint8_t out[32]) {
  if (!p || !in || !out || p->width != 6) return BP_ERR_INVALID_ARGUMENT;
  scm x[4]; for (int i = 0; i < 4; i++) x[i] = load_scalar(in[i]);
  sc_tobytes(out, poseidon_hash_4(*p, x, sbox));
  return BP_OK;
}
static int hash4_common(bp_cs *cs, const bp_poseidon_params *p, const bp_var in[4], const bp_var *statics, uint32_t ns, int32_t sbox, LC &h) {
  if (!cs || !p || !in || !statics) return BP_ERR_INVALID_ARGUMENT;
  std::vector<LC> st; for (uint32_t i = 0; i < ns; i++) { if (!var_ok(cs, statics[i])) return BP_ERR_INVALID_ARGUMENT; st.push_back(LC(statics[i])); }
  LC x[4]; for (int i = 0; i < 4; i++) { if (!var_ok(cs, in[i])) return BP_ERR_INVALID_ARGUMENT; x[i] = LC(in[i]); }
  return poseidon_hash_4_constraints(*cs, *p, x, st, sbox, h);
}
int32_t bp_gadget_poseidon_hash_4(bp_cs *cs, const bp_poseidon_params *p, const bp_var in[4], const bp_var *statics, uint32_t ns, int32_t sbox,
                                  const uint8_t expected[32]) {  // gadget_poseidon.rs:532-551
  if (!expected) return BP_ERR_INVALID_ARGUMENT;
  LC h; int rc = hash4_common(cs, p, in, statics, ns, sbox, h); if (rc) return rc;
  cs->constrain(h - LC::constant(load_scalar(expected)));
  return BP_OK;
}
int32_t bp_gadget_poseidon_hash_4_public(bp_cs *cs, const bp_poseidon_params *p, const bp_var in[4], const bp_var *statics, uint32_t ns, int32_t sbox,
                                         bp_var expected) {
  if (!cs || !var_ok(cs, expected)) return BP_ERR_INVALID_ARGUMENT;
  LC h; int rc = hash4_common(cs, p, in, statics, ns, sbox, h); if (rc) return rc;
  cs->constrain(h - LC(expected));
  return BP_OK;
}
static int vsmt4_args_ok(bp_cs *cs, const bp_poseidon_params *p, uint32_t levels, bp_var leaf, bp_var leaf_index, const bp_var *nodes, const bp_var *statics, uint32_t ns) {
  if (!cs || !p || !nodes || !statics || !var_ok(cs, leaf) || !var_ok(cs, leaf_index) || levels == 0 || levels > 126) return 0;
  for (uint32_t i = 0; i < 3 * levels; i++) if (!var_ok(cs, nodes[i])) return 0;
  for (uint32_t i = 0; i < ns; i++) if (!var_ok(cs, statics[i])) return 0;
  return 1;
}
int32_t bp_gadget_vsmt4_verif(bp_cs *cs, const bp_poseidon_params *p, uint32_t levels, const uint8_t root[32], bp_var leaf, bp_var leaf_index,
                              const uint8_t *index_digits, const bp_var *nodes, const bp_var *statics, uint32_t ns) {
  if (!root || !vsmt4_args_ok(cs, p, levels, leaf, leaf_index, nodes, statics, ns)) return BP_ERR_INVALID_ARGUMENT;
  return vsmt4_verif_gadget(*cs, *p, levels, LC::constant(load_scalar(root)), leaf, leaf_index, index_digits, nodes, statics, ns);
}
int32_t bp_gadget_vsmt4_verif_public(bp_cs *cs, const bp_poseidon_params *p, uint32_t levels, bp_var root, bp_var leaf, bp_var leaf_index,
                                     const uint8_t *index_digits, const bp_var *nodes, const bp_var *statics, uint32_t ns) {
  if (!vsmt4_args_ok(cs, p, levels, leaf, leaf_index, nodes, statics, ns) || !var_ok(cs, root)) return BP_ERR_INVALID_ARGUMENT;
  return vsmt4_verif_gadget(*cs, *p, levels, LC(root), leaf, leaf_index, index_digits, nodes, statics, ns);
}
int32_t bp_gadget_mimc(bp_cs *cs, bp_var left, bp_var right, uint32_t rounds, const uint8_t *constants, const uint8_t image[32]) {
  if (!cs || !constants || !image || !var_ok(cs, left) || !var_ok(cs, right)) return BP_ERR_INVALID_ARGUMENT;
  std::vector<scm> k(rounds); for (uint32_t i = 0; i < rounds; i++) k[i] = load_scalar(constants + 32 * (size_t)i);
  return mimc_gadget(*cs, left, right, rounds, k.data(), LC::constant(load_scalar(image)));
}
int32_t bp_gadget_mimc_public(bp_cs *cs, bp_var left, bp_var right, uint32_t rounds, const uint8_t *constants, bp_var image) {
  if (!cs || !constants || !var_ok(cs, image) || !var_ok(cs, left) || !var_ok(cs, right)) return BP_ERR_INVALID_ARGUMENT;
  std::vector<scm> k(rounds); for (uint32_t i = 0; i < rounds; i++) k[i] = load_scalar(constants + 32 * (size_t)i);
  return mimc_gadget(*cs, left, right, rounds, k.data(), LC(image));
}
int32_t bp_mimc(const uint8_t xl[32], const uint8_t xr[32], uint32_t rounds, const uint8_t *constants, uint8_t out[32]) {
  if (!xl || !xr || !constants || !out) return BP_ERR_INVALID_ARGUMENT;
  std::vector<scm> k(rounds); for (uint32_t i = 0; i < rounds; i++) k[i] = load_scalar(constants + 32 * (size_t)i);
  sc_tobytes(out, mimc_native(load_scalar(xl), load_scalar(xr), rounds, k.data()));
  return BP_OK;
}
int32_t bp_gadget_bound_check(bp_cs *cs, bp_var v, bp_var a, bp_var b, int32_t has, uint64_t vv, uint64_t av, uint64_t bv, uint64_t max,
                              uint64_t min, uint32_t bit_size) {
  if (!cs || !var_ok(cs, v) || !var_ok(cs, a) || !var_ok(cs, b)) return BP_ERR_INVALID_ARGUMENT;
  if (cs->is_prover && !has) return BP_ERR_MISSING_ASSIGNMENT;
  return bound_check_gadget(*cs, v, a, b, cs->is_prover, vv, av, bv, max, min, bit_size);
}

// ------------------------------------------------------------------------------------------------ tier 2
int32_t bp_circuit_compile(const bp_cs *cs, bp_circuit **out) {
  if (!cs || !out) return BP_ERR_INVALID_ARGUMENT;
  BpCircuit *c = nullptr;
  int rc = compile_cs(cs, true, &c);
  if (rc) return rc;
  *out = new bp_circuit{c};
  return BP_OK;
}
int32_t bp_circuit_from_arrays(uint32_t n, uint32_t m, uint32_t q, const uint32_t *cons_ptr, const uint8_t *kind, const uint32_t *idx,
                               const uint8_t *coeff, bp_circuit **out) {
  if (!cons_ptr || !out || cons_ptr[0] != 0) return BP_ERR_INVALID_ARGUMENT;
  for (uint32_t k = 0; k < q; k++) if (cons_ptr[k + 1] < cons_ptr[k]) return BP_ERR_INVALID_ARGUMENT;
  uint32_t nnz = cons_ptr[q];
  if (nnz && (!kind || !idx || !coeff)) return BP_ERR_INVALID_ARGUMENT;
  std::vector<scm> co(nnz ? nnz : 1);
  for (uint32_t t = 0; t < nnz; t++) co[t] = load_scalar(coeff + 32 * (size_t)t);
  BpCircuit *c = nullptr;
  int rc = circuit_create(n, m, 0, q, cons_ptr, kind, idx, co.data(), nullptr, 0, 0, nullptr, nullptr, nullptr, nullptr, nullptr, &c);
  if (rc) return rc;
  *out = new bp_circuit{c};
  return BP_OK;
}
void bp_circuit_free(bp_circuit *c) { if (c) { c->stage[0].release(); c->stage[1].release(); circuit_free(c->c); delete c; } }
int32_t bp_circuit_release_workspace(bp_circuit *c) {
  if (!c) return BP_ERR_INVALID_ARGUMENT;
  int rc = circuit_release_workspace(c->c);
  if (rc) return rc;
  c->stage[0].release(); c->stage[1].release();
  return BP_OK;
}
uint32_t bp_circuit_num_multipliers(const bp_circuit *c) { return c ? c->c->n : 0; }
uint32_t bp_circuit_num_constraints(const bp_circuit *c) { return c ? c->c->q : 0; }
uint32_t bp_circuit_num_commitments(const bp_circuit *c) { return c ? c->c->m : 0; }
uint32_t bp_circuit_num_aux(const bp_circuit *c) { return c ? c->c->naux : 0; }
uint32_t bp_circuit_num_public(const bp_circuit *c) { return c ? c->c->npub : 0; }
int32_t bp_circuit_has_witness_program(const bp_circuit *c) { return c ? c->c->has_tape : 0; }
size_t bp_circuit_proof_len(const bp_circuit *c) { return c ? circuit_proof_len(c->c) : 0; }

// proofs per device chunk: bounded by a workspace budget (bytes) and BP_B200_CHUNK
static uint32_t chunk_size(const BpCircuit *c, uint32_t B) {
  double per = engine_workspace_bytes_per_proof(c) * 1.05;
  // budget for the chunk in flight + the phase-A outputs of every chunk; the generator tables (26 GB at capacity 32768)
  // and the caller's buffers share the 180 GB with it
  double budget = 72e9;
  uint32_t ch = (uint32_t)std::max(1.0, std::min((double)B, budget / per));
  if (ch > 32768) ch = 32768;
  uint32_t nchunks = (B + ch - 1) / ch;
  ch = (B + nchunks - 1) / nchunks;  // equal-sized chunks
  const char *e = getenv("BP_B200_CHUNK");
  if (e && atoi(e) > 0) ch = std::min<uint32_t>(B, (uint32_t)atoi(e));
  return ch;
}

static int prove_args_device(const bp_circuit *c, uint32_t B, const uint8_t *label, size_t label_len, const uint8_t *d_v, const uint8_t *d_vb,
                             const uint8_t *d_entropy, const uint8_t *d_aux, const uint8_t *d_pub, const uint8_t *d_aL, const uint8_t *d_aR,
                             const uint8_t *d_aO, uint8_t *d_V, uint8_t *d_proofs, int32_t *d_status, ProveArgs &a) {
  if (!c || !d_v || !d_vb || !d_entropy || !d_V || !d_proofs || !d_status || (label_len && !label)) return BP_ERR_INVALID_ARGUMENT;
  BpCircuit *cc = c->c;
  if ((d_aL || d_aR || d_aO) && !(d_aL && d_aR && d_aO)) return BP_ERR_INVALID_ARGUMENT;
  if (!d_aL && !cc->has_tape) return BP_ERR_MISSING_ASSIGNMENT;
  if (!d_aL && cc->naux && !d_aux) return BP_ERR_MISSING_ASSIGNMENT;
  a = ProveArgs{}; a.B = (int)B; a.label = label; a.label_len = (int)label_len;
  a.v = d_v; a.vbl = d_vb; a.entropy = d_entropy; a.aux = d_aux; a.pub = d_pub; a.aL = d_aL; a.aR = d_aR; a.aO = d_aO;
  a.V_out = d_V; a.proofs = d_proofs; a.status = d_status;
  return BP_OK;
}
static dev_stream to_stream(void *stream) {
#ifndef BP_HOST_EMUL
  return (dev_stream)stream;
#else
  (void)stream; return 0;
#endif
}

int32_t bp_prove_batch_device(const bp_gens *g, bp_circuit *c, uint32_t B, const uint8_t *label, size_t label_len, const uint8_t *d_v,
                              const uint8_t *d_vb, const uint8_t *d_entropy, const uint8_t *d_aux, const uint8_t *d_pub, const uint8_t *d_aL,
                              const uint8_t *d_aR, const uint8_t *d_aO, uint8_t *d_V, uint8_t *d_proofs, int32_t *d_status, void *stream) {
  if (!g) return BP_ERR_INVALID_ARGUMENT;
  ProveArgs a;
  int rc = prove_args_device(c, B, label, label_len, d_v, d_vb, d_entropy, d_aux, d_pub, d_aL, d_aR, d_aO, d_V, d_proofs, d_status, a);
  if (rc) return rc;
  return engine_prove(g->g, c->c, a, (int)chunk_size(c->c, B), to_stream(stream));
}

// streaming form: see include/bp_b200.h
int32_t bp_prove_stream_begin(const bp_gens *g, bp_circuit *c, int32_t slot, uint32_t B, const uint8_t *label, size_t label_len, const uint8_t *d_v,
                              const uint8_t *d_vb, const uint8_t *d_entropy, const uint8_t *d_aux, const uint8_t *d_pub, const uint8_t *d_aL,
                              const uint8_t *d_aR, const uint8_t *d_aO, uint8_t *d_V, uint8_t *d_proofs, int32_t *d_status, void *stream) {
  if (!g || B == 0) return BP_ERR_INVALID_ARGUMENT;
  ProveArgs a;
  int rc = prove_args_device(c, B, label, label_len, d_v, d_vb, d_entropy, d_aux, d_pub, d_aL, d_aR, d_aO, d_V, d_proofs, d_status, a);
  if (rc) return rc;
  return engine_prove_begin(g->g, c->c, slot, a, (int)chunk_size(c->c, B), to_stream(stream));
}
int32_t bp_prove_stream_finish(const bp_gens *g, bp_circuit *c, int32_t slot, void *stream) {
  if (!g || !c) return BP_ERR_INVALID_ARGUMENT;
  return engine_prove_finish(g->g, c->c, slot, to_stream(stream));
}

// host buffers (pinned for the copies to be asynchronous); witness from the circuit's program only
int32_t bp_prove_stream_begin_host(const bp_gens *g, bp_circuit *c, int32_t slot, uint32_t B, const uint8_t *label, size_t label_len, const uint8_t *v,
                                   const uint8_t *vb, const uint8_t *entropy, const uint8_t *aux, const uint8_t *pub, void *stream) {
  if (!g || !c || !v || !vb || !entropy || B == 0 || slot < 0 || slot > 1 || (label_len && !label)) return BP_ERR_INVALID_ARGUMENT;
  BpCircuit *cc = c->c;
  if (!cc->has_tape || (cc->naux && !aux)) return BP_ERR_MISSING_ASSIGNMENT;
  HostStage &st = c->stage[slot];
  if (st.active) return BP_ERR_INVALID_ARGUMENT;
  const size_t m = cc->m, plen = circuit_proof_len(cc), na = cc->naux, np_ = cc->npub;
  dev_stream s = to_stream(stream);
  if (st.cap < B) {
    if (dev_sync(s)) return BP_ERR_CUDA;
    st.release();
    int bad = 0;
    bad |= dev_malloc((void **)&st.v, B * m * 32); bad |= dev_malloc((void **)&st.vb, B * m * 32); bad |= dev_malloc((void **)&st.e, (size_t)B * 32);
    bad |= dev_malloc((void **)&st.aux, B * na * 32); bad |= dev_malloc((void **)&st.pub, B * np_ * 32); bad |= dev_malloc((void **)&st.V, B * m * 32);
    bad |= dev_malloc((void **)&st.P, B * plen); bad |= dev_malloc((void **)&st.S, B * sizeof(int32_t));
    if (bad) { st.release(); return BP_ERR_OOM; }
    st.cap = B;
  }
  int bad = 0;
  bad |= dev_h2d(st.v, v, B * m * 32, s); bad |= dev_h2d(st.vb, vb, B * m * 32, s); bad |= dev_h2d(st.e, entropy, (size_t)B * 32, s);
  if (na) bad |= dev_h2d(st.aux, aux, B * na * 32, s);
  if (np_ && pub) bad |= dev_h2d(st.pub, pub, B * np_ * 32, s);
  if (bad) return BP_ERR_CUDA;
  int rc = bp_prove_stream_begin(g, c, slot, B, label, label_len, st.v, st.vb, st.e, na ? st.aux : nullptr, (np_ && pub) ? st.pub : nullptr, nullptr, nullptr,
                                 nullptr, st.V, st.P, st.S, stream);
  if (rc) return rc;
  st.B = B; st.active = true;
  return BP_OK;
}
// runs the rest of the slot's batch, copies commitments / proofs / status to the host buffers and waits for them
int32_t bp_prove_stream_finish_host(const bp_gens *g, bp_circuit *c, int32_t slot, uint8_t *V_out, uint8_t *proofs, int32_t *status, void *stream) {
  if (!g || !c || !V_out || !proofs || !status || slot < 0 || slot > 1 || !c->stage[slot].active) return BP_ERR_INVALID_ARGUMENT;
  HostStage &st = c->stage[slot];
  st.active = false;
  int rc = bp_prove_stream_finish(g, c, slot, stream);
  if (rc) return rc;
  dev_stream s = to_stream(stream);
  const size_t B = st.B, m = c->c->m, plen = circuit_proof_len(c->c);
  int bad = 0;
  bad |= dev_d2h(V_out, st.V, B * m * 32, s); bad |= dev_d2h(proofs, st.P, B * plen, s); bad |= dev_d2h(status, st.S, B * sizeof(int32_t), s);
  bad |= dev_sync(s);
  return bad ? BP_ERR_CUDA : BP_OK;
}

int32_t bp_prove_batch(const bp_gens *g, bp_circuit *c, uint32_t B, const uint8_t *label, size_t label_len, const uint8_t *v, const uint8_t *vb,
                       const uint8_t *entropy, const uint8_t *aux, const uint8_t *pub, const uint8_t *aL, const uint8_t *aR, const uint8_t *aO,
                       uint8_t *V_out, uint8_t *proofs, int32_t *status) {
  if (!g || !c || !v || !vb || !entropy || !V_out || !proofs || !status) return BP_ERR_INVALID_ARGUMENT;
  BpCircuit *cc = c->c;
  if ((aL || aR || aO) && !(aL && aR && aO)) return BP_ERR_INVALID_ARGUMENT;
  if (!aL && !cc->has_tape) return BP_ERR_MISSING_ASSIGNMENT;
  if (!aL && cc->naux && !aux) return BP_ERR_MISSING_ASSIGNMENT;
  if (B == 0) return BP_OK;
  const size_t m = cc->m, n = cc->n, plen = circuit_proof_len(cc), na = cc->naux, np_ = cc->npub;
  DevBuf dv, dvb, de, da, dpb, daL, daR, daO, dV, dP, dS;
  if (dpb.alloc(B * np_ * 32 + 32)) return BP_ERR_OOM;
  if (dv.alloc(B * m * 32 + 32) || dvb.alloc(B * m * 32 + 32) || de.alloc((size_t)B * 32) || da.alloc(B * na * 32 + 32) || dV.alloc(B * m * 32 + 32) ||
      dP.alloc(B * plen) || dS.alloc(B * sizeof(int))) return BP_ERR_OOM;
  if (aL && (daL.alloc(B * n * 32 + 32) || daR.alloc(B * n * 32 + 32) || daO.alloc(B * n * 32 + 32))) return BP_ERR_OOM;
  dev_stream s = 0;
  int bad = 0;
  bad |= dev_h2d(dv.p, v, B * m * 32, s); bad |= dev_h2d(dvb.p, vb, B * m * 32, s); bad |= dev_h2d(de.p, entropy, (size_t)B * 32, s);
  if (na && aux) bad |= dev_h2d(da.p, aux, B * na * 32, s);
  if (np_ && pub) bad |= dev_h2d(dpb.p, pub, B * np_ * 32, s);
  if (aL) { bad |= dev_h2d(daL.p, aL, B * n * 32, s); bad |= dev_h2d(daR.p, aR, B * n * 32, s); bad |= dev_h2d(daO.p, aO, B * n * 32, s); }
  if (bad) return BP_ERR_CUDA;
  int rc = bp_prove_batch_device(g, c, B, label, label_len, dv.p, dvb.p, de.p, na ? da.p : nullptr, (np_ && pub) ? dpb.p : nullptr, daL.p, daR.p, daO.p, dV.p, dP.p, (int32_t *)dS.p, nullptr);
  if (rc) return rc;
  bad |= dev_d2h(V_out, dV.p, B * m * 32, s); bad |= dev_d2h(proofs, dP.p, B * plen, s); bad |= dev_d2h(status, dS.p, B * sizeof(int), s);
  bad |= dev_sync(s);
  return bad ? BP_ERR_CUDA : BP_OK;
}

int32_t bp_verify_batch_device(const bp_gens *g, bp_circuit *c, uint32_t B, const uint8_t *label, size_t label_len, const uint8_t *d_V,
                               const uint8_t *d_proofs, const uint8_t *d_entropy, const uint8_t *d_pub, int32_t *d_status, void *stream) {
  if (!g || !c || !d_V || !d_proofs || !d_entropy || !d_status || (label_len && !label)) return BP_ERR_INVALID_ARGUMENT;
  BpCircuit *cc = c->c;
#ifndef BP_HOST_EMUL
  dev_stream s = (dev_stream)stream;
#else
  dev_stream s = 0; (void)stream;
#endif
  const size_t m = cc->m, plen = circuit_proof_len(cc);
  uint32_t ch = chunk_size(cc, B);
  for (uint32_t p0 = 0; p0 < B; p0 += ch) {
    uint32_t bc = std::min(ch, B - p0);
    VerifyArgs a{}; a.B = (int)bc; a.label = label; a.label_len = (int)label_len;
    a.V = d_V + (size_t)p0 * m * 32; a.proofs = d_proofs + (size_t)p0 * plen; a.entropy = d_entropy + (size_t)p0 * 32; a.status = d_status + p0;
    a.pub = d_pub ? d_pub + (size_t)p0 * cc->npub * 32 : nullptr;
    int rc = engine_verify(g->g, cc, a, s);
    if (rc) return rc;
  }
  return BP_OK;
}
int32_t bp_verify_batch_combined_device(const bp_gens *g, bp_circuit *c, uint32_t B, const uint8_t *label, size_t label_len, const uint8_t *d_V,
                                        const uint8_t *d_proofs, const uint8_t *d_entropy, const uint8_t *d_pub, int32_t *d_status, int32_t *d_combined,
                                        void *stream) {
  if (!g || !c || !d_V || !d_proofs || !d_entropy || !d_status || !d_combined || (label_len && !label)) return BP_ERR_INVALID_ARGUMENT;
  BpCircuit *cc = c->c;
#ifndef BP_HOST_EMUL
  dev_stream s = (dev_stream)stream;
#else
  dev_stream s = 0; (void)stream;
#endif
  if (B == 0) return BP_OK;
  if (chunk_size(cc, B) < B) return BP_ERR_INVALID_ARGUMENT;  // one combination per call: the batch must fit one device chunk
  VerifyArgs a{}; a.B = (int)B; a.label = label; a.label_len = (int)label_len;
  a.V = d_V; a.proofs = d_proofs; a.entropy = d_entropy; a.status = d_status; a.pub = d_pub;
  return engine_verify_combined(g->g, cc, a, d_combined, s);
}
int32_t bp_verify_batch_combined(const bp_gens *g, bp_circuit *c, uint32_t B, const uint8_t *label, size_t label_len, const uint8_t *V,
                                 const uint8_t *proofs, const uint8_t *entropy, const uint8_t *pub, int32_t *status, int32_t *combined) {
  if (!g || !c || !V || !proofs || !entropy || !status || !combined) return BP_ERR_INVALID_ARGUMENT;
  *combined = BP_OK;
  if (B == 0) return BP_OK;
  BpCircuit *cc = c->c;
  const size_t m = cc->m, plen = circuit_proof_len(cc);
  DevBuf dV, dP, de, dS, dpb, dC;
  const size_t np_ = cc->npub;
  if (np_ && !pub) return BP_ERR_MISSING_ASSIGNMENT;
  if (dV.alloc(B * m * 32 + 32) || dP.alloc(B * plen) || de.alloc((size_t)B * 32) || dS.alloc(B * sizeof(int)) || dpb.alloc(B * np_ * 32 + 32) || dC.alloc(sizeof(int))) return BP_ERR_OOM;
  dev_stream s = 0;
  if (np_ && dev_h2d(dpb.p, pub, B * np_ * 32, s)) return BP_ERR_CUDA;
  int bad = dev_h2d(dV.p, V, B * m * 32, s) | dev_h2d(dP.p, proofs, B * plen, s) | dev_h2d(de.p, entropy, (size_t)B * 32, s);
  if (bad) return BP_ERR_CUDA;
  int rc = bp_verify_batch_combined_device(g, c, B, label, label_len, dV.p, dP.p, de.p, np_ ? dpb.p : nullptr, (int32_t *)dS.p, (int32_t *)dC.p, nullptr);
  if (rc) return rc;
  bad = dev_d2h(status, dS.p, B * sizeof(int), s) | dev_d2h(combined, dC.p, sizeof(int), s) | dev_sync(s);
  return bad ? BP_ERR_CUDA : BP_OK;
}

// ---- wire format (App. A.7) ----
int64_t bp_proof_to_wire(const uint8_t *proof, size_t proof_len, uint8_t *out, size_t out_cap) {
  if (!proof || !out) return -(int64_t)BP_ERR_INVALID_ARGUMENT;
  if (proof_len < 16 * 32 || proof_len % 32) return -(int64_t)BP_ERR_FORMAT;
  bool one_phase = true;
  for (int i = 96; i < 192; i++) one_phase &= proof[i] == 0;  // A_I2, A_O2, S2 all the identity encoding
  const size_t need = one_phase ? 1 + proof_len - 96 : 1 + proof_len;
  if (out_cap < need) return -(int64_t)BP_ERR_INVALID_ARGUMENT;
  out[0] = one_phase ? 0 : 1;
  memcpy(out + 1, proof, 96);
  if (one_phase) memcpy(out + 97, proof + 192, proof_len - 192); else memcpy(out + 97, proof + 96, proof_len - 96);
  return (int64_t)need;
}
int64_t bp_proof_from_wire(const uint8_t *wire, size_t wire_len, uint8_t *proof_out, size_t out_cap) {
  if (!wire || !proof_out) return -(int64_t)BP_ERR_INVALID_ARGUMENT;
  if (wire_len < 1 || wire[0] > 1) return -(int64_t)BP_ERR_FORMAT;
  const bool one_phase = wire[0] == 0;
  const size_t body = wire_len - 1;
  if (body % 32 || body < (size_t)(one_phase ? 13 : 16) * 32) return -(int64_t)BP_ERR_FORMAT;
  const size_t need = one_phase ? body + 96 : body;
  if ((need / 32 - 14) % 2) return -(int64_t)BP_ERR_FORMAT;  // 14 fixed elements + L/R pairs + a, b
  if (out_cap < need) return -(int64_t)BP_ERR_INVALID_ARGUMENT;
  memcpy(proof_out, wire + 1, 96);
  if (one_phase) { memset(proof_out + 96, 0, 96); memcpy(proof_out + 192, wire + 97, body - 96); } else memcpy(proof_out + 96, wire + 97, body - 96);
  return (int64_t)need;
}

int32_t bp_verify_batch(const bp_gens *g, bp_circuit *c, uint32_t B, const uint8_t *label, size_t label_len, const uint8_t *V, const uint8_t *proofs,
                        const uint8_t *entropy, const uint8_t *pub, int32_t *status) {
  if (!g || !c || !V || !proofs || !entropy || !status) return BP_ERR_INVALID_ARGUMENT;
  if (B == 0) return BP_OK;
  BpCircuit *cc = c->c;
  const size_t m = cc->m, plen = circuit_proof_len(cc);
  DevBuf dV, dP, de, dS, dpb;
  const size_t np_ = cc->npub;
  if (np_ && !pub) return BP_ERR_MISSING_ASSIGNMENT;
  if (dV.alloc(B * m * 32 + 32) || dP.alloc(B * plen) || de.alloc((size_t)B * 32) || dS.alloc(B * sizeof(int)) || dpb.alloc(B * np_ * 32 + 32)) return BP_ERR_OOM;
  dev_stream s = 0;
  if (np_ && dev_h2d(dpb.p, pub, B * np_ * 32, s)) return BP_ERR_CUDA;
  int bad = dev_h2d(dV.p, V, B * m * 32, s) | dev_h2d(dP.p, proofs, B * plen, s) | dev_h2d(de.p, entropy, (size_t)B * 32, s);
  if (bad) return BP_ERR_CUDA;
  int rc = bp_verify_batch_device(g, c, B, label, label_len, dV.p, dP.p, de.p, np_ ? dpb.p : nullptr, (int32_t *)dS.p, nullptr);
  if (rc) return rc;
  bad = dev_d2h(status, dS.p, B * sizeof(int), s) | dev_sync(s);
  return bad ? BP_ERR_CUDA : BP_OK;
}

int32_t bp_msm_gens_device(const bp_gens *g, uint32_t n, const uint8_t *d_scalars, uint8_t *d_out, void *stream) {
  if (!g || !d_scalars || !d_out) return BP_ERR_INVALID_ARGUMENT;
#ifndef BP_HOST_EMUL
  dev_stream s = (dev_stream)stream;
#else
  dev_stream s = 0; (void)stream;
#endif
  return engine_msm_gens(g->g, n, d_scalars, d_out, s);
}

// primitive self-test on the device (tests only); host buffers
int32_t bp_selftest_device(int32_t which, const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len) {
  int rc = bp_device_init();
  if (rc) return rc;
  DevBuf di, dout;
  if (di.alloc(in_len + 16) || dout.alloc(out_len + 16)) return BP_ERR_OOM;
  dev_stream s = 0;
  if (dev_h2d(di.p, in, in_len, s) || dev_memset(dout.p, 0, out_len, s)) return BP_ERR_CUDA;
  if (launch(1, s, KSelfTest{which, di.p, (int)in_len, dout.p, (int)out_len})) return BP_ERR_CUDA;
  if (dev_d2h(out, dout.p, out_len, s) || dev_sync(s)) return BP_ERR_CUDA;
  return BP_OK;
}

// runs the real KLoadScalars + KTsStart kernels on crafted inputs (m commitments): in = V[m][32] | vbl[m][32] | entropy[32] | label
// out = ts state (208 B) | rng state (208 B)
int32_t bp_selftest_tsstart(uint32_t m, const uint8_t *in, const uint8_t *label, size_t label_len, uint8_t *out) {
  int rc = bp_device_init();
  if (rc) return rc;
  DevBuf di, dts, drng, dvbl;
  if (di.alloc(64 * m + 32) || dts.alloc(sizeof(strobe128)) || drng.alloc(sizeof(strobe128)) || dvbl.alloc(32 * m + 32)) return BP_ERR_OOM;
  dev_stream s = 0;
  if (dev_h2d(di.p, in, 64 * m + 32, s)) return BP_ERR_CUDA;
  if (launch(m, s, KLoadScalars{di.p + 32 * m, (scm *)dvbl.p, (int)m, 1})) return BP_ERR_CUDA;
  strobe128 base; ts_init(base, label, (int)label_len);
  const uint8_t r1[7] = {'r', '1', 'c', 's', ' ', 'v', '1'};
  ts_append(base, "dom-sep", r1, 7);
  if (launch(1, s, KTsStart{base, di.p, (int)m, 1, (const scm *)dvbl.p, di.p + 64 * m, (strobe128 *)dts.p, (strobe128 *)drng.p, 1})) return BP_ERR_CUDA;
  if (dev_d2h(out, dts.p, sizeof(strobe128), s) || dev_d2h(out + sizeof(strobe128), drng.p, sizeof(strobe128), s) || dev_sync(s)) return BP_ERR_CUDA;
  return BP_OK;
}

}  // extern "C"
