// Host-side recorder methods and the C++ restatement of the reference's gadget layer.
// Each function cites the reference file:line it mirrors (paths relative to /root/reference/src).
#include "recorder.h"
#include <unordered_map>

LC LC::simplified() const {  // gadget_poseidon.rs:99-112
  LC out;
  std::unordered_map<uint64_t, size_t> pos;
  pos.reserve(terms.size() * 2);
  for (const Term &t : terms) {
    uint64_t key = ((uint64_t)t.var.kind << 32) | t.var.index;
    auto it = pos.find(key);
    if (it == pos.end()) { pos.emplace(key, out.terms.size()); out.terms.push_back(t); }
    else out.terms[it->second].coeff = sc_add(out.terms[it->second].coeff, t.coeff);
  }
  return out;
}

bool bp_cs::eval(const LC &lc, scm &out) const {
  if (!is_prover) return false;
  scm acc = sc_zero();
  for (const Term &t : lc.terms) {
    scm val;
    switch (t.var.kind) {
      case BP_VAR_COMMITTED: val = v[t.var.index]; break;
      case BP_VAR_MULT_LEFT: val = aL[t.var.index]; break;
      case BP_VAR_MULT_RIGHT: val = aR[t.var.index]; break;
      case BP_VAR_MULT_OUT: val = aO[t.var.index]; break;
      case BP_VAR_PUBLIC: val = pub[t.var.index]; break;
      default: val = sc_one(); break;
    }
    acc = sc_add(acc, sc_mul(t.coeff, val));
  }
  out = acc;
  return true;
}

// cs.multiply: allocates one multiplier, evaluates both sides on the prover, adds `left - l = 0`, `right - r = 0`
void bp_cs::multiply(const LC &l, const LC &r, bp_var out[3]) {
  uint32_t i = num_mult++;
  if (is_prover) { scm lv, rv; eval(l, lv); eval(r, rv); aL.push_back(lv); aR.push_back(rv); aO.push_back(sc_mul(lv, rv)); }
  TapeOp op{}; op.opL = W_LC; op.opR = W_LC; op.argL = add_wlc(l); op.argR = add_wlc(r);
  tape.push_back(op);
  out[0] = bp_var{BP_VAR_MULT_LEFT, i}; out[1] = bp_var{BP_VAR_MULT_RIGHT, i}; out[2] = bp_var{BP_VAR_MULT_OUT, i};
  LC lc = l; lc -= LC(out[0]); constrain(lc);
  LC rc = r; rc -= LC(out[1]); constrain(rc);
}

int bp_cs::allocate_multiplier(const scm *l, const scm *r, bp_var out[3]) {
  if (is_prover && (!l || !r)) return BP_ERR_MISSING_ASSIGNMENT;
  uint32_t i = num_mult++;
  if (is_prover) { aL.push_back(*l); aR.push_back(*r); aO.push_back(sc_mul(*l, *r)); aux.push_back(*l); aux.push_back(*r); }
  TapeOp op{}; op.opL = W_AUX; op.opR = W_AUX; op.argL = naux; op.argR = naux + 1; naux += 2;
  tape.push_back(op);
  out[0] = bp_var{BP_VAR_MULT_LEFT, i}; out[1] = bp_var{BP_VAR_MULT_RIGHT, i}; out[2] = bp_var{BP_VAR_MULT_OUT, i};
  return BP_OK;
}

int bp_cs::allocate_single(int how, const scm *value, const LC *lc, bp_var *var, bp_var *out_var, int *has_out) {
  scm val = sc_zero();
  if (is_prover) {
    if (how == 0) { if (!value) return BP_ERR_MISSING_ASSIGNMENT; val = *value; }
    else if (how == 1) eval(*lc, val);
    else { if (pending < 0) return BP_ERR_INVALID_ARGUMENT; val = sc_invert(aL[pending]); }
  }
  if (pending < 0) {
    if (how == 2) return BP_ERR_INVALID_ARGUMENT;
    uint32_t i = num_mult++;
    pending = i;
    if (is_prover) { aL.push_back(val); aR.push_back(sc_zero()); aO.push_back(sc_zero()); }
    TapeOp op{};
    if (how == 1) { op.opL = W_LC; op.argL = add_wlc(*lc); } else { op.opL = W_AUX; op.argL = naux++; if (is_prover) aux.push_back(val); }
    op.opR = W_AUX; op.argR = 0;  // completed by the second call
    tape.push_back(op);
    *var = bp_var{BP_VAR_MULT_LEFT, i};
    if (has_out) *has_out = 0;
  } else {
    uint32_t i = (uint32_t)pending;
    pending = -1;
    if (is_prover) { aR[i] = val; aO[i] = sc_mul(aL[i], val); }
    TapeOp &op = tape[i];
    if (how == 2) { op.opR = W_INV_L; op.argR = 0; }
    else if (how == 1) { op.opR = W_LC; op.argR = add_wlc(*lc); }
    else { op.opR = W_AUX; op.argR = naux++; if (is_prover) aux.push_back(val); }
    *var = bp_var{BP_VAR_MULT_RIGHT, i};
    if (out_var) *out_var = bp_var{BP_VAR_MULT_OUT, i};
    if (has_out) *has_out = 1;
  }
  return BP_OK;
}

// ------------------------------------------------------------------------------------------------ r1cs_utils / zero_nonzero
static void constrain_lc_with_scalar(bp_cs &cs, const LC &lc, const scm &s) {  // r1cs_utils.rs:51-53
  cs.constrain(lc - LC::constant(s));
}
static void is_nonzero_gadget(bp_cs &cs, bp_var x, bp_var x_inv) {  // gadget_zero_nonzero.rs:46-66
  LC x_lc(x), y_lc = LC::constant(sc_one());
  LC one_minus_y = LC(var_one()) - y_lc;
  bp_var o[3];
  cs.multiply(x_lc, one_minus_y, o);
  cs.constrain(LC(o[2]));
  cs.multiply(x_lc, LC(x_inv), o);
  cs.constrain(LC(o[2]) - y_lc);
}

// ------------------------------------------------------------------------------------------------ Poseidon
static scm apply_sbox(const scm &x, int sbox) {  // gadget_poseidon.rs:120-125
  return sbox == BP_SBOX_CUBE ? sc_mul(sc_sqr(x), x) : sc_invert(x);
}
void poseidon_permutation(const bp_poseidon_params &p, std::vector<scm> &st, int sbox) {  // gadget_poseidon.rs:189-280
  const uint32_t w = p.width, total = p.full_rounds_beginning + p.partial_rounds + p.full_rounds_end;
  size_t off = 0;
  for (uint32_t rnd = 0; rnd < total; rnd++) {
    bool full = rnd < p.full_rounds_beginning || rnd >= p.full_rounds_beginning + p.partial_rounds;
    for (uint32_t i = 0; i < w; i++) {
      st[i] = sc_add(st[i], p.round_keys[off++]);
      if (full || i == w - 1) st[i] = apply_sbox(st[i], sbox);
    }
    std::vector<scm> nx(w, sc_zero());
    for (uint32_t i = 0; i < w; i++)
      for (uint32_t j = 0; j < w; j++) nx[i] = sc_add(nx[i], sc_mul(st[j], p.mds[i][j]));
    st = nx;
  }
}
scm poseidon_hash_2(const bp_poseidon_params &p, const scm &xl, const scm &xr, int sbox) {  // gadget_poseidon.rs:428-443
  std::vector<scm> st(p.width, sc_zero());
  st[1] = xl; st[2] = xr; st[3] = sc_from_u64(101);
  poseidon_permutation(p, st, sbox);
  return st[1];
}
static int synthesize_sbox(bp_cs &cs, const LC &input, const scm &round_key, int sbox, bp_var &out) {
  LC inp = input + LC::constant(round_key);
  if (sbox == BP_SBOX_CUBE) {  // gadget_poseidon.rs:141-150
    bp_var a[3], b[3];
    cs.multiply(inp, inp, a);
    cs.multiply(LC(a[2]), LC(a[0]), b);
    out = b[2];
    return BP_OK;
  }
  // gadget_poseidon.rs:153-185
  bp_var var_l, var_r, var_o; int has;
  int rc = cs.allocate_single(1, nullptr, &inp, &var_l, nullptr, &has); if (rc) return rc;
  rc = cs.allocate_single(2, nullptr, nullptr, &var_r, &var_o, &has); if (rc) return rc;
  is_nonzero_gadget(cs, var_l, var_r);
  constrain_lc_with_scalar(cs, LC(var_o), sc_one());
  out = var_r;
  return BP_OK;
}
// block ops of one constraint system share one set of parameters: equal content, whatever object the caller passes
static bool poseidon_params_equal(const bp_poseidon_params &a, const bp_poseidon_params &b) {
  if (a.width != b.width || a.full_rounds_beginning != b.full_rounds_beginning || a.full_rounds_end != b.full_rounds_end ||
      a.partial_rounds != b.partial_rounds || a.round_keys.size() != b.round_keys.size() || a.mds.size() != b.mds.size()) return false;
  if (memcmp(a.round_keys.data(), b.round_keys.data(), a.round_keys.size() * sizeof(scm))) return false;
  for (size_t i = 0; i < a.mds.size(); i++)
    if (a.mds[i].size() != b.mds[i].size() || memcmp(a.mds[i].data(), b.mds[i].data(), a.mds[i].size() * sizeof(scm))) return false;
  return true;
}
int poseidon_permutation_constraints(bp_cs &cs, const bp_poseidon_params &p, std::vector<LC> &st, int sbox) {  // gadget_poseidon.rs:282-399
  const uint32_t w = p.width, total = p.full_rounds_beginning + p.partial_rounds + p.full_rounds_end;
  if (st.size() != w) return BP_ERR_GADGET;
  // witness program: the whole permutation becomes one block op when it has the standard shape
  const bool block = w == POSEIDON_WIDTH && cs.pending < 0 && (cs.pparams == nullptr || poseidon_params_equal(*cs.pparams, p));
  PoseidonBlock blk{};
  if (block) {
    if (!cs.pparams) cs.pparams = std::make_shared<const bp_poseidon_params>(p);
    for (uint32_t i = 0; i < w; i++) blk.in_lc[i] = cs.add_wlc(st[i]);
    blk.sbox = (uint32_t)sbox; blk.first_mult = cs.num_mult;
  }
  size_t off = 0;
  for (uint32_t rnd = 0; rnd < total; rnd++) {
    bool full = rnd < p.full_rounds_beginning || rnd >= p.full_rounds_beginning + p.partial_rounds;
    std::vector<LC> outs(w);
    for (uint32_t i = 0; i < w; i++) {
      const scm &rk = p.round_keys[off++];
      if (full || i == w - 1) { bp_var o; int rc = synthesize_sbox(cs, st[i], rk, sbox, o); if (rc) return rc; outs[i] = LC(o); }
      else outs[i] = st[i] + LC::constant(rk);
    }
    std::vector<LC> nx(w);
    for (uint32_t j = 0; j < w; j++)
      for (uint32_t i = 0; i < w; i++) nx[i] += outs[j] * p.mds[i][j];
    for (uint32_t i = 0; i < w; i++) st[i] = full ? nx[i] : nx[i].simplified();
  }
  if (block && cs.num_mult > blk.first_mult) {
    for (uint32_t i = blk.first_mult; i < cs.num_mult; i++) { cs.tape[i] = TapeOp{}; cs.tape[i].opL = W_SKIP; }
    cs.tape[blk.first_mult].opL = W_POSEIDON; cs.tape[blk.first_mult].argL = (uint32_t)cs.pblocks.size();
    cs.pblocks.push_back(blk);
  }
  return BP_OK;
}
int poseidon_hash_2_constraints(bp_cs &cs, const bp_poseidon_params &p, const LC &xl, const LC &xr, const std::vector<LC> &statics, int sbox, LC &out) {
  if (statics.size() != p.width - 2) return BP_ERR_GADGET;  // gadget_poseidon.rs:445-468
  std::vector<LC> in;
  in.push_back(statics[0]); in.push_back(xl); in.push_back(xr);
  for (size_t i = 1; i < statics.size(); i++) in.push_back(statics[i]);
  int rc = poseidon_permutation_constraints(cs, p, in, sbox);
  if (rc) return rc;
  out = in[1];
  return BP_OK;
}

// ------------------------------------------------------------------------------------------------ VSMT-2
int vsmt2_verif_gadget(bp_cs &cs, const bp_poseidon_params &p, uint32_t depth, const LC &root, bp_var leaf, const bp_var *bits,
                       const bp_var *nodes, const bp_var *statics, uint32_t num_statics) {  // gadget_vsmt_2.rs:171-209
  std::vector<LC> st;
  for (uint32_t i = 0; i < num_statics; i++) st.push_back(LC(statics[i]));
  LC prev;
  for (uint32_t i = 0; i < depth; i++) {
    LC leaf_lc = i == 0 ? LC(leaf) : prev;
    LC one_minus = LC(var_one()) - LC(bits[i]);
    bp_var l1[3], l2[3], r1[3], r2[3];
    cs.multiply(one_minus, leaf_lc, l1);
    cs.multiply(LC(bits[i]), LC(nodes[i]), l2);
    LC left = LC(l1[2]) + LC(l2[2]);
    cs.multiply(LC(bits[i]), leaf_lc, r1);
    cs.multiply(one_minus, LC(nodes[i]), r2);
    LC right = LC(r1[2]) + LC(r2[2]);
    int rc = poseidon_hash_2_constraints(cs, p, left, right, st, BP_SBOX_INVERSE, prev);
    if (rc) return rc;
  }
  cs.constrain(prev - root);  // constrain_lc_with_scalar(cs, prev_hash, root), gadget_vsmt_2.rs:206
  return BP_OK;
}

// ------------------------------------------------------------------------------------------------ Poseidon 4:1, VSMT-4
scm poseidon_hash_4(const bp_poseidon_params &p, const scm in[4], int sbox) {  // gadget_poseidon.rs:488-503
  std::vector<scm> st(p.width, sc_zero());
  for (int i = 0; i < 4; i++) st[1 + i] = in[i];
  st[5] = sc_from_u64(101);
  poseidon_permutation(p, st, sbox);
  return st[1];
}
int poseidon_hash_4_constraints(bp_cs &cs, const bp_poseidon_params &p, const LC in[4], const std::vector<LC> &statics, int sbox, LC &out) {
  if (statics.size() != p.width - 4) return BP_ERR_GADGET;  // gadget_poseidon.rs:505-530
  std::vector<LC> inputs;
  inputs.push_back(statics[0]);
  for (int i = 0; i < 4; i++) inputs.push_back(in[i]);
  for (size_t i = 1; i < statics.size(); i++) inputs.push_back(statics[i]);
  int rc = poseidon_permutation_constraints(cs, p, inputs, sbox);
  if (rc) return rc;
  out = inputs[1];
  return BP_OK;
}
// vanilla_merkle_merkle_tree_4_verif_gadget (gadget_vsmt_4.rs:199-312) for `levels` levels (the reference hard-codes
// 4 * LeafIndexBytes); digits = base-4 digits of the leaf index, least significant first (prover side; NULL otherwise);
// nodes = 3 * levels siblings, the triple (N1, N2, N3) of the first processed level LAST (the reference pops from the end)
int vsmt4_verif_gadget(bp_cs &cs, const bp_poseidon_params &p, uint32_t levels, const LC &root, bp_var leaf, bp_var leaf_index,
                       const uint8_t *digits, const bp_var *nodes, const bp_var *statics, uint32_t num_statics) {
  std::vector<LC> st;
  for (uint32_t i = 0; i < num_statics; i++) st.push_back(LC(statics[i]));
  LC prev(leaf);
  LC index_lc;  // starts as -leaf_index (gadget_vsmt_4.rs:217)
  index_lc.terms.push_back(Term{leaf_index, sc_neg(sc_one())});
  scm exp4 = sc_one();
  const scm two = sc_from_u64(2), four = sc_from_u64(4);
  long top = 3L * levels;
  for (uint32_t lvl = 0; lvl < levels; lvl++) {
    bp_var b0v[3], b1v[3]; int rc;
    for (int which = 0; which < 2; which++) {  // the two bits of this base-4 digit, each with its complement (gadget_vsmt_4.rs:227-241)
      bp_var *o = which ? b1v : b0v;
      if (cs.is_prover) {
        if (!digits) return BP_ERR_MISSING_ASSIGNMENT;
        const uint64_t bit = (digits[lvl] >> which) & 1;
        scm a = sc_from_u64(bit), b = sc_from_u64(1 - bit);
        rc = cs.allocate_multiplier(&a, &b, o);
      } else rc = cs.allocate_multiplier(nullptr, nullptr, o);
      if (rc) return rc;
      cs.constrain(LC(o[2]));
      cs.constrain(LC(o[0]) + (LC(o[1]) - LC::from_u64(1)));
    }
    const bp_var b0 = b0v[0], b0_1 = b0v[1], b1 = b1v[0], b1_1 = b1v[1];
    index_lc.terms.push_back(Term{b1, sc_mul(two, exp4)});
    index_lc.terms.push_back(Term{b0, exp4});
    const LC N3(nodes[--top]), N2(nodes[--top]), N1(nodes[--top]);
    bp_var t[3];
    cs.multiply(LC(b0_1), LC(b1_1), t); const LC b0_1_b1_1(t[2]);
    cs.multiply(LC(b0_1), LC(b1), t);   const LC b0_1_b1(t[2]);
    cs.multiply(LC(b0), LC(b1_1), t);   const LC b0_b1_1(t[2]);
    cs.multiply(LC(b0), LC(b1), t);     const LC b0_b1(t[2]);
    auto mul = [&](const LC &a, const LC &b) { bp_var o[3]; cs.multiply(a, b, o); return LC(o[2]); };
    LC c[4];
    { LC x = mul(b0_1_b1_1, prev), y = mul(LC(b0), N1), z = mul(b0_1_b1, N1); c[0] = x + y + z; }
    { LC x = mul(b0_1_b1_1, N1), y = mul(b0_b1_1, prev), z = mul(b0_1_b1, N2), u = mul(b0_b1, N2); c[1] = x + y + z + u; }
    { LC x = mul(LC(b1_1), N2), y = mul(b0_1_b1, prev), z = mul(b0_b1, N3); c[2] = x + y + z; }
    { LC x = mul(LC(b1_1), N3), y = mul(b0_1_b1, N3), z = mul(b0_b1, prev); c[3] = x + y + z; }
    rc = poseidon_hash_4_constraints(cs, p, c, st, BP_SBOX_INVERSE, prev);
    if (rc) return rc;
    exp4 = sc_mul(exp4, four);
  }
  cs.constrain(index_lc);
  cs.constrain(prev - root);
  return BP_OK;
}

// ------------------------------------------------------------------------------------------------ MiMC
scm mimc_native(const scm &xl_, const scm &xr_, uint32_t rounds, const scm *constants) {  // gadget_mimc.rs:19-39
  scm xl = xl_, xr = xr_;
  for (uint32_t j = 0; j < rounds; j++) {
    scm t = sc_add(xl, constants[j]);
    scm nl = sc_add(sc_mul(sc_sqr(t), t), xr);
    xr = xl; xl = nl;
  }
  return xl;
}
int mimc_gadget(bp_cs &cs, bp_var left, bp_var right, uint32_t rounds, const scm *constants, const LC &image) {  // gadget_mimc.rs:41-79
  LC lv(left), rv(right);
  for (uint32_t j = 0; j < rounds; j++) {
    LC lpc = lv + LC::constant(constants[j]);
    bp_var a[3], b[3];
    cs.multiply(lpc, lpc, a);
    cs.multiply(LC(a[2]), LC(a[0]), b);
    LC tmp = LC(b[2]) + rv;
    rv = lv; lv = tmp;
  }
  cs.constrain(lv - image);  // constrain_lc_with_scalar(cs, res_v, image), gadget_mimc.rs:50
  return BP_OK;
}

// ------------------------------------------------------------------------------------------------ range / bound check
int positive_no_gadget(bp_cs &cs, bp_var v, bool has_assignment, uint64_t value, uint32_t bit_size) {  // r1cs_utils.rs:20-48
  LC cv; cv.terms.push_back(Term{v, sc_neg(sc_one())});
  scm exp2 = sc_one();
  for (uint32_t i = 0; i < bit_size; i++) {
    bp_var o[3]; int rc;
    if (has_assignment) {
      uint64_t bit = i < 64 ? (value >> i) & 1 : 0;
      scm a = sc_from_u64(1 - bit), b = sc_from_u64(bit);
      rc = cs.allocate_multiplier(&a, &b, o);
    } else rc = cs.allocate_multiplier(nullptr, nullptr, o);
    if (rc) return rc;
    cs.constrain(LC(o[2]));
    cs.constrain(LC(o[0]) + (LC(o[1]) - LC::from_u64(1)));
    cv.terms.push_back(Term{o[1], exp2});
    exp2 = sc_add(exp2, exp2);
  }
  cs.constrain(cv);
  return BP_OK;
}
int bound_check_gadget(bp_cs &cs, bp_var v, bp_var a, bp_var b, bool has_assignment, uint64_t vv, uint64_t av, uint64_t bv, uint64_t max,
                       uint64_t min, uint32_t bit_size) {  // gadget_bound_check.rs:18-45
  (void)vv;
  cs.constrain(LC(v) - LC::from_u64(min) - LC(a));
  cs.constrain(LC::from_u64(max) - LC(v) - LC(b));
  constrain_lc_with_scalar(cs, LC(a) + LC(b), sc_from_u64(max - min));
  int rc = positive_no_gadget(cs, a, has_assignment, av, bit_size); if (rc) return rc;
  return positive_no_gadget(cs, b, has_assignment, bv, bit_size);
}
