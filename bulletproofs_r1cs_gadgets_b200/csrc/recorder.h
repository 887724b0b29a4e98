// Host-side constraint-system recorder: the ConstraintSystem / Prover / Verifier surface the
// reference's gadgets are generic over (`bulletproofs::r1cs` of the fork,
// reference Cargo.toml:22-26).  It records multipliers, constraints and -- for batched proving --
// a witness program; all group arithmetic and the proof itself happen on the device (engine.h).
#pragma once
#include <array>
#include <memory>
#include <string>
#include <vector>
#include "../../include/bp_b200.h"
#include "engine.h"

struct Term { bp_var var; scm coeff; };  // coefficient in Montgomery form

// LinearCombination (reference: From<Variable|Scalar|u64>, + - with LC/Variable/Scalar, * Scalar,
// FromIterator<(Variable,Scalar)>, get_terms; e.g. src/r1cs_utils.rs:38-45, src/gadget_poseidon.rs:99-112,299)
struct LC {
  std::vector<Term> terms;
  LC() {}
  LC(bp_var v) { terms.push_back(Term{v, sc_one()}); }
  static LC constant(const scm &c) { LC l; l.terms.push_back(Term{bp_var{BP_VAR_ONE, 0}, c}); return l; }
  static LC from_u64(uint64_t x) { return constant(sc_from_u64(x)); }
  LC &operator+=(const LC &o) { terms.insert(terms.end(), o.terms.begin(), o.terms.end()); return *this; }
  LC &operator-=(const LC &o) { for (const Term &t : o.terms) terms.push_back(Term{t.var, sc_neg(t.coeff)}); return *this; }
  LC operator+(const LC &o) const { LC r = *this; r += o; return r; }
  LC operator-(const LC &o) const { LC r = *this; r -= o; return r; }
  LC operator*(const scm &s) const { LC r; r.terms.reserve(terms.size()); for (const Term &t : terms) r.terms.push_back(Term{t.var, sc_mul(t.coeff, s)}); return r; }
  // simplify_lc (reference src/gadget_poseidon.rs:99-112): merge duplicate variables; first-occurrence order
  LC simplified() const;
};
inline bp_var var_one() { return bp_var{BP_VAR_ONE, 0}; }

struct bp_gens { BpGens *g; };

struct bp_cs {
  const bp_gens *gens = nullptr;
  bool is_prover = false;
  std::vector<uint8_t> label;
  // committed (high-level) variables
  std::vector<scm> v, vbl;                       // prover only
  std::vector<std::array<uint8_t, 32>> V;        // both
  // multipliers; assignments on the prover only
  std::vector<scm> aL, aR, aO;
  uint32_t num_mult = 0;
  long pending = -1;
  // constraints (CSR)
  std::vector<uint32_t> cons_ptr{0};
  std::vector<Term> terms;
  // witness program
  std::vector<TapeOp> tape;
  std::vector<uint32_t> wlc_ptr{0};
  std::vector<Term> wlc_terms;
  std::vector<scm> aux;                          // prover only: values of the auxiliary inputs
  uint32_t naux = 0;
  std::vector<scm> pub;                          // public inputs (values for this cs; zero when recording only)
  std::vector<PoseidonBlock> pblocks;            // block ops of the witness program
  std::shared_ptr<const struct bp_poseidon_params> pparams;  // OWN copy of the parameters shared by every block op (the caller may free its object before compiling)
  // commitments whose value the circuit fixes (allocate_statics: the reference's verifier computes them itself,
  // src/gadget_poseidon.rs:580-608): index among the committed variables and the expected compressed point
  std::vector<uint32_t> fixed_idx;
  std::vector<std::array<uint8_t, 32>> fixed_V;


  uint32_t add_wlc(const LC &lc) {
    wlc_terms.insert(wlc_terms.end(), lc.terms.begin(), lc.terms.end());
    wlc_ptr.push_back((uint32_t)wlc_terms.size());
    return (uint32_t)wlc_ptr.size() - 2;
  }
  bool eval(const LC &lc, scm &out) const;       // Prover: value; Verifier: false (None)
  void constrain(const LC &lc) { terms.insert(terms.end(), lc.terms.begin(), lc.terms.end()); cons_ptr.push_back((uint32_t)terms.size()); }
  void multiply(const LC &l, const LC &r, bp_var out[3]);
  // allocate_multiplier(Some((l,r))) / (None): values become auxiliary inputs of the witness program
  int allocate_multiplier(const scm *l, const scm *r, bp_var out[3]);
  // allocate_single: `how` 0 = caller value (aux), 1 = value of lc, 2 = inverse of the pending left value
  int allocate_single(int how, const scm *value, const LC *lc, bp_var *var, bp_var *out_var, int *has_out);
  size_t num_constraints() const { return cons_ptr.size() - 1; }
};

// ---- gadgets (host language of the reference is Rust; this is the C++ restatement of its gadget layer) ----
struct bp_poseidon_params {
  uint32_t width, full_rounds_beginning, full_rounds_end, partial_rounds;
  std::vector<scm> round_keys;            // total_rounds * width
  std::vector<std::vector<scm>> mds;      // [width][width]
};
void poseidon_permutation(const bp_poseidon_params &p, std::vector<scm> &state, int sbox);
scm poseidon_hash_2(const bp_poseidon_params &p, const scm &xl, const scm &xr, int sbox);
int poseidon_permutation_constraints(bp_cs &cs, const bp_poseidon_params &p, std::vector<LC> &state, int sbox);
scm poseidon_hash_4(const bp_poseidon_params &p, const scm in[4], int sbox);
int poseidon_hash_4_constraints(bp_cs &cs, const bp_poseidon_params &p, const LC in[4], const std::vector<LC> &statics, int sbox, LC &out);
int vsmt4_verif_gadget(bp_cs &cs, const bp_poseidon_params &p, uint32_t levels, const LC &root, bp_var leaf, bp_var leaf_index,
                       const uint8_t *digits, const bp_var *nodes, const bp_var *statics, uint32_t num_statics);
int poseidon_hash_2_constraints(bp_cs &cs, const bp_poseidon_params &p, const LC &xl, const LC &xr, const std::vector<LC> &statics, int sbox, LC &out);
int vsmt2_verif_gadget(bp_cs &cs, const bp_poseidon_params &p, uint32_t depth, const LC &root, bp_var leaf, const bp_var *bits,
                       const bp_var *nodes, const bp_var *statics, uint32_t num_statics);
int mimc_gadget(bp_cs &cs, bp_var left, bp_var right, uint32_t rounds, const scm *constants, const LC &image);
scm mimc_native(const scm &xl, const scm &xr, uint32_t rounds, const scm *constants);
int positive_no_gadget(bp_cs &cs, bp_var v, bool has_assignment, uint64_t value, uint32_t bit_size);
int bound_check_gadget(bp_cs &cs, bp_var v, bp_var a, bp_var b, bool has_assignment, uint64_t vv, uint64_t av, uint64_t bv, uint64_t max,
                       uint64_t min, uint32_t bit_size);
