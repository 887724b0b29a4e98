/* bp_b200 -- C-ABI of the B200-native Bulletproofs R1CS prover / verifier.
 *
 * Drop-in boundary for the gadget crate lovesh/bulletproofs-r1cs-gadgets: every entry point below
 * replaces one item of the `bulletproofs::r1cs` / `bulletproofs::{PedersenGens,BulletproofGens}` /
 * `merlin::Transcript` surface the reference's gadgets call (reference Cargo.toml:18,22-26).  The
 * reference-side call site each one stands in for is cited as file:line under /root/reference.
 *
 * Conventions
 *   - scalars: 32 bytes, little-endian; inputs are reduced mod l, outputs are canonical;
 *   - points: 32-byte ristretto255 encodings (CompressedRistretto);
 *   - every function returns an int32 status (BP_OK = 0); nothing throws or aborts across the ABI;
 *   - the caller owns every buffer; handles are opaque and released by the matching *_free;
 *   - a handle may be used from one host thread at a time (the reference objects are &mut-exclusive);
 *   - the 32 bytes of OS entropy that Prover::prove / Verifier::verify draw internally in the reference
 *     are an explicit argument here, so results are reproducible;
 *   - all proving / verifying arithmetic runs in CUDA kernels on the current device; without a usable
 *     device every entry point that needs one returns BP_ERR_NO_DEVICE.  There is no CPU path.
 *
 * Proof bytes: the R1CSProof field tuple, 32 bytes each, in order
 *   A_I1 A_O1 S1 A_I2 A_O2 S2 T_1 T_3 T_4 T_5 T_6 t_x t_x_blinding e_blinding L_0 R_0 .. L_{k-1} R_{k-1} a b
 * (the reference never serialises a proof; this is the untagged layout).
 */
#ifndef BP_B200_H
#define BP_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* R1CSError (reference src/gadget_poseidon.rs:136, src/gadget_range_proof.rs:28) */
#define BP_OK 0
#define BP_ERR_INVALID_GENERATORS_LENGTH 1
#define BP_ERR_FORMAT 2
#define BP_ERR_VERIFICATION 3
#define BP_ERR_MISSING_ASSIGNMENT 4
#define BP_ERR_GADGET 5
#define BP_ERR_INVALID_ARGUMENT 6
#define BP_ERR_NO_DEVICE 7
#define BP_ERR_CUDA 8
#define BP_ERR_OOM 9

/* Variable (reference src/gadget_vsmt_2.rs:192 Variable::One(), src/gadget_poseidon.rs:101 HashMap key) */
#define BP_VAR_COMMITTED 0
#define BP_VAR_MULT_LEFT 1
#define BP_VAR_MULT_RIGHT 2
#define BP_VAR_MULT_OUT 3
#define BP_VAR_ONE 4
/* a per-proof public input of a compiled circuit (extension for batching: the reference bakes public values such as the
 * Merkle root into the constraint system as constants, src/gadget_vsmt_2.rs:206; a batch shares one circuit) */
#define BP_VAR_PUBLIC 5
typedef struct bp_var { uint32_t kind; uint32_t index; } bp_var;
/* one (Variable, Scalar) term of a LinearCombination (reference src/r1cs_utils.rs:45, get_terms src/gadget_poseidon.rs:102) */
typedef struct bp_term { bp_var var; uint8_t coeff[32]; } bp_term;

typedef struct bp_gens bp_gens;       /* PedersenGens::default() + BulletproofGens::new(capacity, 1) */
typedef struct bp_cs bp_cs;           /* a Prover or Verifier with its Transcript */
typedef struct bp_circuit bp_circuit; /* a compiled constraint system for batched proving */

int32_t bp_version(void);
/* number of CUDA kernels this library has launched in this process (bench.py reports it) */
int64_t bp_launch_count(void);
/* per-kernel device timing: on = 1 brackets every kernel launch with CUDA events on its stream (about 2.5 % overhead on a
 * depth-32 proving step), on = 2 only the launches of the dominant kernel (KBucketAccumulate), 0 = off.
 * bp_profile_report (after the caller synchronised) writes lines "kernel launches total_ms threads" and clears the records. */
void bp_profile_enable(int32_t on);
int32_t bp_profile_report(char *buf, size_t cap);

/* ---- generators --------------------------------------------------------------------------------
 * PedersenGens::default() (reference src/gadget_mimc.rs:99) and BulletproofGens::new(capacity, 1)
 * (reference src/gadget_vsmt_2.rs:290): derived on the device, deterministic. */
int32_t bp_gens_new(uint32_t capacity, bp_gens **out);
void bp_gens_free(bp_gens *g);
uint32_t bp_gens_capacity(const bp_gens *g);
/* compressed B and B_blinding */
int32_t bp_gens_pedersen(const bp_gens *g, uint8_t B[32], uint8_t B_blinding[32]);
/* first `count` generators of chain G (which = 0) or H (which = 1), party 0, compressed */
int32_t bp_gens_export(const bp_gens *g, int32_t which, uint32_t count, uint8_t *out);
/* pc_gens.commit(v, r).compress() (reference src/gadget_poseidon.rs:584-587), `count` at once */
int32_t bp_pc_commit(const bp_gens *g, uint32_t count, const uint8_t *v, const uint8_t *r, uint8_t *out);

/* ---- tier 1: one constraint system at a time (semantic mirror) ----------------------------------
 * Transcript::new(label) + Prover::new(&pc_gens, &mut transcript)   (reference src/gadget_mimc.rs:113-114) */
int32_t bp_prover_new(const bp_gens *g, const uint8_t *label, size_t label_len, bp_cs **out);
/* Transcript::new(label) + Verifier::new(&mut transcript)           (reference src/gadget_mimc.rs:146-147) */
int32_t bp_verifier_new(const bp_gens *g, const uint8_t *label, size_t label_len, bp_cs **out);
void bp_cs_free(bp_cs *cs);
/* prover.commit(v, v_blinding) -> (CompressedRistretto, Variable)   (reference src/gadget_mimc.rs:117) */
int32_t bp_prover_commit(bp_cs *cs, const uint8_t v[32], const uint8_t v_blinding[32], uint8_t V_out[32], bp_var *var);
/* verifier.commit(V) -> Variable                                    (reference src/gadget_mimc.rs:149) */
int32_t bp_verifier_commit(bp_cs *cs, const uint8_t V[32], bp_var *var);
/* cs.multiply(left, right) -> (l, r, o)                             (reference src/gadget_mimc.rs:71) */
int32_t bp_cs_multiply(bp_cs *cs, const bp_term *left, size_t nleft, const bp_term *right, size_t nright, bp_var out[3]);
/* cs.allocate_multiplier(Some((l, r)) | None) -> (l, r, o)          (reference src/r1cs_utils.rs:29) ; l == NULL means None */
int32_t bp_cs_allocate_multiplier(bp_cs *cs, const uint8_t *l, const uint8_t *r, bp_var out[3]);
/* fork API cs.allocate_single(Option<Scalar>) -> (Variable, Option<Variable>)   (reference src/gadget_poseidon.rs:165-166) */
int32_t bp_cs_allocate_single(bp_cs *cs, const uint8_t *value, bp_var *var, bp_var *out_var, int32_t *has_out);
/* fork API cs.evaluate_lc(&lc) -> Option<Scalar>                    (reference src/gadget_poseidon.rs:160); BP_ERR_MISSING_ASSIGNMENT = None */
int32_t bp_cs_evaluate_lc(bp_cs *cs, const bp_term *lc, size_t n, uint8_t out[32]);
/* declares the next public input; `value` is its value for this constraint system (may be NULL on a recording-only cs) */
int32_t bp_cs_public_input(bp_cs *cs, const uint8_t *value, bp_var *var);
/* cs.constrain(lc)                                                  (reference src/r1cs_utils.rs:35) */
int32_t bp_cs_constrain(bp_cs *cs, const bp_term *lc, size_t n);
/* fork API num_constraints / num_multipliers                       (reference src/gadget_vsmt_2.rs:345) */
uint64_t bp_cs_num_constraints(const bp_cs *cs);
uint64_t bp_cs_num_multipliers(const bp_cs *cs);
uint64_t bp_cs_num_commitments(const bp_cs *cs);
size_t bp_cs_proof_len(const bp_cs *cs);
/* prover.prove(&bp_gens) -> R1CSProof                               (reference src/gadget_vsmt_2.rs:347) */
int32_t bp_prover_prove(bp_cs *cs, const uint8_t entropy[32], uint8_t *proof, size_t *proof_len);
/* verifier.verify(&proof, &pc_gens, &bp_gens)                       (reference src/gadget_vsmt_2.rs:395) */
int32_t bp_verifier_verify(bp_cs *cs, const uint8_t *proof, size_t proof_len, const uint8_t entropy[32]);

/* ---- gadgets of the reference, run against a bp_cs (host side of the hot path) -------------------
 * Scalar assignments are prover-only; pass NULL on the verifier side (AllocatedScalar.assignment = None,
 * reference src/r1cs_utils.rs:8-11). */
typedef struct bp_poseidon_params bp_poseidon_params; /* PoseidonParams (reference src/gadget_poseidon.rs:27-94) */
#define BP_SBOX_CUBE 0
#define BP_SBOX_INVERSE 1
/* PoseidonParams::new(width, full_b, full_e, partial) with the constants of src/poseidon_constants.rs;
 * `constants` = 36 MDS entries then round keys, 32-byte LE as loaded by get_scalar_from_hex (src/scalar_utils.rs:232-237) */
int32_t bp_poseidon_params_new(const uint8_t *constants, size_t nconstants, uint32_t width, uint32_t full_rounds_beginning,
                               uint32_t full_rounds_end, uint32_t partial_rounds, bp_poseidon_params **out);
void bp_poseidon_params_free(bp_poseidon_params *p);
/* Poseidon_permutation (src/gadget_poseidon.rs:189-280) and Poseidon_hash_2 (:428-443), evaluated natively */
int32_t bp_poseidon_hash_2(const bp_poseidon_params *p, const uint8_t xl[32], const uint8_t xr[32], int32_t sbox, uint8_t out[32]);
/* allocate_statics_for_prover / _for_verifier (src/gadget_poseidon.rs:554-608) */
int32_t bp_gadget_allocate_statics(bp_cs *cs, uint32_t num_statics, bp_var *out_vars);
/* Poseidon_hash_2_gadget (src/gadget_poseidon.rs:470-486) */
int32_t bp_gadget_poseidon_hash_2(bp_cs *cs, const bp_poseidon_params *p, bp_var xl, bp_var xr, const bp_var *statics,
                                  uint32_t num_statics, int32_t sbox, const uint8_t expected_hash[32]);
int32_t bp_gadget_poseidon_hash_2_public(bp_cs *cs, const bp_poseidon_params *p, bp_var xl, bp_var xr, const bp_var *statics,
                                         uint32_t num_statics, int32_t sbox, bp_var expected_hash);
/* vanilla_merkle_merkle_tree_verif_gadget (src/gadget_vsmt_2.rs:171-209) */
int32_t bp_gadget_vsmt2_verif(bp_cs *cs, const bp_poseidon_params *p, uint32_t depth, const uint8_t root[32], bp_var leaf,
                              const bp_var *leaf_index_bits, const bp_var *proof_nodes, const bp_var *statics, uint32_t num_statics);
/* same circuit with the root as a public-input variable (BP_VAR_PUBLIC) instead of a baked-in constant */
int32_t bp_gadget_vsmt2_verif_public(bp_cs *cs, const bp_poseidon_params *p, uint32_t depth, bp_var root, bp_var leaf,
                                     const bp_var *leaf_index_bits, const bp_var *proof_nodes, const bp_var *statics, uint32_t num_statics);
/* Poseidon_hash_4 / Poseidon_hash_4_gadget (src/gadget_poseidon.rs:488-551): inputs [0, x0..x3, 101], output lane 1; width 6, two statics */
int32_t bp_poseidon_hash_4(const bp_poseidon_params *p, const uint8_t in[4][32], int32_t sbox, uint8_t out[32]);
int32_t bp_gadget_poseidon_hash_4(bp_cs *cs, const bp_poseidon_params *p, const bp_var in[4], const bp_var *statics, uint32_t num_statics,
                                  int32_t sbox, const uint8_t expected_hash[32]);
int32_t bp_gadget_poseidon_hash_4_public(bp_cs *cs, const bp_poseidon_params *p, const bp_var in[4], const bp_var *statics, uint32_t num_statics,
                                         int32_t sbox, bp_var expected_hash);
/* vanilla_merkle_merkle_tree_4_verif_gadget (src/gadget_vsmt_4.rs:199-312), `levels` levels of the 4-ary tree (the reference
 * hard-codes 4 * LeafIndexBytes).  leaf_index is a committed variable; index_digits[levels] are its base-4 digits, least
 * significant first (prover side only; NULL on the verifier -- they become per-proof auxiliary inputs of a batched circuit:
 * for every level the values b0, 1-b0, b1, 1-b1 in that order).  proof_nodes holds 3 * levels siblings, the triple of the
 * first processed (leaf) level LAST, as the reference pops them from the end of its vector. */
int32_t bp_gadget_vsmt4_verif(bp_cs *cs, const bp_poseidon_params *p, uint32_t levels, const uint8_t root[32], bp_var leaf, bp_var leaf_index,
                              const uint8_t *index_digits, const bp_var *proof_nodes, const bp_var *statics, uint32_t num_statics);
int32_t bp_gadget_vsmt4_verif_public(bp_cs *cs, const bp_poseidon_params *p, uint32_t levels, bp_var root, bp_var leaf, bp_var leaf_index,
                                     const uint8_t *index_digits, const bp_var *proof_nodes, const bp_var *statics, uint32_t num_statics);
/* mimc_gadget (src/gadget_mimc.rs:41-79) */
int32_t bp_gadget_mimc(bp_cs *cs, bp_var left, bp_var right, uint32_t rounds, const uint8_t *constants, const uint8_t image[32]);
int32_t bp_gadget_mimc_public(bp_cs *cs, bp_var left, bp_var right, uint32_t rounds, const uint8_t *constants, bp_var image);
/* mimc (src/gadget_mimc.rs:19-39), evaluated natively */
int32_t bp_mimc(const uint8_t xl[32], const uint8_t xr[32], uint32_t rounds, const uint8_t *constants, uint8_t out[32]);
/* bound_check_gadget (src/gadget_bound_check.rs:18-45); v/a/b assignments as u64, has_assignment = 0 on the verifier */
int32_t bp_gadget_bound_check(bp_cs *cs, bp_var v, bp_var a, bp_var b, int32_t has_assignment, uint64_t v_val, uint64_t a_val,
                              uint64_t b_val, uint64_t max, uint64_t min, uint32_t bit_size);

/* ---- tier 2: batched proving over a compiled circuit (what the GPU is for) -----------------------
 * A circuit is the constraint system a gadget recorded on a Verifier-side bp_cs (no assignments) plus the
 * witness program the recorder derived: how each multiplier's (left, right) follows from the committed
 * values -- cs.multiply evaluates its two linear combinations, the inverse S-box inverts its left value,
 * allocate_multiplier reads a caller-supplied auxiliary input.  Every proof of a batch shares it. */
int32_t bp_circuit_compile(const bp_cs *recorded, bp_circuit **out);
/* same, from flat arrays: constraint q holds terms cons_ptr[q]..cons_ptr[q+1] (kinds BP_VAR_*); no witness program */
int32_t bp_circuit_from_arrays(uint32_t n_multipliers, uint32_t n_commitments, uint32_t n_constraints, const uint32_t *cons_ptr,
                               const uint8_t *kind, const uint32_t *idx, const uint8_t *coeff, bp_circuit **out);
void bp_circuit_free(bp_circuit *c);
/* Frees the device workspace the batch calls keep between calls (sized by the largest chunk proved or verified so far: ~26 MB per
 * proof of a chunk at depth 32, 72 GB budget); the next batch call allocates it again.  InvalidArgument while a batch is
 * between bp_prove_stream_begin and _finish. */
int32_t bp_circuit_release_workspace(bp_circuit *c);
uint32_t bp_circuit_num_multipliers(const bp_circuit *c);
uint32_t bp_circuit_num_constraints(const bp_circuit *c);
uint32_t bp_circuit_num_commitments(const bp_circuit *c);
uint32_t bp_circuit_num_aux(const bp_circuit *c);
uint32_t bp_circuit_num_public(const bp_circuit *c);
int32_t bp_circuit_has_witness_program(const bp_circuit *c);
size_t bp_circuit_proof_len(const bp_circuit *c);

/* B independent proofs.  HOST buffers: v, v_blinding [B][m][32]; entropy [B][32]; aux [B][num_aux][32] and
 * pub [B][num_public][32] (NULL when the count is 0); witness aL/aR/aO [B][n][32] or all NULL to run the witness program on the device.
 * Outputs: V [B][m][32], proofs [B][proof_len], status [B].  Copies host<->device inside the call. */
int32_t bp_prove_batch(const bp_gens *g, bp_circuit *c, uint32_t B, const uint8_t *label, size_t label_len, const uint8_t *v,
                       const uint8_t *v_blinding, const uint8_t *entropy, const uint8_t *aux, const uint8_t *pub, const uint8_t *aL,
                       const uint8_t *aR, const uint8_t *aO, uint8_t *V_out, uint8_t *proofs, int32_t *status);
/* same with every buffer already resident in device memory (cudaMalloc'ed by the caller), on `stream`
 * (a cudaStream_t cast to void*; NULL = default stream).  Returns after enqueueing; the caller synchronises. */
int32_t bp_prove_batch_device(const bp_gens *g, bp_circuit *c, uint32_t B, const uint8_t *label, size_t label_len,
                              const uint8_t *d_v, const uint8_t *d_v_blinding, const uint8_t *d_entropy, const uint8_t *d_aux,
                              const uint8_t *d_pub, const uint8_t *d_aL, const uint8_t *d_aR, const uint8_t *d_aO, uint8_t *d_V_out,
                              uint8_t *d_proofs, int32_t *d_status, void *stream);
/* Streaming form of bp_prove_batch_device for a pipeline of batches (same arguments, same bytes).  _begin launches the
 * latency-bound first phase of the batch -- commitments V, the witness program, the 2n+3 sequential blinding draws of the
 * transcript RNG -- on internal streams ordered after the work already enqueued on `stream`, and returns.  _finish enqueues
 * the rest (every MSM, the inner-product argument) on `stream`.  A circuit has two slots (0, 1): enqueueing begin(batch k+1)
 * BEFORE finish(batch k) lets the first phase of the next batch run beside the MSM phase of the current one instead of
 * idling the GPU at the head of every batch.  The buffers of a slot belong to the library from _begin until the work of
 * its _finish has completed.  begin(s) + finish(s) == bp_prove_batch_device. */
int32_t bp_prove_stream_begin(const bp_gens *g, bp_circuit *c, int32_t slot, uint32_t B, const uint8_t *label, size_t label_len,
                              const uint8_t *d_v, const uint8_t *d_v_blinding, const uint8_t *d_entropy, const uint8_t *d_aux,
                              const uint8_t *d_pub, const uint8_t *d_aL, const uint8_t *d_aR, const uint8_t *d_aO, uint8_t *d_V,
                              uint8_t *d_proofs, int32_t *d_status, void *stream);
int32_t bp_prove_stream_finish(const bp_gens *g, bp_circuit *c, int32_t slot, void *stream);
/* the same with HOST buffers (pinned memory makes the copies asynchronous): _begin_host copies the inputs to per-slot device
 * staging and begins; _finish_host finishes, copies V / proofs / status back and waits for them.  Witness from the circuit's
 * witness program only. */
int32_t bp_prove_stream_begin_host(const bp_gens *g, bp_circuit *c, int32_t slot, uint32_t B, const uint8_t *label, size_t label_len,
                                   const uint8_t *v, const uint8_t *v_blinding, const uint8_t *entropy, const uint8_t *aux,
                                   const uint8_t *pub, void *stream);
int32_t bp_prove_stream_finish_host(const bp_gens *g, bp_circuit *c, int32_t slot, uint8_t *V_out, uint8_t *proofs, int32_t *status,
                                    void *stream);

/* B independent verifications of proofs over the same circuit; status[p] = BP_OK or BP_ERR_VERIFICATION / BP_ERR_FORMAT */
int32_t bp_verify_batch(const bp_gens *g, bp_circuit *c, uint32_t B, const uint8_t *label, size_t label_len, const uint8_t *V,
                        const uint8_t *proofs, const uint8_t *entropy, const uint8_t *pub, int32_t *status);
int32_t bp_verify_batch_device(const bp_gens *g, bp_circuit *c, uint32_t B, const uint8_t *label, size_t label_len,
                               const uint8_t *d_V, const uint8_t *d_proofs, const uint8_t *d_entropy, const uint8_t *d_pub, int32_t *d_status,
                               void *stream);

/* Cross-proof batched verification (not part of the reference's dependency, which verifies one
 * proof at a time -- bulletproofs::r1cs::Verifier::verify, reference call sites src/gadget_vsmt_2.rs:395).  The B verification
 * equations are combined with random weights drawn from each proof's verifier RNG, so the 2N+2 generator rows shared by all
 * proofs are paid once per batch.  status[p] reports structural failures only (BP_ERR_FORMAT / BP_ERR_VERIFICATION for a
 * non-canonical scalar, an undecodable or an identity point); such proofs are left out of the combination.
 * *combined = BP_OK when the combined equation of the remaining proofs holds, BP_ERR_VERIFICATION otherwise (then at least one
 * of them is invalid: fall back to bp_verify_batch to find it).  A batch with an invalid proof passes with probability 2^-252.
 * Needs the generators' shift table (capacity <= 2^22). */
int32_t bp_verify_batch_combined(const bp_gens *g, bp_circuit *c, uint32_t B, const uint8_t *label, size_t label_len, const uint8_t *V,
                                 const uint8_t *proofs, const uint8_t *entropy, const uint8_t *pub, int32_t *status, int32_t *combined);
int32_t bp_verify_batch_combined_device(const bp_gens *g, bp_circuit *c, uint32_t B, const uint8_t *label, size_t label_len,
                                        const uint8_t *d_V, const uint8_t *d_proofs, const uint8_t *d_entropy, const uint8_t *d_pub,
                                        int32_t *d_status, int32_t *d_combined, void *stream);

/* ---- wire format (SURVEY.md App. A.7, R1CSProof::to_bytes / from_bytes of the 2.0-era upstream; the reference itself never
 * serialises a proof).  The library's proof buffers are the untagged field tuple (14 + 2k + 2 elements of 32 bytes).  The
 * tagged form is 1 byte (0 = one-phase: A_I2, A_O2, S2 are the identity and are dropped; 1 = two-phase: all six kept) followed
 * by the remaining elements.  Both return the number of bytes written, or a negative BP_ERR_* (FORMAT for a bad tag/length). */
int64_t bp_proof_to_wire(const uint8_t *proof, size_t proof_len, uint8_t *out, size_t out_cap);
int64_t bp_proof_from_wire(const uint8_t *wire, size_t wire_len, uint8_t *proof_out, size_t out_cap);

/* ---- device-side sparse Merkle tree (the step before the hot path) ----------------
 * Batched VanillaSparseMerkleTree (reference src/gadget_vsmt_2.rs:27-166) with every Poseidon hash on the GPU.  Leaf indices
 * are uint64 (depth <= 63; the reference's TreeDepth = 253 stays with the host mirror); values are 32-byte LE scalars.
 * Nodes are stored by position, so the tree keeps its latest state only (the reference's content-addressed map also
 * keeps every older node; `get`, `update` and `root` cannot tell the difference). */
typedef struct bp_vsmt2 bp_vsmt2;
/* VanillaSparseMerkleTree::new (src/gadget_vsmt_2.rs:36-61): empty-subtree hashes, root of the empty tree.  The reference
 * always hashes with SboxType::Inverse (:45,79,85). */
int32_t bp_vsmt2_new(const bp_poseidon_params *p, uint32_t depth, int32_t sbox, bp_vsmt2 **out);
void bp_vsmt2_free(bp_vsmt2 *t);
uint32_t bp_vsmt2_depth(const bp_vsmt2 *t);
uint64_t bp_vsmt2_num_nodes(const bp_vsmt2 *t); /* stored (non-empty) nodes of all levels */
int32_t bp_vsmt2_root(const bp_vsmt2 *t, uint8_t root[32]);
int32_t bp_vsmt2_empty_hashes(const bp_vsmt2 *t, uint8_t *out /* [depth+1][32], [0] = leaf level */);
/* count x VanillaSparseMerkleTree::update (src/gadget_vsmt_2.rs:63-98) in one pass: level by level, one hash per DISTINCT
 * parent.  A repeated index keeps its last value, as sequential updates would.  root_out may be NULL. */
int32_t bp_vsmt2_update_batch(bp_vsmt2 *t, uint32_t count, const uint64_t *idx, const uint8_t *vals, uint8_t root_out[32]);
/* count x VanillaSparseMerkleTree::get (src/gadget_vsmt_2.rs:101-131): leaves [count][32], proofs [count][depth][32] with
 * the siblings root -> leaf, the order `get` pushes them */
int32_t bp_vsmt2_get_batch(const bp_vsmt2 *t, uint32_t count, const uint64_t *idx, uint8_t *leaves, uint8_t *proofs);
/* committed values of count membership proofs in the commit order of the reference's prover (src/gadget_vsmt_2.rs:296-330):
 * v [count][2*depth+5][32] = leaf, index bits LSB first, siblings leaf level first, statics 0,101,0,0 -- the `v` argument of
 * bp_prove_batch[_device] for the circuit of bp_gadget_vsmt2_verif_public; pub [count][32] = the root (may be NULL).
 * _device: device pointers (idx uint64 [count]), asynchronous on `stream`. */
int32_t bp_vsmt2_witness_batch(const bp_vsmt2 *t, uint32_t count, const uint64_t *idx, uint8_t *v, uint8_t *pub);
int32_t bp_vsmt2_witness_batch_device(const bp_vsmt2 *t, uint32_t count, const uint64_t *d_idx, uint8_t *d_v, uint8_t *d_pub, void *stream);
/* The same with 256-bit indices, idx32 [count][32] little-endian integers below 2^depth: the reference's tree is TreeDepth = 253
 * levels deep and keyed by Scalars (src/gadget_vsmt_2.rs:23,63-131); bp_vsmt2_new accepts depth <= 253.  The 64-bit forms above
 * are conveniences for depth <= 63. */
int32_t bp_vsmt2_update_batch_wide(bp_vsmt2 *t, uint32_t count, const uint8_t *idx32, const uint8_t *vals, uint8_t root_out[32]);
int32_t bp_vsmt2_get_batch_wide(const bp_vsmt2 *t, uint32_t count, const uint8_t *idx32, uint8_t *leaves, uint8_t *proofs);
int32_t bp_vsmt2_witness_batch_wide(const bp_vsmt2 *t, uint32_t count, const uint8_t *idx32, uint8_t *v, uint8_t *pub);
int32_t bp_vsmt2_witness_batch_wide_device(const bp_vsmt2 *t, uint32_t count, const uint8_t *d_idx32, uint8_t *d_v, uint8_t *d_pub, void *stream);
/* Poseidon_hash_2 (src/gadget_poseidon.rs:428-443) of count independent pairs on the device; host buffers [count][32] */
int32_t bp_poseidon_hash_2_batch(const bp_poseidon_params *p, int32_t sbox, uint32_t count, const uint8_t *xl, const uint8_t *xr, uint8_t *out);

/* ---- MSM microbenchmark entry (BASELINE.json config 3) -------------------------------------------
 * result = sum_i scalars[i] * points[i] over ristretto255; points are the first n generators of chain G.
 * d_scalars: device, [n][32] canonical LE.  out: device, 32 bytes. */
int32_t bp_msm_gens_device(const bp_gens *g, uint32_t n, const uint8_t *d_scalars, uint8_t *d_out, void *stream);

/* ---- device self-tests of single primitives (used by tests/ to pin the CUDA arithmetic to the oracle) ----
 * which: 0 Merlin transcript KAT, 1 scalar invert, 2 wide scalar reduction, 3 ristretto decode/encode round trip,
 *        4 scalar multiply, 5 ristretto one-way map, 6 transcript RNG draws, 7 Keccak-f[1600].  Host buffers. */
int32_t bp_selftest_device(int32_t which, const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len);

#ifdef __cplusplus
}
#endif
#endif
