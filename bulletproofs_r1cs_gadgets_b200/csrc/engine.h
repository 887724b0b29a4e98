// Host-side engine: device generator tables, compiled circuits, and the batched prover/verifier
// pipelines that string the kernels of kernels.h together on one CUDA stream per engine call.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <vector>
#include "kernel_groups.h"

#include "../../include/bp_b200.h"  // status codes

struct BpGens {
  uint32_t capacity;
  ge_p3 *G_p3, *H_p3;
  ge_niels *G_n, *H_n;
  ge_p3 *pc;            // B, B_blinding
  ge_niels *pc_niels;
  ge_niels *pc_table;   // [2][64][16]
  ge_niels *sg;         // shift table [(2cap+2) + SG_SPARE][SB_WINDOWS]: 2^(15w) * P (sorted-bucket MSM); NULL when disabled
  long pad_n[8], pad_N[8]; int pad_count;
  long merge_slots; uint64_t merge_owner;  // slots after the spare ones hold the generator sums of ONE circuit at a time (KMergeGens)  // spare shift-table slots holding sum_{i=n-N/2}^{N/2-1} H_i of the circuits seen (see engine.cu)
  ge_niels *table;      // fixed-base tables [(2cap+2)][32][128] (see KTableBuild); NULL when disabled
  // fold tables: what KFoldTable materialises the level-J generators from when the 8-bit direct tables do not exist (capacities
  // above ~65k generators).  Same layout with narrower windows, [2][ft_cap][ft_w][2^(ft_bits-1)], built on first use for the
  // first ft_cap generators of each chain (ensure_fold_table); NULL until then
  ge_niels *ftable; int ft_bits; long ft_cap;
  uint8_t pc_c[64];     // compressed B, B_blinding (host copy)
  struct Workspace *msm_ws; uint32_t msm_ws_n;  // scratch of the MSM microbenchmark entry
};

// Poseidon block ops of a witness tape (host side): blocks + the parameters they share
struct HostPoseidonTape { const PoseidonBlock *blocks; uint32_t nblocks; const scm *round_keys; uint32_t nkeys; const scm *mds; uint32_t full_b, partial, full_e; };
struct HostTerm { uint8_t kind; uint32_t idx; uint8_t coeff[32]; };  // kind: 0 committed,1 left,2 right,3 output,4 one

struct Workspace;
struct BpCircuit {
  uint32_t n, N, k, m, q, nnz, nslots, naux, npub;
  uint32_t *d_slot_ptr, *d_tq; scm *d_tcoeff;
  uint32_t nlong, nparts; uint32_t *d_part_beg, *d_part_end, *d_long_slot, *d_long_first;  // slots cut into parts for KFlattenParts (see kernels.h)
  int has_tape;
  TapeOp *d_tape; uint32_t *d_wptr; uint8_t *d_wkind; uint32_t *d_widx; scm *d_wcoeff;
  PoseidonBlock *d_pblocks; scm *d_pos_rk, *d_pos_mds; PoseidonDev pos;
  uint32_t nfixed; uint32_t *d_fixed_idx; uint8_t *d_fixed_V;  // commitments the circuit fixes (statics): slot indices, expected bytes
  int wit_uses_pub;                  // the witness program reads public inputs
  uint64_t serial;                   // unique per circuit_create (a freed circuit's address may be reused)
  std::vector<uint32_t> *merge_src;  // host: groups of A_I rows with equal scalars, 3 generator indices each (G index, or n + H index; ~0 unused)
  Workspace *ws;
};

int bp_device_init();
int gens_create(uint32_t capacity, BpGens **out);
void gens_free(BpGens *g);
int gens_export(const BpGens *g, int which, uint32_t count, uint8_t *out);  // which: 0 G, 1 H -> compressed

// cons_ptr[q+1], terms in constraint order, coefficients in Montgomery form (host).  tape (optional, n entries) + witness LCs.
int circuit_create(uint32_t n, uint32_t m, uint32_t npub, uint32_t q, const uint32_t *cons_ptr, const uint8_t *kind, const uint32_t *idx,
                   const scm *coeff, const TapeOp *tape, uint32_t naux, uint32_t nwlc, const uint32_t *wlc_ptr,
                   const uint8_t *wkind, const uint32_t *widx, const scm *wcoeff, const struct HostPoseidonTape *ptape, BpCircuit **out);
// commitment slots whose compressed value is fixed by the circuit (idx[n], V[n][32], host pointers); checked by the verifiers
int circuit_set_fixed_commitments(BpCircuit *c, uint32_t n, const uint32_t *idx, const uint8_t *V);
void circuit_free(BpCircuit *c);
int circuit_release_workspace(BpCircuit *c);
size_t circuit_proof_len(const BpCircuit *c);
double engine_workspace_bytes_per_proof(const BpCircuit *c);

struct ProveArgs {
  int B;
  const uint8_t *label; int label_len;
  // all pointers are DEVICE pointers, layouts [B][count][32]
  const uint8_t *v, *vbl, *entropy;
  const uint8_t *aL, *aR, *aO;   // explicit witness, or all NULL to run the circuit's witness tape
  const uint8_t *aux;            // [B][naux][32] tape inputs
  const uint8_t *pub;            // [B][npub][32] public inputs (only read by the witness tape on the prover side)
  uint8_t *V_out;                // [B][m][32]
  uint8_t *proofs;               // [B][proof_len]
  int *status;                   // [B]
};
// proves a.B statements; `chunk` = proofs per device chunk (<= 0: one chunk)
int engine_prove(const BpGens *g, BpCircuit *c, const ProveArgs &a, int chunk, dev_stream s);
// the same in two calls, two batches (slot 0 / 1) in flight per circuit: begin = phase A on internal streams, finish = phase B on s
int engine_prove_begin(const BpGens *g, BpCircuit *c, int slot, const ProveArgs &a, int chunk, dev_stream s);
int engine_prove_finish(const BpGens *g, BpCircuit *c, int slot, dev_stream s);

struct VerifyArgs {
  int B;
  const uint8_t *label; int label_len;
  const uint8_t *V;        // [B][m][32] device
  const uint8_t *proofs;   // [B][proof_len] device
  const uint8_t *entropy;  // [B][32] device
  const uint8_t *pub;      // [B][npub][32] device public inputs
  int *status;             // [B] device
};
int engine_verify(const BpGens *g, BpCircuit *c, const VerifyArgs &a, dev_stream s);
// cross-proof batched verification: per-proof structural status + one combined verdict (device int)
int engine_verify_combined(BpGens *g, BpCircuit *c, const VerifyArgs &a, int *d_combined, dev_stream s);

// one-off helper: commitments a*B + b*B_blinding for host scalars
int engine_commit(const BpGens *g, int count, const uint8_t *v, const uint8_t *r, uint8_t *out);
long engine_launch_count();
void engine_profile_enable(int on);
int engine_profile_report(char *buf, size_t cap);
// sum_i scalars[i] * G[i] over the first n generators; d_scalars [n][32] canonical bytes, d_out 32 bytes (device)
int engine_msm_gens(BpGens *g, uint32_t n, const uint8_t *d_scalars, uint8_t *d_out, dev_stream s);
