// Device-side sparse Merkle tree: the batched analogue of VanillaSparseMerkleTree::{new,update,get}
// (reference src/gadget_vsmt_2.rs:36-131).  The reference keeps a content-addressed HashMap hash -> (left, right) and walks
// it one key at a time, 253 Poseidon hashes per update.  Here a node is addressed by its POSITION: heap key
// 2^(depth-level) + (index >> level) (root = 1, leaves = 2^depth + index) in one open-addressing table in HBM
// (8-byte key + 32-byte Montgomery scalar per slot); absent nodes are the empty-subtree hashes of their level.  A batch of
// updates is hashed level by level, one thread per DISTINCT parent, so shared ancestors are hashed once; a batch of
// lookups is one thread per (query, level) because every sibling's key is known from the index alone.
#pragma once
#include "kernels.h"

// ------------------------------------------------------------------------------------------------ native Poseidon
// Poseidon_permutation (reference src/gadget_poseidon.rs:189-280): add round keys to every lane, S-box on all lanes (full
// rounds) or on lane width-1 (partial rounds, :239), then state' = M * state (:217-221).  sbox: 0 cube, 1 inverse
// (Scalar::invert, 0 -> 0).  The six inversions of a full round share one field inversion (Montgomery's trick).
// The loops are kept rolled on the device (POSEIDON_ROLLED, kernels.h: one multiplication site each): fully unrolled, the
// 36 + 17 inlined Montgomery multiplications of a round made a 12 600-instruction kernel (200 KB), more than the instruction
// cache holds; rolled it is 5 200 instructions and 19 % faster (4.8 M hashes/s).
HD void poseidon_permute_native(const PoseidonDev &pos, scm st[POSEIDON_WIDTH], int sbox) {
  uint32_t off = 0;
  const uint32_t total = pos.full_b + pos.partial + pos.full_e;
  POSEIDON_ROLLED
  for (uint32_t rnd = 0; rnd < total; rnd++) {
    const bool full = rnd < pos.full_b || rnd >= pos.full_b + pos.partial;
    const int first = full ? 0 : POSEIDON_WIDTH - 1;
    for (int i = 0; i < POSEIDON_WIDTH; i++) st[i] = sc_add(st[i], pos.round_keys[off + i]);
    off += POSEIDON_WIDTH;
    if (sbox == 0) {
      POSEIDON_ROLLED
      for (int i = first; i < POSEIDON_WIDTH; i++) st[i] = sc_mul(sc_sqr(st[i]), st[i]);
    } else {
      // one inversion per round: of the product of the non-zero lanes (full) or of the last lane (partial)
      scm pre[POSEIDON_WIDTH], acc = sc_one();
      POSEIDON_ROLLED
      for (int i = first; i < POSEIDON_WIDTH; i++) { pre[i] = acc; if (!sc_is_zero(st[i])) acc = sc_mul(acc, st[i]); }
      scm inv = sc_invert(acc);
      POSEIDON_ROLLED
      for (int i = POSEIDON_WIDTH - 1; i >= first; i--) {
        if (sc_is_zero(st[i])) continue;
        const scm x = st[i];
        st[i] = sc_mul(inv, pre[i]); inv = sc_mul(inv, x);
      }
    }
    scm nx[POSEIDON_WIDTH];
    POSEIDON_ROLLED
    for (int i = 0; i < POSEIDON_WIDTH; i++) {
      scm acc = sc_zero();
      POSEIDON_ROLLED
      for (int j = 0; j < POSEIDON_WIDTH; j++) acc = sc_add(acc, sc_mul(st[j], pos.mds[i * POSEIDON_WIDTH + j]));
      nx[i] = acc;
    }
    for (int i = 0; i < POSEIDON_WIDTH; i++) st[i] = nx[i];
  }
}
// Poseidon_hash_2 (reference src/gadget_poseidon.rs:428-443): perm([0, xl, xr, 101, 0, 0])[1]
HD scm poseidon_hash2_native(const PoseidonDev &pos, const scm &xl, const scm &xr, int sbox) {
  scm st[POSEIDON_WIDTH];
  st[0] = sc_zero(); st[1] = xl; st[2] = xr; st[3] = sc_from_u64(101); st[4] = sc_zero(); st[5] = sc_zero();
  poseidon_permute_native(pos, st, sbox);
  return st[1];
}

// ------------------------------------------------------------------------------------------------ node table
// open addressing, linear probing; key 0 = empty slot (heap keys are >= 1), capacity a power of two
struct TreeTable { uint64_t *keys; scm *vals; uint64_t mask; };
HD uint64_t tree_slot(uint64_t k, uint64_t mask) {  // splitmix64 finaliser
  k ^= k >> 30; k *= 0xbf58476d1ce4e5b9ull; k ^= k >> 27; k *= 0x94d049bb133111ebull; k ^= k >> 31;
  return k & mask;
}
HD bool tree_lookup(const TreeTable &t, uint64_t key, scm &out) {
  for (uint64_t s = tree_slot(key, t.mask);; s = (s + 1) & t.mask) {
    uint64_t k = t.keys[s];
    if (k == key) { load_struct(out, &t.vals[s]); return true; }
    if (k == 0) return false;
  }
}
// returns 1 when the key was new.  Keys of one launch are distinct, so the value store needs no ordering.
HD int tree_insert(const TreeTable &t, uint64_t key, const scm &v) {
  for (uint64_t s = tree_slot(key, t.mask);; s = (s + 1) & t.mask) {
#if defined(__CUDA_ARCH__)
    uint64_t old = atomicCAS((unsigned long long *)&t.keys[s], 0ull, (unsigned long long)key);
#else
    uint64_t old = t.keys[s];
    if (old == 0) t.keys[s] = key;
#endif
    if (old == 0 || old == key) { store_struct(&t.vals[s], v); return old == 0; }
  }
}
HD void tree_count_add(unsigned long long *c, unsigned long long n) {
#if defined(__CUDA_ARCH__)
  atomicAdd(c, n);
#else
  *c += n;
#endif
}

// ------------------------------------------------------------------------------------------------ kernels
struct KTreeEmptyChain {  // VanillaSparseMerkleTree::new, reference src/gadget_vsmt_2.rs:41-50: empty[i] = H(empty[i-1], empty[i-1])
  static constexpr int kBlock = 32, kMinBlocks = 1;
  static constexpr const char *kName = "KTreeEmptyChain";
  PoseidonDev pos; int sbox, depth; scm *empty; scm *root;
  HD void operator()(long) const {
    scm cur = sc_zero(); empty[0] = cur;
    for (int i = 1; i <= depth; i++) { cur = poseidon_hash2_native(pos, cur, cur, sbox); empty[i] = cur; }
    *root = cur;
  }
};
struct KTreeLoadLeaves {  // canonical (or any 256-bit) little-endian bytes -> Montgomery scalars, reduced mod l
  static constexpr int kBlock = 128, kMinBlocks = 1;
  static constexpr const char *kName = "KTreeLoadLeaves";
  const uint8_t *bytes; scm *out;
  HD void operator()(long i) const { out[i] = sc_from_bytes_mod_order(bytes + 32 * i); }
};
// One level of a batched update (reference src/gadget_vsmt_2.rs:72-93, for every key of the batch at once): thread j owns
// parent pk[j]; a child is either a node this batch changed (index into the level below) or the stored / empty node.
struct KTreeHashLevel {
  static constexpr int kBlock = 64, kMinBlocks = 1;
  static constexpr const char *kName = "KTreeHashLevel";
  TreeTable t; PoseidonDev pos; int sbox;
  const uint64_t *pk; const int32_t *li, *ri; const scm *child_vals; const scm *empty_child; scm *out;
  HD void operator()(long j) const {
    const uint64_t key = pk[j];
    scm l, r;
    if (li[j] >= 0) load_struct(l, &child_vals[li[j]]); else if (!tree_lookup(t, 2 * key, l)) l = *empty_child;
    if (ri[j] >= 0) load_struct(r, &child_vals[ri[j]]); else if (!tree_lookup(t, 2 * key + 1, r)) r = *empty_child;
    scm h = poseidon_hash2_native(pos, l, r, sbox);
    store_struct(&out[j], h);
  }
};
struct KTreeInsert {
  static constexpr int kBlock = 128, kMinBlocks = 1;
  static constexpr const char *kName = "KTreeInsert";
  TreeTable t; const uint64_t *keys; const scm *vals; unsigned long long *count;
  HD void operator()(long i) const {
    scm v; load_struct(v, &vals[i]);
    if (tree_insert(t, keys[i], v)) tree_count_add(count, 1);
  }
};
struct KTreeRehash {  // thread per slot of the old table
  static constexpr int kBlock = 128, kMinBlocks = 1;
  static constexpr const char *kName = "KTreeRehash";
  TreeTable from, to;
  HD void operator()(long s) const {
    uint64_t k = from.keys[s];
    if (k == 0) return;
    scm v; load_struct(v, &from.vals[s]);
    tree_insert(to, k, v);
  }
};
// VanillaSparseMerkleTree::get for a batch (reference src/gadget_vsmt_2.rs:101-131).  Thread (q, l): l < depth fetches the
// sibling at level l (level 0 = leaves), l == depth the leaf itself.  order 0: siblings root -> leaf as `get` returns them;
// order 1: leaf level first, the commit order of the membership circuit (src/gadget_vsmt_2.rs:319).
struct KTreeGet {
  static constexpr int kBlock = 128, kMinBlocks = 1;
  static constexpr const char *kName = "KTreeGet";
  TreeTable t; const scm *empty; const uint64_t *idx; int depth, order; uint8_t *leaves, *proofs;
  HD void operator()(long tid) const {
    const long q = tid / (depth + 1); const int l = (int)(tid % (depth + 1));
    const uint64_t leaf_key = (1ull << depth) + idx[q];
    scm v;
    if (l == depth) {
      if (!tree_lookup(t, leaf_key, v)) v = empty[0];
      sc_tobytes(leaves + 32 * q, v);
      return;
    }
    if (!tree_lookup(t, (leaf_key >> l) ^ 1, v)) v = empty[l];
    const int pos_ = order == 0 ? depth - 1 - l : l;
    sc_tobytes(proofs + 32 * (q * depth + pos_), v);
  }
};
// Rows of the membership circuit's committed values for a batch of leaves, in the commit order of the reference's prover
// (src/gadget_vsmt_2.rs:296-330): leaf, depth index bits LSB first, depth siblings leaf level first, statics 0, 101, 0, 0
// (src/gadget_poseidon.rs:554-578) -- written straight into the [B][m][32] input of bp_prove_batch_device; row m of a
// thread group writes the proof's public input (the root) when `pub` is given.
struct KTreeWitnessRows {
  static constexpr int kBlock = 128, kMinBlocks = 1;
  static constexpr const char *kName = "KTreeWitnessRows";
  TreeTable t; const scm *empty; const scm *root; const uint64_t *idx; int depth; uint8_t *v; uint8_t *pub;
  HD void operator()(long tid) const {
    const int m = 2 * depth + 5;
    const long q = tid / (m + 1); const int j = (int)(tid % (m + 1));
    const uint64_t leaf_key = (1ull << depth) + idx[q];
    if (j == m) { if (pub) sc_tobytes(pub + 32 * q, *root); return; }
    uint8_t *out = v + 32 * (q * m + j);
    scm val;
    if (j == 0) { if (!tree_lookup(t, leaf_key, val)) val = empty[0]; }
    else if (j <= depth) val = ((idx[q] >> (j - 1)) & 1) ? sc_one() : sc_zero();
    else if (j <= 2 * depth) { const int l = j - depth - 1; if (!tree_lookup(t, (leaf_key >> l) ^ 1, val)) val = empty[l]; }
    else val = j == 2 * depth + 2 ? sc_from_u64(101) : sc_zero();
    sc_tobytes(out, val);
  }
};
// Poseidon_hash_2 of count independent pairs (tree building blocks; also the throughput probe of the hash itself)
struct KPoseidonHash2Batch {
  static constexpr int kBlock = 64, kMinBlocks = 1;
  static constexpr const char *kName = "KPoseidonHash2Batch";
  PoseidonDev pos; int sbox; const uint8_t *xl, *xr; uint8_t *out;
  HD void operator()(long i) const {
    scm h = poseidon_hash2_native(pos, sc_from_bytes_mod_order(xl + 32 * i), sc_from_bytes_mod_order(xr + 32 * i), sbox);
    sc_tobytes(out + 32 * i, h);
  }
};
