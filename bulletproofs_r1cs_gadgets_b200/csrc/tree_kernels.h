// Device-side sparse Merkle tree: the batched analogue of VanillaSparseMerkleTree::{new,update,get}
// (reference src/gadget_vsmt_2.rs:36-131).  The reference keeps a content-addressed HashMap hash -> (left, right) and walks
// it one key at a time, 253 Poseidon hashes per update.  Here a node is addressed by its POSITION: heap key
// 2^(depth-level) + (index >> level) (root = 1, leaves = 2^depth + index) -- a 256-bit integer, since the reference's tree is
// 253 levels deep (TreeDepth, src/gadget_vsmt_2.rs:23) -- in one open-addressing table in HBM (8-byte tag + 32-byte key +
// 32-byte Montgomery scalar per slot); absent nodes are the empty-subtree hashes of their level.  A batch of
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
// 256-bit heap keys (little-endian words; bit 255 is never set: keys are below 2^254)
struct tkey { uint64_t w[4]; };
HD tkey tkey_leaf(const uint64_t idx[4], int depth) {  // 2^depth + idx
  tkey k; for (int i = 0; i < 4; i++) k.w[i] = idx[i];
  k.w[depth >> 6] |= 1ull << (depth & 63);
  return k;
}
HD tkey tkey_shr(const tkey &a, int n) {
  tkey r; const int ws = n >> 6, bs = n & 63;
  for (int i = 0; i < 4; i++) {
    uint64_t v = i + ws < 4 ? a.w[i + ws] >> bs : 0;
    if (bs && i + ws + 1 < 4) v |= a.w[i + ws + 1] << (64 - bs);
    r.w[i] = v;
  }
  return r;
}
HD tkey tkey_child(const tkey &a, int right) {  // 2a + right
  tkey r; r.w[3] = (a.w[3] << 1) | (a.w[2] >> 63); r.w[2] = (a.w[2] << 1) | (a.w[1] >> 63); r.w[1] = (a.w[1] << 1) | (a.w[0] >> 63);
  r.w[0] = (a.w[0] << 1) | (uint64_t)right;
  return r;
}
HD bool tkey_eq(const tkey &a, const tkey &b) { return a.w[0] == b.w[0] && a.w[1] == b.w[1] && a.w[2] == b.w[2] && a.w[3] == b.w[3]; }
HD int tkey_bit(const uint64_t idx[4], int i) { return (int)((idx[i >> 6] >> (i & 63)) & 1); }
HD uint64_t mix64(uint64_t k) {  // splitmix64 finaliser
  k ^= k >> 30; k *= 0xbf58476d1ce4e5b9ull; k ^= k >> 27; k *= 0x94d049bb133111ebull; k ^= k >> 31;
  return k;
}
HD uint64_t tkey_tag(const tkey &k) { return mix64(k.w[0] ^ mix64(k.w[1] ^ mix64(k.w[2] ^ mix64(k.w[3] + 0x9e3779b97f4a7c15ull)))) | 1ull; }  // never 0
// Open addressing, linear probing, capacity a power of two.  A slot is claimed by a compare-and-swap on its 64-bit TAG (a hash of
// the key, 0 = empty); the full key is stored beside it, its last word published LAST with TKEY_VALID set, so a reader that
// meets an equal tag waits for the key before comparing (two different keys with equal tags are a 2^-64 event, handled).
#define TKEY_VALID (1ull << 63)
struct TreeTable { uint64_t *tags; tkey *keys; scm *vals; uint64_t mask; };
HD bool tree_slot_holds(const TreeTable &t, uint64_t s, const tkey &key) {
#if defined(__CUDA_ARCH__)
  volatile uint64_t *kw = t.keys[s].w;
  uint64_t w3;
  while (!((w3 = kw[3]) & TKEY_VALID)) {}
  return kw[0] == key.w[0] && kw[1] == key.w[1] && kw[2] == key.w[2] && (w3 & ~TKEY_VALID) == key.w[3];
#else
  const tkey &k = t.keys[s];
  return k.w[0] == key.w[0] && k.w[1] == key.w[1] && k.w[2] == key.w[2] && (k.w[3] & ~TKEY_VALID) == key.w[3];
#endif
}
HD bool tree_lookup(const TreeTable &t, const tkey &key, scm &out) {
  const uint64_t tag = tkey_tag(key);
  for (uint64_t s = tag & t.mask;; s = (s + 1) & t.mask) {
    const uint64_t k = t.tags[s];
    if (k == 0) return false;
    if (k == tag && tree_slot_holds(t, s, key)) { load_struct(out, &t.vals[s]); return true; }
  }
}
// returns 1 when the key was new.  Keys of one launch are distinct, so the value store needs no ordering.
HD int tree_insert(const TreeTable &t, const tkey &key, const scm &v) {
  const uint64_t tag = tkey_tag(key);
  for (uint64_t s = tag & t.mask;; s = (s + 1) & t.mask) {
#if defined(__CUDA_ARCH__)
    const uint64_t old = atomicCAS((unsigned long long *)&t.tags[s], 0ull, (unsigned long long)tag);
#else
    const uint64_t old = t.tags[s];
    if (old == 0) t.tags[s] = tag;
#endif
    if (old == 0) {
      t.keys[s].w[0] = key.w[0]; t.keys[s].w[1] = key.w[1]; t.keys[s].w[2] = key.w[2];
#if defined(__CUDA_ARCH__)
      __threadfence();
      *(volatile uint64_t *)&t.keys[s].w[3] = key.w[3] | TKEY_VALID;
#else
      t.keys[s].w[3] = key.w[3] | TKEY_VALID;
#endif
      store_struct(&t.vals[s], v);
      return 1;
    }
    if (old == tag && tree_slot_holds(t, s, key)) { store_struct(&t.vals[s], v); return 0; }
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
  const tkey *pk; const int32_t *li, *ri; const scm *child_vals; const scm *empty_child; scm *out;
  HD void operator()(long j) const {
    const tkey key = pk[j];
    scm l, r;
    if (li[j] >= 0) load_struct(l, &child_vals[li[j]]); else if (!tree_lookup(t, tkey_child(key, 0), l)) l = *empty_child;
    if (ri[j] >= 0) load_struct(r, &child_vals[ri[j]]); else if (!tree_lookup(t, tkey_child(key, 1), r)) r = *empty_child;
    scm h = poseidon_hash2_native(pos, l, r, sbox);
    store_struct(&out[j], h);
  }
};
struct KTreeInsert {
  static constexpr int kBlock = 128, kMinBlocks = 1;
  static constexpr const char *kName = "KTreeInsert";
  TreeTable t; const tkey *keys; const scm *vals; unsigned long long *count;
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
    if (from.tags[s] == 0) return;
    tkey k = from.keys[s]; k.w[3] &= ~TKEY_VALID;
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
  TreeTable t; const scm *empty; const uint64_t *idx; int depth, order; uint8_t *leaves, *proofs;  // idx: 4 little-endian words per query
  HD void operator()(long tid) const {
    const long q = tid / (depth + 1); const int l = (int)(tid % (depth + 1));
    const tkey leaf_key = tkey_leaf(idx + 4 * q, depth);
    scm v;
    if (l == depth) {
      if (!tree_lookup(t, leaf_key, v)) v = empty[0];
      sc_tobytes(leaves + 32 * q, v);
      return;
    }
    tkey sib = tkey_shr(leaf_key, l); sib.w[0] ^= 1;
    if (!tree_lookup(t, sib, v)) v = empty[l];
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
    const tkey leaf_key = tkey_leaf(idx + 4 * q, depth);
    if (j == m) { if (pub) sc_tobytes(pub + 32 * q, *root); return; }
    uint8_t *out = v + 32 * (q * m + j);
    scm val;
    if (j == 0) { if (!tree_lookup(t, leaf_key, val)) val = empty[0]; }
    else if (j <= depth) val = tkey_bit(idx + 4 * q, j - 1) ? sc_one() : sc_zero();
    else if (j <= 2 * depth) { const int l = j - depth - 1; tkey sib = tkey_shr(leaf_key, l); sib.w[0] ^= 1; if (!tree_lookup(t, sib, val)) val = empty[l]; }
    else val = j == 2 * depth + 2 ? sc_from_u64(101) : sc_zero();
    sc_tobytes(out, val);
  }
};
struct KTreeWidenIndices {  // 64-bit indices -> 4-word indices
  static constexpr int kBlock = 128, kMinBlocks = 1;
  static constexpr const char *kName = "KTreeWidenIndices";
  const uint64_t *in; uint64_t *out;
  HD void operator()(long i) const { out[4 * i] = in[i]; out[4 * i + 1] = 0; out[4 * i + 2] = 0; out[4 * i + 3] = 0; }
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
