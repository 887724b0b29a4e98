// Device-side binary sparse Merkle tree behind the C-ABI (include/bp_b200.h, bp_vsmt2_*): batched
// VanillaSparseMerkleTree::{new,update,get} of reference src/gadget_vsmt_2.rs:36-131 plus the witness rows of the
// membership circuit (src/gadget_vsmt_2.rs:296-330) written straight into bp_prove_batch_device's input layout.
// Kernels and data layout: tree_kernels.h.  The host side below only orders keys and plans the levels of a batch
// (which distinct parents exist, which of their children the batch changed); every hash runs on the device.
#include "recorder.h"
#include "tree_kernels.h"
#include <algorithm>
#include <array>
#include <numeric>

struct bp_vsmt2 {
  uint32_t depth = 0; int sbox = 1;
  PoseidonDev pos{}; scm *d_rk = nullptr, *d_mds = nullptr;
  scm *d_empty = nullptr;   // depth + 1 empty-subtree hashes, [0] = 0 (leaf level)
  scm *d_root = nullptr;
  unsigned long long *d_count = nullptr;
  TreeTable t{nullptr, nullptr, nullptr, 0};
  uint64_t cap = 0, nodes = 0;
  uint8_t root_bytes[32];
  uint64_t *d_widen = nullptr; uint32_t widen_cap = 0;  // scratch of the 64-bit-index device entry point
};

namespace {
template <class T> int dalloc(T **p, size_t count) { return dev_malloc((void **)p, count * sizeof(T)); }
template <class T> struct Dev {  // scoped device array
  T *p = nullptr;
  int alloc(size_t n) { return dalloc(&p, n); }
  ~Dev() { dev_free(p); }
};
dev_stream as_stream(void *s) {
#ifndef BP_HOST_EMUL
  return (dev_stream)s;
#else
  (void)s; return 0;
#endif
}
int upload_poseidon(const bp_poseidon_params *p, PoseidonDev &pos, scm **d_rk, scm **d_mds, dev_stream s) {
  std::vector<scm> mds(POSEIDON_WIDTH * POSEIDON_WIDTH);
  for (int i = 0; i < POSEIDON_WIDTH; i++) for (int j = 0; j < POSEIDON_WIDTH; j++) mds[i * POSEIDON_WIDTH + j] = p->mds[i][j];
  if (dalloc(d_rk, p->round_keys.size()) || dalloc(d_mds, mds.size())) return BP_ERR_OOM;
  if (dev_h2d(*d_rk, p->round_keys.data(), p->round_keys.size() * sizeof(scm), s) || dev_h2d(*d_mds, mds.data(), mds.size() * sizeof(scm), s) || dev_sync(s))
    return BP_ERR_CUDA;
  pos.round_keys = *d_rk; pos.mds = *d_mds;
  pos.full_b = p->full_rounds_beginning; pos.partial = p->partial_rounds; pos.full_e = p->full_rounds_end;
  return BP_OK;
}
int table_alloc(TreeTable &t, uint64_t cap, dev_stream s) {
  if (dalloc(&t.tags, cap) || dalloc(&t.keys, cap) || dalloc(&t.vals, cap)) {
    dev_free(t.tags); dev_free(t.keys); dev_free(t.vals); t.tags = nullptr; t.keys = nullptr; t.vals = nullptr; return BP_ERR_OOM;
  }
  t.mask = cap - 1;
  return (dev_memset(t.tags, 0, cap * sizeof(uint64_t), s) || dev_memset(t.keys, 0, cap * sizeof(tkey), s)) ? BP_ERR_CUDA : BP_OK;
}
// host-side 256-bit key helpers (little-endian words)
typedef std::array<uint64_t, 4> hkey;
bool hkey_less(const hkey &a, const hkey &b) { for (int i = 3; i >= 0; i--) if (a[i] != b[i]) return a[i] < b[i]; return false; }
hkey hkey_shr1(const hkey &a) { return hkey{(a[0] >> 1) | (a[1] << 63), (a[1] >> 1) | (a[2] << 63), (a[2] >> 1) | (a[3] << 63), a[3] >> 1}; }
bool idx_in_range(const uint64_t *w, uint32_t depth) {  // idx < 2^depth
  for (uint32_t b = depth; b < 256; b += 64 - (b & 63)) { const uint64_t hi = w[b >> 6] >> (b & 63); if (hi) return false; }
  return true;
}
void words_from_bytes(std::vector<uint64_t> &out, const uint8_t *idx32, uint32_t count) {
  out.resize((size_t)count * 4);
  for (size_t i = 0; i < (size_t)count * 4; i++) { uint64_t x = 0; for (int j = 7; j >= 0; j--) x = (x << 8) | idx32[8 * i + j]; out[i] = x; }
}
void words_from_u64(std::vector<uint64_t> &out, const uint64_t *idx, uint32_t count) {
  out.assign((size_t)count * 4, 0);
  for (uint32_t i = 0; i < count; i++) out[4 * (size_t)i] = idx[i];
}
// keeps the load factor at or below 1/2 after `extra` more insertions
int table_reserve(bp_vsmt2 *T, uint64_t extra, dev_stream s) {
  const uint64_t need = 2 * (T->nodes + extra);
  if (need <= T->cap) return BP_OK;
  uint64_t cap = T->cap ? T->cap : 1024;
  while (cap < 2 * need) cap *= 2;  // grow to a quarter full so a stream of batches rehashes rarely
  TreeTable nt{nullptr, nullptr, nullptr, 0};
  int rc = table_alloc(nt, cap, s);
  if (rc) return rc;
  if (T->cap && launch((long)T->cap, s, KTreeRehash{T->t, nt})) return BP_ERR_CUDA;
  if (dev_sync(s)) return BP_ERR_CUDA;
  dev_free(T->t.tags); dev_free(T->t.keys); dev_free(T->t.vals);
  T->t = nt; T->cap = cap;
  return BP_OK;
}
int fetch_root(bp_vsmt2 *T, dev_stream s) {
  scm r; unsigned long long n = 0;
  if (dev_d2h(&r, T->d_root, sizeof(scm), s) || dev_d2h(&n, T->d_count, sizeof(n), s) || dev_sync(s)) return BP_ERR_CUDA;
  sc_tobytes(T->root_bytes, r); T->nodes = n;
  return BP_OK;
}
}  // namespace

extern "C" {

int32_t bp_vsmt2_new(const bp_poseidon_params *p, uint32_t depth, int32_t sbox, bp_vsmt2 **out) {
  if (!p || !out || depth < 1 || depth > 253 || p->width != POSEIDON_WIDTH || (sbox != BP_SBOX_CUBE && sbox != BP_SBOX_INVERSE)) return BP_ERR_INVALID_ARGUMENT;
  int rc = bp_device_init();
  if (rc) return rc;
  dev_stream s = 0;
  bp_vsmt2 *T = new bp_vsmt2();
  T->depth = depth; T->sbox = sbox;
  rc = upload_poseidon(p, T->pos, &T->d_rk, &T->d_mds, s);
  if (!rc && (dalloc(&T->d_empty, depth + 1) || dalloc(&T->d_root, 1) || dalloc(&T->d_count, 1))) rc = BP_ERR_OOM;
  if (!rc && dev_memset(T->d_count, 0, sizeof(unsigned long long), s)) rc = BP_ERR_CUDA;
  if (!rc) rc = table_reserve(T, 1 << 14, s);
  if (!rc && launch(1, s, KTreeEmptyChain{T->pos, T->sbox, (int)depth, T->d_empty, T->d_root})) rc = BP_ERR_CUDA;
  if (!rc) rc = fetch_root(T, s);
  if (rc) { bp_vsmt2_free(T); return rc; }
  *out = T;
  return BP_OK;
}
void bp_vsmt2_free(bp_vsmt2 *T) {
  if (!T) return;
  void *ps[] = {T->d_rk, T->d_mds, T->d_empty, T->d_root, T->d_count, T->t.tags, T->t.keys, T->t.vals, T->d_widen};
  for (void *p : ps) dev_free(p);
  delete T;
}
uint32_t bp_vsmt2_depth(const bp_vsmt2 *T) { return T ? T->depth : 0; }
uint64_t bp_vsmt2_num_nodes(const bp_vsmt2 *T) { return T ? T->nodes : 0; }
int32_t bp_vsmt2_root(const bp_vsmt2 *T, uint8_t root[32]) {
  if (!T || !root) return BP_ERR_INVALID_ARGUMENT;
  memcpy(root, T->root_bytes, 32);
  return BP_OK;
}
int32_t bp_vsmt2_empty_hashes(const bp_vsmt2 *T, uint8_t *out) {
  if (!T || !out) return BP_ERR_INVALID_ARGUMENT;
  std::vector<scm> e(T->depth + 1);
  dev_stream s = 0;
  if (dev_d2h(e.data(), T->d_empty, e.size() * sizeof(scm), s) || dev_sync(s)) return BP_ERR_CUDA;
  for (size_t i = 0; i < e.size(); i++) sc_tobytes(out + 32 * i, e[i]);
  return BP_OK;
}

// Applies count updates (idx[i] -> vals[i]); a key that occurs more than once keeps its LAST value, which is what count
// sequential VanillaSparseMerkleTree::update calls leave behind.  root_out (optional) receives the new root.
// Indices are 256-bit: iw = 4 little-endian words per key (the reference's keys are Scalars, depth 253).
static int32_t update_batch_words(bp_vsmt2 *T, uint32_t count, const uint64_t *iw, const uint8_t *vals, uint8_t root_out[32]) {
  const uint32_t depth = T->depth;
  for (uint32_t i = 0; i < count; i++) if (!idx_in_range(iw + 4 * (size_t)i, depth)) return BP_ERR_INVALID_ARGUMENT;
  if (count == 0) { if (root_out) memcpy(root_out, T->root_bytes, 32); return BP_OK; }
  dev_stream s = 0;
  auto key_of = [&](uint32_t i) { hkey k{iw[4 * (size_t)i], iw[4 * (size_t)i + 1], iw[4 * (size_t)i + 2], iw[4 * (size_t)i + 3]}; return k; };
  // level 0: distinct keys in ascending order, last write wins
  std::vector<uint32_t> order(count);
  std::iota(order.begin(), order.end(), 0u);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return hkey_less(key_of(a), key_of(b)); });
  std::vector<hkey> keys; std::vector<uint8_t> leaf_bytes;
  keys.reserve((size_t)count * 2); leaf_bytes.reserve((size_t)count * 32);
  for (uint32_t i = 0; i < count; i++) {
    if (i + 1 < count && key_of(order[i + 1]) == key_of(order[i])) continue;
    hkey k = key_of(order[i]);
    k[depth >> 6] |= 1ull << (depth & 63);  // heap key 2^depth + idx
    keys.push_back(k);
    leaf_bytes.insert(leaf_bytes.end(), vals + 32 * (size_t)order[i], vals + 32 * (size_t)order[i] + 32);
  }
  // levels 1..depth: distinct parents and, for each, which children this batch changed (index into the level below)
  std::vector<size_t> off{0};
  std::vector<int32_t> li, ri;
  li.assign(keys.size(), -1); ri.assign(keys.size(), -1);
  size_t lo = 0, hi = keys.size();
  for (uint32_t l = 0; l < depth; l++) {
    off.push_back(hi);
    for (size_t i = lo; i < hi;) {
      const hkey parent = hkey_shr1(keys[i]);
      int32_t a = -1, b = -1;
      if (keys[i][0] & 1) b = (int32_t)(i - lo); else a = (int32_t)(i - lo);
      size_t next = i + 1;
      if (next < hi && hkey_shr1(keys[next]) == parent) { b = (int32_t)(next - lo); next++; }
      keys.push_back(parent); li.push_back(a); ri.push_back(b);
      i = next;
    }
    lo = hi; hi = keys.size();
  }
  off.push_back(hi);  // off[l] .. off[l+1] = nodes of level l; level `depth` holds the root alone
  const size_t total = keys.size(), nleaf = off[1];
  int rc = table_reserve(T, total, s);
  if (rc) return rc;
  Dev<tkey> d_keys; Dev<int32_t> d_li, d_ri; Dev<uint8_t> d_bytes; Dev<scm> d_vals;
  if (d_keys.alloc(total) || d_li.alloc(total) || d_ri.alloc(total) || d_bytes.alloc(nleaf * 32) || d_vals.alloc(total)) return BP_ERR_OOM;
  static_assert(sizeof(hkey) == sizeof(tkey), "same layout");
  if (dev_h2d(d_keys.p, keys.data(), total * sizeof(tkey), s) || dev_h2d(d_li.p, li.data(), total * 4, s) || dev_h2d(d_ri.p, ri.data(), total * 4, s) ||
      dev_h2d(d_bytes.p, leaf_bytes.data(), nleaf * 32, s))
    return BP_ERR_CUDA;
  if (launch((long)nleaf, s, KTreeLoadLeaves{d_bytes.p, d_vals.p})) return BP_ERR_CUDA;
  for (uint32_t l = 0; l < depth; l++) {
    const size_t a = off[l + 1], n = off[l + 2] - a;
    if (launch((long)n, s, KTreeHashLevel{T->t, T->pos, T->sbox, d_keys.p + a, d_li.p + a, d_ri.p + a, d_vals.p + off[l], T->d_empty + l, d_vals.p + a}))
      return BP_ERR_CUDA;
  }
  if (launch((long)total, s, KTreeInsert{T->t, d_keys.p, d_vals.p, T->d_count})) return BP_ERR_CUDA;
  if (dev_d2d(T->d_root, d_vals.p + total - 1, sizeof(scm), s)) return BP_ERR_CUDA;
  rc = fetch_root(T, s);
  if (rc) return rc;
  if (root_out) memcpy(root_out, T->root_bytes, 32);
  return BP_OK;
}
int32_t bp_vsmt2_update_batch(bp_vsmt2 *T, uint32_t count, const uint64_t *idx, const uint8_t *vals, uint8_t root_out[32]) {
  if (!T || (count && (!idx || !vals))) return BP_ERR_INVALID_ARGUMENT;
  std::vector<uint64_t> w; words_from_u64(w, idx, count);
  return update_batch_words(T, count, w.data(), vals, root_out);
}
int32_t bp_vsmt2_update_batch_wide(bp_vsmt2 *T, uint32_t count, const uint8_t *idx32, const uint8_t *vals, uint8_t root_out[32]) {
  if (!T || (count && (!idx32 || !vals))) return BP_ERR_INVALID_ARGUMENT;
  std::vector<uint64_t> w; words_from_bytes(w, idx32, count);
  return update_batch_words(T, count, w.data(), vals, root_out);
}

// leaves [count][32]; proofs [count][depth][32], siblings root -> leaf as VanillaSparseMerkleTree::get returns them
static int32_t get_batch_words(const bp_vsmt2 *T, uint32_t count, const uint64_t *iw, uint8_t *leaves, uint8_t *proofs) {
  const int depth = (int)T->depth;
  for (uint32_t i = 0; i < count; i++) if (!idx_in_range(iw + 4 * (size_t)i, T->depth)) return BP_ERR_INVALID_ARGUMENT;
  if (!count) return BP_OK;
  dev_stream s = 0;
  Dev<uint64_t> d_idx; Dev<uint8_t> d_leaves, d_proofs;
  if (d_idx.alloc((size_t)count * 4) || d_leaves.alloc((size_t)count * 32) || d_proofs.alloc((size_t)count * depth * 32)) return BP_ERR_OOM;
  if (dev_h2d(d_idx.p, iw, (size_t)count * 32, s)) return BP_ERR_CUDA;
  if (launch((long)count * (depth + 1), s, KTreeGet{T->t, T->d_empty, d_idx.p, depth, 0, d_leaves.p, d_proofs.p})) return BP_ERR_CUDA;
  if (dev_d2h(leaves, d_leaves.p, (size_t)count * 32, s) || dev_d2h(proofs, d_proofs.p, (size_t)count * depth * 32, s) || dev_sync(s)) return BP_ERR_CUDA;
  return BP_OK;
}
int32_t bp_vsmt2_get_batch(const bp_vsmt2 *T, uint32_t count, const uint64_t *idx, uint8_t *leaves, uint8_t *proofs) {
  if (!T || (count && (!idx || !leaves || !proofs))) return BP_ERR_INVALID_ARGUMENT;
  std::vector<uint64_t> w; words_from_u64(w, idx, count);
  return get_batch_words(T, count, w.data(), leaves, proofs);
}
int32_t bp_vsmt2_get_batch_wide(const bp_vsmt2 *T, uint32_t count, const uint8_t *idx32, uint8_t *leaves, uint8_t *proofs) {
  if (!T || (count && (!idx32 || !leaves || !proofs))) return BP_ERR_INVALID_ARGUMENT;
  std::vector<uint64_t> w; words_from_bytes(w, idx32, count);
  return get_batch_words(T, count, w.data(), leaves, proofs);
}

// device buffers: d_idx32 [count][32] little-endian 256-bit indices, d_v [count][2*depth+5][32], d_pub [count][32] or NULL;
// asynchronous on `stream`.  (A little-endian byte string IS the little-endian word array on this platform.)
int32_t bp_vsmt2_witness_batch_wide_device(const bp_vsmt2 *T, uint32_t count, const uint8_t *d_idx32, uint8_t *d_v, uint8_t *d_pub, void *stream) {
  if (!T || (count && (!d_idx32 || !d_v))) return BP_ERR_INVALID_ARGUMENT;
  const int depth = (int)T->depth;
  if (launch((long)count * (2 * depth + 6), as_stream(stream), KTreeWitnessRows{T->t, T->d_empty, T->d_root, (const uint64_t *)d_idx32, depth, d_v, d_pub})) return BP_ERR_CUDA;
  return BP_OK;
}
// device buffers with 64-bit indices (depth <= 63): widened into a scratch buffer owned by the tree
int32_t bp_vsmt2_witness_batch_device(const bp_vsmt2 *T_, uint32_t count, const uint64_t *d_idx, uint8_t *d_v, uint8_t *d_pub, void *stream) {
  bp_vsmt2 *T = const_cast<bp_vsmt2 *>(T_);
  if (!T || (count && (!d_idx || !d_v))) return BP_ERR_INVALID_ARGUMENT;
  if (T->depth > 63) return BP_ERR_INVALID_ARGUMENT;
  if (!count) return BP_OK;
  if (T->widen_cap < count) {
    dev_free(T->d_widen); T->d_widen = nullptr; T->widen_cap = 0;
    if (dalloc(&T->d_widen, (size_t)count * 4)) return BP_ERR_OOM;
    T->widen_cap = count;
  }
  if (launch((long)count, as_stream(stream), KTreeWidenIndices{d_idx, T->d_widen})) return BP_ERR_CUDA;
  return bp_vsmt2_witness_batch_wide_device(T, count, (const uint8_t *)T->d_widen, d_v, d_pub, stream);
}
static int32_t witness_batch_words(const bp_vsmt2 *T, uint32_t count, const uint64_t *iw, uint8_t *v, uint8_t *pub) {
  const int depth = (int)T->depth; const size_t m = 2 * (size_t)depth + 5;
  for (uint32_t i = 0; i < count; i++) if (!idx_in_range(iw + 4 * (size_t)i, T->depth)) return BP_ERR_INVALID_ARGUMENT;
  if (!count) return BP_OK;
  dev_stream s = 0;
  Dev<uint64_t> d_idx; Dev<uint8_t> d_v, d_pub;
  if (d_idx.alloc((size_t)count * 4) || d_v.alloc(count * m * 32) || d_pub.alloc((size_t)count * 32)) return BP_ERR_OOM;
  if (dev_h2d(d_idx.p, iw, (size_t)count * 32, s)) return BP_ERR_CUDA;
  int rc = bp_vsmt2_witness_batch_wide_device(T, count, (const uint8_t *)d_idx.p, d_v.p, pub ? d_pub.p : nullptr, nullptr);
  if (rc) return rc;
  if (dev_d2h(v, d_v.p, count * m * 32, s) || (pub && dev_d2h(pub, d_pub.p, (size_t)count * 32, s)) || dev_sync(s)) return BP_ERR_CUDA;
  return BP_OK;
}
int32_t bp_vsmt2_witness_batch(const bp_vsmt2 *T, uint32_t count, const uint64_t *idx, uint8_t *v, uint8_t *pub) {
  if (!T || (count && (!idx || !v))) return BP_ERR_INVALID_ARGUMENT;
  std::vector<uint64_t> w; words_from_u64(w, idx, count);
  return witness_batch_words(T, count, w.data(), v, pub);
}
int32_t bp_vsmt2_witness_batch_wide(const bp_vsmt2 *T, uint32_t count, const uint8_t *idx32, uint8_t *v, uint8_t *pub) {
  if (!T || (count && (!idx32 || !v))) return BP_ERR_INVALID_ARGUMENT;
  std::vector<uint64_t> w; words_from_bytes(w, idx32, count);
  return witness_batch_words(T, count, w.data(), v, pub);
}

// Poseidon_hash_2 of count pairs on the device (host buffers; xl, xr, out [count][32])
int32_t bp_poseidon_hash_2_batch(const bp_poseidon_params *p, int32_t sbox, uint32_t count, const uint8_t *xl, const uint8_t *xr, uint8_t *out) {
  if (!p || p->width != POSEIDON_WIDTH || (count && (!xl || !xr || !out)) || (sbox != BP_SBOX_CUBE && sbox != BP_SBOX_INVERSE)) return BP_ERR_INVALID_ARGUMENT;
  int rc = bp_device_init();
  if (rc) return rc;
  if (!count) return BP_OK;
  dev_stream s = 0;
  PoseidonDev pos{}; Dev<scm> rk, mds; Dev<uint8_t> dl, dr, dout;
  rc = upload_poseidon(p, pos, &rk.p, &mds.p, s);
  if (rc) return rc;
  const size_t nb = (size_t)count * 32;
  if (dl.alloc(nb) || dr.alloc(nb) || dout.alloc(nb)) return BP_ERR_OOM;
  if (dev_h2d(dl.p, xl, nb, s) || dev_h2d(dr.p, xr, nb, s)) return BP_ERR_CUDA;
  if (launch((long)count, s, KPoseidonHash2Batch{pos, sbox, dl.p, dr.p, dout.p})) return BP_ERR_CUDA;
  if (dev_d2h(out, dout.p, nb, s) || dev_sync(s)) return BP_ERR_CUDA;
  return BP_OK;
}

}  // extern "C"
