// Keccak-f[1600], SHAKE256 / SHA3-512 sponges, STROBE-128 and the Merlin transcript
// operations the Bulletproofs R1CS protocol uses.  Replaces the `merlin` 2.x and `sha3` 0.8
// crates (reference Cargo.toml:10,18; call sites e.g. src/gadget_vsmt_2.rs:293-294).
// One thread owns one transcript: the state lives in 25 registers during a permutation.
#pragma once
#include "hd.h"

HD uint64_t rol64(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }

// Not inlined on the device: with the permutation inlined into the byte-addressed STROBE code nvcc 12.9
// -O3 produced a wrong transcript-RNG state in KTsStart (caught by the oracle parity test; -G and this
// form are both correct).  A call boundary also keeps every transcript kernel small.
#if defined(__CUDACC__)
static __host__ __device__ __noinline__ void keccak_f1600(uint64_t st[25]) {
#else
HD void keccak_f1600(uint64_t st[25]) {
#endif
  const uint64_t RC[24] = {
      0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL,
      0x000000000000808BULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
      0x000000000000008AULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000AULL,
      0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
      0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
      0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
  uint64_t a[25];
#pragma unroll
  for (int i = 0; i < 25; i++) a[i] = st[i];
#pragma unroll 1
  for (int r = 0; r < 24; r++) {
    uint64_t c[5], d[5], b[25];
#pragma unroll
    for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
#pragma unroll
    for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rol64(c[(x + 1) % 5], 1);
#pragma unroll
    for (int i = 0; i < 25; i++) a[i] ^= d[i % 5];
    // rho + pi: b[y + 5*((2x+3y)%5)] = rol(a[x + 5y], rot[x][y])
    b[0] = a[0];
    b[10] = rol64(a[1], 1);   b[20] = rol64(a[2], 62);  b[5] = rol64(a[3], 28);   b[15] = rol64(a[4], 27);
    b[16] = rol64(a[5], 36);  b[1] = rol64(a[6], 44);   b[11] = rol64(a[7], 6);   b[21] = rol64(a[8], 55);  b[6] = rol64(a[9], 20);
    b[7] = rol64(a[10], 3);   b[17] = rol64(a[11], 10); b[2] = rol64(a[12], 43);  b[12] = rol64(a[13], 25); b[22] = rol64(a[14], 39);
    b[23] = rol64(a[15], 41); b[8] = rol64(a[16], 45);  b[18] = rol64(a[17], 15); b[3] = rol64(a[18], 21);  b[13] = rol64(a[19], 8);
    b[14] = rol64(a[20], 18); b[24] = rol64(a[21], 2);  b[9] = rol64(a[22], 61);  b[19] = rol64(a[23], 56); b[4] = rol64(a[24], 14);
#pragma unroll
    for (int y = 0; y < 25; y += 5) {
#pragma unroll
      for (int x = 0; x < 5; x++) a[y + x] = b[y + x] ^ ((~b[y + (x + 1) % 5]) & b[y + (x + 2) % 5]);
    }
    a[0] ^= RC[r];
  }
#pragma unroll
  for (int i = 0; i < 25; i++) st[i] = a[i];
}

// The same permutation on 25 named lanes, always inlined, every index a compile-time constant: the state never leaves the
// register file.  Used by the steady state of the transcript RNG (KRngDraw), where the STROBE operations of one draw touch
// fixed byte positions and nothing needs byte addressing.
HD void keccak_f1600_lanes(uint64_t (&a)[25]) {
  const uint64_t RC[24] = {
      0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL,
      0x000000000000808BULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
      0x000000000000008AULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000AULL,
      0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
      0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
      0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int r = 0; r < 24; r++) {
    uint64_t c[5], d[5], b[25];
#pragma unroll
    for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
#pragma unroll
    for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rol64(c[(x + 1) % 5], 1);
#pragma unroll
    for (int i = 0; i < 25; i++) a[i] ^= d[i % 5];
    b[0] = a[0];
    b[10] = rol64(a[1], 1);   b[20] = rol64(a[2], 62);  b[5] = rol64(a[3], 28);   b[15] = rol64(a[4], 27);
    b[16] = rol64(a[5], 36);  b[1] = rol64(a[6], 44);   b[11] = rol64(a[7], 6);   b[21] = rol64(a[8], 55);  b[6] = rol64(a[9], 20);
    b[7] = rol64(a[10], 3);   b[17] = rol64(a[11], 10); b[2] = rol64(a[12], 43);  b[12] = rol64(a[13], 25); b[22] = rol64(a[14], 39);
    b[23] = rol64(a[15], 41); b[8] = rol64(a[16], 45);  b[18] = rol64(a[17], 15); b[3] = rol64(a[18], 21);  b[13] = rol64(a[19], 8);
    b[14] = rol64(a[20], 18); b[24] = rol64(a[21], 2);  b[9] = rol64(a[22], 61);  b[19] = rol64(a[23], 56); b[4] = rol64(a[24], 14);
#pragma unroll
    for (int y = 0; y < 25; y += 5) {
#pragma unroll
      for (int x = 0; x < 5; x++) a[y + x] = b[y + x] ^ ((~b[y + (x + 1) % 5]) & b[y + (x + 2) % 5]);
    }
    a[0] ^= RC[r];
  }
}

// ---------------------------------------------------------------- byte access into a lane array
HD uint8_t st_get(const uint64_t *st, int pos) { return (uint8_t)(st[pos >> 3] >> (8 * (pos & 7))); }
HD void st_xor(uint64_t *st, int pos, uint8_t b) { st[pos >> 3] ^= (uint64_t)b << (8 * (pos & 7)); }
HD void st_set(uint64_t *st, int pos, uint8_t b) {
  int sh = 8 * (pos & 7);
  st[pos >> 3] = (st[pos >> 3] & ~((uint64_t)0xff << sh)) | ((uint64_t)b << sh);
}

// ---------------------------------------------------------------- plain sponges (host-side set-up uses these)
struct keccak_xof { uint64_t st[25]; int rate, pos; };
HD void keccak_absorb_all(keccak_xof &k, int rate, uint8_t pad, const uint8_t *in, size_t len) {
  for (int i = 0; i < 25; i++) k.st[i] = 0;
  k.rate = rate;
  int pos = 0;
  for (size_t i = 0; i < len; i++) { st_xor(k.st, pos++, in[i]); if (pos == rate) { keccak_f1600(k.st); pos = 0; } }
  st_xor(k.st, pos, pad); st_xor(k.st, rate - 1, 0x80);
  keccak_f1600(k.st); k.pos = 0;
}
HD void keccak_squeeze(keccak_xof &k, uint8_t *out, size_t len) {
  for (size_t i = 0; i < len; i++) { if (k.pos == k.rate) { keccak_f1600(k.st); k.pos = 0; } out[i] = st_get(k.st, k.pos++); }
}
HD void sha3_512(uint8_t out[64], const uint8_t *in, size_t len) { keccak_xof k; keccak_absorb_all(k, 72, 0x06, in, len); keccak_squeeze(k, out, 64); }
HD void shake256_init(keccak_xof &k, const uint8_t *in, size_t len) { keccak_absorb_all(k, 136, 0x1F, in, len); }

// ---------------------------------------------------------------- STROBE-128 subset used by Merlin
#define STROBE_R 166
enum { SF_I = 1, SF_A = 2, SF_C = 4, SF_T = 8, SF_M = 16, SF_K = 32 };
struct alignas(16) strobe128 { uint64_t st[25]; uint8_t pos, pos_begin, cur_flags, pad[5]; };  // 208 B

HD void strobe_run_f(strobe128 &s) {
  st_xor(s.st, s.pos, s.pos_begin); st_xor(s.st, s.pos + 1, 0x04); st_xor(s.st, STROBE_R + 1, 0x80);
  keccak_f1600(s.st); s.pos = 0; s.pos_begin = 0;
}
HD void strobe_absorb(strobe128 &s, const uint8_t *d, int n) {
  for (int i = 0; i < n; i++) { st_xor(s.st, s.pos, d[i]); if (++s.pos == STROBE_R) strobe_run_f(s); }
}
HD void strobe_overwrite(strobe128 &s, const uint8_t *d, int n) {
  for (int i = 0; i < n; i++) { st_set(s.st, s.pos, d[i]); if (++s.pos == STROBE_R) strobe_run_f(s); }
}
// squeeze whole 64-bit words (pos and n multiples of 8, no rate crossing): the transcript-RNG fast path
HD void strobe_squeeze_words(strobe128 &s, uint64_t *d, int nwords) {
  int w0 = s.pos >> 3;
  for (int i = 0; i < nwords; i++) { d[i] = s.st[w0 + i]; s.st[w0 + i] = 0; }
  s.pos = (uint8_t)(s.pos + 8 * nwords);
}
HD void strobe_squeeze(strobe128 &s, uint8_t *d, int n) {
  for (int i = 0; i < n; i++) { d[i] = st_get(s.st, s.pos); st_set(s.st, s.pos, 0); if (++s.pos == STROBE_R) strobe_run_f(s); }
}
HD void strobe_begin_op(strobe128 &s, uint8_t flags, int more) {
  if (more) return;
  uint8_t hdr[2] = {s.pos_begin, flags};
  s.pos_begin = s.pos + 1; s.cur_flags = flags;
  strobe_absorb(s, hdr, 2);
  if ((flags & (SF_C | SF_K)) && s.pos != 0) strobe_run_f(s);
}
HD void strobe_meta_ad(strobe128 &s, const uint8_t *d, int n, int more) { strobe_begin_op(s, SF_M | SF_A, more); strobe_absorb(s, d, n); }
HD void strobe_ad(strobe128 &s, const uint8_t *d, int n, int more) { strobe_begin_op(s, SF_A, more); strobe_absorb(s, d, n); }
HD void strobe_prf(strobe128 &s, uint8_t *d, int n) { strobe_begin_op(s, SF_I | SF_A | SF_C, 0); strobe_squeeze(s, d, n); }
HD void strobe_key(strobe128 &s, const uint8_t *d, int n) { strobe_begin_op(s, SF_A | SF_C, 0); strobe_overwrite(s, d, n); }
HD void strobe_init(strobe128 &s, const uint8_t *label, int n) {
  for (int i = 0; i < 25; i++) s.st[i] = 0;
  const uint8_t hdr[18] = {1, STROBE_R + 2, 1, 0, 1, 96, 'S', 'T', 'R', 'O', 'B', 'E', 'v', '1', '.', '0', '.', '2'};
  for (int i = 0; i < 18; i++) st_set(s.st, i, hdr[i]);
  keccak_f1600(s.st);
  s.pos = 0; s.pos_begin = 0; s.cur_flags = 0;
  strobe_meta_ad(s, label, n, 0);
}

// global <-> local copies, member by member (no type punning on the mixed u64/u8 struct)
HD void strobe_load(strobe128 &d, const strobe128 *src) {
#pragma unroll
  for (int i = 0; i < 25; i++) d.st[i] = src->st[i];
  d.pos = src->pos; d.pos_begin = src->pos_begin; d.cur_flags = src->cur_flags;
}
HD void strobe_store(strobe128 *dst, const strobe128 &s) {
#pragma unroll
  for (int i = 0; i < 25; i++) dst->st[i] = s.st[i];
  dst->pos = s.pos; dst->pos_begin = s.pos_begin; dst->cur_flags = s.cur_flags;
}

// ---------------------------------------------------------------- Merlin transcript
HD void u32le(uint8_t o[4], uint32_t x) { o[0] = (uint8_t)x; o[1] = (uint8_t)(x >> 8); o[2] = (uint8_t)(x >> 16); o[3] = (uint8_t)(x >> 24); }
template <int LN>
HD void ts_append(strobe128 &t, const char (&label)[LN], const uint8_t *msg, int n) {
  uint8_t l4[4]; u32le(l4, (uint32_t)n);
  uint8_t lb[LN];
  for (int i = 0; i < LN - 1; i++) lb[i] = (uint8_t)label[i];
  strobe_meta_ad(t, lb, LN - 1, 0); strobe_meta_ad(t, l4, 4, 1); strobe_ad(t, msg, n, 0);
}
template <int LN>
HD void ts_challenge_bytes(strobe128 &t, const char (&label)[LN], uint8_t *out, int n) {
  uint8_t l4[4]; u32le(l4, (uint32_t)n);
  uint8_t lb[LN];
  for (int i = 0; i < LN - 1; i++) lb[i] = (uint8_t)label[i];
  strobe_meta_ad(t, lb, LN - 1, 0); strobe_meta_ad(t, l4, 4, 1); strobe_prf(t, out, n);
}
template <int LN>
HD void ts_append_u64(strobe128 &t, const char (&label)[LN], uint64_t x) {
  uint8_t b[8];
  for (int i = 0; i < 8; i++) b[i] = (uint8_t)(x >> (8 * i));
  ts_append(t, label, b, 8);
}
HD void ts_init(strobe128 &t, const uint8_t *label, int n) {
  const uint8_t merlin[11] = {'M', 'e', 'r', 'l', 'i', 'n', ' ', 'v', '1', '.', '0'};
  strobe_init(t, merlin, 11);
  ts_append(t, "dom-sep", label, n);
}
// transcript RNG (merlin TranscriptRngBuilder): clone, rekey with witness bytes, finalize with entropy
template <int LN>
HD void trng_rekey(strobe128 &r, const char (&label)[LN], const uint8_t *w, int n) {
  uint8_t l4[4]; u32le(l4, (uint32_t)n);
  uint8_t lb[LN];
  for (int i = 0; i < LN - 1; i++) lb[i] = (uint8_t)label[i];
  strobe_meta_ad(r, lb, LN - 1, 0); strobe_meta_ad(r, l4, 4, 1); strobe_key(r, w, n);
}
HD void trng_finalize(strobe128 &r, const uint8_t entropy[32]) {
  const uint8_t lb[3] = {'r', 'n', 'g'};
  strobe_meta_ad(r, lb, 3, 0); strobe_key(r, entropy, 32);
}
// 64 uniform bytes as eight little-endian words (what Scalar::random reduces): same stream as trng_fill(r, out, 64)
HD void trng_fill64_words(strobe128 &r, uint64_t w[8]) {
  uint8_t l4[4]; u32le(l4, 64);
  strobe_meta_ad(r, l4, 4, 0);
  strobe_begin_op(r, SF_I | SF_A | SF_C, 0);
  if ((r.pos & 7) == 0 && r.pos + 64 < STROBE_R) strobe_squeeze_words(r, w, 8);
  else {
    uint8_t b[64]; strobe_squeeze(r, b, 64);
    for (int i = 0; i < 8; i++) { uint64_t x = 0; for (int j = 7; j >= 0; j--) x = (x << 8) | b[8 * i + j]; w[i] = x; }
  }
}
HD void trng_fill(strobe128 &r, uint8_t *out, int n) {
  uint8_t l4[4]; u32le(l4, (uint32_t)n);
  strobe_meta_ad(r, l4, 4, 0); strobe_prf(r, out, n);
}
