/* TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C) of Keccak-f[1600], SHAKE256,
 * SHA3-512 (FIPS 202) and the Merlin 2.x transcript over STROBE-128 (reference
 * Cargo.toml:10,18; call sites e.g. gadget_mimc.rs:113-114).  merlin / sha3 are
 * un-vendored dependencies, so this follows the published constructions and is pinned
 * by Merlin's own "test protocol" vector (tests/test_oracle.py).
 */
#ifndef MERLIN_REF_H
#define MERLIN_REF_H
#include "ed25519_ref.h"

static const u64 KECCAK_RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL,
    0x000000000000808BULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008AULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000AULL,
    0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int KECCAK_ROTC[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
static const int KECCAK_PILN[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
#define ROL64(x, n) (((x) << (n)) | ((x) >> (64 - (n))))

static void keccak_f1600(u64 st[25]) {
  u64 bc[5], t;
  for (int r = 0; r < 24; r++) {
    for (int i = 0; i < 5; i++) bc[i] = st[i] ^ st[i + 5] ^ st[i + 10] ^ st[i + 15] ^ st[i + 20];
    for (int i = 0; i < 5; i++) {
      t = bc[(i + 4) % 5] ^ ROL64(bc[(i + 1) % 5], 1);
      for (int j = 0; j < 25; j += 5) st[j + i] ^= t;
    }
    t = st[1];
    for (int i = 0; i < 24; i++) { int j = KECCAK_PILN[i]; bc[0] = st[j]; st[j] = ROL64(t, KECCAK_ROTC[i]); t = bc[0]; }
    for (int j = 0; j < 25; j += 5) {
      for (int i = 0; i < 5; i++) bc[i] = st[j + i];
      for (int i = 0; i < 5; i++) st[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
    }
    st[0] ^= KECCAK_RC[r];
  }
}

/* sponge: absorb all input at once, then squeeze incrementally */
typedef struct { u64 st[25]; int rate, pos; } keccak_xof;
static void keccak_absorb_all(keccak_xof *k, int rate, u8 pad, const u8 *in, size_t len) {
  memset(k, 0, sizeof *k); k->rate = rate;
  u8 *s = (u8 *)k->st; int pos = 0;
  for (size_t i = 0; i < len; i++) { s[pos++] ^= in[i]; if (pos == rate) { keccak_f1600(k->st); pos = 0; } }
  s[pos] ^= pad; s[rate - 1] ^= 0x80;
  keccak_f1600(k->st); k->pos = 0;
}
static void keccak_squeeze(keccak_xof *k, u8 *out, size_t len) {
  u8 *s = (u8 *)k->st;
  for (size_t i = 0; i < len; i++) { if (k->pos == k->rate) { keccak_f1600(k->st); k->pos = 0; } out[i] = s[k->pos++]; }
}
static void sha3_512(u8 out[64], const u8 *in, size_t len) { keccak_xof k; keccak_absorb_all(&k, 72, 0x06, in, len); keccak_squeeze(&k, out, 64); }
static void shake256_init(keccak_xof *k, const u8 *in, size_t len) { keccak_absorb_all(k, 136, 0x1F, in, len); }

/* ------------------------------------------------------------------ STROBE-128 (merlin subset) */
#define STROBE_R 166
enum { FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_T = 8, FLAG_M = 16, FLAG_K = 32 };
typedef struct { u64 st[25]; u8 pos, pos_begin, cur_flags; } strobe128;

static void strobe_run_f(strobe128 *s) {
  u8 *b = (u8 *)s->st;
  b[s->pos] ^= s->pos_begin; b[s->pos + 1] ^= 0x04; b[STROBE_R + 1] ^= 0x80;
  keccak_f1600(s->st); s->pos = 0; s->pos_begin = 0;
}
static void strobe_absorb(strobe128 *s, const u8 *d, size_t n) {
  u8 *b = (u8 *)s->st;
  for (size_t i = 0; i < n; i++) { b[s->pos++] ^= d[i]; if (s->pos == STROBE_R) strobe_run_f(s); }
}
static void strobe_overwrite(strobe128 *s, const u8 *d, size_t n) {
  u8 *b = (u8 *)s->st;
  for (size_t i = 0; i < n; i++) { b[s->pos++] = d[i]; if (s->pos == STROBE_R) strobe_run_f(s); }
}
static void strobe_squeeze(strobe128 *s, u8 *d, size_t n) {
  u8 *b = (u8 *)s->st;
  for (size_t i = 0; i < n; i++) { d[i] = b[s->pos]; b[s->pos++] = 0; if (s->pos == STROBE_R) strobe_run_f(s); }
}
static void strobe_begin_op(strobe128 *s, u8 flags, int more) {
  if (more) return;
  u8 hdr[2] = {s->pos_begin, flags};
  s->pos_begin = s->pos + 1; s->cur_flags = flags;
  strobe_absorb(s, hdr, 2);
  if ((flags & (FLAG_C | FLAG_K)) && s->pos != 0) strobe_run_f(s);
}
static void strobe_meta_ad(strobe128 *s, const void *d, size_t n, int more) { strobe_begin_op(s, FLAG_M | FLAG_A, more); strobe_absorb(s, d, n); }
static void strobe_ad(strobe128 *s, const void *d, size_t n, int more) { strobe_begin_op(s, FLAG_A, more); strobe_absorb(s, d, n); }
static void strobe_prf(strobe128 *s, u8 *d, size_t n) { strobe_begin_op(s, FLAG_I | FLAG_A | FLAG_C, 0); strobe_squeeze(s, d, n); }
static void strobe_key(strobe128 *s, const void *d, size_t n) { strobe_begin_op(s, FLAG_A | FLAG_C, 0); strobe_overwrite(s, d, n); }
static void strobe_init(strobe128 *s, const char *label) {
  memset(s, 0, sizeof *s);
  u8 *b = (u8 *)s->st;
  const u8 hdr[6] = {1, STROBE_R + 2, 1, 0, 1, 96};
  memcpy(b, hdr, 6); memcpy(b + 6, "STROBEv1.0.2", 12);
  keccak_f1600(s->st);
  strobe_meta_ad(s, label, strlen(label), 0);
}

/* ------------------------------------------------------------------ Merlin transcript + bulletproofs protocol labels */
typedef struct { strobe128 s; } transcript;
static void u32le(u8 o[4], u32 x) { o[0] = x; o[1] = x >> 8; o[2] = x >> 16; o[3] = x >> 24; }
static void ts_append(transcript *t, const char *label, const void *msg, size_t n) {
  u8 l[4]; u32le(l, (u32)n);
  strobe_meta_ad(&t->s, label, strlen(label), 0); strobe_meta_ad(&t->s, l, 4, 1); strobe_ad(&t->s, msg, n, 0);
}
static void ts_init(transcript *t, const u8 *label, size_t n) { strobe_init(&t->s, "Merlin v1.0"); ts_append(t, "dom-sep", label, n); }
static void ts_append_u64(transcript *t, const char *label, u64 x) { u8 b[8]; for (int i = 0; i < 8; i++) b[i] = x >> (8 * i); ts_append(t, label, b, 8); }
static void ts_challenge_bytes(transcript *t, const char *label, u8 *out, size_t n) {
  u8 l[4]; u32le(l, (u32)n);
  strobe_meta_ad(&t->s, label, strlen(label), 0); strobe_meta_ad(&t->s, l, 4, 1); strobe_prf(&t->s, out, n);
}
static void ts_append_scalar(transcript *t, const char *label, const sc *s) { u8 b[32]; sc_tobytes(b, s); ts_append(t, label, b, 32); }
static void ts_append_point(transcript *t, const char *label, const u8 p[32]) { ts_append(t, label, p, 32); }
static int ts_validate_and_append_point(transcript *t, const char *label, const u8 p[32]) {
  u8 z = 0; for (int i = 0; i < 32; i++) z |= p[i];
  if (!z) return 0;
  ts_append(t, label, p, 32); return 1;
}
static void ts_challenge_scalar(transcript *t, const char *label, sc *out) { u8 b[64]; ts_challenge_bytes(t, label, b, 64); sc_from_bytes_wide(out, b); }

typedef struct { strobe128 s; } transcript_rng;
static void trng_begin(transcript_rng *r, const transcript *t) { r->s = t->s; }
static void trng_rekey(transcript_rng *r, const char *label, const u8 *w, size_t n) {
  u8 l[4]; u32le(l, (u32)n);
  strobe_meta_ad(&r->s, label, strlen(label), 0); strobe_meta_ad(&r->s, l, 4, 1); strobe_key(&r->s, w, n);
}
static void trng_finalize(transcript_rng *r, const u8 entropy[32]) { strobe_meta_ad(&r->s, "rng", 3, 0); strobe_key(&r->s, entropy, 32); }
static void trng_fill(transcript_rng *r, u8 *out, size_t n) { u8 l[4]; u32le(l, (u32)n); strobe_meta_ad(&r->s, l, 4, 0); strobe_prf(&r->s, out, n); }
static void trng_scalar(transcript_rng *r, sc *out) { u8 b[64]; trng_fill(r, b, 64); sc_from_bytes_wide(out, b); }
#endif
