// Kernel bodies of the batched Bulletproofs R1CS prover/verifier.  Each kernel is a functor
// whose operator()(tid) is the work of one CUDA thread; devrt.h turns it into the named
// kernel run_kernel<Functor>.
//
// Layout conventions (B = proofs in the chunk):
//   * per-proof scalar vectors are proof-minor: element i of proof p lives at v[i*B + p], so a warp of
//     consecutive proofs reads/writes 32 consecutive 32-byte scalars (coalesced);
//   * scalars are in Montgomery form (sc25519.h) while on the device;
//   * MSM digit rows are signed radix-256 (32 int8 per scalar, row r of instance q at dig[q*stride + r*32 + w]) for the
//     direct-table and per-proof bucket kernels, and signed radix-2^15 (17 int16 in a 48-byte row) for the sorted-bucket path;
//   * points are ge_p3 (128 B) when per-proof, ge_niels (96 B, affine) when shared generators.
//
// Protocol references: SURVEY.md App. A (Prover::prove A.3, InnerProductProof::create A.4,
// Verifier::verify A.5) -- the reference takes these from the un-vendored `bulletproofs` fork
// (reference Cargo.toml:22-26; call sites src/gadget_vsmt_2.rs:347,395).
#pragma once
#include "ge25519.h"
#include "sc25519.h"
#include "keccak.h"
// occupancy knobs of the point-arithmetic kernels (blocks of 128 threads per SM the compiler must allow for)
#ifndef BP_OCC_TABLE
#define BP_OCC_TABLE 1
#endif
#ifndef BP_OCC_FOLD
#define BP_OCC_FOLD 1
#endif
#ifndef BP_OCC_BUCKET
#define BP_OCC_BUCKET 3
#endif
#define BP_ERR_FORMAT_ 2
#define BP_ERR_VERIFICATION_ 3

// digit geometry and recoding of the sorted-bucket MSM (kernels further down; the verifier kernels write these rows too)
#ifndef SB_BITS
#define SB_BITS 15        // signed window width: 15 -> 17 windows and 16384 buckets (13 -> 20 windows, 4096 buckets)
#endif
#define SB_WINDOWS ((253 + SB_BITS - 1) / SB_BITS)
#define SB_BUCKETS (1 << (SB_BITS - 1))
#define SB_ROW_BYTES 48   // up to 24 int16 digits (3 x 16 B)
#define SB_SEG_LEN 128
#define SB_SEGS (SB_BUCKETS / SB_SEG_LEN)

HD void sc_recode13(int16_t dig[SB_WINDOWS], const scm &s) {
  uint64_t w[4]; sc_to_canonical(w, s);
  int carry = 0;
#pragma unroll
  for (int i = 0; i < SB_WINDOWS; i++) {
    const int bit = SB_BITS * i, word = bit >> 6, off = bit & 63;
    uint64_t v = word < 4 ? (w[word] >> off) : 0;
    if (off > 64 - SB_BITS && word + 1 < 4) v |= w[word + 1] << (64 - off);
    int d = (int)(v & ((1 << SB_BITS) - 1)) + carry;
    carry = d >= (1 << (SB_BITS - 1));
    d -= carry << SB_BITS;
    dig[i] = (int16_t)d;
  }
}
HD void store_digits13(int8_t *dst, const int16_t dig[SB_WINDOWS]) {
  int16_t tmp[24];
#pragma unroll
  for (int i = 0; i < 24; i++) tmp[i] = i < SB_WINDOWS ? dig[i] : 0;
#if defined(__CUDA_ARCH__)
  uint4 a, b, c;
  memcpy(&a, tmp, 16); memcpy(&b, tmp + 8, 16); memcpy(&c, tmp + 16, 16);
  uint4 *d = reinterpret_cast<uint4 *>(dst); d[0] = a; d[1] = b; d[2] = c;
#else
  memcpy(dst, tmp, SB_ROW_BYTES);
#endif
}
HD void load_digits13(int16_t dig[24], const int8_t *src) {
#if defined(__CUDA_ARCH__)
  const uint4 *s = reinterpret_cast<const uint4 *>(src);
  uint4 a = s[0], b = s[1], c = s[2];
  memcpy(dig, &a, 16); memcpy(dig + 8, &b, 16); memcpy(dig + 16, &c, 16);
#else
  memcpy(dig, src, SB_ROW_BYTES);
#endif
}
// ------------------------------------------------------------------------------------------------
// scalars in / out
// ------------------------------------------------------------------------------------------------
// bytes [B][cnt][32] -> Montgomery [cnt][B]
struct KLoadScalars {
  static constexpr int kBlock = 128, kMinBlocks = 1;
  static constexpr const char *kName = "KLoadScalars";
  const uint8_t *in; scm *out; int cnt, B;
  HD void operator()(long tid) const {
    int p = (int)(tid % B); long i = tid / B;
    out[i * B + p] = sc_from_bytes_mod_order(in + ((long)p * cnt + i) * 32);
  }
};

// signed radix-256 digits of a canonical scalar
HD void sc_recode_bytes(int8_t dig[32], const scm &s) {
  uint64_t w[4]; sc_to_canonical(w, s);
  int carry = 0;
#pragma unroll
  for (int i = 0; i < 32; i++) {
    int d = (int)((w[i >> 3] >> (8 * (i & 7))) & 0xff) + carry;
    carry = d >= 128;  // digits in [-128, 127] fit int8
    d -= carry << 8;
    dig[i] = (int8_t)d;
  }
}
// signed radix-2^bits digits (bits in 4..8) of a canonical scalar: W = ceil(253 / bits) digits in [-2^(bits-1), 2^(bits-1))
HD void sc_recode_win(int8_t *dig, const scm &s, int bits, int W) {
  uint64_t w[4]; sc_to_canonical(w, s);
  const int mask = (1 << bits) - 1, half = 1 << (bits - 1);
  int carry = 0;
  for (int i = 0; i < W; i++) {
    const int bit = bits * i, word = bit >> 6, off = bit & 63;
    uint64_t v = word < 4 ? (w[word] >> off) : 0;
    if (off + bits > 64 && word + 1 < 4) v |= w[word + 1] << (64 - off);
    int d = (int)(v & (uint64_t)mask) + carry;
    carry = d >= half;
    d -= carry << bits;
    dig[i] = (int8_t)d;
  }
}
HD void store_digits(int8_t *dst, const int8_t dig[32]) {
#if defined(__CUDA_ARCH__)
  uint4 a, b;
  memcpy(&a, dig, 16); memcpy(&b, dig + 16, 16);
  reinterpret_cast<uint4 *>(dst)[0] = a; reinterpret_cast<uint4 *>(dst)[1] = b;
#else
  memcpy(dst, dig, 32);
#endif
}
// src [cnt][B] (optionally times mul[p]) -> digit rows row0.. of each instance
struct KRecode {
  static constexpr int kBlock = 128, kMinBlocks = 1;
  static constexpr const char *kName = "KRecode";
  const scm *src; const scm *mul; int cnt, B; int8_t *dig; long inst_stride; int row0;
  HD void operator()(long tid) const {
    int p = (int)(tid % B); long i = tid / B;
    scm s = src[i * B + p];
    if (mul) s = sc_mul(s, mul[p]);
    int8_t d[32]; sc_recode_bytes(d, s);
    store_digits(dig + (long)p * inst_stride + (row0 + i) * 32, d);
  }
};

// ------------------------------------------------------------------------------------------------
// Pedersen-style fixed-base commitments  a*B + b*B_blinding  (PedersenGens::commit)
// table[base][window 0..63][digit 0..15] = digit * 16^window * P  in niels form (digit 0 unused)
// ------------------------------------------------------------------------------------------------
struct KCommit {
  static constexpr int kBlock = 128, kMinBlocks = 1;
  static constexpr const char *kName = "KCommit";
  const scm *a; const scm *b; int cnt, B; const ge_niels *table;
  uint8_t *out; long out_stride_p, out_stride_j; ge_p3 *out_pt;
  HD void operator()(long tid) const {
    int p = (int)(tid % B); long j = tid / B;
    ge_p3 acc; ge_identity(acc);
    for (int base = 0; base < 2; base++) {
      const scm *src = base ? b : a;
      if (!src) continue;
      uint64_t w[4]; sc_to_canonical(w, src[j * B + p]);
      for (int win = 0; win < 64; win++) {
        int d = (int)((w[win >> 4] >> (4 * (win & 15))) & 15);
        if (d) { ge_niels q; load_struct(q, &table[(base * 64 + win) * 16 + d]); ge_madd(acc, acc, q, 0); }
      }
    }
    if (out) ristretto_encode(out + (long)p * out_stride_p + j * out_stride_j, acc);
    if (out_pt) store_struct(&out_pt[j * B + p], acc);
  }
};

// ------------------------------------------------------------------------------------------------
// Multi-scalar multiplication: bucket method, one thread per (instance, 8-bit window).
// A warp is one instance; lane w owns window w and a private set of 128 buckets in global memory
// (signed digits), reads the same base point as the other 31 lanes (broadcast) and byte w of the
// shared 32-byte digit row (one sector per warp).  No atomics, no sorting.
// ------------------------------------------------------------------------------------------------
struct MsmSeg { const void *bases; long inst_stride; int fmt; int count; };  // fmt 0: shared/strided ge_niels, 1: ge_p3
#define MSM_WINDOWS 32
#define MSM_BUCKETS 128

struct KMsmAccumulate {
  static constexpr int kBlock = 128, kMinBlocks = BP_OCC_BUCKET;
  static constexpr const char *kName = "KMsmAccumulate";
  MsmSeg seg[4]; int nseg; int S;  // S = row-range splits per instance (each split owns its own buckets)
  const int8_t *dig; long dig_inst_stride; ge_p3 *buckets; ge_p3 *wsum;
  HD void operator()(long tid) const {
    long q = tid / MSM_WINDOWS; int w = (int)(tid % MSM_WINDOWS);
    long inst = q / S; int sp = (int)(q % S);
    ge_p3 *bk = buckets + tid * MSM_BUCKETS;
    {
      ge_p3 id; ge_identity(id);
      for (int d = 0; d < MSM_BUCKETS; d++) store_struct256(&bk[d], id);
    }
    long total = 0;
    for (int s = 0; s < nseg; s++) total += seg[s].count;
    const long r0 = total * sp / S, r1 = total * (sp + 1) / S;
    const int8_t *drow = dig + inst * dig_inst_stride + w;
    long row = 0;
    for (int s = 0; s < nseg; s++) {
      const long cnt = seg[s].count;
      long lo = (r0 > row ? r0 : row) - row, hi = (r1 < row + cnt ? r1 : row + cnt) - row;
      if (seg[s].fmt == 0) {
        const ge_niels *bases = (const ge_niels *)seg[s].bases + inst * seg[s].inst_stride;
        for (long t = lo; t < hi; t++) {
          int d = drow[(row + t) * 32];
          if (d != 0) {
            int neg = d < 0; int idx = (neg ? -d : d) - 1;
            ge_niels qn; load_struct(qn, &bases[t]);
            ge_p3 acc; load_struct(acc, &bk[idx]);
            ge_madd(acc, acc, qn, neg);  // shared generators without tables: not a hot configuration, keeps the called multiplications
            store_struct(&bk[idx], acc);
          }
        }
      } else {
        const ge_p3 *bases = (const ge_p3 *)seg[s].bases + inst * seg[s].inst_stride;
        for (long t = lo; t < hi; t++) {
          int d = drow[(row + t) * 32];
          if (d != 0) {
            int neg = d < 0; int idx = (neg ? -d : d) - 1;
            ge_p3 qp; load_struct256(qp, &bases[t]);   // 256-bit loads / stores (hd.h): this loop is a read-modify-write of whole points
            ge_cached c; ge_to_cached<true>(c, qp);
            ge_p3 acc; load_struct256(acc, &bk[idx]);
            ge_add_cached_f(acc, acc, c, neg);
            store_struct256(&bk[idx], acc);
          }
        }
      }
      row += cnt;
    }
    // window sum = sum_d (d+1) * bucket[d] by the running-sum trick
    ge_p3 run, tot; ge_identity(run); ge_identity(tot);
    for (int d = MSM_BUCKETS - 1; d >= 0; d--) {
      ge_p3 b; load_struct256(b, &bk[d]);
      ge_add_f(run, run, b);
      ge_add_f(tot, tot, run);
    }
    store_struct(&wsum[tid], tot);
  }
};

// sum the S splits of every (instance, window) in parallel; output compacted to [inst][32] at the front of the same buffer is
// not possible in place, so the sums go to a second array
struct KMsmWindowSum {
  static constexpr int kBlock = 64, kMinBlocks = 1;
  static constexpr const char *kName = "KMsmWindowSum";
  const ge_p3 *wsum; int S; ge_p3 *out;
  HD void operator()(long tid) const {
    long inst = tid / MSM_WINDOWS; int w = (int)(tid % MSM_WINDOWS);
    ge_p3 r; load_struct(r, &wsum[(inst * S) * MSM_WINDOWS + w]);
    for (int s = 1; s < S; s++) { ge_p3 t; load_struct(t, &wsum[(inst * S + s) * MSM_WINDOWS + w]); ge_add(r, r, t); }
    store_struct(&out[tid], r);
  }
};
// sum the splits, Horner over the 32 window sums, then ristretto-encode (mode 0) or test for the identity (mode 1)
struct KMsmFinish {
  static constexpr int kBlock = 64, kMinBlocks = 1;
  static constexpr const char *kName = "KMsmFinish";
  const ge_p3 *wsum; int S; uint8_t *out; long out_stride; int mode; int *status; int fail_code;
  HD void window(ge_p3 &r, long inst, int w) const {
    load_struct(r, &wsum[(inst * S) * MSM_WINDOWS + w]);
    for (int s = 1; s < S; s++) { ge_p3 t; load_struct(t, &wsum[(inst * S + s) * MSM_WINDOWS + w]); ge_add(r, r, t); }
  }
  HD void operator()(long inst) const {
    ge_p3 acc; window(acc, inst, MSM_WINDOWS - 1);
    for (int w = MSM_WINDOWS - 2; w >= 0; w--) {
      for (int i = 0; i < 7; i++) ge_dbl_p2(acc, acc);
      ge_dbl(acc, acc);
      ge_p3 s; window(s, inst, w);
      ge_add(acc, acc, s);
    }
    if (mode == 0) ristretto_encode(out + inst * out_stride, acc);
    else if (!ge_is_identity_ristretto(acc) && status[inst] == 0) status[inst] = fail_code;
  }
};

// ------------------------------------------------------------------------------------------------
// Merlin transcript phases (one thread per proof)
// ------------------------------------------------------------------------------------------------
#define TS_CHALLENGE(t, label, dst) do { uint8_t _b[64]; ts_challenge_bytes(t, label, _b, 64); dst = sc_from_bytes_wide(_b); } while (0)

// base = Transcript::new(label) + "r1cs v1" domain separator (computed once on the host).
// Appends V_0..V_{m-1} and "m"; then builds the prover's transcript RNG (A.3 step 2).
struct KTsStart {
  static constexpr int kBlock = 64, kMinBlocks = 1;
  static constexpr const char *kName = "KTsStart";
  strobe128 base; const uint8_t *V; int m, B; const scm *vbl; const uint8_t *entropy; strobe128 *ts; strobe128 *rng; int prover;
  HD void operator()(long p) const {
    strobe128 t; strobe_load(t, &base);
    for (int j = 0; j < m; j++) ts_append(t, "V", V + ((long)p * m + j) * 32, 32);
    ts_append_u64(t, "m", (uint64_t)m);
    strobe_store(&ts[p], t);
    if (prover) {
      for (int j = 0; j < m; j++) { uint8_t b[32]; sc_tobytes(b, vbl[(long)j * B + p]); trng_rekey(t, "v_blinding", b, 32); }
      trng_finalize(t, entropy + p * 32);
      strobe_store(&rng[p], t);
    }
  }
};
// draws `count` uniform scalars (64 bytes each, wide reduction) into dst[i*B+p]
// Blocks of 128 threads (one warp per SM sub-partition): these latency chains run beside the throughput kernels of the previous
// chunk, whose blocks hold 16 K registers each -- a 128-thread block here displaces about one of them, where four 32-thread
// blocks spread over four SMs would each strand most of a 16 K slot.
struct KRngDraw {
  static constexpr int kBlock = 128, kMinBlocks = 1;
  static constexpr const char *kName = "KRngDraw";
  strobe128 *rng; scm *dst; int count, B;
  // One draw of 64 bytes (merlin TranscriptRng::fill_bytes) is: meta_ad(u32le(64)), begin_op(PRF) -> pad + permute, squeeze 64.
  // After any draw the sponge is at pos = 64, pos_begin = 0, so from the second draw on the operations land on FIXED bytes:
  //   bytes 64..71 ^= [0, M|A, 64, 0, 0, 0, 65, I|A|C]   (the two operation headers and the length)    = lane 8
  //   bytes 72, 73 ^= [71, 0x04],  byte 167 ^= 0x80        (STROBE's padding of run_f)                  = lanes 9 and 20
  // then Keccak-f, lanes 0..7 are the output and are zeroed.  The steady state therefore runs on 25 lanes in registers with no
  // byte addressing at all (the generic path keeps the state in local memory: 36 355 draws per proof at depth 32 took 1.8 s).
  HD void operator()(long p) const {
    strobe128 r; strobe_load(r, &rng[p]);
    int i = 0;
    for (; i < count && !(r.pos == 64 && r.pos_begin == 0); i++) { uint64_t w[8]; trng_fill64_words(r, w); dst[(long)i * B + p] = sc_from_words_wide(w); }
    if (i < count) {
      uint64_t a[25];
#pragma unroll
      for (int j = 0; j < 25; j++) a[j] = r.st[j];
      for (; i < count; i++) {
        a[8] ^= 0x0741000000401200ULL; a[9] ^= 0x0447ULL; a[20] ^= 0x8000000000000000ULL;
        keccak_f1600_lanes(a);
        uint64_t w[8];
#pragma unroll
        for (int j = 0; j < 8; j++) { w[j] = a[j]; a[j] = 0; }
        dst[(long)i * B + p] = sc_from_words_wide(w);
      }
#pragma unroll
      for (int j = 0; j < 25; j++) r.st[j] = a[j];
      r.pos = 64; r.pos_begin = 0; r.cur_flags = SF_I | SF_A | SF_C;
    }
    strobe_store(&rng[p], r);
  }
};
// A_I1,A_O1,S1 + one-phase separator + identity A_I2,A_O2,S2 -> y, z ; also y^-1
struct KTsPhase2 {
  static constexpr int kBlock = 64, kMinBlocks = 1;
  static constexpr const char *kName = "KTsPhase2";
  strobe128 *ts; const uint8_t *proofs; long proof_stride; scm *y, *z, *yinv; int *status; int verifier;
  HD void operator()(long p) const {
    strobe128 t; strobe_load(t, &ts[p]);
    const uint8_t *pf = proofs + p * proof_stride;
    if (verifier) {
      for (int j = 0; j < 3; j++) { uint8_t nz = 0; for (int i = 0; i < 32; i++) nz |= pf[32 * j + i]; if (!nz) status[p] = 3; }
    }
    ts_append(t, "A_I1", pf, 32); ts_append(t, "A_O1", pf + 32, 32); ts_append(t, "S1", pf + 64, 32);
    const uint8_t ph[11] = {'r', '1', 'c', 's', '-', '1', 'p', 'h', 'a', 's', 'e'};
    ts_append(t, "dom-sep", ph, 11);
    ts_append(t, "A_I2", pf + 96, 32); ts_append(t, "A_O2", pf + 128, 32); ts_append(t, "S2", pf + 160, 32);
    scm yy, zz;
    TS_CHALLENGE(t, "y", yy); TS_CHALLENGE(t, "z", zz);
    y[p] = yy; z[p] = zz; yinv[p] = sc_invert(yy);
    strobe_store(&ts[p], t);
  }
};
// T_1,T_3,T_4,T_5,T_6 -> u, x
struct KTsPhase3 {
  static constexpr int kBlock = 64, kMinBlocks = 1;
  static constexpr const char *kName = "KTsPhase3";
  strobe128 *ts; const uint8_t *proofs; long proof_stride; scm *u, *x; int *status; int verifier;
  HD void operator()(long p) const {
    strobe128 t; strobe_load(t, &ts[p]);
    const uint8_t *pf = proofs + p * proof_stride + 192;
    if (verifier) {
      for (int j = 0; j < 5; j++) { uint8_t nz = 0; for (int i = 0; i < 32; i++) nz |= pf[32 * j + i]; if (!nz) status[p] = 3; }
    }
    ts_append(t, "T_1", pf, 32); ts_append(t, "T_3", pf + 32, 32); ts_append(t, "T_4", pf + 64, 32);
    ts_append(t, "T_5", pf + 96, 32); ts_append(t, "T_6", pf + 128, 32);
    scm uu, xx;
    TS_CHALLENGE(t, "u", uu); TS_CHALLENGE(t, "x", xx);
    u[p] = uu; x[p] = xx;
    strobe_store(&ts[p], t);
  }
};
// t_x, t_x_blinding, e_blinding -> w ; then the inner-product domain separator with n = N
struct KTsPhase4 {
  static constexpr int kBlock = 64, kMinBlocks = 1;
  static constexpr const char *kName = "KTsPhase4";
  strobe128 *ts; const uint8_t *proofs; long proof_stride; scm *w; unsigned N;
  HD void operator()(long p) const {
    strobe128 t; strobe_load(t, &ts[p]);
    const uint8_t *pf = proofs + p * proof_stride + 352;
    ts_append(t, "t_x", pf, 32); ts_append(t, "t_x_blinding", pf + 32, 32); ts_append(t, "e_blinding", pf + 64, 32);
    scm ww; TS_CHALLENGE(t, "w", ww); w[p] = ww;
    const uint8_t ipp[6] = {'i', 'p', 'p', ' ', 'v', '1'};
    ts_append(t, "dom-sep", ipp, 6);
    ts_append_u64(t, "n", (uint64_t)N);
    strobe_store(&ts[p], t);
  }
};

// width-4 non-adjacent form of a 256-bit integer: digits in {+-1, +-3, +-5, +-7}, at most one non-zero digit in any four
// consecutive positions (about one addition per five bits instead of one per three for the plain NAF);
// returns the index of the top non-zero digit (or -1)
HD int naf_words(int8_t naf[256], const uint64_t k_in[4]) {
  uint64_t k[4] = {k_in[0], k_in[1], k_in[2], k_in[3]};
  int top = -1;
  for (int i = 0; i < 256; i++) {
    int8_t z = 0;
    if (k[0] & 1) {
      int d = (int)(k[0] & 15);
      if (d > 8) {  // negative digit d - 16: add its magnitude back
        d -= 16;
        uint64_t c = (uint64_t)(-d);
        for (int j = 0; j < 4; j++) { uint64_t t = k[j] + c; c = t < c; k[j] = t; }
      } else {
        k[0] -= (uint64_t)d;
      }
      z = (int8_t)d;
      top = i;
    }
    naf[i] = z;
    k[0] = (k[0] >> 1) | (k[1] << 63); k[1] = (k[1] >> 1) | (k[2] << 63); k[2] = (k[2] >> 1) | (k[3] << 63); k[3] >>= 1;
  }
  return top;
}
HD int sc_naf(int8_t naf[256], const scm &s) { uint64_t k[4]; sc_to_canonical(k, s); return naf_words(naf, k); }

// Half-size decomposition of a fold scalar.  The generator fold G' = G_lo + e G_hi is only needed up to a known factor (the
// factor is carried in the scalars, see KTsIpaRound), so instead of one 253-bit scalar we look for c with BOTH c and
// d = c e mod l short: the extended Euclidean algorithm on (l, e) produces remainders r_i = t_i e (mod l) with |t_i| r_{i-1} <= l;
// stopping at the first r_i < 2^127 gives |c| = |t_i| <= 2^126, 0 <= d = r_i < 2^127.  Then c G_lo + d G_hi = c (G_lo + e G_hi)
// costs ~128 shared doublings and two ~26-addition digit strings instead of 253 doublings and ~51 additions.
// Quotients are taken bit by bit (shift-and-subtract); a remainder moves to r1 only when fully reduced, so (t1, r1) is always
// a true step of the Euclidean sequence.  c comes back as sign + magnitude.
HD int u256_bitlen(const uint64_t a[4]) {
  for (int i = 3; i >= 0; i--) if (a[i]) { int b = 63; while (!((a[i] >> b) & 1)) b--; return 64 * i + b + 1; }
  return 0;
}
HD void u256_shl(uint64_t r[4], const uint64_t a[4], int k) {
  const int ws = k >> 6, bs = k & 63;
  for (int i = 3; i >= 0; i--) {
    uint64_t v = 0;
    if (i - ws >= 0) { v = a[i - ws] << bs; if (bs && i - ws - 1 >= 0) v |= a[i - ws - 1] >> (64 - bs); }
    r[i] = v;
  }
}
HD int u256_ge(const uint64_t a[4], const uint64_t b[4]) {
  for (int i = 3; i >= 0; i--) { if (a[i] != b[i]) return a[i] > b[i]; }
  return 1;
}
HD void u256_sub(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) {  // mod 2^256 (two's complement for the cofactors)
  uint64_t br = 0;
  for (int i = 0; i < 4; i++) { const uint64_t t = a[i] - b[i], t2 = t - br; br = (uint64_t)((a[i] < b[i]) | (t < br)); r[i] = t2; }
}
HD void sc_half_size(uint64_t cmag[4], int &cneg, uint64_t d[4], const scm &e) {
  uint64_t r0[4] = SC_L_LIMBS, r1[4], t0[4] = {0, 0, 0, 0}, t1[4] = {1, 0, 0, 0};
  sc_to_canonical(r1, e);
  while (u256_bitlen(r1) > 127) {
    int k = u256_bitlen(r0) - u256_bitlen(r1);
    uint64_t sh[4]; u256_shl(sh, r1, k);
    if (!u256_ge(r0, sh)) { k--; u256_shl(sh, r1, k); }
    uint64_t ts[4]; u256_shl(ts, t1, k);
    u256_sub(r0, r0, sh); u256_sub(t0, t0, ts);
    if (!u256_ge(r0, r1)) {
      for (int i = 0; i < 4; i++) { uint64_t x = r0[i]; r0[i] = r1[i]; r1[i] = x; x = t0[i]; t0[i] = t1[i]; t1[i] = x; }
    }
  }
  cneg = (int)(t1[3] >> 63);
  if (cneg) { const uint64_t z[4] = {0, 0, 0, 0}; u256_sub(cmag, z, t1); } else { for (int i = 0; i < 4; i++) cmag[i] = t1[i]; }
  for (int i = 0; i < 4; i++) d[i] = r1[i];
}
HD scm sc_from_u256_small(const uint64_t w[4]) { scm x; x.v[0] = w[0]; x.v[1] = w[1]; x.v[2] = w[2]; x.v[3] = w[3]; return sc_montmul_any(x, sc_r2()); }  // w < l

// One inner-product round, transcript side (A.4): append L,R, draw u; derive everything the
// scalar fold and the generator fold of this round need.
//   folded generators are kept un-normalised:  G_true[i] = alpha * Gt[i],  H_true[i] = beta * y^-i * Ht[i]
//   Gt'[i] = cG (Gt[i] + eG * Gt[h+i]) = cG Gt[i] + dG Gt[h+i],  eG = u^2   (times the G-factor class of index h+i in round 0)
//   Ht'[i] = cH (Ht[i] + eH * Ht[h+i]) = cH Ht[i] + dH Ht[h+i],  eH = u^-2 * y^-h  (same class factor)
//   (c, d) = half-size decomposition of e (sc_half_size); round 0 with its two classes keeps c = 1, d = e
//   alpha' = alpha * u^-1 / cG,  beta' = beta * u / cH
// NAF layout: naf[((p*4 + which*2 + cls)*2 + {0: c on the low point, 1: d on the high point}) * 256], tops likewise.
struct KTsIpaRound {
  static constexpr int kBlock = 64, kMinBlocks = 1;
  static constexpr const char *kName = "KTsIpaRound";
  strobe128 *ts; const uint8_t *proofs; long proof_stride; int round; int B; int h;
  const scm *yinvpow; const scm *ufac;  // y^-i table [N][B]; the r1cs challenge u (class factor, round 0 only)
  scm *u, *uinv, *alpha, *beta; int8_t *naf; int *naf_top; int *status; int verifier; int yfree;  // yfree: generators already carry y^-i
  HD void put_naf(long p, int slot, int which, const uint64_t k[4], int neg) const {
    int8_t nf[256];
    const int top = naf_words(nf, k);
    int8_t *dst = naf + ((p * 4 + slot) * 2 + which) * 256;
    for (int i = 0; i <= top; i++) dst[i] = neg ? (int8_t)-nf[i] : nf[i];
    naf_top[(p * 4 + slot) * 2 + which] = top;
  }
  HD void full_size(long p, int slot, const scm &e) const {
    const uint64_t one[4] = {1, 0, 0, 0};
    uint64_t k[4]; sc_to_canonical(k, e);
    put_naf(p, slot, 0, one, 0); put_naf(p, slot, 1, k, 0);
  }
  HD scm half_size(long p, int slot, const scm &e) const {  // writes both digit strings, returns c as a scalar
    uint64_t cm[4], d[4]; int cneg;
    sc_half_size(cm, cneg, d, e);
    put_naf(p, slot, 0, cm, cneg); put_naf(p, slot, 1, d, 0);
    scm c = sc_from_u256_small(cm);
    return cneg ? sc_neg(c) : c;
  }
  HD void operator()(long p) const {
    strobe128 t; strobe_load(t, &ts[p]);
    const uint8_t *pf = proofs + p * proof_stride + 448 + 64 * round;
    if (verifier) {
      for (int j = 0; j < 2; j++) { uint8_t nz = 0; for (int i = 0; i < 32; i++) nz |= pf[32 * j + i]; if (!nz) status[p] = 3; }
    }
    ts_append(t, "L", pf, 32); ts_append(t, "R", pf + 32, 32);
    scm uu; TS_CHALLENGE(t, "u", uu);
    strobe_store(&ts[p], t);
    scm ui = sc_invert(uu);
    u[p] = uu; uinv[p] = ui;
    if (verifier) return;
    scm eG = sc_sqr(uu), eH = sc_sqr(ui);
    if (!yfree) eH = sc_mul(eH, yinvpow[(long)h * B + p]);
    scm cG = sc_one(), cH = sc_one();
    if (round == 0) {
      scm f = ufac[p];
      full_size(p, 0, eG); full_size(p, 1, sc_mul(eG, f));
      full_size(p, 2, eH); full_size(p, 3, sc_mul(eH, f));
    } else {
      cG = half_size(p, 0, eG); cH = half_size(p, 2, eH);
    }
    alpha[p] = sc_mul(sc_mul(alpha[p], ui), sc_invert(cG)); beta[p] = sc_mul(sc_mul(beta[p], uu), sc_invert(cH));
  }
};

// ------------------------------------------------------------------------------------------------
// scalar-vector kernels
// ------------------------------------------------------------------------------------------------
// out[i][p] = base[p]^(i + exp0), i < len; one thread per (chunk of CH exponents, proof)
struct KPowers {
  static constexpr int kBlock = 128, kMinBlocks = 1;
  static constexpr const char *kName = "KPowers";
  const scm *base; scm *out; int len, B, exp0, CH;
  HD void operator()(long tid) const {
    int p = (int)(tid % B); int c = (int)(tid / B);
    int i0 = c * CH, i1 = i0 + CH < len ? i0 + CH : len;
    scm b = base[p];
    scm cur = sc_pow_u32(b, (uint32_t)(i0 + exp0));
    for (int i = i0; i < i1; i++) { out[(long)i * B + p] = cur; cur = sc_mul(cur, b); }
  }
};
HD void fill_scalar(scm *dst, long n, const scm &v) { for (long i = 0; i < n; i++) dst[i] = v; }
struct KFillScalar {
  static constexpr int kBlock = 128, kMinBlocks = 1;
  static constexpr const char *kName = "KFillScalar";
  scm *dst; scm v;
  HD void operator()(long tid) const { dst[tid] = v; }
};

// flattened constraint weights (A.3 step 7) from the slot-major transpose of the constraint matrix:
// slot s in [0,3n+m+1) = wL | wR | wO | wV | wc;  w[s][p] = sum_t coeff_t * z^(q_t+1)   (signs folded into coeff)
struct KFlatten {
  static constexpr int kBlock = 128, kMinBlocks = 1;
  static constexpr const char *kName = "KFlatten";
  const uint32_t *slot_ptr; const uint32_t *t_q; const scm *t_coeff; const scm *zpow; scm *w; int B;
  uint32_t split = 0;  // slots with more terms than this are left to KFlattenParts / KFlattenSum (0: none are)
  HD void operator()(long tid) const {
    int p = (int)(tid % B); long s = tid / B;
    const uint32_t t0 = slot_ptr[s], t1 = slot_ptr[s + 1];
    if (split && t1 - t0 > split) return;
    scm acc = sc_zero();
    for (uint32_t t = t0; t < t1; t++) acc = sc_add(acc, sc_mul(t_coeff[t], zpow[(long)t_q[t] * B + p]));
    w[s * B + p] = acc;
  }
};
// A slot with very many terms -- the constant slot wc carries one term per constraint with a constant (every Poseidon round
// key: 42k terms at depth 32, 340k at depth 253) -- would be one thread's sequential chain per proof; the circuit cuts such slots
// into parts of FLATTEN_PART terms (circuit_create), summed per part here and per slot in KFlattenSum.
#define FLATTEN_PART 256
struct KFlattenParts {
  static constexpr int kBlock = 128, kMinBlocks = 1;
  static constexpr const char *kName = "KFlattenParts";
  const uint32_t *part_beg, *part_end; const uint32_t *t_q; const scm *t_coeff; const scm *zpow; scm *parts; int B;
  HD void operator()(long tid) const {
    int p = (int)(tid % B); long part = tid / B;
    scm acc = sc_zero();
    for (uint32_t t = part_beg[part]; t < part_end[part]; t++) acc = sc_add(acc, sc_mul(t_coeff[t], zpow[(long)t_q[t] * B + p]));
    parts[part * B + p] = acc;
  }
};
struct KFlattenSum {
  static constexpr int kBlock = 128, kMinBlocks = 1;
  static constexpr const char *kName = "KFlattenSum";
  const uint32_t *long_slot, *long_first; const scm *parts; scm *w; int B;  // parts long_first[l] .. long_first[l+1] belong to slot long_slot[l]
  HD void operator()(long tid) const {
    int p = (int)(tid % B); long l = tid / B;
    scm acc = sc_zero();
    for (uint32_t j = long_first[l]; j < long_first[l + 1]; j++) acc = sc_add(acc, parts[(long)j * B + p]);
    w[(long)long_slot[l] * B + p] = acc;
  }
};

struct PolyIn { const scm *aL, *aR, *aO, *sL, *sR, *wL, *wR, *wO, *ypow, *yinvpow; };
struct PolyCoef { scm l1, l2, l3, r0, r1, r3; };
HD void poly_coef(PolyCoef &c, const PolyIn &in, long at) {
  scm ey = in.ypow[at];
  c.l1 = sc_add(in.aL[at], sc_mul(in.yinvpow[at], in.wR[at]));
  c.l2 = in.aO[at];
  c.l3 = in.sL[at];
  c.r0 = sc_sub(in.wO[at], ey);
  c.r1 = sc_add(sc_mul(ey, in.aR[at]), in.wL[at]);
  c.r3 = sc_mul(ey, in.sR[at]);
}
// partial sums of t_1..t_6 (A.3 step 9) over a chunk of multipliers
struct KPolyT {
  static constexpr int kBlock = 128, kMinBlocks = 1;
  static constexpr const char *kName = "KPolyT";
  PolyIn in; int n, B, CH; scm *part;  // part[(c*6 + j)*B + p]
  HD void operator()(long tid) const {
    int p = (int)(tid % B); int c = (int)(tid / B);
    int i0 = c * CH, i1 = i0 + CH < n ? i0 + CH : n;
    scm t1 = sc_zero(), t2 = t1, t3 = t1, t4 = t1, t5 = t1, t6 = t1;
    for (int i = i0; i < i1; i++) {
      PolyCoef k; poly_coef(k, in, (long)i * B + p);
      t1 = sc_add(t1, sc_mul(k.l1, k.r0));
      t2 = sc_add(t2, sc_add(sc_mul(k.l1, k.r1), sc_mul(k.l2, k.r0)));
      t3 = sc_add(t3, sc_add(sc_mul(k.l2, k.r1), sc_mul(k.l3, k.r0)));
      t4 = sc_add(t4, sc_add(sc_mul(k.l1, k.r3), sc_mul(k.l3, k.r1)));
      t5 = sc_add(t5, sc_mul(k.l2, k.r3));
      t6 = sc_add(t6, sc_mul(k.l3, k.r3));
    }
    scm *o = part + ((long)c * 6) * B + p;
    o[0] = t1; o[B] = t2; o[2L * B] = t3; o[3L * B] = t4; o[4L * B] = t5; o[5L * B] = t6;
  }
};
// out[j][p] = sum_c part[(c*nv + j)*B + p]
struct KSumPartials {
  static constexpr int kBlock = 128, kMinBlocks = 1;
  static constexpr const char *kName = "KSumPartials";
  const scm *part; int nchunks, nv, B; scm *out;
  HD void operator()(long tid) const {
    int p = (int)(tid % B); int j = (int)(tid / B);
    scm acc = sc_zero();
    for (int c = 0; c < nchunks; c++) acc = sc_add(acc, part[((long)c * nv + j) * B + p]);
    out[(long)j * B + p] = acc;
  }
};
// l(x), r(x) padded to N (A.3 step 12) -> the a, b vectors of the inner-product argument
struct KPolyEval {
  static constexpr int kBlock = 128, kMinBlocks = 1;
  static constexpr const char *kName = "KPolyEval";
  PolyIn in; int n, B; const scm *x; scm *a, *b;
  HD void operator()(long tid) const {
    int p = (int)(tid % B); long i = tid / B; long at = i * B + p;
    if (i < n) {
      PolyCoef k; poly_coef(k, in, at);
      scm xx = x[p], x2 = sc_sqr(xx), x3 = sc_mul(x2, xx);
      a[at] = sc_add(sc_add(sc_mul(k.l1, xx), sc_mul(k.l2, x2)), sc_mul(k.l3, x3));
      b[at] = sc_add(sc_add(k.r0, sc_mul(k.r1, xx)), sc_mul(k.r3, x3));
    } else {
      a[at] = sc_zero();
      b[at] = sc_neg(in.ypow[at]);
    }
  }
};
// t_x, t_x_blinding, e_blinding (A.3 steps 11-13) -> proof bytes 352..447
struct KProverScalars {
  static constexpr int kBlock = 64, kMinBlocks = 1;
  static constexpr const char *kName = "KProverScalars";
  const scm *t;      // [6][B]: t1..t6
  const scm *tb;     // [5][B]: blindings of T_1,T_3,T_4,T_5,T_6
  const scm *blind;  // [3][B]: i,o,s blindings
  const scm *wV, *vbl; int m, B; const scm *x; uint8_t *proofs; long proof_stride;
  HD void operator()(long p) const {
    scm xx = x[p];
    scm t2b = sc_zero();
    for (int j = 0; j < m; j++) t2b = sc_add(t2b, sc_mul(wV[(long)j * B + p], vbl[(long)j * B + p]));
    scm xp = xx, tx = sc_zero(), txb = sc_zero();
    const int bidx[6] = {0, -1, 1, 2, 3, 4};
    for (int j = 0; j < 6; j++) {
      tx = sc_add(tx, sc_mul(t[(long)j * B + p], xp));
      scm bl = bidx[j] < 0 ? t2b : tb[(long)bidx[j] * B + p];
      txb = sc_add(txb, sc_mul(bl, xp));
      xp = sc_mul(xp, xx);
    }
    scm e = sc_mul(xx, sc_add(blind[p], sc_mul(xx, sc_add(blind[B + p], sc_mul(xx, blind[2L * B + p])))));
    uint8_t *pf = proofs + p * proof_stride + 352;
    sc_tobytes(pf, tx); sc_tobytes(pf + 32, txb); sc_tobytes(pf + 64, e);
  }
};

// c_L = <a_lo, b_hi>, c_R = <a_hi, b_lo> partials
struct KIpaDots {
  static constexpr int kBlock = 128, kMinBlocks = 1;
  static constexpr const char *kName = "KIpaDots";
  const scm *a, *b; int h, B, CH; scm *part;  // part[(c*2+j)*B+p]
  HD void operator()(long tid) const {
    int p = (int)(tid % B); int c = (int)(tid / B);
    int i0 = c * CH, i1 = i0 + CH < h ? i0 + CH : h;
    scm cl = sc_zero(), cr = cl;
    for (int i = i0; i < i1; i++) {
      long lo = (long)i * B + p, hi = (long)(h + i) * B + p;
      cl = sc_add(cl, sc_mul(a[lo], b[hi]));
      cr = sc_add(cr, sc_mul(a[hi], b[lo]));
    }
    part[((long)c * 2) * B + p] = cl; part[((long)c * 2 + 1) * B + p] = cr;
  }
};
// digit rows of the L and R multiscalar multiplications of one round (see KTsIpaRound for the scaling)
//   L rows: [0,h) alpha*a[i]*gf(h+i) on Gt[h+i];  [h,2h) beta*y^-i*b[h+i]*gf(i) on Ht[i];  row 2h: c_L on Q
//   R rows: [0,h) alpha*a[h+i]*gf(i) on Gt[i];    [h,2h) beta*y^-(h+i)*b[i]*gf(h+i) on Ht[h+i]; row 2h: c_R on Q
struct KRecodeIpa {
  static constexpr int kBlock = 128, kMinBlocks = 1;
  static constexpr const char *kName = "KRecodeIpa";
  const scm *a, *b, *alpha, *beta, *yinvpow, *ufac, *clr; int h, B, n, round;
  int8_t *digL, *digR; long inst_stride; int yfree;
  HD void operator()(long tid) const {
    int p = (int)(tid % B); int i = (int)(tid / B);
    long lo = (long)i * B + p, hi = (long)(h + i) * B + p;
    scm al = alpha[p], be = beta[p];
    scm gf_lo = sc_one(), gf_hi = sc_one();
    if (round == 0) { if (i >= n) gf_lo = ufac[p]; if (h + i >= n) gf_hi = ufac[p]; }
    int8_t d[32];
    int8_t *L = digL + (long)p * inst_stride, *R = digR + (long)p * inst_stride;
    sc_recode_bytes(d, sc_mul(sc_mul(al, a[lo]), gf_hi)); store_digits(L + (long)i * 32, d);
    scm ylo = yfree ? sc_one() : yinvpow[lo], yhi = yfree ? sc_one() : yinvpow[hi];
    sc_recode_bytes(d, sc_mul(sc_mul(sc_mul(be, ylo), b[hi]), gf_lo)); store_digits(L + (long)(h + i) * 32, d);
    sc_recode_bytes(d, sc_mul(sc_mul(al, a[hi]), gf_lo)); store_digits(R + (long)i * 32, d);
    sc_recode_bytes(d, sc_mul(sc_mul(sc_mul(be, yhi), b[lo]), gf_hi)); store_digits(R + (long)(h + i) * 32, d);
    if (i == 0) {
      sc_recode_bytes(d, clr[p]); store_digits(L + (long)2 * h * 32, d);
      sc_recode_bytes(d, clr[B + p]); store_digits(R + (long)2 * h * 32, d);
    }
  }
};
// a' = a_lo*u + a_hi*u^-1 ; b' = b_lo*u^-1 + b_hi*u   (in place on the low halves)
struct KFoldAB {
  static constexpr int kBlock = 128, kMinBlocks = 1;
  static constexpr const char *kName = "KFoldAB";
  scm *a, *b; const scm *u, *uinv; int h, B;
  HD void operator()(long tid) const {
    int p = (int)(tid % B); int i = (int)(tid / B);
    long lo = (long)i * B + p, hi = (long)(h + i) * B + p;
    scm uu = u[p], ui = uinv[p];
    a[lo] = sc_add(sc_mul(a[lo], uu), sc_mul(a[hi], ui));
    b[lo] = sc_add(sc_mul(b[lo], ui), sc_mul(b[hi], uu));
  }
};
struct KStoreAB {
  static constexpr int kBlock = 64, kMinBlocks = 1;
  static constexpr const char *kName = "KStoreAB";
  const scm *a, *b; uint8_t *proofs; long proof_stride; long off;
  HD void operator()(long p) const { uint8_t *pf = proofs + p * proof_stride + off; sc_tobytes(pf, a[p]); sc_tobytes(pf + 32, b[p]); }
};

// ------------------------------------------------------------------------------------------------
// generator fold: dst[p][i] = c * lo[i] + d * hi[i], (c, d) given as two NAFs shared by all i of a proof and class (KTsIpaRound):
// one doubling chain for both digit strings (Straus), thread order proof-major so a warp walks the same digits (uniform branches).
// ------------------------------------------------------------------------------------------------
struct KFoldGens {
  static constexpr int kBlock = 128, kMinBlocks = BP_OCC_FOLD;
  static constexpr const char *kName = "KFoldGens";
  const ge_p3 *srcG, *srcH; long src_stride;  // round 0: shared generators (stride 0); later: per-proof
  ge_p3 *dstG, *dstH; long dst_stride;
  const int8_t *naf; const int *naf_top; int h, n, round;
  // odd multiples 1, 3, 5, 7 of a point (one doubling, three additions)
  HD static void odd_multiples(ge_cached c[4], const ge_p3 &pt) {
    ge_p3 p2, t; ge_cached c2;
    ge_dbl_f(p2, pt); ge_to_cached<true>(c2, p2);
    ge_to_cached<true>(c[0], pt);
    ge_add_cached_f(t, pt, c2, 0); ge_to_cached<true>(c[1], t);
    ge_add_cached_f(t, t, c2, 0); ge_to_cached<true>(c[2], t);
    ge_add_cached_f(t, t, c2, 0); ge_to_cached<true>(c[3], t);
  }
  HD void operator()(long tid) const {
    long p = tid / (2 * h); int r = (int)(tid % (2 * h)); int which = r / h; int i = r % h;
    const ge_p3 *src = (which ? srcH : srcG) + p * src_stride;
    ge_p3 *dst = (which ? dstH : dstG) + p * dst_stride;
    int cls = (round == 0 && h + i >= n) ? 1 : 0;
    const long slot = (p * 4 + which * 2 + cls) * 2;
    const int8_t *nfc = naf + slot * 256, *nfd = nfc + 256;
    const int topc = naf_top[slot], topd = naf_top[slot + 1];
    const int top = topc > topd ? topc : topd;
    ge_p3 lo, hi; load_struct(lo, &src[i]); load_struct(hi, &src[h + i]);
    ge_cached c[8];  // [0..4): odd multiples of lo, [4..8): of hi
    odd_multiples(c, lo); odd_multiples(c + 4, hi);
    ge_p3 acc; ge_identity(acc);
    for (int bit = top; bit >= 0; bit--) {
      const int dc = bit <= topc ? nfc[bit] : 0, dd = bit <= topd ? nfd[bit] : 0;
      // the doubling is the hot operation: expanded in place, ONE site; T is only needed when an addition follows or at the end
      if (bit != top) {
        ge_p1p1 t; ge_dbl_p1p1<true>(t, acc);
        ge_p1p1_to_p2<true>(acc, t);
        if (dc != 0 || dd != 0 || bit == 0) fe_mul_x<true>(acc.T, t.X, t.Y);
      }
      if (dc != 0) { const int neg = dc < 0; ge_add_cached_f(acc, acc, c[((neg ? -dc : dc) - 1) >> 1], neg); }
      if (dd != 0) { const int neg = dd < 0; ge_add_cached_f(acc, acc, c[4 + (((neg ? -dd : dd) - 1) >> 1)], neg); }
    }
    store_struct(&dst[i], acc);
  }
};

// ------------------------------------------------------------------------------------------------
// set-up kernels (BulletproofGens::new / PedersenGens::default; outside any timed region)
// ------------------------------------------------------------------------------------------------
// 64 uniform bytes per generator (SHAKE256 stream squeezed on the host) -> ristretto one-way map
struct KGensFromUniform {
  static constexpr int kBlock = 64, kMinBlocks = 1;
  static constexpr const char *kName = "KGensFromUniform";
  const uint8_t *uniform; ge_p3 *out_p3; ge_niels *out_niels;
  HD void operator()(long i) const {
    ge_p3 p, q; ristretto_from_uniform(p, uniform + i * 64);
    ge_normalize(q, p);
    store_struct(&out_p3[i], q);
    ge_niels nl; ge_to_niels(nl, q);
    store_struct(&out_niels[i], nl);
  }
};
// pc[0] = B (decoded from the canonical basepoint encoding), pc[1] = B_blinding (from uniform bytes)
struct KPcBases {
  static constexpr int kBlock = 32, kMinBlocks = 1;
  static constexpr const char *kName = "KPcBases";
  const uint8_t *basepoint_c; const uint8_t *bb_uniform; ge_p3 *pc; ge_niels *pc_niels; uint8_t *pc_c; int *ok;
  HD void operator()(long i) const {
    ge_p3 p, q;
    if (i == 0) { if (!ristretto_decode(p, basepoint_c)) *ok = 0; }
    else ristretto_from_uniform(p, bb_uniform);
    ge_normalize(q, p);
    store_struct(&pc[i], q);
    ge_niels nl; ge_to_niels(nl, q); store_struct(&pc_niels[i], nl);
    ristretto_encode(pc_c + 32 * i, q);
  }
};
// table[(base*64 + win)*16 + d] = d * 16^win * pc[base]
struct KPcTable {
  static constexpr int kBlock = 32, kMinBlocks = 1;
  static constexpr const char *kName = "KPcTable";
  const ge_p3 *pc; ge_niels *table;
  HD void operator()(long tid) const {
    int base = (int)(tid / 64), win = (int)(tid % 64);
    ge_p3 P; load_struct(P, &pc[base]);
    for (int i = 0; i < 4 * win; i++) ge_dbl(P, P);
    ge_p3 acc = P;
    ge_niels id; ge_niels_identity(id);
    store_struct(&table[(base * 64 + win) * 16], id);
    for (int d = 1; d < 16; d++) {
      ge_niels nl; ge_to_niels(nl, acc); store_struct(&table[(base * 64 + win) * 16 + d], nl);
      ge_add(acc, acc, P);
    }
  }
};
struct KEncodePoints {
  static constexpr int kBlock = 64, kMinBlocks = 1;
  static constexpr const char *kName = "KEncodePoints";
  const ge_p3 *pts; uint8_t *out;
  HD void operator()(long i) const { ge_p3 p; load_struct(p, &pts[i]); ristretto_encode(out + 32 * i, p); }
};

// ------------------------------------------------------------------------------------------------
// witness generation: tape interpreter, one thread per proof.
// Each multiplier i has (opL,argL,opR,argR):  W_LC evaluates witness linear combination #arg over
// already-assigned variables (cs.multiply / evaluate_lc), W_INV_L takes the inverse of this
// multiplier's left value (synthesize_inverse_sbox, reference src/gadget_poseidon.rs:160-166),
// W_AUX reads per-proof auxiliary input #arg (allocate_multiplier with caller-computed values,
// reference src/r1cs_utils.rs:29-32).
// ------------------------------------------------------------------------------------------------
enum { W_LC = 0, W_INV_L = 1, W_AUX = 2, W_POSEIDON = 3, W_SKIP = 4 };
struct TapeOp { uint8_t opL, opR, pad[2]; uint32_t argL, argR; };
struct WitnessLcs { const uint32_t *ptr; const uint8_t *kind; const uint32_t *idx; const scm *coeff; };
// Block op: one whole Poseidon permutation (reference src/gadget_poseidon.rs:282-399) evaluated natively on the
// 6-lane state instead of through its ~12k-term linear combinations; fills the multipliers of all its S-boxes in
// circuit order (inverse S-box: (x, 1/x, x/x), (x, 0, 0), (x, 1/x, x/x); cube: (x, x, x^2), (x^2, x, x^3)).
struct PoseidonBlock { uint32_t in_lc[6]; uint32_t sbox; uint32_t first_mult; };
struct PoseidonDev { const scm *round_keys; const scm *mds; uint32_t full_b, partial, full_e; };
#define POSEIDON_WIDTH 6
// keeps a loop rolled on the device so that its body (an inlined Montgomery multiplication) exists once in the instruction stream
#if defined(__CUDA_ARCH__)
#define POSEIDON_ROLLED _Pragma("unroll 1")
#else
#define POSEIDON_ROLLED
#endif

struct KWitnessTape {
  static constexpr int kBlock = 128, kMinBlocks = 1;  // see KRngDraw
  static constexpr const char *kName = "KWitnessTape";
  const TapeOp *tape; WitnessLcs lcs; const PoseidonBlock *pblocks; PoseidonDev pos; int n, B;
  const scm *v; const scm *aux; const scm *pub; scm *aL, *aR, *aO;
  HD scm eval(uint32_t lc, int p) const {
    scm acc = sc_zero();
    for (uint32_t t = lcs.ptr[lc]; t < lcs.ptr[lc + 1]; t++) {
      scm c = lcs.coeff[t]; long at = (long)lcs.idx[t] * B + p;
      const scm *src;  // one multiplication site for every variable kind (instruction-cache footprint)
      switch (lcs.kind[t]) {
        case 0: src = v; break;
        case 1: src = aL; break;
        case 2: src = aR; break;
        case 3: src = aO; break;
        case 5: src = pub; break;
        default: src = nullptr; break;
      }
      acc = sc_add(acc, src ? sc_mul(c, src[at]) : c);
    }
    return acc;
  }
  HD void put(long i, int p, const scm &l, const scm &r, const scm &o) const { long at = i * B + p; aL[at] = l; aR[at] = r; aO[at] = o; }
  HD long sbox_out(long at, int p, scm &x, int sbox) const {  // writes the S-box multipliers, replaces x by the S-box output
    if (sbox == 0) {
      scm sq = sc_sqr(x), cu = sc_mul(sq, x);
      put(at, p, x, x, sq); put(at + 1, p, sq, x, cu);
      x = cu; return at + 2;
    }
    scm inv = sc_invert(x), o = sc_mul(x, inv);
    put(at, p, x, inv, o); put(at + 1, p, x, sc_zero(), sc_zero()); put(at + 2, p, x, inv, o);
    x = inv; return at + 3;
  }
  HD void poseidon(const PoseidonBlock &blk, int p) const {
    scm st[POSEIDON_WIDTH];
    for (int i = 0; i < POSEIDON_WIDTH; i++) st[i] = eval(blk.in_lc[i], p);
    long at = blk.first_mult; uint32_t off = 0;
    const uint32_t total = pos.full_b + pos.partial + pos.full_e;
    for (uint32_t rnd = 0; rnd < total; rnd++) {
      const bool full = rnd < pos.full_b || rnd >= pos.full_b + pos.partial;
      for (int i = 0; i < POSEIDON_WIDTH; i++) st[i] = sc_add(st[i], pos.round_keys[off + i]);
      off += POSEIDON_WIDTH;
      if (blk.sbox == 1) {
        // the inversions of a round (six lanes of a full round, the last lane of a partial one) with ONE field inversion
        // (Montgomery's trick); zero inputs map to zero as Scalar::invert does
        const int first = full ? 0 : POSEIDON_WIDTH - 1;
        scm x[POSEIDON_WIDTH], pre[POSEIDON_WIDTH], acc = sc_one();
        POSEIDON_ROLLED
        for (int i = first; i < POSEIDON_WIDTH; i++) { x[i] = sc_is_zero(st[i]) ? sc_one() : st[i]; pre[i] = acc; acc = sc_mul(acc, x[i]); }
        scm inv = sc_invert(acc);
        POSEIDON_ROLLED
        for (int i = POSEIDON_WIDTH - 1; i >= first; i--) {
          scm xi = sc_mul(inv, pre[i]); inv = sc_mul(inv, x[i]);
          if (sc_is_zero(st[i])) xi = sc_zero();
          x[i] = xi;
        }
        POSEIDON_ROLLED
        for (int i = first; i < POSEIDON_WIDTH; i++) {
          scm o = sc_is_zero(st[i]) ? sc_zero() : sc_one();
          put(at, p, st[i], x[i], o); put(at + 1, p, st[i], sc_zero(), sc_zero()); put(at + 2, p, st[i], x[i], o);
          at += 3; st[i] = x[i];
        }
      } else {
        POSEIDON_ROLLED
        for (int i = full ? 0 : POSEIDON_WIDTH - 1; i < POSEIDON_WIDTH; i++) at = sbox_out(at, p, st[i], 0);
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
  HD void operator()(long p_) const {
    int p = (int)p_;
    for (int i = 0; i < n; i++) {
      TapeOp op = tape[i];
      if (op.opL == W_SKIP) continue;
      if (op.opL == W_POSEIDON) { poseidon(pblocks[op.argL], p); continue; }
      scm l, r;
      if (op.opL == W_LC) l = eval(op.argL, p); else l = aux[(long)op.argL * B + p];
      if (op.opR == W_LC) r = eval(op.argR, p); else if (op.opR == W_INV_L) r = sc_invert(l); else r = aux[(long)op.argR * B + p];
      put(i, p, l, r, sc_mul(l, r));
    }
  }
};

// ------------------------------------------------------------------------------------------------
// primitive self-tests: run one primitive on the device so the test-suite can compare it with the
// oracle through the C-ABI (bp_selftest_device).  in/out are small device byte buffers.
// ------------------------------------------------------------------------------------------------
enum { ST_MERLIN = 0, ST_SC_INVERT = 1, ST_SC_WIDE = 2, ST_RISTRETTO_ROUNDTRIP = 3, ST_SC_MUL = 4, ST_FROM_UNIFORM = 5, ST_RNG = 6, ST_KECCAK = 7, ST_TSSTART = 8 };
struct KSelfTest {
  static constexpr int kBlock = 32, kMinBlocks = 1;
  static constexpr const char *kName = "KSelfTest";
  int which; const uint8_t *in; int in_len; uint8_t *out; int out_len;
  HD void operator()(long) const {
    switch (which) {
      case ST_MERLIN: {  // Transcript::new(in[0..a)); append_message("some label", in[a..)); challenge_bytes("challenge", out)
        int a = in[0];
        strobe128 t; ts_init(t, in + 1, a);
        ts_append(t, "some label", in + 1 + a, in_len - 1 - a);
        ts_challenge_bytes(t, "challenge", out, out_len);
        break;
      }
      case ST_SC_INVERT: sc_tobytes(out, sc_invert(sc_from_bytes_mod_order(in))); break;
      case ST_SC_WIDE: sc_tobytes(out, sc_from_bytes_wide(in)); break;
      case ST_RISTRETTO_ROUNDTRIP: { ge_p3 p; int ok = ristretto_decode(p, in); out[32] = (uint8_t)ok; if (ok) { ge_p3 d; ge_dbl(d, p); ge_sub(d, d, p); ristretto_encode(out, d); } break; }
      case ST_SC_MUL: sc_tobytes(out, sc_mul(sc_from_bytes_mod_order(in), sc_from_bytes_mod_order(in + 32))); break;
      case ST_FROM_UNIFORM: { ge_p3 p; ristretto_from_uniform(p, in); ristretto_encode(out, p); break; }
      case ST_RNG: {  // transcript "rngtest" -> build rng with witness in[0..32), entropy in[32..64) -> out_len/32 scalars
        const uint8_t lbl[7] = {'r', 'n', 'g', 't', 'e', 's', 't'};
        strobe128 t; ts_init(t, lbl, 7);
        trng_rekey(t, "v_blinding", in, 32); trng_finalize(t, in + 32);
        for (int i = 0; i < out_len / 32; i++) { uint8_t b[64]; trng_fill(t, b, 64); sc_tobytes(out + 32 * i, sc_from_bytes_wide(b)); }
        break;
      }
      case ST_KECCAK: {  // keccak-f on 200 bytes
        uint64_t st[25];
        for (int i = 0; i < 25; i++) { uint64_t x = 0; for (int j = 7; j >= 0; j--) x = (x << 8) | in[8 * i + j]; st[i] = x; }
        keccak_f1600(st);
        for (int i = 0; i < 200; i++) out[i] = st_get(st, i);
        break;
      }
      case ST_TSSTART: {  // replica of KTsStart for m = 2: in = V0 V1 vbl0 vbl1 entropy (5 x 32); out = ts state (208) | rng state (208) | 2 draws
        const uint8_t lbl[4] = {'M', 'i', 'M', 'C'};
        strobe128 t; ts_init(t, lbl, 4);
        const uint8_t r1[7] = {'r', '1', 'c', 's', ' ', 'v', '1'};
        ts_append(t, "dom-sep", r1, 7);
        for (int j = 0; j < 2; j++) ts_append(t, "V", in + 32 * j, 32);
        ts_append_u64(t, "m", 2);
        for (int i = 0; i < 200; i++) out[i] = st_get(t.st, i);
        out[200] = t.pos; out[201] = t.pos_begin; out[202] = t.cur_flags;
        for (int j = 0; j < 2; j++) { uint8_t b[32]; sc_tobytes(b, sc_from_bytes_mod_order(in + 64 + 32 * j)); trng_rekey(t, "v_blinding", b, 32); }
        trng_finalize(t, in + 128);
        for (int i = 0; i < 200; i++) out[208 + i] = st_get(t.st, i);
        out[408] = t.pos; out[409] = t.pos_begin; out[410] = t.cur_flags;
        for (int i = 0; i < 2; i++) { uint8_t b[64]; trng_fill(t, b, 64); sc_tobytes(out + 416 + 32 * i, sc_from_bytes_wide(b)); }
        break;
      }
      default: break;
    }
  }
};

// ------------------------------------------------------------------------------------------------
// verifier: transcript replay, verification scalars, one multiscalar check per proof
// ------------------------------------------------------------------------------------------------
// Whole transcript of one verification in one thread: every proof element is known up front.
// chal layout [..][B]: 0 y, 1 z, 2 y^-1, 3 u, 4 x, 5 w, 6 r ; ipa challenges u_j at uj[j*B+p], inverses at ujinv.
struct KTsVerify {
  static constexpr int kBlock = 64, kMinBlocks = 1;
  static constexpr const char *kName = "KTsVerify";
  strobe128 base; const uint8_t *V; int m, B, k; unsigned N; const uint8_t *proofs; long proof_stride; const uint8_t *entropy;
  scm *chal; scm *uj, *ujinv; int *status;
  // combined (cross-proof) mode: digest[p] = 32 bytes bound to EVERYTHING the verification of proof p reads -- its whole
  // transcript (V, every proof element), the inner-product scalars a and b (which the transcript never absorbs), the
  // public inputs and the verifier's entropy.  The batching weights are derived from the digests of ALL proofs (KBatchSeed,
  // KVerifyRho), so no weight can be known before every proof of the batch is fixed.
  uint8_t *digest; const uint8_t *pub; int npub;
  HD static int is_zero32(const uint8_t *b) { uint8_t nz = 0; for (int i = 0; i < 32; i++) nz |= b[i]; return nz == 0; }
  HD void operator()(long p) const {
    strobe128 t; strobe_load(t, &base);
    int st = 0;
    for (int j = 0; j < m; j++) ts_append(t, "V", V + ((long)p * m + j) * 32, 32);
    ts_append_u64(t, "m", (uint64_t)m);
    const uint8_t *pf = proofs + p * proof_stride;
    if (is_zero32(pf) || is_zero32(pf + 32) || is_zero32(pf + 64)) st = BP_ERR_VERIFICATION_;
    ts_append(t, "A_I1", pf, 32); ts_append(t, "A_O1", pf + 32, 32); ts_append(t, "S1", pf + 64, 32);
    const uint8_t ph[11] = {'r', '1', 'c', 's', '-', '1', 'p', 'h', 'a', 's', 'e'};
    ts_append(t, "dom-sep", ph, 11);
    ts_append(t, "A_I2", pf + 96, 32); ts_append(t, "A_O2", pf + 128, 32); ts_append(t, "S2", pf + 160, 32);
    scm y, z, u, x, w, r;
    TS_CHALLENGE(t, "y", y); TS_CHALLENGE(t, "z", z);
    for (int j = 0; j < 5; j++) if (is_zero32(pf + 192 + 32 * j)) st = BP_ERR_VERIFICATION_;
    ts_append(t, "T_1", pf + 192, 32); ts_append(t, "T_3", pf + 224, 32); ts_append(t, "T_4", pf + 256, 32);
    ts_append(t, "T_5", pf + 288, 32); ts_append(t, "T_6", pf + 320, 32);
    TS_CHALLENGE(t, "u", u); TS_CHALLENGE(t, "x", x);
    ts_append(t, "t_x", pf + 352, 32); ts_append(t, "t_x_blinding", pf + 384, 32); ts_append(t, "e_blinding", pf + 416, 32);
    TS_CHALLENGE(t, "w", w);
    const uint8_t ipp[6] = {'i', 'p', 'p', ' ', 'v', '1'};
    ts_append(t, "dom-sep", ipp, 6);
    ts_append_u64(t, "n", (uint64_t)N);
    for (int j = 0; j < k; j++) {
      const uint8_t *lr = pf + 448 + 64 * j;
      if (is_zero32(lr) || is_zero32(lr + 32)) st = BP_ERR_VERIFICATION_;
      ts_append(t, "L", lr, 32); ts_append(t, "R", lr + 32, 32);
      scm c; TS_CHALLENGE(t, "u", c);
      uj[(long)j * B + p] = c; ujinv[(long)j * B + p] = sc_invert(c);
    }
    // scalars of the proof must be canonical (R1CSProof::from_bytes -> FormatError)
    scm tmp;
    if (!sc_from_canonical_bytes(tmp, pf + 352) || !sc_from_canonical_bytes(tmp, pf + 384) || !sc_from_canonical_bytes(tmp, pf + 416) ||
        !sc_from_canonical_bytes(tmp, pf + 448 + 64 * k) || !sc_from_canonical_bytes(tmp, pf + 480 + 64 * k)) st = BP_ERR_FORMAT_;
    trng_finalize(t, entropy + p * 32);
    { uint8_t b[64]; trng_fill(t, b, 64); r = sc_from_bytes_wide(b); }
    chal[p] = y; chal[B + p] = z; chal[2L * B + p] = sc_invert(y); chal[3L * B + p] = u; chal[4L * B + p] = x; chal[5L * B + p] = w; chal[6L * B + p] = r;
    if (digest) {
      ts_append(t, "ipp-a", pf + 448 + 64 * k, 32); ts_append(t, "ipp-b", pf + 480 + 64 * k, 32);
      for (int j = 0; j < npub; j++) ts_append(t, "pub", pub + ((long)p * npub + j) * 32, 32);
      ts_challenge_bytes(t, "batch-digest", digest + p * 32, 32);
    }
    if (st) status[p] = st;
  }
};
// Commitments whose value the circuit fixes (the Poseidon "statics" 0, 101, 0, 0 committed with blinding 0): the reference's
// verifier computes them itself (allocate_statics_for_verifier, src/gadget_poseidon.rs:580-608); a batch verifier takes V from
// the caller, so it compares those slots with the bytes recorded when the circuit was built.  thread = (fixed slot, proof)
struct KCheckFixedCommitments {
  static constexpr int kBlock = 128, kMinBlocks = 1;
  static constexpr const char *kName = "KCheckFixedCommitments";
  const uint8_t *V; int m, B; const uint32_t *idx; const uint8_t *expect; int nfixed; int *status;
  HD void operator()(long tid) const {
    const long p = tid / nfixed; const int j = (int)(tid % nfixed);
    const uint8_t *got = V + (p * m + idx[j]) * 32, *want = expect + 32 * j;
    uint8_t diff = 0;
    for (int i = 0; i < 32; i++) diff |= got[i] ^ want[i];
    if (diff) status[p] = BP_ERR_VERIFICATION_;
  }
};
// s_i = prod_j u_j^(+1 if bit (k-1-j) of i else -1)
struct KVerifyS {
  static constexpr int kBlock = 128, kMinBlocks = 1;
  static constexpr const char *kName = "KVerifyS";
  const scm *uj, *ujinv; int k, B; scm *s;
  HD void operator()(long tid) const {
    int p = (int)(tid % B); long i = tid / B;
    scm acc = sc_one();
    for (int j = 0; j < k; j++) acc = sc_mul(acc, ((i >> (k - 1 - j)) & 1) ? uj[(long)j * B + p] : ujinv[(long)j * B + p]);
    s[i * B + p] = acc;
  }
};
// delta partial sums: sum_{i<n} y^-i * wR_i * wL_i
struct KVerifyDelta {
  static constexpr int kBlock = 128, kMinBlocks = 1;
  static constexpr const char *kName = "KVerifyDelta";
  const scm *wL, *wR, *yinvpow; int n, B, CH; scm *part;
  HD void operator()(long tid) const {
    int p = (int)(tid % B); int c = (int)(tid / B);
    int i0 = c * CH, i1 = i0 + CH < n ? i0 + CH : n;
    scm acc = sc_zero();
    for (int i = i0; i < i1; i++) { long at = (long)i * B + p; acc = sc_add(acc, sc_mul(sc_mul(yinvpow[at], wR[at]), wL[at])); }
    part[(long)c * B + p] = acc;
  }
};
// digit rows of the G and H scalars: rows [2, 2+N) and [2+N, 2+2N)
struct KVerifyGH {
  static constexpr int kBlock = 128, kMinBlocks = 1;
  static constexpr const char *kName = "KVerifyGH";
  const scm *wL, *wR, *wO, *yinvpow, *s, *chal; const uint8_t *proofs; long proof_stride; int n, N, k, B; int8_t *dig; long inst_stride;
  int8_t *wideG, *wideH;  // non-NULL: 15-bit digit rows for the sorted-bucket path, [B][N + 1] rows each (row 0 = B / B_blinding)
  // combined (cross-proof) mode: launched over round_up(B, 32) proofs per row; the rho-weighted scalars of 32 consecutive
  // proofs are summed (warp shuffles) into partG / partH [chunk][N]
  const scm *rho; scm *partG, *partH;
  HD void operator()(long tid) const {
    if (rho) { combined(tid); return; }
    int p = (int)(tid % B); long i = tid / B; long at = i * B + p;
    scm x = chal[4L * B + p], u = chal[3L * B + p];
    const uint8_t *pf = proofs + p * proof_stride + 448 + 64 * k;
    scm a = sc_from_bytes_mod_order(pf), b = sc_from_bytes_mod_order(pf + 32);
    scm yi = yinvpow[at];
    scm g = sc_neg(sc_mul(a, s[at]));
    scm hh = sc_neg(sc_mul(b, s[(long)(N - 1 - i) * B + p]));
    if (i < n) {
      g = sc_add(g, sc_mul(x, sc_mul(yi, wR[at])));
      hh = sc_add(hh, sc_add(sc_mul(x, wL[at]), wO[at]));
    }
    hh = sc_sub(sc_mul(yi, hh), sc_one());
    if (i >= n) { g = sc_mul(g, u); hh = sc_mul(hh, u); }
    if (wideG) {
      int16_t dw[SB_WINDOWS];
      sc_recode13(dw, g); store_digits13(wideG + ((long)p * (N + 1) + 1 + i) * SB_ROW_BYTES, dw);
      sc_recode13(dw, hh); store_digits13(wideH + ((long)p * (N + 1) + 1 + i) * SB_ROW_BYTES, dw);
      return;
    }
    int8_t d[32];
    int8_t *row = dig + (long)p * inst_stride;
    sc_recode_bytes(d, g); store_digits(row + (2 + i) * 32, d);
    sc_recode_bytes(d, hh); store_digits(row + (2 + N + i) * 32, d);
  }
  HD void gh(long i, int p, scm &g, scm &hh) const {
    long at = i * B + p;
    scm x = chal[4L * B + p], u = chal[3L * B + p];
    const uint8_t *pf = proofs + p * proof_stride + 448 + 64 * k;
    scm a = sc_from_bytes_mod_order(pf), b = sc_from_bytes_mod_order(pf + 32);
    scm yi = yinvpow[at];
    g = sc_neg(sc_mul(a, s[at]));
    hh = sc_neg(sc_mul(b, s[(long)(N - 1 - i) * B + p]));
    if (i < n) {
      g = sc_add(g, sc_mul(x, sc_mul(yi, wR[at])));
      hh = sc_add(hh, sc_add(sc_mul(x, wL[at]), wO[at]));
    }
    hh = sc_sub(sc_mul(yi, hh), sc_one());
    if (i >= n) { g = sc_mul(g, u); hh = sc_mul(hh, u); }
  }
  HD void combined(long tid) const {
    const int Bp = (B + 31) & ~31;
    int p = (int)(tid % Bp); long i = tid / Bp;
    scm g = sc_zero(), hh = sc_zero();
    if (p < B) { gh(i, p, g, hh); scm r = rho[p]; g = sc_mul(g, r); hh = sc_mul(hh, r); }
    const long slot = (long)(p >> 5) * N + i;
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      scm tg, th;
#pragma unroll
      for (int j = 0; j < 4; j++) { tg.v[j] = __shfl_down_sync(0xffffffffu, g.v[j], o); th.v[j] = __shfl_down_sync(0xffffffffu, hh.v[j], o); }
      g = sc_add(g, tg); hh = sc_add(hh, th);
    }
    if ((p & 31) == 0) { partG[slot] = g; partH[slot] = hh; }
#else
    // host emulation runs the "threads" one after the other: accumulate in place (buffers zeroed by the caller)
    partG[slot] = sc_add(partG[slot], g); partH[slot] = sc_add(partH[slot], hh);
#endif
  }
};
// combined mode: rows of the single cross-proof MSM = sums over the proof chunks; rows 2N, 2N+1 (B, B_blinding) = sums over proofs
struct KVerifyCombineRows {
  static constexpr int kBlock = 128, kMinBlocks = 1;
  static constexpr const char *kName = "KVerifyCombineRows";
  const scm *partG, *partH, *rowB, *rowBb; int N, B; scm *rows;  // rows[2N + 2]
  HD void operator()(long r) const {
    const int chunks = (B + 31) >> 5;
    scm acc = sc_zero();
    if (r < 2L * N) {
      const scm *part = r < N ? partG : partH; const long i = r < N ? r : r - N;
      for (int c = 0; c < chunks; c++) acc = sc_add(acc, part[(long)c * N + i]);
    } else {
      const scm *src = r == 2L * N ? rowB : rowBb;
      for (int p = 0; p < B; p++) acc = sc_add(acc, src[p]);
    }
    rows[r] = acc;
  }
};
// combined mode, weights.  seed = transcript over the digests of all B proofs, in order (one thread: B x 32 bytes through the
// sponge, ~B / 5 permutations); rho_p = challenge of (seed, p).  A weight therefore depends on every byte of every proof, every
// commitment, every public input and every entropy value of the batch: submitting one proof twice, reusing entropy between
// slots, or knowing the entropy does not let a prover predict or equalise weights (each rho_p is a random-oracle output of the
// finished batch).  Proofs that failed a structural check (status != 0) take no part in the combination.
struct KBatchSeed {
  static constexpr int kBlock = 32, kMinBlocks = 1;
  static constexpr const char *kName = "KBatchSeed";
  const uint8_t *digest; int B; uint8_t *seed;
  HD void operator()(long) const {
    const uint8_t lbl[24] = {'b', 'p', '-', 'b', '2', '0', '0', ' ', 'c', 'o', 'm', 'b', 'i', 'n', 'e', 'd', ' ', 'v', 'e', 'r', 'i', 'f', 'y', '2'};
    strobe128 t; ts_init(t, lbl, 24);
    ts_append_u64(t, "batch", (uint64_t)B);
    for (int p = 0; p < B; p++) ts_append(t, "digest", digest + (long)p * 32, 32);
    ts_challenge_bytes(t, "seed", seed, 32);
  }
};
struct KVerifyRho {
  static constexpr int kBlock = 64, kMinBlocks = 1;
  static constexpr const char *kName = "KVerifyRho";
  const uint8_t *seed; const int *status; scm *rho;
  HD void operator()(long p) const {
    if (status[p] != 0) { rho[p] = sc_zero(); return; }
    const uint8_t lbl[23] = {'b', 'p', '-', 'b', '2', '0', '0', ' ', 'b', 'a', 't', 'c', 'h', ' ', 'w', 'e', 'i', 'g', 'h', 't', ' ', 'v', '2'};
    strobe128 t; ts_init(t, lbl, 23);
    ts_append(t, "seed", seed, 32);
    ts_append_u64(t, "index", (uint64_t)p);
    scm r; TS_CHALLENGE(t, "rho", r);
    rho[p] = r;
  }
};
// remaining scalars: rows 0,1 (B, B_blinding) and the per-proof points after the generators:
// A_I1 A_O1 S1 A_I2 A_O2 S2 | V_0..V_{m-1} | T_1 T_3 T_4 T_5 T_6 | L_0.. | R_0..
struct KVerifyScalars {
  static constexpr int kBlock = 64, kMinBlocks = 1;
  static constexpr const char *kName = "KVerifyScalars";
  const scm *chal, *uj, *ujinv, *wV, *wc, *wP, *pub, *delta; const uint8_t *proofs; long proof_stride; int m, npub, N, k, B; int8_t *dig; long inst_stride;
  int8_t *wideG, *wideH;  // as in KVerifyGH: rows 0 (B, B_blinding) go there when set
  const scm *rho; scm *rowB, *rowBb;  // combined mode: every scalar of proof p is weighted by rho[p]; the B / B_blinding scalars go to rowB / rowBb
  HD void put(long p, int8_t *row, long r, const scm &v) const { int8_t d[32]; sc_recode_bytes(d, rho ? sc_mul(v, rho[p]) : v); store_digits(row + r * 32, d); }
  HD void putw(int8_t *base, long p, const scm &v) const { int16_t dw[SB_WINDOWS]; sc_recode13(dw, v); store_digits13(base + p * (long)(N + 1) * SB_ROW_BYTES, dw); }
  HD void operator()(long p) const {
    const uint8_t *pf = proofs + p * proof_stride;
    scm u = chal[3L * B + p], x = chal[4L * B + p], w = chal[5L * B + p], r = chal[6L * B + p];
    scm t_x = sc_from_bytes_mod_order(pf + 352), t_xb = sc_from_bytes_mod_order(pf + 384), e_b = sc_from_bytes_mod_order(pf + 416);
    scm a = sc_from_bytes_mod_order(pf + 448 + 64 * k), b = sc_from_bytes_mod_order(pf + 480 + 64 * k);
    scm xx = sc_sqr(x), xxx = sc_mul(xx, x), rxx = sc_mul(r, xx);
    int8_t *row = dig + (long)p * inst_stride;
    scm wcv = wc[p];
    for (int i = 0; i < npub; i++) wcv = sc_add(wcv, sc_mul(wP[(long)i * B + p], pub[(long)i * B + p]));
    scm bsc = sc_add(sc_mul(w, sc_sub(t_x, sc_mul(a, b))), sc_mul(r, sc_sub(sc_mul(xx, sc_add(wcv, delta[p])), t_x)));
    if (rho) { rowB[p] = sc_mul(bsc, rho[p]); rowBb[p] = sc_mul(sc_neg(sc_add(e_b, sc_mul(r, t_xb))), rho[p]); }
    else if (wideG) { putw(wideG, p, bsc); putw(wideH, p, sc_neg(sc_add(e_b, sc_mul(r, t_xb)))); }
    else { put(p, row, 0, bsc); put(p, row, 1, sc_neg(sc_add(e_b, sc_mul(r, t_xb)))); }
    long o = (wideG || rho) ? 0 : 2 + 2L * N;  // wide / combined mode: the 8-bit rows hold the per-proof points only
    put(p, row, o + 0, x); put(p, row, o + 1, xx); put(p, row, o + 2, xxx);
    put(p, row, o + 3, sc_mul(u, x)); put(p, row, o + 4, sc_mul(u, xx)); put(p, row, o + 5, sc_mul(u, xxx));
    o += 6;
    for (int j = 0; j < m; j++) put(p, row, o + j, sc_mul(wV[(long)j * B + p], rxx));
    o += m;
    scm rx = sc_mul(r, x);
    put(p, row, o, rx); put(p, row, o + 1, sc_mul(rxx, x)); put(p, row, o + 2, sc_mul(rxx, xx)); put(p, row, o + 3, sc_mul(rxx, xxx)); put(p, row, o + 4, sc_mul(sc_mul(rxx, xx), xx));
    o += 5;
    for (int j = 0; j < k; j++) { put(p, row, o + j, sc_sqr(uj[(long)j * B + p])); put(p, row, o + k + j, sc_sqr(ujinv[(long)j * B + p])); }
  }
};
// decompress the per-proof points in the order KVerifyScalars lays their scalars out
struct KVerifyDecompress {
  static constexpr int kBlock = 64, kMinBlocks = 1;
  static constexpr const char *kName = "KVerifyDecompress";
  const uint8_t *V; const uint8_t *proofs; long proof_stride; int m, k, B; ge_p3 *pts; long pts_stride; int *status;
  HD void operator()(long tid) const {
    int np = 11 + m + 2 * k;
    long p = tid / np; int j = (int)(tid % np);
    const uint8_t *pf = proofs + p * proof_stride, *src;
    if (j < 6) src = pf + 32 * j;
    else if (j < 6 + m) src = V + (p * m + (j - 6)) * 32;
    else if (j < 11 + m) src = pf + 192 + 32 * (j - 6 - m);
    else if (j < 11 + m + k) src = pf + 448 + 64 * (j - 11 - m);
    else src = pf + 448 + 64 * (j - 11 - m - k) + 32;
    ge_p3 pt;
    if (!ristretto_decode(pt, src)) { status[p] = BP_ERR_VERIFICATION_; ge_identity(pt); }
    store_struct(&pts[p * pts_stride + j], pt);
  }
};

// ------------------------------------------------------------------------------------------------
// Fixed-base tables.  All generators (G_0.., H_0.., B, B_blinding) are shared by every proof and never change, so for
// each generator P, window w and digit magnitude e+1 the point (e+1)*2^(8w)*P is precomputed in affine Niels form:
//   table[(gen*32 + w)*128 + e]            (128 B each; 512 KiB per generator)
// A scalar*generator term then costs <= 32 mixed additions into a REGISTER accumulator: no buckets, no bucket traffic,
// no reduction pass, and any subset of (row, window) pairs can be summed by any thread.
// Generator index space: [0,cap) = G, [cap,2cap) = H, 2cap = B, 2cap+1 = B_blinding.
// ------------------------------------------------------------------------------------------------
#define TBL_W 32
#define TBL_E 128
struct KTableBuild {
  static constexpr int kBlock = 64, kMinBlocks = 1;
  static constexpr const char *kName = "KTableBuild";
  const ge_p3 *G, *H, *pc; long cap; ge_niels *table;
  int bits = 8, W = TBL_W, E = TBL_E;  // window width, windows per generator, entries per window (the fold tables of large capacities use narrower windows)
  HD void operator()(long tid) const {
    long gen = tid / W; int w = (int)(tid % W);
    ge_p3 P;
    if (gen < cap) load_struct(P, &G[gen]); else if (gen < 2 * cap) load_struct(P, &H[gen - cap]); else load_struct(P, &pc[gen - 2 * cap]);
    for (int i = 0; i < bits * w; i++) ge_dbl(P, P);
    ge_niels *slot = table + tid * E;
    // pass 1: multiples in projective form parked in the slots, prefix products of Z kept locally
    fe pre[TBL_E];
    ge_p3 acc = P;
    fe run; fe_1(run);
    for (int e = 0; e < E; e++) {
      ge_niels tmp; tmp.ypx = acc.X; tmp.ymx = acc.Y; tmp.xy2d = acc.Z;
      store_struct(&slot[e], tmp);
      pre[e] = run;
      fe_mul(run, run, acc.Z);
      ge_add(acc, acc, P);
    }
    // pass 2: one inversion for all Z (Montgomery's trick), then normalise to affine Niels
    fe inv, d2; fe_invert(inv, run); FE_2D(d2);
    for (int e = E - 1; e >= 0; e--) {
      ge_niels t; load_struct(t, &slot[e]);
      fe zi, x, y;
      fe_mul(zi, inv, pre[e]); fe_mul(inv, inv, t.xy2d);
      fe_mul(x, t.ypx, zi); fe_mul(y, t.ymx, zi);
      ge_niels nl;
      fe_add(nl.ypx, y, x); fe_carry(nl.ypx); fe_sub(nl.ymx, y, x); fe_carry(nl.ymx);
      fe_mul(nl.xy2d, x, y); fe_mul(nl.xy2d, nl.xy2d, d2);
      store_struct(&slot[e], nl);
    }
  }
};

// row -> generator index.  mode 0: explicit map; 1/2: the L / R multiscalar multiplication of an UNFOLDED inner-product
// round over the original generators (nj = current vector length, h = nj/2, N rows of G then H, last row = B).
struct RowMap { int mode; const uint32_t *map; long cap, N, nj, h; long inst_off; long pad_gen; int rel = 0; };  // inst_off: generator offset per instance (split MSM); pad_gen: generator index of row N + 1 (modes 1, 2)
HD long row_gen(const RowMap &m, long r) {
  if (m.mode == 0) return m.map[r];
  if (m.mode == 3) return r;  // (global) row r is generator r; sub-instance `inst` of a split MSM covers rows inst * inst_off ..
  if (m.mode == 4) return r == 0 ? 2 * m.cap + m.nj : m.nj * m.cap + (r - 1);  // verifier half nj (0: B, G_i; 1: B_blinding, H_i)
  if (m.mode == 5) return r < m.N ? r : (r < 2 * m.N ? m.cap + (r - m.N) : 2 * m.cap + (r - 2 * m.N));  // combined check: G_0.., H_0.., B, B_blinding
  if (r == m.N) return 2 * m.cap;  // B
  if (r == m.N + 1) return m.pad_gen;  // sum of the padding generators (round 0, see KRecodeUnfolded13)
  const long half = m.N / 2;
  const bool isH = r >= half;
  const long rr = isH ? r - half : r, blk = rr / m.h, i = rr % m.h;
  // L takes G_hi and H_lo, R takes G_lo and H_hi
  const bool hi = (m.mode == 1) != isH;
  return (isH ? m.cap : 0) + blk * m.nj + (hi ? m.h : 0) + i;
}
// generator index the sort writes into the items of instance `inst`, row r.  rel (mode 3 only, split MSM): relative to the
// instance's first generator inst * inst_off -- KBucketAccumulate adds the base back (SortedView::base_stride) -- so that the items of
// a 2^22-row MSM still fit the 24 index bits of the two-pass sort
HD long row_gen_item(const RowMap &m, long r, long inst) { return m.rel ? r : row_gen(m, r + inst * m.inst_off); }
// table-driven multiscalar multiplication: one thread per (instance, split); partial[tid] = sum over its rows
struct KMsmTable {
  static constexpr int kBlock = 128, kMinBlocks = BP_OCC_TABLE;
  static constexpr const char *kName = "KMsmTable";
  const ge_niels *table; RowMap rmap; const int8_t *dig; long dig_inst_stride; long rows; int S; ge_p3 *partial;
  HD void operator()(long tid) const {
    long inst = tid / S; int sp = (int)(tid % S);
    const long r0 = rows * sp / S, r1 = rows * (sp + 1) / S;
    const int8_t *drow = dig + inst * dig_inst_stride;
    ge_p3 acc; ge_identity(acc);
    for (long r = r0; r < r1; r++) {
      int8_t d[32];
#if defined(__CUDA_ARCH__)
      { const uint4 *src = reinterpret_cast<const uint4 *>(drow + r * 32); uint4 a = src[0], b = src[1]; memcpy(d, &a, 16); memcpy(d + 16, &b, 16); }
#else
      memcpy(d, drow + r * 32, 32);
#endif
      const ge_niels *tg = table + row_gen(rmap, r) * (long)(TBL_W * TBL_E);
#pragma unroll 1
      for (int w = 0; w < TBL_W; w++) {  // rolled: ONE addition site, its field multiplications expanded in place
        int dv = d[w];
        if (dv != 0) {
          int neg = dv < 0; int e = (neg ? -dv : dv) - 1;
          ge_niels q; load_struct(q, &tg[w * TBL_E + e]);
          ge_madd<true>(acc, acc, q, neg);
        }
      }
    }
    store_struct(&partial[tid], acc);
  }
};
// first reduction stage for many splits: thread (inst, j) adds partials j, j+R, j+2R, ... into out[inst*R + j]
struct KMsmTableReduce {
  static constexpr int kBlock = 64, kMinBlocks = 1;
  static constexpr const char *kName = "KMsmTableReduce";
  const ge_p3 *partial; int S, R; ge_p3 *out;
  HD void operator()(long tid) const {
    long inst = tid / R; int j = (int)(tid % R);
    ge_p3 acc; load_struct(acc, &partial[inst * S + j]);
    for (int s = j + R; s < S; s += R) { ge_p3 t; load_struct(t, &partial[inst * S + s]); ge_add(acc, acc, t); }
    store_struct(&out[tid], acc);
  }
};
struct KMsmTableFinish {
  static constexpr int kBlock = 64, kMinBlocks = 1;
  static constexpr const char *kName = "KMsmTableFinish";
  const ge_p3 *partial; int S; uint8_t *out; long out_stride;
  HD void operator()(long inst) const {
    ge_p3 acc; load_struct(acc, &partial[inst * S]);
    for (int s = 1; s < S; s++) { ge_p3 t; load_struct(t, &partial[inst * S + s]); ge_add(acc, acc, t); }
    ristretto_encode(out + inst * out_stride, acc);
  }
};

// verifier: (bucket-method window sums of the per-proof points) + (table partial sums of the generator rows) must be the identity
struct KVerifyCheck {
  static constexpr int kBlock = 64, kMinBlocks = 1;
  static constexpr const char *kName = "KVerifyCheck";
  const ge_p3 *wsum; const ge_p3 *tpart; int S; int *status; long tpart_s_stride;  // partial s of instance i at tpart[i*S + s], or [s*stride + i] when stride != 0
  HD void operator()(long inst) const {
    ge_p3 acc; load_struct(acc, &wsum[inst * MSM_WINDOWS + MSM_WINDOWS - 1]);
    for (int w = MSM_WINDOWS - 2; w >= 0; w--) {
      for (int i = 0; i < 7; i++) ge_dbl_p2(acc, acc);
      ge_dbl(acc, acc);
      ge_p3 sp; load_struct(sp, &wsum[inst * MSM_WINDOWS + w]);
      ge_add(acc, acc, sp);
    }
    for (int s = 0; s < S; s++) { ge_p3 t; load_struct(t, tpart_s_stride ? &tpart[s * tpart_s_stride + inst] : &tpart[inst * S + s]); ge_add(acc, acc, t); }
    if (!ge_is_identity_ristretto(acc) && status[inst] == 0) status[inst] = BP_ERR_VERIFICATION_;
  }
};

// combined (cross-proof) verification: per-proof point part P_p = Horner over the window sums of proof p
struct KVerifyProofPoint {
  static constexpr int kBlock = 64, kMinBlocks = 1;
  static constexpr const char *kName = "KVerifyProofPoint";
  const ge_p3 *wsum; ge_p3 *out;
  HD void operator()(long inst) const {
    ge_p3 acc; load_struct(acc, &wsum[inst * MSM_WINDOWS + MSM_WINDOWS - 1]);
    for (int w = MSM_WINDOWS - 2; w >= 0; w--) {
      for (int i = 0; i < 7; i++) ge_dbl_p2(acc, acc);
      ge_dbl(acc, acc);
      ge_p3 sp; load_struct(sp, &wsum[inst * MSM_WINDOWS + w]);
      ge_add(acc, acc, sp);
    }
    store_struct(&out[inst], acc);
  }
};
// out[t] = sum of pts[t], pts[t + T], ...   (T threads)
struct KSumPointsStrided {
  static constexpr int kBlock = 64, kMinBlocks = 1;
  static constexpr const char *kName = "KSumPointsStrided";
  const ge_p3 *pts; long count, T; ge_p3 *out;
  HD void operator()(long t) const {
    ge_p3 acc; ge_identity(acc);
    for (long j = t; j < count; j += T) { ge_p3 q; load_struct(q, &pts[j]); ge_add(acc, acc, q); }
    store_struct(&out[t], acc);
  }
};
// the one identity test of a combined verification: sum of `count` points (per-proof parts and the shared-generator part)
struct KVerifyCombinedCheck {
  static constexpr int kBlock = 32, kMinBlocks = 1;
  static constexpr const char *kName = "KVerifyCombinedCheck";
  const ge_p3 *pts; int count; int *combined;
  HD void operator()(long) const {
    ge_p3 acc; ge_identity(acc);
    for (int i = 0; i < count; i++) { ge_p3 t; load_struct(t, &pts[i]); ge_add(acc, acc, t); }
    *combined = ge_is_identity_ristretto(acc) ? 0 : BP_ERR_VERIFICATION_;
  }
};

// per-round coefficient tables of the unfolded rounds: UG[b] = prod_t u_t^(+1 if bit t of b else -1), UH[b] = its inverse pattern,
// b in [0, 2^(j+1)) after round j (round t <-> bit (j-t) of b, i.e. round 0 is the most significant bit).
struct KIpaUTable {
  static constexpr int kBlock = 128, kMinBlocks = 1;
  static constexpr const char *kName = "KIpaUTable";
  const scm *u, *uinv; const scm *UGin, *UHin; scm *UGout, *UHout; int B;  // in: [2^j][B], out: [2^(j+1)][B]
  HD void operator()(long tid) const {
    int p = (int)(tid % B); long b = tid / B;
    scm uu = u[p], ui = uinv[p];
    scm g = UGin[(b >> 1) * B + p], h = UHin[(b >> 1) * B + p];
    UGout[b * B + p] = sc_mul(g, (b & 1) ? uu : ui);
    UHout[b * B + p] = sc_mul(h, (b & 1) ? ui : uu);
  }
};
// digit rows of L and R for an unfolded round (layout of RowMap modes 1 / 2); thread = (row pair index, proof)
//   G rows: UG[b]*gf(idx)*a[partner],  H rows: UH[b]*y^-idx*gf(idx)*b[partner],  last row: c*w on B
struct KRecodeUnfolded {
  static constexpr int kBlock = 128, kMinBlocks = 1;
  static constexpr const char *kName = "KRecodeUnfolded";
  const scm *a, *b, *UG, *UH, *yinvpow, *ufac, *clr, *w; long N, nj, h, n; int B; int8_t *digL, *digR; long inst_stride;
  HD void operator()(long tid) const {
    int p = (int)(tid % B); long rr = tid / B;  // rr in [0, N/2): block blk, offset i
    const long blk = rr / h, i = rr % h;
    const long lo = blk * nj + i, hi = lo + h;            // original generator indices of the low / high half entries
    scm ug = UG[blk * B + p], uh = UH[blk * B + p];
    scm gf_lo = lo >= n ? ufac[p] : sc_one(), gf_hi = hi >= n ? ufac[p] : sc_one();
    scm a_lo = a[i * B + p], a_hi = a[(h + i) * B + p], b_lo = b[i * B + p], b_hi = b[(h + i) * B + p];
    int8_t d[32];
    int8_t *L = digL + (long)p * inst_stride, *R = digR + (long)p * inst_stride;
    const long half = N / 2;
    // L: G_hi with a_lo ; H_lo with b_hi.   R: G_lo with a_hi ; H_hi with b_lo.
    sc_recode_bytes(d, sc_mul(sc_mul(ug, gf_hi), a_lo)); store_digits(L + rr * 32, d);
    sc_recode_bytes(d, sc_mul(sc_mul(sc_mul(uh, yinvpow[lo * B + p]), gf_lo), b_hi)); store_digits(L + (half + rr) * 32, d);
    sc_recode_bytes(d, sc_mul(sc_mul(ug, gf_lo), a_hi)); store_digits(R + rr * 32, d);
    sc_recode_bytes(d, sc_mul(sc_mul(sc_mul(uh, yinvpow[hi * B + p]), gf_hi), b_lo)); store_digits(R + (half + rr) * 32, d);
    if (rr == 0) {
      sc_recode_bytes(d, sc_mul(clr[p], w[p])); store_digits(L + N * 32, d);
      sc_recode_bytes(d, sc_mul(clr[B + p], w[p])); store_digits(R + N * 32, d);
    }
  }
};
// digit rows for materialising the folded generators at level J: row idx (G) = UG[b]*gf(idx) / UG[0], row N+idx (H) = UH[b]*y^-idx*gf(idx).
// The G side is kept un-normalised by the factor UG[0] (it becomes alpha, see KTsIpaRound): block 0 of every G output then carries the
// scalar 1 -- its row is all zeros here and KFoldTable adds the generator itself, one addition instead of 32.  (The H side has no
// common factor: its block-0 scalar UH[0]*y^-i depends on the output index.)  UG[0]^-1 = UH[0] (KIpaUTable).
struct KRecodeFoldTable {
  static constexpr int kBlock = 128, kMinBlocks = 1;
  static constexpr const char *kName = "KRecodeFoldTable";
  const scm *UG, *UH, *yinvpow, *ufac; long N, nJ, n; int B; int8_t *dig; long inst_stride;
  int bits = 8, W = TBL_W, rb = 32;  // window width, digits per row, bytes per digit row (multiple of 16)
  HD void put(int8_t *dst, const scm &v) const {
    if (bits == 8) { int8_t d[32]; sc_recode_bytes(d, v); store_digits(dst, d); return; }
    int8_t d[64];
    sc_recode_win(d, v, bits, W);
    for (int i = W; i < rb; i++) d[i] = 0;
    for (int i = 0; i < rb; i += 32) store_digits(dst + i, d + i);
  }
  HD void operator()(long tid) const {
    int p = (int)(tid % B); long idx = tid / B;
    const long blk = idx / nJ;
    scm gf = idx >= n ? ufac[p] : sc_one();
    int8_t *row = dig + (long)p * inst_stride;
    put(row + idx * rb, blk == 0 ? sc_zero() : sc_mul(sc_mul(UG[blk * B + p], gf), UH[p]));
    put(row + (N + idx) * rb, sc_mul(sc_mul(UH[blk * B + p], yinvpow[idx * B + p]), gf));
  }
};
// G_J[i] = G_i + sum_{b>0} (row b*nJ+i) * G_{b*nJ+i} (un-normalised, see KRecodeFoldTable), H_J[i] = sum_b (row N+b*nJ+i) * H_{b*nJ+i}:
// the folded generators after J rounds, straight from the tables.  table[((which*cap + idx) * W + w) * E + e] = (e+1) * 2^(bits*w) * P
// (the 8-bit direct tables of a generator set, or its narrower fold tables when those do not fit: ensure_fold_table in engine.cu)
struct KFoldTable {
  static constexpr int kBlock = 128, kMinBlocks = BP_OCC_TABLE;
  static constexpr const char *kName = "KFoldTable";
  const ge_niels *table; long cap, N, nJ; const int8_t *dig; long dig_inst_stride; ge_p3 *dstG, *dstH; long dst_stride; const ge_p3 *G;
  int W = TBL_W, E = TBL_E, rb = 32;
  HD void operator()(long tid) const {
    long p = tid / (2 * nJ); long r = tid % (2 * nJ); int which = (int)(r / nJ); long i = r % nJ;
    const int8_t *drow = dig + p * dig_inst_stride + (which ? N * (long)rb : 0);
    ge_p3 acc;
    if (which) ge_identity(acc); else load_struct(acc, &G[i]);  // G side: block 0 carries the scalar 1 (see KRecodeFoldTable)
    const long terms = (N / nJ) * W;
    long idx = i; int w = 0;
#pragma unroll 1
    for (long t = 0; t < terms; t++) {  // one flat loop over (block, window): ONE addition site, field multiplications expanded in place
      const int dv = drow[idx * rb + w];
      if (dv != 0) {
        int neg = dv < 0; int e = (neg ? -dv : dv) - 1;
        ge_niels q; load_struct(q, &table[(((which ? cap : 0) + idx) * W + w) * (long)E + e]);
        ge_madd<true>(acc, acc, q, neg);
      }
      if (++w == W) { w = 0; idx += nJ; }
    }
    store_struct(&(which ? dstH : dstG)[p * dst_stride + i], acc);
  }
};

// ------------------------------------------------------------------------------------------------
// Sorted-bucket multiscalar multiplication over the shared generators (13-bit signed windows).
// For every generator P and window w the point 2^(SB_BITS w)*P is tabulated (shift table, SB_WINDOWS x 96 B per generator),
// so all windows of an instance feed ONE set of SB_BUCKETS buckets: 17 additions per term (15-bit windows) instead of 32
// with the 8-bit direct tables, and one running-sum reduction per instance instead of one per window.  Per launch:
//   sort_buckets (block per instance, counting sort in shared memory)  ->  item list grouped by bucket
//   KBucketAccumulate (thread per 128-item segment of the sorted list: register accumulator, partial sums at bucket borders)
//   KBucketReduce (groups of 128 buckets: plain and weighted sums)  ->  KBucketFinishA / B (combine, encode)
// ------------------------------------------------------------------------------------------------
// shift table: sg[gen*20 + w] = 2^(13w) * P_gen in affine Niels form
struct KShiftTableBuild {
  static constexpr int kBlock = 64, kMinBlocks = 1;
  static constexpr const char *kName = "KShiftTableBuild";
  const ge_p3 *G, *H, *pc; long cap; ge_niels *sg;
  HD void operator()(long gen) const {
    ge_p3 P;
    if (gen < cap) load_struct(P, &G[gen]); else if (gen < 2 * cap) load_struct(P, &H[gen - cap]); else load_struct(P, &pc[gen - 2 * cap]);
    for (int w = 0; w < SB_WINDOWS; w++) {
      ge_niels nl; ge_to_niels(nl, P); store_struct(&sg[gen * SB_WINDOWS + w], nl);
      for (int i = 0; i < SB_BITS; i++) ge_dbl(P, P);
    }
  }
};
// shift-table rows of ONE extra point (a circuit's sum of padding generators) at generator slot `gen`
struct KShiftTableOne {
  static constexpr int kBlock = 32, kMinBlocks = 1;
  static constexpr const char *kName = "KShiftTableOne";
  const ge_p3 *pt; long gen; ge_niels *sg;
  HD void operator()(long) const {
    ge_p3 P; load_struct(P, pt);
    for (int w = 0; w < SB_WINDOWS; w++) {
      ge_niels nl; ge_to_niels(nl, P); store_struct(&sg[gen * SB_WINDOWS + w], nl);
      for (int i = 0; i < SB_BITS; i++) ge_dbl(P, P);
    }
  }
};
// src [cnt][B] (optionally times mul[p]) -> 13-bit digit rows row0.. of each instance
struct KRecode13 {
  static constexpr int kBlock = 128, kMinBlocks = 1;
  static constexpr const char *kName = "KRecode13";
  const scm *src; const scm *mul; int cnt, B; int8_t *dig; long inst_stride; int row0;
  const uint8_t *skip;  // optional, indexed by row: rows merged into another row's generator sum get all-zero digits
  HD void operator()(long tid) const {
    int p = (int)(tid % B); long i = tid / B;
    int16_t d[SB_WINDOWS];
    if (skip && skip[row0 + i]) {
#pragma unroll
      for (int j = 0; j < SB_WINDOWS; j++) d[j] = 0;
    } else {
      scm s = src[i * B + p];
      if (mul) s = sc_mul(s, mul[p]);
      sc_recode13(d, s);
    }
    store_digits13(dig + (long)p * inst_stride + (row0 + i) * SB_ROW_BYTES, d);
  }
};
// shift-table rows of generator SUMS: group j = sum of `cnt[j]` generators src[j][0..] (indices into the G | H space) at slot
// slot0 + j.  Used for multipliers whose values are equal by construction (the three left wires and the two non-zero right
// wires of an inverse S-box), so that A_I pays one row per group.
struct KMergeGens {
  static constexpr int kBlock = 64, kMinBlocks = 1;
  static constexpr const char *kName = "KMergeGens";
  const ge_p3 *G, *H; long cap; const uint32_t *src; long slot0; ge_niels *sg;  // src[3*j + {0,1,2}], 0xffffffff = unused
  HD void operator()(long j) const {
    ge_p3 P; ge_identity(P);
    for (int t = 0; t < 3; t++) {
      const uint32_t gidx = src[3 * j + t];
      if (gidx == 0xffffffffu) continue;
      ge_p3 Q; load_struct(Q, gidx < cap ? &G[gidx] : &H[gidx - cap]);
      ge_add(P, P, Q);
    }
    for (int w = 0; w < SB_WINDOWS; w++) {
      ge_niels nl; ge_to_niels(nl, P); store_struct(&sg[(slot0 + j) * SB_WINDOWS + w], nl);
      for (int i = 0; i < SB_BITS; i++) ge_dbl(P, P);
    }
  }
};
// same scalars as KRecodeUnfolded, 13-bit rows
struct KRecodeUnfolded13 {
  static constexpr int kBlock = 128, kMinBlocks = 1;
  static constexpr const char *kName = "KRecodeUnfolded13";
  const scm *a, *b, *UG, *UH, *yinvpow, *ufac, *clr, *w; long N, nj, h, n; int B; int8_t *digL, *digR; long inst_stride;
  // Round 0 of a padded circuit (n < N): the padded entries of r are -y^i, so every H-row of L whose partner index h + i is
  // >= n carries the SAME scalar y^-i * (-y^(h+i)) = -y^h.  With pad_rows set those N - n rows are left empty and one extra
  // row N + 1 holds -y^h for the precomputed point sum_{i = n-h}^{h-1} H_i (ypow_h = y^h per proof).
  int pad_rows; const scm *ypow_h;
  HD void put(int8_t *base, long row, const scm &v) const { int16_t d[SB_WINDOWS]; sc_recode13(d, v); store_digits13(base + row * SB_ROW_BYTES, d); }
  HD void put_zero(int8_t *base, long row) const { int16_t d[SB_WINDOWS]; for (int i = 0; i < SB_WINDOWS; i++) d[i] = 0; store_digits13(base + row * SB_ROW_BYTES, d); }
  HD void operator()(long tid) const {
    int p = (int)(tid % B); long rr = tid / B;
    const long blk = rr / h, i = rr % h;
    const long lo = blk * nj + i, hi = lo + h;
    scm ug = UG[blk * B + p], uh = UH[blk * B + p];
    scm gf_lo = lo >= n ? ufac[p] : sc_one(), gf_hi = hi >= n ? ufac[p] : sc_one();
    scm a_lo = a[i * B + p], a_hi = a[(h + i) * B + p], b_lo = b[i * B + p], b_hi = b[(h + i) * B + p];
    int8_t *L = digL + (long)p * inst_stride, *R = digR + (long)p * inst_stride;
    const long half = N / 2;
    put(L, rr, sc_mul(sc_mul(ug, gf_hi), a_lo));
    if (pad_rows && hi >= n) put_zero(L, half + rr);
    else put(L, half + rr, sc_mul(sc_mul(sc_mul(uh, yinvpow[lo * B + p]), gf_lo), b_hi));
    put(R, rr, sc_mul(sc_mul(ug, gf_lo), a_hi));
    put(R, half + rr, sc_mul(sc_mul(sc_mul(uh, yinvpow[hi * B + p]), gf_hi), b_lo));
    if (rr == 0) {
      put(L, N, sc_mul(clr[p], w[p])); put(R, N, sc_mul(clr[B + p], w[p]));
      if (pad_rows) { put(L, N + 1, sc_neg(ypow_h[p])); put_zero(R, N + 1); }
    }
  }
};
// item = (generator*20 + window) | sign << 31
// boff: [inst][SB_BUCKETS + 1] item offsets; soff: [inst][SB_BUCKETS + 1] slice offsets (a bucket with c items is cut into
// ceil(c / SB_SLICE) slices so that no thread adds more than SB_SLICE points: padded circuits put thousands of identical
// scalars -- hence identical digits -- into a handful of buckets)
#define SB_SLICE 256
struct SortedView { const uint32_t *items; const uint32_t *boff; const uint32_t *soff; long items_stride; long slices_cap; long base_stride = 0; };  // base_stride: shift-table entries between the first generators of consecutive instances (RowMap::rel)
// reference (one thread per instance) counting sort: used by the emulation build and as the fallback for tiny launches
struct KSortBucketsSerial {
  static constexpr int kBlock = 32, kMinBlocks = 1;
  static constexpr const char *kName = "KSortBucketsSerial";
  RowMap rmap; const int8_t *dig; long dig_inst_stride; long rows; uint32_t *items; long items_stride; uint32_t *boff; uint32_t *soff;
  HD void operator()(long inst) const {
    uint32_t *off = boff + inst * (SB_BUCKETS + 1);
    for (int b = 0; b <= SB_BUCKETS; b++) off[b] = 0;
    const int8_t *drow = dig + inst * dig_inst_stride;
    for (long r = 0; r < rows; r++) {
      int16_t d[24]; load_digits13(d, drow + r * SB_ROW_BYTES);
      for (int w = 0; w < SB_WINDOWS; w++) if (d[w]) off[(d[w] < 0 ? -d[w] : d[w])]++;   // count into slot b+1
    }
    for (int b = 0; b < SB_BUCKETS; b++) off[b + 1] += off[b];
    // scatter, walking a cursor per bucket (cursor = off[b] advanced, restored afterwards)
    uint32_t *it = items + inst * items_stride;
    for (long r = 0; r < rows; r++) {
      int16_t d[24]; load_digits13(d, drow + r * SB_ROW_BYTES);
      const uint32_t g = (uint32_t)row_gen_item(rmap, r, inst) * SB_WINDOWS;
      for (int w = 0; w < SB_WINDOWS; w++) if (d[w]) {
        int neg = d[w] < 0; int b = (neg ? -d[w] : d[w]) - 1;
        it[off[b]++] = (g + w) | ((uint32_t)neg << 31);
      }
    }
    for (int b = SB_BUCKETS; b > 0; b--) off[b] = off[b - 1];
    off[0] = 0;
    uint32_t *so = soff + inst * (SB_BUCKETS + 1);
    so[0] = 0;
    for (int b = 0; b < SB_BUCKETS; b++) so[b + 1] = so[b] + (off[b + 1] - off[b] + SB_SLICE - 1) / SB_SLICE;
  }
};
// One thread per SEGMENT of SB_SEG consecutive items of the sorted list (not per bucket): every thread performs the same number
// of additions, whatever the bucket sizes (one-thread-per-bucket lost ~20 % of the lanes to the Poisson spread of bucket
// sizes, and needed a slice table for buckets holding thousands of identical digits).  A segment that crosses a bucket
// boundary stores its partial sum and starts a new one: partial (bucket b, segment s) lives at psum[b + s], which is
// unique because buckets and segments are both ordered along the list.
#ifndef SB_SEG
#define SB_SEG 128
#endif
#ifndef BP_OCC_SORTED
#define BP_OCC_SORTED 4
#endif
#ifndef BP_TMA_STAGE
// 1: table entries staged in shared memory by per-lane cp.async.bulk copies.  MEASURED SLOWER (KBucketAccumulate 934 -> 1135 ms
// per 2731-proof chunk, profiles/r02_tma_entry_staging_negative_result.json): UBLKCP is a warp-uniform instruction, so a per-lane
// 96-byte gather becomes an ELECT loop of ~9 instructions per lane and item.  Kept for the record; the bulk-copy engine is used
// where one thread moves a whole tile (the digit tiles of the bucket sort, k_sorted.cu).
#define BP_TMA_STAGE 0
#endif
template <bool REL>  // REL: item indices are relative to the instance's first generator (split MSM over chain G; a pointer held in registers)
struct KBucketAccumulateT {
  static constexpr int kBlock = 128, kMinBlocks = BP_OCC_SORTED;
  static constexpr const char *kName = REL ? "KBucketAccumulateRel" : "KBucketAccumulate";
  const ge_niels *sg0; SortedView sv; ge_p3 *psum; long segs_cap;  // psum[inst*slices_cap + b + s]
  HD void operator()(long tid) const {
    const long inst = tid / segs_cap; const uint32_t s = (uint32_t)(tid % segs_cap);
    const ge_niels *sg = REL ? sg0 + inst * sv.base_stride : sg0;
    const uint32_t *off = sv.boff + inst * (SB_BUCKETS + 1);
    const uint32_t total = off[SB_BUCKETS];
    const uint32_t k0 = s * SB_SEG;
#if defined(__CUDA_ARCH__) && BP_TMA_STAGE
    staged(inst, s, off, total, k0);
#else
    if (k0 >= total) return;
    const uint32_t k1 = k0 + SB_SEG < total ? k0 + SB_SEG : total;
    int lo = 0, hi = SB_BUCKETS;  // largest b with off[b] <= k0 (then off[b + 1] > k0)
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (off[mid] <= k0) lo = mid; else hi = mid; }
    uint32_t b = (uint32_t)lo, next = off[b + 1];
    const uint32_t *it = sv.items + inst * sv.items_stride;
    ge_p3 *ps = psum + inst * sv.slices_cap + s;
    ge_p3 acc; ge_identity(acc);
    // software pipeline: the (random, mostly L2-missing) table entry of item k+1 is in flight during addition k
    uint32_t item = it[k0];
    ge_niels q; load_struct(q, &sg[item & 0x7fffffffu]);
    for (uint32_t k = k0; k < k1; k++) {
      if (k == next) {
        store_struct(&ps[b], acc); ge_identity(acc);
        do { b++; next = off[b + 1]; } while (next == k);
      }
      const uint32_t cur = item; const ge_niels qc = q;
      if (k + 1 < k1) { item = it[k + 1]; load_struct(q, &sg[item & 0x7fffffffu]); }
      ge_madd<true>(acc, acc, qc, (int)(cur >> 31));  // the kernel's one addition site: field multiplications expanded in place
    }
    store_struct(&ps[b], acc);
#endif
  }
#if defined(__CUDA_ARCH__) && BP_TMA_STAGE
  // Table entries staged in shared memory by the copy engine.  Each thread's next two entries (96 bytes each, at random rows of
  // the shift table) are fetched by cp.async.bulk into a two-stage ring in shared memory and completion is tracked by one
  // mbarrier per warp and stage; the addition of item k runs while the entries of items k+1 and k+2 are in flight.  Against the
  // register-load form: the prefetch needs no registers (24 fewer live across the addition), it is two additions deep instead
  // of one, and the LSU issues one bulk request per entry instead of six 16-byte loads.  Every lane of a warp makes exactly
  // SB_SEG passes (lanes past the end of their list only arrive at the barriers), so barrier phases stay in step.
  __device__ __forceinline__ void staged(long inst, uint32_t s, const uint32_t *off, uint32_t total, uint32_t k0) const {
    const ge_niels *sg = REL ? sg0 + inst * sv.base_stride : sg0;
    __shared__ alignas(128) ge_niels stage[2][kBlock];
    __shared__ alignas(8) uint64_t bars[2][kBlock / 32];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned mask = __activemask();
    const uint32_t bar0 = smem_addr32(&bars[0][warp]), bar1 = smem_addr32(&bars[1][warp]);
    const uint32_t slot0 = smem_addr32(&stage[0][threadIdx.x]), slot1 = smem_addr32(&stage[1][threadIdx.x]);
    if (lane == (unsigned)(__ffs(mask) - 1)) { mbar_init(bar0, __popc(mask)); mbar_init(bar1, __popc(mask)); mbar_fence_init(); }
    __syncwarp(mask);
    const uint32_t k1 = k0 >= total ? k0 : (k0 + SB_SEG < total ? k0 + SB_SEG : total);
    uint32_t b = 0, next = 0;
    if (k1 > k0) {
      int lo = 0, hi = SB_BUCKETS;  // largest b with off[b] <= k0 (then off[b + 1] > k0)
      while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (off[mid] <= k0) lo = mid; else hi = mid; }
      b = (uint32_t)lo; next = off[b + 1];
    }
    const uint32_t *it = sv.items + inst * sv.items_stride;
    ge_p3 *ps = psum + inst * sv.slices_cap + s;
    ge_p3 acc; ge_identity(acc);
    uint32_t item0 = 0, item1 = 0;  // items whose entries sit in stage 0 / 1
    if (k0 < k1) { item0 = it[k0]; mbar_arrive_expect_tx(bar0, (uint32_t)sizeof(ge_niels)); bulk_copy_g2s(slot0, &sg[item0 & 0x7fffffffu], (uint32_t)sizeof(ge_niels), bar0); }
    else mbar_arrive(bar0);
    if (k0 + 1 < k1) { item1 = it[k0 + 1]; mbar_arrive_expect_tx(bar1, (uint32_t)sizeof(ge_niels)); bulk_copy_g2s(slot1, &sg[item1 & 0x7fffffffu], (uint32_t)sizeof(ge_niels), bar1); }
    else mbar_arrive(bar1);
#pragma unroll 1
    for (uint32_t j = 0; j < SB_SEG; j++) {
      const uint32_t k = k0 + j, st = j & 1;
      const uint32_t bar = st ? bar1 : bar0, slot = st ? slot1 : slot0;
      mbar_wait(bar, (j >> 1) & 1);
      const uint32_t cur = st ? item1 : item0;
      ge_niels qc;
      if (k < k1) qc = stage[st][threadIdx.x];  // shared memory: plain copy (load_struct is a global-memory load)
      __syncwarp(mask);
      fence_proxy_async_smem();  // the slot is about to be rewritten by the copy engine
      if (k + 2 < k1) {
        const uint32_t nx = it[k + 2];
        if (st) item1 = nx; else item0 = nx;
        mbar_arrive_expect_tx(bar, (uint32_t)sizeof(ge_niels));
        bulk_copy_g2s(slot, &sg[nx & 0x7fffffffu], (uint32_t)sizeof(ge_niels), bar);
      } else {
        mbar_arrive(bar);
      }
      if (k < k1) {
        if (k == next) {
          store_struct(&ps[b], acc); ge_identity(acc);
          do { b++; next = off[b + 1]; } while (next == k);
        }
        ge_madd<true>(acc, acc, qc, (int)(cur >> 31));  // the kernel's one addition site: field multiplications expanded in place
      }
    }
    if (k1 > k0) store_struct(&ps[b], acc);
  }
#endif
};
using KBucketAccumulate = KBucketAccumulateT<false>;
using KBucketAccumulateRel = KBucketAccumulateT<true>;
// per group of 128 buckets: S = sum b_i, W = sum (local index + 1) * b_i   (running-sum trick); b_i = sum of its partials
#ifndef BP_OCC_REDUCE
#define BP_OCC_REDUCE 8   // 128 registers (292 B of spills) against 176: 16 warps per SM instead of 10 hide the chain of dependent additions (430 -> 384 ms per step)
#endif
struct KBucketReduce {
  static constexpr int kBlock = 64, kMinBlocks = BP_OCC_REDUCE;
  static constexpr const char *kName = "KBucketReduce";
  const ge_p3 *psum; SortedView sv; ge_p3 *seg;  // seg[(inst*32 + s)*2 + {0: S, 1: W}]
  HD void operator()(long tid) const {
    long inst = tid / SB_SEGS; int sgm = (int)(tid % SB_SEGS);
    const uint32_t *off = sv.boff + inst * (SB_BUCKETS + 1);
    const ge_p3 *ps = psum + inst * sv.slices_cap;
    ge_p3 run, tot; ge_identity(run); ge_identity(tot);
    for (int i = SB_SEG_LEN - 1; i >= 0; i--) {
      const int b = sgm * SB_SEG_LEN + i;
      const uint32_t o0 = off[b], o1 = off[b + 1];
      if (o1 > o0) {
        const uint32_t s0 = o0 / SB_SEG, s1 = (o1 - 1) / SB_SEG;
        for (uint32_t sj = s0; sj <= s1; sj++) { ge_p3 t; load_struct(t, &ps[b + sj]); ge_add_f(run, run, t); }
      }
      ge_add_f(tot, tot, run);
    }
    store_struct(&seg[tid * 2], run); store_struct(&seg[tid * 2 + 1], tot);
  }
};
// result = sum_s W_s + 128 * sum_s s * S_s, in two steps: one thread per instance would run 380 point operations one after the
// other, and with a few thousand instances per launch that chain is the launch's whole duration (round 1's KBucketFinish).  Step A: thread (inst, h) folds 16
// segments into Sum_h = sum S, Floc_h = sum j * S_{16h+j}, Wsum_h = sum W; step B: result = sum_h Wsum_h + 128 * (sum_h Floc_h +
// 16 * sum_h h * Sum_h) -- ~50 + ~45 operations deep.
#define SB_FIN_GROUP 16
#define SB_FIN_GROUPS (SB_SEGS / SB_FIN_GROUP)
struct KBucketFinishA {
  static constexpr int kBlock = 64, kMinBlocks = 1;
  static constexpr const char *kName = "KBucketFinishA";
  const ge_p3 *seg; ge_p3 *grp;  // grp[(inst*SB_FIN_GROUPS + h)*3 + {0: Sum, 1: Floc, 2: Wsum}]
  HD void operator()(long tid) const {
    const long inst = tid / SB_FIN_GROUPS; const int h = (int)(tid % SB_FIN_GROUPS);
    const ge_p3 *sgp = seg + (inst * SB_SEGS + (long)h * SB_FIN_GROUP) * 2;
    ge_p3 run, tot, ws; ge_identity(run); ge_identity(tot); ge_identity(ws);
    for (int j = SB_FIN_GROUP - 1; j >= 0; j--) {
      ge_p3 t; load_struct(t, &sgp[j * 2]); ge_add_f(run, run, t);
      if (j > 0) ge_add_f(tot, tot, run);
      load_struct(t, &sgp[j * 2 + 1]); ge_add_f(ws, ws, t);
    }
    store_struct(&grp[tid * 3], run); store_struct(&grp[tid * 3 + 1], tot); store_struct(&grp[tid * 3 + 2], ws);
  }
};
struct KBucketFinishB {
  static constexpr int kBlock = 64, kMinBlocks = 1;
  static constexpr const char *kName = "KBucketFinishB";
  const ge_p3 *grp; uint8_t *out; long out_stride; ge_p3 *out_p3;  // out_p3 != NULL: keep the point
  HD void operator()(long inst) const {
    const ge_p3 *gp = grp + inst * SB_FIN_GROUPS * 3;
    ge_p3 run, T, fl, ws; ge_identity(run); ge_identity(T); ge_identity(fl); ge_identity(ws);
    for (int h = SB_FIN_GROUPS - 1; h >= 0; h--) {
      ge_p3 t;
      if (h > 0) { load_struct(t, &gp[h * 3]); ge_add_f(run, run, t); ge_add_f(T, T, run); }
      load_struct(t, &gp[h * 3 + 1]); ge_add_f(fl, fl, t);
      load_struct(t, &gp[h * 3 + 2]); ge_add_f(ws, ws, t);
    }
    for (int i = 0; i < 4; i++) ge_dbl_f(T, T);  // * SB_FIN_GROUP (16)
    ge_add_f(T, T, fl);
    for (int i = 0; i < 7; i++) ge_dbl_f(T, T);  // * SB_SEG_LEN (128)
    ge_add_f(T, T, ws);
    if (out_p3) store_struct(&out_p3[inst], T); else ristretto_encode(out + inst * out_stride, T);
  }
};
// Few instances (one large MSM split into at most 128 sub-instances): the two kernels above are latency chains -- 128 buckets x 2
// additions per thread, then 380 sequential point operations per instance -- and cost 2.4 ms whatever the size.  Here a thread
// reduces only L buckets and applies its segment's weight by a short double-and-add, so that every segment yields ONE point
//   out[inst*(SB_BUCKETS/L) + s] = sum_{j<L} (s*L + j + 1) * bucket_{s*L+j}
// and the rest is a plain sum (KSumPointsStrided stages) over all segments of all sub-instances: ~10x the additions, spread over
// 16384/L threads per instance, ~100 operations deep in all.
struct KBucketReduceW {
  static constexpr int kBlock = 64, kMinBlocks = 1;
  static constexpr const char *kName = "KBucketReduceW";
  const ge_p3 *psum; SortedView sv; ge_p3 *out; int L;
  HD void operator()(long tid) const {
    const long nseg = SB_BUCKETS / L;
    long inst = tid / nseg; int sgm = (int)(tid % nseg);
    const uint32_t *off = sv.boff + inst * (SB_BUCKETS + 1);
    const ge_p3 *ps = psum + inst * sv.slices_cap;
    ge_p3 run, tot; ge_identity(run); ge_identity(tot);
    for (int i = L - 1; i >= 0; i--) {
      const int b = sgm * L + i;
      const uint32_t o0 = off[b], o1 = off[b + 1];
      if (o1 > o0) {
        const uint32_t s0 = o0 / SB_SEG, s1 = (o1 - 1) / SB_SEG;
        for (uint32_t sj = s0; sj <= s1; sj++) { ge_p3 t; load_struct(t, &ps[b + sj]); ge_add_f(run, run, t); }
      }
      ge_add_f(tot, tot, run);
    }
    const uint32_t k = (uint32_t)sgm * (uint32_t)L;  // + k * (sum of the segment's buckets)
    if (k) {
      int top = 31; while (!((k >> top) & 1)) top--;
      ge_p3 acc = run;
      for (int bit = top - 1; bit >= 0; bit--) { ge_dbl_f(acc, acc); if ((k >> bit) & 1) ge_add_f(acc, acc, run); }
      ge_add_f(tot, tot, acc);
    }
    store_struct(&out[tid], tot);
  }
};
// sum of `count` points -> ristretto encoding (last step of a split MSM; one thread)
struct KSumPointsEncode {
  static constexpr int kBlock = 32, kMinBlocks = 1;
  static constexpr const char *kName = "KSumPointsEncode";
  const ge_p3 *pts; int count; uint8_t *out;
  HD void operator()(long) const {
    ge_p3 acc; ge_identity(acc);
    for (int i = 0; i < count; i++) { ge_p3 t; load_struct(t, &pts[i]); ge_add(acc, acc, t); }
    ristretto_encode(out, acc);
  }
};
