// Batched prover / verifier pipelines.  See engine.h and kernels.h.
#include "engine.h"
#include <algorithm>
#include <mutex>

long g_launch_count = 0;
long engine_launch_count() { return g_launch_count; }

// ------------------------------------------------------------------------------------------------ per-kernel event timing
#ifndef BP_HOST_EMUL
#include <map>
#include <string>
int g_profile_on = 0;
struct ProfRec { const char *name; long threads; cudaEvent_t e0, e1; };
static std::vector<ProfRec> g_prof;
void profile_begin(const char *name, long threads, dev_stream s) {
  ProfRec r{name, threads, nullptr, nullptr};
  cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
  cudaEventRecord(r.e0, s);
  g_prof.push_back(r);
}
void profile_end(dev_stream s) { cudaEventRecord(g_prof.back().e1, s); }
bool profile_selected(const char *name) { return strcmp(name, "KBucketAccumulate") == 0; }
// exact number of point additions of the sorted-bucket path while profiling: sum over instances of the item-list lengths
// (one item = one non-zero digit = one addition in KBucketAccumulate).  Reported as the pseudo-kernel "@sorted_items".
static unsigned long long *g_items_dev = nullptr;
static long g_items_launches = 0;
__global__ void count_items_kernel(const uint32_t *boff, long ninst, unsigned long long *out) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long v = i < ninst ? boff[i * (SB_BUCKETS + 1) + SB_BUCKETS] : 0;
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0 && v) atomicAdd(out, v);
}
static void profile_count_items(const uint32_t *boff, long ninst, dev_stream s) {
  if (!g_items_dev) { if (cudaMalloc(&g_items_dev, 8) != cudaSuccess) { g_items_dev = nullptr; return; } cudaMemsetAsync(g_items_dev, 0, 8, s); }
  count_items_kernel<<<(unsigned)((ninst + 127) / 128), 128, 0, s>>>(boff, ninst, g_items_dev);
  g_items_launches++;
}
void engine_profile_enable(int on) { g_profile_on = on; }
// writes "name launches total_ms threads\n" lines; clears the records.  Caller must have synchronised.
int engine_profile_report(char *buf, size_t cap) {
  struct Agg { long launches = 0; double ms = 0; double threads = 0; };
  std::map<std::string, Agg> agg;
  for (ProfRec &r : g_prof) {
    float ms = 0; cudaEventSynchronize(r.e1); cudaEventElapsedTime(&ms, r.e0, r.e1);
    Agg &a = agg[r.name]; a.launches++; a.ms += ms; a.threads += (double)r.threads;
    cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
  }
  g_prof.clear();
  size_t off = 0;
  if (g_items_dev && g_items_launches) {
    unsigned long long items = 0;
    cudaMemcpy(&items, g_items_dev, 8, cudaMemcpyDeviceToHost); cudaMemset(g_items_dev, 0, 8);
    int w = snprintf(buf, cap, "@sorted_items %ld 0.000000 %llu\n", g_items_launches, items);
    if (w > 0 && (size_t)w < cap) off = (size_t)w;
    g_items_launches = 0;
  }
  for (auto &kv : agg) {
    int w = snprintf(buf + off, off < cap ? cap - off : 0, "%s %ld %.6f %.0f\n", kv.first.c_str(), kv.second.launches, kv.second.ms, kv.second.threads);
    if (w < 0 || off + (size_t)w >= cap) break;
    off += (size_t)w;
  }
  return (int)off;
}
#else
// host emulation (tests): no timing, but the same report lists which kernel bodies ran and how often
#include <map>
#include <string>
int g_profile_on = 0;
static std::map<std::string, std::pair<long, double>> g_prof;
void profile_note(const char *name, long threads) { auto &r = g_prof[name]; r.first++; r.second += (double)threads; }
void engine_profile_enable(int on) { g_profile_on = on; }
int engine_profile_report(char *buf, size_t cap) {
  size_t off = 0;
  for (auto &kv : g_prof) {
    int w = snprintf(buf + off, off < cap ? cap - off : 0, "%s %ld 0.000000 %.0f\n", kv.first.c_str(), kv.second.first, kv.second.second);
    if (w < 0 || off + (size_t)w >= cap) break;
    off += (size_t)w;
  }
  g_prof.clear();
  return (int)off;
}
#endif

#define CK(x) do { int _e = (x); if (_e) { fprintf(stderr, "bp_b200: %s failed at %s:%d\n", #x, __FILE__, __LINE__); return BP_ERR_CUDA; } } while (0)

static const uint8_t BASEPOINT_C[32] = {0xe2, 0xf2, 0xae, 0x0a, 0x6a, 0xbc, 0x4e, 0x71, 0xa8, 0x84, 0xa9, 0x61, 0xc5, 0x00, 0x51, 0x5f,
                                        0x58, 0xe3, 0x0b, 0x6a, 0xa5, 0x82, 0xdd, 0x8d, 0xb6, 0xa6, 0x59, 0x45, 0xe0, 0x8d, 0x2d, 0x76};

int bp_device_init() {
#ifndef BP_HOST_EMUL
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    fprintf(stderr, "bp_b200: no CUDA device (%s); this library has no CPU path\n", e == cudaSuccess ? "count 0" : cudaGetErrorString(e));
    return BP_ERR_NO_DEVICE;
  }
#endif
  return BP_OK;
}

static const int SG_SPARE = 8;  // spare generator slots of the shift table (padding sums of up to 8 circuit shapes)
template <class T>
static int dalloc(T **p, size_t count) { return dev_malloc((void **)p, count * sizeof(T)); }

// Outputs of the latency-bound first phase of a chunk (inputs in Montgomery form, witness, blinding draws, transcript
// and RNG states).  One Front per chunk of a batch so that phase A of EVERY chunk runs concurrently (each on its own
// pair of streams) before the throughput-bound phase B walks the chunks one after the other.
struct Front {
  int B = 0;
  scm *v = 0, *vbl = 0, *aux = 0, *pub = 0, *wit = 0, *rand1 = 0;
  strobe128 *ts = 0, *rng = 0;
  dev_side sideR, sideW;
  void release() {
    void *ps[] = {v, vbl, aux, pub, wit, rand1, ts, rng};
    for (void *p : ps) dev_free(p);
    dev_side_free(sideR); dev_side_free(sideW);
    *this = Front();
  }
};
// a batch between engine_prove_begin and engine_prove_finish: its arguments (label copied) and chunking
struct PendingProve { bool active = false; ProveArgs a{}; int chunk = 0; std::vector<uint8_t> label; };
struct Workspace {
  int B = 0;
  std::vector<Front> fronts[2];  // two slots: phase A of the next batch of a stream runs beside phase B of the current one
  PendingProve pending[2];
  scm *vpub = 0, *uj = 0, *w_all = 0, *zpow = 0, *ypow = 0, *yinvpow = 0, *a = 0, *b = 0, *chal = 0, *t = 0, *tb = 0,
      *clr = 0, *part = 0;
  int8_t *dig = 0; size_t dig_bytes = 0;
  ge_p3 *buckets = 0; size_t bucket_slots = 0;
  ge_p3 *wsum = 0, *Q = 0, *Gt = 0, *Ht = 0, *pts = 0;
  int8_t *naf = 0; int *naf_top = 0;
  scm *utab = 0; uint32_t *rg_as = 0; long rg_cap = -1;
  uint32_t *rg_ai = 0; uint8_t *skip_ai = 0; long rg_ai_cap = -1;  // A_I row map / skip flags with the merged S-box rows
  uint32_t *rg_ver = 0; long rg_ver_cap = -1, rg_ver_N = -1;
  uint32_t *items = 0, *boff = 0, *soff = 0; size_t items_cap = 0, slices_cap = 0; ge_p3 *seg = 0; size_t seg_cap = 0;
  scm *flat_parts = 0;  // partial sums of the long slots (KFlattenParts)
  void release() {
    for (auto &fs : fronts) { for (Front &f : fs) f.release(); fs.clear(); }
    void *ps[] = {flat_parts, items, boff, soff, seg, rg_ver, utab, rg_as, rg_ai, skip_ai, vpub, uj, w_all, zpow, ypow, yinvpow, a, b, chal, t, tb, clr, part, dig, buckets, wsum, Q, Gt, Ht, pts, naf, naf_top};
    for (void *p : ps) dev_free(p);
    *this = Workspace();
  }
};

// ------------------------------------------------------------------------------------------------ generators
int gens_create(uint32_t capacity, BpGens **out) {
  int rc = bp_device_init();
  if (rc) return rc;
  if (capacity == 0) capacity = 1;
  BpGens *g = new BpGens();
  memset(g, 0, sizeof *g);
  g->capacity = capacity;
  dev_stream s = 0;
  if (dalloc(&g->G_p3, capacity) || dalloc(&g->H_p3, capacity) || dalloc(&g->G_n, capacity) || dalloc(&g->H_n, capacity) ||
      dalloc(&g->pc, 2) || dalloc(&g->pc_niels, 2) || dalloc(&g->pc_table, 2 * 64 * 16)) { gens_free(g); return BP_ERR_OOM; }
  // generator chains: SHAKE256("GeneratorsChain" || label) with label = 'G'/'H' || u32le(party 0)
  std::vector<uint8_t> uni((size_t)capacity * 64);
  uint8_t *d_uni = nullptr, *d_small = nullptr; int *d_ok = nullptr;
  CK(dalloc(&d_uni, (size_t)capacity * 64)); CK(dalloc(&d_small, 256)); CK(dalloc(&d_ok, 1));
  for (int which = 0; which < 2; which++) {
    uint8_t seed[20]; memcpy(seed, "GeneratorsChain", 15); seed[15] = which ? 'H' : 'G'; seed[16] = seed[17] = seed[18] = seed[19] = 0;
    keccak_xof k; shake256_init(k, seed, 20);
    keccak_squeeze(k, uni.data(), uni.size());
    CK(dev_h2d(d_uni, uni.data(), uni.size(), s));
    CK(launch(capacity, s, KGensFromUniform{d_uni, which ? g->H_p3 : g->G_p3, which ? g->H_n : g->G_n}));
    CK(dev_sync(s));
  }
  // Pedersen bases
  uint8_t small[96 + 64]; memcpy(small, BASEPOINT_C, 32); sha3_512(small + 32, BASEPOINT_C, 32);
  int ok = 1;
  CK(dev_h2d(d_small, small, 96, s)); CK(dev_h2d(d_ok, &ok, sizeof ok, s));
  CK(launch(2, s, KPcBases{d_small, d_small + 32, g->pc, g->pc_niels, d_small + 96, d_ok}));
  CK(launch(128, s, KPcTable{g->pc, g->pc_table}));
  CK(dev_d2h(g->pc_c, d_small + 96, 64, s)); CK(dev_d2h(&ok, d_ok, sizeof ok, s));
  CK(dev_sync(s));
  dev_free(d_uni); dev_free(d_small); dev_free(d_ok);
  if (!ok) { gens_free(g); return BP_ERR_CUDA; }
  // direct 8-bit tables (384 KiB per generator); BP_B200_NO_TABLE=1 keeps the bucket-method-only pipeline
  const size_t ngen = 2 * (size_t)capacity + 2;
  const size_t table_bytes = ngen * TBL_W * TBL_E * sizeof(ge_niels);
  // BP_B200_NO_DIRECT_TABLE=1: the configuration of capacities above ~65k generators (shift table only) at any size, for tests
  if (!getenv("BP_B200_NO_TABLE") && !getenv("BP_B200_NO_DIRECT_TABLE") && table_bytes <= ((size_t)48 << 30)) {  // capacities above ~65k generators do without
    if (dalloc(&g->table, ngen * TBL_W * TBL_E)) { gens_free(g); return BP_ERR_OOM; }
    CK(launch((long)ngen * TBL_W, s, KTableBuild{g->G_p3, g->H_p3, g->pc, (long)capacity, g->table}));
    CK(dev_sync(s));
  }
  // shift table of the sorted-bucket MSM (SB_WINDOWS x 96 B per generator): kept up to 16 GB, i.e. capacity 2^22
  if (!getenv("BP_B200_NO_TABLE") && !getenv("BP_B200_NO_SORTED") && ngen * SB_WINDOWS * sizeof(ge_niels) <= ((size_t)16 << 30)) {
    g->merge_slots = capacity < 16384 ? capacity : 16384;
    if (dalloc(&g->sg, (ngen + SG_SPARE + g->merge_slots) * SB_WINDOWS)) { gens_free(g); return BP_ERR_OOM; }
    CK(launch((long)ngen, s, KShiftTableBuild{g->G_p3, g->H_p3, g->pc, (long)capacity, g->sg}));
    CK(dev_sync(s));
  }
  *out = g;
  return BP_OK;
}
void gens_free(BpGens *g) {
  if (!g) return;
  if (g->msm_ws) { g->msm_ws->release(); delete g->msm_ws; }
  dev_free(g->G_p3); dev_free(g->H_p3); dev_free(g->G_n); dev_free(g->H_n); dev_free(g->pc); dev_free(g->pc_niels); dev_free(g->pc_table); dev_free(g->table); dev_free(g->ftable); dev_free(g->sg);
  delete g;
}
int gens_export(const BpGens *g, int which, uint32_t count, uint8_t *out) {
  if (count > g->capacity) return BP_ERR_INVALID_GENERATORS_LENGTH;
  uint8_t *d = nullptr; CK(dalloc(&d, (size_t)count * 32));
  CK(launch(count, 0, KEncodePoints{which ? g->H_p3 : g->G_p3, d}));
  CK(dev_d2h(out, d, (size_t)count * 32, 0)); CK(dev_sync(0));
  dev_free(d);
  return BP_OK;
}

// ------------------------------------------------------------------------------------------------ circuits
static uint32_t next_pow2(uint32_t n) { uint32_t N = 1; while (N < n) N <<= 1; return N; }
static uint32_t ilog2(uint32_t N) { uint32_t k = 0; while ((1u << k) < N) k++; return k; }

size_t circuit_proof_len(const BpCircuit *c) { return 32 * (size_t)(14 + 2 * c->k + 2); }

int circuit_create(uint32_t n, uint32_t m, uint32_t npub, uint32_t q, const uint32_t *cons_ptr, const uint8_t *kind, const uint32_t *idx,
                   const scm *coeff, const TapeOp *tape, uint32_t naux, uint32_t nwlc, const uint32_t *wlc_ptr,
                   const uint8_t *wkind, const uint32_t *widx, const scm *wcoeff, const HostPoseidonTape *ptape, BpCircuit **out) {
  int rc = bp_device_init();
  if (rc) return rc;
  BpCircuit *c = new BpCircuit();
  memset(c, 0, sizeof *c);
  { static uint64_t next_serial = 0; c->serial = ++next_serial; }
  c->n = n; c->m = m; c->q = q; c->N = next_pow2(n ? n : 1); c->k = ilog2(c->N);
  c->nnz = cons_ptr[q]; c->nslots = 3 * n + m + 1 + npub; c->naux = naux; c->npub = npub;
  // validate + transpose to slot-major
  std::vector<uint32_t> slot_cnt(c->nslots + 1, 0);
  auto slot_of = [&](uint32_t t) -> long {
    uint32_t i = idx[t];
    switch (kind[t]) {
      case 1: return i < n ? (long)i : -1;
      case 2: return i < n ? (long)n + i : -1;
      case 3: return i < n ? (long)2 * n + i : -1;
      case 0: return i < m ? (long)3 * n + i : -1;
      case 4: return (long)3 * n + m;
      case 5: return i < npub ? (long)3 * n + m + 1 + i : -1;
      default: return -1;
    }
  };
  for (uint32_t t = 0; t < c->nnz; t++) { long s = slot_of(t); if (s < 0) { delete c; return BP_ERR_INVALID_ARGUMENT; } slot_cnt[s + 1]++; }
  for (uint32_t s = 0; s < c->nslots; s++) slot_cnt[s + 1] += slot_cnt[s];
  std::vector<uint32_t> fill(slot_cnt.begin(), slot_cnt.end() - 1), tq(c->nnz ? c->nnz : 1);
  std::vector<scm> tc(c->nnz ? c->nnz : 1);
  for (uint32_t k = 0; k < q; k++)
    for (uint32_t t = cons_ptr[k]; t < cons_ptr[k + 1]; t++) {
      long s = slot_of(t); uint32_t at = fill[s]++;
      tq[at] = k;
      scm co = coeff[t];
      if (kind[t] == 0 || kind[t] == 4 || kind[t] == 5) co = sc_neg(co);  // wV and wc accumulate with a minus sign (A.3 step 7)
      tc[at] = co;
    }
  dev_stream s = 0;
#define CKC(x) do { if (x) { fprintf(stderr, "bp_b200: %s failed at %s:%d\n", #x, __FILE__, __LINE__); circuit_free(c); return BP_ERR_CUDA; } } while (0)
  if (dalloc(&c->d_slot_ptr, c->nslots + 1) || dalloc(&c->d_tq, tq.size()) || dalloc(&c->d_tcoeff, tc.size())) { circuit_free(c); return BP_ERR_OOM; }
  CKC(dev_h2d(c->d_slot_ptr, slot_cnt.data(), (c->nslots + 1) * sizeof(uint32_t), s));
  CKC(dev_h2d(c->d_tq, tq.data(), tq.size() * sizeof(uint32_t), s));
  CKC(dev_h2d(c->d_tcoeff, tc.data(), tc.size() * sizeof(scm), s));
  {  // long slots -> parts of FLATTEN_PART terms
    std::vector<uint32_t> pbeg, pend, lslot, lfirst;
    for (uint32_t sl = 0; sl < c->nslots; sl++) {
      const uint32_t t0 = slot_cnt[sl], t1 = slot_cnt[sl + 1];
      if (t1 - t0 <= 2 * FLATTEN_PART) continue;
      lslot.push_back(sl); lfirst.push_back((uint32_t)pbeg.size());
      for (uint32_t t = t0; t < t1; t += FLATTEN_PART) { pbeg.push_back(t); pend.push_back(std::min(t + FLATTEN_PART, t1)); }
    }
    lfirst.push_back((uint32_t)pbeg.size());
    c->nlong = (uint32_t)lslot.size(); c->nparts = (uint32_t)pbeg.size();
    if (c->nlong) {
      if (dalloc(&c->d_part_beg, pbeg.size()) || dalloc(&c->d_part_end, pend.size()) || dalloc(&c->d_long_slot, lslot.size()) || dalloc(&c->d_long_first, lfirst.size())) {
        circuit_free(c); return BP_ERR_OOM;
      }
      CKC(dev_h2d(c->d_part_beg, pbeg.data(), pbeg.size() * sizeof(uint32_t), s)); CKC(dev_h2d(c->d_part_end, pend.data(), pend.size() * sizeof(uint32_t), s));
      CKC(dev_h2d(c->d_long_slot, lslot.data(), lslot.size() * sizeof(uint32_t), s)); CKC(dev_h2d(c->d_long_first, lfirst.data(), lfirst.size() * sizeof(uint32_t), s));
      CKC(dev_sync(s));  // the host vectors go out of scope
    }
  }
  if (tape) {
    c->has_tape = 1;
    uint32_t wn = wlc_ptr[nwlc];
    for (uint32_t t = 0; t < wn; t++) if (wkind[t] == 5) c->wit_uses_pub = 1;
    for (uint32_t t = 0; t < wn; t++) {
      uint32_t lim = wkind[t] == 0 ? m : (wkind[t] == 4 ? 1 : (wkind[t] == 5 ? npub : n));
      if (wkind[t] > 5 || widx[t] >= lim) { circuit_free(c); return BP_ERR_INVALID_ARGUMENT; }
    }
    for (uint32_t i = 0; i < n; i++) {
      const TapeOp &op = tape[i];
      if (op.opL == W_SKIP) continue;
      if (op.opL == W_POSEIDON) { if (!ptape || op.argL >= ptape->nblocks) { circuit_free(c); return BP_ERR_INVALID_ARGUMENT; } continue; }
      bool okL = (op.opL == W_LC && op.argL < nwlc) || (op.opL == W_AUX && op.argL < naux);
      bool okR = (op.opR == W_LC && op.argR < nwlc) || (op.opR == W_AUX && op.argR < naux) || op.opR == W_INV_L;
      if (!okL || !okR) { circuit_free(c); return BP_ERR_INVALID_ARGUMENT; }
    }
    if (dalloc(&c->d_tape, n ? n : 1) || dalloc(&c->d_wptr, nwlc + 1) || dalloc(&c->d_wkind, wn ? wn : 1) || dalloc(&c->d_widx, wn ? wn : 1) ||
        dalloc(&c->d_wcoeff, wn ? wn : 1)) { circuit_free(c); return BP_ERR_OOM; }
    CKC(dev_h2d(c->d_tape, tape, n * sizeof(TapeOp), s));
    CKC(dev_h2d(c->d_wptr, wlc_ptr, (nwlc + 1) * sizeof(uint32_t), s));
    CKC(dev_h2d(c->d_wkind, wkind, wn, s));
    CKC(dev_h2d(c->d_widx, widx, wn * sizeof(uint32_t), s));
    CKC(dev_h2d(c->d_wcoeff, wcoeff, wn * sizeof(scm), s));
    if (ptape && ptape->nblocks) {
      const uint32_t total = ptape->full_b + ptape->partial + ptape->full_e;
      if (ptape->nkeys < total * POSEIDON_WIDTH) { circuit_free(c); return BP_ERR_INVALID_ARGUMENT; }
      for (uint32_t b = 0; b < ptape->nblocks; b++) {
        const PoseidonBlock &pb = ptape->blocks[b];
        uint32_t per = pb.sbox == 1 ? 3 : 2, cnt = per * (POSEIDON_WIDTH * (ptape->full_b + ptape->full_e) + ptape->partial);
        bool ok = pb.sbox <= 1 && (uint64_t)pb.first_mult + cnt <= n;
        for (int i = 0; i < POSEIDON_WIDTH; i++) ok = ok && pb.in_lc[i] < nwlc;
        if (!ok) { circuit_free(c); return BP_ERR_INVALID_ARGUMENT; }
      }
      if (dalloc(&c->d_pblocks, ptape->nblocks) || dalloc(&c->d_pos_rk, ptape->nkeys) || dalloc(&c->d_pos_mds, POSEIDON_WIDTH * POSEIDON_WIDTH)) { circuit_free(c); return BP_ERR_OOM; }
      CKC(dev_h2d(c->d_pblocks, ptape->blocks, ptape->nblocks * sizeof(PoseidonBlock), s));
      CKC(dev_h2d(c->d_pos_rk, ptape->round_keys, ptape->nkeys * sizeof(scm), s));
      CKC(dev_h2d(c->d_pos_mds, ptape->mds, POSEIDON_WIDTH * POSEIDON_WIDTH * sizeof(scm), s));
      c->pos = PoseidonDev{c->d_pos_rk, c->d_pos_mds, ptape->full_b, ptape->partial, ptape->full_e};
      // inverse S-box multipliers (x, 1/x, .), (x, 0, .), (x, 1/x, .) written by the block op: the three left wires are equal and
      // so are the first and third right wires -> A_I can use the SUM of their generators (one row per group)
      c->merge_src = new std::vector<uint32_t>();
      for (uint32_t b = 0; b < ptape->nblocks; b++) {
        const PoseidonBlock &pb = ptape->blocks[b];
        if (pb.sbox != 1) continue;
        const uint32_t sboxes = POSEIDON_WIDTH * (ptape->full_b + ptape->full_e) + ptape->partial;
        for (uint32_t sb = 0; sb < sboxes; sb++) {
          const uint32_t m0 = pb.first_mult + 3 * sb;
          c->merge_src->insert(c->merge_src->end(), {m0, m0 + 1, m0 + 2});                       // left wires: G_m0 + G_m0+1 + G_m0+2
          c->merge_src->insert(c->merge_src->end(), {n + m0, n + m0 + 2, 0xffffffffu});          // right wires: H_m0 + H_m0+2
        }
      }
    }
  }
  CKC(dev_sync(s));
  c->ws = new Workspace();
  *out = c;
  return BP_OK;
}
int circuit_set_fixed_commitments(BpCircuit *c, uint32_t n, const uint32_t *idx, const uint8_t *V) {
  for (uint32_t i = 0; i < n; i++) if (idx[i] >= c->m) return BP_ERR_INVALID_ARGUMENT;
  dev_free(c->d_fixed_idx); dev_free(c->d_fixed_V); c->d_fixed_idx = nullptr; c->d_fixed_V = nullptr; c->nfixed = 0;
  if (!n) return BP_OK;
  if (dalloc(&c->d_fixed_idx, n) || dalloc(&c->d_fixed_V, (size_t)n * 32)) return BP_ERR_OOM;
  CK(dev_h2d(c->d_fixed_idx, idx, n * sizeof(uint32_t), 0)); CK(dev_h2d(c->d_fixed_V, V, (size_t)n * 32, 0)); CK(dev_sync(0));
  c->nfixed = n;
  return BP_OK;
}
void circuit_free(BpCircuit *c) {
  if (!c) return;
  dev_free(c->d_fixed_idx); dev_free(c->d_fixed_V);
  dev_free(c->d_part_beg); dev_free(c->d_part_end); dev_free(c->d_long_slot); dev_free(c->d_long_first);
  dev_free(c->d_slot_ptr); dev_free(c->d_tq); dev_free(c->d_tcoeff); dev_free(c->d_tape); dev_free(c->d_wptr); dev_free(c->d_wkind);
  dev_free(c->d_widx); dev_free(c->d_wcoeff); dev_free(c->d_pblocks); dev_free(c->d_pos_rk); dev_free(c->d_pos_mds);
  delete c->merge_src;
  if (c->ws) { c->ws->release(); delete c->ws; }
  delete c;
}

// frees the per-batch device workspace (tens of GB for a large batch); the next call allocates it again
int circuit_release_workspace(BpCircuit *c) {
  if (!c || !c->ws) return BP_OK;
  if (c->ws->pending[0].active || c->ws->pending[1].active) return BP_ERR_INVALID_ARGUMENT;  // a batch is between begin and finish
  if (dev_sync_device()) return BP_ERR_CUDA;  // work of earlier calls may still be running on the caller's streams
  c->ws->release();
  return BP_OK;
}

// the sorted-bucket path pays ~54k additions per instance for its 16384-bucket reduction: below this many rows per instance the
// direct 8-bit tables (32 additions per row, no reduction) are cheaper -- the small circuits (bound check, Poseidon 2:1, MiMC)
static long sorted_min_rows() { const char *e = getenv("BP_B200_SORTED_MIN_ROWS"); return e ? atol(e) : 8192; }  // env: tests force either path
#define SORTED_MIN_ROWS sorted_min_rows()
static const int UNFOLD_MAX = 6;
// inner-product rounds computed over the original generators (sorted-bucket MSM on the shift table) before the folded
// generators are materialised; 4 is the measured optimum at N = 32768 (BP_B200_UNFOLD overrides, for experiments)
static int unfold_rounds() {
  static int v = -1;
  if (v < 0) { const char *e = getenv("BP_B200_UNFOLD"); v = e ? atoi(e) : 4; if (v < 0) v = 0; if (v > UNFOLD_MAX) v = UNFOLD_MAX; }
  return v;
}
static const int CH_DOT = 256;  // multipliers per partial-sum thread
static const int CH_POW = 64;   // exponents per powers thread
static long msm_target_warps() { return 148L * 8 * 4; }

// device bytes ensure_workspace() + ensure_front() allocate per proof of a chunk (the MSM bucket floor of small chunks aside)
double engine_workspace_bytes_per_proof(const BpCircuit *c) {
  const double n = c->n, N = c->N, m = c->m, q = c->q, k = c->k;
  const double rows = std::max(std::max(5 * n + 3, 2 * (N + 2)), 2 * N + m + 13 + 2 * k + 2);
  const double items = std::max(2 * n + 1, N + 2) * SB_WINDOWS, slices = std::max(SB_BUCKETS + items / SB_SEG + 2, items * 4 / sizeof(ge_p3) + 1);
  const double nch = std::max(n, N) / CH_DOT + 2;
  double b = 0;
  b += sizeof(scm) * ((double)c->nslots + c->nparts + 1 + (q + 1) + 4 * N + (2 * k + 2) + 40 + nch * 6 + (c->npub + 1));  // w_all, zpow, ypow, yinvpow, a, b, ...
  b += std::max(rows * SB_ROW_BYTES, 2 * N * 64);                                // digit rows
  b += sizeof(ge_p3) * (slices + 2 * (N / 2 + 1) + (m + 12 + 2 * k) + 2 * SB_SEGS + 3 * SB_FIN_GROUPS + 1 + 2 * MSM_WINDOWS);  // partial sums, folded generators, ...
  b += 4 * items + 8 * (SB_BUCKETS + 1) + 8 * 256 + 32;                          // sorted items, offsets, NAFs
  b += sizeof(scm) * (2 * (m + 1) + (c->naux + 1) + (c->npub + 1) + 3 * (n + 1) + (3 + 2 * n)) + 2 * sizeof(strobe128);  // front
  return b;
}

// flattened constraint weights of a chunk: one thread per (slot, proof); slots with very many terms in parts (KFlattenParts)
static int run_flatten(const BpCircuit *c, Workspace *w, int B, dev_stream s) {
  KFlatten kf{c->d_slot_ptr, c->d_tq, c->d_tcoeff, w->zpow, w->w_all, B};
  kf.split = c->nlong ? 2 * FLATTEN_PART : 0;
  CK(launch((long)c->nslots * B, s, kf));
  if (c->nlong) {
    CK(launch((long)c->nparts * B, s, KFlattenParts{c->d_part_beg, c->d_part_end, c->d_tq, c->d_tcoeff, w->zpow, w->flat_parts, B}));
    CK(launch((long)c->nlong * B, s, KFlattenSum{c->d_long_slot, c->d_long_first, w->flat_parts, w->w_all, B}));
  }
  return BP_OK;
}

static int ensure_workspace(BpCircuit *c, int B) {
  Workspace *w = c->ws;
  if (w->B >= B) return BP_OK;
  if (w->pending[0].active || w->pending[1].active) return BP_ERR_INVALID_ARGUMENT;  // growing would free the fronts of a batch in flight
  w->release();
  const size_t n = c->n, N = c->N, m = c->m, q = c->q, Bz = (size_t)B;
  const size_t k = c->k;
  size_t nchunks = (std::max(n, N) + CH_DOT - 1) / CH_DOT + 1;
  size_t rows_as = (2 * n + 1) + (n + 1) + (2 * n + 1), rows_ipa = 2 * (N + 2), rows_ver = 2 * N + m + 13 + 2 * k + 2;
  w->dig_bytes = std::max(std::max(rows_as, rows_ipa), rows_ver) * SB_ROW_BYTES * Bz;  // sized for the wider 13-bit rows
  w->dig_bytes = std::max(w->dig_bytes, 2 * N * (size_t)64 * Bz);  // 2N 64-byte rows of narrow-window digits (KRecodeFoldTable with fold tables)
  // bucket slots: enough for one MSM launch at the largest split the launcher will pick
  size_t max_warps = (size_t)std::max<long>(msm_target_warps(), 2L * B) + B;
  w->items_cap = (size_t)std::max(2 * n + 1, N + 2) * SB_WINDOWS;  // items of one instance
  w->slices_cap = SB_BUCKETS + w->items_cap / SB_SEG + 2;  // partial sums: one per (bucket, segment) crossing
  w->bucket_slots = std::max(max_warps * MSM_WINDOWS * MSM_BUCKETS, Bz * w->slices_cap);
  // the buffer doubles as the scratch of the two-pass bucket sort (one 4-byte item per digit)
  w->bucket_slots = std::max(w->bucket_slots, Bz * ((w->items_cap * sizeof(uint32_t) + sizeof(ge_p3) - 1) / sizeof(ge_p3)));
  int bad = 0;
  bad |= dalloc(&w->vpub, (size_t)(c->npub + 1) * Bz); bad |= dalloc(&w->uj, (2 * k + 2) * Bz);
  bad |= dalloc(&w->w_all, (size_t)c->nslots * Bz); bad |= dalloc(&w->flat_parts, (size_t)(c->nparts + 1) * Bz);
  bad |= dalloc(&w->zpow, (q + 1) * Bz); bad |= dalloc(&w->ypow, N * Bz); bad |= dalloc(&w->yinvpow, N * Bz);
  bad |= dalloc(&w->a, N * Bz); bad |= dalloc(&w->b, N * Bz); bad |= dalloc(&w->chal, 16 * Bz);
  bad |= dalloc(&w->t, 6 * Bz); bad |= dalloc(&w->tb, 5 * Bz); bad |= dalloc(&w->clr, 2 * Bz); bad |= dalloc(&w->part, nchunks * 6 * Bz);
  bad |= dalloc(&w->dig, w->dig_bytes); bad |= dalloc(&w->buckets, w->bucket_slots); bad |= dalloc(&w->wsum, (max_warps + (size_t)B + 64) * MSM_WINDOWS);
  bad |= dalloc(&w->Q, Bz); bad |= dalloc(&w->Gt, (N / 2 + 1) * Bz); bad |= dalloc(&w->Ht, (N / 2 + 1) * Bz);
  bad |= dalloc(&w->pts, (m + 11 + 2 * k + 1) * Bz);
  bad |= dalloc(&w->naf, 8 * 256 * Bz); bad |= dalloc(&w->naf_top, 8 * Bz);
  bad |= dalloc(&w->items, w->items_cap * Bz); bad |= dalloc(&w->boff, (size_t)(SB_BUCKETS + 1) * Bz); bad |= dalloc(&w->soff, (size_t)(SB_BUCKETS + 1) * Bz);
  bad |= dalloc(&w->seg, ((size_t)SB_SEGS * 2 + SB_FIN_GROUPS * 3) * Bz);
  bad |= dalloc(&w->utab, 4 * (size_t)(1 << UNFOLD_MAX) * Bz + 4 * Bz); bad |= dalloc(&w->rg_as, 2 * n + 2); bad |= dalloc(&w->rg_ai, 2 * n + 2); bad |= dalloc(&w->skip_ai, 2 * n + 2);
  if (bad) { w->release(); return BP_ERR_OOM; }
  w->B = B;
  return BP_OK;
}

static int ensure_front(BpCircuit *c, int slot, size_t idx, int B) {
  Workspace *w = c->ws;
  if (w->fronts[slot].size() <= idx) w->fronts[slot].resize(idx + 1);
  Front &f = w->fronts[slot][idx];
  if (f.B >= B) return BP_OK;
  f.release();
  const size_t n = c->n, m = c->m, Bz = (size_t)B;
  int bad = 0;
  bad |= dalloc(&f.v, (m + 1) * Bz); bad |= dalloc(&f.vbl, (m + 1) * Bz); bad |= dalloc(&f.aux, (size_t)(c->naux + 1) * Bz);
  bad |= dalloc(&f.pub, (size_t)(c->npub + 1) * Bz); bad |= dalloc(&f.wit, 3 * (n + 1) * Bz); bad |= dalloc(&f.rand1, (3 + 2 * n) * Bz);
  bad |= dalloc(&f.ts, Bz); bad |= dalloc(&f.rng, Bz);
  bad |= dev_side_init(f.sideR); bad |= dev_side_init(f.sideW);
  if (bad) { f.release(); return BP_ERR_OOM; }
  f.B = B;
  return BP_OK;
}

// one MSM launch: ninst instances, shared digit buffer, result handling in KMsmFinish
static int run_msm(Workspace *w, const MsmSeg *segs, int nseg, long ninst, const int8_t *dig, long dig_inst_stride, uint8_t *out,
                   long out_stride, int mode, int *status, dev_stream s) {
  long rows = 0;
  for (int i = 0; i < nseg; i++) rows += segs[i].count;
  long S = (msm_target_warps() + ninst - 1) / ninst;
  long maxS = rows / 1024; if (maxS < 1) maxS = 1;  // keep the 256-addition bucket reduction of a split below ~25 % of its accumulation
  if (S > maxS) S = maxS;
  if (S < 1) S = 1;
  while ((size_t)(ninst * S) * MSM_WINDOWS * MSM_BUCKETS > w->bucket_slots && S > 1) S--;
  if ((size_t)(ninst * S) * MSM_WINDOWS * MSM_BUCKETS > w->bucket_slots) return BP_ERR_OOM;
  KMsmAccumulate k{};
  for (int i = 0; i < nseg; i++) k.seg[i] = segs[i];
  k.nseg = nseg; k.S = (int)S; k.dig = dig; k.dig_inst_stride = dig_inst_stride; k.buckets = w->buckets; k.wsum = w->wsum;
  CK(launch(ninst * S * MSM_WINDOWS, s, k));
  if (S > 4) {  // many splits (few, large instances): add them up in parallel per window first
    ge_p3 *sums = w->wsum + (size_t)ninst * S * MSM_WINDOWS;
    CK(launch(ninst * MSM_WINDOWS, s, KMsmWindowSum{w->wsum, (int)S, sums}));
    CK(launch(ninst, s, KMsmFinish{sums, 1, out, out_stride, mode, status, BP_ERR_VERIFICATION}));
  } else {
    CK(launch(ninst, s, KMsmFinish{w->wsum, (int)S, out, out_stride, mode, status, BP_ERR_VERIFICATION}));
  }
  return BP_OK;
}

// table-driven MSM over shared generators: ninst instances, rows digit rows each
static int run_msm_table(const BpGens *g, Workspace *w, const RowMap &rmap, long rows, long ninst, const int8_t *dig, long dig_inst_stride,
                         uint8_t *out, long out_stride, dev_stream s) {
  long S = 262144 / (ninst > 0 ? ninst : 1);
  // a handful of instances (the MSM entry point, single proofs): the launch is a latency chain, not a throughput problem -- down to one
  // row (32 additions) per thread, partial sums combined 16 to 1 per stage
  const bool few = ninst <= 8;
  const long min_rows = few ? 1 : 16;
  if (S > rows / min_rows) S = rows / min_rows;
  if (S > (few ? 32768 : 1024)) S = few ? 32768 : 1024;
  if (S < 1) S = 1;
  while ((size_t)(ninst * S) > w->bucket_slots && S > 1) S--;
  CK(launch(ninst * S, s, KMsmTable{g->table, rmap, dig, dig_inst_stride, rows, (int)S, w->buckets}));
  if (few) {
    ge_p3 *cur = w->buckets, *stage = w->buckets + (size_t)ninst * S;
    long count = S;
    while (count > 16) {
      const long R = (count + 15) / 16;
      if ((size_t)(stage - w->buckets) + (size_t)ninst * R > w->bucket_slots) return BP_ERR_OOM;
      CK(launch(ninst * R, s, KMsmTableReduce{cur, (int)count, (int)R, stage}));
      cur = stage; stage += ninst * R; count = R;
    }
    CK(launch(ninst, s, KMsmTableFinish{cur, (int)count, out, out_stride}));
  } else if (S > 64 && ninst < 4096) {  // two-stage reduction of the per-thread partial sums
    const int R = 32;
    ge_p3 *stage = w->buckets + (size_t)ninst * S;
    if ((size_t)ninst * (S + R) > w->bucket_slots) return BP_ERR_OOM;
    CK(launch(ninst * R, s, KMsmTableReduce{w->buckets, (int)S, R, stage}));
    CK(launch(ninst, s, KMsmTableFinish{stage, R, out, out_stride}));
  } else {
    CK(launch(ninst, s, KMsmTableFinish{w->buckets, (int)S, out, out_stride}));
  }
  return BP_OK;
}

// sorted-bucket MSM over shared generators (13-bit digit rows): counting sort -> bucket sums -> segment reduce -> finish
static int run_msm_sorted(const BpGens *g, Workspace *w, const RowMap &rmap, long rows, long ninst, const int8_t *dig, long dig_inst_stride,
                          uint8_t *out, long out_stride, dev_stream s) {
  if ((size_t)rows * SB_WINDOWS > w->items_cap || (size_t)ninst * w->slices_cap > w->bucket_slots) return BP_ERR_OOM;
  // the partial-sum buffer is dead until KBucketAccumulate below writes it: scratch of the two-pass sort
  CK(launch_sort_buckets(rmap, dig, dig_inst_stride, rows, ninst, w->items, (long)w->items_cap, w->boff, w->soff, (uint32_t *)w->buckets,
                         w->bucket_slots * sizeof(ge_p3), 2L * g->capacity + 2 + SG_SPARE + g->merge_slots, s));
  SortedView sv{w->items, w->boff, w->soff, (long)w->items_cap, (long)w->slices_cap};
  const long segs = (rows * SB_WINDOWS + SB_SEG - 1) / SB_SEG;  // segments of this launch's longest possible item list
#ifndef BP_HOST_EMUL
  if (g_profile_on) profile_count_items(w->boff, ninst, s);
#endif
  CK(launch(ninst * segs, s, KBucketAccumulate{g->sg, sv, w->buckets, segs}));
  CK(launch(ninst * SB_SEGS, s, KBucketReduce{w->buckets, sv, w->seg}));
  ge_p3 *grp = w->seg + (size_t)ninst * SB_SEGS * 2;  // behind the segment sums (ensure_workspace sizes the buffer for both)
  CK(launch(ninst * SB_FIN_GROUPS, s, KBucketFinishA{w->seg, grp}));
  CK(launch(ninst, s, KBucketFinishB{grp, out, out_stride, nullptr}));
  return BP_OK;
}

static void base_transcript(strobe128 &t, const uint8_t *label, int label_len) {
  ts_init(t, label, label_len);
  const uint8_t r1[7] = {'r', '1', 'c', 's', ' ', 'v', '1'};
  ts_append(t, "dom-sep", r1, 7);
}

// Round 0 of a padded circuit: the N - n H-rows of L_0 whose partner index is a padding index all carry the scalar -y^h
// (KRecodeUnfolded13), so their generators are summed ONCE per (generators, n, N) into a spare slot of the shift table.
// Returns the generator index of the slot, or -1 when there is no padding / no spare slot (then the rows are kept).
static long ensure_pad_generator(BpGens *g, long n, long N, dev_stream s) {
  if (!g->sg || n >= N || getenv("BP_B200_NO_PADSUM")) return -1;
  for (int k = 0; k < g->pad_count; k++) if (g->pad_n[k] == n && g->pad_N[k] == N) return 2L * g->capacity + 2 + k;
  if (g->pad_count >= SG_SPARE) return -1;
  const long h = N / 2, first = n - h, count = h - first;  // H_first .. H_{h-1}
  if (first < 0 || count <= 0) return -1;
  const long T = count < 256 ? count : 256;
  ge_p3 *tmp = nullptr;
  if (dalloc(&tmp, (size_t)T + 1)) return -1;
  int bad = launch(T, s, KSumPointsStrided{g->H_p3 + first, count, T, tmp});
  bad |= launch(1, s, KSumPointsStrided{tmp, T, 1, tmp + T});
  const int k = g->pad_count;
  bad |= launch(1, s, KShiftTableOne{tmp + T, 2L * g->capacity + 2 + k, g->sg});
  bad |= dev_sync(s);
  dev_free(tmp);
  if (bad) return -1;
  g->pad_n[k] = n; g->pad_N[k] = N; g->pad_count = k + 1;
  return 2L * g->capacity + 2 + k;
}

// Fold tables.  KFoldTable materialises the level-J generators of a proof from fixed-base tables (2^J terms per output, no
// doublings); the first J rounds then never fold generators.  Up to ~61k generators those are the 8-bit direct tables (26 GB at
// capacity 32768).  Above, the direct tables do not fit, but a table with NARROWER windows over the first N generators of each chain
// does: 2 N ceil(253/b) 2^(b-1) 96 B  =  69 GB for b = 6 (41 GB for b = 5) at N = 262144, the reference's own depth-253
// configuration, and a term costs 43 (51) additions instead of 32 -- against folding 2^18 generators with two scalar multiplications per output from round 0.
// Built on first use for the circuit's N (a generator set can serve smaller circuits; capacity itself may be much larger than
// any N: the reference asks for BulletproofGens::new(819200, 1), src/gadget_vsmt_2.rs:290).  BP_B200_FOLD_TABLE_GB: budget
// (default 72, 0 disables); BP_B200_FOLD_BITS: force the window width (tests).
struct FoldTableView { const ge_niels *table; long cap; int bits, W, E, rb; };
static bool fold_table_view(const BpGens *g, long N, FoldTableView &v) {
  if (g->table) { v = {g->table, (long)g->capacity, 8, TBL_W, TBL_E, 32}; return true; }
  if (g->ftable && g->ft_cap >= N) { v = {g->ftable, g->ft_cap, g->ft_bits, (253 + g->ft_bits - 1) / g->ft_bits, 1 << (g->ft_bits - 1), 64}; return true; }
  return false;
}
static int ensure_fold_table(BpGens *g, long N, dev_stream s) {
  const char *force = getenv("BP_B200_FOLD_BITS");
  if (g->table || (g->ftable && g->ft_cap >= N && !(force && atoi(force) != g->ft_bits))) return BP_OK;
  if (getenv("BP_B200_NO_TABLE") || !g->sg || N > (long)g->capacity) return BP_OK;  // the unfolded rounds themselves need the shift table
  double budget = 72.0;
  if (const char *e = getenv("BP_B200_FOLD_TABLE_GB")) budget = atof(e);
  int bits = 0;
  for (int b = 7; b >= 4 && !bits; b--) {
    const double bytes = 2.0 * (double)N * ((253 + b - 1) / b) * (double)(1 << (b - 1)) * sizeof(ge_niels);
    if (bytes <= budget * 1073741824.0) bits = b;
  }
  if (force) { const int b = atoi(force); if (b >= 4 && b <= 7 && budget > 0) bits = b; }
  dev_free(g->ftable); g->ftable = nullptr; g->ft_cap = 0;
  if (!bits) return BP_OK;  // no table: the rounds fold generators from round 0
  // narrower windows when the device has less room left than the budget (the workspace of a large chunk is allocated first)
  while (bits >= 4 && dalloc(&g->ftable, (size_t)2 * N * ((253 + bits - 1) / bits) * ((size_t)1 << (bits - 1)))) { g->ftable = nullptr; bits = force ? 0 : bits - 1; }
  if (bits < 4) return BP_OK;  // not even the 4-bit table fits: same fallback
  const int W = (253 + bits - 1) / bits, E = 1 << (bits - 1);
  KTableBuild kb{g->G_p3, g->H_p3, g->pc, N, g->ftable};
  kb.bits = bits; kb.W = W; kb.E = E;
  CK(launch(2 * N * W, s, kb));
  CK(dev_sync(s));
  g->ft_bits = bits; g->ft_cap = N;
  return BP_OK;
}

// A_I with merged rows: builds (once per generators / circuit pair) the generator sums of the circuit's equal-scalar groups in
// the merge slots of the shift table, the row -> generator map that points each group's first row at its sum, and the skip flags
// of the other rows.  Returns 1 when the merged map is usable.
static int ensure_merged_ai(BpGens *g, BpCircuit *c, dev_stream s) {
  Workspace *w = c->ws;
  if (!g->sg || !c->merge_src || c->merge_src->empty() || getenv("BP_B200_NO_MERGE")) return 0;
  const long groups = (long)c->merge_src->size() / 3, n = c->n, cap = g->capacity;
  if (groups > g->merge_slots) return 0;
  const long slot0 = 2 * cap + 2 + SG_SPARE;
  if (g->merge_owner != c->serial) {
    std::vector<uint32_t> src(*c->merge_src);
    for (uint32_t &x : src) if (x != 0xffffffffu) x = x < (uint32_t)n ? x : (uint32_t)(cap + (x - n));  // circuit index -> generator index
    uint32_t *d_src = nullptr;
    if (dalloc(&d_src, src.size())) return 0;
    int bad = dev_h2d(d_src, src.data(), src.size() * sizeof(uint32_t), s);
    bad |= launch(groups, s, KMergeGens{g->G_p3, g->H_p3, cap, d_src, slot0, g->sg});
    bad |= dev_sync(s);
    dev_free(d_src);
    if (bad) return 0;
    g->merge_owner = c->serial;
    w->rg_ai_cap = -1;
  }
  if (w->rg_ai_cap != cap) {
    std::vector<uint32_t> rg(2 * n + 1);
    std::vector<uint8_t> skip(2 * n + 1, 0);
    rg[0] = 2 * cap + 1;
    for (long i = 0; i < n; i++) { rg[1 + i] = (uint32_t)i; rg[1 + n + i] = (uint32_t)(cap + i); }
    for (long j = 0; j < groups; j++) {
      const uint32_t lead = (*c->merge_src)[3 * j];
      rg[1 + lead] = (uint32_t)(slot0 + j);  // rows are 1 + circuit index (left wires 0..n-1, right wires n..2n-1)
      for (int t = 1; t < 3; t++) { const uint32_t o = (*c->merge_src)[3 * j + t]; if (o != 0xffffffffu) skip[1 + o] = 1; }
    }
    if (dev_h2d(w->rg_ai, rg.data(), rg.size() * sizeof(uint32_t), s) || dev_h2d(w->skip_ai, skip.data(), skip.size(), s) || dev_sync(s)) return 0;
    w->rg_ai_cap = cap;
  }
  return 1;
}

// ------------------------------------------------------------------------------------------------ prover
// Phase A of one chunk (SURVEY A.3 steps 1-3 + witness): everything here is a one-thread-per-proof sequential chain
// (Keccak permutations, field inversions), i.e. latency-bound and nearly free in throughput terms.  It runs on the chunk's
// own two streams -- transcript/RNG on one, witness on the other -- forked from the caller's stream.
static int prove_phase_a(const BpGens *g, BpCircuit *c, Front &f, const ProveArgs &A, dev_stream s) {
  const int B = A.B;
  const long n = c->n, m = c->m;
  scm *aL = f.wit, *aR = f.wit + n * B, *aO = f.wit + 2 * n * B;
  dev_stream sR = dev_side_fork(f.sideR, s);
  CK(launch(m * B, sR, KLoadScalars{A.v, f.v, (int)m, B}));
  CK(launch(m * B, sR, KLoadScalars{A.vbl, f.vbl, (int)m, B}));
  dev_stream sW = dev_side_fork(f.sideW, sR);
  if (A.aL) {
    CK(launch(n * B, sW, KLoadScalars{A.aL, aL, (int)n, B}));
    CK(launch(n * B, sW, KLoadScalars{A.aR, aR, (int)n, B}));
    CK(launch(n * B, sW, KLoadScalars{A.aO, aO, (int)n, B}));
  } else {
    if (c->naux) CK(launch((long)c->naux * B, sW, KLoadScalars{A.aux, f.aux, (int)c->naux, B}));
    if (c->npub && A.pub) CK(launch((long)c->npub * B, sW, KLoadScalars{A.pub, f.pub, (int)c->npub, B}));
    else if (c->npub) CK(dev_memset(f.pub, 0, sizeof(scm) * (size_t)c->npub * B, sW));  // unread (wit_uses_pub checked by the caller), but never uninitialised
    CK(launch(B, sW, KWitnessTape{c->d_tape, WitnessLcs{c->d_wptr, c->d_wkind, c->d_widx, c->d_wcoeff}, c->d_pblocks, c->pos, (int)n, B, f.v, f.aux, f.pub, aL, aR, aO}));
  }
  CK(launch(m * B, sR, KCommit{f.v, f.vbl, (int)m, B, g->pc_table, A.V_out, m * 32, 32, nullptr}));
  strobe128 base; base_transcript(base, A.label, A.label_len);
  CK(launch(B, sR, KTsStart{base, A.V_out, (int)m, B, f.vbl, A.entropy, f.ts, f.rng, 1}));
  CK(launch(B, sR, KRngDraw{f.rng, f.rand1, (int)(3 + 2 * n), B}));
  return BP_OK;
}

static int prove_phase_b(const BpGens *g, BpCircuit *c, Front &f, const ProveArgs &A, dev_stream s);

// Whole batch in two calls.  begin: phase A of every chunk (all chunks concurrently, on the slot's side streams, ordered after
// what is already enqueued on s).  finish: phase B chunk by chunk on the caller's stream.  With two slots the caller can
// enqueue begin(batch k+1) before finish(batch k): the latency-bound phase A of the next batch then overlaps the
// throughput-bound phase B of the current one instead of being exposed at the head of every batch.
static ProveArgs prove_slice(const BpCircuit *c, const ProveArgs &A, int chunk, int ci) {
  const size_t m = c->m, n = c->n, plen = circuit_proof_len(c);
  ProveArgs a = A;
  const size_t p0 = (size_t)ci * chunk;
  a.B = (int)std::min<size_t>(chunk, A.B - p0);
  a.v = A.v + p0 * m * 32; a.vbl = A.vbl + p0 * m * 32; a.entropy = A.entropy + p0 * 32;
  a.aux = A.aux ? A.aux + p0 * c->naux * 32 : nullptr; a.pub = A.pub ? A.pub + p0 * c->npub * 32 : nullptr;
  if (A.aL) { a.aL = A.aL + p0 * n * 32; a.aR = A.aR + p0 * n * 32; a.aO = A.aO + p0 * n * 32; }
  a.V_out = A.V_out + p0 * m * 32; a.proofs = A.proofs + p0 * plen; a.status = A.status + p0;
  return a;
}
int engine_prove_begin(const BpGens *g, BpCircuit *c, int slot, const ProveArgs &A, int chunk, dev_stream s) {
  const int B = A.B;
  if (slot < 0 || slot > 1 || B <= 0) return BP_ERR_INVALID_ARGUMENT;
  if (g->capacity < c->n || g->capacity < c->N) return BP_ERR_INVALID_GENERATORS_LENGTH;
  if (!A.aL && !c->has_tape) return BP_ERR_MISSING_ASSIGNMENT;
  if (!A.aL && c->wit_uses_pub && !A.pub) return BP_ERR_MISSING_ASSIGNMENT;  // the witness program reads public inputs nobody supplied
  if (c->ws->pending[slot].active) return BP_ERR_INVALID_ARGUMENT;
  if (chunk <= 0 || chunk > B) chunk = B;
  int rc = ensure_workspace(c, chunk);
  if (rc) return rc;
  const int nchunks = (B + chunk - 1) / chunk;
  PendingProve &P = c->ws->pending[slot];
  P.label.assign(A.label, A.label + A.label_len);
  P.a = A; P.a.label = P.label.data(); P.chunk = chunk;
  for (int ci = 0; ci < nchunks; ci++) {
    rc = ensure_front(c, slot, ci, chunk); if (rc) return rc;
    rc = prove_phase_a(g, c, c->ws->fronts[slot][ci], prove_slice(c, P.a, chunk, ci), s); if (rc) return rc;
  }
  P.active = true;
  return BP_OK;
}
int engine_prove_finish(const BpGens *g, BpCircuit *c, int slot, dev_stream s) {
  if (slot < 0 || slot > 1 || !c->ws->pending[slot].active) return BP_ERR_INVALID_ARGUMENT;
  PendingProve &P = c->ws->pending[slot];
  P.active = false;
  const ProveArgs &A = P.a;
  const int B = A.B, chunk = P.chunk, nchunks = (B + chunk - 1) / chunk;
  const size_t plen = circuit_proof_len(c);
  int rc;
  CK(dev_memset(A.status, 0, sizeof(int) * B, s));
  CK(dev_memset(A.proofs, 0, plen * B, s));
  for (int ci = 0; ci < nchunks; ci++) {
    Front &f = c->ws->fronts[slot][ci];
#ifndef BP_HOST_EMUL
    if (g_profile_on == 1) profile_begin("@wait_first_phase", 0, s);  // pseudo-kernel: time the caller's stream spends waiting for this chunk's first phase
#endif
    dev_side_join(f.sideR, s); dev_side_join(f.sideW, s);
#ifndef BP_HOST_EMUL
    if (g_profile_on == 1) profile_end(s);
#endif
    rc = prove_phase_b(g, c, f, prove_slice(c, A, chunk, ci), s); if (rc) return rc;
  }
  return BP_OK;
}
int engine_prove(const BpGens *g, BpCircuit *c, const ProveArgs &A, int chunk, dev_stream s) {
  if (A.B <= 0) return BP_OK;
  const int slot = c->ws->pending[0].active ? 1 : 0;  // a streamed batch may be in flight in the other slot
  int rc = engine_prove_begin(g, c, slot, A, chunk, s);
  if (rc) return rc;
  return engine_prove_finish(g, c, slot, s);
}

static int prove_phase_b(const BpGens *g, BpCircuit *c, Front &f, const ProveArgs &A, dev_stream s) {
  const int B = A.B;
  int rc;
  Workspace *w = c->ws;
  const long n = c->n, N = c->N, m = c->m, q = c->q, k = c->k;
  const long plen = (long)circuit_proof_len(c);
  scm *aL = f.wit, *aR = f.wit + n * B, *aO = f.wit + 2 * n * B;
  scm *i_b = f.rand1, *sL = f.rand1 + 3L * B, *sR = f.rand1 + (3 + n) * B;
  scm *wL = w->w_all, *wR = w->w_all + n * B, *wO = w->w_all + 2 * n * B, *wV = w->w_all + 3 * n * B;
  scm *ch_y = w->chal, *ch_z = w->chal + B, *ch_yinv = w->chal + 2L * B, *ch_u = w->chal + 3L * B, *ch_x = w->chal + 4L * B,
      *ch_w = w->chal + 5L * B, *ipa_u = w->chal + 6L * B, *ipa_uinv = w->chal + 7L * B, *alpha = w->chal + 8L * B, *beta = w->chal + 9L * B;
  // 4. A_I1, A_O1, S1 (A.3 step 4)
  {
    const long rowsI = 2 * n + 1, rowsO = n + 1;
    const bool sorted = g->sg != nullptr && rowsI >= SORTED_MIN_ROWS;
    const long rb = sorted ? SB_ROW_BYTES : 32;
    int8_t *dI = w->dig, *dO = dI + rowsI * rb * B, *dS = dO + rowsO * rb * B;
    // witness written by the tape's Poseidon block op: rows with equal scalars by construction share one (summed) generator
    const int merged = (sorted && !A.aL) ? ensure_merged_ai(const_cast<BpGens *>(g), c, s) : 0;
    const uint8_t *skipI = merged ? w->skip_ai : nullptr;
    if (sorted) {
      CK(launch(B, s, KRecode13{i_b, nullptr, 1, B, dI, rowsI * rb, 0, nullptr}));
      CK(launch(n * B, s, KRecode13{aL, nullptr, (int)n, B, dI, rowsI * rb, 1, skipI}));
      CK(launch(n * B, s, KRecode13{aR, nullptr, (int)n, B, dI, rowsI * rb, (int)(1 + n), skipI}));
      if (g->table) {  // A_O stays on the direct tables: with inverse S-boxes its scalars are 0/1 (one addition per row, no bucket reduction)
        CK(launch(B, s, KRecode{i_b + B, nullptr, 1, B, dO, rowsO * 32, 0}));
        CK(launch(n * B, s, KRecode{aO, nullptr, (int)n, B, dO, rowsO * 32, 1}));
      } else {
        CK(launch(B, s, KRecode13{i_b + B, nullptr, 1, B, dO, rowsO * rb, 0, nullptr}));
        CK(launch(n * B, s, KRecode13{aO, nullptr, (int)n, B, dO, rowsO * rb, 1, nullptr}));
      }
      CK(launch(B, s, KRecode13{i_b + 2L * B, nullptr, 1, B, dS, rowsI * rb, 0, nullptr}));
      CK(launch(n * B, s, KRecode13{sL, nullptr, (int)n, B, dS, rowsI * rb, 1, nullptr}));
      CK(launch(n * B, s, KRecode13{sR, nullptr, (int)n, B, dS, rowsI * rb, (int)(1 + n), nullptr}));
    } else {
      CK(launch(B, s, KRecode{i_b, nullptr, 1, B, dI, rowsI * rb, 0}));
      CK(launch(n * B, s, KRecode{aL, nullptr, (int)n, B, dI, rowsI * rb, 1}));
      CK(launch(n * B, s, KRecode{aR, nullptr, (int)n, B, dI, rowsI * rb, (int)(1 + n)}));
      CK(launch(B, s, KRecode{i_b + B, nullptr, 1, B, dO, rowsO * rb, 0}));
      CK(launch(n * B, s, KRecode{aO, nullptr, (int)n, B, dO, rowsO * rb, 1}));
      CK(launch(B, s, KRecode{i_b + 2L * B, nullptr, 1, B, dS, rowsI * rb, 0}));
      CK(launch(n * B, s, KRecode{sL, nullptr, (int)n, B, dS, rowsI * rb, 1}));
      CK(launch(n * B, s, KRecode{sR, nullptr, (int)n, B, dS, rowsI * rb, (int)(1 + n)}));
    }
    if (g->table || sorted) {
      if (w->rg_cap != (long)g->capacity) {
        std::vector<uint32_t> rg(2 * n + 1);
        rg[0] = 2 * g->capacity + 1;
        for (long i = 0; i < n; i++) { rg[1 + i] = (uint32_t)i; rg[1 + n + i] = g->capacity + (uint32_t)i; }
        CK(dev_h2d(w->rg_as, rg.data(), rg.size() * sizeof(uint32_t), s)); CK(dev_sync(s));
        w->rg_cap = g->capacity;
      }
      RowMap rm{0, w->rg_as, (long)g->capacity, 0, 0, 0};
      if (sorted) {
        RowMap rmI{0, merged ? w->rg_ai : w->rg_as, (long)g->capacity, 0, 0, 0};
        rc = run_msm_sorted(g, w, rmI, rowsI, B, dI, rowsI * rb, A.proofs + 0, plen, s); if (rc) return rc;
        if (g->table) { rc = run_msm_table(g, w, rm, rowsO, B, dO, rowsO * 32, A.proofs + 32, plen, s); if (rc) return rc; }
        else { rc = run_msm_sorted(g, w, rm, rowsO, B, dO, rowsO * rb, A.proofs + 32, plen, s); if (rc) return rc; }  // no direct tables: shift table
        rc = run_msm_sorted(g, w, rm, rowsI, B, dS, rowsI * rb, A.proofs + 64, plen, s); if (rc) return rc;
      } else {
        rc = run_msm_table(g, w, rm, rowsI, B, dI, rowsI * rb, A.proofs + 0, plen, s); if (rc) return rc;
        rc = run_msm_table(g, w, rm, rowsO, B, dO, rowsO * rb, A.proofs + 32, plen, s); if (rc) return rc;
        rc = run_msm_table(g, w, rm, rowsI, B, dS, rowsI * rb, A.proofs + 64, plen, s); if (rc) return rc;
      }
    } else {
      MsmSeg segs[3] = {{g->pc_niels + 1, 0, 0, 1}, {g->G_n, 0, 0, (int)n}, {g->H_n, 0, 0, (int)n}};
      rc = run_msm(w, segs, 3, B, dI, rowsI * 32, A.proofs + 0, plen, 0, nullptr, s); if (rc) return rc;
      rc = run_msm(w, segs, 2, B, dO, rowsO * 32, A.proofs + 32, plen, 0, nullptr, s); if (rc) return rc;
      rc = run_msm(w, segs, 3, B, dS, rowsI * 32, A.proofs + 64, plen, 0, nullptr, s); if (rc) return rc;
    }
  }
  // 5. y, z; powers; flattened weights (A.3 steps 5-7)
  CK(launch(B, s, KTsPhase2{f.ts, A.proofs, plen, ch_y, ch_z, ch_yinv, A.status, 0}));
  CK(launch(((q + CH_POW - 1) / CH_POW) * B, s, KPowers{ch_z, w->zpow, (int)q, B, 1, CH_POW}));
  CK(launch(((N + CH_POW - 1) / CH_POW) * B, s, KPowers{ch_y, w->ypow, (int)N, B, 0, CH_POW}));
  CK(launch(((N + CH_POW - 1) / CH_POW) * B, s, KPowers{ch_yinv, w->yinvpow, (int)N, B, 0, CH_POW}));
  rc = run_flatten(c, w, B, s); if (rc) return rc;
  // 6. t(x) coefficients, T commitments (A.3 steps 8-10)
  PolyIn pin{aL, aR, aO, sL, sR, wL, wR, wO, w->ypow, w->yinvpow};
  {
    long nch = (n + CH_DOT - 1) / CH_DOT;
    if (nch == 0) nch = 1;
    CK(launch(nch * B, s, KPolyT{pin, (int)n, B, CH_DOT, w->part}));
    CK(launch(6L * B, s, KSumPartials{w->part, (int)nch, 6, B, w->t}));
  }
  CK(launch(B, s, KRngDraw{f.rng, w->tb, 5, B}));
  {
    // T_1,T_3,T_4,T_5,T_6 use t[0],t[2],t[3],t[4],t[5]: commit rows individually
    const int tj[5] = {0, 2, 3, 4, 5};
    for (int j = 0; j < 5; j++)
      CK(launch(B, s, KCommit{w->t + (long)tj[j] * B, w->tb + (long)j * B, 1, B, g->pc_table, A.proofs + 192 + 32 * j, plen, 0, nullptr}));
  }
  CK(launch(B, s, KTsPhase3{f.ts, A.proofs, plen, ch_u, ch_x, A.status, 0}));
  // 7. evaluate l, r at x; scalars of the proof (A.3 steps 11-13); w and Q (step 14)
  CK(launch(N * B, s, KPolyEval{pin, (int)n, B, ch_x, w->a, w->b}));
  CK(launch(B, s, KProverScalars{w->t, w->tb, i_b, wV, f.vbl, (int)m, B, ch_x, A.proofs, plen}));
  CK(launch(B, s, KTsPhase4{f.ts, A.proofs, plen, ch_w, (unsigned)N}));
  CK(launch(B, s, KCommit{ch_w, nullptr, 1, B, g->pc_table, nullptr, 0, 0, w->Q}));
  // 8. inner-product argument (A.4).  With fixed-base tables the first unfold_rounds() rounds never fold generators: L_j, R_j are
  // multiscalar multiplications over the ORIGINAL generators with scalars a_i * prod u_t^(+-1); the folded generators are then
  // materialised once (KFoldTable) and the remaining, short rounds run on per-proof points (bucket method + NAF fold).
  CK(launch(2L * B, s, KFillScalar{alpha, sc_one()}));  // alpha, beta are adjacent
  FoldTableView ft{};
  const bool can_unfold = g->table || (g->sg != nullptr && N + 1 >= SORTED_MIN_ROWS);  // L_j, R_j over the original generators: direct or shift tables
  if (can_unfold && !g->table && k > 1 && unfold_rounds() > 0) { rc = ensure_fold_table(const_cast<BpGens *>(g), N, s); if (rc) return rc; }
  const int J = can_unfold && fold_table_view(g, N, ft) ? (int)std::min<long>(unfold_rounds(), k) : 0;
  scm *UG[2] = {w->utab, w->utab + ((size_t)1 << UNFOLD_MAX) * B}, *UH[2] = {w->utab + 2 * ((size_t)1 << UNFOLD_MAX) * B, w->utab + 3 * ((size_t)1 << UNFOLD_MAX) * B};
  if (J > 0) { CK(launch(B, s, KFillScalar{UG[0], sc_one()})); CK(launch(B, s, KFillScalar{UH[0], sc_one()})); }
  int yfree = 0;
  long len = N;
  const long gs = N / 2 + 1;
  for (int round = 0; round < k; round++) {
    const long h = len / 2;
    long nch = (h + CH_DOT - 1) / CH_DOT;
    CK(launch(nch * B, s, KIpaDots{w->a, w->b, (int)h, B, CH_DOT, w->part}));
    CK(launch(2L * B, s, KSumPartials{w->part, (int)nch, 2, B, w->clr}));
    if (round < J) {
      const long rows = N + 1;
      const bool sorted = g->sg != nullptr && rows >= SORTED_MIN_ROWS;
      const long rb = sorted ? SB_ROW_BYTES : 32;
      int8_t *dL = w->dig, *dR = w->dig + rows * rb * B;
      const int cur = round & 1;
      RowMap rl{1, nullptr, (long)g->capacity, N, len, h}, rr{2, nullptr, (long)g->capacity, N, len, h};
      if (sorted) {
        // round 0 of a padded circuit: one extra row (N + 1) stands for the N - n H-rows of L_0 that share the scalar -y^h
        const long pad_gen = (round == 0 && (size_t)(N + 2) * SB_WINDOWS <= w->items_cap) ? ensure_pad_generator(const_cast<BpGens *>(g), n, N, s) : -1;
        const long srows = pad_gen >= 0 ? rows + 1 : rows;
        int8_t *dRs = w->dig + srows * rb * B;
        rl.pad_gen = rr.pad_gen = pad_gen;
        CK(launch((N / 2) * B, s, KRecodeUnfolded13{w->a, w->b, UG[cur], UH[cur], w->yinvpow, ch_u, w->clr, ch_w, N, len, h, n, B, dL, dRs, srows * rb,
                                                    pad_gen >= 0 ? 1 : 0, w->ypow + h * B}));
        rc = run_msm_sorted(g, w, rl, srows, B, dL, srows * rb, A.proofs + 448 + 64 * round, plen, s); if (rc) return rc;
        rc = run_msm_sorted(g, w, rr, srows, B, dRs, srows * rb, A.proofs + 448 + 64 * round + 32, plen, s); if (rc) return rc;
      } else {
        CK(launch((N / 2) * B, s, KRecodeUnfolded{w->a, w->b, UG[cur], UH[cur], w->yinvpow, ch_u, w->clr, ch_w, N, len, h, n, B, dL, dR, rows * rb}));
        rc = run_msm_table(g, w, rl, rows, B, dL, rows * rb, A.proofs + 448 + 64 * round, plen, s); if (rc) return rc;
        rc = run_msm_table(g, w, rr, rows, B, dR, rows * rb, A.proofs + 448 + 64 * round + 32, plen, s); if (rc) return rc;
      }
      CK(launch(B, s, KTsIpaRound{f.ts, A.proofs, plen, round, B, (int)h, w->yinvpow, ch_u, ipa_u, ipa_uinv, alpha, beta, w->naf, w->naf_top,
                                 A.status, 1, 0}));  // transcript + u, u^-1 only (verifier mode skips the fold scalars)
      CK(launch(h * B, s, KFoldAB{w->a, w->b, ipa_u, ipa_uinv, (int)h, B}));
      CK(launch((2L << round) * B, s, KIpaUTable{ipa_u, ipa_uinv, UG[cur], UH[cur], UG[cur ^ 1], UH[cur ^ 1], B}));
      if (round == J - 1 && h > 1) {
        // folded generators of level J straight from the tables: H side true (beta = 1, the y^-i factors are inside), G side
        // divided by UG[0] so that the first of its 2^J terms is the generator itself; alpha = UG[0] carries the factor
        if ((size_t)2 * N * ft.rb * B > w->dig_bytes) return BP_ERR_OOM;
        KRecodeFoldTable kr{UG[cur ^ 1], UH[cur ^ 1], w->yinvpow, ch_u, N, h, n, B, w->dig, 2 * N * ft.rb};
        kr.bits = ft.bits; kr.W = ft.W; kr.rb = ft.rb;
        CK(launch(N * B, s, kr));
        KFoldTable kf{ft.table, ft.cap, N, h, w->dig, 2 * N * ft.rb, w->Gt, w->Ht, gs, g->G_p3};
        kf.W = ft.W; kf.E = ft.E; kf.rb = ft.rb;
        CK(launch(2 * h * B, s, kf));
        CK(dev_d2d(alpha, UG[cur ^ 1], sizeof(scm) * B, s));
        yfree = 1;
      }
      len = h;
      continue;
    }
    const long rows = 2 * h + 1;
    int8_t *dL = w->dig, *dR = w->dig + rows * 32 * B;
    CK(launch(h * B, s, KRecodeIpa{w->a, w->b, alpha, beta, w->yinvpow, ch_u, w->clr, (int)h, B, (int)n, round, dL, dR, rows * 32, yfree}));
    MsmSeg sL_[3], sR_[3];
    if (round == 0) {
      sL_[0] = {g->G_n + h, 0, 0, (int)h}; sL_[1] = {g->H_n, 0, 0, (int)h};
      sR_[0] = {g->G_n, 0, 0, (int)h};     sR_[1] = {g->H_n + h, 0, 0, (int)h};
    } else {
      sL_[0] = {w->Gt + h, gs, 1, (int)h}; sL_[1] = {w->Ht, gs, 1, (int)h};
      sR_[0] = {w->Gt, gs, 1, (int)h};     sR_[1] = {w->Ht + h, gs, 1, (int)h};
    }
    sL_[2] = {w->Q, 1, 1, 1}; sR_[2] = sL_[2];
    rc = run_msm(w, sL_, 3, B, dL, rows * 32, A.proofs + 448 + 64 * round, plen, 0, nullptr, s); if (rc) return rc;
    rc = run_msm(w, sR_, 3, B, dR, rows * 32, A.proofs + 448 + 64 * round + 32, plen, 0, nullptr, s); if (rc) return rc;
    CK(launch(B, s, KTsIpaRound{f.ts, A.proofs, plen, round, B, (int)h, w->yinvpow, ch_u, ipa_u, ipa_uinv, alpha, beta, w->naf, w->naf_top,
                               A.status, 0, yfree}));
    CK(launch(h * B, s, KFoldAB{w->a, w->b, ipa_u, ipa_uinv, (int)h, B}));
    if (h > 1) {
      if (round == 0) CK(launch(2 * h * B, s, KFoldGens{g->G_p3, g->H_p3, 0, w->Gt, w->Ht, gs, w->naf, w->naf_top, (int)h, (int)n, round}));
      else CK(launch(2 * h * B, s, KFoldGens{w->Gt, w->Ht, gs, w->Gt, w->Ht, gs, w->naf, w->naf_top, (int)h, (int)n, round}));
    }
    len = h;
  }
  CK(launch(B, s, KStoreAB{w->a, w->b, A.proofs, plen, 448 + 64 * k}));
  if (const char *dump = getenv("BP_B200_DEBUG_DUMP")) {  // developer aid: scalars of the chunk as canonical bytes
    CK(dev_sync(s));
    FILE *fp = fopen(dump, "wb");
    if (fp) {
      auto put = [&](const char *name, const scm *d, long count) {
        std::vector<scm> h(count); dev_d2h(h.data(), d, count * sizeof(scm), s); dev_sync(s);
        std::vector<uint8_t> bts(count * 32); for (long i = 0; i < count; i++) sc_tobytes(bts.data() + 32 * i, h[i]);
        char hdr[32] = {0}; snprintf(hdr, sizeof hdr, "%s", name); fwrite(hdr, 1, 24, fp); uint64_t cnt = count; fwrite(&cnt, 8, 1, fp); fwrite(bts.data(), 1, bts.size(), fp);
      };
      put("rand1", f.rand1, (3 + 2 * n) * B); put("chal", w->chal, 10L * B); put("t", w->t, 6L * B); put("tb", w->tb, 5L * B);
      put("wit", f.wit, 3 * n * B); put("w_all", w->w_all, (long)c->nslots * B); put("v", f.v, m * B); put("vbl", f.vbl, m * B);
      fclose(fp);
    }
  }
  return BP_OK;
}

int engine_commit(const BpGens *g, int count, const uint8_t *v, const uint8_t *r, uint8_t *out) {
  if (count <= 0) return BP_OK;
  uint8_t *d_in = nullptr, *d_out = nullptr; scm *d_s = nullptr;
  CK(dalloc(&d_in, (size_t)count * 64)); CK(dalloc(&d_out, (size_t)count * 32)); CK(dalloc(&d_s, (size_t)count * 2));
  dev_stream s = 0;
  CK(dev_h2d(d_in, v, (size_t)count * 32, s)); CK(dev_h2d(d_in + (size_t)count * 32, r, (size_t)count * 32, s));
  // treat as B = count proofs with one commitment each
  CK(launch(count, s, KLoadScalars{d_in, d_s, 1, count}));
  CK(launch(count, s, KLoadScalars{d_in + (size_t)count * 32, d_s + count, 1, count}));
  CK(launch(count, s, KCommit{d_s, d_s + count, 1, count, g->pc_table, d_out, 32, 0, nullptr}));
  CK(dev_d2h(out, d_out, (size_t)count * 32, s)); CK(dev_sync(s));
  dev_free(d_in); dev_free(d_out); dev_free(d_s);
  return BP_OK;
}

// ------------------------------------------------------------------------------------------------ MSM microbenchmark entry
// One MSM over `nrows` rows of shared generators (row -> generator through `mode`: 3 = chain G in order, 5 = the combined
// verification layout G.., H.., B, B_blinding).  The rows are split into sub-instances of R consecutive rows that go through the
// batched sorted-bucket kernels (sort, accumulate, reduce) like the proofs of a batch do; their partial results are summed.
// scalars: d_bytes (canonical 32-byte scalars) or d_scm (Montgomery form already on the device); result encoded to d_out or
// kept as a point in out_p3.  Uses the generator set's own scratch workspace.
// buckets per thread of KBucketReduceW: short segments while the sub-instances are few (latency), longer ones once there are
// enough threads to fill the device anyway (the weight's double-and-add is paid once per segment)
static int split_seg_len(long S) { int L = 8; while (L < 128 && (long)L * 8 <= S) L *= 2; return L; }
static const int SPLIT_SEG_LEN = 8;  // the shortest: sizes the buffers
static int sorted_split_msm(BpGens *g, long nrows, const uint8_t *d_bytes, const scm *d_scm, int mode, long mapN, uint8_t *d_out, ge_p3 *out_p3,
                            dev_stream s) {
  long R = nrows / 64; if (R < 8192) R = 8192; if (R > 32768) R = 32768; if (R > nrows) R = nrows;
  const long S = (nrows + R - 1) / R;
  if (!g->msm_ws || (long)g->msm_ws_n < nrows || !g->msm_ws->items) {
    if (g->msm_ws) { g->msm_ws->release(); delete g->msm_ws; }
    Workspace *w = g->msm_ws = new Workspace();
    w->items_cap = (size_t)R * SB_WINDOWS;
    w->slices_cap = SB_BUCKETS + w->items_cap / SB_SEG + 2;
    w->bucket_slots = (size_t)S * w->slices_cap + S + 1;
    // the partial-sum buffer doubles as the scratch of the two-pass sort (one 4-byte item per digit)
    w->bucket_slots = std::max(w->bucket_slots, (size_t)S * ((w->items_cap * sizeof(uint32_t) + sizeof(ge_p3) - 1) / sizeof(ge_p3)));
    w->dig_bytes = (size_t)S * R * SB_ROW_BYTES;
    const size_t tree = (size_t)S * (SB_BUCKETS / SPLIT_SEG_LEN);  // one point per segment, then the stages of the plain sum
    if (dalloc(&w->a, (size_t)S * R) || dalloc(&w->dig, w->dig_bytes) || dalloc(&w->buckets, w->bucket_slots) || dalloc(&w->items, w->items_cap * S) ||
        dalloc(&w->boff, (size_t)(SB_BUCKETS + 1) * S) || dalloc(&w->soff, (size_t)(SB_BUCKETS + 1) * S) || dalloc(&w->seg, tree + tree / 8 + 64)) {
      w->release(); delete w; g->msm_ws = nullptr; return BP_ERR_OOM;
    }
    w->seg_cap = tree + tree / 8 + 64;
    g->msm_ws_n = (uint32_t)nrows;
  }
  Workspace *w = g->msm_ws;
  if ((size_t)S * (SB_BUCKETS / SPLIT_SEG_LEN) * 9 / 8 + 64 > w->seg_cap || (size_t)S * R * SB_ROW_BYTES > w->dig_bytes || (size_t)R * SB_WINDOWS > w->items_cap || (size_t)S * w->slices_cap + S + 1 > w->bucket_slots) {
    g->msm_ws_n = 0;  // geometry of an earlier, different size: rebuild
    return sorted_split_msm(g, nrows, d_bytes, d_scm, mode, mapN, d_out, out_p3, s);
  }
  CK(dev_memset(w->dig, 0, (size_t)S * R * SB_ROW_BYTES, s));  // rows past nrows stay all-zero digits
  if (d_bytes) { CK(launch(nrows, s, KLoadScalars{d_bytes, w->a, (int)nrows, 1})); d_scm = w->a; }
  CK(launch(nrows, s, KRecode13{d_scm, nullptr, (int)nrows, 1, w->dig, 0, 0, nullptr}));  // B = 1: row i at dig + i * SB_ROW_BYTES = sub-instance i / R, row i % R
  RowMap rm{mode, nullptr, (long)g->capacity, mapN, 0, 0, R};
  // chain G in order (mode 3): items carry the generator index relative to the sub-instance's first row, so 2^22 rows still sort in two passes
  rm.rel = mode == 3 && !getenv("BP_B200_NO_REL");
  const long gen_slots = rm.rel ? R : 2L * g->capacity + 2 + SG_SPARE + g->merge_slots;
  CK(launch_sort_buckets(rm, w->dig, R * SB_ROW_BYTES, R, S, w->items, (long)w->items_cap, w->boff, w->soff, (uint32_t *)w->buckets,
                         w->bucket_slots * sizeof(ge_p3), gen_slots, s));
  SortedView sv{w->items, w->boff, w->soff, (long)w->items_cap, (long)w->slices_cap};
  sv.base_stride = rm.rel ? R * SB_WINDOWS : 0;
  const long segs = (R * SB_WINDOWS + SB_SEG - 1) / SB_SEG;
  if (rm.rel) CK(launch(S * segs, s, KBucketAccumulateRel{g->sg, sv, w->buckets, segs}));
  else CK(launch(S * segs, s, KBucketAccumulate{g->sg, sv, w->buckets, segs}));
  // at most 128 sub-instances: short segments with their weights applied in place, then a plain sum over all of them (see KBucketReduceW)
  const int L = split_seg_len(S);
  long count = S * (SB_BUCKETS / L);
  ge_p3 *bufA = w->seg, *bufB = w->seg + count, *cur = bufA;  // stage outputs shrink 16x: ping-pong between the two
  CK(launch(count, s, KBucketReduceW{w->buckets, sv, cur, L}));
  while (count > 16) {
    const long T = (count + 15) / 16;
    ge_p3 *dst = cur == bufA ? bufB : bufA;
    CK(launch(T, s, KSumPointsStrided{cur, count, T, dst}));
    cur = dst; count = T;
  }
  if (out_p3) CK(launch(1, s, KSumPointsStrided{cur, count, 1, out_p3}));
  else CK(launch(1, s, KSumPointsEncode{cur, (int)count, d_out}));
  return BP_OK;
}
static int msm_gens_sorted(BpGens *g, uint32_t n, const uint8_t *d_scalars, uint8_t *d_out, dev_stream s) {
  return sorted_split_msm(g, n, d_scalars, nullptr, 3, 0, d_out, nullptr, s);
}
int engine_msm_gens(BpGens *g, uint32_t n, const uint8_t *d_scalars, uint8_t *d_out, dev_stream s) {
  if (n == 0 || n > g->capacity) return BP_ERR_INVALID_GENERATORS_LENGTH;
  if (g->sg && n >= 32768 && !getenv("BP_B200_MSM_BUCKET")) return msm_gens_sorted(g, n, d_scalars, d_out, s);
  if (!g->msm_ws || g->msm_ws_n < n || !g->msm_ws->wsum) {
    if (g->msm_ws) { g->msm_ws->release(); delete g->msm_ws; }
    Workspace *w = g->msm_ws = new Workspace();
    size_t max_warps = (size_t)msm_target_warps() + 1;
    w->bucket_slots = max_warps * MSM_WINDOWS * MSM_BUCKETS;
    if (dalloc(&w->a, n) || dalloc(&w->dig, (size_t)n * 32) || dalloc(&w->buckets, w->bucket_slots) || dalloc(&w->wsum, (max_warps + 64) * MSM_WINDOWS)) {
      w->release(); return BP_ERR_OOM;
    }
    g->msm_ws_n = n;
  }
  Workspace *w = g->msm_ws;
  CK(launch(n, s, KLoadScalars{d_scalars, w->a, (int)n, 1}));
  CK(launch(n, s, KRecode{w->a, nullptr, (int)n, 1, w->dig, (long)n * 32, 0}));
  if (g->table && !getenv("BP_B200_MSM_BUCKET")) {  // fixed-base tables cover these generators
    if (!w->rg_as || w->rg_cap < (long)n) {
      dev_free(w->rg_as); w->rg_as = nullptr;
      std::vector<uint32_t> id(n); for (uint32_t i = 0; i < n; i++) id[i] = i;
      if (dalloc(&w->rg_as, n)) return BP_ERR_OOM;
      CK(dev_h2d(w->rg_as, id.data(), n * sizeof(uint32_t), s)); CK(dev_sync(s));
      w->rg_cap = n;
    }
    RowMap rm{0, w->rg_as, (long)g->capacity, 0, 0, 0};
    return run_msm_table(g, w, rm, n, 1, w->dig, (long)n * 32, d_out, 32, s);
  }
  MsmSeg seg[1] = {{g->G_n, 0, 0, (int)n}};
  return run_msm(w, seg, 1, 1, w->dig, (long)n * 32, d_out, 32, 0, nullptr, s);
}

// ------------------------------------------------------------------------------------------------ combined verification
// Cross-proof batched verification: the B verification equations  sum_j s_{p,j} P_{p,j} = 0  are combined with
// random weights rho_p drawn from each proof's verifier RNG (so they depend on the verifier's entropy and on the whole
// transcript of the proof):  sum_p rho_p * (equation p) = 0.  The generator rows (G_i, H_i, B, B_blinding) are shared by all
// proofs, so their weighted scalars are SUMMED over the batch and the 2N + 2 rows are paid once per batch instead of once per
// proof; only the ~110 points that belong to a proof (A_*, S, V_j, T_*, L_j, R_j) remain per-proof work.  A batch of valid
// proofs always passes; a batch containing an invalid proof passes with probability 2^-252.  Proofs failing a structural check
// (non-canonical scalar, undecodable or identity point) get their status set and weight zero.  *combined = 0 or BP_ERR_VERIFICATION.
int engine_verify_combined(BpGens *g, BpCircuit *c, const VerifyArgs &A, int *d_combined, dev_stream s) {
  const int B = A.B;
  if (B <= 0) return BP_OK;
  if (g->capacity < c->N) return BP_ERR_INVALID_GENERATORS_LENGTH;
  if (c->npub && !A.pub) return BP_ERR_MISSING_ASSIGNMENT;
  if (!g->sg) return BP_ERR_INVALID_ARGUMENT;  // needs the shift table of the sorted-bucket path
  int rc = ensure_workspace(c, B);
  if (rc) return rc;
  Workspace *w = c->ws;
  const long n = c->n, N = c->N, m = c->m, q = c->q, k = c->k;
  const long plen = (long)circuit_proof_len(c);
  const long npts = 11 + m + 2 * k, chunks = (B + 31) / 32, Bp = chunks * 32;
  if ((size_t)B * MSM_WINDOWS * MSM_BUCKETS > w->bucket_slots || (size_t)B * npts * 32 > w->dig_bytes) return BP_ERR_OOM;
  scm *wL = w->w_all, *wR = w->w_all + n * B, *wO = w->w_all + 2 * n * B, *wV = w->w_all + 3 * n * B, *wc = w->w_all + (3 * n + m) * B,
      *wP = w->w_all + (3 * n + m + 1) * B;
  scm *ch_z = w->chal + B, *ch_yinv = w->chal + 2L * B, *rho = w->chal + 7L * B, *rowB = w->chal + 8L * B, *rowBb = w->chal + 9L * B;
  scm *uj = w->uj, *ujinv = w->uj + (k + 1) * B;
  // the N x chunks partial sums of the G and H scalars reuse the b vector and the y^i table (not needed by a verifier); the
  // 2N + 2 combined rows sit behind the 8-bit digit rows of the per-proof points
  const size_t dig8_bytes = (((size_t)B * npts * 32) + 31) & ~(size_t)31;
  if (dig8_bytes + (size_t)(2 * N + 2) * sizeof(scm) > w->dig_bytes) return BP_ERR_OOM;
  scm *partG = w->b, *partH = w->ypow, *rows = (scm *)(w->dig + dig8_bytes);
  CK(dev_memset(A.status, 0, sizeof(int) * B, s));
  if (c->nfixed) CK(launch((long)c->nfixed * B, s, KCheckFixedCommitments{A.V, (int)m, B, c->d_fixed_idx, c->d_fixed_V, (int)c->nfixed, A.status}));
  strobe128 base; base_transcript(base, A.label, A.label_len);
  // per-proof digests and the batch seed live in the (not yet used) partial-sum buffer
  uint8_t *digest = (uint8_t *)w->part, *seed = digest + (size_t)B * 32;
  CK(launch(B, s, KTsVerify{base, A.V, (int)m, B, (int)k, (unsigned)N, A.proofs, plen, A.entropy, w->chal, uj, ujinv, A.status, digest, A.pub, (int)c->npub}));
  CK(launch(npts * B, s, KVerifyDecompress{A.V, A.proofs, plen, (int)m, (int)k, B, w->pts, npts, A.status}));
  CK(launch(1, s, KBatchSeed{digest, B, seed}));
  CK(launch(B, s, KVerifyRho{seed, A.status, rho}));
  CK(launch(((q + CH_POW - 1) / CH_POW) * B, s, KPowers{ch_z, w->zpow, (int)q, B, 1, CH_POW}));
  CK(launch(((N + CH_POW - 1) / CH_POW) * B, s, KPowers{ch_yinv, w->yinvpow, (int)N, B, 0, CH_POW}));
  rc = run_flatten(c, w, B, s); if (rc) return rc;
  if (c->npub) CK(launch((long)c->npub * B, s, KLoadScalars{A.pub, w->vpub, (int)c->npub, B}));
  CK(launch(N * B, s, KVerifyS{uj, ujinv, (int)k, B, w->a}));
  {
    long nch = (n + CH_DOT - 1) / CH_DOT; if (nch == 0) nch = 1;
    CK(launch(nch * B, s, KVerifyDelta{wL, wR, w->yinvpow, (int)n, B, CH_DOT, w->part}));
    CK(launch(B, s, KSumPartials{w->part, (int)nch, 1, B, w->clr}));
  }
#ifdef BP_HOST_EMUL
  CK(dev_memset(partG, 0, sizeof(scm) * chunks * N, s)); CK(dev_memset(partH, 0, sizeof(scm) * chunks * N, s));
#endif
  CK(launch(N * Bp, s, KVerifyGH{wL, wR, wO, w->yinvpow, w->a, w->chal, A.proofs, plen, (int)n, (int)N, (int)k, B, nullptr, 0, nullptr, nullptr, rho, partG, partH}));
  CK(launch(B, s, KVerifyScalars{w->chal, uj, ujinv, wV, wc, wP, w->vpub, w->clr, A.proofs, plen, (int)m, (int)c->npub, (int)N, (int)k, B, w->dig, npts * 32,
                                 nullptr, nullptr, rho, rowB, rowBb}));
  // rows of the single shared-generator MSM (KFlatten is done with the z^q table by now)
  CK(launch(2 * N + 2, s, KVerifyCombineRows{partG, partH, rowB, rowBb, (int)N, B, rows}));
  // per-proof points: bucket method per proof, Horner per proof, then a two-level sum
  MsmSeg seg[1] = {{w->pts, npts, 1, (int)npts}};
  KMsmAccumulate ka{};
  ka.seg[0] = seg[0]; ka.nseg = 1; ka.S = 1; ka.dig = w->dig; ka.dig_inst_stride = npts * 32; ka.buckets = w->buckets; ka.wsum = w->wsum;
  CK(launch((long)B * MSM_WINDOWS, s, ka));
  ge_p3 *pp = w->buckets + (size_t)B * MSM_WINDOWS * MSM_BUCKETS;  // [B] per-proof parts, then [T + 1] partial sums, behind the buckets in use
  const long T = B < 64 ? B : 64;
  if ((size_t)B * MSM_WINDOWS * MSM_BUCKETS + (size_t)B + T + 1 > w->bucket_slots) return BP_ERR_OOM;
  CK(launch(B, s, KVerifyProofPoint{w->wsum, pp}));
  CK(launch(T, s, KSumPointsStrided{pp, B, T, pp + B}));
  rc = sorted_split_msm(g, 2 * N + 2, nullptr, rows, 5, N, nullptr, pp + B + T, s); if (rc) return rc;
  CK(launch(1, s, KVerifyCombinedCheck{pp + B, (int)(T + 1), d_combined}));
  return BP_OK;
}

// ------------------------------------------------------------------------------------------------ verifier (A.5)
int engine_verify(const BpGens *g, BpCircuit *c, const VerifyArgs &A, dev_stream s) {
  const int B = A.B;
  if (B <= 0) return BP_OK;
  if (g->capacity < c->N) return BP_ERR_INVALID_GENERATORS_LENGTH;
  if (c->npub && !A.pub) return BP_ERR_MISSING_ASSIGNMENT;
  int rc = ensure_workspace(c, B);
  if (rc) return rc;
  Workspace *w = c->ws;
  const long n = c->n, N = c->N, m = c->m, q = c->q, k = c->k;
  const long plen = (long)circuit_proof_len(c);
  scm *wL = w->w_all, *wR = w->w_all + n * B, *wO = w->w_all + 2 * n * B, *wV = w->w_all + 3 * n * B, *wc = w->w_all + (3 * n + m) * B,
      *wP = w->w_all + (3 * n + m + 1) * B;
  scm *ch_z = w->chal + B, *ch_yinv = w->chal + 2L * B;
  scm *uj = w->uj, *ujinv = w->uj + (k + 1) * B;
  CK(dev_memset(A.status, 0, sizeof(int) * B, s));
  if (c->nfixed) CK(launch((long)c->nfixed * B, s, KCheckFixedCommitments{A.V, (int)m, B, c->d_fixed_idx, c->d_fixed_V, (int)c->nfixed, A.status}));
  strobe128 base; base_transcript(base, A.label, A.label_len);
  CK(launch(B, s, KTsVerify{base, A.V, (int)m, B, (int)k, (unsigned)N, A.proofs, plen, A.entropy, w->chal, uj, ujinv, A.status, nullptr, nullptr, 0}));
  CK(launch(((q + CH_POW - 1) / CH_POW) * B, s, KPowers{ch_z, w->zpow, (int)q, B, 1, CH_POW}));
  CK(launch(((N + CH_POW - 1) / CH_POW) * B, s, KPowers{ch_yinv, w->yinvpow, (int)N, B, 0, CH_POW}));
  rc = run_flatten(c, w, B, s); if (rc) return rc;
  if (c->npub) CK(launch((long)c->npub * B, s, KLoadScalars{A.pub, w->vpub, (int)c->npub, B}));
  CK(launch(N * B, s, KVerifyS{uj, ujinv, (int)k, B, w->a}));
  {
    long nch = (n + CH_DOT - 1) / CH_DOT; if (nch == 0) nch = 1;
    CK(launch(nch * B, s, KVerifyDelta{wL, wR, w->yinvpow, (int)n, B, CH_DOT, w->part}));
    CK(launch(B, s, KSumPartials{w->part, (int)nch, 1, B, w->clr}));
  }
  const long npts = 11 + m + 2 * k, rows = 2 + 2 * N + npts;
  // with the shift table the generator rows (B, G_i | B_blinding, H_i) are two sorted-bucket instances of N + 1 rows per proof;
  // their 15-bit digit rows live in front of the 8-bit rows of the ~110 per-proof points
  const bool sorted = g->sg != nullptr && N + 1 >= SORTED_MIN_ROWS && (size_t)(N + 1) * SB_WINDOWS <= w->items_cap &&
                      (size_t)B * (2 * (N + 1) * SB_ROW_BYTES + npts * 32) <= w->dig_bytes && (size_t)B * w->slices_cap <= w->bucket_slots &&
                      (size_t)B * MSM_WINDOWS * MSM_BUCKETS <= w->bucket_slots && N / 2 + 1 >= 2;
  int8_t *wideG = sorted ? w->dig : nullptr, *wideH = sorted ? w->dig + (size_t)B * (N + 1) * SB_ROW_BYTES : nullptr;
  int8_t *dig8 = sorted ? w->dig + (size_t)B * 2 * (N + 1) * SB_ROW_BYTES : w->dig;
  const long stride8 = sorted ? npts * 32 : rows * 32;
  CK(launch(N * B, s, KVerifyGH{wL, wR, wO, w->yinvpow, w->a, w->chal, A.proofs, plen, (int)n, (int)N, (int)k, B, dig8, stride8, wideG, wideH, nullptr, nullptr, nullptr}));
  CK(launch(B, s, KVerifyScalars{w->chal, uj, ujinv, wV, wc, wP, w->vpub, w->clr, A.proofs, plen, (int)m, (int)c->npub, (int)N, (int)k, B, dig8, stride8, wideG, wideH, nullptr, nullptr, nullptr}));
  CK(launch(npts * B, s, KVerifyDecompress{A.V, A.proofs, plen, (int)m, (int)k, B, w->pts, npts, A.status}));
  if (sorted) {
    // the ~110 per-proof points first (bucket method, result in wsum), then the two generator halves through the sorted path;
    // both use the bucket area in turn, the two half results go to the (otherwise idle) folded-generator buffer
    MsmSeg seg[1] = {{w->pts, npts, 1, (int)npts}};
    KMsmAccumulate k{};
    k.seg[0] = seg[0]; k.nseg = 1; k.S = 1; k.dig = dig8; k.dig_inst_stride = stride8; k.buckets = w->buckets; k.wsum = w->wsum;
    CK(launch((long)B * MSM_WINDOWS, s, k));
    ge_p3 *tpart = w->Gt;
    for (int half = 0; half < 2; half++) {
      RowMap rm{4, nullptr, (long)g->capacity, 0, half, 0, 0};
      const int8_t *dg = half ? wideH : wideG;
      // (the bucket area is idle between KMsmAccumulate above, which leaves its window sums in wsum, and KBucketAccumulate below: sort scratch)
      CK(launch_sort_buckets(rm, dg, (N + 1) * SB_ROW_BYTES, N + 1, B, w->items, (long)w->items_cap, w->boff, w->soff, (uint32_t *)w->buckets,
                             w->bucket_slots * sizeof(ge_p3), 2L * g->capacity + 2 + SG_SPARE + g->merge_slots, s));
      SortedView sv{w->items, w->boff, w->soff, (long)w->items_cap, (long)w->slices_cap};
      const long segs = ((N + 1) * SB_WINDOWS + SB_SEG - 1) / SB_SEG;
      CK(launch(B * segs, s, KBucketAccumulate{g->sg, sv, w->buckets, segs}));
      CK(launch((long)B * SB_SEGS, s, KBucketReduce{w->buckets, sv, w->seg}));
      ge_p3 *grp = w->seg + (size_t)B * SB_SEGS * 2;
      CK(launch((long)B * SB_FIN_GROUPS, s, KBucketFinishA{w->seg, grp}));
      CK(launch(B, s, KBucketFinishB{grp, nullptr, 0, tpart + (size_t)half * B}));
    }
    CK(launch(B, s, KVerifyCheck{w->wsum, tpart, 2, A.status, (long)B}));
    return BP_OK;
  }
  if (g->table) {
    // generator rows (B, B_blinding, G, H) from the fixed-base tables; the ~110 per-proof points by the bucket method; the two
    // partial results are added in KVerifyCheck
    if (w->rg_ver_cap != (long)g->capacity || w->rg_ver_N != N) {
      std::vector<uint32_t> rg(2 + 2 * N);
      rg[0] = 2 * g->capacity; rg[1] = 2 * g->capacity + 1;
      for (long i = 0; i < N; i++) { rg[2 + i] = (uint32_t)i; rg[2 + N + i] = g->capacity + (uint32_t)i; }
      dev_free(w->rg_ver); w->rg_ver = nullptr;
      if (dalloc(&w->rg_ver, rg.size())) return BP_ERR_OOM;
      CK(dev_h2d(w->rg_ver, rg.data(), rg.size() * sizeof(uint32_t), s)); CK(dev_sync(s));
      w->rg_ver_cap = g->capacity; w->rg_ver_N = N;
    }
    const long trows = 2 + 2 * N;
    long S = 262144 / B; if (S > trows / 16) S = trows / 16; if (S > 1024) S = 1024; if (S < 1) S = 1;
    // partial sums of the table part live behind the bucket area used by the per-proof-point MSM
    const size_t bucket_need = (size_t)B * MSM_WINDOWS * MSM_BUCKETS;
    if (bucket_need + (size_t)B * S > w->bucket_slots) return BP_ERR_OOM;
    ge_p3 *tpart = w->buckets + bucket_need;
    RowMap rm{0, w->rg_ver, (long)g->capacity, 0, 0, 0};
    CK(launch(B * S, s, KMsmTable{g->table, rm, w->dig, rows * 32, trows, (int)S, tpart}));
    MsmSeg seg[1] = {{w->pts, npts, 1, (int)npts}};
    KMsmAccumulate k{};
    k.seg[0] = seg[0]; k.nseg = 1; k.S = 1; k.dig = w->dig + trows * 32; k.dig_inst_stride = rows * 32; k.buckets = w->buckets; k.wsum = w->wsum;
    CK(launch((long)B * MSM_WINDOWS, s, k));
    CK(launch(B, s, KVerifyCheck{w->wsum, tpart, (int)S, A.status, 0}));
    return BP_OK;
  }
  MsmSeg segs[4] = {{g->pc_niels, 0, 0, 2}, {g->G_n, 0, 0, (int)N}, {g->H_n, 0, 0, (int)N}, {w->pts, npts, 1, (int)npts}};
  return run_msm(w, segs, 4, B, w->dig, rows * 32, nullptr, 0, 1, A.status, s);
}
