// Sorted-bucket MSM group (see kernel_groups.h) + the one block-cooperative kernel of the library: the per-instance counting
// sort of (row, window) digit items by bucket, with the 4096 counters and cursors in shared memory.
#define KGROUP_DEFINING
#include "kernel_groups.h"
KGROUP_SORTED(KDEFINE)

#ifndef BP_HOST_EMUL
// One block per instance.  Pass 0 counts the items of every bucket (shared-memory atomics) and scans the counts into the
// bucket offsets.  Pass 1 walks the rows in tiles of SORT_THREADS rows: the tile's items are ranked per bucket in shared
// memory, staged there in bucket order and copied out one bucket run per warp step, so that global memory sees short
// CONTIGUOUS runs appended to each bucket's region instead of one random 4-byte store per item (that scatter cost a
// 32-byte read-modify-write in DRAM per item once the item lists of all resident blocks outgrew the L2: 16x amplification).
#define SORT_THREADS 1024
#define SORT_TILE_ITEMS (SORT_THREADS * SB_WINDOWS)
#define SORT_SMEM_BYTES ((2 * SB_BUCKETS + 32 + SORT_TILE_ITEMS) * 4)
// in-place exclusive scan of the SB_BUCKETS counters at c[]; returns the total to every thread
__device__ __forceinline__ uint32_t sort_block_scan(uint32_t *c, uint32_t *warp_tot) {
  constexpr int PER = SB_BUCKETS / SORT_THREADS;
  const int tid = threadIdx.x;
  uint32_t loc[PER], sum = 0;
#pragma unroll
  for (int i = 0; i < PER; i++) { loc[i] = sum; sum += c[tid * PER + i]; }
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += v; }
  if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
  __syncthreads();
  if (tid < 32) {
    uint32_t v = warp_tot[tid], inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (tid >= o) inc += t; }
    warp_tot[tid] = inc - v;      // exclusive prefix of the warp totals
    if (tid == 31) warp_tot[32] = inc;
  }
  __syncthreads();
  const uint32_t base = warp_tot[tid >> 5] + (incl - sum);
#pragma unroll
  for (int i = 0; i < PER; i++) c[tid * PER + i] = base + loc[i];
  const uint32_t total = warp_tot[32];
  __syncthreads();
  return total;
}
__global__ void __launch_bounds__(SORT_THREADS, 1) sort_buckets_kernel(RowMap rmap, const int8_t *dig, long dig_inst_stride, long rows,
                                                                      uint32_t *items, long items_stride, uint32_t *boff) {
  extern __shared__ uint32_t sort_sm[];
  uint32_t *cur = sort_sm;                         // [SB_BUCKETS] next free slot of every bucket's global region
  uint32_t *tcnt = sort_sm + SB_BUCKETS;           // [SB_BUCKETS + 1] counts, then offsets, of the tile in flight
  uint32_t *stage = sort_sm + 2 * SB_BUCKETS + 32; // [SORT_TILE_ITEMS] the tile's items in bucket order
  __shared__ uint32_t warp_tot[33];
  const long inst = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int8_t *drow = dig + inst * dig_inst_stride;
  uint32_t *it = items + inst * items_stride;
  // pass 0: histogram of the whole instance -> bucket offsets
  for (int b = tid; b < SB_BUCKETS; b += SORT_THREADS) tcnt[b] = 0;
  __syncthreads();
  for (long r = tid; r < rows; r += SORT_THREADS) {
    int16_t d[24]; load_digits13(d, drow + r * SB_ROW_BYTES);
#pragma unroll
    for (int w = 0; w < SB_WINDOWS; w++) if (d[w]) atomicAdd(&tcnt[(d[w] < 0 ? -d[w] : d[w]) - 1], 1u);
  }
  __syncthreads();
  {
    const uint32_t total = sort_block_scan(tcnt, warp_tot);
    uint32_t *off = boff + inst * (SB_BUCKETS + 1);
    for (int b = tid; b < SB_BUCKETS; b += SORT_THREADS) { const uint32_t o = tcnt[b]; off[b] = o; cur[b] = o; }
    if (tid == 0) off[SB_BUCKETS] = total;
  }
  __syncthreads();
  // pass 1: tiles of SORT_THREADS rows, one row per thread
  for (long t0 = 0; t0 < rows; t0 += SORT_THREADS) {
    for (int b = tid; b < SB_BUCKETS; b += SORT_THREADS) tcnt[b] = 0;
    __syncthreads();
    const long r = t0 + tid;
    uint32_t code[SB_WINDOWS];  // sign << 31 | bucket << 16 | rank within (tile, bucket); ~0 = no item
    if (r < rows) {
      int16_t d[24]; load_digits13(d, drow + r * SB_ROW_BYTES);
#pragma unroll
      for (int w = 0; w < SB_WINDOWS; w++) {
        code[w] = 0xffffffffu;
        if (d[w]) {
          const uint32_t neg = d[w] < 0; const uint32_t b = (uint32_t)(neg ? -d[w] : d[w]) - 1;
          code[w] = (neg << 31) | (b << 16) | atomicAdd(&tcnt[b], 1u);
        }
      }
    } else {
#pragma unroll
      for (int w = 0; w < SB_WINDOWS; w++) code[w] = 0xffffffffu;
    }
    __syncthreads();
    const uint32_t ttotal = sort_block_scan(tcnt, warp_tot);
    if (tid == 0) tcnt[SB_BUCKETS] = ttotal;
    if (r < rows) {
      const uint32_t g = (uint32_t)row_gen_item(rmap, r, inst) * SB_WINDOWS;
#pragma unroll
      for (int w = 0; w < SB_WINDOWS; w++) if (code[w] != 0xffffffffu) {
        const uint32_t b = (code[w] >> 16) & 0x7fffu;
        stage[tcnt[b] + (code[w] & 0xffffu)] = (g + w) | (code[w] & 0x80000000u);
      }
    }
    __syncthreads();
    // copy-out: every thread appends the runs of its SB_BUCKETS / SORT_THREADS buckets (a few items each) to their global regions
    for (int b = tid; b < SB_BUCKETS; b += SORT_THREADS) {
      const uint32_t s0 = tcnt[b], c = tcnt[b + 1] - s0, dst = cur[b];
      for (uint32_t j = 0; j < c; j++) it[dst + j] = stage[s0 + j];
      cur[b] = dst + c;
    }
    __syncthreads();
  }
}
// Variant without staging: after the histogram pass every item is stored straight at its bucket's cursor (a shared-memory
// atomic).  With 15-bit windows a tile of 1024 rows holds about one item per bucket, so staging buys no contiguity and its
// per-tile zero / scan / copy-out over 16384 buckets is pure overhead; one block per SM keeps the open 32-byte sectors of
// all resident blocks (16384 per block) inside the L2, where the partial writes are merged.
__global__ void __launch_bounds__(SORT_THREADS, 1) sort_buckets_direct_kernel(RowMap rmap, const int8_t *dig, long dig_inst_stride, long rows,
                                                                             uint32_t *items, long items_stride, uint32_t *boff) {
  extern __shared__ uint32_t sort_sm[];
  uint32_t *cur = sort_sm;  // [SB_BUCKETS] counts, then offsets = cursors
  __shared__ uint32_t warp_tot[33];
  const long inst = blockIdx.x;
  const int tid = threadIdx.x;
  const int8_t *drow = dig + inst * dig_inst_stride;
  uint32_t *it = items + inst * items_stride;
  for (int b = tid; b < SB_BUCKETS; b += SORT_THREADS) cur[b] = 0;
  __syncthreads();
  for (long r = tid; r < rows; r += SORT_THREADS) {
    int16_t d[24]; load_digits13(d, drow + r * SB_ROW_BYTES);
#pragma unroll
    for (int w = 0; w < SB_WINDOWS; w++) if (d[w]) atomicAdd(&cur[(d[w] < 0 ? -d[w] : d[w]) - 1], 1u);
  }
  __syncthreads();
  {
    const uint32_t total = sort_block_scan(cur, warp_tot);
    uint32_t *off = boff + inst * (SB_BUCKETS + 1);
    for (int b = tid; b < SB_BUCKETS; b += SORT_THREADS) off[b] = cur[b];
    if (tid == 0) off[SB_BUCKETS] = total;
  }
  __syncthreads();
  for (long r = tid; r < rows; r += SORT_THREADS) {
    int16_t d[24]; load_digits13(d, drow + r * SB_ROW_BYTES);
    const uint32_t g = (uint32_t)row_gen_item(rmap, r, inst) * SB_WINDOWS;
#pragma unroll
    for (int w = 0; w < SB_WINDOWS; w++) if (d[w]) {
      const uint32_t neg = d[w] < 0; const uint32_t b = (uint32_t)(neg ? -d[w] : d[w]) - 1;
      it[atomicAdd(&cur[b], 1u)] = (g + w) | (neg << 31);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------------
// Two-pass (most-significant-digit first) variant, the default.  The one-pass kernels above append 4-byte items to 16384
// bucket regions per instance; with ~1 item per bucket and tile that is one partial 32-byte sector per item, and the ncu
// capture (profiles/) shows it: 37.7 GB of DRAM traffic for 8.4 GB of digits + items, at the ~2 TB/s that scattered
// sectors reach.  Here pass 1 partitions the items of an instance into SORT_GROUPS coarse groups of SORT_FINE buckets (runs
// of ~70 items per group and tile leave shared memory as whole lines) and pass 2 sorts one coarse group per block inside
// shared memory and writes it back linearly.  Every global access is coalesced; items cross HBM three times instead of
// being read-modified-written sector by sector.  The low bucket bits ride in bits 24..30 of the intermediate item, so
// generator slot * SB_WINDOWS must stay below 2^24 (the launcher falls back to the staged kernel otherwise).
#define SORT2_THREADS 512
#define SORT2_WARPS (SORT2_THREADS / 32)
#define SORT_FINE_BITS 7
#define SORT_FINE (1 << SORT_FINE_BITS)
#define SORT_GROUPS (SB_BUCKETS / SORT_FINE)
#define SORT2_TILE_ITEMS (SORT2_THREADS * SB_WINDOWS)
#define SORT2_BUF_WORDS (SB_BUCKETS > SORT2_TILE_ITEMS ? SB_BUCKETS : SORT2_TILE_ITEMS)
#define SORT2_WC_STRIDE (SORT2_WARPS + 1)  // odd stride: the 32 lanes of a warp (same warp slot, different groups) hit different banks
#define SORT2_WC_PER ((SORT_GROUPS * SORT2_WC_STRIDE + SORT2_THREADS - 1) / SORT2_THREADS)
#define SORT2_WC_WORDS (SORT2_WC_PER * SORT2_THREADS)
#define SORT2_SMEM_BYTES ((SORT2_BUF_WORDS + SORT2_WC_WORDS + SORT_GROUPS + 8) * 4)
#define SORT_ITEM_BITS 24   // bits 24..30 carry the fine bucket bits, bit 31 the sign: up to 986k generator slots (capacity ~493k)
#define SORT_ITEM_MASK (0x80000000u | ((1u << SORT_ITEM_BITS) - 1))
#define SORT_FINE_CAP 6016   // items of one coarse group sorted inside shared memory (larger groups scatter directly)

// in-place exclusive scan of THREADS * PER counters, thread-contiguous; returns the total to every thread
template <int PER, int THREADS>
__device__ __forceinline__ uint32_t block_scan_excl(uint32_t *c, uint32_t *warp_tot) {
  const int tid = threadIdx.x, lane = tid & 31;
  uint32_t sum = 0;
#pragma unroll 4
  for (int i = 0; i < PER; i++) sum += c[tid * PER + i];
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) warp_tot[tid >> 5] = incl;
  __syncthreads();
  if (tid < 32) {
    uint32_t v = tid < THREADS / 32 ? warp_tot[tid] : 0, inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (tid >= o) inc += t; }
    if (tid < THREADS / 32) warp_tot[tid] = inc - v;
    if (tid == 31) warp_tot[32] = inc;
  }
  __syncthreads();
  uint32_t run = warp_tot[tid >> 5] + (incl - sum);
#pragma unroll 4
  for (int i = 0; i < PER; i++) { const uint32_t v = c[tid * PER + i]; c[tid * PER + i] = run; run += v; }
  const uint32_t total = warp_tot[32];
  __syncthreads();
  return total;
}

__global__ void __launch_bounds__(SORT2_THREADS, 2) sort_coarse_kernel(RowMap rmap, const int8_t *dig, long dig_inst_stride, long rows,
                                                                      uint32_t *tmp, long items_stride, uint32_t *boff) {
  extern __shared__ uint32_t sort_sm[];
  uint32_t *buf = sort_sm;                                  // pass 0: SB_BUCKETS fine counters; afterwards the tile's staged items
  uint32_t *wc = sort_sm + SORT2_BUF_WORDS;                 // [SORT_GROUPS][SORT2_WC_STRIDE] items of (coarse group, warp) in the tile
  uint32_t *ccur = wc + SORT2_WC_WORDS;                     // [SORT_GROUPS] next free slot of every coarse group's region
  __shared__ uint32_t warp_tot[33];
  const long inst = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int8_t *drow = dig + inst * dig_inst_stride;
  uint32_t *out = tmp + inst * items_stride;
  // pass 0: fine histogram of the whole instance -> bucket offsets (global) and coarse cursors
  // Both passes stream the instance's digit rows through shared memory in tiles of SORT2_THREADS rows (x 48 contiguous bytes):
  // ONE asynchronous bulk copy per tile (cp.async.bulk, the TMA engine's 1-D form) issued by thread 0 and tracked by an
  // mbarrier; the copy of the next tile is issued as soon as every thread has taken its row into registers, so it is in flight
  // while the block works on the current one.
  __shared__ alignas(128) int8_t trows[SORT2_THREADS * SB_ROW_BYTES];
  __shared__ alignas(8) uint64_t tbar_mem;
  const uint32_t tbar = smem_addr32(&tbar_mem), tdst = smem_addr32(trows);
  uint32_t tphase = 0;
  auto fetch_tile = [&](long t0) {  // thread 0 only
    const long left = rows - t0, cnt = left < SORT2_THREADS ? left : SORT2_THREADS;
    mbar_arrive_expect_tx(tbar, (uint32_t)(cnt * SB_ROW_BYTES));
    bulk_copy_g2s(tdst, drow + t0 * SB_ROW_BYTES, (uint32_t)(cnt * SB_ROW_BYTES), tbar);
  };
  if (tid == 0) { mbar_init(tbar, 1); mbar_fence_init(); if (rows > 0) fetch_tile(0); }
  for (int b = tid; b < SB_BUCKETS; b += SORT2_THREADS) buf[b] = 0;
  __syncthreads();
  for (long t0 = 0; t0 < rows; t0 += SORT2_THREADS) {
    mbar_wait(tbar, tphase); tphase ^= 1;
    int16_t d[24];
    const bool have = t0 + tid < rows;
    if (have) load_digits13(d, trows + (long)tid * SB_ROW_BYTES);
    __syncthreads();  // rows are in registers: the buffer is free
    if (tid == 0) {   // next tile of this pass, or the first tile of pass 1
      fence_proxy_async_smem();
      fetch_tile(t0 + SORT2_THREADS < rows ? t0 + SORT2_THREADS : 0);
    }
    if (have) {
#pragma unroll
      for (int w = 0; w < SB_WINDOWS; w++) if (d[w]) atomicAdd(&buf[(d[w] < 0 ? -d[w] : d[w]) - 1], 1u);
    }
  }
  __syncthreads();
  {
    const uint32_t total = block_scan_excl<SB_BUCKETS / SORT2_THREADS, SORT2_THREADS>(buf, warp_tot);
    uint32_t *off = boff + inst * (SB_BUCKETS + 1);
    for (int b = tid; b < SB_BUCKETS; b += SORT2_THREADS) off[b] = buf[b];
    if (tid == 0) off[SB_BUCKETS] = total;
    if (tid < SORT_GROUPS) ccur[tid] = buf[tid * SORT_FINE];
  }
  __syncthreads();
  // pass 1: tiles of SORT2_THREADS rows, one row per thread (the first tile's copy was issued at the end of pass 0)
  for (long t0 = 0; t0 < rows; t0 += SORT2_THREADS) {
    for (int i = tid; i < SORT2_WC_WORDS; i += SORT2_THREADS) wc[i] = 0;
    __syncthreads();
    const long r = t0 + tid;
    uint32_t code[SB_WINDOWS];  // sign << 31 | bucket << 16 | rank within (tile, coarse group, warp); ~0 = no item
#pragma unroll
    for (int w = 0; w < SB_WINDOWS; w++) code[w] = 0xffffffffu;
    mbar_wait(tbar, tphase); tphase ^= 1;
    if (r < rows) {
      int16_t d[24]; load_digits13(d, trows + (long)tid * SB_ROW_BYTES);
#pragma unroll
      for (int w = 0; w < SB_WINDOWS; w++) if (d[w]) {
        const uint32_t neg = d[w] < 0; const uint32_t b = (uint32_t)(neg ? -d[w] : d[w]) - 1;
        code[w] = (neg << 31) | (b << 16) | atomicAdd(&wc[(b >> SORT_FINE_BITS) * SORT2_WC_STRIDE + wid], 1u);
      }
    }
    __syncthreads();  // every row of the tile is in registers: the buffer is free for the next tile
    if (tid == 0 && t0 + SORT2_THREADS < rows) { fence_proxy_async_smem(); fetch_tile(t0 + SORT2_THREADS); }
    const uint32_t ttotal = block_scan_excl<SORT2_WC_PER, SORT2_THREADS>(wc, warp_tot);
    if (r < rows) {
      const uint32_t g = (uint32_t)row_gen_item(rmap, r, inst) * SB_WINDOWS;
#pragma unroll
      for (int w = 0; w < SB_WINDOWS; w++) if (code[w] != 0xffffffffu) {
        const uint32_t b = (code[w] >> 16) & 0x7fffu;
        buf[wc[(b >> SORT_FINE_BITS) * SORT2_WC_STRIDE + wid] + (code[w] & 0xffffu)] =
            (g + w) | ((b & (SORT_FINE - 1)) << SORT_ITEM_BITS) | (code[w] & 0x80000000u);
      }
    }
    __syncthreads();
    // copy-out: one warp per coarse group, the group's run of the tile leaves as consecutive 128-byte lines
    for (int c = wid; c < SORT_GROUPS; c += SORT2_WARPS) {
      const uint32_t s0 = wc[c * SORT2_WC_STRIDE], s1 = c + 1 < SORT_GROUPS ? wc[(c + 1) * SORT2_WC_STRIDE] : ttotal, dst = ccur[c];
      for (uint32_t j = lane; j < s1 - s0; j += 32) out[dst + j] = buf[s0 + j];
      __syncwarp();
      if (lane == 0) ccur[c] = dst + (s1 - s0);
    }
    __syncthreads();
  }
}
// pass 2: block (instance, coarse group) orders the group's items by their low bucket bits
__global__ void __launch_bounds__(256, 8) sort_fine_kernel(const uint32_t *tmp, uint32_t *items, long items_stride, const uint32_t *boff) {
  __shared__ uint32_t cur[SORT_FINE];
  __shared__ uint32_t stage[SORT_FINE_CAP];
  const long inst = blockIdx.x / SORT_GROUPS; const int c = (int)(blockIdx.x % SORT_GROUPS);
  const int tid = threadIdx.x;
  const uint32_t *off = boff + inst * (SB_BUCKETS + 1) + c * SORT_FINE;
  const uint32_t start = off[0], n = off[SORT_FINE] - start;
  if (n == 0) return;
  if (tid < SORT_FINE) cur[tid] = off[tid] - start;
  __syncthreads();
  const uint32_t *in = tmp + inst * items_stride + start;
  uint32_t *out = items + inst * items_stride + start;
  if (n <= SORT_FINE_CAP) {
    for (uint32_t k0 = 0; k0 < n; k0 += 4 * 256) {  // four independent loads in flight per thread
      uint32_t x[4];
#pragma unroll
      for (int j = 0; j < 4; j++) { const uint32_t k = k0 + j * 256 + tid; x[j] = k < n ? in[k] : 0xffffffffu; }
#pragma unroll
      for (int j = 0; j < 4; j++) if (k0 + j * 256 + tid < n) stage[atomicAdd(&cur[(x[j] >> SORT_ITEM_BITS) & (SORT_FINE - 1)], 1u)] = x[j] & SORT_ITEM_MASK;
    }
    __syncthreads();
    for (uint32_t k = tid; k < n; k += 256) out[k] = stage[k];
  } else {
    for (uint32_t k = tid; k < n; k += 256) {
      const uint32_t x = in[k];
      out[atomicAdd(&cur[(x >> SORT_ITEM_BITS) & (SORT_FINE - 1)], 1u)] = x & SORT_ITEM_MASK;
    }
  }
}
#endif

int launch_sort_buckets(const RowMap &rmap, const int8_t *dig, long dig_inst_stride, long rows, long ninst, uint32_t *items, long items_stride,
                        uint32_t *boff, uint32_t *soff, uint32_t *tmp, size_t tmp_bytes, long gen_slots, dev_stream s) {
#ifndef BP_HOST_EMUL
  if (ninst <= 0) return 0;
  static int mode = -1;  // 2: two-pass radix (default), 0: staged tiles (BP_B200_SORT=staged), 1: direct scatter (BP_B200_SORT=direct)
  if (mode < 0) {
    const char *e = getenv("BP_B200_SORT");
    mode = e ? (e[0] == 'd' ? 1 : (e[0] == 's' ? 0 : 2)) : 2;
    cudaFuncSetAttribute(sort_buckets_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_SMEM_BYTES);
    cudaFuncSetAttribute(sort_buckets_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_SMEM_BYTES);
    cudaFuncSetAttribute(sort_coarse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT2_SMEM_BYTES);
  }
  // the two-pass form needs scratch for the coarse-partitioned items and 22 bits for generator slot * SB_WINDOWS + window
  const bool radix = mode == 2 && tmp && tmp_bytes >= (size_t)ninst * items_stride * sizeof(uint32_t) && gen_slots > 0 &&
                     (size_t)gen_slots * SB_WINDOWS < (1u << SORT_ITEM_BITS) && ninst * SORT_GROUPS < (1L << 31);
  if (radix) {
    if (g_profile_on == 1) profile_begin("sort_coarse_kernel", ninst * SORT2_THREADS, s);
    sort_coarse_kernel<<<(unsigned)ninst, SORT2_THREADS, SORT2_SMEM_BYTES, s>>>(rmap, dig, dig_inst_stride, rows, tmp, items_stride, boff);
    if (g_profile_on == 1) profile_end(s);
    if (g_profile_on == 1) profile_begin("sort_fine_kernel", ninst * SORT_GROUPS * 256, s);
    sort_fine_kernel<<<(unsigned)(ninst * SORT_GROUPS), 256, 0, s>>>(tmp, items, items_stride, boff);
    if (g_profile_on == 1) profile_end(s);
    g_launch_count += 2;
  } else {
    if (g_profile_on == 1) profile_begin("sort_buckets_kernel", ninst * SORT_THREADS, s);
    // the direct variant asks for the same (large) shared-memory carve-out on purpose: one block per SM bounds the open sectors
    if (mode == 1) sort_buckets_direct_kernel<<<(unsigned)ninst, SORT_THREADS, SORT_SMEM_BYTES, s>>>(rmap, dig, dig_inst_stride, rows, items, items_stride, boff);
    else sort_buckets_kernel<<<(unsigned)ninst, SORT_THREADS, SORT_SMEM_BYTES, s>>>(rmap, dig, dig_inst_stride, rows, items, items_stride, boff);
    if (g_profile_on == 1) profile_end(s);
    g_launch_count++;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { fprintf(stderr, "bp_b200: sort_buckets launch failed: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
#else
  (void)tmp; (void)tmp_bytes; (void)gen_slots;
  return launch(ninst, s, KSortBucketsSerial{rmap, dig, dig_inst_stride, rows, items, items_stride, boff, soff});
#endif
}
