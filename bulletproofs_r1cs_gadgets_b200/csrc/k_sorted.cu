// Sorted-bucket MSM group (see kernel_groups.h) + the one block-cooperative kernel of the library: the per-instance counting
// sort of (row, window) digit items by bucket, with the 4096 counters and cursors in shared memory.
#define KGROUP_DEFINING
#include "kernel_groups.h"
KGROUP_SORTED(KDEFINE)

#ifndef BP_HOST_EMUL
#define SORT_THREADS 512
__global__ void __launch_bounds__(SORT_THREADS) sort_buckets_kernel(RowMap rmap, const int8_t *dig, long dig_inst_stride, long rows,
                                                                   uint32_t *items, long items_stride, uint32_t *boff, uint32_t *soff) {
  __shared__ uint32_t cnt[SB_BUCKETS];
  __shared__ uint32_t warp_tot[SORT_THREADS / 32];
  const long inst = blockIdx.x;
  const int tid = threadIdx.x;
  for (int b = tid; b < SB_BUCKETS; b += SORT_THREADS) cnt[b] = 0;
  __syncthreads();
  const int8_t *drow = dig + inst * dig_inst_stride;
  // pass 1: histogram
  for (long r = tid; r < rows; r += SORT_THREADS) {
    int16_t d[24]; load_digits13(d, drow + r * SB_ROW_BYTES);
#pragma unroll
    for (int w = 0; w < SB_WINDOWS; w++) if (d[w]) atomicAdd(&cnt[(d[w] < 0 ? -d[w] : d[w]) - 1], 1u);
  }
  __syncthreads();
  // exclusive scan of the 4096 counters: 8 consecutive counters per thread, warp scan, block scan
  constexpr int PER = SB_BUCKETS / SORT_THREADS;
  uint32_t loc[PER], sum = 0, sloc[PER], ssum = 0;
#pragma unroll
  for (int i = 0; i < PER; i++) { uint32_t c = cnt[tid * PER + i]; loc[i] = sum; sum += c; sloc[i] = ssum; ssum += (c + SB_SLICE - 1) / SB_SLICE; }
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += v; }
  if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
  __syncthreads();
  if (tid < 32) {
    uint32_t v = tid < SORT_THREADS / 32 ? warp_tot[tid] : 0, inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (tid >= o) inc += t; }
    if (tid < SORT_THREADS / 32) warp_tot[tid] = inc - v;  // exclusive prefix of the warp totals
  }
  __syncthreads();
  const uint32_t base = warp_tot[tid >> 5] + (incl - sum);
  uint32_t *off = boff + inst * (SB_BUCKETS + 1);
#pragma unroll
  for (int i = 0; i < PER; i++) { uint32_t o = base + loc[i]; off[tid * PER + i] = o; }
  if (tid == SORT_THREADS - 1) off[SB_BUCKETS] = base + sum;
  __syncthreads();
  {  // same block scan for the slice counts
    uint32_t sincl = ssum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, sincl, o); if ((tid & 31) >= o) sincl += v; }
    if ((tid & 31) == 31) warp_tot[tid >> 5] = sincl;
    __syncthreads();
    if (tid < 32) {
      uint32_t v = tid < SORT_THREADS / 32 ? warp_tot[tid] : 0, inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (tid >= o) inc += t; }
      if (tid < SORT_THREADS / 32) warp_tot[tid] = inc - v;
    }
    __syncthreads();
    const uint32_t sbase = warp_tot[tid >> 5] + (sincl - ssum);
    uint32_t *so = soff + inst * (SB_BUCKETS + 1);
#pragma unroll
    for (int i = 0; i < PER; i++) so[tid * PER + i] = sbase + sloc[i];
    if (tid == SORT_THREADS - 1) so[SB_BUCKETS] = sbase + ssum;
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < PER; i++) cnt[tid * PER + i] = base + loc[i];  // cursors
  __syncthreads();
  // pass 2: scatter
  uint32_t *it = items + inst * items_stride;
  for (long r = tid; r < rows; r += SORT_THREADS) {
    int16_t d[24]; load_digits13(d, drow + r * SB_ROW_BYTES);
    const uint32_t g = (uint32_t)row_gen(rmap, r) * SB_WINDOWS;
#pragma unroll
    for (int w = 0; w < SB_WINDOWS; w++) if (d[w]) {
      const int neg = d[w] < 0; const int b = (neg ? -d[w] : d[w]) - 1;
      const uint32_t pos = atomicAdd(&cnt[b], 1u);
      it[pos] = (g + w) | ((uint32_t)neg << 31);
    }
  }
}
#endif

int launch_sort_buckets(const RowMap &rmap, const int8_t *dig, long dig_inst_stride, long rows, long ninst, uint32_t *items, long items_stride,
                        uint32_t *boff, uint32_t *soff, dev_stream s) {
#ifndef BP_HOST_EMUL
  if (ninst <= 0) return 0;
  if (g_profile_on) profile_begin("sort_buckets_kernel", ninst * SORT_THREADS, s);
  sort_buckets_kernel<<<(unsigned)ninst, SORT_THREADS, 0, s>>>(rmap, dig, dig_inst_stride, rows, items, items_stride, boff, soff);
  if (g_profile_on) profile_end(s);
  g_launch_count++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { fprintf(stderr, "bp_b200: sort_buckets launch failed: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
#else
  return launch(ninst, s, KSortBucketsSerial{rmap, dig, dig_inst_stride, rows, items, items_stride, boff, soff});
#endif
}
