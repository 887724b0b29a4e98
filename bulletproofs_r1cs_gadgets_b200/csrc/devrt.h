// Launch / memory shim.  Product build: CUDA runtime on the caller's stream, every kernel is a
// named instantiation run_kernel<Functor> (that is the name ncu shows).  Emulation build
// (-DBP_HOST_EMUL, tests/emul only): the same functor bodies run in a host loop so kernel-body
// unit tests work in the CPU-only container.  There is no runtime switch between the two.
#pragma once
#include "hd.h"
#include <stdio.h>
#include <stdlib.h>

#ifndef BP_HOST_EMUL
#include <cuda_runtime.h>
typedef cudaStream_t dev_stream;

template <class K>
__global__ void __launch_bounds__(K::kBlock, K::kMinBlocks) run_kernel(long n, K k) {
  long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid < n) k(tid);
}
extern long g_launch_count;
// optional per-kernel timing (bp_profile_*): CUDA events recorded on the launching stream around every kernel
extern int g_profile_on;
void profile_begin(const char *name, long threads, dev_stream s);
bool profile_selected(const char *name);
void profile_end(dev_stream s);
template <class K>
int launch(long n, dev_stream s, const K &k) {
  if (n <= 0) return 0;
  long blocks = (n + K::kBlock - 1) / K::kBlock;
  // g_profile_on: 1 = every launch, 2 = only the launches of profile_selected() kernels (the events around ALL ~700 launches of a
  // step cost ~2.5 % of the step; the timed region of bench.py brackets the dominant kernel only)
  const bool prof = g_profile_on == 1 || (g_profile_on == 2 && profile_selected(K::kName));
  if (prof) profile_begin(K::kName, n, s);
  run_kernel<K><<<(unsigned)blocks, K::kBlock, 0, s>>>(n, k);
  if (prof) profile_end(s);
  g_launch_count++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { fprintf(stderr, "bp_b200: launch failed: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
inline int dev_malloc(void **p, size_t n) {
  if (cudaMalloc(p, n ? n : 16) == cudaSuccess) return 0;
  cudaGetLastError();  // an allocation failure is reported to the caller (BP_ERR_OOM or a fallback); it must not surface again at the next launch check
  *p = nullptr;
  return 1;
}
inline void dev_free(void *p) { if (p) cudaFree(p); }
inline int dev_h2d(void *d, const void *h, size_t n, dev_stream s) { return n ? cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s) != cudaSuccess : 0; }
inline int dev_d2h(void *h, const void *d, size_t n, dev_stream s) { return n ? cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s) != cudaSuccess : 0; }
inline int dev_d2d(void *d, const void *s_, size_t n, dev_stream s) { return n ? cudaMemcpyAsync(d, s_, n, cudaMemcpyDeviceToDevice, s) != cudaSuccess : 0; }
inline int dev_memset(void *d, int v, size_t n, dev_stream s) { return n ? cudaMemsetAsync(d, v, n, s) != cudaSuccess : 0; }
// fork/join helpers for the second internal stream (witness generation runs beside the transcript RNG)
struct dev_side { cudaStream_t s = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
inline int dev_side_init(dev_side &d) {
  if (d.s) return 0;
  // highest priority: the side streams carry the latency chains of a batch's first phase (a few hundred small blocks that run
  // for most of a second); the block scheduler must place them as soon as an SM has room, ahead of the queued blocks of the
  // throughput kernels on the caller's stream -- otherwise they wait for a grid's tail and the next batch stalls on them
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  if (cudaStreamCreateWithPriority(&d.s, cudaStreamNonBlocking, hi) != cudaSuccess) return 1;
  if (cudaEventCreateWithFlags(&d.fork, cudaEventDisableTiming) != cudaSuccess) return 1;
  if (cudaEventCreateWithFlags(&d.join, cudaEventDisableTiming) != cudaSuccess) return 1;
  return 0;
}
inline void dev_side_free(dev_side &d) { if (d.s) { cudaStreamDestroy(d.s); cudaEventDestroy(d.fork); cudaEventDestroy(d.join); d = dev_side(); } }
inline dev_stream dev_side_fork(dev_side &d, dev_stream main) { cudaEventRecord(d.fork, main); cudaStreamWaitEvent(d.s, d.fork, 0); return d.s; }
inline void dev_side_join(dev_side &d, dev_stream main) { cudaEventRecord(d.join, d.s); cudaStreamWaitEvent(main, d.join, 0); }
inline int dev_sync(dev_stream s) {
  cudaError_t e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) { fprintf(stderr, "bp_b200: stream sync failed: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
inline int dev_sync_device() { return cudaDeviceSynchronize() != cudaSuccess; }
#else
typedef int dev_stream;
extern long g_launch_count;
extern int g_profile_on;
void profile_note(const char *name, long threads);
template <class K>
int launch(long n, dev_stream, const K &k) {
  if (g_profile_on) profile_note(K::kName, n);
  for (long t = 0; t < n; t++) k(t);
  g_launch_count++;
  return 0;
}
inline int dev_malloc(void **p, size_t n) { *p = malloc(n ? n : 16); return *p == NULL; }
inline void dev_free(void *p) { free(p); }
inline int dev_h2d(void *d, const void *h, size_t n, dev_stream) { if (n) memcpy(d, h, n); return 0; }
inline int dev_d2h(void *h, const void *d, size_t n, dev_stream) { if (n) memcpy(h, d, n); return 0; }
inline int dev_d2d(void *d, const void *s_, size_t n, dev_stream) { if (n) memcpy(d, s_, n); return 0; }
inline int dev_memset(void *d, int v, size_t n, dev_stream) { if (n) memset(d, v, n); return 0; }
struct dev_side { int s = 0; };
inline int dev_side_init(dev_side &) { return 0; }
inline void dev_side_free(dev_side &) {}
inline dev_stream dev_side_fork(dev_side &, dev_stream main) { return main; }
inline void dev_side_join(dev_side &, dev_stream) {}
inline int dev_sync(dev_stream) { return 0; }
inline int dev_sync_device() { return 0; }
#endif
