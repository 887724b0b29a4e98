// Host/device plumbing shared by every arithmetic header.
//
// The arithmetic (field, scalar, point, Keccak) is written once as __host__ __device__ inline
// functions.  The product library compiles them with nvcc for sm_100a and only ever runs them
// inside CUDA kernels; tests/emul builds the same bodies with g++ (-DBP_HOST_EMUL) so that
// kernel-body unit tests can run in the CPU-only container.  The emulation build is test
// infrastructure: nothing in the shipped library or the Python API can reach it.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#define HDNI __host__ __device__ __noinline__
#define DEVCONST __device__ __constant__
#else
#define HD inline
#define HDNI
#define DEVCONST
#endif

// 64x64 -> 128
HD void mul64wide(uint64_t a, uint64_t b, uint64_t &hi, uint64_t &lo) {
#if defined(__CUDA_ARCH__)
  lo = a * b;
  hi = __umul64hi(a, b);
#else
  unsigned __int128 p = (unsigned __int128)a * b;
  lo = (uint64_t)p;
  hi = (uint64_t)(p >> 64);
#endif
}

// 16-byte vector copies of plain structs (global <-> registers); sizes are multiples of 16 or 8
template <typename T>
HD void load_struct(T &dst, const T *src) {
#if defined(__CUDA_ARCH__)
  if constexpr (sizeof(T) % 16 == 0) {
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    uint4 *d = reinterpret_cast<uint4 *>(&dst);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = s[i];
  } else {
    const uint2 *s = reinterpret_cast<const uint2 *>(src);
    uint2 *d = reinterpret_cast<uint2 *>(&dst);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 8); i++) d[i] = s[i];
  }
#else
  dst = *src;
#endif
}
template <typename T>
HD void store_struct(T *dst, const T &src) {
#if defined(__CUDA_ARCH__)
  if constexpr (sizeof(T) % 16 == 0) {
    uint4 *d = reinterpret_cast<uint4 *>(dst);
    const uint4 *s = reinterpret_cast<const uint4 *>(&src);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = s[i];
  } else {
    uint2 *d = reinterpret_cast<uint2 *>(dst);
    const uint2 *s = reinterpret_cast<const uint2 *>(&src);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 8); i++) d[i] = s[i];
  }
#else
  *dst = src;
#endif
}
// The same in 256-bit granules -- LDG.E.256 / STG.E.256, a width that exists from sm_100 on (PTX ld.global.v8.u32) -- for structs
// of 32-byte alignment in GLOBAL memory.  Used where a kernel streams whole points through a read-modify-write (the per-proof
// bucket kernel: -9 %); measured neutral for the table gathers of the sorted path and worse for the chained loads of the bucket
// reduction, which keep the 128-bit form.
template <typename T>
HD void load_struct256(T &dst, const T *src) {
#if defined(__CUDA_ARCH__)
  static_assert(sizeof(T) % 32 == 0 && alignof(T) >= 32, "32-byte granules");
  uint32_t *d = reinterpret_cast<uint32_t *>(&dst);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 32); i++)
    asm volatile("ld.global.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(d[8 * i]), "=r"(d[8 * i + 1]), "=r"(d[8 * i + 2]), "=r"(d[8 * i + 3]), "=r"(d[8 * i + 4]), "=r"(d[8 * i + 5]), "=r"(d[8 * i + 6]), "=r"(d[8 * i + 7])
                 : "l"(reinterpret_cast<const char *>(src) + 32 * i) : "memory");
#else
  dst = *src;
#endif
}
template <typename T>
HD void store_struct256(T *dst, const T &src) {
#if defined(__CUDA_ARCH__)
  static_assert(sizeof(T) % 32 == 0 && alignof(T) >= 32, "32-byte granules");
  const uint32_t *s = reinterpret_cast<const uint32_t *>(&src);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 32); i++)
    asm volatile("st.global.v8.u32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(reinterpret_cast<char *>(dst) + 32 * i), "r"(s[8 * i]), "r"(s[8 * i + 1]),
                 "r"(s[8 * i + 2]), "r"(s[8 * i + 3]), "r"(s[8 * i + 4]), "r"(s[8 * i + 5]), "r"(s[8 * i + 6]), "r"(s[8 * i + 7]) : "memory");
#else
  *dst = src;
#endif
}
// ---- asynchronous bulk copies (the TMA engine's 1-D form, cp.async.bulk) with mbarrier completion: sm_90+ PTX ----
#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t smem_addr32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completion is signalled on `bar` (complete_tx)
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif
