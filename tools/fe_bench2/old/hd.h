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
