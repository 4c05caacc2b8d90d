// Shared helpers for the sm_100a kernels behind include/reid_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/reid_b200.h"

namespace reid {

void set_error(const char* fmt, ...);

#define REID_CHECK_ARG(cond, ...)                 \
  do {                                            \
    if (!(cond)) {                                \
      ::reid::set_error(__VA_ARGS__);             \
      return REID_ERR_INVALID_ARG;                \
    }                                             \
  } while (0)

#define REID_CUDA(call)                                                              \
  do {                                                                               \
    cudaError_t e__ = (call);                                                        \
    if (e__ != cudaSuccess) {                                                        \
      ::reid::set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return REID_ERR_CUDA;                                                          \
    }                                                                                \
  } while (0)

// Developer switches (skip parts of a kernel to time the rest) exist only in builds with -DREID_DEV; a release
// library has no environment variable that changes what a kernel computes.
#ifdef REID_DEV
#define REID_DBG(p) ((p).dbg)
inline int dev_env(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
#else
#define REID_DBG(p) 0
inline int dev_env(const char*, int dflt) { return dflt; }
#endif

extern unsigned long long g_launches;  // kernels launched by this library in this process

#define REID_LAUNCH_CHECK()        \
  do {                             \
    ++::reid::g_launches;          \
    REID_CUDA(cudaGetLastError()); \
  } while (0)

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    T u = __shfl_xor_sync(kFull, v, o);
    v = u > v ? u : v;
  }
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_min(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    T u = __shfl_xor_sync(kFull, v, o);
    v = u < v ? u : v;
  }
  return v;
}

// Monotone map float -> uint32 (larger float => larger uint).
__device__ __forceinline__ uint32_t float_ord(float f) {
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord_float(uint32_t o) {
  uint32_t b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(b);
}
// Composite selection key: larger = better (higher similarity, then lower index).
__device__ __forceinline__ uint64_t sel_key(float key, int idx) {
  return ((uint64_t)float_ord(key) << 32) | (uint32_t)(0xffffffffu - (uint32_t)idx);
}
__device__ __forceinline__ int sel_key_idx(uint64_t k) { return (int)(0xffffffffu - (uint32_t)k); }
__device__ __forceinline__ float sel_key_val(uint64_t k) { return ord_float((uint32_t)(k >> 32)); }

inline int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace reid
