// Host-side plumbing shared by every translation unit of libalm_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/alm_b200.h"

namespace alm {

// Every kernel launch of this library goes through ALM_LAUNCHED() so the host can report
// "gpu_launches" per step (bench.py) without a profiler.
extern unsigned long long g_launch_count;
#define ALM_LAUNCHED(n) (::alm::g_launch_count += (n))

#define ALM_CUDA_OK(expr)                                                                         \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      fprintf(stderr, "[alm] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(_e), __FILE__, __LINE__, \
              cudaGetErrorString(_e));                                                            \
      return ALM_ERR_CUDA;                                                                        \
    }                                                                                             \
  } while (0)

#define ALM_REQUIRE(cond, code)                                                            \
  do {                                                                                     \
    if (!(cond)) {                                                                         \
      fprintf(stderr, "[alm] argument check failed (%s) at %s:%d\n", #cond, __FILE__, __LINE__); \
      return (code);                                                                       \
    }                                                                                      \
  } while (0)

#define ALM_CHECK_LAUNCH() ALM_CUDA_OK(cudaGetLastError())

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

// Build a (up to) 3-D bf16/fp32 tiled tensor map with 128-B swizzle. dims/strides innermost first;
// strides in BYTES for dims 1.. (dim 0 is contiguous). Returns ALM_OK or an error code.
int make_tensor_map(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128);

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) {
  return (a + b - 1) / b;
}

// ---- device helpers ------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

}  // namespace alm
