// Shared helpers for the dfmir_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define DFMIR_OK 0
#define DFMIR_ERR_ARG -1
#define DFMIR_ERR_CUDA -2
#define DFMIR_ERR_UNSUPPORTED -3

// thread-local error string, exported through dfmir_last_error()
void dfmir_set_error(const char* fmt, ...);
// every kernel launch of the library is counted (dfmir_launch_count, used by bench.py)
void dfmir_count_launch();

#define DFMIR_CHECK_ARG(cond, ...)              \
  do {                                          \
    if (!(cond)) {                              \
      dfmir_set_error(__VA_ARGS__);             \
      return DFMIR_ERR_ARG;                     \
    }                                           \
  } while (0)

#define DFMIR_CHECK_LAUNCH(name)                                           \
  do {                                                                     \
    cudaError_t e__ = cudaGetLastError();                                  \
    dfmir_count_launch();                                                  \
    if (e__ != cudaSuccess) {                                              \
      dfmir_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return DFMIR_ERR_CUDA;                                               \
    }                                                                      \
  } while (0)

#define DFMIR_CUDA(call)                                                   \
  do {                                                                     \
    cudaError_t e__ = (call);                                              \
    if (e__ != cudaSuccess) {                                              \
      dfmir_set_error("%s: %s", #call, cudaGetErrorString(e__));           \
      return DFMIR_ERR_CUDA;                                               \
    }                                                                      \
  } while (0)

static inline int dfmir_ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// Number of SMs on the current device (cached). Grids for grid-stride kernels are sized
// as a multiple of this (148 on B200).
int dfmir_num_sms();

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum (blockDim.x multiple of 32, <= 1024). Result valid in thread 0.
__device__ __forceinline__ float block_sum(float v, float* smem32) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem32[w] = v;
  __syncthreads();
  if (w == 0) {
    int nw = (blockDim.x + 31) >> 5;
    v = lane < nw ? smem32[lane] : 0.f;
    v = warp_sum(v);
  }
  return v;
}

// Exact unsigned division by a run-time constant without the ~20-instruction integer divide (Granlund & Montgomery):
// q = n / d for every 32-bit n, d >= 1.  Host side: dfmir_fastdiv(d); device: .div(n), .divmod(n, r).
struct DfmirFastDiv {
  uint32_t d, m, s1, s2;
  __device__ __forceinline__ uint32_t div(uint32_t n) const {
    const uint32_t t = __umulhi(m, n);
    return (t + ((n - t) >> s1)) >> s2;
  }
  __device__ __forceinline__ uint32_t divmod(uint32_t n, uint32_t& r) const {
    const uint32_t q = div(n);
    r = n - q * d;
    return q;
  }
};
static inline DfmirFastDiv dfmir_fastdiv(uint32_t d) {
  DfmirFastDiv f;
  f.d = d;
  uint32_t s = 0;
  while ((1ull << s) < d) ++s;                     // ceil(log2 d)
  f.m = (uint32_t)(((1ull << 32) * ((1ull << s) - d)) / d + 1);
  f.s1 = s < 1 ? s : 1;
  f.s2 = s - f.s1;
  return f;
}

// streaming 128-bit load that does not pollute L1 (read-once data)
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
