// Shared helpers for the mmdfn_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/mmdfn_b200.h"

namespace mmdfn {

typedef long long i64;

// process-wide count of kernel launches issued by this library (bench.py reports it as gpu_launches)
extern unsigned long long g_launch_count;

#define MMDFN_LAUNCH_CHECK()                              \
  do {                                                    \
    cudaError_t e__ = cudaGetLastError();                 \
    if (e__ != cudaSuccess) return (int)e__;              \
    ++::mmdfn::g_launch_count;                            \
  } while (0)

#define MMDFN_TRY(expr)                                   \
  do {                                                    \
    int r__ = (expr);                                     \
    if (r__ != 0) return r__;                             \
  } while (0)

#define MMDFN_CUDA(expr)                                  \
  do {                                                    \
    cudaError_t e__ = (expr);                             \
    if (e__ != cudaSuccess) return (int)e__;              \
  } while (0)

// argument errors (MMDFN_E*, negative so they never collide with cudaError_t) come from the public header

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline i64 ceil_div64(i64 a, i64 b) { return (a + b - 1) / b; }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming 128-bit global accesses (read-once data: keep it out of L1)
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

}  // namespace mmdfn
