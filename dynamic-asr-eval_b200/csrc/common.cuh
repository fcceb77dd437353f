// Shared device/host helpers for libdae.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math_constants.h>
#include <atomic>
#include "../../include/dae.h"

namespace dae {

extern std::atomic<int64_t> g_launches;   // bumped once per kernel launch (dae_launch_count)

constexpr int kNumSMs = 148;              // B200: 2 dies x 74 SMs

#define DAE_LAUNCH_OK()                                                    \
  do {                                                                     \
    ::dae::g_launches.fetch_add(1, std::memory_order_relaxed);             \
    cudaError_t e__ = cudaGetLastError();                                  \
    if (e__ != cudaSuccess) return (int)e__;                               \
  } while (0)

#define DAE_CUDA(call)                                                     \
  do {                                                                     \
    cudaError_t e__ = (call);                                              \
    if (e__ != cudaSuccess) return (int)e__;                               \
  } while (0)

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Streaming 128-bit load that does not allocate in L1 (data is touched once per CTA).
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float ld_stream1(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
// Streaming 128-bit store (write-once outputs).
__device__ __forceinline__ void st_stream4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Raise a kernel's dynamic shared-memory limit, once per (kernel, device, size): the driver call costs a few
// microseconds of host time, which shows when the GPU is waiting for the launch.  The cache is guarded by a mutex.
cudaError_t ensure_dyn_smem_impl(const void* kern, int bytes);
template <typename Kern>
inline cudaError_t ensure_dyn_smem(Kern kern, int bytes) {
  return ensure_dyn_smem_impl(reinterpret_cast<const void*>(kern), bytes);
}

// SM count of the current device (cached per device; 0 on error).
int sm_count();

__host__ __device__ __forceinline__ bool aligned16(const void* p) {
  return (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
}

}  // namespace dae
