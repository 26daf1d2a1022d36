// Shared helpers for libcompyute_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/compyute_b200.h"

namespace cpt {

// thread-local error message returned by cpt_last_error()
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// number of SMs of the current device (cached)
int sm_count();

// kernels launched by this library since load (what bench.py reports as gpu_launches)
void count_launch();

#define CPT_REQUIRE(cond, code, ...)   \
  do {                                 \
    if (!(cond)) {                     \
      ::cpt::set_error(__VA_ARGS__);   \
      return (code);                   \
    }                                  \
  } while (0)

#define CPT_CUDA(expr)                                          \
  do {                                                          \
    cudaError_t _e = (expr);                                    \
    if (_e != cudaSuccess) return ::cpt::cuda_fail(_e, #expr);  \
  } while (0)

#define CPT_LAUNCH_CHECK(name)                                       \
  do {                                                               \
    ::cpt::count_launch();                                           \
    cudaError_t _e = cudaGetLastError();                             \
    if (_e != cudaSuccess) return ::cpt::cuda_fail(_e, name);        \
  } while (0)

// grid for a grid-stride elementwise kernel: enough CTAs for ~8 resident per SM, multiple of SM count
inline int ew_grid(int64_t work_items, int block) {
  int64_t need = (work_items + block - 1) / block;
  int64_t cap = (int64_t)sm_count() * 8;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// ---------------------------------------------------------------- device helpers
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

// block-wide sum; result valid in every thread. `sh` must hold >= 32 floats.
__device__ __forceinline__ float block_sum(float v, float* sh) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  float r = (lane < nw) ? sh[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

// streaming 128-bit loads/stores (read-once data: keep it out of L1)
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(float4* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// BatchNorm apply / backward-apply passes that also emit the channels-last bf16 copy for the neighbouring tensor-core
// convolution (fused_cl.cuh, compiled in tc_host.cu; called by bn.cu)
namespace tc {
int bn_apply_cl(const float* x, const float* w, const float* b, const float* mean, const float* rstd, float* y, void* y_cl, int N,
                int C, int HW, int act, cudaStream_t st);
size_t bn_bwd_apply_cl_ws(int N, int C, int HW);
int bn_bwd_apply_cl(const float* x, const float* dy, const float* w, const float* b, const float* mean, const float* rstd,
                    const float* coef, float count, float* dx, void* dx_cl, float* dx_chan_sum, int N, int C, int HW, int act,
                    void* ws, size_t ws_bytes, cudaStream_t st);
}  // namespace tc

}  // namespace cpt
