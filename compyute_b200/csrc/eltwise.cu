// Memory-bound elementwise kernels + library plumbing (error string, device info).
// All kernels: 128-bit vectorised main body, scalar tail, grid-stride over a grid sized in multiples
// of the SM count.  Roofline bound: HBM (bytes per element stated per kernel).
#include <stdarg.h>

#include <cuda_bf16.h>

#include "common.cuh"

namespace cpt {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return CPT_ERR_CUDA;
}

static unsigned long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }

int sm_count() {
  static int cached[16] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// ------------------------------------------------------------------ ReLU (8 B/elem + 1/8 B mask)
// Mask layout (private to this pair of kernels): elements are processed in chunks of 1024 by one warp; lane l handles the
// float4s at (j*32 + l), j = 0..7, of the chunk and owns mask word chunk*32 + l (bit 4j+i = element i of its j-th float4),
// so both the eight 128-bit loads per thread and the 128-byte mask store per warp are fully coalesced.  The tail
// (< 1024 elements) uses plain bit order after the last full chunk.
// `lp` (optional): the same values as bf16, same linear order — the pre-cast operand of a Linear layer that consumes the result
// (row pitch == row length, i.e. the layout of cpt_cast_bf16 for a feature count that is a multiple of 8).
__device__ __forceinline__ uint2 pack_bf16x4(float a, float b, float c, float d) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  return make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
}

__global__ void __launch_bounds__(256) relu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                       uint32_t* __restrict__ mask, int64_t n_chunks, int64_t n,
                                                       __nv_bfloat16* __restrict__ lp) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t c = warp0; c < n_chunks; c += nwarps) {
    const float4* xp = reinterpret_cast<const float4*>(x) + c * 256 + lane;
    float4* yp = reinterpret_cast<float4*>(y) + c * 256 + lane;
    float4 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = ld_stream(xp + j * 32);
    uint32_t bits = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float e[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float r = (e[i] != e[i]) ? e[i] : fmaxf(e[i], 0.f);  // numpy.maximum propagates NaN
        bits |= (r > 0.f ? 1u : 0u) << (4 * j + i);
        e[i] = r;
      }
      st_stream(yp + j * 32, make_float4(e[0], e[1], e[2], e[3]));
      if (lp) reinterpret_cast<uint2*>(lp)[c * 256 + lane + j * 32] = pack_bf16x4(e[0], e[1], e[2], e[3]);
    }
    if (mask) mask[c * 32 + lane] = bits;
  }
  // tail: one warp, plain bit order
  if (warp0 == 0) {
    const int64_t t0 = n_chunks * 1024;
    for (int64_t base = t0; base < n; base += 32) {
      const int64_t i = base + lane;
      float r = 0.f;
      if (i < n) {
        const float xv = x[i];
        r = (xv != xv) ? xv : fmaxf(xv, 0.f);
        y[i] = r;
        if (lp) lp[i] = __float2bfloat16_rn(r);
      }
      const uint32_t b = __ballot_sync(0xffffffffu, i < n && r > 0.f);
      if (mask && lane == 0) mask[n_chunks * 32 + (base - t0) / 32] = b;
    }
  }
}

__global__ void __launch_bounds__(256) relu_bwd_kernel(const float* __restrict__ dy, const uint32_t* __restrict__ mask,
                                                       float* __restrict__ dx, int64_t n_chunks, int64_t n,
                                                       __nv_bfloat16* __restrict__ lp) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t c = warp0; c < n_chunks; c += nwarps) {
    const float4* gp = reinterpret_cast<const float4*>(dy) + c * 256 + lane;
    float4* xp = reinterpret_cast<float4*>(dx) + c * 256 + lane;
    float4 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = ld_stream(gp + j * 32);
    const uint32_t bits = mask[c * 32 + lane];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // dy * mask (keeps -0.0 / NaN like numpy's float * bool)
      const float4 o = make_float4(v[j].x * (float)((bits >> (4 * j)) & 1u), v[j].y * (float)((bits >> (4 * j + 1)) & 1u),
                                   v[j].z * (float)((bits >> (4 * j + 2)) & 1u), v[j].w * (float)((bits >> (4 * j + 3)) & 1u));
      st_stream(xp + j * 32, o);
      if (lp) reinterpret_cast<uint2*>(lp)[c * 256 + lane + j * 32] = pack_bf16x4(o.x, o.y, o.z, o.w);
    }
  }
  if (warp0 == 0) {
    const int64_t t0 = n_chunks * 1024;
    for (int64_t base = t0; base < n; base += 32) {
      const int64_t i = base + lane;
      const uint32_t b = mask[n_chunks * 32 + (base - t0) / 32];
      if (i < n) {
        const float o = dy[i] * (float)((b >> lane) & 1u);
        dx[i] = o;
        if (lp) lp[i] = __float2bfloat16_rn(o);
      }
    }
  }
}

// dx = dy * mask for a mask in PLAIN bit order (element e = bit e % 32 of word e / 32), as written by the fused
// Linear + ReLU epilogue (cpt_linear_relu_fwd_bf16); lp as in relu_bwd_kernel.
__global__ void __launch_bounds__(256) relu_bwd_plain_kernel(const float* __restrict__ dy, const uint32_t* __restrict__ mask,
                                                             float* __restrict__ dx, int64_t n4, int64_t n,
                                                             __nv_bfloat16* __restrict__ lp) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = t; i < n4; i += stride) {
    const float4 v = ld_stream(reinterpret_cast<const float4*>(dy) + i);
    const uint32_t b = __ldg(mask + (i >> 3)) >> ((i & 7) * 4);
    const float4 o = make_float4(v.x * (float)(b & 1u), v.y * (float)((b >> 1) & 1u), v.z * (float)((b >> 2) & 1u),
                                 v.w * (float)((b >> 3) & 1u));
    st_stream(reinterpret_cast<float4*>(dx) + i, o);
    if (lp) reinterpret_cast<uint2*>(lp)[i] = pack_bf16x4(o.x, o.y, o.z, o.w);
  }
  for (int64_t i = n4 * 4 + t; i < n; i += stride) {
    const float o = dy[i] * (float)((__ldg(mask + (i >> 5)) >> (i & 31)) & 1u);
    dx[i] = o;
    if (lp) lp[i] = __float2bfloat16_rn(o);
  }
}

// ------------------------------------------------------------------ a += b (12 B/elem), y = alpha x (+y)
__global__ void __launch_bounds__(256) add_inplace_kernel(float* __restrict__ a, const float* __restrict__ b,
                                                          int64_t n4, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = t; i < n4; i += stride) {
    float4 u = reinterpret_cast<float4*>(a)[i];
    float4 v = ld_stream(reinterpret_cast<const float4*>(b) + i);
    u.x += v.x; u.y += v.y; u.z += v.z; u.w += v.w;
    reinterpret_cast<float4*>(a)[i] = u;
  }
  for (int64_t i = n4 * 4 + t; i < n; i += stride) a[i] += b[i];
}

__global__ void __launch_bounds__(256) axpby_kernel(float* __restrict__ y, const float* __restrict__ x, float alpha,
                                                    int accumulate, int64_t n4, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = t; i < n4; i += stride) {
    float4 v = ld_stream(reinterpret_cast<const float4*>(x) + i);
    float4 u = make_float4(alpha * v.x, alpha * v.y, alpha * v.z, alpha * v.w);
    if (accumulate) {
      float4 o = reinterpret_cast<float4*>(y)[i];
      u.x += o.x; u.y += o.y; u.z += o.z; u.w += o.w;
    }
    reinterpret_cast<float4*>(y)[i] = u;
  }
  for (int64_t i = n4 * 4 + t; i < n; i += stride) y[i] = alpha * x[i] + (accumulate ? y[i] : 0.f);
}

__global__ void __launch_bounds__(256) fill_kernel(float* __restrict__ a, float value, int64_t n4, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const float4 v = make_float4(value, value, value, value);
  for (int64_t i = t; i < n4; i += stride) reinterpret_cast<float4*>(a)[i] = v;
  for (int64_t i = n4 * 4 + t; i < n; i += stride) a[i] = value;
}

// ------------------------------------------------------------------ NaN flag (4 B/elem) and sum
__global__ void __launch_bounds__(256) isnan_kernel(const float* __restrict__ x, int64_t n4, int64_t n,
                                                    int* __restrict__ flag) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int bad = 0;
  for (int64_t i = t; i < n4; i += stride) {
    float4 v = ld_stream(reinterpret_cast<const float4*>(x) + i);
    bad |= (v.x != v.x) | (v.y != v.y) | (v.z != v.z) | (v.w != v.w);
  }
  for (int64_t i = n4 * 4 + t; i < n; i += stride) bad |= (x[i] != x[i]);
  bad = __any_sync(0xffffffffu, bad);
  if (bad && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

__global__ void __launch_bounds__(256) sum_kernel(const float* __restrict__ x, int64_t n4, int64_t n,
                                                  float* __restrict__ out) {
  __shared__ float sh[32];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float s = 0.f;
  for (int64_t i = t; i < n4; i += stride) {
    float4 v = ld_stream(reinterpret_cast<const float4*>(x) + i);
    s += (v.x + v.y) + (v.z + v.w);
  }
  for (int64_t i = n4 * 4 + t; i < n; i += stride) s += x[i];
  s = block_sum(s, sh);
  if (threadIdx.x == 0) atomicAdd(out, s);
}

}  // namespace cpt

using namespace cpt;

extern "C" {

const char* cpt_last_error(void) { return g_err; }
int cpt_version(void) { return 100; }
uint64_t cpt_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int cpt_device_info(int device, int* sm, int* major, int* minor, size_t* smem_optin, size_t* total_mem) {
  cudaDeviceProp p;
  CPT_CUDA(cudaGetDeviceProperties(&p, device));
  if (sm) *sm = p.multiProcessorCount;
  if (major) *major = p.major;
  if (minor) *minor = p.minor;
  if (smem_optin) *smem_optin = p.sharedMemPerBlockOptin;
  if (total_mem) *total_mem = p.totalGlobalMem;
  return CPT_OK;
}

int cpt_relu_fwd(const float* x, float* y, uint8_t* mask, int64_t n, void* stream) {
  return cpt_relu_fwd_lp(x, y, mask, nullptr, n, stream);
}
int cpt_relu_bwd(const float* dy, const uint8_t* mask, float* dx, int64_t n, void* stream) {
  return cpt_relu_bwd_lp(dy, mask, dx, nullptr, n, stream);
}

int cpt_relu_fwd_lp(const float* x, float* y, uint8_t* mask, void* y_bf16, int64_t n, void* stream) {
  CPT_REQUIRE(n >= 0 && x && y, CPT_ERR_INVALID, "relu_fwd: bad arguments");
  CPT_REQUIRE(!y_bf16 || (reinterpret_cast<uintptr_t>(y_bf16) & 7) == 0, CPT_ERR_INVALID, "relu_fwd: y_bf16 must be 8-byte aligned");
  if (n == 0) return CPT_OK;
  CPT_REQUIRE(aligned16(x) && aligned16(y) && (!mask || (reinterpret_cast<uintptr_t>(mask) & 3) == 0), CPT_ERR_INVALID,
              "relu_fwd: x, y must be 16-byte aligned and mask 4-byte aligned");
  const int64_t n_chunks = n / 1024;
  relu_fwd_kernel<<<ew_grid((n_chunks > 0 ? n_chunks : 1) * 32, 256), 256, 0, as_stream(stream)>>>(
      x, y, reinterpret_cast<uint32_t*>(mask), n_chunks, n, reinterpret_cast<__nv_bfloat16*>(y_bf16));
  CPT_LAUNCH_CHECK("relu_fwd");
  return CPT_OK;
}

int cpt_relu_bwd_lp(const float* dy, const uint8_t* mask, float* dx, void* dx_bf16, int64_t n, void* stream) {
  CPT_REQUIRE(n >= 0 && dy && dx && mask, CPT_ERR_INVALID, "relu_bwd: bad arguments");
  CPT_REQUIRE(!dx_bf16 || (reinterpret_cast<uintptr_t>(dx_bf16) & 7) == 0, CPT_ERR_INVALID, "relu_bwd: dx_bf16 must be 8-byte aligned");
  if (n == 0) return CPT_OK;
  CPT_REQUIRE(aligned16(dy) && aligned16(dx), CPT_ERR_INVALID, "relu_bwd: pointers must be 16-byte aligned");
  const int64_t n_chunks = n / 1024;
  relu_bwd_kernel<<<ew_grid((n_chunks > 0 ? n_chunks : 1) * 32, 256), 256, 0, as_stream(stream)>>>(
      dy, reinterpret_cast<const uint32_t*>(mask), dx, n_chunks, n, reinterpret_cast<__nv_bfloat16*>(dx_bf16));
  CPT_LAUNCH_CHECK("relu_bwd");
  return CPT_OK;
}

int cpt_relu_bwd_plain(const float* dy, const uint8_t* mask, float* dx, void* dx_bf16, int64_t n, void* stream) {
  CPT_REQUIRE(n >= 0 && dy && dx && mask, CPT_ERR_INVALID, "relu_bwd_plain: bad arguments");
  CPT_REQUIRE((reinterpret_cast<uintptr_t>(mask) & 3) == 0 && (!dx_bf16 || (reinterpret_cast<uintptr_t>(dx_bf16) & 7) == 0), CPT_ERR_INVALID,
              "relu_bwd_plain: mask must be 4-byte, dx_bf16 8-byte aligned");
  if (n == 0) return CPT_OK;
  const int64_t n4 = (aligned16(dy) && aligned16(dx)) ? n / 4 : 0;
  relu_bwd_plain_kernel<<<ew_grid(n4 > 0 ? n4 : n, 256), 256, 0, as_stream(stream)>>>(dy, reinterpret_cast<const uint32_t*>(mask), dx, n4, n,
                                                                                     reinterpret_cast<__nv_bfloat16*>(dx_bf16));
  CPT_LAUNCH_CHECK("relu_bwd_plain");
  return CPT_OK;
}

int cpt_add_inplace(float* a, const float* b, int64_t n, void* stream) {
  CPT_REQUIRE(n >= 0 && a && b, CPT_ERR_INVALID, "add_inplace: bad arguments");
  if (n == 0) return CPT_OK;
  int64_t n4 = (aligned16(a) && aligned16(b)) ? n / 4 : 0;
  add_inplace_kernel<<<ew_grid(n4 > 0 ? n4 : n, 256), 256, 0, as_stream(stream)>>>(a, b, n4, n);
  CPT_LAUNCH_CHECK("add_inplace");
  return CPT_OK;
}

int cpt_axpby(float* y, const float* x, float alpha, int accumulate, int64_t n, void* stream) {
  CPT_REQUIRE(n >= 0 && y && x, CPT_ERR_INVALID, "axpby: bad arguments");
  if (n == 0) return CPT_OK;
  int64_t n4 = (aligned16(y) && aligned16(x)) ? n / 4 : 0;
  axpby_kernel<<<ew_grid(n4 > 0 ? n4 : n, 256), 256, 0, as_stream(stream)>>>(y, x, alpha, accumulate, n4, n);
  CPT_LAUNCH_CHECK("axpby");
  return CPT_OK;
}

int cpt_fill(float* a, float value, int64_t n, void* stream) {
  CPT_REQUIRE(n >= 0 && a, CPT_ERR_INVALID, "fill: bad arguments");
  if (n == 0) return CPT_OK;
  int64_t n4 = aligned16(a) ? n / 4 : 0;
  fill_kernel<<<ew_grid(n4 > 0 ? n4 : n, 256), 256, 0, as_stream(stream)>>>(a, value, n4, n);
  CPT_LAUNCH_CHECK("fill");
  return CPT_OK;
}

int cpt_isnan_flag(const float* x, int64_t n, int* flag, void* stream) {
  CPT_REQUIRE(n >= 0 && x && flag, CPT_ERR_INVALID, "isnan_flag: bad arguments");
  if (n == 0) return CPT_OK;
  int64_t n4 = aligned16(x) ? n / 4 : 0;
  isnan_kernel<<<ew_grid(n4 > 0 ? n4 : n, 256), 256, 0, as_stream(stream)>>>(x, n4, n, flag);
  CPT_LAUNCH_CHECK("isnan_flag");
  return CPT_OK;
}

int cpt_sum(const float* x, int64_t n, float* out, void* stream) {
  CPT_REQUIRE(n >= 0 && x && out, CPT_ERR_INVALID, "sum: bad arguments");
  CPT_CUDA(cudaMemsetAsync(out, 0, sizeof(float), as_stream(stream)));
  if (n == 0) return CPT_OK;
  int64_t n4 = aligned16(x) ? n / 4 : 0;
  int grid = ew_grid(n4 > 0 ? n4 : n, 256);
  if (grid > sm_count() * 2) grid = sm_count() * 2;
  sum_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, n4, n, out);
  CPT_LAUNCH_CHECK("sum");
  return CPT_OK;
}

}  // extern "C"
