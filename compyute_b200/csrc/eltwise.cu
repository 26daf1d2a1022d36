// Memory-bound elementwise kernels + library plumbing (error string, device info).
// All kernels: 128-bit vectorised main body, scalar tail, grid-stride over a grid sized in multiples
// of the SM count.  Roofline bound: HBM (bytes per element stated per kernel).
#include <stdarg.h>

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
__global__ void __launch_bounds__(256) relu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                       uint8_t* __restrict__ mask, int64_t n8, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const float4* xp = reinterpret_cast<const float4*>(x) + 2 * i;
    float4 a = ld_stream(xp), b = ld_stream(xp + 1);
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    unsigned bits = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[j] = fmaxf(v[j], 0.f);  // numpy.maximum(x, 0): NaN propagates in numpy; fmaxf drops it -> fix below
      bits |= (v[j] > 0.f ? 1u : 0u) << j;
    }
    // numpy.maximum propagates NaN; restore that behaviour
    float w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (w[j] != w[j]) v[j] = w[j];
    float4* yp = reinterpret_cast<float4*>(y) + 2 * i;
    st_stream(yp, make_float4(v[0], v[1], v[2], v[3]));
    st_stream(yp + 1, make_float4(v[4], v[5], v[6], v[7]));
    if (mask) mask[i] = (uint8_t)bits;
  }
  // tail (< 8 elements): one thread
  if (blockIdx.x == 0 && threadIdx.x == 0 && (n8 * 8 < n)) {
    unsigned bits = 0;
    for (int64_t i = n8 * 8; i < n; ++i) {
      float xv = x[i];
      float r = (xv != xv) ? xv : fmaxf(xv, 0.f);
      y[i] = r;
      bits |= (r > 0.f ? 1u : 0u) << (i - n8 * 8);
    }
    if (mask) mask[n8] = (uint8_t)bits;
  }
}

__global__ void __launch_bounds__(256) relu_bwd_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ mask,
                                                       float* __restrict__ dx, int64_t n8, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const float4* gp = reinterpret_cast<const float4*>(dy) + 2 * i;
    float4 a = ld_stream(gp), b = ld_stream(gp + 1);
    unsigned bits = mask[i];
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = v[j] * (float)((bits >> j) & 1u);  // dy * mask (keeps -0.0 / NaN like numpy)
    float4* xp = reinterpret_cast<float4*>(dx) + 2 * i;
    st_stream(xp, make_float4(v[0], v[1], v[2], v[3]));
    st_stream(xp + 1, make_float4(v[4], v[5], v[6], v[7]));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && (n8 * 8 < n)) {
    unsigned bits = mask[n8];
    for (int64_t i = n8 * 8; i < n; ++i) dx[i] = dy[i] * (float)((bits >> (i - n8 * 8)) & 1u);
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
  CPT_REQUIRE(n >= 0 && x && y, CPT_ERR_INVALID, "relu_fwd: bad arguments");
  if (n == 0) return CPT_OK;
  CPT_REQUIRE(aligned16(x) && aligned16(y), CPT_ERR_INVALID, "relu_fwd: pointers must be 16-byte aligned");
  int64_t n8 = n / 8;
  relu_fwd_kernel<<<ew_grid(n8 > 0 ? n8 : 1, 256), 256, 0, as_stream(stream)>>>(x, y, mask, n8, n);
  CPT_LAUNCH_CHECK("relu_fwd");
  return CPT_OK;
}

int cpt_relu_bwd(const float* dy, const uint8_t* mask, float* dx, int64_t n, void* stream) {
  CPT_REQUIRE(n >= 0 && dy && dx && mask, CPT_ERR_INVALID, "relu_bwd: bad arguments");
  if (n == 0) return CPT_OK;
  CPT_REQUIRE(aligned16(dy) && aligned16(dx), CPT_ERR_INVALID, "relu_bwd: pointers must be 16-byte aligned");
  int64_t n8 = n / 8;
  relu_bwd_kernel<<<ew_grid(n8 > 0 ? n8 : 1, 256), 256, 0, as_stream(stream)>>>(dy, mask, dx, n8, n);
  CPT_LAUNCH_CHECK("relu_bwd");
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
