// Fused multi-tensor optimizer steps — HBM-bound (Adam 28 B/param: r p,g,m,v  w p,m,v; SGD+momentum 20 B/param).
// Reference: compyute/nn/optimizers.py:152-176 (SGD), :241-271 (Adam), :335-362 (AdamW).
// One launch updates every parameter: blockIdx.y selects the table entry, blockIdx.x grid-strides inside it.
// grad_scale folds the data-parallel 1/world averaging of the all-reduced (summed) gradient into the update.
#include "common.cuh"

namespace cpt {

__global__ void __launch_bounds__(256) adam_kernel(const cpt_param_entry* __restrict__ table, float lr, float beta1,
                                                   float beta2, float eps, float wd, float m_div, float v_div,
                                                   float grad_scale, int decoupled, const float* __restrict__ live) {
  if (live) { lr = live[0]; m_div = live[1]; v_div = live[2]; }  // per-step scalars from device memory (CUDA-graph replay)
  const cpt_param_entry e = table[blockIdx.y];
  const int64_t n = e.n;
  if (n == 0) return;
  float* __restrict__ p = e.p;
  const float* __restrict__ g = e.g;
  float* __restrict__ m = e.m;
  float* __restrict__ v = e.v;
  const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  const int64_t n4 = vec ? n / 4 : 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;

  auto upd = [&](float& pv, float gv, float& mv, float& vv) {
    gv *= grad_scale;
    if (decoupled) pv *= 1.0f - lr * wd;       // AdamW :344
    else if (wd != 0.0f) gv = gv + wd * pv;   // Adam  :250-253
    mv = beta1 * mv + omb1 * gv;               // :256-258
    vv = beta2 * vv + omb2 * (gv * gv);        // :261-263
    const float mh = mv / m_div, vh = vv / v_div;  // :265-266
    pv -= lr * mh / (sqrtf(vh) + eps);         // :268-269
  };

  for (int64_t i = t; i < n4; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = ld_stream(reinterpret_cast<const float4*>(g) + i);
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    upd(pv.x, gv.x, mv.x, vv.x);
    upd(pv.y, gv.y, mv.y, vv.y);
    upd(pv.z, gv.z, mv.z, vv.z);
    upd(pv.w, gv.w, mv.w, vv.w);
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (int64_t i = n4 * 4 + t; i < n; i += stride) {
    float pv = p[i], mv = m[i], vv = v[i];
    upd(pv, g[i], mv, vv);
    p[i] = pv; m[i] = mv; v[i] = vv;
  }
}

// NAdam.step optimizers.py:437-475: m_hat = mu_next * m / m_div + (1 - mu) * g / g_div, v_hat = v / v_div
__global__ void __launch_bounds__(256) nadam_kernel(const cpt_param_entry* __restrict__ table, float lr, float beta1, float beta2,
                                                    float eps, float wd, float mu, float mu_next, float m_div, float g_div,
                                                    float v_div, float grad_scale, const float* __restrict__ live) {
  if (live) { lr = live[0]; m_div = live[1]; v_div = live[2]; mu = live[3]; mu_next = live[4]; g_div = live[5]; }
  const cpt_param_entry e = table[blockIdx.y];
  const int64_t n = e.n;
  if (n == 0) return;
  float* __restrict__ p = e.p;
  const float* __restrict__ g = e.g;
  float* __restrict__ m = e.m;
  float* __restrict__ v = e.v;
  const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2, omu = 1.0f - mu;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float pv = p[i], gv = g[i] * grad_scale;
    if (wd != 0.0f) gv = gv + wd * pv;
    const float mv = beta1 * m[i] + omb1 * gv;
    const float vv = beta2 * v[i] + omb2 * (gv * gv);
    m[i] = mv;
    v[i] = vv;
    const float mh = mu_next * mv / m_div + omu * gv / g_div;
    const float vh = vv / v_div;
    p[i] = pv - lr * mh / (sqrtf(vh) + eps);
  }
}

__global__ void __launch_bounds__(256) sgd_kernel(const cpt_param_entry* __restrict__ table, float lr, float momentum,
                                                  int nesterov, float wd, float grad_scale, const float* __restrict__ live) {
  if (live) lr = live[0];
  const cpt_param_entry e = table[blockIdx.y];
  const int64_t n = e.n;
  if (n == 0) return;
  float* __restrict__ p = e.p;
  const float* __restrict__ g = e.g;
  float* __restrict__ vel = e.m;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float pv = p[i];
    float gv = g[i] * grad_scale;
    if (wd > 0.0f) gv += wd * pv;              // :160-161
    if (momentum > 0.0f) {
      const float vv = momentum * vel[i] + gv;  // :163-166 (first step: v_prev = 0)
      vel[i] = vv;
      gv = nesterov ? gv + momentum * vv : vv;  // :168-171
    }
    p[i] = pv - lr * gv;                        // :173-174
  }
}

struct Live8 { float v[8]; };
__global__ void set_live_kernel(float* __restrict__ dst, Live8 vals, int n) {
  if ((int)threadIdx.x < n) dst[threadIdx.x] = vals.v[threadIdx.x];
}
__global__ void set_u64_kernel(unsigned long long* __restrict__ dst, unsigned long long v) { *dst = v; }

static dim3 mt_grid(int n_entries, int64_t max_n) {
  int64_t gx = (max_n + 256 * 4 - 1) / (256 * 4);
  const int64_t cap = (int64_t)sm_count() * 4;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  return dim3((unsigned)gx, (unsigned)n_entries);
}

}  // namespace cpt

using namespace cpt;

extern "C" {

int cpt_adam_step(const cpt_param_entry* table, int n_entries, int64_t max_n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, float m_div, float v_div, float grad_scale, int decoupled,
                  const float* live_scalars, void* stream) {
  CPT_REQUIRE(table && n_entries >= 0 && n_entries <= 65535, CPT_ERR_INVALID, "adam_step: bad table");
  if (n_entries == 0) return CPT_OK;
  adam_kernel<<<mt_grid(n_entries, max_n), 256, 0, as_stream(stream)>>>(table, lr, beta1, beta2, eps, weight_decay, m_div,
                                                                      v_div, grad_scale, decoupled, live_scalars);
  CPT_LAUNCH_CHECK("adam_step");
  return CPT_OK;
}

int cpt_nadam_step(const cpt_param_entry* table, int n_entries, int64_t max_n, float lr, float beta1, float beta2, float eps,
                   float weight_decay, float mu, float mu_next, float m_div, float g_div, float v_div, float grad_scale,
                   const float* live_scalars, void* stream) {
  CPT_REQUIRE(table && n_entries >= 0 && n_entries <= 65535, CPT_ERR_INVALID, "nadam_step: bad table");
  if (n_entries == 0) return CPT_OK;
  nadam_kernel<<<mt_grid(n_entries, max_n), 256, 0, as_stream(stream)>>>(table, lr, beta1, beta2, eps, weight_decay, mu, mu_next,
                                                                       m_div, g_div, v_div, grad_scale, live_scalars);
  CPT_LAUNCH_CHECK("nadam_step");
  return CPT_OK;
}

int cpt_set_live_scalars(float* dst, const float* host_values, int n, void* stream) {
  CPT_REQUIRE(dst && host_values && n >= 1 && n <= 8, CPT_ERR_INVALID, "set_live_scalars: bad arguments");
  Live8 v{};
  for (int i = 0; i < n; ++i) v.v[i] = host_values[i];  // copied into the launch parameters: no host-buffer lifetime issue
  set_live_kernel<<<1, 32, 0, as_stream(stream)>>>(dst, v, n);
  CPT_LAUNCH_CHECK("set_live_scalars");
  return CPT_OK;
}

int cpt_set_u64(uint64_t* dst, uint64_t value, void* stream) {
  CPT_REQUIRE(dst, CPT_ERR_INVALID, "set_u64: null pointer");
  set_u64_kernel<<<1, 1, 0, as_stream(stream)>>>(reinterpret_cast<unsigned long long*>(dst), (unsigned long long)value);
  CPT_LAUNCH_CHECK("set_u64");
  return CPT_OK;
}

int cpt_sgd_step(const cpt_param_entry* table, int n_entries, int64_t max_n, float lr, float momentum, int nesterov,
                 float weight_decay, float grad_scale, const float* live_scalars, void* stream) {
  CPT_REQUIRE(table && n_entries >= 0 && n_entries <= 65535, CPT_ERR_INVALID, "sgd_step: bad table");
  if (n_entries == 0) return CPT_OK;
  sgd_kernel<<<mt_grid(n_entries, max_n), 256, 0, as_stream(stream)>>>(table, lr, momentum, nesterov, weight_decay,
                                                                     grad_scale, live_scalars);
  CPT_LAUNCH_CHECK("sgd_step");
  return CPT_OK;
}

}  // extern "C"
