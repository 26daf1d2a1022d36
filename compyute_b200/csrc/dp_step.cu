// Data-parallel optimizer step FUSED with the gradient exchange over NVLink 5 / NVSwitch peer memory.
//
// The reference updates every parameter on one device (compyute/nn/optimizers.py:152-176 SGD, :241-271 Adam, :335-362
// AdamW).  In batch-sharded data parallelism every rank holds a replica and the summed gradient is needed before the update;
// the classic form is all-reduce(grad) -> identical update on every rank.  Here the gradient arena G and the parameter arena
// P live in symmetric memory (same offsets on every rank, mapped into every peer and — on NVSwitch — into a multicast
// object), rank r owns the contiguous shard [r*S, (r+1)*S) of the arenas, and ONE kernel per rank
//     1. reads the SUM of its gradient shard over all ranks — `multimem.ld_reduce` (the reduction happens in the switch, one
//        response per 16 bytes) or, without multicast support, `world` peer loads added in rank order;
//     2. applies the update to its shard (the moments exist only for the shard: optimizer state and update traffic / world);
//     3. writes the new parameters to EVERY replica — `multimem.st` (the switch replicates) or `world` peer stores.
// Bytes over each GPU's links: (world-1)/world * 4P in and out — the all-reduce optimum — with no separate collective, no
// second pass over the gradients, and replicas that are bit-identical by construction (one writer per element).
// Cross-rank ordering (all gradients written before step 1; all parameters landed before the next forward) is the caller's:
// a symmetric-memory barrier on the launching stream before and after (compyute_b200/distributed.py SymmetricArena.barrier).
#include <stdlib.h>

#include "common.cuh"

namespace cpt {

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc_ptr) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(mc_ptr)
               : "memory");
  return r;
}
__device__ __forceinline__ void multimem_st(float* mc_ptr, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_ptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float4 ld_sys(const float* p) {
  float4 r;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void st_sys(float* p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}


// summed gradient of 4 consecutive arena elements at float offset `off`
template <bool MC>
__device__ __forceinline__ float4 grad_sum(const cpt_dp_view& d, int64_t off) {
  if (d.pre_reduced) return *reinterpret_cast<const float4*>(d.g_local + off);  // sync_grads() already exchanged and averaged
  if (MC) return multimem_ld_reduce_add(d.g_mc + off);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = 0; r < d.world; ++r) {  // fixed rank order: deterministic; only the owner computes this sum
    const float4 g = ld_sys(reinterpret_cast<const float*>(d.g_peers[r]) + off);
    s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
  }
  return s;
}
template <bool MC>
__device__ __forceinline__ void param_bcast(const cpt_dp_view& d, int64_t off, float4 p) {
  if (MC) { multimem_st(d.p_mc + off, p); return; }
  for (int r = 0; r < d.world; ++r) st_sys(reinterpret_cast<float*>(d.p_peers[r]) + off, p);
}

template <bool MC, int DP_UNROLL>
__global__ void __launch_bounds__(256) dp_adam_kernel(const cpt_dp_view d, float* __restrict__ m, float* __restrict__ v, float lr,
                                                      float beta1, float beta2, float eps, float wd, float m_div, float v_div,
                                                      float grad_scale, int decoupled, const float* __restrict__ live) {
  if (live) { lr = live[0]; m_div = live[1]; v_div = live[2]; }
  const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
  auto upd = [&](float& pv, float gv, float& mv, float& vv) {  // same expression order as adam_kernel (optim.cu)
    gv *= grad_scale;
    if (decoupled) pv *= 1.0f - lr * wd;
    else if (wd != 0.0f) gv = gv + wd * pv;
    mv = beta1 * mv + omb1 * gv;
    vv = beta2 * vv + omb2 * (gv * gv);
    const float mh = mv / m_div, vh = vv / v_div;
    pv -= lr * mh / (sqrtf(vh) + eps);
  };
  // U independent 16-byte gradient sums per thread are in flight before the first is used: a multimem.ld_reduce (or a peer
  // load) takes a round trip through the NVSwitch (~2-3 us), so link bandwidth needs megabytes outstanding per SM
  const int64_t n4 = d.shard_elems / 4, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += stride * DP_UNROLL) {
    float4 g[DP_UNROLL], p[DP_UNROLL], mv[DP_UNROLL], vv[DP_UNROLL];
#pragma unroll
    for (int u = 0; u < DP_UNROLL; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < n4) g[u] = grad_sum<MC>(d, d.shard_off + 4 * i);
    }
#pragma unroll
    for (int u = 0; u < DP_UNROLL; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < n4) {
        p[u] = *reinterpret_cast<const float4*>(d.p_local + d.shard_off + 4 * i);
        mv[u] = reinterpret_cast<float4*>(m)[i];
        vv[u] = reinterpret_cast<float4*>(v)[i];
      }
    }
#pragma unroll
    for (int u = 0; u < DP_UNROLL; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < n4) {
        upd(p[u].x, g[u].x, mv[u].x, vv[u].x); upd(p[u].y, g[u].y, mv[u].y, vv[u].y);
        upd(p[u].z, g[u].z, mv[u].z, vv[u].z); upd(p[u].w, g[u].w, mv[u].w, vv[u].w);
        reinterpret_cast<float4*>(m)[i] = mv[u];
        reinterpret_cast<float4*>(v)[i] = vv[u];
        param_bcast<MC>(d, d.shard_off + 4 * i, p[u]);
      }
    }
  }
  __threadfence_system();
}

template <bool MC, int DP_UNROLL>
__global__ void __launch_bounds__(256) dp_sgd_kernel(const cpt_dp_view d, float* __restrict__ vel, float lr, float momentum, int nesterov,
                                                     float wd, float grad_scale, const float* __restrict__ live) {
  if (live) lr = live[0];
  auto upd = [&](float& pv, float gv, float& vv) {  // same expression order as sgd_kernel (optim.cu)
    gv *= grad_scale;
    if (wd > 0.0f) gv += wd * pv;
    if (momentum > 0.0f) {
      vv = momentum * vv + gv;
      gv = nesterov ? gv + momentum * vv : vv;
    }
    pv = pv - lr * gv;
  };
  const int64_t n4 = d.shard_elems / 4, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += stride * DP_UNROLL) {
    float4 g[DP_UNROLL], p[DP_UNROLL], vv[DP_UNROLL];
#pragma unroll
    for (int u = 0; u < DP_UNROLL; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < n4) g[u] = grad_sum<MC>(d, d.shard_off + 4 * i);
    }
#pragma unroll
    for (int u = 0; u < DP_UNROLL; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < n4) {
        p[u] = *reinterpret_cast<const float4*>(d.p_local + d.shard_off + 4 * i);
        vv[u] = vel ? reinterpret_cast<float4*>(vel)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < DP_UNROLL; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < n4) {
        upd(p[u].x, g[u].x, vv[u].x); upd(p[u].y, g[u].y, vv[u].y); upd(p[u].z, g[u].z, vv[u].z); upd(p[u].w, g[u].w, vv[u].w);
        if (vel) reinterpret_cast<float4*>(vel)[i] = vv[u];
        param_bcast<MC>(d, d.shard_off + 4 * i, p[u]);
      }
    }
  }
  __threadfence_system();
}

static int check_view(const cpt_dp_view* d, const char* who) {
  CPT_REQUIRE(d && d->p_local && d->g_local && d->world >= 1 && d->world <= 64, CPT_ERR_INVALID, "%s: bad view", who);
  CPT_REQUIRE((d->p_mc && d->g_mc) || (d->p_peers && d->g_peers), CPT_ERR_INVALID, "%s: needs multicast or peer pointers", who);
  CPT_REQUIRE(d->shard_off >= 0 && d->shard_elems >= 0 && d->shard_off % 4 == 0 && d->shard_elems % 4 == 0, CPT_ERR_INVALID,
              "%s: shard offset / size must be multiples of 4 floats", who);
  return CPT_OK;
}
// launch shape: CTAs per SM and 16-byte gradient sums in flight per thread (tools/dp_step_bench.py sweeps them on the box:
// CPT_DP_CTAS_PER_SM, CPT_DP_UNROLL)
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
static int dp_unroll() {
  static const int u = env_int("CPT_DP_UNROLL", 8);  // 537 MB at 8 GPUs (tools/dp_step_bench.py): 1.48 ms with 4, 1.24 with 8
  return u >= 8 ? 8 : (u <= 2 ? 2 : 4);
}
static int dp_grid(int64_t elems, int unroll, int max_ctas_per_sm) {
  static const int per_sm_env = env_int("CPT_DP_CTAS_PER_SM", 4);
  const int per_sm = max_ctas_per_sm > 0 ? max_ctas_per_sm : per_sm_env;
  int64_t g = (elems / 4 + 256 * unroll - 1) / (256 * unroll), cap = (int64_t)sm_count() * (per_sm < 1 ? 1 : per_sm);
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}

}  // namespace cpt

using namespace cpt;

extern "C" {

int cpt_dp_adam_step(const cpt_dp_view* view, float* m, float* v, float lr, float beta1, float beta2, float eps, float weight_decay,
                     float m_div, float v_div, float grad_scale, int decoupled, const float* live_scalars, void* stream) {
  if (int e = check_view(view, "dp_adam_step")) return e;
  CPT_REQUIRE(m && v, CPT_ERR_INVALID, "dp_adam_step: moment shards are NULL");
  if (view->shard_elems == 0) return CPT_OK;
  const bool mc = view->p_mc && view->g_mc;
  const int u = dp_unroll();
  const int grid = dp_grid(view->shard_elems, u, view->max_ctas_per_sm);
#define CPT_DP_ADAM(MC, U) dp_adam_kernel<MC, U><<<grid, 256, 0, as_stream(stream)>>>(*view, m, v, lr, beta1, beta2, eps, weight_decay, m_div, v_div, grad_scale, decoupled, live_scalars)
  if (mc) { if (u == 8) CPT_DP_ADAM(true, 8); else if (u == 2) CPT_DP_ADAM(true, 2); else CPT_DP_ADAM(true, 4); }
  else { if (u == 8) CPT_DP_ADAM(false, 8); else if (u == 2) CPT_DP_ADAM(false, 2); else CPT_DP_ADAM(false, 4); }
#undef CPT_DP_ADAM
  CPT_LAUNCH_CHECK("dp_adam_step");
  return CPT_OK;
}

int cpt_dp_sgd_step(const cpt_dp_view* view, float* velocity, float lr, float momentum, int nesterov, float weight_decay,
                    float grad_scale, const float* live_scalars, void* stream) {
  if (int e = check_view(view, "dp_sgd_step")) return e;
  CPT_REQUIRE(velocity || momentum <= 0.0f, CPT_ERR_INVALID, "dp_sgd_step: momentum needs the velocity shard");
  if (view->shard_elems == 0) return CPT_OK;
  const bool mc = view->p_mc && view->g_mc;
  const int u = dp_unroll();
  const int grid = dp_grid(view->shard_elems, u, view->max_ctas_per_sm);
#define CPT_DP_SGD(MC, U) dp_sgd_kernel<MC, U><<<grid, 256, 0, as_stream(stream)>>>(*view, velocity, lr, momentum, nesterov, weight_decay, grad_scale, live_scalars)
  if (mc) { if (u == 8) CPT_DP_SGD(true, 8); else if (u == 2) CPT_DP_SGD(true, 2); else CPT_DP_SGD(true, 4); }
  else { if (u == 8) CPT_DP_SGD(false, 8); else if (u == 2) CPT_DP_SGD(false, 2); else CPT_DP_SGD(false, 4); }
#undef CPT_DP_SGD
  CPT_LAUNCH_CHECK("dp_sgd_step");
  return CPT_OK;
}

}  // extern "C"
