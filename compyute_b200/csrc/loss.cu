// Softmax cross-entropy forward/backward, accuracy count, dropout — small memory-bound kernels that
// close the train step on the device.  Reference: compyute/nn/functional/loss_funcs.py:53-69,
// activation_funcs.py:276-284 (softmax), metric_funcs.py:10-25, regularization_funcs.py:11-32.
#include "common.cuh"

namespace cpt {

// one warp per row (NC up to a few thousand); probs = exp(x - max) / Σ; row_loss[row] = -log(p_t + eta)
__global__ void __launch_bounds__(256) softmax_ce_fwd_kernel(const float* __restrict__ logits, const int32_t* __restrict__ targets,
                                                             float* __restrict__ probs, float* __restrict__ row_loss, int B,
                                                             int NC, float eta) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  const float* x = logits + (int64_t)row * NC;
  float* p = probs + (int64_t)row * NC;
  float mx = -INFINITY;
  for (int j = lane; j < NC; j += 32) mx = fmaxf(mx, x[j]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < NC; j += 32) {
    const float e = expf(x[j] - mx);
    p[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  for (int j = lane; j < NC; j += 32) p[j] = p[j] / sum;
  __syncwarp();
  if (lane == 0) {
    const int t = targets[row];
    const float pt = (t >= 0 && t < NC) ? p[t] : 0.f;
    row_loss[row] = -logf(pt + eta);
  }
}

// Same computation with the row held in REGISTERS between the three phases (max, exp + sum, normalise): one read of the
// logits and one write of the probabilities, 8 B/element — the row-per-warp form above reads x twice and p once more
// (ncu round 1: 25 % of HBM peak at 8192 x 4096).  GROUP threads share a row (a whole block of 256, or one warp), each
// holding up to V float4; needs NC % 4 == 0, 16-byte aligned rows and NC <= GROUP * 4 * V.
template <int GROUP, int V>
__global__ void __launch_bounds__(256) softmax_ce_fwd_reg_kernel(const float* __restrict__ logits, const int32_t* __restrict__ targets,
                                                                 float* __restrict__ probs, float* __restrict__ row_loss, int B,
                                                                 int NC, float eta) {
  constexpr int ROWS = 256 / GROUP;  // rows per block
  __shared__ float red[ROWS][GROUP / 32 > 0 ? GROUP / 32 : 1];
  const int g = threadIdx.x / GROUP, lid = threadIdx.x % GROUP, lane = threadIdx.x & 31, w = lid >> 5;
  const int row = blockIdx.x * ROWS + g;
  const bool live = row < B;
  const int n4 = NC >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(logits + (int64_t)(live ? row : 0) * NC);
  float4 v[V];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int j = lid + i * GROUP;
    if (live && j < n4) {
      v[i] = ld_stream(x4 + j);
      mx = fmaxf(fmaxf(mx, fmaxf(v[i].x, v[i].y)), fmaxf(v[i].z, v[i].w));
    }
  }
  auto group_reduce = [&](float val, bool is_max) -> float {
    val = is_max ? warp_max(val) : warp_sum(val);
    if (GROUP > 32) {
      __syncthreads();
      if (lane == 0) red[g][w] = val;
      __syncthreads();
      float r = is_max ? -INFINITY : 0.f;
#pragma unroll
      for (int k = 0; k < GROUP / 32; ++k) r = is_max ? fmaxf(r, red[g][k]) : r + red[g][k];  // fixed order
      val = r;
    }
    return val;
  };
  mx = group_reduce(mx, true);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int j = lid + i * GROUP;
    if (live && j < n4) {
      v[i].x = expf(v[i].x - mx); v[i].y = expf(v[i].y - mx); v[i].z = expf(v[i].z - mx); v[i].w = expf(v[i].w - mx);
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  sum = group_reduce(sum, false);
  if (!live) return;
  float4* p4 = reinterpret_cast<float4*>(probs + (int64_t)row * NC);
  const int t = targets[row];
  float pt = (t >= 0 && t < NC) ? -1.f : 0.f;  // -1: not seen yet (exactly one thread of the group holds the target element)
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int j = lid + i * GROUP;
    if (j < n4) {
      float4 q;
      q.x = v[i].x / sum; q.y = v[i].y / sum; q.z = v[i].z / sum; q.w = v[i].w / sum;
      st_stream(p4 + j, q);
      if ((t >> 2) == j) { const int e = t & 3; pt = e == 0 ? q.x : (e == 1 ? q.y : (e == 2 ? q.z : q.w)); }
    }
  }
  if (t < 0 || t >= NC) { if (lid == 0) row_loss[row] = -logf(eta); }
  else if (pt >= 0.f) row_loss[row] = -logf(pt + eta);
}

// loss = mean(row_loss): one block, fixed summation order -> run-to-run (and graph-replay) deterministic
__global__ void __launch_bounds__(1024) ce_loss_mean_kernel(const float* __restrict__ row_loss, float* __restrict__ loss, int B) {
  __shared__ float sh[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) s += row_loss[i];
  s = block_sum(s, sh);
  if (threadIdx.x == 0) *loss = s / (float)B;
}

__global__ void __launch_bounds__(256) softmax_ce_bwd_kernel(const float* __restrict__ probs, const int32_t* __restrict__ targets,
                                                             float* __restrict__ dlogits, int64_t total, int NC, float bsz) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t row = i / NC;
    const int j = (int)(i - row * NC);
    const float onehot = (__ldg(targets + row) == j) ? 1.f : 0.f;
    dlogits[i] = (probs[i] - onehot) / bsz;  // loss_funcs.py:69
  }
}

__global__ void __launch_bounds__(256) accuracy_kernel(const float* __restrict__ logits, const int32_t* __restrict__ targets,
                                                       int* __restrict__ correct, int B, int NC) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  const float* x = logits + (int64_t)row * NC;
  float best = -INFINITY;
  int arg = NC;
  for (int j = lane; j < NC; j += 32) {
    const float v = x[j];
    if (v > best) { best = v; arg = j; }  // first maximum within the lane's stride
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }  // numpy argmax: first occurrence
  }
  if (lane == 0 && arg == targets[row]) atomicAdd(correct, 1);
}

// counter-based RNG (splitmix64 finaliser over (seed, index)): stateless, reproducible, one draw per element
__device__ __forceinline__ float u01(uint64_t seed, uint64_t idx) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (float)(z >> 40) * (1.0f / 16777216.0f);  // 24 random bits -> [0, 1)
}

__global__ void __launch_bounds__(256) dropout_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                          int8_t* __restrict__ mask, int64_t n, float keep, uint64_t seed,
                                                          const uint64_t* __restrict__ live_seed) {
  if (live_seed) seed ^= *live_seed * 0xD6E8FEB86659FD93ull;  // per-replay stream when the launch is inside a CUDA graph
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int8_t mk = u01(seed, (uint64_t)i) < keep ? 1 : 0;  // bernoulli(1-p): random() < p_keep (random.py:264)
    mask[i] = mk;
    y[i] = x[i] * (float)mk / keep;  // regularization_funcs.py:22
  }
}

__global__ void __launch_bounds__(256) dropout_bwd_kernel(const float* __restrict__ dy, const int8_t* __restrict__ mask,
                                                          float* __restrict__ dx, int64_t n, float keep) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dx[i] = dy[i] * (float)mask[i] / keep;
}

}  // namespace cpt

using namespace cpt;

extern "C" {

int cpt_softmax_ce_fwd(const float* logits, const int32_t* targets, float* probs, float* loss, float* row_loss, int B, int NC,
                       float eta, void* stream) {
  CPT_REQUIRE(B > 0 && NC > 0 && logits && targets && probs && loss && row_loss, CPT_ERR_INVALID, "softmax_ce_fwd: bad arguments");
  cudaStream_t st = as_stream(stream);
  const bool vec = NC % 4 == 0 && aligned16(logits) && aligned16(probs);
  if (vec && NC > 1024 && NC <= 8192) softmax_ce_fwd_reg_kernel<256, 8><<<B, 256, 0, st>>>(logits, targets, probs, row_loss, B, NC, eta);
  else if (vec && NC > 128 && NC <= 1024) softmax_ce_fwd_reg_kernel<32, 8><<<(B + 7) / 8, 256, 0, st>>>(logits, targets, probs, row_loss, B, NC, eta);
  else softmax_ce_fwd_kernel<<<(B + 7) / 8, 256, 0, st>>>(logits, targets, probs, row_loss, B, NC, eta);
  CPT_LAUNCH_CHECK("softmax_ce_fwd");
  ce_loss_mean_kernel<<<1, 1024, 0, st>>>(row_loss, loss, B);
  CPT_LAUNCH_CHECK("ce_loss_mean");
  return CPT_OK;
}

int cpt_softmax_ce_bwd(const float* probs, const int32_t* targets, float* dlogits, int B, int NC, void* stream) {
  CPT_REQUIRE(B > 0 && NC > 0 && probs && targets && dlogits, CPT_ERR_INVALID, "softmax_ce_bwd: bad arguments");
  const int64_t total = (int64_t)B * NC;
  softmax_ce_bwd_kernel<<<ew_grid(total, 256), 256, 0, as_stream(stream)>>>(probs, targets, dlogits, total, NC,
                                                                          (float)B);
  CPT_LAUNCH_CHECK("softmax_ce_bwd");
  return CPT_OK;
}

int cpt_accuracy_count(const float* logits, const int32_t* targets, int* correct, int B, int NC, void* stream) {
  CPT_REQUIRE(B > 0 && NC > 0 && logits && targets && correct, CPT_ERR_INVALID, "accuracy_count: bad arguments");
  cudaStream_t st = as_stream(stream);
  CPT_CUDA(cudaMemsetAsync(correct, 0, sizeof(int), st));
  accuracy_kernel<<<(B + 7) / 8, 256, 0, st>>>(logits, targets, correct, B, NC);
  CPT_LAUNCH_CHECK("accuracy_count");
  return CPT_OK;
}

int cpt_dropout_fwd(const float* x, float* y, int8_t* mask, int64_t n, float p, uint64_t seed, const uint64_t* live_seed,
                    void* stream) {
  CPT_REQUIRE(n >= 0 && x && y && mask && p >= 0.f && p < 1.f, CPT_ERR_INVALID, "dropout_fwd: bad arguments");
  if (n == 0) return CPT_OK;
  dropout_fwd_kernel<<<ew_grid(n, 256), 256, 0, as_stream(stream)>>>(x, y, mask, n, 1.0f - p, seed, live_seed);
  CPT_LAUNCH_CHECK("dropout_fwd");
  return CPT_OK;
}

int cpt_dropout_bwd(const float* dy, const int8_t* mask, float* dx, int64_t n, float p, void* stream) {
  CPT_REQUIRE(n >= 0 && dy && dx && mask && p >= 0.f && p < 1.f, CPT_ERR_INVALID, "dropout_bwd: bad arguments");
  if (n == 0) return CPT_OK;
  dropout_bwd_kernel<<<ew_grid(n, 256), 256, 0, as_stream(stream)>>>(dy, mask, dx, n, 1.0f - p);
  CPT_LAUNCH_CHECK("dropout_bwd");
  return CPT_OK;
}

}  // extern "C"
