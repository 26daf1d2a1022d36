// BatchNorm 1-D / 2-D forward (train, eval) and backward over x viewed as (N, C, HW) — HBM-bound.
// Reference: compyute/nn/functional/normalization_funcs.py:10-177.
//
// Training forward = 3 launches: partial statistics (grid C x S) -> finalize (per channel, fixed-order
// merge: deterministic) -> apply.  x is read twice, y written once: 12 B/elem algorithmic.
// Backward = partial sums (Σdy, Σdy·x̂) -> finalize (dw, db, coefficients) -> apply: reads dy, x twice each,
// writes dx: 20 B/elem.  x̂ is recomputed from (x, mean, rstd) instead of being cached like the reference does.
//
// Variance: shifted sums Σ(x-K), Σ(x-K)² with K = first element of the channel, so the single pass does not
// suffer E[x²]-E[x]² cancellation when |mean| >> std.
#include "common.cuh"

namespace cpt {

constexpr int BN_THREADS = 256;

// ReLU fused behind the affine output (Sequential peephole BatchNorm -> ReLU): the activation's mask is recomputed in
// backward from the SAME expression the forward evaluates, y = fmaf(w, (x - mean) * rstd, b) > 0, so nothing is cached
// for it and dy * mask keeps numpy's float * bool semantics (-0.0, NaN).
__device__ __forceinline__ float relu_fwd_val(float v) { return (v != v) ? v : fmaxf(v, 0.f); }  // numpy.maximum keeps NaN
__device__ __forceinline__ float relu_masked(float g, float x, float mu, float rs, float ww, float bb) {
  return g * (fmaf(ww, (x - mu) * rs, bb) > 0.f ? 1.f : 0.f);
}

// partial sums for channel c = blockIdx.x, split s = blockIdx.y.  MODE 0: stats (a = x-K, b = (x-K)²);
// MODE 1: backward (a = dy, b = dy * x̂); MODE 2: backward behind a fused ReLU (dy masked first).
template <int MODE, int VEC>
__global__ void __launch_bounds__(BN_THREADS) bn_partial_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                 const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                 const float* __restrict__ w, const float* __restrict__ b,
                                                                 float2* __restrict__ partial, int N, int C, int HW) {
  __shared__ float sh[32];
  const int c = blockIdx.x, S = gridDim.y, s = blockIdx.y;
  const int HWv = HW / VEC;
  const int64_t items = (int64_t)N * HWv;
  const int64_t per = (items + S - 1) / S;
  const int64_t lo = per * s, hi = (lo + per < items) ? lo + per : items;
  float k0, k1, ww = 0.f, bb = 0.f;
  if (MODE == 0) { k0 = __ldg(x + (int64_t)c * HW); k1 = 0.f; }
  else { k0 = __ldg(mean + c); k1 = __ldg(rstd + c); }
  if (MODE == 2) { ww = __ldg(w + c); bb = __ldg(b + c); }
  float sa = 0.f, sb = 0.f;
  for (int64_t it = lo + threadIdx.x; it < hi; it += BN_THREADS) {
    const int64_t n = it / HWv;
    const int j = (int)(it - n * HWv);
    const int64_t off = (n * C + c) * (int64_t)HW + (int64_t)j * VEC;
    float xv[VEC], gv[VEC];
    if (VEC == 4) {
      float4 v = ld_stream(reinterpret_cast<const float4*>(x + off));
      xv[0] = v.x; xv[1] = v.y; xv[2] = v.z; xv[3] = v.w;
      if (MODE >= 1) {
        float4 g = ld_stream(reinterpret_cast<const float4*>(dy + off));
        gv[0] = g.x; gv[1] = g.y; gv[2] = g.z; gv[3] = g.w;
      }
    } else {
      xv[0] = x[off];
      if (MODE >= 1) gv[0] = dy[off];
    }
#pragma unroll
    for (int u = 0; u < VEC; ++u) {
      if (MODE == 0) {
        const float d = xv[u] - k0;
        sa += d;
        sb = fmaf(d, d, sb);
      } else {
        const float g = MODE == 2 ? relu_masked(gv[u], xv[u], k0, k1, ww, bb) : gv[u];
        sa += g;
        sb = fmaf(g, (xv[u] - k0) * k1, sb);
      }
    }
  }
  sa = block_sum(sa, sh);
  sb = block_sum(sb, sh);
  if (threadIdx.x == 0) partial[(int64_t)c * S + s] = make_float2(sa, sb);
}

// HW == 1 (BatchNorm1D on (N, C)): threads map across channels so loads stay coalesced.
template <int MODE>
__global__ void __launch_bounds__(256) bn_partial_hw1_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                             const float* __restrict__ mean, const float* __restrict__ rstd,
                                                             const float* __restrict__ w, const float* __restrict__ b,
                                                             float2* __restrict__ partial, int N, int C) {
  __shared__ float sa_s[8][33], sb_s[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx, S = gridDim.y, s = blockIdx.y;
  float sa = 0.f, sb = 0.f;
  if (c < C) {
    float k0, k1, ww = 0.f, bb = 0.f;
    if (MODE == 0) { k0 = __ldg(x + c); k1 = 0.f; }
    else { k0 = __ldg(mean + c); k1 = __ldg(rstd + c); }
    if (MODE == 2) { ww = __ldg(w + c); bb = __ldg(b + c); }
    for (int n = s * 8 + ty; n < N; n += S * 8) {
      const float xv = x[(int64_t)n * C + c];
      if (MODE == 0) {
        const float d = xv - k0;
        sa += d;
        sb = fmaf(d, d, sb);
      } else {
        float g = dy[(int64_t)n * C + c];
        if (MODE == 2) g = relu_masked(g, xv, k0, k1, ww, bb);
        sa += g;
        sb = fmaf(g, (xv - k0) * k1, sb);
      }
    }
  }
  sa_s[ty][tx] = sa;
  sb_s[ty][tx] = sb;
  __syncthreads();
  if (ty == 0 && c < C) {
#pragma unroll
    for (int r = 1; r < 8; ++r) { sa += sa_s[r][tx]; sb += sb_s[r][tx]; }
    partial[(int64_t)c * S + s] = make_float2(sa, sb);
  }
}

__global__ void bn_fwd_finalize_kernel(const float* __restrict__ x, const float2* __restrict__ partial, int S,
                                       const float* __restrict__ rmean, const float* __restrict__ rvar,
                                       float* __restrict__ rmean_out, float* __restrict__ rvar_out,
                                       float* __restrict__ save_mean, float* __restrict__ save_rstd, int C, int HW,
                                       float count, float m, float eps) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float sa = 0.f, sb = 0.f;
  for (int s = 0; s < S; ++s) {  // fixed order -> run-to-run deterministic
    const float2 p = partial[(int64_t)c * S + s];
    sa += p.x;
    sb += p.y;
  }
  const float k0 = x[(int64_t)c * HW];
  const float dm = sa / count;
  const float mean = k0 + dm;
  float var = sb / count - dm * dm;  // biased (ddof = 0), used for normalisation
  var = fmaxf(var, 0.f);
  save_mean[c] = mean;
  save_rstd[c] = 1.0f / sqrtf(var + eps);
  // running stats use the unbiased variance (normalization_funcs.py:147); count == 1 -> nan like numpy
  const float var_unbiased = var * count / (count - 1.0f);
  rmean_out[c] = rmean[c] * (1.0f - m) + mean * m;
  rvar_out[c] = rvar[c] * (1.0f - m) + var_unbiased * m;
}

// Batch statistics from the column sums a convolution's epilogue left behind (cpt_conv2d_fprop_*_stats): block per 32
// channels, lanes = channels (coalesced float2 loads), the 32 warps split the slots and their partials are combined in fixed
// order -> deterministic; sums in double; var = E[a²] - E[a]² on the bias-free accumulators a (the convolution's bias only
// shifts the mean).
__global__ void __launch_bounds__(1024) bn_fwd_finalize_presum_kernel(const float* __restrict__ stats, int slots,
                                                                      const float* __restrict__ conv_bias,
                                                                      const float* __restrict__ rmean, const float* __restrict__ rvar,
                                                                      float* __restrict__ rmean_out, float* __restrict__ rvar_out,
                                                                      float* __restrict__ save_mean, float* __restrict__ save_rstd,
                                                                      int C, float count, float m, float eps) {
  __shared__ double sh1[32][33], sh2[32][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  double s1 = 0.0, s2 = 0.0;
  if (c < C) {
    for (int s = warp; s < slots; s += 32) {
      const float2 v = reinterpret_cast<const float2*>(stats)[(int64_t)s * C + c];
      s1 += (double)v.x;
      s2 += (double)v.y;
    }
  }
  sh1[warp][lane] = s1;
  sh2[warp][lane] = s2;
  __syncthreads();
  if (warp != 0 || c >= C) return;
  s1 = s2 = 0.0;
#pragma unroll
  for (int w = 0; w < 32; ++w) { s1 += sh1[w][lane]; s2 += sh2[w][lane]; }
  const double n = (double)count, mu = s1 / n;
  double var = s2 / n - mu * mu;
  if (var < 0.0) var = 0.0;
  const float mean = (float)mu + (conv_bias ? conv_bias[c] : 0.f);
  save_mean[c] = mean;
  save_rstd[c] = 1.0f / sqrtf((float)var + eps);
  const float var_unbiased = (float)(var * n / (n - 1.0));
  rmean_out[c] = rmean[c] * (1.0f - m) + mean * m;
  rvar_out[c] = rvar[c] * (1.0f - m) + var_unbiased * m;
}

// ---- synchronised BatchNorm (statistics over the global batch of all data-parallel ranks) ----
// local (mean, M2 = Σ(x - mean)², count) of this rank's shard from the shifted partial sums -> stats[3][C]
__global__ void bn_local_stats_kernel(const float* __restrict__ x, const float2* __restrict__ partial, int S,
                                      float* __restrict__ stats, int C, int HW, float count) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double sa = 0.0, sb = 0.0;
  for (int s = 0; s < S; ++s) {
    const float2 p = partial[(int64_t)c * S + s];
    sa += (double)p.x;
    sb += (double)p.y;
  }
  const double dm = sa / (double)count;
  double m2 = sb - (double)count * dm * dm;
  stats[c] = (float)((double)x[(int64_t)c * HW] + dm);
  stats[C + c] = (float)(m2 > 0.0 ? m2 : 0.0);
  stats[2 * C + c] = count;
}

// merges the ranks' (mean, M2, count) in rank order with the pairwise update of Chan et al. — every rank runs the same
// arithmetic on the same gathered values, so all replicas get bit-identical statistics
__global__ void bn_fwd_finalize_merged_kernel(const float* __restrict__ gathered, int world, const float* __restrict__ rmean,
                                              const float* __restrict__ rvar, float* __restrict__ rmean_out,
                                              float* __restrict__ rvar_out, float* __restrict__ save_mean,
                                              float* __restrict__ save_rstd, float* __restrict__ global_count, int C, float m,
                                              float eps) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double n = 0.0, mean = 0.0, m2 = 0.0;
  for (int r = 0; r < world; ++r) {
    const float* g = gathered + (int64_t)r * 3 * C;
    const double nb = (double)g[2 * C + c];
    if (nb <= 0.0) continue;
    const double delta = (double)g[c] - mean, tot = n + nb;
    mean += delta * nb / tot;
    m2 += (double)g[C + c] + delta * delta * n * nb / tot;
    n = tot;
  }
  const double var = m2 / n;
  if (c == 0) *global_count = (float)n;
  save_mean[c] = (float)mean;
  save_rstd[c] = 1.0f / sqrtf((float)var + eps);
  rmean_out[c] = rmean[c] * (1.0f - m) + (float)mean * m;
  rvar_out[c] = rvar[c] * (1.0f - m) + (float)(m2 / (n - 1.0)) * m;
}

// local backward sums -> sums[2][C] = (Σdy, Σdy·x̂) (to be all-reduced), dw / db (local: the gradient exchange averages them)
__global__ void bn_bwd_local_sums_kernel(const float2* __restrict__ partial, int S, float* __restrict__ sums,
                                         float* __restrict__ dw, float* __restrict__ db, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float sa = 0.f, sb = 0.f;
  for (int s = 0; s < S; ++s) {
    const float2 p = partial[(int64_t)c * S + s];
    sa += p.x;
    sb += p.y;
  }
  sums[c] = sa;
  sums[C + c] = sb;
  db[c] = sa;
  dw[c] = sb;
}

// the global element count lives on the device (written by the merged forward), so the apply kernels run with count = 1 and
// coefficients pre-divided by n: dx = w·rstd·(dy - Σdy/n - x̂·Σ(dy·x̂)/n)
__global__ void bn_bwd_coef_kernel(const float* __restrict__ sums, const float* __restrict__ w, const float* __restrict__ rstd,
                                   float* __restrict__ coef, int C, const float* __restrict__ global_count) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float n = *global_count;
  coef[3 * c + 0] = w[c] * rstd[c];
  coef[3 * c + 1] = sums[c] / n;
  coef[3 * c + 2] = sums[C + c] / n;
}

__global__ void bn_eval_stats_kernel(const float* __restrict__ rmean, const float* __restrict__ rvar,
                                     float* __restrict__ save_mean, float* __restrict__ save_rstd, int C, float eps) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  save_mean[c] = rmean[c];
  save_rstd[c] = 1.0f / sqrtf(rvar[c] + eps);
}

// y = w * ((x - mean) * rstd) + b   (RELU: followed by max(., 0))
template <int VEC, bool RELU>
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ b, const float* __restrict__ mean,
                                                       const float* __restrict__ rstd, float* __restrict__ y,
                                                       int64_t items, int C, int HWv) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += stride) {
    const int c = (int)((it / HWv) % C);
    const float mu = __ldg(mean + c), rs = __ldg(rstd + c), ww = __ldg(w + c), bb = __ldg(b + c);
    if (VEC == 4) {
      float4 v = ld_stream(reinterpret_cast<const float4*>(x) + it);
      v.x = fmaf(ww, (v.x - mu) * rs, bb);
      v.y = fmaf(ww, (v.y - mu) * rs, bb);
      v.z = fmaf(ww, (v.z - mu) * rs, bb);
      v.w = fmaf(ww, (v.w - mu) * rs, bb);
      if (RELU) { v.x = relu_fwd_val(v.x); v.y = relu_fwd_val(v.y); v.z = relu_fwd_val(v.z); v.w = relu_fwd_val(v.w); }
      st_stream(reinterpret_cast<float4*>(y) + it, v);
    } else {
      const float r = fmaf(ww, (x[it] - mu) * rs, bb);
      y[it] = RELU ? relu_fwd_val(r) : r;
    }
  }
}

// per channel: dw = Σ dy x̂, db = Σ dy, coef = (A, Bc, Cc) with dx = A*dy - Bc - x̂*Cc
__global__ void bn_bwd_finalize_kernel(const float2* __restrict__ partial, int S, const float* __restrict__ w,
                                       const float* __restrict__ rstd, float* __restrict__ dw, float* __restrict__ db,
                                       float* __restrict__ coef, int C, float count) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float sa = 0.f, sb = 0.f;
  for (int s = 0; s < S; ++s) {
    const float2 p = partial[(int64_t)c * S + s];
    sa += p.x;
    sb += p.y;
  }
  db[c] = sa;
  dw[c] = sb;
  // reference: w / (std * n) * (n * dy - dy_sum - x_norm * dy_x_norm_sum)
  const float g = w[c] * rstd[c] / count;
  coef[3 * c + 0] = g;
  coef[3 * c + 1] = sa;
  coef[3 * c + 2] = sb;
}

template <int VEC, bool RELU>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           const float* __restrict__ w, const float* __restrict__ b,
                                                           const float* __restrict__ coef, float* __restrict__ dx,
                                                           int64_t items, int C, int HWv, float count) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += stride) {
    const int c = (int)((it / HWv) % C);
    const float mu = __ldg(mean + c), rs = __ldg(rstd + c);
    const float g = __ldg(coef + 3 * c), s1 = __ldg(coef + 3 * c + 1), s2 = __ldg(coef + 3 * c + 2);
    float ww = 0.f, bb = 0.f;
    if (RELU) { ww = __ldg(w + c); bb = __ldg(b + c); }
    if (VEC == 4) {
      const float4 xv = ld_stream(reinterpret_cast<const float4*>(x) + it);
      float4 gv = ld_stream(reinterpret_cast<const float4*>(dy) + it);
      if (RELU) {
        gv.x = relu_masked(gv.x, xv.x, mu, rs, ww, bb); gv.y = relu_masked(gv.y, xv.y, mu, rs, ww, bb);
        gv.z = relu_masked(gv.z, xv.z, mu, rs, ww, bb); gv.w = relu_masked(gv.w, xv.w, mu, rs, ww, bb);
      }
      float4 o;
      o.x = g * (count * gv.x - s1 - (xv.x - mu) * rs * s2);
      o.y = g * (count * gv.y - s1 - (xv.y - mu) * rs * s2);
      o.z = g * (count * gv.z - s1 - (xv.z - mu) * rs * s2);
      o.w = g * (count * gv.w - s1 - (xv.w - mu) * rs * s2);
      st_stream(reinterpret_cast<float4*>(dx) + it, o);
    } else {
      const float gd = RELU ? relu_masked(dy[it], x[it], mu, rs, ww, bb) : dy[it];
      dx[it] = g * (count * gd - s1 - (x[it] - mu) * rs * s2);
    }
  }
}

// y = relu(bn(x) + skip): the tail of a residual block (BatchNorm2D -> `y += skip` -> ReLU; containers.py:153-157 followed by
// activations.py:114-120) in one pass — 12 1/8 B/elem instead of 8 + 12 + 8 1/8.  Same arithmetic in the same order as the three
// separate kernels (fmaf(w, (x-mu)*rstd, b); + skip; NaN-propagating max), so results are bit-identical.  Thread mapping and
// mask layout are those of relu_fwd_kernel (eltwise.cu): warp per 1024-element chunk, lane l owns the float4s at j*32 + l and
// mask word chunk*32 + l; tail in plain bit order — cpt_relu_bwd consumes the mask unchanged.  Requires HW % 4 == 0.
__global__ void __launch_bounds__(256) bn_add_relu_kernel(const float* __restrict__ x, const float* __restrict__ skip,
                                                          const float* __restrict__ w, const float* __restrict__ b,
                                                          const float* __restrict__ mean, const float* __restrict__ rstd,
                                                          float* __restrict__ y, uint32_t* __restrict__ mask, uint32_t n_chunks,
                                                          uint32_t n, uint32_t C, uint32_t HW) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t c = warp0; c < n_chunks; c += nwarps) {
    const float4* xp = reinterpret_cast<const float4*>(x) + (size_t)c * 256 + lane;
    const float4* sp = reinterpret_cast<const float4*>(skip) + (size_t)c * 256 + lane;
    float4* yp = reinterpret_cast<float4*>(y) + (size_t)c * 256 + lane;
    float4 v[8], k[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { v[j] = ld_stream(xp + j * 32); k[j] = ld_stream(sp + j * 32); }
    uint32_t bits = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t e0 = (c * 256 + lane + j * 32) * 4;  // first element of this float4 (one channel: HW % 4 == 0)
      const uint32_t ch = (e0 / HW) % C;
      const float mu = __ldg(mean + ch), rs = __ldg(rstd + ch), ww = __ldg(w + ch), bb = __ldg(b + ch);
      float e[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
      const float sk[4] = {k[j].x, k[j].y, k[j].z, k[j].w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float r = relu_fwd_val(fmaf(ww, (e[i] - mu) * rs, bb) + sk[i]);
        bits |= (r > 0.f ? 1u : 0u) << (4 * j + i);
        e[i] = r;
      }
      st_stream(yp + j * 32, make_float4(e[0], e[1], e[2], e[3]));
    }
    if (mask) mask[c * 32 + lane] = bits;
  }
  if (warp0 == 0) {
    const uint32_t t0 = n_chunks * 1024;
    for (uint32_t base = t0; base < n; base += 32) {
      const uint32_t i = base + lane;
      float r = 0.f;
      if (i < n) {
        const uint32_t ch = (i / HW) % C;
        r = relu_fwd_val(fmaf(__ldg(w + ch), (x[i] - __ldg(mean + ch)) * __ldg(rstd + ch), __ldg(b + ch)) + skip[i]);
        y[i] = r;
      }
      const uint32_t bl = __ballot_sync(0xffffffffu, i < n && r > 0.f);
      if (mask && lane == 0) mask[n_chunks * 32 + (base - t0) / 32] = bl;
    }
  }
}

// ---- BatchNorm2D -> ReLU -> MaxPooling2D(2) as one forward pass and one backward pass pair -------------------------------
// (normalizations.py:150-171, activations.py:114-120, pooling_funcs.py:71-82.)  The full-resolution activation
// a = relu(bn(x)) is never written: forward reads x and writes the pooled maximum; backward recomputes a for the four
// elements of a window from x, rebuilds the pooling tie mask (a == max, every tie receives the gradient) and the ReLU
// mask (bn(x) > 0), and feeds g = dy_pool * tie * relu into the usual two BatchNorm backward passes.  Same expressions
// and the same summation order as the three separate layers -> bit-identical results.  Requires H even, W % 4 == 0.
__device__ __forceinline__ float max_nan_bn(float a, float b) { return (a > b || a != a) ? a : b; }

struct BnPoolWin {  // one 2 x 4 patch: activations of both rows, the two window maxima
  float a0[4], a1[4], m[2];
};
__device__ __forceinline__ BnPoolWin bn_pool_window(const float4 x0, const float4 x1, float mu, float rs, float ww, float bb) {
  BnPoolWin r;
  const float r0[4] = {x0.x, x0.y, x0.z, x0.w}, r1[4] = {x1.x, x1.y, x1.z, x1.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    r.a0[i] = relu_fwd_val(fmaf(ww, (r0[i] - mu) * rs, bb));
    r.a1[i] = relu_fwd_val(fmaf(ww, (r1[i] - mu) * rs, bb));
  }
  r.m[0] = max_nan_bn(max_nan_bn(r.a0[0], r.a0[1]), max_nan_bn(r.a1[0], r.a1[1]));  // order of maxpool_fwd_kernel
  r.m[1] = max_nan_bn(max_nan_bn(r.a0[2], r.a0[3]), max_nan_bn(r.a1[2], r.a1[3]));
  return r;
}
// gradient w.r.t. bn(x) of element (row, i): dy_pool * (max == a) [pooling_funcs.py:81-82] * (bn(x) > 0) [activation_funcs.py:28,33]
__device__ __forceinline__ float bn_pool_grad(float g, float mx, float a) { return (g * (mx == a ? 1.0f : 0.0f)) * (a > 0.f ? 1.f : 0.f); }

// item = (image-channel bc, output row p, group of 4 input columns)
__global__ void __launch_bounds__(256) bn_relu_pool2_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                const float* __restrict__ b, const float* __restrict__ mean,
                                                                const float* __restrict__ rstd, float* __restrict__ y,
                                                                int64_t n_items, int C, int H, int W) {
  const int Wv = W >> 2, Ho = H >> 1, Wo = W >> 1;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < n_items; it += stride) {
    const int wv = (int)(it % Wv);
    const int64_t t = it / Wv;
    const int p = (int)(t % Ho);
    const int64_t bc = t / Ho;
    const int c = (int)(bc % C);
    const float* src = x + (bc * H + 2 * (int64_t)p) * W + 4 * wv;
    const float4 x0 = ld_stream(reinterpret_cast<const float4*>(src)), x1 = ld_stream(reinterpret_cast<const float4*>(src + W));
    const BnPoolWin r = bn_pool_window(x0, x1, __ldg(mean + c), __ldg(rstd + c), __ldg(w + c), __ldg(b + c));
    *reinterpret_cast<float2*>(y + (bc * Ho + p) * Wo + 2 * wv) = make_float2(r.m[0], r.m[1]);
  }
}

// Backward partial sums (Σg, Σg·x̂) per channel: the iteration space and accumulation order of bn_partial_kernel<2, 4> (items =
// float4s of the channel's planes, thread-strided, fixed split), so the sums are bit-identical to the unfused layers; the
// other row of each window comes through L1 (the same block touches it within the same step).
__global__ void __launch_bounds__(BN_THREADS) bn_pool2_partial_kernel(const float* __restrict__ x, const float* __restrict__ dyp,
                                                                       const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                       const float* __restrict__ w, const float* __restrict__ b,
                                                                       float2* __restrict__ partial, int N, int C, int H, int W) {
  __shared__ float sh[32];
  const int c = blockIdx.x, S = gridDim.y, s = blockIdx.y;
  const int HW = H * W, HWv = HW >> 2, Wv = W >> 2, Wo = W >> 1, Ho = H >> 1;
  const int64_t items = (int64_t)N * HWv;
  const int64_t per = (items + S - 1) / S;
  const int64_t lo = per * s, hi = (lo + per < items) ? lo + per : items;
  const float mu = __ldg(mean + c), rs = __ldg(rstd + c), ww = __ldg(w + c), bb = __ldg(b + c);
  float sa = 0.f, sb = 0.f;
  for (int64_t it = lo + threadIdx.x; it < hi; it += BN_THREADS) {
    const int64_t n = it / HWv;
    const int j = (int)(it - n * HWv), h = j / Wv, wv = j - h * Wv;
    const int64_t plane = (n * C + c) * (int64_t)HW;
    const float* rowp = x + plane + (int64_t)(h & ~1) * W + 4 * wv;
    const float4 x0 = ld_stream(reinterpret_cast<const float4*>(rowp)), x1 = ld_stream(reinterpret_cast<const float4*>(rowp + W));
    const float2 g = *reinterpret_cast<const float2*>(dyp + ((n * C + c) * (int64_t)Ho + (h >> 1)) * Wo + 2 * wv);
    const BnPoolWin r = bn_pool_window(x0, x1, mu, rs, ww, bb);
    const float* a = (h & 1) ? r.a1 : r.a0;
    const float xr[4] = {(h & 1) ? x1.x : x0.x, (h & 1) ? x1.y : x0.y, (h & 1) ? x1.z : x0.z, (h & 1) ? x1.w : x0.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float gg = bn_pool_grad(u < 2 ? g.x : g.y, r.m[u >> 1], a[u]);
      sa += gg;
      sb = fmaf(gg, (xr[u] - mu) * rs, sb);
    }
  }
  sa = block_sum(sa, sh);
  sb = block_sum(sb, sh);
  if (threadIdx.x == 0) partial[(int64_t)c * S + s] = make_float2(sa, sb);
}

// dx = coef0 * (count * g - coef1 - x̂ * coef2) for both rows of the window (expression of bn_bwd_apply_kernel)
__global__ void __launch_bounds__(256) bn_pool2_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dyp,
                                                                 const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                 const float* __restrict__ w, const float* __restrict__ b,
                                                                 const float* __restrict__ coef, float* __restrict__ dx,
                                                                 int64_t n_items, int C, int H, int W, float count) {
  const int Wv = W >> 2, Ho = H >> 1, Wo = W >> 1;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < n_items; it += stride) {
    const int wv = (int)(it % Wv);
    const int64_t t = it / Wv;
    const int p = (int)(t % Ho);
    const int64_t bc = t / Ho;
    const int c = (int)(bc % C);
    const int64_t off = (bc * H + 2 * (int64_t)p) * W + 4 * wv;
    const float4 x0 = ld_stream(reinterpret_cast<const float4*>(x + off)), x1 = ld_stream(reinterpret_cast<const float4*>(x + off + W));
    const float2 g = *reinterpret_cast<const float2*>(dyp + (bc * Ho + p) * Wo + 2 * wv);
    const float mu = __ldg(mean + c), rs = __ldg(rstd + c);
    const BnPoolWin r = bn_pool_window(x0, x1, mu, rs, __ldg(w + c), __ldg(b + c));
    const float k0 = __ldg(coef + 3 * c), s1 = __ldg(coef + 3 * c + 1), s2 = __ldg(coef + 3 * c + 2);
    const float r0[4] = {x0.x, x0.y, x0.z, x0.w}, r1[4] = {x1.x, x1.y, x1.z, x1.w};
    float o0[4], o1[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float gd = u < 2 ? g.x : g.y;
      o0[u] = k0 * (count * bn_pool_grad(gd, r.m[u >> 1], r.a0[u]) - s1 - (r0[u] - mu) * rs * s2);
      o1[u] = k0 * (count * bn_pool_grad(gd, r.m[u >> 1], r.a1[u]) - s1 - (r1[u] - mu) * rs * s2);
    }
    st_stream(reinterpret_cast<float4*>(dx + off), make_float4(o0[0], o0[1], o0[2], o0[3]));
    st_stream(reinterpret_cast<float4*>(dx + off + W), make_float4(o1[0], o1[1], o1[2], o1[3]));
  }
}

static int bn_splits(int N, int C, int HW) {
  // aim for >= 4 CTAs per SM in total, each with at least ~4K elements
  const int64_t per_channel = (int64_t)N * HW;
  int64_t want = ((int64_t)sm_count() * 4 + C - 1) / C;
  int64_t maxs = per_channel / 4096;
  if (maxs < 1) maxs = 1;
  if (want > maxs) want = maxs;
  if (want > 64) want = 64;
  if (want < 1) want = 1;
  return (int)want;
}

static int bn_check(const char* name, int N, int C, int HW) {
  CPT_REQUIRE(N > 0 && C > 0 && HW > 0, CPT_ERR_INVALID, "%s: non-positive dimension", name);
  return CPT_OK;
}

template <int MODE>
static void launch_partial(const float* x, const float* dy, const float* mean, const float* rstd, const float* w, const float* b,
                           float2* partial, int N, int C, int HW, int S, cudaStream_t st) {
  if (HW == 1) {
    dim3 grid((C + 31) / 32, S);
    bn_partial_hw1_kernel<MODE><<<grid, 256, 0, st>>>(x, dy, mean, rstd, w, b, partial, N, C);
  } else if (HW % 4 == 0 && aligned16(x) && (MODE == 0 || aligned16(dy))) {
    bn_partial_kernel<MODE, 4><<<dim3(C, S), BN_THREADS, 0, st>>>(x, dy, mean, rstd, w, b, partial, N, C, HW);
  } else {
    bn_partial_kernel<MODE, 1><<<dim3(C, S), BN_THREADS, 0, st>>>(x, dy, mean, rstd, w, b, partial, N, C, HW);
  }
}

static int check_act(const char* name, int act) {
  CPT_REQUIRE(act == CPT_ACT_NONE || act == CPT_ACT_RELU, CPT_ERR_INVALID, "%s: unknown activation %d", name, act);
  return CPT_OK;
}

static void launch_apply(const float* x, const float* w, const float* b, const float* mean, const float* rstd, float* y, int N, int C,
                         int HW, int act, cudaStream_t st) {
  const int64_t total = (int64_t)N * C * HW;
  const bool vec = HW % 4 == 0 && aligned16(x) && aligned16(y);
  if (vec) {
    if (act) bn_apply_kernel<4, true><<<ew_grid(total / 4, 256), 256, 0, st>>>(x, w, b, mean, rstd, y, total / 4, C, HW / 4);
    else bn_apply_kernel<4, false><<<ew_grid(total / 4, 256), 256, 0, st>>>(x, w, b, mean, rstd, y, total / 4, C, HW / 4);
  } else {
    if (act) bn_apply_kernel<1, true><<<ew_grid(total, 256), 256, 0, st>>>(x, w, b, mean, rstd, y, total, C, HW);
    else bn_apply_kernel<1, false><<<ew_grid(total, 256), 256, 0, st>>>(x, w, b, mean, rstd, y, total, C, HW);
  }
}

}  // namespace cpt

using namespace cpt;

extern "C" {

size_t cpt_bn_workspace_size(int N, int C, int HW) {
  if (N <= 0 || C <= 0 || HW <= 0) return 0;
  // partials [C][S] float2 + coef [C][3]
  return (size_t)C * 64 * sizeof(float2) + (size_t)C * 3 * sizeof(float) + 256;
}

int cpt_bn_act_fwd_train(const float* x, const float* w, const float* b, const float* rmean, const float* rvar, float* y,
                         float* rmean_out, float* rvar_out, float* save_mean, float* save_rstd, int N, int C, int HW,
                         float m, float eps, int act, void* ws, size_t ws_bytes, void* stream) {
  if (int e = bn_check("bn_fwd_train", N, C, HW)) return e;
  if (int e = check_act("bn_fwd_train", act)) return e;
  CPT_REQUIRE(ws && ws_bytes >= cpt_bn_workspace_size(N, C, HW), CPT_ERR_WORKSPACE, "bn_fwd_train: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int S = (HW == 1) ? (int)((N / 8 / 64 > 0) ? ((N / 8 / 64 > 64) ? 64 : N / 8 / 64) : 1) : bn_splits(N, C, HW);
  float2* partial = reinterpret_cast<float2*>(ws);
  launch_partial<0>(x, nullptr, nullptr, nullptr, nullptr, nullptr, partial, N, C, HW, S, st);
  CPT_LAUNCH_CHECK("bn_stats");
  const float count = (float)((int64_t)N * HW);
  bn_fwd_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(x, partial, S, rmean, rvar, rmean_out, rvar_out, save_mean,
                                                          save_rstd, C, HW, count, m, eps);
  CPT_LAUNCH_CHECK("bn_fwd_finalize");
  if (!y) return CPT_OK;  // statistics only: the caller applies them itself (cpt_bn_add_relu_apply)
  launch_apply(x, w, b, save_mean, save_rstd, y, N, C, HW, act, st);
  CPT_LAUNCH_CHECK("bn_apply");
  return CPT_OK;
}

size_t cpt_bn_cl_workspace_size(int N, int C, int HW) {
  if (N <= 0 || C <= 0 || HW <= 0) return 0;
  return cpt_bn_workspace_size(N, C, HW) + tc::bn_bwd_apply_cl_ws(N, C, HW) + 256;
}

int cpt_bn_act_fwd_train_cl(const float* x, const float* w, const float* b, const float* rmean, const float* rvar, float* y,
                            void* y_cl, float* rmean_out, float* rvar_out, float* save_mean, float* save_rstd, int N, int C,
                            int HW, float m, float eps, int act, void* ws, size_t ws_bytes, void* stream) {
  if (int e = bn_check("bn_fwd_train_cl", N, C, HW)) return e;
  if (int e = check_act("bn_fwd_train_cl", act)) return e;
  CPT_REQUIRE(y_cl, CPT_ERR_INVALID, "bn_fwd_train_cl: y_cl is NULL");
  CPT_REQUIRE(ws && ws_bytes >= cpt_bn_workspace_size(N, C, HW), CPT_ERR_WORKSPACE, "bn_fwd_train_cl: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int S = (HW == 1) ? (int)((N / 8 / 64 > 0) ? ((N / 8 / 64 > 64) ? 64 : N / 8 / 64) : 1) : bn_splits(N, C, HW);
  float2* partial = reinterpret_cast<float2*>(ws);
  launch_partial<0>(x, nullptr, nullptr, nullptr, nullptr, nullptr, partial, N, C, HW, S, st);
  CPT_LAUNCH_CHECK("bn_stats");
  const float count = (float)((int64_t)N * HW);
  bn_fwd_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(x, partial, S, rmean, rvar, rmean_out, rvar_out, save_mean,
                                                          save_rstd, C, HW, count, m, eps);
  CPT_LAUNCH_CHECK("bn_fwd_finalize");
  return tc::bn_apply_cl(x, w, b, save_mean, save_rstd, y, y_cl, N, C, HW, act, st);
}

int cpt_bn_act_fwd_train_presum(const float* x, const float* w, const float* b, const float* rmean, const float* rvar, float* y,
                                void* y_cl, float* rmean_out, float* rvar_out, float* save_mean, float* save_rstd, int N, int C,
                                int HW, float m, float eps, int act, const float* stats, int stat_slots, const float* conv_bias,
                                void* stream) {
  if (int e = bn_check("bn_fwd_train_presum", N, C, HW)) return e;
  if (int e = check_act("bn_fwd_train_presum", act)) return e;
  CPT_REQUIRE(stats && stat_slots > 0, CPT_ERR_INVALID, "bn_fwd_train_presum: no statistics");
  cudaStream_t st = as_stream(stream);
  const float count = (float)((int64_t)N * HW);
  bn_fwd_finalize_presum_kernel<<<(C + 31) / 32, 1024, 0, st>>>(stats, stat_slots, conv_bias, rmean, rvar, rmean_out, rvar_out,
                                                                save_mean, save_rstd, C, count, m, eps);
  CPT_LAUNCH_CHECK("bn_fwd_finalize_presum");
  if (!y) return CPT_OK;
  if (y_cl) return tc::bn_apply_cl(x, w, b, save_mean, save_rstd, y, y_cl, N, C, HW, act, st);
  launch_apply(x, w, b, save_mean, save_rstd, y, N, C, HW, act, st);
  CPT_LAUNCH_CHECK("bn_apply");
  return CPT_OK;
}

int cpt_bn_act_fwd_eval_cl(const float* x, const float* w, const float* b, const float* rmean, const float* rvar, float* y,
                           void* y_cl, float* save_mean, float* save_rstd, int N, int C, int HW, float eps, int act,
                           void* stream) {
  if (int e = bn_check("bn_fwd_eval_cl", N, C, HW)) return e;
  if (int e = check_act("bn_fwd_eval_cl", act)) return e;
  CPT_REQUIRE(y_cl, CPT_ERR_INVALID, "bn_fwd_eval_cl: y_cl is NULL");
  cudaStream_t st = as_stream(stream);
  bn_eval_stats_kernel<<<(C + 127) / 128, 128, 0, st>>>(rmean, rvar, save_mean, save_rstd, C, eps);
  CPT_LAUNCH_CHECK("bn_eval_stats");
  return tc::bn_apply_cl(x, w, b, save_mean, save_rstd, y, y_cl, N, C, HW, act, st);
}

int cpt_bn_act_bwd_cl(const float* x, const float* dy, const float* w, const float* b, const float* save_mean,
                      const float* save_rstd, float* dx, void* dx_cl, float* dx_chan_sum, float* dw, float* db, int N, int C,
                      int HW, int act, void* ws, size_t ws_bytes, void* stream) {
  if (int e = bn_check("bn_bwd_cl", N, C, HW)) return e;
  if (int e = check_act("bn_bwd_cl", act)) return e;
  CPT_REQUIRE(dx_cl, CPT_ERR_INVALID, "bn_bwd_cl: dx_cl is NULL");
  CPT_REQUIRE(act == CPT_ACT_NONE || b, CPT_ERR_INVALID, "bn_bwd_cl: the fused ReLU mask needs the bias");
  CPT_REQUIRE(ws && ws_bytes >= cpt_bn_cl_workspace_size(N, C, HW), CPT_ERR_WORKSPACE, "bn_bwd_cl: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int S = (HW == 1) ? (int)((N / 8 / 64 > 0) ? ((N / 8 / 64 > 64) ? 64 : N / 8 / 64) : 1) : bn_splits(N, C, HW);
  float2* partial = reinterpret_cast<float2*>(ws);
  float* coef = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + (size_t)C * 64 * sizeof(float2));
  char* ws2 = reinterpret_cast<char*>(ws) + ((cpt_bn_workspace_size(N, C, HW) + 255) / 256) * 256;
  if (act) launch_partial<2>(x, dy, save_mean, save_rstd, w, b, partial, N, C, HW, S, st);
  else launch_partial<1>(x, dy, save_mean, save_rstd, w, b, partial, N, C, HW, S, st);
  CPT_LAUNCH_CHECK("bn_bwd_partial");
  const float count = (float)((int64_t)N * HW);
  bn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(partial, S, w, save_rstd, dw, db, coef, C, count);
  CPT_LAUNCH_CHECK("bn_bwd_finalize");
  return tc::bn_bwd_apply_cl(x, dy, w, b, save_mean, save_rstd, coef, count, dx, dx_cl, dx_chan_sum, N, C, HW, act, ws2,
                             ws_bytes - (size_t)(ws2 - reinterpret_cast<char*>(ws)), st);
}

int cpt_bn_fwd_train(const float* x, const float* w, const float* b, const float* rmean, const float* rvar, float* y,
                     float* rmean_out, float* rvar_out, float* save_mean, float* save_rstd, int N, int C, int HW,
                     float m, float eps, void* ws, size_t ws_bytes, void* stream) {
  return cpt_bn_act_fwd_train(x, w, b, rmean, rvar, y, rmean_out, rvar_out, save_mean, save_rstd, N, C, HW, m, eps, CPT_ACT_NONE, ws,
                              ws_bytes, stream);
}

int cpt_bn_act_fwd_eval(const float* x, const float* w, const float* b, const float* rmean, const float* rvar, float* y,
                        float* save_mean, float* save_rstd, int N, int C, int HW, float eps, int act, void* stream) {
  if (int e = bn_check("bn_fwd_eval", N, C, HW)) return e;
  if (int e = check_act("bn_fwd_eval", act)) return e;
  cudaStream_t st = as_stream(stream);
  bn_eval_stats_kernel<<<(C + 127) / 128, 128, 0, st>>>(rmean, rvar, save_mean, save_rstd, C, eps);
  CPT_LAUNCH_CHECK("bn_eval_stats");
  if (!y) return CPT_OK;
  launch_apply(x, w, b, save_mean, save_rstd, y, N, C, HW, act, st);
  CPT_LAUNCH_CHECK("bn_apply");
  return CPT_OK;
}

int cpt_bn_fwd_eval(const float* x, const float* w, const float* b, const float* rmean, const float* rvar, float* y,
                    float* save_mean, float* save_rstd, int N, int C, int HW, float eps, void* stream) {
  return cpt_bn_act_fwd_eval(x, w, b, rmean, rvar, y, save_mean, save_rstd, N, C, HW, eps, CPT_ACT_NONE, stream);
}

int cpt_bn_act_bwd(const float* x, const float* dy, const float* w, const float* b, const float* save_mean, const float* save_rstd,
                   float* dx, float* dw, float* db, int N, int C, int HW, int act, void* ws, size_t ws_bytes, void* stream) {
  if (int e = bn_check("bn_bwd", N, C, HW)) return e;
  if (int e = check_act("bn_bwd", act)) return e;
  CPT_REQUIRE(act == CPT_ACT_NONE || b, CPT_ERR_INVALID, "bn_bwd: the fused ReLU mask needs the bias");
  CPT_REQUIRE(ws && ws_bytes >= cpt_bn_workspace_size(N, C, HW), CPT_ERR_WORKSPACE, "bn_bwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int S = (HW == 1) ? (int)((N / 8 / 64 > 0) ? ((N / 8 / 64 > 64) ? 64 : N / 8 / 64) : 1) : bn_splits(N, C, HW);
  float2* partial = reinterpret_cast<float2*>(ws);
  float* coef = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + (size_t)C * 64 * sizeof(float2));
  if (act) launch_partial<2>(x, dy, save_mean, save_rstd, w, b, partial, N, C, HW, S, st);
  else launch_partial<1>(x, dy, save_mean, save_rstd, w, b, partial, N, C, HW, S, st);
  CPT_LAUNCH_CHECK("bn_bwd_partial");
  const float count = (float)((int64_t)N * HW);
  bn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(partial, S, w, save_rstd, dw, db, coef, C, count);
  CPT_LAUNCH_CHECK("bn_bwd_finalize");
  const int64_t total = (int64_t)N * C * HW;
  if (HW % 4 == 0 && aligned16(x) && aligned16(dy) && aligned16(dx)) {
    if (act) bn_bwd_apply_kernel<4, true><<<ew_grid(total / 4, 256), 256, 0, st>>>(x, dy, save_mean, save_rstd, w, b, coef, dx, total / 4, C, HW / 4, count);
    else bn_bwd_apply_kernel<4, false><<<ew_grid(total / 4, 256), 256, 0, st>>>(x, dy, save_mean, save_rstd, w, b, coef, dx, total / 4, C, HW / 4, count);
  } else {
    if (act) bn_bwd_apply_kernel<1, true><<<ew_grid(total, 256), 256, 0, st>>>(x, dy, save_mean, save_rstd, w, b, coef, dx, total, C, HW, count);
    else bn_bwd_apply_kernel<1, false><<<ew_grid(total, 256), 256, 0, st>>>(x, dy, save_mean, save_rstd, w, b, coef, dx, total, C, HW, count);
  }
  CPT_LAUNCH_CHECK("bn_bwd_apply");
  return CPT_OK;
}

int cpt_bn_bwd(const float* x, const float* dy, const float* w, const float* save_mean, const float* save_rstd,
               float* dx, float* dw, float* db, int N, int C, int HW, void* ws, size_t ws_bytes, void* stream) {
  return cpt_bn_act_bwd(x, dy, w, nullptr, save_mean, save_rstd, dx, dw, db, N, C, HW, CPT_ACT_NONE, ws, ws_bytes, stream);
}

// ---- synchronised BatchNorm: the three-step forms of the passes above (the caller runs the collective in between) ----
static int bn_S(int N, int C, int HW) {
  return (HW == 1) ? (int)((N / 8 / 64 > 0) ? ((N / 8 / 64 > 64) ? 64 : N / 8 / 64) : 1) : bn_splits(N, C, HW);
}

int cpt_bn_local_stats(const float* x, float* stats, int N, int C, int HW, void* ws, size_t ws_bytes, void* stream) {
  if (int e = bn_check("bn_local_stats", N, C, HW)) return e;
  CPT_REQUIRE(x && stats, CPT_ERR_INVALID, "bn_local_stats: null pointer");
  CPT_REQUIRE(ws && ws_bytes >= cpt_bn_workspace_size(N, C, HW), CPT_ERR_WORKSPACE, "bn_local_stats: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int S = bn_S(N, C, HW);
  float2* partial = reinterpret_cast<float2*>(ws);
  launch_partial<0>(x, nullptr, nullptr, nullptr, nullptr, nullptr, partial, N, C, HW, S, st);
  CPT_LAUNCH_CHECK("bn_stats");
  bn_local_stats_kernel<<<(C + 127) / 128, 128, 0, st>>>(x, partial, S, stats, C, HW, (float)((int64_t)N * HW));
  CPT_LAUNCH_CHECK("bn_local_stats");
  return CPT_OK;
}

int cpt_bn_act_fwd_train_merged(const float* x, const float* w, const float* b, const float* rmean, const float* rvar, float* y,
                                void* y_cl, float* rmean_out, float* rvar_out, float* save_mean, float* save_rstd, int N, int C,
                                int HW, float m, float eps, int act, const float* gathered, int world, float* global_count,
                                void* stream) {
  if (int e = bn_check("bn_fwd_train_merged", N, C, HW)) return e;
  if (int e = check_act("bn_fwd_train_merged", act)) return e;
  CPT_REQUIRE(gathered && world > 0 && global_count, CPT_ERR_INVALID, "bn_fwd_train_merged: no gathered statistics");
  cudaStream_t st = as_stream(stream);
  bn_fwd_finalize_merged_kernel<<<(C + 127) / 128, 128, 0, st>>>(gathered, world, rmean, rvar, rmean_out, rvar_out, save_mean,
                                                                 save_rstd, global_count, C, m, eps);
  CPT_LAUNCH_CHECK("bn_fwd_finalize_merged");
  if (!y) return CPT_OK;
  if (y_cl) return tc::bn_apply_cl(x, w, b, save_mean, save_rstd, y, y_cl, N, C, HW, act, st);
  launch_apply(x, w, b, save_mean, save_rstd, y, N, C, HW, act, st);
  CPT_LAUNCH_CHECK("bn_apply");
  return CPT_OK;
}

int cpt_bn_act_bwd_local_sums(const float* x, const float* dy, const float* w, const float* b, const float* save_mean,
                              const float* save_rstd, float* sums, float* dw, float* db, int N, int C, int HW, int act, void* ws,
                              size_t ws_bytes, void* stream) {
  if (int e = bn_check("bn_bwd_local_sums", N, C, HW)) return e;
  if (int e = check_act("bn_bwd_local_sums", act)) return e;
  CPT_REQUIRE(act == CPT_ACT_NONE || b, CPT_ERR_INVALID, "bn_bwd_local_sums: the fused ReLU mask needs the bias");
  CPT_REQUIRE(ws && ws_bytes >= cpt_bn_workspace_size(N, C, HW), CPT_ERR_WORKSPACE, "bn_bwd_local_sums: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int S = bn_S(N, C, HW);
  float2* partial = reinterpret_cast<float2*>(ws);
  if (act) launch_partial<2>(x, dy, save_mean, save_rstd, w, b, partial, N, C, HW, S, st);
  else launch_partial<1>(x, dy, save_mean, save_rstd, w, b, partial, N, C, HW, S, st);
  CPT_LAUNCH_CHECK("bn_bwd_partial");
  bn_bwd_local_sums_kernel<<<(C + 127) / 128, 128, 0, st>>>(partial, S, sums, dw, db, C);
  CPT_LAUNCH_CHECK("bn_bwd_local_sums");
  return CPT_OK;
}

int cpt_bn_act_bwd_apply_global(const float* x, const float* dy, const float* w, const float* b, const float* save_mean,
                                const float* save_rstd, const float* sums, const float* global_count, float* dx, void* dx_cl,
                                float* dx_chan_sum, int N, int C, int HW, int act, void* ws, size_t ws_bytes, void* stream) {
  if (int e = bn_check("bn_bwd_apply_global", N, C, HW)) return e;
  if (int e = check_act("bn_bwd_apply_global", act)) return e;
  CPT_REQUIRE(sums && global_count, CPT_ERR_INVALID, "bn_bwd_apply_global: no sums / count");
  CPT_REQUIRE(ws && ws_bytes >= (dx_cl ? cpt_bn_cl_workspace_size(N, C, HW) : cpt_bn_workspace_size(N, C, HW)), CPT_ERR_WORKSPACE,
              "bn_bwd_apply_global: workspace too small");
  cudaStream_t st = as_stream(stream);
  float* coef = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + (size_t)C * 64 * sizeof(float2));
  bn_bwd_coef_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, w, save_rstd, coef, C, global_count);
  CPT_LAUNCH_CHECK("bn_bwd_coef");
  if (dx_cl) {
    char* ws2 = reinterpret_cast<char*>(ws) + ((cpt_bn_workspace_size(N, C, HW) + 255) / 256) * 256;
    return tc::bn_bwd_apply_cl(x, dy, w, b, save_mean, save_rstd, coef, 1.0f, dx, dx_cl, dx_chan_sum, N, C, HW, act, ws2,
                               ws_bytes - (size_t)(ws2 - reinterpret_cast<char*>(ws)), st);
  }
  const int64_t total = (int64_t)N * C * HW;
  if (HW % 4 == 0 && aligned16(x) && aligned16(dy) && aligned16(dx)) {
    if (act) bn_bwd_apply_kernel<4, true><<<ew_grid(total / 4, 256), 256, 0, st>>>(x, dy, save_mean, save_rstd, w, b, coef, dx, total / 4, C, HW / 4, 1.0f);
    else bn_bwd_apply_kernel<4, false><<<ew_grid(total / 4, 256), 256, 0, st>>>(x, dy, save_mean, save_rstd, w, b, coef, dx, total / 4, C, HW / 4, 1.0f);
  } else {
    if (act) bn_bwd_apply_kernel<1, true><<<ew_grid(total, 256), 256, 0, st>>>(x, dy, save_mean, save_rstd, w, b, coef, dx, total, C, HW, 1.0f);
    else bn_bwd_apply_kernel<1, false><<<ew_grid(total, 256), 256, 0, st>>>(x, dy, save_mean, save_rstd, w, b, coef, dx, total, C, HW, 1.0f);
  }
  CPT_LAUNCH_CHECK("bn_bwd_apply");
  return CPT_OK;
}

int cpt_bn_add_relu_apply(const float* x, const float* skip, const float* w, const float* b, const float* save_mean,
                          const float* save_rstd, float* y, uint8_t* mask, int N, int C, int HW, void* stream) {
  if (int e = bn_check("bn_add_relu_apply", N, C, HW)) return e;
  CPT_REQUIRE(x && skip && w && b && save_mean && save_rstd && y, CPT_ERR_INVALID, "bn_add_relu_apply: null pointer");
  const int64_t n = (int64_t)N * C * HW;
  CPT_REQUIRE(HW % 4 == 0 && n < (1LL << 31), CPT_ERR_UNSUPPORTED, "bn_add_relu_apply: needs H*W %% 4 == 0 and < 2^31 elements");
  CPT_REQUIRE(aligned16(x) && aligned16(skip) && aligned16(y) && (!mask || (reinterpret_cast<uintptr_t>(mask) & 3) == 0), CPT_ERR_INVALID,
              "bn_add_relu_apply: x, skip, y must be 16-byte aligned and mask 4-byte aligned");
  const uint32_t n_chunks = (uint32_t)(n / 1024);
  bn_add_relu_kernel<<<ew_grid((int64_t)(n_chunks > 0 ? n_chunks : 1) * 32, 256), 256, 0, as_stream(stream)>>>(
      x, skip, w, b, save_mean, save_rstd, y, reinterpret_cast<uint32_t*>(mask), n_chunks, (uint32_t)n, (uint32_t)C, (uint32_t)HW);
  CPT_LAUNCH_CHECK("bn_add_relu_apply");
  return CPT_OK;
}

int cpt_bn_relu_pool2_fwd(const float* x, const float* w, const float* b, const float* save_mean, const float* save_rstd, float* y,
                          int N, int C, int H, int W, void* stream) {
  if (int e = bn_check("bn_relu_pool2_fwd", N, C, H * W)) return e;
  CPT_REQUIRE(x && w && b && save_mean && save_rstd && y, CPT_ERR_INVALID, "bn_relu_pool2_fwd: null pointer");
  CPT_REQUIRE(H % 2 == 0 && W % 4 == 0, CPT_ERR_UNSUPPORTED, "bn_relu_pool2_fwd: needs H even and W %% 4 == 0");
  CPT_REQUIRE(aligned16(x) && (reinterpret_cast<uintptr_t>(y) & 7) == 0, CPT_ERR_INVALID, "bn_relu_pool2_fwd: x must be 16-byte, y 8-byte aligned");
  const int64_t items = (int64_t)N * C * (H / 2) * (W / 4);
  bn_relu_pool2_fwd_kernel<<<ew_grid(items, 256), 256, 0, as_stream(stream)>>>(x, w, b, save_mean, save_rstd, y, items, C, H, W);
  CPT_LAUNCH_CHECK("bn_relu_pool2_fwd");
  return CPT_OK;
}

int cpt_bn_relu_pool2_bwd(const float* x, const float* dy_pool, const float* w, const float* b, const float* save_mean,
                          const float* save_rstd, float* dx, float* dw, float* db, int N, int C, int H, int W, void* ws,
                          size_t ws_bytes, void* stream) {
  const int HW = H * W;
  if (int e = bn_check("bn_relu_pool2_bwd", N, C, HW)) return e;
  CPT_REQUIRE(x && dy_pool && w && b && save_mean && save_rstd && dx && dw && db, CPT_ERR_INVALID, "bn_relu_pool2_bwd: null pointer");
  CPT_REQUIRE(H % 2 == 0 && W % 4 == 0, CPT_ERR_UNSUPPORTED, "bn_relu_pool2_bwd: needs H even and W %% 4 == 0");
  CPT_REQUIRE(aligned16(x) && aligned16(dx) && (reinterpret_cast<uintptr_t>(dy_pool) & 7) == 0, CPT_ERR_INVALID,
              "bn_relu_pool2_bwd: x, dx must be 16-byte, dy_pool 8-byte aligned");
  CPT_REQUIRE(ws && ws_bytes >= cpt_bn_workspace_size(N, C, HW), CPT_ERR_WORKSPACE, "bn_relu_pool2_bwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int S = bn_splits(N, C, HW);  // the split of bn_partial_kernel: same partial boundaries, same sums
  float2* partial = reinterpret_cast<float2*>(ws);
  float* coef = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + (size_t)C * 64 * sizeof(float2));
  bn_pool2_partial_kernel<<<dim3(C, S), BN_THREADS, 0, st>>>(x, dy_pool, save_mean, save_rstd, w, b, partial, N, C, H, W);
  CPT_LAUNCH_CHECK("bn_pool2_partial");
  const float count = (float)((int64_t)N * HW);
  bn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(partial, S, w, save_rstd, dw, db, coef, C, count);
  CPT_LAUNCH_CHECK("bn_bwd_finalize");
  const int64_t items = (int64_t)N * C * (H / 2) * (W / 4);
  bn_pool2_bwd_apply_kernel<<<ew_grid(items, 256), 256, 0, st>>>(x, dy_pool, save_mean, save_rstd, w, b, coef, dx, items, C, H, W, count);
  CPT_LAUNCH_CHECK("bn_pool2_bwd_apply");
  return CPT_OK;
}

}  // extern "C"
