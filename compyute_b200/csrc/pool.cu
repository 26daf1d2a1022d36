// MaxPooling2D / AvgPooling2D forward + backward (stride = kernel, no padding, floor) — HBM-bound.
// Reference: compyute/nn/functional/pooling_funcs.py:67-121.
//   fwd bytes: 4N read + 4N/k² write;  bwd (max): 4N (x) + 4N (dx) + 8N/k² (y, dy re-read through L1/L2).
#include "common.cuh"

namespace cpt {

// numpy max semantics: NaN in the window -> NaN
__device__ __forceinline__ float max_nan(float a, float b) { return (a > b || a != a) ? a : b; }

// one thread per output element; k = 2 fast path uses 64-bit loads (a warp reads 256 contiguous bytes per row)
template <int KS>
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n_out,
                                                          int H, int W, int Ho, int Wo, int k, int vec2) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n_out; o += stride) {
    const int q = (int)(o % Wo);
    const int64_t t = o / Wo;
    const int p = (int)(t % Ho);
    const int64_t bc = t / Ho;
    const float* src = x + (bc * H + (int64_t)p * k) * W + (int64_t)q * k;
    float m;
    if (KS == 2 && vec2) {
      float2 r0 = *reinterpret_cast<const float2*>(src);
      float2 r1 = *reinterpret_cast<const float2*>(src + W);
      m = max_nan(max_nan(r0.x, r0.y), max_nan(r1.x, r1.y));
    } else {
      m = src[0];
      for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) m = max_nan(m, src[(int64_t)i * W + j]);
    }
    y[o] = m;
  }
}

// one thread per 4 consecutive input columns (float4) when W % 4 == 0, else per element
template <int VEC>
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                          const float* __restrict__ dy, float* __restrict__ dx,
                                                          int64_t n_items, int H, int W, int Ho, int Wo, int k) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int Wv = W / VEC;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < n_items; it += stride) {
    const int wv = (int)(it % Wv);
    const int64_t t = it / Wv;
    const int h = (int)(t % H);
    const int64_t bc = t / H;
    const int p = h / k;
    float xv[VEC], out[VEC];
    const int64_t base = (bc * H + h) * W + (int64_t)wv * VEC;
    if (VEC == 4) {
      float4 v = ld_stream(reinterpret_cast<const float4*>(x + base));
      xv[0] = v.x; xv[1] = v.y; xv[2] = v.z; xv[3] = v.w;
    } else {
      xv[0] = x[base];
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int w = wv * VEC + j;
      const int q = w / k;
      if (p < Ho && q < Wo) {
        const int64_t oi = (bc * Ho + p) * Wo + q;
        // reference: upsample(dy) * (upsample(y) == x): bool mask multiplies dy (dy*0 keeps the sign of dy)
        out[j] = __ldg(dy + oi) * ((__ldg(y + oi) == xv[j]) ? 1.0f : 0.0f);
      } else {
        out[j] = 0.0f;  // zero-padded tail of the upsampled dy
      }
    }
    if (VEC == 4) st_stream(reinterpret_cast<float4*>(dx + base), make_float4(out[0], out[1], out[2], out[3]));
    else dx[base] = out[0];
  }
}

// k = 2 fast paths (every MaxPooling2D of the BASELINE configs): W % 4 == 0, H even, 16-byte aligned rows.
// Work item = (image-channel bc, output row p, group of 4 input columns): two 128-bit loads of x (rows 2p and 2p+1) serve two
// outputs; two items per thread and loop iteration keep four independent 128-bit loads in flight.  Same comparison order as
// the generic kernel (bit-exact, NaN in the window -> NaN).
struct Pool2Item {
  int64_t row0;  // element offset of x[bc][2p][4*wv]
  int64_t out;   // element offset of y[bc][p][2*wv]
};
__device__ __forceinline__ Pool2Item pool2_item(int64_t it, int Wv, int Ho, int H, int W, int Wo) {
  int64_t t;
  int wv, p;
  if (it < 0x7fffffff) {
    const uint32_t u = (uint32_t)it, t32 = u / (uint32_t)Wv;
    wv = (int)(u - t32 * (uint32_t)Wv);
    const uint32_t bc = t32 / (uint32_t)Ho;
    p = (int)(t32 - bc * (uint32_t)Ho);
    t = bc;
  } else {
    const int64_t t64 = it / Wv;
    wv = (int)(it - t64 * Wv);
    t = t64 / Ho;
    p = (int)(t64 - t * Ho);
  }
  Pool2Item r;
  r.row0 = (t * H + 2 * (int64_t)p) * W + 4 * (int64_t)wv;
  r.out = (t * Ho + p) * Wo + 2 * (int64_t)wv;
  return r;
}

__global__ void __launch_bounds__(256) maxpool2_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n_items,
                                                           int H, int W, int Ho, int Wo) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int Wv = W / 4;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < n_items; it += 2 * stride) {
    const bool two = it + stride < n_items;
    const Pool2Item a = pool2_item(it, Wv, Ho, H, W, Wo), b = pool2_item(two ? it + stride : it, Wv, Ho, H, W, Wo);
    const float4 a0 = ld_stream(reinterpret_cast<const float4*>(x + a.row0)), a1 = ld_stream(reinterpret_cast<const float4*>(x + a.row0 + W));
    const float4 b0 = ld_stream(reinterpret_cast<const float4*>(x + b.row0)), b1 = ld_stream(reinterpret_cast<const float4*>(x + b.row0 + W));
    *reinterpret_cast<float2*>(y + a.out) = make_float2(max_nan(max_nan(a0.x, a0.y), max_nan(a1.x, a1.y)),
                                                        max_nan(max_nan(a0.z, a0.w), max_nan(a1.z, a1.w)));
    if (two)
      *reinterpret_cast<float2*>(y + b.out) = make_float2(max_nan(max_nan(b0.x, b0.y), max_nan(b1.x, b1.y)),
                                                          max_nan(max_nan(b0.z, b0.w), max_nan(b1.z, b1.w)));
  }
}

__device__ __forceinline__ float4 pool2_mask(float4 xv, float2 yv, float2 g) {
  // reference: upsample(dy) * (upsample(y) == x): the bool mask multiplies dy (dy * 0 keeps the sign of dy)
  return make_float4(g.x * (yv.x == xv.x ? 1.0f : 0.0f), g.x * (yv.x == xv.y ? 1.0f : 0.0f), g.y * (yv.y == xv.z ? 1.0f : 0.0f),
                     g.y * (yv.y == xv.w ? 1.0f : 0.0f));
}

__global__ void __launch_bounds__(256) maxpool2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                           const float* __restrict__ dy, float* __restrict__ dx, int64_t n_items,
                                                           int H, int W, int Ho, int Wo) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int Wv = W / 4;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < n_items; it += stride) {
    const Pool2Item a = pool2_item(it, Wv, Ho, H, W, Wo);
    const float4 x0 = ld_stream(reinterpret_cast<const float4*>(x + a.row0)), x1 = ld_stream(reinterpret_cast<const float4*>(x + a.row0 + W));
    const float2 yv = *reinterpret_cast<const float2*>(y + a.out), g = *reinterpret_cast<const float2*>(dy + a.out);
    st_stream(reinterpret_cast<float4*>(dx + a.row0), pool2_mask(x0, yv, g));
    st_stream(reinterpret_cast<float4*>(dx + a.row0 + W), pool2_mask(x1, yv, g));
  }
}

__global__ void __launch_bounds__(256) avgpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n_out,
                                                          int H, int W, int Ho, int Wo, int k) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float inv = 1.0f / (float)(k * k);
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n_out; o += stride) {
    const int q = (int)(o % Wo);
    const int64_t t = o / Wo;
    const int p = (int)(t % Ho);
    const int64_t bc = t / Ho;
    const float* src = x + (bc * H + (int64_t)p * k) * W + (int64_t)q * k;
    float s = 0.f;
    for (int i = 0; i < k; ++i)
      for (int j = 0; j < k; ++j) s += src[(int64_t)i * W + j];
    y[o] = s * inv;
  }
}

// Global average (Ho = Wo = 1, the 7x7 head of the ResNet-shaped config): a warp per (image, channel) plane reads its H*W
// contiguous floats coalesced and reduces them with shuffles — the generic kernel's thread-per-output reads each plane with
// one lane (44 % of HBM peak, round 1).  Covers the k x k window at the top-left like the generic form (tail rows / columns
// outside the window do not contribute, pooling_funcs.py:111-121).
__global__ void __launch_bounds__(256) avgpool_global_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t planes,
                                                                 int H, int W, int k) {
  const int lane = threadIdx.x & 31;
  const int64_t wstride = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float inv = 1.0f / (float)(k * k);
  const bool full = (k == H && k == W);
  for (int64_t pl = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); pl < planes; pl += wstride) {
    const float* src = x + pl * H * W;
    float s = 0.f;
    if (full) {
      for (int i = lane; i < H * W; i += 32) s += src[i];
    } else {
      for (int i = lane; i < k * k; i += 32) s += src[(i / k) * W + (i % k)];
    }
    s = warp_sum(s);
    if (lane == 0) y[pl] = s * inv;
  }
}

// k = 2, W % 4 == 0: a thread makes two outputs from two 128-bit loads (rows 2p and 2p+1), like the max-pooling fast path
__global__ void __launch_bounds__(256) avgpool2_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n_pairs,
                                                           int H, int W, int Ho, int Wo) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int wp = W >> 2;  // output pairs per row
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n_pairs; o += stride) {
    const int q2 = (int)(o % wp);
    const int64_t t = o / wp;
    const int p = (int)(t % Ho);
    const int64_t bc = t / Ho;
    const float4* r0 = reinterpret_cast<const float4*>(x + (bc * H + 2 * (int64_t)p) * W) + q2;
    const float4 a = ld_stream(r0), b = ld_stream(r0 + wp);
    float2 out;
    out.x = ((a.x + a.y) + (b.x + b.y)) * 0.25f;
    out.y = ((a.z + a.w) + (b.z + b.w)) * 0.25f;
    *reinterpret_cast<float2*>(y + (bc * Ho + p) * Wo + 2 * q2) = out;
  }
}

__global__ void __launch_bounds__(256) avgpool_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int64_t n_in,
                                                          int H, int W, int Ho, int Wo, int k) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float div = (float)(k * k);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_in; i += stride) {
    const int w = (int)(i % W);
    const int64_t t = i / W;
    const int h = (int)(t % H);
    const int64_t bc = t / H;
    const int p = h / k, q = w / k;
    dx[i] = (p < Ho && q < Wo) ? __ldg(dy + (bc * Ho + p) * Wo + q) / div : 0.0f;
  }
}

static int check_pool(const char* name, int B, int C, int H, int W, int k) {
  CPT_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && k > 0, CPT_ERR_INVALID, "%s: non-positive dimension", name);
  CPT_REQUIRE(H >= k && W >= k, CPT_ERR_INVALID, "%s: kernel %d larger than input %dx%d", name, k, H, W);
  return CPT_OK;
}

}  // namespace cpt

using namespace cpt;

extern "C" {

int cpt_maxpool2d_fwd(const float* x, float* y, int B, int C, int H, int W, int k, void* stream) {
  if (int e = check_pool("maxpool2d_fwd", B, C, H, W, k)) return e;
  const int Ho = H / k, Wo = W / k;
  const int64_t n_out = (int64_t)B * C * Ho * Wo;
  const int grid = ew_grid(n_out, 256);
  if (k == 2 && W % 4 == 0 && H % 2 == 0 && aligned16(x) && (reinterpret_cast<uintptr_t>(y) & 7) == 0) {
    const int64_t items = (int64_t)B * C * Ho * (W / 4);
    maxpool2_fwd_kernel<<<ew_grid((items + 1) / 2, 256), 256, 0, as_stream(stream)>>>(x, y, items, H, W, Ho, Wo);
  } else if (k == 2) {
    const int vec2 = (W % 2 == 0) && ((reinterpret_cast<uintptr_t>(x) & 7) == 0);
    maxpool_fwd_kernel<2><<<grid, 256, 0, as_stream(stream)>>>(x, y, n_out, H, W, Ho, Wo, k, vec2);
  } else {
    maxpool_fwd_kernel<0><<<grid, 256, 0, as_stream(stream)>>>(x, y, n_out, H, W, Ho, Wo, k, 0);
  }
  CPT_LAUNCH_CHECK("maxpool2d_fwd");
  return CPT_OK;
}

int cpt_maxpool2d_bwd(const float* x, const float* y, const float* dy, float* dx, int B, int C, int H, int W, int k,
                      void* stream) {
  if (int e = check_pool("maxpool2d_bwd", B, C, H, W, k)) return e;
  const int Ho = H / k, Wo = W / k;
  if (k == 2 && W % 4 == 0 && H % 2 == 0 && aligned16(x) && aligned16(dx) && ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(dy)) & 7) == 0) {
    const int64_t items = (int64_t)B * C * Ho * (W / 4);
    maxpool2_bwd_kernel<<<ew_grid(items, 256), 256, 0, as_stream(stream)>>>(x, y, dy, dx, items, H, W, Ho, Wo);
  } else if (W % 4 == 0 && aligned16(x) && aligned16(dx)) {
    const int64_t items = (int64_t)B * C * H * (W / 4);
    maxpool_bwd_kernel<4><<<ew_grid(items, 256), 256, 0, as_stream(stream)>>>(x, y, dy, dx, items, H, W, Ho, Wo, k);
  } else {
    const int64_t items = (int64_t)B * C * H * W;
    maxpool_bwd_kernel<1><<<ew_grid(items, 256), 256, 0, as_stream(stream)>>>(x, y, dy, dx, items, H, W, Ho, Wo, k);
  }
  CPT_LAUNCH_CHECK("maxpool2d_bwd");
  return CPT_OK;
}

int cpt_avgpool2d_fwd(const float* x, float* y, int B, int C, int H, int W, int k, void* stream) {
  if (int e = check_pool("avgpool2d_fwd", B, C, H, W, k)) return e;
  const int Ho = H / k, Wo = W / k;
  const int64_t n_out = (int64_t)B * C * Ho * Wo;
  if (Ho == 1 && Wo == 1 && k * k >= 16) {
    const int64_t planes = (int64_t)B * C;
    avgpool_global_fwd_kernel<<<ew_grid(planes * 32, 256), 256, 0, as_stream(stream)>>>(x, y, planes, H, W, k);
  } else if (k == 2 && W % 4 == 0 && aligned16(x) && (reinterpret_cast<uintptr_t>(y) & 7) == 0) {
    const int64_t n_pairs = n_out / 2;
    avgpool2_fwd_kernel<<<ew_grid(n_pairs, 256), 256, 0, as_stream(stream)>>>(x, y, n_pairs, H, W, Ho, Wo);
  } else {
    avgpool_fwd_kernel<<<ew_grid(n_out, 256), 256, 0, as_stream(stream)>>>(x, y, n_out, H, W, Ho, Wo, k);
  }
  CPT_LAUNCH_CHECK("avgpool2d_fwd");
  return CPT_OK;
}

int cpt_avgpool2d_bwd(const float* dy, float* dx, int B, int C, int H, int W, int k, void* stream) {
  if (int e = check_pool("avgpool2d_bwd", B, C, H, W, k)) return e;
  const int Ho = H / k, Wo = W / k;
  const int64_t n_in = (int64_t)B * C * H * W;
  avgpool_bwd_kernel<<<ew_grid(n_in, 256), 256, 0, as_stream(stream)>>>(dy, dx, n_in, H, W, Ho, Wo, k);
  CPT_LAUNCH_CHECK("avgpool2d_bwd");
  return CPT_OK;
}

}  // extern "C"
