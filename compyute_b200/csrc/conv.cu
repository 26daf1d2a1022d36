// Conv2D forward / dgrad / wgrad(+db): C-ABI entry points and the exact-fp32 (FFMA) problem definitions.
// Reference: compyute/nn/functional/convolution_funcs.py:218-410; direct-form equations in SURVEY Appendix D.
// Tensor-core modes (TF32/BF16) are dispatched to tc_conv.cu.
#include "simt_gemm.cuh"
#include "tc.cuh"

namespace cpt {

struct ConvGeom {
  int B, Ci, H, W, Co, K, P, S, D, Ho, Wo;
};

static int make_geom(const cpt_conv2d_desc* d, ConvGeom& g, const char* who) {
  CPT_REQUIRE(d != nullptr, CPT_ERR_INVALID, "%s: null descriptor", who);
  CPT_REQUIRE(d->B > 0 && d->Ci > 0 && d->H > 0 && d->W > 0 && d->Co > 0 && d->K > 0, CPT_ERR_INVALID,
              "%s: non-positive dimension", who);
  CPT_REQUIRE(d->pad >= 0 && d->stride >= 1 && d->dil >= 1, CPT_ERR_INVALID, "%s: bad padding/stride/dilation", who);
  g = {d->B, d->Ci, d->H, d->W, d->Co, d->K, d->pad, d->stride, d->dil, 0, 0};
  const int keff = d->dil * (d->K - 1) + 1;
  CPT_REQUIRE(d->H + 2 * d->pad >= keff && d->W + 2 * d->pad >= keff, CPT_ERR_INVALID,
              "%s: dilated kernel (%d) larger than padded input (%dx%d)", who, keff, d->H + 2 * d->pad, d->W + 2 * d->pad);
  g.Ho = (d->H + 2 * d->pad - keff) / d->stride + 1;
  g.Wo = (d->W + 2 * d->pad - keff) / d->stride + 1;
  CPT_REQUIRE((int64_t)d->B * g.Ho * g.Wo < (1LL << 31) && (int64_t)d->B * d->H * d->W < (1LL << 31) &&
                  (int64_t)d->Ci * d->K * d->K < (1LL << 31) && (int64_t)d->Co * d->K * d->K < (1LL << 31),
              CPT_ERR_UNSUPPORTED, "%s: GEMM extent exceeds int32", who);
  return CPT_OK;
}

// ---- fprop: M = B*Ho*Wo (pixels), N = Co, K = Ci*K*K
struct ConvFpropP {
  const float* x; const float* w; const float* bias; float* y;
  ConvGeom g; int M, N, K, KK, HoWo, vec;
  static constexpr bool A_MN_CONTIG = true, B_MN_CONTIG = false, OUT_M_CONTIG = true;
  struct RowA { const float* base; int ih0, iw0; };
  struct ColB { const float* base; };
  __device__ RowA rowA(int m) const {
    RowA r;
    if (m >= M) { r.base = nullptr; r.ih0 = r.iw0 = 0; return r; }
    const int b = m / HoWo, rem = m - b * HoWo, p = rem / g.Wo, q = rem - p * g.Wo;
    r.base = x + (int64_t)b * g.Ci * g.H * g.W;
    r.ih0 = p * g.S - g.P;
    r.iw0 = q * g.S - g.P;
    return r;
  }
  __device__ float loadA(const RowA& r, int k) const {
    if (!r.base) return 0.f;
    const int ci = k / KK, t = k - ci * KK, j = t / g.K, kk = t - j * g.K;
    const int ih = r.ih0 + j * g.D, iw = r.iw0 + kk * g.D;
    if ((unsigned)ih >= (unsigned)g.H || (unsigned)iw >= (unsigned)g.W) return 0.f;
    return __ldg(r.base + ((int64_t)ci * g.H + ih) * g.W + iw);
  }
  __device__ ColB colB(int n) const { return ColB{n < N ? w + (int64_t)n * K : nullptr}; }
  __device__ float loadB(const ColB& c, int k) const { return c.base ? __ldg(c.base + k) : 0.f; }
  __device__ void store4(int, int m, int n, const float v[4]) const {
    if (n >= N || m >= M) return;
    const float bv = bias ? __ldg(bias + n) : 0.f;
    const int b = m / HoWo, rem = m - b * HoWo;
    float* dst = y + ((int64_t)b * N + n) * HoWo + rem;
    if (vec && (HoWo & 3) == 0 && m + 3 < M) {  // m % 4 == 0 and HoWo % 4 == 0: the 4 pixels share an image, 16B aligned
      *reinterpret_cast<float4*>(dst) = make_float4(v[0] + bv, v[1] + bv, v[2] + bv, v[3] + bv);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int mm = m + i;
        if (mm < M) {
          const int b2 = mm / HoWo, r2 = mm - b2 * HoWo;
          y[((int64_t)b2 * N + n) * HoWo + r2] = v[i] + bv;
        }
      }
    }
  }
};

// ---- dgrad, decomposed by stride residue class (rh, rw): the input pixels h = rh + S*hs, w = rw + S*ws form a sub-grid
// on which dgrad is a dense stride-1 correlation over only the taps with (h + P - j*D) % S == 0 — no multiplications by
// the zeros of the stride-dilated dy (RawConv2DFn.backward :391 materialises them).  S = 1 has one class = all taps.
//   M = B*Hs*Ws (sub-grid pixels), N = Ci, K = Co * ntaps
struct ConvDgradP {
  const float* dy; const float* w; float* dx;
  ConvGeom g; int M, N, K, KK, Hs, Ws, rh, rw, ntaps, vec;
  unsigned char tj[64], tk[64];  // taps of this class
  static constexpr bool A_MN_CONTIG = true, B_MN_CONTIG = false, OUT_M_CONTIG = true;
  struct RowA { const float* base; int u, v; };
  struct ColB { const float* base; };
  __device__ RowA rowA(int m) const {
    RowA r;
    if (m >= M) { r.base = nullptr; r.u = r.v = 0; return r; }
    const int hw = Hs * Ws, b = m / hw, rem = m - b * hw, hs = rem / Ws, ws = rem - hs * Ws;
    r.base = dy + (int64_t)b * g.Co * g.Ho * g.Wo;
    r.u = rh + g.S * hs + g.P;
    r.v = rw + g.S * ws + g.P;
    return r;
  }
  __device__ float loadA(const RowA& r, int k) const {
    if (!r.base) return 0.f;
    const int co = k / ntaps, t = k - co * ntaps;
    const int nu = r.u - (int)tj[t] * g.D, nv = r.v - (int)tk[t] * g.D;
    if (nu < 0 || nv < 0) return 0.f;
    const int p = nu / g.S, q = nv / g.S;  // exact: the class guarantees divisibility
    if (p >= g.Ho || q >= g.Wo) return 0.f;
    return __ldg(r.base + ((int64_t)co * g.Ho + p) * g.Wo + q);
  }
  __device__ ColB colB(int n) const { return ColB{n < N ? w + (int64_t)n * KK : nullptr}; }
  __device__ float loadB(const ColB& c, int k) const {
    if (!c.base) return 0.f;
    const int co = k / ntaps, t = k - co * ntaps;
    return __ldg(c.base + (int64_t)co * N * KK + (int)tj[t] * g.K + (int)tk[t]);  // w[co][ci][j][kk]
  }
  __device__ void store4(int, int m, int n, const float v[4]) const {
    if (n >= N || m >= M) return;
    if (g.S == 1 && vec && ((g.H * g.W) & 3) == 0 && m + 3 < M) {  // dense case: 4 consecutive pixels of one image
      const int HW = g.H * g.W, b = m / HW, rem = m - b * HW;
      *reinterpret_cast<float4*>(dx + ((int64_t)b * N + n) * HW + rem) = make_float4(v[0], v[1], v[2], v[3]);
      return;
    }
    const int hw = Hs * Ws;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int mm = m + i;
      if (mm < M) {
        const int b = mm / hw, rem = mm - b * hw, hs = rem / Ws, ws = rem - hs * Ws;
        dx[(((int64_t)b * N + n) * g.H + rh + g.S * hs) * g.W + rw + g.S * ws] = v[i];
      }
    }
  }
};

// ---- dgrad for tiny Ci (e.g. the 3-channel stem): a GEMM formulation would waste a 128-wide N tile on <= 8 columns.
// One thread per input pixel of one residue class, CI accumulators in registers, filter in shared memory as
// [co][tap][ci]; lanes run along ws so dy reads are coalesced and the tap set is warp-uniform.
struct SmallTaps { unsigned char tj[64], tk[64]; };  // passed by value: nothing to upload, CUDA-graph safe

template <int CI>
__global__ void __launch_bounds__(256) dgrad_small_ci_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                             float* __restrict__ dx, ConvGeom g, int Hs, int Ws, int rh, int rw,
                                                             int ntaps, const SmallTaps taps) {
  extern __shared__ float wsm[];  // [Co][ntaps][CI]
  __shared__ int s_tj[64], s_tk[64];
  for (int i = threadIdx.x; i < ntaps; i += blockDim.x) { s_tj[i] = taps.tj[i]; s_tk[i] = taps.tk[i]; }
  __syncthreads();
  const int KK = g.K * g.K;
  for (int i = threadIdx.x; i < g.Co * ntaps * CI; i += blockDim.x) {
    const int ci = i % CI, t = (i / CI) % ntaps, co = i / (CI * ntaps);
    wsm[i] = ci < g.Ci ? w[((int64_t)co * g.Ci + ci) * KK + s_tj[t] * g.K + s_tk[t]] : 0.f;
  }
  __syncthreads();
  const int64_t total = (int64_t)g.B * Hs * Ws;
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < total; m += (int64_t)gridDim.x * blockDim.x) {
    const int ws = (int)(m % Ws);
    const int64_t r = m / Ws;
    const int hs = (int)(r % Hs), b = (int)(r / Hs);
    const int u = rh + g.S * hs + g.P, v = rw + g.S * ws + g.P;
    float acc[CI];
#pragma unroll
    for (int c = 0; c < CI; ++c) acc[c] = 0.f;
    const float* src = dy + (int64_t)b * g.Co * g.Ho * g.Wo;
    for (int t = 0; t < ntaps; ++t) {
      const int nu = u - s_tj[t] * g.D, nv = v - s_tk[t] * g.D;
      if (nu < 0 || nv < 0) continue;
      const int p = nu / g.S, q = nv / g.S;
      if (p >= g.Ho || q >= g.Wo) continue;
      const float* sp = src + (int64_t)p * g.Wo + q;
      const float* wp = wsm + t * CI;
#pragma unroll 4
      for (int co = 0; co < g.Co; ++co) {
        const float d = __ldg(sp + (int64_t)co * g.Ho * g.Wo);
#pragma unroll
        for (int c = 0; c < CI; ++c) acc[c] = fmaf(d, wp[(co * ntaps) * CI + c], acc[c]);
      }
    }
    const int h = rh + g.S * hs, wv = rw + g.S * ws;
#pragma unroll
    for (int c = 0; c < CI; ++c)
      if (c < g.Ci) dx[(((int64_t)b * g.Ci + c) * g.H + h) * g.W + wv] = acc[c];
  }
}

// taps (j, kk) contributing to residue class (rh, rw); returns the count
static int class_taps(const ConvGeom& g, int rh, int rw, unsigned char* tj, unsigned char* tk) {
  int n = 0;
  for (int j = 0; j < g.K; ++j) {
    if (((rh + g.P - j * g.D) % g.S + g.S) % g.S != 0) continue;
    for (int kk = 0; kk < g.K; ++kk) {
      if (((rw + g.P - kk * g.D) % g.S + g.S) % g.S != 0) continue;
      if (n < 64) { tj[n] = (unsigned char)j; tk[n] = (unsigned char)kk; }
      ++n;
    }
  }
  return n;
}

// exact-fp32 dgrad driver (all strides): one launch per non-empty residue class
static int dgrad_fp32(const ConvGeom& g, const float* dy, const float* w, float* dx, void* /*ws*/, size_t /*ws_bytes*/, cudaStream_t st) {
  const int classes = g.S * g.S;
  bool any_empty = false;
  for (int c = 0; c < classes && !any_empty; ++c) {
    unsigned char a[64], b2[64];
    const int rh = c / g.S, rw = c % g.S;
    if (rh < g.H && rw < g.W && class_taps(g, rh, rw, a, b2) == 0) any_empty = true;
  }
  if (any_empty)  // input pixels no output window touches (e.g. 1x1 stride 2) get dx = 0
    CPT_CUDA(cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)g.B * g.Ci * g.H * g.W, st));
  for (int c = 0; c < classes; ++c) {
    const int rh = c / g.S, rw = c % g.S;
    if (rh >= g.H || rw >= g.W) continue;
    ConvDgradP p;
    p.ntaps = class_taps(g, rh, rw, p.tj, p.tk);
    if (p.ntaps == 0) continue;
    CPT_REQUIRE(p.ntaps <= 64, CPT_ERR_UNSUPPORTED, "conv2d_dgrad: more than 64 taps per stride class (K=%d)", g.K);
    const int Hs = (g.H - rh + g.S - 1) / g.S, Ws = (g.W - rw + g.S - 1) / g.S;
    const size_t wbytes = (size_t)g.Co * p.ntaps * (g.Ci <= 4 ? 4 : 8) * sizeof(float);
    if (g.Ci <= 8 && wbytes <= 96 * 1024) {
      SmallTaps dtaps;
      for (int i = 0; i < 64; ++i) { dtaps.tj[i] = p.tj[i]; dtaps.tk[i] = p.tk[i]; }
      const int64_t total = (int64_t)g.B * Hs * Ws;
      const int grid = ew_grid(total, 256);
      if (g.Ci <= 4) {
        static bool cfg4 = false;
        if (!cfg4) { CPT_CUDA(cudaFuncSetAttribute(dgrad_small_ci_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); cfg4 = true; }
        dgrad_small_ci_kernel<4><<<grid, 256, wbytes, st>>>(dy, w, dx, g, Hs, Ws, rh, rw, p.ntaps, dtaps);
      } else {
        static bool cfg8 = false;
        if (!cfg8) { CPT_CUDA(cudaFuncSetAttribute(dgrad_small_ci_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); cfg8 = true; }
        dgrad_small_ci_kernel<8><<<grid, 256, wbytes, st>>>(dy, w, dx, g, Hs, Ws, rh, rw, p.ntaps, dtaps);
      }
      CPT_LAUNCH_CHECK("dgrad_small_ci");
      continue;
    }
    p.dy = dy; p.w = w; p.dx = dx; p.g = g; p.vec = aligned16(dx);
    p.KK = g.K * g.K; p.Hs = Hs; p.Ws = Ws; p.rh = rh; p.rw = rw;
    p.M = g.B * Hs * Ws; p.N = g.Ci; p.K = g.Co * p.ntaps;
    CPT_CUDA(sg_launch(p, 1, st));
  }
  return CPT_OK;
}

// ---- wgrad: M = Co, N = Ci*K*K, K(reduction) = B*Ho*Wo; split-K partials [split][M][N]
struct ConvWgradP {
  const float* x; const float* dy; float* out;  // out: partial buffer or dw itself when splits == 1
  ConvGeom g; int M, N, K, KK, HoWo;
  static constexpr bool A_MN_CONTIG = false, B_MN_CONTIG = false, OUT_M_CONTIG = false;
  struct RowA { const float* base; };
  struct ColB { const float* base; int oh, ow; };
  __device__ RowA rowA(int m) const { return RowA{m < M ? dy + (int64_t)m * HoWo : nullptr}; }
  __device__ float loadA(const RowA& r, int k) const {
    if (!r.base) return 0.f;
    const int b = k / HoWo, rem = k - b * HoWo;
    return __ldg(r.base + (int64_t)b * M * HoWo + rem);
  }
  __device__ ColB colB(int n) const {
    ColB c;
    if (n >= N) { c.base = nullptr; c.oh = c.ow = 0; return c; }
    const int ci = n / KK, t = n - ci * KK, j = t / g.K, kk = t - j * g.K;
    c.base = x + (int64_t)ci * g.H * g.W;
    c.oh = j * g.D - g.P;
    c.ow = kk * g.D - g.P;
    return c;
  }
  __device__ float loadB(const ColB& c, int k) const {
    if (!c.base) return 0.f;
    const int b = k / HoWo, rem = k - b * HoWo, p = rem / g.Wo, q = rem - p * g.Wo;
    const int ih = p * g.S + c.oh, iw = q * g.S + c.ow;
    if ((unsigned)ih >= (unsigned)g.H || (unsigned)iw >= (unsigned)g.W) return 0.f;
    return __ldg(c.base + ((int64_t)b * g.Ci * g.H + ih) * g.W + iw);
  }
  __device__ void store4(int split, int m, int n, const float v[4]) const {
    if (m >= M) return;
    float* dst = out + ((int64_t)split * M + m) * N + n;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (n + i < N) dst[i] = v[i];
  }
};

__global__ void __launch_bounds__(256) reduce_splits_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                                            int64_t n, int splits) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += partial[(int64_t)k * n + i];
    out[i] = s;
  }
}

void launch_reduce_splits(const float* partial, float* out, int64_t n, int splits, cudaStream_t st) {
  reduce_splits_kernel<<<ew_grid(n, 256), 256, 0, st>>>(partial, out, n, splits);
}

// ---- per-channel sum (db): partial (grid C x S) + fixed-order finalize
__global__ void __launch_bounds__(256) chan_sum_partial_kernel(const float* __restrict__ x, float* __restrict__ partial, int N,
                                                               int C, int HW) {
  __shared__ float sh[32];
  const int c = blockIdx.x, S = gridDim.y, s = blockIdx.y;
  const int64_t items = (int64_t)N * HW;
  const int64_t per = (items + S - 1) / S;
  const int64_t lo = per * s, hi = (lo + per < items) ? lo + per : items;
  float acc = 0.f;
  for (int64_t it = lo + threadIdx.x; it < hi; it += 256) {
    const int64_t n = it / HW;
    const int j = (int)(it - n * HW);
    acc += x[(n * C + c) * (int64_t)HW + j];
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) partial[(int64_t)c * S + s] = acc;
}
// HW == 1: x is (N, C); lanes across channels
__global__ void __launch_bounds__(256) col_sum_partial_kernel(const float* __restrict__ x, float* __restrict__ partial, int N,
                                                              int C) {
  __shared__ float sm[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx, S = gridDim.y, s = blockIdx.y;
  float acc = 0.f;
  if (c < C)
    for (int n = s * 8 + ty; n < N; n += S * 8) acc += x[(int64_t)n * C + c];
  sm[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && c < C) {
#pragma unroll
    for (int r = 1; r < 8; ++r) acc += sm[r][tx];
    partial[(int64_t)c * S + s] = acc;
  }
}
__global__ void chan_sum_final_kernel(const float* __restrict__ partial, float* __restrict__ out, int C, int S) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int k = 0; k < S; ++k) s += partial[(int64_t)c * S + k];
  out[c] = s;
}

int channel_sum(const float* x, float* out, int N, int C, int HW, void* ws, cudaStream_t st) {
  float* partial = reinterpret_cast<float*>(ws);
  int S;
  if (HW == 1) {
    S = N / 512; if (S > 64) S = 64; if (S < 1) S = 1;
    col_sum_partial_kernel<<<dim3((C + 31) / 32, S), 256, 0, st>>>(x, partial, N, C);
  } else {
    int64_t want = ((int64_t)sm_count() * 4 + C - 1) / C, maxs = ((int64_t)N * HW) / 4096;
    if (maxs < 1) maxs = 1;
    if (want > maxs) want = maxs;
    if (want > 64) want = 64;
    S = (int)want;
    chan_sum_partial_kernel<<<dim3(C, S), 256, 0, st>>>(x, partial, N, C, HW);
  }
  CPT_LAUNCH_CHECK("channel_sum_partial");
  chan_sum_final_kernel<<<(C + 127) / 128, 128, 0, st>>>(partial, out, C, S);
  CPT_LAUNCH_CHECK("channel_sum_final");
  return CPT_OK;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static size_t chan_sum_ws(int C) { return align_up((size_t)C * 64 * sizeof(float), 256); }

}  // namespace cpt

using namespace cpt;

extern "C" {

int cpt_conv2d_out_shape(const cpt_conv2d_desc* d, int* Ho, int* Wo) {
  ConvGeom g;
  if (int e = make_geom(d, g, "conv2d_out_shape")) return e;
  if (Ho) *Ho = g.Ho;
  if (Wo) *Wo = g.Wo;
  return CPT_OK;
}

size_t cpt_conv2d_workspace_size(int op, const cpt_conv2d_desc* d, int mode) {
  ConvGeom g;
  if (make_geom(d, g, "conv2d_workspace_size")) return 0;
  if (mode != CPT_MODE_FP32) return tc::conv_workspace_size(op, d, mode);
  if (op == CPT_OP_WGRAD) {
    const int M = g.Co, N = g.Ci * g.K * g.K, K = g.B * g.Ho * g.Wo;
    const int splits = sg_pick_splits(M, N, K);
    return align_up((size_t)(splits > 1 ? splits : 0) * M * N * sizeof(float), 256) + chan_sum_ws(g.Co) + 256;
  }
  if (op == CPT_OP_DGRAD) return (size_t)g.S * g.S * 128 + 256;  // tap tables of the small-Ci kernel
  return 256;
}

int cpt_conv2d_fprop(const cpt_conv2d_desc* d, const float* x, const float* w, const float* bias, float* y, int mode,
                     void* ws, size_t ws_bytes, void* stream) {
  ConvGeom g;
  if (int e = make_geom(d, g, "conv2d_fprop")) return e;
  CPT_REQUIRE(x && w && y, CPT_ERR_INVALID, "conv2d_fprop: null tensor");
  if (mode != CPT_MODE_FP32) return tc::conv_fprop(d, x, w, bias, y, mode, ws, ws_bytes, as_stream(stream));
  ConvFpropP p;
  p.x = x; p.w = w; p.bias = bias; p.y = y; p.g = g; p.vec = aligned16(y);
  p.KK = g.K * g.K; p.HoWo = g.Ho * g.Wo;
  p.M = g.B * p.HoWo; p.N = g.Co; p.K = g.Ci * p.KK;
  CPT_CUDA(sg_launch(p, 1, as_stream(stream)));
  return CPT_OK;
}

int cpt_conv2d_dgrad(const cpt_conv2d_desc* d, const float* dy, const float* w, float* dx, int mode, void* ws,
                     size_t ws_bytes, void* stream) {
  ConvGeom g;
  if (int e = make_geom(d, g, "conv2d_dgrad")) return e;
  CPT_REQUIRE(dy && w && dx, CPT_ERR_INVALID, "conv2d_dgrad: null tensor");
  if (mode != CPT_MODE_FP32) return tc::conv_dgrad(d, dy, w, dx, mode, ws, ws_bytes, as_stream(stream));
  return dgrad_fp32(g, dy, w, dx, ws, ws_bytes, as_stream(stream));
}

int cpt_conv2d_wgrad(const cpt_conv2d_desc* d, const float* x, const float* dy, float* dw, float* db, int mode,
                     void* ws, size_t ws_bytes, void* stream) {
  ConvGeom g;
  if (int e = make_geom(d, g, "conv2d_wgrad")) return e;
  CPT_REQUIRE(x && dy && dw, CPT_ERR_INVALID, "conv2d_wgrad: null tensor");
  if (mode != CPT_MODE_FP32) return tc::conv_wgrad(d, x, dy, dw, db, mode, ws, ws_bytes, as_stream(stream));
  CPT_REQUIRE(ws && ws_bytes >= cpt_conv2d_workspace_size(CPT_OP_WGRAD, d, mode), CPT_ERR_WORKSPACE,
              "conv2d_wgrad: workspace too small (%zu < %zu)", ws_bytes, cpt_conv2d_workspace_size(CPT_OP_WGRAD, d, mode));
  cudaStream_t st = as_stream(stream);
  ConvWgradP p;
  p.x = x; p.dy = dy; p.g = g;
  p.KK = g.K * g.K; p.HoWo = g.Ho * g.Wo;
  p.M = g.Co; p.N = g.Ci * p.KK; p.K = g.B * p.HoWo;
  const int splits = sg_pick_splits(p.M, p.N, p.K);
  const size_t part_bytes = align_up((size_t)(splits > 1 ? splits : 0) * p.M * p.N * sizeof(float), 256);
  p.out = splits > 1 ? reinterpret_cast<float*>(ws) : dw;
  CPT_CUDA(sg_launch(p, splits, st));
  if (splits > 1) {
    launch_reduce_splits(reinterpret_cast<float*>(ws), dw, (int64_t)p.M * p.N, splits, st);
    CPT_LAUNCH_CHECK("conv2d_wgrad reduce");
  }
  if (db) return channel_sum(dy, db, g.B, g.Co, p.HoWo, reinterpret_cast<char*>(ws) + part_bytes, st);
  return CPT_OK;
}

}  // extern "C"
