// Generic device-tensor operator set: what the reference reaches through `Tensor.data <op>` and
// `device.module.<numpy fn>` (compyute/tensors.py:196-292, 552-682; tensor_ops/*.py), as memory-bound kernels.
//   * elementwise binary / compare with NumPy broadcasting (strides of 0 on broadcast dims, <= CPT_MAX_DIMS merged dims),
//     128-bit fast path for same-shape and tensor-scalar operands;
//   * elementwise unary;
//   * reductions over any subset of axes (row form: block per output slice; column form: lanes along the contiguous kept
//     dim), split over the grid with a fixed-order second pass — deterministic, no atomics;
//   * strided copy (slicing, permute, flip, concat, pad, broadcast_to, __setitem__), row gather (batch indexing), casts,
//     arange, counter-based uniform / normal.
// Roofline bound: HBM for all of them.
#include <math.h>

#include <type_traits>

#include "common.cuh"

namespace cpt {

struct Dims {
  int nd;
  int64_t d[CPT_MAX_DIMS];
  int64_t sa[CPT_MAX_DIMS];
  int64_t sb[CPT_MAX_DIMS];
};

// ------------------------------------------------------------------------------------------------ operators
template <int OP>
__device__ __forceinline__ float bin(float a, float b) {
  if constexpr (OP == CPT_EW_ADD) return a + b;
  if constexpr (OP == CPT_EW_SUB) return a - b;
  if constexpr (OP == CPT_EW_MUL) return a * b;
  if constexpr (OP == CPT_EW_DIV) return a / b;
  if constexpr (OP == CPT_EW_POW) return powf(a, b);
  if constexpr (OP == CPT_EW_MAX) return (a != a || b != b) ? (a != a ? a : b) : fmaxf(a, b);  // numpy.maximum: NaN wins
  if constexpr (OP == CPT_EW_MIN) return (a != a || b != b) ? (a != a ? a : b) : fminf(a, b);
  if constexpr (OP == CPT_EW_FLOORDIV) return floorf(a / b);
  if constexpr (OP == CPT_EW_MOD) {  // numpy: result has the sign of the divisor
    float r = fmodf(a, b);
    return (r != 0.f && ((r < 0.f) != (b < 0.f))) ? r + b : r;
  }
  if constexpr (OP == CPT_EW_LT) return a < b ? 1.f : 0.f;
  if constexpr (OP == CPT_EW_GT) return a > b ? 1.f : 0.f;
  if constexpr (OP == CPT_EW_LE) return a <= b ? 1.f : 0.f;
  if constexpr (OP == CPT_EW_GE) return a >= b ? 1.f : 0.f;
  if constexpr (OP == CPT_EW_EQ) return a == b ? 1.f : 0.f;
  if constexpr (OP == CPT_EW_NE) return a != b ? 1.f : 0.f;
  return 0.f;
}
constexpr bool is_cmp(int op) { return op >= CPT_EW_LT && op <= CPT_EW_NE; }

template <int OP>
__device__ __forceinline__ float una(float a, float p0, float p1) {
  if constexpr (OP == CPT_UN_NEG) return -a;
  if constexpr (OP == CPT_UN_ABS) return fabsf(a);
  if constexpr (OP == CPT_UN_EXP) return expf(a);
  if constexpr (OP == CPT_UN_LOG) return logf(a);
  if constexpr (OP == CPT_UN_LOG2) return log2f(a);
  if constexpr (OP == CPT_UN_LOG10) return log10f(a);
  if constexpr (OP == CPT_UN_SQRT) return sqrtf(a);
  if constexpr (OP == CPT_UN_TANH) return tanhf(a);
  if constexpr (OP == CPT_UN_SIN) return sinf(a);
  if constexpr (OP == CPT_UN_COS) return cosf(a);
  if constexpr (OP == CPT_UN_TAN) return tanf(a);
  if constexpr (OP == CPT_UN_SINH) return sinhf(a);
  if constexpr (OP == CPT_UN_COSH) return coshf(a);
  if constexpr (OP == CPT_UN_CLIP) return (a != a) ? a : fminf(fmaxf(a, p0), p1);  // numpy.clip keeps NaN
  if constexpr (OP == CPT_UN_ISNAN) return a != a ? 1.f : 0.f;
  if constexpr (OP == CPT_UN_ROUND) return rintf(a * p0) / p0;  // numpy.round(x, d): half-to-even at scale 10^d (p0)
  if constexpr (OP == CPT_UN_SQUARE) return a * a;
  if constexpr (OP == CPT_UN_RECIP) return 1.f / a;
  return a;
}
constexpr bool un_is_bool(int op) { return op == CPT_UN_ISNAN; }

template <typename F>
static bool dispatch_bin(int op, F&& f) {
  switch (op) {
#define C(X) case X: f(std::integral_constant<int, X>{}); return true;
    C(CPT_EW_ADD) C(CPT_EW_SUB) C(CPT_EW_MUL) C(CPT_EW_DIV) C(CPT_EW_POW) C(CPT_EW_MAX) C(CPT_EW_MIN) C(CPT_EW_FLOORDIV)
    C(CPT_EW_MOD) C(CPT_EW_LT) C(CPT_EW_GT) C(CPT_EW_LE) C(CPT_EW_GE) C(CPT_EW_EQ) C(CPT_EW_NE)
#undef C
  }
  return false;
}
template <typename F>
static bool dispatch_un(int op, F&& f) {
  switch (op) {
#define C(X) case X: f(std::integral_constant<int, X>{}); return true;
    C(CPT_UN_NEG) C(CPT_UN_ABS) C(CPT_UN_EXP) C(CPT_UN_LOG) C(CPT_UN_LOG2) C(CPT_UN_LOG10) C(CPT_UN_SQRT) C(CPT_UN_TANH)
    C(CPT_UN_SIN) C(CPT_UN_COS) C(CPT_UN_TAN) C(CPT_UN_SINH) C(CPT_UN_COSH) C(CPT_UN_CLIP) C(CPT_UN_ISNAN) C(CPT_UN_ROUND)
    C(CPT_UN_SQUARE) C(CPT_UN_RECIP)
#undef C
  }
  return false;
}

__device__ __forceinline__ void store_out(float* o, int64_t i, float v) { o[i] = v; }
__device__ __forceinline__ void store_out(uint8_t* o, int64_t i, float v) { o[i] = v != 0.f; }
__device__ __forceinline__ void store4(float* o, int64_t i4, float4 v) { st_stream(reinterpret_cast<float4*>(o) + i4, v); }
__device__ __forceinline__ void store4(uint8_t* o, int64_t i4, float4 v) {
  reinterpret_cast<uchar4*>(o)[i4] = make_uchar4(v.x != 0.f, v.y != 0.f, v.z != 0.f, v.w != 0.f);
}

// MODE 0: a[i] op b[i]   1: a[i] op s   2: s op a[i]      (same shape, contiguous)
template <int OP, int MODE, typename OutT>
__global__ void __launch_bounds__(256) ew_bin_flat_kernel(OutT* __restrict__ out, const float* __restrict__ a,
                                                          const float* __restrict__ b, float s, int64_t n4, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = t; i < n4; i += stride) {
    const float4 u = ld_stream(reinterpret_cast<const float4*>(a) + i);
    float4 v = make_float4(s, s, s, s);
    if constexpr (MODE == 0) v = ld_stream(reinterpret_cast<const float4*>(b) + i);
    float4 r;
    if constexpr (MODE == 2) r = make_float4(bin<OP>(v.x, u.x), bin<OP>(v.y, u.y), bin<OP>(v.z, u.z), bin<OP>(v.w, u.w));
    else r = make_float4(bin<OP>(u.x, v.x), bin<OP>(u.y, v.y), bin<OP>(u.z, v.z), bin<OP>(u.w, v.w));
    store4(out, i, r);
  }
  for (int64_t i = n4 * 4 + t; i < n; i += stride) {
    const float u = a[i], v = MODE == 0 ? b[i] : s;
    store_out(out, i, MODE == 2 ? bin<OP>(v, u) : bin<OP>(u, v));
  }
}

// general broadcasting: flat output index -> per-operand offsets (dims merged on the host; innermost dim last)
template <int OP, typename OutT>
__global__ void __launch_bounds__(256) ew_bin_bcast_kernel(OutT* __restrict__ out, const float* __restrict__ a,
                                                           const float* __restrict__ b, Dims D, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int64_t r = i, oa = 0, ob = 0;
#pragma unroll
    for (int k = CPT_MAX_DIMS - 1; k >= 0; --k) {
      if (k < D.nd) {
        int64_t q, c;
        if (r < 0x7fffffff && D.d[k] < 0x7fffffff) { q = (uint32_t)r / (uint32_t)D.d[k]; c = (uint32_t)r - (uint32_t)q * (uint32_t)D.d[k]; }
        else { q = r / D.d[k]; c = r - q * D.d[k]; }
        oa += c * D.sa[k]; ob += c * D.sb[k]; r = q;
      }
    }
    store_out(out, i, bin<OP>(a[oa], b[ob]));
  }
}

// Row form of the broadcast kernel: the innermost merged dim (length L) is contiguous in the output and walked with stride
// 0 or 1 by each operand.  Work item = (row of the outer dims, ROW_SEG-element segment of that row), one warp per item: the
// outer index is decomposed once per item, a broadcast operand is loaded once per item, the rest streams — with 128-bit
// accesses when every row start is 16-byte aligned (VEC).
constexpr int ROW_SEG = 2048;
struct RowItem {
  int64_t row, oa, ob, e0, e1;
};
__device__ __forceinline__ RowItem row_item(int64_t w, int segs, const Dims& D, int64_t L) {
  RowItem it;
  it.row = (w < 0x7fffffff) ? (int64_t)((uint32_t)w / (uint32_t)segs) : w / segs;
  const int seg = (int)(w - it.row * segs);
  int64_t r = it.row;
  it.oa = it.ob = 0;
#pragma unroll
  for (int k = CPT_MAX_DIMS - 1; k >= 0; --k) {
    if (k < D.nd) {
      int64_t q, c;
      if (r < 0x7fffffff && D.d[k] < 0x7fffffff) { q = (uint32_t)r / (uint32_t)D.d[k]; c = (uint32_t)r - (uint32_t)q * (uint32_t)D.d[k]; }
      else { q = r / D.d[k]; c = r - q * D.d[k]; }
      it.oa += c * D.sa[k]; it.ob += c * D.sb[k]; r = q;
    }
  }
  it.e0 = (int64_t)seg * ROW_SEG;
  it.e1 = it.e0 + ROW_SEG < L ? it.e0 + ROW_SEG : L;
  return it;
}

template <int OP, typename OutT, bool VEC>
__global__ void __launch_bounds__(256) ew_bin_rows_kernel(OutT* __restrict__ out, const float* __restrict__ a,
                                                          const float* __restrict__ b, Dims D, int64_t L, int sa_in, int sb_in,
                                                          int64_t items, int segs) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t w = w0; w < items; w += nw) {
    const RowItem it = row_item(w, segs, D, L);
    const float* ap = a + it.oa;
    const float* bp = b + it.ob;
    OutT* op = out + it.row * L;
    const float a0 = sa_in ? 0.f : __ldg(ap), b0 = sb_in ? 0.f : __ldg(bp);
    if constexpr (VEC) {
      for (int64_t i = it.e0 + lane * 4; i < it.e1; i += 128) {
        const float4 u = sa_in ? ld_stream(reinterpret_cast<const float4*>(ap + i)) : make_float4(a0, a0, a0, a0);
        const float4 v = sb_in ? ld_stream(reinterpret_cast<const float4*>(bp + i)) : make_float4(b0, b0, b0, b0);
        store4(op, i >> 2, make_float4(bin<OP>(u.x, v.x), bin<OP>(u.y, v.y), bin<OP>(u.z, v.z), bin<OP>(u.w, v.w)));
      }
    } else {
      for (int64_t i = it.e0 + lane; i < it.e1; i += 32) store_out(op, i, bin<OP>(sa_in ? ap[i] : a0, sb_in ? bp[i] : b0));
    }
  }
}

template <int OP, typename OutT>
__global__ void __launch_bounds__(256) ew_un_kernel(OutT* __restrict__ out, const float* __restrict__ a, float p0, float p1,
                                                    int64_t n4, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = t; i < n4; i += stride) {
    const float4 u = ld_stream(reinterpret_cast<const float4*>(a) + i);
    store4(out, i, make_float4(una<OP>(u.x, p0, p1), una<OP>(u.y, p0, p1), una<OP>(u.z, p0, p1), una<OP>(u.w, p0, p1)));
  }
  for (int64_t i = n4 * 4 + t; i < n; i += stride) store_out(out, i, una<OP>(a[i], p0, p1));
}

// boolean (uint8 0/1) logic: 0 and, 1 or, 2 xor, 3 not(a)
__global__ void __launch_bounds__(256) logic_kernel(uint8_t* __restrict__ out, const uint8_t* __restrict__ a,
                                                    const uint8_t* __restrict__ b, int op, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const bool x = a[i] != 0, y = b ? b[i] != 0 : false;
    out[i] = op == 0 ? (x && y) : op == 1 ? (x || y) : op == 2 ? (x != y) : !x;
  }
}

// ------------------------------------------------------------------------------------------------ reductions
// Accumulator: value (+ flat position for argmax).  Input element type T is float or uint8 (bool).
struct Acc {
  float v;
  long long i;  // argmax: position along the reduced index; count: integer sum
};

template <int OP>
__device__ __forceinline__ Acc red_init() {
  Acc a;
  a.i = OP == CPT_RED_ARGMAX ? 0x7fffffffffffffffll : 0;
  a.v = (OP == CPT_RED_MAX || OP == CPT_RED_ARGMAX) ? -INFINITY : OP == CPT_RED_MIN ? INFINITY
        : (OP == CPT_RED_PROD || OP == CPT_RED_ALL) ? 1.f : 0.f;
  return a;
}
__device__ __forceinline__ bool gt_nan(float a, float b) { return (a != a) ? (b == b) : a > b; }  // NaN is the largest
template <int OP>
__device__ __forceinline__ void red_step(Acc& a, float x, long long pos) {
  if constexpr (OP == CPT_RED_SUM) a.v += x;
  if constexpr (OP == CPT_RED_SUMSQ) a.v = fmaf(x, x, a.v);
  if constexpr (OP == CPT_RED_PROD) a.v *= x;
  if constexpr (OP == CPT_RED_MAX) a.v = (a.v != a.v || x != x) ? NAN : fmaxf(a.v, x);
  if constexpr (OP == CPT_RED_MIN) a.v = (a.v != a.v || x != x) ? NAN : fminf(a.v, x);
  if constexpr (OP == CPT_RED_ANY) a.v = (a.v != 0.f || x != 0.f) ? 1.f : 0.f;
  if constexpr (OP == CPT_RED_ALL) a.v = (a.v != 0.f && x != 0.f) ? 1.f : 0.f;
  if constexpr (OP == CPT_RED_COUNT) a.i += (x != 0.f);
  if constexpr (OP == CPT_RED_ARGMAX) {
    if (gt_nan(x, a.v) || (a.i == 0x7fffffffffffffffll) || (!gt_nan(a.v, x) && pos < a.i)) { a.v = x; a.i = pos; }
  }
}
template <int OP>
__device__ __forceinline__ void red_merge(Acc& a, const Acc& b) {
  if constexpr (OP == CPT_RED_SUM || OP == CPT_RED_SUMSQ) a.v += b.v;
  else if constexpr (OP == CPT_RED_COUNT) a.i += b.i;
  else if constexpr (OP == CPT_RED_ARGMAX) {
    if (b.i != 0x7fffffffffffffffll && (a.i == 0x7fffffffffffffffll || gt_nan(b.v, a.v) || (!gt_nan(a.v, b.v) && b.i < a.i))) a = b;
  } else red_step<OP>(a, b.v, 0);
}
template <int OP>
__device__ __forceinline__ Acc warp_red(Acc a) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Acc b;
    b.v = __shfl_xor_sync(0xffffffffu, a.v, o);
    b.i = __shfl_xor_sync(0xffffffffu, a.i, o);
    red_merge<OP>(a, b);
  }
  return a;
}
template <int OP>
__device__ __forceinline__ void red_store(void* out, int64_t o, const Acc& a, float scale) {
  if constexpr (OP == CPT_RED_ARGMAX || OP == CPT_RED_COUNT) reinterpret_cast<long long*>(out)[o] = a.i;
  else if constexpr (OP == CPT_RED_ANY || OP == CPT_RED_ALL) reinterpret_cast<uint8_t*>(out)[o] = a.v != 0.f;
  else reinterpret_cast<float*>(out)[o] = (OP == CPT_RED_SUM) ? a.v * scale : a.v;
}

// kept dims (kd, ks) and reduced dims (rd, rs), element strides, innermost last
struct RedGeom {
  int nk, nr;
  int64_t kd[3], ks[3], rd[3], rs[3];
  int64_t K, R;  // products
};
__device__ __forceinline__ int64_t offs(int64_t idx, int n, const int64_t* d, const int64_t* s) {
  if (n <= 1) return n == 1 ? idx * s[0] : 0;
  int64_t o = 0;
#pragma unroll
  for (int k = 2; k >= 0; --k) {
    if (k < n) {
      int64_t q, c;
      if (idx < 0x7fffffff && d[k] < 0x7fffffff) { q = (uint32_t)idx / (uint32_t)d[k]; c = (uint32_t)idx - (uint32_t)q * (uint32_t)d[k]; }
      else { q = idx / d[k]; c = idx - q * d[k]; }
      o += c * s[k]; idx = q;
    }
  }
  return o;
}

// Row form: block (o, split) reduces R/S consecutive reduced indices of output o; consecutive threads read consecutive
// reduced indices (coalesced when the innermost reduced dim has stride 1).  VEC: 128-bit loads when every row start and
// length is a multiple of 4 elements.
template <int OP, typename T, bool VEC>
__global__ void __launch_bounds__(256) reduce_rows_kernel(const T* __restrict__ x, void* __restrict__ out, Acc* __restrict__ part,
                                                          RedGeom G, int S, float scale) {
  __shared__ Acc sh[8];
  const int64_t o = blockIdx.x;
  const int s = blockIdx.y;
  const int64_t chunk = ((G.R + S - 1) / S + 3) & ~3ll, r0 = s * chunk, r1 = min(G.R, r0 + chunk);
  const T* xo = x + offs(o, G.nk, G.kd, G.ks);
  Acc a = red_init<OP>();
  if constexpr (VEC) {
    int64_t r = r0 + threadIdx.x * 4;
    for (; r + 3 * 1024 < r1; r += 4 * 1024) {  // four independent 128-bit loads in flight per thread
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = ld_stream(reinterpret_cast<const float4*>(xo + offs(r + u * 1024, G.nr, G.rd, G.rs)));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t q = r + u * 1024;
        red_step<OP>(a, v[u].x, q); red_step<OP>(a, v[u].y, q + 1); red_step<OP>(a, v[u].z, q + 2); red_step<OP>(a, v[u].w, q + 3);
      }
    }
    for (; r < r1; r += 1024) {
      const float4 v = ld_stream(reinterpret_cast<const float4*>(xo + offs(r, G.nr, G.rd, G.rs)));
      red_step<OP>(a, v.x, r); red_step<OP>(a, v.y, r + 1); red_step<OP>(a, v.z, r + 2); red_step<OP>(a, v.w, r + 3);
    }
  } else {
    int64_t r = r0 + threadIdx.x;
    for (; r + 3 * 256 < r1; r += 4 * 256) {
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = (float)xo[offs(r + u * 256, G.nr, G.rd, G.rs)];
#pragma unroll
      for (int u = 0; u < 4; ++u) red_step<OP>(a, v[u], r + u * 256);
    }
    for (; r < r1; r += 256) red_step<OP>(a, (float)xo[offs(r, G.nr, G.rd, G.rs)], r);
  }
  a = warp_red<OP>(a);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) red_merge<OP>(a, sh[w]);
    if (S == 1) red_store<OP>(out, o, a, scale);
    else part[(int64_t)s * G.K + o] = a;
  }
}

// Row form for short reductions (R <= a few thousand, many outputs): one warp per output, no shared memory, no split.
template <int OP, typename T, bool VEC>
__global__ void __launch_bounds__(256) reduce_rows_warp_kernel(const T* __restrict__ x, void* __restrict__ out, RedGeom G, float scale) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t o = w0; o < G.K; o += nw) {
    const T* xo = x + offs(o, G.nk, G.kd, G.ks);
    Acc a = red_init<OP>();
    if constexpr (VEC) {
      for (int64_t r = lane * 4; r < G.R; r += 128) {
        const float4 v = ld_stream(reinterpret_cast<const float4*>(xo + offs(r, G.nr, G.rd, G.rs)));
        red_step<OP>(a, v.x, r); red_step<OP>(a, v.y, r + 1); red_step<OP>(a, v.z, r + 2); red_step<OP>(a, v.w, r + 3);
      }
    } else {
      for (int64_t r = lane; r < G.R; r += 32) red_step<OP>(a, (float)xo[offs(r, G.nr, G.rd, G.rs)], r);
    }
    a = warp_red<OP>(a);
    if (lane == 0) red_store<OP>(out, o, a, scale);
  }
}

// Column form: innermost kept dim is contiguous.  Block = 32 lanes along it x 8 rows along the reduced index.
template <int OP, typename T>
__global__ void __launch_bounds__(256) reduce_cols_kernel(const T* __restrict__ x, void* __restrict__ out, Acc* __restrict__ part,
                                                          RedGeom G, int S, float scale, int64_t inner_blocks) {
  __shared__ Acc sh[8][33];
  const int lane = threadIdx.x & 31, row = threadIdx.x >> 5;
  const int64_t kin = G.kd[G.nk - 1];
  const int64_t ko = blockIdx.x / inner_blocks, ki = (blockIdx.x % inner_blocks) * 32 + lane;
  const int s = blockIdx.y;
  const int64_t chunk = (G.R + S - 1) / S, r0 = s * chunk, r1 = min(G.R, r0 + chunk);
  Acc a = red_init<OP>();
  if (ki < kin) {
    const T* xo = x + offs(ko, G.nk - 1, G.kd, G.ks) + ki;
    int64_t r = r0 + row;
    for (; r + 24 < r1; r += 32) {  // four independent loads in flight per thread
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = (float)xo[offs(r + 8 * u, G.nr, G.rd, G.rs)];
#pragma unroll
      for (int u = 0; u < 4; ++u) red_step<OP>(a, v[u], r + 8 * u);
    }
    for (; r < r1; r += 8) red_step<OP>(a, (float)xo[offs(r, G.nr, G.rd, G.rs)], r);
  }
  sh[row][lane] = a;
  __syncthreads();
  if (row == 0 && ki < kin) {
    for (int w = 1; w < 8; ++w) red_merge<OP>(a, sh[w][lane]);
    const int64_t o = ko * kin + ki;
    if (S == 1) red_store<OP>(out, o, a, scale);
    else part[(int64_t)s * G.K + o] = a;
  }
}

template <int OP>
__global__ void __launch_bounds__(256) reduce_finish_kernel(const Acc* __restrict__ part, void* __restrict__ out, int64_t K, int S,
                                                            float scale) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= K) return;
  Acc a = part[o];
  for (int s = 1; s < S; ++s) red_merge<OP>(a, part[(int64_t)s * K + o]);  // fixed order: deterministic
  red_store<OP>(out, o, a, scale);
}

template <typename F>
static bool dispatch_red(int op, F&& f) {
  switch (op) {
#define C(X) case X: f(std::integral_constant<int, X>{}); return true;
    C(CPT_RED_SUM) C(CPT_RED_SUMSQ) C(CPT_RED_PROD) C(CPT_RED_MAX) C(CPT_RED_MIN) C(CPT_RED_ANY) C(CPT_RED_ALL) C(CPT_RED_COUNT)
    C(CPT_RED_ARGMAX)
#undef C
  }
  return false;
}

// ------------------------------------------------------------------------------------------------ data movement
template <typename T>
__global__ void __launch_bounds__(256) strided_copy_kernel(T* __restrict__ dst, const T* __restrict__ src, Dims D, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int64_t r = i, od = 0, os = 0;
#pragma unroll
    for (int k = CPT_MAX_DIMS - 1; k >= 0; --k) {
      if (k < D.nd) {
        int64_t q, c;
        if (r < 0x7fffffff && D.d[k] < 0x7fffffff) { q = (uint32_t)r / (uint32_t)D.d[k]; c = (uint32_t)r - (uint32_t)q * (uint32_t)D.d[k]; }
        else { q = r / D.d[k]; c = r - q * D.d[k]; }
        od += c * D.sa[k]; os += c * D.sb[k]; r = q;
      }
    }
    dst[od] = src[os];
  }
}

// Row form of the strided copy (innermost merged dim contiguous in dst, stride 0 or 1 in src): warp per (row, segment).
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) strided_copy_rows_kernel(T* __restrict__ dst, const T* __restrict__ src, Dims D, int64_t L,
                                                                int ss_in, int64_t items, int segs) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t w = w0; w < items; w += nw) {
    const RowItem it = row_item(w, segs, D, L);
    T* dp = dst + it.oa;
    const T* sp = src + it.ob;
    if constexpr (VEC) {  // T is 4 bytes, rows 16-byte aligned on both sides
      for (int64_t i = it.e0 + lane * 4; i < it.e1; i += 128) *reinterpret_cast<uint4*>(dp + i) = *reinterpret_cast<const uint4*>(sp + i);
    } else if (ss_in) {
      for (int64_t i = it.e0 + lane; i < it.e1; i += 32) dp[i] = sp[i];
    } else {
      const T v = sp[0];
      for (int64_t i = it.e0 + lane; i < it.e1; i += 32) dp[i] = v;
    }
  }
}

// Batched 2-D transposition of 4-byte elements: dst[bt][r][c] (C-contiguous) = src[bt*sB + r + c*sC] — 32x32 tiles through
// shared memory, coalesced on both sides (permute / transpose / NCHW <-> NHWC of the generic operator set).
__global__ void __launch_bounds__(256) transpose_tile_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, int64_t R,
                                                             int64_t Cc, int64_t sB, int64_t sC, int64_t tiles_r, int64_t tiles_c,
                                                             int64_t total_tiles) {
  __shared__ uint32_t t[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int64_t bt = tile / (tiles_r * tiles_c), rem = tile - bt * tiles_r * tiles_c;
    const int64_t tr = rem / tiles_c, tc = rem - tr * tiles_c;
    const int64_t r0 = tr * 32, c0 = tc * 32;
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
      const int64_t r = r0 + tx, c = c0 + j;
      if (r < R && c < Cc) t[j][tx] = src[bt * sB + r + c * sC];
    }
    __syncthreads();
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
      const int64_t r = r0 + j, c = c0 + tx;
      if (r < R && c < Cc) dst[(bt * R + r) * Cc + c] = t[tx][j];
    }
    __syncthreads();
  }
}

// dst[i, :] = src[idx[i], :]: warp per (index, 8 KB segment of the row); W = widest word the row size / alignment allows
constexpr int GATHER_SEG_BYTES = 8192;
template <typename I, typename W>
__global__ void __launch_bounds__(256) gather_rows_kernel(W* __restrict__ dst, const W* __restrict__ src, const I* __restrict__ idx,
                                                          int64_t row_w, int64_t n_src_rows, int* __restrict__ err, int64_t items,
                                                          int segs) {
  constexpr int SEG_W = GATHER_SEG_BYTES / (int)sizeof(W);
  const int lane = threadIdx.x & 31;
  const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t w = w0; w < items; w += nw) {
    const int64_t r = w / segs;
    const int seg = (int)(w - r * segs);
    int64_t j = (int64_t)idx[r];
    if (j < 0) j += n_src_rows;
    if (j < 0 || j >= n_src_rows) {
      if (err && lane == 0) *err = 1;
      j = j < 0 ? 0 : n_src_rows - 1;
    }
    const int64_t e0 = (int64_t)seg * SEG_W, e1 = e0 + SEG_W < row_w ? e0 + SEG_W : row_w;
    const W* sp = src + j * row_w;
    W* dp = dst + r * row_w;
    for (int64_t i = e0 + lane; i < e1; i += 32) dp[i] = sp[i];
  }
}

template <typename D, typename S>
__global__ void __launch_bounds__(256) cast_kernel(D* __restrict__ dst, const S* __restrict__ src, int64_t n, int to_bool) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = to_bool ? (D)(src[i] != (S)0) : (D)src[i];
}

template <typename T>
__global__ void __launch_bounds__(256) arange_kernel(T* __restrict__ dst, double start, double step, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = (T)(start + step * (double)i);
}

__device__ __forceinline__ uint64_t mix64(uint64_t seed, uint64_t idx) {  // splitmix64 finaliser (as loss.cu's dropout RNG)
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// kind 0: uniform [lo, hi)   1: normal(mean=lo, std=hi) by Box-Muller   2: integers in [lo, hi) stored as float
__global__ void __launch_bounds__(256) random_kernel(float* __restrict__ dst, int64_t n, float lo, float hi, uint64_t seed, int kind) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint64_t z = mix64(seed, (uint64_t)i);
    const float u = (float)(z >> 40) * (1.0f / 16777216.0f);  // [0, 1)
    // uniform: drawn in double (53 random bits) and cast, like numpy.random.uniform(...).astype(float32) in the reference
    // (random.py:118-150): every float32 of the interval can occur, so exact duplicates — ties in a max-pooling window, where
    // the reference and torch disagree by design — are as rare as on the NumPy path (a 24-bit draw repeats values ~60x more often)
    if (kind == 0) dst[i] = fminf((float)((double)lo + ((double)hi - (double)lo) * ((double)(z >> 11) * (1.0 / 9007199254740992.0))), nextafterf(hi, lo));
    else if (kind == 2) dst[i] = fminf(floorf(lo + (hi - lo) * u), hi - 1.f);
    else {
      const float u2 = (float)((z >> 16) & 0xffffff) * (1.0f / 16777216.0f);
      dst[i] = lo + hi * sqrtf(-2.f * logf(1.f - u)) * cospif(2.f * u2);
    }
  }
}

// grid of 256-thread CTAs for warp-per-item kernels: ~8 resident CTAs per SM, never more warps than items
static int warp_grid(int64_t items) {
  int64_t need = (items + 7) / 8, cap = (int64_t)sm_count() * 8;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// splits the innermost merged dim off D (nd >= 1): returns its length and leaves the outer dims in D
static int64_t pop_inner(Dims& D, int64_t& sa_in, int64_t& sb_in) {
  const int k = D.nd - 1;
  const int64_t L = D.d[k];
  sa_in = D.sa[k]; sb_in = D.sb[k];
  D.d[k] = 1; D.sa[k] = D.sb[k] = 0;
  D.nd = k;
  return L;
}

static int merge_and_check(int ndim, const int64_t* dims, const int64_t* sa, const int64_t* sb, Dims& D, int64_t& n, const char* who) {
  CPT_REQUIRE(ndim >= 0 && ndim <= CPT_MAX_DIMS && (ndim == 0 || (dims && sa && sb)), CPT_ERR_INVALID,
              "%s: 0 <= ndim <= %d and dims/strides required", who, CPT_MAX_DIMS);
  n = 1;
  D.nd = 0;
  for (int k = 0; k < ndim; ++k) {
    CPT_REQUIRE(dims[k] >= 0, CPT_ERR_INVALID, "%s: negative dimension", who);
    n *= dims[k];
    if (dims[k] == 1) continue;
    // merge with the previous dim when both operands step contiguously across the boundary
    if (D.nd > 0 && D.sa[D.nd - 1] == sa[k] * dims[k] && D.sb[D.nd - 1] == sb[k] * dims[k]) {
      D.d[D.nd - 1] *= dims[k]; D.sa[D.nd - 1] = sa[k]; D.sb[D.nd - 1] = sb[k];
    } else {
      D.d[D.nd] = dims[k]; D.sa[D.nd] = sa[k]; D.sb[D.nd] = sb[k]; ++D.nd;
    }
  }
  for (int k = D.nd; k < CPT_MAX_DIMS; ++k) { D.d[k] = 1; D.sa[k] = D.sb[k] = 0; }
  return CPT_OK;
}

}  // namespace cpt

using namespace cpt;

extern "C" {

int cpt_ew_binary(int op, void* out, const float* a, const float* b, float scalar, int scalar_mode, int ndim, const int64_t* dims,
                  const int64_t* sa, const int64_t* sb, void* stream) {
  CPT_REQUIRE(out && a && (b || scalar_mode == 1 || scalar_mode == 2), CPT_ERR_INVALID, "ew_binary: bad arguments");
  Dims D;
  int64_t n;
  static const int64_t zeros[CPT_MAX_DIMS] = {0, 0, 0, 0, 0, 0};
  if (!b) sb = zeros;
  int rc = merge_and_check(ndim, dims, sa, sb, D, n, "ew_binary");
  if (rc != CPT_OK) return rc;
  if (n == 0) return CPT_OK;
  cudaStream_t st = as_stream(stream);
  // flat: both operands walk the output contiguously (or the second one is the scalar)
  const bool flat_a = D.nd == 0 || (D.nd == 1 && D.sa[0] == 1);
  const bool flat = flat_a && (!b || D.nd == 0 || D.sb[0] == 1);
  bool ok;
  if (flat) {
    const bool al = aligned16(a) && (!b || aligned16(b)) && aligned16(out);
    ok = dispatch_bin(op, [&](auto tag) {
      constexpr int OP = decltype(tag)::value;
      using OutT = std::conditional_t<is_cmp(OP), uint8_t, float>;
      const int64_t n4 = (al && (is_cmp(OP) ? (reinterpret_cast<uintptr_t>(out) & 3) == 0 : true)) ? n / 4 : 0;
      const int grid = ew_grid(n4 > 0 ? n4 : n, 256);
      OutT* o = reinterpret_cast<OutT*>(out);
      if (b) ew_bin_flat_kernel<OP, 0, OutT><<<grid, 256, 0, st>>>(o, a, b, 0.f, n4, n);
      else if (scalar_mode == 1) ew_bin_flat_kernel<OP, 1, OutT><<<grid, 256, 0, st>>>(o, a, nullptr, scalar, n4, n);
      else ew_bin_flat_kernel<OP, 2, OutT><<<grid, 256, 0, st>>>(o, a, nullptr, scalar, n4, n);
    });
  } else {
    CPT_REQUIRE(b, CPT_ERR_INVALID, "ew_binary: a scalar operand needs a contiguous tensor operand");
    const int ki = D.nd - 1;
    const bool rows_ok = D.d[ki] >= 32 && (D.sa[ki] == 0 || D.sa[ki] == 1) && (D.sb[ki] == 0 || D.sb[ki] == 1);
    if (rows_ok) {
      Dims Do = D;
      int64_t sa_in, sb_in;
      const int64_t L = pop_inner(Do, sa_in, sb_in);
      const int segs = (int)((L + ROW_SEG - 1) / ROW_SEG);
      const int64_t items = (n / L) * segs;
      bool vec = L % 4 == 0 && aligned16(a) && aligned16(b) && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
      for (int k = 0; k < Do.nd; ++k) vec = vec && (sa_in == 0 || Do.sa[k] % 4 == 0) && (sb_in == 0 || Do.sb[k] % 4 == 0);
      ok = dispatch_bin(op, [&](auto tag) {
        constexpr int OP = decltype(tag)::value;
        using OutT = std::conditional_t<is_cmp(OP), uint8_t, float>;
        OutT* o = reinterpret_cast<OutT*>(out);
        if (vec) ew_bin_rows_kernel<OP, OutT, true><<<warp_grid(items), 256, 0, st>>>(o, a, b, Do, L, (int)sa_in, (int)sb_in, items, segs);
        else ew_bin_rows_kernel<OP, OutT, false><<<warp_grid(items), 256, 0, st>>>(o, a, b, Do, L, (int)sa_in, (int)sb_in, items, segs);
      });
    } else {
      ok = dispatch_bin(op, [&](auto tag) {
        constexpr int OP = decltype(tag)::value;
        using OutT = std::conditional_t<is_cmp(OP), uint8_t, float>;
        ew_bin_bcast_kernel<OP, OutT><<<ew_grid(n, 256), 256, 0, st>>>(reinterpret_cast<OutT*>(out), a, b, D, n);
      });
    }
  }
  CPT_REQUIRE(ok, CPT_ERR_UNSUPPORTED, "ew_binary: unknown op %d", op);
  CPT_LAUNCH_CHECK("ew_binary");
  return CPT_OK;
}

int cpt_ew_unary(int op, void* out, const float* a, float p0, float p1, int64_t n, void* stream) {
  CPT_REQUIRE(out && a && n >= 0, CPT_ERR_INVALID, "ew_unary: bad arguments");
  if (n == 0) return CPT_OK;
  const int64_t n4 = (aligned16(a) && aligned16(out)) ? n / 4 : 0;
  const bool ok = dispatch_un(op, [&](auto tag) {
    constexpr int OP = decltype(tag)::value;
    using OutT = std::conditional_t<un_is_bool(OP), uint8_t, float>;
    ew_un_kernel<OP, OutT><<<ew_grid(n4 > 0 ? n4 : n, 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<OutT*>(out), a, p0, p1, n4, n);
  });
  CPT_REQUIRE(ok, CPT_ERR_UNSUPPORTED, "ew_unary: unknown op %d", op);
  CPT_LAUNCH_CHECK("ew_unary");
  return CPT_OK;
}

int cpt_logic(int op, uint8_t* out, const uint8_t* a, const uint8_t* b, int64_t n, void* stream) {
  CPT_REQUIRE(out && a && n >= 0 && op >= 0 && op <= 3 && (b || op == 3), CPT_ERR_INVALID, "logic: bad arguments");
  if (n == 0) return CPT_OK;
  logic_kernel<<<ew_grid(n, 256), 256, 0, as_stream(stream)>>>(out, a, b, op, n);
  CPT_LAUNCH_CHECK("logic");
  return CPT_OK;
}

size_t cpt_reduce_workspace_size(int64_t n_out) {
  // splits * outputs <= 16 CTAs per SM * 32 outputs per CTA (column form), never less than one split per output
  return (size_t)((int64_t)sm_count() * 16 * 32 + (n_out > 0 ? n_out : 1)) * sizeof(Acc) + 256;
}

int cpt_reduce(int op, void* out, const void* x, int x_dtype, int ndim, const int64_t* dims, const int32_t* reduced, float scale,
               void* ws, size_t ws_bytes, void* stream) {
  CPT_REQUIRE(out && x && ndim >= 0 && ndim <= CPT_MAX_DIMS && (ndim == 0 || (dims && reduced)), CPT_ERR_INVALID,
              "reduce: bad arguments");
  CPT_REQUIRE(x_dtype == CPT_DT_F32 || x_dtype == CPT_DT_U8, CPT_ERR_UNSUPPORTED, "reduce: float32 or bool input only");
  // merge adjacent dims of the same kind (the input is C-contiguous); drop size-1 dims
  RedGeom G;
  G.nk = G.nr = 0; G.K = G.R = 1;
  int64_t md[CPT_MAX_DIMS], ms[CPT_MAX_DIMS];
  int mk[CPT_MAX_DIMS], nm = 0;
  int64_t stride = 1;
  for (int k = ndim - 1; k >= 0; --k) {
    CPT_REQUIRE(dims[k] >= 0, CPT_ERR_INVALID, "reduce: negative dimension");
    const int kind = reduced[k] ? 1 : 0;
    if (dims[k] != 1) {
      if (nm > 0 && mk[nm - 1] == kind) md[nm - 1] *= dims[k];
      else { md[nm] = dims[k]; ms[nm] = stride; mk[nm] = kind; ++nm; }
    }
    stride *= dims[k];
  }
  // md[] is innermost-first; geometry arrays are innermost-last
  int nk = 0, nr = 0;
  for (int j = 0; j < nm; ++j) (mk[j] ? nr : nk)++;
  CPT_REQUIRE(nk <= 3 && nr <= 3, CPT_ERR_UNSUPPORTED, "reduce: more than 3 separate kept / reduced axis groups");
  G.nk = nk; G.nr = nr;
  int ik = nk, ir = nr;
  for (int j = 0; j < nm; ++j) {
    if (mk[j]) { --ir; G.rd[ir] = md[j]; G.rs[ir] = ms[j]; G.R *= md[j]; }
    else { --ik; G.kd[ik] = md[j]; G.ks[ik] = ms[j]; G.K *= md[j]; }
  }
  for (int j = nk; j < 3; ++j) { G.kd[j] = 1; G.ks[j] = 0; }
  for (int j = nr; j < 3; ++j) { G.rd[j] = 1; G.rs[j] = 0; }
  if (G.K == 0) return CPT_OK;
  CPT_REQUIRE(G.R > 0 || (op != CPT_RED_MAX && op != CPT_RED_MIN && op != CPT_RED_ARGMAX), CPT_ERR_INVALID,
              "reduce: zero-size reduction has no identity");
  CPT_REQUIRE(op != CPT_RED_ARGMAX || nr <= 1, CPT_ERR_UNSUPPORTED, "argmax: one axis (or the flattened tensor) only");
  cudaStream_t st = as_stream(stream);
  // column form when the innermost merged dim is kept and wide enough to fill the lanes
  const bool cols = nm > 0 && mk[0] == 0 && md[0] >= 16;
  const int64_t inner_blocks = cols ? (G.kd[nk - 1] + 31) / 32 : 1;
  const int64_t blocks = cols ? (G.K / G.kd[nk - 1]) * inner_blocks : G.K;
  CPT_REQUIRE(blocks < 0x7fffffff, CPT_ERR_UNSUPPORTED, "reduce: too many output slices");
  // split the reduced range over grid.y until ~2 waves of CTAs are in flight (each split at least 2048 elements)
  int64_t S = 1;
  const int64_t want = (int64_t)sm_count() * 16;
  if (blocks < want) {
    S = want / blocks;
    const int64_t maxS = (G.R + (cols ? 63 : 2047)) / (cols ? 64 : 2048);
    if (S > maxS) S = maxS;
    if (S > 65535) S = 65535;
    if (S < 1) S = 1;
  }
  if (S > 1) {
    const size_t need = (size_t)(S * G.K) * sizeof(Acc);
    if (!ws || ws_bytes < need) {
      S = ws ? (int64_t)(ws_bytes / ((size_t)G.K * sizeof(Acc))) : 1;
      if (S < 1) S = 1;
    }
  }
  Acc* part = reinterpret_cast<Acc*>(ws);
  CPT_REQUIRE(S == 1 || (reinterpret_cast<uintptr_t>(ws) & 15) == 0, CPT_ERR_INVALID, "reduce: workspace must be 16-byte aligned");
  const bool f32 = x_dtype == CPT_DT_F32;
  bool vec = f32 && !cols && aligned16(x) && G.rs[nr > 0 ? nr - 1 : 0] == 1 && (G.rd[nr > 0 ? nr - 1 : 0] % 4 == 0);
  for (int j = 0; j < 3; ++j) vec = vec && (j >= nk || G.ks[j] % 4 == 0 || G.kd[j] == 1) && (j >= nr - 1 || G.rs[j] % 4 == 0);
  // short rows, many outputs: one warp per output instead of one CTA (no split, no second pass)
  const bool warp_rows = !cols && G.R <= 4096 && G.K >= (int64_t)sm_count() * 16;
  if (warp_rows) S = 1;
  const bool ok = dispatch_red(op, [&](auto tag) {
    constexpr int OP = decltype(tag)::value;
    dim3 grid((unsigned)blocks, (unsigned)S);
    if (cols) {
      if (f32) reduce_cols_kernel<OP, float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(x), out, part, G, (int)S, scale, inner_blocks);
      else reduce_cols_kernel<OP, uint8_t><<<grid, 256, 0, st>>>(reinterpret_cast<const uint8_t*>(x), out, part, G, (int)S, scale, inner_blocks);
    } else if (warp_rows) {
      const int wg = warp_grid(G.K);
      if (!f32) reduce_rows_warp_kernel<OP, uint8_t, false><<<wg, 256, 0, st>>>(reinterpret_cast<const uint8_t*>(x), out, G, scale);
      else if (vec) reduce_rows_warp_kernel<OP, float, true><<<wg, 256, 0, st>>>(reinterpret_cast<const float*>(x), out, G, scale);
      else reduce_rows_warp_kernel<OP, float, false><<<wg, 256, 0, st>>>(reinterpret_cast<const float*>(x), out, G, scale);
    } else if (f32) {
      if (vec) reduce_rows_kernel<OP, float, true><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(x), out, part, G, (int)S, scale);
      else reduce_rows_kernel<OP, float, false><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(x), out, part, G, (int)S, scale);
    } else {
      reduce_rows_kernel<OP, uint8_t, false><<<grid, 256, 0, st>>>(reinterpret_cast<const uint8_t*>(x), out, part, G, (int)S, scale);
    }
    count_launch();
    if (S > 1) reduce_finish_kernel<OP><<<(unsigned)((G.K + 255) / 256), 256, 0, st>>>(part, out, G.K, (int)S, scale);
  });
  CPT_REQUIRE(ok, CPT_ERR_UNSUPPORTED, "reduce: unknown op %d", op);
  CPT_LAUNCH_CHECK("reduce");
  return CPT_OK;
}

int cpt_strided_copy(void* dst, const void* src, int elem_size, int ndim, const int64_t* dims, const int64_t* dst_strides,
                     const int64_t* src_strides, void* stream) {
  CPT_REQUIRE(dst && src, CPT_ERR_INVALID, "strided_copy: bad arguments");
  Dims D;
  int64_t n;
  int rc = merge_and_check(ndim, dims, dst_strides, src_strides, D, n, "strided_copy");
  if (rc != CPT_OK) return rc;
  if (n == 0) return CPT_OK;
  cudaStream_t st = as_stream(stream);
  CPT_REQUIRE(elem_size == 1 || elem_size == 4 || elem_size == 8, CPT_ERR_UNSUPPORTED, "strided_copy: element size %d", elem_size);
  const int ki = D.nd - 1;
  // (1) batched 2-D transposition (dst contiguous, src contiguous along the second-to-last dim): tiled through shared memory
  if (elem_size == 4 && (D.nd == 2 || D.nd == 3) && D.sa[ki] == 1 && D.sa[ki - 1] == D.d[ki] && D.sb[ki - 1] == 1 &&
      (D.nd == 2 || D.sa[0] == D.d[1] * D.d[2]) && D.d[ki] >= 8 && D.d[ki - 1] >= 8 && D.sb[ki] > 0 && (D.nd == 2 || D.sb[0] >= 0)) {
    const int64_t R = D.d[ki - 1], Cc = D.d[ki], Bt = D.nd == 3 ? D.d[0] : 1, sB = D.nd == 3 ? D.sb[0] : 0;
    const int64_t tr = (R + 31) / 32, tc = (Cc + 31) / 32, total = Bt * tr * tc;
    const int64_t cap = (int64_t)sm_count() * 16;
    transpose_tile_kernel<<<(unsigned)(total < cap ? total : cap), 256, 0, st>>>((uint32_t*)dst, (const uint32_t*)src, R, Cc, sB, D.sb[ki], tr,
                                                                                tc, total);
    CPT_LAUNCH_CHECK("transpose_tile");
    return CPT_OK;
  }
  // (2) rows: innermost dim contiguous in dst and stride 0 / 1 in src
  if (D.nd >= 1 && D.d[ki] >= 32 && D.sa[ki] == 1 && (D.sb[ki] == 0 || D.sb[ki] == 1)) {
    Dims Do = D;
    int64_t sd_in, ss_in;
    const int64_t L = pop_inner(Do, sd_in, ss_in);
    const int segs = (int)((L + ROW_SEG - 1) / ROW_SEG);
    const int64_t items = (n / L) * segs;
    bool vec = elem_size == 4 && ss_in == 1 && L % 4 == 0 && aligned16(dst) && aligned16(src);
    for (int k = 0; k < Do.nd; ++k) vec = vec && Do.sa[k] % 4 == 0 && Do.sb[k] % 4 == 0;
    const int g = warp_grid(items);
    if (elem_size == 1) strided_copy_rows_kernel<uint8_t, false><<<g, 256, 0, st>>>((uint8_t*)dst, (const uint8_t*)src, Do, L, (int)ss_in, items, segs);
    else if (elem_size == 8) strided_copy_rows_kernel<uint64_t, false><<<g, 256, 0, st>>>((uint64_t*)dst, (const uint64_t*)src, Do, L, (int)ss_in, items, segs);
    else if (vec) strided_copy_rows_kernel<uint32_t, true><<<g, 256, 0, st>>>((uint32_t*)dst, (const uint32_t*)src, Do, L, (int)ss_in, items, segs);
    else strided_copy_rows_kernel<uint32_t, false><<<g, 256, 0, st>>>((uint32_t*)dst, (const uint32_t*)src, Do, L, (int)ss_in, items, segs);
    CPT_LAUNCH_CHECK("strided_copy_rows");
    return CPT_OK;
  }
  // (3) anything else: one index decomposition per element
  const int grid = ew_grid(n, 256);
  if (elem_size == 1) strided_copy_kernel<uint8_t><<<grid, 256, 0, st>>>((uint8_t*)dst, (const uint8_t*)src, D, n);
  else if (elem_size == 4) strided_copy_kernel<uint32_t><<<grid, 256, 0, st>>>((uint32_t*)dst, (const uint32_t*)src, D, n);
  else strided_copy_kernel<uint64_t><<<grid, 256, 0, st>>>((uint64_t*)dst, (const uint64_t*)src, D, n);
  CPT_LAUNCH_CHECK("strided_copy");
  return CPT_OK;
}

int cpt_gather_rows(void* dst, const void* src, const void* idx, int idx_dtype, int64_t n_idx, int64_t row_bytes, int64_t n_src_rows,
                    int* err_flag, void* stream) {
  CPT_REQUIRE(dst && src && idx && n_idx >= 0 && row_bytes > 0 && n_src_rows > 0, CPT_ERR_INVALID, "gather_rows: bad arguments");
  CPT_REQUIRE(idx_dtype == CPT_DT_I32 || idx_dtype == CPT_DT_I64, CPT_ERR_UNSUPPORTED, "gather_rows: int32 / int64 indices only");
  if (n_idx == 0) return CPT_OK;
  cudaStream_t st = as_stream(stream);
  const uintptr_t al = reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src) | (uintptr_t)row_bytes;
  const int wbytes = (al & 15) == 0 ? 16 : (al & 3) == 0 ? 4 : 1;
  const int64_t row_w = row_bytes / wbytes;
  const int segs = (int)((row_bytes + GATHER_SEG_BYTES - 1) / GATHER_SEG_BYTES);
  const int64_t items = n_idx * segs;
  const int grid = warp_grid(items);
#define GO(I, W) gather_rows_kernel<I, W><<<grid, 256, 0, st>>>((W*)dst, (const W*)src, (const I*)idx, row_w, n_src_rows, err_flag, items, segs)
  if (idx_dtype == CPT_DT_I32) { if (wbytes == 16) GO(int32_t, uint4); else if (wbytes == 4) GO(int32_t, uint32_t); else GO(int32_t, uint8_t); }
  else { if (wbytes == 16) GO(int64_t, uint4); else if (wbytes == 4) GO(int64_t, uint32_t); else GO(int64_t, uint8_t); }
#undef GO
  CPT_LAUNCH_CHECK("gather_rows");
  return CPT_OK;
}

int cpt_cast(void* dst, int dst_dtype, const void* src, int src_dtype, int64_t n, void* stream) {
  CPT_REQUIRE(dst && src && n >= 0, CPT_ERR_INVALID, "cast: bad arguments");
  if (n == 0) return CPT_OK;
  cudaStream_t st = as_stream(stream);
  const int grid = ew_grid(n, 256);
  bool ok = false;
#define ROW(SD, ST)                                                                                                         \
  if (src_dtype == SD) {                                                                                                    \
    const ST* s = (const ST*)src;                                                                                           \
    if (dst_dtype == CPT_DT_F32) { cast_kernel<float, ST><<<grid, 256, 0, st>>>((float*)dst, s, n, 0); ok = true; }         \
    if (dst_dtype == CPT_DT_I32) { cast_kernel<int32_t, ST><<<grid, 256, 0, st>>>((int32_t*)dst, s, n, 0); ok = true; }     \
    if (dst_dtype == CPT_DT_I64) { cast_kernel<int64_t, ST><<<grid, 256, 0, st>>>((int64_t*)dst, s, n, 0); ok = true; }     \
    if (dst_dtype == CPT_DT_U8) { cast_kernel<uint8_t, ST><<<grid, 256, 0, st>>>((uint8_t*)dst, s, n, 1); ok = true; }      \
    if (dst_dtype == CPT_DT_F64) { cast_kernel<double, ST><<<grid, 256, 0, st>>>((double*)dst, s, n, 0); ok = true; }       \
  }
  ROW(CPT_DT_F32, float) ROW(CPT_DT_I32, int32_t) ROW(CPT_DT_I64, int64_t) ROW(CPT_DT_U8, uint8_t) ROW(CPT_DT_F64, double)
#undef ROW
  CPT_REQUIRE(ok, CPT_ERR_UNSUPPORTED, "cast: dtype %d -> %d", src_dtype, dst_dtype);
  CPT_LAUNCH_CHECK("cast");
  return CPT_OK;
}

int cpt_arange(void* dst, int dtype, double start, double step, int64_t n, void* stream) {
  CPT_REQUIRE(dst && n >= 0, CPT_ERR_INVALID, "arange: bad arguments");
  if (n == 0) return CPT_OK;
  cudaStream_t st = as_stream(stream);
  const int grid = ew_grid(n, 256);
  if (dtype == CPT_DT_F32) arange_kernel<float><<<grid, 256, 0, st>>>((float*)dst, start, step, n);
  else if (dtype == CPT_DT_I32) arange_kernel<int32_t><<<grid, 256, 0, st>>>((int32_t*)dst, start, step, n);
  else if (dtype == CPT_DT_I64) arange_kernel<int64_t><<<grid, 256, 0, st>>>((int64_t*)dst, start, step, n);
  else CPT_REQUIRE(false, CPT_ERR_UNSUPPORTED, "arange: dtype %d", dtype);
  CPT_LAUNCH_CHECK("arange");
  return CPT_OK;
}

int cpt_random_fill(float* dst, int64_t n, int kind, float p0, float p1, uint64_t seed, void* stream) {
  CPT_REQUIRE(dst && n >= 0 && kind >= 0 && kind <= 2, CPT_ERR_INVALID, "random_fill: bad arguments");
  if (n == 0) return CPT_OK;
  random_kernel<<<ew_grid(n, 256), 256, 0, as_stream(stream)>>>(dst, n, p0, p1, seed, kind);
  CPT_LAUNCH_CHECK("random_fill");
  return CPT_OK;
}

}  // extern "C"
