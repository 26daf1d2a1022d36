// Host side of the tensor-core back end: TMA descriptor construction (tiled + im2col), operand staging
// kernels (NCHW fp32 -> NHWC bf16/tf32, weight re-layout, casts), and the per-op drivers that launch
// tc_kernel<> for Conv2D fprop / dgrad / wgrad and Linear fwd / dgrad / wgrad.
#include <cuda_bf16.h>

#include "simt_gemm.cuh"
#include <stdlib.h>

#include "tc.cuh"
#include "tc_launch.cuh"
#include "strip_params.cuh"

namespace cpt {
namespace tc {

// ------------------------------------------------------------------ driver entry points (no -lcuda needed)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode_tiled = nullptr;
static EncodeIm2colFn g_encode_im2col = nullptr;

static int load_driver() {
  if (g_encode_tiled && g_encode_im2col) return CPT_OK;
  void* f1 = nullptr;
  void* f2 = nullptr;
  cudaDriverEntryPointQueryResult q;
  CPT_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f1, cudaEnableDefault, &q));
  CPT_REQUIRE(f1 && q == cudaDriverEntryPointSuccess, CPT_ERR_CUDA, "cuTensorMapEncodeTiled not available");
  CPT_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f2, cudaEnableDefault, &q));
  CPT_REQUIRE(f2 && q == cudaDriverEntryPointSuccess, CPT_ERR_CUDA, "cuTensorMapEncodeIm2col not available");
  g_encode_tiled = reinterpret_cast<EncodeTiledFn>(f1);
  g_encode_im2col = reinterpret_cast<EncodeIm2colFn>(f2);
  return CPT_OK;
}

// CPT_MODE_FP32X3: fp32-exact contraction on the tensor cores — every operand is staged as TWO tf32 planes (hi = rna_tf32(a),
// lo = rna_tf32(a - hi), the lo plane `plane bytes` after the hi plane) and the kernel issues three MMAs per k-step
static inline bool is_x3(int mode) { return mode == CPT_MODE_FP32X3; }
static inline int planes(int mode) { return is_x3(mode) ? 2 : 1; }
static inline int esize(int mode) { return mode == CPT_MODE_BF16 ? 2 : 4; }
static inline int kc_of(int mode) { return 128 / esize(mode); }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int round_up(int v, int a) { return (v + a - 1) / a * a; }

// 2-D tiled map over a row-major [rows][cols] matrix with `pitch` elements per row; box = (box_cols, box_rows)
// mn_major: the tile feeds an MN-major UMMA operand; for 32-bit elements that needs the 32-byte-atom 128B swizzle
static inline CUtensorMapSwizzle swizzle_of(int mode, bool mn_major) {
  return (mn_major && mode != CPT_MODE_BF16) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
}

static int make_map_2d(CUtensorMap* m, const void* base, int mode, uint64_t cols, uint64_t rows, uint64_t pitch, int box_cols,
                       int box_rows, bool mn_major = false) {
  if (int e = load_driver()) return e;
  const int es = esize(mode);
  CPT_REQUIRE(((uintptr_t)base & 15) == 0 && (pitch * es) % 16 == 0, CPT_ERR_UNSUPPORTED,
              "TMA needs 16-byte aligned base and row pitch (pitch=%llu elems)", (unsigned long long)pitch);
  CPT_REQUIRE(box_cols * es == 128 && box_rows >= 1 && box_rows <= 256, CPT_ERR_INVALID, "bad TMA box %dx%d", box_cols, box_rows);
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch * es};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(m, mode == CPT_MODE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                              const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              swizzle_of(mode, mn_major), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CPT_REQUIRE(r == CUDA_SUCCESS, CPT_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) cols=%llu rows=%llu pitch=%llu box=%dx%d", (int)r,
              (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)pitch, box_cols, box_rows);
  return CPT_OK;
}

// im2col map over a channels-last activation tensor (N, H, W, C) [C innermost, Cp channels per pixel].
// Bounding box of the *base pixel* in input coordinates: [lower, extent - 1 + upper] per spatial dim, walked with
// `stride`; the filter-tap offset is added per load (PTX {off_w, off_h}).  Out-of-bounds reads are zero-filled.
static int make_map_im2col(CUtensorMap* m, const void* base, int mode, int Cp, int W, int H, int N, int lower_w, int lower_h,
                           int upper_w, int upper_h, int stride, int channels_per_pixel, int pixels_per_column,
                           bool mn_major = false) {
  if (int e = load_driver()) return e;
  const int es = esize(mode);
  CPT_REQUIRE(((uintptr_t)base & 15) == 0 && ((size_t)Cp * es) % 16 == 0, CPT_ERR_UNSUPPORTED, "im2col TMA alignment");
  CPT_REQUIRE(lower_w >= -128 && lower_w <= 127 && lower_h >= -128 && lower_h <= 127 && upper_w >= -128 && upper_w <= 127 &&
                  upper_h >= -128 && upper_h <= 127,
              CPT_ERR_UNSUPPORTED, "im2col corner out of the 8-bit range");
  CPT_REQUIRE(stride >= 1 && stride <= 8, CPT_ERR_UNSUPPORTED, "im2col traversal stride %d unsupported", stride);
  cuuint64_t dims[4] = {(cuuint64_t)Cp, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)Cp * es, (cuuint64_t)W * Cp * es, (cuuint64_t)H * W * Cp * es};
  int lo[2] = {lower_w, lower_h};
  int hi[2] = {upper_w, upper_h};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = g_encode_im2col(m, mode == CPT_MODE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                               const_cast<void*>(base), dims, strides, lo, hi, (cuuint32_t)channels_per_pixel,
                               (cuuint32_t)pixels_per_column, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_of(mode, mn_major),
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CPT_REQUIRE(r == CUDA_SUCCESS, CPT_ERR_CUDA,
              "cuTensorMapEncodeIm2col failed (%d) C=%d W=%d H=%d N=%d lo=%d,%d hi=%d,%d stride=%d cpp=%d ppc=%d", (int)r, Cp, W, H, N,
              lower_w, lower_h, upper_w, upper_h, stride, channels_per_pixel, pixels_per_column);
  // Known driver issue (also worked around by CUTLASS, copy_traits_sm90_im2col.hpp): for tensors < 128 KiB the
  // encoder sets a descriptor bit that makes small im2col loads fault; clear it.
  int drv = 0;
  cudaDriverGetVersion(&drv);
  if (drv <= 13010 && (size_t)Cp * W * H * N * es < 131072) reinterpret_cast<uint64_t*>(m)[1] &= ~(1ull << 21);
  return CPT_OK;
}

// ------------------------------------------------------------------ staging kernels
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// NCHW fp32 -> NHWC (C padded to Cp) bf16 / tf32-rounded fp32 — HBM-bound staging pass (6 B/elem bf16, 8 B/elem tf32).
// Tile = 64 channels x 128 pixels through shared memory, every access conflict-free and 128-byte coalesced:
//   load : lanes run along pixels (each warp reads 4 x 128 B of one channel row; 32 independent loads per thread)
//   store: lanes run along channels (each warp writes the 128 B / 2 x 128 B of one pixel)
// BF16 packs channel pairs into 32-bit words before the transpose, so the smem tile is [32 pairs][129] words and
// both phases hit 32 distinct banks.  chan_sum (optional) fuses db = dy.sum((0,2,3)): warp-shuffle tree + one atomic
// per (block, channel).
constexpr int CL_PX = 128, CL_CH = 64;

// Zero-padded destination (strip convolution path): pixel (b, h, w) goes to row b*Hp*Wp + (h + P)*Wp + (w + P)
struct PadGeom { int W, H, P; };

// GATE: every value is multiplied by (gate > 0) first, gate = a tensor of src's shape — the backward pass of a ReLU folded into
// the staging of dy (dy * mask, activation_funcs.py:32-34, with the mask taken from the ReLU's own output): 10 B/elem instead
// of 8 1/8 (ReLU backward) + 6 (staging); the channel sums (db) are those of the gated values.
template <bool BF16, bool PAD = false, bool GATE = false>
__global__ void __launch_bounds__(256, 4) nchw_to_nhwc_kernel(const float* __restrict__ src, void* __restrict__ dst, int C, int HW,
                                                           int Cp, float* __restrict__ chan_sum, float* __restrict__ partial,
                                                           int64_t Q, float* __restrict__ dst_lo, PadGeom pg = PadGeom{0, 0, 0},
                                                           const float* __restrict__ gate = nullptr) {
  // pixel tiles run over the flattened (image, pixel) index q in [0, Q = B*HW): small feature maps (HW < 128) fill the
  // 128-pixel tile with pixels of several images instead of leaving lanes idle
  __shared__ uint32_t tile[BF16 ? 32 : 64][CL_PX + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.y * CL_CH;
  const int64_t n_tiles = (Q + CL_PX - 1) / CL_PX;
  uint32_t* d = reinterpret_cast<uint32_t*>(dst);
  float acc[8];  // per-thread channel partial sums, reduced once per block (not once per tile)
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;

  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int64_t p0 = t * CL_PX;
    // element offset of (image, channel 0, pixel) for this lane's four pixels (32-bit: the host checks B*C*HW < 2^32 - 2^20)
    uint32_t soff[4];
    constexpr uint32_t PAST_END = 0xFFFFFFFFu;
#pragma unroll
    for (int pi = 0; pi < 4; ++pi) {
      const int64_t q = p0 + lane + 32 * pi;
      const uint32_t b = (uint32_t)q / (uint32_t)HW;
      soff[pi] = q < Q ? b * (uint32_t)(C * HW) + ((uint32_t)q - b * (uint32_t)HW) : PAST_END;
    }
    const float* s = src;
    if (BF16) {
      float v0[4][4], v1[4][4];
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        const int c = c0 + 2 * (warp + 8 * ci);
#pragma unroll
        for (int pi = 0; pi < 4; ++pi) {
          const bool okp = soff[pi] != PAST_END;
          v0[ci][pi] = (okp && c < C) ? s[soff[pi] + (uint32_t)(c * HW)] : 0.f;
          v1[ci][pi] = (okp && c + 1 < C) ? s[soff[pi] + (uint32_t)((c + 1) * HW)] : 0.f;
        }
      }
      if (GATE) {
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int c = c0 + 2 * (warp + 8 * ci);
#pragma unroll
          for (int pi = 0; pi < 4; ++pi) {
            const bool okp = soff[pi] != PAST_END;
            const float g0 = (okp && c < C) ? gate[soff[pi] + (uint32_t)(c * HW)] : 0.f;
            const float g1 = (okp && c + 1 < C) ? gate[soff[pi] + (uint32_t)((c + 1) * HW)] : 0.f;
            v0[ci][pi] *= g0 > 0.f ? 1.f : 0.f;  // a product, like numpy's float * bool: NaN / inf gradients stay visible
            v1[ci][pi] *= g1 > 0.f ? 1.f : 0.f;
          }
        }
      }
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
#pragma unroll
        for (int pi = 0; pi < 4; ++pi) {
          __nv_bfloat162 h = __floats2bfloat162_rn(v0[ci][pi], v1[ci][pi]);
          tile[warp + 8 * ci][lane + 32 * pi] = *reinterpret_cast<uint32_t*>(&h);
        }
        acc[2 * ci] += (v0[ci][0] + v0[ci][1]) + (v0[ci][2] + v0[ci][3]);
        acc[2 * ci + 1] += (v1[ci][0] + v1[ci][1]) + (v1[ci][2] + v1[ci][3]);
      }
      __syncthreads();
      const int c = c0 + 2 * lane;  // this lane's channel pair
      if (c < Cp) {
        if (PAD) {
          // destination row of pixel q = p0 + warp + 8 i, advanced incrementally (one division per tile, not per store)
          const int Wp = pg.W + 2 * pg.P, HpWp = (pg.H + 2 * pg.P) * Wp;
          int64_t q = p0 + warp;
          int b = (int)(q / HW), r = (int)(q - (int64_t)b * HW), h = r / pg.W, w = r - h * pg.W;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (q < Q) {
              const int64_t row = (int64_t)b * HpWp + (int64_t)(h + pg.P) * Wp + (w + pg.P);
              d[(row * Cp + c) >> 1] = tile[lane][warp + 8 * i];
            }
            q += 8; w += 8;
            while (w >= pg.W) { w -= pg.W; ++h; }
            while (h >= pg.H) { h -= pg.H; ++b; }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int pp = warp + 8 * i;
            const int64_t q = p0 + pp;
            if (q < Q) d[(q * Cp + c) >> 1] = tile[lane][pp];
          }
        }
      }
    } else {
      float v[8][4];
#pragma unroll
      for (int ci = 0; ci < 8; ++ci) {
        const int c = c0 + warp + 8 * ci;
#pragma unroll
        for (int pi = 0; pi < 4; ++pi) v[ci][pi] = (soff[pi] != PAST_END && c < C) ? s[soff[pi] + (uint32_t)(c * HW)] : 0.f;
      }
      if (GATE) {
#pragma unroll
        for (int ci = 0; ci < 8; ++ci) {
          const int c = c0 + warp + 8 * ci;
#pragma unroll
          for (int pi = 0; pi < 4; ++pi) {
            const float g = (soff[pi] != PAST_END && c < C) ? gate[soff[pi] + (uint32_t)(c * HW)] : 0.f;
            v[ci][pi] *= g > 0.f ? 1.f : 0.f;
          }
        }
      }
#pragma unroll
      for (int ci = 0; ci < 8; ++ci) {
#pragma unroll
        for (int pi = 0; pi < 4; ++pi) tile[warp + 8 * ci][lane + 32 * pi] = __float_as_uint(v[ci][pi]);  // rounded at the store
        acc[ci] += (v[ci][0] + v[ci][1]) + (v[ci][2] + v[ci][3]);
      }
      __syncthreads();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = c0 + lane + 32 * h;
        if (c < Cp) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int pp = warp + 8 * i;
            const int64_t q = p0 + pp;
            if (q < Q) {
              const float a = __uint_as_float(tile[lane + 32 * h][pp]), hi = round_tf32(a);
              d[q * Cp + c] = __float_as_uint(hi);
              if (dst_lo) dst_lo[q * Cp + c] = round_tf32(a - hi);   // FP32X3: a - hi is exact in fp32
            }
          }
        }
      }
    }
    __syncthreads();  // tile is reused by the next pixel tile
  }

  if (chan_sum || partial) {  // warp-shuffle tree, then one value per (block, channel)
    const int64_t blk = blockIdx.x;  // row of the partial matrix [gx][groups*64]
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float t = warp_sum(acc[i]);
      // BF16: acc[2ci], acc[2ci+1] <-> channels c0 + 2(warp + 8ci) + {0,1};  TF32: acc[ci] <-> channel c0 + warp + 8ci
      const int c = BF16 ? c0 + 2 * (warp + 8 * (i >> 1)) + (i & 1) : c0 + warp + 8 * i;
      if (lane == 0 && c < C) {
        if (partial) partial[blk * ((int64_t)gridDim.y * CL_CH) + c] = t;  // deterministic: reduced in fixed order below
        else atomicAdd(chan_sum + c, t);
      }
    }
  }
}

// out[c] = Σ_rows partial[row][c]: block per 32 channels, 32 warps split the rows (lanes = channels: 128 B coalesced),
// then the 32 warp partials are added in fixed order -> deterministic
__global__ void __launch_bounds__(1024) chan_partial_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out, int C,
                                                                  int64_t rows, int pitch) {
  __shared__ float sm[32][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float s0 = 0.f, s1 = 0.f;
  if (c < C) {
    int64_t r = warp;
    for (; r + 32 < rows; r += 64) {
      s0 += partial[r * pitch + c];
      s1 += partial[(r + 32) * pitch + c];
    }
    if (r < rows) s0 += partial[r * pitch + c];
  }
  sm[warp][lane] = s0 + s1;
  __syncthreads();
  if (warp == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 32; ++w) t += sm[w][lane];
    out[c] = t;
  }
}

// w (Co, Ci, K, K) fp32 -> fprop weight matrix [Co][T][Ck] (zero padded, Ck = round_up(Ci, KC))
// dst_lo (FP32X3 only): the tf32 lo plane, same layout
template <bool BF16>
__global__ void w_fprop_kernel(const float* __restrict__ w, void* __restrict__ dst, int Co, int Ci, int T, int Ck, float* __restrict__ dst_lo) {
  const int64_t n = (int64_t)Co * T * Ck;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Ck);
    const int64_t r = i / Ck;
    const int tap = (int)(r % T), co = (int)(r / T);
    const float v = c < Ci ? w[((int64_t)co * Ci + c) * T + tap] : 0.f;
    if (BF16) reinterpret_cast<__nv_bfloat16*>(dst)[i] = __float2bfloat16_rn(v);
    else {
      const float hi = round_tf32(v);
      reinterpret_cast<float*>(dst)[i] = hi;
      if (dst_lo) dst_lo[i] = round_tf32(v - hi);
    }
  }
}
// dgrad weight matrix [Ci][nt][Cok] for the taps of one stride class: w'[ci][t][co] = w[co][ci][tap_idx[t]]
struct TapIdx { unsigned char idx[64]; };
template <bool BF16>
__global__ void w_dgrad_kernel(const float* __restrict__ w, void* __restrict__ dst, int Co, int Ci, int T, int nt, int Cok,
                               const TapIdx taps, float* __restrict__ dst_lo) {
  const int64_t n = (int64_t)Ci * nt * Cok;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cok);
    const int64_t r = i / Cok;
    const int tap = (int)(r % nt), ci = (int)(r / nt);
    const float v = co < Co ? w[((int64_t)co * Ci + ci) * T + taps.idx[tap]] : 0.f;
    if (BF16) reinterpret_cast<__nv_bfloat16*>(dst)[i] = __float2bfloat16_rn(v);
    else {
      const float hi = round_tf32(v);
      reinterpret_cast<float*>(dst)[i] = hi;
      if (dst_lo) dst_lo[i] = round_tf32(v - hi);
    }
  }
}
// dw[co][ci][tap] = Σ_split partial[split][co][tap][ci]   (fixed order)
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int Co, int Ci, int T, int splits) {
  const int64_t n = (int64_t)Co * Ci * T;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Ci);
    const int64_t r = i / Ci;
    const int tap = (int)(r % T), co = (int)(r / T);
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += partial[(int64_t)k * n + i];
    dw[((int64_t)co * Ci + ci) * T + tap] = s;
  }
}
// Same reduction, one block per output channel: partial[k][co] is a contiguous [tap][ci] matrix and dw[co] a contiguous
// [ci][tap] one, so the split sum is read coalesced, transposed through shared memory (odd T: conflict-free) and written
// coalesced — no per-element 64-bit divisions, no stride-T scatter.  Needs Ci*T floats of dynamic shared memory.
__global__ void __launch_bounds__(256) wgrad_reduce_rows_kernel(const float* __restrict__ partial, float* __restrict__ dw, int Co,
                                                                int Ci, int T, int splits) {
  extern __shared__ float row[];  // [ci][tap]
  const int co = blockIdx.x, n_row = Ci * T;
  const int64_t n = (int64_t)Co * n_row;
  const float* src = partial + (int64_t)co * n_row;
  for (int idx = threadIdx.x; idx < n_row; idx += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += src[(int64_t)k * n + idx];
    const int tap = idx / Ci, ci = idx - tap * Ci;
    row[ci * T + tap] = s;
  }
  __syncthreads();
  float* dst = dw + (int64_t)co * n_row;
  for (int idx = threadIdx.x; idx < n_row; idx += blockDim.x) dst[idx] = row[idx];
}
static int launch_wgrad_reduce(const float* partial, float* dw, int Co, int Ci, int T, int splits, cudaStream_t st) {
  const size_t smem = (size_t)Ci * T * sizeof(float);
  if (smem <= 48 * 1024 && Co >= 256) {  // one block per output channel: needs enough channels to fill the GPU
    wgrad_reduce_rows_kernel<<<Co, 256, smem, st>>>(partial, dw, Co, Ci, T, splits);
  } else {
    const int64_t n = (int64_t)Co * Ci * T;
    wgrad_reduce_kernel<<<ew_grid(n, 256), 256, 0, st>>>(partial, dw, Co, Ci, T, splits);
  }
  CPT_LAUNCH_CHECK("wgrad_reduce");
  return CPT_OK;
}
// [R][C] fp32 -> [R][Cp] bf16 (zero padded columns)
__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t R, int C, int Cp) {
  const int64_t n = R * (Cp / 2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / (Cp / 2);
    const int c = (int)(i - r * (Cp / 2)) * 2;
    const float v0 = c < C ? src[r * C + c] : 0.f, v1 = (c + 1) < C ? src[r * C + c + 1] : 0.f;
    *reinterpret_cast<__nv_bfloat162*>(dst + r * Cp + c) = __floats2bfloat162_rn(v0, v1);
  }
}

// [R][C] fp32 -> tf32 hi / lo planes [R][C] each (FP32X3 operands of the Linear GEMMs); C % 4 == 0, 16-byte aligned
__global__ void split_tf32_kernel(const float4* __restrict__ src, float4* __restrict__ hi, float4* __restrict__ lo, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = src[i];
    float4 h, l;
    h.x = round_tf32(v.x); h.y = round_tf32(v.y); h.z = round_tf32(v.z); h.w = round_tf32(v.w);
    l.x = round_tf32(v.x - h.x); l.y = round_tf32(v.y - h.y); l.z = round_tf32(v.z - h.z); l.w = round_tf32(v.w - h.w);
    hi[i] = h; lo[i] = l;
  }
}

// ------------------------------------------------------------------ kernel launch
__device__ int g_tc_status = 0;
// SMs left free by the persistent tensor-core kernels (cpt_tc_reserve_sms): room for a concurrent NCCL all-reduce in
// data-parallel backward passes.  A persistent grid that owns every SM (and nearly every register) serialises with it.
static int g_reserved_sms = 0;

// 2-CTA (cta_group::2) is used when the tile grid is large enough to keep all SM pairs busy; CPT_TC_2CTA=0/1 overrides.
// The drivers ask this BEFORE building the B tensor map (its box covers BN/2 columns per CTA in 2-CTA mode).
static bool want_2cta(int BN, int64_t m_tiles128) {
  static int forced = -2;
  if (forced == -2) {
    const char* e = getenv("CPT_TC_2CTA");
    forced = e ? atoi(e) : -1;
  }
  if (BN < 128 || m_tiles128 < 2) return false;
  if (forced >= 0) return forced != 0;
  return true;  // measured: C=512 fprop 2.85 -> 2.65 ms, never slower for BN >= 128
}

// 64-column tiles of the K-major kernels (convolution fprop / dgrad with <= 64 output channels).  Such a tile is bound by the
// ONE thread that issues its MMAs (~450 cycles per filter tap, twice what the tensor pipe needs); with
// cta_group::2 the same instruction stream drives the tensor cores of two SMs (a 256 x 64 tile), halving the issue cost per pixel.
static bool want_2cta_kmajor(int BN, int64_t m_tiles128, int mode) {
  if (BN >= 128) return want_2cta(BN, m_tiles128);
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("CPT_TC_2CTA_64");
    on = e ? atoi(e) : 1;
  }
  return on && BN == 64 && m_tiles128 >= 2 && !is_x3(mode);
}

template <bool A_MN, bool B_MN, int OP>
static int launch_bn(TcParams p, int mode, int BN, bool use2, cudaStream_t st) {
  if (use2) p.m_tiles = (p.m_tiles + 1) / 2;  // 256-row tiles
  LaunchSel sel{A_MN, B_MN, OP, BN, use2, (sm_count() - g_reserved_sms) / (use2 ? 2 : 1)};
  if (mode == CPT_MODE_BF16) return launch_bf16(p, sel, st);
  if (is_x3(mode)) return launch_x3(p, sel, st);
  return launch_tf32(p, sel, st);
}

static int pick_bn(int64_t n) { return n > 128 ? 256 : (n > 64 ? 128 : 64); }
// FP32X3 accumulates the tile in registers (chunked reduction): at most 128 columns
static int pick_bn_mode(int64_t n, int mode) { const int bn = pick_bn(n); return (is_x3(mode) && bn > 128) ? 128 : bn; }

static int get_status_ptr(int** ptr) {
  CPT_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(ptr), g_tc_status));
  return CPT_OK;
}

// splits so that tiles*splits fills whole waves of the persistent grid; every split keeps >= min_iters iterations
static int pick_splits(int tiles, int k_iters, int min_iters) {
  const int sms = sm_count();
  int best = 1;
  double best_eff = 0.0;
  {  // a split costs a partial write + a reduce pass: not worth it when one pass already fills >= 80 % of its waves
    const int waves1 = (tiles + sms - 1) / sms;
    if ((double)tiles / ((double)waves1 * sms) >= 0.8) return 1;
  }
  for (int s = 1; s <= 64; ++s) {
    if (s > 1 && k_iters / s < min_iters) break;
    const int total = tiles * s;
    const int waves = (total + sms - 1) / sms;
    const double eff = (double)total / ((double)waves * sms);
    if (eff > best_eff + 0.02) { best_eff = eff; best = s; }
  }
  return best;
}

// ------------------------------------------------------------------ Conv2D
struct G {
  int B, Ci, H, W, Co, K, P, S, D, Ho, Wo, T;
};
static G geom(const cpt_conv2d_desc* d) {
  G g{d->B, d->Ci, d->H, d->W, d->Co, d->K, d->pad, d->stride, d->dil, 0, 0, d->K * d->K};
  const int keff = d->dil * (d->K - 1) + 1;
  g.Ho = (d->H + 2 * d->pad - keff) / d->stride + 1;
  g.Wo = (d->W + 2 * d->pad - keff) / d->stride + 1;
  return g;
}
// one plane of a channels-last activation tensor / of a re-laid-out filter matrix; FP32X3 tensors hold two (hi, then lo)
static size_t cl_plane_bytes(int B, int C, int H, int W, int mode) {
  return align_up((size_t)B * H * W * round_up(C, 8) * esize(mode), 1024);
}
static size_t cl_bytes(int B, int C, int H, int W, int mode) { return planes(mode) * cl_plane_bytes(B, C, H, W, mode); }
static size_t wmat_plane_bytes(int rows, int T, int C, int mode) {
  return align_up((size_t)rows * T * round_up(C, kc_of(mode)) * esize(mode), 1024);
}
static size_t wmat_bytes(int rows, int T, int C, int mode) { return planes(mode) * wmat_plane_bytes(rows, T, C, mode); }
static inline const void* lo_plane(const void* base, size_t plane_bytes, int mode) {
  return is_x3(mode) ? reinterpret_cast<const char*>(base) + plane_bytes : nullptr;
}
static bool dgrad_tc_ok(const G& g);

static int wgrad_splits(const G& g, int mode) {
  const int bk = kc_of(mode);
  const int k_iters = (int)(((int64_t)g.B * g.Ho * g.Wo + bk - 1) / bk);
  const int tiles = ((g.Ci + 127) / 128) * ((g.Co + pick_bn_mode(g.Co, mode) - 1) / pick_bn_mode(g.Co, mode)) * g.T;
  return pick_splits(tiles, k_iters, 8);
}

static void cl_grid(int B, int C, int H, int W, int& gx, int& groups) {
  const int Cp = round_up(C, 8);
  // each block walks several 128-pixel tiles of its 64-channel group (tiles run over the flattened (image, pixel) index) so
  // the per-channel partial sums are reduced once per block; ~16 blocks per SM overall (several waves)
  const int64_t n_tiles = ((int64_t)B * H * W + CL_PX - 1) / CL_PX;
  groups = (Cp + CL_CH - 1) / CL_CH;
  int64_t want = (16LL * sm_count() + groups - 1) / groups;
  if (want < 1) want = 1;
  if (want > n_tiles) want = n_tiles;
  gx = (int)want;
}

size_t to_channels_last_ws(int B, int C, int H, int W) {
  int gx, groups;
  cl_grid(B, C, H, W, gx, groups);
  return align_up((size_t)gx * groups * CL_CH * sizeof(float), 256);
}

// chan_sum != NULL: per-channel sums of src.  With a workspace they are reduced deterministically (per-block partials +
// fixed-order pass) and chan_sum is overwritten; without one they are atomically accumulated into chan_sum (pre-zeroed).
int to_channels_last(const float* src, void* dst, int B, int C, int H, int W, int mode, float* chan_sum, void* ws, size_t ws_bytes,
                     cudaStream_t st, const float* gate = nullptr) {
  const int Cp = round_up(C, 8), HW = H * W;
  int gx, groups;
  cl_grid(B, C, H, W, gx, groups);
  dim3 grid(gx, groups, 1);
  CPT_REQUIRE(grid.y <= 65535 && (int64_t)B * HW < (1LL << 31) && (int64_t)B * C * HW < (1LL << 32) - (1 << 20),
              CPT_ERR_UNSUPPORTED, "to_channels_last: tensor too large");
  float* partial = nullptr;
  if (chan_sum && ws && ws_bytes >= to_channels_last_ws(B, C, H, W)) partial = reinterpret_cast<float*>(ws);
  const int64_t Q = (int64_t)B * HW;
  float* dst_lo = is_x3(mode) ? reinterpret_cast<float*>(reinterpret_cast<char*>(dst) + cl_plane_bytes(B, C, H, W, mode)) : nullptr;
  if (gate) {
    if (mode == CPT_MODE_BF16)
      nchw_to_nhwc_kernel<true, false, true><<<grid, 256, 0, st>>>(src, dst, C, HW, Cp, chan_sum, partial, Q, nullptr, PadGeom{0, 0, 0}, gate);
    else
      nchw_to_nhwc_kernel<false, false, true><<<grid, 256, 0, st>>>(src, dst, C, HW, Cp, chan_sum, partial, Q, dst_lo, PadGeom{0, 0, 0}, gate);
  } else if (mode == CPT_MODE_BF16) nchw_to_nhwc_kernel<true><<<grid, 256, 0, st>>>(src, dst, C, HW, Cp, chan_sum, partial, Q, nullptr);
  else nchw_to_nhwc_kernel<false><<<grid, 256, 0, st>>>(src, dst, C, HW, Cp, chan_sum, partial, Q, dst_lo);
  CPT_LAUNCH_CHECK("nchw_to_nhwc");
  if (partial) {
    chan_partial_reduce_kernel<<<(C + 31) / 32, 1024, 0, st>>>(partial, chan_sum, C, (int64_t)gx, groups * CL_CH);
    CPT_LAUNCH_CHECK("chan_partial_reduce");
  }
  return CPT_OK;
}

#include "fused_cl.cuh"

// Geometry of one implicit-GEMM launch over a channels-last activation tensor.
struct ConvPlan {
  int ntaps;
  unsigned short off_w[64], off_h[64];  // im2col offsets per tap (>= 0, relative to the lower corner)
  int lower_w, lower_h, upper_w, upper_h, trav;  // bounding box of the base pixel, traversal stride
  int sub_H, sub_W;                     // grid of base pixels == GEMM lanes per image
  int out_H, out_W, out_s, out_r0, out_c0;  // where lane (r, c) lands in the NCHW output plane
};

// out[b, n, out_r0 + out_s*r, out_c0 + out_s*c] = Σ_{t, ch} act_cl[b, lower_h + r*trav + off_h[t], lower_w + c*trav + off_w[t], ch]
//                                                          * wmat[n][t][ch]     (+ bias[n])
static size_t stats_bytes(int Ncols) { return (size_t)sm_count() * 4 * Ncols * 2 * sizeof(float); }

// FP32X3: act_cl holds the hi and lo planes back to back (cl_bytes layout); wmat_lo is the lo plane of the filter matrix
static int conv_im2col_gemm(const void* act_cl, int B, int Cact, int Hin, int Win, const void* wmat, const void* wmat_lo, int Ncols,
                            const ConvPlan& pl, const float* bias, float* out, int mode, cudaStream_t st, float* stats = nullptr,
                            bool relu = false) {
  const int kc = kc_of(mode), Cp = round_up(Cact, 8), Ck = round_up(Cact, kc), T = pl.ntaps;
  const int BN = pick_bn_mode(Ncols, mode);
  TcParams p{};
  const int64_t M = (int64_t)B * pl.sub_H * pl.sub_W;
  CPT_REQUIRE(M < (1LL << 31), CPT_ERR_UNSUPPORTED, "conv: pixel count exceeds int32");
  const bool use2 = want_2cta_kmajor(BN, (M + 127) / 128, mode);
  if (int e = make_map_im2col(&p.tmA, act_cl, mode, Cp, Win, Hin, B, pl.lower_w, pl.lower_h, pl.upper_w, pl.upper_h, pl.trav, kc, 128)) return e;
  if (int e = make_map_2d(&p.tmB, wmat, mode, (uint64_t)T * Ck, Ncols, (uint64_t)T * Ck, kc, use2 ? BN / 2 : BN)) return e;
  if (is_x3(mode)) {
    const void* act_lo = lo_plane(act_cl, cl_plane_bytes(B, Cact, Hin, Win, mode), mode);
    if (int e = make_map_im2col(&p.tmA2, act_lo, mode, Cp, Win, Hin, B, pl.lower_w, pl.lower_h, pl.upper_w, pl.upper_h, pl.trav, kc, 128)) return e;
    if (int e = make_map_2d(&p.tmB2, wmat_lo, mode, (uint64_t)T * Ck, Ncols, (uint64_t)T * Ck, kc, use2 ? BN / 2 : BN)) return e;
  }
  p.out = out;
  p.bias = bias;
  p.bias_mode = bias ? BIAS_COL : BIAS_NONE;
  p.relu = relu ? 3 : 0;
  if (stats) {  // slots of CTAs / N-tiles that never run stay zero
    CPT_CUDA(cudaMemsetAsync(stats, 0, stats_bytes(Ncols), st));
    p.stats = stats;
  }
  if (int e = get_status_ptr(&p.status)) return e;
  p.M = (int)M;
  p.N = Ncols;
  p.m_tiles = (int)((M + 127) / 128);
  p.n_tiles = (Ncols + BN - 1) / BN;
  p.z_tiles = 1;
  p.cchunks = Ck / kc;
  p.k_iters_total = T * p.cchunks;
  p.k_iters_per_split = p.k_iters_total;
  p.col_stride = (long long)pl.out_H * pl.out_W;
  p.lane_is_pixel = 1;
  p.px_per_img = pl.sub_H * pl.sub_W;
  p.img_stride = (long long)Ncols * pl.out_H * pl.out_W;
  p.out_W = pl.out_W; p.out_s = pl.out_s; p.out_r0 = pl.out_r0; p.out_c0 = pl.out_c0;
  p.Wo = pl.sub_W;
  p.trav = pl.trav; p.lower_w = pl.lower_w; p.lower_h = pl.lower_h;
  p.taps = T;
  p.wk_cols = Ck;
  for (int t = 0; t < T; ++t) { p.tap_w[t] = pl.off_w[t]; p.tap_h[t] = pl.off_h[t]; }
  return launch_bn<false, false, OP_CONV>(p, mode, BN, use2, st);
}

static bool fprop_tc_ok(const G& g) {
  const int upper = g.P - (g.K - 1) * g.D;
  return g.T <= 64 && g.P <= 128 && upper >= -128 && upper <= 127 && (g.K - 1) * g.D <= 255 && g.S <= 8;
}

int conv_fprop_cl(const cpt_conv2d_desc* d, const void* x_cl, const float* w, const float* bias, float* y, int mode, void* ws,
                  size_t ws_bytes, cudaStream_t st, float* stats = nullptr, bool relu = false) {
  const G g = geom(d);
  CPT_REQUIRE(fprop_tc_ok(g), CPT_ERR_UNSUPPORTED, "conv2d_fprop_cl: kernel %d / padding %d / dilation %d outside the TMA im2col limits", g.K, g.P, g.D);
  const size_t need = wmat_bytes(g.Co, g.T, g.Ci, mode);
  CPT_REQUIRE(ws && ws_bytes >= need, CPT_ERR_WORKSPACE, "conv2d_fprop_cl: workspace too small (%zu < %zu)", ws_bytes, need);
  const int Ck = round_up(g.Ci, kc_of(mode));
  const int64_t n = (int64_t)g.Co * g.T * Ck;
  float* w_lo = const_cast<float*>(reinterpret_cast<const float*>(lo_plane(ws, wmat_plane_bytes(g.Co, g.T, g.Ci, mode), mode)));
  if (mode == CPT_MODE_BF16) w_fprop_kernel<true><<<ew_grid(n, 256), 256, 0, st>>>(w, ws, g.Co, g.Ci, g.T, Ck, nullptr);
  else w_fprop_kernel<false><<<ew_grid(n, 256), 256, 0, st>>>(w, ws, g.Co, g.Ci, g.T, Ck, w_lo);
  CPT_LAUNCH_CHECK("w_fprop");
  ConvPlan pl{};
  pl.ntaps = g.T;
  for (int j = 0; j < g.K; ++j)
    for (int kk = 0; kk < g.K; ++kk) { pl.off_h[j * g.K + kk] = (unsigned short)(j * g.D); pl.off_w[j * g.K + kk] = (unsigned short)(kk * g.D); }
  pl.lower_w = pl.lower_h = -g.P;
  // upper corner = P - (K-1) D (same rule as CUTLASS detail.hpp compute_upper_corner_whd, fprop): base pixels run from
  // -P to extent-1+upper in steps of S, i.e. exactly Wo (Ho) positions per row (column)
  pl.upper_w = pl.upper_h = g.P - (g.K - 1) * g.D;
  pl.trav = g.S;
  pl.sub_H = g.Ho; pl.sub_W = g.Wo;
  pl.out_H = g.Ho; pl.out_W = g.Wo; pl.out_s = 1; pl.out_r0 = pl.out_c0 = 0;
  return conv_im2col_gemm(x_cl, g.B, g.Ci, g.H, g.W, ws, w_lo, g.Co, pl, bias, y, mode, st, stats, relu);
}

// taps (j, kk) of stride class (rh, rw) — same rule as the exact path (conv.cu class_taps)
static int tc_class_taps(const G& g, int rh, int rw, int* tj, int* tk) {
  int n = 0;
  for (int j = 0; j < g.K; ++j) {
    if (((rh + g.P - j * g.D) % g.S + g.S) % g.S != 0) continue;
    for (int kk = 0; kk < g.K; ++kk) {
      if (((rw + g.P - kk * g.D) % g.S + g.S) % g.S != 0) continue;
      if (n < 64) { tj[n] = j; tk[n] = kk; }
      ++n;
    }
  }
  return n;
}

static inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// Plans dgrad for one stride class; returns false if it cannot run on the im2col path.
static bool dgrad_class_plan(const G& g, int rh, int rw, ConvPlan& pl, TapIdx& ti) {
  int tj[64], tk[64];
  const int nt = tc_class_taps(g, rh, rw, tj, tk);
  if (nt > 64) return false;
  pl = ConvPlan{};
  pl.ntaps = nt;
  pl.sub_H = (g.H - rh + g.S - 1) / g.S;
  pl.sub_W = (g.W - rw + g.S - 1) / g.S;
  pl.out_H = g.H; pl.out_W = g.W; pl.out_s = g.S; pl.out_r0 = rh; pl.out_c0 = rw;
  pl.trav = 1;
  if (nt == 0) return true;
  // dy row read by tap j for sub-grid row hs: p = hs + (rh + P - j*D) / S  (exact division inside a class)
  int lo_h = 1 << 30, lo_w = 1 << 30, hi_h = -(1 << 30), hi_w = -(1 << 30);
  for (int t = 0; t < nt; ++t) {
    const int oh = floordiv(rh + g.P - tj[t] * g.D, g.S), ow = floordiv(rw + g.P - tk[t] * g.D, g.S);
    lo_h = oh < lo_h ? oh : lo_h; hi_h = oh > hi_h ? oh : hi_h;
    lo_w = ow < lo_w ? ow : lo_w; hi_w = ow > hi_w ? ow : hi_w;
  }
  for (int t = 0; t < nt; ++t) {
    pl.off_h[t] = (unsigned short)(floordiv(rh + g.P - tj[t] * g.D, g.S) - lo_h);
    pl.off_w[t] = (unsigned short)(floordiv(rw + g.P - tk[t] * g.D, g.S) - lo_w);
    ti.idx[t] = (unsigned char)(tj[t] * g.K + tk[t]);
  }
  pl.lower_h = lo_h; pl.lower_w = lo_w;
  // base pixels lower .. lower + sub - 1 over a (Ho, Wo) tensor: upper = lower + sub - extent
  pl.upper_h = lo_h + pl.sub_H - g.Ho;
  pl.upper_w = lo_w + pl.sub_W - g.Wo;
  auto in8 = [](int v) { return v >= -128 && v <= 127; };
  return in8(pl.lower_h) && in8(pl.lower_w) && in8(pl.upper_h) && in8(pl.upper_w) && hi_h - lo_h <= 255 && hi_w - lo_w <= 255;
}

static bool dgrad_tc_ok(const G& g) {
  if (g.S > 8) return false;
  for (int c = 0; c < g.S * g.S; ++c) {
    const int rh = c / g.S, rw = c % g.S;
    if (rh >= g.H || rw >= g.W) continue;
    ConvPlan pl; TapIdx ti;
    if (!dgrad_class_plan(g, rh, rw, pl, ti)) return false;
  }
  return true;
}

int conv_dgrad_cl(const cpt_conv2d_desc* d, const void* dy_cl, const float* w, float* dx, int mode, void* ws, size_t ws_bytes,
                  cudaStream_t st) {
  const G g = geom(d);
  CPT_REQUIRE(dgrad_tc_ok(g), CPT_ERR_UNSUPPORTED, "conv2d_dgrad_cl: geometry (K=%d, stride=%d, pad=%d, dil=%d) outside the TMA im2col limits",
              g.K, g.S, g.P, g.D);
  const int Cok = round_up(g.Co, kc_of(mode));
  const size_t per_plane = align_up((size_t)g.Ci * g.T * Cok * esize(mode), 1024), per_class = per_plane * planes(mode);
  const size_t need = per_class * g.S * g.S;
  CPT_REQUIRE(ws && ws_bytes >= need, CPT_ERR_WORKSPACE, "conv2d_dgrad_cl: workspace too small (%zu < %zu)", ws_bytes, need);
  bool any_empty = false;
  for (int c = 0; c < g.S * g.S; ++c) {
    const int rh = c / g.S, rw = c % g.S;
    if (rh >= g.H || rw >= g.W) continue;
    ConvPlan pl; TapIdx ti;
    dgrad_class_plan(g, rh, rw, pl, ti);
    if (pl.ntaps == 0) any_empty = true;
  }
  if (any_empty) CPT_CUDA(cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)g.B * g.Ci * g.H * g.W, st));
  for (int c = 0; c < g.S * g.S; ++c) {
    const int rh = c / g.S, rw = c % g.S;
    if (rh >= g.H || rw >= g.W) continue;
    ConvPlan pl; TapIdx ti{};
    dgrad_class_plan(g, rh, rw, pl, ti);
    if (pl.ntaps == 0) continue;
    void* wm = reinterpret_cast<char*>(ws) + per_class * c;
    const int64_t n = (int64_t)g.Ci * pl.ntaps * Cok;
    float* wm_lo = const_cast<float*>(reinterpret_cast<const float*>(lo_plane(wm, per_plane, mode)));
    if (mode == CPT_MODE_BF16) w_dgrad_kernel<true><<<ew_grid(n, 256), 256, 0, st>>>(w, wm, g.Co, g.Ci, g.T, pl.ntaps, Cok, ti, nullptr);
    else w_dgrad_kernel<false><<<ew_grid(n, 256), 256, 0, st>>>(w, wm, g.Co, g.Ci, g.T, pl.ntaps, Cok, ti, wm_lo);
    CPT_LAUNCH_CHECK("w_dgrad");
    if (int e = conv_im2col_gemm(dy_cl, g.B, g.Co, g.Ho, g.Wo, wm, wm_lo, g.Ci, pl, nullptr, dx, mode, st)) return e;
  }
  return CPT_OK;
}

int conv_wgrad_cl(const cpt_conv2d_desc* d, const void* x_cl, const void* dy_cl, float* dw, int mode, void* ws, size_t ws_bytes,
                  cudaStream_t st) {
  const G g = geom(d);
  const int kc = kc_of(mode), bk = kc;
  const int splits_req = wgrad_splits(g, mode);
  const int64_t pixels = (int64_t)g.B * g.Ho * g.Wo;
  const int k_iters = (int)((pixels + bk - 1) / bk);
  const int kps = (k_iters + splits_req - 1) / splits_req;
  const int splits = (k_iters + kps - 1) / kps;  // no empty split
  const size_t need = align_up((size_t)splits * g.Co * g.T * g.Ci * sizeof(float), 1024);
  CPT_REQUIRE(ws && ws_bytes >= need, CPT_ERR_WORKSPACE, "conv2d_wgrad_cl: workspace too small (%zu < %zu)", ws_bytes, need);
  const int BN = pick_bn_mode(g.Co, mode);
  TcParams p{};
  const int upper_w = g.P - (g.K - 1) * g.D, upper_h = upper_w;
  // A: x_cl through im2col, lanes = input channels (MN-major), reduction = output pixels
  if (int e = make_map_im2col(&p.tmA, x_cl, mode, round_up(g.Ci, 8), g.W, g.H, g.B, -g.P, -g.P, upper_w, upper_h, g.S, kc, bk, true)) return e;
  // B: dy_cl as [pixels][Cop], columns = output channels (MN-major)
  const int Cop = round_up(g.Co, 8);
  if (int e = make_map_2d(&p.tmB, dy_cl, mode, Cop, (uint64_t)pixels, Cop, kc, bk, true)) return e;
  if (is_x3(mode)) {
    const void* x_lo = lo_plane(x_cl, cl_plane_bytes(g.B, g.Ci, g.H, g.W, mode), mode);
    const void* dy_lo = lo_plane(dy_cl, cl_plane_bytes(g.B, g.Co, g.Ho, g.Wo, mode), mode);
    if (int e = make_map_im2col(&p.tmA2, x_lo, mode, round_up(g.Ci, 8), g.W, g.H, g.B, -g.P, -g.P, upper_w, upper_h, g.S, kc, bk, true)) return e;
    if (int e = make_map_2d(&p.tmB2, dy_lo, mode, Cop, (uint64_t)pixels, Cop, kc, bk, true)) return e;
  }
  p.out = reinterpret_cast<float*>(ws);
  p.bias = nullptr;
  p.bias_mode = BIAS_NONE;
  if (int e = get_status_ptr(&p.status)) return e;
  p.M = g.Ci;
  p.N = g.Co;
  p.m_tiles = (g.Ci + 127) / 128;
  p.n_tiles = (g.Co + BN - 1) / BN;
  p.z_tiles = g.T * splits;
  p.k_iters_total = k_iters;
  p.k_iters_per_split = kps;
  p.col_stride = (long long)g.T * g.Ci;
  p.split_stride = (long long)g.Co * g.T * g.Ci;
  p.tap_stride = g.Ci;
  p.lane_is_pixel = 0;
  p.px_per_img = g.Ho * g.Wo;
  p.Wo = g.Wo;
  p.Ho = g.Ho;
  p.conv_stride = g.S;
  p.pad = g.P;
  p.dil = g.D;
  p.Kw = g.K;
  p.taps = g.T;
  p.out_s = 1;
  if (int e = launch_bn<true, true, OP_WGRAD>(p, mode, BN, want_2cta(BN, (g.Ci + 127) / 128), st)) return e;
  return launch_wgrad_reduce(reinterpret_cast<float*>(ws), dw, g.Co, g.Ci, g.T, splits, st);
}

size_t conv_workspace_size(int op, const cpt_conv2d_desc* d, int mode) {
  const G g = geom(d);
  if (op == CPT_OP_FPROP) return cl_bytes(g.B, g.Ci, g.H, g.W, mode) + wmat_bytes(g.Co, g.T, g.Ci, mode) + 1024;
  if (op == CPT_OP_DGRAD) {
    if (!dgrad_tc_ok(g)) return 8192;  // exact path: tap tables of the stride classes
    return cl_bytes(g.B, g.Co, g.Ho, g.Wo, mode) + (size_t)g.S * g.S * wmat_bytes(g.Ci, g.T, g.Co, mode) + 2048;
  }
  size_t part = align_up((size_t)wgrad_splits(g, mode) * g.Co * g.T * g.Ci * sizeof(float), 1024);
  const size_t csum = to_channels_last_ws(g.B, g.Co, g.Ho, g.Wo);
  if (part < csum) part = csum;
  return cl_bytes(g.B, g.Ci, g.H, g.W, mode) + cl_bytes(g.B, g.Co, g.Ho, g.Wo, mode) + part + 1024;
}

int conv_fprop(const cpt_conv2d_desc* d, const float* x, const float* w, const float* bias, float* y, int mode, void* ws,
               size_t ws_bytes, cudaStream_t st) {
  const G g = geom(d);
  CPT_REQUIRE(ws && ws_bytes >= conv_workspace_size(CPT_OP_FPROP, d, mode), CPT_ERR_WORKSPACE, "conv2d_fprop: workspace too small");
  const size_t xb = cl_bytes(g.B, g.Ci, g.H, g.W, mode);
  if (int e = to_channels_last(x, ws, g.B, g.Ci, g.H, g.W, mode, nullptr, nullptr, 0, st)) return e;
  return conv_fprop_cl(d, ws, w, bias, y, mode, reinterpret_cast<char*>(ws) + xb, ws_bytes - xb, st);
}

int conv_dgrad(const cpt_conv2d_desc* d, const float* dy, const float* w, float* dx, int mode, void* ws, size_t ws_bytes,
               cudaStream_t st) {
  const G g = geom(d);
  if (!dgrad_tc_ok(g))  // strided dgrad: exact FFMA path (more accurate than the requested mode, never less)
    return cpt_conv2d_dgrad(d, dy, w, dx, CPT_MODE_FP32, ws, ws_bytes, st);
  CPT_REQUIRE(ws && ws_bytes >= conv_workspace_size(CPT_OP_DGRAD, d, mode), CPT_ERR_WORKSPACE, "conv2d_dgrad: workspace too small");
  const size_t yb = cl_bytes(g.B, g.Co, g.Ho, g.Wo, mode);
  if (int e = to_channels_last(dy, ws, g.B, g.Co, g.Ho, g.Wo, mode, nullptr, nullptr, 0, st)) return e;
  return conv_dgrad_cl(d, ws, w, dx, mode, reinterpret_cast<char*>(ws) + yb, ws_bytes - yb, st);
}

int conv_wgrad(const cpt_conv2d_desc* d, const float* x, const float* dy, float* dw, float* db, int mode, void* ws,
               size_t ws_bytes, cudaStream_t st) {
  const G g = geom(d);
  CPT_REQUIRE(ws && ws_bytes >= conv_workspace_size(CPT_OP_WGRAD, d, mode), CPT_ERR_WORKSPACE, "conv2d_wgrad: workspace too small");
  char* base = reinterpret_cast<char*>(ws);
  const size_t xb = cl_bytes(g.B, g.Ci, g.H, g.W, mode), yb = cl_bytes(g.B, g.Co, g.Ho, g.Wo, mode);
  if (int e = to_channels_last(x, base, g.B, g.Ci, g.H, g.W, mode, nullptr, nullptr, 0, st)) return e;
  // db fused into the staging pass of dy (deterministic partials in the part of ws that wgrad uses afterwards)
  if (int e = to_channels_last(dy, base + xb, g.B, g.Co, g.Ho, g.Wo, mode, db, base + xb + yb, ws_bytes - xb - yb, st)) return e;
  return conv_wgrad_cl(d, base, base + xb, dw, mode, base + xb + yb, ws_bytes - xb - yb, st);
}

// ------------------------------------------------------------------ Linear (lanes = the output's contiguous dim)
static size_t cast_bytes(int64_t R, int C) { return align_up((size_t)R * round_up(C, 8) * 2, 1024); }

static int cast_to_bf16(const float* src, void* dst, int64_t R, int C, cudaStream_t st) {
  const int Cp = round_up(C, 8);
  cast_bf16_kernel<<<ew_grid(R * (Cp / 2), 256), 256, 0, st>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), R, C, Cp);
  CPT_LAUNCH_CHECK("cast_bf16");
  return CPT_OK;
}

static bool tf32_direct_ok(const void* a, const void* b, int In, int Out) {
  return In % 4 == 0 && Out % 4 == 0 && aligned16(a) && aligned16(b);
}

// FP32X3: hi + lo tf32 planes of an [R][C] fp32 matrix (same pitch as the source)
static size_t split_bytes(int64_t R, int C) { return 2 * align_up((size_t)R * C * 4, 1024); }
static int split_to_tf32(const float* src, void* dst, int64_t R, int C, cudaStream_t st) {
  const int64_t n4 = R * C / 4;
  float4* hi = reinterpret_cast<float4*>(dst);
  float4* lo = reinterpret_cast<float4*>(reinterpret_cast<char*>(dst) + split_bytes(R, C) / 2);
  split_tf32_kernel<<<ew_grid(n4, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(src), hi, lo, n4);
  CPT_LAUNCH_CHECK("split_tf32");
  return CPT_OK;
}
static inline const void* split_lo(const void* hi, int64_t R, int C) { return reinterpret_cast<const char*>(hi) + split_bytes(R, C) / 2; }

size_t linear_workspace_size(int op, int64_t N, int In, int Out, int mode) {
  size_t s = 1024;
  if (mode == CPT_MODE_BF16) {
    if (op == CPT_OP_FPROP) s += cast_bytes(N, In) + cast_bytes(Out, In);
    else if (op == CPT_OP_DGRAD) s += cast_bytes(N, Out) + cast_bytes(Out, In);
    else s += cast_bytes(N, In) + cast_bytes(N, Out);
  } else if (is_x3(mode)) {
    if (op == CPT_OP_FPROP) s += split_bytes(N, In) + split_bytes(Out, In);
    else if (op == CPT_OP_DGRAD) s += split_bytes(N, Out) + split_bytes(Out, In);
    else s += split_bytes(N, In) + split_bytes(N, Out);
  }
  if (op == CPT_OP_WGRAD) {
    // split-K partials + db partials
    s += align_up((size_t)16 * Out * In * sizeof(float), 1024);  // up to 16 split-K partials
    s += align_up((size_t)Out * 64 * sizeof(float), 256);
  }
  return s;
}

// *_lo: the tf32 lo planes of the two operands (FP32X3 mode only)
int linear_fwd_lp(const void* xa, const void* wa, const float* bias, float* y, int64_t N, int In, int Out, int mode, cudaStream_t st,
                  int relu = 0, void* y_lp = nullptr, unsigned int* mask = nullptr, const void* xa_lo = nullptr, const void* wa_lo = nullptr);
int linear_dgrad_lp(const void* ga, const void* wa, float* dx, int64_t N, int In, int Out, int mode, cudaStream_t st,
                    const unsigned int* in_mask = nullptr, void* dx_lp = nullptr, const void* ga_lo = nullptr, const void* wa_lo = nullptr);
int linear_wgrad_lp(const void* xa, const void* ga, float* dw, int64_t N, int In, int Out, int mode, void* ws, size_t ws_bytes,
                    cudaStream_t st, int max_splits = 16, const void* xa_lo = nullptr, const void* ga_lo = nullptr);

int linear_fwd(const float* x, const float* w, const float* bias, float* y, int64_t N, int In, int Out, int mode, void* ws,
               size_t ws_bytes, cudaStream_t st) {
  const void* xa = x;
  const void* wa = w;
  if (mode == CPT_MODE_BF16) {
    CPT_REQUIRE(ws && ws_bytes >= linear_workspace_size(CPT_OP_FPROP, N, In, Out, mode), CPT_ERR_WORKSPACE, "linear_fwd: workspace too small");
    char* base = reinterpret_cast<char*>(ws);
    if (int e = cast_to_bf16(x, base, N, In, st)) return e;
    if (int e = cast_to_bf16(w, base + cast_bytes(N, In), Out, In, st)) return e;
    xa = base; wa = base + cast_bytes(N, In);
  } else if (!tf32_direct_ok(x, w, In, 4)) {
    return cpt_linear_fwd(x, w, bias, y, N, In, Out, CPT_MODE_FP32, ws, ws_bytes, st);
  } else if (is_x3(mode)) {
    CPT_REQUIRE(ws && ws_bytes >= linear_workspace_size(CPT_OP_FPROP, N, In, Out, mode), CPT_ERR_WORKSPACE, "linear_fwd: workspace too small");
    char* base = reinterpret_cast<char*>(ws);
    if (int e = split_to_tf32(x, base, N, In, st)) return e;
    if (int e = split_to_tf32(w, base + split_bytes(N, In), Out, In, st)) return e;
    xa = base; wa = base + split_bytes(N, In);
    return linear_fwd_lp(xa, wa, bias, y, N, In, Out, mode, st, 0, nullptr, nullptr, split_lo(xa, N, In), split_lo(wa, Out, In));
  }
  return linear_fwd_lp(xa, wa, bias, y, N, In, Out, mode, st);
}

// operands already in the mode's element type: bf16 with row pitch round_up(C, 8), or fp32 (tf32) with pitch C
int linear_fwd_lp(const void* xa, const void* wa, const float* bias, float* y, int64_t N, int In, int Out, int mode, cudaStream_t st,
                  int relu, void* y_lp, unsigned int* mask, const void* xa_lo, const void* wa_lo) {
  const int pitch = mode == CPT_MODE_BF16 ? round_up(In, 8) : In;
  const int kc = kc_of(mode), BN = pick_bn_mode(N, mode);
  TcParams p{};
  // y[n][o]: lanes = o.  A = w [Out][In] K-major, B = x [N][In] K-major
  if (int e = make_map_2d(&p.tmA, wa, mode, In, Out, pitch, kc, 128)) return e;
  const bool use2 = want_2cta(BN, (Out + 127) / 128);
  if (int e = make_map_2d(&p.tmB, xa, mode, In, (uint64_t)N, pitch, kc, use2 ? BN / 2 : BN)) return e;
  if (is_x3(mode)) {
    if (int e = make_map_2d(&p.tmA2, wa_lo, mode, In, Out, pitch, kc, 128)) return e;
    if (int e = make_map_2d(&p.tmB2, xa_lo, mode, In, (uint64_t)N, pitch, kc, use2 ? BN / 2 : BN)) return e;
  }
  p.out = y; p.bias = bias; p.bias_mode = bias ? BIAS_LANE : BIAS_NONE;
  p.relu = relu; p.relu_lp = y_lp; p.relu_mask = mask;
  if (int e = get_status_ptr(&p.status)) return e;
  p.M = Out; p.N = (int)N;
  p.m_tiles = (Out + 127) / 128; p.n_tiles = (int)((N + BN - 1) / BN); p.z_tiles = 1;
  p.k_iters_total = (In + kc - 1) / kc; p.k_iters_per_split = p.k_iters_total;
  p.col_stride = Out; p.taps = 1;
  return launch_bn<false, false, OP_GEMM>(p, mode, BN, use2, st);
}

int linear_dgrad(const float* dy, const float* w, float* dx, int64_t N, int In, int Out, int mode, void* ws, size_t ws_bytes,
                 cudaStream_t st) {
  const void* ga = dy;
  const void* wa = w;
  if (mode == CPT_MODE_BF16) {
    CPT_REQUIRE(ws && ws_bytes >= linear_workspace_size(CPT_OP_DGRAD, N, In, Out, mode), CPT_ERR_WORKSPACE, "linear_dgrad: workspace too small");
    char* base = reinterpret_cast<char*>(ws);
    if (int e = cast_to_bf16(dy, base, N, Out, st)) return e;
    if (int e = cast_to_bf16(w, base + cast_bytes(N, Out), Out, In, st)) return e;
    ga = base; wa = base + cast_bytes(N, Out);
  } else if (!tf32_direct_ok(dy, w, In, Out)) {
    return cpt_linear_dgrad(dy, w, dx, N, In, Out, CPT_MODE_FP32, ws, ws_bytes, st);
  } else if (is_x3(mode)) {
    CPT_REQUIRE(ws && ws_bytes >= linear_workspace_size(CPT_OP_DGRAD, N, In, Out, mode), CPT_ERR_WORKSPACE, "linear_dgrad: workspace too small");
    char* base = reinterpret_cast<char*>(ws);
    if (int e = split_to_tf32(dy, base, N, Out, st)) return e;
    if (int e = split_to_tf32(w, base + split_bytes(N, Out), Out, In, st)) return e;
    ga = base; wa = base + split_bytes(N, Out);
    return linear_dgrad_lp(ga, wa, dx, N, In, Out, mode, st, nullptr, nullptr, split_lo(ga, N, Out), split_lo(wa, Out, In));
  }
  return linear_dgrad_lp(ga, wa, dx, N, In, Out, mode, st);
}

int linear_dgrad_lp(const void* ga, const void* wa, float* dx, int64_t N, int In, int Out, int mode, cudaStream_t st,
                    const unsigned int* in_mask, void* dx_lp, const void* ga_lo, const void* wa_lo) {
  const int gpitch = mode == CPT_MODE_BF16 ? round_up(Out, 8) : Out, wpitch = mode == CPT_MODE_BF16 ? round_up(In, 8) : In;
  const int kc = kc_of(mode), BN = pick_bn_mode(N, mode);
  TcParams p{};
  // dx[n][i]: lanes = i.  A(m=i, k=o) = w[o][i]: MN-major over the [Out][In] matrix; B = dy [N][Out] K-major
  if (int e = make_map_2d(&p.tmA, wa, mode, In, Out, wpitch, kc, kc, true)) return e;
  const bool use2 = want_2cta(BN, (In + 127) / 128);
  if (int e = make_map_2d(&p.tmB, ga, mode, Out, (uint64_t)N, gpitch, kc, use2 ? BN / 2 : BN)) return e;
  if (is_x3(mode)) {
    if (int e = make_map_2d(&p.tmA2, wa_lo, mode, In, Out, wpitch, kc, kc, true)) return e;
    if (int e = make_map_2d(&p.tmB2, ga_lo, mode, Out, (uint64_t)N, gpitch, kc, use2 ? BN / 2 : BN)) return e;
  }
  p.out = dx; p.bias = nullptr; p.bias_mode = BIAS_NONE;
  if (in_mask) { p.relu = 2; p.relu_mask = const_cast<unsigned int*>(in_mask); p.relu_lp = dx_lp; }
  if (int e = get_status_ptr(&p.status)) return e;
  p.M = In; p.N = (int)N;
  p.m_tiles = (In + 127) / 128; p.n_tiles = (int)((N + BN - 1) / BN); p.z_tiles = 1;
  p.k_iters_total = (Out + kc - 1) / kc; p.k_iters_per_split = p.k_iters_total;
  p.col_stride = In; p.taps = 1;
  return launch_bn<true, false, OP_GEMM>(p, mode, BN, use2, st);
}

int linear_wgrad(const float* x, const float* dy, float* dw, float* db, int64_t N, int In, int Out, int mode, void* ws,
                 size_t ws_bytes, cudaStream_t st) {
  CPT_REQUIRE(ws && ws_bytes >= linear_workspace_size(CPT_OP_WGRAD, N, In, Out, mode), CPT_ERR_WORKSPACE, "linear_wgrad: workspace too small");
  char* base = reinterpret_cast<char*>(ws);
  const void* xa = x;
  const void* ga = dy;
  size_t off = 0;
  if (mode == CPT_MODE_BF16) {
    if (int e = cast_to_bf16(x, base, N, In, st)) return e;
    if (int e = cast_to_bf16(dy, base + cast_bytes(N, In), N, Out, st)) return e;
    xa = base; ga = base + cast_bytes(N, In);
    off = cast_bytes(N, In) + cast_bytes(N, Out);
  } else if (!tf32_direct_ok(x, dy, In, Out)) {
    return cpt_linear_wgrad(x, dy, dw, db, N, In, Out, CPT_MODE_FP32, ws, ws_bytes, st);
  } else if (is_x3(mode)) {
    if (int e = split_to_tf32(x, base, N, In, st)) return e;
    if (int e = split_to_tf32(dy, base + split_bytes(N, In), N, Out, st)) return e;
    xa = base; ga = base + split_bytes(N, In);
    off = split_bytes(N, In) + split_bytes(N, Out);
  }
  if (int e = linear_wgrad_lp(xa, ga, dw, N, In, Out, mode, base + off, ws_bytes - off, st, 16,
                              is_x3(mode) ? split_lo(xa, N, In) : nullptr, is_x3(mode) ? split_lo(ga, N, Out) : nullptr)) return e;
  if (db) {
    const size_t part = align_up((size_t)16 * Out * In * sizeof(float), 1024);
    return channel_sum(dy, db, (int)N, Out, 1, base + off + part, st);
  }
  return CPT_OK;
}

// ws: split-K partials (up to 16 x Out x In floats)
int linear_wgrad_lp(const void* xa, const void* ga, float* dw, int64_t N, int In, int Out, int mode, void* ws, size_t ws_bytes,
                    cudaStream_t st, int max_splits, const void* xa_lo, const void* ga_lo) {
  char* base = reinterpret_cast<char*>(ws);
  const size_t off = 0;
  const int xpitch = mode == CPT_MODE_BF16 ? round_up(In, 8) : In, gpitch = mode == CPT_MODE_BF16 ? round_up(Out, 8) : Out;
  const int kc = kc_of(mode), bk = kc, BN = pick_bn_mode(Out, mode);
  const int k_iters = (int)((N + bk - 1) / bk);
  const int tiles = ((In + 127) / 128) * ((Out + BN - 1) / BN);
  int splits_req = pick_splits(tiles, k_iters, 8);
  if (splits_req > max_splits) splits_req = max_splits;
  const int kps = (k_iters + splits_req - 1) / splits_req;
  const int splits = (k_iters + kps - 1) / kps;
  float* partial = reinterpret_cast<float*>(base + off);
  CPT_REQUIRE(splits == 1 || (ws && ws_bytes >= (size_t)splits * Out * In * sizeof(float)), CPT_ERR_WORKSPACE,
              "linear_wgrad: workspace too small for %d split-K partials", splits);
  TcParams p{};
  // dw[o][i]: lanes = i.  A(m=i, k=n) = x[n][i] MN-major; B(col=o, k=n) = dy[n][o] MN-major
  if (int e = make_map_2d(&p.tmA, xa, mode, In, (uint64_t)N, xpitch, kc, bk, true)) return e;
  if (int e = make_map_2d(&p.tmB, ga, mode, Out, (uint64_t)N, gpitch, kc, bk, true)) return e;
  if (is_x3(mode)) {
    if (int e = make_map_2d(&p.tmA2, xa_lo, mode, In, (uint64_t)N, xpitch, kc, bk, true)) return e;
    if (int e = make_map_2d(&p.tmB2, ga_lo, mode, Out, (uint64_t)N, gpitch, kc, bk, true)) return e;
  }
  p.out = splits > 1 ? partial : dw; p.bias = nullptr; p.bias_mode = BIAS_NONE;
  if (int e = get_status_ptr(&p.status)) return e;
  p.M = In; p.N = Out;
  p.m_tiles = (In + 127) / 128; p.n_tiles = (Out + BN - 1) / BN; p.z_tiles = splits;
  p.k_iters_total = k_iters; p.k_iters_per_split = kps;
  p.col_stride = In; p.split_stride = (long long)Out * In; p.taps = 1;
  if (int e = launch_bn<true, true, OP_GEMM>(p, mode, BN, want_2cta(BN, (In + 127) / 128), st)) return e;
  if (splits > 1) {
    launch_reduce_splits(partial, dw, (int64_t)Out * In, splits, st);
    CPT_LAUNCH_CHECK("linear_wgrad reduce");
  }
  return CPT_OK;
}

// ------------------------------------------------------------------ packed-K convolution (first layers: tiny Ci)
// The im2col TMA path spends one 64-channel k-iteration per filter tap, so a layer with Ci = 3 (ResNet stem 7x7: 49 taps,
// VGG conv1) or Ci = 1 (MNIST conv1) wastes > 90 % of every tile's bytes and MMAs, and re-reads the activations once per
// tap through L2.  For those layers the patches are written out ONCE as an explicit bf16 matrix
//     col[(b, ho, wo)][k],  k = c*T + j*K + kk  (the OIHW order of one filter row),  row pitch Kp = round_up(Ci*T, 8)
// and the three passes become plain GEMMs over it with K = Ci*T packed densely:
//     fprop  y[b][co][px]   = col[px][:] . w[co][:] + bias      (lanes = pixels -> NCHW stores)
//     wgrad  dw[co][k]      = Σ_px dy_cl[px][co] col[px][k]      (== Linear wgrad, split-K)
//     dgrad  dcol[b][k][px] = dy_cl[px][:] . w[:][k]             (lanes = pixels), then a gather (col2im) into dx
struct PackGeom {
  int Kdim, Kp, Wpad;
  int64_t px;
  size_t smem;
};
static PackGeom pack_geom(const G& g) {
  PackGeom q{};
  q.Kdim = g.Ci * g.T;
  q.Kp = round_up(q.Kdim, 8);
  q.Wpad = (g.Wo - 1) * g.S + (g.K - 1) * g.D + 1;
  q.px = (int64_t)g.B * g.Ho * g.Wo;
  q.smem = (size_t)g.Ci * g.K * q.Wpad * sizeof(float);
  return q;
}
static bool packed_ok(const G& g, int mode) {
  if (mode != CPT_MODE_BF16 || g.Ho < 1 || g.Wo < 1) return false;
  const PackGeom q = pack_geom(g);
  // worth it when the tap-per-k-iteration path would run mostly empty; limits: one input-row patch in shared memory,
  // int32 pixel index, grid.y
  return g.Ci <= 16 && q.Kdim <= 1024 && q.smem <= 160 * 1024 && q.px < (1LL << 31) && g.B <= 65535 && g.H <= 65535 &&
         (int64_t)g.B * g.Ci <= 65535 &&
         (int64_t)g.B * q.Kdim * g.Ho * g.Wo < (1LL << 40);
}

// One block per (image, output row): the Ci*K input rows that row needs are loaded once (coalesced, zero-filled outside the
// image) into a patch [Ci*K][Wpad]; each warp then writes whole output rows (Kp bf16 = one contiguous run), every lane
// owning fixed k-pairs whose patch offsets it looked up once — no integer division in either loop.
template <int MAXQ>
__global__ void __launch_bounds__(256) im2col_pack_kernel(const float* __restrict__ x, uint32_t* __restrict__ col, int Ci, int H,
                                                          int W, int K, int P, int S, int D, int Ho, int Wo, int Kdim, int Kp,
                                                          int Wpad) {
  extern __shared__ float patch[];
  const int ho = blockIdx.x, b = blockIdx.y, T = K * K;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const float* xb = x + (int64_t)b * Ci * H * W;
  for (int r = warp; r < Ci * K; r += nwarps) {
    const int c = r / K, j = r - c * K, h = ho * S - P + j * D;
    const bool row_ok = h >= 0 && h < H;
    const float* src = xb + ((int64_t)c * H + (row_ok ? h : 0)) * W;
    float* dst = patch + (size_t)r * Wpad;
    for (int wp = lane; wp < Wpad; wp += 32) {
      const int w = wp - P;
      dst[wp] = (row_ok && w >= 0 && w < W) ? __ldg(src + w) : 0.f;
    }
  }
  // this lane's k-pairs: k = 2*(lane + 32 q), q < MAXQ = ceil(Kp / 64) rounded up to a power of two (Kp <= 1024 -> <= 16)
  int l0[MAXQ], l1[MAXQ];
  const int pairs = Kp >> 1;
#pragma unroll
  for (int q = 0; q < MAXQ; ++q) {
    const int k = 2 * (lane + 32 * q);
    l0[q] = l1[q] = -1;
    if (k < Kdim) { const int c = k / T, t = k - c * T, j = t / K, kk = t - j * K; l0[q] = (c * K + j) * Wpad + kk * D; }
    if (k + 1 < Kdim) { const int c = (k + 1) / T, t = k + 1 - c * T, j = t / K, kk = t - j * K; l1[q] = (c * K + j) * Wpad + kk * D; }
  }
  __syncthreads();
  uint32_t* dst = col + ((int64_t)b * Ho + ho) * Wo * pairs;
  for (int wo = warp; wo < Wo; wo += nwarps) {
    const float* pw = patch + wo * S;
    uint32_t* drow = dst + (int64_t)wo * pairs;
#pragma unroll
    for (int q = 0; q < MAXQ; ++q) {
      const int k2 = lane + 32 * q;
      if (k2 < pairs) {
        const float v0 = l0[q] >= 0 ? pw[l0[q]] : 0.f, v1 = l1[q] >= 0 ? pw[l1[q]] : 0.f;
        __nv_bfloat162 h2 = __floats2bfloat162_rn(v0, v1);
        drow[k2] = *reinterpret_cast<uint32_t*>(&h2);
      }
    }
  }
}

// Vectorised form of the write phase (Kp % 4 == 0, i.e. always for Kp = round_up(Ci K², 8)): the block's output — Wo rows of
// Kp bf16 — is one contiguous run, walked as 8-byte chunks of four consecutive k by all 256 threads (no idle lanes at row
// ends); the four patch offsets of a chunk come from a table in shared memory (one 128-bit load), the (row, chunk) index is
// advanced incrementally.  ~16 instructions per 8 bytes instead of ~15 per 4 bytes plus idle predicated iterations: the
// kernel was issue-bound (ncu: SM throughput 88 %, DRAM 26 %).
__global__ void __launch_bounds__(256) im2col_pack4_kernel(const float* __restrict__ x, uint2* __restrict__ col, int Ci, int H, int W,
                                                           int K, int P, int S, int D, int Ho, int Wo, int Kdim, int Kp, int Wpad) {
  extern __shared__ float patch[];                                     // [Ci*K][Wpad] fp32, then int4 offs[Kp / 4]
  const int nch = Kp >> 2;                                              // chunks per output row
  int4* offs = reinterpret_cast<int4*>(patch + (((size_t)Ci * K * Wpad + 3) & ~(size_t)3));
  const int ho = blockIdx.x, b = blockIdx.y, T = K * K;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const float* xb = x + (int64_t)b * Ci * H * W;
  for (int r = warp; r < Ci * K; r += nwarps) {
    const int c = r / K, j = r - c * K, h = ho * S - P + j * D;
    const bool row_ok = h >= 0 && h < H;
    const float* src = xb + ((int64_t)c * H + (row_ok ? h : 0)) * W;
    float* dst = patch + (size_t)r * Wpad;
    for (int wp = lane; wp < Wpad; wp += 32) {
      const int w = wp - P;
      dst[wp] = (row_ok && w >= 0 && w < W) ? __ldg(src + w) : 0.f;
    }
  }
  for (int ch = threadIdx.x; ch < nch; ch += blockDim.x) {
    int o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = 4 * ch + e;
      o[e] = -1;
      if (k < Kdim) { const int c = k / T, t = k - c * T, j = t / K, kk = t - j * K; o[e] = (c * K + j) * Wpad + kk * D; }
    }
    offs[ch] = make_int4(o[0], o[1], o[2], o[3]);
  }
  __syncthreads();
  uint2* dst = col + ((int64_t)b * Ho + ho) * Wo * nch;
  const int total = Wo * nch;
  // idx = wo * nch + ch, advanced by blockDim.x = dq * nch + dr per step
  const int dq = blockDim.x / nch, dr = blockDim.x - dq * nch;
  int wo = threadIdx.x / nch, ch = threadIdx.x - wo * nch;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int4 o = offs[ch];
    const float* pw = patch + wo * S;
    const float v0 = o.x >= 0 ? pw[o.x] : 0.f, v1 = o.y >= 0 ? pw[o.y] : 0.f, v2 = o.z >= 0 ? pw[o.z] : 0.f, v3 = o.w >= 0 ? pw[o.w] : 0.f;
    __nv_bfloat162 lo = __floats2bfloat162_rn(v0, v1), hi = __floats2bfloat162_rn(v2, v3);
    dst[idx] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    wo += dq; ch += dr;
    if (ch >= nch) { ch -= nch; ++wo; }
  }
}

// wT[k][co] = w[co][k] as bf16, row pitch Cop (zero padded): the K-major B operand of the dgrad GEMM
__global__ void w_packT_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int Co, int Kdim, int Cop) {
  const int64_t n = (int64_t)Kdim * Cop;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cop), k = (int)(i / Cop);
    dst[i] = __float2bfloat16_rn(co < Co ? w[(int64_t)co * Kdim + k] : 0.f);
  }
}

// dx[b][c][h][w] = Σ_{j, kk : (h + P - j D) = S ho, (w + P - kk D) = S wo} dcol[b][c*T + j*K + kk][ho][wo]   (fixed order)
// One thread per dx element, consecutive threads along w (lanes of equal parity read consecutive wo of one dcol row).
// The loops run over the output positions (ho, wo) that can reach (h, w) — about (K/S)^2 of them — instead of over all
// K^2 taps; DIL1 removes the divisibility test of the dilated case.
template <bool DIL1>
__global__ void __launch_bounds__(256) col2im_kernel(const __nv_bfloat16* __restrict__ dcol, float* __restrict__ dx, int Ci, int H, int W,
                                                     int K, int P, int S, int D, int Ho, int Wo) {
  const int T = K * K;
  const int64_t plane = (int64_t)Ho * Wo;
  const int w = blockIdx.x * blockDim.x + threadIdx.x, h = blockIdx.y;
  const int bc = blockIdx.z;  // b * Ci + c
  if (w >= W) return;
  const __nv_bfloat16* base = dcol + (int64_t)bc * T * plane;
  const int span = (K - 1) * D;
  // ho in [ceil((h + P - span) / S), floor((h + P) / S)] ∩ [0, Ho)
  const int hp = h + P, wp = w + P;
  int ho_lo = hp - span > 0 ? (hp - span + S - 1) / S : 0, ho_hi = hp / S;
  int wo_lo = wp - span > 0 ? (wp - span + S - 1) / S : 0, wo_hi = wp / S;
  if (ho_hi > Ho - 1) ho_hi = Ho - 1;
  if (wo_hi > Wo - 1) wo_hi = Wo - 1;
  float acc = 0.f;
  // ascending j == descending ho: keeps the summation order of the tap loop.  The wo loop is walked four positions at a
  // time with predicated loads so that several independent loads are in flight per thread (the kernel is latency-bound
  // otherwise: ~16 dependent 4-byte loads per output).
  for (int ho = ho_hi; ho >= ho_lo; --ho) {
    const int hh = hp - ho * S;
    int j = hh;
    if (!DIL1) { j = hh / D; if (j * D != hh) continue; }
    const __nv_bfloat16* rowp = base + (int64_t)(j * K) * plane + (int64_t)ho * Wo;
    for (int wo0 = wo_hi; wo0 >= wo_lo; wo0 -= 4) {
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int wo = wo0 - u, ww = wp - wo * S;
        int kk = ww;
        bool ok = wo >= wo_lo;
        if (!DIL1) { kk = ww / D; ok = ok && kk * D == ww; }
        v[u] = ok ? __bfloat162float(rowp[(int64_t)kk * plane + wo]) : 0.f;
      }
      acc += v[0]; acc += v[1]; acc += v[2]; acc += v[3];
    }
  }
  dx[((int64_t)bc * H + h) * W + w] = acc;
}

// Same sum, staged through shared memory: one block per (image-channel, dx row h).  The <= ceil(K/S) output rows ho that
// reach h contribute K dcol rows each ((j, kk) planes, Wo contiguous bf16): they are copied once, coalesced, into
// rows[hi][kk][Wo + 1] (fp32, odd pitch: the two lane parities of a stride-2 gather land in different banks), then every
// thread sums its (ho, wo) pairs from shared memory in the order of col2im_kernel (bit-identical).  Every dcol element is
// read from HBM exactly once by a 4-byte coalesced load instead of a 2-byte gather (the stem's col2im went from 21 % of
// HBM bandwidth to the streaming rate of the 0.94 GB it has to read).
template <bool DIL1, int SS>
__global__ void __launch_bounds__(256) col2im_rows_kernel(const __nv_bfloat16* __restrict__ dcol, float* __restrict__ dx, int H, int W,
                                                          int K, int P, int S_rt, int D, int Ho, int Wo, int nrows_max) {
  extern __shared__ float rows[];                                                       // [nho][K][Wo + 1] fp32
  const __nv_bfloat16** rowptr = reinterpret_cast<const __nv_bfloat16**>(rows + (size_t)nrows_max * (Wo + 1) + ((nrows_max * (Wo + 1)) & 1));
  const int S = SS ? SS : S_rt;  // compile-time stride: the index divisions below become shifts
  const int h = blockIdx.x, bc = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int pitch = Wo + 1, span = (K - 1) * D, hp = h + P;
  const int64_t plane = (int64_t)Ho * Wo;
  const int ho_lo = hp - span > 0 ? (hp - span + S - 1) / S : 0;
  int ho_hi = hp / S;
  if (ho_hi > Ho - 1) ho_hi = Ho - 1;
  const int nho = ho_hi - ho_lo + 1, nrows = nho * K;
  // source row of every staged row (NULL: this ho does not reach h under dilation), one thread per row
  if ((int)threadIdx.x < nrows) {
    const int r = threadIdx.x, hi = r / K, kk = r - hi * K, ho = ho_hi - hi, hh = hp - ho * S;
    int j = hh;
    bool ok = true;
    if (!DIL1) { j = hh / D; ok = j * D == hh; }
    rowptr[r] = ok ? dcol + ((int64_t)bc * K * K + j * K + kk) * plane + (int64_t)ho * Wo : nullptr;
  }
  __syncthreads();
  if ((Wo & 1) == 0) {
    // each warp stages four rows at a time, two 4-byte words per lane and row: eight independent loads in flight per lane
    const int wpr = Wo >> 1;
    for (int r0 = warp; r0 < nrows; r0 += nwarps * 4) {
      for (int i0 = 0; i0 < wpr; i0 += 64) {
        uint32_t v[4][2];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = r0 + u * nwarps;
          const uint32_t* p = r < nrows ? reinterpret_cast<const uint32_t*>(rowptr[r]) : nullptr;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int i = i0 + lane + 32 * e;
            v[u][e] = (p && i < wpr) ? __ldg(p + i) : 0u;
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = r0 + u * nwarps;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int i = i0 + lane + 32 * e;
            if (r < nrows && i < wpr) {
              rows[r * pitch + 2 * i] = __uint_as_float(v[u][e] << 16);
              rows[r * pitch + 2 * i + 1] = __uint_as_float(v[u][e] & 0xffff0000u);
            }
          }
        }
      }
    }
  } else {
    for (int r = warp; r < nrows; r += nwarps) {
      const __nv_bfloat16* src = rowptr[r];
      if (!src) continue;
      for (int i = lane; i < Wo; i += 32) rows[r * pitch + i] = __bfloat162float(src[i]);
    }
  }
  __syncthreads();
  for (int w = threadIdx.x; w < W; w += blockDim.x) {
    const int wp = w + P;
    const int wo_lo = wp - span > 0 ? (wp - span + S - 1) / S : 0;
    int wo_hi = wp / S;
    if (wo_hi > Wo - 1) wo_hi = Wo - 1;
    float acc = 0.f;
    for (int hi = 0; hi < nho; ++hi) {
      if (!DIL1 && rowptr[hi * K] == nullptr) continue;
      const float* rp = rows + hi * K * pitch;
      for (int wo = wo_hi; wo >= wo_lo; --wo) {
        const int ww = wp - wo * S;
        int kk = ww;
        if (!DIL1) { kk = ww / D; if (kk * D != ww) continue; }
        acc += rp[kk * pitch + wo];
      }
    }
    dx[((int64_t)bc * H + h) * W + w] = acc;
  }
}

size_t conv_packed_bytes(const cpt_conv2d_desc* d, int mode) {
  const G g = geom(d);
  if (!packed_ok(g, mode)) return 0;
  const PackGeom q = pack_geom(g);
  return align_up((size_t)q.px * q.Kp * 2, 1024);
}

size_t conv_packed_workspace_size(int op, const cpt_conv2d_desc* d) {
  const G g = geom(d);
  const PackGeom q = pack_geom(g);
  if (op == CPT_OP_FPROP) return cast_bytes(g.Co, q.Kdim) + 1024;
  if (op == CPT_OP_DGRAD)
    return align_up((size_t)q.Kdim * round_up(g.Co, 8) * 2, 1024) + align_up((size_t)g.B * q.Kdim * g.Ho * g.Wo * 2, 1024) + 1024;
  return align_up((size_t)64 * g.Co * q.Kdim * sizeof(float), 1024) + 1024;
}

int conv_im2col_pack(const cpt_conv2d_desc* d, const float* x, void* col, cudaStream_t st) {
  const G g = geom(d);
  CPT_REQUIRE(packed_ok(g, CPT_MODE_BF16), CPT_ERR_UNSUPPORTED, "conv2d_im2col_pack: geometry not covered by the packed-K path");
  const PackGeom q = pack_geom(g);
  const int nq = (q.Kp / 2 + 31) / 32;
  static const bool pair_form = getenv("CPT_IM2COL_PAIRS") != nullptr;  // A/B switch for tools/stem_bench.py
  const size_t smem4 = ((((size_t)g.Ci * g.K * q.Wpad + 3) & ~(size_t)3)) * sizeof(float) + (size_t)(q.Kp / 4) * sizeof(int4);
  if (!pair_form && q.Kp % 4 == 0 && (reinterpret_cast<uintptr_t>(col) & 7) == 0 && smem4 <= 200 * 1024) {
    if (smem4 > 48 * 1024) CPT_CUDA(cudaFuncSetAttribute(im2col_pack4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    im2col_pack4_kernel<<<dim3(g.Ho, g.B), 256, smem4, st>>>(x, reinterpret_cast<uint2*>(col), g.Ci, g.H, g.W, g.K, g.P, g.S, g.D, g.Ho,
                                                            g.Wo, q.Kdim, q.Kp, q.Wpad);
    CPT_LAUNCH_CHECK("im2col_pack4");
    return CPT_OK;
  }
  auto launch = [&](auto kern) -> int {
    if (q.smem > 48 * 1024) CPT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    kern<<<dim3(g.Ho, g.B), 256, q.smem, st>>>(x, reinterpret_cast<uint32_t*>(col), g.Ci, g.H, g.W, g.K, g.P, g.S, g.D, g.Ho, g.Wo,
                                               q.Kdim, q.Kp, q.Wpad);
    return CPT_OK;
  };
  int e;
  if (nq <= 1) e = launch(im2col_pack_kernel<1>);
  else if (nq <= 2) e = launch(im2col_pack_kernel<2>);
  else if (nq <= 4) e = launch(im2col_pack_kernel<4>);
  else if (nq <= 8) e = launch(im2col_pack_kernel<8>);
  else e = launch(im2col_pack_kernel<16>);
  if (e) return e;
  CPT_LAUNCH_CHECK("im2col_pack");
  return CPT_OK;
}

int conv_fprop_packed(const cpt_conv2d_desc* d, const void* col, const float* w, const float* bias, float* y, void* ws,
                      size_t ws_bytes, cudaStream_t st, float* stats = nullptr, bool relu = false) {
  const G g = geom(d);
  const int mode = CPT_MODE_BF16;
  CPT_REQUIRE(packed_ok(g, mode), CPT_ERR_UNSUPPORTED, "conv2d_fprop_packed: geometry not covered by the packed-K path");
  CPT_REQUIRE(ws && ws_bytes >= conv_packed_workspace_size(CPT_OP_FPROP, d), CPT_ERR_WORKSPACE, "conv2d_fprop_packed: workspace too small");
  const PackGeom q = pack_geom(g);
  if (int e = cast_to_bf16(w, ws, g.Co, q.Kdim, st)) return e;  // [Co][Kp]: OIHW rows are already in k order
  const int kc = kc_of(mode), BN = pick_bn(g.Co);
  TcParams p{};
  const bool use2 = want_2cta(BN, (q.px + 127) / 128);
  if (int e = make_map_2d(&p.tmA, col, mode, q.Kdim, (uint64_t)q.px, q.Kp, kc, 128)) return e;
  if (int e = make_map_2d(&p.tmB, ws, mode, q.Kdim, g.Co, q.Kp, kc, use2 ? BN / 2 : BN)) return e;
  p.out = y; p.bias = bias; p.bias_mode = bias ? BIAS_COL : BIAS_NONE;
  p.relu = relu ? 3 : 0;
  if (stats) {
    CPT_CUDA(cudaMemsetAsync(stats, 0, stats_bytes(g.Co), st));
    p.stats = stats;
  }
  if (int e = get_status_ptr(&p.status)) return e;
  p.M = (int)q.px; p.N = g.Co;
  p.m_tiles = (int)((q.px + 127) / 128); p.n_tiles = (g.Co + BN - 1) / BN; p.z_tiles = 1;
  p.k_iters_total = (q.Kdim + kc - 1) / kc; p.k_iters_per_split = p.k_iters_total;
  p.col_stride = (long long)g.Ho * g.Wo;
  p.lane_is_pixel = 1; p.px_per_img = g.Ho * g.Wo; p.img_stride = (long long)g.Co * g.Ho * g.Wo;
  p.out_W = g.Wo; p.out_s = 1; p.Wo = g.Wo; p.taps = 1;
  return launch_bn<false, false, OP_GEMM>(p, mode, BN, use2, st);
}

int conv_wgrad_packed(const cpt_conv2d_desc* d, const void* col, const void* dy_cl, float* dw, void* ws, size_t ws_bytes,
                      cudaStream_t st) {
  const G g = geom(d);
  CPT_REQUIRE(packed_ok(g, CPT_MODE_BF16), CPT_ERR_UNSUPPORTED, "conv2d_wgrad_packed: geometry not covered by the packed-K path");
  CPT_REQUIRE(ws && ws_bytes >= conv_packed_workspace_size(CPT_OP_WGRAD, d), CPT_ERR_WORKSPACE, "conv2d_wgrad_packed: workspace too small");
  const PackGeom q = pack_geom(g);
  // dw[co][k] = Σ_px dy_cl[px][co] * col[px][k]: the Linear weight gradient with x = col, In = Ci*T, Out = Co
  return linear_wgrad_lp(col, dy_cl, dw, q.px, q.Kdim, g.Co, CPT_MODE_BF16, ws, ws_bytes, st, 64);
}

int conv_dgrad_packed(const cpt_conv2d_desc* d, const void* dy_cl, const float* w, float* dx, void* ws, size_t ws_bytes,
                      cudaStream_t st) {
  const G g = geom(d);
  const int mode = CPT_MODE_BF16;
  CPT_REQUIRE(packed_ok(g, mode), CPT_ERR_UNSUPPORTED, "conv2d_dgrad_packed: geometry not covered by the packed-K path");
  CPT_REQUIRE(ws && ws_bytes >= conv_packed_workspace_size(CPT_OP_DGRAD, d), CPT_ERR_WORKSPACE, "conv2d_dgrad_packed: workspace too small");
  const PackGeom q = pack_geom(g);
  const int Cop = round_up(g.Co, 8), kc = kc_of(mode);
  char* base = reinterpret_cast<char*>(ws);
  __nv_bfloat16* wT = reinterpret_cast<__nv_bfloat16*>(base);
  // dcol is kept in bf16: each of its entries is one tap's contribution to a dx element (a 64..512-term fp32 dot product rounded
  // once); dx sums <= ceil(K/S)^2 of them in fp32.  Halves the traffic of the two passes that dominate this path.
  __nv_bfloat16* dcol = reinterpret_cast<__nv_bfloat16*>(base + align_up((size_t)q.Kdim * Cop * 2, 1024));
  w_packT_kernel<<<ew_grid((int64_t)q.Kdim * Cop, 256), 256, 0, st>>>(w, wT, g.Co, q.Kdim, Cop);
  CPT_LAUNCH_CHECK("w_packT");
  const int BN = pick_bn(q.Kdim);
  TcParams p{};
  const bool use2 = want_2cta(BN, (q.px + 127) / 128);
  // dcol[b][k][ho][wo]: lanes = pixels (A = dy_cl [px][Cop], K-major), columns = k (B = wT [Kdim][Cop], K-major)
  if (int e = make_map_2d(&p.tmA, dy_cl, mode, g.Co, (uint64_t)q.px, Cop, kc, 128)) return e;
  if (int e = make_map_2d(&p.tmB, wT, mode, g.Co, q.Kdim, Cop, kc, use2 ? BN / 2 : BN)) return e;
  p.out = reinterpret_cast<float*>(dcol); p.out_bf16 = 1; p.bias = nullptr; p.bias_mode = BIAS_NONE;
  if (int e = get_status_ptr(&p.status)) return e;
  p.M = (int)q.px; p.N = q.Kdim;
  p.m_tiles = (int)((q.px + 127) / 128); p.n_tiles = (q.Kdim + BN - 1) / BN; p.z_tiles = 1;
  p.k_iters_total = (g.Co + kc - 1) / kc; p.k_iters_per_split = p.k_iters_total;
  p.col_stride = (long long)g.Ho * g.Wo;
  p.lane_is_pixel = 1; p.px_per_img = g.Ho * g.Wo; p.img_stride = (long long)q.Kdim * g.Ho * g.Wo;
  p.out_W = g.Wo; p.out_s = 1; p.Wo = g.Wo; p.taps = 1;
  if (int e = launch_bn<false, false, OP_GEMM>(p, mode, BN, use2, st)) return e;
  {
    CPT_REQUIRE(g.H <= 65535 && (int64_t)g.B * g.Ci <= 65535, CPT_ERR_UNSUPPORTED, "conv2d_dgrad_packed: grid too large");
    const int nho_max = ((g.K - 1) * g.D) / g.S + 1, nrows_max = nho_max * g.K;
    const size_t smem = ((size_t)nrows_max * (g.Wo + 1) + 1) * sizeof(float) + (size_t)nrows_max * sizeof(void*);
    static const bool gather_form = getenv("CPT_COL2IM_GATHER") != nullptr;  // A/B switch for tools/stem_bench.py
    if (smem <= 48 * 1024 && nrows_max <= 256 && !gather_form) {  // rows of one dx line fit in shared memory: streaming form
      int bx = g.W >= 256 ? 256 : ((g.W + 31) / 32) * 32;
      if (bx < nrows_max) bx = ((nrows_max + 31) / 32) * 32;
      dim3 grid(g.H, g.B * g.Ci);
#define CPT_C2I(DIL, SS) col2im_rows_kernel<DIL, SS><<<grid, bx, smem, st>>>(dcol, dx, g.H, g.W, g.K, g.P, g.S, g.D, g.Ho, g.Wo, nrows_max)
      if (g.D == 1) { if (g.S == 1) CPT_C2I(true, 1); else if (g.S == 2) CPT_C2I(true, 2); else CPT_C2I(true, 0); }
      else CPT_C2I(false, 0);
#undef CPT_C2I
    } else {
      const int bx = g.W >= 256 ? 256 : (g.W >= 128 ? 128 : (g.W >= 64 ? 64 : 32));
      dim3 grid((g.W + bx - 1) / bx, g.H, g.B * g.Ci);
      if (g.D == 1) col2im_kernel<true><<<grid, bx, 0, st>>>(dcol, dx, g.Ci, g.H, g.W, g.K, g.P, g.S, g.D, g.Ho, g.Wo);
      else col2im_kernel<false><<<grid, bx, 0, st>>>(dcol, dx, g.Ci, g.H, g.W, g.K, g.P, g.S, g.D, g.Ho, g.Wo);
    }
  }
  CPT_LAUNCH_CHECK("col2im");
  return CPT_OK;
}


// ------------------------------------------------------------------ strip ("shared halo") path: see strip_kernel.cuh
// Stride-1 / dilation-1 / same-padded odd-K layers in bf16 mode whose channel counts are multiples of 64.  Activations are
// staged zero-padded (cpt_to_channels_last_padded); fprop and dgrad run strip_conv_kernel, wgrad reads the same padded
// tensors through im2col maps whose bounding box is the interior.
struct StripPlan {
  int Hp, Wp, box_rows, n_loads, unit_bytes, n_units, b_stages, resident, BN, m_tiles, n_tiles;
  long long M_lanes;
  size_t smem;
};
static int strip_max_channels() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CPT_STRIP_MAX_C");
    v = e ? atoi(e) : 128;   // larger layers are tensor-pipe bound on the im2col path already
  }
  return v;
}
static bool strip_plan(int B, int Cact, int H, int W, int K, int Ncols, StripPlan& q) {
  const int P = (K - 1) / 2;
  q.Hp = H + 2 * P; q.Wp = W + 2 * P;
  const int SR = 128 + (K - 1) * (q.Wp + 1);
  if (SR <= 256) { q.n_loads = 1; q.box_rows = round_up(SR, 8); }
  else { q.n_loads = 2; q.box_rows = round_up((SR + 1) / 2, 8); }
  if (q.box_rows > 256) return false;
  q.unit_bytes = q.n_loads * q.box_rows * 128;
  q.BN = pick_bn(Ncols);
  q.n_tiles = (Ncols + q.BN - 1) / q.BN;
  q.M_lanes = (long long)(B - 1) * q.Hp * q.Wp + (long long)(H - 1) * q.Wp + W;
  if (q.M_lanes >= (1LL << 31) - 256 || (long long)B * q.Hp * q.Wp >= (1LL << 31)) return false;
  q.m_tiles = (int)((q.M_lanes + 127) / 128);
  const int T = K * K, cchunks = Cact / 64, b_bytes = q.BN * 128;
  const int avail = 232448 - 2048 - 4 * q.BN * 2 * 4 - 512;   // barriers + alignment slack, static stats accumulators
  q.resident = q.n_tiles == 1 && T * cchunks <= STRIP_MAX_BSTAGES && 2 * q.unit_bytes + T * cchunks * b_bytes <= avail;
  if (q.resident) {
    q.b_stages = T * cchunks;
    q.n_units = (avail - q.b_stages * b_bytes) / q.unit_bytes;
    if (q.n_units > STRIP_MAX_UNITS) q.n_units = STRIP_MAX_UNITS;
  } else {
    q.n_units = 3;
    if (avail - 3 * q.unit_bytes < 6 * b_bytes) q.n_units = 2;
    q.b_stages = (avail - q.n_units * q.unit_bytes) / b_bytes;
    if (q.b_stages > STRIP_MAX_BSTAGES) q.b_stages = STRIP_MAX_BSTAGES;
    if (q.b_stages < 3) return false;
  }
  q.smem = 1024 + 1024 + (size_t)q.n_units * q.unit_bytes + (size_t)q.b_stages * b_bytes;
  return q.n_units >= 2;
}

// Opt-in (cpt_conv2d_set_strip_enabled / CPT_STRIP=1): measured on B200 it removes 84 % of the L2->SM traffic of a C = 64
// layer (1.39 GB -> 0.23 GB) but is not faster than the im2col kernel — both are bound by the 128x64x16 SS-mode MMA (64 cycles
// each, operand fetch from shared memory) issued from one thread whose descriptor registers are recycled per tap.
static int g_strip_enabled = -1;
static bool strip_enabled() {
  if (g_strip_enabled < 0) {
    const char* e = getenv("CPT_STRIP");
    g_strip_enabled = (e && atoi(e) != 0) ? 1 : 0;
  }
  return g_strip_enabled != 0;
}

bool strip_ok(const G& g, int mode) {
  if (!strip_enabled() || mode != CPT_MODE_BF16 || g.S != 1 || g.D != 1 || g.K < 3 || (g.K & 1) == 0 || g.K > 7 || g.P != (g.K - 1) / 2) return false;
  if (g.Ci % 64 != 0 || g.Co % 64 != 0 || g.Ci > strip_max_channels() || g.Co > strip_max_channels()) return false;
  StripPlan a, b;
  return strip_plan(g.B, g.Ci, g.H, g.W, g.K, g.Co, a) && strip_plan(g.B, g.Co, g.H, g.W, g.K, g.Ci, b);
}

size_t cl_padded_bytes(int B, int C, int H, int W, int P) {
  return align_up((size_t)B * (H + 2 * P) * (W + 2 * P) * round_up(C, 8) * 2, 1024);
}

// zeroes the pad pixels of act_pad[B][Hp][Wp][Cp] (bf16): one warp per border pixel, 16 bytes per lane and step
__global__ void __launch_bounds__(256) zero_border_kernel(uint4* __restrict__ dst, int B, int H, int W, int P, int vec_per_px) {
  const int Wp = W + 2 * P, Hp = H + 2 * P;
  const int per_img = Hp * Wp - H * W, top = P * Wp;
  const int64_t total = (int64_t)B * per_img;
  const int lane = threadIdx.x & 31;
  for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < total; i += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    const int b = (int)(i / per_img), r = (int)(i - (int64_t)b * per_img);
    int hp, wp;
    if (r < top) { hp = r / Wp; wp = r - hp * Wp; }
    else if (r < 2 * top) { const int rr = r - top; hp = H + P + rr / Wp; wp = rr % Wp; }
    else { const int rr = r - 2 * top, row = rr / (2 * P), k = rr - row * (2 * P); hp = P + row; wp = k < P ? k : W + k; }
    uint4* px = dst + ((int64_t)b * Hp * Wp + (int64_t)hp * Wp + wp) * vec_per_px;
    for (int v = lane; v < vec_per_px; v += 32) px[v] = make_uint4(0u, 0u, 0u, 0u);
  }
}

int to_channels_last_padded(const float* src, void* dst, int B, int C, int H, int W, int P, float* chan_sum, void* ws, size_t ws_bytes,
                            cudaStream_t st) {
  const int Cp = round_up(C, 8), HW = H * W;
  int gx, groups;
  cl_grid(B, C, H, W, gx, groups);
  dim3 grid(gx, groups, 1);
  CPT_REQUIRE(grid.y <= 65535 && (int64_t)B * (H + 2 * P) * (W + 2 * P) < (1LL << 31) && (int64_t)B * C * HW < (1LL << 32) - (1 << 20),
              CPT_ERR_UNSUPPORTED, "to_channels_last_padded: tensor too large");
  if (P > 0) {
    const int64_t border = (int64_t)B * ((H + 2 * P) * (W + 2 * P) - HW);
    zero_border_kernel<<<ew_grid(border * 32, 256), 256, 0, st>>>(reinterpret_cast<uint4*>(dst), B, H, W, P, Cp * 2 / 16);
    CPT_LAUNCH_CHECK("zero_border");
  }
  float* partial = nullptr;
  if (chan_sum && ws && ws_bytes >= to_channels_last_ws(B, C, H, W)) partial = reinterpret_cast<float*>(ws);
  const int64_t Q = (int64_t)B * HW;
  nchw_to_nhwc_kernel<true, true><<<grid, 256, 0, st>>>(src, dst, C, HW, Cp, chan_sum, partial, Q, nullptr, PadGeom{W, H, P});
  CPT_LAUNCH_CHECK("nchw_to_nhwc_padded");
  if (partial) {
    chan_partial_reduce_kernel<<<(C + 31) / 32, 1024, 0, st>>>(partial, chan_sum, C, (int64_t)gx, groups * CL_CH);
    CPT_LAUNCH_CHECK("chan_partial_reduce");
  }
  return CPT_OK;
}

// out[b, n, h, w] = Σ_{t=(j,k), ch} act_pad[b, h + j, w + k, ch] * wmat[n][t][ch]  (+ bias[n])
static int conv_strip_gemm(const void* act_pad, int B, int Cact, int H, int W, int K, const void* wmat, int Ncols, const float* bias,
                           float* out, cudaStream_t st, float* stats) {
  StripPlan q;
  CPT_REQUIRE(strip_plan(B, Cact, H, W, K, Ncols, q), CPT_ERR_UNSUPPORTED, "strip convolution: geometry does not fit shared memory");
  const int mode = CPT_MODE_BF16, T = K * K, Ck = round_up(Cact, 64);
  StripParams p{};
  if (int e = make_map_2d(&p.tmA, act_pad, mode, Cact, (uint64_t)B * q.Hp * q.Wp, round_up(Cact, 8), 64, q.box_rows)) return e;
  if (int e = make_map_2d(&p.tmB, wmat, mode, (uint64_t)T * Ck, Ncols, (uint64_t)T * Ck, 64, q.BN)) return e;
  p.out = out; p.bias = bias;
  if (int e = get_status_ptr(&p.status)) return e;
  if (stats) {
    CPT_CUDA(cudaMemsetAsync(stats, 0, stats_bytes(Ncols), st));
    p.stats = stats;
  }
  p.M_lanes = (int)q.M_lanes; p.N = Ncols; p.m_tiles = q.m_tiles; p.n_tiles = q.n_tiles;
  p.T = T; p.cchunks = Cact / 64; p.wk_cols = Ck;
  p.H = H; p.W = W; p.Wp = q.Wp; p.HpWp = q.Hp * q.Wp;
  p.box_rows = q.box_rows; p.n_loads = q.n_loads; p.unit_bytes = q.unit_bytes; p.n_units = q.n_units; p.b_stages = q.b_stages;
  p.resident = q.resident;
  p.col_stride = (long long)H * W; p.img_stride = (long long)Ncols * H * W;
  for (int j = 0; j < K; ++j)
    for (int k = 0; k < K; ++k) p.tap_off[j * K + k] = j * q.Wp + k;
  int grid = sm_count() - g_reserved_sms;
  const int total = q.m_tiles * q.n_tiles;
  if (grid > total) grid = total;
  if (grid < 1) return CPT_OK;
  return launch_strip(p, q.BN, grid, q.smem, st);
}

size_t conv_strip_workspace_size(int op, const cpt_conv2d_desc* d) {
  const G g = geom(d);
  const int mode = CPT_MODE_BF16;
  if (op == CPT_OP_FPROP) return wmat_bytes(g.Co, g.T, g.Ci, mode) + 1024;
  if (op == CPT_OP_DGRAD) return wmat_bytes(g.Ci, g.T, g.Co, mode) + 1024;
  return align_up((size_t)wgrad_splits(g, mode) * g.Co * g.T * g.Ci * sizeof(float), 1024) + 1024;
}

int conv_fprop_strip(const cpt_conv2d_desc* d, const void* x_pad, const float* w, const float* bias, float* y, void* ws, size_t ws_bytes,
                     cudaStream_t st, float* stats) {
  const G g = geom(d);
  CPT_REQUIRE(strip_ok(g, CPT_MODE_BF16), CPT_ERR_UNSUPPORTED, "conv2d_fprop_strip: geometry not covered by the strip path");
  CPT_REQUIRE(ws && ws_bytes >= conv_strip_workspace_size(CPT_OP_FPROP, d), CPT_ERR_WORKSPACE, "conv2d_fprop_strip: workspace too small");
  const int Ck = round_up(g.Ci, 64);
  const int64_t n = (int64_t)g.Co * g.T * Ck;
  w_fprop_kernel<true><<<ew_grid(n, 256), 256, 0, st>>>(w, ws, g.Co, g.Ci, g.T, Ck, nullptr);
  CPT_LAUNCH_CHECK("w_fprop");
  return conv_strip_gemm(x_pad, g.B, g.Ci, g.H, g.W, g.K, ws, g.Co, bias, y, st, stats);
}

int conv_dgrad_strip(const cpt_conv2d_desc* d, const void* dy_pad, const float* w, float* dx, void* ws, size_t ws_bytes, cudaStream_t st) {
  const G g = geom(d);
  CPT_REQUIRE(strip_ok(g, CPT_MODE_BF16), CPT_ERR_UNSUPPORTED, "conv2d_dgrad_strip: geometry not covered by the strip path");
  CPT_REQUIRE(ws && ws_bytes >= conv_strip_workspace_size(CPT_OP_DGRAD, d), CPT_ERR_WORKSPACE, "conv2d_dgrad_strip: workspace too small");
  // dx[h][w] = Σ_{j',k'} dy_pad[h + j'][w + k'] * w[:, :, K-1-j', K-1-k']: the forward strip over dy with the taps reversed
  const int Cok = round_up(g.Co, 64);
  TapIdx ti{};
  for (int t = 0; t < g.T; ++t) ti.idx[t] = (unsigned char)(g.T - 1 - t);
  const int64_t n = (int64_t)g.Ci * g.T * Cok;
  w_dgrad_kernel<true><<<ew_grid(n, 256), 256, 0, st>>>(w, ws, g.Co, g.Ci, g.T, g.T, Cok, ti, nullptr);
  CPT_LAUNCH_CHECK("w_dgrad");
  return conv_strip_gemm(dy_pad, g.B, g.Co, g.H, g.W, g.K, ws, g.Ci, nullptr, dx, st, nullptr);
}

// wgrad over the padded tensors: A = x_pad through an im2col map with padding 0 (the pad is in the data), B = dy_pad through an
// im2col map whose bounding box is the interior; both walk the H x W output grid in the same order
int conv_wgrad_padded(const cpt_conv2d_desc* d, const void* x_pad, const void* dy_pad, float* dw, void* ws, size_t ws_bytes, cudaStream_t st) {
  const G g = geom(d);
  const int mode = CPT_MODE_BF16;
  CPT_REQUIRE(strip_ok(g, mode), CPT_ERR_UNSUPPORTED, "conv2d_wgrad_padded: geometry not covered by the strip path");
  const int kc = kc_of(mode), bk = kc, P = g.P, Hp = g.H + 2 * P, Wp = g.W + 2 * P;
  const int splits_req = wgrad_splits(g, mode);
  const int64_t pixels = (int64_t)g.B * g.Ho * g.Wo;
  const int k_iters = (int)((pixels + bk - 1) / bk);
  const int kps = (k_iters + splits_req - 1) / splits_req;
  const int splits = (k_iters + kps - 1) / kps;
  const size_t need = align_up((size_t)splits * g.Co * g.T * g.Ci * sizeof(float), 1024);
  CPT_REQUIRE(ws && ws_bytes >= need, CPT_ERR_WORKSPACE, "conv2d_wgrad_padded: workspace too small (%zu < %zu)", ws_bytes, need);
  const int BN = pick_bn(g.Co);
  TcParams p{};
  if (int e = make_map_im2col(&p.tmA, x_pad, mode, round_up(g.Ci, 8), Wp, Hp, g.B, 0, 0, -(g.K - 1), -(g.K - 1), 1, kc, bk, true)) return e;
  if (int e = make_map_im2col(&p.tmB, dy_pad, mode, round_up(g.Co, 8), Wp, Hp, g.B, P, P, -P, -P, 1, kc, bk, true)) return e;
  p.b_im2col = 1; p.b_pad = P;
  p.out = reinterpret_cast<float*>(ws);
  p.bias = nullptr; p.bias_mode = BIAS_NONE;
  if (int e = get_status_ptr(&p.status)) return e;
  p.M = g.Ci; p.N = g.Co;
  p.m_tiles = (g.Ci + 127) / 128; p.n_tiles = (g.Co + BN - 1) / BN; p.z_tiles = g.T * splits;
  p.k_iters_total = k_iters; p.k_iters_per_split = kps;
  p.col_stride = (long long)g.T * g.Ci; p.split_stride = (long long)g.Co * g.T * g.Ci; p.tap_stride = g.Ci;
  p.lane_is_pixel = 0; p.px_per_img = g.Ho * g.Wo; p.Wo = g.Wo; p.Ho = g.Ho;
  p.conv_stride = 1; p.pad = 0; p.dil = 1; p.Kw = g.K; p.taps = g.T; p.out_s = 1;
  if (int e = launch_bn<true, true, OP_WGRAD>(p, mode, BN, want_2cta(BN, (g.Ci + 127) / 128), st)) return e;
  return launch_wgrad_reduce(reinterpret_cast<float*>(ws), dw, g.Co, g.Ci, g.T, splits, st);
}

}  // namespace tc
}  // namespace cpt

using namespace cpt;

extern "C" {

static int check_tc_desc(const cpt_conv2d_desc* d) {
  CPT_REQUIRE(d && d->B > 0 && d->Ci > 0 && d->H > 0 && d->W > 0 && d->Co > 0 && d->K > 0 && d->pad >= 0 && d->stride >= 1 && d->dil >= 1,
              CPT_ERR_INVALID, "bad convolution descriptor");
  return CPT_OK;
}

/* strip ("shared halo") path for stride-1 same-padded small-channel layers: see include/compyute_b200.h */
int cpt_conv2d_set_strip_enabled(int enabled) {
  const int prev = tc::strip_enabled() ? 1 : 0;
  tc::g_strip_enabled = enabled ? 1 : 0;
  return prev;
}
int cpt_conv2d_strip_supported(const cpt_conv2d_desc* d, int mode) {
  if (!d || d->B <= 0 || d->Ci <= 0 || d->H <= 0 || d->W <= 0 || d->Co <= 0 || d->K <= 0 || d->pad < 0 || d->stride < 1 || d->dil < 1) return 0;
  return tc::strip_ok(tc::geom(d), mode) ? 1 : 0;
}
size_t cpt_channels_last_padded_bytes(int B, int C, int H, int W, int pad) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || pad < 0) return 0;
  return tc::cl_padded_bytes(B, C, H, W, pad);
}
int cpt_to_channels_last_padded(const float* src, void* dst, int B, int C, int H, int W, int pad, float* chan_sum, void* ws,
                                size_t ws_bytes, void* stream) {
  CPT_REQUIRE(src && dst && B > 0 && C > 0 && H > 0 && W > 0 && pad >= 0 && C % 8 == 0, CPT_ERR_INVALID, "to_channels_last_padded: bad arguments");
  return tc::to_channels_last_padded(src, dst, B, C, H, W, pad, chan_sum, ws, ws_bytes, as_stream(stream));
}
size_t cpt_conv2d_strip_workspace_size(int op, const cpt_conv2d_desc* d) {
  if (check_tc_desc(d)) return 0;
  return tc::conv_strip_workspace_size(op, d);
}
int cpt_conv2d_fprop_strip(const cpt_conv2d_desc* d, const void* x_pad, const float* w, const float* bias, float* y, float* stats,
                           void* ws, size_t ws_bytes, void* stream) {
  if (int e = check_tc_desc(d)) return e;
  CPT_REQUIRE(x_pad && w && y, CPT_ERR_INVALID, "conv2d_fprop_strip: null pointer");
  return tc::conv_fprop_strip(d, x_pad, w, bias, y, ws, ws_bytes, as_stream(stream), stats);
}
int cpt_conv2d_dgrad_strip(const cpt_conv2d_desc* d, const void* dy_pad, const float* w, float* dx, void* ws, size_t ws_bytes,
                           void* stream) {
  if (int e = check_tc_desc(d)) return e;
  CPT_REQUIRE(dy_pad && w && dx, CPT_ERR_INVALID, "conv2d_dgrad_strip: null pointer");
  return tc::conv_dgrad_strip(d, dy_pad, w, dx, ws, ws_bytes, as_stream(stream));
}
int cpt_conv2d_wgrad_padded(const cpt_conv2d_desc* d, const void* x_pad, const void* dy_pad, float* dw, void* ws, size_t ws_bytes,
                            void* stream) {
  if (int e = check_tc_desc(d)) return e;
  CPT_REQUIRE(x_pad && dy_pad && dw, CPT_ERR_INVALID, "conv2d_wgrad_padded: null pointer");
  return tc::conv_wgrad_padded(d, x_pad, dy_pad, dw, ws, ws_bytes, as_stream(stream));
}

size_t cpt_channels_last_bytes(int B, int C, int H, int W, int mode) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || mode == CPT_MODE_FP32) return 0;
  return tc::cl_bytes(B, C, H, W, mode);
}

size_t cpt_to_channels_last_workspace_size(int B, int C, int H, int W) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
  return tc::to_channels_last_ws(B, C, H, W);
}

int cpt_to_channels_last(const float* src, void* dst, int B, int C, int H, int W, int mode, float* chan_sum, void* ws,
                         size_t ws_bytes, void* stream) {
  CPT_REQUIRE(src && dst && B > 0 && C > 0 && H > 0 && W > 0, CPT_ERR_INVALID, "to_channels_last: bad arguments");
  CPT_REQUIRE(mode == CPT_MODE_TF32 || mode == CPT_MODE_BF16 || mode == CPT_MODE_FP32X3, CPT_ERR_INVALID,
              "to_channels_last: mode must be TF32, BF16 or FP32X3");
  return tc::to_channels_last(src, dst, B, C, H, W, mode, chan_sum, ws, ws_bytes, as_stream(stream));
}

static int check_tc(const cpt_conv2d_desc* d, int mode, const char* who) {
  CPT_REQUIRE(d && d->B > 0 && d->Ci > 0 && d->H > 0 && d->W > 0 && d->Co > 0 && d->K > 0 && d->pad >= 0 && d->stride >= 1 &&
                  d->dil >= 1,
              CPT_ERR_INVALID, "%s: bad descriptor", who);
  CPT_REQUIRE(mode == CPT_MODE_TF32 || mode == CPT_MODE_BF16 || mode == CPT_MODE_FP32X3, CPT_ERR_INVALID,
              "%s: mode must be TF32, BF16 or FP32X3", who);
  return CPT_OK;
}

int cpt_conv2d_fprop_cl(const cpt_conv2d_desc* d, const void* x_cl, const float* w, const float* bias, float* y, int mode,
                        void* ws, size_t ws_bytes, void* stream) {
  if (int e = check_tc(d, mode, "conv2d_fprop_cl")) return e;
  return tc::conv_fprop_cl(d, x_cl, w, bias, y, mode, ws, ws_bytes, as_stream(stream));
}
int cpt_conv2d_dgrad_cl(const cpt_conv2d_desc* d, const void* dy_cl, const float* w, float* dx, int mode, void* ws,
                        size_t ws_bytes, void* stream) {
  if (int e = check_tc(d, mode, "conv2d_dgrad_cl")) return e;
  return tc::conv_dgrad_cl(d, dy_cl, w, dx, mode, ws, ws_bytes, as_stream(stream));
}
int cpt_conv2d_dgrad_cl_supported(const cpt_conv2d_desc* d, int mode) {
  if (check_tc(d, mode, "conv2d_dgrad_cl_supported")) return 0;
  return tc::dgrad_tc_ok(tc::geom(d)) ? 1 : 0;
}
int cpt_conv2d_wgrad_cl(const cpt_conv2d_desc* d, const void* x_cl, const void* dy_cl, float* dw, int mode, void* ws,
                        size_t ws_bytes, void* stream) {
  if (int e = check_tc(d, mode, "conv2d_wgrad_cl")) return e;
  return tc::conv_wgrad_cl(d, x_cl, dy_cl, dw, mode, ws, ws_bytes, as_stream(stream));
}

/* packed-K path for first layers (tiny Ci): see include/compyute_b200.h */
size_t cpt_conv2d_packed_bytes(const cpt_conv2d_desc* d, int mode) {
  if (check_tc(d, mode, "conv2d_packed_bytes")) return 0;
  return tc::conv_packed_bytes(d, mode);
}
size_t cpt_conv2d_packed_workspace_size(int op, const cpt_conv2d_desc* d) {
  if (check_tc(d, CPT_MODE_BF16, "conv2d_packed_workspace_size")) return 0;
  return tc::conv_packed_workspace_size(op, d);
}
int cpt_conv2d_im2col_pack(const cpt_conv2d_desc* d, const float* x, void* col, void* stream) {
  if (int e = check_tc(d, CPT_MODE_BF16, "conv2d_im2col_pack")) return e;
  CPT_REQUIRE(x && col, CPT_ERR_INVALID, "conv2d_im2col_pack: null pointer");
  return tc::conv_im2col_pack(d, x, col, as_stream(stream));
}
int cpt_conv2d_fprop_packed(const cpt_conv2d_desc* d, const void* col, const float* w, const float* bias, float* y, void* ws,
                            size_t ws_bytes, void* stream) {
  if (int e = check_tc(d, CPT_MODE_BF16, "conv2d_fprop_packed")) return e;
  return tc::conv_fprop_packed(d, col, w, bias, y, ws, ws_bytes, as_stream(stream));
}
int cpt_conv2d_dgrad_packed(const cpt_conv2d_desc* d, const void* dy_cl, const float* w, float* dx, void* ws, size_t ws_bytes,
                            void* stream) {
  if (int e = check_tc(d, CPT_MODE_BF16, "conv2d_dgrad_packed")) return e;
  return tc::conv_dgrad_packed(d, dy_cl, w, dx, ws, ws_bytes, as_stream(stream));
}
int cpt_conv2d_wgrad_packed(const cpt_conv2d_desc* d, const void* col, const void* dy_cl, float* dw, void* ws, size_t ws_bytes,
                            void* stream) {
  if (int e = check_tc(d, CPT_MODE_BF16, "conv2d_wgrad_packed")) return e;
  return tc::conv_wgrad_packed(d, col, dy_cl, dw, ws, ws_bytes, as_stream(stream));
}

/* forward passes that also leave the batch statistics of their output for a BatchNorm that consumes it */
size_t cpt_conv2d_stats_bytes(const cpt_conv2d_desc* d) {
  if (!d || d->Co <= 0) return 0;
  return tc::stats_bytes(d->Co);
}
int cpt_conv2d_stats_slots(void) { return sm_count() * 4; }
int cpt_conv2d_fprop_cl_stats(const cpt_conv2d_desc* d, const void* x_cl, const float* w, const float* bias, float* y, float* stats,
                              int mode, void* ws, size_t ws_bytes, void* stream) {
  if (int e = check_tc(d, mode, "conv2d_fprop_cl_stats")) return e;
  CPT_REQUIRE(stats, CPT_ERR_INVALID, "conv2d_fprop_cl_stats: stats is NULL");
  return tc::conv_fprop_cl(d, x_cl, w, bias, y, mode, ws, ws_bytes, as_stream(stream), stats);
}
int cpt_conv2d_fprop_packed_stats(const cpt_conv2d_desc* d, const void* col, const float* w, const float* bias, float* y,
                                  float* stats, void* ws, size_t ws_bytes, void* stream) {
  if (int e = check_tc(d, CPT_MODE_BF16, "conv2d_fprop_packed_stats")) return e;
  CPT_REQUIRE(stats, CPT_ERR_INVALID, "conv2d_fprop_packed_stats: stats is NULL");
  return tc::conv_fprop_packed(d, col, w, bias, y, ws, ws_bytes, as_stream(stream), stats);
}

/* Conv2D -> ReLU pairs: the forward epilogue applies max(. , 0) (convolution_funcs.py:237-238 + activation_funcs.py:26-29);
 * the backward pass stages dy * (y > 0) in one pass (activation_funcs.py:32-34 folded into the operand staging of dgrad / wgrad) */
int cpt_conv2d_fprop_cl_relu(const cpt_conv2d_desc* d, const void* x_cl, const float* w, const float* bias, float* y, int mode,
                             void* ws, size_t ws_bytes, void* stream) {
  if (int e = check_tc(d, mode, "conv2d_fprop_cl_relu")) return e;
  return tc::conv_fprop_cl(d, x_cl, w, bias, y, mode, ws, ws_bytes, as_stream(stream), nullptr, true);
}
int cpt_conv2d_fprop_packed_relu(const cpt_conv2d_desc* d, const void* col, const float* w, const float* bias, float* y, void* ws,
                                 size_t ws_bytes, void* stream) {
  if (int e = check_tc(d, CPT_MODE_BF16, "conv2d_fprop_packed_relu")) return e;
  return tc::conv_fprop_packed(d, col, w, bias, y, ws, ws_bytes, as_stream(stream), nullptr, true);
}
int cpt_to_channels_last_gated(const float* src, const float* gate, void* dst, int B, int C, int H, int W, int mode, float* chan_sum,
                               void* ws, size_t ws_bytes, void* stream) {
  CPT_REQUIRE(src && gate && dst && B > 0 && C > 0 && H > 0 && W > 0, CPT_ERR_INVALID, "to_channels_last_gated: bad arguments");
  CPT_REQUIRE(mode == CPT_MODE_TF32 || mode == CPT_MODE_BF16 || mode == CPT_MODE_FP32X3, CPT_ERR_INVALID,
              "to_channels_last_gated: mode must be TF32, BF16 or FP32X3");
  return tc::to_channels_last(src, dst, B, C, H, W, mode, chan_sum, ws, ws_bytes, as_stream(stream), gate);
}

size_t cpt_cast_bf16_bytes(int64_t rows, int cols) { return rows > 0 && cols > 0 ? tc::cast_bytes(rows, cols) : 0; }
int cpt_cast_bf16(const float* src, void* dst, int64_t rows, int cols, void* stream) {
  CPT_REQUIRE(src && dst && rows > 0 && cols > 0, CPT_ERR_INVALID, "cast_bf16: bad arguments");
  return tc::cast_to_bf16(src, dst, rows, cols, as_stream(stream));
}
static int check_lp(int64_t N, int In, int Out, int mode, const char* who) {
  CPT_REQUIRE(N > 0 && In > 0 && Out > 0 && N < (1LL << 31), CPT_ERR_INVALID, "%s: bad dimensions", who);
  CPT_REQUIRE(mode == CPT_MODE_BF16, CPT_ERR_INVALID, "%s: pre-cast operands are bf16 (mode must be CPT_MODE_BF16)", who);
  return CPT_OK;
}
int cpt_linear_fwd_bf16(const void* x_bf, const void* w_bf, const float* bias, float* y, int64_t N, int In, int Out, void* stream) {
  if (int e = check_lp(N, In, Out, CPT_MODE_BF16, "linear_fwd_bf16")) return e;
  return tc::linear_fwd_lp(x_bf, w_bf, bias, y, N, In, Out, CPT_MODE_BF16, as_stream(stream));
}
int cpt_linear_relu_fwd_bf16(const void* x_bf, const void* w_bf, const float* bias, float* y, void* y_bf16, uint8_t* mask, int64_t N,
                             int In, int Out, void* stream) {
  if (int e = check_lp(N, In, Out, CPT_MODE_BF16, "linear_relu_fwd_bf16")) return e;
  CPT_REQUIRE(Out % 32 == 0, CPT_ERR_UNSUPPORTED, "linear_relu_fwd_bf16: the fused ReLU needs Out %% 32 == 0");
  CPT_REQUIRE(!mask || (reinterpret_cast<uintptr_t>(mask) & 3) == 0, CPT_ERR_INVALID, "linear_relu_fwd_bf16: mask must be 4-byte aligned");
  return tc::linear_fwd_lp(x_bf, w_bf, bias, y, N, In, Out, CPT_MODE_BF16, as_stream(stream), 1, y_bf16,
                           reinterpret_cast<unsigned int*>(mask));
}
int cpt_linear_dgrad_bf16(const void* dy_bf, const void* w_bf, float* dx, int64_t N, int In, int Out, void* stream) {
  if (int e = check_lp(N, In, Out, CPT_MODE_BF16, "linear_dgrad_bf16")) return e;
  return tc::linear_dgrad_lp(dy_bf, w_bf, dx, N, In, Out, CPT_MODE_BF16, as_stream(stream));
}
int cpt_linear_dgrad_relu_bf16(const void* dy_bf, const void* w_bf, const uint8_t* mask, float* dx, void* dx_bf16, int64_t N, int In,
                               int Out, void* stream) {
  if (int e = check_lp(N, In, Out, CPT_MODE_BF16, "linear_dgrad_relu_bf16")) return e;
  CPT_REQUIRE(mask && In % 32 == 0, CPT_ERR_UNSUPPORTED, "linear_dgrad_relu_bf16: needs the mask and In %% 32 == 0");
  CPT_REQUIRE((reinterpret_cast<uintptr_t>(mask) & 3) == 0, CPT_ERR_INVALID, "linear_dgrad_relu_bf16: mask must be 4-byte aligned");
  return tc::linear_dgrad_lp(dy_bf, w_bf, dx, N, In, Out, CPT_MODE_BF16, as_stream(stream), reinterpret_cast<const unsigned int*>(mask),
                             dx_bf16);
}
int cpt_linear_wgrad_bf16(const void* x_bf, const void* dy_bf, float* dw, int64_t N, int In, int Out, void* ws, size_t ws_bytes,
                          void* stream) {
  if (int e = check_lp(N, In, Out, CPT_MODE_BF16, "linear_wgrad_bf16")) return e;
  return tc::linear_wgrad_lp(x_bf, dy_bf, dw, N, In, Out, CPT_MODE_BF16, ws, ws_bytes, as_stream(stream));
}
/* db = dy.sum(leading) on its own (linear_funcs.py:33): out[c] = sum_n x[n][c][hw] */
int cpt_channel_sum(const float* x, float* out, int N, int C, int HW, void* ws, size_t ws_bytes, void* stream) {
  CPT_REQUIRE(x && out && N > 0 && C > 0 && HW > 0, CPT_ERR_INVALID, "channel_sum: bad arguments");
  CPT_REQUIRE(ws && ws_bytes >= (size_t)C * 64 * sizeof(float), CPT_ERR_WORKSPACE, "channel_sum: workspace too small");
  return channel_sum(x, out, N, C, HW, ws, as_stream(stream));
}

int cpt_tc_reserve_sms(int n) {
  CPT_REQUIRE(n >= 0 && n < sm_count(), CPT_ERR_INVALID, "tc_reserve_sms: %d outside [0, %d)", n, sm_count());
  tc::g_reserved_sms = n & ~1;  // keep SM pairs whole for the 2-CTA kernels
  return CPT_OK;
}

// Debug/test helper: synchronises the device and returns the pipeline-timeout flag of the tensor-core kernels
// (0 = healthy), clearing it.
int cpt_tc_check_status(void) {
  int v = 0, zero = 0;
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(&v, tc::g_tc_status, sizeof(int)) != cudaSuccess) return -1;
  cudaMemcpyToSymbol(tc::g_tc_status, &zero, sizeof(int));
  return v;
}

}  // extern "C"
