// BatchNorm apply / backward-apply passes that ALSO emit the channels-last bf16 copy the neighbouring tensor-core convolution
// consumes (included by tc_host.cu, which owns the staging helpers it shares).
//
// Without this, a Conv -> BN -> ReLU -> Conv chain pays a separate NCHW->NHWC staging pass (read 4 B + write 2 B per
// element) in front of every convolution, forward (x_cl) and backward (dy_cl).  Here the producer of the fp32 NCHW
// tensor — the BN(+ReLU) apply kernel in forward, the BN backward-apply kernel in backward — transposes the values it
// already holds in registers through shared memory and writes the bf16 NHWC tile as well (+2 B per element), so the staging
// pass of the consumer disappears.  Tile, thread mapping and bank-conflict-free transpose are those of nchw_to_nhwc_kernel.
//
//   OP 0 (forward):  y  = act(fmaf(w, (x - mean) * rstd, b))                       -> y  (fp32 NCHW) + y_cl  (bf16 NHWC)
//   OP 1 (backward): dx = g * (count * dy' - s1 - (x - mean) * rstd * s2),  dy' = dy * [y > 0] if RELU
//                                                                                   -> dx (fp32 NCHW) + dx_cl (bf16 NHWC)
//                    + optional per-channel sums of dx (the bias gradient of the convolution that produced x)
// The arithmetic expressions are the ones of bn.cu, so results are bit-identical to the unfused kernels.
// (textually included inside namespace cpt::tc of tc_host.cu, after the staging helpers)
#pragma once

__device__ __forceinline__ float cl_relu_fwd(float v) { return (v != v) ? v : fmaxf(v, 0.f); }

struct BnTileParams {
  const float* x;
  const float* dy;      // OP 1
  float* out;           // y / dx, fp32 NCHW
  uint32_t* out_cl;     // bf16 pairs, NHWC with Cp channels per pixel
  const float *mean, *rstd, *w, *b;
  const float* coef;    // OP 1: [C][3] = (g, s1, s2) from bn_bwd_finalize_kernel
  float count;
  int C, HW, Cp;
  long long Q;          // B * HW: pixel tiles run over the flattened (image, pixel) index
  float* partial;       // SUMS: [blocks][groups * 64] per-block channel sums of `out`
};

template <int OP, bool RELU, bool SUMS>
__global__ void __launch_bounds__(256, 2) bn_tile_cl_kernel(const BnTileParams p) {
  __shared__ uint32_t tile[32][CL_PX + 1];
  __shared__ float prm[7][CL_CH];  // mean, rstd, w, b, g, s1, s2 of this block's 64 channels
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.y * CL_CH;
  const int C = p.C, HW = p.HW, Cp = p.Cp;
  const int64_t Q = p.Q, n_tiles = (Q + CL_PX - 1) / CL_PX;
  if (threadIdx.x < CL_CH) {
    const int c = c0 + threadIdx.x;
    const bool ok = c < C;
    prm[0][threadIdx.x] = ok ? p.mean[c] : 0.f;
    prm[1][threadIdx.x] = ok ? p.rstd[c] : 0.f;
    prm[2][threadIdx.x] = ok ? p.w[c] : 0.f;
    prm[3][threadIdx.x] = (ok && p.b) ? p.b[c] : 0.f;
    if (OP == 1) {
      prm[4][threadIdx.x] = ok ? p.coef[3 * c] : 0.f;
      prm[5][threadIdx.x] = ok ? p.coef[3 * c + 1] : 0.f;
      prm[6][threadIdx.x] = ok ? p.coef[3 * c + 2] : 0.f;
    }
  }
  __syncthreads();
  const float* xs = p.x;
  const float* gs = OP == 1 ? p.dy : nullptr;
  float* os = p.out;
  float acc[8];
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
    float v0[4][4], v1[4][4];
    float g0[4][4], g1[4][4];
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) {
      const int c = c0 + 2 * (warp + 8 * ci);
#pragma unroll
      for (int pi = 0; pi < 4; ++pi) {
        const bool okp = soff[pi] != PAST_END;
        v0[ci][pi] = (okp && c < C) ? xs[soff[pi] + (uint32_t)(c * HW)] : 0.f;
        v1[ci][pi] = (okp && c + 1 < C) ? xs[soff[pi] + (uint32_t)((c + 1) * HW)] : 0.f;
        if (OP == 1) {
          g0[ci][pi] = (okp && c < C) ? gs[soff[pi] + (uint32_t)(c * HW)] : 0.f;
          g1[ci][pi] = (okp && c + 1 < C) ? gs[soff[pi] + (uint32_t)((c + 1) * HW)] : 0.f;
        }
      }
    }
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) {
      const int cl = 2 * (warp + 8 * ci), c = c0 + cl;  // channel pair (c, c + 1), local index cl
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float mu = prm[0][cl + h], rs = prm[1][cl + h], ww = prm[2][cl + h], bb = prm[3][cl + h];
        const float gc = OP == 1 ? prm[4][cl + h] : 0.f, s1 = OP == 1 ? prm[5][cl + h] : 0.f, s2 = OP == 1 ? prm[6][cl + h] : 0.f;
        float(&v)[4] = h ? v1[ci] : v0[ci];
        float(&g)[4] = h ? g1[ci] : g0[ci];
#pragma unroll
        for (int pi = 0; pi < 4; ++pi) {
          const float xv = v[pi];
          float r;
          if (OP == 0) {
            r = fmaf(ww, (xv - mu) * rs, bb);
            if (RELU) r = cl_relu_fwd(r);
          } else {
            float gd = g[pi];
            if (RELU) gd = gd * (fmaf(ww, (xv - mu) * rs, bb) > 0.f ? 1.f : 0.f);
            r = gc * (p.count * gd - s1 - (xv - mu) * rs * s2);
          }
          if (soff[pi] != PAST_END && c + h < C) os[soff[pi] + (uint32_t)((c + h) * HW)] = r;
          else r = 0.f;  // padding channels / pixels of the bf16 tile are zero
          v[pi] = r;
        }
        if (SUMS) acc[2 * ci + h] += (v[0] + v[1]) + (v[2] + v[3]);
      }
#pragma unroll
      for (int pi = 0; pi < 4; ++pi) {
        __nv_bfloat162 h2 = __floats2bfloat162_rn(v0[ci][pi], v1[ci][pi]);
        tile[warp + 8 * ci][lane + 32 * pi] = *reinterpret_cast<uint32_t*>(&h2);
      }
    }
    __syncthreads();
    const int cpair = c0 + 2 * lane;  // this lane's channel pair in the transposed store
    if (cpair < Cp) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int pp = warp + 8 * i;
        const int64_t q = p0 + pp;
        if (q < Q) p.out_cl[(q * Cp + cpair) >> 1] = tile[lane][pp];
      }
    }
    __syncthreads();  // tile is reused by the next pixel tile
  }

  if (SUMS) {
    const int64_t blk = blockIdx.x;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float t = warp_sum(acc[i]);
      const int c = c0 + 2 * (warp + 8 * (i >> 1)) + (i & 1);
      if (lane == 0 && c < C) p.partial[blk * ((int64_t)gridDim.y * CL_CH) + c] = t;
    }
  }
}

// host launchers (called from bn.cu through the declarations in common.cuh)
int bn_apply_cl(const float* x, const float* w, const float* b, const float* mean, const float* rstd, float* y, void* y_cl, int N,
                int C, int HW, int act, cudaStream_t st) {
  BnTileParams p{};
  p.x = x; p.out = y; p.out_cl = reinterpret_cast<uint32_t*>(y_cl);
  p.mean = mean; p.rstd = rstd; p.w = w; p.b = b;
  p.C = C; p.HW = HW; p.Cp = round_up(C, 8); p.Q = (long long)N * HW;
  int gx, groups;
  cl_grid(N, C, HW, 1, gx, groups);
  dim3 grid(gx, groups, 1);
  CPT_REQUIRE(grid.y <= 65535 && p.Q < (1LL << 31) && (int64_t)N * C * HW < (1LL << 32) - (1 << 20), CPT_ERR_UNSUPPORTED,
              "bn_apply_cl: tensor too large");
  if (act) bn_tile_cl_kernel<0, true, false><<<grid, 256, 0, st>>>(p);
  else bn_tile_cl_kernel<0, false, false><<<grid, 256, 0, st>>>(p);
  CPT_LAUNCH_CHECK("bn_apply_cl");
  return CPT_OK;
}

size_t bn_bwd_apply_cl_ws(int N, int C, int HW) { return to_channels_last_ws(N, C, HW, 1); }

int bn_bwd_apply_cl(const float* x, const float* dy, const float* w, const float* b, const float* mean, const float* rstd,
                    const float* coef, float count, float* dx, void* dx_cl, float* dx_chan_sum, int N, int C, int HW, int act,
                    void* ws, size_t ws_bytes, cudaStream_t st) {
  BnTileParams p{};
  p.x = x; p.dy = dy; p.out = dx; p.out_cl = reinterpret_cast<uint32_t*>(dx_cl);
  p.mean = mean; p.rstd = rstd; p.w = w; p.b = b; p.coef = coef; p.count = count;
  p.C = C; p.HW = HW; p.Cp = round_up(C, 8); p.Q = (long long)N * HW;
  int gx, groups;
  cl_grid(N, C, HW, 1, gx, groups);
  dim3 grid(gx, groups, 1);
  CPT_REQUIRE(grid.y <= 65535 && p.Q < (1LL << 31) && (int64_t)N * C * HW < (1LL << 32) - (1 << 20), CPT_ERR_UNSUPPORTED,
              "bn_bwd_apply_cl: tensor too large");
  if (dx_chan_sum) {
    CPT_REQUIRE(ws && ws_bytes >= bn_bwd_apply_cl_ws(N, C, HW), CPT_ERR_WORKSPACE, "bn_bwd_apply_cl: workspace too small");
    p.partial = reinterpret_cast<float*>(ws);
    if (act) bn_tile_cl_kernel<1, true, true><<<grid, 256, 0, st>>>(p);
    else bn_tile_cl_kernel<1, false, true><<<grid, 256, 0, st>>>(p);
    CPT_LAUNCH_CHECK("bn_bwd_apply_cl");
    chan_partial_reduce_kernel<<<(C + 31) / 32, 1024, 0, st>>>(p.partial, dx_chan_sum, C, (int64_t)gx, groups * CL_CH);
    CPT_LAUNCH_CHECK("chan_partial_reduce");
  } else {
    if (act) bn_tile_cl_kernel<1, true, false><<<grid, 256, 0, st>>>(p);
    else bn_tile_cl_kernel<1, false, false><<<grid, 256, 0, st>>>(p);
    CPT_LAUNCH_CHECK("bn_bwd_apply_cl");
  }
  return CPT_OK;
}
