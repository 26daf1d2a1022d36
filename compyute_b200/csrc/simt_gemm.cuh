// Exact-fp32 implicit-GEMM engine on CUDA cores (FFMA), used by CPT_MODE_FP32 for Conv2D and Linear and as
// the general fallback for shapes the tcgen05 kernels do not cover.  128x128x8 CTA tile, 256 threads, 8x8
// register micro-tile, double-buffered shared memory.  A "problem" functor supplies operand gathers and the
// output scatter, so fprop / dgrad / wgrad / plain GEMM share one mainloop (SURVEY Appendix D mappings).
#pragma once
#include "common.cuh"

namespace cpt {

constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 8, SG_THREADS = 256;

// Problem concept:
//   int M, N, K;                         GEMM extents (K = reduction)
//   static constexpr bool A_MN_CONTIG;   consecutive m are contiguous in memory for fixed k (else consecutive k)
//   static constexpr bool B_MN_CONTIG;   same for B / n
//   static constexpr bool OUT_M_CONTIG;  store4 receives 4 values along m (else along n)
//   RowA rowA(int m); float loadA(const RowA&, int k);
//   ColB colB(int n); float loadB(const ColB&, int k);
//   void store4(int split, int m, int n, const float v[4]);
template <class P>
__global__ void __launch_bounds__(SG_THREADS, 2) simt_gemm_kernel(const P p, int k_per_split) {
  __shared__ __align__(16) float As[2][SG_BK][SG_BM];
  __shared__ __align__(16) float Bs[2][SG_BK][SG_BN];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * SG_BM, n0 = blockIdx.y * SG_BN, split = blockIdx.z;
  const int kb = split * k_per_split;
  const int ke = min(p.K, kb + k_per_split);

  // ---- loader roles
  typename P::RowA ra[P::A_MN_CONTIG ? 1 : 4];
  typename P::ColB cb[P::B_MN_CONTIG ? 1 : 4];
  if (P::A_MN_CONTIG) ra[0] = p.rowA(m0 + (tid & 127));
  else {
#pragma unroll
    for (int i = 0; i < 4; ++i) ra[P::A_MN_CONTIG ? 0 : i] = p.rowA(m0 + (tid >> 3) + 32 * i);
  }
  if (P::B_MN_CONTIG) cb[0] = p.colB(n0 + (tid & 127));
  else {
#pragma unroll
    for (int i = 0; i < 4; ++i) cb[P::B_MN_CONTIG ? 0 : i] = p.colB(n0 + (tid >> 3) + 32 * i);
  }

  float a_reg[4], b_reg[4];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (P::A_MN_CONTIG) {
        const int k = k0 + (tid >> 7) + 2 * i;
        a_reg[i] = (k < ke) ? p.loadA(ra[0], k) : 0.f;
      } else {
        const int k = k0 + (tid & 7);
        a_reg[i] = (k < ke) ? p.loadA(ra[P::A_MN_CONTIG ? 0 : i], k) : 0.f;
      }
      if (P::B_MN_CONTIG) {
        const int k = k0 + (tid >> 7) + 2 * i;
        b_reg[i] = (k < ke) ? p.loadB(cb[0], k) : 0.f;
      } else {
        const int k = k0 + (tid & 7);
        b_reg[i] = (k < ke) ? p.loadB(cb[P::B_MN_CONTIG ? 0 : i], k) : 0.f;
      }
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (P::A_MN_CONTIG) As[buf][(tid >> 7) + 2 * i][tid & 127] = a_reg[i];
      else As[buf][tid & 7][(tid >> 3) + 32 * i] = a_reg[i];
      if (P::B_MN_CONTIG) Bs[buf][(tid >> 7) + 2 * i][tid & 127] = b_reg[i];
      else Bs[buf][tid & 7][(tid >> 3) + 32 * i] = b_reg[i];
    }
  };

  // ---- compute roles: lanes run along the output's contiguous dimension
  const int mi = P::OUT_M_CONTIG ? (tid & 15) : (tid >> 4);
  const int ni = P::OUT_M_CONTIG ? (tid >> 4) : (tid & 15);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  if (kb < ke) {
    gload(kb);
    sstore(0);
  }
  __syncthreads();
  int buf = 0;
  for (int k0 = kb; k0 < ke; k0 += SG_BK) {
    const bool has_next = (k0 + SG_BK) < ke;
    if (has_next) gload(k0 + SG_BK);
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][mi * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + mi * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][ni * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + ni * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (has_next) {
      sstore(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }

  // ---- epilogue
#pragma unroll
  for (int rh = 0; rh < 2; ++rh)
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      const int mb = m0 + rh * 64 + mi * 4, nb = n0 + ch * 64 + ni * 4;
      if (P::OUT_M_CONTIG) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float v[4] = {acc[rh * 4 + 0][ch * 4 + j], acc[rh * 4 + 1][ch * 4 + j], acc[rh * 4 + 2][ch * 4 + j],
                              acc[rh * 4 + 3][ch * 4 + j]};
          p.store4(split, mb, nb + j, v);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float v[4] = {acc[rh * 4 + i][ch * 4 + 0], acc[rh * 4 + i][ch * 4 + 1], acc[rh * 4 + i][ch * 4 + 2],
                              acc[rh * 4 + i][ch * 4 + 3]};
          p.store4(split, mb + i, nb, v);
        }
      }
    }
}

// number of K-splits so that a small output still fills the machine (deterministic: partials are reduced in
// a fixed order by reduce_splits_kernel, never with float atomics)
inline int sg_pick_splits(int M, int N, int K) {
  const int64_t tiles = (int64_t)((M + SG_BM - 1) / SG_BM) * ((N + SG_BN - 1) / SG_BN);
  int64_t s = (2LL * sm_count() + tiles - 1) / tiles;
  const int64_t max_by_k = (K + 255) / 256;  // at least 256 reduction steps per split
  if (s > max_by_k) s = max_by_k;
  if (s > 128) s = 128;
  if (s < 1) s = 1;
  return (int)s;
}

template <class P>
inline cudaError_t sg_launch(const P& p, int splits, cudaStream_t st) {
  if (p.M <= 0 || p.N <= 0) return cudaSuccess;
  int k_tiles = (p.K + SG_BK - 1) / SG_BK;
  int k_per_split = ((k_tiles + splits - 1) / splits) * SG_BK;
  if (k_per_split < SG_BK) k_per_split = SG_BK;
  dim3 grid((p.M + SG_BM - 1) / SG_BM, (p.N + SG_BN - 1) / SG_BN, splits);
  simt_gemm_kernel<P><<<grid, SG_THREADS, 0, st>>>(p, k_per_split);
  count_launch();
  return cudaGetLastError();
}

// out[i] = Σ_s partial[s][i]  (+ optional layout-preserving), fixed order
void launch_reduce_splits(const float* partial, float* out, int64_t n, int splits, cudaStream_t st);

// out[c] = Σ_{n,hw} x[n][c][hw]; ws needs C*64 float2
int channel_sum(const float* x, float* out, int N, int C, int HW, void* ws, cudaStream_t st);

}  // namespace cpt
