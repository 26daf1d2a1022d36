// Launch parameters of strip_conv_kernel (strip_kernel.cuh), shared with the host driver in tc_host.cu.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "tc_params.cuh"

namespace cpt {
namespace tc {

constexpr int STRIP_MAX_UNITS = 4;
constexpr int STRIP_MAX_BSTAGES = 40;

struct alignas(64) StripParams {
  CUtensorMap tmA;   // act_pad as [R][Cp] bf16, box = {64 channels, box_rows}
  CUtensorMap tmB;   // filters [N][T*Ck] bf16, box = {64, BN}
  float* out;        // NCHW fp32 (B, N, H, W)
  const float* bias; // per output channel or NULL
  int* status;
  float* stats;      // optional [gridDim.x * 4][N][2] per-epilogue-warp column sums (see tc_kernel)
  int M_lanes;       // lanes that can be valid: (B-1)*HpWp + (H-1)*Wp + W
  int N;
  int m_tiles, n_tiles;
  int T, cchunks, wk_cols;   // taps, 64-channel chunks, filter-matrix columns per tap
  int H, W, Wp, HpWp;
  int box_rows, n_loads;     // a strip = n_loads boxes of box_rows rows (>= 128 + (K-1)*(Wp+1))
  int unit_bytes;            // n_loads * box_rows * 128, multiple of 1024
  int n_units, b_stages;     // strip buffers, filter-tile slots
  int resident;              // 1: filter tiles are loaded once (slot = chunk*T + tap)
  long long col_stride, img_stride;  // H*W, N*H*W
  int tap_off[64];           // j*Wp + k
};

int launch_strip(const StripParams& p, int BN, int grid, size_t smem_bytes, cudaStream_t st);

}  // namespace tc
}  // namespace cpt
