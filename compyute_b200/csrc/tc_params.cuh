// Launch parameters and tile configuration of the tcgen05 GEMM / implicit-GEMM kernel (shared by the host drivers in
// tc_host.cu and the per-mode instantiation units tc_inst_*.cu).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace cpt {
namespace tc {

enum { OP_GEMM = 0, OP_CONV = 1, OP_WGRAD = 2 };
enum { BIAS_NONE = 0, BIAS_COL = 1, BIAS_LANE = 2 };

struct alignas(64) TcParams {
  CUtensorMap tmA;  // operand A (M side)
  CUtensorMap tmB;  // operand B (N side)
  CUtensorMap tmA2, tmB2;  // X3 mode: the tf32 "lo" planes of A and B (same geometry as tmA / tmB)
  float* out;
  const float* bias;
  int* status;      // device int: set non-zero on a pipeline timeout
  int bias_mode;
  int out_bf16;     // 1: `out` is bf16 (no bias): the packed-K dgrad intermediate dcol
  // fused ReLU of a Linear layer's forward (lanes = output features, columns = rows of the batch; M % 32 == 0): out receives
  // max(acc + bias, 0); relu_lp (optional) the same values as bf16 rows for the next Linear layer; relu_mask (optional) the
  // bits (out > 0) in plain order — element e = col * M + m is bit e % 32 of word e / 32 (consumed by cpt_relu_bwd_plain)
  // relu == 2 is the backward counterpart on a Linear layer's dgrad (lanes = input features): out = acc * mask, with the mask
  // (READ here) of the ReLU that produced this layer's input, relu_lp = the same values as bf16 rows (dy operand of the
  // previous Linear layer's backward)
  // relu == 3: ReLU behind a convolution (lanes = pixels, column bias): out = max(acc + bias, 0), nothing else is written — the
  // backward pass takes the mask from the output itself (out > 0), see cpt_to_channels_last_gated
  int relu;
  void* relu_lp;
  unsigned int* relu_mask;
  float* stats;     // optional [gridDim.x * 4][N][2]: per-epilogue-warp column sums (Σ acc, Σ acc²) of the raw accumulators
                    // (without bias) over the valid lanes — the batch statistics of a BatchNorm that consumes the output
  int M, N;         // valid extents of the lane / column dimensions
  int m_tiles, n_tiles, z_tiles;  // m_tiles counts 128-row (1-CTA) or 256-row (2-CTA) tiles; z = split / tap*split
  int k_iters_total;              // k iterations of the whole reduction (GEMM / WGRAD) or per tile (CONV)
  int k_iters_per_split;
  // epilogue addressing: dst = out + z_off + lane_off(m) + col * col_stride
  long long col_stride, split_stride, tap_stride;
  int lane_is_pixel;              // 1: lane m -> (image b, sub-grid row r, col c): lane_off = b*img_stride + (out_r0 + out_s*r)*out_W + out_c0 + out_s*c
  int px_per_img;                 // pixels per image of the lane / reduction grid (CONV lanes: sub_H*sub_W, WGRAD: Ho*Wo)
  long long img_stride;
  int out_W, out_s, out_r0, out_c0;  // output scatter of CONV lanes (dense fprop/dgrad: out_W = Wo, out_s = 1, r0 = c0 = 0)
  // convolution geometry (im2col coordinates): base pixel of grid position (r, c) = (lower_h + r*trav, lower_w + c*trav)
  int Wo, Ho;                      // width (and, WGRAD, height) of the lane / reduction pixel grid
  int trav, lower_w, lower_h;      // traversal stride and lower corner of the im2col bounding box
  int conv_stride, pad, dil, Kw;   // WGRAD: conv geometry for the tap of this tile
  int b_im2col, b_pad;             // WGRAD: operand B (dy) is a zero-padded channels-last tensor read through an im2col map
                                   // whose bounding box is the un-padded interior (base pixel = output pixel + b_pad)
  int taps, cchunks, wk_cols;      // wk_cols: weight-matrix columns per tap (padded C)
  unsigned short tap_w[64], tap_h[64];  // CONV: im2col offsets of tap t (fprop: kk*dil, j*dil; dgrad: class offsets)
};

template <bool BF16>
struct Elem {
  static constexpr int BYTES = BF16 ? 2 : 4;
  static constexpr int KC = 128 / BYTES;      // elements per 128-byte swizzle row: 64 bf16 / 32 tf32
  static constexpr int UMMA_K = 32 / BYTES;   // 16 bf16 / 8 tf32
  static constexpr int BK = KC;               // reduction elements per stage (also k-rows of an MN-major stage)
};

// Epilogue warps per CTA.  A warp drains its 32 TMEM lanes one column at a time (load, bias, predicated 128-byte store,
// pointer step: ~5 dependent-issue instructions per column and ONE warp per scheduler), ~45 cycles per column: a 128 x 64 tile
// costs ~2800 cycles of epilogue against 9 k-iterations of MMA at C = 64 — the small-channel layers were EPILOGUE-bound
// (ncu round 2: epilogue warps busy 60-70 % of the kernel, tensor pipe 24-30 %).  Tiles of <= 128 columns therefore use two
// warps per TMEM lane quarter, each taking half of the columns; 256-column tiles (tensor-pipe bound, and 160+ registers per
// thread) keep four.  The X3 epilogue accumulates the whole tile in registers and keeps four as well.
template <int BN, bool X3 = false>
struct EpiCfg {
  static constexpr int WARPS = (BN <= 128 && !X3) ? 8 : 4;
  static constexpr int THREADS = 128 + 32 * WARPS;
};

template <int BN, bool CTA2, bool X3 = false>
struct StageCfg {
  static constexpr int A_BYTES = 128 * 128;   // 128 lanes x 128 B (K-major) == (128/KC chunks) x BK rows x 128 B
  static constexpr int BN_CTA = CTA2 ? BN / 2 : BN;  // B columns staged by one CTA
  static constexpr int B_BYTES = BN_CTA * 128;
  static constexpr int PLANE_BYTES = A_BYTES + B_BYTES;
  // X3 (fp32 operands split into tf32 hi + lo planes, three MMAs per k-step): a stage holds [A_hi][B_hi][A_lo][B_lo]
  static constexpr int STAGE_BYTES = PLANE_BYTES * (X3 ? 2 : 1);
  static constexpr int STAGES = (196608 / STAGE_BYTES) > 8 ? 8 : (196608 / STAGE_BYTES);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

}  // namespace tc
}  // namespace cpt
