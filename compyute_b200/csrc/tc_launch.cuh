// Per-mode launch units of the tcgen05 kernel: each compute mode's ~25 instantiations of tc_kernel<> live in their own
// translation unit (tc_inst_bf16.cu / tc_inst_tf32.cu / tc_inst_x3.cu) so they compile in parallel.
#pragma once
#include "common.cuh"
#include "tc_params.cuh"

namespace cpt {
namespace tc {

struct LaunchSel {
  bool a_mn, b_mn;   // operand layouts (MN-major = the reduction dimension is the strided one)
  int op;            // OP_GEMM / OP_CONV / OP_WGRAD
  int BN;            // 64 / 128 / 256
  bool use2;         // cta_group::2 (256-row tiles); p.m_tiles already counts 256-row tiles
  int groups_max;    // CTAs (1-CTA) or CTA pairs (2-CTA) the persistent grid may use
};

int launch_bf16(const TcParams& p, const LaunchSel& s, cudaStream_t st);
int launch_tf32(const TcParams& p, const LaunchSel& s, cudaStream_t st);
int launch_x3(const TcParams& p, const LaunchSel& s, cudaStream_t st);   // fp32 operands as tf32 hi + lo planes, 3 MMAs

}  // namespace tc
}  // namespace cpt
