// Instantiation + launch of tc_kernel<> for one compute mode (included by tc_inst_*.cu only).
#pragma once
#include "tc_kernel.cuh"
#include "tc_launch.cuh"

namespace cpt {
namespace tc {

template <bool BF16, bool X3, bool A_MN, bool B_MN, int BN, int OP, bool CTA2>
static int launch_inst(const TcParams& p, const LaunchSel& s, cudaStream_t st) {
  using S = StageCfg<BN, CTA2, X3>;
  static bool configured = false;
  auto kern = tc_kernel<BF16, X3, A_MN, B_MN, BN, OP, CTA2>;
  if (!configured) {
    CPT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::SMEM_BYTES));
    configured = true;
  }
  const int total = p.m_tiles * p.n_tiles * p.z_tiles;  // tiles (1-CTA) or tile pairs (2-CTA)
  const int ncta = CTA2 ? 2 : 1;
  int groups = s.groups_max;
  if (groups < 1) groups = 1;
  if (total < groups) groups = total;
  if (groups < 1) return CPT_OK;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(groups * ncta);
  cfg.blockDim = dim3(EpiCfg<BN, X3>::THREADS);
  cfg.dynamicSmemBytes = S::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = ncta;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CTA2 ? 1 : 0;
  CPT_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  CPT_LAUNCH_CHECK("tc_kernel");
  return CPT_OK;
}

template <bool BF16, bool X3, bool A_MN, bool B_MN, int OP>
static int launch_layout(const TcParams& p, const LaunchSel& s, cudaStream_t st) {
  if constexpr (!X3) {  // the X3 epilogue keeps BN accumulators per thread in registers: BN <= 128 (see pick_bn_mode)
    if (s.BN == 256) return s.use2 ? launch_inst<BF16, X3, A_MN, B_MN, 256, OP, true>(p, s, st) : launch_inst<BF16, X3, A_MN, B_MN, 256, OP, false>(p, s, st);
  } else {
    CPT_REQUIRE(s.BN <= 128, CPT_ERR_INVALID, "tc launch: FP32X3 tiles are at most 128 columns wide");
  }
  if constexpr (!A_MN && !B_MN && !X3) {  // K-major operands only: half of a 64-column MN-major B tile is less than one swizzle chunk
    if (s.use2 && s.BN == 64) return launch_inst<BF16, X3, A_MN, B_MN, 64, OP, true>(p, s, st);
  }
  CPT_REQUIRE(!s.use2 || s.BN == 128, CPT_ERR_INVALID, "tc launch: no cta_group::2 instantiation for %d-column tiles of this layout", s.BN);
  if (s.use2) return launch_inst<BF16, X3, A_MN, B_MN, 128, OP, true>(p, s, st);
  if (s.BN == 128) return launch_inst<BF16, X3, A_MN, B_MN, 128, OP, false>(p, s, st);
  return launch_inst<BF16, X3, A_MN, B_MN, 64, OP, false>(p, s, st);
}

// the five (layout, op) combinations the host drivers use
template <bool BF16, bool X3>
static int launch_mode(const TcParams& p, const LaunchSel& s, cudaStream_t st) {
  if (s.op == OP_CONV && !s.a_mn && !s.b_mn) return launch_layout<BF16, X3, false, false, OP_CONV>(p, s, st);
  if (s.op == OP_WGRAD && s.a_mn && s.b_mn) return launch_layout<BF16, X3, true, true, OP_WGRAD>(p, s, st);
  if (s.op == OP_GEMM && !s.a_mn && !s.b_mn) return launch_layout<BF16, X3, false, false, OP_GEMM>(p, s, st);
  if (s.op == OP_GEMM && s.a_mn && !s.b_mn) return launch_layout<BF16, X3, true, false, OP_GEMM>(p, s, st);
  if (s.op == OP_GEMM && s.a_mn && s.b_mn) return launch_layout<BF16, X3, true, true, OP_GEMM>(p, s, st);
  CPT_REQUIRE(false, CPT_ERR_INVALID, "tc launch: no instantiation for op %d layouts %d/%d", s.op, (int)s.a_mn, (int)s.b_mn);
}

}  // namespace tc
}  // namespace cpt
