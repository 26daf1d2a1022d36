// strip_conv_kernel<> instantiations (see strip_kernel.cuh)
#include "common.cuh"
#include "strip_kernel.cuh"

namespace cpt {
namespace tc {

template <int BN>
static int launch_strip_bn(const StripParams& p, int grid, size_t smem_bytes, cudaStream_t st) {
  static bool configured = false;
  auto kern = strip_conv_kernel<BN>;
  if (!configured) {
    CPT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 4 * BN * 2 * 4));
    configured = true;
  }
  kern<<<grid, EpiCfg<BN>::THREADS, smem_bytes, st>>>(p);
  CPT_LAUNCH_CHECK("strip_conv_kernel");
  return CPT_OK;
}

int launch_strip(const StripParams& p, int BN, int grid, size_t smem_bytes, cudaStream_t st) {
  if (BN == 256) return launch_strip_bn<256>(p, grid, smem_bytes, st);
  if (BN == 128) return launch_strip_bn<128>(p, grid, smem_bytes, st);
  return launch_strip_bn<64>(p, grid, smem_bytes, st);
}

}  // namespace tc
}  // namespace cpt
