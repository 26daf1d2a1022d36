// tc_kernel<> instantiations of one compute mode (see tc_launch.cuh)
#include "tc_inst.cuh"

namespace cpt {
namespace tc {
int launch_bf16(const TcParams& p, const LaunchSel& s, cudaStream_t st) { return launch_mode<true, false>(p, s, st); }
}  // namespace tc
}  // namespace cpt
