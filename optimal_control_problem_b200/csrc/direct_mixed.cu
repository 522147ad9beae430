// Instantiation of admm_direct_kernel<false> (see direct_launch.h).
#include "admm_direct_kernel.cuh"
#include "direct_launch.h"

namespace ocpb200 {
namespace direct {

cudaError_t kernel_info_mixed(KernelInfo* out) {
  cudaFuncAttributes fa{};
  cudaError_t e = cudaFuncGetAttributes(&fa, admm_direct_kernel<false>);
  if (e == cudaSuccess) { out->static_smem = static_cast<int>(fa.sharedSizeBytes); out->regs = fa.numRegs; }
  return e;
}
cudaError_t set_max_dynamic_smem_mixed(int bytes) {
  return cudaFuncSetAttribute(admm_direct_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}
cudaError_t occupancy_mixed(int threads, int dyn_smem, int* per_sm) {
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, admm_direct_kernel<false>, threads, dyn_smem);
}
cudaError_t launch_mixed(int grid, int threads, int dyn_smem, cudaStream_t st, const PatternDev& P,
                      const ocp_b200_settings& S, const SolveArgs& A, uint32_t smem_mask) {
  admm_direct_kernel<false><<<grid, threads, dyn_smem, st>>>(P, S, A, smem_mask);
  return cudaGetLastError();
}

}  // namespace direct
}  // namespace ocpb200
