// Instantiation of admm_direct_kernel<true> (see direct_launch.h).
#include "admm_direct_kernel.cuh"
#include "direct_launch.h"

namespace ocpb200 {
namespace direct {

cudaError_t kernel_info_smem(KernelInfo* out) {
  cudaFuncAttributes fa{};
  cudaError_t e = cudaFuncGetAttributes(&fa, admm_direct_kernel<true>);
  if (e == cudaSuccess) { out->static_smem = static_cast<int>(fa.sharedSizeBytes); out->regs = fa.numRegs; }
  return e;
}
cudaError_t set_max_dynamic_smem_smem(int bytes) {
  return cudaFuncSetAttribute(admm_direct_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}
cudaError_t occupancy_smem(int threads, int dyn_smem, int* per_sm) {
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, admm_direct_kernel<true>, threads, dyn_smem);
}
cudaError_t launch_smem(int grid, int threads, int dyn_smem, cudaStream_t st, const PatternDev& P,
                      const ocp_b200_settings& S, const SolveArgs& A, uint32_t smem_mask) {
  admm_direct_kernel<true><<<grid, threads, dyn_smem, st>>>(P, S, A, smem_mask);
  return cudaGetLastError();
}

size_t plan_array_doubles(const PatternDev& P, int id) { return array_doubles(P, id); }
int plan_array_count() { return AR_COUNT; }
int block_threads() { return kDirectThreads; }

cudaError_t kernel_info_mixed(KernelInfo* out);
cudaError_t set_max_dynamic_smem_mixed(int bytes);
cudaError_t occupancy_mixed(int threads, int dyn_smem, int* per_sm);
cudaError_t launch_mixed(int grid, int threads, int dyn_smem, cudaStream_t st, const PatternDev& P,
                         const ocp_b200_settings& S, const SolveArgs& A, uint32_t smem_mask);

cudaError_t kernel_info(bool all_smem, KernelInfo* out) { return all_smem ? kernel_info_smem(out) : kernel_info_mixed(out); }
cudaError_t set_max_dynamic_smem(bool all_smem, int bytes) {
  return all_smem ? set_max_dynamic_smem_smem(bytes) : set_max_dynamic_smem_mixed(bytes);
}
cudaError_t occupancy(bool all_smem, int threads, int dyn_smem, int* per_sm) {
  return all_smem ? occupancy_smem(threads, dyn_smem, per_sm) : occupancy_mixed(threads, dyn_smem, per_sm);
}
cudaError_t launch(bool all_smem, int grid, int threads, int dyn_smem, cudaStream_t st, const PatternDev& P,
                   const ocp_b200_settings& S, const SolveArgs& A, uint32_t smem_mask) {
  return all_smem ? launch_smem(grid, threads, dyn_smem, st, P, S, A, smem_mask)
                  : launch_mixed(grid, threads, dyn_smem, st, P, S, A, smem_mask);
}

}  // namespace direct
}  // namespace ocpb200
