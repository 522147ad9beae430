// Instantiations of compact::admm_compact_kernel<BS, 128, 4> (admm_compact_kernel.cuh): the throughput
// plan for OCP-shaped QPs that are small enough for three or four CTAs per SM.
#include "admm_compact_kernel.cuh"
#include "direct_launch.h"

namespace ocpb200 {
namespace compact {

constexpr int kThreads = 128, kBlocks = 4;

template <int BS>
static cudaError_t info_t(direct::KernelInfo* out) {
  cudaFuncAttributes fa{};
  cudaError_t e = cudaFuncGetAttributes(&fa, admm_compact_kernel<BS, kThreads, kBlocks>);
  if (e == cudaSuccess) { out->static_smem = static_cast<int>(fa.sharedSizeBytes); out->regs = fa.numRegs; out->threads = kThreads; }
  return e;
}

cudaError_t kernel_info(int bs, direct::KernelInfo* out) { return bs == 16 ? info_t<16>(out) : info_t<20>(out); }
cudaError_t set_max_dynamic_smem(int bs, int bytes) {
  return bs == 16 ? cudaFuncSetAttribute(admm_compact_kernel<16, kThreads, kBlocks>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)
                  : cudaFuncSetAttribute(admm_compact_kernel<20, kThreads, kBlocks>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}
cudaError_t occupancy(int bs, int dyn_smem, int* per_sm) {
  return bs == 16 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, admm_compact_kernel<16, kThreads, kBlocks>, kThreads, dyn_smem)
                  : cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, admm_compact_kernel<20, kThreads, kBlocks>, kThreads, dyn_smem);
}
cudaError_t launch(int bs, int grid, int dyn_smem, cudaStream_t st, const PatternDev& P, const CompactIdx& C,
                   const ocp_b200_settings& S, const SolveArgs& A) {
  if (bs == 16) admm_compact_kernel<16, kThreads, kBlocks><<<grid, kThreads, dyn_smem, st>>>(P, C, S, A);
  else admm_compact_kernel<20, kThreads, kBlocks><<<grid, kThreads, dyn_smem, st>>>(P, C, S, A);
  return cudaGetLastError();
}
void plan_sizes(const PatternDev& P, int arena_words, size_t* smem_doubles, size_t* slab_doubles, bool* ok) {
  const Layout L = make_layout(P, arena_words);
  *smem_doubles = L.smem_doubles; *slab_doubles = L.slab_doubles; *ok = L.ok;
}

}  // namespace compact
}  // namespace ocpb200
