// direct_launch.h -- host-side entry points of the two instantiations of admm_direct_kernel,
// each compiled in its own translation unit (they are large; nvcc builds them in parallel).
#pragma once

#include "admm_common.cuh"

namespace ocpb200 {
namespace direct {

struct KernelInfo { int static_smem; int regs; };

// kAllSmem = true: every per-instance array in shared memory (LDS/STS code); false: mixed placement
cudaError_t kernel_info(bool all_smem, KernelInfo* out);
cudaError_t set_max_dynamic_smem(bool all_smem, int bytes);
cudaError_t occupancy(bool all_smem, int threads, int dyn_smem, int* per_sm);
cudaError_t launch(bool all_smem, int grid, int threads, int dyn_smem, cudaStream_t st, const PatternDev& P,
                   const ocp_b200_settings& S, const SolveArgs& A, uint32_t smem_mask);

// plan helpers (host)
size_t plan_array_doubles(const PatternDev& P, int id);
int plan_array_count();
int block_threads();   // threads per CTA the kernel is compiled for

}  // namespace direct
}  // namespace ocpb200
