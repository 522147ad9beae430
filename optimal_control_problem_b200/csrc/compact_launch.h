// compact_launch.h -- host-side entry points of the compact throughput kernel (direct_compact.cu).
#pragma once
#include "admm_common.cuh"
#include "direct_launch.h"
#include "periodic_index.h"

namespace ocpb200 {
namespace compact {

// variant 0: 128 threads x 4 CTAs/SM, variant 1: 192 threads x 3 CTAs/SM; layout_flags: compact::kQInSmem | kLuInSmem
cudaError_t kernel_info(int bs, int variant, direct::KernelInfo* out);
cudaError_t set_max_dynamic_smem(int bs, int variant, int bytes);
cudaError_t occupancy(int bs, int variant, int dyn_smem, int* per_sm);
cudaError_t launch(int bs, int variant, int grid, int dyn_smem, cudaStream_t st, const PatternDev& P, const CompactIdx& C,
                   const ocp_b200_settings& S, const SolveArgs& A, int layout_flags);
void plan_sizes(const PatternDev& P, int arena_words, int layout_flags, size_t* smem_doubles, size_t* slab_doubles, bool* ok);

}  // namespace compact
}  // namespace ocpb200
