// direct_multi_variant.h -- one instantiation of the throughput-plan kernel
// admm_direct_kernel<PLACE_MULTI, THREADS, BLOCKS_PER_SM>, compiled in its own translation unit.
// DIRECT_MULTI_DEFINE(tag, threads, blocks) defines kernel_info_<tag>, set_max_dynamic_smem_<tag>,
// occupancy_<tag>, launch_<tag>; direct_multi.cu picks one of them per process (OCP_B200_MULTI_VARIANT).
#pragma once
#include "admm_direct_kernel.cuh"
#include "direct_launch.h"

#define DIRECT_MULTI_DEFINE(tag, THREADS, BLOCKS)                                                                     \
  namespace ocpb200 { namespace direct {                                                                              \
  cudaError_t kernel_info_##tag(KernelInfo* out) {                                                                    \
    cudaFuncAttributes fa{};                                                                                          \
    cudaError_t e = cudaFuncGetAttributes(&fa, admm_direct_kernel<PLACE_MULTI, THREADS, BLOCKS>);                     \
    if (e == cudaSuccess) { out->static_smem = static_cast<int>(fa.sharedSizeBytes); out->regs = fa.numRegs; out->threads = THREADS; } \
    return e;                                                                                                         \
  }                                                                                                                   \
  cudaError_t set_max_dynamic_smem_##tag(int bytes) {                                                                 \
    return cudaFuncSetAttribute(admm_direct_kernel<PLACE_MULTI, THREADS, BLOCKS>,                                     \
                                cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);                                  \
  }                                                                                                                   \
  cudaError_t occupancy_##tag(int dyn_smem, int* per_sm) {                                                            \
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, admm_direct_kernel<PLACE_MULTI, THREADS, BLOCKS>,    \
                                                         THREADS, dyn_smem);                                          \
  }                                                                                                                   \
  cudaError_t launch_##tag(int grid, int dyn_smem, cudaStream_t st, const PatternDev& P, const ocp_b200_settings& S,  \
                           const SolveArgs& A, uint32_t smem_mask) {                                                  \
    admm_direct_kernel<PLACE_MULTI, THREADS, BLOCKS><<<grid, THREADS, dyn_smem, st>>>(P, S, A, smem_mask);            \
    return cudaGetLastError();                                                                                        \
  }                                                                                                                   \
  } }

#define DIRECT_MULTI_DECLARE(tag)                                                                                     \
  cudaError_t kernel_info_##tag(KernelInfo* out);                                                                     \
  cudaError_t set_max_dynamic_smem_##tag(int bytes);                                                                  \
  cudaError_t occupancy_##tag(int dyn_smem, int* per_sm);                                                             \
  cudaError_t launch_##tag(int grid, int dyn_smem, cudaStream_t st, const PatternDev& P, const ocp_b200_settings& S,  \
                           const SolveArgs& A, uint32_t smem_mask);
