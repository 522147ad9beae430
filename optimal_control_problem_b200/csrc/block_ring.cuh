// block_ring.cuh -- asynchronous block ring of the sweeps over a slab-resident factor.
#pragma once
#include <cstdint>

namespace ocpb200 {
namespace direct {

// ---- asynchronous block ring -------------------------------------------------------------
// When the factor lives in the global slab and a block row no longer fits a three-deep register
// prefetch, each chain warp streams its blocks through a ring of shared-memory slots with 1-D
// bulk copies (cp.async.bulk, completion counted on one mbarrier per slot): lane 0 keeps
// ring_slots copies in flight, every lane waits on the slot's barrier before reading it.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void ring_issue(double* dst, const double* src, uint32_t bytes, unsigned long long* bar) {
  const uint32_t b = smem_addr(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void ring_wait(unsigned long long* bar, uint32_t parity) {
  const uint32_t b = smem_addr(bar);
  uint32_t ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(b), "r"(parity) : "memory");
  } while (!ok);
}

}  // namespace direct
}  // namespace ocpb200
