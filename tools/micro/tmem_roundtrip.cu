// tmem_roundtrip.cu -- Tensor Memory as per-warp scratch for FP64 data (no MMA involved):
// tcgen05.alloc / tcgen05.st.32x32b / tcgen05.ld.32x32b / tcgen05.dealloc, four warps, each in its own lane quadrant.
// Checks the round trip bit for bit and times a dependent tcgen05.ld -> FMA chain against the same chain fed from
// shared memory.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_roundtrip tmem_roundtrip.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kCols = 128;        // columns allocated per CTA (power of two >= 32)
constexpr int kBlocks = 8;        // 16-column blocks per quadrant

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
                 "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(128, 4) roundtrip(int* errors, long long* cycles, double* sink) {
  __shared__ uint32_t tmem_base;
  __shared__ __align__(16) double sm_blk[4][kBlocks / 2][32][8];   // (static shared memory limit: half of the blocks keep a copy)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(&tmem_base))), "n"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_base + (static_cast<uint32_t>(32 * warp) << 16);
  // store: block j of this warp, lane's 8 doubles
  for (int j = 0; j < kBlocks; ++j) {
    uint32_t v[16];
    for (int c = 0; c < 8; ++c) {
      const double d = 1.0 + 1e-3 * (blockIdx.x * 100000 + warp * 10000 + j * 1000 + lane * 10 + c);
      sm_blk[warp][j & 3][lane][c] = d;
      v[2 * c] = __double2loint(d); v[2 * c + 1] = __double2hiint(d);
    }
    tmem_st16(base + 16 * j, v);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  __syncwarp();
  int bad = 0;
  for (int j = 0; j < kBlocks; ++j) {
    uint32_t v[16];
    tmem_ld16(base + 16 * j, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int c = 0; c < 8; ++c) {
      const double d = __hiloint2double(v[2 * c + 1], v[2 * c]);
      const double want = 1.0 + 1e-3 * (blockIdx.x * 100000 + warp * 10000 + j * 1000 + lane * 10 + c);
      if (__double_as_longlong(d) != __double_as_longlong(want)) ++bad;
    }
  }
  if (bad) atomicAdd(errors, bad);
  // timing: dependent chain  acc = fma(L[c], acc, ...) over the 8 blocks, TMEM-fed vs shared-memory-fed
  double acc = 1.0;
  __syncthreads();
  long long t0 = clock64();
  for (int rep = 0; rep < 64; ++rep)
    for (int j = 0; j < kBlocks; ++j) {
      uint32_t v[16];
      tmem_ld16(base + 16 * j, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int c = 0; c < 8; ++c) acc = fma(__hiloint2double(v[2 * c + 1], v[2 * c]), 1e-9, acc);
    }
  long long t1 = clock64();
  for (int rep = 0; rep < 64; ++rep)
    for (int j = 0; j < kBlocks; ++j) {
      const double2* p = reinterpret_cast<const double2*>(&sm_blk[warp][j & 3][lane][0]);
#pragma unroll
      for (int c = 0; c < 4; ++c) { const double2 d = p[c]; acc = fma(d.x, 1e-9, acc); acc = fma(d.y, 1e-9, acc); }
    }
  long long t2 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { cycles[0] = (t1 - t0) / (64 * kBlocks); cycles[1] = (t2 - t1) / (64 * kBlocks); }
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kCols) : "memory");
}

int main() {
  int* err; long long* cyc; double* sink;
  const int grid = 148 * 4;
  cudaMalloc(&err, sizeof(int)); cudaMalloc(&cyc, 2 * sizeof(long long)); cudaMalloc(&sink, grid * 128 * sizeof(double));
  cudaMemset(err, 0, sizeof(int));
  for (int g : {1, grid}) {
    roundtrip<<<g, 128>>>(err, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    int h = -1; long long c[2] = {0, 0};
    cudaMemcpy(&h, err, sizeof(int), cudaMemcpyDeviceToHost); cudaMemcpy(c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    printf("grid %4d: %s, mismatches %d, per 16-column block (8 doubles per lane + 8 dependent DFMA): tcgen05.ld %lld cycles, ld.shared %lld cycles\n",
           g, cudaGetErrorString(e), h, c[0], c[1]);
  }
  return 0;
}
