// Microbenchmark: dependent-chain latency and single-warp throughput of DFMA / DADD / LDS / SHFL /
// shared-memory store->load round trip on the target GPU.  nvcc -arch=sm_100a -O3 fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma_chain(double* out, double a, double b, int n, long long* cyc) {
  double x = out[threadIdx.x];
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; ++i) x = fma(x, a, b);
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_dfma_ilp4(double* out, double a, double b, int n, long long* cyc) {
  double x0 = out[threadIdx.x], x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
  long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < n; ++i) { x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b); }
  long long t1 = clock64();
  out[threadIdx.x] = x0 + x1 + x2 + x3;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_dfma_ilp16(double* out, double a, double b, int n, long long* cyc) {
  double x[16];
  for (int j = 0; j < 16; ++j) x[j] = out[threadIdx.x] + j;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] = fma(x[j], a, b);
  }
  long long t1 = clock64();
  double s = 0;
  for (int j = 0; j < 16; ++j) s += x[j];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_ffma_chain(float* out, float a, float b, int n, long long* cyc) {
  float x = out[threadIdx.x];
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; ++i) x = fmaf(x, a, b);
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_lds_chain(int* out, int n, long long* cyc) {
  __shared__ int s[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (i * 7 + 1) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; ++i) p = s[p];
  long long t1 = clock64();
  out[threadIdx.x] = p;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_shfl_chain(double* out, int n, long long* cyc) {
  double x = out[threadIdx.x];
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; ++i) x += __shfl_xor_sync(0xffffffffu, x, 1);
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_sts_lds_roundtrip(double* out, int n, long long* cyc) {
  __shared__ double s[64];
  double x = out[threadIdx.x];
  s[threadIdx.x] = x;
  __syncwarp();
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    s[threadIdx.x] = x;
    __syncwarp();
    x = s[(threadIdx.x + 1) & 31] + 1.0;
    __syncwarp();
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_bar(double* out, int n, long long* cyc) {
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) { *cyc = t1 - t0; out[0] = 1; }
}
__global__ void k_drcp_chain(double* out, int n, long long* cyc) {
  double x = out[threadIdx.x] + 1.5;
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < n; ++i) x = 1.0 / x + 1.0;
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  double* d; float* f; int* ip; long long* c; long long h;
  cudaMalloc(&d, 4096 * 8); cudaMalloc(&f, 4096 * 4); cudaMalloc(&ip, 4096 * 4); cudaMalloc(&c, 8);
  cudaMemset(d, 0, 4096 * 8); cudaMemset(f, 0, 4096 * 4);
  const int n = 4096;
  auto report = [&](const char* name, double ops) {
    cudaDeviceSynchronize(); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("%-44s %8.2f cycles/op\n", name, double(h) / ops);
  };
  for (int rep = 0; rep < 2; ++rep) {
    k_dfma_chain<<<1, 32>>>(d, 1.0000001, 1e-9, n, c); if (rep) report("DFMA dependent chain (1 warp)", n);
    k_dfma_ilp4<<<1, 32>>>(d, 1.0000001, 1e-9, n, c); if (rep) report("DFMA 4 independent chains, per DFMA", 4.0 * n);
    k_dfma_ilp16<<<1, 32>>>(d, 1.0000001, 1e-9, n, c); if (rep) report("DFMA 16 independent chains, per DFMA", 16.0 * n);
    k_dfma_ilp16<<<1, 256>>>(d, 1.0000001, 1e-9, n, c); if (rep) report("DFMA 16 chains x 8 warps, per warp-DFMA", 16.0 * n);
    k_ffma_chain<<<1, 32>>>(f, 1.0000001f, 1e-9f, n, c); if (rep) report("FFMA dependent chain", n);
    k_lds_chain<<<1, 32>>>(ip, n, c); if (rep) report("LDS dependent chain", n);
    k_shfl_chain<<<1, 32>>>(d, n, c); if (rep) report("SHFL(double)+DADD dependent chain", n);
    k_sts_lds_roundtrip<<<1, 32>>>(d, n, c); if (rep) report("STS->syncwarp->LDS+DADD->syncwarp round trip", n);
    k_bar<<<1, 256>>>(d, n, c); if (rep) report("__syncthreads (256 threads)", n);
    k_bar<<<1, 512>>>(d, n, c); if (rep) report("__syncthreads (512 threads)", n);
    k_drcp_chain<<<1, 32>>>(d, n, c); if (rep) report("1.0/x + 1.0 dependent chain", n);
  }
  return 0;
}
