// Microbenchmark of the sequential block sweep y_k = b_k - L_k y_{k-1} (16x16 blocks, one warp,
// shared memory) in a few formulations, cycles per stage.  nvcc -arch=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
constexpr int BS = 16, LD = 18, NB = 20;

// V0: as in tri_fast.cuh (2 lanes per row, shuffle combine, read-modify-write of dst)
__global__ void v0(const double* Lg, double* bg, long long* cyc, int reps) {
  __shared__ __align__(16) double L[NB * BS * LD];
  __shared__ __align__(16) double b[NB * BS];
  for (int i = threadIdx.x; i < NB * BS * LD; i += 32) L[i] = Lg[i];
  for (int i = threadIdx.x; i < NB * BS; i += 32) b[i] = bg[i];
  __syncwarp();
  const int lane = threadIdx.x, row = lane >> 1, half = lane & 1;
  long long t0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
    double La[8];
    { const double2* p = reinterpret_cast<const double2*>(L + 1 * BS * LD + row * LD + half * 8);
#pragma unroll
      for (int i = 0; i < 4; ++i) { double2 v = p[i]; La[2 * i] = v.x; La[2 * i + 1] = v.y; } }
    for (int k = 1; k < NB; ++k) {
      double Ln[8];
      if (k + 1 < NB) { const double2* p = reinterpret_cast<const double2*>(L + (k + 1) * BS * LD + row * LD + half * 8);
#pragma unroll
        for (int i = 0; i < 4; ++i) { double2 v = p[i]; Ln[2 * i] = v.x; Ln[2 * i + 1] = v.y; } }
      const double* src = b + (k - 1) * BS + half * 8;
      double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
      for (int c = 0; c < 8; c += 4) { s0 = fma(La[c], src[c], s0); s1 = fma(La[c + 1], src[c + 1], s1); s2 = fma(La[c + 2], src[c + 2], s2); s3 = fma(La[c + 3], src[c + 3], s3); }
      double s = (s0 + s1) + (s2 + s3);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      if (half == 0) b[k * BS + row] -= s;
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) La[i] = Ln[i];
    }
  }
  long long t1 = clock64();
  if (lane == 0) *cyc = t1 - t0;
  for (int i = threadIdx.x; i < NB * BS; i += 32) bg[i] = b[i];
}

// V1: 1 lane per row (16 lanes), 4 chains of 4, dst value preloaded, y read as double2 broadcast
__global__ void v1(const double* Lg, double* bg, long long* cyc, int reps) {
  __shared__ __align__(16) double L[NB * BS * LD];
  __shared__ __align__(16) double b[NB * BS];
  for (int i = threadIdx.x; i < NB * BS * LD; i += 32) L[i] = Lg[i];
  for (int i = threadIdx.x; i < NB * BS; i += 32) b[i] = bg[i];
  __syncwarp();
  const int lane = threadIdx.x, row = lane & 15;
  long long t0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
    double La[16];
    { const double2* p = reinterpret_cast<const double2*>(L + 1 * BS * LD + row * LD);
#pragma unroll
      for (int i = 0; i < 8; ++i) { double2 v = p[i]; La[2 * i] = v.x; La[2 * i + 1] = v.y; } }
    double bk = b[1 * BS + row];
    for (int k = 1; k < NB; ++k) {
      double Ln[16]; double bn = 0;
      if (k + 1 < NB) { const double2* p = reinterpret_cast<const double2*>(L + (k + 1) * BS * LD + row * LD);
#pragma unroll
        for (int i = 0; i < 8; ++i) { double2 v = p[i]; Ln[2 * i] = v.x; Ln[2 * i + 1] = v.y; }
        bn = b[(k + 1) * BS + row]; }
      const double2* src = reinterpret_cast<const double2*>(b + (k - 1) * BS);
      double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
      for (int c = 0; c < 8; c += 2) { double2 y0 = src[c], y1 = src[c + 1];
        s0 = fma(La[2 * c], y0.x, s0); s1 = fma(La[2 * c + 1], y0.y, s1); s2 = fma(La[2 * c + 2], y1.x, s2); s3 = fma(La[2 * c + 3], y1.y, s3); }
      const double y = bk - ((s0 + s1) + (s2 + s3));
      if (lane < 16) b[k * BS + row] = y;
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 16; ++i) La[i] = Ln[i];
      bk = bn;
    }
  }
  long long t1 = clock64();
  if (lane == 0) *cyc = t1 - t0;
  for (int i = threadIdx.x; i < NB * BS; i += 32) bg[i] = b[i];
}

// V2: like V1 but y broadcast through shuffles from registers (no shared-memory round trip)
__global__ void v2(const double* Lg, double* bg, long long* cyc, int reps) {
  __shared__ __align__(16) double L[NB * BS * LD];
  __shared__ __align__(16) double b[NB * BS];
  for (int i = threadIdx.x; i < NB * BS * LD; i += 32) L[i] = Lg[i];
  for (int i = threadIdx.x; i < NB * BS; i += 32) b[i] = bg[i];
  __syncwarp();
  const int lane = threadIdx.x, row = lane & 15;
  long long t0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
    double La[16];
    { const double2* p = reinterpret_cast<const double2*>(L + 1 * BS * LD + row * LD);
#pragma unroll
      for (int i = 0; i < 8; ++i) { double2 v = p[i]; La[2 * i] = v.x; La[2 * i + 1] = v.y; } }
    double bk = b[1 * BS + row];
    double y = b[row];   // y_0
    for (int k = 1; k < NB; ++k) {
      double Ln[16]; double bn = 0;
      if (k + 1 < NB) { const double2* p = reinterpret_cast<const double2*>(L + (k + 1) * BS * LD + row * LD);
#pragma unroll
        for (int i = 0; i < 8; ++i) { double2 v = p[i]; Ln[2 * i] = v.x; Ln[2 * i + 1] = v.y; }
        bn = b[(k + 1) * BS + row]; }
      double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
      for (int c = 0; c < 16; c += 4) {
        s0 = fma(La[c], __shfl_sync(0xffffffffu, y, c), s0); s1 = fma(La[c + 1], __shfl_sync(0xffffffffu, y, c + 1), s1);
        s2 = fma(La[c + 2], __shfl_sync(0xffffffffu, y, c + 2), s2); s3 = fma(La[c + 3], __shfl_sync(0xffffffffu, y, c + 3), s3); }
      y = bk - ((s0 + s1) + (s2 + s3));
      if (lane < 16) b[k * BS + row] = y;
#pragma unroll
      for (int i = 0; i < 16; ++i) La[i] = Ln[i];
      bk = bn;
    }
    __syncwarp();
  }
  long long t1 = clock64();
  if (lane == 0) *cyc = t1 - t0;
  for (int i = threadIdx.x; i < NB * BS; i += 32) bg[i] = b[i];
}

int main() {
  double *L, *b; long long* c; long long h;
  cudaMalloc(&L, NB * BS * LD * 8); cudaMalloc(&b, NB * BS * 8); cudaMalloc(&c, 8);
  cudaMemset(L, 0, NB * BS * LD * 8); cudaMemset(b, 0, NB * BS * 8);
  const int reps = 200;
  for (int w = 0; w < 2; ++w) {
    v0<<<1, 32>>>(L, b, c, reps); cudaDeviceSynchronize(); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    if (w) printf("V0 (2 lanes/row, shuffle, RMW dst)          %7.1f cycles/stage\n", double(h) / reps / (NB - 1));
    v1<<<1, 32>>>(L, b, c, reps); cudaDeviceSynchronize(); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    if (w) printf("V1 (1 lane/row, double2 y, dst preloaded)   %7.1f cycles/stage\n", double(h) / reps / (NB - 1));
    v2<<<1, 32>>>(L, b, c, reps); cudaDeviceSynchronize(); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    if (w) printf("V2 (1 lane/row, y via shuffles)             %7.1f cycles/stage\n", double(h) / reps / (NB - 1));
  }
  return 0;
}
