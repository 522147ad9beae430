// Microbenchmark: cycles per pivot of the one-warp register-resident Gauss-Jordan inverse (16x16).
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include "ocp_b200.h"
namespace ocpb200 { struct PatternDev; namespace direct { struct Work; } }
// minimal stand-ins so that tri_fast.cuh's inverse can be included on its own
namespace ocpb200 { namespace direct {
template <int BS, int LD = BS + 2>
__device__ __forceinline__ void inv_a(double* M, int lane, double* piv) {   // current formulation
  constexpr int ld = LD;
  const bool act = lane < BS; const int cc = act ? lane : 0;
  double col[BS];
#pragma unroll
  for (int r = 0; r < BS; ++r) col[r] = M[r * ld + cc];
  const double2* piv2 = reinterpret_cast<const double2*>(piv);
#pragma unroll
  for (int k = 0; k < BS; ++k) {
    const double ck = col[k];
    const double ipiv = 1.0 / __shfl_sync(0xffffffffu, ck, k);
    const bool mine = lane == k;
    if (act) piv[lane] = mine ? ipiv : (lane < k ? ck * ipiv : -ck * ipiv);
    __syncwarp();
#pragma unroll
    for (int r = 0; r < BS / 2; ++r) {
      const double2 f = piv2[r];
      const int r0 = 2 * r, r1 = 2 * r + 1;
      col[r0] = mine ? f.x : (r0 == k ? ck * f.x : fma(f.x, ck, col[r0]));
      col[r1] = mine ? f.y : (r1 == k ? ck * f.y : fma(f.y, ck, col[r1]));
    }
    __syncwarp();
  }
  if (act) {
#pragma unroll
    for (int r = 0; r < BS; ++r) M[r * ld + lane] = col[r];
  }
}
// variant: no selects -- lane k's column is fixed up after the loop body with predicated moves
template <int BS, int LD = BS + 2>
__device__ __forceinline__ void inv_b(double* M, int lane, double* piv) {
  constexpr int ld = LD;
  const bool act = lane < BS; const int cc = act ? lane : 0;
  double col[BS];
#pragma unroll
  for (int r = 0; r < BS; ++r) col[r] = M[r * ld + cc];
  const double2* piv2 = reinterpret_cast<const double2*>(piv);
#pragma unroll
  for (int k = 0; k < BS; ++k) {
    const double ck0 = col[k];
    const double ipiv = 1.0 / __shfl_sync(0xffffffffu, ck0, k);
    const bool mine = lane == k;
    if (act) piv[lane] = mine ? ipiv : (lane < k ? ck0 * ipiv : -ck0 * ipiv);
    // lane k: col <- 0 and ck <- 1, so that the common update col[r] += f[r] * ck yields f[r]
    const double ck = mine ? 1.0 : ck0;
    if (mine) {
#pragma unroll
      for (int r = 0; r < BS; ++r) col[r] = 0.0;
    }
    col[k] = 0.0;   // row k: new value = ck * f[k] (= ck / pivot; lane k: 1 / pivot)
    __syncwarp();
#pragma unroll
    for (int r = 0; r < BS / 2; ++r) {
      const double2 f = piv2[r];
      col[2 * r] = fma(f.x, ck, col[2 * r]);
      col[2 * r + 1] = fma(f.y, ck, col[2 * r + 1]);
    }
    __syncwarp();
  }
  if (act) {
#pragma unroll
    for (int r = 0; r < BS; ++r) M[r * ld + lane] = col[r];
  }
}
}}
using namespace ocpb200::direct;
template <int V>
__global__ void bench(const double* Mg, double* out, long long* cyc, int reps) {
  __shared__ __align__(16) double M[16 * 18];
  __shared__ __align__(16) double piv[32];
  const int lane = threadIdx.x;
  long long total = 0, first = 0;
  for (int rep = 0; rep < reps; ++rep) {
    for (int i = lane; i < 16 * 18; i += 32) M[i] = Mg[i];
    __syncwarp();
    long long t0 = clock64();
    if (V == 0) inv_a<16>(M, lane, piv); else inv_b<16>(M, lane, piv);
    __syncwarp();
    const long long dt = clock64() - t0;
    if (rep == 0) first = dt; else total += dt;
  }
  if (lane == 0) { cyc[0] = total; cyc[1] = first; }
  for (int i = lane; i < 16 * 18; i += 32) out[i] = M[i];
}
int main() {
  double h[16 * 18] = {0}, o0[16 * 18], o1[16 * 18];
  for (int r = 0; r < 16; ++r) for (int c = 0; c < 16; ++c) h[r * 18 + c] = (r == c ? 20.0 : 0.0) + 1.0 / (1 + r + c) + 0.3 * ((r * 7 + c * 7) % 5);
  for (int r = 0; r < 16; ++r) for (int c = 0; c < r; ++c) h[r * 18 + c] = h[c * 18 + r];
  double *M, *out; long long *c, hc[2];
  cudaMalloc(&M, sizeof(h)); cudaMalloc(&out, sizeof(h)); cudaMalloc(&c, 16);
  cudaMemcpy(M, h, sizeof(h), cudaMemcpyHostToDevice);
  const int reps = 100;
  for (int w = 0; w < 2; ++w) {
    bench<0><<<1, 32>>>(M, out, c, reps); cudaDeviceSynchronize(); cudaMemcpy(hc, c, 16, cudaMemcpyDeviceToHost); cudaMemcpy(o0, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("inverse A (selects)     launch %d: warm %7.1f cycles/pivot, first call in the launch %7.1f\n", w, double(hc[0]) / (reps - 1) / 16, double(hc[1]) / 16);
    bench<1><<<1, 32>>>(M, out, c, reps); cudaDeviceSynchronize(); cudaMemcpy(hc, c, 16, cudaMemcpyDeviceToHost); cudaMemcpy(o1, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("inverse B (no selects)  launch %d: warm %7.1f cycles/pivot, first call in the launch %7.1f\n", w, double(hc[0]) / (reps - 1) / 16, double(hc[1]) / 16);
  }
  // check: M * inv = I
  double err0 = 0, err1 = 0;
  for (int r = 0; r < 16; ++r) for (int c = 0; c < 16; ++c) {
    double s0 = 0, s1 = 0;
    for (int t = 0; t < 16; ++t) { s0 += h[r * 18 + t] * o0[t * 18 + c]; s1 += h[r * 18 + t] * o1[t * 18 + c]; }
    err0 = fmax(err0, fabs(s0 - (r == c))); err1 = fmax(err1, fabs(s1 - (r == c)));
  }
  printf("max |M inv - I|: A %.2e  B %.2e\n", err0, err1);
  return 0;
}
