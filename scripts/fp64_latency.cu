// Dependent-issue latency of the fp64 instructions the serial phases of the selection kernel are made of
// (DFMA, DADD, 64-bit shuffle + DADD round, LDS.64 + DFMA), one warp, clock64 around a chain of N.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/exp/fp64_latency scripts/fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int N = 4096;
__global__ void k(double* out, long long* cyc, double a, double b) {
  __shared__ double sm[64];
  sm[threadIdx.x & 63] = a;
  __syncthreads();
  double x = a + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fma(x, b, a);
  long long t1 = clock64();
  double y = x;
#pragma unroll 16
  for (int i = 0; i < N; ++i) y = __dadd_rn(y, b);
  long long t2 = clock64();
  double z = y;
#pragma unroll 16
  for (int i = 0; i < N; ++i) z += __shfl_xor_sync(0xffffffffu, z, 1 + (i & 15));
  long long t3 = clock64();
  double w = z;
#pragma unroll 16
  for (int i = 0; i < N; ++i) w = fma(sm[(i + threadIdx.x) & 63], b, w);
  long long t4 = clock64();
  // four independent DFMA chains: issue-bound
  double c0 = w, c1 = w + 1, c2 = w + 2, c3 = w + 3;
#pragma unroll 16
  for (int i = 0; i < N; ++i) { c0 = fma(c0, b, a); c1 = fma(c1, b, a); c2 = fma(c2, b, a); c3 = fma(c3, b, a); }
  long long t5 = clock64();
  out[threadIdx.x + blockIdx.x * blockDim.x] = c0 + c1 + c2 + c3;
  if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; }
}
int main() {
  double* out; long long* cyc; long long h[5];
  cudaMalloc(&out, 8 * 1024 * 148); cudaMalloc(&cyc, 40);
  for (int warps = 1; warps <= 8; warps *= 2) {
    k<<<1, 32 * warps>>>(out, cyc, 1.0000001, 0.9999999);
    cudaMemcpy(h, cyc, 40, cudaMemcpyDeviceToHost);
    printf("{\"warps_per_sm\": %d, \"dfma_dep_cycles\": %.2f, \"dadd_dep_cycles\": %.2f, \"shfl64_dadd_round_cycles\": %.2f, "
           "\"lds_dfma_chain_cycles\": %.2f, \"dfma_4_chains_cycles_per_dfma\": %.2f}\n", warps, (double)h[0] / N, (double)h[1] / N,
           (double)h[2] / N, (double)h[3] / N, (double)h[4] / (4.0 * N));
  }
  return 0;
}
