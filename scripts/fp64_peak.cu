// Measures the fp64 FMA-pipe and fp64 mma.sync (DMMA) peak of the device, for the roofline
// denominator of the fp64 GEMM-shaped kernels (MEASURED_PEAKS.json has no fp64 figure).
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters) {
  double a[16];
  const double x = 1.0000001 + threadIdx.x * 1e-9, y = 1e-9;
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = i * 0.5;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], x, y);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE>
__global__ void __launch_bounds__(256) k_dmma(double* out, int iters) {
  double c[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = 0.0;
  double a0 = 1.0 + threadIdx.x * 1e-6, a1 = 0.5, a2 = 0.25, a3 = 0.125, b0 = 1e-3, b1 = 2e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (SHAPE == 0) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                     : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a0), "d"(b0));
      } else if (SHAPE == 1) {
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                     : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a0), "d"(a1), "d"(b0));
      } else if (SHAPE == 2) {
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                     : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                     : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
double time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 20000;
  double* out; cudaMalloc(&out, sizeof(double) * blocks * threads);
  double ms = time_ms([&] { k_dfma<<<blocks, threads>>>(out, iters); });
  double fl = 2.0 * 16 * iters * (double)blocks * threads;
  printf("{\"dfma_tflops\": %.2f", fl / ms / 1e9);
  ms = time_ms([&] { k_dmma<0><<<blocks, threads>>>(out, iters); });
  fl = 2.0 * 8 * 8 * 4 * 8 * iters * (double)blocks * threads / 32;
  printf(", \"dmma_m8n8k4_tflops\": %.2f", fl / ms / 1e9);
  ms = time_ms([&] { k_dmma<1><<<blocks, threads>>>(out, iters); });
  fl = 2.0 * 16 * 8 * 4 * 8 * iters * (double)blocks * threads / 32;
  printf(", \"dmma_m16n8k4_tflops\": %.2f", fl / ms / 1e9);
  ms = time_ms([&] { k_dmma<2><<<blocks, threads>>>(out, iters); });
  fl = 2.0 * 16 * 8 * 8 * 8 * iters * (double)blocks * threads / 32;
  printf(", \"dmma_m16n8k8_tflops\": %.2f", fl / ms / 1e9);
  printf(", \"sms\": %d, \"cuda_err\": \"%s\"}\n", p.multiProcessorCount, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
