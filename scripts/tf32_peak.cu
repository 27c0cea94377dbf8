// Measures the dense tcgen05 kind::tf32 peak of the device (one CTA per SM, cta_group::1, M = 128, N = 128 or 256,
// K = 8 per instruction; operands resident in shared memory, accumulators in TMEM, no loads in the loop) -- the
// roofline denominator of the TF32 mode of the lattice update GEMM (MEASURED_PEAKS.json has bf16 only).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tf32_peak scripts/tf32_peak.cu && /tmp/tf32_peak
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(unsigned saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B, K-major
  return d;
}
__device__ __forceinline__ void mma_tf32(unsigned tmem_d, uint64_t da, uint64_t db, unsigned idesc, unsigned accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate));
}

template <int N>
__global__ void __launch_bounds__(128, 1) k_peak(int iters, float* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* A = reinterpret_cast<float*>(smem);              // 128 x 32 tf32, SW128
  float* B = A + 128 * 32;                                // N x 32
  __shared__ unsigned s_tmem;
  __shared__ __align__(8) unsigned long long s_bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 + N) * 32; i += 128) A[i] = 0.0f;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"(smem_u32(&s_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(&s_tmem)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const unsigned tmem = s_tmem;
  const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((128u >> 4) << 24);
  if (tid == 0) {
    unsigned phase = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 16; ++u) {                      // 16 x 4 MMAs, accumulators rotate over the TMEM columns
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          mma_tf32(tmem + (unsigned)((u * N) & 511), make_desc(smem_u32(A) + ks * 32), make_desc(smem_u32(B) + ks * 32), idesc, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(smem_u32(&s_bar)) : "memory");
      asm volatile(
          "{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D1;\nbra W1;\nD1:\n}\n"
          :: "r"(smem_u32(&s_bar)), "r"(phase) : "memory");
      phase ^= 1;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "r"(512));
  if (sink && tid == 0) sink[blockIdx.x] = A[0];
}

template <int N>
double run(int sms, int iters) {
  const size_t smem = (size_t)(128 + N) * 32 * 4 + 1024;
  cudaFuncSetAttribute(k_peak<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_peak<N><<<sms, 128, smem>>>(iters / 8, nullptr);
  cudaDeviceSynchronize();
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    k_peak<N><<<sms, 128, smem>>>(iters, nullptr);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tf = (double)sms * iters * 64.0 * 2.0 * 128.0 * N * 8.0 / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  return best;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  const double t128 = run<128>(sms, 20000), t256 = run<256>(sms, 10000);
  cudaError_t e = cudaDeviceSynchronize();
  printf("{\"device\": \"%s\", \"sms\": %d, \"cuda\": \"%s\", \"tcgen05_tf32_m128n128k8_tflops\": %.1f, "
         "\"tcgen05_tf32_m128n256k8_tflops\": %.1f, \"how\": \"one CTA per SM, cta_group::1, 64 back-to-back tcgen05.mma kind::tf32 per commit, "
         "operands in shared memory (SWIZZLE_128B K-major), accumulators rotating over 512 TMEM columns, best of 5, CUDA events\"}\n",
         p.name, sms, cudaGetErrorString(e), t128, t256);
  return 0;
}
