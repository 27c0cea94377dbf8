// Stand-alone validation of the tcgen05 building blocks used by the TF32 mode of the grid GEMM:
// K-major SWIZZLE_128B operand tiles written by CUDA cores (3xTF32 hi/lo split), UMMA shared-memory
// and instruction descriptors, tcgen05.mma kind::tf32 with TMEM accumulators, tcgen05.ld epilogue.
// D[128 x 128] = A[128 x K] * B[128 x K]^T, compared with an fp64 host reference.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float to_tf32(float x) {
  unsigned r;
  asm volatile("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// K-major SWIZZLE_128B tile: row pitch 128 B (32 tf32), 8-row atoms of 1024 B
__device__ __forceinline__ unsigned sw128_off(int row, int k) {
  return (unsigned)((row >> 3) * 1024 + (row & 7) * 128 + ((((k * 4) >> 4) ^ (row & 7)) << 4) + ((k * 4) & 15));
}
__device__ __forceinline__ uint64_t make_desc(unsigned saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address
  d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset: 8-row atom pitch
  d |= (uint64_t)1 << 46;                          // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void mma_tf32(unsigned tmem_d, uint64_t da, uint64_t db, unsigned idesc, unsigned accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate));
}

__global__ void __launch_bounds__(128, 1) k_test(const float* A, const float* B, float* D, int K) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* Ahi = reinterpret_cast<float*>(smem);
  float* Alo = Ahi + 128 * 32;
  float* Bhi = Alo + 128 * 32;
  float* Blo = Bhi + 128 * 32;
  __shared__ unsigned s_tmem;
  __shared__ __align__(8) unsigned long long s_bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"(smem_u32(&s_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(&s_tmem)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const unsigned tmem = s_tmem;
  // instruction descriptor: D=f32, A=B=tf32, both K-major, N=128, M=128
  const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
  unsigned phase = 0;
  for (int k0 = 0; k0 < K; k0 += 32) {
    for (int k = 0; k < 32; ++k) {
      const float a = A[(size_t)tid * K + k0 + k], b = B[(size_t)tid * K + k0 + k];
      const float ah = to_tf32(a), bh = to_tf32(b);
      const unsigned off = sw128_off(tid, k) / 4;
      Ahi[off] = ah; Alo[off] = to_tf32(a - ah);
      Bhi[off] = bh; Blo[off] = to_tf32(b - bh);
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      for (int ks = 0; ks < 4; ++ks) {
        const unsigned adv = ks * 32;   // 8 tf32 = 32 bytes along K inside the 128-byte row
        const uint64_t dah = make_desc(smem_u32(Ahi) + adv), dal = make_desc(smem_u32(Alo) + adv);
        const uint64_t dbh = make_desc(smem_u32(Bhi) + adv), dbl = make_desc(smem_u32(Blo) + adv);
        mma_tf32(tmem, dal, dbh, idesc, (k0 | ks) ? 1u : 0u);
        mma_tf32(tmem, dah, dbl, idesc, 1u);
        mma_tf32(tmem, dah, dbh, idesc, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(smem_u32(&s_bar)) : "memory");
    }
    // everyone waits for the MMAs to finish reading the operand tiles
    asm volatile(
        "{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D1;\nbra W1;\nD1:\n}\n"
        :: "r"(smem_u32(&s_bar)), "r"(phase) : "memory");
    phase ^= 1;
  }
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  // epilogue: warp w owns TMEM lanes 32w..32w+31 (= rows of D)
  for (int c0 = 0; c0 < 128; c0 += 32) {
    unsigned r[32];
    const unsigned taddr = tmem + ((unsigned)(warp * 32) << 16) + c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    for (int j = 0; j < 32; ++j) D[(size_t)tid * 128 + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "r"(128));
}

int main() {
  const int K = 96;
  std::vector<float> A(128 * K), B(128 * K), D(128 * 128);
  srand(1);
  for (auto& x : A) x = (float)rand() / RAND_MAX * 2 - 1;
  for (auto& x : B) x = (float)rand() / RAND_MAX * 2 - 1;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, D.size() * 4);
  const size_t smem = 4 * 128 * 32 * 4 + 1024;
  cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_test<<<1, 128, smem>>>(dA, dB, dD, K);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 128; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * (double)B[n * K + k];
      maxerr = fmax(maxerr, fabs(s - D[m * 128 + n]));
      maxref = fmax(maxref, fabs(s));
    }
  printf("{\"cuda\": \"%s\", \"max_abs_err\": %.3e, \"max_ref\": %.3f, \"D00\": %.6f}\n", cudaGetErrorString(e), maxerr, maxref, D[0]);
  return 0;
}
