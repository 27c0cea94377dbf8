// TF32 mode of the grid backend's update GEMM: tcgen05.mma kind::tf32 with TMEM accumulators.
//
// Same contraction and the same stream-K split as k_grid_gemm (gpis_grid.cuh), 32 model rows per
// chunk.  Accuracy comes from the 3xTF32 split (hi = tf32(x), lo = tf32(x - hi); hi*hi + hi*lo +
// lo*hi accumulated in fp32), i.e. ~2^-21 relative per product instead of TF32's 2^-11; the
// factor, the posterior state and the epilogue stay fp64.
//
// Warp roles (448 threads, 1 CTA/SM):
//   warp 0        producer: bulk async copies (UBLKCP) of fp32 axis-table rows + the per-chunk
//                 coefficients c[rho] = W[r, rho] * Tz[rho][jz]
//   warp 1        MMA issuer (one thread) and TMEM allocation (2 x 128 accumulator columns)
//   warps 2-5     A transform: row jy of the tile <- c[rho] * Ty[rho][jy], hi/lo, written K-major
//                 into SWIZZLE_128B shared-memory tiles (the layout tcgen05 reads)
//   warps 6-9     B transform: row jx <- Tx[rho][jx], hi/lo
//   warps 10-13   epilogue: tcgen05.ld of a finished accumulator -> fp64 partials in the split-K
//                 workspace, in the fragment order k_grid_epilogue expects
#pragma once
#include "gpis_grid.cuh"

namespace gpis {

constexpr int TKC = 32;                       // model rows per chunk (= 128 bytes of tf32: one swizzle row)
constexpr int T_RAW = 3, T_OPS = 2, T_ACC = 2;
constexpr int TF32_THREADS = 448;
constexpr int T_RAW_BYTES = 2 * TKC * GT * 4 + TKC * 4;          // Ty, Tx chunk (fp32) + coefficients
constexpr int T_OP_BYTES = 4 * GT * TKC * 4;                     // A hi, A lo, B hi, B lo
constexpr size_t TF32_SMEM = 1024 + (size_t)T_OPS * T_OP_BYTES + (size_t)T_RAW * T_RAW_BYTES + 256;

__device__ __forceinline__ float to_tf32(float x) {
  unsigned r;
  asm volatile("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
// x rounded to nearest at 11 significant bits (a tf32 value): Veltkamp split with 2^13 + 1
__device__ __forceinline__ float split_hi(float x) {
  const float c = __fmul_rn(x, 8193.0f);
  return __fsub_rn(c, __fsub_rn(c, x));
}
// UMMA shared-memory descriptor: K-major, SWIZZLE_128B, 8-row atoms of 1024 bytes
__device__ __forceinline__ uint64_t umma_desc_sw128(unsigned saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void umma_tf32(unsigned tmem_d, uint64_t da, uint64_t db, unsigned idesc, unsigned acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
               :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

// row `row` of a K-major SW128 tile, 16-byte chunk j (4 tf32) -> byte offset
__device__ __forceinline__ unsigned sw128_chunk(int row, int j) {
  return (unsigned)((row >> 3) * 1024 + (row & 7) * 128 + ((j ^ (row & 7)) << 4));
}

template <int NB>
__global__ void __launch_bounds__(TF32_THREADS, 1) k_grid_gemm_tf32(EngineParams P) {
  extern __shared__ unsigned char t_smem_raw[];
  const DevState* st = P.st;
  if (st->done) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r0 = st->r0;
  const int nx = P.glen[0], ny = P.glen[1];
  const int tiles_x = (nx + GT - 1) / GT, tiles_y = (ny + GT - 1) / GT;
  const int T = tiles_x * tiles_y * P.glen[2];
  const int C = (r0 + NB + TKC - 1) / TKC;
  const long long U = (long long)T * C;
  const int G = (int)min((long long)gridDim.x, U), g = blockIdx.x;
  if (g >= G) return;
  const long long u0 = (long long)g * U / G, u1 = (long long)(g + 1) * U / G;

  // carve shared memory: operand tiles need 1024-byte alignment
  unsigned char* base = t_smem_raw + ((1024 - (smem_u32(t_smem_raw) & 1023)) & 1023);
  unsigned char* ops = base;                                     // [T_OPS][Ahi|Alo|Bhi|Blo]
  unsigned char* raws = ops + (size_t)T_OPS * T_OP_BYTES;        // [T_RAW][Ty|Tx|cs]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(raws + (size_t)T_RAW * T_RAW_BYTES);
  unsigned long long* raw_full = bars;                 // [T_RAW]
  unsigned long long* raw_empty = raw_full + T_RAW;    // [T_RAW]
  unsigned long long* op_full = raw_empty + T_RAW;     // [T_OPS]
  unsigned long long* op_empty = op_full + T_OPS;      // [T_OPS]
  unsigned long long* acc_full = op_empty + T_OPS;     // [T_ACC]
  unsigned long long* acc_empty = acc_full + T_ACC;    // [T_ACC]
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(acc_empty + T_ACC);

  if (tid == 0) {
    for (int i = 0; i < T_RAW; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_empty[i], 8); }
    for (int i = 0; i < T_OPS; ++i) { mbar_init(&op_full[i], 8); mbar_init(&op_empty[i], 1); }
    for (int i = 0; i < T_ACC; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                 :: "r"(smem_u32(tmem_slot)), "r"(T_ACC * GT));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------ producer
    int rs = 0;
    unsigned rph = 0;
    for (int a = 0; a < NB; ++a) {
      const double* wrow = P.W + (long long)(r0 + a) * P.ldW;
      const int klen = r0 + a + 1;
      for (long long u = u0; u < u1; ++u) {
        const int tile = (int)(u / C), c = (int)(u % C);
        const int tz = tile / (tiles_x * tiles_y), ty = (tile / tiles_x) % tiles_y, tx = tile % tiles_x;
        mbar_wait(&raw_empty[rs], rph ^ 1);
        float* rTy = reinterpret_cast<float*>(raws + (size_t)rs * T_RAW_BYTES);
        float* rTx = rTy + TKC * GT;
        float* cs = rTx + TKC * GT;
        const long long rho = (long long)c * TKC + lane;
        double cv = rho < klen ? wrow[rho] : 0.0;
        cv *= P.T[2][rho * P.ldt[2] + P.gz0 + tz];
        cs[lane] = (float)cv;
        __syncwarp();
        if (lane == 0) mbar_arrive_expect_tx(&raw_full[rs], 2 * TKC * GT * 4);
        __syncwarp();
        // a 128-wide axis makes the 32 table rows of a chunk one contiguous 16 KB block: one bulk
        // copy instead of 32 (the per-copy cost of the TMA engine is what bounds a short chunk)
        const long long rho_c = (long long)c * TKC;
        if (P.ldt[1] == GT) { if (lane == 0) bulk_g2s(rTy, P.Tf[1] + rho_c * GT, TKC * GT * 4, &raw_full[rs]); }
        else bulk_g2s(rTy + lane * GT, P.Tf[1] + rho * P.ldt[1] + ty * GT, GT * 4, &raw_full[rs]);
        if (P.ldt[0] == GT) { if (lane == 1) bulk_g2s(rTx, P.Tf[0] + rho_c * GT, TKC * GT * 4, &raw_full[rs]); }
        else bulk_g2s(rTx + lane * GT, P.Tf[0] + rho * P.ldt[0] + tx * GT, GT * 4, &raw_full[rs]);
        if (++rs == T_RAW) { rs = 0; rph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(GT >> 3) << 17) | ((unsigned)(GT >> 4) << 24);
      int os = 0, as = 0;
      unsigned oph = 0, aph = 0;
      for (int a = 0; a < NB; ++a) {
        long long u = u0;
        while (u < u1) {
          const int tile = (int)(u / C);
          const long long uend = min(u1, (long long)(tile + 1) * C);
          mbar_wait(&acc_empty[as], aph ^ 1);              // epilogue has drained this accumulator
          tc_fence_after();
          const unsigned d_tmem = tmem + (unsigned)(as * GT);
          bool first = true;
          for (; u < uend; ++u) {
            mbar_wait(&op_full[os], oph);
            tc_fence_after();
            const unsigned ob = smem_u32(ops + (size_t)os * T_OP_BYTES);
#pragma unroll
            for (int ks = 0; ks < TKC / 8; ++ks) {
              const unsigned adv = ks * 32;
              const uint64_t dah = umma_desc_sw128(ob + adv), dal = umma_desc_sw128(ob + GT * TKC * 4 + adv);
              const uint64_t dbh = umma_desc_sw128(ob + 2 * GT * TKC * 4 + adv), dbl = umma_desc_sw128(ob + 3 * GT * TKC * 4 + adv);
              umma_tf32(d_tmem, dal, dbh, idesc, first ? 0u : 1u);
              umma_tf32(d_tmem, dah, dbl, idesc, 1u);
              umma_tf32(d_tmem, dah, dbh, idesc, 1u);
              first = false;
            }
            umma_commit(&op_empty[os]);                    // operand stage free once these MMAs retire
            if (++os == T_OPS) { os = 0; oph ^= 1; }
          }
          umma_commit(&acc_full[as]);                      // accumulator complete
          if (++as == T_ACC) { as = 0; aph ^= 1; }
        }
      }
    }
  } else if (warp < 10) {
    // ------------------------------------------------ operand transform (A: warps 2-5, B: warps 6-9)
    const bool is_a = warp < 6;
    const int row = (is_a ? tid - 64 : tid - 192);         // 0..127: jy (A) or jx (B) inside the tile
    int rs = 0, os = 0;
    unsigned rph = 0, oph = 0;
    for (int a = 0; a < NB; ++a) {
      for (long long u = u0; u < u1; ++u) {
        mbar_wait(&raw_full[rs], rph);
        mbar_wait(&op_empty[os], oph ^ 1);
        const float* raw = reinterpret_cast<const float*>(raws + (size_t)rs * T_RAW_BYTES) + (is_a ? 0 : TKC * GT);
        const float* cs = reinterpret_cast<const float*>(raws + (size_t)rs * T_RAW_BYTES) + 2 * TKC * GT;
        unsigned char* hi = ops + (size_t)os * T_OP_BYTES + (is_a ? 0 : 2 * GT * TKC * 4);
        unsigned char* lo = hi + GT * TKC * 4;
#pragma unroll
        for (int j = 0; j < TKC / 4; ++j) {
          float4 h, l;
          float v0 = raw[(4 * j + 0) * GT + row], v1 = raw[(4 * j + 1) * GT + row];
          float v2 = raw[(4 * j + 2) * GT + row], v3 = raw[(4 * j + 3) * GT + row];
          if (is_a) { v0 *= cs[4 * j]; v1 *= cs[4 * j + 1]; v2 *= cs[4 * j + 2]; v3 *= cs[4 * j + 3]; }
          // round-to-nearest hi (Veltkamp split on the FMA pipe, 3 full-rate ops instead of the
          // conversion unit's cvt.rna); lo = x - hi is exact in fp32 and the tensor core reads its
          // top 11 bits.  A truncating hi (AND mask) loses a bit and changed the pick sequence.
          h.x = split_hi(v0); h.y = split_hi(v1); h.z = split_hi(v2); h.w = split_hi(v3);
          l.x = v0 - h.x; l.y = v1 - h.y; l.z = v2 - h.z; l.w = v3 - h.w;
          const unsigned off = sw128_chunk(row, j);
          *reinterpret_cast<float4*>(hi + off) = h;
          *reinterpret_cast<float4*>(lo + off) = l;
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncwarp();
        if (lane == 0) { mbar_arrive(&op_full[os]); mbar_arrive(&raw_empty[rs]); }
        if (++rs == T_RAW) { rs = 0; rph ^= 1; }
        if (++os == T_OPS) { os = 0; oph ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------ epilogue (warps 10-13)
    const int quarter = warp & 3;                          // TMEM lanes this warp may read
    const int my = quarter * 32 + lane;                    // row (jy) inside the tile
    int as = 0;
    unsigned aph = 0;
    for (int a = 0; a < NB; ++a) {
      long long u = u0;
      while (u < u1) {
        const int tile = (int)(u / C);
        u = min(u1, (long long)(tile + 1) * C);
        const int seg = g - unit_owner((long long)tile * C, U, G);
        double2* ws = reinterpret_cast<double2*>(P.gws + (((long long)a * T + tile) * P.gws_segmax + seg) * GTILE_ELEMS);
        mbar_wait(&acc_full[as], aph);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < GT; c0 += 32) {
          unsigned r[32];
          const unsigned taddr = tmem + ((unsigned)(quarter * 32) << 16) + (unsigned)(as * GT + c0);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
              "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
              : "r"(taddr));
          asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
          // (my, mx) -> fragment slot of the DMMA layout the epilogue kernel reads
#pragma unroll
          for (int t = 0; t < 16; ++t) {
            const int mx = c0 + 2 * t;
            const int gwarp = (my >> 5) * 2 + (mx >> 6);
            const int frag = (gwarp * 32 + ((my & 31) >> 3) * 8 + ((mx & 63) >> 3)) * 32 + (my & 7) * 4 + ((mx & 7) >> 1);
            ws[frag] = make_double2((double)__uint_as_float(r[2 * t]), (double)__uint_as_float(r[2 * t + 1]));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[as]);
        if (++as == T_ACC) { as = 0; aph ^= 1; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "r"(T_ACC * GT));
  }
}

}  // namespace gpis
