// TF32 mode of the grid backend's update GEMM: tcgen05.mma kind::tf32 with TMEM accumulators.
//
// Same contraction and the same stream-K split as k_grid_gemm (gpis_grid.cuh), 32 model rows per
// chunk.  Accuracy comes from the 3xTF32 split (hi = tf32(x), lo = tf32(x - hi); hi*hi + hi*lo +
// lo*hi accumulated in fp32), i.e. ~2^-21 relative per product instead of TF32's 2^-11; the
// factor, the posterior state and the epilogue stay fp64.
//
// A CTA tile is a PAIR of z-slices (TZ = 2) of one (jy, jx) tile: the Ty chunk and the B tile are
// the same for both, only the coefficients differ, so the operand bytes fetched from L2 per MMA are
// halved -- at one slice per tile the kernel was bound by the L2 latency x bandwidth product (a
// 128x128x32 unit needs 48 KB for 0.39 us of MMA time; the rings that fit in shared memory cannot
// cover ~2 us of L2 latency at that rate).
//
// Warp roles (480 threads, 1 CTA/SM):
//   warp 0        raw producer: bulk async copies (UBLKCP) of the fp32 Ty rows of a chunk + the
//                 per-chunk coefficients c[rho] = W[r, rho] * Tz[rho][jz]
//   warp 1        MMA issuer (one thread) and TMEM allocation (all 512 columns: 2 stages x TZ slices)
//   warps 2-9     A transform: row jy of the tile <- c[rho] * Ty[rho][jy], hi/lo, written K-major
//                 into SWIZZLE_128B shared-memory tiles (the layout tcgen05 reads); 16 values/thread
//   warps 10-13   epilogue: tcgen05.ld of a finished accumulator -> fp64 partials in the split-K
//                 workspace, in the fragment order k_grid_epilogue expects
//   warp 14       B producer: the hi/lo Tx tile of the chunk does not depend on the step or the
//                 z-slice; it is kept in global memory already in the MMA's shared-memory image
//                 (grid_write_tables) and bulk-copied into its own 3-deep ring
#pragma once
#include "gpis_grid.cuh"

namespace gpis {

constexpr int TKC = 32;                       // model rows per chunk (= 128 bytes of tf32: one swizzle row)
constexpr int TZ = 2;                         // z-slices per CTA tile: they share the Ty chunk and the B tile
constexpr int T_RAW = 2, T_AOP = 2, T_BOP = 2, T_ACC = 2;
constexpr int TF32_THREADS = 480;
constexpr int T_TILE_BYTES = GT * TKC * 4;                       // one swizzled operand tile (16 KB)
constexpr int T_RAW_BYTES = TKC * GT * 4 + TZ * TKC * 4;         // Ty chunk (fp32) + coefficients of each slice
constexpr int T_A_BYTES = TZ * 2 * T_TILE_BYTES;                 // A hi/lo of each slice
constexpr size_t TF32_SMEM = 1024 + (size_t)T_AOP * T_A_BYTES + (size_t)T_BOP * 2 * T_TILE_BYTES +
                             (size_t)T_RAW * T_RAW_BYTES + 256;

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
// The 12 MMAs of one 32-row chunk (4 K-steps x {lo*hi, hi*lo, hi*hi}) in ONE asm block: the
// descriptors differ only in the 14-bit start-address field, so each is one 32-bit add and one
// mov.b64.  The MMAs are issued by a single thread; building each 64-bit descriptor in C++ cost
// ~30 instructions per MMA and made that thread, not the tensor pipe, the bottleneck.
//   a_hi/a_lo/b_hi/b_lo: low descriptor words of the four operand tiles (start address >> 4 | LBO)
__device__ __forceinline__ void umma_tf32_chunk(unsigned tmem_d, unsigned a_hi, unsigned a_lo, unsigned b_hi,
                                                unsigned b_lo, unsigned desc_hi, unsigned idesc, unsigned acc_first) {
  asm volatile(
      "{\n\t"
      ".reg .pred p0, p1;\n\t"
      ".reg .b32 ah, al, bh, bl;\n\t"
      ".reg .b64 dah, dal, dbh, dbl;\n\t"
      "setp.ne.b32 p0, %7, 0;\n\t"
      "setp.eq.b32 p1, 0, 0;\n\t"
      "mov.b64 dah, {%1, %5}; mov.b64 dal, {%2, %5}; mov.b64 dbh, {%3, %5}; mov.b64 dbl, {%4, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], dal, dbh, %6, p0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], dah, dbl, %6, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], dah, dbh, %6, p1;\n\t"
      "add.u32 ah, %1, 2; add.u32 al, %2, 2; add.u32 bh, %3, 2; add.u32 bl, %4, 2;\n\t"
      "mov.b64 dah, {ah, %5}; mov.b64 dal, {al, %5}; mov.b64 dbh, {bh, %5}; mov.b64 dbl, {bl, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], dal, dbh, %6, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], dah, dbl, %6, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], dah, dbh, %6, p1;\n\t"
      "add.u32 ah, %1, 4; add.u32 al, %2, 4; add.u32 bh, %3, 4; add.u32 bl, %4, 4;\n\t"
      "mov.b64 dah, {ah, %5}; mov.b64 dal, {al, %5}; mov.b64 dbh, {bh, %5}; mov.b64 dbl, {bl, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], dal, dbh, %6, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], dah, dbl, %6, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], dah, dbh, %6, p1;\n\t"
      "add.u32 ah, %1, 6; add.u32 al, %2, 6; add.u32 bh, %3, 6; add.u32 bl, %4, 6;\n\t"
      "mov.b64 dah, {ah, %5}; mov.b64 dal, {al, %5}; mov.b64 dbh, {bh, %5}; mov.b64 dbl, {bl, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], dal, dbh, %6, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], dah, dbl, %6, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], dah, dbh, %6, p1;\n\t"
      "}\n"
      :: "r"(tmem_d), "r"(a_hi), "r"(a_lo), "r"(b_hi), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(acc_first)
      : "memory");
}
__device__ __forceinline__ unsigned umma_desc_lo(unsigned saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
constexpr unsigned UMMA_DESC_HI_SW128 = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO, version, SWIZZLE_128B
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
  const int nx = P.glen[0], ny = P.glen[1], nz = P.glen[2];
  const int tiles_x = (nx + GT - 1) / GT, tiles_y = (ny + GT - 1) / GT, txy = tiles_x * tiles_y;
  const int T = txy * nz;                                  // real 128x128 tiles (workspace index)
  const int zg = P.gemm_zgroup;                            // 1 or TZ slices per work tile (host's choice)
  const int Tg = txy * ((nz + zg - 1) / zg);               // z-grouped tiles (work units)
  const int C = (r0 + NB + TKC - 1) / TKC;
  const long long U = (long long)Tg * C;
  const int G = (int)min((long long)gridDim.x, U), g = blockIdx.x;
  if (g >= G) return;
  const long long u0 = (long long)g * U / G, u1 = (long long)(g + 1) * U / G;

  // carve shared memory: operand tiles need 1024-byte alignment
  unsigned char* base = t_smem_raw + ((1024 - (smem_u32(t_smem_raw) & 1023)) & 1023);
  unsigned char* aops = base;                                          // [T_AOP][slice][Ahi|Alo]
  unsigned char* bops = aops + (size_t)T_AOP * T_A_BYTES;              // [T_BOP][Bhi|Blo]
  unsigned char* raws = bops + (size_t)T_BOP * 2 * T_TILE_BYTES;       // [T_RAW][Ty|cs0|cs1]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(raws + (size_t)T_RAW * T_RAW_BYTES);
  unsigned long long* raw_full = bars;                  // [T_RAW]
  unsigned long long* raw_empty = raw_full + T_RAW;     // [T_RAW]
  unsigned long long* a_full = raw_empty + T_RAW;       // [T_AOP]
  unsigned long long* a_empty = a_full + T_AOP;         // [T_AOP]
  unsigned long long* b_full = a_empty + T_AOP;         // [T_BOP]
  unsigned long long* b_empty = b_full + T_BOP;         // [T_BOP]
  unsigned long long* acc_full = b_empty + T_BOP;       // [T_ACC]
  unsigned long long* acc_empty = acc_full + T_ACC;     // [T_ACC]
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(acc_empty + T_ACC);

  if (tid == 0) {
    for (int i = 0; i < T_RAW; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_empty[i], 8); }
    for (int i = 0; i < T_AOP; ++i) { mbar_init(&a_full[i], 8); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < T_BOP; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < T_ACC; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                 :: "r"(smem_u32(tmem_slot)), "r"(T_ACC * TZ * GT));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------ raw producer
    int rs = 0;
    unsigned rph = 0;
    for (int a = 0; a < NB; ++a) {
      const double* wrow = P.W + (long long)(r0 + a) * P.ldW;
      const int klen = r0 + a + 1;
      for (long long u = u0; u < u1; ++u) {
        const int gt = (int)(u / C), c = (int)(u % C);
        const int tz0 = (gt / txy) * zg, ty = (gt / tiles_x) % tiles_y;
        mbar_wait(&raw_empty[rs], rph ^ 1);
        float* rTy = reinterpret_cast<float*>(raws + (size_t)rs * T_RAW_BYTES);
        float* cs = rTy + TKC * GT;
        const long long rho = (long long)c * TKC + lane;
        const double w = rho < klen ? wrow[rho] : 0.0;
#pragma unroll
        for (int z = 0; z < TZ; ++z)
          cs[z * TKC + lane] = (z < zg && tz0 + z < nz) ? (float)(w * P.T[2][rho * P.ldt[2] + P.gz0 + tz0 + z]) : 0.0f;
        __syncwarp();
        if (lane == 0) mbar_arrive_expect_tx(&raw_full[rs], TKC * GT * 4);
        __syncwarp();
        // a 128-wide axis makes the 32 table rows of a chunk one contiguous 16 KB block: one bulk
        // copy instead of 32 (the per-copy cost of the TMA engine is what bounds a short chunk)
        if (P.ldt[1] == GT) { if (lane == 0) bulk_g2s(rTy, P.Tf[1] + (long long)c * TKC * GT, TKC * GT * 4, &raw_full[rs]); }
        else bulk_g2s(rTy + lane * GT, P.Tf[1] + rho * P.ldt[1] + ty * GT, GT * 4, &raw_full[rs]);
        if (++rs == T_RAW) { rs = 0; rph ^= 1; }
      }
    }
  } else if (warp == 14) {
    // ------------------------------------------------ B producer (one thread)
    if (lane == 0) {
      int bs = 0;
      unsigned bph = 0;
      for (int a = 0; a < NB; ++a) {
        for (long long u = u0; u < u1; ++u) {
          const int gt = (int)(u / C), c = (int)(u % C);
          const int tx = gt % tiles_x;
          mbar_wait(&b_empty[bs], bph ^ 1);
          mbar_arrive_expect_tx(&b_full[bs], 2 * T_TILE_BYTES);
          bulk_g2s(bops + (size_t)bs * 2 * T_TILE_BYTES,
                   P.bimg + (((long long)tx * P.bimg_cmax + c) * 2) * (GT * TKC), 2 * T_TILE_BYTES, &b_full[bs]);
          if (++bs == T_BOP) { bs = 0; bph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(GT >> 3) << 17) | ((unsigned)(GT >> 4) << 24);
      int as_ = 0, bs = 0, cs_ = 0;
      unsigned aph = 0, bph = 0, cph = 0;
      for (int a = 0; a < NB; ++a) {
        long long u = u0;
        while (u < u1) {
          const int gt = (int)(u / C);
          const int nsl = min(zg, nz - (gt / txy) * zg);    // valid slices of this group
          const long long uend = min(u1, (long long)(gt + 1) * C);
          mbar_wait(&acc_empty[cs_], cph ^ 1);             // epilogue has drained this accumulator stage
          tc_fence_after();
          bool first = true;
          for (; u < uend; ++u) {
            mbar_wait(&a_full[as_], aph);
            mbar_wait(&b_full[bs], bph);
            tc_fence_after();
            const unsigned oa = smem_u32(aops + (size_t)as_ * T_A_BYTES);
            const unsigned ob = smem_u32(bops + (size_t)bs * 2 * T_TILE_BYTES);
            const unsigned b_hi = umma_desc_lo(ob), b_lo = umma_desc_lo(ob + T_TILE_BYTES);
            for (int z = 0; z < nsl; ++z) {
              const unsigned oz = oa + z * 2 * T_TILE_BYTES;
              umma_tf32_chunk(tmem + (unsigned)((cs_ * TZ + z) * GT), umma_desc_lo(oz), umma_desc_lo(oz + T_TILE_BYTES),
                              b_hi, b_lo, UMMA_DESC_HI_SW128, idesc, first ? 0u : 1u);
            }
            first = false;
            umma_commit(&a_empty[as_]);                    // stages free once these MMAs retire
            umma_commit(&b_empty[bs]);
            if (++as_ == T_AOP) { as_ = 0; aph ^= 1; }
            if (++bs == T_BOP) { bs = 0; bph ^= 1; }
          }
          umma_commit(&acc_full[cs_]);                     // accumulators of the group complete
          if (++cs_ == T_ACC) { cs_ = 0; cph ^= 1; }
        }
      }
    }
  } else if (warp < 10) {
    // ------------------------------------------------ A transform (warps 2-9): 128 rows x 2 halves of the chunk
    const int t = tid - 64;
    const int row = t & (GT - 1), half = t >> 7;           // jy inside the tile; chunks 4*half .. 4*half + 3
    int rs = 0, as_ = 0;
    unsigned rph = 0, aph = 0;
    for (int a = 0; a < NB; ++a) {
      for (long long u = u0; u < u1; ++u) {
        mbar_wait(&raw_full[rs], rph);
        mbar_wait(&a_empty[as_], aph ^ 1);
        const float* raw = reinterpret_cast<const float*>(raws + (size_t)rs * T_RAW_BYTES);
        const float* cs = raw + TKC * GT;
        unsigned char* abase = aops + (size_t)as_ * T_A_BYTES;
        // all shared-memory loads of the thread's 16 table values and 2 x 16 coefficients first, so
        // the FMA pipe is not stalled behind one LDS per multiply
        float tv[TKC / 2];
        float4 cz[TZ][TKC / 8];
#pragma unroll
        for (int i = 0; i < TKC / 2; ++i) tv[i] = raw[(half * (TKC / 2) + i) * GT + row];
#pragma unroll
        for (int z = 0; z < TZ; ++z)
#pragma unroll
          for (int jj = 0; jj < TKC / 8; ++jj)
            cz[z][jj] = *reinterpret_cast<const float4*>(cs + z * TKC + half * (TKC / 2) + 4 * jj);
#pragma unroll
        for (int jj = 0; jj < TKC / 8; ++jj) {
          const int j = half * (TKC / 8) + jj;
          const unsigned off = sw128_chunk(row, j);
#pragma unroll
          for (int z = 0; z < TZ; ++z) {
            const float v0 = tv[4 * jj] * cz[z][jj].x, v1 = tv[4 * jj + 1] * cz[z][jj].y;
            const float v2 = tv[4 * jj + 2] * cz[z][jj].z, v3 = tv[4 * jj + 3] * cz[z][jj].w;
            float4 h, l;
            // round-to-nearest hi (Veltkamp split on the FMA pipe, 3 full-rate ops instead of the
            // conversion unit's cvt.rna); lo = x - hi is exact in fp32 and the tensor core reads its
            // top 11 bits.  A truncating hi (AND mask) loses a bit and changed the pick sequence.
            h.x = split_hi(v0); h.y = split_hi(v1); h.z = split_hi(v2); h.w = split_hi(v3);
            l.x = v0 - h.x; l.y = v1 - h.y; l.z = v2 - h.z; l.w = v3 - h.w;
            *reinterpret_cast<float4*>(abase + z * 2 * T_TILE_BYTES + off) = h;
            *reinterpret_cast<float4*>(abase + z * 2 * T_TILE_BYTES + T_TILE_BYTES + off) = l;
          }
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncwarp();
        if (lane == 0) { mbar_arrive(&a_full[as_]); mbar_arrive(&raw_empty[rs]); }
        if (++rs == T_RAW) { rs = 0; rph ^= 1; }
        if (++as_ == T_AOP) { as_ = 0; aph ^= 1; }
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
        const int gt = (int)(u / C);
        u = min(u1, (long long)(gt + 1) * C);
        const int seg = g - unit_owner((long long)gt * C, U, G);
        const int tz0 = (gt / txy) * zg, tyx = gt % txy;
        const int nsl = min(zg, nz - tz0);
        mbar_wait(&acc_full[as], aph);
        tc_fence_after();
        for (int z = 0; z < nsl; ++z) {
          const int tile = (tz0 + z) * txy + tyx;
          double2* ws = reinterpret_cast<double2*>(P.gws + (((long long)a * T + tile) * P.gws_segmax + seg) * GTILE_ELEMS);
#pragma unroll 1
          for (int c0 = 0; c0 < GT; c0 += 32) {
            unsigned r[32];
            const unsigned taddr = tmem + ((unsigned)(quarter * 32) << 16) + (unsigned)((as * TZ + z) * GT + c0);
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "r"(T_ACC * TZ * GT));
  }
}

}  // namespace gpis
