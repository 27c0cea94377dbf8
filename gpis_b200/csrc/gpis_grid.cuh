// Grid backend: candidates on a regular axis-aligned lattice (the reference's SelectFromGrid
// case, gpu_active_set_selector.cpp:170-201, and every BASELINE.json config).
//
// The SE kernel separates over axes, so the covariance of model row rho with lattice point
// (jx, jy, jz) is Tx[rho][jx] * Ty[rho][jy] * Tz[rho][jz] (gradient rows carry the factor
// (x_j - x_rho)/ell^2 in their own axis table).  With W = L^-1 maintained explicitly, the new
// row of V = L^-1 K(rows, candidates) is
//     v[jy, jx] = sum_rho (W[r, rho] Tz[rho][jz]) * Ty[rho][jy] * Tx[rho][jx]
// -- a dense [ny x r] x [r x nx] contraction per z-slice, run on the fp64 tensor pipe (DMMA).
// V is never stored: the per-step HBM traffic drops from 8*N*r bytes (panel backend) to the
// posterior arrays (~42 bytes per candidate), and the step becomes compute-bound.
#pragma once
#include "gpis_launch.cuh"
#include "gpis_common.cuh"
#include "gpis_select.cuh"

namespace gpis {

constexpr int GT = 128;         // tile edge (jy x jx) of the update GEMM
constexpr int GKC = 16;         // model rows per pipeline stage
constexpr int GLD = GT + 4;     // padded smem row (conflict-free DMMA fragment loads)


constexpr int GSOLVE_THREADS = 256;
constexpr int GSU = 16;         // independent 256-byte row segments in flight per warp in the W passes


#ifndef GPIS_EMULATE   // tests/emu supplies stand-ins for the inline-PTX helpers (test infrastructure only)
__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
#endif
// Internal layout of the per-candidate state (mu, var, ambiguity, flags) in the grid backend:
// tile-major, and inside a TE x TE tile (TE = P.tile_e: 128 for the per-step kernels, 64 for the persistent
// selection kernel on 1-D / 2-D lattices; 16 = its 16 x 16 x 16 cubes on 3-D lattices, below) the order of the update GEMM's accumulator fragments -- 8 warps as 4 (jy) x 2 (jx), each
// warp tile (TE/4) x (TE/2) in 8 x 8 DMMA blocks -- so that the epilogue streams it with fully coalesced
// 16-byte accesses.
__host__ __device__ inline int grid_tile_edge(const EngineParams& P) { return P.tile_e > 0 ? P.tile_e : GT; }
template <int TE>
__host__ __device__ inline long long grid_slot_te(const EngineParams& P, int jx, int jy, int tz) {
  constexpr int WY = TE / 4, WX = TE / 2, NI = WY / 8, NJ = WX / 8;
  const int tiles_x = (P.glen[0] + TE - 1) / TE, tiles_y = (P.glen[1] + TE - 1) / TE;
  const long long tile = ((long long)tz * tiles_y + jy / TE) * tiles_x + jx / TE;
  const int my = jy % TE, mx = jx % TE;
  const int gwarp = (my / WY) * 2 + mx / WX;
  const int frag = ((gwarp * NI + (my % WY) / 8) * NJ + (mx % WX) / 8) * 32 + (my % 8) * 4 + (mx % 8) / 2;
  return tile * (TE * TE) + frag * 2 + (mx & 1);
}
// tile_e == 16: the persistent kernel's 16 x 16 x 16 cubes of a 3-D lattice.  Warp w of the update GEMM holds the
// z-slices 2w, 2w + 1 of a cube as 4 x 2 DMMA blocks: fragment ((w * 4 + i) * 2 + j) * 32 + lane with
// i = (jz & 1) * 2 + (jy >> 3), j = jx >> 3 (gpis_grid_persist.cuh: persist_pair_index<true>)
constexpr int GCUBE = 16;
__host__ __device__ inline long long grid_slot_cube(const EngineParams& P, int jx, int jy, int jz) {
  const int tiles_x = (P.glen[0] + GCUBE - 1) / GCUBE, tiles_y = (P.glen[1] + GCUBE - 1) / GCUBE;
  const long long tile = ((long long)(jz / GCUBE) * tiles_y + jy / GCUBE) * tiles_x + jx / GCUBE;
  const int mz = jz % GCUBE, my = jy % GCUBE, mx = jx % GCUBE;
  const int frag = (((mz >> 1) * 4 + (mz & 1) * 2 + (my >> 3)) * 2 + (mx >> 3)) * 32 + (my & 7) * 4 + (mx & 7) / 2;
  return tile * (GCUBE * GCUBE * GCUBE) + frag * 2 + (mx & 1);
}
__host__ __device__ inline bool grid_unslot_cube(const EngineParams& P, long long slot, int& jx, int& jy, int& jz) {
  const int tiles_x = (P.glen[0] + GCUBE - 1) / GCUBE, tiles_y = (P.glen[1] + GCUBE - 1) / GCUBE;
  const long long tile = slot / (GCUBE * GCUBE * GCUBE);
  const int w = (int)(slot % (GCUBE * GCUBE * GCUBE));
  const int e = w & 1, frag = w >> 1;
  const int lane = frag & 31, j = (frag >> 5) & 1, i = (frag >> 6) & 3, gwarp = frag >> 8;
  jz = (int)(tile / (tiles_x * tiles_y)) * GCUBE + 2 * gwarp + (i >> 1);
  jy = (int)((tile / tiles_x) % tiles_y) * GCUBE + (i & 1) * 8 + (lane >> 2);
  jx = (int)(tile % tiles_x) * GCUBE + j * 8 + (lane & 3) * 2 + e;
  return jx < P.glen[0] && jy < P.glen[1] && jz < P.glen[2];
}
__host__ __device__ inline long long grid_slot(const EngineParams& P, int jx, int jy, int tz) {
  const int te = grid_tile_edge(P);
  return te == GCUBE ? grid_slot_cube(P, jx, jy, tz) : te == 64 ? grid_slot_te<64>(P, jx, jy, tz) : grid_slot_te<GT>(P, jx, jy, tz);
}
__host__ __device__ inline long long grid_slot_of_local(const EngineParams& P, long long jl) {
  const int nx = P.glen[0], ny = P.glen[1];
  return grid_slot(P, (int)(jl % nx), (int)((jl / nx) % ny), (int)(jl / ((long long)nx * ny)));
}
// inverse: slot -> (jx, jy, tz); returns false for padding slots
template <int TE>
__host__ __device__ inline bool grid_unslot_te(const EngineParams& P, long long slot, int& jx, int& jy, int& tz) {
  constexpr int WY = TE / 4, WX = TE / 2, NI = WY / 8, NJ = WX / 8;
  const int tiles_x = (P.glen[0] + TE - 1) / TE, tiles_y = (P.glen[1] + TE - 1) / TE;
  const long long tile = slot / (TE * TE);
  const int w = (int)(slot % (TE * TE));
  const int e = w & 1, frag = w >> 1;
  const int lane = frag & 31, t = frag >> 5;
  const int j = t % NJ, i = (t / NJ) % NI, gwarp = t / (NJ * NI);
  const int my = (gwarp >> 1) * WY + i * 8 + (lane >> 2);
  const int mx = (gwarp & 1) * WX + j * 8 + (lane & 3) * 2 + e;
  tz = (int)(tile / (tiles_x * tiles_y));
  jy = (int)((tile / tiles_x) % tiles_y) * TE + my;
  jx = (int)(tile % tiles_x) * TE + mx;
  return jx < P.glen[0] && jy < P.glen[1];
}
__host__ __device__ inline bool grid_unslot(const EngineParams& P, long long slot, int& jx, int& jy, int& tz) {
  const int te = grid_tile_edge(P);
  return te == GCUBE ? grid_unslot_cube(P, slot, jx, jy, tz)
                     : te == 64 ? grid_unslot_te<64>(P, slot, jx, jy, tz) : grid_unslot_te<GT>(P, slot, jx, jy, tz);
}
// slots of the whole (tile-padded) lattice in the layout of this selection
__host__ __device__ inline long long grid_slot_count(const EngineParams& P) {
  const int TE = grid_tile_edge(P);
  if (TE == GCUBE)
    return (long long)((P.glen[0] + TE - 1) / TE) * ((P.glen[1] + TE - 1) / TE) * ((P.glen[2] + TE - 1) / TE) * TE * TE * TE;
  return (long long)((P.glen[0] + TE - 1) / TE) * ((P.glen[1] + TE - 1) / TE) * P.glen[2] * TE * TE;
}

__global__ void k_grid_init_state(EngineParams P, long long nslots) {
  const long long slot = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (slot >= nslots) return;
  int jx, jy, tz;
  if (grid_unslot(P, slot, jx, jy, tz)) {
    const int jd[3] = {jx, jy, tz + P.gz0};
    double m = P.mean_c;
    for (int d = 0; d < MAXD; ++d) m += P.mean_w[d] * (P.gorg[d] + P.gh[d] * (double)jd[d]);
    P.mu[slot] = m;
    P.var[slot] = P.sf2;
    P.flags[slot] = 0;
  } else {                         // padding of partial tiles: never scored
    P.mu[slot] = 0.0;
    P.var[slot] = 0.0;
    P.flags[slot] = F_HIGH | F_LOW;
  }
  P.amb[slot] = 0.0;
}

// natural-order copies of the state for the getters
__global__ void k_grid_export(EngineParams P, double* mu, double* var, double* amb, unsigned char* flags) {
  const long long jl = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (jl >= P.N) return;
  const long long slot = grid_slot_of_local(P, jl);
  if (mu) mu[jl] = P.mu[slot];
  if (var) var[jl] = P.var[slot];
  if (amb) {      // straddle score of the current posterior (select_active_set_straddle.m:106-108)
    double beta = P.beta;
    const double it = (double)(P.st->iter > 0 ? P.st->iter : 1);
    if (P.beta_mode == 1) beta = 2.0 * log((double)P.N_total * it * 9.869604401089358 / (6.0 * P.beta_delta));
    else if (P.beta_mode == 2) beta = 2.0 * log((double)P.N_total * 9.869604401089358 * it * it / (6.0 * P.beta_delta));
    double ve = P.var[slot];
    if (P.variance_mode == 1) ve += P.noise ? P.noise[jl] : P.noise_scalar;
    const double w = sqrt(beta) * ve, m = P.mu[slot];
    amb[jl] = fmin(m + w - P.level, P.level - (m - w));
  }
  if (flags) flags[jl] = P.flags[slot];
}

#define GPIS_ISSUE_FENCE() asm volatile("" ::: "memory")
#ifndef GPIS_EMULATE
__device__ __forceinline__ void g_cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void g_cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
// Non-coherent 8-byte load the compiler may not sink below the fence that follows a batch of them:
// all GSU loads of a batch are issued before the first dependent FMA (one memory round trip per
// batch instead of one per ~6 loads).
__device__ __forceinline__ double ld_nc_f64(const double* p) {
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];\n" : "=d"(v) : "l"(p));
  return v;
}

template <int N>
__device__ __forceinline__ void g_cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// after mbarrier.init, before the barriers are used by the async proxy
__device__ __forceinline__ void mbar_init_fence() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
#endif

__device__ __forceinline__ int pick_winner(const EngineParams& P, const double* recs) {
  int win = -1;
  double bs = 0.0;
  long long bid = 0;
  for (int w = 0; w < P.world; ++w) {
    const double* pay = recs + (long long)w * P.pay_stride;
    const long long* pl = reinterpret_cast<const long long*>(pay);
    if (pl[PAY_LOCAL] < 0) continue;
    if (win < 0 || better(pay[PAY_SCORE], pl[PAY_ID], bs, bid)) { win = w; bs = pay[PAY_SCORE]; bid = pl[PAY_ID]; }
  }
  return win;
}

// Axis-table rows of the NB new model rows of the point at nx (whole block cooperates):
//   T_d[row][j] = exp(-(x_j - p_d)^2 / (2 ell^2)), times (x_j - p_d)/ell^2 on the row's own axis for
//   gradient rows ([d_a f(x_rho), f(x_j)]); sf2 is folded into axis 0.  TF32 mode also keeps fp32
//   copies and, for axis 0, the hi/lo split already in the K-major SWIZZLE_128B image the tcgen05
//   B operand is read from (one 2 x 16 KB image per (x-tile, 32-row chunk); it does not depend on
//   the step or the z-slice, so it is written once per row and bulk-copied by every CTA).
template <int NB, int D>
__device__ __forceinline__ void grid_write_tables(const EngineParams& P, int r0, const double* nx) {
  for (int d = 0; d < D; ++d) {
    const int len = P.gtlen[d];
    for (int e = threadIdx.x; e < len * NB; e += blockDim.x) {
      const int a = e / len, j = e % len;
      const double diff = (P.gorg[d] + P.gh[d] * (double)j) - nx[d];
      double v = exp(-0.5 * diff * diff * P.inv_ell2);
      if (a > 0 && a - 1 == d) v *= diff * P.inv_ell2;
      if (d == 0) v *= P.sf2;
      const long long row = r0 + a;
      P.T[d][row * P.ldt[d] + j] = v;
      if (P.Tf[d]) P.Tf[d][row * P.ldt[d] + j] = (float)v;
      if (d == 0 && P.bimg) {
        const float x = (float)v;
        const float c = __fmul_rn(x, 8193.0f);
        const float hi = __fsub_rn(c, __fsub_rn(c, x));           // Veltkamp: x rounded to a tf32
        const int k = (int)(row & 31), n = j & (GT - 1);
        float* img = P.bimg + (((long long)(j / GT) * P.bimg_cmax + (row >> 5)) * 2) * (GT * 32);
        const unsigned off = (unsigned)((n >> 3) * 1024 + (n & 7) * 128 + (((k >> 2) ^ (n & 7)) << 4) + ((k & 3) << 2)) >> 2;
        img[off] = hi;
        img[GT * 32 + off] = x - hi;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// K0 (one CTA): winner, its covariance with the model rows (kvec), its axis-table rows.
// ---------------------------------------------------------------------------------------
template <int NB, int D>
__global__ void __launch_bounds__(1024, 1) k_grid_point(EngineParams P, int finalize_only) {
  DevState* st = P.st;
  __shared__ int s_win;
  if (st->done) {
    if (finalize_only) signal_selection_done(P, st);
    return;
  }
  const int tid = threadIdx.x;
  const int gen = st->iter;
  if (!wait_records(P, st, gen)) return;
  const double* recs = records_ptr(P, gen);
  if (tid == 0) {
    int win = pick_winner(P, recs);
    if (win < 0) st->done = 1;
    else if (!finalize_only && st->m >= P.M_cap) { st->done = 1; win = -1; }
    s_win = win;
  }
  __syncthreads();
  const int win = s_win;
  if (win < 0) {
    if (finalize_only) signal_selection_done(P, st);
    return;
  }
  const double* pay = recs + (long long)win * P.pay_stride;
  const long long* pl = reinterpret_cast<const long long*>(pay);
  const int r0 = st->r, m = st->m;
  if (tid == 0) {
    if (win == P.rank) P.flags[grid_slot_of_local(P, pl[PAY_LOCAL])] |= F_ACTIVE;
    P.sel_ids[m] = pl[PAY_ID];
    st->n_sel = m + 1;
  }
  if (finalize_only) {
    signal_selection_done(P, st);
    return;
  }
  double nx[D];
#pragma unroll
  for (int d = 0; d < D; ++d) nx[d] = pay[PAY_X + d];
  // covariance of the new rows with the old rows -> kvec[a][rho]
  for (int i = tid; i < m; i += blockDim.x) {
    double z[D], c[NB][NB];
#pragma unroll
    for (int d = 0; d < D; ++d) z[d] = P.AX[d][i];
    cov_rows_rows<NB, D>(P, nx, z, c);
#pragma unroll
    for (int a = 0; a < NB; ++a)
#pragma unroll
      for (int b = 0; b < NB; ++b) P.kvec[(long long)a * P.R_cap + i * NB + b] = c[a][b];
  }
  grid_write_tables<NB, D>(P, r0, nx);
  if (tid == 0) {
#pragma unroll
    for (int d = 0; d < D; ++d) { st->nx[d] = nx[d]; P.AX[d][m] = nx[d]; }
    st->nnoise = pay[PAY_NOISE];
    st->r0 = r0;
    // residual targets of the new rows, consumed by k_grid_wrow
    double mf = P.mean_c;
#pragma unroll
    for (int d = 0; d < D; ++d) mf += P.mean_w[d] * nx[d];
    st->gnew[0] = pay[PAY_TSDF] - mf;
    for (int a = 1; a < NB; ++a) st->gnew[a] = pay[PAY_NRM + a - 1] - P.mean_w[a - 1];
  }
}

// ---------------------------------------------------------------------------------------
// K1: X = W[0:r0, 0:r0] kvec  (= L^-1 Q(old, new)); one warp per row, written into L's new rows.
// FUSED = true folds K0 into this launch: every CTA waits for the records, picks the winner and
// computes kvec into shared memory itself (r0 kernel evaluations, cheaper than a dependent
// launch); one extra CTA (the last) does K0's bookkeeping and table rows instead of rows of W.
// ---------------------------------------------------------------------------------------
template <int NB, int D, bool FUSED>
__global__ void __launch_bounds__(GSOLVE_THREADS) k_grid_solve(EngineParams P) {
  GPIS_DYN_SMEM_F64(s_kv);                           // FUSED: [NB][R_cap]
  DevState* st = P.st;
  if (st->done) return;
  const int tid = threadIdx.x, lane = tid & 31;
  int r0, m = 0;
  const double* kv = P.kvec;
  long long kvld = P.R_cap;
  int nctas = gridDim.x;
  if (FUSED) {
    __shared__ int s_win;
    const int gen = st->iter;
    if (!wait_records(P, st, gen)) return;
    const double* recs = records_ptr(P, gen);
    m = st->m;
    if (tid == 0) {
      int win = pick_winner(P, recs);
      if (win >= 0 && m >= P.M_cap) win = -1;
      s_win = win;
    }
    __syncthreads();
    const int win = s_win;
    const bool book = (blockIdx.x == gridDim.x - 1);
    if (win < 0) {
      if (book && tid == 0) st->done = 1;            // nothing unclassified / capacity
      return;
    }
    const double* pay = recs + (long long)win * P.pay_stride;
    const long long* pl = reinterpret_cast<const long long*>(pay);
    r0 = st->r;
    double nx[D];
#pragma unroll
    for (int d = 0; d < D; ++d) nx[d] = pay[PAY_X + d];
    if (book) {
      if (tid == 0) {
        if (win == P.rank) P.flags[grid_slot_of_local(P, pl[PAY_LOCAL])] |= F_ACTIVE;
        P.sel_ids[m] = pl[PAY_ID];
        st->n_sel = m + 1;
#pragma unroll
        for (int d = 0; d < D; ++d) { st->nx[d] = nx[d]; P.AX[d][m] = nx[d]; }
        st->nnoise = pay[PAY_NOISE];
        st->r0 = r0;
        double mf = P.mean_c;
#pragma unroll
        for (int d = 0; d < D; ++d) mf += P.mean_w[d] * nx[d];
        st->gnew[0] = pay[PAY_TSDF] - mf;
        for (int a = 1; a < NB; ++a) st->gnew[a] = pay[PAY_NRM + a - 1] - P.mean_w[a - 1];
      }
      grid_write_tables<NB, D>(P, r0, nx);
      return;
    }
    for (int i = tid; i < m; i += blockDim.x) {
      double z[D], c[NB][NB];
#pragma unroll
      for (int d = 0; d < D; ++d) z[d] = P.AX[d][i];
      cov_rows_rows<NB, D>(P, nx, z, c);
#pragma unroll
      for (int a = 0; a < NB; ++a)
#pragma unroll
        for (int b = 0; b < NB; ++b) s_kv[(long long)a * P.R_cap + i * NB + b] = c[a][b];
    }
    __syncthreads();
    kv = s_kv;
    nctas = gridDim.x - 1;
  } else {
    r0 = st->r0;
  }
  const int wid = (blockIdx.x * blockDim.x + tid) >> 5, nw = (nctas * blockDim.x) >> 5;
  for (int s = wid; s < r0; s += nw) {
    const double* wrow = P.W + (long long)s * P.ldW;
    double acc[NB];
#pragma unroll
    for (int a = 0; a < NB; ++a) acc[a] = 0.0;
    int c = lane;
    for (; c + 32 * (GSU - 1) <= s; c += 32 * GSU) {
      double w[GSU];
#pragma unroll
      for (int u = 0; u < GSU; ++u) w[u] = ld_nc_f64(wrow + c + 32 * u);
      GPIS_ISSUE_FENCE();
#pragma unroll
      for (int u = 0; u < GSU; ++u)
#pragma unroll
        for (int a = 0; a < NB; ++a) acc[a] += w[u] * kv[(long long)a * kvld + c + 32 * u];
    }
    for (; c <= s; c += 32) {
      const double w = __ldg(wrow + c);
#pragma unroll
      for (int a = 0; a < NB; ++a) acc[a] += w * kv[(long long)a * kvld + c];
    }
#pragma unroll
    for (int a = 0; a < NB; ++a) {
      double v = acc[a];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) P.L[(long long)(r0 + a) * P.R_cap + s] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------
// One-pass variant of K1/K2 for function-only models with R_cap <= 4096 (the headline configs):
//   l = W kvec and Y = W' l need W only ONCE: a CTA takes rows s = cta, cta + G, ... in order;
//   for each row it forms l[s] (block reduction) and immediately accumulates l[s] * W[s, :] into
//   per-thread column accumulators while the row is still in registers (the next row is already
//   in flight).  Each CTA leaves a partial Y; k_grid_wrow1 sums the partials in a fixed order, so
//   every rank computes bit-identical rows of W.  Half the bytes of the two-pass form and no W'.
// ---------------------------------------------------------------------------------------
constexpr int GS1_CPT = 16;                        // columns per thread: R_cap <= 16 * 256
constexpr int GS1_RMAX = GS1_CPT * GSOLVE_THREADS;

template <int D>
__global__ void __launch_bounds__(GSOLVE_THREADS, 2) k_grid_solve1(EngineParams P) {
  GPIS_DYN_SMEM_F64(s_kv);                           // [GS1_RMAX]
  __shared__ double s_red[2][GSOLVE_THREADS / 32];
  __shared__ int s_win;
  DevState* st = P.st;
  if (st->done) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gen = st->iter;
  if (!wait_records(P, st, gen)) return;
  const double* recs = records_ptr(P, gen);
  const int m = st->m;
  if (tid == 0) {
    int win = pick_winner(P, recs);
    if (win >= 0 && m >= P.M_cap) win = -1;
    s_win = win;
  }
  __syncthreads();
  const int win = s_win;
  const bool book = (blockIdx.x == gridDim.x - 1);
  if (win < 0) {
    if (book && tid == 0) st->done = 1;
    return;
  }
  const double* pay = recs + (long long)win * P.pay_stride;
  const long long* pl = reinterpret_cast<const long long*>(pay);
  const int r0 = st->r;
  double nx[D];
#pragma unroll
  for (int d = 0; d < D; ++d) nx[d] = pay[PAY_X + d];
  if (book) {
    if (tid == 0) {
      if (win == P.rank) P.flags[grid_slot_of_local(P, pl[PAY_LOCAL])] |= F_ACTIVE;
      P.sel_ids[m] = pl[PAY_ID];
      st->n_sel = m + 1;
#pragma unroll
      for (int d = 0; d < D; ++d) { st->nx[d] = nx[d]; P.AX[d][m] = nx[d]; }
      st->nnoise = pay[PAY_NOISE];
      st->r0 = r0;
      double mf = P.mean_c;
#pragma unroll
      for (int d = 0; d < D; ++d) mf += P.mean_w[d] * nx[d];
      st->gnew[0] = pay[PAY_TSDF] - mf;
    }
    grid_write_tables<1, D>(P, r0, nx);
    return;
  }
#pragma unroll 4
  for (int i = tid; i < GS1_RMAX; i += GSOLVE_THREADS) {
    double k = 0.0;
    if (i < m) {
      double d2 = 0.0;
#pragma unroll
      for (int d = 0; d < D; ++d) { const double df = P.AX[d][i] - nx[d]; d2 += df * df; }
      k = P.sf2 * exp(-0.5 * d2 * P.inv_ell2);
    }
    s_kv[i] = k;
  }
  __syncthreads();
  const int G = gridDim.x - 1;
  double acc[GS1_CPT], wa[GS1_CPT], wb[GS1_CPT];
#pragma unroll
  for (int i = 0; i < GS1_CPT; ++i) acc[i] = 0.0;
  auto load_row = [&](int row, double* w) {
    const double* wrow = P.W + (long long)row * P.ldW + tid;
#pragma unroll
    for (int i = 0; i < GS1_CPT; ++i) w[i] = (tid + GSOLVE_THREADS * i <= row) ? ld_nc_f64(wrow + GSOLVE_THREADS * i) : 0.0;
  };
  int par = 0;
  auto do_row = [&](int row, const double* w) {
    double p = 0.0;
#pragma unroll
    for (int i = 0; i < GS1_CPT; ++i) p += w[i] * s_kv[tid + GSOLVE_THREADS * i];
    for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
    if (lane == 0) s_red[par][warp] = p;
    __syncthreads();
    double ell = 0.0;
#pragma unroll
    for (int w8 = 0; w8 < GSOLVE_THREADS / 32; ++w8) ell += s_red[par][w8];
    par ^= 1;
    if (tid == 0) P.L[(long long)r0 * P.R_cap + row] = ell;
#pragma unroll
    for (int i = 0; i < GS1_CPT; ++i) acc[i] += ell * w[i];
  };
  int s = blockIdx.x;
  if (s < r0) load_row(s, wa);
  while (s < r0) {
    if (s + G < r0) load_row(s + G, wb);
    do_row(s, wa);
    s += G;
    if (s >= r0) break;
    if (s + G < r0) load_row(s + G, wa);
    do_row(s, wb);
    s += G;
  }
  if ((int)blockIdx.x < r0) {
    double* yp = P.ypart + (long long)blockIdx.x * GS1_RMAX + tid;
#pragma unroll
    for (int i = 0; i < GS1_CPT; ++i)
      if (tid + GSOLVE_THREADS * i < r0) yp[GSOLVE_THREADS * i] = acc[i];
  }
}

// new row of W from the partial Y's; d, g_new, L's diagonal entry.  Four columns per warp,
// eight lanes per column over the partials (fixed assignment + fixed shuffle tree).
__device__ __forceinline__ void grid_wrow1_body(const EngineParams& P, DevState* st, int r0, int n_solve_ctas) {
  __shared__ double s_gram[GSOLVE_THREADS / 32][2];
  __shared__ double s_dinv;
  const int tid = threadIdx.x, lane = tid & 31;
  const double* X = P.L + (long long)r0 * P.R_cap;
  double p0 = 0.0, p1 = 0.0;
  for (int s = tid; s < r0; s += GSOLVE_THREADS) { const double x = X[s]; p0 += x * x; p1 += x * P.g[s]; }
  for (int o = 16; o > 0; o >>= 1) { p0 += __shfl_xor_sync(0xffffffffu, p0, o); p1 += __shfl_xor_sync(0xffffffffu, p1, o); }
  if (lane == 0) { s_gram[tid >> 5][0] = p0; s_gram[tid >> 5][1] = p1; }
  __syncthreads();
  if (tid == 0) {
    double xx = 0.0, xg = 0.0;
    for (int w = 0; w < GSOLVE_THREADS / 32; ++w) { xx += s_gram[w][0]; xg += s_gram[w][1]; }
    double sdiag = P.sf2 + st->nnoise - xx;
    bool ok = sdiag > 0.0;
    if (!ok) sdiag = 1.0;
    const double d = sqrt(sdiag), dinv = 1.0 / d;
    s_dinv = dinv;
    if (blockIdx.x == 0) {
      const double gn = (st->gnew[0] - xg) / d;
      P.g[r0] = gn;
      P.L[(long long)r0 * P.R_cap + r0] = d;
      P.W[(long long)r0 * P.ldW + r0] = dinv;
      st->gnew[0] = gn;
      st->r = r0 + 1;
      st->m = st->m + 1;
      if (!ok) { st->err = -5; st->done = 1; }
    }
  }
  __syncthreads();
  const double dinv = s_dinv;
  const int nparts = min(n_solve_ctas, r0);
  const int gw = (blockIdx.x * blockDim.x + tid) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const int sub = lane >> 2, cq = lane & 3;
  for (int c0 = gw * 4; c0 < r0; c0 += nw * 4) {
    const int c = c0 + cq;
    double y = 0.0;
    if (c < r0) {
      const double* yp = P.ypart + c;
      // every solve CTA < nparts wrote all r0 columns of its partial (zeros included): plain
      // strided sum, five independent loads in flight per lane
      int part = sub;
      for (; part + 32 < nparts; part += 40) {
        const double a0 = __ldcg(yp + (long long)part * GS1_RMAX);
        const double a1 = __ldcg(yp + (long long)(part + 8) * GS1_RMAX);
        const double a2 = __ldcg(yp + (long long)(part + 16) * GS1_RMAX);
        const double a3 = __ldcg(yp + (long long)(part + 24) * GS1_RMAX);
        const double a4 = __ldcg(yp + (long long)(part + 32) * GS1_RMAX);
        y += a0; y += a1; y += a2; y += a3; y += a4;
      }
      for (; part < nparts; part += 8) y += __ldcg(yp + (long long)part * GS1_RMAX);
    }
    y += __shfl_xor_sync(0xffffffffu, y, 4);
    y += __shfl_xor_sync(0xffffffffu, y, 8);
    y += __shfl_xor_sync(0xffffffffu, y, 16);
    if (sub == 0 && c < r0) P.W[(long long)r0 * P.ldW + c] = -y * dinv;
  }
}

// Same pass with the rows of W streamed through a shared-memory ring by bulk async copies (one
// copy per row, GS2_STAGES rows in flight per CTA, one CTA per SM): no register staging, so the
// depth of the prefetch no longer competes with the column accumulators.
constexpr int GS2_STAGES = 5;
constexpr size_t GS2_SMEM = sizeof(double) * GS1_RMAX * (1 + GS2_STAGES) + 8 * GS2_STAGES;

template <int D>
__global__ void __launch_bounds__(GSOLVE_THREADS, 1) k_grid_solve2(EngineParams P) {
  GPIS_DYN_SMEM_F64A(s_dyn);
  double* s_kv = s_dyn;                               // [GS1_RMAX]
  double* ring = s_dyn + GS1_RMAX;                    // [GS2_STAGES][GS1_RMAX]
  unsigned long long* full = reinterpret_cast<unsigned long long*>(ring + (size_t)GS2_STAGES * GS1_RMAX);
  __shared__ double s_red[2][GSOLVE_THREADS / 32];
  __shared__ int s_win;
  DevState* st = P.st;
  if (st->done) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gen = st->iter;
  if (!wait_records(P, st, gen)) return;
  const double* recs = records_ptr(P, gen);
  const int m = st->m;
  if (tid == 0) {
    int win = pick_winner(P, recs);
    if (win >= 0 && m >= P.M_cap) win = -1;
    s_win = win;
  }
  __syncthreads();
  const int win = s_win;
  const bool book = (blockIdx.x == gridDim.x - 1);
  if (win < 0) {
    if (book && tid == 0) st->done = 1;
    return;
  }
  const double* pay = recs + (long long)win * P.pay_stride;
  const long long* pl = reinterpret_cast<const long long*>(pay);
  const int r0 = st->r;
  double nx[D];
#pragma unroll
  for (int d = 0; d < D; ++d) nx[d] = pay[PAY_X + d];
  if (book) {
    if (tid == 0) {
      if (win == P.rank) P.flags[grid_slot_of_local(P, pl[PAY_LOCAL])] |= F_ACTIVE;
      P.sel_ids[m] = pl[PAY_ID];
      st->n_sel = m + 1;
#pragma unroll
      for (int d = 0; d < D; ++d) { st->nx[d] = nx[d]; P.AX[d][m] = nx[d]; }
      st->nnoise = pay[PAY_NOISE];
      st->r0 = r0;
      double mf = P.mean_c;
#pragma unroll
      for (int d = 0; d < D; ++d) mf += P.mean_w[d] * nx[d];
      st->gnew[0] = pay[PAY_TSDF] - mf;
    }
    grid_write_tables<1, D>(P, r0, nx);
  }
  const int G = gridDim.x - 1;
  const int my_rows = (!book && (int)blockIdx.x < r0) ? (r0 - 1 - (int)blockIdx.x) / G + 1 : 0;   // rows blockIdx.x + i*G
  if (tid == 0) {
    for (int i = 0; i < GS2_STAGES; ++i) mbar_init(&full[i], 1);
    mbar_init_fence();
  }
  __syncthreads();
  auto issue = [&](int i) {          // thread 0: row i of this CTA into ring slot i % STAGES
    const int row = (int)blockIdx.x + i * G;
    const unsigned bytes = (unsigned)(((row + 2) >> 1) << 4);          // row + 1 doubles, rounded up to 16 B
    unsigned long long* bar = &full[i % GS2_STAGES];
    mbar_arrive_expect_tx(bar, bytes);
    bulk_g2s(ring + (size_t)(i % GS2_STAGES) * GS1_RMAX, P.W + (long long)row * P.ldW, bytes, bar);
  };
  if (tid == 0)
    for (int i = 0; i < GS2_STAGES && i < my_rows; ++i) issue(i);
#pragma unroll 4
  for (int i = tid; i < GS1_RMAX; i += GSOLVE_THREADS) {
    double k = 0.0;
    if (i < m) {
      double d2 = 0.0;
#pragma unroll
      for (int d = 0; d < D; ++d) { const double df = P.AX[d][i] - nx[d]; d2 += df * df; }
      k = P.sf2 * exp(-0.5 * d2 * P.inv_ell2);
    }
    s_kv[i] = k;
  }
  __syncthreads();
  double acc[GS1_CPT];
#pragma unroll
  for (int i = 0; i < GS1_CPT; ++i) acc[i] = 0.0;
  int par = 0;
  for (int i = 0; i < my_rows; ++i) {
    const int row = (int)blockIdx.x + i * G;
    const double* buf = ring + (size_t)(i % GS2_STAGES) * GS1_RMAX;
    mbar_wait(&full[i % GS2_STAGES], (unsigned)((i / GS2_STAGES) & 1));
    double w[GS1_CPT];
    double p = 0.0;
#pragma unroll
    for (int c = 0; c < GS1_CPT; ++c) {
      const int col = tid + GSOLVE_THREADS * c;
      w[c] = col <= row ? buf[col] : 0.0;
      p += w[c] * s_kv[col];
    }
    for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
    if (lane == 0) s_red[par][warp] = p;
    __syncthreads();                 // also: every thread has read this row's slot
    double ell = 0.0;
#pragma unroll
    for (int w8 = 0; w8 < GSOLVE_THREADS / 32; ++w8) ell += s_red[par][w8];
    par ^= 1;
    if (tid == 0) {
      P.L[(long long)r0 * P.R_cap + row] = ell;
      if (i + GS2_STAGES < my_rows) issue(i + GS2_STAGES);   // refill the slot just drained
    }
#pragma unroll
    for (int c = 0; c < GS1_CPT; ++c) acc[c] += ell * w[c];
  }
  if (!book && (int)blockIdx.x < r0) {
    double* yp = P.ypart + (long long)blockIdx.x * GS1_RMAX + tid;
#pragma unroll
    for (int c = 0; c < GS1_CPT; ++c)
      if (tid + GSOLVE_THREADS * c < r0) yp[GSOLVE_THREADS * c] = acc[c];
  }
  // ---- grid barrier (one CTA per SM, all resident), then the former k_grid_wrow1 in the same launch:
  //      the partial Y's and the new row of L are complete, every CTA finishes a slice of W's new row
  __threadfence();
  __syncthreads();
  __shared__ int s_bar_to;
  if (tid == 0) {
    s_bar_to = 0;
    const unsigned target = (unsigned)(gen + 1) * gridDim.x;
    atomicAdd(&st->grid_bar, 1u);
    const long long t0 = clock64();
    while (*(volatile unsigned*)&st->grid_bar < target) {
      if (clock64() - t0 > P.spin_cycles) { s_bar_to = 1; break; }
      __nanosleep(64);
    }
    __threadfence();
  }
  __syncthreads();
  if (s_bar_to) {
    if (tid == 0) { st->err = -7; st->done = 1; }
    return;
  }
  grid_wrow1_body(P, st, r0, G);
}

__global__ void __launch_bounds__(GSOLVE_THREADS) k_grid_wrow1(EngineParams P, int n_solve_ctas) {
  DevState* st = P.st;
  if (st->done) return;
  grid_wrow1_body(P, st, st->r0, n_solve_ctas);
}

// W' from W (the grid backend's one-pass mode does not maintain it): 32 x 32 tiles
__global__ void __launch_bounds__(256) k_transpose_lower(const double* __restrict__ W, double* Wt, long long ld, int R) {
  __shared__ double t[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;      // Wt[bx + i][by + j] = W[by + j][bx + i]
  if (bx > by + 31) return;                                   // strictly upper block of W: zero
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int row = by + i, col = bx + threadIdx.x;
    t[i][threadIdx.x] = (row < R && col < R) ? W[(long long)row * ld + col] : 0.0;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int row = bx + i, col = by + threadIdx.x;
    if (row < R && col < R) Wt[(long long)row * ld + col] = t[threadIdx.x][i];
  }
}

// ---------------------------------------------------------------------------------------
// K2: S = Q(new,new) - X'X = Lnn Lnn';  new rows of W = Lnn^-1 [-X'W | I]  (and of W');
//     g_new = Lnn^-1 (y - m - X'g).  Every CTA recomputes the small Gram block.
// ---------------------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(GSOLVE_THREADS) k_grid_wrow(EngineParams P) {
  DevState* st = P.st;
  if (st->done) return;
  __shared__ double s_dot[NB][NB + 1];
  __shared__ double s_linv[NB][NB];
  __shared__ int s_ok;
  const int tid = threadIdx.x, lane = tid & 31;
  const int r0 = st->r0;
  const double* X[NB];
#pragma unroll
  for (int a = 0; a < NB; ++a) X[a] = P.L + (long long)(r0 + a) * P.R_cap;
  {
    // Gram block X'X (lower) and X'g in ONE sweep over the rows, one block reduction
    constexpr int ND = NB * (NB + 1) / 2 + NB;
    __shared__ double s_gram[GSOLVE_THREADS / 32][ND];
    double p[ND];
#pragma unroll
    for (int i = 0; i < ND; ++i) p[i] = 0.0;
    for (int s = tid; s < r0; s += GSOLVE_THREADS) {
      double x[NB];
#pragma unroll
      for (int a = 0; a < NB; ++a) x[a] = X[a][s];
      const double gs = P.g[s];
      int i = 0;
#pragma unroll
      for (int a = 0; a < NB; ++a) {
#pragma unroll
        for (int b = 0; b <= a; ++b) p[i++] += x[a] * x[b];
        p[i++] += x[a] * gs;
      }
    }
#pragma unroll
    for (int i = 0; i < ND; ++i)
      for (int o = 16; o > 0; o >>= 1) p[i] += __shfl_xor_sync(0xffffffffu, p[i], o);
    if (lane == 0)
#pragma unroll
      for (int i = 0; i < ND; ++i) s_gram[tid >> 5][i] = p[i];
    __syncthreads();
    if (tid == 0) {
      int i = 0;
      for (int a = 0; a < NB; ++a) {
        for (int b = 0; b <= a + 1; ++b) {
          double t = 0.0;
          for (int w = 0; w < GSOLVE_THREADS / 32; ++w) t += s_gram[w][i];
          ++i;
          if (b <= a) s_dot[a][b] = t; else s_dot[a][NB] = t;
        }
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    const double nnoise = st->nnoise;
    double Qn[NB][NB], Ln[NB][NB], Li[NB][NB];
    for (int a = 0; a < NB; ++a)
      for (int b = 0; b < NB; ++b) Qn[a][b] = 0.0;
    Qn[0][0] = P.sf2 + nnoise;
    for (int d = 1; d < NB; ++d) Qn[d][d] = (P.sf2 + nnoise) * P.inv_ell2;   // noise leak, se_cov_derivative.m:28-35
    bool ok = true;
    for (int a = 0; a < NB; ++a) {
      for (int b = 0; b <= a; ++b) {
        double s = Qn[a][b] - s_dot[a][b];
        for (int t = 0; t < b; ++t) s -= Ln[a][t] * Ln[b][t];
        if (a == b) { if (!(s > 0.0)) { ok = false; s = 1.0; } Ln[a][a] = sqrt(s); }
        else Ln[a][b] = s / Ln[b][b];
      }
      for (int b = a + 1; b < NB; ++b) Ln[a][b] = 0.0;
    }
    for (int c = 0; c < NB; ++c)                       // Li = Ln^-1 (lower), column by column
      for (int a = 0; a < NB; ++a) {
        double s = (a == c) ? 1.0 : 0.0;
        for (int t = 0; t < a; ++t) s -= Ln[a][t] * Li[t][c];
        Li[a][c] = (a >= c) ? s / Ln[a][a] : 0.0;
      }
    for (int a = 0; a < NB; ++a)
      for (int b = 0; b < NB; ++b) s_linv[a][b] = Li[a][b];
    s_ok = ok;
    if (blockIdx.x == 0) {
      double gn[NB];
      for (int a = 0; a < NB; ++a) {
        double s = st->gnew[a] - s_dot[a][NB];
        for (int t = 0; t < a; ++t) s -= Ln[a][t] * gn[t];
        gn[a] = s / Ln[a][a];
      }
      for (int a = 0; a < NB; ++a) {
        P.g[r0 + a] = gn[a];
        for (int b = 0; b <= a; ++b) {
          P.L[(long long)(r0 + a) * P.R_cap + r0 + b] = Ln[a][b];
          P.W[(long long)(r0 + a) * P.ldW + r0 + b] = Li[a][b];
          P.Wt[(long long)(r0 + b) * P.ldW + r0 + a] = Li[a][b];
        }
      }
      for (int a = 0; a < NB; ++a) st->gnew[a] = gn[a];
      st->r = r0 + NB;
      st->m = st->m + 1;
      if (!ok) { st->err = -5; st->done = 1; }
    }
  }
  __syncthreads();
  const int wid = (blockIdx.x * blockDim.x + tid) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int rho = wid; rho < r0; rho += nw) {
    const double* wt = P.Wt + (long long)rho * P.ldW;     // Wt[rho][s] = W[s][rho], s >= rho
    double acc[NB];
#pragma unroll
    for (int a = 0; a < NB; ++a) acc[a] = 0.0;
    int s = rho + lane;
    for (; s + 32 * (GSU - 1) < r0; s += 32 * GSU) {
      double w[GSU];
#pragma unroll
      for (int u = 0; u < GSU; ++u) w[u] = ld_nc_f64(wt + s + 32 * u);
      GPIS_ISSUE_FENCE();
#pragma unroll
      for (int u = 0; u < GSU; ++u)
#pragma unroll
        for (int a = 0; a < NB; ++a) acc[a] += w[u] * X[a][s + 32 * u];
    }
    for (; s < r0; s += 32) {
      const double w = __ldg(wt + s);
#pragma unroll
      for (int a = 0; a < NB; ++a) acc[a] += w * X[a][s];
    }
#pragma unroll
    for (int a = 0; a < NB; ++a)
      for (int o = 16; o > 0; o >>= 1) acc[a] += __shfl_xor_sync(0xffffffffu, acc[a], o);
    if (lane == 0) {
#pragma unroll
      for (int a = 0; a < NB; ++a) {
        double v = 0.0;
#pragma unroll
        for (int b = 0; b < NB; ++b)
          if (b <= a) v -= s_linv[a][b] * acc[b];
        P.W[(long long)(r0 + a) * P.ldW + rho] = v;
        P.Wt[(long long)rho * P.ldW + r0 + a] = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// K3a: the per-step update as a split-K GEMM on the fp64 tensor pipe (DMMA.8x8x4).
//
// Persistent: one CTA per SM.  Work unit = (128x128 tile of one z-slice, 16-row chunk of the
// model); the flattened unit range is cut into gridDim.x equal pieces (stream-K style), so
// every SM gets the same share whatever the tile count is (128 tiles on 148 SMs; 16 tiles per
// GPU when the lattice is sharded 8 ways).  Each CTA writes the partial accumulators of the
// (at most a few) tile segments it covered to a workspace; K3b sums them and runs the epilogue.
//
// Warp-specialised: the last warp is the producer -- one bulk async copy (TMA engine, UBLKCP) per
// 1 KB axis-table row into padded shared-memory rows, completion on an mbarrier; warps 0-7
// consume (32x64 warp tiles = 4x8 DMMA tiles) and hand stages back through "empty" mbarriers.
// No CTA-wide barrier in the main loop.
// ---------------------------------------------------------------------------------------
constexpr int GSTAGE_DOUBLES = 2 * GKC * GLD + GKC;
constexpr int GPIPE = 4;
constexpr int GEMM_CWARPS = 8;                       // consumer warps: 4 (jy) x 2 (jx), 32 x 64 each
constexpr int GEMM_THREADS = (GEMM_CWARPS + 1) * 32;   // + 1 producer warp
constexpr size_t GEMM_SMEM = (size_t)GPIPE * GSTAGE_DOUBLES * sizeof(double) + 2 * GPIPE * sizeof(unsigned long long);
constexpr int GTILE_ELEMS = GT * GT;

// owner CTA of flattened unit u when U units are cut into G ranges [g*U/G, (g+1)*U/G)
__device__ __forceinline__ int unit_owner(long long u, long long U, int G) {
  return (int)(((u + 1) * G + U - 1) / U) - 1;
}

// exp(-d^2 / (2 ell^2)) < 2^-70  <=>  d^2 / (2 ell^2) > 70 ln 2: a model row farther than that from every point of a
// tile cannot change a double-precision sum over the tile (the dropped term is below 2^-70 |W|)
constexpr double GRID_REACH_CUT = 48.52030263919617;
constexpr int GLIST_TMAX = 2048;      // most 128 x 128 tiles per rank the list form of the per-step GEMM takes

// Per-step path with ROW LISTS (models the persistent kernel does not take: gradient observations, more than 4096
// rows): after the factor append, the NB rows of the new point join the list of every 128 x 128 tile they can reach
// and the exclusive prefix of the tiles' 16-row chunk counts is rebuilt.  One CTA.
template <int NB, int D>
__global__ void __launch_bounds__(1024, 1) k_grid_lists(EngineParams P) {
  const DevState* st = P.st;
  if (st->done) return;
  __shared__ int s_part[1024];
  const int tid = threadIdx.x;
  const int r0 = st->r0;
  const int nx = P.glen[0], ny = P.glen[1];
  const int tiles_x = (nx + GT - 1) / GT, tiles_y = (ny + GT - 1) / GT, txy = tiles_x * tiles_y;
  const int T = txy * P.glen[2];
  double p[D];
#pragma unroll
  for (int d = 0; d < D; ++d) p[d] = st->nx[d];
  const int per = (T + 1023) / 1024;
  int loc = 0;
  for (int k = 0; k < per; ++k) {
    const int t = tid * per + k;
    if (t >= T) break;
    const int tz = t / txy, ty = (t / tiles_x) % tiles_y, tx = t % tiles_x;
    double d2 = 0.0;
    {
      const double lo = P.gorg[0] + P.gh[0] * (double)(tx * GT), hi = P.gorg[0] + P.gh[0] * (double)(min(tx * GT + GT, nx) - 1);
      const double a = fmin(lo, hi), b = fmax(lo, hi);
      const double df = p[0] < a ? a - p[0] : (p[0] > b ? p[0] - b : 0.0);
      d2 += df * df;
    }
    if (D > 1) {
      const double q = p[D > 1 ? 1 : 0];
      const double lo = P.gorg[1] + P.gh[1] * (double)(ty * GT), hi = P.gorg[1] + P.gh[1] * (double)(min(ty * GT + GT, ny) - 1);
      const double a = fmin(lo, hi), b = fmax(lo, hi);
      const double df = q < a ? a - q : (q > b ? q - b : 0.0);
      d2 += df * df;
    }
    if (D > 2) {
      const double df = (P.gorg[2] + P.gh[2] * (double)(P.gz0 + tz)) - p[D > 2 ? 2 : 0];
      d2 += df * df;
    }
    int n = P.rcnt[t];
    if (0.5 * d2 * P.inv_ell2 <= GRID_REACH_CUT) {
      for (int a = 0; a < NB; ++a) P.rlist[(size_t)t * P.rlist_stride + n + a] = (unsigned short)(r0 + a);
      n += NB;
      P.rcnt[t] = n;
    }
    const int c = (n + GKC - 1) / GKC;
    loc += c > 0 ? c : 1;
  }
  // exclusive prefix over the tiles (thread tid owns tiles tid*per .. tid*per + per - 1)
  s_part[tid] = loc;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int v = tid >= o ? s_part[tid - o] : 0;
    __syncthreads();
    s_part[tid] += v;
    __syncthreads();
  }
  int run = s_part[tid] - loc;
  for (int k = 0; k < per; ++k) {
    const int t = tid * per + k;
    if (t >= T) break;
    P.upref[t] = run;
    const int c = (P.rcnt[t] + GKC - 1) / GKC;
    run += c > 0 ? c : 1;
  }
  if (tid == 1023) P.upref[T] = s_part[1023];
}

template <int NB, bool LISTS>
__global__ void __launch_bounds__(GEMM_THREADS, 1) k_grid_gemm(EngineParams P) {
  GPIS_DYN_SMEM_F64A(g_smem);
  const DevState* st = P.st;
  if (st->done) return;
  __shared__ int s_upref[LISTS ? GLIST_TMAX + 1 : 1];
  unsigned long long* full = reinterpret_cast<unsigned long long*>(g_smem + (size_t)GPIPE * GSTAGE_DOUBLES);
  unsigned long long* empty = full + GPIPE;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int i = 0; i < GPIPE; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], GEMM_CWARPS); }
    mbar_init_fence();
  }
  __syncthreads();
  const int r0 = st->r0;
  const int nx = P.glen[0], ny = P.glen[1];
  const int tiles_x = (nx + GT - 1) / GT, tiles_y = (ny + GT - 1) / GT;
  const int T = tiles_x * tiles_y * P.glen[2];
  const int C = (r0 + NB + GKC - 1) / GKC;           // chunks per tile (all passes share the count)
  if (LISTS) {       // per-tile chunk counts: their exclusive prefix (k_grid_lists) into shared memory
    for (int i = tid; i <= T; i += GEMM_THREADS) s_upref[i] = P.upref[i];
    __syncthreads();
  }
  const long long U = LISTS ? (long long)s_upref[T] : (long long)T * C;
  // fewer units than CTAs: only the first U CTAs work, so every working CTA's range is non-empty
  const int G = (int)min((long long)gridDim.x, U), g = blockIdx.x;
  if (g >= G) return;
  const long long u0 = (long long)g * U / G, u1 = (long long)(g + 1) * U / G;
  int tile_first = 0;
  if (LISTS) {       // last tile whose first unit is <= u0
    int lo = 0, hi = T - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if ((long long)s_upref[mid] <= u0) lo = mid; else hi = mid - 1; }
    tile_first = lo;
  }
  int stage = 0;
  unsigned phase = 0;

  if (warp == GEMM_CWARPS) {
    // ------------------------------- producer -------------------------------
    for (int a = 0; a < NB; ++a) {
      const double* wrow = P.W + (long long)(r0 + a) * P.ldW;
      const int klen = r0 + a + 1;
      int ltile = tile_first;
      for (long long u = u0; u < u1; ++u) {
        int tile, c;
        if (LISTS) {
          while ((long long)s_upref[ltile + 1] <= u) ++ltile;
          tile = ltile;
          c = (int)(u - s_upref[ltile]);
        } else {
          tile = (int)(u / C);
          c = (int)(u % C);
        }
        const int tz = tile / (tiles_x * tiles_y), ty = (tile / tiles_x) % tiles_y, tx = tile % tiles_x;
        mbar_wait(&empty[stage], phase ^ 1);
        double* As = g_smem + (size_t)stage * GSTAGE_DOUBLES;
        double* Bs = As + GKC * GLD;
        double* cs = Bs + GKC * GLD;
        const long long rho0 = (long long)c * GKC;
        long long rho = rho0 + (lane & (GKC - 1));      // LISTS: the chunk's row of the tile's list (row 0 past its end)
        bool live = true;
        if (LISTS) {
          const int k = (int)rho;
          live = k < P.rcnt[tile];
          rho = live ? (long long)P.rlist[(size_t)tile * P.rlist_stride + k] : 0;
        }
        if (lane < GKC) {
          double cv = (live && rho < klen) ? wrow[rho] : 0.0;
          cv *= P.T[2][rho * P.ldt[2] + P.gz0 + tz];
          cs[lane] = cv;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_expect_tx(&full[stage], 2 * GKC * GT * (unsigned)sizeof(double));
        __syncwarp();
        if (lane < GKC)
          bulk_g2s(As + lane * GLD, P.T[1] + rho * P.ldt[1] + ty * GT, GT * sizeof(double), &full[stage]);
        else
          bulk_g2s(Bs + (lane - GKC) * GLD, P.T[0] + rho * P.ldt[0] + tx * GT, GT * sizeof(double), &full[stage]);
        if (++stage == GPIPE) { stage = 0; phase ^= 1; }
      }
    }
    return;
  }

  // --------------------------------- consumers ---------------------------------
  const int wm = warp >> 1, wn = warp & 1;
  for (int a = 0; a < NB; ++a) {
    long long u = u0;
    int ltile = tile_first;
    while (u < u1) {
      if (LISTS) while ((long long)s_upref[ltile + 1] <= u) ++ltile;
      const int tile = LISTS ? ltile : (int)(u / C);
      const long long uend = min(u1, LISTS ? (long long)s_upref[tile + 1] : (long long)(tile + 1) * C);
      double acc[4][8][2];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
      for (; u < uend; ++u) {
        mbar_wait(&full[stage], phase);
        const double* As = g_smem + (size_t)stage * GSTAGE_DOUBLES;
        const double* Bs = As + GKC * GLD;
        const double* cs = Bs + GKC * GLD;
#pragma unroll
        for (int k4 = 0; k4 < GKC / 4; ++k4) {
          const int kb = k4 * 4 + (lane & 3);
          const double cv = cs[kb];
          double af[4], bf[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) af[i] = As[kb * GLD + wm * 32 + i * 8 + (lane >> 2)] * cv;
#pragma unroll
          for (int j = 0; j < 8; ++j) bf[j] = Bs[kb * GLD + wn * 64 + j * 8 + (lane >> 2)];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) dmma8x8x4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == GPIPE) { stage = 0; phase ^= 1; }
      }
      // partial accumulators of this (tile, segment) -> workspace, fragment-native layout
      const int seg = g - unit_owner(LISTS ? (long long)s_upref[tile] : (long long)tile * C, U, G);
      // LISTS: the first segment of a tile in the tile's slot, a later one (always the first tile of its CTA's range)
      // in the CTA's slot -- no bound on the segments per tile is needed
      double2* ws = reinterpret_cast<double2*>(
          LISTS ? P.gws + (seg == 0 ? ((long long)a * T + tile) : ((long long)NB * T + (long long)a * gridDim.x + g)) * GTILE_ELEMS
                : P.gws + (((long long)a * T + tile) * P.gws_segmax + seg) * GTILE_ELEMS);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          ws[(warp * 32 + i * 8 + j) * 32 + lane] = make_double2(acc[i][j][0], acc[i][j][1]);
    }
  }
}

// ---------------------------------------------------------------------------------------
// K3b: sum the split-K partials, update the posterior, straddle score, classify, arg-max.
// One thread per accumulator fragment slot (two x-adjacent candidates): every access to the
// workspace and to the per-candidate state is a coalesced 16-byte (flags: 2-byte) access.
// ---------------------------------------------------------------------------------------
constexpr int GEPI_THREADS = 256;
constexpr int GEPI_SPT = 4;                                         // fragment slots per thread
constexpr int GEPI_CTAS_PER_TILE = (GTILE_ELEMS / 2) / (GEPI_THREADS * GEPI_SPT);

template <int NB>
__global__ void __launch_bounds__(GEPI_THREADS) k_grid_epilogue(EngineParams P) {
  DevState* st = P.st;
  if (st->done) return;
  const int tid = threadIdx.x, lane = tid & 31, ew = tid >> 5;
  const int tile = blockIdx.x / GEPI_CTAS_PER_TILE;
  const int q0 = (blockIdx.x % GEPI_CTAS_PER_TILE) * (GEPI_THREADS * GEPI_SPT) + tid;   // + k * GEPI_THREADS
  const int nx = P.glen[0], ny = P.glen[1];
  const int tiles_x = (nx + GT - 1) / GT, tiles_y = (ny + GT - 1) / GT;
  const int T = tiles_x * tiles_y * P.glen[2];
  const double NEG_INF = -__longlong_as_double(0x7ff0000000000000LL);
  __shared__ int s_nseg, s_gfirst;
  __shared__ double s_sb, s_gn[NB];
  if (tid == 0) {                       // CTA-uniform: split-K segment count of this tile, sqrt(beta_t)
    const int r0 = st->r0;
    const int C = (r0 + NB + P.gemm_kc - 1) / P.gemm_kc;
    // the GEMM's work tile is a group of gemm_zgroup z-slices (1 for DMMA, 2 for tcgen05)
    const int txy = tiles_x * tiles_y, zg = P.gemm_zgroup;
    const int gtile = (tile / txy / zg) * txy + tile % txy;
    const long long U = (long long)txy * ((P.glen[2] + zg - 1) / zg) * C;
    const int G = (int)min((long long)P.gemm_ctas, U);
    int g_first = unit_owner((long long)gtile * C, U, G);
    s_nseg = unit_owner((long long)(gtile + 1) * C - 1, U, G) - g_first + 1;
    if (P.upref) {      // row lists: chunk counts differ per tile
      const long long Ul = P.upref[T];
      const int Gl = (int)min((long long)P.gemm_ctas, Ul);
      g_first = unit_owner((long long)P.upref[tile], Ul, Gl);
      s_nseg = unit_owner((long long)P.upref[tile + 1] - 1, Ul, Gl) - g_first + 1;
    }
    s_gfirst = g_first;
    double beta = P.beta;
    if (P.beta_mode == 1) beta = 2.0 * log((double)P.N_total * (double)(st->iter + 1) * 9.869604401089358 / (6.0 * P.beta_delta));
    else if (P.beta_mode == 2) { const double k1 = (double)(st->iter + 1); beta = 2.0 * log((double)P.N_total * 9.869604401089358 * k1 * k1 / (6.0 * P.beta_delta)); }
    s_sb = sqrt(beta);
    for (int a = 0; a < NB; ++a) s_gn[a] = st->gnew[a];
  }
  __syncthreads();
  const int nseg = s_nseg;
  const double sb = s_sb;

  // phase 1: all loads of the thread's slots in flight together
  double dv[GEPI_SPT][2], dm[GEPI_SPT][2];
#pragma unroll
  for (int k = 0; k < GEPI_SPT; ++k) dv[k][0] = dv[k][1] = dm[k][0] = dm[k][1] = 0.0;
#pragma unroll
  for (int a = 0; a < NB; ++a) {
    const bool lists = P.upref != nullptr;
    const double2* ws = reinterpret_cast<const double2*>(
        P.gws + (lists ? ((long long)a * T + tile) : (((long long)a * T + tile) * P.gws_segmax)) * GTILE_ELEMS) + q0;
    // row lists: segments after the first sit in the slots of the CTAs that computed them
    const double2* wsx = reinterpret_cast<const double2*>(
        P.gws + ((long long)NB * T + (long long)a * P.gemm_ctas + s_gfirst) * GTILE_ELEMS) + q0;
    double2 v[GEPI_SPT];
#pragma unroll
    for (int k = 0; k < GEPI_SPT; ++k) v[k] = __ldcg(ws + k * GEPI_THREADS);
    for (int sgi = 1; sgi < nseg; ++sgi) {
      double2 pz[GEPI_SPT];
#pragma unroll
      for (int k = 0; k < GEPI_SPT; ++k) pz[k] = __ldcg((lists ? wsx : ws) + (long long)sgi * (GTILE_ELEMS / 2) + k * GEPI_THREADS);
#pragma unroll
      for (int k = 0; k < GEPI_SPT; ++k) { v[k].x += pz[k].x; v[k].y += pz[k].y; }
    }
    const double ga = s_gn[a];
#pragma unroll
    for (int k = 0; k < GEPI_SPT; ++k) {
      dv[k][0] += v[k].x * v[k].x; dv[k][1] += v[k].y * v[k].y;
      dm[k][0] += v[k].x * ga;     dm[k][1] += v[k].y * ga;
    }
  }
  const long long slot_base = (long long)tile * (GTILE_ELEMS / 2) + q0;
  double2 mu[GEPI_SPT], var[GEPI_SPT];
  uchar2 fl[GEPI_SPT];
#pragma unroll
  for (int k = 0; k < GEPI_SPT; ++k) {
    mu[k] = reinterpret_cast<const double2*>(P.mu)[slot_base + k * GEPI_THREADS];
    var[k] = reinterpret_cast<const double2*>(P.var)[slot_base + k * GEPI_THREADS];
    fl[k] = reinterpret_cast<const uchar2*>(P.flags)[slot_base + k * GEPI_THREADS];
  }

  double best_s = NEG_INF;
  long long best_id = 0x7fffffffffffffffLL, best_j = -1;
  int n_scored = 0;
#pragma unroll
  for (int k = 0; k < GEPI_SPT; ++k) {
    const long long slot2 = slot_base + k * GEPI_THREADS;
    mu[k].x += dm[k][0]; mu[k].y += dm[k][1];
    var[k].x -= dv[k][0]; var[k].y -= dv[k][1];
    reinterpret_cast<double2*>(P.mu)[slot2] = mu[k];
    reinterpret_cast<double2*>(P.var)[slot2] = var[k];
    unsigned char f[2] = {fl[k].x, fl[k].y};
    if ((f[0] & (F_ACTIVE | F_HIGH | F_LOW)) && (f[1] & (F_ACTIVE | F_HIGH | F_LOW))) continue;
    const double mus[2] = {mu[k].x, mu[k].y}, vars[2] = {var[k].x, var[k].y};
    int jx, jy, tz;
    grid_unslot_te<GT>(P, slot2 * 2, jx, jy, tz);       // the per-step kernels work on 128 x 128 tiles
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      if (f[e] & (F_ACTIVE | F_HIGH | F_LOW)) continue;
      ++n_scored;
      const long long jl = (long long)(jx + e) + (long long)nx * (jy + (long long)ny * tz);
      double ve = vars[e];
      if (P.variance_mode == 1) ve = __dadd_rn(ve, P.noise ? P.noise[jl] : P.noise_scalar);
      const double w = __dmul_rn(sb, ve);                  // variance, not sigma: straddle.m:106-107
      const double hi = __dadd_rn(mus[e], w), lo = __dadd_rn(mus[e], -w);
      const double sc = fmin(__dadd_rn(hi, -P.level), __dadd_rn(P.level, -lo));
      if (__dadd_rn(lo, P.eps) > P.level) f[e] |= F_HIGH;   // :122-125
      if (__dadd_rn(hi, -P.eps) < P.level) f[e] |= F_LOW;
      if (sc == sc) {
        const long long id = cand_id(P, jl);
        if (better(sc, id, best_s, best_id)) { best_s = sc; best_id = id; best_j = jl; }
      }
    }
    // the ambiguity is not stored per step (16 bytes of traffic per candidate): the getter
    // recomputes it from the running posterior
    if (f[0] != fl[k].x || f[1] != fl[k].y) reinterpret_cast<uchar2*>(P.flags)[slot2] = make_uchar2(f[0], f[1]);
  }
  const int r0 = st->r0;

  __shared__ BlockBest s_bb[GEPI_THREADS / 32];
  __shared__ int s_cnt[GEPI_THREADS / 32];
  __shared__ int s_last;
  for (int o = 16; o > 0; o >>= 1) {
    const double os = __shfl_xor_sync(0xffffffffu, best_s, o);
    const long long oid = __shfl_xor_sync(0xffffffffu, best_id, o);
    const long long oj = __shfl_xor_sync(0xffffffffu, best_j, o);
    if (oj >= 0 && better(os, oid, best_s, best_id)) { best_s = os; best_id = oid; best_j = oj; }
    n_scored += __shfl_xor_sync(0xffffffffu, n_scored, o);
  }
  if (lane == 0) { s_bb[ew].score = best_s; s_bb[ew].id = best_id; s_bb[ew].local = best_j; s_cnt[ew] = n_scored; }
  __syncthreads();
  if (tid == 0) {
    BlockBest b = s_bb[0];
    int cnt = s_cnt[0];
    for (int w = 1; w < GEPI_THREADS / 32; ++w) {
      cnt += s_cnt[w];
      if (s_bb[w].local >= 0 && better(s_bb[w].score, s_bb[w].id, b.score, b.id)) b = s_bb[w];
    }
    P.partials[blockIdx.x] = b;
    (void)cnt;      // the scored count is not accumulated here: every lattice point is updated each step
    __threadfence();
    const unsigned ticket = atomicAdd(&st->blocks_done, 1u);
    s_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  BlockBest b;
  b.score = NEG_INF; b.id = 0x7fffffffffffffffLL; b.local = -1;
  for (int i = tid; i < (int)gridDim.x; i += GEPI_THREADS) {
    const volatile BlockBest* pp = &P.partials[i];
    const double ps = pp->score;
    const long long pid = pp->id, ploc = pp->local;
    if (ploc >= 0 && better(ps, pid, b.score, b.id)) { b.score = ps; b.id = pid; b.local = ploc; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double os = __shfl_xor_sync(0xffffffffu, b.score, o);
    const long long oid = __shfl_xor_sync(0xffffffffu, b.id, o);
    const long long ol = __shfl_xor_sync(0xffffffffu, b.local, o);
    if (ol >= 0 && better(os, oid, b.score, b.id)) { b.score = os; b.id = oid; b.local = ol; }
  }
  __syncthreads();
  if (lane == 0) s_bb[ew] = b;
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < GEPI_THREADS / 32; ++w)
      if (s_bb[w].local >= 0 && better(s_bb[w].score, s_bb[w].id, b.score, b.id)) b = s_bb[w];
    s_bb[0] = b;
    const int it = st->iter;
    if (it <= P.M_cap) {
      P.hist[it] = r0;
      P.hist[(P.M_cap + 1) + it] = T;
      P.hist[2 * (P.M_cap + 1) + it] = (int)min(P.N, (long long)0x7fffffff);
    }
    st->n_scored = 0;
    st->blocks_done = 0;
    st->iter = it + 1;
  }
  __syncthreads();
  b = s_bb[0];
  write_record(P, b.score, b.local, 0);
  publish_record(P, st, st->iter, 0);
}

}  // namespace gpis
