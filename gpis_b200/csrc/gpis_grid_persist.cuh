// Grid backend, ONE persistent cooperative kernel per selection (function-only models, R_cap <= 4096).
//
// Replaces the per-step launch chain k_grid_solve2 -> k_grid_gemm -> k_grid_epilogue (3 x 3999
// dependent launches at BASELINE config 3) and, in the reference, the host loop of
// gpu_active_set_selector.cpp:499-560 (kernel vector, TRSM, GEMV, norm, arg-max, full re-factorisation
// per step).  One CTA per SM stays resident for the whole greedy loop of
// select_active_set_straddle.m:62-126; the phases of a step are separated by in-kernel grid
// barriers (cooperative launch: all CTAs are co-resident by construction):
//
//   A  wait for the exchange records of this generation (local flags; peers store them over NVLink),
//      every CTA picks the same winner
//   B  factor append, W pass: rows of W = L^-1 stream through a shared-memory ring by bulk async
//      copies; three rows per batch: l[s] = W[s,:] k (ONE block reduction per batch, not per row) and,
//      while the rows are still in registers, Y += l[s] W[s,:] into per-thread column accumulators.
//      The last CTA also writes the winner's axis-table rows and the bookkeeping.         -- barrier
//   C  d, g_new and the new row of W = [-Y/d, 1/d] from the per-CTA partial Y's (fixed order)  -- barrier
//   D  update GEMM v[jy,jx] = sum_rho (W[r,rho] Tz[rho][jz]) Ty[rho][jy] Tx[rho][jx] on the fp64 tensor
//      pipe (DMMA.8x8x4), stream-K over all CTAs, with the epilogue FUSED: once the segments of a tile
//      have arrived (per-tile counter), its contributors share the tile's epilogue -- sum of the
//      split-K partials in segment order, posterior update, straddle score (straddle.m:106-108),
//      classification (:122-125), arg-max.  CTA 0 reduces the per-CTA winners and publishes the next
//      record (which is also what releases the other CTAs into the next step).
//
// mbarriers and the pipeline are set up inside the resident kernel; nothing returns to the host
// until the selection has stopped.
#pragma once
#include "gpis_grid.cuh"

namespace gpis {

// lap timer of CTA 0 (globaltimer ns): each stamp books the time since the previous stamp
enum { PCLK_RECWAIT = 0, PCLK_WPASS = 1, PCLK_BAR1 = 2, PCLK_FINISH = 3, PCLK_BAR2 = 4, PCLK_GEMM = 5, PCLK_EPI = 6,
       PCLK_ARRIVE = 7, PCLK_PUBLISH = 8, PCLK_STEPS = 9, PCLK_N = 12 };

#ifndef GPIS_EMULATE

constexpr int PB = 3;                                   // rows of W per batch in phase B (3 x 16 + 16 doubles per thread)
constexpr int PRING_SLOTS_MAX = 24;
constexpr int PRING_DOUBLES = 6 * GS1_RMAX;             // 192 KB ring (aliases the GEMM stages of phase D)
constexpr int PCW = GEMM_CWARPS * 32;                   // 256 compute threads (warps 0-7); warp 8 produces
constexpr size_t PERSIST_SMEM =
    sizeof(double) * ((size_t)GS1_RMAX + PRING_DOUBLES) + 8 * (2 * PRING_SLOTS_MAX + 2 * GPIPE);
static_assert((size_t)GPIPE * GSTAGE_DOUBLES <= (size_t)PRING_DOUBLES, "GEMM stages alias the ring");
static_assert(GS1_CPT * PCW == GS1_RMAX, "column ownership of the W pass");

#define PSTAMP(slot)                                                       \
  do {                                                                     \
    if (cta == 0 && tid == 0) {                                            \
      const unsigned long long _now = gtimer_ns();                         \
      s_clk[slot] += _now - s_tk[0];                                       \
      s_tk[0] = _now;                                                      \
    }                                                                      \
  } while (0)

__device__ __forceinline__ unsigned long long gtimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbar_inval(unsigned long long* bar) {
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}
// generic-proxy writes to global memory (rows of W, axis tables) are later read by bulk async copies
// issued from other CTAs: order them across the proxies on both sides of the grid barrier
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;\n" ::: "memory"); }

__device__ __forceinline__ void persist_fail(DevState* st) {
  atomicCAS(&st->err, 0, -7);
  *(volatile int*)&st->done = 1;
}

// spin until *p >= target; false on timeout or when another CTA has raised an error
__device__ __forceinline__ bool spin_ge(const unsigned* p, unsigned target, const EngineParams& P, DevState* st) {
  const volatile unsigned* vp = p;
  if (*vp >= target) return true;
  const long long t0 = clock64();
  unsigned n = 0;
  while (*vp < target) {
    if ((++n & 255u) == 0u) {
      if (*(volatile int*)&st->err != 0) return false;
      if (clock64() - t0 > P.spin_cycles) return false;
    }
  }
  return true;
}

// grid-wide barrier on a monotonic counter (all CTAs resident: cooperative launch); `passed` counts
// the barriers this CTA has gone through.  Returns false when the wait was abandoned.
__device__ __forceinline__ bool grid_barrier(const EngineParams& P, DevState* st, unsigned& passed, int* s_flag) {
  __syncthreads();
  ++passed;
  if (threadIdx.x == 0) {
    fence_proxy_async_all();
    __threadfence();
    atomicAdd(&st->grid_bar, 1u);
    const bool ok = spin_ge(&st->grid_bar, passed * gridDim.x, P, st);
    __threadfence();
    fence_proxy_async_all();
    *s_flag = ok ? 1 : 0;
  }
  __syncthreads();
  return *s_flag != 0;
}

template <int D>
__global__ void __launch_bounds__(GEMM_THREADS, 1) k_grid_persist(EngineParams P, int iters) {
  GPIS_DYN_SMEM_F64A(p_smem);
  double* s_kv = p_smem;                                  // [GS1_RMAX] covariance of the winner with the model rows
  double* ring = p_smem + GS1_RMAX;                       // phase B: rows of W; phase D: GEMM stages
  unsigned long long* rfull = reinterpret_cast<unsigned long long*>(ring + PRING_DOUBLES);
  unsigned long long* rempty = rfull + PRING_SLOTS_MAX;
  unsigned long long* gfull = rempty + PRING_SLOTS_MAX;
  unsigned long long* gempty = gfull + GPIPE;
  __shared__ double s_red[2][PB][GEMM_CWARPS];
  __shared__ double s_rec[PAY_HDR];
  __shared__ double s_gram[GEMM_THREADS / 32][2];
  __shared__ double s_fin[3];                             // dinv, g_new, ok
  __shared__ BlockBest s_bb[GEMM_CWARPS];
  __shared__ int s_win, s_flag;
  __shared__ unsigned long long s_clk[PCLK_N], s_tk[1];      // CTA 0, thread 0 only: lap timer (globaltimer ns)

  DevState* st = P.st;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cta = blockIdx.x, G = gridDim.x;
  const double NEG_INF = -__longlong_as_double(0x7ff0000000000000LL);
  if (*(volatile int*)&st->done) return;
  // a function-only selection started by k_begin: rows == points == generation, so ONE counter (r0) carries
  // the loop state and the barrier count (3 per step) is derived from it
  int r0 = 0;
  if (*(volatile int*)&st->r != 0 || *(volatile int*)&st->m != 0 || *(volatile int*)&st->iter != 0) return;
  for (int i = tid; i < GS1_RMAX; i += GEMM_THREADS) s_kv[i] = 0.0;
  if (tid == 0) {
    for (int i = 0; i < PRING_SLOTS_MAX; ++i) { mbar_init(&rfull[i], 1); mbar_init(&rempty[i], GEMM_CWARPS); }
    for (int i = 0; i < GPIPE; ++i) { mbar_init(&gfull[i], 1); mbar_init(&gempty[i], GEMM_CWARPS); }
    mbar_init_fence();
  }
  __syncthreads();

  const int nx = P.glen[0], ny = P.glen[1];
  const int tiles_x = (nx + GT - 1) / GT, tiles_y = (ny + GT - 1) / GT, txy = tiles_x * tiles_y;
  const int T = txy * P.glen[2];
  if (tid < PCLK_N) s_clk[tid] = 0;
  if (tid == 0) s_tk[0] = gtimer_ns();
  int stop = 0;          // 1: nothing unclassified / capacity (done), 2: not positive definite

  for (; r0 < iters;) {
    const int m = r0, gen = r0;
    unsigned passed = 3u * (unsigned)r0;
    // ------------------------------------------------------------------ A: records, winner
    if (!wait_records(P, st, gen)) return;
    PSTAMP(PCLK_RECWAIT);
    const double* recs = records_ptr(P, gen);
    if (tid == 0) {
      int win = -1;
      double bs = 0.0;
      long long bid = 0;
      for (int w = 0; w < P.world; ++w) {
        const double* pay = recs + (long long)w * P.pay_stride;
        const long long loc = __double_as_longlong(__ldcg(pay + PAY_LOCAL));
        if (loc < 0) continue;
        const double sc = __ldcg(pay + PAY_SCORE);
        const long long id = __double_as_longlong(__ldcg(pay + PAY_ID));
        if (win < 0 || better(sc, id, bs, bid)) { win = w; bs = sc; bid = id; }
      }
      if (win >= 0 && m >= P.M_cap) win = -1;
      s_win = win;
    }
    __syncthreads();
    const int win = s_win;
    if (win < 0) { stop = 1; break; }
    if (tid < PAY_HDR) s_rec[tid] = __ldcg(recs + (long long)win * P.pay_stride + tid);
    __syncthreads();

    // ------------------------------------------------------------------ B: W pass
    // rows s = cta + i*G of W; slot stride fits the longest row of this step
    const int my_rows = cta < r0 ? (r0 - 1 - cta) / G + 1 : 0;
    const int stride = (r0 + 2) & ~1;                                   // doubles, 16-byte multiple
    const int nslots = min(PRING_SLOTS_MAX, PRING_DOUBLES / (stride > 0 ? stride : 2));
    if (tid == 0) {      // fresh phase bits every step (the slot count changes with the row length)
      for (int i = 0; i < PRING_SLOTS_MAX; ++i) {
        mbar_inval(&rfull[i]); mbar_inval(&rempty[i]);
        mbar_init(&rfull[i], 1); mbar_init(&rempty[i], GEMM_CWARPS);
      }
      mbar_init_fence();
    }
    __syncthreads();
    double nxp[D];
#pragma unroll
    for (int d = 0; d < D; ++d) nxp[d] = s_rec[PAY_X + d];
    if (warp == GEMM_CWARPS) {
      // producer warp: one bulk copy per row
      if (lane == 0) {
        for (int i = 0; i < my_rows; ++i) {
          const int slot = i % nslots, use = i / nslots;
          if (use > 0) mbar_wait(&rempty[slot], (unsigned)((use - 1) & 1));
          const int row = cta + i * G;
          const unsigned bytes = (unsigned)(((row + 2) >> 1) << 4);    // row + 1 doubles, rounded up to 16 B
          mbar_arrive_expect_tx(&rfull[slot], bytes);
          bulk_g2s(ring + (size_t)slot * stride, P.W + (long long)row * P.ldW, bytes, &rfull[slot]);
        }
      }
      // bookkeeping + the winner's axis-table rows: last CTA, this warp (the others stream rows)
      if (cta == G - 1) {
        if (lane == 0) {
          const long long loc = __double_as_longlong(s_rec[PAY_LOCAL]);
          if (win == P.rank) {
            unsigned char* f = P.flags + grid_slot_of_local(P, loc);
            *f = (unsigned char)(__ldcg(f) | F_ACTIVE);
          }
          P.sel_ids[m] = __double_as_longlong(s_rec[PAY_ID]);
#pragma unroll
          for (int d = 0; d < D; ++d) P.AX[d][m] = nxp[d];
        }
        for (int d = 0; d < D; ++d) {
          const int len = P.gtlen[d];
          for (int j = lane; j < len; j += 32) {
            const double diff = (P.gorg[d] + P.gh[d] * (double)j) - nxp[d];
            double v = exp(-0.5 * diff * diff * P.inv_ell2);
            if (d == 0) v *= P.sf2;
            P.T[d][(long long)r0 * P.ldt[d] + j] = v;
          }
        }
      }
    } else {
      // covariance of the winner with the model rows (entries >= m stay zero)
      for (int i = tid; i < m; i += PCW) {
        double d2 = 0.0;
#pragma unroll
        for (int d = 0; d < D; ++d) { const double df = __ldcg(P.AX[d] + i) - nxp[d]; d2 += df * df; }
        s_kv[i] = P.sf2 * exp(-0.5 * d2 * P.inv_ell2);
      }
      named_sync(1, PCW);
      double acc[GS1_CPT];
#pragma unroll
      for (int c = 0; c < GS1_CPT; ++c) acc[c] = 0.0;
      int par = 0;
      for (int i0 = 0; i0 < my_rows; i0 += PB) {
        double w[PB][GS1_CPT];
        double p[PB];
#pragma unroll
        for (int b = 0; b < PB; ++b) {
          p[b] = 0.0;
          const int i = i0 + b;
          if (i < my_rows) {
            const int slot = i % nslots, row = cta + i * G;
            mbar_wait(&rfull[slot], (unsigned)((i / nslots) & 1));
            const double* buf = ring + (size_t)slot * stride;
#pragma unroll
            for (int c = 0; c < GS1_CPT; ++c) {
              const int col = tid + PCW * c;
              w[b][c] = col <= row ? buf[col] : 0.0;
            }
          } else {
#pragma unroll
            for (int c = 0; c < GS1_CPT; ++c) w[b][c] = 0.0;
          }
        }
        __syncwarp();
        if (lane == 0) {
#pragma unroll
          for (int b = 0; b < PB; ++b)
            if (i0 + b < my_rows) mbar_arrive(&rempty[(i0 + b) % nslots]);
        }
#pragma unroll
        for (int b = 0; b < PB; ++b) {
#pragma unroll
          for (int c = 0; c < GS1_CPT; ++c) p[b] += w[b][c] * s_kv[tid + PCW * c];
          for (int o = 16; o > 0; o >>= 1) p[b] += __shfl_xor_sync(0xffffffffu, p[b], o);
        }
        if (lane == 0) {
#pragma unroll
          for (int b = 0; b < PB; ++b) s_red[par][b][warp] = p[b];
        }
        named_sync(1, PCW);
        double ell[PB];
#pragma unroll
        for (int b = 0; b < PB; ++b) {
          double e = 0.0;
#pragma unroll
          for (int w8 = 0; w8 < GEMM_CWARPS; ++w8) e += s_red[par][b][w8];
          ell[b] = e;
        }
#pragma unroll
        for (int b = 0; b < PB; ++b)
          if (tid == b && i0 + b < my_rows) P.L[(long long)r0 * P.R_cap + cta + (i0 + b) * G] = ell[b];
        par ^= 1;
#pragma unroll
        for (int b = 0; b < PB; ++b)
#pragma unroll
          for (int c = 0; c < GS1_CPT; ++c) acc[c] += ell[b] * w[b][c];
      }
      if (my_rows > 0) {
        double* yp = P.ypart + (long long)cta * GS1_RMAX + tid;
#pragma unroll
        for (int c = 0; c < GS1_CPT; ++c)
          if (tid + PCW * c < r0) yp[PCW * c] = acc[c];
      }
    }
    PSTAMP(PCLK_WPASS);
    if (!grid_barrier(P, st, passed, &s_flag)) { if (tid == 0) persist_fail(st); return; }
    PSTAMP(PCLK_BAR1);

    // ------------------------------------------------------------------ C: d, g_new, new row of W
    {
      const double* X = P.L + (long long)r0 * P.R_cap;
      double p0 = 0.0, p1 = 0.0;
      for (int s = tid; s < r0; s += GEMM_THREADS) { const double x = __ldcg(X + s); p0 += x * x; p1 += x * __ldcg(P.g + s); }
      for (int o = 16; o > 0; o >>= 1) { p0 += __shfl_xor_sync(0xffffffffu, p0, o); p1 += __shfl_xor_sync(0xffffffffu, p1, o); }
      if (lane == 0) { s_gram[warp][0] = p0; s_gram[warp][1] = p1; }
      __syncthreads();
      if (tid == 0) {
        double xx = 0.0, xg = 0.0;
        for (int w = 0; w < GEMM_THREADS / 32; ++w) { xx += s_gram[w][0]; xg += s_gram[w][1]; }
        double mf = P.mean_c;
#pragma unroll
        for (int d = 0; d < D; ++d) mf += P.mean_w[d] * s_rec[PAY_X + d];
        const double resid = s_rec[PAY_TSDF] - mf;
        double sdiag = P.sf2 + s_rec[PAY_NOISE] - xx;
        const bool ok = sdiag > 0.0;
        if (!ok) sdiag = 1.0;
        const double d = sqrt(sdiag), dinv = 1.0 / d, gn = (resid - xg) / d;
        s_fin[0] = dinv;
        s_fin[1] = gn;
        s_fin[2] = ok ? 1.0 : 0.0;
        if (cta == 0) {
          P.g[r0] = gn;
          P.L[(long long)r0 * P.R_cap + r0] = d;
          P.W[(long long)r0 * P.ldW + r0] = dinv;
        }
      }
      __syncthreads();
      const double dinv = s_fin[0];
      const int nparts = min(G, r0);
      const int gw = (cta * GEMM_THREADS + tid) >> 5, nw = (G * GEMM_THREADS) >> 5;
      const int sub = lane >> 2, cq = lane & 3;
      for (int c0 = gw * 4; c0 < r0; c0 += nw * 4) {
        const int c = c0 + cq;
        double y = 0.0;
        if (c < r0) {
          const double* yp = P.ypart + c;
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
    if (s_fin[2] == 0.0) { stop = 2; break; }          // every CTA computes the same pivot: uniform exit
    PSTAMP(PCLK_FINISH);
    if (!grid_barrier(P, st, passed, &s_flag)) { if (tid == 0) persist_fail(st); return; }
    PSTAMP(PCLK_BAR2);

    // ------------------------------------------------------------------ D: update GEMM + fused epilogue
    const int C = (r0 + 1 + GKC - 1) / GKC;              // chunks per tile
    const long long U = (long long)T * C;
    const int Gw = (int)min((long long)G, U);
    const bool working = cta < Gw;
    const long long u0 = working ? (long long)cta * U / Gw : 0, u1 = working ? (long long)(cta + 1) * U / Gw : 0;
    const int n_units = (int)(u1 - u0), tile0 = (int)(u0 / C), c0 = (int)(u0 % C);
    unsigned* tcnt = P.tile_cnt + (size_t)(gen & 1) * T;
    unsigned* tcnt_next = P.tile_cnt + (size_t)((gen + 1) & 1) * T;
    if (tid == 0) {
      for (int i = 0; i < GPIPE; ++i) {
        mbar_inval(&gfull[i]); mbar_inval(&gempty[i]);
        mbar_init(&gfull[i], 1); mbar_init(&gempty[i], GEMM_CWARPS);
      }
      mbar_init_fence();
    }
    __syncthreads();
    if (warp == GEMM_CWARPS) {
      // ------------------------------- producer -------------------------------
      int stage = 0;
      unsigned phase = 0;
      const double* wrow = P.W + (long long)r0 * P.ldW;
      const int klen = r0 + 1;
      int tile = tile0, c = c0;
      int tz = tile / txy, ty = (tile / tiles_x) % tiles_y, tx = tile % tiles_x;
      for (int ul = 0; ul < n_units; ++ul) {
        mbar_wait(&gempty[stage], phase ^ 1);
        double* As = ring + (size_t)stage * GSTAGE_DOUBLES;
        double* Bs = As + GKC * GLD;
        double* cs = Bs + GKC * GLD;
        const int rho0 = c * GKC;
        if (lane < GKC) {
          const int rho = rho0 + lane;
          double cv = rho < klen ? __ldcg(wrow + rho) : 0.0;
          cv *= __ldcg(P.T[2] + (long long)rho * P.ldt[2] + P.gz0 + tz);
          cs[lane] = cv;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_expect_tx(&gfull[stage], 2 * GKC * GT * (unsigned)sizeof(double));
        __syncwarp();
        if (lane < GKC)
          bulk_g2s(As + lane * GLD, P.T[1] + (long long)(rho0 + lane) * P.ldt[1] + ty * GT, GT * sizeof(double), &gfull[stage]);
        else
          bulk_g2s(Bs + (lane - GKC) * GLD, P.T[0] + (long long)(rho0 + lane - GKC) * P.ldt[0] + tx * GT, GT * sizeof(double), &gfull[stage]);
        if (++c == C) {
          c = 0;
          ++tile;
          if (++tx == tiles_x) { tx = 0; if (++ty == tiles_y) { ty = 0; ++tz; } }
        }
        if (++stage == GPIPE) { stage = 0; phase ^= 1; }
      }
    } else {
      // --------------------------------- consumers ---------------------------------
      int stage = 0;
      unsigned phase = 0;
      const int wm = warp >> 1, wn = warp & 1;
      int ul = 0, tile = tile0, c = c0;
      while (ul < n_units) {
        const int seg_units = min(n_units - ul, C - c);
        double acc[4][8][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        for (int q = 0; q < seg_units; ++q) {
          mbar_wait(&gfull[stage], phase);
          const double* As = ring + (size_t)stage * GSTAGE_DOUBLES;
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
          if (lane == 0) mbar_arrive(&gempty[stage]);
          if (++stage == GPIPE) { stage = 0; phase ^= 1; }
        }
        ul += seg_units;
        // partial accumulators of this (tile, segment) -> workspace (fragment order), then arrive on the tile
        const int seg = cta - unit_owner((long long)tile * C, U, Gw);
        double2* ws = reinterpret_cast<double2*>(P.gws + ((long long)tile * P.gws_segmax + seg) * GTILE_ELEMS);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j)
            __stcg(&ws[(warp * 32 + i * 8 + j) * 32 + lane], make_double2(acc[i][j][0], acc[i][j][1]));
        __threadfence();
        named_sync(1, PCW);
        if (tid == 0) {
          if (seg == 0) *(volatile unsigned*)&tcnt_next[tile] = 0u;      // reset the other parity for the next step
          atomicAdd(&tcnt[tile], 1u);
        }
        ++tile;
        c = 0;
      }
      PSTAMP(PCLK_GEMM);

      // ---- fused epilogue: the contributors of a tile share it, 256-slot units split by segment index
      const double sb = [&]() {
        double beta = P.beta;
        if (P.beta_mode == 1) beta = 2.0 * log((double)P.N_total * (double)(gen + 1) * 9.869604401089358 / (6.0 * P.beta_delta));
        else if (P.beta_mode == 2) { const double k1 = (double)(gen + 1); beta = 2.0 * log((double)P.N_total * 9.869604401089358 * k1 * k1 / (6.0 * P.beta_delta)); }
        return sqrt(beta);
      }();
      const double gn1 = s_fin[1];
      double best_s = NEG_INF;
      long long best_id = 0x7fffffffffffffffLL, best_j = -1;
      bool ok = true;
      const int tile_last = n_units > 0 ? (int)((u1 - 1) / C) : tile0 - 1;
      for (int tile = tile0; tile <= tile_last && ok; ++tile) {
        const int g_first = unit_owner((long long)tile * C, U, Gw);
        const int nseg = unit_owner((long long)(tile + 1) * C - 1, U, Gw) - g_first + 1;
        const int seg = cta - g_first;
        if (tid == 0) s_flag = spin_ge(&tcnt[tile], (unsigned)nseg, P, st) ? 1 : 0;
        named_sync(1, PCW);
        if (!s_flag) { ok = false; break; }
        __threadfence();
        constexpr int UNITS = GTILE_ELEMS / 2 / PCW;                   // 32 units of 256 fragment slots
        const int un0 = seg * UNITS / nseg, un1 = (seg + 1) * UNITS / nseg;
        const double2* wsb = reinterpret_cast<const double2*>(P.gws + (long long)tile * P.gws_segmax * GTILE_ELEMS);
        for (int ub = un0; ub < un1; ub += GEPI_SPT) {
          double2 v[GEPI_SPT], mu[GEPI_SPT], var[GEPI_SPT];
          uchar2 fl[GEPI_SPT];
#pragma unroll
          for (int k = 0; k < GEPI_SPT; ++k) {
            const int q = (ub + k < un1 ? ub + k : un1 - 1) * PCW + tid;
            v[k] = __ldcg(wsb + q);
          }
          for (int sgi = 1; sgi < nseg; ++sgi) {
            double2 pz[GEPI_SPT];
#pragma unroll
            for (int k = 0; k < GEPI_SPT; ++k) {
              const int q = (ub + k < un1 ? ub + k : un1 - 1) * PCW + tid;
              pz[k] = __ldcg(wsb + (long long)sgi * (GTILE_ELEMS / 2) + q);
            }
#pragma unroll
            for (int k = 0; k < GEPI_SPT; ++k) { v[k].x += pz[k].x; v[k].y += pz[k].y; }
          }
#pragma unroll
          for (int k = 0; k < GEPI_SPT; ++k) {
            const long long slot2 = (long long)tile * (GTILE_ELEMS / 2) + (ub + k < un1 ? ub + k : un1 - 1) * PCW + tid;
            mu[k] = __ldcg(reinterpret_cast<const double2*>(P.mu) + slot2);
            var[k] = __ldcg(reinterpret_cast<const double2*>(P.var) + slot2);
            fl[k] = __ldcg(reinterpret_cast<const uchar2*>(P.flags) + slot2);
          }
#pragma unroll
          for (int k = 0; k < GEPI_SPT; ++k) {
            if (ub + k >= un1) continue;
            const long long slot2 = (long long)tile * (GTILE_ELEMS / 2) + (ub + k) * PCW + tid;
            mu[k].x += v[k].x * gn1; mu[k].y += v[k].y * gn1;
            var[k].x -= v[k].x * v[k].x; var[k].y -= v[k].y * v[k].y;
            __stcg(reinterpret_cast<double2*>(P.mu) + slot2, mu[k]);
            __stcg(reinterpret_cast<double2*>(P.var) + slot2, var[k]);
            unsigned char f[2] = {fl[k].x, fl[k].y};
            if ((f[0] & (F_ACTIVE | F_HIGH | F_LOW)) && (f[1] & (F_ACTIVE | F_HIGH | F_LOW))) continue;
            const double mus[2] = {mu[k].x, mu[k].y}, vars[2] = {var[k].x, var[k].y};
            int jx, jy, tz;
            grid_unslot(P, slot2 * 2, jx, jy, tz);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              if (f[e] & (F_ACTIVE | F_HIGH | F_LOW)) continue;
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
            if (f[0] != fl[k].x || f[1] != fl[k].y)
              __stcg(reinterpret_cast<uchar2*>(P.flags) + slot2, make_uchar2(f[0], f[1]));
          }
        }
      }
      if (!ok && tid == 0) persist_fail(st);
      // CTA winner
      for (int o = 16; o > 0; o >>= 1) {
        const double os = __shfl_xor_sync(0xffffffffu, best_s, o);
        const long long oid = __shfl_xor_sync(0xffffffffu, best_id, o);
        const long long oj = __shfl_xor_sync(0xffffffffu, best_j, o);
        if (oj >= 0 && better(os, oid, best_s, best_id)) { best_s = os; best_id = oid; best_j = oj; }
      }
      if (lane == 0) { s_bb[warp].score = best_s; s_bb[warp].id = best_id; s_bb[warp].local = best_j; }
      named_sync(1, PCW);
      if (tid == 0) {
        BlockBest b = s_bb[0];
        for (int w = 1; w < GEMM_CWARPS; ++w)
          if (s_bb[w].local >= 0 && (b.local < 0 || better(s_bb[w].score, s_bb[w].id, b.score, b.id))) b = s_bb[w];
        BlockBest* dst = &P.partials[cta];
        __stcg(&dst->score, b.score);
        __stcg(&dst->id, b.id);
        __stcg(&dst->local, b.local);
      }
      PSTAMP(PCLK_EPI);
    }
    // ---- arrive; CTA 0 waits for everyone, reduces the CTA winners and publishes the next record
    __syncthreads();
    if (tid == 0) s_win = *(volatile int*)&st->err;      // one reader: the exit must be uniform
    __syncthreads();
    if (s_win != 0) return;
    ++passed;
    if (tid == 0) {
      __threadfence();
      atomicAdd(&st->grid_bar, 1u);
    }
    r0 += 1;
    if (cta == 0) {
      if (tid == 0) {
        s_flag = spin_ge(&st->grid_bar, passed * G, P, st) ? 1 : 0;
        __threadfence();
      }
      __syncthreads();
      if (!s_flag) { if (tid == 0) persist_fail(st); return; }
      PSTAMP(PCLK_ARRIVE);
      BlockBest b;
      b.score = NEG_INF; b.id = 0x7fffffffffffffffLL; b.local = -1;
      if (warp == 0) {
        for (int i = lane; i < G; i += 32) {
          const BlockBest* pp = &P.partials[i];
          const double ps = __ldcg(&pp->score);
          const long long pid = __ldcg(&pp->id), ploc = __ldcg(&pp->local);
          if (ploc >= 0 && better(ps, pid, b.score, b.id)) { b.score = ps; b.id = pid; b.local = ploc; }
        }
        for (int o = 16; o > 0; o >>= 1) {
          const double os = __shfl_xor_sync(0xffffffffu, b.score, o);
          const long long oid = __shfl_xor_sync(0xffffffffu, b.id, o);
          const long long ol = __shfl_xor_sync(0xffffffffu, b.local, o);
          if (ol >= 0 && better(os, oid, b.score, b.id)) { b.score = os; b.id = oid; b.local = ol; }
        }
        if (lane == 0) s_bb[0] = b;
      }
      __syncthreads();
      b = s_bb[0];
      if (tid == 0) {
        const int itx = r0 - 1;
        if (itx <= P.M_cap) {
          P.hist[itx] = itx;
          P.hist[(P.M_cap + 1) + itx] = T;
          P.hist[2 * (P.M_cap + 1) + itx] = (int)min(P.N, (long long)0x7fffffff);
        }
        // state the host (and the finalize kernel) reads; iter tags the record published next
        st->r = r0; st->r0 = r0 - 1; st->m = r0; st->n_sel = r0; st->iter = r0; st->gnew[0] = s_fin[1];
        st->nnoise = s_rec[PAY_NOISE];
#pragma unroll
        for (int d = 0; d < D; ++d) st->nx[d] = s_rec[PAY_X + d];
      }
      __syncthreads();
      write_record(P, b.score, b.local, 0);
      publish_record(P, st, r0, 0);
      PSTAMP(PCLK_PUBLISH);
      if (tid == 0) s_clk[PCLK_STEPS] += 1;
    }
  }
  if (cta == 0 && tid == 0) {
    if (stop == 1) st->done = 1;
    if (stop == 2) { st->err = -5; st->done = 1; }
    if (P.phase_clk)
      for (int i = 0; i < PCLK_N; ++i) P.phase_clk[i] = s_clk[i];
  }
}

#endif  // GPIS_EMULATE

}  // namespace gpis
