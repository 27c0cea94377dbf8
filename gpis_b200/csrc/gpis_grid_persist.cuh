// Grid backend, ONE persistent cooperative kernel per selection (function-only models, R_cap <= 4096).
//
// Replaces the per-step launch chain k_grid_solve2 -> k_grid_gemm -> k_grid_epilogue (3 x 3999
// dependent launches at BASELINE config 3) and, in the reference, the host loop of
// gpu_active_set_selector.cpp:499-560 (kernel vector, TRSM, GEMV, norm, arg-max, full re-factorisation
// per step).  One CTA per SM stays resident for the whole greedy loop of
// select_active_set_straddle.m:62-126; the phases of a step are separated by in-kernel grid
// barriers (cooperative launch: all CTAs are co-resident by construction):
//
//   A  the winner of the previous step: on one GPU every CTA reduces the per-CTA winners itself (no
//      record, no flag); sharded, every CTA waits for the ranks' exchange records (local flags; peers
//      store them over NVLink) and picks the same winner
//   B  factor append, W pass: rows of W = L^-1 stream through a shared-memory ring by bulk async
//      copies; three rows per batch: l[s] = W[s,:] k (ONE block reduction per batch, not per row) and,
//      while the rows are still in registers, Y += l[s] W[s,:] into per-thread column accumulators.
//      The last CTA also writes the winner's axis-table rows and the bookkeeping; every CTA appends
//      the new model row to the row lists of the lattice tiles it can reach.          -- barrier
//   C  d, g_new and the new row of W = [-Y/d, 1/d] from the per-CTA partial Y's (fixed order)  -- barrier
//   D  update GEMM v[jy,jx] = sum_rho (W[r,rho] Tz[rho][jz]) Ty[rho][jy] Tx[rho][jx] on the fp64 tensor
//      pipe (DMMA.8x8x4) over 64 x 64 lattice tiles.  A tile only sums the model rows of ITS list: the
//      rows whose SE factor over the tile's box is >= 2^-70 -- the dropped terms are below 2^-70 |W| each,
//      far under one ulp of the sum, so the result is the double-precision sum of all rows.  Stream-K over
//      the (tile, 32-row chunk) units of all lists; a tile that lies inside one CTA's range (most do)
//      goes from the accumulator registers straight into the epilogue (posterior update, straddle score
//      straddle.m:106-108, classification :122-125, arg-max) while the producer warp already fills the
//      ring with the next tile's rows; a tile cut by a range boundary goes through a per-CTA workspace and its contributors
//      share its epilogue.                                                            -- barrier
//
// mbarriers and the pipelines are set up once in the resident kernel; nothing returns to the host
// until the selection has stopped.
#pragma once
#include "gpis_grid.cuh"

namespace gpis {

// lap timer of CTA 0 (globaltimer ns): each stamp books the time since the previous stamp
enum { PCLK_RECWAIT = 0, PCLK_WSETUP = 1, PCLK_WPASS = 2, PCLK_BAR1 = 3, PCLK_FINISH = 4, PCLK_BAR2 = 5, PCLK_PREFIX = 6,
       PCLK_GEMM = 7, PCLK_EPI = 8, PCLK_BAR3 = 9, PCLK_WINNER = 10, PCLK_STEPS = 11,
       // inside PCLK_GEMM (booked twice): coefficient precompute, partial tiles to the workspace, in-register epilogues; the
       // (tile, 32-row chunk) units of ALL blocks (a count: x 2 * 64 * 64 * 32 = the flops the MMAs executed)
       PCLK_PRECOMP = 12, PCLK_WAITFULL = 13, PCLK_EPIREG = 14, PCLK_UNITS = 15, PCLK_N = 16 };

constexpr int QT = 64;                                  // lattice tile edge (jy x jx) of the persistent kernel
constexpr int QKC = 32;                                 // model rows per pipeline stage (a "unit" = one tile x 32 rows)
constexpr int QLD = QT + 4;                             // padded smem row (conflict-free DMMA fragment loads)
constexpr int QSTAGE_DOUBLES = 2 * QKC * QLD;           // Ty rows, Tx rows
constexpr int QPIPE = 4;
// 3-D lattices: 16 x 16 x 16 tiles (same 4096 points, same accumulator fragments).  A model row reaches far fewer
// cubes than 64 x 64 x 1 slabs (config 3: 15.0 % of the dense (tile, row) pairs instead of 27.6 %), the contraction is
// [(jz, jy) = 256 x rows] x [rows x (jx) = 16] with A = (W[r,rho] Tz[rho][jz]) Ty[rho][jy] formed in registers.
constexpr int QC = 16;                                  // cube edge
constexpr int QLDC = QC + 4;                            // padded smem row of a cube stage (conflict-free fragment loads)
constexpr int QSTAGE_CUBE = 3 * QKC * QLDC;             // Ty rows, Tx rows, Tz rows (18 doubles copied: even-aligned start)
#ifndef GPIS_QPIPE_CUBE
#define GPIS_QPIPE_CUBE 6
#endif
constexpr int QPIPE_CUBE = GPIS_QPIPE_CUBE;
constexpr int QPIPE_MAX = 10;
constexpr int QTILE_ELEMS = QT * QT;
static_assert(QC * QC * QC == QT * QT, "both tile shapes hold the same number of lattice points");
constexpr int QPAIRS = QTILE_ELEMS / 2;                 // accumulator fragment slots (x-adjacent pairs) per tile
constexpr int QT_MAX = 2048;                            // most tiles per rank the kernel takes (prefix table in smem)
// exp(-d^2 / (2 ell^2)) < 2^-70  <=>  d^2 / (2 ell^2) > 70 ln 2
constexpr double QCUT = 48.52030263919617;
// Stream-K splits the COST of a step evenly: a tile costs its chunks plus QEPI unit times for its epilogue (a chunk is
// 1.04 us of MMAs, an epilogue about 2 us), so blocks with many short tiles are not the last to arrive.
#ifndef GPIS_PROD_EPI
#define GPIS_PROD_EPI 1      // cubes: the producer warps run the epilogues of whole tiles (0: the MMA warps, from registers)
#endif
#ifndef GPIS_QEPI
#define GPIS_QEPI 2
#endif
constexpr int QEPI = GPIS_QEPI;
constexpr int QBLK = 128;                               // units whose coefficients sit in shared memory at a time

// lattice index of the first point of fragment slot q (0 .. QPAIRS-1) of the tile at (tx, ty, tz)
template <bool CUBE>
__host__ __device__ inline long long persist_pair_index(int q, int tx, int ty, int tz, int nx, int ny) {
  if (CUBE) {      // warp w holds z-slices 2w, 2w + 1 of the cube: fragment f = i * 2 + j, i = (jz & 1) * 2 + (jy >> 3), j = jx >> 3
    const int l = q & 31, j = (q >> 5) & 1, i = (q >> 6) & 3, w = q >> 8;
    const int jz = tz * QC + 2 * w + (i >> 1);
    const int jy = ty * QC + (i & 1) * 8 + (l >> 2);
    const int jx = tx * QC + j * 8 + (l & 3) * 2;
    return (long long)jx + (long long)nx * (jy + (long long)ny * jz);
  }
  const int l = q & 31, j = (q >> 5) & 3, i = (q >> 7) & 1, w = q >> 8;
  const int jy = ty * QT + (w >> 1) * 16 + i * 8 + (l >> 2);
  const int jx = tx * QT + (w & 1) * 32 + j * 8 + (l & 3) * 2;
  return (long long)jx + (long long)nx * (jy + (long long)ny * tz);
}

#ifndef GPIS_EMULATE
#ifndef GPIS_EXP
#define GPIS_EXP 0
#endif

constexpr int PB = 3;                                   // rows of W per batch in phase B (3 x 16 + 16 doubles per thread)
constexpr int PRING_SLOTS_MAX = 24;
constexpr int PRING_DOUBLES = 5 * GS1_RMAX;             // 160 KB ring (aliases the GEMM stages of phase D)
constexpr int PCW = GEMM_CWARPS * 32;                   // 256 compute threads (warps 0-7)
// Warps 8..11 produce.  A bulk copy with a per-lane address compiles to a serial elect-one loop (R2UR + UBLKCP per
// lane, ~40 cycles each), so ONE warp issuing the 32 row copies of a unit cannot keep up with 0.5 us of DMMA work
// per unit: four producer warps take every fourth unit each.  (Registers are allocated per 4 warps: 12 warps cost
// the same 168 registers per thread as 9.)
constexpr int QPROD = 4;
constexpr int PTHREADS = (GEMM_CWARPS + QPROD) * 32;
constexpr size_t PERSIST_SMEM =
    sizeof(double) * ((size_t)GS1_RMAX + PRING_DOUBLES) + 8 * (2 * PRING_SLOTS_MAX + 2 * QPIPE_MAX);
static_assert((size_t)QPIPE * QSTAGE_DOUBLES <= (size_t)PRING_DOUBLES, "GEMM stages alias the ring");
static_assert((size_t)QPIPE_CUBE * QSTAGE_CUBE + 2 * (size_t)QTILE_ELEMS <= (size_t)PRING_DOUBLES,
              "GEMM stages and the two accumulator staging tiles alias the ring");
static_assert(GS1_CPT * PCW == GS1_RMAX, "column ownership of the W pass");
static_assert(QBLK * QKC <= GS1_RMAX, "the coefficients of a block alias the covariance vector");
static_assert(QKC == 32, "a producer warp issues one Ty row and one Tx row per lane");

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
__device__ __forceinline__ bool mbar_test(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
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

// tagged-slot records of the sharded persistent loop: 8 doubles per rank
enum { PLL_SCORE = 0, PLL_ID = 1, PLL_LOCAL = 2, PLL_TSDF = 3, PLL_NOISE = 4, PLL_X = 5 };
constexpr int PLL_WORLD_MAX = 16;
__device__ __forceinline__ unsigned ll_tag(unsigned epoch, int gen) {
  return 0x80000000u | ((epoch & 0x7fu) << 24) | ((unsigned)(gen + 1) & 0xffffffu);
}

struct PBest {
  double s;
  long long id, j;
};

// best candidate of the tile a thread is working on: its score and 2 * fragment slot + element inside the tile (-1: none).
// A thread meets its cells of a tile in ascending lattice order, so a strict comparison keeps the lowest index on a tie.
struct PTileBest {
  double s;
  int q;
};

// straddle rule for one cell (straddle.m:106-108, 122-125), without branches: returns the cell's flags after the
// classification and offers its score to the tile's best when the cell is still unclassified
__device__ __forceinline__ unsigned persist_cell(double mu, double ve, unsigned f0, bool live, int qe, double sb, double level, double eps,
                                                 PTileBest& b) {
  const double w = __dmul_rn(sb, ve);                  // variance, not sigma: straddle.m:106-107
  const double hi = __dadd_rn(mu, w), lo = __dadd_rn(mu, -w);
  const double sc = fmin(__dadd_rn(hi, -level), __dadd_rn(level, -lo));
  const bool un = live && (f0 & (unsigned)(F_ACTIVE | F_HIGH | F_LOW)) == 0u;
  unsigned add = 0u;
  if (__dadd_rn(lo, eps) > level) add |= (unsigned)F_HIGH;   // :122-125
  if (__dadd_rn(hi, -eps) < level) add |= (unsigned)F_LOW;
  const bool take = un && sc > b.s;                    // (false for NaN)
  b.s = take ? sc : b.s;
  b.q = take ? qe : b.q;
  return un ? (f0 | add) : f0;
}

// Posterior update + straddle rule for NP accumulator fragment slots (two x-adjacent lattice points each) of one tile,
// branch-free per cell so that the NP independent dependency chains interleave (the branchy per-cell form was bound by
// fixed-latency waits and instruction fetch: ncu, profiles/r02_k_grid_persist_cube_ncu.csv).
// q[k]: the fragment slot inside the tile; live[k]: the slot exists (the last batch of a shared tile may be short).
// nz: predictive-variance mode only, the cells' noise (else null).
__device__ __forceinline__ void persist_pairs(const int NP, const EngineParams& P, long long tbase, const int* q, const bool* live, const double2* v,
                                              double2* mu, double2* var, const uchar2* fl, const double2* nz, double sb, double gn1,
                                              PTileBest& b) {
  const unsigned CLS = F_ACTIVE | F_HIGH | F_LOW;
  bool open = false;
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    mu[k].x += v[k].x * gn1; mu[k].y += v[k].y * gn1;
    var[k].x -= v[k].x * v[k].x; var[k].y -= v[k].y * v[k].y;
    if (live[k]) {
      __stcg(reinterpret_cast<double2*>(P.mu) + tbase + q[k], mu[k]);
      __stcg(reinterpret_cast<double2*>(P.var) + tbase + q[k], var[k]);
    }
    open = open || (live[k] && (!(fl[k].x & CLS) || !(fl[k].y & CLS)));
  }
  if (!__any_sync(0xffffffffu, open)) return;      // every cell of the warp's slots is classified: nothing to score
  const double level = P.level, eps = P.eps;
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const unsigned f0 = fl[k].x, f1 = fl[k].y;
    const unsigned g0 = persist_cell(mu[k].x, nz ? __dadd_rn(var[k].x, nz[k].x) : var[k].x, f0, live[k], 2 * q[k], sb, level, eps, b);
    const unsigned g1 = persist_cell(mu[k].y, nz ? __dadd_rn(var[k].y, nz[k].y) : var[k].y, f1, live[k], 2 * q[k] + 1, sb, level, eps, b);
    if (g0 != f0 || g1 != f1)
      __stcg(reinterpret_cast<uchar2*>(P.flags) + tbase + q[k], make_uchar2((unsigned char)g0, (unsigned char)g1));
  }
}

// predictive-variance mode: the noise of the two cells of fragment slot q (padding cells are classified: not read)
template <bool CUBE>
__device__ __forceinline__ double2 persist_cell_noise(const EngineParams& P, int q, int tx, int ty, int tz, int nx, int ny, uchar2 f) {
  if (!P.noise) return make_double2(P.noise_scalar, P.noise_scalar);
  const long long jl = persist_pair_index<CUBE>(q, tx, ty, tz, nx, ny);
  const unsigned cls = F_ACTIVE | F_HIGH | F_LOW;
  return make_double2((f.x & cls) ? 0.0 : P.noise[jl], (f.y & cls) ? 0.0 : P.noise[jl + 1]);
}

// the tile's best candidate joins the CTA's running best of the step (lowest id on a tie)
template <bool CUBE>
__device__ __forceinline__ void persist_merge_best(const EngineParams& P, const PTileBest& tb, int tx, int ty, int tz, int nx, int ny,
                                                   PBest& best) {
  if (tb.q < 0) return;
  const long long jl = persist_pair_index<CUBE>(tb.q >> 1, tx, ty, tz, nx, ny) + (tb.q & 1);
  const long long id = cand_id(P, jl);
  if (better(tb.s, id, best.s, best.id)) { best.s = tb.s; best.id = id; best.j = jl; }
}

template <int D, bool CUBE>
__global__ void __launch_bounds__(PTHREADS, 1) k_grid_persist(EngineParams P, int iters) {
  static_assert(!CUBE || D == 3, "cube tiles are for 3-D lattices");
  constexpr int NPIPE = CUBE ? QPIPE_CUBE : QPIPE;
  constexpr int TE = CUBE ? QC : QT;                      // tile edge along x and y
  constexpr int STAGE = CUBE ? QSTAGE_CUBE : QSTAGE_DOUBLES;
  GPIS_DYN_SMEM_F64A(p_smem);
  double* s_kv = p_smem;                                  // [GS1_RMAX] covariance of the winner with the model rows
  double* ring = p_smem + GS1_RMAX;                       // phase B: rows of W; phase D: GEMM stages
  unsigned long long* rfull = reinterpret_cast<unsigned long long*>(ring + PRING_DOUBLES);
  unsigned long long* rempty = rfull + PRING_SLOTS_MAX;
  unsigned long long* gfull = rempty + PRING_SLOTS_MAX;
  unsigned long long* gempty = gfull + QPIPE_MAX;
  __shared__ double s_red[2][PB][GEMM_CWARPS];
  __shared__ double s_rec[PAY_HDR];
  __shared__ double s_gram[PTHREADS / 32][2];
  __shared__ double s_fin[3];                             // dinv, g_new, ok
  __shared__ BlockBest s_bb[PTHREADS / 32];
  __shared__ double s_wred[PTHREADS / 32][2];             // winner reduction: the warps' candidates' target and noise
  __shared__ unsigned long long s_sbar[4];                // cubes: accumulator staging tiles, full[2] / free[2]
  __shared__ int s_stile[2];                              // the tiles they hold
  __shared__ double s_next_aux[2];                        // its target and noise
  __shared__ BlockBest s_next;                            // the winner every CTA reduced at the end of the previous step
  __shared__ int s_win, s_flag;
  __shared__ int s_pref[QT_MAX + 1];                      // exclusive prefix of the tiles' chunk counts
  __shared__ int s_wsum[GEMM_CWARPS];
  __shared__ unsigned short s_rho[QBLK * QKC];            // model rows of the units of the running block
  __shared__ __align__(8) unsigned s_ll[16 * PLL_WORLD_MAX];   // sharded: the ranks' records of this step, 8 doubles each
  __shared__ int s_uoff[QBLK];                            // their tiles' offsets into the axis-table rows: jy0 | jx0 << 16
  __shared__ unsigned long long s_clk[PCLK_N], s_tk[1];   // CTA 0, thread 0 only: lap timer (globaltimer ns)

  DevState* st = P.st;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cta = blockIdx.x, G = gridDim.x;
  const double NEG_INF = -__longlong_as_double(0x7ff0000000000000LL);
  if (*(volatile int*)&st->done) return;
  // a function-only selection started by k_begin: rows == points == generation, so ONE counter (r0) carries
  // the loop state and the barrier count (3 per step) is derived from it
  if (*(volatile int*)&st->r != 0 || *(volatile int*)&st->m != 0 || *(volatile int*)&st->iter != 0) return;
  for (int i = tid; i < GS1_RMAX; i += PTHREADS) s_kv[i] = 0.0;
  if (tid == 0) {
    for (int i = 0; i < NPIPE; ++i) { mbar_init(&gfull[i], CUBE ? 32 : 1); mbar_init(&gempty[i], GEMM_CWARPS); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_sbar[i], GEMM_CWARPS); mbar_init(&s_sbar[2 + i], QPROD); }
    for (int i = 0; i < PRING_SLOTS_MAX; ++i) { mbar_init(&rfull[i], 1); mbar_init(&rempty[i], GEMM_CWARPS); }
    mbar_init_fence();
    s_next.score = NEG_INF; s_next.id = 0x7fffffffffffffffLL; s_next.local = -1;
  }
  __syncthreads();

  const int nx = P.glen[0], ny = P.glen[1];
  const int tiles_x = (nx + TE - 1) / TE, tiles_y = (ny + TE - 1) / TE, txy = tiles_x * tiles_y;
  const int T = txy * (CUBE ? (P.glen[2] + QC - 1) / QC : P.glen[2]);
  const int zodd = CUBE ? (P.gz0 & 1) : 0;                // the Tz rows are copied from the even index below the cube's first slice
  if (tid < PCLK_N) s_clk[tid] = 0;
  if (tid == 0) s_tk[0] = gtimer_ns();
  unsigned long long d_ns = 0, dwait_ns = 0, dt0 = 0;      // thread 0 of every CTA: its time in phase D / (experiments 5, 6) waiting for its producer warps, for shared tiles
  int stop = 0;          // 1: nothing unclassified / capacity (done), 2: not positive definite
  unsigned nst = 0;      // cubes: whole tiles staged (consumers) / taken from the staging tiles (producers) so far
  unsigned gcount = 0;   // units this CTA has put through the GEMM ring so far (the barriers' phases run on across steps)
  int w_nslots = -1;     // W ring: slots of the current geometry and rows streamed through it so far
  unsigned wseq = 0;
  const bool exchange = P.world > 1;
  int r0 = 0;

  for (; r0 < iters; ++r0) {
    const int m = r0, gen = r0;
    unsigned passed = 3u * (unsigned)r0;
    // ------------------------------------------------------------------ A: the winner
    if (exchange && r0 > 0) {
      // sharded: every rank stored its record of this generation into our region as 16 tagged 8-byte slots
      // ({low or high half of a double, tag}): a slot is valid as soon as its tag is this generation's -- no
      // fence, no separate flag.  Every CTA polls the slots itself and picks the same winner.
      const unsigned tag = ll_tag(st->epoch, gen);
      const volatile unsigned long long* ll = reinterpret_cast<const volatile unsigned long long*>(
          reinterpret_cast<const char*>(P.xchg) + P.xll_off) + (size_t)(gen & 1) * P.world * 16;
      if (tid == 0) s_flag = 1;
      __syncthreads();
      if (tid < 16 * P.world && tid < 16 * PLL_WORLD_MAX) {
        const long long t0 = clock64();
        unsigned long long v = ll[tid];
        unsigned n = 0;
        while ((unsigned)(v >> 32) != tag) {
          if ((++n & 63u) == 0u && (clock64() - t0 > P.spin_cycles || *(volatile int*)&st->err != 0)) { s_flag = 0; break; }
          v = ll[tid];
        }
        s_ll[tid] = (unsigned)v;
      }
      __syncthreads();
      if (!s_flag) { if (tid == 0) persist_fail(st); return; }
      if (tid == 0) {
        int win = -1;
        double bs = 0.0;
        long long bid = 0;
        const double* recd = reinterpret_cast<const double*>(s_ll);
        for (int w = 0; w < P.world; ++w) {
          const double* pay = recd + w * 8;
          const long long loc = __double_as_longlong(pay[PLL_LOCAL]);
          if (loc < 0) continue;
          const long long id = __double_as_longlong(pay[PLL_ID]);
          if (win < 0 || better(pay[PLL_SCORE], id, bs, bid)) { win = w; bs = pay[PLL_SCORE]; bid = id; }
        }
        if (win >= 0 && m >= P.M_cap) win = -1;
        s_win = win;
        if (win >= 0) {
          const double* pay = recd + win * 8;
          s_rec[PAY_SCORE] = pay[PLL_SCORE]; s_rec[PAY_ID] = pay[PLL_ID]; s_rec[PAY_LOCAL] = pay[PLL_LOCAL];
          s_rec[PAY_TSDF] = pay[PLL_TSDF]; s_rec[PAY_NOISE] = pay[PLL_NOISE];
          for (int d = 0; d < MAXD; ++d) s_rec[PAY_X + d] = pay[PLL_X + d];
        }
      }
    } else if (r0 == 0) {
      // the first point: k_begin's record(s), in the record format of the per-step kernels
      if (!wait_records(P, st, gen)) return;
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
      if (s_win >= 0 && tid < PAY_HDR) s_rec[tid] = __ldcg(recs + (long long)s_win * P.pay_stride + tid);
    } else {
      // one GPU: the per-CTA winners were reduced by every CTA after the last barrier of the previous step
      if (tid == 0) {
        const long long jl = s_next.local;
        const int win = (jl >= 0 && m < P.M_cap) ? 0 : -1;
        s_win = win;
        if (win >= 0) {
          s_rec[PAY_SCORE] = s_next.score;
          s_rec[PAY_ID] = __longlong_as_double(s_next.id);
          s_rec[PAY_LOCAL] = __longlong_as_double(jl);
          s_rec[PAY_TSDF] = s_next_aux[0];
          s_rec[PAY_NOISE] = s_next_aux[1];
          long long rem = jl;
#pragma unroll
          for (int d = 0; d < MAXD; ++d) {
            const long long len = P.glen[d] > 0 ? P.glen[d] : 1;
            const long long jd = rem % len + (d == 2 ? P.gz0 : 0);
            rem /= len;
            s_rec[PAY_X + d] = P.gorg[d] + P.gh[d] * (double)jd;
          }
        }
      }
    }
    __syncthreads();
    const int win = s_win;
    if (win < 0) { stop = 1; break; }
    PSTAMP(PCLK_RECWAIT);

    // ------------------------------------------------------------------ B: W pass
    // rows s = cta + i*G of W; slot stride fits the longest row of this step
    const int my_rows = cta < r0 ? (r0 - 1 - cta) / G + 1 : 0;
    const int stride = (r0 + 2) & ~1;                                   // doubles, 16-byte multiple
    const int nslots = min(PRING_SLOTS_MAX, PRING_DOUBLES / (stride > 0 ? stride : 2));
    if (nslots != w_nslots) {      // new ring geometry (a few times per selection): fresh phase bits
      if (tid == 0) {
        for (int i = 0; i < PRING_SLOTS_MAX; ++i) {
          mbar_inval(&rfull[i]); mbar_inval(&rempty[i]);
          mbar_init(&rfull[i], 1); mbar_init(&rempty[i], GEMM_CWARPS);
        }
        mbar_init_fence();
      }
      __syncthreads();
      w_nslots = nslots;
      wseq = 0;
    }
    double nxp[D];
#pragma unroll
    for (int d = 0; d < D; ++d) nxp[d] = s_rec[PAY_X + d];
    if (warp > GEMM_CWARPS) {
      // producer warps 9..11 have nothing to do in this phase
    } else if (warp == GEMM_CWARPS) {
      // producer warp, lane 0: one bulk copy per row
      if (lane == 0) {
        for (int i = 0; i < my_rows; ++i) {
          const unsigned q = wseq + (unsigned)i;
          const int slot = (int)(q % (unsigned)nslots);
          const unsigned use = q / (unsigned)nslots;
          if (use > 0) mbar_wait(&rempty[slot], (use - 1) & 1u);
          const int row = cta + i * G;
          const unsigned bytes = (unsigned)(((row + 2) >> 1) << 4);    // row + 1 doubles, rounded up to 16 B
          mbar_arrive_expect_tx(&rfull[slot], bytes);
          bulk_g2s(ring + (size_t)slot * stride, P.W + (long long)row * P.ldW, bytes, &rfull[slot]);
        }
      } else {
        // lanes 1..31: the new model row joins the row list of every tile whose box it reaches
        // (tiles cta, cta + G, ..: one writer per tile), and the tiles' split-K counters are cleared
        for (int t = cta + (lane - 1) * G; t < T; t += 31 * G) {
          const int tz = t / txy, ty = (t / tiles_x) % tiles_y, tx = t % tiles_x;
          double d2 = 0.0;
          {
            const double lo = P.gorg[0] + P.gh[0] * (double)(tx * TE), hi = P.gorg[0] + P.gh[0] * (double)(min(tx * TE + TE, nx) - 1);
            const double a = fmin(lo, hi), b = fmax(lo, hi);
            const double df = nxp[0] < a ? a - nxp[0] : (nxp[0] > b ? nxp[0] - b : 0.0);
            d2 += df * df;
          }
          if (D > 1) {
            const double p1 = nxp[D > 1 ? 1 : 0];
            const double lo = P.gorg[1] + P.gh[1] * (double)(ty * TE), hi = P.gorg[1] + P.gh[1] * (double)(min(ty * TE + TE, ny) - 1);
            const double a = fmin(lo, hi), b = fmax(lo, hi);
            const double df = p1 < a ? a - p1 : (p1 > b ? p1 - b : 0.0);
            d2 += df * df;
          }
          if (D > 2) {
            const double p2 = nxp[D > 2 ? 2 : 0];
            if (CUBE) {
              const double lo = P.gorg[2] + P.gh[2] * (double)(P.gz0 + tz * QC),
                           hi = P.gorg[2] + P.gh[2] * (double)(P.gz0 + min(tz * QC + QC, P.glen[2]) - 1);
              const double a = fmin(lo, hi), b = fmax(lo, hi);
              const double df = p2 < a ? a - p2 : (p2 > b ? p2 - b : 0.0);
              d2 += df * df;
            } else {
              const double df = (P.gorg[2] + P.gh[2] * (double)(P.gz0 + tz)) - p2;
              d2 += df * df;
            }
          }
          if (0.5 * d2 * P.inv_ell2 <= QCUT) {
            const int n = __ldcg(P.rcnt + t);
            P.rlist[(size_t)t * P.rlist_stride + n] = (unsigned short)r0;
            if (!CUBE) {
              // the row's z-axis factor at this tile's slice (what the axis-2 table holds for it)
              double tzf = 1.0;
              if (D > 2) {
                const double df = (P.gorg[2] + P.gh[2] * (double)(P.gz0 + tz)) - nxp[D > 2 ? 2 : 0];
                tzf = exp(-0.5 * df * df * P.inv_ell2);
              }
              P.rtz[(size_t)t * P.rlist_stride + n] = tzf;
            }
            __stcg(P.rcnt + t, n + 1);
          }
          P.tile_cnt[t] = 0u;
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
            P.TT[d][(long long)j * P.ldtt + r0] = v;
          }
        }
      }
    } else {
      // covariance of the winner with the model rows (entries >= m stay zero)
      // k(x_i, x_w) = sf2 exp(-|x_i - x_w|^2 / 2 ell^2) is the product of row i's three axis-table entries at the
      // winner's lattice index (the very factors the update GEMM multiplies), read from the transposed tables
      // TT[d][index][row]: three coalesced loads per row instead of an exp(), all of the thread's rows in flight
      // together.  Entries >= m stay zero.
      {
        int wj[3] = {0, 0, 0};
        {
          long long rem = __double_as_longlong(s_rec[PAY_LOCAL]);
          if (win != P.rank) {      // a peer's pick: its lattice index from the coordinates
#pragma unroll
            for (int d = 0; d < D; ++d) wj[d] = (int)llrint((nxp[d] - P.gorg[d]) / P.gh[d]);
          } else {
#pragma unroll
            for (int d = 0; d < D; ++d) {
              const long long len = P.glen[d] > 0 ? P.glen[d] : 1;
              wj[d] = (int)(rem % len) + (d == 2 ? P.gz0 : 0);
              rem /= len;
            }
          }
        }
        for (int i0 = 0; i0 < m; i0 += 8 * PCW) {
          double tv[8][D];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int i = i0 + tid + PCW * k;
#pragma unroll
            for (int d = 0; d < D; ++d) tv[k][d] = i < m ? __ldcg(P.TT[d] + (long long)wj[d] * P.ldtt + i) : 0.0;
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int i = i0 + tid + PCW * k;
            if (i < m) {
              double kv = tv[k][0];
#pragma unroll
              for (int d = 1; d < D; ++d) kv *= tv[k][d];
              s_kv[i] = kv;
            }
          }
        }
      }
      named_sync(1, PCW);
      PSTAMP(PCLK_WSETUP);
      double acc[GS1_CPT];
#pragma unroll
      for (int c = 0; c < GS1_CPT; ++c) acc[c] = 0.0;
      int par = 0;
      for (int i0 = 0; i0 < my_rows; i0 += PB) {
#if GPIS_EXP == 7      // experiment: where a W-pass batch spends its time: [12] waits + loads, [13] dots + shuffles, [14] barrier
        unsigned long long t7 = (cta == 0 && tid == 0) ? gtimer_ns() : 0ull;
#define GPIS_LAP7(slot) do { if (cta == 0 && tid == 0) { const unsigned long long n7 = gtimer_ns(); s_clk[slot] += n7 - t7; t7 = n7; } } while (0)
#else
#define GPIS_LAP7(slot) do { } while (0)
#endif
        double w[PB][GS1_CPT];
        double p[PB];
        // column slots of this batch's longest row (its last), in groups of four: the slots beyond hold nothing of
        // any row of the batch and are skipped (uniform branches) -- on average two thirds of them
        const int cgmax = min(GS1_CPT / 4, ((cta + min(i0 + PB - 1, my_rows - 1) * G) / (4 * PCW)) + 1);
#pragma unroll
        for (int b = 0; b < PB; ++b) {
          p[b] = 0.0;
          const int i = i0 + b;
          if (i < my_rows) {
            const unsigned q = wseq + (unsigned)i;
            const int slot = (int)(q % (unsigned)nslots), row = cta + i * G;
            mbar_wait(&rfull[slot], (q / (unsigned)nslots) & 1u);
            const double* buf = ring + (size_t)slot * stride;
#pragma unroll
            for (int cg = 0; cg < GS1_CPT / 4; ++cg) {
              if (cg < cgmax) {
#pragma unroll
                for (int c = cg * 4; c < cg * 4 + 4; ++c) {
                  const int col = tid + PCW * c;
                  w[b][c] = col <= row ? buf[col] : 0.0;
                }
              } else {
#pragma unroll
                for (int c = cg * 4; c < cg * 4 + 4; ++c) w[b][c] = 0.0;
              }
            }
          } else {
#pragma unroll
            for (int c = 0; c < GS1_CPT; ++c) w[b][c] = 0.0;
          }
        }
        GPIS_LAP7(PCLK_PRECOMP);
        __syncwarp();
        if (lane == 0) {
#pragma unroll
          for (int b = 0; b < PB; ++b)
            if (i0 + b < my_rows) mbar_arrive(&rempty[(int)((wseq + (unsigned)(i0 + b)) % (unsigned)nslots)]);
        }
#pragma unroll
        for (int cg = 0; cg < GS1_CPT / 4; ++cg) {
          if (cg < cgmax) {
#pragma unroll
            for (int c = cg * 4; c < cg * 4 + 4; ++c) {
              const double kv = s_kv[tid + PCW * c];
#pragma unroll
              for (int b = 0; b < PB; ++b) p[b] += w[b][c] * kv;
            }
          }
        }
#pragma unroll
        for (int b = 0; b < PB; ++b)
          for (int o = 16; o > 0; o >>= 1) p[b] += __shfl_xor_sync(0xffffffffu, p[b], o);
        if (lane == 0) {
#pragma unroll
          for (int b = 0; b < PB; ++b) s_red[par][b][warp] = p[b];
        }
        GPIS_LAP7(PCLK_WAITFULL);
        named_sync(1, PCW);
        GPIS_LAP7(PCLK_EPIREG);
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
        for (int cg = 0; cg < GS1_CPT / 4; ++cg) {
          if (cg < cgmax) {
#pragma unroll
            for (int b = 0; b < PB; ++b)
#pragma unroll
              for (int c = cg * 4; c < cg * 4 + 4; ++c) acc[c] += ell[b] * w[b][c];
          }
        }
      }
      if (my_rows > 0) {
        double* yp = P.ypart + (long long)cta * GS1_RMAX + tid;
#pragma unroll
        for (int c = 0; c < GS1_CPT; ++c)
          if (tid + PCW * c < r0) yp[PCW * c] = acc[c];
      }
    }
    wseq += (unsigned)my_rows;
    PSTAMP(PCLK_WPASS);
    if (!grid_barrier(P, st, passed, &s_flag)) { if (tid == 0) persist_fail(st); return; }
    PSTAMP(PCLK_BAR1);

    // ------------------------------------------------------------------ C: d, g_new, new row of W
    {
      // The sum of the CTAs' partial Y's for this thread's column does not depend on d: its loads go out together with
      // those of |l|^2 and l.g (one L2 round trip instead of two), only the scaling by 1/d waits for the block sums.
      const int nparts = min(G, r0);
      const int gw = (cta * PTHREADS + tid) >> 5;      // (G * 12 warps x 4 columns >= 4096 rows: one pass)
      const int sub = lane >> 2, cq = lane & 3;
      const int c = gw * 4 + cq;
      double a[24];
      {
        const double* yp = P.ypart + c;
        // the CTAs' partials in a fixed order (bit-identical on every rank), 24 loads in flight
#pragma unroll
        for (int q = 0; q < 24; ++q) {
          const int part = sub + 8 * q;
          a[q] = (c < r0 && part < nparts) ? __ldcg(yp + (long long)part * GS1_RMAX) : 0.0;
        }
      }
      const double* X = P.L + (long long)r0 * P.R_cap;
      double p0 = 0.0, p1 = 0.0;
      for (int s0 = 0; s0 < r0; s0 += 8 * PTHREADS) {      // 16 loads in flight per thread
        double xv[8], gv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int si = s0 + tid + PTHREADS * k;
          xv[k] = si < r0 ? __ldcg(X + si) : 0.0;
          gv[k] = si < r0 ? __ldcg(P.g + si) : 0.0;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) { p0 += xv[k] * xv[k]; p1 += xv[k] * gv[k]; }
      }
      for (int o = 16; o > 0; o >>= 1) { p0 += __shfl_xor_sync(0xffffffffu, p0, o); p1 += __shfl_xor_sync(0xffffffffu, p1, o); }
      if (lane == 0) { s_gram[warp][0] = p0; s_gram[warp][1] = p1; }
      __syncthreads();
      if (tid == 0) {
        double xx = 0.0, xg = 0.0;
        for (int w = 0; w < PTHREADS / 32; ++w) { xx += s_gram[w][0]; xg += s_gram[w][1]; }
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
      double y = 0.0;
#pragma unroll
      for (int q = 0; q < 24; ++q) y += a[q];
      if (c < r0)
        for (int part = sub + 8 * 24; part < nparts; part += 8) y += __ldcg(P.ypart + c + (long long)part * GS1_RMAX);
      y += __shfl_xor_sync(0xffffffffu, y, 4);
      y += __shfl_xor_sync(0xffffffffu, y, 8);
      y += __shfl_xor_sync(0xffffffffu, y, 16);
      __syncthreads();
      if (sub == 0 && c < r0) P.W[(long long)r0 * P.ldW + c] = -y * s_fin[0];
    }
    // (the row lists are final since the first barrier: the unit prefix of phase D is formed here, off its critical path)
    // exclusive prefix of the tiles' COSTS in unit times: their 32-row chunks (at least one: a tile's candidates are
    // scored every step even when no model row reaches it) + QEPI for the tile's epilogue
    {
      const int per = (T + PCW - 1) / PCW;             // <= QT_MAX / 256 = 8
      int cbuf[QT_MAX / PCW], loc = 0;
      if (tid < PCW) {
#pragma unroll
        for (int k = 0; k < QT_MAX / PCW; ++k) {
          const int t = tid * per + k;
          int c = 0;
          if (k < per && t < T) { c = (__ldcg(P.rcnt + t) + QKC - 1) / QKC; c = (c > 0 ? c : 1) + QEPI; }
          cbuf[k] = c;
          loc += c;
        }
      }
      int inc = loc;
      for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += n; }
      if (tid < PCW && lane == 31) s_wsum[warp] = inc;
      __syncthreads();
      if (tid < PCW) {
        int run = inc - loc;
        for (int w = 0; w < warp; ++w) run += s_wsum[w];
#pragma unroll
        for (int k = 0; k < QT_MAX / PCW; ++k) {
          const int t = tid * per + k;
          if (k < per && t < T) s_pref[t] = run;
          run += cbuf[k];
        }
        if (tid == PCW - 1) s_pref[T] = run;
      }
      __syncthreads();
    }
    // This CTA's share [u0, u1) of the cost line; inside tile t the positions [0, C_t) are its chunks and
    // [C_t, C_t + QEPI) stand for its epilogue.  cpref(t) = s_pref[t] - QEPI t is the prefix of the chunks alone.
    const long long U = s_pref[T];
    const int Gw = (int)min((long long)G, U);
    const bool working = cta < Gw;
    const long long u0 = working ? (long long)cta * U / Gw : 0, u1 = working ? (long long)(cta + 1) * U / Gw : 0;
    int tile0 = 0, c0 = 0, n_units = 0;
    if (working) {
      int lo = 0, hi = T - 1;      // last tile whose first position is <= u0
      while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if ((long long)s_pref[mid] <= u0) lo = mid; else hi = mid - 1; }
      tile0 = lo;
      c0 = (int)(u0 - s_pref[tile0]);
      if (c0 >= s_pref[tile0 + 1] - s_pref[tile0] - QEPI) { ++tile0; c0 = 0; }      // u0 falls into the epilogue part
      // chunks up to u1: those of the tile holding u1 - 1 count up to that position
      lo = 0; hi = T - 1;
      while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if ((long long)s_pref[mid] <= u1 - 1) lo = mid; else hi = mid - 1; }
      const int tl = lo, Cl = s_pref[tl + 1] - s_pref[tl] - QEPI;
      const long long q1 = ((long long)s_pref[tl] - (long long)QEPI * tl) + min((long long)Cl, u1 - s_pref[tl]);
      const long long q0 = tile0 < T ? ((long long)s_pref[tile0] - (long long)QEPI * tile0) + c0 : q1;
      n_units = (int)max(0LL, q1 - q0);
      if (tile0 >= T) tile0 = T - 1;
    }
    const long long q_first = ((long long)s_pref[tile0] - (long long)QEPI * tile0) + c0;      // first chunk, in chunk space
    PBest best;
    best.s = NEG_INF; best.id = 0x7fffffffffffffffLL; best.j = -1;
    const double sb = [&]() {
      double beta = P.beta;
      if (P.beta_mode == 1) beta = 2.0 * log((double)P.N_total * (double)(gen + 1) * 9.869604401089358 / (6.0 * P.beta_delta));
      else if (P.beta_mode == 2) { const double k1 = (double)(gen + 1); beta = 2.0 * log((double)P.N_total * 9.869604401089358 * k1 * k1 / (6.0 * P.beta_delta)); }
      return sqrt(beta);
    }();
    const double gn1 = s_fin[1];
    // predictive-variance mode (max_subset_buffers.cu:113-115): the cells' noise joins the variance before scoring
    const bool pred_var = P.variance_mode == 1;
    double* s_cs = s_kv;                                  // [QKC][QBLK], aliases the winner's covariance vector (phase B only)
    // The units of this CTA go through the ring in blocks of QBLK.  Before a block starts every thread takes one
    // unit and leaves, in shared memory, its 16 model-row indices, its tile's offsets into the axis-table rows and
    // the rows' z factors (part A: nothing of it depends on this step's row of W) and then multiplies the
    // coefficients W[r, rho] in (part B), so that neither the producers nor the MMA warps ever wait on global
    // memory for them.
    auto stage_part_a = [&](int b0, int bn) {
      for (int ul = tid; ul < bn; ul += PTHREADS) {
        const long long u = q_first + b0 + ul;      // chunk-space index of this unit
        int lo = tile0, hi = T - 1;
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if ((long long)s_pref[mid] - (long long)QEPI * mid <= u) lo = mid; else hi = mid - 1;
        }
        const int t = lo, c = (int)(u - ((long long)s_pref[t] - (long long)QEPI * t));
        const int cnt = __ldcg(P.rcnt + t);
        const uint4* rl4 = reinterpret_cast<const uint4*>(P.rlist + (size_t)t * P.rlist_stride + c * QKC);
        unsigned rw[QKC / 2];
#pragma unroll
        for (int k = 0; k < QKC / 8; ++k) {
          const uint4 r4 = __ldcg(rl4 + k);
          rw[4 * k] = r4.x; rw[4 * k + 1] = r4.y; rw[4 * k + 2] = r4.z; rw[4 * k + 3] = r4.w;
        }
        double tv[QKC];
        if (CUBE) {      // the z factors come from the axis-2 table with the stage
#pragma unroll
          for (int k = 0; k < QKC; ++k) tv[k] = 1.0;
        } else {
          const double2* tz2 = reinterpret_cast<const double2*>(P.rtz + (size_t)t * P.rlist_stride + c * QKC);
#pragma unroll
          for (int k = 0; k < QKC; k += 2) { const double2 z2 = __ldcg(tz2 + (k >> 1)); tv[k] = z2.x; tv[k + 1] = z2.y; }
        }
#pragma unroll
        for (int k = 0; k < QKC; ++k) {
          const bool live = c * QKC + k < cnt;
          const int rr = (int)((rw[k >> 1] >> ((k & 1) * 16)) & 0xffffu);
          s_rho[k * QBLK + ul] = (unsigned short)(live ? rr : 0);      // [k][unit]: conflict-free stores
          s_cs[k * QBLK + (ul ^ ((k & 3) << 2))] = live ? tv[k] : 0.0;      // (swizzled: see the MMA loop)
        }
        if (CUBE) s_uoff[ul] = (t % tiles_x) | ((t / tiles_x) % tiles_y) << 10 | (t / txy) << 20;      // tile indices
        else s_uoff[ul] = (((t / tiles_x) % tiles_y) * QT) | ((t % tiles_x) * QT) << 16;
      }
    };
    auto stage_part_b = [&](int bn) {
      // row r0 of W is new this step: no line of it can be stale in L1, so these loads may be cached there
      const double* wrow = P.W + (long long)r0 * P.ldW;
      for (int ul = tid; ul < bn; ul += PTHREADS) {
        double wv[QKC];
#pragma unroll
        for (int k = 0; k < QKC; ++k) wv[k] = wrow[s_rho[k * QBLK + ul]];
#pragma unroll
        for (int k = 0; k < QKC; ++k) s_cs[k * QBLK + (ul ^ ((k & 3) << 2))] *= wv[k];
      }
    };
    // producers: the 64 row copies of unit ul of the block at b0 (the unit's stage and phase follow from the
    // running unit count)
    auto issue_unit = [&](int b0, int ul) {
      const unsigned q = gcount + (unsigned)(b0 + ul);
      const int stage = (int)(q % NPIPE);
      mbar_wait(&gempty[stage], ((q / NPIPE) & 1u) ^ 1u);
      double* As = ring + (size_t)stage * STAGE;
      const int rho = (int)s_rho[lane * QBLK + ul];
      const int uo = s_uoff[ul];
      if (CUBE) {
        // 16-byte asynchronous copies (LDGSTS), eight lanes per 128-byte row segment: 25 instructions per lane and
        // unit, completion counted by the stage's mbarrier (cp.async.mbarrier.arrive.noinc, 32 arrivals).  A bulk
        // copy per row segment (96 per unit) costs a serial elect-one loop of ~40 cycles each: 2 us per unit and warp,
        // which left the producer warps no time for the epilogues they run now.
        double* Bs = As + QKC * QLDC;
        double* Zs = Bs + QKC * QLDC;
        const int rsel = lane >> 3, pc = (lane & 7) * 2;
        const double* sy = P.T[1] + ((uo >> 10) & 1023) * QC + pc;
        const double* sx = P.T[0] + (uo & 1023) * QC + pc;
        const double* sz = P.T[D > 2 ? 2 : 0] + ((P.gz0 + (uo >> 20) * QC) & ~1);
        const long long l1 = P.ldt[1], l0 = P.ldt[0], l2 = P.ldt[D > 2 ? 2 : 0];
#pragma unroll
        for (int i = 0; i < QKC / 4; ++i) {
          const int row = i * 4 + rsel;
          const long long rr = (long long)s_rho[row * QBLK + ul];
          g_cp_async16(As + row * QLDC + pc, sy + rr * l1);
          g_cp_async16(Bs + row * QLDC + pc, sx + rr * l0);
          g_cp_async16(Zs + row * QLDC + pc, sz + rr * l2 + pc);
        }
        g_cp_async16(Zs + lane * QLDC + QC, sz + (long long)rho * l2 + QC);      // the z rows' ninth piece (even-aligned start)
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(&gfull[stage])) : "memory");
      } else {
        double* Bs = As + QKC * QLD;
        if (lane == 0) mbar_arrive_expect_tx(&gfull[stage], 2 * QKC * QT * (unsigned)sizeof(double));
        __syncwarp();
        bulk_g2s(As + lane * QLD, P.T[1] + (long long)rho * P.ldt[1] + (uo & 0xffff), QT * sizeof(double), &gfull[stage]);
        bulk_g2s(Bs + lane * QLD, P.T[0] + (long long)rho * P.ldt[0] + (uo >> 16), QT * sizeof(double), &gfull[stage]);
      }
    };
    // Cubes: a tile that lies inside this CTA's range leaves the MMA warps as a 32 KB staging tile in shared memory
    // (two of them, full / free mbarriers) and the four PRODUCER warps -- idle nine tenths of the time -- run its
    // epilogue (posterior update, straddle score, classification, arg-max) a quarter each, while the MMA warps are
    // already on the next tile.  whole_upto(u): such tiles among the first u units of this CTA.
    auto whole_upto = [&](int uend) {
      int n = 0, t = tile0, c = c0, pos = 0;
      while (pos < uend && t < T) {
        const int Ct = s_pref[t + 1] - s_pref[t] - QEPI, take = min(Ct - c, uend - pos);
        pos += take;
        if (c == 0 && take == Ct) ++n;
        ++t; c = 0;
      }
      return n;
    };
    double2* stg = reinterpret_cast<double2*>(ring + (size_t)NPIPE * STAGE);      // [2][QPAIRS]
    int ptaken = 0, pb = 0;      // producers: staged tiles of THIS step finished, batches of the tile at hand done
    PTileBest ptb;
    ptb.s = NEG_INF; ptb.q = -1;
    // one batch (GEPI_SPT fragment slots per lane: a quarter of this warp's quarter) of the staged tile at hand
    auto service_batch = [&]() {
      const int slot = (int)(nst & 1u);
      const int t = *(volatile int*)&s_stile[slot];
      const double2* sv = stg + (size_t)slot * QPAIRS;
      const long long tbase = (long long)t * QPAIRS;
      const int etz = t / txy, ety = (t / tiles_x) % tiles_y, etx = t % tiles_x;
      int pq[GEPI_SPT];
      bool plive[GEPI_SPT];
      double2 pv[GEPI_SPT], pmu[GEPI_SPT], pvar[GEPI_SPT], pnz[GEPI_SPT];
      uchar2 pfl[GEPI_SPT];
#pragma unroll
      for (int k = 0; k < GEPI_SPT; ++k) {
        pq[k] = ((warp - GEMM_CWARPS) * 16 + pb * GEPI_SPT + k) * 32 + lane;
        plive[k] = true;
        pmu[k] = __ldcg(reinterpret_cast<const double2*>(P.mu) + tbase + pq[k]);
        pvar[k] = __ldcg(reinterpret_cast<const double2*>(P.var) + tbase + pq[k]);
        pfl[k] = __ldcg(reinterpret_cast<const uchar2*>(P.flags) + tbase + pq[k]);
      }
#pragma unroll
      for (int k = 0; k < GEPI_SPT; ++k) pv[k] = sv[pq[k]];
      if (pred_var) {
#pragma unroll
        for (int k = 0; k < GEPI_SPT; ++k) pnz[k] = persist_cell_noise<CUBE>(P, pq[k], etx, ety, etz, nx, ny, pfl[k]);
      }
      const double2* pnzp = pnz;
      persist_pairs(GEPI_SPT, P, tbase, pq, plive, pv, pmu, pvar, pfl, pred_var ? pnzp : nullptr, sb, gn1, ptb);
      if (++pb == 16 / GEPI_SPT) {
        persist_merge_best<CUBE>(P, ptb, etx, ety, etz, nx, ny, best);
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_sbar[2 + slot]);
        ++nst;
        ++ptaken;
        pb = 0;
        ptb.s = NEG_INF; ptb.q = -1;
      }
    };
    // producers, main loop: like issue_unit, but while the stage is still in use they take staged tiles
    auto issue_unit_serving = [&](int b0, int ul, int n_whole) {
      const unsigned q = gcount + (unsigned)(b0 + ul);
      const int stage = (int)(q % NPIPE);
      while (!mbar_test(&gempty[stage], ((q / NPIPE) & 1u) ^ 1u)) {      // copies first, a batch of epilogue while the ring is full
        if (pb > 0 || (ptaken < n_whole && mbar_test(&s_sbar[nst & 1u], (nst >> 1) & 1u))) service_batch();
      }
      issue_unit(b0, ul);
    };
    if (s_fin[2] == 0.0) { stop = 2; break; }          // every CTA computes the same pivot: uniform exit
    PSTAMP(PCLK_FINISH);
    if (!grid_barrier(P, st, passed, &s_flag)) { if (tid == 0) persist_fail(st); return; }
    PSTAMP(PCLK_BAR2);
    // Block 0's part A, then the first QPIPE units' copies go out while part B multiplies the coefficients in.
    // (Measured: doing part A and these copies BEFORE the barrier gains nothing -- the wait-for-data time of the MMA
    // warps is the cost of polling a completed barrier once per unit, not a pipeline-fill stall.)
    const int bn0 = min(QBLK, n_units), npre = min(NPIPE, bn0);
    stage_part_a(0, bn0);
    __syncthreads();
    if (warp >= GEMM_CWARPS)
      for (int ul = warp - GEMM_CWARPS; ul < npre; ul += QPROD) issue_unit(0, ul);

    // ------------------------------------------------------------------ D: update GEMM + epilogue
    PSTAMP(PCLK_PREFIX);
    if (tid == 0) dt0 = gtimer_ns();
    const int wm = warp >> 1, wn = warp & 1;
    // walking state of the consumers: tile, chunk inside it, its chunk count
    int wt = tile0, wc = c0, wCt = s_pref[tile0 + 1] - s_pref[tile0] - QEPI;
    int seg_c0 = c0;                                      // first chunk of the running segment
    int def_t[2], ndef = 0;
    // accumulator fragments f = 0..7: 64 x 64 tiles f = i * 4 + j (2 x 4 blocks of the warp's 16 x 32), cubes
    // f = i * 2 + j (4 x 2 blocks of the warp's [2 slices x 16 jy] x 16 jx); fragment slot q = (warp * 8 + f) * 32 + lane
    double acc[8][2];
#pragma unroll
    for (int f = 0; f < 8; ++f) acc[f][0] = acc[f][1] = 0.0;
    for (int b0 = 0; b0 < n_units; b0 += QBLK) {
      const int bn = min(QBLK, n_units - b0);
      unsigned long long tk0 = 0;
      if (cta == 0 && tid == 0) tk0 = gtimer_ns();
      if (b0 > 0) {
        __syncthreads();                                  // the previous block's coefficients are no longer read
        stage_part_a(b0, bn);
        __syncthreads();
      }
      stage_part_b(bn);
      __syncthreads();
      if (cta == 0 && tid == 0) { s_clk[PCLK_PRECOMP] += gtimer_ns() - tk0; if (b0 == 0) s_clk[PCLK_UNITS] += (unsigned long long)(U - (long long)QEPI * T); }
      if (warp >= GEMM_CWARPS) {
        // ------------------------------- producers: every QPROD-th unit each -------------------------------
        if (CUBE && GPIS_PROD_EPI) {
          const int n_whole = whole_upto(b0 + bn);       // staged by the end of this block of units
          for (int ul = (b0 == 0 ? npre : 0) + warp - GEMM_CWARPS; ul < bn; ul += QPROD) issue_unit_serving(b0, ul, n_whole);
          while (ptaken < n_whole) {                     // every copy of the block is out: the MMA warps cannot be waiting for us
            if (pb == 0) mbar_wait(&s_sbar[nst & 1u], (nst >> 1) & 1u);
            service_batch();
          }
        } else {
          for (int ul = (b0 == 0 ? npre : 0) + warp - GEMM_CWARPS; ul < bn; ul += QPROD) issue_unit(b0, ul);
        }
      } else {
        // --------------------------------- consumers ---------------------------------
        for (int ul = 0; ul < bn; ++ul) {
          const unsigned q = gcount + (unsigned)(b0 + ul);
          const int gstage = (int)(q % NPIPE);
          mbar_wait(&gfull[gstage], (q / NPIPE) & 1u);
          const double* As = ring + (size_t)gstage * STAGE;
          // a k-step reads the coefficients of four consecutive rows of the unit: their column is XOR-swizzled by the
          // row (rows are QBLK doubles = a whole number of bank cycles apart) so that the four land in different banks
          const double* cs = s_cs + (ul ^ ((lane & 3) << 2));
#if GPIS_EXP == 1      // experiment: the MMA loop runs twice (second time with zero coefficients): the rise is its cost
          for (int rep = 0; rep < 2; ++rep)
#endif
          if (CUBE) {
            const double* Bs = As + QKC * QLDC;
            const double* Zs = Bs + QKC * QLDC + zodd + 2 * warp;
#pragma unroll
            for (int k4 = 0; k4 < QKC / 4; ++k4) {
              const int kb = k4 * 4 + (lane & 3);
#if GPIS_EXP == 1
              const double cv = rep ? 0.0 : cs[kb * QBLK];
#else
              const double cv = cs[kb * QBLK];
#endif
              // A[(jz, jy)][rho] = (W[r,rho] Tz[rho][jz]) Ty[rho][jy]: the same product, in the same order, as the
              // 64 x 64 form's (tz w) ty
              const double z0 = Zs[kb * QLDC] * cv, z1 = Zs[kb * QLDC + 1] * cv;
              const double y0 = As[kb * QLDC + (lane >> 2)], y1 = As[kb * QLDC + 8 + (lane >> 2)];
              const double af[4] = {y0 * z0, y1 * z0, y0 * z1, y1 * z1};
              double bf[2];
#pragma unroll
              for (int j = 0; j < 2; ++j) bf[j] = Bs[kb * QLDC + j * 8 + (lane >> 2)];
#pragma unroll
              for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) dmma8x8x4(acc[i * 2 + j][0], acc[i * 2 + j][1], af[i], bf[j]);
            }
          } else {
            const double* Bs = As + QKC * QLD;
#pragma unroll
            for (int k4 = 0; k4 < QKC / 4; ++k4) {
              const int kb = k4 * 4 + (lane & 3);
#if GPIS_EXP == 1
              const double cv = rep ? 0.0 : cs[kb * QBLK];
#else
              const double cv = cs[kb * QBLK];
#endif
              double af[2], bf[4];
#pragma unroll
              for (int i = 0; i < 2; ++i) af[i] = As[kb * QLD + wm * 16 + i * 8 + (lane >> 2)] * cv;
#pragma unroll
              for (int j = 0; j < 4; ++j) bf[j] = Bs[kb * QLD + wn * 32 + j * 8 + (lane >> 2)];
#pragma unroll
              for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma8x8x4(acc[i * 4 + j][0], acc[i * 4 + j][1], af[i], bf[j]);
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&gempty[gstage]);
          ++wc;
          if (wc < wCt && b0 + ul + 1 < n_units) continue;
          // ---- the running segment of tile wt ends here
          const bool whole = (seg_c0 == 0) && (wc == wCt);
          const long long tbase = (long long)wt * QPAIRS;
          if (cta == 0 && tid == 0) tk0 = gtimer_ns();
          if (whole && CUBE && GPIS_PROD_EPI) {      // to a staging tile in shared memory: the producer warps run its epilogue
            const int slot = (int)(nst & 1u);
            mbar_wait(&s_sbar[2 + slot], ((nst >> 1) & 1u) ^ 1u);
            double2* sv = stg + (size_t)slot * QPAIRS;
#pragma unroll
            for (int f = 0; f < 8; ++f) sv[(warp * 8 + f) * 32 + lane] = make_double2(acc[f][0], acc[f][1]);
            if (tid == 0) s_stile[slot] = wt;
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_sbar[slot]);
            ++nst;
          } else if (whole) {       // straight from the accumulator registers: all of the tile's state loads in flight first
            double2 pmu[8], pvar[8];
            uchar2 pfl[8];
#pragma unroll
            for (int f = 0; f < 8; ++f) {
              const long long s2 = tbase + (warp * 8 + f) * 32 + lane;
              pmu[f] = __ldcg(reinterpret_cast<const double2*>(P.mu) + s2);
              pvar[f] = __ldcg(reinterpret_cast<const double2*>(P.var) + s2);
              pfl[f] = __ldcg(reinterpret_cast<const uchar2*>(P.flags) + s2);
            }
            const int etz = wt / txy, ety = (wt / tiles_x) % tiles_y, etx = wt % tiles_x;
            int pq[8];
            bool plive[8];
            double2 pv[8], pnz[8];
#pragma unroll
            for (int f = 0; f < 8; ++f) {
              pq[f] = (warp * 8 + f) * 32 + lane;
              plive[f] = true;
              pv[f] = make_double2(acc[f][0], acc[f][1]);
            }
            if (pred_var) {
#pragma unroll
              for (int f = 0; f < 8; ++f) pnz[f] = persist_cell_noise<CUBE>(P, pq[f], etx, ety, etz, nx, ny, pfl[f]);
            }
            PTileBest tb;
            tb.s = NEG_INF; tb.q = -1;
            const double2* pnzp = pnz;
            persist_pairs(8, P, tbase, pq, plive, pv, pmu, pvar, pfl, pred_var ? pnzp : nullptr, sb, gn1, tb);
            persist_merge_best<CUBE>(P, tb, etx, ety, etz, nx, ny, best);
          } else {
            // a tile cut by this CTA's range: partial accumulators -> the CTA's workspace slot (0: the range's
            // first tile, 1: its last), then arrive on the tile
            double2* ws = reinterpret_cast<double2*>(P.gws) + ((size_t)cta * 2 + (wt == tile0 ? 0 : 1)) * QPAIRS;
#pragma unroll
            for (int f = 0; f < 8; ++f) __stcg(&ws[(warp * 8 + f) * 32 + lane], make_double2(acc[f][0], acc[f][1]));
            named_sync(1, PCW);      // the MMA warps' stores happen before thread 0's fence (cumulative) and its arrival
            if (tid == 0) { __threadfence(); atomicAdd(&P.tile_cnt[wt], 1u); }
            def_t[ndef++] = wt;
          }
          if (cta == 0 && tid == 0) s_clk[whole ? PCLK_EPIREG : PCLK_WAITFULL] += gtimer_ns() - tk0;
#pragma unroll
          for (int f = 0; f < 8; ++f) acc[f][0] = acc[f][1] = 0.0;
          ++wt;
          wc = 0;
          seg_c0 = 0;
          if (wt < T) wCt = s_pref[wt + 1] - s_pref[wt] - QEPI;
        }
      }
    }
    gcount += (unsigned)n_units;
    if (warp < GEMM_CWARPS) {
      PSTAMP(PCLK_GEMM);

      // ---- tiles shared with neighbouring CTAs: wait for all segments, split the epilogue by segment index
      bool ok = true;
      for (int d = 0; d < ndef && ok; ++d) {
        const int t2 = def_t[d];
        const int g_first = unit_owner((long long)s_pref[t2], U, Gw);                      // owner of the tile's first chunk
        const int nseg = unit_owner((long long)s_pref[t2 + 1] - QEPI - 1, U, Gw) - g_first + 1;      // .. of its last chunk
        const int seg = cta - g_first;
        if (tid == 0) {
          const unsigned long long tw0 = (GPIS_EXP == 6) ? gtimer_ns() : 0ull;
          s_flag = spin_ge(&P.tile_cnt[t2], (unsigned)nseg, P, st) ? 1 : 0;
          if (GPIS_EXP == 6) dwait_ns += gtimer_ns() - tw0;      // per-CTA view: waiting for a shared tile's other contributors
        }
        named_sync(1, PCW);
        if (!s_flag) { ok = false; break; }
        __threadfence();
        constexpr int UNITS = QPAIRS / PCW;                              // 8 units of 256 fragment slots
        const int un0 = seg * UNITS / nseg, un1 = (seg + 1) * UNITS / nseg;
        // the first contributor's partial sits in its slot 1 unless the tile is the first of its range
        // (it starts its range when that range begins at the tile's first chunk or inside the previous tile's epilogue part)
        const int slot_first = ((long long)g_first * U / Gw >= (long long)s_pref[t2] - QEPI) ? 0 : 1;
        const double2* wsb = reinterpret_cast<const double2*>(P.gws);
        const long long tbase = (long long)t2 * QPAIRS;
        const int e2x = t2 % tiles_x, e2y = (t2 / tiles_x) % tiles_y, e2z = t2 / txy;
        PTileBest tb;
        tb.s = NEG_INF; tb.q = -1;
        for (int ub = un0; ub < un1; ub += GEPI_SPT) {
          double2 v[GEPI_SPT], mu[GEPI_SPT], var[GEPI_SPT];
          uchar2 fl[GEPI_SPT];
#pragma unroll
          for (int k = 0; k < GEPI_SPT; ++k) {
            const int q = (ub + k < un1 ? ub + k : un1 - 1) * PCW + tid;
            v[k] = __ldcg(wsb + ((size_t)g_first * 2 + slot_first) * QPAIRS + q);
          }
          for (int sgi = 1; sgi < nseg; ++sgi) {
            double2 pz[GEPI_SPT];
#pragma unroll
            for (int k = 0; k < GEPI_SPT; ++k) {
              const int q = (ub + k < un1 ? ub + k : un1 - 1) * PCW + tid;
              pz[k] = __ldcg(wsb + ((size_t)(g_first + sgi) * 2) * QPAIRS + q);
            }
#pragma unroll
            for (int k = 0; k < GEPI_SPT; ++k) { v[k].x += pz[k].x; v[k].y += pz[k].y; }
          }
#pragma unroll
          for (int k = 0; k < GEPI_SPT; ++k) {
            const long long s2 = tbase + (ub + k < un1 ? ub + k : un1 - 1) * PCW + tid;
            mu[k] = __ldcg(reinterpret_cast<const double2*>(P.mu) + s2);
            var[k] = __ldcg(reinterpret_cast<const double2*>(P.var) + s2);
            fl[k] = __ldcg(reinterpret_cast<const uchar2*>(P.flags) + s2);
          }
          int pq[GEPI_SPT];
          bool plive[GEPI_SPT];
          double2 pnz[GEPI_SPT];
#pragma unroll
          for (int k = 0; k < GEPI_SPT; ++k) {
            plive[k] = ub + k < un1;
            pq[k] = (plive[k] ? ub + k : un1 - 1) * PCW + tid;
          }
          if (pred_var) {
#pragma unroll
            for (int k = 0; k < GEPI_SPT; ++k) pnz[k] = persist_cell_noise<CUBE>(P, pq[k], e2x, e2y, e2z, nx, ny, fl[k]);
          }
          const double2* pnzp = pnz;
          persist_pairs(GEPI_SPT, P, tbase, pq, plive, v, mu, var, fl, pred_var ? pnzp : nullptr, sb, gn1, tb);
        }
        persist_merge_best<CUBE>(P, tb, e2x, e2y, e2z, nx, ny, best);
      }
      if (!ok && tid == 0) persist_fail(st);
    }
    // CTA winner (cubes: the producer warps hold the best of the tiles they finished)
    {
      for (int o = 16; o > 0; o >>= 1) {
        const double os = __shfl_xor_sync(0xffffffffu, best.s, o);
        const long long oid = __shfl_xor_sync(0xffffffffu, best.id, o);
        const long long oj = __shfl_xor_sync(0xffffffffu, best.j, o);
        if (oj >= 0 && better(os, oid, best.s, best.id)) { best.s = os; best.id = oid; best.j = oj; }
      }
      if (lane == 0) { s_bb[warp].score = best.s; s_bb[warp].id = best.id; s_bb[warp].local = best.j; }
      const unsigned long long tp0 = (GPIS_EXP == 5 && tid == 0) ? gtimer_ns() : 0ull;
      __syncthreads();
      if (GPIS_EXP == 5 && tid == 0) dwait_ns += gtimer_ns() - tp0;      // per-CTA view: waiting for the producer warps' last epilogues
      if (tid == 0) {
        BlockBest b = s_bb[0];
        for (int w = 1; w < PTHREADS / 32; ++w)
          if (s_bb[w].local >= 0 && (b.local < 0 || better(s_bb[w].score, s_bb[w].id, b.score, b.id))) b = s_bb[w];
        BlockBest* dst = &P.partials[cta];
        __stcg(&dst->score, b.score);
        __stcg(&dst->id, b.id);
        __stcg(&dst->local, b.local);
        // the candidate's target and noise travel with it (entries G + cta): whoever wins, the next step needs
        // no dependent load for them
        BlockBest* aux = &P.partials[G + cta];
        __stcg(&aux->score, b.local >= 0 ? P.tsdf[b.local] : 0.0);
        __stcg(reinterpret_cast<double*>(&aux->id), b.local >= 0 ? (P.noise ? P.noise[b.local] : P.noise_scalar) : 0.0);
      }
      PSTAMP(PCLK_EPI);
      if (tid == 0) d_ns += gtimer_ns() - dt0;
    }
    // ---- every CTA waits for all, then reduces the CTA winners itself: the same winner everywhere
    if (!grid_barrier(P, st, passed, &s_flag)) { if (tid == 0) persist_fail(st); return; }
    PSTAMP(PCLK_BAR3);
    // one record per lane over the first ceil(G / 32) warps (ONE L2 round trip, not one per 32 records), a shuffle tree
    // per warp, then the warps' bests in order: the same winner in every CTA
    {
      const int nwr = (G + 31) >> 5;
      if (warp < nwr) {
        BlockBest b;
        b.score = NEG_INF; b.id = 0x7fffffffffffffffLL; b.local = -1;
        double bt = 0.0, bn = 0.0;
        const int i = warp * 32 + lane;
        if (i < G) {
          const BlockBest* pp = &P.partials[i];
          const double ps = __ldcg(&pp->score);
          const long long pid = __ldcg(&pp->id), ploc = __ldcg(&pp->local);
          const double pt = __ldcg(&P.partials[G + i].score), pn = __ldcg(reinterpret_cast<const double*>(&P.partials[G + i].id));
          if (ploc >= 0) { b.score = ps; b.id = pid; b.local = ploc; bt = pt; bn = pn; }
        }
        for (int o = 16; o > 0; o >>= 1) {
          const double os = __shfl_xor_sync(0xffffffffu, b.score, o);
          const long long oid = __shfl_xor_sync(0xffffffffu, b.id, o);
          const long long ol = __shfl_xor_sync(0xffffffffu, b.local, o);
          const double ot = __shfl_xor_sync(0xffffffffu, bt, o), on = __shfl_xor_sync(0xffffffffu, bn, o);
          if (ol >= 0 && (b.local < 0 || better(os, oid, b.score, b.id))) { b.score = os; b.id = oid; b.local = ol; bt = ot; bn = on; }
        }
        if (lane == 0) { s_bb[warp] = b; s_wred[warp][0] = bt; s_wred[warp][1] = bn; }
      }
      __syncthreads();
      if (tid == 0) {
        BlockBest b = s_bb[0];
        double bt = s_wred[0][0], bn = s_wred[0][1];
        for (int w = 1; w < nwr; ++w)
          if (s_bb[w].local >= 0 && (b.local < 0 || better(s_bb[w].score, s_bb[w].id, b.score, b.id))) {
            b = s_bb[w]; bt = s_wred[w][0]; bn = s_wred[w][1];
          }
        s_next = b; s_next_aux[0] = bt; s_next_aux[1] = bn;
      }
    }
    __syncthreads();
    if (cta == 0) {
      if (tid == 0) {
        const int itx = r0;
        if (itx <= P.M_cap) {
          P.hist[itx] = itx;
          P.hist[(P.M_cap + 1) + itx] = T;
          P.hist[2 * (P.M_cap + 1) + itx] = (int)min(P.N, (long long)0x7fffffff);
        }
        // state the host (and the finalize kernel) reads; iter tags the record published next
        st->r = r0 + 1; st->r0 = r0; st->m = r0 + 1; st->n_sel = r0 + 1; st->iter = r0 + 1; st->gnew[0] = s_fin[1];
        st->nnoise = s_rec[PAY_NOISE];
#pragma unroll
        for (int d = 0; d < D; ++d) st->nx[d] = s_rec[PAY_X + d];
      }
      // sharded: this rank's record of every step goes to all peers as tagged slots; the last winner is also
      // published in the record format of the per-step kernels (the finalize kernel reads it)
      if (exchange && r0 + 1 < iters) {
        const BlockBest b = s_next;
        if (tid == 0) {
          double* pay = reinterpret_cast<double*>(s_ll);
          const long long jl = b.local;
          pay[PLL_SCORE] = b.score;
          pay[PLL_ID] = __longlong_as_double(jl >= 0 ? cand_id(P, jl) : -1);
          pay[PLL_LOCAL] = __longlong_as_double(jl);
          pay[PLL_TSDF] = s_next_aux[0];
          pay[PLL_NOISE] = s_next_aux[1];
          long long rem = jl >= 0 ? jl : 0;
#pragma unroll
          for (int d = 0; d < MAXD; ++d) {
            const long long len = P.glen[d] > 0 ? P.glen[d] : 1;
            const long long jd = rem % len + (d == 2 ? P.gz0 : 0);
            rem /= len;
            pay[PLL_X + d] = P.gorg[d] + P.gh[d] * (double)jd;
          }
        }
        __syncthreads();
        const unsigned tag = ll_tag(st->epoch, r0 + 1);
        for (int q = tid; q < 16 * P.world; q += PTHREADS) {
          const int p = q >> 4, k = q & 15;
          volatile unsigned long long* dst = reinterpret_cast<volatile unsigned long long*>(
              reinterpret_cast<char*>(P.peer_xchg[p]) + P.xll_off) + ((size_t)((r0 + 1) & 1) * P.world + P.rank) * 16 + k;
          *dst = ((unsigned long long)tag << 32) | (unsigned long long)s_ll[k];
        }
        __syncthreads();
      }
      if (r0 + 1 == iters) {
        __syncthreads();
        const BlockBest b = s_next;
        write_record(P, b.score, b.local, 0);
        publish_record(P, st, r0 + 1, 0);
      }
      PSTAMP(PCLK_WINNER);
      if (tid == 0) s_clk[PCLK_STEPS] += 1;
    }
  }
  if (cta == 0 && tid == 0) {
    if (stop == 1) st->done = 1;
    if (stop == 2) { st->err = -5; st->done = 1; }
    if (P.phase_clk)
      for (int i = 0; i < PCLK_N; ++i) P.phase_clk[i] = s_clk[i];
  }
  if (tid == 0 && P.phase_clk) {      // per-CTA view of phase D (load balance): [16 + 2 cta] total, [17 + 2 cta] waiting for data
    P.phase_clk[PCLK_N + 2 * cta] = d_ns;
    P.phase_clk[PCLK_N + 2 * cta + 1] = dwait_ns;
  }
}

#endif  // GPIS_EMULATE

}  // namespace gpis
