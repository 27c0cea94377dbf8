// Batch of independent small objects (BASELINE config 5: 512 x 32^3 object GPIS fits): the WHOLE greedy
// level-set selection of one object runs inside one thread block, and one launch advances every object
// of the batch -- no per-step launches, no inter-block synchronisation, nothing on the host.
//
// Same algorithm and the same arithmetic rules as the lattice backend (select_active_set_straddle.m:62-130
// through the incremental factor): W = L^-1 explicit, new entry of V = L^-1 K(rows, candidates) from the
// separable axis tables,
//     v[jz][jy][jx] = sum_rho (w[rho] Tz[rho][jz]) * Ty[rho][jy] * Tx[rho][jx],
// straddle score without FMA contraction, lowest index on ties, last pick selected but not in the model.
// What differs is the mapping: the axis tables of the object (3 x 256 x 32 doubles = 192 KB) live in the
// block's shared memory, every thread owns a 2 x 2 column of the (jy, jx) plane and walks the z-slices
// eight at a time with a 2 x 2 x 8 register tile (plain DFMA: on B200 the fp64 FMA pipe and the fp64
// tensor pipe have the same measured peak, profiles/r01_fp64_peak.json), and the posterior state of the
// object (mu, var, flags: 17 bytes per lattice point) streams through L2 once per step.
// Limits: function values only, scalar noise, fixed beta, lattice <= 32 per axis, <= 256 points.
#pragma once
#include "gpis_launch.cuh"
#include "gpis_common.cuh"

namespace gpis {

constexpr int BT = 256;          // threads per object
constexpr int BRMAX = 256;       // model rows (= points) per object
constexpr int BAX = 32;          // lattice points per axis
constexpr int BZB = 8;           // z-slices per register tile
constexpr size_t BATCH_SMEM = sizeof(double) * (3 * BRMAX * BAX + 3 * BRMAX + 3 * BRMAX + BRMAX);

struct BatchParams {
  int nx, ny, nz;
  long long N;                   // nx * ny * nz
  int n_objects, K;
  long long first;               // first pick (straddle.m:36-37), the same for every object
  double org[3], h[3];           // x = org + h * index
  double inv_ell2, sf2, noise, level, eps, beta;
  int variance_mode;
  double mean_w[3], mean_c;
  const double* tsdf;            // [n_objects][N]
  double* mu;                    // [n_objects][N]  running posterior mean (output)
  double* var;                   // [n_objects][N]  running posterior variance (output)
  unsigned char* flags;          // [n_objects][N]  F_ACTIVE | F_HIGH | F_LOW (output)
  long long* ids;                // [n_objects][K]  picks in selection order (output)
  int* n_sel;                    // [n_objects]
  int* n_model;                  // [n_objects]
  int* err;                      // [n_objects]     0 or GPIS_ERR_NOT_PD (-5)
  double* W;                     // [gridDim.x][BRMAX][BRMAX]   row-major lower L^-1, per-block workspace
  double* Wt;                    // [gridDim.x][BRMAX][BRMAX]   its transpose
};

__device__ __forceinline__ bool batch_better(double s, long long id, double bs, long long bid) {
  return (s > bs) || (s == bs && id < bid);
}

// FAST selects the tuned forms of the three inner loops (vector shared-memory loads and w folded into the
// Ty factor in the update; the state of eight lattice points loaded before any of them is processed; four
// independent partial sums in the two triangular passes of the append).  Both forms are kept: the plain
// one is the reference the tuned one is tested against (same picks; sums grouped differently).
template <bool FAST>
__global__ void __launch_bounds__(BT, 1) k_batch_select(BatchParams P) {
  GPIS_DYN_SMEM(double, sm);
  double* Tx = sm;                           // [BRMAX][BAX]
  double* Ty = Tx + BRMAX * BAX;
  double* Tz = Ty + BRMAX * BAX;             // carries sf2
  double* s_ax = Tz + BRMAX * BAX;           // [3][BRMAX] active points
  double* s_w = s_ax + 3 * BRMAX;            // [BRMAX] new row of W
  double* s_k = s_w + BRMAX;                 // [BRMAX] covariance of the new point with the model rows
  double* s_l = s_k + BRMAX;                 // [BRMAX] new row of L
  double* s_g = s_l + BRMAX;                 // [BRMAX] g = L^-1 (y - m)
  __shared__ double s_red[2][BT / 32];
  __shared__ double s_bs[BT / 32];
  __shared__ long long s_bj[BT / 32];
  __shared__ double s_d, s_gnew;
  __shared__ long long s_pending;
  __shared__ int s_stop;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nx = P.nx, ny = P.ny, nz = P.nz;
  const double NEG_INF = -1.7976931348623157e308;          // scores are finite; best_j < 0 marks "none"
  const double sb = sqrt(P.beta);
  double* W = P.W + (long long)blockIdx.x * BRMAX * BRMAX;
  double* Wt = P.Wt + (long long)blockIdx.x * BRMAX * BRMAX;
  const int ty2 = (tid >> 4) * 2, tx2 = (tid & 15) * 2;          // this thread's 2 x 2 column of the plane

  for (int obj = blockIdx.x; obj < P.n_objects; obj += gridDim.x) {
    const double* tsdf = P.tsdf + (long long)obj * P.N;
    double* mu = P.mu + (long long)obj * P.N;
    double* var = P.var + (long long)obj * P.N;
    unsigned char* flags = P.flags + (long long)obj * P.N;
    long long* ids = P.ids + (long long)obj * P.K;

    // prior state: mu = m(x), var = sf2 (k_grid_init_state)
    for (long long j = tid; j < P.N; j += BT) {
      const int jx = (int)(j % nx), jy = (int)((j / nx) % ny), jz = (int)(j / ((long long)nx * ny));
      const int jd[3] = {jx, jy, jz};
      double m = P.mean_c;
      for (int d = 0; d < 3; ++d) m += P.mean_w[d] * (P.org[d] + P.h[d] * (double)jd[d]);
      mu[j] = m;
      var[j] = P.sf2;
      flags[j] = (j == P.first) ? F_ACTIVE : 0;
    }
    if (tid == 0) {
      ids[0] = P.first;
      s_pending = P.first;
      s_stop = 0;
    }
    int r = 0, n_sel = 1, bad = 0;
    __syncthreads();

    for (int step = 1; step < P.K; ++step) {
      // ---------------------------------------------------------------- append the pending pick
      const long long pj = s_pending;
      const int pjd[3] = {(int)(pj % nx), (int)((pj / nx) % ny), (int)(pj / ((long long)nx * ny))};
      double xn[3];
      for (int d = 0; d < 3; ++d) xn[d] = P.org[d] + P.h[d] * (double)pjd[d];
      if (tid < r) {
        double d2 = 0.0;
        for (int d = 0; d < 3; ++d) { const double df = s_ax[d * BRMAX + tid] - xn[d]; d2 += df * df; }
        s_k[tid] = P.sf2 * exp(-0.5 * d2 * P.inv_ell2);
      }
      __syncthreads();
      // l = W k: thread s walks column entries Wt[c][s], c <= s (coalesced across s), ascending c
      double lv = 0.0;
      if (tid < r) {
        if (FAST) {
          double l0 = 0.0, l1 = 0.0, l2 = 0.0, l3 = 0.0;
          int c = 0;
          for (; c + 3 <= tid; c += 4) {
            l0 += Wt[(long long)c * BRMAX + tid] * s_k[c];
            l1 += Wt[(long long)(c + 1) * BRMAX + tid] * s_k[c + 1];
            l2 += Wt[(long long)(c + 2) * BRMAX + tid] * s_k[c + 2];
            l3 += Wt[(long long)(c + 3) * BRMAX + tid] * s_k[c + 3];
          }
          for (; c <= tid; ++c) l0 += Wt[(long long)c * BRMAX + tid] * s_k[c];
          lv = (l0 + l1) + (l2 + l3);
        } else {
          for (int c = 0; c <= tid; ++c) lv += Wt[(long long)c * BRMAX + tid] * s_k[c];
        }
        s_l[tid] = lv;
      }
      // |l|^2 and l.g: fixed shuffle tree per warp, then the 8 warp sums in order
      double q = (tid < r) ? lv * lv : 0.0, lg = (tid < r) ? lv * s_g[tid] : 0.0;
      for (int o = 16; o > 0; o >>= 1) {
        q += __shfl_xor_sync(0xffffffffu, q, o);
        lg += __shfl_xor_sync(0xffffffffu, lg, o);
      }
      if (lane == 0) { s_red[0][warp] = q; s_red[1][warp] = lg; }
      __syncthreads();
      if (tid == 0) {
        double qs = 0.0, ls = 0.0;
        for (int w8 = 0; w8 < BT / 32; ++w8) { qs += s_red[0][w8]; ls += s_red[1][w8]; }
        const double d2 = (P.sf2 + P.noise) - qs;
        if (!(d2 > 0.0)) s_stop = 2;
        const double dd = sqrt(d2);
        double m = P.mean_c;
        for (int d = 0; d < 3; ++d) m += P.mean_w[d] * xn[d];
        s_d = dd;
        s_gnew = ((tsdf[pj] - m) - ls) / dd;
      }
      __syncthreads();
      if (s_stop == 2) { bad = 1; break; }
      const double dd = s_d, gnew = s_gnew;
      // new row of W: w[c] = -(sum_{s >= c} l[s] W[s][c]) / d, w[r] = 1 / d
      if (tid <= r) {
        double wv;
        if (tid < r) {
          double y = 0.0;
          if (FAST) {
            double y1 = 0.0, y2 = 0.0, y3 = 0.0;
            int s = tid;
            for (; s + 3 < r; s += 4) {
              y += s_l[s] * W[(long long)s * BRMAX + tid];
              y1 += s_l[s + 1] * W[(long long)(s + 1) * BRMAX + tid];
              y2 += s_l[s + 2] * W[(long long)(s + 2) * BRMAX + tid];
              y3 += s_l[s + 3] * W[(long long)(s + 3) * BRMAX + tid];
            }
            for (; s < r; ++s) y += s_l[s] * W[(long long)s * BRMAX + tid];
            y = (y + y1) + (y2 + y3);
          } else {
            for (int s = tid; s < r; ++s) y += s_l[s] * W[(long long)s * BRMAX + tid];
          }
          wv = -y / dd;
        } else {
          wv = 1.0 / dd;
        }
        s_w[tid] = wv;
        W[(long long)r * BRMAX + tid] = wv;
        Wt[(long long)tid * BRMAX + r] = wv;
      }
      // axis-table rows of the new point, zero beyond the lattice
      if (tid < 3 * BAX) {
        const int a = tid / BAX, j = tid % BAX;
        const int len = (a == 0) ? nx : (a == 1) ? ny : nz;
        double t = 0.0;
        if (j < len) {
          const double df = (P.org[a] + P.h[a] * (double)j) - xn[a];
          t = exp(-0.5 * df * df * P.inv_ell2);
          if (a == 2) t *= P.sf2;
        }
        (a == 0 ? Tx : a == 1 ? Ty : Tz)[r * BAX + j] = t;
      }
      if (tid < 3) s_ax[tid * BRMAX + r] = xn[tid];
      if (tid == 0) s_g[r] = gnew;
      r += 1;
      __syncthreads();

      // ---------------------------------------------------------------- update, score, classify, arg-max
      double best_s = NEG_INF;
      long long best_j = -1;
      for (int z0 = 0; z0 < nz; z0 += BZB) {
        double acc[BZB][2][2];
#pragma unroll
        for (int z = 0; z < BZB; ++z) acc[z][0][0] = acc[z][0][1] = acc[z][1][0] = acc[z][1][1] = 0.0;
        if (FAST) {
          for (int rho = 0; rho < r; ++rho) {
            const double w = s_w[rho];
            const double2 yy = *reinterpret_cast<const double2*>(Ty + rho * BAX + ty2);
            const double2 xx = *reinterpret_cast<const double2*>(Tx + rho * BAX + tx2);
            const double y0 = w * yy.x, y1 = w * yy.y;
            const double p00 = y0 * xx.x, p01 = y0 * xx.y, p10 = y1 * xx.x, p11 = y1 * xx.y;
            const double2* tz2 = reinterpret_cast<const double2*>(Tz + rho * BAX + z0);
#pragma unroll
            for (int zq = 0; zq < BZB / 2; ++zq) {
              const double2 t = tz2[zq];
              acc[2 * zq][0][0] += t.x * p00;
              acc[2 * zq][0][1] += t.x * p01;
              acc[2 * zq][1][0] += t.x * p10;
              acc[2 * zq][1][1] += t.x * p11;
              acc[2 * zq + 1][0][0] += t.y * p00;
              acc[2 * zq + 1][0][1] += t.y * p01;
              acc[2 * zq + 1][1][0] += t.y * p10;
              acc[2 * zq + 1][1][1] += t.y * p11;
            }
          }
        } else {
          for (int rho = 0; rho < r; ++rho) {
            const double w = s_w[rho];
            const double y0 = Ty[rho * BAX + ty2], y1 = Ty[rho * BAX + ty2 + 1];
            const double x0 = Tx[rho * BAX + tx2], x1 = Tx[rho * BAX + tx2 + 1];
            const double p00 = y0 * x0, p01 = y0 * x1, p10 = y1 * x0, p11 = y1 * x1;
            const double* tz = Tz + rho * BAX + z0;
#pragma unroll
            for (int z = 0; z < BZB; ++z) {
              const double cz = w * tz[z];
              acc[z][0][0] += cz * p00;
              acc[z][0][1] += cz * p01;
              acc[z][1][0] += cz * p10;
              acc[z][1][1] += cz * p11;
            }
          }
        }
        // posterior update, straddle score, classification of the thread's 2 x 2 x 8 lattice points, two
        // z-slices (eight points) at a time; FAST loads the state of all eight before touching any
        constexpr int ZP = 2;
#pragma unroll
        for (int zb = 0; zb < BZB; zb += ZP) {
          double m0[ZP][2][2], v0[ZP][2][2];
          unsigned char f0[ZP][2][2];
          bool ok[ZP][2][2];
#pragma unroll
          for (int zz = 0; zz < ZP; ++zz)
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
              for (int b = 0; b < 2; ++b) {
                const int jz = z0 + zb + zz, jy = ty2 + a, jx = tx2 + b;
                ok[zz][a][b] = jz < nz && jy < ny && jx < nx;
                if (FAST && ok[zz][a][b]) {
                  const long long j = (long long)jx + (long long)nx * (jy + (long long)ny * jz);
                  m0[zz][a][b] = mu[j];
                  v0[zz][a][b] = var[j];
                  f0[zz][a][b] = flags[j];
                }
              }
#pragma unroll
          for (int zz = 0; zz < ZP; ++zz)
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
              for (int b = 0; b < 2; ++b) {
                if (!ok[zz][a][b]) continue;
                const int jz = z0 + zb + zz, jy = ty2 + a, jx = tx2 + b;
                const long long j = (long long)jx + (long long)nx * (jy + (long long)ny * jz);
                const double v = acc[zb + zz][a][b];
                const double mold = FAST ? m0[zz][a][b] : mu[j];
                const double vold = FAST ? v0[zz][a][b] : var[j];
                const double m1 = __dadd_rn(mold, __dmul_rn(v, gnew));
                const double v1 = __dadd_rn(vold, -__dmul_rn(v, v));
                mu[j] = m1;
                var[j] = v1;
                const unsigned char f = FAST ? f0[zz][a][b] : flags[j];
                if (f & (F_ACTIVE | F_HIGH | F_LOW)) continue;
                double ve = v1;
                if (P.variance_mode == 1) ve = __dadd_rn(ve, P.noise);
                const double wd = __dmul_rn(sb, ve);                  // variance, not sigma: straddle.m:106-107
                const double hi = __dadd_rn(m1, wd), lo = __dadd_rn(m1, -wd);
                const double sc = fmin(__dadd_rn(hi, -P.level), __dadd_rn(P.level, -lo));
                unsigned char f1 = f;
                if (__dadd_rn(lo, P.eps) > P.level) f1 |= F_HIGH;      // :122-125
                if (__dadd_rn(hi, -P.eps) < P.level) f1 |= F_LOW;
                if (f1 != f) flags[j] = f1;
                if (sc == sc && (best_j < 0 || batch_better(sc, j, best_s, best_j))) { best_s = sc; best_j = j; }
              }
        }
      }
      for (int o = 16; o > 0; o >>= 1) {
        const double os = __shfl_xor_sync(0xffffffffu, best_s, o);
        const long long oj = __shfl_xor_sync(0xffffffffu, best_j, o);
        if (oj >= 0 && (best_j < 0 || batch_better(os, oj, best_s, best_j))) { best_s = os; best_j = oj; }
      }
      if (lane == 0) { s_bs[warp] = best_s; s_bj[warp] = best_j; }
      __syncthreads();
      if (tid == 0) {
        double bs = s_bs[0];
        long long bj = s_bj[0];
        for (int w8 = 1; w8 < BT / 32; ++w8)
          if (s_bj[w8] >= 0 && (bj < 0 || batch_better(s_bs[w8], s_bj[w8], bs, bj))) { bs = s_bs[w8]; bj = s_bj[w8]; }
        if (bj < 0) {
          s_stop = 1;                                   // nothing left unclassified (straddle.m:84-86)
        } else {
          s_pending = bj;
          ids[n_sel] = bj;
          flags[bj] |= F_ACTIVE;
        }
      }
      __syncthreads();
      if (s_stop) break;
      n_sel += 1;
    }
    if (tid == 0) {
      P.n_sel[obj] = n_sel;
      P.n_model[obj] = r;
      P.err[obj] = bad ? -5 : 0;
    }
    __syncthreads();                                    // s_stop / s_pending are rewritten for the next object
  }
}

inline void launch_batch_select(const BatchParams& P, int n_blocks, bool fast, cudaStream_t s) {
  auto k0 = k_batch_select<false>;
  auto k1 = k_batch_select<true>;
  if (fast) GPIS_LAUNCH_SMEM(k1, dim3((unsigned)n_blocks), dim3(BT), BATCH_SMEM, s, P);
  else GPIS_LAUNCH_SMEM(k0, dim3((unsigned)n_blocks), dim3(BT), BATCH_SMEM, s, P);
}

}  // namespace gpis
