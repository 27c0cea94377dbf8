// Selection kernels: incremental Cholesky block append (one CTA) and the HBM-streaming
// candidate update + straddle scoring + arg-max kernel (persistent grid).
//
// Reference behaviour restated (not ported): select_active_set_straddle.m:62-126 for the
// rule, se_cov_derivative.m for the covariance blocks; the reference GPU path
// (gpu_active_set_selector.cpp:499-560) refactorises and re-solves everything each step.
#pragma once
#include "gpis_launch.cuh"
#include "gpis_common.cuh"

namespace gpis {

__device__ __forceinline__ bool better(double s, long long id, double bs, long long bid) {
  return (s > bs) || (s == bs && id < bid);
}

// Covariance of rows [f, d_1 f .. d_D f] at x (first argument) against the same rows at z.
// diff = z - x (se_cov_derivative.m:45-47).  out[a][b]: a = row kind at x, b = kind at z.
template <int NB, int D>
__device__ __forceinline__ void cov_rows_rows(const EngineParams& P, const double* x, const double* z,
                                              double out[NB][NB]) {
  double diff[D], d2 = 0.0;
#pragma unroll
  for (int d = 0; d < D; ++d) { diff[d] = z[d] - x[d]; d2 += diff[d] * diff[d]; }
  const double k = P.sf2 * exp(-0.5 * d2 * P.inv_ell2);
  out[0][0] = k;
  if (NB > 1) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      out[1 + d][0] = diff[d] * k * P.inv_ell2;    // [d_d f(x), f(z)]      (:47)
      out[0][1 + d] = -diff[d] * k * P.inv_ell2;   // [f(x), d_d f(z)]      (:63)
    }
#pragma unroll
    for (int di = 0; di < D; ++di)
#pragma unroll
      for (int dj = 0; dj < D; ++dj)              // [d_dj f(x), d_di f(z)] (:85-87)
        out[1 + dj][1 + di] = ((di == dj ? 1.0 : 0.0) - diff[di] * diff[dj] * P.inv_ell2) * k * P.inv_ell2;
  }
}

__device__ __forceinline__ double block_reduce_sum(double v, double* scratch) {
  // deterministic: fixed shuffle tree, then warp partials summed in order by every thread
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[w] = v;
  __syncthreads();
  double s = 0.0;
  for (int i = 0; i < nw; ++i) s += scratch[i];
  return s;
}

// ---------------------------------------------------------------------------------------
// k_init: candidate state + first exchange record
// ---------------------------------------------------------------------------------------
__global__ void k_init_candidates(EngineParams P) {
  const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (j < P.N) {
    double m = P.mean_c;
    for (int d = 0; d < MAXD; ++d)
      if (P.cx[d]) m += P.mean_w[d] * P.cx[d][j];
    P.mu[j] = m;
    P.var[j] = P.sf2;
    P.amb[j] = 0.0;
    P.flags[j] = 0;
  }
  if (j < P.n_tiles) P.live[0][j] = (int)j;
}

__device__ inline void write_record(const EngineParams& P, double score, long long j, int rows) {
  double* pay = P.pay_send;
  if (threadIdx.x == 0) {
    pay[PAY_SCORE] = score;
    long long* pl = reinterpret_cast<long long*>(pay);
    pl[PAY_LOCAL] = j;
    pl[PAY_ID] = j >= 0 ? cand_id(P, j) : -1;
    if (j >= 0) {
      pay[PAY_TSDF] = P.tsdf[j];
      pay[PAY_NOISE] = P.noise ? P.noise[j] : P.noise_scalar;
      long long rem = j;
      for (int d = 0; d < MAXD; ++d) {
        if (P.glen[0] > 0) {       // grid backend: coordinates from the lattice index
          const long long len = P.glen[d] > 0 ? P.glen[d] : 1;
          const long long jd = rem % len + (d == 2 ? P.gz0 : 0);
          rem /= len;
          pay[PAY_X + d] = P.gorg[d] + P.gh[d] * (double)jd;
        } else {
          pay[PAY_X + d] = P.cx[d] ? P.cx[d][j] : 0.0;
        }
        pay[PAY_NRM + d] = P.nrm[d] ? P.nrm[d][j] : 0.0;
      }
    }
  }
  if (j >= 0 && rows > 0) {
    const double* col = P.V + (j / TW) * P.tile_stride + (j % TW);
    for (int s = threadIdx.x; s < rows; s += blockDim.x) pay[PAY_HDR + s] = __ldcg(col + (long long)s * TW);
  }
}


// ---------------------------------------------------------------------------------------
// Exchange records.  Collective mode: the host all-gathers pay_send into pay_recv between the
// update and the append kernels.  Peer mode: the producing kernel stores the record into every
// rank's exchange region over NVLink and raises a generation flag; the consuming kernel spins on
// its local flags.  Generation g of a selection is consumed by update iteration g; regions are
// double-buffered by parity (a rank can be at most one generation ahead of a peer).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ const double* records_ptr(const EngineParams& P, int gen) {
  return P.peer_mode ? P.xchg + (long long)(gen & 1) * P.world * P.pay_stride : P.pay_recv;
}

__device__ __forceinline__ unsigned long long record_tag(unsigned epoch, int gen) {
  return ((unsigned long long)epoch << 32) | (unsigned)(gen + 1);
}

// whole block; returns false on timeout (a peer died): the caller stops the selection
__device__ inline bool wait_records(const EngineParams& P, DevState* st, int gen) {
  if (!P.peer_mode) return true;
  __shared__ int s_to;
  if (threadIdx.x == 0) s_to = 0;
  __syncthreads();
  if ((int)threadIdx.x < P.world) {
    const unsigned long long want = record_tag(st->epoch, gen);
    volatile unsigned long long* f = P.xflag + (gen & 1) * P.world + threadIdx.x;
    const long long t0 = clock64();
    while (*f != want) {
      if (clock64() - t0 > P.spin_cycles || *(volatile int*)&st->err != 0) { s_to = 1; break; }
      __nanosleep(100);
    }
  }
  __threadfence_system();
  __syncthreads();
  if (s_to) {
    if (threadIdx.x == 0) { atomicCAS(&st->err, 0, -7); st->done = 1; }
    return false;
  }
  return true;
}

// whole block, after pay_send holds the record (header + `rows` column entries)
__device__ inline void publish_record(const EngineParams& P, const DevState* st, int gen, int rows) {
  if (!P.peer_mode) return;
  __threadfence();
  __syncthreads();
  const int n = PAY_HDR + rows;
  for (int p = 0; p < P.world; ++p) {
    double* dst = P.peer_xchg[p] + ((long long)(gen & 1) * P.world + P.rank) * P.pay_stride;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = P.pay_send[i];
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < P.world) {
    volatile unsigned long long* f = P.peer_xflag[threadIdx.x] + (gen & 1) * P.world + P.rank;
    *f = record_tag(st->epoch, gen);
  }
}

// whole block, when this rank has consumed the last records of its selection: tell every peer
__device__ inline void signal_selection_done(const EngineParams& P, const DevState* st) {
  if (!P.peer_mode || !P.peer_xdone) return;
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < P.world) {
    volatile unsigned long long* f = P.peer_xdone[threadIdx.x] + P.rank;
    *f = (unsigned long long)st->epoch;
  }
}

__global__ void k_begin(EngineParams P, long long first_local, unsigned epoch) {
  DevState* st = P.st;
  // peer mode: the regions are double-buffered by generation parity only, so selection `epoch` may not
  // write into a peer's region while that peer still reads the last generation of selection epoch - 1
  if (P.peer_mode && P.xdone && (int)threadIdx.x < P.world && epoch > 0) {
    volatile unsigned long long* f = P.xdone + threadIdx.x;
    const long long t0 = clock64();
    while (*f + 1 < (unsigned long long)epoch) {
      if (clock64() - t0 > P.spin_cycles) break;      // the records' own tags still protect the readers
      __nanosleep(200);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    st->r = st->r0 = st->m = st->n_sel = st->done = st->iter = st->err = 0;
    st->n_live[0] = P.n_tiles;
    st->n_live[1] = 0;
    st->blocks_done = 0;
    st->n_scored = 0;
    st->grid_bar = 0;
    st->pending_id = -1;
    st->epoch = epoch;
  }
  __syncthreads();
  write_record(P, first_local >= 0 ? __longlong_as_double(0x7ff0000000000000LL) : -__longlong_as_double(0x7ff0000000000000LL),
               first_local, 0);
  publish_record(P, st, 0, 0);
}

// ---------------------------------------------------------------------------------------
// k_append: choose the global winner among the exchange records, append its NB rows to L.
//   X = L^-1 Q(old rows, new rows)   (column 0 is the winner's stored column of V;
//                                      gradient columns by a blocked triangular solve)
//   S = Q(new,new) - X'X = Lnn Lnn',  g_new = Lnn^-1 (y - m - X'g)
// EXTERNAL = true: the point comes from a record with no stored column (gpis_fit).
// ---------------------------------------------------------------------------------------
template <int NB, int D>
__global__ void __launch_bounds__(APP_THREADS, 1) k_append(EngineParams P, int external, int finalize_only) {
  DevState* st = P.st;
  __shared__ int s_win;
  __shared__ double s_blk[32][33];
  __shared__ double s_x[NB][32];
  __shared__ double s_red[APP_THREADS / 32];
  __shared__ double s_dot[NB][NB + 1];
  if (st->done) {
    if (finalize_only && !external) signal_selection_done(P, st);
    return;
  }
  const int tid = threadIdx.x;
  const int gen = st->iter;
  if (!external && !wait_records(P, st, gen)) return;
  const double* recs = external ? P.pay_recv : records_ptr(P, gen);
  if (tid == 0) {
    int win = -1;
    double bs = 0.0;
    long long bid = 0;
    for (int w = 0; w < P.world; ++w) {
      const double* pay = recs + (long long)w * P.pay_stride;
      const long long* pl = reinterpret_cast<const long long*>(pay);
      if (pl[PAY_LOCAL] < 0) continue;
      if (win < 0 || better(pay[PAY_SCORE], pl[PAY_ID], bs, bid)) { win = w; bs = pay[PAY_SCORE]; bid = pl[PAY_ID]; }
    }
    if (win < 0) st->done = 1;                       // nothing unclassified: straddle.m:84-86
    else if (!finalize_only && st->m >= P.M_cap) { st->done = 1; win = -1; }
    s_win = win;
  }
  __syncthreads();
  const int win = s_win;
  if (win < 0) {
    if (finalize_only && !external) signal_selection_done(P, st);
    return;
  }
  const double* pay = recs + (long long)win * P.pay_stride;
  const long long* pl = reinterpret_cast<const long long*>(pay);
  const int r0 = st->r, m = st->m;
  if (tid == 0) {
    if (win == P.rank && !external) P.flags[pl[PAY_LOCAL]] |= F_ACTIVE;   // straddle.m:121
    P.sel_ids[m] = pl[PAY_ID];
    st->n_sel = m + 1;
  }
  if (finalize_only) {
    if (!external) signal_selection_done(P, st);
    return;
  }

  double nx[D];
#pragma unroll
  for (int d = 0; d < D; ++d) nx[d] = pay[PAY_X + d];
  const double nnoise = pay[PAY_NOISE];
  const int ldl = P.R_cap;
  double* Lr[NB];
#pragma unroll
  for (int a = 0; a < NB; ++a) Lr[a] = P.L + (long long)(r0 + a) * ldl;

  // right-hand sides: covariance of the new rows with the old rows
  const int a0 = external ? 0 : 1;   // rows that need the triangular solve
  if (!external)
    for (int s = tid; s < r0; s += APP_THREADS) Lr[0][s] = pay[PAY_HDR + s];
  if (a0 < NB) {
    constexpr int ONB = NB;
    for (int i = tid; i < m; i += APP_THREADS) {
      double z[D], c[NB][NB];
#pragma unroll
      for (int d = 0; d < D; ++d) z[d] = P.AX[d][i];
      cov_rows_rows<NB, D>(P, nx, z, c);
#pragma unroll
      for (int a = 0; a < NB; ++a)
        if (a >= a0)
#pragma unroll
          for (int b = 0; b < ONB; ++b) Lr[a][i * ONB + b] = c[a][b];
    }
    __syncthreads();
    // blocked forward substitution, in place, NB - a0 right-hand sides
    for (int jb = 0; jb < r0; jb += 32) {
      const int nbk = min(32, r0 - jb);
      for (int e = tid; e < 32 * 32; e += APP_THREADS) {
        const int i = e >> 5, t = e & 31;
        s_blk[i][t] = (i < nbk && t < nbk) ? P.L[(long long)(jb + i) * ldl + jb + t] : (i == t ? 1.0 : 0.0);
      }
      __syncthreads();
      if (tid < 32) {
        const int lane = tid;
        double q[NB];
#pragma unroll
        for (int a = 0; a < NB; ++a) q[a] = (a >= a0 && lane < nbk) ? Lr[a][jb + lane] : 0.0;
        for (int t = 0; t < nbk; ++t) {
          const double dinv = 1.0 / s_blk[t][t];
#pragma unroll
          for (int a = 0; a < NB; ++a) {
            double xt = q[a] * dinv;
            xt = __shfl_sync(0xffffffffu, xt, t);
            if (lane == t) q[a] = xt;
            else if (lane > t) q[a] -= s_blk[lane][t] * xt;
          }
        }
#pragma unroll
        for (int a = 0; a < NB; ++a)
          if (a >= a0) {
            s_x[a][lane] = q[a];
            if (lane < nbk) Lr[a][jb + lane] = q[a];
          }
      }
      __syncthreads();
      for (int s = jb + nbk + tid; s < r0; s += APP_THREADS) {
        const double* lrow = P.L + (long long)s * ldl + jb;
        double acc[NB];
#pragma unroll
        for (int a = 0; a < NB; ++a) acc[a] = 0.0;
        for (int t = 0; t < nbk; ++t) {
          const double l = lrow[t];
#pragma unroll
          for (int a = 0; a < NB; ++a) acc[a] += l * s_x[a][t];
        }
#pragma unroll
        for (int a = 0; a < NB; ++a)
          if (a >= a0) Lr[a][s] -= acc[a];
      }
      __syncthreads();
    }
  }
  __syncthreads();
  // Gram block X'X and X'g
  for (int a = 0; a < NB; ++a)
    for (int b = 0; b <= a + 1; ++b) {
      double p = 0.0;
      if (b <= a) { for (int s = tid; s < r0; s += APP_THREADS) p += Lr[a][s] * Lr[b][s]; }
      else        { for (int s = tid; s < r0; s += APP_THREADS) p += Lr[a][s] * P.g[s]; }
      const double tot = block_reduce_sum(p, s_red);
      if (tid == 0) { if (b <= a) s_dot[a][b] = tot; else s_dot[a][NB] = tot; }
    }
  __syncthreads();
  if (tid == 0) {
    double Qn[NB][NB];
#pragma unroll
    for (int a = 0; a < NB; ++a)
#pragma unroll
      for (int b = 0; b < NB; ++b) Qn[a][b] = 0.0;
    Qn[0][0] = P.sf2 + nnoise;
    // noise leak of se_cov_derivative.m:28-35,85-87: gradient diagonal = (sf2 + noise) / ell^2
    for (int d = 1; d < NB; ++d) Qn[d][d] = (P.sf2 + nnoise) * P.inv_ell2;
    double Ln[NB][NB];
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
    double y[NB], gn[NB];
    double mf = P.mean_c;
#pragma unroll
    for (int d = 0; d < D; ++d) mf += P.mean_w[d] * nx[d];
    y[0] = pay[PAY_TSDF] - mf;
    for (int d = 1; d < NB; ++d) y[d] = pay[PAY_NRM + d - 1] - P.mean_w[d - 1];
    for (int a = 0; a < NB; ++a) {
      double s = y[a] - s_dot[a][NB];
      for (int t = 0; t < a; ++t) s -= Ln[a][t] * gn[t];
      gn[a] = s / Ln[a][a];
    }
    for (int a = 0; a < NB; ++a) {
      for (int b = 0; b < NB; ++b) {
        st->Lnn[a * MAXNB + b] = Ln[a][b];
        if (b <= a) Lr[a][r0 + b] = Ln[a][b];
      }
      st->gnew[a] = gn[a];
      P.g[r0 + a] = gn[a];
    }
#pragma unroll
    for (int d = 0; d < D; ++d) { st->nx[d] = nx[d]; P.AX[d][m] = nx[d]; }
    st->nnoise = nnoise;
    st->r0 = r0;
    st->r = r0 + NB;
    st->m = m + 1;
    if (!ok) { st->err = -5; st->done = 1; }
  }
}

// ---------------------------------------------------------------------------------------
// k_update: one selection step over the live candidate tiles.
//   v_new = Lnn^-1 (k(new rows, x_j) - X' V[0:r0, j]);  V[r0.., j] = v_new
//   var_j -= |v_new|^2 ; mu_j += v_new . g_new ; straddle score, classification, arg-max.
// HBM-bound: reads r0 rows of every live tile once (8*TW bytes each), writes NB rows.
// ---------------------------------------------------------------------------------------
template <int NB, int D>
__global__ void __launch_bounds__(UPD_THREADS, 2) k_update(EngineParams P) {
  GPIS_DYN_SMEM_F64(s_ell);                         // [NB][ell_chunk]
  __shared__ double s_red[UPD_WARPS][NB][TW];
  __shared__ BlockBest s_best[2];
  __shared__ int s_flag, s_cnt;
  DevState* st = P.st;
  if (st->done) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_cnt = 0;
  const int r0 = st->r0;
  const int cur = st->iter & 1;
  const int n_live = st->n_live[cur];
  const int* live = P.live[cur];
  const int SC = P.ell_chunk;
  const double NEG_INF = -__longlong_as_double(0x7ff0000000000000LL);

  // straddle scale sqrt(beta_t)
  double beta = P.beta;
  if (P.beta_mode == 1) beta = 2.0 * log((double)P.N_total * (double)(st->iter + 1) * 9.869604401089358 / (6.0 * P.beta_delta));
  else if (P.beta_mode == 2) { const double k1 = (double)(st->iter + 1); beta = 2.0 * log((double)P.N_total * 9.869604401089358 * k1 * k1 / (6.0 * P.beta_delta)); }
  const double sb = sqrt(beta);

  double nx[D], Ln[NB][NB], gn[NB];
#pragma unroll
  for (int d = 0; d < D; ++d) nx[d] = st->nx[d];
#pragma unroll
  for (int a = 0; a < NB; ++a) {
    gn[a] = st->gnew[a];
#pragma unroll
    for (int b = 0; b < NB; ++b) Ln[a][b] = st->Lnn[a * MAXNB + b];
  }

  BlockBest best;
  best.score = NEG_INF; best.id = 0x7fffffffffffffffLL; best.local = -1;
  const bool single_chunk = r0 <= SC;
  bool ell_loaded = false;

  for (int t = blockIdx.x; t < n_live; t += gridDim.x) {
    const int tile = live[t];
    const double* Vt = P.V + (long long)tile * P.tile_stride;
    double acc[NB][2];
#pragma unroll
    for (int a = 0; a < NB; ++a) acc[a][0] = acc[a][1] = 0.0;

    for (int c0 = 0; c0 < r0; c0 += SC) {
      const int len = min(SC, r0 - c0);
      if (!(single_chunk && ell_loaded)) {
        __syncthreads();
#pragma unroll
        for (int a = 0; a < NB; ++a) {
          const double* lrow = P.L + (long long)(r0 + a) * P.R_cap + c0;
          for (int s = tid; s < len; s += UPD_THREADS) s_ell[a * SC + s] = lrow[s];
        }
        __syncthreads();
        ell_loaded = true;
      }
      const double2* vp = reinterpret_cast<const double2*>(Vt + (long long)c0 * TW) + lane;
      constexpr int U = 8;
      int s = warp;
      for (; s + UPD_WARPS * (U - 1) < len; s += UPD_WARPS * U) {
        double2 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = vp[(long long)(s + UPD_WARPS * u) * (TW / 2)];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int a = 0; a < NB; ++a) {
            const double l = s_ell[a * SC + s + UPD_WARPS * u];
            acc[a][0] += l * v[u].x;
            acc[a][1] += l * v[u].y;
          }
      }
      for (; s < len; s += UPD_WARPS) {
        const double2 v = vp[(long long)s * (TW / 2)];
#pragma unroll
        for (int a = 0; a < NB; ++a) {
          const double l = s_ell[a * SC + s];
          acc[a][0] += l * v.x;
          acc[a][1] += l * v.y;
        }
      }
    }
#pragma unroll
    for (int a = 0; a < NB; ++a) {
      s_red[warp][a][2 * lane] = acc[a][0];
      s_red[warp][a][2 * lane + 1] = acc[a][1];
    }
    if (tid == 0) s_flag = 0;
    __syncthreads();

    if (tid < TW) {
      const int c = tid;
      const long long j = (long long)tile * TW + c;
      double sc = NEG_INF;
      long long id = 0x7fffffffffffffffLL;
      bool still = false, was_unc = false;
      if (j < P.N) {
        double tt[NB];
#pragma unroll
        for (int a = 0; a < NB; ++a) {
          double s = 0.0;
#pragma unroll
          for (int w = 0; w < UPD_WARPS; ++w) s += s_red[w][a][c];
          tt[a] = s;
        }
        double diff[D], d2 = 0.0, xm = 0.0;
#pragma unroll
        for (int d = 0; d < D; ++d) { diff[d] = P.cx[d][j] - nx[d]; d2 += diff[d] * diff[d]; }
        (void)xm;
        const double k = P.sf2 * exp(-0.5 * d2 * P.inv_ell2);
        tt[0] = k - tt[0];
#pragma unroll
        for (int a = 1; a < NB; ++a) tt[a] = diff[a - 1] * k * P.inv_ell2 - tt[a];   // [d_a f(x_new), f(x_j)]
        double v[NB], dv = 0.0, dm = 0.0;
#pragma unroll
        for (int a = 0; a < NB; ++a) {
          double s = tt[a];
#pragma unroll
          for (int b = 0; b < NB; ++b)
            if (b < a) s -= Ln[a][b] * v[b];
          v[a] = s / Ln[a][a];
          dv += v[a] * v[a];
          dm += v[a] * gn[a];
        }
        double* Vw = P.V + (long long)tile * P.tile_stride + (long long)r0 * TW + c;
#pragma unroll
        for (int a = 0; a < NB; ++a) Vw[a * TW] = v[a];
        const double var = P.var[j] - dv;
        const double mu = P.mu[j] + dm;
        P.var[j] = var;
        P.mu[j] = mu;
        unsigned char f = P.flags[j];
        if (!(f & (F_ACTIVE | F_HIGH | F_LOW))) {
          was_unc = true;
          double ve = var;
          if (P.variance_mode == 1) ve = __dadd_rn(var, P.noise ? P.noise[j] : P.noise_scalar);
          const double w = __dmul_rn(sb, ve);                 // variance, not sigma: straddle.m:106-107
          const double hi = __dadd_rn(mu, w), lo = __dadd_rn(mu, -w);
          sc = fmin(__dadd_rn(hi, -P.level), __dadd_rn(P.level, -lo));   // :108
          id = cand_id(P, j);
          P.amb[j] = sc;
          if (__dadd_rn(lo, P.eps) > P.level) f |= F_HIGH;    // :122-125
          if (__dadd_rn(hi, -P.eps) < P.level) f |= F_LOW;
          P.flags[j] = f;
          still = !(f & (F_HIGH | F_LOW));
          if (!(sc == sc)) sc = NEG_INF;                       // NaN never wins
        }
      }
      // arg-max over the 64 columns: two warps, shuffle tree on (score, lowest id)
      long long jj = j;
      for (int o = 16; o > 0; o >>= 1) {
        const double os = __shfl_xor_sync(0xffffffffu, sc, o);
        const long long oid = __shfl_xor_sync(0xffffffffu, id, o);
        const long long oj = __shfl_xor_sync(0xffffffffu, jj, o);
        if (better(os, oid, sc, id)) { sc = os; id = oid; jj = oj; }
      }
      const unsigned any_still = __ballot_sync(0xffffffffu, still);
      const unsigned scored = __ballot_sync(0xffffffffu, was_unc);
      if (lane == 0) {
        s_best[warp].score = sc; s_best[warp].id = id; s_best[warp].local = jj;
        if (any_still) atomicOr(&s_flag, 1);
        if (scored) atomicAdd(&s_cnt, __popc(scored));
      }
    }
    __syncthreads();
    if (tid == 0) {
      for (int w = 0; w < 2; ++w)
        if (s_best[w].score > NEG_INF && better(s_best[w].score, s_best[w].id, best.score, best.id)) best = s_best[w];
      if (s_flag) {
        const int pos = atomicAdd(&st->n_live[cur ^ 1], 1);
        P.live[cur ^ 1][pos] = tile;
      }
    }
  }

  // grid-level arg-max: last CTA to finish reduces the per-CTA partials
  __shared__ int s_last;
  if (tid == 0) {
    P.partials[blockIdx.x] = best;
    if (s_cnt) atomicAdd(&st->n_scored, s_cnt);
    __threadfence();
    const unsigned ticket = atomicAdd(&st->blocks_done, 1u);
    s_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  __shared__ BlockBest s_part[UPD_WARPS];
  BlockBest b;
  b.score = NEG_INF; b.id = 0x7fffffffffffffffLL; b.local = -1;
  for (int i = tid; i < (int)gridDim.x; i += UPD_THREADS) {
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
  if (lane == 0) s_part[warp] = b;
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < UPD_WARPS; ++w)
      if (s_part[w].local >= 0 && better(s_part[w].score, s_part[w].id, b.score, b.id)) b = s_part[w];
    s_part[0] = b;
    const int it = st->iter;
    if (it <= P.M_cap) {
      P.hist[it] = r0;
      P.hist[(P.M_cap + 1) + it] = n_live;
      P.hist[2 * (P.M_cap + 1) + it] = *(volatile int*)&st->n_scored;
    }
    st->n_scored = 0;
    st->n_live[cur] = 0;
    st->blocks_done = 0;
    st->iter = st->iter + 1;
  }
  __syncthreads();
  b = s_part[0];
  write_record(P, b.score, b.local, r0 + NB);
  publish_record(P, st, st->iter, b.local >= 0 ? r0 + NB : 0);
}

}  // namespace gpis
