// Full posterior covariance and posterior shape sampling.
//
//   MEAN = gp_mean(model, X*, derivatives)                          (gp_mean.m:1-12)
//   COV  = K** - K*' Q^-1 K*, symmetrised                           (gp_cov.m:1-21, full matrix)
//   draw = MEAN + chol(COV + jitter I)' z                           (sample_shapes.m:20, mvnrnd)
//   logp = -|z|^2/2 - sum log diag(chol) - d/2 log(2 pi)            (sample_shapes.m:31, mvnpdf)
//
// Values are ordered like the reference's joint vector: dimension-major, [f(all q); d1 f(all q); ..]
// (se_cov_derivative.m block layout).  Everything is fp64; the three dense contractions
// (V = W K*, COV -= V'V, the trailing updates of the blocked Cholesky and the draw itself) run on the
// fp64 tensor pipe (DMMA.8x8x4) through one k-major GEMM kernel.
//
// The launch sequences live here too (posterior_full_device, chol_upper_device, draw_device) so that
// tests/emu can compile this header with g++ against a CPU emulation of the execution model
// (-DGPIS_EMULATE, test infrastructure only: the shipped library never defines it) and check the
// indexing of every kernel and the pointer arithmetic of every launch without a GPU.
#pragma once
#include "gpis_launch.cuh"
#include "gpis_common.cuh"

namespace gpis {

#ifndef GPIS_EMULATE
// D[8x8] += A[8x4] B[4x8] on the fp64 tensor pipe (SASS DMMA.8x8x4): lane l holds A[l/4][l%4],
// B[l%4][l/4] and D[l/4][2*(l%4) + {0,1}].
__device__ __forceinline__ void s_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
#endif

struct FullCovParams {
  const double* AX[MAXD];   // active points, SoA
  const double* q[MAXD];    // query points, SoA
  long long nq;             // query points
  int nt;                   // value kinds per query point: 1 (f) or D + 1 (f, d1 f, .., dD f)
  int R;                    // model rows
  double inv_ell2, sf2;
  double mean_w[MAXD], mean_c;
};

// K*[k][c]: covariance of model row k (point i = k / NB, kind a = k % NB) with query value
// c = b * nq + q (kind b at x_q).  diff = x_q - X_i (column point minus row point, se_cov_derivative.m:45-46):
//   [f, f] = K;  [d_a, f] = diff_a K / l^2 (:47);  [f, d_b] = -diff_b K / l^2 (:63);
//   [d_a, d_b] = (delta_ab - diff_a diff_b / l^2) K / l^2 (:85-87).
template <int NB, int D>
__global__ void __launch_bounds__(256) k_cross_full(FullCovParams P, double* Ks, long long ld) {
  const long long q = blockIdx.x * 256LL + threadIdx.x;
  const int k = blockIdx.y;
  if (q >= P.nq || k >= P.R) return;
  const int i = k / NB, a = k % NB;
  double diff[D], d2 = 0.0;
#pragma unroll
  for (int d = 0; d < D; ++d) { diff[d] = P.q[d][q] - P.AX[d][i]; d2 += diff[d] * diff[d]; }
  const double kv = P.sf2 * exp(-0.5 * d2 * P.inv_ell2);
  double da = 0.0;                       // diff along the row's derivative axis
  if (NB > 1 && a > 0) {
#pragma unroll
    for (int d = 0; d < D; ++d) if (d == a - 1) da = diff[d];
  }
  double* out = Ks + (long long)k * ld + q;
  out[0] = (NB > 1 && a > 0) ? da * kv * P.inv_ell2 : kv;
  if (P.nt > 1) {
#pragma unroll
    for (int b = 1; b <= D; ++b) {
      double val;
      if (NB > 1 && a > 0) val = ((a == b ? 1.0 : 0.0) - da * diff[b - 1] * P.inv_ell2) * kv * P.inv_ell2;
      else val = -diff[b - 1] * kv * P.inv_ell2;
      out[(long long)b * P.nq] = val;
    }
  }
}

// K**[(a, p)][(b, q)] with beta = 0 (gp_cov.m:6): the same block formulas between query points,
// diff = x_q - x_p.
template <int D>
__global__ void __launch_bounds__(256) k_prior_full(FullCovParams P, double* S, long long ld) {
  const long long q = blockIdx.x * 256LL + threadIdx.x;
  const long long p = blockIdx.y;
  if (q >= P.nq || p >= P.nq) return;
  double diff[D], d2 = 0.0;
#pragma unroll
  for (int d = 0; d < D; ++d) { diff[d] = P.q[d][q] - P.q[d][p]; d2 += diff[d] * diff[d]; }
  const double kv = P.sf2 * exp(-0.5 * d2 * P.inv_ell2);
  S[p * ld + q] = kv;
  if (P.nt > 1) {
#pragma unroll
    for (int b = 1; b <= D; ++b) S[p * ld + b * P.nq + q] = -diff[b - 1] * kv * P.inv_ell2;
#pragma unroll
    for (int a = 1; a <= D; ++a) {
      double* row = S + (a * P.nq + p) * ld;
      row[q] = diff[a - 1] * kv * P.inv_ell2;
#pragma unroll
      for (int b = 1; b <= D; ++b)
        row[b * P.nq + q] = ((a == b ? 1.0 : 0.0) - diff[a - 1] * diff[b - 1] * P.inv_ell2) * kv * P.inv_ell2;
    }
  }
}

// MEAN[c] = m(c) + sum_k alpha[k] K*[k][c]; m = w'x + c0 for function values, w_d for d_d f
// (linear_mean_derivative.m:8,13).  One thread per value, fixed summation order.
template <int D>
__global__ void __launch_bounds__(256) k_full_mean(FullCovParams P, const double* __restrict__ Ks, long long ld,
                                                   const double* __restrict__ alpha, double* mean) {
  const long long c = blockIdx.x * 256LL + threadIdx.x;
  const long long nv = P.nq * P.nt;
  if (c >= nv) return;
  const int b = (int)(c / P.nq);
  const long long q = c - (long long)b * P.nq;
  double m = 0.0;
  if (b == 0) {
    m = P.mean_c;
#pragma unroll
    for (int d = 0; d < D; ++d) m += P.mean_w[d] * P.q[d][q];
  } else {
#pragma unroll
    for (int d = 0; d < D; ++d) if (d == b - 1) m = P.mean_w[d];
  }
  double s = 0.0;
  for (int k = 0; k < P.R; ++k) s += alpha[k] * Ks[(long long)k * ld + c];
  mean[c] = m + s;
}

// C[M x N] = beta C + alpha sum_k A[k][m] B[k][n]: both operands k-major (A is [K][lda], B is [K][ldb]).
// CTA tile 64 x 128, 8 warps as 2 x 4, each a 32 x 32 warp tile of 4 x 4 DMMA.8x8x4; k-chunks of 16
// staged through registers into padded shared rows (same fragment layout as k_predict_f64).
// upper != 0: tiles lying entirely below the diagonal (n < m) are skipped (symmetric rank-k update of an
// upper-stored matrix).  All bounds are guarded, no alignment is assumed.
constexpr int SGM = 64, SGN = 128, SGK = 16, SGLA = SGM + 4, SGLB = SGN + 4;

__global__ void __launch_bounds__(256) k_gemm_tn_f64(const double* A, long long lda, const double* B, long long ldb,
                                                     double* C, long long ldc, int M, int N, int K,
                                                     double alpha, double beta, int upper) {
  __shared__ double As[SGK * SGLA];
  __shared__ double Bs[SGK * SGLB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m0 = blockIdx.y * SGM, n0 = blockIdx.x * SGN;
  if (upper && n0 + SGN <= m0) return;
  const int wm = warp & 1, wn = warp >> 1;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  double ra[4], rb[8];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256, kk = e >> 6, m = e & 63;
      const int gk = k0 + kk, gm = m0 + m;
      ra[i] = (gk < K && gm < M) ? A[(long long)gk * lda + gm] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int e = tid + i * 256, kk = e >> 7, n = e & 127;
      const int gk = k0 + kk, gn = n0 + n;
      rb[i] = (gk < K && gn < N) ? B[(long long)gk * ldb + gn] : 0.0;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int i = 0; i < 4; ++i) { const int e = tid + i * 256; As[(e >> 6) * SGLA + (e & 63)] = ra[i]; }
#pragma unroll
    for (int i = 0; i < 8; ++i) { const int e = tid + i * 256; Bs[(e >> 7) * SGLB + (e & 127)] = rb[i]; }
  };
  const int nch = (K + SGK - 1) / SGK;
  fetch(0);
  stash();
  __syncthreads();
  const double* Ab = As + wm * 32 + (lane >> 2);
  const double* Bb = Bs + wn * 32 + (lane >> 2);
  for (int c = 0; c < nch; ++c) {
    const bool more = c + 1 < nch;
    if (more) fetch((c + 1) * SGK);
#pragma unroll
    for (int k4 = 0; k4 < SGK / 4; ++k4) {
      const int kb = k4 * 4 + (lane & 3);
      double af[4], bf[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = Ab[kb * SGLA + i * 8];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = Bb[kb * SGLB + j * 8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s_dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    __syncthreads();
    if (more) {
      stash();
      __syncthreads();
    }
  }
  // fragment element (i, j, e): row i*8 + lane/4, column j*8 + 2*(lane%4) + e of the warp tile
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + wm * 32 + i * 8 + (lane >> 2);
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = n0 + wn * 32 + j * 8 + 2 * (lane & 3) + e;
        if (col >= N) continue;
        double* cp = C + (long long)row * ldc + col;
        const double v = alpha * acc[i][j][e];
        *cp = (beta != 0.0) ? beta * (*cp) + v : v;
      }
  }
}

// S = (S + S') / 2 (gp_cov.m:19-20).
__global__ void __launch_bounds__(256) k_symmetrise(double* S, long long ld, int n) {
  const int j = blockIdx.x * 16 + (threadIdx.x & 15);
  const int i = blockIdx.y * 16 + (threadIdx.x >> 4);
  if (i < j && j < n) {
    const double v = 0.5 * (S[(long long)i * ld + j] + S[(long long)j * ld + i]);
    S[(long long)i * ld + j] = v;
    S[(long long)j * ld + i] = v;
  }
}

// diagonal: + jitter inside the matrix, 1 on the padding (the padded matrix is blockdiag(COV, I), so
// every block of the factorisation is a full 64 x 64 block).
__global__ void __launch_bounds__(256) k_prepare_diag(double* S, long long ld, int n, int npad, double jitter) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < n) S[(long long)i * ld + i] += jitter;
  else if (i < npad) S[(long long)i * ld + i] = 1.0;
}

constexpr int CHB = 64;            // Cholesky block
constexpr int CHLD = CHB + 1;

// Upper Cholesky (T'T = S) of the 64 x 64 diagonal block at offset o.  The block lives in registers:
// thread (ti, tj) of a 16 x 16 arrangement owns the 4 x 4 sub-block at rows 4ti.., columns 4tj.. (threads
// below the diagonal idle).  Per step k the owners of row k publish it through a double-buffered
// shared row, everyone scales what it needs by 1/sqrt(pivot) and applies the rank-1 update to its own
// registers: one barrier per step.  A non-positive pivot records its (1-based) global index in *err.
__global__ void __launch_bounds__(256) k_chol_diag(double* S, long long ld, int o, int* err) {
  __shared__ double rowbuf[2][CHB];
  const int tid = threadIdx.x, ti = tid >> 4, tj = tid & 15;
  const bool mine = tj >= ti;
  double a[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) a[r][c] = mine ? S[(long long)(o + 4 * ti + r) * ld + o + 4 * tj + c] : 0.0;
  for (int k = 0; k < CHB; ++k) {
    const int kb = k >> 2, kr = k & 3;
    double* rb = rowbuf[k & 1];
    if (ti == kb && mine) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (r == kr) {
#pragma unroll
          for (int c = 0; c < 4; ++c) rb[4 * tj + c] = a[r][c];
        }
    }
    __syncthreads();
    const double piv = rb[k];
    const double dg = sqrt(piv);
    const double inv = 1.0 / dg;
    if (!(piv > 0.0) && tid == 0) atomicCAS(err, 0, o + k + 1);
    if (!mine) continue;                                  // idle threads only keep the barrier count
    double ri[4], cj[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) ri[r] = rb[4 * ti + r] * inv;
#pragma unroll
    for (int c = 0; c < 4; ++c) cj[c] = rb[4 * tj + c] * inv;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int i = 4 * ti + r, j = 4 * tj + c;
        if (i == k && j >= k) a[r][c] = (j == k) ? dg : cj[c];          // the finished row k of the factor
        else if (i > k && j >= i) a[r][c] -= ri[r] * cj[c];
      }
  }
  if (mine) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (4 * tj + c >= 4 * ti + r) S[(long long)(o + 4 * ti + r) * ld + o + 4 * tj + c] = a[r][c];
  }
}

// Block row of the factor right of the diagonal block: T_jj' X = S[o:o+64, o+64:n], one thread per
// column holding its 64 entries in registers.  Column-oriented forward substitution: once x[k] is final
// (times the reciprocal pivot) it is eliminated from all later entries -- 63-k independent FMAs per step,
// the diagonal block's row k broadcast from shared memory.
__global__ void __launch_bounds__(128) k_chol_panel(double* S, long long ld, int o, int n) {
  __shared__ double T[CHB * CHLD];
  __shared__ double rinv[CHB];
  const int tid = threadIdx.x;
  for (int e = tid; e < CHB * CHB; e += 128) {
    const int i = e >> 6, j = e & 63;
    T[i * CHLD + j] = S[(long long)(o + i) * ld + o + j];     // only j >= i entries are read below
  }
  __syncthreads();
  if (tid < CHB) rinv[tid] = 1.0 / T[tid * CHLD + tid];
  __syncthreads();
  const long long m = (long long)o + CHB + blockIdx.x * 128LL + tid;
  if (m >= n) return;
  double x[CHB];
#pragma unroll
  for (int k = 0; k < CHB; ++k) x[k] = S[(long long)(o + k) * ld + m];
#pragma unroll
  for (int k = 0; k < CHB; ++k) {
    x[k] *= rinv[k];
#pragma unroll
    for (int j = k + 1; j < CHB; ++j) x[j] -= T[k * CHLD + j] * x[k];
  }
#pragma unroll
  for (int k = 0; k < CHB; ++k) S[(long long)(o + k) * ld + m] = x[k];
}

__global__ void __launch_bounds__(256) k_zero_lower(double* S, long long ld, int n) {
  const int j = blockIdx.x * 16 + (threadIdx.x & 15);
  const int i = blockIdx.y * 16 + (threadIdx.x >> 4);
  if (i < n && j < i) S[(long long)i * ld + j] = 0.0;
}

// Zt[k][s] = z[s][k] (the draws arrive one sample per row, the GEMM wants them k-major).
__global__ void __launch_bounds__(256) k_draws_kmajor(const double* __restrict__ z, int ns, long long nv, double* Zt) {
  const long long e = blockIdx.x * 256LL + threadIdx.x;
  if (e >= nv * ns) return;
  const long long k = e / ns;
  const int s = (int)(e - k * ns);
  Zt[e] = z[(long long)s * nv + k];
}

// out[s][c] = mean[c]
__global__ void __launch_bounds__(256) k_rows_from_mean(const double* __restrict__ mean, int ns, long long nv, double* out) {
  const long long e = blockIdx.x * 256LL + threadIdx.x;
  if (e >= nv * ns) return;
  out[e] = mean[e % nv];
}

// logp[s] = -|z_s|^2 / 2 - sum_k log T[k][k] - nv/2 log(2 pi); one CTA per sample, fixed reduction order.
__global__ void __launch_bounds__(256) k_logpdf(const double* __restrict__ z, long long nv, const double* __restrict__ T,
                                                long long ld, double* logp) {
  __shared__ double red[256];
  const int tid = threadIdx.x;
  const double* zs = z + (long long)blockIdx.x * nv;
  double a = 0.0;
  for (long long k = tid; k < nv; k += 256) a += 0.5 * zs[k] * zs[k] + log(T[k * ld + k]);
  red[tid] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) red[tid] += red[tid + o];
    __syncthreads();
  }
  if (tid == 0) logp[blockIdx.x] = -red[0] - 0.5 * (double)nv * 1.8378770664093454835606594728112;
}

// ---------------------------------------------------------------------------------------------
// launch sequences (host)
// ---------------------------------------------------------------------------------------------
inline void launch_gemm_tn(const double* A, long long lda, const double* B, long long ldb, double* Cm, long long ldc,
                           int M, int N, int K, double alpha, double beta, int upper, cudaStream_t s) {
  const dim3 g((unsigned)((N + SGN - 1) / SGN), (unsigned)((M + SGM - 1) / SGM));
  GPIS_LAUNCH(k_gemm_tn_f64, g, dim3(256), s, A, lda, B, ldb, Cm, ldc, M, N, K, alpha, beta, upper);
}

template <int NB, int D>
inline void launch_full_cov_kernels(const FullCovParams& F, double* Ks, long long ldk, double* S, long long lds,
                                    const double* alpha, double* mean, cudaStream_t s) {
  const dim3 gp((unsigned)((F.nq + 255) / 256), (unsigned)F.nq);
  const dim3 gc((unsigned)((F.nq + 255) / 256), (unsigned)(F.R > 0 ? F.R : 1));
  const dim3 gm((unsigned)((F.nq * F.nt + 255) / 256));
  auto kp = k_prior_full<D>;
  auto kc = k_cross_full<NB, D>;
  auto km = k_full_mean<D>;
  GPIS_LAUNCH(kp, gp, dim3(256), s, F, S, lds);
  if (F.R > 0) GPIS_LAUNCH(kc, gc, dim3(256), s, F, Ks, ldk);
  GPIS_LAUNCH(km, gm, dim3(256), s, F, Ks, ldk, alpha, mean);
}

// MEAN [nv] and COV (the nv x nv corner of S [npad][npad], S zeroed by the caller) of the joint vector.
// Wt[k][row] = (L^-1)[row][k] (zero above the diagonal), alpha = Q^-1 (y - m); Ks, V: [R][nv] scratch.
inline void posterior_full_device(int D, bool grad, const FullCovParams& F, const double* Wt, long long ldw,
                                  const double* alpha, double* Ks, double* V, double* S, long long npad, double* mean,
                                  cudaStream_t s, long long* launches) {
  const long long nv = F.nq * F.nt;
  switch (D * 2 + (grad ? 1 : 0)) {
    case 2: launch_full_cov_kernels<1, 1>(F, Ks, nv, S, npad, alpha, mean, s); break;
    case 3: launch_full_cov_kernels<2, 1>(F, Ks, nv, S, npad, alpha, mean, s); break;
    case 4: launch_full_cov_kernels<1, 2>(F, Ks, nv, S, npad, alpha, mean, s); break;
    case 5: launch_full_cov_kernels<3, 2>(F, Ks, nv, S, npad, alpha, mean, s); break;
    case 6: launch_full_cov_kernels<1, 3>(F, Ks, nv, S, npad, alpha, mean, s); break;
    default: launch_full_cov_kernels<4, 3>(F, Ks, nv, S, npad, alpha, mean, s); break;
  }
  *launches += F.R > 0 ? 3 : 2;
  if (F.R > 0) {
    // V = W K*,  COV = K** - V'V
    launch_gemm_tn(Wt, ldw, Ks, nv, V, nv, F.R, (int)nv, F.R, 1.0, 0.0, 0, s);
    launch_gemm_tn(V, nv, V, nv, S, npad, (int)nv, (int)nv, F.R, -1.0, 1.0, 0, s);
    *launches += 2;
  }
  const unsigned t16 = (unsigned)((nv + 15) / 16);
  GPIS_LAUNCH(k_symmetrise, dim3(t16, t16), dim3(256), s, S, npad, (int)nv);
  *launches += 1;
}

// Blocked right-looking upper Cholesky of blockdiag(COV + jitter I, I) in place: T'T = COV + jitter I in
// the upper triangle, zeros below.  *d_err (zeroed by the caller) receives the first bad pivot.
// Two levels of blocking: 64-wide block rows (diagonal block in one CTA, block row by forward
// substitution) update only the rest of their 256-wide outer block row; the trailing matrix is updated
// once per outer block with K = 256, which divides its HBM traffic (the cost that bounds a K = 64 update)
// by four.
constexpr int CHO = 256;           // outer block
inline void chol_upper_device(double* S, long long npad, long long nv, double jitter, int* d_err, cudaStream_t s,
                              long long* launches) {
  GPIS_LAUNCH(k_prepare_diag, dim3((unsigned)((npad + 255) / 256)), dim3(256), s, S, npad, (int)nv, (int)npad, jitter);
  *launches += 1;
  for (long long O = 0; O < npad; O += CHO) {
    const long long oe = (O + CHO < npad) ? O + CHO : npad;
    for (long long o = O; o < oe; o += CHB) {
      GPIS_LAUNCH(k_chol_diag, dim3(1), dim3(256), s, S, npad, (int)o, d_err);
      *launches += 1;
      const long long rest = npad - o - CHB;
      if (rest <= 0) continue;
      GPIS_LAUNCH(k_chol_panel, dim3((unsigned)((rest + 127) / 128)), dim3(128), s, S, npad, (int)o, (int)npad);
      *launches += 1;
      const long long mi = oe - (o + CHB);         // rows of this outer block row still to be factored
      if (mi > 0) {
        const double* pan = S + o * npad + o + CHB;
        launch_gemm_tn(pan, npad, pan, npad, S + (o + CHB) * npad + o + CHB, npad, (int)mi, (int)rest, CHB, -1.0, 1.0,
                       1, s);
        *launches += 1;
      }
    }
    const long long rem = npad - oe;
    if (rem > 0) {
      const double* pan = S + O * npad + oe;       // rows O..oe of the factor, columns oe..npad
      launch_gemm_tn(pan, npad, pan, npad, S + oe * npad + oe, npad, (int)rem, (int)rem, (int)(oe - O), -1.0, 1.0, 1, s);
      *launches += 1;
    }
  }
  const unsigned t16 = (unsigned)((npad + 15) / 16);
  GPIS_LAUNCH(k_zero_lower, dim3(t16, t16), dim3(256), s, S, npad, (int)npad);
  *launches += 1;
}

// out[s][c] = MEAN[c] + sum_k z[s][k] T[k][c];  logp[s].  zt: [nv][ns] scratch.
inline void draw_device(const double* z, double* zt, const double* mean, const double* T, long long npad, int ns,
                        long long nv, double* out, double* logp, cudaStream_t s, long long* launches) {
  const dim3 ge((unsigned)(((long long)ns * nv + 255) / 256));
  GPIS_LAUNCH(k_draws_kmajor, ge, dim3(256), s, z, ns, nv, zt);
  GPIS_LAUNCH(k_rows_from_mean, ge, dim3(256), s, mean, ns, nv, out);
  launch_gemm_tn(zt, ns, T, npad, out, nv, ns, (int)nv, (int)nv, 1.0, 1.0, 0, s);
  GPIS_LAUNCH(k_logpdf, dim3((unsigned)ns), dim3(256), s, z, nv, T, npad, logp);
  *launches += 4;
}

}  // namespace gpis
