// Dense posterior prediction: W = L^-1 (explicit, transposed), alpha = W' g, and the fused
// covariance-tile + triangular-GEMM + column-norm kernel.
//
//   mean_q = m(x_q) + sum_k alpha_k K*[k, q]                      (gp_mean.m:11)
//   var_q  = sf2 - | W K*[:, q] |^2                               (diag of gp_cov.m:16-17)
//   nrm_q  = w + sum_k alpha_k cov(d f(x_q), row k)               (predict_2d_grid.m:12-13)
//
// K* (rows of the active set x query points) is generated tile by tile in shared memory and
// never stored; the reference materialises it and runs cublasStrsm on it
// (gpu_active_set_selector.cpp:858-871).
#pragma once
#include "gpis_launch.cuh"
#include "gpis_common.cuh"

namespace gpis {

constexpr int PRED_THREADS = 256;
constexpr int PQ = 32;     // queries per CTA
constexpr int PSB = 256;   // factor rows per super-block (register tile: 8 rows x 4 queries per thread)
constexpr int PKC = 16;    // k-chunk

struct PredictParams {
  const double* Wt;        // [Rk_pad][ldw]: Wt[k][row] = (L^-1)[row][k], zero-padded
  long long ldw;           // padded row count (multiple of PSB)
  const double* alpha;     // [R]
  const double* AX[MAXD];  // active points SoA
  const double* q[MAXD];   // query SoA
  long long nq;
  double* mean;
  double* var;
  double* nrm[MAXD];       // or null
  int R;                   // rows in the model
  double inv_ell2, sf2;
  double mean_w[MAXD], mean_c;
};

// One warp per column k of W: forward substitution L w = e_k, stored as row k of Wt.
__global__ void __launch_bounds__(256) k_invert_factor(const double* __restrict__ L, int ldl, int R,
                                                       double* Wt, long long ldw) {
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (k >= R) return;
  double* w = Wt + (long long)k * ldw;
  for (int row = k; row < R; ++row) {
    const double* lrow = L + (long long)row * ldl;
    double p = 0.0;
    for (int s = k + lane; s < row; s += 32) p += lrow[s] * __ldcg(w + s);
    for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
    const double val = ((row == k ? 1.0 : 0.0) - p) / lrow[row];
    if (lane == 0) w[row] = val;
    __syncwarp();
  }
}

// alpha_k = sum_{row >= k} Wt[k][row] g[row]   (alpha = L^-T g)
__global__ void __launch_bounds__(256) k_alpha(const double* __restrict__ Wt, long long ldw,
                                               const double* __restrict__ g, int R, double* alpha) {
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (k >= R) return;
  const double* w = Wt + (long long)k * ldw;
  double p = 0.0;
  for (int s = k + lane; s < R; s += 32) p += w[s] * g[s];
  for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
  if (lane == 0) alpha[k] = p;
}

#ifndef GPIS_EMULATE   // tests/emu supplies synchronous stand-ins (test infrastructure only)
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
#endif

// Covariance of model row k (point i = k / NB, kind a = k % NB, at X_i) with f(x_q), plus what
// the mean's gradient needs.  diff = x_q - X_i.
template <int NB, int D>
__device__ __forceinline__ double kstar(const PredictParams& P, int k, const double* xq, bool want_grad,
                                        double* gradc /*[D] cov(d_d f(x_q), row k)*/) {
  const int i = k / NB, a = k % NB;
  double diff[D], d2 = 0.0;
#pragma unroll
  for (int d = 0; d < D; ++d) { diff[d] = xq[d] - P.AX[d][i]; d2 += diff[d] * diff[d]; }
  const double kv = P.sf2 * exp(-0.5 * d2 * P.inv_ell2);
  double val = kv;
  if (NB > 1 && a > 0) val = diff[a - 1] * kv * P.inv_ell2;          // [d_a f(X_i), f(x_q)]
  if (want_grad) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      if (NB == 1 || a == 0) gradc[d] = -diff[d] * kv * P.inv_ell2;  // [d_d f(x_q), f(X_i)]
      else gradc[d] = ((d == a - 1 ? 1.0 : 0.0) - diff[d] * diff[a - 1] * P.inv_ell2) * kv * P.inv_ell2;
    }
  }
  return val;
}

// fp64 tensor pipe: D[256 rows x 32 queries] += Wt-tile' * K*-tile, DMMA.8x8x4; 8 warps, each a
// 32 x 32 warp tile (4 x 4 DMMA tiles).  Shared rows are padded by 4 doubles so that the four
// k-rows of a fragment load fall on disjoint bank groups.
constexpr int PLDA = PSB + 4;
constexpr int PLDB = PQ + 4;

#ifndef GPIS_EMULATE
__device__ __forceinline__ void p_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
#endif

template <int NB, int D>
__global__ void __launch_bounds__(PRED_THREADS, 2) k_predict_f64(PredictParams P) {
  GPIS_DYN_SMEM(double, smem);
  double* As = smem;                                  // [2][PKC][PLDA]
  double* Bs = smem + 2 * PKC * PLDA;                 // [2][PKC][PLDB]
  double* Red = Bs + 2 * PKC * PLDB;                  // [8][PQ] reductions
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long q0 = (long long)blockIdx.x * PQ;
  const int R = P.R;

  // this thread generates K* elements (kk = tid/32 + 8*i, query tid%32)
  const int gq = tid & 31, gk = tid >> 5;
  const long long myq = q0 + gq;
  const bool qvalid = myq < P.nq;
  double xq[D];
#pragma unroll
  for (int d = 0; d < D; ++d) xq[d] = qvalid ? P.q[d][myq] : 0.0;
  const bool want_nrm = P.nrm[0] != nullptr;
  double mu_part = 0.0, nr_part[D];
#pragma unroll
  for (int d = 0; d < D; ++d) nr_part[d] = 0.0;
  double vsq[4][2];
#pragma unroll
  for (int j = 0; j < 4; ++j) vsq[j][0] = vsq[j][1] = 0.0;

  const int n_sb = (R + PSB - 1) / PSB;
  for (int sb = 0; sb < n_sb; ++sb) {
    const bool last_sb = (sb == n_sb - 1);
    const int k_end = min(R, (sb + 1) * PSB);
    const int n_chunks = (k_end + PKC - 1) / PKC;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto issue_A = [&](int c, int buf) {
      // tile rows k = c*PKC .. +15, columns sb*PSB .. +255 of Wt: 16 x 2 KB
      const double* src = P.Wt + (long long)(c * PKC) * P.ldw + (long long)sb * PSB;
      double* dst = As + buf * PKC * PLDA;
#pragma unroll
      for (int i = 0; i < (PKC * PSB * 8 / 16) / PRED_THREADS; ++i) {
        const int e = tid + i * PRED_THREADS;        // 16-byte element
        const int kk = e / (PSB / 2), col2 = e % (PSB / 2);
        cp_async16(dst + kk * PLDA + col2 * 2, src + (long long)kk * P.ldw + col2 * 2);
      }
      cp_async_commit();
    };
    auto gen_B = [&](int c, double out[2]) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int k = c * PKC + gk + 8 * i;
        double val = 0.0;
        if (k < R && qvalid) {
          double gc[D];
          val = kstar<NB, D>(P, k, xq, last_sb && want_nrm, gc);
          if (last_sb) {
            const double al = P.alpha[k];
            mu_part += al * val;
            if (want_nrm) {
#pragma unroll
              for (int d = 0; d < D; ++d) nr_part[d] += al * gc[d];
            }
          }
        }
        out[i] = val;
      }
    };

    double bnext[2];
    issue_A(0, 0);
    gen_B(0, bnext);
    Bs[gk * PLDB + gq] = bnext[0];
    Bs[(gk + 8) * PLDB + gq] = bnext[1];
    cp_async_wait_all();
    __syncthreads();
    for (int c = 0; c < n_chunks; ++c) {
      const int buf = c & 1;
      const bool more = c + 1 < n_chunks;
      if (more) { issue_A(c + 1, buf ^ 1); gen_B(c + 1, bnext); }
      const double* Ab = As + buf * PKC * PLDA + warp * 32 + (lane >> 2);
      const double* Bb = Bs + buf * PKC * PLDB + (lane >> 2);
      // W = L^-1 is lower triangular: a chunk of columns k only reaches rows >= k, so in the super-block on the
      // diagonal a warp skips the chunks beyond its last row (their tiles of Wt are zero) and leaves the pipe to the
      // other warps / the SM's other CTA
      if (c * PKC <= sb * PSB + warp * 32 + 31)
#pragma unroll
      for (int k4 = 0; k4 < PKC / 4; ++k4) {
        const int kb = k4 * 4 + (lane & 3);
        double af[4], bf[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) af[i] = Ab[kb * PLDA + i * 8];
#pragma unroll
        for (int j = 0; j < 4; ++j) bf[j] = Bb[kb * PLDB + j * 8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) p_dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      }
      if (more) {
        Bs[(buf ^ 1) * PKC * PLDB + gk * PLDB + gq] = bnext[0];
        Bs[(buf ^ 1) * PKC * PLDB + (gk + 8) * PLDB + gq] = bnext[1];
        cp_async_wait_all();
      }
      __syncthreads();
    }
    // |v|^2 over this super-block's rows: fragment element (i, j, e) is row i*8 + lane/4 of the
    // warp's 32 rows, query j*8 + 2*(lane%4) + e
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        vsq[j][0] += acc[i][j][0] * acc[i][j][0];
        vsq[j][1] += acc[i][j][1] * acc[i][j][1];
      }
  }

  // reductions: |v|^2 over the 8 row lanes of a warp (fixed shuffle tree) and the 8 warps; mean /
  // gradient parts over the 8 generator rows
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      double v = vsq[j][e];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if ((lane >> 2) == 0) Red[warp * PQ + j * 8 + 2 * (lane & 3) + e] = v;
    }
  __syncthreads();
  double vs = 0.0;
  if (tid < PQ) for (int i = 0; i < 8; ++i) vs += Red[i * PQ + tid];
  __syncthreads();
  Red[gk * PQ + gq] = mu_part;
  __syncthreads();
  double mu = 0.0;
  if (tid < PQ) for (int i = 0; i < 8; ++i) mu += Red[i * PQ + tid];
  double nr[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    nr[d] = 0.0;
    if (want_nrm) {
      __syncthreads();
      Red[gk * PQ + gq] = nr_part[d];
      __syncthreads();
      if (tid < PQ) for (int i = 0; i < 8; ++i) nr[d] += Red[i * PQ + tid];
    }
  }
  if (tid < PQ && qvalid) {
    double m0 = P.mean_c;
#pragma unroll
    for (int d = 0; d < D; ++d) m0 += P.mean_w[d] * xq[d];
    P.mean[myq] = m0 + mu;
    P.var[myq] = P.sf2 - vs;
    if (want_nrm) {
#pragma unroll
      for (int d = 0; d < D; ++d) P.nrm[d][myq] = P.mean_w[d] + nr[d];
    }
  }
}

}  // namespace gpis
