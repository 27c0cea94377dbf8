// TSDF builder: the step BEFORE the selection path -- matlab/core/auto_tsdf.m:133-262
// from the rasterised shape on: 3x3 dilate / erode masks -> five-level TSDF (:133-157), per-pixel measurement
// noise (:159-218), central-difference normals (:221), and the subsampling to the measured points (:229-236).
//
// HBM-bound byte work on images (the reference's are 25 x 25; nothing here assumes that): one thread per
// pixel, every pixel's 5 x 5 (or (2 edgeWin + 1)^2) window read through L1/L2, no shared-memory staging needed
// at these sizes.  Images are MATLAB matrices: (i, j) = (row, column) = (y, x), element i + j * rows.
#pragma once
#include "gpis_launch.cuh"
#include <stdint.h>

namespace gpis {

struct TsdfParams {
  int rows, cols;
  int edge_win, specular, grad_mode;
  double noise_scale, occlusion_scale, interior_rate, sparsity_rate, horiz_scale, vert_scale;
  double y_low[3], y_high[3], x_low[3], x_high[3];
};

// :133-157.  imdilate = window max (255, outside, wins), imerode = window min (102, inside, wins); the border
// is handled as those functions do (only pixels of the image take part); J_2d = the 5 x 5 window.
__global__ void k_tsdf_levels(const unsigned char* __restrict__ inside, double* __restrict__ tsdf, int rows, int cols) {
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= (long long)rows * cols) return;
  const int i = (int)(p % rows), j = (int)(p / rows);
  bool any_out3 = false, any_in3 = false, any_out5 = false;
  for (int dj = -2; dj <= 2; ++dj) {
    const int jj = j + dj;
    if (jj < 0 || jj >= cols) continue;
    for (int di = -2; di <= 2; ++di) {
      const int ii = i + di;
      if (ii < 0 || ii >= rows) continue;
      const bool in = inside[ii + (long long)jj * rows] != 0;
      any_out5 |= !in;
      if (di >= -1 && di <= 1 && dj >= -1 && dj <= 1) { any_out3 |= !in; any_in3 |= in; }
    }
  }
  const bool in0 = inside[p] != 0;
  double t = 0.0;
  if (any_out3 && in0) t = 0.0;            // surfaceMask     = outsideMaskDi  & insideMaskOrig
  if (any_out5 && !any_out3) t = -0.5;     // surfaceMaskIn   = outsideMaskDi2 & insideMaskDi
  if (!in0 && any_in3) t = 0.5;            // surfaceMaskOut  = outsideMaskOrig & insideMaskEr
  if (!any_out5) t = -1.0;                 // insideMask      = insideMaskDi2
  if (!any_in3) t = 1.0;                   // outsideMask     = outsideMaskEr
  tsdf[p] = t;
}

// :159-221.  u: [3][rows*cols] uniform draws (interior, specular scale, sparsity) or null (then 0.5).
__global__ void k_tsdf_noise_normals(const double* __restrict__ tsdf, const double* __restrict__ u, double* __restrict__ noise,
                                     double* __restrict__ gx, double* __restrict__ gy, unsigned char* __restrict__ valid, TsdfParams P) {
  const long long n = (long long)P.rows * P.cols;
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int i0 = (int)(p % P.rows), j0 = (int)(p / P.rows);
  const int i = i0 + 1, j = j0 + 1;                       // the reference's 1-based indices
  const double t = tsdf[p];
  bool occluded = false;
#pragma unroll
  for (int k = 0; k < 3; ++k)
    occluded |= (double)i > P.y_low[k] && (double)i <= P.y_high[k] && (double)j > P.x_low[k] && (double)j <= P.x_high[k];
  double nz;
  if (t < 0.6 && occluded) {
    nz = P.occlusion_scale;
  } else if (t < -0.5) {
    const double u0 = u ? u[p] : 0.5;
    nz = u0 > 1.0 - P.interior_rate ? P.noise_scale : P.occlusion_scale;
  } else {
    double nv = 1.0;
    if (P.specular) {
      double mn = 1e300;
      for (int jj = max(0, j0 - P.edge_win); jj <= min(P.cols - 1, j0 + P.edge_win); ++jj)
        for (int ii = max(0, i0 - P.edge_win); ii <= min(P.rows - 1, i0 + P.edge_win); ++ii)
          mn = fmin(mn, fabs(tsdf[ii + (long long)jj * P.rows]));
      if (mn < 0.6) {
        nv = u ? u[n + p] : 0.5;
        if ((u ? u[2 * n + p] : 0.5) > 1.0 - P.sparsity_rate) nv = P.occlusion_scale / P.noise_scale;
      }
    }
    const double ns = P.noise_scale;
    // (products and sums in the reference's order, no contraction: the values are compared bit for bit)
    switch (P.grad_mode) {
      case 1: nz = __dmul_rn(nv, __dadd_rn(__dmul_rn(__dmul_rn(ns, P.horiz_scale), (double)j), __dmul_rn(__dmul_rn(ns, P.vert_scale), (double)i))); break;
      case 2: nz = __dmul_rn(nv, __dadd_rn(__dmul_rn(__dmul_rn(ns, P.horiz_scale), (double)(P.cols - j + 1)), __dmul_rn(__dmul_rn(ns, P.vert_scale), (double)i))); break;
      case 3: nz = __dmul_rn(nv, __dadd_rn(__dmul_rn(__dmul_rn(ns, P.horiz_scale), (double)j), __dmul_rn(__dmul_rn(ns, P.vert_scale), (double)(P.rows - i + 1)))); break;
      case 4: nz = __dmul_rn(nv, __dadd_rn(__dmul_rn(__dmul_rn(ns, P.horiz_scale), (double)(P.cols - j + 1)), __dmul_rn(__dmul_rn(ns, P.vert_scale), (double)(P.rows - i + 1)))); break;
      default: nz = __dmul_rn(nv, ns);
    }
  }
  noise[p] = nz;
  valid[p] = nz < P.occlusion_scale ? 1 : 0;              // :229
  // imgradientxy 'CentralDifference': (f(k+1) - f(k-1)) / 2 inside, one-sided at the border
  double dx = 0.0, dy = 0.0;
  if (P.cols > 1) {
    if (j0 == 0) dx = tsdf[p + P.rows] - t;
    else if (j0 == P.cols - 1) dx = t - tsdf[p - P.rows];
    else dx = (tsdf[p + P.rows] - tsdf[p - P.rows]) / 2.0;
  }
  if (P.rows > 1) {
    if (i0 == 0) dy = tsdf[p + 1] - t;
    else if (i0 == P.rows - 1) dy = t - tsdf[p - 1];
    else dy = (tsdf[p + 1] - tsdf[p - 1]) / 2.0;
  }
  gx[p] = dx;
  gy[p] = dy;
}

// :229-236: linear (column-major) indices of the measured pixels, in order.  One block; ordered block scan per
// chunk of 1024 pixels (the images of this path are small; this is not on the hot loop).
__global__ void __launch_bounds__(1024) k_tsdf_compact(const unsigned char* __restrict__ valid, long long n, int* __restrict__ idx,
                                                       int* __restrict__ n_valid) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (long long c0 = 0; c0 < n; c0 += 1024) {
    const long long p = c0 + tid;
    const int v = p < n ? (int)valid[p] : 0;
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) { const int x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += x; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    if (v) idx[off + inc - 1] = (int)p;
    __syncthreads();
    if (tid == 1023) s_base = off + inc;
    __syncthreads();
  }
  if (tid == 0) *n_valid = s_base;
}

}  // namespace gpis
