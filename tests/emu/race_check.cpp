// TEST INFRASTRUCTURE ONLY: runs the emulated kernels (cuda_emu.h) on small inputs under ThreadSanitizer.
// In the emulation every CUDA thread is an OS thread and __syncthreads() / warp collectives are the only
// synchronisation, so a missing barrier between a shared-memory (or same-block global-memory) write and a
// read by another thread is a data race TSan reports.  Build: g++ -std=c++20 -O1 -g -fsanitize=thread.
#include "emu_kernels.cpp"

#include <cstdio>
#include <random>
#include <string>

static void small_model(int m, std::vector<double>& ax, std::vector<double>& Wt, long long ldw, std::vector<double>& alpha,
                        double ell, double noise, std::vector<double>* L_out = nullptr) {
  std::mt19937 rng(5);
  std::uniform_real_distribution<double> u(0.0, 6.0);
  ax.assign(2 * m, 0.0);
  for (auto& v : ax) v = u(rng);
  std::vector<double> L((size_t)m * m, 0.0);
  for (int i = 0; i < m; ++i)
    for (int j = 0; j <= i; ++j) {
      const double dx = ax[i] - ax[j], dy = ax[m + i] - ax[m + j];
      double s = std::exp(-0.5 * (dx * dx + dy * dy) / (ell * ell)) + (i == j ? noise : 0.0);
      for (int k = 0; k < j; ++k) s -= L[(size_t)i * m + k] * L[(size_t)j * m + k];
      L[(size_t)i * m + j] = (i == j) ? std::sqrt(s) : s / L[(size_t)j * m + j];
    }
  Wt.assign((size_t)m * ldw, 0.0);                 // Wt[k][row] = (L^-1)[row][k]
  for (int k = 0; k < m; ++k)
    for (int row = k; row < m; ++row) {
      double s = (row == k) ? 1.0 : 0.0;
      for (int c = k; c < row; ++c) s -= L[(size_t)row * m + c] * Wt[(size_t)k * ldw + c];
      Wt[(size_t)k * ldw + row] = s / L[(size_t)row * m + row];
    }
  alpha.assign(m, 0.0);
  for (int i = 0; i < m; ++i) alpha[i] = 0.1 * std::sin(1.0 + i);
  if (L_out) *L_out = L;
}

// Detector self-test: a kernel with a barrier missing between a shared write and the other threads' reads.
static void k_missing_barrier(double* out) {
  __shared__ double s_val;
  if (threadIdx.x == 0) s_val = 3.0;
  // __syncthreads() belongs here
  out[threadIdx.x] = s_val + (double)threadIdx.x;
}

int main(int argc, char** argv) {
  if (argc > 1 && std::string(argv[1]) == "--self-test") {
    std::vector<double> out(64);
    double* o = out.data();
    emu::launch_kernel(dim3(1), dim3(64), 0, k_missing_barrier, o);
    std::printf("self-test done\n");
    return 0;
  }
  // 1. posterior covariance + sampling: 20-row model, 30 query points with derivatives (90 values, two
  //    64-blocks: diagonal kernel, block row, inner update, draws)
  {
    const int m = 20, nq = 30, ns = 2;
    const long long ldw = 256, nv = 3LL * nq;
    std::vector<double> ax, Wt, alpha, q(2 * nq), z((size_t)ns * nv), mean(nv), cov((size_t)nv * nv), chol((size_t)nv * nv),
        samples((size_t)ns * nv), lp(ns);
    small_model(m, ax, Wt, ldw, alpha, 1.7, 0.05);
    for (int i = 0; i < nq; ++i) { q[i] = 0.2 * i; q[nq + i] = 6.0 - 0.19 * i; }
    for (size_t i = 0; i < z.size(); ++i) z[i] = std::cos(0.37 * (double)i);
    const double mw[3] = {0.01, -0.02, 0.0};
    long long launches = 0;
    const int rc = emu_sample_shapes(2, 0, m, ax.data(), m, Wt.data(), ldw, alpha.data(), q.data(), nq, 1, 1.7, 1.0, mw, 0.1,
                                     ns, z.data(), 1e-9, mean.data(), cov.data(), chol.data(), samples.data(), lp.data(),
                                     &launches);
    std::printf("sampling rc=%d launches=%lld logpdf=%.6f\n", rc, launches, lp[0]);
    if (rc != 0) return 1;
  }
  // 2. two-level Cholesky across an outer block boundary (280 -> 320 padded: K = 256 trailing update)
  {
    const long long n = 280;
    std::vector<double> A((size_t)n * n), T((size_t)n * n);
    for (long long i = 0; i < n; ++i)
      for (long long j = 0; j < n; ++j)
        A[i * n + j] = std::exp(-0.02 * (double)((i - j) * (i - j))) + (i == j ? 0.5 : 0.0);
    long long launches = 0;
    const int rc = emu_chol_upper(A.data(), n, 1e-9, T.data(), &launches);
    std::printf("cholesky rc=%d launches=%lld T00=%.6f\n", rc, launches, T[0]);
    if (rc != 0) return 1;
  }
  // 3. object-batched selection, all three forms of the inner loops: 2 objects on one block, two z-groups
  for (int fast = 0; fast < 3; ++fast) {
    const int dims[3] = {6, 5, 9}, B = 2, K = 8;
    const long long N = 6 * 5 * 9;
    const double org[3] = {0, 0, 0}, sp[3] = {1, 1, 1}, mw[3] = {0, 0, 0};
    std::vector<double> tsdf((size_t)B * N), mu((size_t)B * N), var((size_t)B * N);
    std::vector<unsigned char> flags((size_t)B * N);
    std::vector<long long> ids((size_t)B * K);
    std::vector<int> n_sel(B), n_model(B), err(B);
    for (int b = 0; b < B; ++b)
      for (long long j = 0; j < N; ++j) {
        const double x = (double)(j % 6) - 2.6 - 0.3 * b, y = (double)((j / 6) % 5) - 2.1, zc = (double)(j / 30) - 4.2;
        double sd = (std::sqrt(x * x + y * y + zc * zc) - 2.4) / 2.0;
        sd = sd > 1 ? 1 : (sd < -1 ? -1 : sd);
        tsdf[(size_t)b * N + j] = sd + 1e-3 * std::sin(12.9898 * (double)(j + 7 * b));
      }
    emu_batch_select(dims, org, sp, B, tsdf.data(), 17, K, 2.0, 1.0, 0.05, 0.0, 1e-2, 10.0, 0, mw, 0.0, 1, fast, ids.data(),
                     n_sel.data(), n_model.data(), err.data(), mu.data(), var.data(), flags.data());
    std::printf("batch fast=%d n_sel=%d,%d picks %lld %lld %lld\n", fast, n_sel[0], n_sel[1], ids[1], ids[2], ids[K + 1]);
    if (err[0] || err[1]) return 1;
  }
  // 4. dense prediction: factor inverse by columns (warp-synchronous), alpha, the double-buffered fused
  //    prediction kernel; 40 rows (partial k-chunk), 45 queries (two query tiles, one partial)
  {
    const int m = 40, nq = 45;
    std::vector<double> ax, Wt, alpha, L, g(m), q(2 * nq), mean(nq), var(nq), nrm(2 * nq);
    small_model(m, ax, Wt, 256, alpha, 1.7, 0.05, &L);
    for (int i = 0; i < m; ++i) g[i] = 0.3 * std::cos(0.7 * i);
    for (int i = 0; i < nq; ++i) { q[i] = 0.13 * i; q[nq + i] = 5.9 - 0.12 * i; }
    const double mw[3] = {0.01, -0.02, 0.0};
    emu_predict(2, 0, m, m, ax.data(), L.data(), m, g.data(), q.data(), nq, 1.7, 1.0, mw, 0.1, 1, mean.data(), var.data(),
                nrm.data(), nullptr);
    std::printf("predict mean0=%.6f var0=%.6f\n", mean[0], var[0]);
  }
  // 5. panel backend: k_init_candidates, k_begin, (k_append, k_update) x 7, final k_append; 2-D points with
  //    gradient observations (3 rows per point: blocked forward substitution), 150 candidates in 3 tiles
  {
    const long long N = 150;
    const int K = 8;
    std::vector<double> pts(2 * N), tsdf(N), nrm(2 * N), mu(N), var(N), L((size_t)K * 3 * K * 3), g(K * 3);
    std::vector<unsigned char> flags(N);
    std::vector<long long> ids(K + 1);
    for (long long j = 0; j < N; ++j) {
      const double x = 0.37 * (double)(j % 15) + 0.01 * std::sin(3.1 * j), y = 0.41 * (double)(j / 15) + 0.01 * std::cos(1.7 * j);
      const double rr = std::sqrt((x - 2.6) * (x - 2.6) + (y - 2.0) * (y - 2.0));
      pts[j] = x;
      pts[N + j] = y;
      double sd = (rr - 1.4) / 1.2;
      sd = sd > 1 ? 1 : (sd < -1 ? -1 : sd);
      tsdf[j] = sd + 1e-3 * std::sin(12.9898 * (double)j);
      nrm[j] = rr > 1e-9 ? (x - 2.6) / rr / 1.2 : 0.0;
      nrm[N + j] = rr > 1e-9 ? (y - 2.0) / rr / 1.2 : 0.0;
    }
    const double mw[3] = {0, 0, 0};
    int n_sel = 0, n_model = 0, err = 0;
    emu_panel_select(2, 1, N, pts.data(), tsdf.data(), nrm.data(), nullptr, K, 17, 1.8, 1.0, mw, 0.0, 0.01, 0.0, 1e-2, 10.0, 2,
                     ids.data(), &n_sel, &n_model, &err, mu.data(), var.data(), flags.data(), L.data(), g.data());
    std::printf("panel n_sel=%d n_model=%d err=%d pick1=%lld\n", n_sel, n_model, err, ids[1]);
    if (err) return 1;
  }
  std::printf("race_check done\n");
  return 0;
}
