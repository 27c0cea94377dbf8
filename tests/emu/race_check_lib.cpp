// TEST INFRASTRUCTURE ONLY: the whole C ABI (gpis_capi.cu under -DGPIS_EMULATE) on tiny lattice selections
// under ThreadSanitizer -- the lattice backend's kernels (one-pass and two-pass append, the warp-specialised
// bulk-copy / mbarrier / DMMA update GEMM, the epilogue) with every CUDA thread an OS thread.
// Build: g++ -std=c++20 -O1 -g -fsanitize=thread -x c++ -DGPIS_EMULATE.
#include "../../gpis_b200/csrc/gpis_capi.cu"

#include <cstdio>

static int run(int D, const int* dims, int grad, int K, long long first) {
  long long N = 1;
  for (int d = 0; d < D; ++d) N *= dims[d];
  std::vector<double> tsdf(N), nrm((size_t)D * N, 0.0), mu(N), var(N);
  for (long long j = 0; j < N; ++j) {
    long long rem = j;
    double x[3] = {0, 0, 0}, r2 = 0.0;
    for (int d = 0; d < D; ++d) { x[d] = (double)(rem % dims[d]) - 0.47 * dims[d]; rem /= dims[d]; r2 += x[d] * x[d]; }
    const double rr = std::sqrt(r2);
    double sd = (rr - 0.3 * dims[0]) / 1.5;
    const bool in_band = sd > -1 && sd < 1;
    sd = sd > 1 ? 1 : (sd < -1 ? -1 : sd);
    tsdf[j] = sd + 1e-3 * std::sin(12.9898 * (double)j);
    for (int d = 0; d < D; ++d) nrm[(size_t)d * N + j] = (in_band && rr > 1e-9) ? x[d] / rr / 1.5 : 0.0;
  }
  gpis_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.struct_size = sizeof(cfg);
  cfg.dim = D;
  cfg.num_candidates = N;
  cfg.max_active = K;
  cfg.use_gradients = grad;
  cfg.world_size = 1;
  cfg.device = -1;
  cfg.ell = 1.8;
  cfg.sf2 = 1.0;
  cfg.noise = 0.03;
  cfg.eps = 1e-2;
  cfg.beta = 10.0;
  gpis_handle h = nullptr;
  if (gpis_create(&h, &cfg) != GPIS_OK) return 1;
  const double org[3] = {0, 0, 0}, sp[3] = {1, 1, 1};
  int rc = gpis_set_grid_candidates(h, dims, org, sp, 0, D == 3 ? dims[2] : 1, tsdf.data(), grad ? nrm.data() : nullptr, nullptr, nullptr);
  if (rc == GPIS_OK) rc = gpis_select(h, first, K, nullptr);
  std::vector<int64_t> ids(K + 1);
  int32_t n_sel = 0, n_model = 0;
  if (rc == GPIS_OK) rc = gpis_get_active(h, ids.data(), &n_sel, &n_model, nullptr);
  if (rc == GPIS_OK) rc = gpis_get_candidate_posterior(h, mu.data(), var.data(), nullptr);
  std::printf("D=%d grad=%d rc=%d n_sel=%d n_model=%d pick1=%lld mu0=%.6f\n", D, grad, rc, n_sel, n_model, (long long)ids[1], mu[0]);
  gpis_destroy(h);
  return rc;
}

int main() {
  const int d3[3] = {6, 5, 3}, d2[3] = {6, 5, 1};
  if (run(3, d3, 0, 6, 17)) return 1;        // function values: one-pass append
  if (run(2, d2, 1, 4, 9)) return 1;         // gradient observations: two-pass append, 3 rows per point
  std::printf("race_check_lib done\n");
  return 0;
}
