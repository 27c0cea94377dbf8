// TEST INFRASTRUCTURE ONLY: the kernels and launch sequences of gpis_sample.cuh compiled for the CPU
// emulation in cuda_emu.h, behind a small C entry point that follows gpis_posterior_cov /
// gpis_sample_shapes in gpis_capi.cu step by step (allocate, zero, launch, copy out).
#define GPIS_EMULATE 1
#include "../../gpis_b200/csrc/gpis_sample.cuh"

#include <vector>

using namespace gpis;

extern "C" int emu_sample_shapes(int D, int grad, int R, const double* AX, int m, const double* Wt, long long ldw,
                                 const double* alpha, const double* query, long long nq, int deriv, double ell,
                                 double sf2, const double* mean_w, double mean_c, int ns, const double* z,
                                 double jitter, double* mean_out, double* cov_out, double* chol_out, double* samples,
                                 double* logpdf, long long* launches) {
  const int nt = deriv ? D + 1 : 1;
  const long long nv = nq * nt, npad = (nv + CHB - 1) / CHB * CHB;
  std::vector<double> S((size_t)npad * npad, 0.0), mean(nv), Ks((size_t)(R > 0 ? R : 1) * nv), V((size_t)(R > 0 ? R : 1) * nv);
  FullCovParams F;
  memset(&F, 0, sizeof(F));
  for (int d = 0; d < D; ++d) { F.AX[d] = AX + (size_t)d * m; F.q[d] = query + (size_t)d * nq; }
  F.nq = nq;
  F.nt = nt;
  F.R = R;
  F.inv_ell2 = 1.0 / (ell * ell);
  F.sf2 = sf2;
  for (int d = 0; d < D; ++d) F.mean_w[d] = mean_w[d];
  F.mean_c = mean_c;
  *launches = 0;
  posterior_full_device(D, grad != 0, F, Wt, ldw, alpha, R > 0 ? Ks.data() : nullptr, R > 0 ? V.data() : nullptr,
                        S.data(), npad, mean.data(), nullptr, launches);
  memcpy(mean_out, mean.data(), sizeof(double) * nv);
  for (long long i = 0; i < nv; ++i) memcpy(cov_out + i * nv, S.data() + i * npad, sizeof(double) * nv);
  int err = 0;
  chol_upper_device(S.data(), npad, nv, jitter, &err, nullptr, launches);
  if (err) return err;
  for (long long i = 0; i < nv; ++i) memcpy(chol_out + i * nv, S.data() + i * npad, sizeof(double) * nv);
  if (ns > 0) {
    std::vector<double> zt((size_t)ns * nv);
    draw_device(z, zt.data(), mean.data(), S.data(), npad, ns, nv, samples, logpdf, nullptr, launches);
  }
  return 0;
}

// chol_upper_device alone on a given symmetric matrix A [n][n]: T (upper, T'T = A + jitter I) into T_out.
extern "C" int emu_chol_upper(const double* A, long long n, double jitter, double* T_out, long long* launches) {
  const long long npad = (n + CHB - 1) / CHB * CHB;
  std::vector<double> S((size_t)npad * npad, 0.0);
  for (long long i = 0; i < n; ++i) memcpy(S.data() + i * npad, A + i * n, sizeof(double) * n);
  int err = 0;
  *launches = 0;
  chol_upper_device(S.data(), npad, n, jitter, &err, nullptr, launches);
  if (err) return err;
  for (long long i = 0; i < n; ++i) memcpy(T_out + i * n, S.data() + i * npad, sizeof(double) * n);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// gpis_batch.cuh: the one-block-per-object whole-selection kernel, driven like gpis_batch_select does
// ---------------------------------------------------------------------------------------------
#include "../../gpis_b200/csrc/gpis_batch.cuh"

extern "C" int emu_batch_select(const int* dims, const double* origin, const double* spacing, int n_objects,
                                const double* tsdf, long long first, int K, double ell, double sf2, double noise,
                                double level, double eps, double beta, int variance_mode, const double* mean_w,
                                double mean_c, int n_blocks, int fast, long long* ids, int* n_sel, int* n_model, int* err,
                                double* mu, double* var, unsigned char* flags) {
  BatchParams P;
  memset(&P, 0, sizeof(P));
  P.nx = dims[0]; P.ny = dims[1]; P.nz = dims[2];
  P.N = (long long)P.nx * P.ny * P.nz;
  P.n_objects = n_objects; P.K = K; P.first = first;
  for (int d = 0; d < 3; ++d) { P.org[d] = origin[d]; P.h[d] = spacing[d]; P.mean_w[d] = mean_w[d]; }
  P.inv_ell2 = 1.0 / (ell * ell); P.sf2 = sf2; P.noise = noise; P.level = level; P.eps = eps; P.beta = beta;
  P.variance_mode = variance_mode; P.mean_c = mean_c;
  P.tsdf = tsdf; P.mu = mu; P.var = var; P.flags = flags; P.ids = ids; P.n_sel = n_sel; P.n_model = n_model; P.err = err;
  std::vector<double> W((size_t)n_blocks * BRMAX * BRMAX, -7.0), Wt((size_t)n_blocks * BRMAX * BRMAX, -7.0);   // poisoned
  P.W = W.data(); P.Wt = Wt.data();
  launch_batch_select(P, n_blocks, fast != 0, nullptr);
  return 0;
}
