// TEST INFRASTRUCTURE ONLY: kernel headers of gpis_b200/csrc compiled for the CPU emulation in cuda_emu.h,
// each behind a small C entry point that sets up buffers and launches the way gpis_capi.cu does
// (gpis_sample.cuh shares its launch sequences with the shipped library; the others replicate them).
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
  launch_batch_select(P, n_blocks, fast, nullptr);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// gpis_predict.cuh: W' = L^-T by columns, alpha = W' g, fused covariance-tile + triangular GEMM +
// column-norm prediction kernel -- launched as ensure_inverse / gpis_predict in gpis_capi.cu do
// ---------------------------------------------------------------------------------------------
#include "../../gpis_b200/csrc/gpis_predict.cuh"

template <int NB, int D>
static void emu_predict_launch(const PredictParams& Q, long long nq) {
  const size_t smem = (2 * PKC * PLDA + 2 * PKC * PLDB + 8 * PQ) * sizeof(double);      // PRED_SMEM of gpis_capi.cu
  auto k = k_predict_f64<NB, D>;
  GPIS_LAUNCH_SMEM(k, dim3((unsigned)((nq + PQ - 1) / PQ)), dim3(PRED_THREADS), smem, nullptr, Q);
}

extern "C" int emu_predict(int D, int grad, int R, int m, const double* AX, const double* L, int ldl, const double* g,
                           const double* query, long long nq, double ell, double sf2, const double* mean_w,
                           double mean_c, int want_nrm, double* mean, double* var, double* nrm, double* alpha_out) {
  const long long ldw = ((long long)R + PSB - 1) / PSB * PSB, rk_pad = ((long long)R + PKC - 1) / PKC * PKC;
  std::vector<double> Wt((size_t)ldw * rk_pad, 0.0), alpha(R > 0 ? R : 1, 0.0);
  if (R > 0) {
    const int blocks = (R * 32 + 255) / 256;
    GPIS_LAUNCH(k_invert_factor, dim3((unsigned)blocks), dim3(256), nullptr, L, ldl, R, Wt.data(), ldw);
    GPIS_LAUNCH(k_alpha, dim3((unsigned)blocks), dim3(256), nullptr, Wt.data(), ldw, g, R, alpha.data());
  }
  PredictParams Q;
  memset(&Q, 0, sizeof(Q));
  Q.Wt = Wt.data();
  Q.ldw = ldw;
  Q.alpha = alpha.data();
  for (int d = 0; d < D; ++d) { Q.AX[d] = AX + (size_t)d * m; Q.q[d] = query + (size_t)d * nq; }
  Q.nq = nq;
  Q.mean = mean;
  Q.var = var;
  for (int d = 0; d < D; ++d) Q.nrm[d] = want_nrm ? nrm + (size_t)d * nq : nullptr;
  Q.R = R;
  Q.inv_ell2 = 1.0 / (ell * ell);
  Q.sf2 = sf2;
  for (int d = 0; d < D; ++d) Q.mean_w[d] = mean_w[d];
  Q.mean_c = mean_c;
  switch (D * 2 + (grad ? 1 : 0)) {
    case 2: emu_predict_launch<1, 1>(Q, nq); break;
    case 3: emu_predict_launch<2, 1>(Q, nq); break;
    case 4: emu_predict_launch<1, 2>(Q, nq); break;
    case 5: emu_predict_launch<3, 2>(Q, nq); break;
    case 6: emu_predict_launch<1, 3>(Q, nq); break;
    default: emu_predict_launch<4, 3>(Q, nq); break;
  }
  if (alpha_out) memcpy(alpha_out, alpha.data(), sizeof(double) * R);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// gpis_select.cuh: the panel backend's selection loop (arbitrary candidate points, factor panels in
// memory) -- buffers and launches as gpis_create / gpis_set_candidates / gpis_select set them up
// ---------------------------------------------------------------------------------------------
#include "../../gpis_b200/csrc/gpis_select.cuh"

template <int NB, int D>
static void emu_panel_steps(const EngineParams& P, int K, int n_ctas, size_t upd_smem) {
  auto ka = k_append<NB, D>;
  auto ku = k_update<NB, D>;
  for (int it = 0; it < K - 1; ++it) {
    GPIS_LAUNCH(ka, dim3(1), dim3(APP_THREADS), nullptr, P, 0, 0);
    GPIS_LAUNCH_SMEM(ku, dim3((unsigned)n_ctas), dim3(UPD_THREADS), upd_smem, nullptr, P);
  }
  GPIS_LAUNCH(ka, dim3(1), dim3(APP_THREADS), nullptr, P, 0, 1);      // gpis_select_end: record the last pick
}

extern "C" int emu_panel_select(int D, int grad, long long N, const double* pts, const double* tsdf,
                                const double* normals, const double* noise, int K, long long first, double ell,
                                double sf2, const double* mean_w, double mean_c, double noise_scalar, double level,
                                double eps, double beta, int n_ctas, long long* ids, int* n_sel, int* n_model, int* err,
                                double* mu, double* var, unsigned char* flags, double* L_out, double* g_out) {
  const int NB = grad ? D + 1 : 1;
  EngineParams P;
  memset(&P, 0, sizeof(P));
  P.N = N;
  P.N_total = N;
  P.n_tiles = (int)((N + TW - 1) / TW);
  P.M_cap = K;
  P.R_cap = K * NB;
  P.tile_stride = (long long)P.R_cap * TW;
  P.world = 1;
  P.inv_ell2 = 1.0 / (ell * ell);
  P.sf2 = sf2;
  P.level = level;
  P.eps = eps;
  P.beta = beta;
  P.noise_scalar = noise_scalar;
  for (int d = 0; d < D; ++d) P.mean_w[d] = mean_w[d];
  P.mean_c = mean_c;
  P.pay_stride = PAY_HDR + P.R_cap;
  std::vector<double> V((size_t)P.n_tiles * P.tile_stride, 0.0), amb(N), L((size_t)P.R_cap * P.R_cap, 0.0), g(P.R_cap, 0.0),
      ax((size_t)MAXD * P.M_cap, 0.0), pay(P.pay_stride, 0.0);
  std::vector<int> live0(P.n_tiles + 1), live1(P.n_tiles + 1), hist(3 * ((size_t)P.M_cap + 1), 0);
  std::vector<long long> sel(P.M_cap + 1, -1);
  std::vector<BlockBest> partials(n_ctas);
  DevState st;
  memset(&st, 0, sizeof(st));
  for (int d = 0; d < D; ++d) {
    P.cx[d] = pts + (size_t)d * N;
    P.nrm[d] = normals ? normals + (size_t)d * N : nullptr;
    P.AX[d] = ax.data() + (size_t)d * P.M_cap;
  }
  P.tsdf = tsdf;
  P.noise = noise;
  P.V = V.data();
  P.mu = mu;
  P.var = var;
  P.amb = amb.data();
  P.flags = flags;
  P.live[0] = live0.data();
  P.live[1] = live1.data();
  P.L = L.data();
  P.g = g.data();
  P.sel_ids = sel.data();
  P.st = &st;
  P.hist = hist.data();
  P.partials = partials.data();
  P.pay_send = pay.data();
  P.pay_recv = pay.data();
  const size_t max_ell = (size_t)96 * 1024 / (sizeof(double) * NB);
  P.ell_chunk = (int)std::min<size_t>((size_t)P.R_cap, max_ell);
  const size_t upd_smem = sizeof(double) * NB * (size_t)P.ell_chunk;
  const long long n = std::max<long long>(P.N, P.n_tiles);
  GPIS_LAUNCH(k_init_candidates, dim3((unsigned)((n + 255) / 256)), dim3(256), nullptr, P);
  GPIS_LAUNCH(k_begin, dim3(1), dim3(32), nullptr, P, first, 1u);
  switch (D * 2 + (grad ? 1 : 0)) {
    case 2: emu_panel_steps<1, 1>(P, K, n_ctas, upd_smem); break;
    case 3: emu_panel_steps<2, 1>(P, K, n_ctas, upd_smem); break;
    case 4: emu_panel_steps<1, 2>(P, K, n_ctas, upd_smem); break;
    case 5: emu_panel_steps<3, 2>(P, K, n_ctas, upd_smem); break;
    case 6: emu_panel_steps<1, 3>(P, K, n_ctas, upd_smem); break;
    default: emu_panel_steps<4, 3>(P, K, n_ctas, upd_smem); break;
  }
  *n_sel = st.n_sel;
  *n_model = st.m;
  *err = st.err;
  for (int i = 0; i < st.n_sel; ++i) ids[i] = sel[i];
  if (L_out) memcpy(L_out, L.data(), sizeof(double) * L.size());
  if (g_out) memcpy(g_out, g.data(), sizeof(double) * g.size());
  return 0;
}
