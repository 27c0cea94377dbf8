/* Plain-C consumer of include/gpis_b200.h: proves the header is valid C99, the struct layout the
 * Python binding assumes, and that the library refuses to run without a GPU (no CPU fallback).
 * With a GPU (argv[1] == "gpu") it runs a tiny lattice selection entirely through the C ABI with
 * HOST buffers, the way a C/C++/MATLAB caller of the reference would. */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "gpis_b200.h"

#define SYM(name) name##_t p_##name = (name##_t)dlsym(lib, #name); if (!p_##name) { printf("missing %s\n", #name); return 2; }
typedef gpis_status (*gpis_create_t)(gpis_handle*, const gpis_config*);
typedef gpis_status (*gpis_destroy_t)(gpis_handle);
typedef gpis_status (*gpis_set_grid_candidates_t)(gpis_handle, const int32_t*, const double*, const double*, int32_t, int32_t,
                                                  const double*, const double*, const double*, void*);
typedef gpis_status (*gpis_select_t)(gpis_handle, int64_t, int32_t, void*);
typedef gpis_status (*gpis_get_active_t)(gpis_handle, int64_t*, int32_t*, int32_t*, void*);
typedef gpis_status (*gpis_get_candidate_posterior_t)(gpis_handle, double*, double*, void*);
typedef const char* (*gpis_version_t)(void);
typedef gpis_status (*gpis_set_grid_candidates_f32_t)(gpis_handle, const int32_t*, const double*, const double*, int32_t, int32_t,
                                                      const float*, const float*, const float*, void*);
typedef gpis_status (*gpis_get_active_f32_t)(gpis_handle, int32_t, float*, float*, int32_t*, int32_t*, void*);
typedef gpis_status (*gpis_build_tsdf_t)(const uint8_t*, int32_t, int32_t, const gpis_tsdf_params*, const double*, double*, double*,
                                         double*, int32_t*, int32_t*, int32_t, void*);

int main(int argc, char** argv) {
  if (argc < 2) { printf("usage: abi_check <lib> [gpu]\n"); return 2; }
  void* lib = dlopen(argv[1], RTLD_NOW);
  if (!lib) { printf("dlopen failed: %s\n", dlerror()); return 2; }
  SYM(gpis_create) SYM(gpis_destroy) SYM(gpis_set_grid_candidates) SYM(gpis_select) SYM(gpis_get_active)
  SYM(gpis_get_candidate_posterior) SYM(gpis_version) SYM(gpis_set_grid_candidates_f32) SYM(gpis_get_active_f32) SYM(gpis_build_tsdf)
  printf("sizeof(gpis_config)=%zu version=%s\n", sizeof(gpis_config), p_gpis_version());
  enum { G = 10, N = G * G * G, K = 30 };
  gpis_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.struct_size = (int32_t)sizeof(cfg);
  cfg.dim = 3; cfg.num_candidates = N; cfg.max_active = K; cfg.world_size = 1; cfg.device = -1;
  cfg.ell = 2.0; cfg.sf2 = 1.0; cfg.noise = 0.05; cfg.level = 0.0; cfg.eps = 1e-2; cfg.beta = 10.0;
  gpis_handle h = NULL;
  gpis_status st = p_gpis_create(&h, &cfg);
  if (argc < 3) {                      /* CPU box: must refuse, never fall back */
    printf("create status=%d\n", st);
    /* a caller compiled against the version-0.1 header passes the shorter struct: accepted (same answer);
     * a size that was never published is rejected as such */
    cfg.struct_size = GPIS_CONFIG_SIZE_V1;
    gpis_status st1 = p_gpis_create(&h, &cfg);
    cfg.struct_size = 40;
    gpis_status st2 = p_gpis_create(&h, &cfg);
    printf("create(v1 struct)=%d create(bad size)=%d\n", st1, st2);
    return (st == GPIS_ERR_NO_DEVICE && st1 == GPIS_ERR_NO_DEVICE && st2 == GPIS_ERR_INVALID) ? 0 : 1;
  }
  if (st != GPIS_OK) { printf("create failed %d\n", st); return 1; }
  static double tsdf[N], mu[N], var[N];
  for (int k = 0; k < G; ++k) for (int j = 0; j < G; ++j) for (int i = 0; i < G; ++i) {
    double d = sqrt((i - 4.6) * (i - 4.6) + (j - 4.3) * (j - 4.3) + (k - 4.8) * (k - 4.8)) - 3.1;
    tsdf[i + G * (j + G * k)] = d / 2 > 1 ? 1 : (d / 2 < -1 ? -1 : d / 2);
  }
  int32_t dims[3] = {G, G, G};
  double org[3] = {0, 0, 0}, sp[3] = {1, 1, 1};
  if (p_gpis_set_grid_candidates(h, dims, org, sp, 0, G, tsdf, NULL, NULL, NULL) != GPIS_OK) return 1;
  if (p_gpis_select(h, 123, K, NULL) != GPIS_OK) return 1;
  int64_t ids[K + 1]; int32_t ns = 0, nm = 0;
  if (p_gpis_get_active(h, ids, &ns, &nm, NULL) != GPIS_OK) return 1;
  if (p_gpis_get_candidate_posterior(h, mu, var, NULL) != GPIS_OK) return 1;
  double vmin = 1e9, vmax = -1e9;
  for (int i = 0; i < N; ++i) { if (var[i] < vmin) vmin = var[i]; if (var[i] > vmax) vmax = var[i]; }
  printf("selected=%d in_model=%d first=%lld var=[%.4f, %.4f]\n", ns, nm, (long long)ids[0], vmin, vmax);
  /* the reference's own fp32 signatures (gpu_active_set_selector.hpp:41-46): float TSDF in, float active set out */
  static float tsdf32[N], ain[3 * K], atg[K];
  for (int i = 0; i < N; ++i) tsdf32[i] = (float)tsdf[i];
  if (p_gpis_set_grid_candidates_f32(h, dims, org, sp, 0, G, tsdf32, NULL, NULL, NULL) != GPIS_OK) return 1;
  if (p_gpis_select(h, 123, K, NULL) != GPIS_OK) return 1;
  int32_t ns32 = 0, nm32 = 0;
  if (p_gpis_get_active_f32(h, K, ain, atg, &ns32, &nm32, NULL) != GPIS_OK) return 1;
  int f32_ok = ns32 == K && nm32 == K - 1 && ain[0] == 3.0f && ain[K] == 2.0f && ain[2 * K] == 1.0f && atg[0] == tsdf32[123];
  printf("f32: selected=%d in_model=%d first point=(%g, %g, %g) target=%g ok=%d\n", ns32, nm32, ain[0], ain[K], ain[2 * K], atg[0], f32_ok);
  p_gpis_destroy(h);
  /* auto_tsdf.m:133-262 on a 12 x 9 image (rows x columns, column-major) holding a filled rectangle */
  enum { R = 12, C = 9 };
  static uint8_t inside[R * C];
  static double t2[R * C], nrm2[2 * R * C], nz2[R * C];
  static int32_t vidx[R * C];
  for (int j = 0; j < C; ++j) for (int i = 0; i < R; ++i) inside[i + j * R] = (i >= 3 && i < 9 && j >= 2 && j < 7);
  gpis_tsdf_params tp;
  memset(&tp, 0, sizeof(tp));
  tp.struct_size = (int32_t)sizeof(tp);
  tp.noise_scale = 0.05; tp.occlusion_scale = 1000.0; tp.interior_rate = 0.4; tp.horiz_scale = 1.0; tp.vert_scale = 1.0;
  for (int k = 0; k < 3; ++k) { tp.occ_y_low[k] = R; tp.occ_x_low[k] = C; }      /* no occlusion window */
  int32_t nvalid = 0;
  st = p_gpis_build_tsdf(inside, R, C, &tp, NULL, t2, nrm2, nz2, vidx, &nvalid, -1, NULL);
  /* centre of the rectangle is two pixels from the border: -1; its border pixels: 0; the image corner: +1 */
  int tsdf_ok = st == GPIS_OK && t2[5 + 4 * R] == -1.0 && t2[3 + 2 * R] == 0.0 && t2[0] == 1.0 && t2[2 + 2 * R] == 0.5 &&
                nrm2[2 + 4 * R] == 0.0 && nvalid > 0 && nvalid <= R * C && nz2[0] == 0.05;
  printf("tsdf: status=%d valid=%d centre=%g border=%g corner=%g ok=%d\n", st, nvalid, t2[5 + 4 * R], t2[3 + 2 * R], t2[0], tsdf_ok);
  return (ns == K && nm == K - 1 && ids[0] == 123 && vmin > -1e-9 && vmax <= 1.0 + 1e-12 && f32_ok && tsdf_ok) ? 0 : 1;
}
