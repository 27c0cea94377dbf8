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

int main(int argc, char** argv) {
  if (argc < 2) { printf("usage: abi_check <lib> [gpu]\n"); return 2; }
  void* lib = dlopen(argv[1], RTLD_NOW);
  if (!lib) { printf("dlopen failed: %s\n", dlerror()); return 2; }
  SYM(gpis_create) SYM(gpis_destroy) SYM(gpis_set_grid_candidates) SYM(gpis_select) SYM(gpis_get_active)
  SYM(gpis_get_candidate_posterior) SYM(gpis_version)
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
  p_gpis_destroy(h);
  return (ns == K && nm == K - 1 && ids[0] == 123 && vmin > -1e-9 && vmax <= 1.0 + 1e-12) ? 0 : 1;
}
