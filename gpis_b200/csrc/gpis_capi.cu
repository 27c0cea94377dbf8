// C ABI of the GPIS engine (include/gpis_b200.h): host-side orchestration only.
#ifdef GPIS_EMULATE   // tests/emu: this file compiled by g++ against a CPU emulation of the CUDA runtime and
#include "cuda_rt_emu.h"   // execution model -- test infrastructure only, the shipped library never defines it
#else
#include <cuda_runtime.h>
#endif

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/gpis_b200.h"
#include "gpis_common.cuh"
#include "gpis_grid.cuh"
#ifndef GPIS_EMULATE
#include "gpis_grid_tf32.cuh"
#else   // tcgen05 / TMEM are not emulated: GPIS_PREC_TF32 is rejected by gpis_create in that build
namespace gpis {
constexpr int TF32_THREADS = 32, TKC = 32, TZ = 2;
constexpr size_t TF32_SMEM = 0;
template <int NB>
void k_grid_gemm_tf32(EngineParams) {}
constexpr size_t PERSIST_SMEM = 0;      // in-kernel grid barriers are not emulated either: the per-step path runs
constexpr int PTHREADS = 32, PLL_WORLD_MAX = 16;
template <int D, bool CUBE>
void k_grid_persist(EngineParams, int) {}
}  // namespace gpis
#endif
#include "gpis_grid_persist.cuh"
#include "gpis_predict.cuh"
#include "gpis_batch.cuh"
#include "gpis_sample.cuh"
#include "gpis_select.cuh"
#include "gpis_tsdf.cuh"

using namespace gpis;

struct gpis_engine {
  gpis_config cfg;
  int NB, D;
  int device;
  int num_sms;
  EngineParams P;
  std::vector<void*> allocs;
  // device buffers
  double* d_pts = nullptr;      // [D][N]
  double* d_tsdf = nullptr;
  double* d_nrm = nullptr;      // [D][N]
  double* d_noise = nullptr;
  long long* d_ids = nullptr;
  double* d_Wt = nullptr;
  double* d_alpha = nullptr;
  long long ldw = 0, rk_pad = 0;
  bool have_candidates = false, have_noise = false, have_ids = false;
  int backend = 0;              // 0 none, 1 panel (arbitrary points), 2 grid (lattice, separable)
  int grid_ctas = 0, solve_ctas = 0;
  bool fused_solve = false, onepass = false, solve_ring = false;
  bool persist = false;            // gpis_select runs the whole greedy loop in one cooperative kernel
  bool self_peer = false;          // world_size 1: the exchange region is this engine's own (flags release the CTAs)
  unsigned long long* d_phase_clk = nullptr;
  size_t tile_cnt_cap = 0, rlist_cap = 0, tt_cap[MAXD] = {0, 0, 0}, upref_cap = 0, rtz_cap = 0;
  int* d_upref = nullptr;
  bool persist_cube = false;       // persistent kernel on 16 x 16 x 16 tiles (3-D lattices)
  bool step_lists = false;         // per-step path sums per-tile row lists (gradient / large fp64 models on a lattice)
  bool persist_grid_ok = false;    // the lattice of the last gpis_set_grid_candidates fits the persistent kernel
  double persist_ms[16] = {0};     // CTA 0's lap times of the last persistent selection (PCLK_* slots), ms
  int solve1_ctas = 0;
  size_t solve_smem = 0;
  size_t table_bytes[MAXD] = {0, 0, 0};
  size_t bimg_floats = 0;
  size_t w_bytes = 0, gws_bytes = 0, partials_cap = 0, state_cap = 0, grid_slots = 0;
  bool selecting = false;
  int inverse_rows = -1;        // rows Wt/alpha are valid for
  int upd_grid = 0;
  size_t upd_smem = 0;
  cudaGraphExec_t step_graph = nullptr;
  cudaStream_t graph_stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double select_ms = 0.0, predict_ms = 0.0, update_ms = 0.0, append_ms = 0.0, main_ms = 0.0, epilogue_ms = 0.0;
  double cov_ms = 0.0, chol_ms = 0.0, draw_ms = 0.0;   // last gpis_posterior_cov / gpis_sample_shapes
  cudaEvent_t ev2 = nullptr, ev3 = nullptr;
  long long update_launches = 0;
  bool profiling = false;
  std::vector<cudaEvent_t> prof_events;
  long long launches = 0;
  std::string err;
  DevState host_state;
  unsigned epoch = 0;
  void* xchg_base = nullptr;       // peer exchange region (cudaMalloc, IPC-exported)
  size_t xchg_bytes = 0;
  std::vector<void*> ipc_opened;
  // scratch blocks of the prediction / sampling calls, kept across calls (a 25^3 draw needs ~2 GB of them and
  // cudaMalloc / cudaFree of those dominated its host-side time); released by gpis_destroy
  struct PoolBlock { void* p; size_t bytes; bool busy; };
  std::vector<PoolBlock> pool;
};

namespace {

const char* kVersion = "gpis_b200 0.1 (sm_100a)";

thread_local std::string g_last_error;      // failures without a handle (gpis_create, gpis_batch_select)

gpis_status fail(gpis_engine* e, gpis_status s, const std::string& msg) {
  if (e) e->err = msg;
  else g_last_error = msg;
  return s;
}

// Every entry point runs on the engine's device and leaves the caller's current device as it found it.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int dev) {
    if (dev >= 0 && cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};
#define GUARD(e) DeviceGuard _guard((e)->device)

// gpis_config grew after version 0.1: accept any size from the v1 prefix up, zero-fill the rest
bool read_config(const gpis_config* in, gpis_config* out) {
  if (!in || in->struct_size < (int32_t)GPIS_CONFIG_SIZE_V1 || in->struct_size > (int32_t)sizeof(gpis_config)) return false;
  memset(out, 0, sizeof(*out));
  memcpy(out, in, (size_t)in->struct_size);
  out->struct_size = (int32_t)sizeof(gpis_config);
  return true;
}

#ifdef GPIS_EMULATE
template <class K>
cudaError_t launch_cooperative(K, unsigned, unsigned, size_t, cudaStream_t, void**) { return cudaErrorNotSupported; }
#else
template <class K>
cudaError_t launch_cooperative(K kern, unsigned grid, unsigned block, size_t smem, cudaStream_t s, void** args) {
  cudaLaunchConfig_t lc;
  memset(&lc, 0, sizeof(lc));
  lc.gridDim = dim3(grid);
  lc.blockDim = dim3(block);
  lc.dynamicSmemBytes = smem;
  lc.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  lc.attrs = at;
  lc.numAttrs = 1;
  return cudaLaunchKernelExC(&lc, (const void*)kern, args);
}
#endif

#define CU(call)                                                                         \
  do {                                                                                   \
    cudaError_t _e = (call);                                                             \
    if (_e != cudaSuccess)                                                               \
      return fail(e, _e == cudaErrorMemoryAllocation ? GPIS_ERR_NOMEM : GPIS_ERR_CUDA,   \
                  std::string(#call) + ": " + cudaGetErrorString(_e));                   \
  } while (0)

template <class T>
cudaError_t dalloc(gpis_engine* e, T** p, size_t n) {
  void* q = nullptr;
  cudaError_t r = cudaMalloc(&q, n * sizeof(T) > 0 ? n * sizeof(T) : 16);
  if (r == cudaSuccess) { e->allocs.push_back(q); *p = static_cast<T*>(q); }
  return r;
}

struct Dispatch {
  void (*append)(EngineParams, int, int);
  void (*update)(EngineParams);
  void (*predict)(PredictParams);
  void (*g_point)(EngineParams, int);
  void (*g_solve)(EngineParams);
  void (*g_solve_fused)(EngineParams);
  void (*g_solve1)(EngineParams);
  void (*g_solve2)(EngineParams);
  void (*g_wrow)(EngineParams);
  void (*g_gemm)(EngineParams);
  void (*g_gemm_lists)(EngineParams);
  void (*g_lists)(EngineParams);
  void (*g_gemm_tf32)(EngineParams);
  void (*g_epi)(EngineParams);
  void (*g_persist)(EngineParams, int);
  void (*g_persist_cube)(EngineParams, int);
};

template <int D, bool CUBE>
auto persist_kernel() -> void (*)(EngineParams, int) {
  if constexpr (!CUBE || D == 3) return k_grid_persist<D, CUBE>;
  else return nullptr;
}

template <int NB, int D>
Dispatch make_dispatch() {
  return Dispatch{k_append<NB, D>, k_update<NB, D>, k_predict_f64<NB, D>,
                  k_grid_point<NB, D>, k_grid_solve<NB, D, false>, k_grid_solve<NB, D, true>, k_grid_solve1<D>, k_grid_solve2<D>, k_grid_wrow<NB>, k_grid_gemm<NB, false>, k_grid_gemm<NB, true>, k_grid_lists<NB, D>, k_grid_gemm_tf32<NB>, k_grid_epilogue<NB>, persist_kernel<D, false>(), persist_kernel<D, true>()};
}

Dispatch get_dispatch(int D, bool grad) {
  switch (D * 2 + (grad ? 1 : 0)) {
    case 2: return make_dispatch<1, 1>();
    case 3: return make_dispatch<2, 1>();
    case 4: return make_dispatch<1, 2>();
    case 5: return make_dispatch<3, 2>();
    case 6: return make_dispatch<1, 3>();
    default: return make_dispatch<4, 3>();
  }
}

constexpr size_t PRED_SMEM = (2 * PKC * PLDA + 2 * PKC * PLDB + 8 * PQ) * sizeof(double);

gpis_status launch_append(gpis_engine* e, cudaStream_t s, int external, int finalize_only) {
  Dispatch d = get_dispatch(e->D, e->NB > 1);
  if (e->backend == 2 && !external) {
    if (finalize_only) {
      GPIS_LAUNCH_SMEM(d.g_point, 1, 1024, 0, s, e->P, 1);
      e->launches++;
    } else {
      if (e->onepass) {
        if (e->solve_ring) {      // solve + grid barrier + new row of W in ONE launch (one CTA per SM)
          void* args[] = {(void*)&e->P};
          CU(launch_cooperative(d.g_solve2, (unsigned)(e->solve1_ctas + 1), GSOLVE_THREADS, GS2_SMEM, s, args));
          e->launches += 1;
        } else {
          GPIS_LAUNCH_SMEM(d.g_solve1, e->solve1_ctas + 1, GSOLVE_THREADS, sizeof(double) * GS1_RMAX, s, e->P);
          GPIS_LAUNCH_SMEM(k_grid_wrow1, e->num_sms, GSOLVE_THREADS, 0, s, e->P, e->solve1_ctas);
          e->launches += 2;
        }
        if (e->P.upref) {      // the new point's row joins the tiles' row lists; unit prefix of the update GEMM
          GPIS_LAUNCH_SMEM(d.g_lists, 1, 1024, 0, s, e->P);
          e->launches += 1;
        }
        CU(cudaGetLastError());
        return GPIS_OK;
      }
      if (e->fused_solve) {
        GPIS_LAUNCH_SMEM(d.g_solve_fused, e->solve_ctas + 1, GSOLVE_THREADS, e->solve_smem, s, e->P);
      } else {
        GPIS_LAUNCH_SMEM(d.g_point, 1, 1024, 0, s, e->P, 0);
        GPIS_LAUNCH_SMEM(d.g_solve, e->solve_ctas, GSOLVE_THREADS, 0, s, e->P);
        e->launches++;
      }
      GPIS_LAUNCH_SMEM(d.g_wrow, e->solve_ctas, GSOLVE_THREADS, 0, s, e->P);
      e->launches += 2;
      if (e->P.upref) {      // the new point's rows join the tiles' row lists; unit prefix of the update GEMM
        GPIS_LAUNCH_SMEM(d.g_lists, 1, 1024, 0, s, e->P);
        e->launches += 1;
      }
    }
  } else {
    GPIS_LAUNCH_SMEM(d.append, 1, APP_THREADS, 0, s, e->P, external, finalize_only);
    e->launches++;
  }
  CU(cudaGetLastError());
  return GPIS_OK;
}

gpis_status launch_update(gpis_engine* e, cudaStream_t s, cudaEvent_t mid = nullptr) {
  Dispatch d = get_dispatch(e->D, e->NB > 1);
  if (e->backend == 2) {
    if (e->cfg.precision == GPIS_PREC_TF32) GPIS_LAUNCH_SMEM(d.g_gemm_tf32, e->P.gemm_ctas, TF32_THREADS, TF32_SMEM, s, e->P);
    else if (e->P.upref) GPIS_LAUNCH_SMEM(d.g_gemm_lists, e->P.gemm_ctas, GEMM_THREADS, GEMM_SMEM, s, e->P);
    else GPIS_LAUNCH_SMEM(d.g_gemm, e->P.gemm_ctas, GEMM_THREADS, GEMM_SMEM, s, e->P);
    if (mid) cudaEventRecord(mid, s);
    GPIS_LAUNCH_SMEM(d.g_epi, GEPI_CTAS_PER_TILE * e->grid_ctas, GEPI_THREADS, 0, s, e->P);
    e->launches += 2;
  } else {
    GPIS_LAUNCH_SMEM(d.update, e->upd_grid, UPD_THREADS, e->upd_smem, s, e->P);
    e->launches++;
  }
  CU(cudaGetLastError());
  return GPIS_OK;
}

gpis_status read_state(gpis_engine* e, cudaStream_t s) {
  CU(cudaMemcpyAsync(&e->host_state, e->P.st, sizeof(DevState), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  if (e->host_state.err == GPIS_ERR_NOT_PD)
    return fail(e, GPIS_ERR_NOT_PD, "covariance block lost positive definiteness while appending a point");
  if (e->host_state.err == GPIS_ERR_TIMEOUT)
    return fail(e, GPIS_ERR_TIMEOUT, "an in-kernel wait (a peer's exchange record or a grid barrier) exceeded spin_timeout_ms");
  return GPIS_OK;
}

// W = L^-1 (transposed) and alpha for the current model; cached per row count.
gpis_status ensure_inverse(gpis_engine* e, cudaStream_t s) {
  gpis_status rc = read_state(e, s);
  if (rc != GPIS_OK) return rc;
  const int R = e->host_state.r;
  if (e->inverse_rows == R) return GPIS_OK;
  const bool have_w = e->backend == 2 && e->selecting;    // the grid backend maintains W (and W' unless one-pass)
  const bool have_wt = have_w && !e->onepass;
  if (!have_wt) CU(cudaMemsetAsync(e->d_Wt, 0, sizeof(double) * e->rk_pad * e->ldw, s));
  if (R > 0) {
    const int blocks = (R * 32 + 255) / 256;
    if (have_w && !have_wt) {
      dim3 tg((R + 31) / 32, (R + 31) / 32);
      GPIS_LAUNCH_SMEM(k_transpose_lower, tg, dim3(32, 8), 0, s, e->P.W, e->d_Wt, e->ldw, R);
      CU(cudaGetLastError());
      e->launches++;
    } else if (!have_wt) {
      GPIS_LAUNCH_SMEM(k_invert_factor, blocks, 256, 0, s, e->P.L, e->P.R_cap, R, e->d_Wt, e->ldw);
      CU(cudaGetLastError());
      e->launches++;
    }
    GPIS_LAUNCH_SMEM(k_alpha, blocks, 256, 0, s, e->d_Wt, e->ldw, e->P.g, R, e->d_alpha);
    CU(cudaGetLastError());
    e->launches += 1;
  }
  e->inverse_rows = R;
  return GPIS_OK;
}

}  // namespace

extern "C" {

const char* gpis_version(void) { return kVersion; }

const char* gpis_status_string(gpis_status s) {
  switch (s) {
    case GPIS_OK: return "ok";
    case GPIS_ERR_INVALID: return "invalid argument";
    case GPIS_ERR_CUDA: return "CUDA error";
    case GPIS_ERR_NOMEM: return "out of device memory";
    case GPIS_ERR_STATE: return "call out of order";
    case GPIS_ERR_NOT_PD: return "matrix not positive definite";
    case GPIS_ERR_NO_DEVICE: return "no CUDA device (there is no CPU path)";
    case GPIS_ERR_TIMEOUT: return "in-kernel wait timed out";
    default: return "unknown status";
  }
}

const char* gpis_last_error(gpis_handle h) { return h ? h->err.c_str() : g_last_error.c_str(); }

gpis_status gpis_create(gpis_handle* out, const gpis_config* cfg_in) {
  if (!out || !cfg_in) return GPIS_ERR_INVALID;
  *out = nullptr;
  gpis_config cfg_v, *cfg = &cfg_v;
  if (!read_config(cfg_in, cfg)) return fail(nullptr, GPIS_ERR_INVALID, "gpis_config.struct_size is not a known size");
  if (cfg->loop_mode < 0 || cfg->loop_mode > GPIS_LOOP_STEPS || cfg->append_mode < 0 || cfg->append_mode > GPIS_APPEND_TWO_PASS ||
      cfg->spin_timeout_ms < 0)
    return fail(nullptr, GPIS_ERR_INVALID, "loop_mode / append_mode / spin_timeout_ms out of range");
  if (cfg->dim < 1 || cfg->dim > MAXD || cfg->num_candidates < 0 || cfg->max_active < 1) return GPIS_ERR_INVALID;
  if (!(cfg->ell > 0.0) || !(cfg->sf2 > 0.0) || cfg->world_size < 1 || cfg->rank < 0 || cfg->rank >= cfg->world_size)
    return GPIS_ERR_INVALID;
  if (cfg->precision != GPIS_PREC_FP64 && cfg->precision != GPIS_PREC_TF32) return GPIS_ERR_INVALID;
#ifdef GPIS_EMULATE
  if (cfg->precision == GPIS_PREC_TF32) return GPIS_ERR_INVALID;      // tcgen05 is not emulated
#endif
  if (cfg->num_candidates > 0x7fffffffLL * TW / 2) return GPIS_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return GPIS_ERR_NO_DEVICE; }
  gpis_engine* e = new (std::nothrow) gpis_engine();
  if (!e) return GPIS_ERR_NOMEM;
  e->cfg = *cfg;
  e->D = cfg->dim;
  e->NB = cfg->use_gradients ? cfg->dim + 1 : 1;
  gpis_status rc = GPIS_OK;
  int caller_device = -1;
  cudaGetDevice(&caller_device);
  auto body = [&]() -> gpis_status {
    if (cfg->device >= 0) CU(cudaSetDevice(cfg->device));
    CU(cudaGetDevice(&e->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, e->device));
    if (prop.major < 10) return fail(e, GPIS_ERR_NO_DEVICE, "device is not sm_100 class");
    e->num_sms = prop.multiProcessorCount;
    EngineParams& P = e->P;
    memset(&P, 0, sizeof(P));
    const long long N = cfg->num_candidates;
    const int D = e->D, NB = e->NB;
    P.N = N;
    P.N_total = cfg->total_candidates > 0 ? cfg->total_candidates : N;
    P.id_base = cfg->id_base;
    P.n_tiles = (int)((N + TW - 1) / TW);
    P.M_cap = cfg->max_active;
    P.R_cap = cfg->max_active * NB;
    P.tile_stride = (long long)P.R_cap * TW;
    P.rank = cfg->rank;
    P.world = cfg->world_size;
    P.inv_ell2 = 1.0 / (cfg->ell * cfg->ell);
    P.sf2 = cfg->sf2;
    P.level = cfg->level;
    P.eps = cfg->eps;
    P.beta = cfg->beta;
    P.beta_delta = cfg->beta_delta;
    P.noise_scalar = cfg->noise;
    for (int d = 0; d < MAXD; ++d) P.mean_w[d] = d < D ? cfg->mean_w[d] : 0.0;
    P.mean_c = cfg->mean_c;
    P.beta_mode = cfg->beta_mode;
    P.variance_mode = cfg->variance_mode;
    P.pay_stride = PAY_HDR + P.R_cap;

    const size_t Nn = (size_t)(N > 0 ? N : 1);
    CU(dalloc(e, &e->d_pts, Nn * D));
    CU(dalloc(e, &e->d_tsdf, Nn));
    CU(dalloc(e, &e->d_nrm, Nn * D));
    CU(dalloc(e, &e->d_noise, Nn));
    CU(dalloc(e, &e->d_ids, Nn));
    CU(dalloc(e, &P.mu, Nn));
    CU(dalloc(e, &P.var, Nn));
    CU(dalloc(e, &P.amb, Nn));
    CU(dalloc(e, &P.flags, Nn));
    CU(dalloc(e, &P.live[0], (size_t)P.n_tiles + 1));
    CU(dalloc(e, &P.live[1], (size_t)P.n_tiles + 1));
    CU(dalloc(e, &P.L, (size_t)P.R_cap * P.R_cap));
    CU(dalloc(e, &P.g, (size_t)P.R_cap));
    for (int d = 0; d < D; ++d) CU(dalloc(e, &P.AX[d], (size_t)P.M_cap));
    CU(dalloc(e, &P.sel_ids, (size_t)P.M_cap + 1));
    CU(dalloc(e, &P.st, 1));
    CU(dalloc(e, &P.hist, 3 * ((size_t)P.M_cap + 1)));
    CU(cudaMemset(P.hist, 0, sizeof(int) * 3 * ((size_t)P.M_cap + 1)));
    CU(cudaMemset(P.st, 0, sizeof(DevState)));
    CU(cudaMemset(P.L, 0, sizeof(double) * (size_t)P.R_cap * P.R_cap));
    CU(dalloc(e, &P.pay_send, (size_t)P.pay_stride));
    if (P.world > 1) CU(dalloc(e, &P.pay_recv, (size_t)P.pay_stride * P.world));
    else P.pay_recv = P.pay_send;
    CU(cudaMemset(P.pay_send, 0, sizeof(double) * P.pay_stride));
    e->ldw = ((long long)P.R_cap + PSB - 1) / PSB * PSB;
    e->rk_pad = ((long long)P.R_cap + PKC - 1) / PKC * PKC;
    CU(dalloc(e, &e->d_Wt, (size_t)e->ldw * e->rk_pad));
    CU(dalloc(e, &e->d_alpha, (size_t)P.R_cap));
    CU(dalloc(e, &P.kvec, (size_t)MAXNB * P.R_cap));
    P.Wt = e->d_Wt;
    P.ldW = e->ldw;
    for (int d = 0; d < D; ++d) { P.cx[d] = e->d_pts + (size_t)d * Nn; }
    P.tsdf = e->d_tsdf;

    // update kernel launch shape: persistent, a whole number of CTAs per SM
    Dispatch disp = get_dispatch(D, NB > 1);
    size_t max_ell = (size_t)96 * 1024 / (sizeof(double) * NB);
    P.ell_chunk = (int)std::min<size_t>((size_t)P.R_cap, max_ell);
    if (P.ell_chunk < 1) P.ell_chunk = 1;
    e->upd_smem = sizeof(double) * NB * (size_t)P.ell_chunk;
    CU(cudaFuncSetAttribute((const void*)disp.update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->upd_smem));
    CU(cudaFuncSetAttribute((const void*)disp.predict, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PRED_SMEM));
    CU(cudaFuncSetAttribute((const void*)disp.g_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM));
    CU(cudaFuncSetAttribute((const void*)disp.g_gemm_lists, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM));
    CU(cudaFuncSetAttribute((const void*)disp.g_gemm_tf32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TF32_SMEM));
    P.gemm_kc = cfg->precision == GPIS_PREC_TF32 ? TKC : GKC;
    P.gemm_zgroup = cfg->precision == GPIS_PREC_TF32 ? TZ : 1;
    e->solve_ctas = e->num_sms * 4;
    e->solve_smem = sizeof(double) * (size_t)NB * P.R_cap;
    e->fused_solve = e->solve_smem <= 48 * 1024;      // kvec of every CTA in shared memory, 4 CTAs/SM
    {
      int khz = 0;
      CU(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, e->device));
      const long long ms = cfg->spin_timeout_ms > 0 ? cfg->spin_timeout_ms : 30000;
      P.spin_cycles = ms * (long long)(khz > 0 ? khz : 2000000);
    }
    e->onepass = (NB == 1) && P.R_cap <= GS1_RMAX && cfg->append_mode != GPIS_APPEND_TWO_PASS;
    e->solve_ring = cfg->append_mode == GPIS_APPEND_AUTO;       // rows of W through a bulk-copy smem ring + grid barrier
    e->persist = e->onepass && cfg->loop_mode == GPIS_LOOP_AUTO && cfg->precision == GPIS_PREC_FP64;
#ifdef GPIS_EMULATE
    e->solve_ring = false;      // in-kernel grid barriers: emulated blocks run one after another
    e->persist = false;
#endif
    if (e->onepass) {
      int coop = 0;
      CU(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, e->device));
      CU(cudaFuncSetAttribute((const void*)disp.g_solve2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GS2_SMEM));
      CU(cudaFuncSetAttribute((const void*)disp.g_solve1, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)(sizeof(double) * GS1_RMAX)));
#ifndef GPIS_EMULATE
      CU(cudaFuncSetAttribute((const void*)disp.g_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PERSIST_SMEM));
      if (disp.g_persist_cube)
        CU(cudaFuncSetAttribute((const void*)disp.g_persist_cube, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PERSIST_SMEM));
      // kernels with in-kernel grid barriers need every CTA resident: one per SM, or not at all
      int occ2 = 0, occp = 0;
      CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, (const void*)disp.g_solve2, GSOLVE_THREADS, GS2_SMEM));
      CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occp, (const void*)disp.g_persist, PTHREADS, PERSIST_SMEM));
      if (!coop || occ2 < 1) e->solve_ring = false;
      if (disp.g_persist_cube && occp >= 1)
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occp, (const void*)disp.g_persist_cube, PTHREADS, PERSIST_SMEM));
      if (!coop || occp < 1) e->persist = false;
      // phase C of the persistent kernel sums the partial Y's with ONE column per thread group: every model row needs one
      if ((long long)e->num_sms * (PTHREADS / 32) * 4 < (long long)P.R_cap) e->persist = false;
      if (e->num_sms > 32 * (PTHREADS / 32)) e->persist = false;      // its winner reduction: one record per lane of the CTA's warps
#endif
      e->solve1_ctas = e->solve_ring ? e->num_sms - 1 : e->num_sms * 2;
      CU(dalloc(e, &P.ypart, (size_t)std::max(e->solve1_ctas, e->num_sms) * GS1_RMAX));
      CU(dalloc(e, &e->d_phase_clk, (size_t)16 + 2 * (size_t)e->num_sms));
      CU(cudaMemset(e->d_phase_clk, 0, (16 + 2 * (size_t)e->num_sms) * sizeof(unsigned long long)));
      P.phase_clk = e->d_phase_clk;
    }
    if (e->fused_solve)
      CU(cudaFuncSetAttribute((const void*)disp.g_solve_fused, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)e->solve_smem));
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)disp.update, UPD_THREADS, e->upd_smem));
    if (occ < 1) return fail(e, GPIS_ERR_CUDA, "update kernel does not fit on an SM");
    e->upd_grid = e->num_sms * occ;
    e->partials_cap = (size_t)std::max(e->upd_grid, 2 * e->num_sms);      // persistent kernel: winner + its target / noise per CTA
    CU(dalloc(e, &P.partials, e->partials_cap));
    e->state_cap = Nn;
    CU(cudaEventCreate(&e->ev0));
    CU(cudaEventCreate(&e->ev1));
    CU(cudaEventCreate(&e->ev2));
    CU(cudaEventCreate(&e->ev3));
    return GPIS_OK;
  };
  rc = body();
  if (rc == GPIS_OK && e->P.world == 1 && e->persist) {
    // world_size 1: the exchange region is the engine's own; its generation flags are what release the
    // CTAs of the persistent kernel into the next step
    char h64[64];
    rc = gpis_peer_export(e, h64);
    void* self = e->xchg_base;
    if (rc == GPIS_OK) rc = gpis_peer_import(e, &self, 1);
    e->self_peer = rc == GPIS_OK;
  }
  if (caller_device >= 0) cudaSetDevice(caller_device);      // the caller's current device is not ours to change
  if (rc != GPIS_OK) {
    g_last_error = e->err;        // reachable through gpis_last_error(NULL)
    gpis_destroy(e);
    return rc;
  }
  *out = e;
  return GPIS_OK;
}

gpis_status gpis_destroy(gpis_handle e) {
  if (!e) return GPIS_ERR_INVALID;
  GUARD(e);
  if (e->step_graph) cudaGraphExecDestroy(e->step_graph);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  if (e->ev2) cudaEventDestroy(e->ev2);
  if (e->ev3) cudaEventDestroy(e->ev3);
  for (cudaEvent_t ev : e->prof_events) cudaEventDestroy(ev);
  for (void* p : e->ipc_opened) cudaIpcCloseMemHandle(p);
  for (void* p : e->allocs) cudaFree(p);
  for (auto& b : e->pool) cudaFree(b.p);
  delete e;
  return GPIS_OK;
}

gpis_status gpis_set_candidates(gpis_handle e, const double* pts, const double* tsdf, const double* normals,
                                const double* noise, const int64_t* ids, void* stream) {
  if (!e) return GPIS_ERR_INVALID;
  GUARD(e);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t N = (size_t)e->P.N;
  if (e->cfg.precision == GPIS_PREC_TF32)
    return fail(e, GPIS_ERR_INVALID, "GPIS_PREC_TF32 is implemented for lattice candidates (gpis_set_grid_candidates) only");
  if (N && (!pts || !tsdf)) return fail(e, GPIS_ERR_INVALID, "pts and tsdf are required");
  if (e->NB > 1 && N && !normals) return fail(e, GPIS_ERR_INVALID, "use_gradients needs normals");
  if (N) {
    CU(cudaMemcpyAsync(e->d_pts, pts, sizeof(double) * N * e->D, cudaMemcpyDefault, s));
    CU(cudaMemcpyAsync(e->d_tsdf, tsdf, sizeof(double) * N, cudaMemcpyDefault, s));
    if (normals) CU(cudaMemcpyAsync(e->d_nrm, normals, sizeof(double) * N * e->D, cudaMemcpyDefault, s));
    if (noise) CU(cudaMemcpyAsync(e->d_noise, noise, sizeof(double) * N, cudaMemcpyDefault, s));
    if (ids) CU(cudaMemcpyAsync(e->d_ids, ids, sizeof(long long) * N, cudaMemcpyDefault, s));
  }
  for (int d = 0; d < e->D; ++d) e->P.nrm[d] = normals ? e->d_nrm + (size_t)d * (N ? N : 1) : nullptr;
  e->P.noise = noise ? e->d_noise : nullptr;
  e->P.ids = ids ? e->d_ids : nullptr;
  if (!e->P.V) CU(dalloc(e, &e->P.V, (size_t)e->P.n_tiles * (size_t)e->P.tile_stride));
  for (int d = 0; d < MAXD; ++d) e->P.glen[d] = 0;
  for (int d = 0; d < e->D; ++d) e->P.cx[d] = e->d_pts + (size_t)d * (N ? N : 1);
  e->backend = 1;
  e->have_candidates = true;
  e->selecting = false;
  if (e->step_graph) { cudaGraphExecDestroy(e->step_graph); e->step_graph = nullptr; }
  return GPIS_OK;
}

gpis_status gpis_set_grid_candidates(gpis_handle e, const int32_t* dims, const double* origin,
                                     const double* spacing, int32_t z0, int32_t nz_total, const double* tsdf,
                                     const double* normals, const double* noise, void* stream) {
  if (!e) return GPIS_ERR_INVALID;
  GUARD(e);
  if (!dims || !origin || !spacing || !tsdf) return fail(e, GPIS_ERR_INVALID, "dims, origin, spacing and tsdf are required");
  cudaStream_t s = (cudaStream_t)stream;
  const int D = e->D;
  long long n = 1;
  for (int d = 0; d < D; ++d) { if (dims[d] < 1) return fail(e, GPIS_ERR_INVALID, "dims must be >= 1"); n *= dims[d]; }
  if (n != e->P.N) return fail(e, GPIS_ERR_INVALID, "product of dims differs from num_candidates");
  if (e->NB > 1 && !normals) return fail(e, GPIS_ERR_INVALID, "use_gradients needs normals");
  if (D < 3) { z0 = 0; nz_total = 1; }
  if (nz_total < dims[D - 1] + z0 && D == 3) nz_total = dims[2] + z0;
  EngineParams& P = e->P;
  const size_t N = (size_t)P.N;
  CU(cudaMemcpyAsync(e->d_tsdf, tsdf, sizeof(double) * N, cudaMemcpyDefault, s));
  if (normals) CU(cudaMemcpyAsync(e->d_nrm, normals, sizeof(double) * N * D, cudaMemcpyDefault, s));
  if (noise) CU(cudaMemcpyAsync(e->d_noise, noise, sizeof(double) * N, cudaMemcpyDefault, s));
  for (int d = 0; d < D; ++d) P.nrm[d] = normals ? e->d_nrm + (size_t)d * N : nullptr;
  P.noise = noise ? e->d_noise : nullptr;
  P.ids = nullptr;
  for (int d = 0; d < MAXD; ++d) {
    P.cx[d] = nullptr;
    P.glen[d] = d < D ? dims[d] : 1;
    P.gorg[d] = d < D ? origin[d] : 0.0;
    P.gh[d] = d < D ? spacing[d] : 1.0;
    // axis 2's table spans the global z range; tiles read whole 128-wide rows of axes 0 and 1
    const long long len = (d == 2 && D == 3) ? nz_total : P.glen[d];
    P.gtlen[d] = (int)len;
    const int ld = (int)((len + GT - 1) / GT * GT);
    const size_t bytes = sizeof(double) * (size_t)(e->rk_pad + GKC) * ld;
    if (bytes > e->table_bytes[d]) {
      CU(dalloc(e, &P.T[d], bytes / sizeof(double)));
      if (e->cfg.precision == GPIS_PREC_TF32 && d < 2) CU(dalloc(e, &P.Tf[d], bytes / sizeof(double)));
      e->table_bytes[d] = bytes;
    }
    P.ldt[d] = ld;
  }
  P.gz0 = z0;
  if (!P.W) {
    CU(dalloc(e, &P.W, (size_t)e->rk_pad * e->ldw));
    e->w_bytes = sizeof(double) * (size_t)e->rk_pad * e->ldw;
  }
  if (e->cfg.precision == GPIS_PREC_TF32) {
    P.bimg_cmax = (int)((e->rk_pad + GKC + TKC - 1) / TKC);
    const size_t need = (size_t)((P.glen[0] + GT - 1) / GT) * P.bimg_cmax * 2 * GT * TKC;
    if (need > e->bimg_floats) {
      CU(dalloc(e, &P.bimg, need));
      e->bimg_floats = need;
    }
  }
  const int tiles = ((P.glen[0] + GT - 1) / GT) * ((P.glen[1] + GT - 1) / GT) * P.glen[2];
  e->grid_ctas = tiles;
  if ((size_t)GEPI_CTAS_PER_TILE * tiles > e->partials_cap) {
    CU(dalloc(e, &P.partials, (size_t)GEPI_CTAS_PER_TILE * tiles));
    e->partials_cap = (size_t)GEPI_CTAS_PER_TILE * tiles;
  }
  // per-candidate state in tile/fragment order (padded to whole tiles)
  e->grid_slots = (size_t)tiles * GTILE_ELEMS;
  if (D == 3)      // the persistent kernel's 16 x 16 x 16 cube layout pads the z axis too
    e->grid_slots = std::max(e->grid_slots, (size_t)((P.glen[0] + QC - 1) / QC) * ((P.glen[1] + QC - 1) / QC) *
                                                ((P.glen[2] + QC - 1) / QC) * QTILE_ELEMS);
  if (e->grid_slots > e->state_cap) {
    CU(dalloc(e, &P.mu, e->grid_slots));
    CU(dalloc(e, &P.var, e->grid_slots));
    CU(dalloc(e, &P.amb, e->grid_slots));
    CU(dalloc(e, &P.flags, e->grid_slots));
    e->state_cap = e->grid_slots;
  }
  P.gemm_ctas = e->num_sms;
  {
    const int txy = ((P.glen[0] + GT - 1) / GT) * ((P.glen[1] + GT - 1) / GT);
    // TF32 GEMM: pair z-slices per work tile only while that still leaves about one work tile per
    // two SMs; on a small shard (multi-GPU slabs) pairing would split each tile over twice as many
    // CTAs and double the partials the epilogue has to sum
    if (e->cfg.precision == GPIS_PREC_TF32)
      P.gemm_zgroup = (txy * ((P.glen[2] + TZ - 1) / TZ) * 3 >= e->num_sms) ? TZ : 1;
    const int groups = txy * ((P.glen[2] + P.gemm_zgroup - 1) / P.gemm_zgroup);
    P.gws_segmax = P.gemm_ctas / groups + 2;
  }
  // persistent selection kernel: 64 x 64 lattice tiles, a row list and a split-K counter per tile, two
  // workspace slots per CTA
  // (3-D lattices: 16 x 16 x 16 cubes, unless the slab is so thin that padding its z axis to whole cubes wastes more
  // than the shorter row lists save)
  e->persist_cube = D == 3 && 3 * (((long long)P.glen[2] + QC - 1) / QC * QC) <= 4 * (long long)P.glen[2];
  const long long tiles_q = e->persist_cube
      ? (long long)((P.glen[0] + QC - 1) / QC) * ((P.glen[1] + QC - 1) / QC) * ((P.glen[2] + QC - 1) / QC)
      : (long long)((P.glen[0] + QT - 1) / QT) * ((P.glen[1] + QT - 1) / QT) * P.glen[2];
  e->persist_grid_ok = e->persist && tiles_q <= QT_MAX && P.R_cap <= 65535;
  // per-step fp64 path (the models the persistent kernel does not take -- gradient observations, more than 4096 rows --
  // and loop_mode = GPIS_LOOP_STEPS): row lists per 128 x 128 tile, one workspace slot per tile + one per CTA
  e->step_lists = e->cfg.precision == GPIS_PREC_FP64 && tiles <= GLIST_TMAX && P.R_cap <= 65535;
  {
    size_t need = sizeof(double) * (size_t)e->NB * tiles * P.gws_segmax * GTILE_ELEMS;
    if (e->step_lists) need = sizeof(double) * (size_t)e->NB * ((size_t)tiles + P.gemm_ctas) * GTILE_ELEMS;
    if (e->persist_grid_ok) need = std::max(need, sizeof(double) * (size_t)e->num_sms * 2 * QTILE_ELEMS);
    if (need > e->gws_bytes) {
      CU(dalloc(e, &P.gws, need / sizeof(double)));
      e->gws_bytes = need;
    }
  }
  P.upref = nullptr;
  if (e->step_lists) {
    if ((size_t)tiles + 1 > e->upref_cap) {
      CU(dalloc(e, &e->d_upref, (size_t)tiles + 1));
      e->upref_cap = (size_t)tiles + 1;
    }
    P.upref = e->d_upref;
  }
  if (e->persist_grid_ok || e->step_lists) {
    // (the per-step row-list path keeps its lists of 128 x 128 tiles in the same arrays)
    const size_t list_tiles = (size_t)std::max<long long>(e->persist_grid_ok ? tiles_q : 0, e->step_lists ? tiles : 0);
    if (list_tiles > e->tile_cnt_cap) {
      CU(dalloc(e, &P.tile_cnt, list_tiles));
      CU(dalloc(e, &P.rcnt, list_tiles));
      e->tile_cnt_cap = list_tiles;
    }
    P.ldtt = e->rk_pad;
    for (int d = 0; d < D && e->persist_grid_ok; ++d) {
      const size_t need = (size_t)P.gtlen[d] * (size_t)P.ldtt;
      if (need > e->tt_cap[d]) {
        CU(dalloc(e, &P.TT[d], need));
        e->tt_cap[d] = need;
      }
    }
    P.rlist_stride = (P.R_cap + QKC - 1) / QKC * QKC;      // whole chunks, 64-byte aligned
    if (list_tiles * P.rlist_stride > e->rlist_cap) {
      CU(dalloc(e, &P.rlist, list_tiles * P.rlist_stride));
      e->rlist_cap = list_tiles * P.rlist_stride;
    }
    if (e->persist_grid_ok && !e->persist_cube && (size_t)tiles_q * P.rlist_stride > e->rtz_cap) {
      CU(dalloc(e, &P.rtz, (size_t)tiles_q * P.rlist_stride));
      e->rtz_cap = (size_t)tiles_q * P.rlist_stride;
    }
  }
  P.tile_e = GT;
  e->backend = 2;
  e->have_candidates = true;
  e->selecting = false;
  if (e->step_graph) { cudaGraphExecDestroy(e->step_graph); e->step_graph = nullptr; }
  return GPIS_OK;
}

namespace {
__global__ void k_fill_ones_col0(double* T, float* Tf, int ld, long long rows) {
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r < rows) {
    T[r * ld] = 1.0;
    if (Tf) Tf[r * ld] = 1.0f;
  }
}
}  // namespace

namespace {
gpis_status select_begin_impl(gpis_handle e, int64_t first_local, void* stream, int tile_e);
}

gpis_status gpis_select_begin(gpis_handle e, int64_t first_local, void* stream) {
  if (!e) return GPIS_ERR_INVALID;
  GUARD(e);
  return select_begin_impl(e, first_local, stream, GT);      // the per-step kernels work on 128 x 128 tiles
}

namespace {
gpis_status select_begin_impl(gpis_handle e, int64_t first_local, void* stream, int tile_e) {
  if (!e->have_candidates) return fail(e, GPIS_ERR_STATE, "gpis_set_candidates has not been called");
  if (first_local >= e->P.N) return fail(e, GPIS_ERR_INVALID, "first index out of range");
  cudaStream_t s = (cudaStream_t)stream;
  const long long n = std::max<long long>(e->P.N, e->P.n_tiles);
  if (e->backend == 2) {
    e->P.tile_e = tile_e;             // layout of the per-candidate state for THIS selection
    const long long nslots = grid_slot_count(e->P);
    GPIS_LAUNCH_SMEM(k_grid_init_state, (unsigned)((nslots + 255) / 256), 256, 0, s, e->P, nslots);
    CU(cudaGetLastError());
    e->launches++;
  } else if (n > 0) {
    GPIS_LAUNCH_SMEM(k_init_candidates, (unsigned)((n + 255) / 256), 256, 0, s, e->P);
    CU(cudaGetLastError());
    e->launches++;
  }
  if (e->backend == 2) {
    CU(cudaMemsetAsync(e->P.W, 0, e->w_bytes, s));
    if (e->P.tile_cnt) {
      CU(cudaMemsetAsync(e->P.tile_cnt, 0, sizeof(unsigned) * e->tile_cnt_cap, s));
      CU(cudaMemsetAsync(e->P.rcnt, 0, sizeof(int) * e->tile_cnt_cap, s));
    }
    if (e->P.bimg) CU(cudaMemsetAsync(e->P.bimg, 0, sizeof(float) * e->bimg_floats, s));
    CU(cudaMemsetAsync(e->d_Wt, 0, sizeof(double) * e->rk_pad * e->ldw, s));
    for (int d = 0; d < MAXD; ++d) {
      CU(cudaMemsetAsync(e->P.T[d], 0, e->table_bytes[d], s));
      if (e->P.Tf[d]) CU(cudaMemsetAsync(e->P.Tf[d], 0, e->table_bytes[d] / 2, s));
      if (d >= e->D) {    // unused axes: constant-one tables
        const long long rows = e->rk_pad + GKC;
        GPIS_LAUNCH_SMEM(k_fill_ones_col0, (unsigned)((rows + 255) / 256), 256, 0, s, e->P.T[d], e->P.Tf[d], e->P.ldt[d], rows);
        e->launches++;
      }
    }
  }
  GPIS_LAUNCH_SMEM(k_begin, 1, 32, 0, s, e->P, first_local < 0 ? -1 : first_local, ++e->epoch);
  CU(cudaGetLastError());
  e->launches++;
  e->selecting = true;
  e->inverse_rows = -1;
  return GPIS_OK;
}
}  // namespace

gpis_status gpis_step_append(gpis_handle e, void* stream) {
  if (!e) return GPIS_ERR_INVALID;
  GUARD(e);
  if (!e->selecting) return fail(e, GPIS_ERR_STATE, "gpis_select_begin has not been called");
  e->inverse_rows = -1;
  return launch_append(e, (cudaStream_t)stream, 0, 0);
}

gpis_status gpis_step_update(gpis_handle e, void* stream) {
  if (!e) return GPIS_ERR_INVALID;
  GUARD(e);
  if (!e->selecting) return fail(e, GPIS_ERR_STATE, "gpis_select_begin has not been called");
  return launch_update(e, (cudaStream_t)stream);
}

gpis_status gpis_select_end(gpis_handle e, void* stream) {
  if (!e) return GPIS_ERR_INVALID;
  GUARD(e);
  if (!e->selecting) return fail(e, GPIS_ERR_STATE, "gpis_select_begin has not been called");
  return launch_append(e, (cudaStream_t)stream, 0, 1);
}

gpis_status gpis_exchange_buffers(gpis_handle e, void** send, void** recv, int64_t* bytes_per_rank) {
  if (!e) return GPIS_ERR_INVALID;
  GUARD(e);
  if (send) *send = e->P.pay_send;
  if (recv) *recv = e->P.pay_recv;
  if (bytes_per_rank) *bytes_per_rank = (int64_t)(e->P.pay_stride * sizeof(double));
  return GPIS_OK;
}

gpis_status gpis_peer_export(gpis_handle e, void* handle64) {
  if (!e || !handle64) return GPIS_ERR_INVALID;
  GUARD(e);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  EngineParams& P = e->P;
  if (!e->xchg_base) {
    const size_t rec = sizeof(double) * 2 * (size_t)P.world * P.pay_stride;
    // records, generation flags [2][world], done [world], tagged per-step slots [2][world][16]
    P.xll_off = (long long)(rec + sizeof(unsigned long long) * 3 * P.world);
    e->xchg_bytes = (size_t)P.xll_off + sizeof(unsigned long long) * 2 * P.world * 16;
    CU(dalloc(e, reinterpret_cast<char**>(&e->xchg_base), e->xchg_bytes));
    CU(cudaMemset(e->xchg_base, 0, e->xchg_bytes));
    P.xchg = static_cast<double*>(e->xchg_base);
    P.xflag = reinterpret_cast<unsigned long long*>(static_cast<char*>(e->xchg_base) + rec);
    P.xdone = P.xflag + 2 * P.world;
  }
  cudaIpcMemHandle_t hd;
  CU(cudaIpcGetMemHandle(&hd, e->xchg_base));
  memcpy(handle64, &hd, 64);
  return GPIS_OK;
}

gpis_status gpis_peer_import(gpis_handle e, const void* handles, int32_t same_process) {
  if (!e || !handles) return GPIS_ERR_INVALID;
  GUARD(e);
  EngineParams& P = e->P;
  if (!e->xchg_base) return fail(e, GPIS_ERR_STATE, "gpis_peer_export has not been called");
  const size_t rec = sizeof(double) * 2 * (size_t)P.world * P.pay_stride;
  std::vector<double*> xp(P.world);
  std::vector<unsigned long long*> fp(P.world), dn(P.world);
  for (int r = 0; r < P.world; ++r) {
    void* base = nullptr;
    if (r == P.rank) base = e->xchg_base;
    else if (same_process) base = static_cast<void* const*>(handles)[r];
    else {
      cudaIpcMemHandle_t hd;
      memcpy(&hd, static_cast<const char*>(handles) + 64 * (size_t)r, 64);
      CU(cudaIpcOpenMemHandle(&base, hd, cudaIpcMemLazyEnablePeerAccess));
      e->ipc_opened.push_back(base);
    }
    xp[r] = static_cast<double*>(base);
    fp[r] = reinterpret_cast<unsigned long long*>(static_cast<char*>(base) + rec);
    dn[r] = fp[r] + 2 * P.world;
  }
  if (!P.peer_xchg) {
    CU(dalloc(e, &P.peer_xchg, (size_t)P.world));
    CU(dalloc(e, &P.peer_xflag, (size_t)P.world));
    CU(dalloc(e, &P.peer_xdone, (size_t)P.world));
  }
  CU(cudaMemcpy(P.peer_xdone, dn.data(), sizeof(unsigned long long*) * P.world, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(P.peer_xchg, xp.data(), sizeof(double*) * P.world, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(P.peer_xflag, fp.data(), sizeof(unsigned long long*) * P.world, cudaMemcpyHostToDevice));
  P.peer_mode = 1;
  if (e->step_graph) { cudaGraphExecDestroy(e->step_graph); e->step_graph = nullptr; }
  return GPIS_OK;
}

gpis_status gpis_select(gpis_handle e, int64_t first_local, int32_t K, void* stream) {
  if (!e) return GPIS_ERR_INVALID;
  GUARD(e);
  if (e->P.world != 1 && !e->P.peer_mode)
    return fail(e, GPIS_ERR_STATE, "gpis_select with world_size > 1 needs the peer exchange (gpis_peer_import); "
                                   "otherwise drive the step calls and all-gather the exchange buffer");
  if (K < 1 || K > e->P.M_cap) return fail(e, GPIS_ERR_INVALID, "K must be in 1..max_active");
  if (first_local < 0 && e->P.world == 1) return fail(e, GPIS_ERR_INVALID, "first index out of range");
  cudaStream_t s = (cudaStream_t)stream;
  CU(cudaEventRecord(e->ev0, s));
  // one persistent cooperative kernel for the whole greedy loop (lattice candidates, function values, fp64)
  const bool persist = e->persist && e->backend == 2 && e->persist_grid_ok && !e->profiling && e->P.peer_mode && e->P.tile_cnt &&
                       e->P.world <= PLL_WORLD_MAX;
  gpis_status rc = select_begin_impl(e, first_local, stream, persist ? (e->persist_cube ? QC : QT) : GT);
  if (rc != GPIS_OK) return rc;
  const int iters = K - 1;
  bool persist_ran = persist;
  for (double& v : e->persist_ms) v = 0.0;      // of THIS selection (stay zero on the per-step path)
  if (persist && iters > 0) {
    Dispatch d = get_dispatch(e->D, false);
    int it_arg = iters;
    void* args[] = {(void*)&e->P, (void*)&it_arg};
    const cudaError_t ce = launch_cooperative(e->persist_cube ? d.g_persist_cube : d.g_persist, (unsigned)e->num_sms, PTHREADS, PERSIST_SMEM, s, args);
    if (ce == cudaSuccess) {
      e->launches += 1;
    } else if (e->P.world == 1) {
      // the driver refused the cooperative launch (its blocks cannot all be resident: another context or an SM limit
      // on the device): the per-step kernels take the selection, on their own state layout, and keep taking it
      cudaGetLastError();
      e->persist = false;
      persist_ran = false;
      rc = select_begin_impl(e, first_local, stream, GT);
      if (rc != GPIS_OK) return rc;
    } else {
      return fail(e, GPIS_ERR_CUDA, std::string("cooperative launch of the persistent selection kernel: ") + cudaGetErrorString(ce));
    }
  }
  if (persist_ran) {
    rc = gpis_select_end(e, stream);
    if (rc != GPIS_OK) return rc;
    CU(cudaEventRecord(e->ev1, s));
    CU(cudaEventSynchronize(e->ev1));
    float pms = 0.f;
    CU(cudaEventElapsedTime(&pms, e->ev0, e->ev1));
    e->select_ms = pms;
    unsigned long long clk[PCLK_N];
    memset(clk, 0, sizeof(clk));
    if (iters > 0) CU(cudaMemcpy(clk, e->d_phase_clk, sizeof(clk), cudaMemcpyDeviceToHost));
    for (int i = 0; i < PCLK_N; ++i) e->persist_ms[i] = (double)clk[i] * 1e-6;
    e->persist_ms[PCLK_STEPS] = (double)clk[PCLK_STEPS];
    e->persist_ms[PCLK_UNITS] = (double)clk[PCLK_UNITS];
    // CTA 0's view of the phases, summed over the steps: W pass + finish (with their barriers) = append,
    // GEMM, fused epilogue
    e->append_ms = e->persist_ms[PCLK_WSETUP] + e->persist_ms[PCLK_WPASS] + e->persist_ms[PCLK_BAR1] + e->persist_ms[PCLK_FINISH] +
                   e->persist_ms[PCLK_BAR2];
    e->main_ms = e->persist_ms[PCLK_PREFIX] + e->persist_ms[PCLK_GEMM];
    e->epilogue_ms = e->persist_ms[PCLK_EPI];
    e->update_ms = e->main_ms + e->epilogue_ms;
    e->update_launches = (long long)clk[PCLK_STEPS];
    return read_state(e, s);
  }
  if (iters > 0 && !e->step_graph && !e->profiling) {
    // capture one (append, update) pair; every argument is constant, all state is on the device
    cudaStream_t cs = nullptr;
    cudaGraph_t g = nullptr;
    CU(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    bool ok = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
      const long long l0 = e->launches;
      ok = launch_append(e, cs, 0, 0) == GPIS_OK && launch_update(e, cs) == GPIS_OK;
      e->launches = l0;
      ok = (cudaStreamEndCapture(cs, &g) == cudaSuccess) && ok && g;
      if (ok) ok = cudaGraphInstantiate(&e->step_graph, g, 0) == cudaSuccess;
      if (g) cudaGraphDestroy(g);
    }
    cudaStreamDestroy(cs);
    if (!ok) { cudaGetLastError(); e->step_graph = nullptr; }
  }
  const bool prof = e->profiling;
  if (prof) {
    while ((int)e->prof_events.size() < 4 * iters) {
      cudaEvent_t ev;
      CU(cudaEventCreate(&ev));
      e->prof_events.push_back(ev);
    }
  }
  for (int it = 0; it < iters; ++it) {
    if (prof) {
      CU(cudaEventRecord(e->prof_events[4 * it], s));
      rc = launch_append(e, s, 0, 0);
      if (rc != GPIS_OK) return rc;
      CU(cudaEventRecord(e->prof_events[4 * it + 1], s));
      rc = launch_update(e, s, e->backend == 2 ? e->prof_events[4 * it + 2] : nullptr);
      if (rc != GPIS_OK) return rc;
      if (e->backend != 2) CU(cudaEventRecord(e->prof_events[4 * it + 2], s));
      CU(cudaEventRecord(e->prof_events[4 * it + 3], s));
    } else if (e->step_graph) {
      CU(cudaGraphLaunch(e->step_graph, s));
      e->launches += e->backend == 2 ? ((e->onepass && e->solve_ring) ? 3 : (e->fused_solve || e->onepass) ? 4 : 5) : 2;
    } else {
      rc = launch_append(e, s, 0, 0);
      if (rc != GPIS_OK) return rc;
      rc = launch_update(e, s);
      if (rc != GPIS_OK) return rc;
    }
  }
  rc = gpis_select_end(e, stream);
  if (rc != GPIS_OK) return rc;
  CU(cudaEventRecord(e->ev1, s));
  CU(cudaEventSynchronize(e->ev1));
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
  e->select_ms = ms;
  e->update_ms = e->append_ms = e->main_ms = e->epilogue_ms = 0.0;
  e->update_launches = 0;
  if (prof) {
    for (int it = 0; it < iters; ++it) {
      float t = 0.f;
      CU(cudaEventElapsedTime(&t, e->prof_events[4 * it], e->prof_events[4 * it + 1]));
      e->append_ms += t;
      CU(cudaEventElapsedTime(&t, e->prof_events[4 * it + 1], e->prof_events[4 * it + 2]));
      e->main_ms += t;
      CU(cudaEventElapsedTime(&t, e->prof_events[4 * it + 2], e->prof_events[4 * it + 3]));
      e->epilogue_ms += t;
      CU(cudaEventElapsedTime(&t, e->prof_events[4 * it + 1], e->prof_events[4 * it + 3]));
      e->update_ms += t;
    }
    e->update_launches = iters;
  }
  return read_state(e, s);
}

gpis_status gpis_get_active(gpis_handle e, int64_t* ids, int32_t* n_selected, int32_t* n_in_model, void* stream) {
  if (!e) return GPIS_ERR_INVALID;
  GUARD(e);
  cudaStream_t s = (cudaStream_t)stream;
  gpis_status rc = read_state(e, s);
  if (rc != GPIS_OK) return rc;
  if (n_selected) *n_selected = e->host_state.n_sel;
  if (n_in_model) *n_in_model = e->host_state.m;
  if (ids && e->host_state.n_sel > 0) {
    CU(cudaMemcpyAsync(ids, e->P.sel_ids, sizeof(long long) * e->host_state.n_sel, cudaMemcpyDefault, s));
    CU(cudaStreamSynchronize(s));
  }
  return GPIS_OK;
}

gpis_status gpis_get_factor(gpis_handle e, double* L, int64_t ldl, double* alpha, int32_t* rows, void* stream) {
  if (!e) return GPIS_ERR_INVALID;
  GUARD(e);
  cudaStream_t s = (cudaStream_t)stream;
  gpis_status rc = ensure_inverse(e, s);
  if (rc != GPIS_OK) return rc;
  const int R = e->host_state.r;
  if (rows) *rows = R;
  if (L && R > 0) {
    if (ldl < R) return fail(e, GPIS_ERR_INVALID, "ldl smaller than the row count");
    CU(cudaMemcpy2DAsync(L, sizeof(double) * ldl, e->P.L, sizeof(double) * e->P.R_cap, sizeof(double) * R, R,
                         cudaMemcpyDefault, s));
  }
  if (alpha && R > 0) CU(cudaMemcpyAsync(alpha, e->d_alpha, sizeof(double) * R, cudaMemcpyDefault, s));
  CU(cudaStreamSynchronize(s));
  return GPIS_OK;
}

gpis_status gpis_fit(gpis_handle e, const double* pts, const double* tsdf, const double* normals,
                     const double* noise, int32_t m, void* stream) {
  if (!e) return GPIS_ERR_INVALID;
  GUARD(e);
  if (m < 0 || m > e->P.M_cap) return fail(e, GPIS_ERR_INVALID, "m must be in 0..max_active");
  if (m && (!pts || !tsdf)) return fail(e, GPIS_ERR_INVALID, "pts and tsdf are required");
  if (e->NB > 1 && m && !normals) return fail(e, GPIS_ERR_INVALID, "use_gradients needs normals");
  cudaStream_t s = (cudaStream_t)stream;
  const int D = e->D;
  // stage the training set on the host side as exchange records, one point per append
  std::vector<double> hp((size_t)m * D), ht(m), hn((size_t)m * D, 0.0), hz(m, e->cfg.noise);
  CU(cudaMemcpyAsync(hp.data(), pts, sizeof(double) * m * D, cudaMemcpyDefault, s));
  CU(cudaMemcpyAsync(ht.data(), tsdf, sizeof(double) * m, cudaMemcpyDefault, s));
  if (normals) CU(cudaMemcpyAsync(hn.data(), normals, sizeof(double) * m * D, cudaMemcpyDefault, s));
  if (noise) CU(cudaMemcpyAsync(hz.data(), noise, sizeof(double) * m, cudaMemcpyDefault, s));
  CU(cudaStreamSynchronize(s));
  {
    EngineParams q = e->P;
    q.peer_mode = 0;
    GPIS_LAUNCH_SMEM(k_begin, 1, 32, 0, s, q, -1, 0u);
  }
  CU(cudaGetLastError());
  e->launches++;
  e->selecting = false;
  e->inverse_rows = -1;
  // records live in a staging array on the device; each append reads record 0 of pay_recv,
  // so copy them in one by one (stream-ordered)
  double* d_recs = nullptr;
  CU(cudaMalloc(&d_recs, sizeof(double) * PAY_HDR * (size_t)(m > 0 ? m : 1)));
  std::vector<double> recs((size_t)PAY_HDR * m, 0.0);
  for (int i = 0; i < m; ++i) {
    double* r = &recs[(size_t)i * PAY_HDR];
    r[PAY_SCORE] = 0.0;
    long long idv = i, loc = 0;
    memcpy(&r[PAY_ID], &idv, 8);
    memcpy(&r[PAY_LOCAL], &loc, 8);
    r[PAY_TSDF] = ht[i];
    r[PAY_NOISE] = hz[i];
    for (int d = 0; d < D; ++d) { r[PAY_X + d] = hp[(size_t)d * m + i]; r[PAY_NRM + d] = hn[(size_t)d * m + i]; }
  }
  gpis_status rc = GPIS_OK;
  cudaError_t ce = cudaMemcpyAsync(d_recs, recs.data(), sizeof(double) * recs.size(), cudaMemcpyHostToDevice, s);
  EngineParams saved = e->P;
  for (int i = 0; i < m && ce == cudaSuccess && rc == GPIS_OK; ++i) {
    e->P.pay_recv = d_recs + (size_t)i * PAY_HDR;
    e->P.world = 1;
    e->P.rank = 0;
    e->P.peer_mode = 0;
    rc = launch_append(e, s, 1, 0);
  }
  e->P = saved;
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
  cudaFree(d_recs);
  if (ce != cudaSuccess) return fail(e, GPIS_ERR_CUDA, cudaGetErrorString(ce));
  if (rc != GPIS_OK) return rc;
  return read_state(e, s);
}

gpis_status gpis_predict(gpis_handle e, const double* query, int64_t nq, double* mean, double* var,
                         double* normals, void* stream) {
  if (!e) return GPIS_ERR_INVALID;
  GUARD(e);
  if (nq < 0 || (nq && (!query || !mean || !var))) return fail(e, GPIS_ERR_INVALID, "query, mean and var are required");
  if (nq == 0) return GPIS_OK;
  cudaStream_t s = (cudaStream_t)stream;
  gpis_status rc = ensure_inverse(e, s);
  if (rc != GPIS_OK) return rc;
  const int D = e->D;
  // query / outputs may be host memory: stage through device buffers
  cudaPointerAttributes at;
  auto on_device = [&](const void* p) {
    if (!p) return true;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
  };
  const bool qd = on_device(query), od = on_device(mean) && on_device(var) && on_device(normals);
  double *dq = nullptr, *dout = nullptr;
  const size_t n = (size_t)nq;
  auto cleanup = [&]() { if (dq) cudaFree(dq); if (dout) cudaFree(dout); };
  if (!qd) {
    cudaError_t ce = cudaMalloc(&dq, sizeof(double) * n * D);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(dq, query, sizeof(double) * n * D, cudaMemcpyHostToDevice, s);
    if (ce != cudaSuccess) { cleanup(); return fail(e, GPIS_ERR_CUDA, cudaGetErrorString(ce)); }
  }
  if (!od) {
    cudaError_t ce = cudaMalloc(&dout, sizeof(double) * n * (2 + D));
    if (ce != cudaSuccess) { cleanup(); return fail(e, GPIS_ERR_NOMEM, cudaGetErrorString(ce)); }
  }
  PredictParams Q;
  memset(&Q, 0, sizeof(Q));
  Q.Wt = e->d_Wt;
  Q.ldw = e->ldw;
  Q.alpha = e->d_alpha;
  const double* qsrc = qd ? query : dq;
  for (int d = 0; d < D; ++d) { Q.AX[d] = e->P.AX[d]; Q.q[d] = qsrc + (size_t)d * n; }
  Q.nq = nq;
  Q.mean = od ? mean : dout;
  Q.var = od ? var : dout + n;
  for (int d = 0; d < D; ++d) Q.nrm[d] = normals ? (od ? normals + (size_t)d * n : dout + (2 + d) * n) : nullptr;
  Q.R = e->host_state.r;
  Q.inv_ell2 = e->P.inv_ell2;
  Q.sf2 = e->P.sf2;
  for (int d = 0; d < MAXD; ++d) Q.mean_w[d] = e->P.mean_w[d];
  Q.mean_c = e->P.mean_c;
  Dispatch disp = get_dispatch(D, e->NB > 1);
  cudaEventRecord(e->ev0, s);
  GPIS_LAUNCH_SMEM(disp.predict, (unsigned)((nq + PQ - 1) / PQ), PRED_THREADS, PRED_SMEM, s, Q);
  cudaError_t ce = cudaGetLastError();
  e->launches++;
  cudaEventRecord(e->ev1, s);
  if (ce == cudaSuccess && !od) {
    ce = cudaMemcpyAsync(mean, dout, sizeof(double) * n, cudaMemcpyDeviceToHost, s);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(var, dout + n, sizeof(double) * n, cudaMemcpyDeviceToHost, s);
    if (ce == cudaSuccess && normals)
      ce = cudaMemcpyAsync(normals, dout + 2 * n, sizeof(double) * n * D, cudaMemcpyDeviceToHost, s);
  }
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
  cleanup();
  if (ce != cudaSuccess) return fail(e, GPIS_ERR_CUDA, cudaGetErrorString(ce));
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, e->ev0, e->ev1) == cudaSuccess) e->predict_ms = ms;
  return GPIS_OK;
}

}  // extern "C"

namespace {

// per-call device scratch, released on every return path
struct Scratch {
  std::vector<void*> p;
  gpis_engine* owner = nullptr;         // with an owner the blocks come from (and go back to) the handle's pool
  cudaStream_t stream = nullptr;
  Scratch() {}
  Scratch(gpis_engine* e, cudaStream_t s) : owner(e), stream(s) {}
  ~Scratch() {
    if (!owner) { for (void* q : p) cudaFree(q); return; }
    cudaStreamSynchronize(stream);      // (cudaFree used to synchronise here) the blocks may be handed out again at once
    for (void* q : p)
      for (auto& b : owner->pool)
        if (b.p == q) b.busy = false;
  }
  template <class T>
  cudaError_t get(T** out, size_t n) {
    const size_t bytes = (n ? n : 1) * sizeof(T);
    if (owner) {
      int best = -1;
      for (int i = 0; i < (int)owner->pool.size(); ++i) {
        const auto& b = owner->pool[i];
        if (!b.busy && b.bytes >= bytes && (best < 0 || b.bytes < owner->pool[best].bytes)) best = i;
      }
      if (best >= 0 && owner->pool[best].bytes <= 2 * bytes + (1 << 20)) {      // no 2 GB block for a 1 KB request
        owner->pool[best].busy = true;
        p.push_back(owner->pool[best].p);
        *out = static_cast<T*>(owner->pool[best].p);
        return cudaSuccess;
      }
    }
    void* q = nullptr;
    cudaError_t r = cudaMalloc(&q, bytes);
    if (r == cudaSuccess) {
      p.push_back(q);
      *out = static_cast<T*>(q);
      if (owner) owner->pool.push_back({q, bytes, true});
    }
    return r;
  }
};

struct FullPosterior {
  double* mean = nullptr;   // [nv]
  double* S = nullptr;      // [npad][npad], zero outside the nv x nv corner
  long long nq = 0, nv = 0, npad = 0;
};

// MEAN and COV (gp_mean.m, gp_cov.m in full) of the joint vector at nq query points, on the device.
gpis_status posterior_full(gpis_engine* e, const double* query, int64_t nq, int deriv, cudaStream_t s, Scratch& sc,
                           FullPosterior* out) {
  gpis_status rc = ensure_inverse(e, s);
  if (rc != GPIS_OK) return rc;
  const int D = e->D, R = e->host_state.r;
  const int nt = deriv ? D + 1 : 1;
  const long long nv = (long long)nq * nt, npad = (nv + CHB - 1) / CHB * CHB;
  if (nq > 65535 || npad > 1000000) return fail(e, GPIS_ERR_INVALID, "too many query points for a full covariance (nq <= 65535)");
  double *dq = nullptr, *Ks = nullptr, *V = nullptr;
  CU(sc.get(&dq, (size_t)nq * D));
  CU(sc.get(&out->mean, (size_t)nv));
  CU(sc.get(&out->S, (size_t)npad * npad));
  if (R > 0) {
    CU(sc.get(&Ks, (size_t)R * nv));
    CU(sc.get(&V, (size_t)R * nv));
  }
  CU(cudaMemcpyAsync(dq, query, sizeof(double) * nq * D, cudaMemcpyDefault, s));
  CU(cudaMemsetAsync(out->S, 0, sizeof(double) * npad * npad, s));
  FullCovParams F;
  memset(&F, 0, sizeof(F));
  for (int d = 0; d < D; ++d) { F.AX[d] = e->P.AX[d]; F.q[d] = dq + (size_t)d * nq; }
  F.nq = nq;
  F.nt = nt;
  F.R = R;
  F.inv_ell2 = e->P.inv_ell2;
  F.sf2 = e->P.sf2;
  for (int d = 0; d < MAXD; ++d) F.mean_w[d] = e->P.mean_w[d];
  F.mean_c = e->P.mean_c;
  posterior_full_device(D, e->NB > 1, F, e->d_Wt, e->ldw, e->d_alpha, Ks, V, out->S, npad, out->mean, s, &e->launches);
  CU(cudaGetLastError());
  out->nq = nq;
  out->nv = nv;
  out->npad = npad;
  return GPIS_OK;
}

}  // namespace

extern "C" {

gpis_status gpis_posterior_cov(gpis_handle e, const double* query, int64_t nq, int32_t use_derivative, double* mean,
                               double* cov, int64_t ldc, void* stream) {
  if (!e) return GPIS_ERR_INVALID;
  GUARD(e);
  if (nq < 0 || (nq && !query)) return fail(e, GPIS_ERR_INVALID, "query is required");
  if (nq == 0) return GPIS_OK;
  const long long nv = (long long)nq * (use_derivative ? e->D + 1 : 1);
  if (cov && ldc < nv) return fail(e, GPIS_ERR_INVALID, "ldc must be >= the number of values");
  cudaStream_t s = (cudaStream_t)stream;
  Scratch sc(e, s);
  FullPosterior fp;
  cudaEventRecord(e->ev0, s);
  gpis_status rc = posterior_full(e, query, nq, use_derivative, s, sc, &fp);
  if (rc != GPIS_OK) return rc;
  cudaEventRecord(e->ev1, s);
  if (mean) CU(cudaMemcpyAsync(mean, fp.mean, sizeof(double) * nv, cudaMemcpyDefault, s));
  if (cov)
    CU(cudaMemcpy2DAsync(cov, sizeof(double) * ldc, fp.S, sizeof(double) * fp.npad, sizeof(double) * nv, (size_t)nv,
                         cudaMemcpyDefault, s));
  CU(cudaStreamSynchronize(s));
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, e->ev0, e->ev1) == cudaSuccess) e->cov_ms = ms;
  return GPIS_OK;
}

gpis_status gpis_sample_shapes(gpis_handle e, const double* query, int64_t nq, int32_t use_derivative,
                               int32_t num_samples, const double* z, double jitter, double* samples, double* logpdf,
                               double* chol, int64_t ldt, void* stream) {
  if (!e) return GPIS_ERR_INVALID;
  GUARD(e);
  if (nq <= 0 || !query) return fail(e, GPIS_ERR_INVALID, "query is required");
  if (num_samples < 0 || (num_samples && (!z || !samples))) return fail(e, GPIS_ERR_INVALID, "z and samples are required");
  if (!(jitter >= 0.0)) return fail(e, GPIS_ERR_INVALID, "jitter must be >= 0");
  const long long nv = (long long)nq * (use_derivative ? e->D + 1 : 1);
  if (chol && ldt < nv) return fail(e, GPIS_ERR_INVALID, "ldt must be >= the number of values");
  cudaStream_t s = (cudaStream_t)stream;
  Scratch sc(e, s);
  FullPosterior fp;
  cudaEventRecord(e->ev0, s);
  gpis_status rc = posterior_full(e, query, nq, use_derivative, s, sc, &fp);
  if (rc != GPIS_OK) return rc;
  cudaEventRecord(e->ev1, s);
  const long long npad = fp.npad;
  const int ns = num_samples;
  int* d_err = nullptr;
  double *dz = nullptr, *dzt = nullptr, *dout = nullptr, *dlp = nullptr;
  CU(sc.get(&d_err, 1));
  CU(cudaMemsetAsync(d_err, 0, sizeof(int), s));
  if (ns > 0) {
    CU(sc.get(&dz, (size_t)ns * nv));
    CU(sc.get(&dzt, (size_t)ns * nv));
    CU(sc.get(&dout, (size_t)ns * nv));
    CU(sc.get(&dlp, (size_t)ns));
    CU(cudaMemcpyAsync(dz, z, sizeof(double) * ns * nv, cudaMemcpyDefault, s));
  }
  chol_upper_device(fp.S, npad, nv, jitter, d_err, s, &e->launches);
  CU(cudaGetLastError());
  cudaEventRecord(e->ev2, s);
  if (ns > 0) {
    draw_device(dz, dzt, fp.mean, fp.S, npad, ns, nv, dout, dlp, s, &e->launches);
    CU(cudaGetLastError());
  }
  cudaEventRecord(e->ev3, s);
  int h_err = 0;
  CU(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  if (h_err != 0) {
    char buf[160];
    snprintf(buf, sizeof(buf), "posterior covariance + jitter is not positive definite (pivot %d of %lld): raise jitter",
             h_err, nv);
    return fail(e, GPIS_ERR_NOT_PD, buf);
  }
  if (ns > 0) {
    CU(cudaMemcpyAsync(samples, dout, sizeof(double) * ns * nv, cudaMemcpyDefault, s));
    if (logpdf) CU(cudaMemcpyAsync(logpdf, dlp, sizeof(double) * ns, cudaMemcpyDefault, s));
  }
  if (chol)
    CU(cudaMemcpy2DAsync(chol, sizeof(double) * ldt, fp.S, sizeof(double) * npad, sizeof(double) * nv, (size_t)nv,
                         cudaMemcpyDefault, s));
  CU(cudaStreamSynchronize(s));
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, e->ev0, e->ev1) == cudaSuccess) e->cov_ms = ms;
  if (cudaEventElapsedTime(&ms, e->ev1, e->ev2) == cudaSuccess) e->chol_ms = ms;
  if (cudaEventElapsedTime(&ms, e->ev2, e->ev3) == cudaSuccess) e->draw_ms = ms;
  return GPIS_OK;
}

gpis_status gpis_batch_select(const gpis_config* cfg_in, const int32_t* dims, const double* origin, const double* spacing,
                              int32_t n_objects, const double* tsdf, int64_t first_index, int32_t K, int64_t* ids,
                              int32_t* n_selected, int32_t* n_in_model, double* mean, double* var, uint8_t* flags,
                              double* device_ms, void* stream) {
  gpis_config cfg_v, *cfg = &cfg_v;
  if (!read_config(cfg_in, cfg) || !dims || !tsdf || !ids || !n_selected) return GPIS_ERR_INVALID;
  if (cfg->batch_kernel < 0 || cfg->batch_kernel > GPIS_BATCH_DFMA) return GPIS_ERR_INVALID;
  const int D = cfg->dim;
  if (D < 2 || D > 3 || cfg->use_gradients || cfg->precision != GPIS_PREC_FP64 || cfg->beta_mode != GPIS_BETA_FIXED ||
      cfg->world_size > 1)
    return GPIS_ERR_INVALID;            // function values, fp64, fixed beta, one rank (shard the OBJECTS across ranks)
  BatchParams P;
  memset(&P, 0, sizeof(P));
  P.nx = dims[0];
  P.ny = dims[1];
  P.nz = D == 3 ? dims[2] : 1;
  if (P.nx < 1 || P.ny < 1 || P.nz < 1 || P.nx > BAX || P.ny > BAX || P.nz > BAX) return GPIS_ERR_INVALID;
  P.N = (long long)P.nx * P.ny * P.nz;
  if (n_objects < 0 || K < 1 || K > BRMAX || first_index < 0 || first_index >= P.N) return GPIS_ERR_INVALID;
  if (!(cfg->ell > 0.0) || !(cfg->sf2 > 0.0) || !(cfg->noise >= 0.0) || !(cfg->beta >= 0.0)) return GPIS_ERR_INVALID;
  if (n_objects == 0) return GPIS_OK;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); return GPIS_ERR_NO_DEVICE; }
  gpis_engine* e = nullptr;             // the CU() macro records messages on a handle; there is none here
  if (cfg->device >= 0) CU(cudaSetDevice(cfg->device));
  int dev = 0;
  CU(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, dev));
  if (prop.major < 10) return GPIS_ERR_NO_DEVICE;
  cudaStream_t s = (cudaStream_t)stream;
  // One block per SM, but no more blocks than the rounds need: objects cost the same, so the launch takes
  // ceil(n / SMs) rounds whatever the grid, and the fewest blocks that still finish in that many rounds keep the fewest
  // objects' state (557 KB each at 32^3) competing for L2 at a time (512 objects on 148 SMs: 4 rounds of 128).
  const int sms = prop.multiProcessorCount;
  const int rounds = (n_objects + sms - 1) / sms;
  const int n_blocks = n_objects < sms ? n_objects : (n_objects + rounds - 1) / rounds;
  const size_t BN = (size_t)n_objects * (size_t)P.N;
  P.n_objects = n_objects;
  P.K = K;
  P.first = first_index;
  for (int d = 0; d < 3; ++d) {
    P.org[d] = (origin && d < D) ? origin[d] : 0.0;
    P.h[d] = (spacing && d < D) ? spacing[d] : 1.0;
    P.mean_w[d] = d < D ? cfg->mean_w[d] : 0.0;
  }
  P.inv_ell2 = 1.0 / (cfg->ell * cfg->ell);
  P.sf2 = cfg->sf2;
  P.noise = cfg->noise;
  P.level = cfg->level;
  P.eps = cfg->eps;
  P.beta = cfg->beta;
  P.variance_mode = cfg->variance_mode;
  P.mean_c = cfg->mean_c;
  Scratch sc;
  double *d_tsdf = nullptr, *d_mu = nullptr, *d_var = nullptr;
  unsigned char* d_flags = nullptr;
  long long* d_ids = nullptr;
  int* d_cnt = nullptr;                 // [3][n_objects]: n_sel, n_model, err
  CU(sc.get(&d_tsdf, BN));
  CU(sc.get(&d_mu, BN));
  CU(sc.get(&d_var, BN));
  CU(sc.get(&d_flags, BN));
  CU(sc.get(&d_ids, (size_t)n_objects * K));
  CU(sc.get(&d_cnt, (size_t)3 * n_objects));
  CU(sc.get(&P.W, (size_t)n_blocks * BRMAX * BRMAX));
  CU(sc.get(&P.Wt, (size_t)n_blocks * BRMAX * BRMAX));
  CU(cudaMemcpyAsync(d_tsdf, tsdf, sizeof(double) * BN, cudaMemcpyDefault, s));
  CU(cudaMemsetAsync(d_ids, 0xff, sizeof(long long) * (size_t)n_objects * K, s));        // -1: not selected
  P.tsdf = d_tsdf;
  P.mu = d_mu;
  P.var = d_var;
  P.flags = d_flags;
  P.ids = d_ids;
  P.n_sel = d_cnt;
  P.n_model = d_cnt + n_objects;
  P.err = d_cnt + 2 * (size_t)n_objects;
  // inner loops of the kernel (gpis_batch.cuh): cfg->batch_kernel; GPIS_BATCH_AUTO = the tensor-pipe (DMMA) form,
  // the fastest of the three on a B200 (512 x 32^3 x 256: 93 ms against 126 ms DFMA-tuned and 169 ms plain,
  // profiles/r02_batch_variants.json; same picks in all)
  const int fast = cfg->batch_kernel == GPIS_BATCH_PLAIN ? 0 : cfg->batch_kernel == GPIS_BATCH_DFMA ? 1 : 2;
  CU(cudaFuncSetAttribute((const void*)k_batch_select<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BATCH_SMEM));
  CU(cudaFuncSetAttribute((const void*)k_batch_select<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BATCH_SMEM));
  CU(cudaFuncSetAttribute((const void*)k_batch_select_dmma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BATCH_SMEM));
  struct EventPair {                    // released on every return path
    cudaEvent_t a = nullptr, b = nullptr;
    ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
  } ev;
  CU(cudaEventCreate(&ev.a));
  CU(cudaEventCreate(&ev.b));
  cudaEvent_t e0 = ev.a, e1 = ev.b;
  cudaEventRecord(e0, s);
  launch_batch_select(P, n_blocks, fast, s);
  cudaError_t ce = cudaGetLastError();
  cudaEventRecord(e1, s);
  std::vector<int> h_cnt((size_t)3 * n_objects);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(h_cnt.data(), d_cnt, sizeof(int) * h_cnt.size(), cudaMemcpyDeviceToHost, s);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(ids, d_ids, sizeof(long long) * (size_t)n_objects * K, cudaMemcpyDefault, s);
  if (ce == cudaSuccess && mean) ce = cudaMemcpyAsync(mean, d_mu, sizeof(double) * BN, cudaMemcpyDefault, s);
  if (ce == cudaSuccess && var) ce = cudaMemcpyAsync(var, d_var, sizeof(double) * BN, cudaMemcpyDefault, s);
  if (ce == cudaSuccess && flags) ce = cudaMemcpyAsync(flags, d_flags, BN, cudaMemcpyDefault, s);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
  float ms = 0.f;
  if (ce == cudaSuccess && device_ms && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) *device_ms = ms;
  if (ce != cudaSuccess) return GPIS_ERR_CUDA;
  bool not_pd = false;
  for (int k = 0; k < n_objects; ++k) {
    n_selected[k] = h_cnt[k];
    if (n_in_model) n_in_model[k] = h_cnt[(size_t)n_objects + k];
    not_pd = not_pd || h_cnt[2 * (size_t)n_objects + k] != 0;
  }
  return not_pd ? GPIS_ERR_NOT_PD : GPIS_OK;
}

gpis_status gpis_get_sampling_timing(gpis_handle e, double* cov_ms, double* chol_ms, double* draw_ms) {
  if (!e) return GPIS_ERR_INVALID;
  GUARD(e);
  if (cov_ms) *cov_ms = e->cov_ms;
  if (chol_ms) *chol_ms = e->chol_ms;
  if (draw_ms) *draw_ms = e->draw_ms;
  return GPIS_OK;
}

namespace {
__global__ void k_split_flags(const unsigned char* f, long long n, unsigned char* hi, unsigned char* lo,
                              unsigned char* ac) {
  const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (j < n) {
    const unsigned char v = f[j];
    if (hi) hi[j] = (v & F_HIGH) ? 1 : 0;
    if (lo) lo[j] = (v & F_LOW) ? 1 : 0;
    if (ac) ac[j] = (v & F_ACTIVE) ? 1 : 0;
  }
}
}  // namespace

gpis_status gpis_get_levelset(gpis_handle e, double* ambiguity, uint8_t* high, uint8_t* low, uint8_t* active,
                              void* stream) {
  if (!e) return GPIS_ERR_INVALID;
  GUARD(e);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t N = (size_t)e->P.N;
  if (!N) return GPIS_OK;
  double* tmp_amb = nullptr;
  unsigned char* tmp = nullptr;
  cudaError_t ce = cudaMalloc(&tmp, 4 * N);
  if (ce == cudaSuccess && e->backend == 2) ce = cudaMalloc(&tmp_amb, sizeof(double) * N);
  const double* amb_src = e->P.amb;
  const unsigned char* flag_src = e->P.flags;
  if (ce == cudaSuccess && e->backend == 2) {     // state is stored in tile/fragment order: export it
    GPIS_LAUNCH_SMEM(k_grid_export, (unsigned)((N + 255) / 256), 256, 0, s, e->P, nullptr, nullptr, tmp_amb, tmp + 3 * N);
    e->launches++;
    amb_src = tmp_amb;
    flag_src = tmp + 3 * N;
    ce = cudaGetLastError();
  }
  if (ce == cudaSuccess && ambiguity) ce = cudaMemcpyAsync(ambiguity, amb_src, sizeof(double) * N, cudaMemcpyDefault, s);
  if (ce == cudaSuccess && (high || low || active)) {
    GPIS_LAUNCH_SMEM(k_split_flags, (unsigned)((N + 255) / 256), 256, 0, s, flag_src, (long long)N, tmp, tmp + N, tmp + 2 * N);
    e->launches++;
    ce = cudaGetLastError();
    if (ce == cudaSuccess && high) ce = cudaMemcpyAsync(high, tmp, N, cudaMemcpyDefault, s);
    if (ce == cudaSuccess && low) ce = cudaMemcpyAsync(low, tmp + N, N, cudaMemcpyDefault, s);
    if (ce == cudaSuccess && active) ce = cudaMemcpyAsync(active, tmp + 2 * N, N, cudaMemcpyDefault, s);
  }
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
  if (tmp) cudaFree(tmp);
  if (tmp_amb) cudaFree(tmp_amb);
  if (ce != cudaSuccess) return fail(e, GPIS_ERR_CUDA, cudaGetErrorString(ce));
  return GPIS_OK;
}

gpis_status gpis_get_candidate_posterior(gpis_handle e, double* mean, double* var, void* stream) {
  if (!e) return GPIS_ERR_INVALID;
  GUARD(e);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t N = (size_t)e->P.N;
  if (!N) return GPIS_OK;
  if (e->backend == 2) {
    // export from tile/fragment order; straight into the caller's buffers when they are device memory
    cudaPointerAttributes at;
    auto on_device = [&](const void* p) {
      if (!p) return true;
      if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
      return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
    };
    const bool dev = on_device(mean) && on_device(var);
    double* tmp = nullptr;
    if (!dev) CU(cudaMalloc(&tmp, sizeof(double) * 2 * N));
    GPIS_LAUNCH_SMEM(k_grid_export, (unsigned)((N + 255) / 256), 256, 0, s, e->P, dev ? mean : tmp, dev ? var : tmp + N, nullptr, nullptr);
    e->launches++;
    cudaError_t ce = cudaGetLastError();
    if (!dev) {
      if (ce == cudaSuccess && mean) ce = cudaMemcpyAsync(mean, tmp, sizeof(double) * N, cudaMemcpyDeviceToHost, s);
      if (ce == cudaSuccess && var) ce = cudaMemcpyAsync(var, tmp + N, sizeof(double) * N, cudaMemcpyDeviceToHost, s);
    }
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
    if (tmp) cudaFree(tmp);
    if (ce != cudaSuccess) return fail(e, GPIS_ERR_CUDA, cudaGetErrorString(ce));
    return GPIS_OK;
  }
  if (mean) CU(cudaMemcpyAsync(mean, e->P.mu, sizeof(double) * N, cudaMemcpyDefault, s));
  if (var) CU(cudaMemcpyAsync(var, e->P.var, sizeof(double) * N, cudaMemcpyDefault, s));
  CU(cudaStreamSynchronize(s));
  return GPIS_OK;
}

gpis_status gpis_get_stats(gpis_handle e, int64_t* kernel_launches, int64_t* live_tiles, int64_t* iterations,
                           int64_t* panel_bytes) {
  if (!e) return GPIS_ERR_INVALID;
  GUARD(e);
  if (kernel_launches) *kernel_launches = e->launches;
  if (live_tiles || iterations) {
    CU(cudaDeviceSynchronize());        // the state may have been produced on any stream
    gpis_status rc = read_state(e, 0);
    if (rc != GPIS_OK && rc != GPIS_ERR_NOT_PD) return rc;
    if (live_tiles) *live_tiles = e->host_state.n_live[e->host_state.iter & 1];
    if (iterations) *iterations = e->host_state.iter;
  }
  if (panel_bytes) *panel_bytes = (int64_t)((size_t)e->P.n_tiles * (size_t)e->P.tile_stride * sizeof(double));
  return GPIS_OK;
}

gpis_status gpis_get_timing(gpis_handle e, double* select_ms, double* predict_ms, double* update_ms,
                            int64_t* update_launches) {
  if (!e) return GPIS_ERR_INVALID;
  GUARD(e);
  if (select_ms) *select_ms = e->select_ms;
  if (predict_ms) *predict_ms = e->predict_ms;
  if (update_ms) *update_ms = e->update_ms;
  if (update_launches) *update_launches = e->update_launches;
  return GPIS_OK;
}

gpis_status gpis_get_phase_timing(gpis_handle e, double* append_ms, double* main_ms, double* epilogue_ms) {
  if (!e) return GPIS_ERR_INVALID;
  GUARD(e);
  if (append_ms) *append_ms = e->append_ms;
  if (main_ms) *main_ms = e->main_ms;
  if (epilogue_ms) *epilogue_ms = e->epilogue_ms;
  return GPIS_OK;
}

// ---------------------------------------------------------------------------------------------------
// float entry points: the reference's ABI is fp32 (include/gpu_active_set_selector.hpp:41-46,
// include/active_set_selection_types.h:16-41).  The engine computes in fp64; these convert on the device.
// ---------------------------------------------------------------------------------------------------
namespace {
__global__ void k_f32_to_f64(const float* __restrict__ a, double* __restrict__ b, size_t n) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) b[i] = (double)a[i];
}
__global__ void k_f64_to_f32(const double* __restrict__ a, float* __restrict__ b, size_t n) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) b[i] = (float)a[i];
}
// y_a = m(x_a) + (L g)[row of f(x_a)]: the targets of the model points from the replicated factor
__global__ void k_active_targets(EngineParams P, int m, int NB, int D, float* inputs, float* targets, int ld_out) {
  const int a = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (a >= m) return;
  const long long row = (long long)a * NB;
  double acc = 0.0;
  for (long long c = lane; c <= row; c += 32) acc += P.L[row * P.R_cap + c] * P.g[c];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    double mf = P.mean_c;
    for (int d = 0; d < D; ++d) mf += P.mean_w[d] * P.AX[d][a];
    if (targets) targets[a] = (float)(mf + acc);
    if (inputs)
      for (int d = 0; d < D; ++d) inputs[a + (long long)d * ld_out] = (float)P.AX[d][a];      // active_inputs[a + d*max_active]
  }
}
bool ptr_on_device(const void* p) {
  if (!p) return true;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}
// n floats at a host or device pointer -> doubles on the device (scratch)
cudaError_t stage_in_f32(Scratch& sc, const float* src, size_t n, cudaStream_t s, double** out) {
  *out = nullptr;
  if (!src || !n) return cudaSuccess;
  const float* dsrc = src;
  float* tmp = nullptr;
  cudaError_t ce = cudaSuccess;
  if (!ptr_on_device(src)) {
    if ((ce = sc.get(&tmp, n)) != cudaSuccess) return ce;
    if ((ce = cudaMemcpyAsync(tmp, src, sizeof(float) * n, cudaMemcpyHostToDevice, s)) != cudaSuccess) return ce;
    dsrc = tmp;
  }
  if ((ce = sc.get(out, n)) != cudaSuccess) return ce;
  GPIS_LAUNCH_SMEM(k_f32_to_f64, (unsigned)((n + 255) / 256), 256, 0, s, dsrc, *out, n);
  return cudaGetLastError();
}
// n doubles on the device -> floats at a host or device pointer
cudaError_t stage_out_f32(Scratch& sc, const double* dsrc, float* dst, size_t n, cudaStream_t s) {
  if (!dst || !n) return cudaSuccess;
  float* ddst = dst;
  cudaError_t ce = cudaSuccess;
  const bool dev = ptr_on_device(dst);
  if (!dev && (ce = sc.get(&ddst, n)) != cudaSuccess) return ce;
  GPIS_LAUNCH_SMEM(k_f64_to_f32, (unsigned)((n + 255) / 256), 256, 0, s, dsrc, ddst, n);
  if ((ce = cudaGetLastError()) != cudaSuccess) return ce;
  if (!dev) ce = cudaMemcpyAsync(dst, ddst, sizeof(float) * n, cudaMemcpyDeviceToHost, s);
  return ce;
}
}  // namespace

gpis_status gpis_set_candidates_f32(gpis_handle e, const float* pts, const float* tsdf, const float* normals, const float* noise,
                                    const int64_t* ids, void* stream) {
  if (!e) return GPIS_ERR_INVALID;
  GUARD(e);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t N = (size_t)e->P.N;
  Scratch sc;
  double *dp = nullptr, *dt = nullptr, *dn = nullptr, *dz = nullptr;
  CU(stage_in_f32(sc, pts, N * e->D, s, &dp));
  CU(stage_in_f32(sc, tsdf, N, s, &dt));
  CU(stage_in_f32(sc, normals, N * e->D, s, &dn));
  CU(stage_in_f32(sc, noise, N, s, &dz));
  gpis_status rc = gpis_set_candidates(e, pts ? dp : nullptr, tsdf ? dt : nullptr, dn, dz, ids, stream);
  CU(cudaStreamSynchronize(s));
  return rc;
}

gpis_status gpis_set_grid_candidates_f32(gpis_handle e, const int32_t* dims, const double* origin, const double* spacing,
                                         int32_t z0, int32_t nz_total, const float* tsdf, const float* normals,
                                         const float* noise, void* stream) {
  if (!e) return GPIS_ERR_INVALID;
  GUARD(e);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t N = (size_t)e->P.N;
  Scratch sc;
  double *dt = nullptr, *dn = nullptr, *dz = nullptr;
  CU(stage_in_f32(sc, tsdf, N, s, &dt));
  CU(stage_in_f32(sc, normals, N * e->D, s, &dn));
  CU(stage_in_f32(sc, noise, N, s, &dz));
  gpis_status rc = gpis_set_grid_candidates(e, dims, origin, spacing, z0, nz_total, tsdf ? dt : nullptr, dn, dz, stream);
  CU(cudaStreamSynchronize(s));
  return rc;
}

gpis_status gpis_get_active_f32(gpis_handle e, int32_t max_size, float* active_inputs, float* active_targets,
                                int32_t* n_selected, int32_t* n_in_model, void* stream) {
  if (!e) return GPIS_ERR_INVALID;
  GUARD(e);
  cudaStream_t s = (cudaStream_t)stream;
  gpis_status rc = read_state(e, s);
  if (rc != GPIS_OK) return rc;
  const int m = e->host_state.m;
  if (n_selected) *n_selected = e->host_state.n_sel;
  if (n_in_model) *n_in_model = m;
  if (m == 0 || (!active_inputs && !active_targets)) return GPIS_OK;
  if (max_size < m) return fail(e, GPIS_ERR_INVALID, "max_size smaller than the number of model points");
  Scratch sc;
  float *din = nullptr, *dtg = nullptr;
  CU(sc.get(&din, (size_t)max_size * e->D));
  CU(sc.get(&dtg, (size_t)max_size));
  CU(cudaMemsetAsync(din, 0, sizeof(float) * (size_t)max_size * e->D, s));
  CU(cudaMemsetAsync(dtg, 0, sizeof(float) * (size_t)max_size, s));
  GPIS_LAUNCH_SMEM(k_active_targets, (unsigned)((m + 7) / 8), 256, 0, s, e->P, m, e->NB, e->D, din, dtg, (int)max_size);
  CU(cudaGetLastError());
  e->launches++;
  if (active_inputs) CU(cudaMemcpyAsync(active_inputs, din, sizeof(float) * (size_t)max_size * e->D, cudaMemcpyDefault, s));
  if (active_targets) CU(cudaMemcpyAsync(active_targets, dtg, sizeof(float) * (size_t)max_size, cudaMemcpyDefault, s));
  CU(cudaStreamSynchronize(s));
  return GPIS_OK;
}

gpis_status gpis_predict_f32(gpis_handle e, const float* query, int64_t nq, float* mean, float* var, float* normals, void* stream) {
  if (!e) return GPIS_ERR_INVALID;
  GUARD(e);
  if (nq < 0 || (nq && (!query || !mean || !var))) return fail(e, GPIS_ERR_INVALID, "query, mean and var are required");
  if (nq == 0) return GPIS_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = (size_t)nq;
  Scratch sc;
  double *dq = nullptr, *dout = nullptr;
  CU(stage_in_f32(sc, query, n * e->D, s, &dq));
  CU(sc.get(&dout, n * (2 + e->D)));
  gpis_status rc = gpis_predict(e, dq, nq, dout, dout + n, normals ? dout + 2 * n : nullptr, stream);
  if (rc != GPIS_OK) return rc;
  CU(stage_out_f32(sc, dout, mean, n, s));
  CU(stage_out_f32(sc, dout + n, var, n, s));
  if (normals) CU(stage_out_f32(sc, dout + 2 * n, normals, n * e->D, s));
  CU(cudaStreamSynchronize(s));
  return GPIS_OK;
}

gpis_status gpis_get_candidate_posterior_f32(gpis_handle e, float* mean, float* var, void* stream) {
  if (!e) return GPIS_ERR_INVALID;
  GUARD(e);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t N = (size_t)e->P.N;
  if (!N) return GPIS_OK;
  Scratch sc;
  double* d = nullptr;
  CU(sc.get(&d, 2 * N));
  gpis_status rc = gpis_get_candidate_posterior(e, d, d + N, stream);
  if (rc != GPIS_OK) return rc;
  CU(stage_out_f32(sc, d, mean, N, s));
  CU(stage_out_f32(sc, d + N, var, N, s));
  CU(cudaStreamSynchronize(s));
  return GPIS_OK;
}

gpis_status gpis_build_tsdf(const uint8_t* inside, int32_t rows, int32_t cols, const gpis_tsdf_params* prm, const double* u,
                            double* tsdf, double* normals, double* noise, int32_t* valid_idx, int32_t* n_valid,
                            int32_t device, void* stream) {
  gpis_engine* e = nullptr;             // the CU() macro records messages on a handle; there is none here
  if (!inside || !prm || rows < 1 || cols < 1 || (long long)rows * cols > 0x7fffffffLL) return fail(e, GPIS_ERR_INVALID, "bad image");
  if (prm->struct_size != (int32_t)sizeof(gpis_tsdf_params)) return fail(e, GPIS_ERR_INVALID, "gpis_tsdf_params.struct_size");
  if (prm->edge_win < 0 || prm->noise_grad_mode < 0 || prm->noise_grad_mode > 4 || !(prm->noise_scale > 0.0))
    return fail(e, GPIS_ERR_INVALID, "edge_win / noise_grad_mode / noise_scale out of range");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); return GPIS_ERR_NO_DEVICE; }
  int dev = -1;
  CU(cudaGetDevice(&dev));
  DeviceGuard guard(device >= 0 ? device : dev);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = (size_t)rows * cols;
  TsdfParams P;
  P.rows = rows; P.cols = cols;
  P.edge_win = prm->edge_win; P.specular = prm->specular_noise; P.grad_mode = prm->noise_grad_mode;
  P.noise_scale = prm->noise_scale; P.occlusion_scale = prm->occlusion_scale; P.interior_rate = prm->interior_rate;
  P.sparsity_rate = prm->sparsity_rate; P.horiz_scale = prm->horiz_scale; P.vert_scale = prm->vert_scale;
  for (int k = 0; k < 3; ++k) {
    P.y_low[k] = prm->occ_y_low[k]; P.y_high[k] = prm->occ_y_high[k];
    P.x_low[k] = prm->occ_x_low[k]; P.x_high[k] = prm->occ_x_high[k];
  }
  Scratch sc;
  unsigned char *d_in = nullptr, *d_valid = nullptr;
  double *d_t = nullptr, *d_nz = nullptr, *d_g = nullptr, *d_u = nullptr;
  int *d_idx = nullptr, *d_cnt = nullptr;
  CU(sc.get(&d_in, n));
  CU(sc.get(&d_valid, n));
  CU(sc.get(&d_t, n));
  CU(sc.get(&d_nz, n));
  CU(sc.get(&d_g, 2 * n));
  CU(sc.get(&d_idx, n));
  CU(sc.get(&d_cnt, 1));
  CU(cudaMemcpyAsync(d_in, inside, n, cudaMemcpyDefault, s));
  if (u) {
    CU(sc.get(&d_u, 3 * n));
    CU(cudaMemcpyAsync(d_u, u, sizeof(double) * 3 * n, cudaMemcpyDefault, s));
  }
  const unsigned blocks = (unsigned)((n + 255) / 256);
  GPIS_LAUNCH_SMEM(k_tsdf_levels, blocks, 256, 0, s, d_in, d_t, rows, cols);
  GPIS_LAUNCH_SMEM(k_tsdf_noise_normals, blocks, 256, 0, s, d_t, d_u, d_nz, d_g, d_g + n, d_valid, P);
  GPIS_LAUNCH_SMEM(k_tsdf_compact, 1, 1024, 0, s, d_valid, (long long)n, d_idx, d_cnt);
  CU(cudaGetLastError());
  if (tsdf) CU(cudaMemcpyAsync(tsdf, d_t, sizeof(double) * n, cudaMemcpyDefault, s));
  if (normals) CU(cudaMemcpyAsync(normals, d_g, sizeof(double) * 2 * n, cudaMemcpyDefault, s));
  if (noise) CU(cudaMemcpyAsync(noise, d_nz, sizeof(double) * n, cudaMemcpyDefault, s));
  if (valid_idx) CU(cudaMemcpyAsync(valid_idx, d_idx, sizeof(int) * n, cudaMemcpyDefault, s));
  if (n_valid) CU(cudaMemcpyAsync(n_valid, d_cnt, sizeof(int), cudaMemcpyDefault, s));
  CU(cudaStreamSynchronize(s));         // the scratch is released on return
  return GPIS_OK;
}

gpis_status gpis_get_persist_balance(gpis_handle e, double* ms, int32_t capacity, int32_t* n_blocks) {
  if (!e || !n_blocks) return GPIS_ERR_INVALID;
  GUARD(e);
  *n_blocks = e->d_phase_clk ? e->num_sms : 0;
  if (!ms || !e->d_phase_clk) return GPIS_OK;
  if (capacity < 2 * e->num_sms) return fail(e, GPIS_ERR_INVALID, "capacity smaller than 2 * blocks");
  std::vector<unsigned long long> clk(2 * (size_t)e->num_sms);
  CU(cudaMemcpy(clk.data(), e->d_phase_clk + 16, sizeof(unsigned long long) * clk.size(), cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < clk.size(); ++i) ms[i] = (double)clk[i] * 1e-6;
  return GPIS_OK;
}

gpis_status gpis_get_persist_timing(gpis_handle e, double* ms16) {
  if (!e || !ms16) return GPIS_ERR_INVALID;
  for (int i = 0; i < 16; ++i) ms16[i] = e->persist_ms[i];
  return GPIS_OK;
}

gpis_status gpis_set_profiling(gpis_handle e, int32_t on) {
  if (!e) return GPIS_ERR_INVALID;
  GUARD(e);
  e->profiling = on != 0;
  return GPIS_OK;
}

gpis_status gpis_get_history(gpis_handle e, int32_t* rows, int32_t* live_tiles, int32_t* scored, int32_t* n) {
  if (!e) return GPIS_ERR_INVALID;
  GUARD(e);
  CU(cudaDeviceSynchronize());          // the history may have been produced on any stream
  gpis_status rc = read_state(e, 0);
  if (rc != GPIS_OK && rc != GPIS_ERR_NOT_PD) return rc;
  const int it = std::min(e->host_state.iter, e->P.M_cap + 1);
  if (n) *n = it;
  const size_t cap = (size_t)e->P.M_cap + 1;
  if (it > 0) {
    if (rows) CU(cudaMemcpy(rows, e->P.hist, sizeof(int) * it, cudaMemcpyDefault));
    if (live_tiles) CU(cudaMemcpy(live_tiles, e->P.hist + cap, sizeof(int) * it, cudaMemcpyDefault));
    if (scored) CU(cudaMemcpy(scored, e->P.hist + 2 * cap, sizeof(int) * it, cudaMemcpyDefault));
  }
  return GPIS_OK;
}

}  // extern "C"
