/* gpis_b200.h -- C ABI of the B200-native GPIS active-set-selection + prediction engine.
 *
 * Drop-in boundary for the reference's accelerated path.  Each entry point names the
 * reference interface it replaces (paths relative to the reference tree):
 *
 *   include/active_set_buffers.h:7-16, include/max_subset_buffers.h:7-12,
 *   include/classification_buffers.h:7-8      construct_* / free_* / update_* / find_best_*
 *   include/gpu_active_set_selector.hpp:32-47   GpuActiveSetSelector::SelectFromGrid / SelectChol
 *   matlab/set_selection/select_active_set.m:1-2, select_active_set_straddle.m:1-147
 *   matlab/core/create_gpis.m, gp_mean.m, gp_cov.m, predict_2d_grid.m
 *
 * Conventions (they fix the reference's: it printf()s and exit(0)s on error,
 * include/cuda_macros.h:10-35, and never fills its output arrays):
 *   - every call returns a gpis_status; nothing prints, nothing exits;
 *   - every pointer argument may be a HOST or a DEVICE pointer (copies are
 *     cudaMemcpyAsync(..., cudaMemcpyDefault) on the given stream); the engine owns all
 *     of its device memory;
 *   - layouts are struct-of-arrays exactly like the reference: pts[i + d*N]
 *     (src/gpu_active_set_selector.cpp:96-101); normals likewise;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     calls are asynchronous on it unless they return values through host pointers,
 *     in which case they synchronise the stream before returning;
 *   - there is no CPU fallback: without a CUDA device gpis_create fails;
 *   - kernels that wait on one another inside a launch (the persistent selection kernel, the fused
 *     factor append) are COOPERATIVE launches: the driver guarantees their blocks are co-resident or
 *     refuses the launch, in which case the engine uses its plain multi-kernel forms.  To run several
 *     handles concurrently on one device (gpis_b200/batch.py does) ask for those forms up front:
 *     gpis_config.loop_mode = GPIS_LOOP_STEPS, append_mode = GPIS_APPEND_TWO_KERNELS.
 *   - behaviour is selected by gpis_config fields only; the library reads no environment variable.
 */
#ifndef GPIS_B200_H
#define GPIS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gpis_engine* gpis_handle;
typedef int32_t gpis_status;

enum {
  GPIS_OK = 0,
  GPIS_ERR_INVALID = -1,   /* bad argument / config                          */
  GPIS_ERR_CUDA = -2,      /* CUDA runtime error (see gpis_last_error)       */
  GPIS_ERR_NOMEM = -3,     /* device allocation failed                       */
  GPIS_ERR_STATE = -4,     /* call out of order (e.g. select before set)     */
  GPIS_ERR_NOT_PD = -5,    /* covariance block lost positive definiteness    */
  GPIS_ERR_NO_DEVICE = -6, /* no usable CUDA device: there is no CPU path    */
  GPIS_ERR_TIMEOUT = -7    /* an in-kernel wait (peer record, grid barrier) timed out */
};

enum { GPIS_BETA_FIXED = 0,      /* select_active_set_straddle.m:10 (varScale fixed)            */
       GPIS_BETA_GOTOVOS = 1,    /* matlab/lse/discrete_gp_lse.m:87  2 log(N t pi^2/(6 delta))  */
       GPIS_BETA_SELECTCHOL = 2  /* gpu_active_set_selector.cpp:493,550  2 log(N pi^2 (k+1)^2/(6 tol)) */ };
enum { GPIS_VAR_LATENT = 0,      /* gp_cov.m:6 (beta = 0 on the test block)                     */
       GPIS_VAR_PREDICTIVE = 1   /* max_subset_buffers.cu:113-115 (k(x,x) + noise - reduction)  */ };
enum { GPIS_PREC_FP64 = 0, GPIS_PREC_TF32 = 1 };
/* How gpis_select drives the greedy loop of a lattice (grid) engine. */
enum { GPIS_LOOP_AUTO = 0,       /* one persistent cooperative kernel per selection when eligible   */
       GPIS_LOOP_STEPS = 1       /* per-step launches (captured CUDA graph)                         */ };
/* Factor-append form of the per-step path (function-only lattice models up to 4096 rows). */
enum { GPIS_APPEND_AUTO = 0,         /* fused: W pass + in-kernel grid barrier + new row, one cooperative launch */
       GPIS_APPEND_TWO_KERNELS = 1,  /* the same pass as two plain launches (safe next to other work on the GPU) */
       GPIS_APPEND_TWO_PASS = 2      /* rows of W, then of W' (the form gradient / larger models always use)     */ };
/* Inner loops of gpis_batch_select. */
enum { GPIS_BATCH_AUTO = 0, GPIS_BATCH_PLAIN = 1, GPIS_BATCH_DMMA = 2, GPIS_BATCH_DFMA = 3 };

/* Replaces GaussianProcessHyperparams + the scalar arguments of SelectChol
 * (include/active_set_selection_types.h:10-13, gpu_active_set_selector.hpp:41-46) and the
 * MATLAB trainingParams / hyp structs (matlab/grasp_metric/setup_and_run_gpis_experiment.m:59-73). */
typedef struct gpis_config {
  int32_t struct_size;      /* = sizeof(gpis_config); ABI guard                              */
  int32_t dim;              /* D, 1..3 (MAX_DIM_INPUT is 10 in the reference; GPIS uses 2,3) */
  int64_t num_candidates;   /* N on THIS rank (a shard when world_size > 1)                  */
  int64_t total_candidates; /* N over all ranks (beta schedules use it); 0 = num_candidates  */
  int64_t id_base;          /* id of local candidate 0 when no explicit ids are given        */
  int32_t max_active;       /* capacity in points (activeSetSize / maxSize)                  */
  int32_t use_gradients;    /* normals observed: D+1 rows per point (se_cov_derivative.m)    */
  int32_t precision;        /* GPIS_PREC_*                                                   */
  int32_t beta_mode;        /* GPIS_BETA_*                                                   */
  int32_t variance_mode;    /* GPIS_VAR_*                                                    */
  int32_t rank, world_size; /* candidate sharding; the factor is replicated                  */
  int32_t device;           /* CUDA device ordinal, -1 = current                             */
  double ell, sf2;          /* covSEiso: sf2 * exp(-|x-z|^2 / (2 ell^2))                     */
  double mean_w[3], mean_c; /* meanSum{meanLinear, meanConst}                                */
  double noise;             /* scalar noise variance (exp(2*hyp.lik)) when no per-point noise */
  double level, eps, beta;  /* trainingParams.levelSet / eps / beta                          */
  double beta_delta;        /* delta (GOTOVOS) or tolerance (SELECTCHOL)                     */
  /* -- fields added after version 0.1; a caller compiled against the shorter struct passes its own
   *    sizeof in struct_size and gets the defaults (all zero) for what it does not know -- */
  int32_t loop_mode;        /* GPIS_LOOP_*                                                   */
  int32_t append_mode;      /* GPIS_APPEND_*                                                 */
  int32_t batch_kernel;     /* GPIS_BATCH_* (gpis_batch_select only)                         */
  int32_t spin_timeout_ms;  /* longest in-kernel wait for a peer's record or a grid barrier before the
                             * selection is abandoned with GPIS_ERR_TIMEOUT; 0 = 30000              */
} gpis_config;
/* sizeof(gpis_config) of version 0.1 (up to and including beta_delta) */
#define GPIS_CONFIG_SIZE_V1 152

const char* gpis_version(void);
const char* gpis_status_string(gpis_status s);
/* Text of the last error on this handle ("" if none); h == NULL: the message of the last failed
 * gpis_create / gpis_batch_select on the calling thread. */
const char* gpis_last_error(gpis_handle h);

/* construct_active_set_buffers + construct_max_subset_buffers + construct_classification_buffers
 * (active_set_buffers.cu:12-37, max_subset_buffers.cu:12-33, classification_buffers.cu:6-14). */
gpis_status gpis_create(gpis_handle* out, const gpis_config* cfg);
/* free_*_buffers (same files). */
gpis_status gpis_destroy(gpis_handle h);

/* The H2D copy inside construct_max_subset_buffers (max_subset_buffers.cu:17-26) /
 * shapeParams.{points,tsdf,normals,noise} of select_active_set_straddle.m:4-7.
 * normals may be NULL iff use_gradients == 0; noise NULL = scalar cfg.noise;
 * ids NULL = id_base + local index.  ids break ties (lowest wins) and are what
 * gpis_get_active reports.  Resets any previous selection. */
gpis_status gpis_set_candidates(gpis_handle h, const double* pts, const double* tsdf,
                                const double* normals, const double* noise,
                                const int64_t* ids, void* stream);

/* GpuActiveSetSelector::SelectFromGrid (gpu_active_set_selector.cpp:170-201, ReadCsv :75-113):
 * the candidates are ALL points of a regular lattice, x fastest (IJK_TO_LINEAR,
 * active_set_selection_types.h:6): point (i,j,k) = origin + spacing * (i, j, z0 + k),
 * dims = {nx, ny, nz_local}.  tsdf / normals / noise are indexed by the linear lattice index.
 * z0 / nz_total describe this rank's z-slab of the whole lattice (0 / dims[2] when unsharded).
 * This selects the grid backend: the SE kernel's separability over axes turns each selection
 * step into a dense contraction on the fp64 tensor pipe and the running posterior is exact for
 * every lattice point (gpis_get_candidate_posterior is then the full posterior grid). */
gpis_status gpis_set_grid_candidates(gpis_handle h, const int32_t* dims, const double* origin,
                                     const double* spacing, int32_t z0, int32_t nz_total,
                                     const double* tsdf, const double* normals, const double* noise,
                                     void* stream);

/* GpuActiveSetSelector::SelectChol (gpu_active_set_selector.cpp:408-611) with the semantics
 * of select_active_set_straddle.m:62-126: greedy loop entirely on the device, at most K
 * points, stops early when nothing is unclassified.  first_local = local index of the first
 * point (straddle.m:36-37 draws it at random; main.cpp:77 seeds rand()), or -1 on the ranks
 * that do not own it.  With world_size > 1 this needs the peer exchange (gpis_peer_import);
 * without it sharded runs drive the step calls below and all-gather the exchange buffer
 * between gpis_step_update and gpis_step_append. */
gpis_status gpis_select(gpis_handle h, int64_t first_local, int32_t K, void* stream);

gpis_status gpis_select_begin(gpis_handle h, int64_t first_local_or_neg, void* stream);
/* update_active_set_buffers + SolveLinearSystemChol, as one incremental row append
 * (active_set_buffers.cu:212-240, gpu_active_set_selector.cpp:892-917). */
gpis_status gpis_step_append(gpis_handle h, void* stream);
/* GpPredictCholBatch over all candidates + find_best_active_set_candidate, fused
 * (gpu_active_set_selector.cpp:844-890, max_subset_buffers.cu:197-225). */
gpis_status gpis_step_update(gpis_handle h, void* stream);
/* Records the last pick (it is selected but, as in straddle.m:129-130, not in the model). */
gpis_status gpis_select_end(gpis_handle h, void* stream);
/* Per-rank exchange record: send = this rank's best candidate (score, id, point, its column
 * of L^-1 K); recv = world_size such records, rank-major.  With world_size == 1 recv == send. */
gpis_status gpis_exchange_buffers(gpis_handle h, void** send, void** recv, int64_t* bytes_per_rank);

/* Peer exchange (one process per GPU, world_size > 1): instead of a collective call per step,
 * every rank stores its record straight into all peers' exchange regions over NVLink (P2P stores
 * from the last CTA of the update kernel) and the next kernel waits on local flags -- the whole
 * loop then runs inside gpis_select with no host involvement.  gpis_peer_export writes this
 * rank's 64-byte CUDA IPC handle; gather the handles of all ranks (any transport) and pass the
 * world_size x 64 bytes, rank-major, to gpis_peer_import.  same_process != 0: `handles` is an
 * array of world_size device pointers (regions exported by engines of the same process). */
gpis_status gpis_peer_export(gpis_handle h, void* handle64);
gpis_status gpis_peer_import(gpis_handle h, const void* handles, int32_t same_process);

/* What SelectChol should have written to activeInputs/activeTargets and alpha.csv
 * (gpu_active_set_selector.cpp:584-586; the arrays are never filled there).
 * ids: selection order, n_selected entries (capacity max_active);
 * n_in_model <= n_selected rows-in-points of the factor (straddle.m:129-130). */
gpis_status gpis_get_active(gpis_handle h, int64_t* ids, int32_t* n_selected, int32_t* n_in_model,
                            void* stream);
/* Lower factor (Q = L L', point-major rows [f, d1 f, .., dD f] per point) into L[row*ldl + col],
 * alpha = Q^-1 (y - m) in the same row order.  Either may be NULL.  rows = n_in_model*(D+1 or 1). */
gpis_status gpis_get_factor(gpis_handle h, double* L, int64_t ldl, double* alpha, int32_t* rows,
                            void* stream);

/* create_gpis.m:1-101 on a given training set (create_dense_gpis.m): replaces the model
 * with a dense fit on m points, appended in the given order.  SoA pts / normals. */
gpis_status gpis_fit(gpis_handle h, const double* pts, const double* tsdf, const double* normals,
                     const double* noise, int32_t m, void* stream);

/* gp_mean.m + diag(gp_cov.m) (+ the mean's gradient, predict_2d_grid.m:12-16) at nq
 * arbitrary points, SoA query[i + d*nq].  var is the latent variance.  normals
 * (SoA D x nq) may be NULL. */
gpis_status gpis_predict(gpis_handle h, const double* query, int64_t nq, double* mean, double* var,
                         double* normals, void* stream);

/* gp_mean.m:1-12 and gp_cov.m:1-21 IN FULL (the matrix, not its diagonal) at nq points, SoA query.
 * Values are ordered like the reference's joint vector, dimension-major: value c = b*nq + q is
 * f(x_q) for b = 0 and d_b f(x_q) for b = 1..D (se_cov_derivative.m block layout); nv = nq values
 * when use_derivative == 0, nq*(D+1) otherwise.  mean[nv], cov[row*ldc + col] (ldc >= nv),
 * symmetrised as gp_cov.m:19-20 does, latent (beta = 0 on K**, gp_cov.m:6).  Either output may be
 * NULL.  nq <= 65535. */
gpis_status gpis_posterior_cov(gpis_handle h, const double* query, int64_t nq, int32_t use_derivative,
                               double* mean, double* cov, int64_t ldc, void* stream);

/* mc_sampling/sample_shapes.m:10-31 (mvnrnd + mvnpdf on MEAN, COV + jitter I) and
 * Gpis3D.sample_sdfs (src/grasp_selection/gpis.py:98-112) at nq points:
 *   samples[s*nv + c] = MEAN[c] + sum_k z[s*nv + k] T[k][c],   T'T = COV + jitter I (T upper),
 *   logpdf[s] = -|z_s|^2/2 - sum_k log T[k][k] - nv/2 log(2 pi)   (the reference's pdfs = exp(logpdf)).
 * z holds the standard-normal draws, one sample per row ([num_samples][nv]); the caller supplies them
 * so that runs are reproducible (MATLAB's randn stream is not).  The reference uses jitter = 1e-12
 * (sample_shapes.m:20).  logpdf may be NULL; chol (NULL or [nv][ldt], ldt >= nv) receives T.
 * num_samples may be 0 (factor only).  GPIS_ERR_NOT_PD when a pivot is not positive: raise jitter
 * (mvnrnd falls back to an eigendecomposition there; this engine reports it instead). */
gpis_status gpis_sample_shapes(gpis_handle h, const double* query, int64_t nq, int32_t use_derivative,
                               int32_t num_samples, const double* z, double jitter, double* samples,
                               double* logpdf, double* chol, int64_t ldt, void* stream);
/* Device time (CUDA events on the call's stream), ms, of the last gpis_posterior_cov /
 * gpis_sample_shapes: mean + covariance build, Cholesky, draws.  Any pointer may be NULL. */
gpis_status gpis_get_sampling_timing(gpis_handle h, double* cov_ms, double* chol_ms, double* draw_ms);

/* A batch of INDEPENDENT small objects on the same lattice (BASELINE config 5: 512 x 32^3 object fits;
 * the Dex-Net-style caller loops Gpis3D / select_active_set over objects): the whole greedy selection
 * of an object (select_active_set_straddle.m:62-130) runs inside one thread block and one launch
 * advances every object -- no handle, no per-step launches.  cfg supplies dim (2 or 3), ell, sf2,
 * mean, noise, level, eps, beta, variance_mode, device; it must ask for function values only
 * (use_gradients 0), fp64, GPIS_BETA_FIXED, world_size <= 1 (shard the objects across ranks).
 * dims <= 32 per axis, K <= 256.  tsdf[k*N + j]: object k, lattice index j (x fastest).
 * Outputs (host or device pointers): ids[k*K + i] picks in selection order (-1 beyond
 * n_selected[k]), n_in_model[k] (NULL allowed), and -- each may be NULL -- the posterior mean /
 * variance of every lattice point and the level-set flags (bit 0 active, bit 1 high, bit 2 low),
 * [n_objects][N].  device_ms (NULL allowed): CUDA-event time of the launch.  cfg->batch_kernel selects
 * the form of the kernel's inner loops (GPIS_BATCH_PLAIN: untuned, GPIS_BATCH_DMMA: the update on the fp64
 * tensor pipe -- the fastest, and what AUTO selects --, GPIS_BATCH_DFMA: the tuned register-tile form; same picks in all).  Returns
 * GPIS_ERR_NOT_PD if any object's factor lost positive definiteness (its counts tell where). */
gpis_status gpis_batch_select(const gpis_config* cfg, const int32_t* dims, const double* origin,
                              const double* spacing, int32_t n_objects, const double* tsdf,
                              int64_t first_index, int32_t K, int64_t* ids, int32_t* n_selected,
                              int32_t* n_in_model, double* mean, double* var, uint8_t* flags,
                              double* device_ms, void* stream);

/* float entry points.  The reference's ABI is fp32 throughout (SelectChol(maxSize, float* inputPoints, float* targetPoints,
 * .., float* activeInputs, float* activeTargets), include/gpu_active_set_selector.hpp:41-46; the float* members of
 * ActiveSetBuffers / MaxSubsetBuffers, include/active_set_selection_types.h:16-41).  These take and return float arrays
 * with the same layouts as their double counterparts above (host or device pointers); the engine converts on the device
 * and computes in fp64.
 *   gpis_get_active_f32 fills what SelectChol's activeInputs / activeTargets were meant to receive
 *   (gpu_active_set_selector.cpp:584-586 never writes them): active_inputs[a + d*max_size], active_targets[a] for the
 *   n_in_model model points in selection order (max_size >= n_in_model; the rest is zeroed); the targets are read back
 *   from the replicated factor (y = m + L g), so every rank of a sharded run returns all of them. */
gpis_status gpis_set_candidates_f32(gpis_handle h, const float* pts, const float* tsdf, const float* normals,
                                    const float* noise, const int64_t* ids, void* stream);
gpis_status gpis_set_grid_candidates_f32(gpis_handle h, const int32_t* dims, const double* origin, const double* spacing,
                                         int32_t z0, int32_t nz_total, const float* tsdf, const float* normals,
                                         const float* noise, void* stream);
gpis_status gpis_get_active_f32(gpis_handle h, int32_t max_size, float* active_inputs, float* active_targets,
                                int32_t* n_selected, int32_t* n_in_model, void* stream);
gpis_status gpis_predict_f32(gpis_handle h, const float* query, int64_t nq, float* mean, float* var, float* normals,
                             void* stream);
gpis_status gpis_get_candidate_posterior_f32(gpis_handle h, float* mean, float* var, void* stream);

/* matlab/core/auto_tsdf.m:133-262, the step BEFORE the selection path: from the rasterised shape to shapeParams.
 *   inside[i + j*rows] != 0: pixel (row i, column j) is inside the shape (the reference's image value 102; 255 outside).
 *   :133-157  3x3 imdilate / imerode masks (the 5x5 window for the double dilation) -> TSDF in {-1, -0.5, 0, 0.5, 1};
 *   :159-218  measurement noise per pixel: occlusionScale inside the three occlusion rectangles (1-based, low < i <= high)
 *             where tsdf < 0.6; interior points (tsdf < -0.5) measured with probability interiorRate; elsewhere noiseScale,
 *             scaled by a specular factor near the surface (edgeWin window) and by the image position (noiseGradMode);
 *   :221      imgradientxy 'CentralDifference' normals;   :229-236  validIndices = find(noise < occlusionScale).
 * The reference draws rand() in a data-dependent order that cannot be reproduced: the uniform draws are the input
 * u[3][rows*cols] (interior, specular scale, sparsity; NULL = 0.5 everywhere).  The first stage of auto_tsdf.m
 * (vision.ShapeInserter, :129-131) needs MATLAB's Computer Vision toolbox and is not part of this call.
 * Outputs (host or device pointers, any may be NULL), all column-major like the reference's tsdf(:):
 *   tsdf[n], normals[2][n] (Gx then Gy), noise[n], valid_idx[n] (first *n_valid entries: 0-based linear indices of the
 *   measured pixels, ascending).  device: CUDA ordinal or -1 = current.  Synchronises `stream` before returning. */
typedef struct gpis_tsdf_params {
  int32_t struct_size;          /* = sizeof(gpis_tsdf_params) */
  int32_t edge_win;             /* varParams.edgeWin */
  int32_t specular_noise;       /* varParams.specularNoise */
  int32_t noise_grad_mode;      /* varParams.noiseGradMode: 0 none, 1 TLBR, 2 TRBL, 3 BLTR, 4 BRTL */
  double noise_scale, occlusion_scale, interior_rate, sparsity_rate, horiz_scale, vert_scale;
  double occ_y_low[3], occ_y_high[3], occ_x_low[3], occ_x_high[3];   /* varParams.{y,x}_thresh{1,2,3}_{low,high} */
} gpis_tsdf_params;
gpis_status gpis_build_tsdf(const uint8_t* inside, int32_t rows, int32_t cols, const gpis_tsdf_params* prm, const double* u,
                            double* tsdf, double* normals, double* noise, int32_t* valid_idx, int32_t* n_valid,
                            int32_t device, void* stream);

/* Level-set state of the candidates after selection: the high / low masks (straddle.m:121-125; ClassificationBuffers)
 * and the straddle score (straddle.m:106-108).  For a candidate that is still unclassified (not active, not high, not
 * low) `ambiguity` is the score it was last scored with = the score of the current posterior, on both backends.  For
 * a classified candidate the backends keep different things and say so: the panel backend stops updating it (the
 * reference only predicts on unclassified points, straddle.m:79-81) and returns the score it had when it was
 * classified; the lattice backend updates every lattice point every step and returns the score of the CURRENT
 * posterior with the last step's beta. */
gpis_status gpis_get_levelset(gpis_handle h, double* ambiguity, uint8_t* high, uint8_t* low,
                              uint8_t* active, void* stream);
/* Running posterior of the candidates (exact for candidates still unclassified). */
gpis_status gpis_get_candidate_posterior(gpis_handle h, double* mean, double* var, void* stream);

/* Counters: kernels launched by this handle, live candidate tiles now, iterations run,
 * bytes of the factor panels.  Any pointer may be NULL. */
gpis_status gpis_get_stats(gpis_handle h, int64_t* kernel_launches, int64_t* live_tiles,
                           int64_t* iterations, int64_t* panel_bytes);
/* Device-side time of the last gpis_select / gpis_predict on their stream (CUDA events), ms;
 * with profiling on (gpis_set_profiling) also the summed device time of the update kernel's
 * launches in the last gpis_select and their count. */
gpis_status gpis_get_timing(gpis_handle h, double* select_ms, double* predict_ms, double* update_ms,
                            int64_t* update_launches);
/* Profiling only: summed device time of the last gpis_select's factor-append kernels, of the
 * main update kernel (k_update / k_grid_gemm) and of the grid backend's epilogue kernel. */
gpis_status gpis_get_phase_timing(gpis_handle h, double* append_ms, double* main_ms, double* epilogue_ms);
/* The persistent selection kernel's own lap timer (block 0, %globaltimer), summed over the steps of the
 * last gpis_select, ms: [0] winner / wait for the exchange records, [1] covariance of the winner with the model
 * rows, [2] W pass of the factor append, [3] grid barrier, [4] new row of W, [5] grid barrier, [6] unit prefix,
 * [7] update GEMM (3-D lattices: finished cubes are handed to the producer warps, whose epilogues run meanwhile;
 *     1-D / 2-D lattices: with the in-register epilogues of whole tiles), [8] epilogues of tiles shared between blocks
 *     + the wait for the producer warps' last epilogues,
 * [9] grid barrier, [10] winner reduction (+ record publication when sharded), [11] = number of steps (a count);
 * inside [7]: [12] coefficient precompute, [13] partial tiles to the workspace (store, fence, arrival), [14] whole tiles:
 *     staging for the producer warps (3-D) or in-register epilogues, [15] = number of (tile, 32-row chunk) units all
 *     blocks processed (a count; x 2*4096*32 = executed MMA flops: a tile is 4096 lattice points).
 * 16 doubles. */
gpis_status gpis_get_persist_timing(gpis_handle h, double* ms16);
/* Load balance of the persistent kernel's update phase in the last gpis_select: for every block, ms[2 b] = its time in
 * the phase (MMAs + epilogues), ms[2 b + 1] = reserved (0; experiment builds book a wait there); summed over the steps.  *n_blocks
 * receives the block count (ms may be NULL to query it); capacity = doubles available at ms. */
gpis_status gpis_get_persist_balance(gpis_handle h, double* ms, int32_t capacity, int32_t* n_blocks);
/* on != 0: gpis_select brackets every update-kernel launch with CUDA events on the launching
 * stream (plain launches instead of the captured graph). */
gpis_status gpis_set_profiling(gpis_handle h, int32_t on);
/* Per-iteration history of the last selection: factor rows read, live tiles processed and
 * candidates scored (unclassified at the start of the iteration).  Arrays of capacity
 * max_active + 1; n = iterations recorded. */
gpis_status gpis_get_history(gpis_handle h, int32_t* rows, int32_t* live_tiles, int32_t* scored, int32_t* n);

#ifdef __cplusplus
}
#endif
#endif /* GPIS_B200_H */
