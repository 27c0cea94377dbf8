// Batch of independent small objects (BASELINE config 5: 512 x 32^3 object GPIS fits): the WHOLE greedy
// level-set selection of one object runs inside one thread block, and one launch advances every object
// of the batch -- no per-step launches, no inter-block synchronisation, nothing on the host.
//
// Same algorithm and the same arithmetic rules as the lattice backend (select_active_set_straddle.m:62-130
// through the incremental factor): W = L^-1 explicit, new entry of V = L^-1 K(rows, candidates) from the
// separable axis tables,
//     v[jz][jy][jx] = sum_rho (w[rho] Tz[rho][jz]) * Ty[rho][jy] * Tx[rho][jx],
// straddle score without FMA contraction, lowest index on ties, last pick selected but not in the model.
// What differs is the mapping: the axis tables of the object (3 x 256 x 32 doubles = 192 KB) live in the
// block's shared memory, every thread owns a 2 x 2 column of the (jy, jx) plane and walks the z-slices
// eight at a time with a 2 x 2 x 8 register tile (plain DFMA: on B200 the fp64 FMA pipe and the fp64
// tensor pipe have the same measured peak, profiles/r01_fp64_peak.json), and the posterior state of the
// object (mu, var, flags: 17 bytes per lattice point) streams through L2 once per step.
// Limits: function values only, scalar noise, fixed beta, lattice <= 32 per axis, <= 256 points.
#pragma once
#include "gpis_launch.cuh"
#include "gpis_common.cuh"

namespace gpis {

constexpr int BT = 256;          // threads per object
constexpr int BRMAX = 256;       // model rows (= points) per object
constexpr int BAX = 32;          // lattice points per axis
constexpr int BZB = 8;           // z-slices per register tile
constexpr size_t BATCH_SMEM = sizeof(double) * (3 * BRMAX * BAX + 3 * BRMAX + 3 * BRMAX + BRMAX);

struct BatchParams {
  int nx, ny, nz;
  long long N;                   // nx * ny * nz
  int n_objects, K;
  long long first;               // first pick (straddle.m:36-37), the same for every object
  double org[3], h[3];           // x = org + h * index
  double inv_ell2, sf2, noise, level, eps, beta;
  int variance_mode;
  double mean_w[3], mean_c;
  const double* tsdf;            // [n_objects][N]
  double* mu;                    // [n_objects][N]  running posterior mean (output)
  double* var;                   // [n_objects][N]  running posterior variance (output)
  unsigned char* flags;          // [n_objects][N]  F_ACTIVE | F_HIGH | F_LOW (output)
  long long* ids;                // [n_objects][K]  picks in selection order (output)
  int* n_sel;                    // [n_objects]
  int* n_model;                  // [n_objects]
  int* err;                      // [n_objects]     0 or GPIS_ERR_NOT_PD (-5)
  double* W;                     // [gridDim.x][BRMAX][BRMAX]   row-major lower L^-1, per-block workspace
  double* Wt;                    // [gridDim.x][BRMAX][BRMAX]   its transpose
};

__device__ __forceinline__ bool batch_better(double s, long long id, double bs, long long bid) {
  return (s > bs) || (s == bs && id < bid);
}

#ifndef GPIS_EMULATE
// D[8x8] += A[8x4] B[4x8] on the fp64 tensor pipe (SASS DMMA.8x8x4): lane l holds A[l/4][l%4],
// B[l%4][l/4] and D[l/4][2*(l%4) + {0,1}].
__device__ __forceinline__ void b_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
#endif

// FAST selects the tuned forms of the three inner loops (vector shared-memory loads and w folded into the
// Ty factor in the update; the state of eight lattice points loaded before any of them is processed; four
// independent partial sums in the two triangular passes of the append).  Both forms are kept: the plain
// one is the reference the tuned one is tested against (same picks; sums grouped differently).
#define GPIS_BATCH_KERNEL_NAME k_batch_select
#define GPIS_BATCH_BT 256
#include "gpis_batch_kernel.inc"
#undef GPIS_BATCH_BT
#undef GPIS_BATCH_KERNEL_NAME

// The update on the fp64 tensor pipe (opt-in, GPIS_BATCH_KERNEL=dmma; verified on the CPU emulation only so
// far): warp w takes the z-slices w, w+8, .., each slice a [32 x r] x [r x 32] product as 4 x 4 DMMA.8x8x4
// tiles with the A fragment scaled by w[rho] Tz[rho][jz] (29 instructions per 128 FMAs per lane instead of 47
// per 32); the Tx / Ty rows are stored with their column index XOR-swizzled by the row so that the fragment
// loads are conflict-free without padding (there is no shared memory left to pad with).
#define GPIS_BATCH_KERNEL_NAME k_batch_select_dmma
#define GPIS_BATCH_DMMA 1
#define GPIS_BATCH_BT 512      // 16 warps: more slices in flight, the tensor pipe overlaps other warps' epilogues
#include "gpis_batch_kernel.inc"
#undef GPIS_BATCH_BT
#undef GPIS_BATCH_DMMA
#undef GPIS_BATCH_KERNEL_NAME
constexpr int BT_DMMA = 512;

inline void launch_batch_select(const BatchParams& P, int n_blocks, int variant, cudaStream_t s) {
  auto k0 = k_batch_select<false>;
  auto k1 = k_batch_select<true>;
  auto k2 = k_batch_select_dmma<true>;
  if (variant == 2) GPIS_LAUNCH_SMEM(k2, dim3((unsigned)n_blocks), dim3(BT_DMMA), BATCH_SMEM, s, P);
  else if (variant == 1) GPIS_LAUNCH_SMEM(k1, dim3((unsigned)n_blocks), dim3(BT), BATCH_SMEM, s, P);
  else GPIS_LAUNCH_SMEM(k0, dim3((unsigned)n_blocks), dim3(BT), BATCH_SMEM, s, P);
}

}  // namespace gpis
