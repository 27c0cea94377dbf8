// Shared device/host definitions of the GPIS engine (sm_100a).
#pragma once
#ifndef GPIS_EMULATE   // tests/emu supplies the execution model on the CPU (test infrastructure only)
#include <cuda_runtime.h>
#endif
#include <stdint.h>

namespace gpis {

constexpr int TW = 64;            // candidates per tile (one 512-byte fp64 row segment)
constexpr int UPD_THREADS = 256;  // update kernel: 8 warps, each owns every 8th factor row
constexpr int UPD_WARPS = UPD_THREADS / 32;
constexpr int APP_THREADS = 512;  // append kernel (single CTA)
constexpr int PAY_HDR = 16;       // exchange record header, in 8-byte slots
constexpr int MAXD = 3;
constexpr int MAXNB = MAXD + 1;

constexpr unsigned char F_ACTIVE = 1, F_HIGH = 2, F_LOW = 4;

// exchange record header slots
enum { PAY_SCORE = 0, PAY_ID = 1, PAY_LOCAL = 2, PAY_TSDF = 3, PAY_NOISE = 4, PAY_X = 5, PAY_NRM = 8 };

struct DevState {
  int r;        // rows in the factor
  int r0;       // rows before the block appended last (what the update kernel reads)
  int m;        // points in the model
  int n_sel;    // points selected (m, or m + 1 while the last pick is pending)
  int done;     // selection stopped (nothing unclassified / capacity)
  int iter;     // update iterations executed
  int err;      // sticky device-side error (GPIS_ERR_*)
  int n_live[2];
  unsigned int blocks_done;
  int n_scored;  // candidates scored (unclassified at the start) in the running iteration
  unsigned int grid_bar;  // arrivals at the in-kernel grid barrier of the fused solve kernel (monotonic per selection)
  unsigned int epoch;  // selection counter: tags the peer-exchange flags of this selection
  long long pending_id;
  double nx[MAXD];          // point appended last
  double Lnn[MAXNB * MAXNB];  // its diagonal block of L
  double gnew[MAXNB];       // its entries of g = L^-1 (y - m)
  double nnoise;
};

struct BlockBest {
  double score;
  long long id;
  long long local;
};

struct EngineParams {
  // candidates (this rank's shard)
  const double* cx[MAXD];
  const double* tsdf;
  const double* nrm[MAXD];
  const double* noise;     // per-point or null
  const long long* ids;    // or null
  long long id_base;
  long long N;
  long long N_total;
  int n_tiles;
  // panels of V = L^-1 K(rows, candidates): [tile][row][TW]
  double* V;
  long long tile_stride;   // R_cap * TW
  double* mu;
  double* var;
  double* amb;
  unsigned char* flags;
  int* live[2];
  // factor
  double* L;               // [R_cap][R_cap] lower, row-major
  double* g;               // L^-1 (y - m)
  double* AX[MAXD];        // active points, SoA [M_cap]
  long long* sel_ids;      // [M_cap + 1]
  int R_cap, M_cap;
  DevState* st;
  int* hist;               // [3][M_cap + 1]: rows read, live tiles, candidates scored, per iteration
  BlockBest* partials;
  double* pay_send;
  double* pay_recv;
  long long pay_stride;    // slots per record
  int rank, world;
  // peer exchange over NVLink (peer_mode): every rank stores its record straight into all peers'
  // exchange regions [2 parities][world slots][pay_stride] and raises a flag there; no collective call
  int peer_mode;
  double* xchg;
  unsigned long long* xflag;       // [2][world]
  double** peer_xchg;              // device array of world pointers (own region included)
  unsigned long long** peer_xflag;
  // selection fence: xdone[p] = last selection (epoch) rank p has finished reading records of; a rank
  // publishes nothing of selection e + 1 into a peer's region before every peer has finished e
  // persistent selection kernel, sharded: per-step records as 8-byte {data half, tag} slots (no fence, no separate
  // flag: a slot is valid when its tag is this step's) [2 parities][world][16 slots], at byte offset xll_off of
  // every rank's exchange region
  long long xll_off;
  unsigned long long* xdone;       // [world], written by the peers
  unsigned long long** peer_xdone;
  int ell_chunk;           // factor rows staged in shared memory per pass
  // grid backend (glen[0] > 0): lattice geometry, separable axis tables, explicit W = L^-1
  int glen[MAXD];          // candidates per axis on THIS rank (z: the local slab)
  int gz0;                 // first global z index of the local slab
  int gtlen[MAXD];         // axis-table lengths (axis 2: the GLOBAL z count)
  double gorg[MAXD], gh[MAXD];   // x = gorg + gh * global index
  double* T[MAXD];         // [rows][ldt]: axis tables (axis 2 spans the GLOBAL z range)
  int ldt[MAXD];
  float* Tf[MAXD];         // fp32 copies of the axis tables (TF32 mode), same leading dimensions
  float* bimg;             // TF32 mode: hi/lo axis-0 table in the MMA's shared-memory image [x-tiles][bimg_cmax][2][128*32]
  int bimg_cmax;
  int gemm_zgroup;         // z-slices per work tile of the update GEMM (1 DMMA, 2 tcgen05)
  int gemm_kc;             // model rows per split-K chunk of the update GEMM (16 DMMA, 32 tcgen05)
  double* W;               // [rows][ldW] lower triangular L^-1
  double* Wt;              // its transpose (shared with the predict kernel)
  long long ldW;
  double* kvec;            // [MAXNB][R_cap] covariance of the new rows with the model rows
  double* ypart;           // one-pass solve: per-CTA partial W' l vectors [solve CTAs][4096]
  double* gws;             // split-K workspace [NB][tiles][gws_segmax][128*128]
  int gws_segmax;
  int gemm_ctas;           // persistent grid of the update GEMM
  unsigned int* tile_cnt;  // persistent selection kernel: [2][tiles] split-K arrival counters (step parity)
  // persistent selection kernel: tile edge of the per-candidate state layout of THIS selection (128 for the
  // per-step kernels, 64 for the persistent kernel) and, per 64 x 64 tile, the list of model rows whose
  // kernel factor over the tile's box is not below 2^-70 (the others cannot change a double-precision sum)
  int tile_e;
  unsigned short* rlist;   // [tiles][rlist_stride] model-row indices, append order (ascending)
  double* TT[MAXD];        // transposed axis tables [axis index][ldtt rows] (the winner's covariance with the model rows)
  long long ldtt;
  double* rtz;             // [tiles][rlist_stride] the rows' z-axis factor at the tile's slice
  int* rcnt;               // [tiles]
  int rlist_stride;
  int* upref;              // per-step path with row lists: exclusive prefix of the 128 x 128 tiles' chunk counts [tiles + 1], else null
  unsigned long long* phase_clk;   // persistent selection kernel: CTA 0's phase times (globaltimer ns), [8]
  long long spin_cycles;   // SM clocks an in-kernel wait (peer record, grid barrier) may take before it gives up
  // model
  double inv_ell2, sf2, level, eps, beta, beta_delta, noise_scalar;
  double mean_w[MAXD], mean_c;
  int beta_mode, variance_mode;
};

__host__ __device__ inline long long cand_id(const EngineParams& P, long long j) {
  return P.ids ? P.ids[j] : P.id_base + j;
}

}  // namespace gpis
