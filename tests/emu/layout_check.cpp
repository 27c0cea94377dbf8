// TEST INFRASTRUCTURE ONLY: the per-candidate state layouts of the grid backend (gpis_grid.cuh: grid_slot /
// grid_unslot / grid_slot_count for tile_e = 128, 64 and the 16 x 16 x 16 cubes) and the persistent kernel's
// fragment-slot -> lattice-index map (gpis_grid_persist.cuh: persist_pair_index) checked against each other on the
// host: every lattice point has its own slot below the slot count, unslot inverts slot, padding slots say so, and
// slot / 2 of a point is the fragment slot whose pair index is the point (the epilogue's addressing).
// g++ -std=c++20 -DGPIS_EMULATE -Itests/emu tests/emu/layout_check.cpp -o build/layout_check
#define GPIS_EMULATE 1
#include "cuda_rt_emu.h"
#include <cstdio>
#include <vector>
#include "../../gpis_b200/csrc/gpis_common.cuh"
#include "../../gpis_b200/csrc/gpis_grid.cuh"
#include "../../gpis_b200/csrc/gpis_grid_persist.cuh"

using namespace gpis;

static long long check(int nx, int ny, int nz, int tile_e) {
  EngineParams P;
  memset(&P, 0, sizeof(P));
  P.glen[0] = nx; P.glen[1] = ny; P.glen[2] = nz;
  P.tile_e = tile_e;
  const long long nslots = grid_slot_count(P);
  std::vector<char> seen((size_t)nslots, 0);
  long long bad = 0;
  const int te = tile_e;
  for (int jz = 0; jz < nz; ++jz)
    for (int jy = 0; jy < ny; ++jy)
      for (int jx = 0; jx < nx; ++jx) {
        const long long s = grid_slot(P, jx, jy, jz);
        if (s < 0 || s >= nslots || seen[(size_t)s]) { ++bad; continue; }
        seen[(size_t)s] = 1;
        int ux, uy, uz;
        if (!grid_unslot(P, s, ux, uy, uz) || ux != jx || uy != jy || uz != jz) ++bad;
        const long long jl = (long long)jx + (long long)nx * (jy + (long long)ny * jz);
        if (grid_slot_of_local(P, jl) != s) ++bad;
        if (te == 16 || te == 64) {      // the persistent kernel's addressing: fragment slot q of tile t holds this pair
          const long long tile = s / QTILE_ELEMS;
          const int q = (int)((s % QTILE_ELEMS) >> 1), e = (int)(s & 1);
          const int tiles_x = (nx + te - 1) / te, tiles_y = (ny + te - 1) / te;
          const int tx = (int)(tile % tiles_x), ty = (int)((tile / tiles_x) % tiles_y), tz = (int)(tile / ((long long)tiles_x * tiles_y));
          const long long pj = (te == 16 ? persist_pair_index<true>(q, tx, ty, tz, nx, ny) : persist_pair_index<false>(q, tx, ty, tz, nx, ny)) + e;
          if (pj != jl) ++bad;
        }
      }
  for (long long s = 0; s < nslots; ++s)
    if (!seen[(size_t)s]) {
      int ux, uy, uz;
      if (grid_unslot(P, s, ux, uy, uz)) ++bad;      // a padding slot must not claim a lattice point
    }
  return bad;
}

int main() {
  const int dims[][3] = {{16, 16, 16}, {24, 20, 24}, {48, 48, 45}, {33, 17, 19}, {128, 64, 32}, {5, 7, 1}, {70, 130, 3}};
  long long bad = 0;
  for (const auto& d : dims)
    for (int te : {128, 64, 16}) {
      if (te == 16 && d[2] == 1 && false) continue;
      const long long b = check(d[0], d[1], d[2], te);
      if (b) printf("layout mismatch: %d x %d x %d, tile_e %d: %lld\n", d[0], d[1], d[2], te, b);
      bad += b;
    }
  printf("%s\n", bad ? "FAILED" : "layouts ok");
  return bad ? 1 : 0;
}
