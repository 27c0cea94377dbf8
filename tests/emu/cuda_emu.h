// TEST INFRASTRUCTURE ONLY -- a CPU emulation of the CUDA execution model, just enough of it to run the
// kernels and launch sequences of gpis_b200/csrc/gpis_sample.cuh under g++ (-DGPIS_EMULATE) so that their
// indexing and pointer arithmetic are checked where there is no GPU.  One OS thread per CUDA thread,
// thread blocks one after another, __syncthreads() = a std::barrier, __shared__ = a function-static
// array, mma.sync.m8n8k4.f64 = a warp-wide exchange through a staging box.  Nothing under gpis_b200/
// uses this; the shipped library is compiled by nvcc without GPIS_EMULATE and has no CPU path.
#pragma once
#include <barrier>
#include <cmath>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __shared__ static
#define __forceinline__ inline
#define __launch_bounds__(...)

struct uint3 { unsigned x, y, z; };
struct double2 { double x, y; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
typedef void* cudaStream_t;
typedef int cudaError_t;
constexpr int cudaSuccess = 0;

namespace emu {
struct WarpBox {
  std::barrier<> bar{32};
  double a[32], b[32];
  unsigned long long slot[32];
};
inline thread_local uint3 t_idx, b_idx;
inline thread_local WarpBox* my_warp;
inline thread_local int my_lane;
inline dim3 g_dim, b_dim;
inline std::barrier<>* cta_bar;
inline unsigned char* dyn_smem;          // the block's dynamic shared memory (GPIS_DYN_SMEM)

template <class F>
void launch(dim3 grid, dim3 block, size_t smem_bytes, F body) {
  const unsigned nth = block.x * block.y * block.z;
  g_dim = grid;
  b_dim = block;
  std::vector<unsigned char> smem(smem_bytes + 16);
  dyn_smem = smem.data();
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        std::barrier<> bar(nth);
        cta_bar = &bar;
        std::vector<std::unique_ptr<WarpBox>> warps;
        for (unsigned w = 0; w < (nth + 31) / 32; ++w) warps.emplace_back(new WarpBox);
        std::vector<std::thread> th;
        th.reserve(nth);
        for (unsigned t = 0; t < nth; ++t)
          th.emplace_back([&, t]() {
            t_idx = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
            b_idx = uint3{bx, by, bz};
            my_warp = warps[t / 32].get();
            my_lane = (int)(t % 32);
            body();
            bar.arrive_and_drop();          // a thread that returned no longer takes part in barriers
          });
        for (auto& x : th) x.join();
      }
}
}  // namespace emu

#define threadIdx emu::t_idx
#define blockIdx emu::b_idx
#define blockDim emu::b_dim
#define gridDim emu::g_dim

using std::max;
using std::min;
inline void __syncthreads() { emu::cta_bar->arrive_and_wait(); }
inline void __syncwarp() { emu::my_warp->bar.arrive_and_wait(); }
template <class T>
inline T __ldcg(const T* p) { return *p; }
inline void __threadfence_block() {}
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
// warp collectives over all 32 lanes: every lane publishes its value, then reads what it needs
template <class T>
inline T __shfl_sync(unsigned, T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
  emu::WarpBox* w = emu::my_warp;
  unsigned long long bits = 0;
  memcpy(&bits, &v, sizeof(T));
  w->slot[emu::my_lane] = bits;
  w->bar.arrive_and_wait();
  const unsigned long long other = w->slot[src_lane & 31];
  w->bar.arrive_and_wait();
  T out;
  memcpy(&out, &other, sizeof(T));
  return out;
}
inline unsigned __ballot_sync(unsigned, bool pred) {
  emu::WarpBox* w = emu::my_warp;
  w->slot[emu::my_lane] = pred ? 1ull : 0ull;
  w->bar.arrive_and_wait();
  unsigned m = 0;
  for (int l = 0; l < 32; ++l) m |= (unsigned)(w->slot[l] & 1ull) << l;
  w->bar.arrive_and_wait();
  return m;
}
// warp shuffle over all 32 lanes: every lane publishes its value, reads its partner's
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
  emu::WarpBox* w = emu::my_warp;
  const int l = emu::my_lane;
  unsigned long long bits = 0;
  memcpy(&bits, &v, sizeof(T));
  w->slot[l] = bits;
  w->bar.arrive_and_wait();
  const unsigned long long other = w->slot[l ^ lane_mask];
  w->bar.arrive_and_wait();
  T out;
  memcpy(&out, &other, sizeof(T));
  return out;
}
inline int atomicCAS(int* p, int cmp, int val) { return __sync_val_compare_and_swap(p, cmp, val); }
inline int atomicAdd(int* p, int v) { return __sync_fetch_and_add(p, v); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __sync_fetch_and_add(p, v); }
inline int atomicOr(int* p, int v) { return __sync_fetch_and_or(p, v); }
inline void __threadfence() { __sync_synchronize(); }
inline void __threadfence_system() { __sync_synchronize(); }
inline long long clock64() { return 0; }
inline void __nanosleep(unsigned) {}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }

namespace gpis {
// cp.async stand-ins: the copy happens at once, commit / wait are no-ops (a superset of the allowed
// behaviours: a missing wait cannot be detected here)
inline void cp_async16(void* smem, const void* gmem) { memcpy(smem, gmem, 16); }
inline void cp_async_commit() {}
inline void cp_async_wait_all() {}
// mma.sync.aligned.m8n8k4.row.col.f64: lane l supplies A[l/4][l%4] and B[l%4][l/4] and owns
// D[l/4][2*(l%4) + {0,1}] (PTX ISA, "Matrix fragments for mma.m8n8k4 with .f64").
inline void s_dmma(double& c0, double& c1, double a, double b) {
  emu::WarpBox* w = emu::my_warp;
  const int l = emu::my_lane;
  w->a[l] = a;
  w->b[l] = b;
  w->bar.arrive_and_wait();
  const int m = l >> 2, n = 2 * (l & 3);
  for (int k = 0; k < 4; ++k) {
    c0 = std::fma(w->a[m * 4 + k], w->b[n * 4 + k], c0);
    c1 = std::fma(w->a[m * 4 + k], w->b[(n + 1) * 4 + k], c1);
  }
  w->bar.arrive_and_wait();
}
inline void p_dmma(double& c0, double& c1, double a, double b) { s_dmma(c0, c1, a, b); }
}  // namespace gpis
