// TEST INFRASTRUCTURE ONLY -- a CPU emulation of the CUDA execution model, just enough of it to run the
// kernels and launch sequences of gpis_b200/csrc/*.cuh under g++ (-DGPIS_EMULATE) so that their indexing,
// barrier structure and pointer arithmetic are checked where there is no GPU.  One OS thread per CUDA
// thread, thread blocks one after another, __syncthreads() = a std::barrier, __shared__ = a
// function-static array, warp collectives and mma.sync.m8n8k4.f64 = exchanges through a per-warp staging
// box, mbarriers = a packed atomic word, asynchronous copies = immediate copies.  Nothing under gpis_b200/
// uses this; the shipped library is compiled by nvcc without GPIS_EMULATE and has no CPU path.
#pragma once
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __shared__ static
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n)

struct uint3 { unsigned x, y, z; };
struct double2 { double x, y; };
inline double2 make_double2(double a, double b) { return double2{a, b}; }
struct uchar2 { unsigned char x, y; };
inline uchar2 make_uchar2(unsigned char a, unsigned char b) { return uchar2{a, b}; }
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
typedef void* cudaStream_t;
typedef int cudaError_t;
constexpr int cudaSuccess = 0;

namespace emu {
// Staging box of a warp's collectives.  Double-buffered by a per-lane generation counter, so one barrier
// per collective suffices: a lane can run at most one collective ahead of the slowest lane (it has to pass
// the barrier of the collective in between), and that one writes the other buffer.
struct WarpBox {
  std::barrier<> bar{32};
  double a[2][32], b[2][32];
  unsigned long long slot[2][32];
};
// Barrier of a thread block whose participants may leave (a CUDA thread that returned no longer takes part
// in __syncthreads()); reset between blocks.
struct BlockBarrier {
  std::mutex m;
  std::condition_variable cv;
  int expected = 0, arrived = 0;
  unsigned gen = 0;
  void reset(int n) { std::lock_guard<std::mutex> lk(m); expected = n; arrived = 0; }
  void wait() {
    std::unique_lock<std::mutex> lk(m);
    const unsigned g = gen;
    if (++arrived == expected) { arrived = 0; ++gen; cv.notify_all(); }
    else cv.wait(lk, [&] { return gen != g; });
  }
  void drop() {
    std::lock_guard<std::mutex> lk(m);
    --expected;
    if (expected > 0 && arrived == expected) { arrived = 0; ++gen; cv.notify_all(); }
  }
};
inline thread_local uint3 t_idx, b_idx;
inline thread_local WarpBox* my_warp;
inline thread_local int my_lane;
inline thread_local unsigned my_gen;     // collectives executed by this thread (parity selects the staging buffer)
inline dim3 g_dim, b_dim;
inline BlockBarrier* cta_bar;
inline unsigned char* dyn_smem;          // the block's dynamic shared memory (GPIS_DYN_SMEM*)

// One OS thread per CUDA thread of a block, kept for the whole launch: the threads walk the grid's blocks
// together, one block after another (__shared__ arrays are function-statics, so blocks cannot overlap).
template <class F>
void launch(dim3 grid, dim3 block, size_t smem_bytes, F body) {
  const unsigned nth = block.x * block.y * block.z;
  g_dim = grid;
  b_dim = block;
  std::vector<unsigned char> smem(smem_bytes + 64);
  dyn_smem = smem.data() + (64 - reinterpret_cast<uintptr_t>(smem.data()) % 64) % 64;
  BlockBarrier cta;
  cta.reset((int)nth);
  cta_bar = &cta;
  std::barrier<> block_end((std::ptrdiff_t)nth);
  std::vector<std::unique_ptr<WarpBox>> warps;
  for (unsigned w = 0; w < (nth + 31) / 32; ++w) warps.emplace_back(new WarpBox);
  std::vector<std::thread> th;
  th.reserve(nth);
  for (unsigned t = 0; t < nth; ++t)
    th.emplace_back([&, t]() {
      t_idx = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
      my_warp = warps[t / 32].get();
      my_lane = (int)(t % 32);
      for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
          for (unsigned bx = 0; bx < grid.x; ++bx) {
            b_idx = uint3{bx, by, bz};
            my_gen = 0;
            body();
            cta.drop();                       // a thread that returned no longer takes part in __syncthreads()
            block_end.arrive_and_wait();      // the whole block has finished
            if (t == 0) cta.reset((int)nth);
            block_end.arrive_and_wait();      // ... and the barrier is armed for the next block
          }
    });
  for (auto& x : th) x.join();
}

// kernel<<<grid, block, smem>>>(args...): the arguments are evaluated ONCE, by the launching thread (an
// argument expression may have side effects), and every emulated thread gets a copy
template <class K, class... A>
void launch_kernel(dim3 grid, dim3 block, size_t smem_bytes, K kern, A... args) {
  launch(grid, block, smem_bytes, [&]() { kern(args...); });
}

// one exchange among the 32 lanes of the calling warp
inline unsigned long long exchange(unsigned long long bits, int from_lane, unsigned* ballot) {
  WarpBox* w = my_warp;
  const int buf = (int)(my_gen++ & 1u);
  w->slot[buf][my_lane] = bits;
  w->bar.arrive_and_wait();
  if (ballot) {
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) m |= (unsigned)(w->slot[buf][l] & 1ull) << l;
    *ballot = m;
    return 0;
  }
  return w->slot[buf][from_lane & 31];
}
}  // namespace emu

#define threadIdx emu::t_idx
#define blockIdx emu::b_idx
#define blockDim emu::b_dim
#define gridDim emu::g_dim

using std::max;
using std::min;
inline void __syncthreads() { emu::cta_bar->wait(); }
inline void __syncwarp() { emu::my_warp->bar.arrive_and_wait(); }
inline void __threadfence_block() {}
inline void __threadfence() { __sync_synchronize(); }
inline void __threadfence_system() { __sync_synchronize(); }
inline long long clock64() { return 0; }
inline void __nanosleep(unsigned) { std::this_thread::yield(); }
inline int atomicCAS(int* p, int cmp, int val) { return __sync_val_compare_and_swap(p, cmp, val); }
inline int atomicAdd(int* p, int v) { return __sync_fetch_and_add(p, v); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __sync_fetch_and_add(p, v); }
inline int atomicOr(int* p, int v) { return __sync_fetch_and_or(p, v); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
template <class T>
inline T __ldcg(const T* p) { return *p; }
template <class T>
inline T __ldg(const T* p) { return *p; }

// warp collectives over all 32 lanes: every lane publishes its value, then reads what it needs
template <class T>
inline T __shfl_sync(unsigned, T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
  unsigned long long bits = 0;
  memcpy(&bits, &v, sizeof(T));
  const unsigned long long other = emu::exchange(bits, src_lane, nullptr);
  T out;
  memcpy(&out, &other, sizeof(T));
  return out;
}
template <class T>
inline T __shfl_xor_sync(unsigned m, T v, int lane_mask) { return __shfl_sync(m, v, emu::my_lane ^ lane_mask); }
// lanes below `delta` keep their own value, like the hardware
template <class T>
inline T __shfl_up_sync(unsigned m, T v, int delta) { return __shfl_sync(m, v, emu::my_lane >= delta ? emu::my_lane - delta : emu::my_lane); }
inline unsigned __ballot_sync(unsigned, bool pred) {
  unsigned m = 0;
  emu::exchange(pred ? 1ull : 0ull, 0, &m);
  return m;
}

namespace gpis {
// mma.sync.aligned.m8n8k4.row.col.f64: lane l supplies A[l/4][l%4] and B[l%4][l/4] and owns
// D[l/4][2*(l%4) + {0,1}] (PTX ISA, "Matrix fragments for mma.m8n8k4 with .f64").
inline void s_dmma(double& c0, double& c1, double a, double b) {
  emu::WarpBox* w = emu::my_warp;
  const int l = emu::my_lane, buf = (int)(emu::my_gen++ & 1u);
  w->a[buf][l] = a;
  w->b[buf][l] = b;
  w->bar.arrive_and_wait();
  const int m = l >> 2, n = 2 * (l & 3);
  for (int k = 0; k < 4; ++k) {
    c0 = std::fma(w->a[buf][m * 4 + k], w->b[buf][n * 4 + k], c0);
    c1 = std::fma(w->a[buf][m * 4 + k], w->b[buf][(n + 1) * 4 + k], c1);
  }
}
inline void p_dmma(double& c0, double& c1, double a, double b) { s_dmma(c0, c1, a, b); }
inline void dmma8x8x4(double& c0, double& c1, double a, double b) { s_dmma(c0, c1, a, b); }
inline void b_dmma(double& c0, double& c1, double a, double b) { s_dmma(c0, c1, a, b); }

// cp.async stand-ins: the copy happens at once, commit / wait are no-ops (a superset of the allowed
// behaviours: a missing wait cannot be detected here)
inline void cp_async16(void* smem, const void* gmem) { memcpy(smem, gmem, 16); }
inline void cp_async_commit() {}
inline void cp_async_wait_all() {}
inline void g_cp_async16(void* smem, const void* gmem) { memcpy(smem, gmem, 16); }
inline void g_cp_commit() {}
template <int N>
inline void g_cp_wait() {}
inline double ld_nc_f64(const double* p) { return *p; }

// mbarrier (shared::cta, 64-bit): bits 0..19 pending arrivals, 20..39 the initial count, bit 40 the phase,
// bits 41..63 outstanding transaction bytes.  A phase completes when both counts reach zero.
namespace mbar_emu {
constexpr unsigned long long CNT = (1ull << 20) - 1;
inline void update(unsigned long long* bar, unsigned arrivals, long long tx_delta) {
  unsigned long long old = __atomic_load_n(bar, __ATOMIC_ACQUIRE), neu;
  do {
    unsigned long long pending = old & CNT, init = (old >> 20) & CNT, phase = (old >> 40) & 1ull;
    long long tx = (long long)(old >> 41);
    if (tx & (1ll << 22)) tx -= (1ll << 23);             // 23-bit two's complement: a copy may complete before expect_tx
    tx += tx_delta;
    pending -= arrivals;
    if (pending == 0 && tx == 0) { phase ^= 1ull; pending = init; }
    neu = pending | (init << 20) | (phase << 40) | (((unsigned long long)tx & ((1ull << 23) - 1)) << 41);
  } while (!__atomic_compare_exchange_n(bar, &old, neu, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE));
}
}  // namespace mbar_emu
inline void mbar_init(unsigned long long* bar, unsigned count) {
  __atomic_store_n(bar, (unsigned long long)count | ((unsigned long long)count << 20), __ATOMIC_RELEASE);
}
inline void mbar_arrive(unsigned long long* bar) { mbar_emu::update(bar, 1, 0); }
inline void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) { mbar_emu::update(bar, 1, (long long)bytes); }
inline void mbar_wait(unsigned long long* bar, unsigned parity) {
  while (((__atomic_load_n(bar, __ATOMIC_ACQUIRE) >> 40) & 1ull) == (parity & 1u)) std::this_thread::yield();
}
inline void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  memcpy(dst, src, bytes);
  mbar_emu::update(bar, 0, -(long long)bytes);
}
inline void mbar_init_fence() {}
}  // namespace gpis
