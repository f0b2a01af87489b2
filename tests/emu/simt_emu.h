// Host SIMT emulator for CPU tests of hand-written CUDA kernels (test infrastructure, never on the product path).
// Every CUDA thread of ONE block is an OS thread; __syncthreads / __syncwarp are barriers; warp shuffles go through a
// per-warp exchange buffer between two warp barriers, so a kernel's barrier and shuffle structure is exercised as
// written (a missing __syncthreads shows up as a data race / wrong result, a divergent shuffle as a deadlock).
// Blocks of a grid run one after the other.  Block sizes must be multiples of 32.
#pragma once
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

struct EmuDim3 { int x, y, z; };
static thread_local EmuDim3 threadIdx{0, 0, 0};
static thread_local EmuDim3 blockIdx{0, 0, 0};
static EmuDim3 blockDim{1, 1, 1}, gridDim{1, 1, 1};
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __align__(x) __attribute__((aligned(x)))
#define __shared__

static std::unique_ptr<std::barrier<>> emu_block_bar;
static std::vector<std::unique_ptr<std::barrier<>>> emu_warp_bar;
static uint32_t emu_slots[32][32];

inline void __syncthreads() { emu_block_bar->arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu_warp_bar[threadIdx.x >> 5]->arrive_and_wait(); }
template <class T>
inline T emu_exchange(T v, int src_lane) {
  static_assert(sizeof(T) == 4, "32-bit shuffles only");
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  std::memcpy(&emu_slots[w][l], &v, 4);
  emu_warp_bar[w]->arrive_and_wait();
  T r;
  std::memcpy(&r, &emu_slots[w][src_lane & 31], 4);
  emu_warp_bar[w]->arrive_and_wait();
  return r;
}
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int off) { return emu_exchange(v, (threadIdx.x & 31) ^ off); }
template <class T>
inline T __shfl_sync(unsigned, T v, int src) { return emu_exchange(v, src); }
inline unsigned __ballot_sync(unsigned, bool pred) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  emu_slots[w][l] = pred ? 1u : 0u;
  emu_warp_bar[w]->arrive_and_wait();
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= emu_slots[w][i] << i;
  emu_warp_bar[w]->arrive_and_wait();
  return r;
}
inline int __ffs(unsigned x) { return x ? __builtin_ctz(x) + 1 : 0; }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline float __frcp_rn(float x) { return 1.0f / x; }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
inline int __float_as_int(float f) { int x; std::memcpy(&x, &f, 4); return x; }
template <class T>
inline T __ldg(const T* p) { return *p; }
inline float __int_as_float(int x) { float f; std::memcpy(&f, &x, 4); return f; }

// Runs kernel(args...) for every block of the grid with `threads` OS threads per block.
inline void emu_launch(int grid, int threads, const std::function<void()>& body) {
  blockDim = {threads, 1, 1};
  gridDim = {grid, 1, 1};
  for (int b = 0; b < grid; ++b) {
    emu_block_bar = std::make_unique<std::barrier<>>(threads);
    emu_warp_bar.clear();
    for (int w = 0; w < threads / 32; ++w) emu_warp_bar.push_back(std::make_unique<std::barrier<>>(32));
    std::vector<std::thread> pool;
    pool.reserve(threads);
    for (int t = 0; t < threads; ++t)
      pool.emplace_back([&, t, b] {
        threadIdx = {t, 0, 0};
        blockIdx = {b, 0, 0};
        body();
      });
    for (auto& th : pool) th.join();
  }
}
