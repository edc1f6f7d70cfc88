// Shared helpers for the ciaosr_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/ciaosr_b200.h"

namespace ciaosr {

// ---- error plumbing ------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launch_count;

#define CIAOSR_CUDA_OK(expr)                                                     \
  do {                                                                           \
    cudaError_t _e = (expr);                                                     \
    if (_e != cudaSuccess) {                                                     \
      ::ciaosr::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,          \
                          cudaGetErrorString(_e));                               \
      return CIAOSR_E_CUDA;                                                      \
    }                                                                            \
  } while (0)

#define CIAOSR_REQUIRE(cond, code, ...)                                          \
  do {                                                                           \
    if (!(cond)) {                                                               \
      ::ciaosr::set_error(__VA_ARGS__);                                          \
      return (code);                                                             \
    }                                                                            \
  } while (0)

// Every kernel launch goes through this so `ciaosr_launch_count()` is honest.
#define CIAOSR_LAUNCH(kernel, grid, block, smem, stream, ...)                    \
  do {                                                                           \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                  \
    ::ciaosr::g_launch_count.fetch_add(1, std::memory_order_relaxed);            \
    cudaError_t _e = cudaGetLastError();                                         \
    if (_e != cudaSuccess) {                                                     \
      ::ciaosr::set_error("%s:%d: launch of %s failed: %s", __FILE__, __LINE__,  \
                          #kernel, cudaGetErrorString(_e));                      \
      return CIAOSR_E_CUDA;                                                      \
    }                                                                            \
  } while (0)

// ---- opt-in to > 48 KB of dynamic shared memory --------------------------------------------------
// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of (function, DEVICE): a process that uses
// cuda:1 after cuda:0 must opt in again there.  One DynSmemOptIn per kernel remembers, per device, the
// largest size already granted; concurrent callers may both set the attribute (harmless, same value).
constexpr int CIAOSR_MAX_DEVICES = 64;
struct DynSmemOptIn {
  std::atomic<int> granted[CIAOSR_MAX_DEVICES] = {};
  template <class Kernel>
  int ensure(Kernel kernel, int bytes) {
    int dev = 0;
    CIAOSR_CUDA_OK(cudaGetDevice(&dev));
    const bool tracked = dev >= 0 && dev < CIAOSR_MAX_DEVICES;
    if (tracked && granted[dev].load(std::memory_order_acquire) >= bytes) return CIAOSR_OK;
    CIAOSR_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    if (tracked) {
      int cur = granted[dev].load(std::memory_order_relaxed);
      while (cur < bytes && !granted[dev].compare_exchange_weak(cur, bytes, std::memory_order_release)) {}
    }
    return CIAOSR_OK;
  }
};

// ---- stage timing ---------------------------------------------------------
struct StageScope {          // RAII: records an event pair around a stage when profiling is on
  int stage; cudaStream_t st; cudaEvent_t a, b; bool on;
  StageScope(int stage, cudaStream_t st);
  ~StageScope();
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// Bump allocator over a caller-provided device buffer.
struct Arena {
  char* base; size_t cap; size_t off; bool ok;
  Arena(void* p, size_t n) : base((char*)p), cap(n), off(0), ok(true) {}
  template <class T> T* take(size_t count) {
    off = align_up(off, 256);
    size_t bytes = count * sizeof(T);
    if (base != nullptr && off + bytes > cap) ok = false;
    T* r = base ? (T*)(base + off) : nullptr;
    off += bytes;
    return r;
  }
  size_t used() const { return align_up(off, 256); }
};

// ---- coordinate arithmetic ------------------------------------------------
// The gather indices are the only integer results on the path and must match
// ATen bit for bit, so every step is an explicitly rounded fp32 operation
// (no FMA contraction), in the order ATen's grid_sampler uses:
//   u = ((c + 1) * n - 1) / 2 ; index = nearbyint(u)   (align_corners=False)
__device__ __forceinline__ float unnormalize(float c, int n) {
  return __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(c, 1.0f), (float)n), 1.0f), 2.0f);
}
__device__ __forceinline__ int nearest_index(float c, int n) {
  return (int)rintf(unnormalize(c, n));   // round-half-even, like std::nearbyint
}

struct LaunchGeom { int B, H, W, C; };

}  // namespace ciaosr
