// Shared helpers for libxpgnn_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "knobs.cuh"
#include "xpgnn_b200.h"

namespace xpgnn {

extern thread_local std::string g_last_error;
extern std::atomic<int64_t> g_launches;

inline int fail(const char* what, const char* file, int line, cudaError_t e = cudaSuccess) {
  char buf[512];
  if (e != cudaSuccess)
    snprintf(buf, sizeof buf, "%s:%d %s: %s", file, line, what, cudaGetErrorString(e));
  else
    snprintf(buf, sizeof buf, "%s:%d %s", file, line, what);
  g_last_error = buf;
  return 1;
}

#define XP_CHECK(expr)                                                     \
  do {                                                                     \
    cudaError_t _e = (expr);                                               \
    if (_e != cudaSuccess) return ::xpgnn::fail(#expr, __FILE__, __LINE__, _e); \
  } while (0)

#define XP_REQUIRE(cond, msg)                                              \
  do {                                                                     \
    if (!(cond)) return ::xpgnn::fail(msg, __FILE__, __LINE__);            \
  } while (0)

// every kernel launch of the library goes through this so gpu_launches can be reported
#define XP_LAUNCH(kernel, grid, block, smem, stream, ...)                  \
  do {                                                                     \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);            \
    ::xpgnn::g_launches.fetch_add(1, std::memory_order_relaxed);           \
    XP_CHECK(cudaGetLastError());                                          \
  } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

constexpr int kNumSMs = 148;  // B200

// stream-ordered scratch that frees itself
struct Scratch {
  void* p = nullptr;
  cudaStream_t s;
  explicit Scratch(cudaStream_t st) : s(st) {}
  cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&p, bytes ? bytes : 1, s); }
  ~Scratch() {
    if (p) cudaFreeAsync(p, s);
  }
  template <class T>
  T* as() { return reinterpret_cast<T*>(p); }
};

}  // namespace xpgnn
