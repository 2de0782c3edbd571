// Shared helpers for the NSFF sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>

#include "nsff_b200.h"

namespace nsff {

extern thread_local cudaError_t g_last_cuda_error;
extern std::atomic<uint64_t> g_launch_count;

// Call right after a kernel launch: records launch-time errors and counts the launch.
inline int after_launch(int n = 1) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    g_last_cuda_error = e;
    return NSFF_ECUDA;
  }
  g_launch_count.fetch_add((uint64_t)n, std::memory_order_relaxed);
  return NSFF_OK;
}

inline int cuda_fail(cudaError_t e) {
  g_last_cuda_error = e;
  return NSFF_ECUDA;
}

#define NSFF_CUDA_TRY(expr)                          \
  do {                                               \
    cudaError_t _e = (expr);                         \
    if (_e != cudaSuccess) return nsff::cuda_fail(_e); \
  } while (0)

constexpr float kZFar = 0.95f;             // rendering.py:316
constexpr float kLastDeltaStatic = 100.f;  // rendering.py:203
constexpr float kLastDeltaTrans = 1e-3f;   // rendering.py:204
constexpr float kGreyBg = 0.8f;            // rendering.py:265

__device__ __forceinline__ float softplus_f(float x) {
  // torch.nn.Softplus(beta=1, threshold=20)   (rendering.py:305)
  return x > 20.f ? x : log1pf(expf(x));
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace nsff
