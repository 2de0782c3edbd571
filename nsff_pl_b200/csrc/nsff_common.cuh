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

// sin / cos of the next octave from the current one: sin 2a = 2 sin a cos a, cos 2a = 1 - 2 sin^2 a.  The reference
// evaluates sin(2^f x) with 2^f x exact in fp32, so the recurrence computes the same quantity; its absolute error
// roughly doubles per octave (<= ~5e-5 at 2^9), far below the bf16 rounding (2^-9) applied to the result - and it costs
// 3 instructions per value pair where sincosf costs ~50 (the field kernel's prologue used to take ~13k cycles per tile).  Shared by the forward
// prologue and the backward's input images so that both see bit-identical bf16 encodings.
__device__ __forceinline__ void pe_next_octave(float (&sn)[3], float (&cs)[3]) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float s2 = 2.f * sn[c];
    const float nc = fmaf(-s2, sn[c], 1.f);
    sn[c] = s2 * cs[c];
    cs[c] = nc;
  }
}

inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace nsff
