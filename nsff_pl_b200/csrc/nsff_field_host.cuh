// Host-side launchers shared between the fp32 verification path (nsff_field_fp32.cu, where the kernels live) and the
// tensor-core backward (nsff_field_bwd_tc.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nsff_b200.h"

namespace nsff {
// d_code[r, k] = sum over the S samples of ray r of dX[(r,s), col0 + k]
int launch_ray_reduce(const float* dX, int ld, int col0, int n, int S, int64_t R, float* d_code, cudaStream_t st);
// autograd of the sinusoidal encoding w.r.t. the position (nerf.py:26-28)
int launch_pe_bwd(const float* dX, int ld, const float* pos, int n_freqs, float* d_pos, int64_t M, cudaStream_t st);
}  // namespace nsff
