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

// ---- backward chain (nsff_field_tc.cu: field_forward_kernel<.., kBwd = true>) -----------------------------------------
// One launch walks every 128-point tile through the whole dgrad chain of a branch set: dZ of a layer stays in tensor
// memory as the A operand of the next (earlier) layer's MMAs; only the bf16 images the weight-gradient GEMMs need and the
// fp32 input-gradient rows go to global memory.  The caller describes the chain layer by layer.
namespace tc {
enum ChainEpi : int32_t { CHAIN_LIN = 0, CHAIN_MASK = 1, CHAIN_ROWS = 2 };
enum ChainChunk : uint8_t { CC_H0 = 0, CC_H1, CC_H2, CC_H3, CC_DP_STATIC = 4, CC_DP_TRANS = 5 };
struct ChainLayer {
  uint32_t w_off;        // byte offset (in `pack`) of block 0; blocks are [block_rows x 64] bf16, 128B-swizzled, K-major
  int32_t block_rows;    // 256, or 128 for the trunk-input blocks
  int32_t row0;          // first block row used (CHAIN_ROWS: 0 or 64)
  int32_t n;             // 256 (hidden layer = dZ image) or 64 (CHAIN_ROWS)
  int32_t nchunks;
  uint8_t chunk[6];      // A source of K-chunk kc (ChainChunk); hidden chunks = dZ of the previous chain layer
  uint8_t blk[6];        // block index of K-chunk kc
  int32_t epi;           // ChainEpi
  uint64_t save_off;     // hidden: offset of the dZ tile image in `save` (~0: not saved)
  uint64_t bits_off;     // CHAIN_MASK: offset of the forward's ReLU bit plane in `mask`
  int32_t out_col0;      // CHAIN_ROWS: first destination column in `out_rows`; slot 0/1 = out_col0 / 64
  int32_t beta;          // CHAIN_ROWS: 1 = add to what is there
};
int launch_backward_chain(const ChainLayer* layers, int n_layers, const void* pack, const uint8_t* dp_static,
                          const uint8_t* dp_trans, int64_t B, const uint8_t* mask, uint8_t* save, float* out_rows,
                          int32_t out_ld, cudaStream_t st);
}  // namespace tc
}  // namespace nsff
