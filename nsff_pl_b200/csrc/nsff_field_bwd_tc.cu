// Tensor-core backward of the NeRF field (autograd of models/nerf.py:118-213, reference train.py:178-198).
//
// Layer-wise bf16 GEMMs on tcgen05 with fp32 accumulation in TMEM.  Every activation / gradient plane lives in HBM
// as a "tile image": [tile of 128 points][64-column chunk][128 rows x 128 B, 128-byte swizzled] - the exact shared
// memory operand layout - so every operand load is ONE 1-D bulk copy (cp.async.bulk + mbarrier complete_tx), and the
// SAME bytes serve both GEMM flavours:
//   * img_gemm_kernel   out[P, N]   = epi( A[P, K] . B[N, K]^T )      A, B K-major   (forward recompute, dgrad)
//   * img_wgrad_kernel  dW^T[N, M] += sum_p B[p, n] . A[p, m]         A, B MN-major  (wgrad: K = points)
// The point tile that is a K-major A operand of the dgrad ([128 pts x 64 ch], rows = points) is, read as
// "MN-major with K = points", exactly the canonical SWIZZLE_128B MN-major atom (64 MN elements contiguous, 8 K rows of
// 128 B per 1024-B group): no transposed copy of any activation is ever made.
//
// Both kernels are HBM-bound at this shape (a 256x256 layer moves 134-200 MB per 131k points = 20-30 us, its MMAs take
// 8 us), so they are written for streaming: weights resident in smem (img_gemm), deep bulk-copy rings, accumulators
// double-buffered in TMEM so the epilogue of tile i overlaps the loads/MMAs of tile i+1.
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "nsff_common.cuh"
#include "nsff_field_host.cuh"
#include "nsff_field_save.cuh"
#include "nsff_sm100.cuh"

#define TRY(expr)                   \
  do {                              \
    int _rc = (expr);               \
    if (_rc != NSFF_OK) return _rc; \
  } while (0)

namespace nsff {
namespace bwd {
using namespace sm100;

constexpr uint32_t kChunkBytes = 128 * 128;   // 128 rows x 64 bf16
constexpr uint32_t kSmemLimit = 232448;       // 227 KB opt-in
constexpr int kMaxK = 6;                      // K chunks of one GEMM

// bulk copy with an L2 evict-first policy: operand tiles are consumed once per kernel, the output of the kernel should
// be what stays in the 126 MB L2 for the next launch
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst_smem),
      "l"(src), "r"(bytes), "r"(bar), "l"(pol)
      : "memory");
}

// Programmatic dependent launch: the next GEMM of the chain may start its preamble (barrier init, TMEM allocation,
// weight blocks into shared memory) while this one drains; nothing that depends on the previous kernel's output is
// touched before griddep_wait().  Both are no-ops when the launch does not carry the PDL attribute.
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ uint32_t row_off(int row) { return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128); }

// ------------------------------------------------------------------------------------------------------------------
// small image kernels
// ------------------------------------------------------------------------------------------------------------------
// fp32 row-major columns [col0, col0+ncols) -> one chunk of an image (zero padded to 64 columns / 128-row tiles)
__global__ void rows_to_img_kernel(const float* __restrict__ src, int ld, int col0, int ncols, int64_t n_rows,
                                   uint8_t* __restrict__ img, uint32_t tile_stride, int n_tiles) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;       // one thread per (row, 16-byte unit)
  if (idx >= (int64_t)n_tiles * 128 * 8) return;
  const int unit = (int)(idx & 7);
  const int64_t r = idx >> 3;
  const int row = (int)(r & 127);
  const int64_t tile = r >> 7;
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float v[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c = unit * 8 + i * 2 + j;
      v[j] = (r < n_rows && c < ncols) ? src[r * ld + col0 + c] : 0.f;
    }
    w[i] = pack_bf16x2(v[0], v[1]);
  }
  uint4* dst = reinterpret_cast<uint4*>(img + tile * tile_stride + row_off(row) + (((unit ^ row) & 7) << 4));
  *dst = make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void img_to_rows_kernel(const uint8_t* __restrict__ img, uint32_t tile_stride, float* __restrict__ dst,
                                   int ld, int col0, int ncols, int64_t n_rows) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_rows * 64) return;
  const int c = (int)(idx & 63);
  const int64_t r = idx >> 6;
  if (c >= ncols) return;
  const int row = (int)(r & 127);
  const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(img + (r >> 7) * tile_stride + row_off(row) +
                                                                  ((((c >> 3) ^ row) & 7) << 4)) + (c & 7);
  dst[r * ld + col0 + c] = __bfloat162float(*p);
}

// weight block [N rows x 64 cols] bf16, 128B-swizzled: blk[n][dk0 + k] = T ? src[(k0+k)*ld + n0+n] : src[(n0+n)*ld + k0+k]
struct PackJob {
  const float* src;
  uint8_t* dst;
  int32_t ld, transpose, n0, k0, n_valid, k_valid, dk0, dn0;
};
constexpr int kPackJobsPerLaunch = 64;
struct PackJobs {
  PackJob job[kPackJobsPerLaunch];
};
__global__ void pack_blocks_kernel(const __grid_constant__ PackJobs J) {
  const PackJob& j = J.job[blockIdx.x];
  const int total = j.n_valid * j.k_valid;
  for (int e = threadIdx.x + blockIdx.y * blockDim.x; e < total; e += blockDim.x * gridDim.y) {
    int n, k;
    if (j.transpose) n = e % j.n_valid, k = e / j.n_valid;     // consecutive threads walk the contiguous source dim
    else k = e % j.k_valid, n = e / j.k_valid;
    const float v = j.transpose ? j.src[(int64_t)(j.k0 + k) * j.ld + j.n0 + n] : j.src[(int64_t)(j.n0 + n) * j.ld + j.k0 + k];
    const int row = j.dn0 + n, col = j.dk0 + k;
    *reinterpret_cast<__nv_bfloat16*>(j.dst + row_off(row) + ((((col >> 3) ^ row) & 7) << 4) + (col & 7) * 2) =
        __float2bfloat16_rn(v);
  }
}

// pre-activation gradients of all head columns of one branch: fp32 (B,16) is not kept, only the bf16 image chunk
// (columns >= n and rows >= B are zero) and the bias gradients.
struct HeadSpec {
  int8_t src[16];        // image column i <- column src[i] of out / d_out (-1: zero column)
  uint8_t act[16];       // 0 identity, 1 sigmoid, 2 flow
  float* db[16];         // where the bias gradient of column i accumulates (NULL = none)
};
__global__ void __launch_bounds__(128) heads_dpre_kernel(const float* __restrict__ d_out, const float* __restrict__ out,
                                                         int ld, const __grid_constant__ HeadSpec H,
                                                         const float* __restrict__ zs, float scale, int64_t B,
                                                         uint8_t* __restrict__ img, uint32_t tile_stride, int n_tiles) {
  __shared__ float red_s[4][16];
  const int row = threadIdx.x;                       // one thread per image row; the block strides over the tiles
  float sum[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) sum[i] = 0.f;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t r = (int64_t)tile * 128 + row;
    float g[16];
    const bool valid = r < B;
    const bool far = valid && zs && zs[r] > kZFar;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float v = 0.f;
      if (valid && H.src[i] >= 0) {
        const float go = d_out[r * ld + H.src[i]], y = out[r * ld + H.src[i]];
        if (H.act[i] == 1) v = go * y * (1.f - y);
        else if (H.act[i] == 2) {
          const float t = y / scale;
          v = far ? 0.f : go * scale * (1.f - t * t);
        } else v = go;
      }
      g[i] = v;
      sum[i] += v;
    }
    uint8_t* rowp = img + (size_t)tile * tile_stride + row_off(row);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      uint4 q = make_uint4(0, 0, 0, 0);
      if (u < 2) q = make_uint4(pack_bf16x2(g[u * 8], g[u * 8 + 1]), pack_bf16x2(g[u * 8 + 2], g[u * 8 + 3]),
                                pack_bf16x2(g[u * 8 + 4], g[u * 8 + 5]), pack_bf16x2(g[u * 8 + 6], g[u * 8 + 7]));
      *reinterpret_cast<uint4*>(rowp + (((u ^ row) & 7) << 4)) = q;
    }
  }
  // bias gradients: warp reduce, block reduce, ONE atomic per column and block (every block hits the same <= 10 words)
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float v = sum[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red_s[threadIdx.x >> 5][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < 16 && H.src[threadIdx.x] >= 0 && H.db[threadIdx.x])
    atomicAdd(H.db[threadIdx.x], red_s[0][threadIdx.x] + red_s[1][threadIdx.x] + red_s[2][threadIdx.x] + red_s[3][threadIdx.x]);
}

// The MLP input row of every point (rendering.py:164-173) as bf16 chunk images, one launch: PE(xyz) (63 -> 64 columns),
// time code (48), PE(dir) (27), appearance code (48).  One thread per (row, kind): a full 128-byte image row each.
struct EmbedImgArgs {
  const float* xyz;       // (B,3)
  const float* dirs;      // (R,3)
  const float* t_table;   // (R,in_t) per-ray rows or NULL
  const float* a_table;   // (R,in_a) per-ray rows or NULL
  int64_t B;
  int32_t S, n_tiles, in_xyz, in_dir, in_t, in_a;
  uint8_t* img[4];        // PE, T, DIR, A (NULL = skip)
};
// sin/cos of one frequency for the three components, written to columns 3+6f .. 3+6f+5 of a 64-wide row buffer
template <int kNFreq>
__device__ __forceinline__ void pe_row(const float (&b)[3], float (&row)[64]) {
  // same evaluation as the forward prologue (nsff_field_tc.cu: one sincosf at octave 0, then the doubling recurrence)
  float sn[3], cs[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    row[c] = b[c];
    sincosf(b[c], &sn[c], &cs[c]);                 // freqs = 2^linspace(0, N-1, N)  (nerf.py:13)
  }
#pragma unroll
  for (int f = 0; f < kNFreq; ++f) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      row[3 + 6 * f + c] = sn[c];
      row[3 + 6 * f + 3 + c] = cs[c];
    }
    if (f + 1 < kNFreq) pe_next_octave(sn, cs);
  }
}
__global__ void __launch_bounds__(128) embed_img_kernel(const __grid_constant__ EmbedImgArgs E) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // one thread per (row, kind)
  if (idx >= (int64_t)E.n_tiles * 128 * 4) return;
  const int kind = (int)(idx & 3);
  const int64_t r = idx >> 2;
  if (!E.img[kind]) return;
  const int row = (int)(r & 127);
  float v[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = 0.f;
  if (r < E.B) {
    const int64_t ray = r / E.S;
    if (kind == 0) {
      const float b[3] = {E.xyz[r * 3], E.xyz[r * 3 + 1], E.xyz[r * 3 + 2]};
      pe_row<10>(b, v);
    } else if (kind == 2) {
      const float b[3] = {E.dirs[ray * 3], E.dirs[ray * 3 + 1], E.dirs[ray * 3 + 2]};
      pe_row<4>(b, v);
    } else {
      const float* tab = kind == 1 ? E.t_table : E.a_table;
      if (tab) {
#pragma unroll
        for (int i = 0; i < 48; ++i) v[i] = tab[ray * 48 + i];
      }
    }
  }
  uint8_t* rowp = E.img[kind] + (r >> 7) * kChunkBytes + row_off(row);
#pragma unroll
  for (int u = 0; u < 8; ++u)
    *reinterpret_cast<uint4*>(rowp + (((u ^ row) & 7) << 4)) =
        make_uint4(pack_bf16x2(v[u * 8], v[u * 8 + 1]), pack_bf16x2(v[u * 8 + 2], v[u * 8 + 3]),
                   pack_bf16x2(v[u * 8 + 4], v[u * 8 + 5]), pack_bf16x2(v[u * 8 + 6], v[u * 8 + 7]));
}

// scatter of a wgrad result: acc_t is [(64*nb) x 256] fp32 with acc_t[n*256 + m] = sum_p B[p,n] A[p,m].
//   transposed mode (layers): dW[m*ldw + dcol0[j] + c] += acc_t[(64j + c)*256 + m]   for c < ncols[j]
//   direct mode (heads, operands swapped): row n of acc_t is the gradient of head row n: dst[n][m] += acc_t[n*256 + m]
struct FinishJob {
  const float* acc_t;
  float* dW;
  int32_t ldw, nb, direct;
  int32_t dcol0[4], ncols[4];
  float* drow[16];           // direct mode: destination of acc_t row n (256 floats) or NULL
};
constexpr int kFinishBatch = 16;
struct FinishJobs {
  FinishJob job[kFinishBatch];
};
__global__ void wgrad_finish_kernel(const __grid_constant__ FinishJobs JJ) {
  const FinishJob& J = JJ.job[blockIdx.y];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = idx >> 8, m = idx & 255;
  if (J.direct) {
    if (n < 16 && J.drow[n]) J.drow[n][m] += J.acc_t[n * 256 + m];
    return;
  }
  if (n >= J.nb * 64) return;
  const int j = n >> 6, c = n & 63;
  if (c < J.ncols[j]) J.dW[(int64_t)m * J.ldw + J.dcol0[j] + c] += J.acc_t[n * 256 + m];
}

// ------------------------------------------------------------------------------------------------------------------
// img_gemm_kernel: K-major GEMM, weights resident in shared memory
// ------------------------------------------------------------------------------------------------------------------
enum GemmEpi : int32_t { GE_RELU_IMG = 0, GE_LIN_IMG = 1, GE_MASK_IMG = 2, GE_F32_ROWS = 3, GE_MASK_BITS = 4 };
// ReLU bit planes: per 128-row tile [8 words][128 rows] uint32, bit i of word w = (activation[row][32w + i] > 0).
// The recompute writes them next to each activation image; the dgrad epilogue reads 16 B per row instead of 512 B.
constexpr uint32_t kBitsPerTile = 8 * 128;   // uint32 per tile

struct GemmParams {
  const uint8_t* a_base[kMaxK];   // chunk c of tile t: a_base[c] + t * a_stride[c]
  uint32_t a_stride[kMaxK];
  uint8_t kcnt[8];                // K=16 MMAs issued from chunk c (1..4)
  int32_t n_tiles, n_chunks, N, stages, epi;
  const uint8_t* b_blocks;        // n_chunks blocks of N x 64 bf16
  const float* bias;              // N floats or NULL
  uint8_t* out_img;               // N/64 chunks per tile
  uint32_t out_stride;
  const uint8_t* mask_img;        // same geometry as out_img (GE_MASK_IMG): out = acc where mask > 0 else 0
  uint32_t mask_stride;
  float* out_f32;                 // GE_F32_ROWS: (n_rows, ld), columns [0, ncols)
  int32_t ld, ncols, beta;
  int64_t n_rows;
  uint32_t* relu_bits_out;        // image epilogues, optional: bit plane of (out > 0)
  const uint32_t* mask_bits;      // GE_MASK_BITS: out = acc where bit set else 0
  float* colsum;                  // image epilogues, optional: colsum[c] += sum over rows of out[:, c]  (fp32, pre-rounding)
  int32_t col0;                   // first output column of this launch inside the (<= 256 wide) output plane
  int32_t l2_hint;                // 1: operand tiles are loaded with an L2 evict-first policy
};

constexpr int kGemmEpiWarps = 16;    // 4 TMEM lane quadrants x 4 column parts: 4 warps per SMSP hide the TMEM-load and
                                     // dependent-arithmetic latency of the epilogue (8 warps: 6 cycles per instruction)
constexpr int kGemmThreads = (kGemmEpiWarps + 2) * 32;
constexpr int kGemmMaxStages = 8;
enum { GB_B_FULL = 0, GB_A_FULL = 1, GB_A_EMPTY = GB_A_FULL + kGemmMaxStages, GB_ACC_FULL = GB_A_EMPTY + kGemmMaxStages,
       GB_ACC_EMPTY = GB_ACC_FULL + 2, GB_COUNT = GB_ACC_EMPTY + 2 };
constexpr uint32_t kGemmAux = 1024 /*bias*/ + 256 /*barriers + tmem ptr*/;

template <int kEpi>
__global__ void __launch_bounds__(kGemmThreads, 1) img_gemm_kernel(const __grid_constant__ GemmParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t blk_bytes = (uint32_t)P.N * 128u;
  const uint32_t sB = base;
  const uint32_t sA = sB + (uint32_t)P.n_chunks * blk_bytes;
  const uint32_t sBias = sA + (uint32_t)P.stages * kChunkBytes;
  const uint32_t sBar = sBias + 1024;
  const uint32_t sTmemPtr = sBar + GB_COUNT * 8;
  auto bar = [&](int i) { return sBar + (uint32_t)i * 8u; };
  float* bias_s = reinterpret_cast<float*>(smem_raw + (sBias - raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(bar(GB_B_FULL), 1);
    for (int i = 0; i < kGemmMaxStages; ++i) {
      mbar_init(bar(GB_A_FULL + i), 1);
      mbar_init(bar(GB_A_EMPTY + i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(GB_ACC_FULL + i), 1);
      mbar_init(bar(GB_ACC_EMPTY + i), kGemmEpiWarps);
    }
    fence_mbar_init();
  }
  griddep_launch_dependents();
  if (warp == kGemmEpiWarps + 1) tmem_alloc(sTmemPtr, 512);
  for (int i = threadIdx.x; i < P.N; i += blockDim.x) bias_s[i] = P.bias ? P.bias[i] : 0.f;
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(sTmemPtr));

  if (warp == kGemmEpiWarps) {
    // ===== producer: weights once, then the A chunks of this CTA's tiles through the ring =====
    if (elect_one()) {
      mbar_arrive_expect_tx(bar(GB_B_FULL), (uint32_t)P.n_chunks * blk_bytes);
      for (int c = 0; c < P.n_chunks; ++c)
        bulk_g2s(sB + (uint32_t)c * blk_bytes, P.b_blocks + (size_t)c * blk_bytes, blk_bytes, bar(GB_B_FULL));
      griddep_wait();                       // weights are constants of the backward; the A tiles are not
      uint32_t stage = 0, par = 1;
      const uint64_t pol = policy_evict_first();
      for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
        for (int c = 0; c < P.n_chunks; ++c) {
          mbar_wait(bar(GB_A_EMPTY + stage), par);
          mbar_arrive_expect_tx(bar(GB_A_FULL + stage), kChunkBytes);
          if (P.l2_hint)
            bulk_g2s_hint(sA + stage * kChunkBytes, P.a_base[c] + (size_t)tile * P.a_stride[c], kChunkBytes,
                          bar(GB_A_FULL + stage), pol);
          else
            bulk_g2s(sA + stage * kChunkBytes, P.a_base[c] + (size_t)tile * P.a_stride[c], kChunkBytes,
                     bar(GB_A_FULL + stage));
          if (++stage == (uint32_t)P.stages) stage = 0, par ^= 1u;
        }
      }
    }
  } else if (warp == kGemmEpiWarps + 1) {
    // ===== MMA issuer =====
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(128, (uint32_t)P.N);
      mbar_wait(bar(GB_B_FULL), 0);
      uint32_t stage = 0, par = 0, local = 0;
      for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++local) {
        const uint32_t slot = local & 1u, use = local >> 1;
        mbar_wait(bar(GB_ACC_EMPTY + slot), (use & 1u) ^ 1u);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + slot * 256u;
        for (int c = 0; c < P.n_chunks; ++c) {
          mbar_wait(bar(GB_A_FULL + stage), par);
          tc_fence_after_sync();
          const uint32_t a0 = sA + stage * kChunkBytes, b0 = sB + (uint32_t)c * blk_bytes;
          const int kc = P.kcnt[c];
          for (int k = 0; k < kc; ++k)
            mma_bf16_ss(d_tmem, make_desc_k128(a0 + (uint32_t)k * 32u), make_desc_k128(b0 + (uint32_t)k * 32u), idesc,
                        (c | k) != 0);
          mma_commit(bar(GB_A_EMPTY + stage));
          if (++stage == (uint32_t)P.stages) stage = 0, par ^= 1u;
        }
        mma_commit(bar(GB_ACC_FULL + slot));
      }
    }
  } else {
    // ===== epilogue: warp (q, part) drains rows 32q.. of a quarter of the columns (32 at least) =====
    griddep_wait();                         // reads bit planes / accumulates into rows written by earlier kernels
    const int q = warp & 3, part = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t roff = row_off(row);
    const int rsw = row & 7;
    const int cols_per_part = P.N >= 128 ? P.N / 4 : 32;
    const int n_batches = part * cols_per_part < P.N ? cols_per_part / 32 : 0;       // 32-column batches: 0, 1 or 2
    const int c_begin = part * cols_per_part;
    const int word0 = (P.col0 + c_begin) >> 5;
    float colacc[2] = {0.f, 0.f};
    uint32_t local = 0;
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++local) {
      const uint32_t slot = local & 1u, use = local >> 1;
      // the ReLU bits of this row are fetched before the accumulator is awaited: off the critical path
      uint32_t mbits[2] = {~0u, ~0u};
      if constexpr (kEpi == GE_MASK_BITS) {
#pragma unroll
        for (int b = 0; b < 2; ++b)
          if (b < n_batches) mbits[b] = P.mask_bits[(size_t)tile * kBitsPerTile + (size_t)(word0 + b) * 128 + row];
      }
      mbar_wait(bar(GB_ACC_FULL + slot), use & 1u);
      tc_fence_after_sync();
      const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + slot * 256u;
      const int64_t p = (int64_t)tile * 128 + row;
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        if (b >= n_batches) break;
        const int cb = c_begin + 32 * b;
        uint32_t v[32];
        tmem_ld32(t_acc + (uint32_t)cb, v);
        tmem_ld_wait();
        if constexpr (kEpi == GE_F32_ROWS) {
          if (p < P.n_rows) {
            float* o = P.out_f32 + p * P.ld + cb;
            const bool vec = ((P.ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(P.out_f32) & 15) == 0);
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              if (vec && cb + i + 4 <= P.ncols) {
                float4 x = make_float4(__uint_as_float(v[i]) + bias_s[cb + i], __uint_as_float(v[i + 1]) + bias_s[cb + i + 1],
                                       __uint_as_float(v[i + 2]) + bias_s[cb + i + 2], __uint_as_float(v[i + 3]) + bias_s[cb + i + 3]);
                if (P.beta) {
                  const float4 y = *reinterpret_cast<const float4*>(o + i);
                  x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
                }
                *reinterpret_cast<float4*>(o + i) = x;
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  if (cb + i + j < P.ncols) {
                    const float x = __uint_as_float(v[i + j]) + bias_s[cb + i + j];
                    o[i + j] = P.beta ? o[i + j] + x : x;
                  }
                }
              }
            }
          }
        } else {
          const int gc = P.col0 + cb;                                   // column inside the output plane
          const size_t off = (size_t)tile * P.out_stride + (size_t)(gc >> 6) * kChunkBytes + roff;
          const int unit0 = (gc & 63) >> 3;
          uint32_t mb = mbits[b];
          if constexpr (kEpi == GE_MASK_IMG) {
            // ReLU mask given as an activation image: max(x,0) in bf16, so "> 0" == "any magnitude bit set"
            mb = 0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const uint4 mk = *reinterpret_cast<const uint4*>(P.mask_img + (size_t)tile * P.mask_stride +
                                                               (size_t)(gc >> 6) * kChunkBytes + roff +
                                                               ((uint32_t)(((unit0 + u) ^ rsw) & 7) << 4));
              const uint32_t w4[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                mb |= ((w4[k] & 0x7FFFu) ? 1u : 0u) << (u * 8 + 2 * k);
                mb |= ((w4[k] & 0x7FFF0000u) ? 1u : 0u) << (u * 8 + 2 * k + 1);
              }
            }
          }
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 bb = *reinterpret_cast<const float4*>(bias_s + cb + i);
            f[i] = __uint_as_float(v[i]) + bb.x;
            f[i + 1] = __uint_as_float(v[i + 1]) + bb.y;
            f[i + 2] = __uint_as_float(v[i + 2]) + bb.z;
            f[i + 3] = __uint_as_float(v[i + 3]) + bb.w;
          }
          if constexpr (kEpi == GE_RELU_IMG) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
          }
          if constexpr (kEpi == GE_MASK_IMG || kEpi == GE_MASK_BITS) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = ((mb >> i) & 1u) ? f[i] : 0.f;
          }
          // two aligned 32-byte sectors per row: the swizzle only swaps the 16-byte units inside each pair
          const bool swap = rsw & 1;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int u = 2 * j;
            const uint4 a = make_uint4(pack_bf16x2(f[u * 8], f[u * 8 + 1]), pack_bf16x2(f[u * 8 + 2], f[u * 8 + 3]),
                                       pack_bf16x2(f[u * 8 + 4], f[u * 8 + 5]), pack_bf16x2(f[u * 8 + 6], f[u * 8 + 7]));
            const uint4 b2 = make_uint4(pack_bf16x2(f[u * 8 + 8], f[u * 8 + 9]), pack_bf16x2(f[u * 8 + 10], f[u * 8 + 11]),
                                        pack_bf16x2(f[u * 8 + 12], f[u * 8 + 13]), pack_bf16x2(f[u * 8 + 14], f[u * 8 + 15]));
            st_global_v8(P.out_img + off + ((uint32_t)(((unit0 + u) ^ rsw) & 6) << 4), swap ? b2 : a, swap ? a : b2);
          }
          if (P.relu_bits_out) {
            uint32_t pos_bits = 0;
#pragma unroll
            for (int i = 0; i < 32; ++i) pos_bits |= (f[i] > 0.f ? 1u : 0u) << i;
            P.relu_bits_out[(size_t)tile * kBitsPerTile + (size_t)(gc >> 5) * 128 + row] = pos_bits;
          }
          if (P.colsum) {
            // transpose-reduce over the warp's 32 rows: after the butterfly lane L holds the sum of column cb + L
#pragma unroll
            for (int o2 = 16; o2 >= 1; o2 >>= 1) {
              const bool upper = (lane & o2) != 0;
#pragma unroll
              for (int i = 0; i < o2; ++i) {
                const float keep = upper ? f[i + o2] : f[i];
                const float send = upper ? f[i] : f[i + o2];
                f[i] = keep + __shfl_xor_sync(0xffffffffu, send, o2);
              }
            }
            colacc[b] += f[0];
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(GB_ACC_EMPTY + slot));
    }
    if constexpr (kEpi != GE_F32_ROWS) {
      if (P.colsum) {
#pragma unroll
        for (int b = 0; b < 2; ++b)
          if (b < n_batches) atomicAdd(P.colsum + c_begin + 32 * b + lane, colacc[b]);
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == kGemmEpiWarps + 1) {
    tc_fence_after_sync();
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// img_wgrad_kernel: both operands MN-major (K = points), split over the point range, fp32 RED into acc_t
// ------------------------------------------------------------------------------------------------------------------
// One launch runs up to kWgMaxJobs independent weight-gradient GEMMs side by side (the chained backward has all dZ images
// before the first one starts): each job owns a contiguous range of CTAs sized by its bytes per tile, so a CTA walks
// 4-8 times more tiles per launch ramp and per accumulator flush than with one job per launch.
struct WgradJob {
  const uint8_t* a_base;          // 4 chunks (M = 256 channels)
  uint32_t a_stride;
  int32_t nb;
  const uint8_t* b_base[4];       // nb chunks (N = 64*nb channels)
  uint32_t b_stride[4];
  float* acc_t;                   // [(64*nb) x 256]
  float* colsum;                  // optional: colsum[m] += sum over the points of A[:, m] (bias gradient of the layer whose
                                  // dZ image is A), computed by the otherwise idle epilogue warps from the staged tiles
};
constexpr int kWgMaxJobs = 8;
struct WgradParams {
  WgradJob job[kWgMaxJobs];
  int32_t cta0[kWgMaxJobs + 1];   // job j runs on CTAs [cta0[j], cta0[j+1])
  int32_t n_jobs, n_tiles, l2_hint;
};
constexpr int kWgEpiWarps = 8;
constexpr int kWgThreads = (kWgEpiWarps + 2) * 32;
constexpr int kWgStages = 3;
constexpr uint32_t kHalfChunk = 8192;                 // 64 points x 128 B
constexpr uint32_t kWgStageBytes = 8 * kHalfChunk;    // 4 A + up to 4 B half chunks
enum { WB_FULL = 0, WB_EMPTY = WB_FULL + kWgStages, WB_ACC = WB_EMPTY + kWgStages, WB_COUNT };
constexpr uint32_t kWgSmem = kWgStages * kWgStageBytes + 256 + 1024;

// Shared-memory descriptor, MN-major operand, 128-byte swizzle (canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in
// 16-byte units): 64 MN elements contiguous; LBO = byte distance between 64-element MN groups; SBO = byte distance
// between groups of 8 K rows.
__device__ __forceinline__ uint64_t make_desc_mn128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1u << 46;
  d |= (uint64_t)2u << 61;
  return d;
}

__global__ void __launch_bounds__(kWgThreads, 1) img_wgrad_kernel(const __grid_constant__ WgradParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sBar = base + kWgStages * kWgStageBytes;
  const uint32_t sTmemPtr = sBar + WB_COUNT * 8;
  auto bar = [&](int i) { return sBar + (uint32_t)i * 8u; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int ji = 0;
  while (ji + 1 < P.n_jobs && (int)blockIdx.x >= P.cta0[ji + 1]) ++ji;
  const WgradJob& J = P.job[ji];
  const int bi = (int)blockIdx.x - P.cta0[ji], nbi = P.cta0[ji + 1] - P.cta0[ji];
  // contiguous tile range of this CTA inside its job
  const int t0 = (int)((int64_t)P.n_tiles * bi / nbi);
  const int t1 = (int)((int64_t)P.n_tiles * (bi + 1) / nbi);
  const int n_half = (t1 - t0) * 2;

#ifdef NSFF_WATCHDOG
  if (threadIdx.x == 0 && blockIdx.x == 0) printf("NSFF_WATCHDOG wgrad: barriers @%u (FULL[3] EMPTY[3] ACC), colsum %d\n", sBar, J.colsum != nullptr);
#endif
  if (threadIdx.x == 0) {
    for (int i = 0; i < kWgStages; ++i) {
      mbar_init(bar(WB_FULL + i), 1);
      mbar_init(bar(WB_EMPTY + i), J.colsum ? 1 + kWgEpiWarps : 1);    // MMA commit (+ the column-sum warps)
    }
    mbar_init(bar(WB_ACC), 1);
    fence_mbar_init();
  }
  griddep_launch_dependents();
  if (warp == kWgEpiWarps + 1) tmem_alloc(sTmemPtr, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(sTmemPtr));
  const uint32_t N = (uint32_t)J.nb * 64u;
  if (warp != kWgEpiWarps + 1) griddep_wait();       // operand tiles and the RED target belong to earlier kernels

  if (n_half > 0) {
    if (warp == kWgEpiWarps) {
      if (elect_one()) {
        uint32_t stage = 0, par = 1;
        const uint64_t pol = policy_evict_first();
        const uint32_t bytes = (4u + (uint32_t)J.nb) * kHalfChunk;
        for (int h = 0; h < n_half; ++h) {
          const size_t tile = (size_t)(t0 + (h >> 1));
          const size_t hoff = (size_t)(h & 1) * kHalfChunk;
          mbar_wait(bar(WB_EMPTY + stage), par);
          mbar_arrive_expect_tx(bar(WB_FULL + stage), bytes);
          const uint32_t s0 = base + stage * kWgStageBytes;
          for (int c = 0; c < 4 + J.nb; ++c) {
            const uint8_t* src = c < 4 ? J.a_base + tile * J.a_stride + (size_t)c * kChunkBytes + hoff
                                       : J.b_base[c - 4] + tile * J.b_stride[c - 4] + hoff;
            if (P.l2_hint) bulk_g2s_hint(s0 + (uint32_t)c * kHalfChunk, src, kHalfChunk, bar(WB_FULL + stage), pol);
            else bulk_g2s(s0 + (uint32_t)c * kHalfChunk, src, kHalfChunk, bar(WB_FULL + stage));
          }
          if (++stage == kWgStages) stage = 0, par ^= 1u;
        }
      }
    } else if (warp == kWgEpiWarps + 1) {
      if (elect_one()) {
        // M = 128 channels of A per MMA, N = 64*nb channels of B, K = 16 points; both MN-major (bits 15, 16)
        const uint32_t idesc = make_idesc_bf16(128, N) | (1u << 15) | (1u << 16);
        const uint32_t lbo = kHalfChunk, sbo = 1024u;     // MN groups (64 channels) 8 KB apart, K groups (8 points) 1 KB apart
        uint32_t stage = 0, par = 0;
        for (int h = 0; h < n_half; ++h) {
          mbar_wait(bar(WB_FULL + stage), par);
          tc_fence_after_sync();
          const uint32_t s0 = base + stage * kWgStageBytes;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t bdesc = make_desc_mn128(s0 + 4u * kHalfChunk + (uint32_t)kk * 2048u, lbo, sbo);
#pragma unroll
            for (int m = 0; m < 2; ++m) {
              const uint64_t adesc = make_desc_mn128(s0 + (uint32_t)(2 * m) * kHalfChunk + (uint32_t)kk * 2048u, lbo, sbo);
              mma_bf16_ss(tmem_base + (uint32_t)m * 256u, adesc, bdesc, idesc, (h | kk) != 0);
            }
          }
          mma_commit(bar(WB_EMPTY + stage));
          if (++stage == kWgStages) stage = 0, par ^= 1u;
        }
        mma_commit(bar(WB_ACC));
      }
    } else {
      const int q = warp & 3, mh = warp >> 2;
      if (J.colsum) {
        // column sums of the A image (256 channels, one per thread) from the staged half tiles: 64 points x 2 bytes per
        // thread and stage; the 32 lanes of a warp read 64 contiguous (swizzled) bytes of each row
        const int ch = threadIdx.x, c = ch >> 6, col = ch & 63;
        float sum = 0.f;
        uint32_t stage = 0, par = 0;
        for (int h = 0; h < n_half; ++h) {
          mbar_wait(bar(WB_FULL + stage), par);
          const uint32_t a0 = base + stage * kWgStageBytes + (uint32_t)c * kHalfChunk + (uint32_t)(col & 7) * 2u;
#pragma unroll 8
          for (int r = 0; r < 64; ++r) {
            uint16_t v;
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a0 + (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((col >> 3) ^ r) & 7) << 4))));
            sum += __uint_as_float((uint32_t)v << 16);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(WB_EMPTY + stage));
          if (++stage == kWgStages) stage = 0, par ^= 1u;
        }
        atomicAdd(J.colsum + ch, sum);
      }
      mbar_wait(bar(WB_ACC), 0);
      tc_fence_after_sync();
      const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)mh * 256u;
      float* dst = J.acc_t + mh * 128 + q * 32 + lane;
      // every CTA adds into the same 64*nb x 256 words: start each CTA at a different column so that the L2 atomic
      // units see different addresses at any one time
      const uint32_t rot = ((uint32_t)bi * 32u) % N;
      for (uint32_t c0 = 0; c0 < N; c0 += 32) {
        const uint32_t cb = (c0 + rot) % N;
        uint32_t v[32];
        tmem_ld32(t_acc + cb, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) atomicAdd(dst + (size_t)(cb + i) * 256, __uint_as_float(v[i]));
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kWgEpiWarps + 1) {
    tc_fence_after_sync();
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
static int g_sm_count = 0;
static int sm_count() {
  if (!g_sm_count) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      return 148;
    g_sm_count = n;
  }
  return g_sm_count;
}

static int check_arch() {
  int dev = 0, major = 0;
  NSFF_CUDA_TRY(cudaGetDevice(&dev));
  NSFF_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  return major == 10 ? NSFF_OK : NSFF_EARCH;
}

static int l2_hint_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("NSFF_BWD_L2HINT");     // A/B knob; measured 6.97 -> 6.82 ms per training step
    v = e ? atoi(e) : 1;
  }
  return v;
}

static int pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("NSFF_BWD_PDL");        // A/B knob.  Measured on the 1024 x 128 training step: backward phase
    v = e ? atoi(e) : 0;                           // 4.19 -> 4.00 ms, but median step 4.92 -> 5.08 ms (more jitter): off
  }
  return v;
}

template <typename Kern, typename Params>
static int launch_pdl(Kern kern, int grid, int threads, uint32_t smem, cudaStream_t st, const Params& P) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  NSFF_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, P));
  return after_launch();
}

template <int kEpi>
static int launch_gemm_t(const GemmParams& P, int grid, uint32_t smem, cudaStream_t st) {
  NSFF_CUDA_TRY(cudaFuncSetAttribute(img_gemm_kernel<kEpi>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
  return launch_pdl(img_gemm_kernel<kEpi>, grid, kGemmThreads, smem, st, P);
}

static int launch_gemm(GemmParams& P, cudaStream_t st) {
  if (P.n_tiles <= 0) return NSFF_OK;
  if (P.n_chunks < 1 || P.n_chunks > kMaxK || (P.N != 256 && P.N != 128 && P.N != 64)) return NSFF_EINVAL;
  const uint32_t fixed = (uint32_t)P.n_chunks * (uint32_t)P.N * 128u + kGemmAux + 1024u;
  int stages = (int)((kSmemLimit - fixed) / kChunkBytes);
  if (stages > kGemmMaxStages) stages = kGemmMaxStages;
  if (stages < 2) return NSFF_EUNSUPPORTED;
  P.stages = stages;
  P.l2_hint = l2_hint_enabled();
  const uint32_t smem = fixed + (uint32_t)stages * kChunkBytes;
  int grid = sm_count();
  if (grid > P.n_tiles) grid = P.n_tiles;
  switch (P.epi) {
    case GE_RELU_IMG: return launch_gemm_t<GE_RELU_IMG>(P, grid, smem, st);
    case GE_LIN_IMG: return launch_gemm_t<GE_LIN_IMG>(P, grid, smem, st);
    case GE_MASK_IMG: return launch_gemm_t<GE_MASK_IMG>(P, grid, smem, st);
    case GE_F32_ROWS: return launch_gemm_t<GE_F32_ROWS>(P, grid, smem, st);
    case GE_MASK_BITS: return launch_gemm_t<GE_MASK_BITS>(P, grid, smem, st);
  }
  return NSFF_EINVAL;
}

static int launch_wgrad(WgradParams& P, cudaStream_t st) {
  if (P.n_tiles <= 0 || P.n_jobs <= 0) return NSFF_OK;
  if (P.n_jobs > kWgMaxJobs) return NSFF_EINVAL;
  int weight = 0;
  for (int j = 0; j < P.n_jobs; ++j) {
    if (P.job[j].nb < 1 || P.job[j].nb > 4) return NSFF_EINVAL;
    weight += 4 + P.job[j].nb;                      // bytes per tile: 4 A chunks + nb B chunks
  }
  NSFF_CUDA_TRY(cudaFuncSetAttribute(img_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWgSmem));
  P.l2_hint = l2_hint_enabled();
  const int sms = sm_count();
  int at = 0;
  for (int j = 0; j < P.n_jobs; ++j) {
    int n = (int)((int64_t)sms * (4 + P.job[j].nb) / weight);
    if (n < 1) n = 1;
    if (n > P.n_tiles) n = P.n_tiles;
    P.cta0[j] = at;
    at += n;
  }
  P.cta0[P.n_jobs] = at;
  return launch_pdl(img_wgrad_kernel, at, kWgThreads, kWgSmem, st, P);
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
constexpr int kTinCol0 = 64;     // column of the first time-code gradient in the (B,128) trunk-input gradient rows

// ---- weight pack of the backward ----------------------------------------------------------------------------------
// All blocks are N x 64 bf16 (N = 256, or 128 for the "trunk input" extras), zero where no weight maps.
struct Blk {
  uint32_t off;
  int n;        // number of consecutive blocks (per half when split)
  int N;
  int split;    // 1: the 256 output columns are packed as two halves of n blocks of 128 x 64 each (two launches): used
                // for 6-chunk layers, whose 192 KB of resident weights would leave only a 2-deep A ring
};
struct BwdPack {
  // forward (recompute): W blocks, K chunks in the order the GEMM consumes them
  Blk s_fwd[8], s_fin, s_dir, t_fwd[8], t_fin;
  // dgrad: W^T blocks, K chunks = output-channel chunks of the layer
  Blk s_hidT[8], s_finT /*4 + sigma-head block*/, s_dirT /*4: feat part*/, s_dirxT /*4, N=128: [dir|a] part*/, s_rgbT /*1*/;
  Blk t_hidT[8], t_finT, t_headT /*1*/, t_in1T /*4, N=128*/, t_in5T /*4, N=128*/;
  size_t bytes;
};

struct PackBuilder {
  std::vector<PackJob> jobs;   // only filled when dst != nullptr
  uint8_t* dst = nullptr;
  uint32_t off = 0;
  Blk alloc(int n, int N, bool split = false) {
    Blk b{off, n, split ? 128 : N, split ? 1 : 0};
    off += (uint32_t)n * (uint32_t)N * 128u;
    return b;
  }
  void job(const Blk& b, int idx, const float* src, int ld, bool T, int n0, int k0, int n_valid, int k_valid, int dk0 = 0,
           int dn0 = 0) {
    if (!dst || !src || n_valid <= 0 || k_valid <= 0) return;
    PackJob j{src, dst + b.off + (size_t)idx * b.N * 128, ld, T ? 1 : 0, n0, k0, n_valid, k_valid, dk0, dn0};
    jobs.push_back(j);
  }
  // forward-type block idx (K chunk) of a possibly split layer: rows [0,256) of src, columns [k0, k0+k_valid)
  void job_fwd(const Blk& b, int idx, const float* src, int ld, int k0, int k_valid) {
    if (!b.split) return job(b, idx, src, ld, false, 0, k0, 256, k_valid);
    for (int h = 0; h < 2; ++h) job(b, h * b.n + idx, src, ld, false, 128 * h, k0, 128, k_valid);
  }
};

static int flush_pack_jobs(PackBuilder& pb, cudaStream_t st) {
  for (size_t i = 0; i < pb.jobs.size(); i += kPackJobsPerLaunch) {
    PackJobs J;
    memset(&J, 0, sizeof(J));
    const int n = (int)std::min<size_t>(kPackJobsPerLaunch, pb.jobs.size() - i);
    for (int k = 0; k < n; ++k) J.job[k] = pb.jobs[i + k];
    pack_blocks_kernel<<<dim3((unsigned)n, 8), 256, 0, st>>>(J);      // 8 blocks per 16K-element job
    TRY(after_launch());
  }
  pb.jobs.clear();
  return NSFF_OK;
}

static int supported(const nsff_field_desc* d) {
  if (d->W != 256 || d->D != 8 || d->skip != 4 || d->in_xyz != 63) return NSFF_EUNSUPPORTED;
  if (d->use_viewdir && d->in_dir != 27) return NSFF_EUNSUPPORTED;
  if (d->in_a != 0 && (d->in_a != 48 || !d->use_viewdir)) return NSFF_EUNSUPPORTED;
  if (d->has_transient && d->in_t != 48) return NSFF_EUNSUPPORTED;
  return NSFF_OK;
}

// Lays the blocks out (always) and, when pb.dst is set, emits the pack jobs reading `p`.
static void build_pack(const nsff_field_desc* d, const nsff_field_params* p, BwdPack& L, PackBuilder& pb) {
  memset(&L, 0, sizeof(L));
  const int W = 256, nx = d->in_xyz, nt = d->in_t, sk = d->skip;
  auto trunk = [&](bool trans, Blk* fwd, Blk* hidT) {
    const int n_in = trans ? nx + nt : nx;
    const float* const* w = p ? (trans ? p->trans_w : p->static_w) : nullptr;
    for (int i = 0; i < 8; ++i) {
      const bool first = i == 0, skip = i == sk;
      const int ld = first ? n_in : (skip ? n_in + W : W);
      const int n_extra = (first || skip) ? (trans ? 2 : 1) : 0;
      const int n_hid = first ? 0 : 4;
      fwd[i] = pb.alloc(n_extra + n_hid, 256, n_extra + n_hid > 5);
      const float* src = w ? w[i] : nullptr;
      if (n_extra) {
        pb.job_fwd(fwd[i], 0, src, ld, 0, nx);
        if (trans) pb.job_fwd(fwd[i], 1, src, ld, nx, nt);
      }
      for (int j = 0; j < n_hid; ++j) pb.job_fwd(fwd[i], n_extra + j, src, ld, (skip ? n_in : 0) + 64 * j, 64);
      if (!first) {
        hidT[i] = pb.alloc(4, 256);
        for (int j = 0; j < 4; ++j) pb.job(hidT[i], j, src, ld, true, skip ? n_in : 0, 64 * j, 256, 64);
      } else {
        hidT[i] = Blk{0, 0, 0, 0};
      }
    }
  };
  trunk(false, L.s_fwd, L.s_hidT);
  L.s_fin = pb.alloc(4, 256);
  L.s_finT = pb.alloc(5, 256);
  for (int j = 0; j < 4; ++j) {
    pb.job(L.s_fin, j, p ? p->static_final_w : nullptr, W, false, 0, 64 * j, 256, 64);
    pb.job(L.s_finT, j, p ? p->static_final_w : nullptr, W, true, 0, 64 * j, 256, 64);
  }
  pb.job(L.s_finT, 4, p ? p->static_sigma_w : nullptr, W, true, 0, 0, 256, 1, /*dk0=*/3);
  L.s_rgbT = pb.alloc(1, 256);
  pb.job(L.s_rgbT, 0, p ? p->static_rgb_w : nullptr, W, true, 0, 0, 256, 3, 0);
  if (d->use_viewdir) {
    const int ldd = W + d->in_dir + d->in_a;
    const float* wd = p ? p->static_dir_w : nullptr;
    L.s_dir = pb.alloc(4 + 1 + (d->in_a > 0 ? 1 : 0), 256, d->in_a > 0);
    L.s_dirT = pb.alloc(4, 256);
    for (int j = 0; j < 4; ++j) {
      pb.job_fwd(L.s_dir, j, wd, ldd, 64 * j, 64);
      pb.job(L.s_dirT, j, wd, ldd, true, 0, 64 * j, 256, 64);
    }
    pb.job_fwd(L.s_dir, 4, wd, ldd, W, d->in_dir);
    if (d->in_a > 0) {
      pb.job_fwd(L.s_dir, 5, wd, ldd, W + d->in_dir, d->in_a);
      L.s_dirxT = pb.alloc(4, 128);
      for (int j = 0; j < 4; ++j) pb.job(L.s_dirxT, j, wd, ldd, true, W, 64 * j, d->in_dir + d->in_a, 64);
    }
  }
  if (d->has_transient) {
    trunk(true, L.t_fwd, L.t_hidT);
    L.t_fin = pb.alloc(4, 256);
    L.t_finT = pb.alloc(4, 256);
    for (int j = 0; j < 4; ++j) {
      pb.job(L.t_fin, j, p ? p->trans_final_w : nullptr, W, false, 0, 64 * j, 256, 64);
      pb.job(L.t_finT, j, p ? p->trans_final_w : nullptr, W, true, 0, 64 * j, 256, 64);
    }
    L.t_headT = pb.alloc(1, 256);
    pb.job(L.t_headT, 0, p ? p->trans_rgb_w : nullptr, W, true, 0, 0, 256, 3, 0);
    pb.job(L.t_headT, 0, p ? p->trans_sigma_w : nullptr, W, true, 0, 0, 256, 1, 3);
    pb.job(L.t_headT, 0, p ? p->trans_fw_w : nullptr, W, true, 0, 0, 256, 3, 4);
    pb.job(L.t_headT, 0, p ? p->trans_bw_w : nullptr, W, true, 0, 0, 256, 3, 7);
    L.t_in1T = pb.alloc(4, 128);
    L.t_in5T = pb.alloc(4, 128);
    // rows 0..62: PE(xyz) columns, rows 64..111: time-code columns (kTinCol0) - the two 64-row halves are separate
    // N = 64 blocks of the backward chain, so a pass that only needs the time-code gradient issues one of them
    for (int j = 0; j < 4; ++j) {
      pb.job(L.t_in1T, j, p ? p->trans_w[0] : nullptr, nx + nt, true, 0, 64 * j, nx, 64);
      pb.job(L.t_in1T, j, p ? p->trans_w[0] : nullptr, nx + nt, true, nx, 64 * j, nt, 64, 0, kTinCol0);
      pb.job(L.t_in5T, j, p ? p->trans_w[sk] : nullptr, nx + nt + W, true, 0, 64 * j, nx, 64);
      pb.job(L.t_in5T, j, p ? p->trans_w[sk] : nullptr, nx + nt + W, true, nx, 64 * j, nt, 64, 0, kTinCol0);
    }
  }
  L.bytes = pb.off;
}

// ---- scratch --------------------------------------------------------------------------------------------------------
struct Scratch {
  uint8_t *PE, *T, *DIR, *A, *SH[8], *SFEAT, *SDIRO, *TH[8], *TFEAT, *G0, *G1, *DP;
  // backward chain: dZ image of every layer (the weight-gradient GEMMs run after the chain), second head-gradient chunk
  uint8_t *GS[8], *GS_FIN, *GS_DIR, *GT[8], *GT_FIN, *DP2, *base;
  uint32_t *SHB[8], *SDIROB, *THB[8];       // ReLU bit planes of the activations above
  float *DIN, *ACC;
  size_t acc_bytes, total;
};
constexpr size_t kAccSlots = 96;          // 64-row x 256 fp32 wgrad result slots per call
// with_acts = false: the activation images come from the forward's save buffer (nsff_field_forward_save)
static Scratch carve(const nsff_field_desc* d, int64_t B, uint8_t* base, bool with_acts) {
  Scratch s;
  const size_t nt = (size_t)cdiv64(B, 128);
  const size_t one = nt * kChunkBytes, four = 4 * one;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    uint8_t* p = base ? base + o : nullptr;
    o += align_up(bytes, 1024);
    return p;
  };
  const size_t bits = nt * kBitsPerTile * sizeof(uint32_t);
  s.PE = take(one); s.T = take(one); s.DIR = take(one); s.A = take(one);
  const size_t a4 = with_acts ? four : 0, ab = with_acts ? bits : 0, t4 = d->has_transient ? a4 : 0,
               tb = d->has_transient ? ab : 0;
  for (int i = 0; i < 8; ++i) s.SH[i] = take(a4), s.SHB[i] = reinterpret_cast<uint32_t*>(take(ab));
  s.SFEAT = take(a4); s.SDIRO = take(a4); s.SDIROB = reinterpret_cast<uint32_t*>(take(ab));
  for (int i = 0; i < 8; ++i) {
    s.TH[i] = take(t4);
    s.THB[i] = reinterpret_cast<uint32_t*>(take(tb));
  }
  s.TFEAT = take(t4);
  s.G0 = take(four); s.G1 = take(four); s.DP = take(one);
  s.base = base;
  {
    const size_t c4 = with_acts ? 0 : four, ct4 = d->has_transient ? c4 : 0;     // chain: only with a save buffer
    for (int i = 0; i < 8; ++i) s.GS[i] = take(c4);
    s.GS_FIN = take(c4); s.GS_DIR = take(d->use_viewdir ? c4 : 0);
    for (int i = 0; i < 8; ++i) s.GT[i] = take(ct4);
    s.GT_FIN = take(ct4);
    s.DP2 = take(with_acts ? 0 : one);
  }
  s.DIN = reinterpret_cast<float*>(take((size_t)B * 128 * sizeof(float)));
  s.acc_bytes = kAccSlots * 64 * 256 * sizeof(float);
  s.ACC = reinterpret_cast<float*>(take(s.acc_bytes));
  s.total = o;
  return s;
}

// ---- orchestration helpers --------------------------------------------------------------------------------------------
struct Src {
  const uint8_t* p;
  uint32_t stride;
  int kcnt;
};
static inline Src chunk_of(const uint8_t* img, int chunks_per_tile, int c, int kcnt = 4) {
  return Src{img + (size_t)c * kChunkBytes, (uint32_t)chunks_per_tile * kChunkBytes, kcnt};
}
struct Ctx {
  cudaStream_t st;
  int n_tiles;
  int64_t B;
  const uint8_t* pack;
  float* acc;
  size_t acc_used, acc_cap;   // in floats
  std::vector<FinishJob> finish;   // scatter jobs of the wgrad results, launched in batches (each owns its acc slot)
  bool batch_wgrads = false;       // chained backward: every dZ image exists up front, so the weight-gradient GEMMs are
  std::vector<WgradJob> pending;   // collected and launched kWgMaxJobs at a time
};

static int flush_wgrads(Ctx& c) {
  for (size_t i = 0; i < c.pending.size(); i += kWgMaxJobs) {
    WgradParams P;
    memset(&P, 0, sizeof(P));
    P.n_jobs = (int)std::min<size_t>(kWgMaxJobs, c.pending.size() - i);
    for (int j = 0; j < P.n_jobs; ++j) P.job[j] = c.pending[i + j];
    P.n_tiles = c.n_tiles;
    TRY(launch_wgrad(P, c.st));
  }
  c.pending.clear();
  return NSFF_OK;
}

static int flush_finish(Ctx& c) {
  TRY(flush_wgrads(c));
  for (size_t i = 0; i < c.finish.size(); i += kFinishBatch) {
    FinishJobs JJ;
    memset(&JJ, 0, sizeof(JJ));
    const int n = (int)std::min<size_t>(kFinishBatch, c.finish.size() - i);
    for (int k = 0; k < n; ++k) JJ.job[k] = c.finish[i + k];
    wgrad_finish_kernel<<<dim3(256, (unsigned)n), 256, 0, c.st>>>(JJ);
    TRY(after_launch());
  }
  c.finish.clear();
  return NSFF_OK;
}

struct GemmOut {
  int epi;
  uint8_t* out_img;                      // 4-chunk (or N/64-chunk) image
  const float* bias = nullptr;
  const uint32_t* mask_bits = nullptr;   // GE_MASK_BITS
  uint32_t* relu_bits = nullptr;         // optional bit plane of (out > 0)
  float* colsum = nullptr;               // optional column sums (bias gradient)
};
static int gemm_img(Ctx& c, const std::vector<Src>& a, const Blk& blk, const GemmOut& o) {
  if ((int)a.size() != blk.n) return NSFF_EINVAL;
  GemmParams P;
  memset(&P, 0, sizeof(P));
  for (size_t i = 0; i < a.size(); ++i) P.a_base[i] = a[i].p, P.a_stride[i] = a[i].stride, P.kcnt[i] = (uint8_t)a[i].kcnt;
  P.n_tiles = c.n_tiles; P.n_chunks = blk.n; P.N = blk.N; P.epi = o.epi;
  P.out_img = o.out_img; P.out_stride = (uint32_t)(blk.split ? 4 : blk.N / 64) * kChunkBytes;
  P.mask_bits = o.mask_bits; P.relu_bits_out = o.relu_bits;
  P.n_rows = c.B;
  for (int h = 0; h < (blk.split ? 2 : 1); ++h) {
    P.b_blocks = c.pack + blk.off + (size_t)h * blk.n * blk.N * 128;
    P.bias = o.bias ? o.bias + h * 128 : nullptr;
    P.colsum = o.colsum ? o.colsum + h * 128 : nullptr;
    P.col0 = h * 128;
    TRY(launch_gemm(P, c.st));
  }
  return NSFF_OK;
}
static int gemm_rows(Ctx& c, const std::vector<Src>& a, const Blk& blk, float* out, int ld, int ncols, int beta) {
  if ((int)a.size() != blk.n) return NSFF_EINVAL;
  GemmParams P;
  memset(&P, 0, sizeof(P));
  for (size_t i = 0; i < a.size(); ++i) P.a_base[i] = a[i].p, P.a_stride[i] = a[i].stride, P.kcnt[i] = (uint8_t)a[i].kcnt;
  P.n_tiles = c.n_tiles; P.n_chunks = blk.n; P.N = blk.N; P.epi = GE_F32_ROWS;
  P.b_blocks = c.pack + blk.off;
  P.out_f32 = out; P.ld = ld; P.ncols = ncols; P.beta = beta; P.n_rows = c.B;
  return launch_gemm(P, c.st);
}
static std::vector<Src> four(const uint8_t* img) {
  return {chunk_of(img, 4, 0), chunk_of(img, 4, 1), chunk_of(img, 4, 2), chunk_of(img, 4, 3)};
}
// acc_t rows = channels of the B chunks, columns = the 256 channels of `a_img`
static int wgrad_img(Ctx& c, const uint8_t* a_img, const std::vector<Src>& b, float** acc_out, float* colsum = nullptr) {
  const size_t need = b.size() * 64 * 256;
  if (b.empty() || b.size() > 4 || c.acc_used + need > c.acc_cap) return NSFF_EINVAL;
  WgradJob J;
  memset(&J, 0, sizeof(J));
  J.a_base = a_img; J.a_stride = 4 * kChunkBytes;
  for (size_t j = 0; j < b.size(); ++j) J.b_base[j] = b[j].p, J.b_stride[j] = b[j].stride;
  J.nb = (int)b.size();
  J.acc_t = c.acc + c.acc_used;
  J.colsum = colsum;
  c.acc_used += need;
  *acc_out = J.acc_t;
  c.pending.push_back(J);
  if (c.batch_wgrads) return NSFF_OK;               // launched by flush_wgrads (at the latest from flush_finish)
  return flush_wgrads(c);
}
static int finish_cols(Ctx& c, const float* acc, float* dW, int ldw, std::initializer_list<int> dcol0,
                       std::initializer_list<int> ncols) {
  if (!dW) return NSFF_EINVAL;
  FinishJob J;
  memset(&J, 0, sizeof(J));
  J.acc_t = acc; J.dW = dW; J.ldw = ldw; J.nb = (int)dcol0.size();
  int i = 0;
  for (int v : dcol0) J.dcol0[i++] = v;
  i = 0;
  for (int v : ncols) J.ncols[i++] = v;
  c.finish.push_back(J);
  return NSFF_OK;
}
static int finish_rows(Ctx& c, const float* acc, float* const (&drow)[16]) {
  FinishJob J;
  memset(&J, 0, sizeof(J));
  J.acc_t = acc; J.direct = 1;
  for (int i = 0; i < 16; ++i) J.drow[i] = drow[i];
  c.finish.push_back(J);
  return NSFF_OK;
}
struct Trunk {
  bool trans;
  const Blk* fwd;
  const Blk* hidT;
  const float* const* b;
  float* const* gw;
  float* const* gb;
  uint8_t* const* H;
  uint32_t* const* HB;     // ReLU bit planes of H
  std::vector<Src> in;     // trunk input chunks ([PE] or [PE, T])
  int n_in;
};

static int trunk_forward(Ctx& c, const nsff_field_desc* d, const Trunk& t) {
  for (int i = 0; i < 8; ++i) {
    std::vector<Src> a;
    if (i == 0 || i == d->skip) a = t.in;
    if (i > 0) {
      auto h = four(t.H[i - 1]);
      a.insert(a.end(), h.begin(), h.end());
    }
    GemmOut o{GE_RELU_IMG, t.H[i]};
    o.bias = t.b[i];
    o.relu_bits = t.HB[i];
    TRY(gemm_img(c, a, t.fwd[i], o));
  }
  return NSFF_OK;
}

// g_cur holds dZ_8 (gradient w.r.t. the pre-activation of the last trunk layer; its column sums - the bias gradient
// of layer 8 - were accumulated by the GEMM that produced it).  d_in (fp32 (B,128)) receives the
// gradient w.r.t. the trunk input columns [PE | t], accumulated over the skip layer and layer 1.
static int trunk_backward(Ctx& c, const nsff_field_desc* d, const Trunk& t, uint8_t* g_cur, uint8_t* g_next,
                          const Blk* in1T, const Blk* in5T, float* d_in) {
  const int W = 256;
  for (int i = 7; i >= 0; --i) {
    const bool first = i == 0, skip = i == d->skip;
    const int in_k = first ? t.n_in : (skip ? t.n_in + W : W);
    if (!t.gw[i] || !t.gb[i]) return NSFF_EINVAL;
    float* acc;
    if (!first) {
      TRY(wgrad_img(c, g_cur, four(t.H[i - 1]), &acc));
      const int h0 = skip ? t.n_in : 0;
      TRY(finish_cols(c, acc, t.gw[i], in_k, {h0, h0 + 64, h0 + 128, h0 + 192}, {64, 64, 64, 64}));
    }
    if (first || skip) {
      TRY(wgrad_img(c, g_cur, t.in, &acc));
      if (t.trans) TRY(finish_cols(c, acc, t.gw[i], in_k, {0, d->in_xyz}, {d->in_xyz, d->in_t}));
      else TRY(finish_cols(c, acc, t.gw[i], in_k, {0}, {d->in_xyz}));
      if (d_in) TRY(gemm_rows(c, four(g_cur), first ? *in1T : *in5T, d_in, 128, kTinCol0 + d->in_t, first ? 1 : 0));
    }
    if (!first) {
      if (!t.gb[i - 1]) return NSFF_EINVAL;
      GemmOut o{GE_MASK_BITS, g_next};
      o.mask_bits = t.HB[i - 1];
      o.colsum = t.gb[i - 1];                                   // bias gradient of layer i-1 = column sums of dZ_{i-1}
      TRY(gemm_img(c, four(g_cur), t.hidT[i], o));
      uint8_t* tmp = g_cur; g_cur = g_next; g_next = tmp;
    }
  }
  return NSFF_OK;
}

// ---- chained backward ------------------------------------------------------------------------------------------------
// The dgrad GEMMs of all layers as ONE launch of the forward's tile machinery (nsff_field_tc.cu, kBwd): per tile, dZ of a
// layer stays in tensor memory as the A operand of the next (earlier) layer; the launch writes the dZ image of every
// layer and the fp32 trunk-input gradient rows.  The weight-gradient GEMMs then run over the saved images - their idle
// epilogue warps also sum the columns of dZ (bias gradients).  Needs the forward's save buffer; appearance codes use the
// layer-wise path.
static int chain_enabled() {
  const char* e = getenv("NSFF_BWD_CHAIN");
  return e ? atoi(e) != 0 : 1;
}

static int backward_chain(Ctx& c, const nsff_field_desc* d, const BwdPack& L, Scratch& s, const SaveLayout& SL,
                          const uint8_t* saved, const nsff_points* pts, bool want_s, bool want_t, bool want_fw, bool want_bw,
                          const float* out, const float* d_out, int32_t out_ld, const nsff_field_grads* g, float* d_pos,
                          float* d_t) {
  using namespace nsff::tc;
  const int W = 256;
  const int64_t B = c.B, R = B / pts->S;
  cudaStream_t st = c.st;
  const Src pe = chunk_of(s.PE, 1, 0, 4), tcode = chunk_of(s.T, 1, 0, 3), dirc = chunk_of(s.DIR, 1, 0, 2);
  const Src dps = chunk_of(s.DP, 1, 0, 1), dpt = chunk_of(s.DP2, 1, 0, 1);
  auto heads = [&](const HeadSpec& H, uint8_t* img) {
    const int blocks = c.n_tiles < 4 * sm_count() ? c.n_tiles : 4 * sm_count();
    heads_dpre_kernel<<<(unsigned)blocks, 128, 0, st>>>(d_out, out, out_ld, H, pts->zs, d->flow_scale, B, img, kChunkBytes,
                                                        c.n_tiles);
    return after_launch();
  };
  auto off_of = [&](const uint8_t* p) { return (uint64_t)(p - s.base); };
  std::vector<ChainLayer> chain;
  auto hidden = [&](const Blk& blk, std::initializer_list<uint8_t> chunks, std::initializer_list<uint8_t> blks, bool mask,
                    uint64_t bits_off, uint8_t* save) {
    ChainLayer l;
    memset(&l, 0, sizeof(l));
    l.w_off = blk.off; l.block_rows = 256; l.row0 = 0; l.n = 256;
    l.nchunks = (int)chunks.size();
    int k = 0;
    for (uint8_t ch : chunks) l.chunk[k++] = ch;
    k = 0;
    for (uint8_t b : blks) l.blk[k++] = b;
    l.epi = mask ? CHAIN_MASK : CHAIN_LIN;
    l.bits_off = bits_off;
    l.save_off = off_of(save);
    chain.push_back(l);
  };
  auto rows = [&](const Blk& blk, int half, int beta) {
    ChainLayer l;
    memset(&l, 0, sizeof(l));
    l.w_off = blk.off; l.block_rows = 128; l.row0 = 64 * half; l.n = 64; l.nchunks = 4;
    for (int k = 0; k < 4; ++k) l.chunk[k] = (uint8_t)(CC_H3 - k), l.blk[k] = (uint8_t)(3 - k);    // last chunk first (drain proof)
    l.epi = CHAIN_ROWS;
    l.save_off = l.bits_off = ~0ull;
    l.out_col0 = 64 * half;
    l.beta = beta;
    chain.push_back(l);
  };
  const std::initializer_list<uint8_t> H4 = {CC_H0, CC_H1, CC_H2, CC_H3}, B4 = {0, 1, 2, 3};

  int col = 0;
  HeadSpec Hs, Ht;
  float* t_rows[16] = {nullptr};
  if (want_s) {
    memset(&Hs, 0, sizeof(Hs));
    for (int i = 0; i < 16; ++i) Hs.src[i] = -1;
    for (int i = 0; i < 3; ++i) Hs.src[i] = (int8_t)(col + i), Hs.act[i] = 1, Hs.db[i] = g->static_rgb_b ? g->static_rgb_b + i : nullptr;
    Hs.src[3] = (int8_t)(col + 3), Hs.act[3] = 0, Hs.db[3] = g->static_sigma_b;
    TRY(heads(Hs, s.DP));
    if (!g->static_rgb_w || !g->static_sigma_w || !g->static_final_b || !g->static_final_w) return NSFF_EINVAL;
    if (d->use_viewdir) {
      if (!g->static_dir_w || !g->static_dir_b) return NSFF_EINVAL;
      hidden(L.s_rgbT, {CC_DP_STATIC}, {0}, true, SL.s_dirb, s.GS_DIR);                               // dZ_dir
      hidden(L.s_dirT, H4, B4, false, ~0ull, s.GS_FIN);                                                // dFEAT
    } else {
      hidden(L.s_rgbT, {CC_DP_STATIC}, {0}, false, ~0ull, s.GS_FIN);
    }
    hidden(L.s_finT, {CC_DP_STATIC, CC_H0, CC_H1, CC_H2, CC_H3}, {4, 0, 1, 2, 3}, true, SL.s_hb[7], s.GS[7]);   // dZ_8 (+ sigma head)
    for (int i = 7; i >= 1; --i) hidden(L.s_hidT[i], H4, B4, true, SL.s_hb[i - 1], s.GS[i - 1]);
    col += 4;
  }
  if (want_t) {
    memset(&Ht, 0, sizeof(Ht));
    for (int i = 0; i < 16; ++i) Ht.src[i] = -1;
    for (int i = 0; i < 3; ++i) Ht.src[i] = (int8_t)(col + i), Ht.act[i] = 1, Ht.db[i] = g->trans_rgb_b ? g->trans_rgb_b + i : nullptr;
    Ht.src[3] = (int8_t)(col + 3), Ht.act[3] = 0, Ht.db[3] = g->trans_sigma_b;
    if (!g->trans_rgb_w || !g->trans_sigma_w || !g->trans_final_b || !g->trans_final_w) return NSFF_EINVAL;
    t_rows[0] = g->trans_rgb_w, t_rows[1] = g->trans_rgb_w + W, t_rows[2] = g->trans_rgb_w + 2 * W, t_rows[3] = g->trans_sigma_w;
    int cc = col + 4;
    if (want_fw) {
      if (!g->trans_fw_w) return NSFF_EINVAL;
      for (int i = 0; i < 3; ++i) {
        Ht.src[4 + i] = (int8_t)(cc + i), Ht.act[4 + i] = 2, Ht.db[4 + i] = g->trans_fw_b ? g->trans_fw_b + i : nullptr;
        t_rows[4 + i] = g->trans_fw_w + i * W;
      }
      cc += 3;
    }
    if (want_bw) {
      if (!g->trans_bw_w) return NSFF_EINVAL;
      for (int i = 0; i < 3; ++i) {
        Ht.src[7 + i] = (int8_t)(cc + i), Ht.act[7 + i] = 2, Ht.db[7 + i] = g->trans_bw_b ? g->trans_bw_b + i : nullptr;
        t_rows[7 + i] = g->trans_bw_w + i * W;
      }
      cc += 3;
    }
    TRY(heads(Ht, s.DP2));
    hidden(L.t_headT, {CC_DP_TRANS}, {0}, false, ~0ull, s.GT_FIN);                                     // dTFEAT
    hidden(L.t_finT, H4, B4, true, SL.t_hb[7], s.GT[7]);                                               // dZ_8
    const bool rows_pe = d_pos != nullptr, rows_t = d_t != nullptr && d->in_t > 0;
    for (int i = 7; i >= 1; --i) {
      if (i == d->skip) {                          // chain state = dZ of the skip layer: its trunk-input columns
        if (rows_pe) rows(L.t_in5T, 0, 0);
        if (rows_t) rows(L.t_in5T, 1, 0);
      }
      hidden(L.t_hidT[i], H4, B4, true, SL.t_hb[i - 1], s.GT[i - 1]);
    }
    if (rows_pe) rows(L.t_in1T, 0, 1);
    if (rows_t) rows(L.t_in1T, 1, 1);
  }
  TRY(launch_backward_chain(chain.data(), (int)chain.size(), c.pack, want_s ? s.DP : nullptr, want_t ? s.DP2 : nullptr, B,
                            saved, s.base, s.DIN, 128, st));

  // ---- weight gradients over the saved images (any order: collected and launched kWgMaxJobs at a time) ----
  c.batch_wgrads = true;
  float* acc;
  auto trunk_wgrads = [&](bool trans, uint8_t* const* G, uint8_t* const* H, float* const* gw, float* const* gb) -> int {
    const int n_in = trans ? d->in_xyz + d->in_t : d->in_xyz;
    for (int i = 7; i >= 0; --i) {
      const bool first = i == 0, skip = i == d->skip;
      const int in_k = first ? n_in : (skip ? n_in + W : W);
      if (!gw[i] || !gb[i]) return NSFF_EINVAL;
      if (!first) {
        TRY(wgrad_img(c, G[i], four(H[i - 1]), &acc, gb[i]));
        const int h0 = skip ? n_in : 0;
        TRY(finish_cols(c, acc, gw[i], in_k, {h0, h0 + 64, h0 + 128, h0 + 192}, {64, 64, 64, 64}));
      }
      if (first || skip) {
        if (trans) {
          TRY(wgrad_img(c, G[i], {pe, tcode}, &acc, first ? gb[i] : nullptr));
          TRY(finish_cols(c, acc, gw[i], in_k, {0, d->in_xyz}, {d->in_xyz, d->in_t}));
        } else {
          TRY(wgrad_img(c, G[i], {pe}, &acc, first ? gb[i] : nullptr));
          TRY(finish_cols(c, acc, gw[i], in_k, {0}, {d->in_xyz}));
        }
      }
    }
    return NSFF_OK;
  };
  if (want_s) {
    const uint8_t* rgb_in = d->use_viewdir ? s.SDIRO : s.SFEAT;
    float* rows1[16] = {g->static_rgb_w, g->static_rgb_w + W, g->static_rgb_w + 2 * W};
    TRY(wgrad_img(c, rgb_in, {dps}, &acc));
    TRY(finish_rows(c, acc, rows1));
    float* rows2[16] = {nullptr, nullptr, nullptr, g->static_sigma_w};
    TRY(wgrad_img(c, s.SH[7], {dps}, &acc));
    TRY(finish_rows(c, acc, rows2));
    if (d->use_viewdir) {
      const int kd = W + d->in_dir;
      TRY(wgrad_img(c, s.GS_DIR, four(s.SFEAT), &acc, g->static_dir_b));
      TRY(finish_cols(c, acc, g->static_dir_w, kd, {0, 64, 128, 192}, {64, 64, 64, 64}));
      TRY(wgrad_img(c, s.GS_DIR, {dirc}, &acc));
      TRY(finish_cols(c, acc, g->static_dir_w, kd, {W}, {d->in_dir}));
    }
    TRY(wgrad_img(c, s.GS_FIN, four(s.SH[7]), &acc, g->static_final_b));
    TRY(finish_cols(c, acc, g->static_final_w, W, {0, 64, 128, 192}, {64, 64, 64, 64}));
    TRY(trunk_wgrads(false, s.GS, s.SH, g->static_w, g->static_b));
  }
  if (want_t) {
    TRY(wgrad_img(c, s.TFEAT, {dpt}, &acc));
    TRY(finish_rows(c, acc, t_rows));
    TRY(wgrad_img(c, s.GT_FIN, four(s.TH[7]), &acc, g->trans_final_b));
    TRY(finish_cols(c, acc, g->trans_final_w, W, {0, 64, 128, 192}, {64, 64, 64, 64}));
    TRY(trunk_wgrads(true, s.GT, s.TH, g->trans_w, g->trans_b));
    if (d_t && d->in_t > 0) TRY(launch_ray_reduce(s.DIN, 128, kTinCol0, d->in_t, pts->S, R, d_t, st));
    if (d_pos) TRY(launch_pe_bwd(s.DIN, 128, pts->xyz, (d->in_xyz - 3) / 6, d_pos, B, st));
  }
  return flush_finish(c);
}

}  // namespace bwd
}  // namespace nsff

using namespace nsff;
using namespace nsff::bwd;

extern "C" {

size_t nsff_field_bwd_tc_pack_bytes(const nsff_field_desc* d) {
  if (!d || supported(d) != NSFF_OK) return 0;
  BwdPack L;
  PackBuilder pb;
  build_pack(d, nullptr, L, pb);
  return L.bytes;
}

int nsff_field_bwd_tc_pack(const nsff_field_desc* d, const nsff_field_params* p, void* pack, void* stream) {
  if (!d || !p || !pack) return NSFF_EINVAL;
  TRY(supported(d));
  cudaStream_t st = (cudaStream_t)stream;
  BwdPack L;
  PackBuilder pb;
  pb.dst = reinterpret_cast<uint8_t*>(pack);
  build_pack(d, p, L, pb);
  NSFF_CUDA_TRY(cudaMemsetAsync(pack, 0, L.bytes, st));
  return flush_pack_jobs(pb, st);
}

size_t nsff_field_bwd_tc_scratch_bytes(const nsff_field_desc* d, int64_t B, int with_saved) {
  if (!d || B < 0 || supported(d) != NSFF_OK) return 0;
  return carve(d, B, nullptr, !with_saved).total;
}

int nsff_field_backward_tc(const nsff_field_desc* d, const nsff_field_params* p, const void* pack,
                           const nsff_points* pts, int out_mask,
                           const float* out, const float* d_out, int32_t out_ld, const nsff_field_grads* g,
                           float* d_pos, float* d_t, float* d_a, const void* saved, void* scratch, void* stream) {
  if (!d || !p || !pack || !pts || !out || !d_out || !g || !scratch) return NSFF_EINVAL;
  TRY(supported(d));
  if (out_mask & NSFF_OUT_SIGMA_ONLY) return NSFF_EUNSUPPORTED;
  if (pts->t_index || pts->a_index || pts->offset) return NSFF_EINVAL;   // per-ray code rows, final positions
  const bool want_s = out_mask & NSFF_OUT_STATIC, want_t = out_mask & NSFF_OUT_TRANSIENT;
  const bool want_fw = out_mask & NSFF_OUT_FLOW_FW, want_bw = out_mask & NSFF_OUT_FLOW_BW;
  if (!want_s && !want_t) return NSFF_EINVAL;
  if (want_t && !d->has_transient) return NSFF_EINVAL;
  const int64_t B = pts->B;
  if (B <= 0 || pts->S < 1 || B % pts->S) return B == 0 ? NSFF_OK : NSFF_EINVAL;
  if (want_t && d->in_t > 0 && !pts->t_table) return NSFF_EINVAL;
  TRY(check_arch());
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t R = B / pts->S;
  const int W = 256;
  BwdPack L;
  {
    PackBuilder pb;
    build_pack(d, nullptr, L, pb);
  }
  Scratch s = carve(d, B, reinterpret_cast<uint8_t*>(scratch), saved == nullptr);
  if (saved) {
    // activations written by nsff_field_forward_save for the same (desc, B, out_mask): nothing is recomputed
    if ((uintptr_t)saved & 15u) return NSFF_EINVAL;
    const SaveLayout SL = save_layout(d, B, out_mask);
    uint8_t* sv = const_cast<uint8_t*>(reinterpret_cast<const uint8_t*>(saved));
    auto at = [&](uint64_t off) { return off == kNoSave ? nullptr : sv + off; };
    for (int i = 0; i < 8; ++i) {
      s.SH[i] = at(SL.s_h[i]); s.SHB[i] = reinterpret_cast<uint32_t*>(at(SL.s_hb[i]));
      s.TH[i] = at(SL.t_h[i]); s.THB[i] = reinterpret_cast<uint32_t*>(at(SL.t_hb[i]));
    }
    s.SFEAT = at(SL.s_feat); s.SDIRO = at(SL.s_dir); s.SDIROB = reinterpret_cast<uint32_t*>(at(SL.s_dirb));
    s.TFEAT = at(SL.t_feat);
  }
  const bool recompute = saved == nullptr;
  Ctx c{st, (int)cdiv64(B, 128), B, reinterpret_cast<const uint8_t*>(pack), s.ACC, 0, s.acc_bytes / sizeof(float), {}};
  NSFF_CUDA_TRY(cudaMemsetAsync(s.ACC, 0, s.acc_bytes, st));
  if (d_pos) NSFF_CUDA_TRY(cudaMemsetAsync(d_pos, 0, (size_t)B * 3 * sizeof(float), st));
  if (d_t) NSFF_CUDA_TRY(cudaMemsetAsync(d_t, 0, (size_t)R * d->in_t * sizeof(float), st));
  if (d_a) NSFF_CUDA_TRY(cudaMemsetAsync(d_a, 0, (size_t)R * d->in_a * sizeof(float), st));

  // ---- inputs: the embedded row of every point as bf16 chunk images ----
  {
    EmbedImgArgs E;
    memset(&E, 0, sizeof(E));
    E.xyz = pts->xyz; E.dirs = pts->dirs; E.t_table = pts->t_table; E.a_table = pts->a_table;
    E.B = B; E.S = pts->S; E.n_tiles = c.n_tiles;
    E.in_xyz = d->in_xyz; E.in_dir = d->in_dir; E.in_t = d->in_t; E.in_a = d->in_a;
    E.img[0] = s.PE;
    E.img[1] = want_t ? s.T : nullptr;
    E.img[2] = (want_s && d->use_viewdir) ? s.DIR : nullptr;
    E.img[3] = (want_s && d->in_a > 0) ? s.A : nullptr;
    if (E.img[2] && !pts->dirs) return NSFF_EINVAL;
    embed_img_kernel<<<(unsigned)cdiv64((int64_t)c.n_tiles * 128 * 4, 128), 128, 0, st>>>(E);
    TRY(after_launch());
  }
  if (saved && d->in_a == 0 && chain_enabled())
    return backward_chain(c, d, L, s, save_layout(d, B, out_mask), reinterpret_cast<const uint8_t*>(saved), pts, want_s,
                          want_t, want_fw, want_bw, out, d_out, out_ld, g, d_pos, d_t);
  const Src pe = chunk_of(s.PE, 1, 0, 4), tc = chunk_of(s.T, 1, 0, 3), dirc = chunk_of(s.DIR, 1, 0, 2),
            ac = chunk_of(s.A, 1, 0, 3), dp = chunk_of(s.DP, 1, 0, 1);
  int col = 0;
  auto heads = [&](const HeadSpec& H) {
    const int blocks = c.n_tiles < 4 * sm_count() ? c.n_tiles : 4 * sm_count();
    heads_dpre_kernel<<<(unsigned)blocks, 128, 0, st>>>(d_out, out, out_ld, H, pts->zs, d->flow_scale, B, s.DP,
                                                        kChunkBytes, c.n_tiles);
    return after_launch();
  };
  float* acc;

  if (want_s) {
    Trunk t{false, L.s_fwd, L.s_hidT, p->static_b, g->static_w, g->static_b, s.SH, s.SHB, {pe}, d->in_xyz};
    // ---- recompute, keeping every layer (unless the forward saved them) ----
    if (recompute) {
      TRY(trunk_forward(c, d, t));
      GemmOut o{GE_LIN_IMG, s.SFEAT};
      o.bias = p->static_final_b;
      TRY(gemm_img(c, four(s.SH[7]), L.s_fin, o));
    }
    const uint8_t* rgb_in = s.SFEAT;
    if (d->use_viewdir) {
      if (recompute) {
        std::vector<Src> a = four(s.SFEAT);
        a.push_back(dirc);
        if (d->in_a > 0) a.push_back(ac);
        GemmOut o{GE_RELU_IMG, s.SDIRO};
        o.bias = p->static_dir_b;
        o.relu_bits = s.SDIROB;
        TRY(gemm_img(c, a, L.s_dir, o));
      }
      rgb_in = s.SDIRO;
    }
    // ---- heads: image columns 0..2 = rgb (sigmoid), 3 = sigma ----
    HeadSpec H;
    memset(&H, 0, sizeof(H));
    for (int i = 0; i < 16; ++i) H.src[i] = -1;
    for (int i = 0; i < 3; ++i) H.src[i] = (int8_t)(col + i), H.act[i] = 1, H.db[i] = g->static_rgb_b ? g->static_rgb_b + i : nullptr;
    H.src[3] = (int8_t)(col + 3), H.act[3] = 0, H.db[3] = g->static_sigma_b;
    TRY(heads(H));
    if (!g->static_rgb_w || !g->static_sigma_w) return NSFF_EINVAL;
    {
      float* rows[16] = {g->static_rgb_w, g->static_rgb_w + W, g->static_rgb_w + 2 * W};
      TRY(wgrad_img(c, rgb_in, {dp}, &acc));
      TRY(finish_rows(c, acc, rows));
      float* rows2[16] = {nullptr, nullptr, nullptr, g->static_sigma_w};
      TRY(wgrad_img(c, s.SH[7], {dp}, &acc));
      TRY(finish_rows(c, acc, rows2));
    }
    uint8_t *g_fin, *other;
    if (d->use_viewdir) {
      const int kd = W + d->in_dir + d->in_a;
      if (!g->static_dir_w || !g->static_dir_b) return NSFF_EINVAL;
      {
        GemmOut o{GE_MASK_BITS, s.G0};                                                     // dZ_dir
        o.mask_bits = s.SDIROB;
        o.colsum = g->static_dir_b;
        TRY(gemm_img(c, {dp}, L.s_rgbT, o));
      }
      TRY(wgrad_img(c, s.G0, four(s.SFEAT), &acc));
      TRY(finish_cols(c, acc, g->static_dir_w, kd, {0, 64, 128, 192}, {64, 64, 64, 64}));
      if (d->in_a > 0) {
        TRY(wgrad_img(c, s.G0, {dirc, ac}, &acc));
        TRY(finish_cols(c, acc, g->static_dir_w, kd, {W, W + d->in_dir}, {d->in_dir, d->in_a}));
        if (d_a) {
          TRY(gemm_rows(c, four(s.G0), L.s_dirxT, s.DIN, 128, d->in_dir + d->in_a, 0));
          TRY(launch_ray_reduce(s.DIN, 128, d->in_dir, d->in_a, pts->S, R, d_a, st));
        }
      } else {
        TRY(wgrad_img(c, s.G0, {dirc}, &acc));
        TRY(finish_cols(c, acc, g->static_dir_w, kd, {W}, {d->in_dir}));
      }
      {
        GemmOut o{GE_LIN_IMG, s.G1};                                                       // dFEAT = dZ_final
        o.colsum = g->static_final_b;
        TRY(gemm_img(c, four(s.G0), L.s_dirT, o));
      }
      g_fin = s.G1, other = s.G0;
    } else {
      GemmOut o{GE_LIN_IMG, s.G0};
      o.colsum = g->static_final_b;
      TRY(gemm_img(c, {dp}, L.s_rgbT, o));
      g_fin = s.G0, other = s.G1;
    }
    if (!g->static_final_b || !g->static_final_w || !g->static_b[7]) return NSFF_EINVAL;
    TRY(wgrad_img(c, g_fin, four(s.SH[7]), &acc));
    TRY(finish_cols(c, acc, g->static_final_w, W, {0, 64, 128, 192}, {64, 64, 64, 64}));
    {
      std::vector<Src> a = four(g_fin);
      a.push_back(dp);
      GemmOut o{GE_MASK_BITS, other};                                                      // dZ_8 (incl. sigma head)
      o.mask_bits = s.SHB[7];
      o.colsum = g->static_b[7];
      TRY(gemm_img(c, a, L.s_finT, o));
    }
    // static positions never carry gradient (the warped passes evaluate the dynamic branch only)
    TRY(trunk_backward(c, d, t, other, g_fin, nullptr, nullptr, nullptr));
    col += 4;
  }
  if (want_t) {
    Trunk t{true, L.t_fwd, L.t_hidT, p->trans_b, g->trans_w, g->trans_b, s.TH, s.THB, {pe, tc}, d->in_xyz + d->in_t};
    if (recompute) {
      TRY(trunk_forward(c, d, t));
      GemmOut o{GE_LIN_IMG, s.TFEAT};
      o.bias = p->trans_final_b;
      TRY(gemm_img(c, four(s.TH[7]), L.t_fin, o));
    }
    // ---- heads: image columns 0..2 rgb, 3 sigma, 4..6 flow fw, 7..9 flow bw (fixed, whatever the mask) ----
    HeadSpec H;
    memset(&H, 0, sizeof(H));
    for (int i = 0; i < 16; ++i) H.src[i] = -1;
    for (int i = 0; i < 3; ++i) H.src[i] = (int8_t)(col + i), H.act[i] = 1, H.db[i] = g->trans_rgb_b ? g->trans_rgb_b + i : nullptr;
    H.src[3] = (int8_t)(col + 3), H.act[3] = 0, H.db[3] = g->trans_sigma_b;
    float* rows[16] = {g->trans_rgb_w, g->trans_rgb_w + W, g->trans_rgb_w + 2 * W, g->trans_sigma_w};
    if (!g->trans_rgb_w || !g->trans_sigma_w) return NSFF_EINVAL;
    int cc = col + 4;
    if (want_fw) {
      if (!g->trans_fw_w) return NSFF_EINVAL;
      for (int i = 0; i < 3; ++i) {
        H.src[4 + i] = (int8_t)(cc + i), H.act[4 + i] = 2, H.db[4 + i] = g->trans_fw_b ? g->trans_fw_b + i : nullptr;
        rows[4 + i] = g->trans_fw_w + i * W;
      }
      cc += 3;
    }
    if (want_bw) {
      if (!g->trans_bw_w) return NSFF_EINVAL;
      for (int i = 0; i < 3; ++i) {
        H.src[7 + i] = (int8_t)(cc + i), H.act[7 + i] = 2, H.db[7 + i] = g->trans_bw_b ? g->trans_bw_b + i : nullptr;
        rows[7 + i] = g->trans_bw_w + i * W;
      }
      cc += 3;
    }
    TRY(heads(H));
    TRY(wgrad_img(c, s.TFEAT, {dp}, &acc));
    TRY(finish_rows(c, acc, rows));
    if (!g->trans_final_b || !g->trans_final_w || !g->trans_b[7]) return NSFF_EINVAL;
    {
      GemmOut o{GE_LIN_IMG, s.G0};                                                         // dTFEAT = dZ_final
      o.colsum = g->trans_final_b;
      TRY(gemm_img(c, {dp}, L.t_headT, o));
    }
    TRY(wgrad_img(c, s.G0, four(s.TH[7]), &acc));
    TRY(finish_cols(c, acc, g->trans_final_w, W, {0, 64, 128, 192}, {64, 64, 64, 64}));
    {
      GemmOut o{GE_MASK_BITS, s.G1};                                                       // dZ_8
      o.mask_bits = s.THB[7];
      o.colsum = g->trans_b[7];
      TRY(gemm_img(c, four(s.G0), L.t_finT, o));
    }
    float* d_in = (d_pos || d_t) ? s.DIN : nullptr;
    TRY(trunk_backward(c, d, t, s.G1, s.G0, &L.t_in1T, &L.t_in5T, d_in));
    if (d_t && d->in_t > 0) TRY(launch_ray_reduce(d_in, 128, kTinCol0, d->in_t, pts->S, R, d_t, st));
    if (d_pos) TRY(launch_pe_bwd(d_in, 128, pts->xyz, (d->in_xyz - 3) / 6, d_pos, B, st));
  }
  return flush_finish(c);
}

// ---- generic entry points (unit tests; hosts that keep their own bf16 tile images) -------------------------------
size_t nsff_img_bytes(int64_t n_rows, int n_chunks) {
  if (n_rows < 0 || n_chunks < 1) return 0;
  return (size_t)cdiv64(n_rows, 128) * (size_t)n_chunks * kChunkBytes;
}

int nsff_img_from_rows(const float* src, int32_t ld, int32_t col0, int32_t ncols, int64_t n_rows, void* img,
                       int32_t img_chunks, int32_t chunk, void* stream) {
  if (!src || !img || ncols < 0 || ncols > 64 || chunk < 0 || chunk >= img_chunks || n_rows < 0) return NSFF_EINVAL;
  const int n_tiles = (int)cdiv64(n_rows, 128);
  if (!n_tiles) return NSFF_OK;
  const int64_t n = (int64_t)n_tiles * 128 * 8;
  rows_to_img_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, (cudaStream_t)stream>>>(
      src, ld, col0, ncols, n_rows, reinterpret_cast<uint8_t*>(img) + (size_t)chunk * kChunkBytes,
      (uint32_t)img_chunks * kChunkBytes, n_tiles);
  return after_launch();
}

int nsff_img_to_rows(const void* img, int32_t img_chunks, int32_t chunk, float* dst, int32_t ld, int32_t col0,
                     int32_t ncols, int64_t n_rows, void* stream) {
  if (!dst || !img || ncols < 0 || ncols > 64 || chunk < 0 || chunk >= img_chunks || n_rows < 0) return NSFF_EINVAL;
  if (!n_rows) return NSFF_OK;
  img_to_rows_kernel<<<(unsigned)cdiv64(n_rows * 64, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint8_t*>(img) + (size_t)chunk * kChunkBytes, (uint32_t)img_chunks * kChunkBytes, dst, ld,
      col0, ncols, n_rows);
  return after_launch();
}

int nsff_img_pack_weight(const float* src, int32_t ld, int32_t transpose, int32_t n0, int32_t k0, int32_t n_valid,
                         int32_t k_valid, int32_t N, void* block, void* stream) {
  if (!src || !block || (N != 256 && N != 128 && N != 64) || n_valid < 0 || n_valid > N || k_valid < 0 || k_valid > 64)
    return NSFF_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  NSFF_CUDA_TRY(cudaMemsetAsync(block, 0, (size_t)N * 128, st));
  if (!n_valid || !k_valid) return NSFF_OK;
  PackJobs J;
  memset(&J, 0, sizeof(J));
  J.job[0] = PackJob{src, reinterpret_cast<uint8_t*>(block), ld, transpose ? 1 : 0, n0, k0, n_valid, k_valid, 0, 0};
  pack_blocks_kernel<<<1, 256, 0, st>>>(J);
  return after_launch();
}

int nsff_img_gemm(const nsff_img_gemm_args* a, void* stream) {
  if (!a || !a->a_img || !a->b_blocks || a->a_chunks < 1 || a->a_chunks > kMaxK || a->n_rows < 0) return NSFF_EINVAL;
  if (a->epi < 0 || a->epi > 4) return NSFF_EINVAL;
  if (a->epi == GE_F32_ROWS ? !a->out_rows : !a->out_img) return NSFF_EINVAL;
  if (a->epi == GE_MASK_IMG && !a->mask_img) return NSFF_EINVAL;
  if (a->epi == GE_MASK_BITS && !a->mask_bits) return NSFF_EINVAL;
  TRY(check_arch());
  GemmParams P;
  memset(&P, 0, sizeof(P));
  for (int i = 0; i < a->a_chunks; ++i) {
    P.a_base[i] = reinterpret_cast<const uint8_t*>(a->a_img) + (size_t)i * kChunkBytes;
    P.a_stride[i] = (uint32_t)a->a_chunks * kChunkBytes;
    P.kcnt[i] = 4;
  }
  P.n_tiles = (int)cdiv64(a->n_rows, 128); P.n_chunks = a->a_chunks; P.N = a->N; P.epi = a->epi;
  P.b_blocks = reinterpret_cast<const uint8_t*>(a->b_blocks); P.bias = a->bias;
  P.out_img = reinterpret_cast<uint8_t*>(a->out_img); P.out_stride = (uint32_t)(a->N / 64) * kChunkBytes;
  P.mask_img = reinterpret_cast<const uint8_t*>(a->mask_img); P.mask_stride = P.out_stride;
  P.out_f32 = a->out_rows; P.ld = a->ld; P.ncols = a->ncols; P.beta = a->beta; P.n_rows = a->n_rows;
  P.relu_bits_out = a->relu_bits; P.mask_bits = a->mask_bits; P.colsum = a->colsum;
  return launch_gemm(P, (cudaStream_t)stream);
}

int nsff_img_wgrad(const nsff_img_wgrad_args* a, void* stream) {
  if (!a || !a->a_img || !a->b_img || a->b_chunks < 1 || a->b_chunks > 4 || !a->acc_t || a->n_rows < 0) return NSFF_EINVAL;
  TRY(check_arch());
  WgradParams P;
  memset(&P, 0, sizeof(P));
  WgradJob& J = P.job[0];
  J.a_base = reinterpret_cast<const uint8_t*>(a->a_img); J.a_stride = 4 * kChunkBytes;
  for (int j = 0; j < a->b_chunks; ++j) {
    J.b_base[j] = reinterpret_cast<const uint8_t*>(a->b_img) + (size_t)j * kChunkBytes;
    J.b_stride[j] = (uint32_t)a->b_chunks * kChunkBytes;
  }
  J.nb = a->b_chunks; J.acc_t = a->acc_t;
  P.n_jobs = 1; P.n_tiles = (int)cdiv64(a->n_rows, 128);
  return launch_wgrad(P, (cudaStream_t)stream);
}
}
