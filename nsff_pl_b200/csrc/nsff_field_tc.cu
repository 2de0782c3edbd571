// a3/a4/a5/a7/a10: fused positional-encoding + NeRF MLP forward on the 5th-gen tensor cores.
//
//   reference: models/nerf.py:118-213 (NeRF.forward, all call modes) fed by
//              models/rendering.py:113-120,152-175 (PE(xyz) | PE(dir) | a | t input rows)
//
// One persistent CTA PAIR (cluster of 2 = one TPC, cta_group::2) walks pairs of 128-point tiles.  For each tile the
// whole layer program of the requested heads (static trunk -> sigma head -> final -> dir layer -> rgb head, then the
// dynamic trunk -> final -> rgb/sigma/flow heads) runs without the activations ever leaving the SM - and, since
// round 2, without the hidden activation ever touching SHARED memory either:
//
//   TMEM (512 columns / CTA): two 256-column regions X, Y.  A layer accumulates D (128 x 256 fp32) into the region that
//       does not hold its input.  The epilogue drains D with tcgen05.ld, adds the bias, applies ReLU, rounds to bf16 and
//       writes the rows back IN PLACE with tcgen05.st, packed two per column: K-step j of 64-channel chunk c lands at
//       column (c>>1)*H + 16 j + (c&1)*8 of the same region (H = 64, or 128 in split mode) - each warp only overwrites
//       columns it has itself drained.  The next layer reads that image as its A operand straight from tensor memory
//       (tcgen05.mma, .ts form).  Compared with round 1 (hidden chunks in swizzled shared memory) this removes the
//       st.shared, the fence.proxy.async (measured 140 cycles per 64-column slice, profiles/r02a_ubench_tmem.txt), the
//       shared-memory A-operand fetch of every MMA and 64 KB of shared memory.
//       Heads (N = 16) accumulate into the free upper columns of the region that holds their input; their K-chunks run
//       last-chunk-first so that "chunk 3 published" proves the whole region has been drained.
//   split mode (NSFF_TC_SPLIT=1): every 256-wide layer is issued as two N = 128 halves with their own accumulator
//       columns and barriers, so the epilogue of half 0 runs under the MMAs of half 1 and the first two K-chunks of the
//       next layer are ready when half 1 completes: no accumulator-drain bubble between layers.
//   smem  A operand : only the "extra" chunks written once per tile by the prologue: PE(xyz) (63 -> 64 cols), time code
//                     (48), PE(dir) (27 -> 32) [+ appearance code (48)]  (128 rows x 64 bf16, 128B-swizzled, K-major)
//   smem  B operand : 10-stage ring of weight pieces ((N/2) x 64 bf16 per CTA) streamed from the L2-resident field pack
//                     with cp.async.bulk + mbarrier complete_tx
//   warps 0-15      : prologue (encoding) + epilogue (4 TMEM lane quadrants x 4 column parts)
//   warp  16        : weight + bias producer (one elected lane)
//   warp  17        : leader CTA: tcgen05.mma issuer (one elected lane walks a flat, host-built list of K-chunk ops);
//                     peer CTA: relays "my weight half has landed" to the leader's barriers
//
// Ordering notes.  (1) tcgen05.st -> tcgen05.wait::st -> tcgen05.fence::before_thread_sync -> mbarrier arrive |
// mbarrier wait -> tcgen05.fence::after_thread_sync -> tcgen05.mma makes the written rows visible to the MMA.
// (2) A region is overwritten by the accumulator of the layer after next.  Its first MMA may be issued while MMAs of
// the previous layer that still READ the region are in flight (extras-first K order, split mode): tcgen05.mma
// instructions of one issuing thread execute in order and read their operands before a later instruction writes.
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <initializer_list>

#include "nsff_common.cuh"
#include "nsff_field_save.cuh"
#include "nsff_sm100.cuh"

namespace nsff {
namespace tc {

using namespace sm100;

// ---------------------------------------------------------------------------------------------
// pack layout (shared by nsff_field_pack_bytes / nsff_field_pack / the forward launcher)
// ---------------------------------------------------------------------------------------------
enum ChunkId : uint8_t { CH_H0 = 0, CH_H1, CH_H2, CH_H3, CH_PE, CH_T, CH_DIR, CH_A, CH_NONE = 255 };
enum LayerId {
  L_S1 = 0, L_S2, L_S3, L_S4, L_S5, L_S6, L_S7, L_S8, L_SSIG, L_SFIN, L_SDIR, L_SRGB,
  L_T1, L_T2, L_T3, L_T4, L_T5, L_T6, L_T7, L_T8, L_TFIN, L_THEAD, L_COUNT
};
enum Epi : uint8_t { EPI_RELU_HID = 0, EPI_LIN_HID = 1, EPI_HEAD = 2 };
enum HeadAct : uint8_t { HA_NONE = 0, HA_SIGMOID = 1, HA_FLOW = 2 };

constexpr int kMaxChunks = 6;

struct PackLayer {
  bool present;
  uint32_t w_off;      // byte offset of the first weight block
  uint32_t bias_off;   // float index into the bias table
  int n;               // 256 or 16
  int nchunks;
  uint8_t chunk[kMaxChunks];
  uint8_t kcnt[kMaxChunks];   // number of K=16 MMAs issued from this chunk
};

struct PackLayout {
  PackLayer layer[L_COUNT];
  uint32_t bias_base;   // byte offset of the fp32 bias table
  uint32_t n_bias;      // floats
  size_t bytes;
};

static int build_layout(const nsff_field_desc& d, PackLayout& L) {
  if (d.W != 256 || d.D != 8 || d.skip != 4 || d.in_xyz != 63) return NSFF_EUNSUPPORTED;
  if (d.use_viewdir && d.in_dir != 27) return NSFF_EUNSUPPORTED;
  if (d.in_a != 0 && (d.in_a != 48 || !d.use_viewdir)) return NSFF_EUNSUPPORTED;   // NeRF-W appearance code: 48 wide
  if (d.has_transient && d.in_t != 48) return NSFF_EUNSUPPORTED;
  memset(&L, 0, sizeof(L));
  uint32_t off = 0, boff = 0;
  auto add = [&](int id, int n, std::initializer_list<uint8_t> chunks) {
    PackLayer& p = L.layer[id];
    p.present = true;
    p.n = n;
    p.nchunks = 0;
    for (uint8_t c : chunks) {
      p.chunk[p.nchunks] = c;
      p.kcnt[p.nchunks] = (c == CH_T || c == CH_A) ? 3 : (c == CH_DIR) ? 2 : 4;
      ++p.nchunks;
    }
    p.w_off = off;
    p.bias_off = boff;
    off += (uint32_t)p.nchunks * (uint32_t)n * 128u;
    boff += (uint32_t)n;
  };
  // extras first so the tensor pipe has work while the previous epilogue is still producing hidden chunk 0;
  // head blocks read the hidden chunks LAST FIRST: waiting for chunk 3 proves that every epilogue warp has drained the
  // whole accumulator, whose upper columns the head's own accumulator reuses (TMEM map in the kernel header)
  add(L_S1, 256, {CH_PE});
  for (int i = L_S2; i <= L_S8; ++i) {
    if (i == L_S5) add(i, 256, {CH_PE, CH_H0, CH_H1, CH_H2, CH_H3});
    else add(i, 256, {CH_H0, CH_H1, CH_H2, CH_H3});
  }
  add(L_SSIG, 16, {CH_H3, CH_H2, CH_H1, CH_H0});
  add(L_SFIN, 256, {CH_H0, CH_H1, CH_H2, CH_H3});
  if (d.use_viewdir) {
    if (d.in_a > 0) add(L_SDIR, 256, {CH_DIR, CH_A, CH_H0, CH_H1, CH_H2, CH_H3});
    else add(L_SDIR, 256, {CH_DIR, CH_H0, CH_H1, CH_H2, CH_H3});
  }
  add(L_SRGB, 16, {CH_H3, CH_H2, CH_H1, CH_H0});
  if (d.has_transient) {
    add(L_T1, 256, {CH_PE, CH_T});
    for (int i = L_T2; i <= L_T8; ++i) {
      if (i == L_T5) add(i, 256, {CH_PE, CH_T, CH_H0, CH_H1, CH_H2, CH_H3});
      else add(i, 256, {CH_H0, CH_H1, CH_H2, CH_H3});
    }
    add(L_TFIN, 256, {CH_H0, CH_H1, CH_H2, CH_H3});
    add(L_THEAD, 16, {CH_H3, CH_H2, CH_H1, CH_H0});
  }
  L.bias_base = off;
  L.n_bias = boff;
  L.bytes = (size_t)off + (size_t)boff * 4;
  return NSFF_OK;
}

// ---------------------------------------------------------------------------------------------
// pack kernel: fp32 (out,in) nn.Linear weights -> bf16 blocks [n rows x 64 cols], 128B-swizzled
// ---------------------------------------------------------------------------------------------
struct PackJob {
  const float* src;   // weight (rows, ld) row-major, or bias (n) when is_bias
  uint32_t dst_off;   // byte offset in the pack (weights) / float index (bias)
  int32_t ld;         // source row stride
  int32_t row0;       // first destination row inside the block
  int32_t rows;       // rows to write
  int32_t col0;       // first source column
  int32_t cols;       // valid source columns (the rest of the 64 are zero)
  int32_t is_bias;
};
constexpr int kMaxPackJobs = 160;
struct PackJobs {
  PackJob job[kMaxPackJobs];
};

__device__ __forceinline__ uint32_t swz_off(int row, int col /*bf16 column 0..63*/) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((col >> 3) ^ (row & 7)) & 7) << 4) + (col & 7) * 2);
}

__global__ void pack_kernel(const PackJobs* __restrict__ jobs_dev, uint8_t* __restrict__ pack, uint32_t bias_base) {
  const PackJob j = jobs_dev->job[blockIdx.x];
  if (j.is_bias) {
    float* b = reinterpret_cast<float*>(pack + bias_base) + j.dst_off;
    for (int i = threadIdx.x; i < j.rows; i += blockDim.x) b[i] = j.src ? j.src[i] : 0.f;
    return;
  }
  for (int e = threadIdx.x; e < j.rows * 64; e += blockDim.x) {
    int r = e >> 6, c = e & 63;
    float v = (c < j.cols) ? j.src[(int64_t)r * j.ld + j.col0 + c] : 0.f;
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    *reinterpret_cast<__nv_bfloat16*>(pack + j.dst_off + swz_off(j.row0 + r, c)) = h;
  }
}

// ---------------------------------------------------------------------------------------------
// forward kernel
// ---------------------------------------------------------------------------------------------
constexpr int kTileM = 128;
constexpr int kEpiWarps = 16;        // 4 TMEM lane quadrants x 4 column parts
constexpr int kThreads = (kEpiWarps + 2) * 32;
constexpr uint32_t kChunkBytes = 128 * 128;          // 128 rows x 64 bf16
constexpr uint32_t kStageBytes = 128 * 128;          // (256 / 2) rows x 64 bf16: this CTA's half of a 256-wide block
constexpr int kBiasSlots = 4;
constexpr uint32_t kSmemBias = kBiasSlots * 1024;    // fp32 bias rows of the sub-steps in flight
constexpr uint32_t kSmemBar = 512;
constexpr int kMaxStages = 10;
constexpr int kMaxSubs = 44;                         // sub-steps per tile (main pass, split: 19 x 2 + 3 heads = 41)
constexpr int kMaxOps = 168;                         // K-chunk ops per tile (main pass, split, appearance: 164)
constexpr int kAccIds = 5;                           // accumulators with their own full/empty barriers: X.h0 X.h1 Y.h0 Y.h1 head
constexpr int kHeadAcc = 4;

// kAppear: one more prologue chunk (the 48-wide NeRF-W appearance code feeding the dir layer), paid for with one
//          ring stage.
template <bool kAppear>
struct Cfg {
  static constexpr int kStages = kAppear ? 9 : 10;
  static constexpr uint32_t kSmemA = (kAppear ? 4 : 3) * kChunkBytes;   // PE + T + DIR (+ A)
  static constexpr uint32_t kSmemB = kStages * kStageBytes;
  static constexpr uint32_t kSmemTotal = kSmemA + kSmemB + kSmemBias + kSmemBar + 1024;   // + alignment slack
};

enum : uint8_t { OP_TS = 1, OP_FIRST = 2, OP_LAST = 4, OP_ADV_HID = 8 };

// One K-chunk (<= 4 tcgen05.mma of K = 16) of one sub-step: what the producer copies and what the issuer issues.
struct ChunkOp {
  uint32_t w_src;      // byte offset in the pack of the leader CTA's rows of this weight piece (peer: + bytes)
  uint16_t a;          // OP_TS: TMEM column of K-step 0 (+16 per K-step); else smem offset / 16 of the extra chunk
  uint16_t d_col;      // TMEM column of the accumulator
  uint16_t bytes;      // weight bytes per CTA = (N / 2) * 128
  uint8_t n16;         // N / 16
  uint8_t kcnt;        // K = 16 steps in this chunk (2..4)
  uint8_t flags;       // OP_*
  uint8_t wait_hid;    // 1 + hidden-chunk barrier to wait for before issuing (0: none; only the first use in a layer)
  uint8_t acc_id;
  uint8_t wait_empty;  // OP_FIRST: bit mask of accumulators that must have been drained
};
static_assert(sizeof(ChunkOp) == 16, "ChunkOp is read with one 16-byte constant load");

// One accumulator's worth of work for the epilogue warps (a whole layer, or one N = 128 half of it in split mode).
struct SubStep {
  uint32_t bias_off;       // float index of the first bias of this sub-step
  uint16_t n;              // 256 / 128 (hidden) or 16 (head)
  uint16_t d_col;          // TMEM column of the accumulator (and of the in-place bf16 image)
  uint8_t acc_id;
  uint8_t epi;
  uint8_t slice0;          // first 64-column slice (= hidden chunk) this sub-step produces: 0 or 2
  uint8_t nslices;         // 4, or 2 in split mode; 0 for heads
  int8_t head_dst[16];
  uint8_t head_act[16];
  uint64_t save_off;       // training mode: byte offset of this layer's activation image in the save buffer (kNoSave: none)
  uint64_t bits_off;       // ... and of its ReLU bit plane
};

struct FwdParams {
  const uint8_t* pack;
  const float* bias;
  nsff_points pts;          // raw-point mode
  const float* x_emb;       // embedded mode (NeRF.forward drop-in); NULL in raw-point mode
  int32_t x_ld, c_dir, c_t; // embedded-row column offsets
  float* out;
  int32_t out_ld;
  int32_t n_subs, n_ops;
  int32_t n_tiles;
  int32_t need_t, need_dir, need_a, c_a;
  int32_t last_extra_sub;   // last sub-step whose K-chunks include PE / time code / PE(dir): once its accumulator is
                            // complete the prologue chunks are dead and the NEXT tile's prologue can be written early
  float flow_scale;
  uint8_t* save_base;       // training mode (kSave kernels): activation save buffer, see nsff_field_save.cuh
  unsigned long long* dbg;  // NSFF_TIMELINE builds only: (tag, clock64) event log of CTA 0
  SubStep subs[kMaxSubs];
  ChunkOp ops[kMaxOps];
};

// Timeline instrumentation (compiled in only with -DNSFF_TIMELINE; see tools/timeline.py).
#ifdef NSFF_TIMELINE
constexpr int kDbgPerRole = 8192;
#define NSFF_TL_DECL(role) unsigned long long* tl_ptr = (P.dbg && blockIdx.x == 0) ? P.dbg + (role) * 2 * kDbgPerRole : nullptr; int tl_n = 0
#define NSFF_TL(tag)                                                          \
  do {                                                                        \
    if (tl_ptr && tl_n < kDbgPerRole && lane == 0) {                          \
      tl_ptr[2 * tl_n] = (unsigned long long)(tag);                           \
      tl_ptr[2 * tl_n + 1] = (unsigned long long)clock64();                   \
      ++tl_n;                                                                 \
    }                                                                         \
  } while (0)
#else
#define NSFF_TL_DECL(role) do {} while (0)
#define NSFF_TL(tag) do {} while (0)
#endif

// barrier slots (8 bytes each) inside the barrier region
enum {
  BAR_W_FULL = 0,                           // [kMaxStages] weight stage landed (tx + partner relay)
  BAR_W_EMPTY = BAR_W_FULL + kMaxStages,    // [kMaxStages] MMAs that read the stage completed (commit)
  BAR_HID_FULL = BAR_W_EMPTY + kMaxStages,  // [4]  bf16 hidden chunk written to TMEM (by both CTAs)
  BAR_EXTRA_FULL = BAR_HID_FULL + 4,        // [1]  prologue chunks written
  BAR_ACC_FULL = BAR_EXTRA_FULL + 1,        // [kAccIds] accumulator complete (commit)
  BAR_ACC_EMPTY = BAR_ACC_FULL + kAccIds,   // [kAccIds] accumulator drained
  BAR_BIAS_FULL = BAR_ACC_EMPTY + kAccIds,  // [kBiasSlots]
  BAR_BIAS_EMPTY = BAR_BIAS_FULL + kBiasSlots,   // [kBiasSlots]
  BAR_COUNT = BAR_BIAS_EMPTY + kBiasSlots
};
static_assert(BAR_COUNT * 8 + 8 <= kSmemBar, "barrier region too small");

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// Writes 8 consecutive bf16 columns [8*unit, 8*unit+8) of `row` into a swizzled 64-column chunk.
__device__ __forceinline__ void store_unit(uint32_t chunk_base, int row, int unit, const float (&f)[8]) {
  uint32_t addr = chunk_base + (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + (((unit ^ (row & 7)) & 7) << 4));
  st_shared_v4(addr, pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
               pack_bf16x2(f[6], f[7]));
}
// Same for a row whose swizzle-row offset is precomputed; ReLU (if kRelu) rides on the bf16 conversion.
template <bool kRelu>
__device__ __forceinline__ void store_unit_at(uint32_t row_base, int rsw, int unit, const float (&f)[8]) {
  const uint32_t addr = row_base + (uint32_t)(((unit ^ rsw) & 7) << 4);
  if constexpr (kRelu)
    st_shared_v4(addr, pack_bf16x2_relu(f[0], f[1]), pack_bf16x2_relu(f[2], f[3]), pack_bf16x2_relu(f[4], f[5]),
                 pack_bf16x2_relu(f[6], f[7]));
  else
    st_shared_v4(addr, pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                 pack_bf16x2(f[6], f[7]));
}

// column c (0..63) of the xyz encoding row [x, sin(2^0 x), cos(2^0 x), ...] (nerf.py:17-30); col 63 = pad
template <int kNFreq>
__device__ __forceinline__ float pe_col(int c, const float (&base)[3], const float (&sn)[kNFreq][3],
                                        const float (&cs)[kNFreq][3], int f0) {
  if (c < 3) return base[c];
  int k = c - 3, f = k / 6, r = k % 6;
  if (f - f0 < 0 || f - f0 >= kNFreq) return 0.f;
  return r < 3 ? sn[f - f0][r] : cs[f - f0][r - 3];
}

// PE(xyz) columns [16*kPart, 16*kPart+16) of one row: the frequencies those columns touch are
// part 0: 2^0..2^2, part 1: 2^2..2^4, part 2: 2^4..2^7, part 3: 2^7..2^9 (+ the zero pad column 63).
template <int kPart>
__device__ __forceinline__ void encode_pe_part(uint32_t chPE, int row, const float (&base)[3]) {
  constexpr int f0 = kPart == 0 ? 0 : kPart == 1 ? 2 : kPart == 2 ? 4 : 7;
  constexpr int nf = kPart == 2 ? 4 : 3;
  float sn[nf][3], cs[nf][3];
#pragma unroll
  for (int f = 0; f < nf; ++f)
#pragma unroll
    for (int c = 0; c < 3; ++c) sincosf(__fmul_rn((float)(1 << (f0 + f)), base[c]), &sn[f][c], &cs[f][c]);
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = (kPart * 2 + u) * 8 + i;
      v[i] = c == 63 ? 0.f : pe_col<nf>(c, base, sn, cs, f0);
    }
    store_unit(chPE, row, kPart * 2 + u, v);
  }
}

// Prologue of one tile: encode the inputs of `row` into the PE / time-code / PE(dir) chunks (bf16, swizzled).
// Four threads (part 0..3) share a row: each writes 2 of the 8 PE units; parts 0-2 also write 2 of the 6
// time-code units, part 3 the 4 PE(dir) units.
__device__ __forceinline__ void encode_tile_row(const FwdParams& P, uint32_t sA, int row, int part, int64_t p,
                                                bool valid) {
  const uint32_t chPE = sA + 0 * kChunkBytes, chT = sA + 1 * kChunkBytes, chD = sA + 2 * kChunkBytes;
  if (P.x_emb == nullptr) {
    const int64_t ray = p / P.pts.S;
    float base[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = P.pts.xyz[p * 3 + c];
      if (P.pts.offset) v = __fadd_rn(v, P.pts.offset[p * P.pts.offset_ld + c]);
      base[c] = v;
    }
    if (part == 0 && valid && P.pts.xyz_out) {
#pragma unroll
      for (int c = 0; c < 3; ++c) P.pts.xyz_out[p * 3 + c] = base[c];
    }
    if (part == 0) encode_pe_part<0>(chPE, row, base);
    else if (part == 1) encode_pe_part<1>(chPE, row, base);
    else if (part == 2) encode_pe_part<2>(chPE, row, base);
    else encode_pe_part<3>(chPE, row, base);
    if (P.need_t && part < 3) {
      int64_t trow = ray;
      if (P.pts.t_index) {
        trow = P.pts.t_index[ray] + P.pts.t_shift;
        trow = trow < 0 ? 0 : (trow > P.pts.max_t ? P.pts.max_t : trow);
      }
      const float4* tp = reinterpret_cast<const float4*>(P.pts.t_table + trow * 48) + part * 4;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        float4 a = __ldg(tp + 2 * u), b = __ldg(tp + 2 * u + 1);
        float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        store_unit(chT, row, part * 2 + u, v);
      }
    }
    if (P.need_a && part < 3) {
      const int64_t arow = P.pts.a_index ? P.pts.a_index[ray] : ray;
      const float4* ap = reinterpret_cast<const float4*>(P.pts.a_table + arow * 48) + part * 4;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        float4 a = __ldg(ap + 2 * u), b = __ldg(ap + 2 * u + 1);
        float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        store_unit(sA + 3 * kChunkBytes, row, part * 2 + u, v);
      }
    }
    if (P.need_dir && part == 3) {
      float d[3] = {P.pts.dirs[ray * 3 + 0], P.pts.dirs[ray * 3 + 1], P.pts.dirs[ray * 3 + 2]};
      float sn[4][3], cs[4][3];
#pragma unroll
      for (int f = 0; f < 4; ++f)
#pragma unroll
        for (int c = 0; c < 3; ++c) sincosf(__fmul_rn((float)(1 << f), d[c]), &sn[f][c], &cs[f][c]);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (u * 8 + i >= 27) ? 0.f : pe_col<4>(u * 8 + i, d, sn, cs, 0);
        store_unit(chD, row, u, v);
      }
    }
  } else {
    // embedded mode: x row = [xyz63 | dir27 | a | t48] already encoded by the caller
    const float* xr = P.x_emb + p * P.x_ld;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int c = (part * 2 + u) * 8 + i;
        v[i] = c < 63 ? xr[c] : 0.f;
      }
      store_unit(chPE, row, part * 2 + u, v);
    }
    if (P.need_t && part < 3) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = xr[P.c_t + (part * 2 + u) * 8 + i];
        store_unit(chT, row, part * 2 + u, v);
      }
    }
    if (P.need_a && part < 3) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = xr[P.c_a + (part * 2 + u) * 8 + i];
        store_unit(sA + 3 * kChunkBytes, row, part * 2 + u, v);
      }
    }
    if (P.need_dir && part == 3) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (u * 8 + i >= 27) ? 0.f : xr[P.c_dir + u * 8 + i];
        store_unit(chD, row, u, v);
      }
    }
  }
}

template <bool kAppear, bool kSave>
__global__ void __launch_bounds__(kThreads, 1) field_forward_kernel(const __grid_constant__ FwdParams P) {
  using C = Cfg<kAppear>;
  constexpr int kStages = C::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base;                       // extra chunks
  const uint32_t sB = smem_base + C::kSmemA;           // weight ring
  const uint32_t sBias = sB + C::kSmemB;
  const uint32_t sBar = sBias + kSmemBias;
  const uint32_t sTmemPtr = sBar + BAR_COUNT * 8;
  auto bar = [&](int i) { return sBar + (uint32_t)i * 8u; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  // barriers the MMA issuer waits on live in the leader CTA; both CTAs of the pair arrive there
  auto arrive_leader = [&](int i) { mbar_arrive_cluster(mapa_cluster(bar(i), 0)); };
  const int group = blockIdx.x >> 1, n_groups = gridDim.x >> 1;
  const int n_units = (P.n_tiles + 1) >> 1;            // tile pairs

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      // leader: own half (producer arrive + tx) AND the partner's relay arrive complete a stage
      mbar_init(bar(BAR_W_FULL + i), leader ? 2 : 1);
      mbar_init(bar(BAR_W_EMPTY + i), 1);
    }
    for (int i = 0; i < 4; ++i) mbar_init(bar(BAR_HID_FULL + i), kEpiWarps * 2);
    mbar_init(bar(BAR_EXTRA_FULL), kEpiWarps * 2);
    for (int i = 0; i < kAccIds; ++i) {
      mbar_init(bar(BAR_ACC_FULL + i), 1);
      mbar_init(bar(BAR_ACC_EMPTY + i), kEpiWarps * 2);
    }
    for (int i = 0; i < kBiasSlots; ++i) {
      mbar_init(bar(BAR_BIAS_FULL + i), 1);
      mbar_init(bar(BAR_BIAS_EMPTY + i), kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == kEpiWarps + 1) tmem_alloc_pair(sTmemPtr, 512);
  tc_fence_before_sync();
  cluster_sync_all();              // also: the peer's barriers are initialised before any remote arrive
  tc_fence_after_sync();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(sTmemPtr));

  const int n_subs = P.n_subs, n_ops = P.n_ops;

  if (warp == kEpiWarps) {
    // ===================== weight + bias producer (whole warp walks the op list, one elected lane copies) ====
    uint32_t stage = 0, epar = 1, gsub = 0;        // epar: parity that means "stage is free" (fresh barrier: 1)
    NSFF_TL_DECL(0);
    for (int unit = group; unit < n_units; unit += n_groups) {
      int sub = 0;
      for (int i = 0; i < n_ops; ++i) {
        const ChunkOp op = P.ops[i];
        if (op.flags & OP_FIRST) {
          const SubStep& ss = P.subs[sub];
          const uint32_t bslot = gsub & (kBiasSlots - 1), buse = gsub / kBiasSlots;
          mbar_wait_sleep(bar(BAR_BIAS_EMPTY + bslot), (buse & 1u) ^ 1u);
          if (elect_one()) {
            mbar_arrive_expect_tx(bar(BAR_BIAS_FULL + bslot), (uint32_t)ss.n * 4u);
            bulk_g2s(sBias + bslot * 1024u, P.bias + ss.bias_off, (uint32_t)ss.n * 4u, bar(BAR_BIAS_FULL + bslot));
          }
          __syncwarp();
          ++sub, ++gsub;
        }
        mbar_wait_sleep(bar(BAR_W_EMPTY + stage), epar);
        NSFF_TL(((gsub - 1) << 8) | 1);                    // producer: stage free, issuing copy
        if (elect_one()) {
          mbar_arrive_expect_tx(bar(BAR_W_FULL + stage), op.bytes);
          bulk_g2s(sB + stage * kStageBytes, P.pack + op.w_src + cta_rank * op.bytes, op.bytes, bar(BAR_W_FULL + stage));
        }
        if (++stage == (uint32_t)kStages) stage = 0, epar ^= 1u;
      }
    }
  } else if (warp == kEpiWarps + 1) {
    if (leader) {
      // ===================== MMA issuer =====================
      // One elected lane walks the flat op list (the other 31 park at the reconvergence point).  The loop body is the
      // serial critical path of the kernel - one instruction stream feeds the tensor pipe, 64 cycles per MMA in split
      // mode - so everything that can be precomputed lives in the 16-byte op, and the barrier probes of op i+1 are
      // issued before the MMAs of op i so that their round trip overlaps the issue work.
      const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);          // SBO | version | SWIZZLE_128B
      const uint32_t a_lo0 = ((sA & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t b_lo0 = ((sB & 0x3FFFFu) >> 4) | (1u << 16);
      auto desc = [&](uint32_t lo) { return ((uint64_t)desc_hi << 32) | (uint64_t)lo; };
      if (elect_one()) {
        uint32_t stage = 0, wpar = 0, hid_gen = 0, unit_iter = 0, gsub = 0;
        uint32_t fill_par = 0;             // bit a: (number of fills of accumulator a so far) & 1
        NSFF_TL_DECL(1);
        for (int unit = group; unit < n_units; unit += n_groups, ++unit_iter) {
          mbar_wait(bar(BAR_EXTRA_FULL), unit_iter & 1u);
          NSFF_TL((gsub << 8) | 15);                           // mma: prologue chunks ready
          ChunkOp op = P.ops[0];
          bool hid_ok = op.wait_hid == 0 || mbar_try_wait(bar(BAR_HID_FULL + op.wait_hid - 1), (hid_gen - 1u) & 1u);
          bool w_ok = mbar_try_wait(bar(BAR_W_FULL + stage), wpar);
          int kc = 0;
          for (int i = 0; i < n_ops; ++i) {
            if (op.flags & OP_FIRST) {
              kc = 0;
              uint32_t m = op.wait_empty;
              while (m) {                                      // accumulators this sub-step overwrites: drained?
                const int a = __ffs((int)m) - 1;
                m &= m - 1u;
                mbar_wait(bar(BAR_ACC_EMPTY + a), ((fill_par >> a) & 1u) ^ 1u);
              }
              NSFF_TL((gsub << 8) | 14);                       // mma: accumulator free
            }
            NSFF_TL((gsub << 8) | (kc << 4) | 2);              // mma: chunk loop top
            if (!hid_ok) mbar_wait(bar(BAR_HID_FULL + op.wait_hid - 1), (hid_gen - 1u) & 1u);
            NSFF_TL((gsub << 8) | (kc << 4) | 3);              // mma: A chunk ready
            if (!w_ok) mbar_wait(bar(BAR_W_FULL + stage), wpar);
            NSFF_TL((gsub << 8) | (kc << 4) | 4);              // mma: weights ready -> issue
            const ChunkOp cur = op;
            const uint32_t b_lo = b_lo0 + stage * (kStageBytes >> 4);
            const uint32_t empty_bar = bar(BAR_W_EMPTY + stage);
            if (++stage == (uint32_t)kStages) stage = 0, wpar ^= 1u;
            const uint32_t hid_next = hid_gen + ((cur.flags & OP_ADV_HID) ? 1u : 0u);
            if (i + 1 < n_ops) {                               // early probes for the next op
              op = P.ops[i + 1];
              hid_ok = op.wait_hid == 0 || mbar_try_wait(bar(BAR_HID_FULL + op.wait_hid - 1), (hid_next - 1u) & 1u);
              w_ok = mbar_try_wait(bar(BAR_W_FULL + stage), wpar);
            }
            tc_fence_after_sync();
            const uint32_t d_tmem = tmem_base + cur.d_col;
            const uint32_t idesc = make_idesc_bf16(2 * kTileM, (uint32_t)cur.n16 * 16u);
            const uint32_t acc0 = (cur.flags & OP_FIRST) ? 0u : 1u;
            const uint32_t kcnt = cur.kcnt;
            if (cur.flags & OP_TS) {
              const uint32_t a_t = tmem_base + cur.a;          // 8 columns (16 bf16) per K-step, 16 columns apart
              mma_bf16_ts_pair(d_tmem, a_t, desc(b_lo), idesc, acc0);
              mma_bf16_ts_pair(d_tmem, a_t + 16, desc(b_lo + 2), idesc, 1u);
              if (kcnt > 2) mma_bf16_ts_pair(d_tmem, a_t + 32, desc(b_lo + 4), idesc, 1u);
              if (kcnt > 3) mma_bf16_ts_pair(d_tmem, a_t + 48, desc(b_lo + 6), idesc, 1u);
            } else {
              const uint32_t a_lo = a_lo0 + cur.a;             // +32 B per K-step inside the 128 B swizzle row
              mma_bf16_ss_pair(d_tmem, desc(a_lo), desc(b_lo), idesc, acc0);
              mma_bf16_ss_pair(d_tmem, desc(a_lo + 2), desc(b_lo + 2), idesc, 1u);
              if (kcnt > 2) mma_bf16_ss_pair(d_tmem, desc(a_lo + 4), desc(b_lo + 4), idesc, 1u);
              if (kcnt > 3) mma_bf16_ss_pair(d_tmem, desc(a_lo + 6), desc(b_lo + 6), idesc, 1u);
            }
            mma_commit_pair_mc(empty_bar, 3);                  // frees the stage in BOTH CTAs
            if (cur.flags & OP_LAST) {
              mma_commit_pair_mc(bar(BAR_ACC_FULL + cur.acc_id), 3);
              fill_par ^= 1u << cur.acc_id;
              ++gsub;
            }
            NSFF_TL(((gsub - ((cur.flags & OP_LAST) ? 1u : 0u)) << 8) | (kc << 4) | 11);   // mma: committed
            hid_gen = hid_next;
            ++kc;
          }
        }
      }
      __syncwarp();
    } else {
      // ===================== peer CTA: relay "my weight half has landed" to the leader ==============
      uint32_t stage = 0, wpar = 0;
      const uint32_t leader_w_full = mapa_cluster(bar(BAR_W_FULL), 0);
      for (int unit = group; unit < n_units; unit += n_groups) {
        for (int i = 0; i < n_ops; ++i) {
          mbar_wait_sleep(bar(BAR_W_FULL + stage), wpar);
          if (lane == 0) mbar_arrive_cluster(leader_w_full + stage * 8u);
          if (++stage == (uint32_t)kStages) stage = 0, wpar ^= 1u;
        }
      }
    }
  } else {
    // ===================== prologue + epilogue warps =====================
    const int q = warp & 3, part = warp >> 2;      // TMEM lane quadrant, 16-column part of every 64-column slice
    const int row = q * 32 + lane;
    const uint32_t lane_field = (uint32_t)(q * 32) << 16;
    const uint32_t row_off = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
    const int rsw = row & 7;
    uint32_t gsub = 0;
    uint32_t drain_par = 0;                        // bit a: (number of drains of accumulator a so far) & 1
#ifdef NSFF_TIMELINE
    unsigned long long* tl_ptr = (P.dbg && blockIdx.x == 0 && warp == 0) ? P.dbg + 2 * 2 * kDbgPerRole : nullptr;
    int tl_n = 0;
#endif
    // Prologue: encode a tile's inputs into the extra chunks.  Only the first tile pays for it up front; for every
    // later tile it is run during the previous tile, right after the last sub-step that reads those chunks
    // (P.last_extra_sub), i.e. hidden behind the remaining layers instead of idling the tensor pipe.
    auto prologue = [&](int unit_) {
      const int64_t pr = (int64_t)(unit_ * 2 + (int)cta_rank) * kTileM + row;
      const bool ok = pr < P.pts.B;
      encode_tile_row(P, sA, row, part, ok ? pr : P.pts.B - 1, ok);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) arrive_leader(BAR_EXTRA_FULL);
    };
    if (group < n_units) prologue(group);
    for (int unit = group; unit < n_units; unit += n_groups) {
      NSFF_TL((gsub << 8) | 13);                             // epi: tile start
      const int tile = unit * 2 + (int)cta_rank;
      const int64_t p_raw = (int64_t)tile * kTileM + row;
      const bool valid = p_raw < P.pts.B;
      const int64_t p = valid ? p_raw : P.pts.B - 1;

      float zrow = 0.f;
      if (P.pts.zs && P.x_emb == nullptr) zrow = P.pts.zs[p];

      // ---------------- epilogues, one per sub-step ----------------
      for (int s = 0; s < n_subs; ++s, ++gsub) {
        const SubStep& ss = P.subs[s];
        const uint32_t acc = ss.acc_id;
        const uint32_t bslot = gsub & (kBiasSlots - 1);
        mbar_wait(bar(BAR_BIAS_FULL + bslot), (gsub / kBiasSlots) & 1u);   // landed long ago: issued a sub-step ahead
        const uint32_t s_bias = sBias + bslot * 1024u;
        const uint32_t t_acc = tmem_base + lane_field + ss.d_col;
        if (ss.epi != EPI_HEAD) {
          // bias of the first slice is fetched before the accumulator is awaited: it is off the critical path
          float4 b[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) lds_v4(s_bias + (uint32_t)((part * 16 + i * 4) * 4), b[i]);
          NSFF_TL((gsub << 8) | 5);                            // epi: waiting for the accumulator
          mbar_wait(bar(BAR_ACC_FULL + acc), (drain_par >> acc) & 1u);
          NSFF_TL((gsub << 8) | 6);                            // epi: accumulator complete
          tc_fence_after_sync();
          const bool relu = ss.epi == EPI_RELU_HID;
          const int nsl = ss.nslices;
          // software pipeline over the 64-column slices: TMEM load of slice c+1 is in flight while slice c is biased /
          // activated / packed / written back to TMEM as the next layer's A operand.
          uint32_t v[2][16];
          tmem_ld16(t_acc + (uint32_t)(part * 16), v[0]);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (c < nsl) {
              tmem_ld_wait();
              if (c + 1 < nsl) {
                tmem_ld16(t_acc + (uint32_t)((c + 1) * 64 + part * 16), v[(c + 1) & 1]);
              } else {
                // accumulator fully drained into registers: its columns may be overwritten by the layer after next
                // (which cannot start before the bf16 image written below has been published and consumed)
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) arrive_leader(BAR_ACC_EMPTY + acc);
              }
              uint32_t w[8];
              uint32_t pos16 = 0;
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                float f[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 bb = b[u * 2 + (i >> 2)];
                  const float bv = (i & 3) == 0 ? bb.x : (i & 3) == 1 ? bb.y : (i & 3) == 2 ? bb.z : bb.w;
                  f[i] = __uint_as_float(v[c & 1][u * 8 + i]) + bv;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  w[u * 4 + i] = relu ? pack_bf16x2_relu(f[2 * i], f[2 * i + 1]) : pack_bf16x2(f[2 * i], f[2 * i + 1]);
                if constexpr (kSave) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) pos16 |= (f[i] > 0.f ? 1u : 0u) << (u * 8 + i);
                }
              }
              // K-step `part` of hidden chunk slice0+c: 16 bf16 = 8 columns, in place (columns this warp has drained)
              tmem_st8(t_acc + (uint32_t)(((c >> 1) << 6) + ((c & 1) << 3) + part * 16), w[0], w[1], w[2], w[3], w[4],
                       w[5], w[6], w[7]);
              tmem_st_wait();
              tc_fence_before_sync();
              __syncwarp();
              if (lane == 0) arrive_leader(BAR_HID_FULL + ss.slice0 + c);
              if constexpr (kSave) {
                // training mode: the bf16 rows just published are also the backward's tile image (128B-swizzled chunk
                // [tile][chunk][128 rows x 128 B]); this warp owns one aligned 32-byte sector of its row
                if (ss.save_off != kNoSave && tile < P.n_tiles) {
                  const int cg = ss.slice0 + c;
                  uint8_t* g = P.save_base + ss.save_off + (size_t)tile * kSaveImgPerTile + (size_t)cg * kChunkBytes +
                               (row_off);
                  const bool swap = rsw & 1;
                  const uint4 lo = make_uint4(w[0], w[1], w[2], w[3]), hi = make_uint4(w[4], w[5], w[6], w[7]);
                  st_global_v8(g + ((((part * 2) ^ rsw) & 6) << 4), swap ? hi : lo, swap ? lo : hi);
                  if (relu)
                    *reinterpret_cast<uint16_t*>(P.save_base + ss.bits_off + (size_t)tile * kSaveBitsPerTile +
                                                 ((size_t)(2 * cg + (part >> 1)) * 128 + row) * 4 + (part & 1) * 2) =
                        (uint16_t)pos16;
                }
              }
              NSFF_TL((gsub << 8) | (c << 4) | 7);             // epi: hidden chunk published
              if (c + 1 < nsl) {
#pragma unroll
                for (int i = 0; i < 4; ++i) lds_v4(s_bias + (uint32_t)(((c + 1) * 64 + part * 16 + i * 4) * 4), b[i]);
              } else {
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(BAR_BIAS_EMPTY + bslot));      // bias row fully consumed
              }
            }
          }
        } else {
          mbar_wait(bar(BAR_ACC_FULL + acc), (drain_par >> acc) & 1u);
          tc_fence_after_sync();
          uint32_t v[16];
          float hb[16];
          if (part == 0) {
            tmem_ld16(t_acc, v);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float4 t4;
              lds_v4(s_bias + (uint32_t)(i * 16), t4);
              hb[4 * i] = t4.x, hb[4 * i + 1] = t4.y, hb[4 * i + 2] = t4.z, hb[4 * i + 3] = t4.w;
            }
            tmem_ld_wait();
          }
          // accumulator and bias row are in registers: hand both back before the activations / global stores
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) {
            arrive_leader(BAR_ACC_EMPTY + acc);
            mbar_arrive(bar(BAR_BIAS_EMPTY + bslot));
          }
          if (part == 0 && valid) {
            float* orow = P.out + p * P.out_ld;
            float fl[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};      // flow fw / bw = head columns 4..9 of the dynamic block
            bool has_flow = false;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int dst = ss.head_dst[i];
              if (dst < 0) continue;
              float x = __uint_as_float(v[i]) + hb[i];
              const int act = ss.head_act[i];
              if (act == HA_SIGMOID) x = sigmoid_f(x);
              else if (act == HA_FLOW) {
                x = (zrow > kZFar) ? 0.f : P.flow_scale * tanhf(x);
                if (i >= 4 && i < 10) fl[i - 4] = x, has_flow = true;
              }
              orow[dst] = x;
            }
            // points of the warped passes, written here so those passes do not gather 12 B out of every output row
            if (has_flow && P.x_emb == nullptr && P.pts.offset == nullptr) {
              const float* x3 = P.pts.xyz + p * 3;
              if (P.pts.xyz_fw_out) {
#pragma unroll
                for (int c = 0; c < 3; ++c) P.pts.xyz_fw_out[p * 3 + c] = __fadd_rn(x3[c], fl[c]);
              }
              if (P.pts.xyz_bw_out) {
#pragma unroll
                for (int c = 0; c < 3; ++c) P.pts.xyz_bw_out[p * 3 + c] = __fadd_rn(x3[c], fl[3 + c]);
              }
            }
          }
        }
        drain_par ^= 1u << acc;
        // this sub-step's accumulator was complete => every MMA up to and including it has finished reading smem
        if (s == P.last_extra_sub && unit + n_groups < n_units) prologue(unit + n_groups);
      }
    }
  }

  tc_fence_before_sync();
  cluster_sync_all();              // no CTA of the pair may exit while its partner can still signal it
  if (warp == kEpiWarps + 1) {
    tc_fence_after_sync();
    __syncwarp();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

}  // namespace tc
}  // namespace nsff

using namespace nsff;
using namespace nsff::tc;

extern "C" int nsff_field_pack_bytes(const nsff_field_desc* desc, size_t* bytes) {
  if (!desc || !bytes) return NSFF_EINVAL;
  PackLayout L;
  int rc = build_layout(*desc, L);
  if (rc != NSFF_OK) return rc;
  // the job table of nsff_field_pack() is staged right after the image
  *bytes = ((L.bytes + 1023) & ~(size_t)1023) + sizeof(PackJobs);
  return NSFF_OK;
}

extern "C" int nsff_field_pack(const nsff_field_desc* desc, const nsff_field_params* prm, void* pack,
                               size_t pack_bytes, void* stream) {
  if (!desc || !prm || !pack) return NSFF_EINVAL;
  if (((uintptr_t)pack & 127u) != 0) return NSFF_EINVAL;   // bulk copies need 16 B; keep lines whole
  PackLayout L;
  int rc = build_layout(*desc, L);
  if (rc != NSFF_OK) return rc;
  const size_t jobs_off = (L.bytes + 1023) & ~(size_t)1023;
  if (pack_bytes < jobs_off + sizeof(PackJobs)) return NSFF_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;

  static thread_local PackJobs jobs;   // host staging; copied synchronously w.r.t. the host below
  int nj = 0;
  bool missing = false;
  auto wjob = [&](int lid, int kc, const float* w, int ld, int row0, int rows, int col0, int cols) {
    const PackLayer& pl = L.layer[lid];
    if (!w) { missing = true; return; }
    PackJob& j = jobs.job[nj++];
    j.src = w; j.ld = ld; j.row0 = row0; j.rows = rows; j.col0 = col0; j.cols = cols; j.is_bias = 0;
    j.dst_off = pl.w_off + (uint32_t)kc * (uint32_t)pl.n * 128u;
  };
  auto bjob = [&](int lid, int idx0, const float* b, int n) {
    if (!b) { missing = true; return; }
    PackJob& j = jobs.job[nj++];
    j.src = b; j.rows = n; j.is_bias = 1; j.dst_off = L.layer[lid].bias_off + idx0;
    j.ld = j.row0 = j.col0 = j.cols = 0;
  };
  // a 256-wide layer whose K runs over [pre-columns..., hidden 256]; `pre` = {(chunk-slot, col0, cols)}
  auto trunk_layer = [&](int lid, const float* w, const float* b, int ld, int hidden_col0, int n_pre,
                         const int (*pre)[2]) {
    int kc = 0;
    for (int i = 0; i < n_pre; ++i, ++kc) wjob(lid, kc, w, ld, 0, 256, pre[i][0], pre[i][1]);
    if (hidden_col0 >= 0)
      for (int h = 0; h < 4; ++h, ++kc) wjob(lid, kc, w, ld, 0, 256, hidden_col0 + h * 64, 64);
    bjob(lid, 0, b, 256);
  };
  const int pre_pe[1][2] = {{0, 63}};
  const int pre_pe_t[2][2] = {{0, 63}, {63, 48}};
  const int pre_dir[1][2] = {{256, 27}};
  // static branch (nerf.py:74-91)
  trunk_layer(L_S1, prm->static_w[0], prm->static_b[0], 63, -1, 1, pre_pe);
  for (int i = 1; i < 8; ++i) {
    if (i == 4) trunk_layer(L_S1 + i, prm->static_w[i], prm->static_b[i], 319, 63, 1, pre_pe);
    else trunk_layer(L_S1 + i, prm->static_w[i], prm->static_b[i], 256, 0, 0, nullptr);
  }
  for (int h = 0; h < 4; ++h) wjob(L_SSIG, 3 - h, prm->static_sigma_w, 256, 0, 1, h * 64, 64);   // K-chunk order H3..H0
  bjob(L_SSIG, 0, prm->static_sigma_b, 1);
  trunk_layer(L_SFIN, prm->static_final_w, prm->static_final_b, 256, 0, 0, nullptr);
  const int pre_dir_a[2][2] = {{256, 27}, {283, 48}};
  if (desc->use_viewdir) {
    if (desc->in_a > 0) trunk_layer(L_SDIR, prm->static_dir_w, prm->static_dir_b, 283 + 48, 0, 2, pre_dir_a);
    else trunk_layer(L_SDIR, prm->static_dir_w, prm->static_dir_b, 283, 0, 1, pre_dir);
  }
  for (int h = 0; h < 4; ++h) wjob(L_SRGB, 3 - h, prm->static_rgb_w, 256, 0, 3, h * 64, 64);
  bjob(L_SRGB, 0, prm->static_rgb_b, 3);
  if (desc->has_transient) {   // nerf.py:93-116
    trunk_layer(L_T1, prm->trans_w[0], prm->trans_b[0], 111, -1, 2, pre_pe_t);
    for (int i = 1; i < 8; ++i) {
      if (i == 4) trunk_layer(L_T1 + i, prm->trans_w[i], prm->trans_b[i], 367, 111, 2, pre_pe_t);
      else trunk_layer(L_T1 + i, prm->trans_w[i], prm->trans_b[i], 256, 0, 0, nullptr);
    }
    trunk_layer(L_TFIN, prm->trans_final_w, prm->trans_final_b, 256, 0, 0, nullptr);
    // head block rows: 0-2 rgb, 3 sigma, 4-6 flow fw, 7-9 flow bw
    for (int h = 0; h < 4; ++h) {
      wjob(L_THEAD, 3 - h, prm->trans_rgb_w, 256, 0, 3, h * 64, 64);
      wjob(L_THEAD, 3 - h, prm->trans_sigma_w, 256, 3, 1, h * 64, 64);
      if (desc->has_flow) {
        wjob(L_THEAD, 3 - h, prm->trans_fw_w, 256, 4, 3, h * 64, 64);
        wjob(L_THEAD, 3 - h, prm->trans_bw_w, 256, 7, 3, h * 64, 64);
      }
    }
    bjob(L_THEAD, 0, prm->trans_rgb_b, 3);
    bjob(L_THEAD, 3, prm->trans_sigma_b, 1);
    if (desc->has_flow) {
      bjob(L_THEAD, 4, prm->trans_fw_b, 3);
      bjob(L_THEAD, 7, prm->trans_bw_b, 3);
    }
  }
  if (missing) return NSFF_EINVAL;
  if (nj > kMaxPackJobs) return NSFF_EUNSUPPORTED;

  uint8_t* p8 = static_cast<uint8_t*>(pack);
  NSFF_CUDA_TRY(cudaMemsetAsync(p8, 0, jobs_off, st));   // zero padding rows/cols and unused bias slots
  NSFF_CUDA_TRY(cudaMemcpyAsync(p8 + jobs_off, &jobs, sizeof(PackJob) * nj, cudaMemcpyHostToDevice, st));
  // (pageable-source cudaMemcpyAsync returns once `jobs` has been staged, so the host copy may be reused)
  pack_kernel<<<nj, 256, 0, st>>>(reinterpret_cast<const PackJobs*>(p8 + jobs_off), p8, L.bias_base);
  return after_launch();
}

// Builds the per-tile program: the sub-step table the epilogue warps walk and the flat K-chunk op list the producer
// and the MMA issuer walk.  `split`: every 256-wide layer becomes two N = 128 sub-steps (see the file header).
static int build_program(const nsff_field_desc& d, const PackLayout& L, int out_mask, bool split, FwdParams& P,
                         int* n_out_cols, const SaveLayout* SL) {
  const bool sigma_only = out_mask & NSFF_OUT_SIGMA_ONLY;
  const bool want_s = out_mask & NSFF_OUT_STATIC, want_t = out_mask & NSFF_OUT_TRANSIENT;
  const bool want_fw = out_mask & NSFF_OUT_FLOW_FW, want_bw = out_mask & NSFF_OUT_FLOW_BW;
  if (!want_s && !want_t) return NSFF_EINVAL;
  if (want_t && !d.has_transient) return NSFF_EINVAL;
  if ((want_fw || want_bw) && (!want_t || !d.has_flow || sigma_only)) return NSFF_EINVAL;
  int n_subs = 0, n_ops = 0, col = 0;
  int cur = -1;                        // region (0 = X, 1 = Y) that holds the current hidden activation; -1: none yet
  bool overflow = false;
  P.last_extra_sub = 0;
  const uint32_t half_cols = split ? 128u : 64u;     // TMEM columns between the bf16 images of channels 0-127 / 128-255
  auto push = [&](int lid, uint8_t epi) -> SubStep* {
    const PackLayer& pl = L.layer[lid];
    const bool head = epi == EPI_HEAD;
    const int parts = (!head && split) ? 2 : 1;
    const int region = head ? cur : (cur < 0 ? 0 : 1 - cur);
    if (head && cur < 0) { overflow = true; return nullptr; }
    SubStep* first = nullptr;
    for (int h = 0; h < parts; ++h) {
      if (n_subs >= kMaxSubs) { overflow = true; return nullptr; }
      SubStep& s = P.subs[n_subs];
      if (!first) first = &s;
      const int n = head ? pl.n : pl.n / parts;
      s.bias_off = pl.bias_off + (uint32_t)(h * n);
      s.n = (uint16_t)n;
      s.epi = epi;
      if (head) {
        // free columns of the region that holds the input: above the two in-place bf16 images
        s.d_col = (uint16_t)(region * 256 + (split ? 64 : 128));
        s.acc_id = kHeadAcc;
        s.slice0 = 0, s.nslices = 0;
      } else {
        s.d_col = (uint16_t)(region * 256 + h * 128);
        s.acc_id = (uint8_t)(region * 2 + h);
        s.slice0 = (uint8_t)(2 * h);
        s.nslices = (uint8_t)(4 / parts);
      }
      for (int i = 0; i < 16; ++i) s.head_dst[i] = -1, s.head_act[i] = HA_NONE;
      s.save_off = s.bits_off = kNoSave;
      if (SL) {
        if (lid >= L_S1 && lid <= L_S8) s.save_off = SL->s_h[lid - L_S1], s.bits_off = SL->s_hb[lid - L_S1];
        else if (lid == L_SFIN) s.save_off = SL->s_feat;
        else if (lid == L_SDIR) s.save_off = SL->s_dir, s.bits_off = SL->s_dirb;
        else if (lid >= L_T1 && lid <= L_T8) s.save_off = SL->t_h[lid - L_T1], s.bits_off = SL->t_hb[lid - L_T1];
        else if (lid == L_TFIN) s.save_off = SL->t_feat;
      }
      for (int kc = 0; kc < pl.nchunks; ++kc) {
        if (n_ops >= kMaxOps) { overflow = true; return nullptr; }
        ChunkOp& o = P.ops[n_ops++];
        const int ch = pl.chunk[kc];
        const uint32_t block_bytes = (uint32_t)pl.n * 128u;              // whole N x 64 block of this K-chunk
        o.bytes = (uint16_t)((uint32_t)n * 128u / 2u);                   // this CTA's N/2 rows
        o.w_src = pl.w_off + (uint32_t)kc * block_bytes + (uint32_t)h * (uint32_t)n * 128u;
        o.d_col = s.d_col;
        o.n16 = (uint8_t)(n / 16);
        o.kcnt = pl.kcnt[kc];
        o.acc_id = s.acc_id;
        o.flags = 0;
        o.wait_hid = 0;
        o.wait_empty = 0;
        if (ch < CH_PE) {
          if (cur < 0) { overflow = true; return nullptr; }
          o.flags |= OP_TS;
          o.a = (uint16_t)(cur * 256 + (ch >> 1) * (int)half_cols + (ch & 1) * 8);
          if (h == 0) o.wait_hid = (uint8_t)(1 + ch);                    // later uses in the same layer: already awaited
        } else {
          o.a = (uint16_t)((ch - CH_PE) * (int)(kChunkBytes >> 4));
          P.last_extra_sub = n_subs;
        }
        if (kc == 0) {
          o.flags |= OP_FIRST;
          o.wait_empty = (uint8_t)((1u << s.acc_id) | (1u << kHeadAcc));
        }
        if (kc == pl.nchunks - 1) {
          o.flags |= OP_LAST;
          if (!head && h == parts - 1) o.flags |= OP_ADV_HID;
        }
      }
      ++n_subs;
    }
    if (!head) cur = region;
    return first;
  };
  P.need_dir = 0;
  P.need_t = 0;
  P.need_a = 0;
  if (want_s) {
    for (int i = L_S1; i <= L_S8; ++i) push(i, EPI_RELU_HID);
    if (sigma_only) {
      SubStep* s = push(L_SSIG, EPI_HEAD);
      if (s) s->head_dst[0] = (int8_t)col;
      col += 1;
    } else {
      SubStep* sg = push(L_SSIG, EPI_HEAD);
      if (sg) sg->head_dst[0] = (int8_t)(col + 3);
      push(L_SFIN, EPI_LIN_HID);
      if (d.use_viewdir) {
        push(L_SDIR, EPI_RELU_HID);
        P.need_dir = 1;
        P.need_a = d.in_a > 0;
      }
      SubStep* rg = push(L_SRGB, EPI_HEAD);
      if (rg)
        for (int i = 0; i < 3; ++i) rg->head_dst[i] = (int8_t)(col + i), rg->head_act[i] = HA_SIGMOID;
      col += 4;
    }
  }
  if (want_t) {
    P.need_t = 1;
    for (int i = L_T1; i <= L_T8; ++i) push(i, EPI_RELU_HID);
    push(L_TFIN, EPI_LIN_HID);
    SubStep* h = push(L_THEAD, EPI_HEAD);
    if (h) {
      if (sigma_only) {
        h->head_dst[3] = (int8_t)col;
        col += 1;
      } else {
        for (int i = 0; i < 3; ++i) h->head_dst[i] = (int8_t)(col + i), h->head_act[i] = HA_SIGMOID;
        h->head_dst[3] = (int8_t)(col + 3);
        col += 4;
        if (want_fw) {
          for (int i = 0; i < 3; ++i) h->head_dst[4 + i] = (int8_t)(col + i), h->head_act[4 + i] = HA_FLOW;
          col += 3;
        }
        if (want_bw) {
          for (int i = 0; i < 3; ++i) h->head_dst[7 + i] = (int8_t)(col + i), h->head_act[7 + i] = HA_FLOW;
          col += 3;
        }
      }
    }
  }
  if (overflow) return NSFF_EUNSUPPORTED;
  P.n_subs = n_subs;
  P.n_ops = n_ops;
  *n_out_cols = col;
  return NSFF_OK;
}

static bool env_flag(const char* name, bool dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) != 0 : dflt;
}

static int launch_forward(const nsff_field_desc* desc, const void* pack, FwdParams& P, int out_mask, int64_t B,
                          float* out, int32_t out_ld, void* stream, void* save = nullptr) {
  if (!desc || !pack || !out) return NSFF_EINVAL;
  if (B < 0 || B > ((int64_t)1 << 31) - 256) return NSFF_EINVAL;
  if (B == 0) return NSFF_OK;
  PackLayout L;
  int rc = build_layout(*desc, L);
  if (rc != NSFF_OK) return rc;
  int ncols = 0;
  SaveLayout SL;
  if (save) {
    if (out_mask & NSFF_OUT_SIGMA_ONLY) return NSFF_EUNSUPPORTED;
    SL = save_layout(desc, B, out_mask);
    P.save_base = static_cast<uint8_t*>(save);
  }
  // NSFF_TC_SPLIT=0 selects whole-layer accumulators (N = 256 MMAs; A/B experiments)
  rc = build_program(*desc, L, out_mask, env_flag("NSFF_TC_SPLIT", true), P, &ncols, save ? &SL : nullptr);
  if (rc != NSFF_OK) return rc;
  if (out_ld < ncols) return NSFF_EINVAL;
  int dev = 0, major = 0, sms = 0;
  NSFF_CUDA_TRY(cudaGetDevice(&dev));
  NSFF_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) return NSFF_EARCH;
  NSFF_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  P.pack = static_cast<const uint8_t*>(pack);
  P.bias = reinterpret_cast<const float*>(P.pack + L.bias_base);
  P.out = out;
  P.out_ld = out_ld;
  P.n_tiles = (int)cdiv64(B, kTileM);
  P.flow_scale = desc->flow_scale;
  int max_ctas = sms;
  if (const char* cap = getenv("NSFF_MAX_CTAS")) {          // experiment knob: fewer resident CTAs
    int c = atoi(cap);
    if (c > 0 && c < max_ctas) max_ctas = c;
  }
  cudaStream_t st = (cudaStream_t)stream;
#ifdef NSFF_TIMELINE
  const char* tl_path = getenv("NSFF_TIMELINE_OUT");
  unsigned long long* tl_dev = nullptr;
  const size_t tl_bytes = (size_t)3 * 2 * kDbgPerRole * sizeof(unsigned long long);
  if (tl_path) {
    NSFF_CUDA_TRY(cudaMalloc(&tl_dev, tl_bytes));
    NSFF_CUDA_TRY(cudaMemsetAsync(tl_dev, 0, tl_bytes, st));
    P.dbg = tl_dev;
  }
#endif
  const bool appear = P.need_a != 0;
  auto launch_pair = [&](auto kern, uint32_t smem) -> int {
    NSFF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int n_pairs = (P.n_tiles + 1) / 2;
    int clusters = max_ctas / 2;
    if (clusters < 1) clusters = 1;
    if (n_pairs < clusters) clusters = n_pairs;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    NSFF_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, P));
    return NSFF_OK;
  };
  int lrc;
  if (save) {
    lrc = appear ? launch_pair(field_forward_kernel<true, true>, Cfg<true>::kSmemTotal)
                 : launch_pair(field_forward_kernel<false, true>, Cfg<false>::kSmemTotal);
  } else {
    lrc = appear ? launch_pair(field_forward_kernel<true, false>, Cfg<true>::kSmemTotal)
                 : launch_pair(field_forward_kernel<false, false>, Cfg<false>::kSmemTotal);
  }
  if (lrc != NSFF_OK) return lrc;
#ifdef NSFF_TIMELINE
  if (tl_dev) {
    static int tl_seq = 0;
    unsigned long long* host = (unsigned long long*)malloc(tl_bytes);
    NSFF_CUDA_TRY(cudaStreamSynchronize(st));
    NSFF_CUDA_TRY(cudaMemcpy(host, tl_dev, tl_bytes, cudaMemcpyDeviceToHost));
    char name[512];
    snprintf(name, sizeof(name), "%s.%d.bin", tl_path, tl_seq++);
    if (FILE* f = fopen(name, "wb")) {
      fwrite(host, 1, tl_bytes, f);
      fclose(f);
    }
    free(host);
    cudaFree(tl_dev);
  }
#endif
  return after_launch();
}

extern "C" int nsff_field_forward(const nsff_field_desc* desc, const void* pack, const nsff_points* pts,
                                  int out_mask, float* out, int32_t out_ld, void* stream) {
  if (!pts || !desc) return NSFF_EINVAL;
  if (pts->B == 0) return NSFF_OK;                       // empty batch: data pointers of empty tensors are NULL
  if (pts->B > 0 && (!pts->xyz || pts->S < 1)) return NSFF_EINVAL;
  if ((out_mask & NSFF_OUT_TRANSIENT) && !pts->t_table) return NSFF_EINVAL;
  if ((out_mask & NSFF_OUT_STATIC) && !(out_mask & NSFF_OUT_SIGMA_ONLY) && desc->use_viewdir && !pts->dirs)
    return NSFF_EINVAL;
  if (pts->t_table && ((uintptr_t)pts->t_table & 15u)) return NSFF_EINVAL;
  if ((out_mask & NSFF_OUT_STATIC) && !(out_mask & NSFF_OUT_SIGMA_ONLY) && desc->use_viewdir && desc->in_a > 0 &&
      (!pts->a_table || ((uintptr_t)pts->a_table & 15u)))
    return NSFF_EINVAL;
  FwdParams P;
  memset(&P, 0, sizeof(P));
  P.pts = *pts;
  P.x_emb = nullptr;
  return launch_forward(desc, pack, P, out_mask, pts->B, out, out_ld, stream);
}

extern "C" size_t nsff_field_save_bytes(const nsff_field_desc* desc, int64_t B, int out_mask) {
  if (!desc || B < 0 || (out_mask & NSFF_OUT_SIGMA_ONLY)) return 0;
  return (size_t)save_layout(desc, B, out_mask).total;
}

extern "C" int nsff_field_forward_save(const nsff_field_desc* desc, const void* pack, const nsff_points* pts,
                                       int out_mask, float* out, int32_t out_ld, void* save, void* stream) {
  if (!pts || !desc) return NSFF_EINVAL;
  if (pts->B == 0) return NSFF_OK;
  if (!save) return NSFF_EINVAL;
  if (pts->B > 0 && (!pts->xyz || pts->S < 1)) return NSFF_EINVAL;
  if ((out_mask & NSFF_OUT_TRANSIENT) && !pts->t_table) return NSFF_EINVAL;
  if ((out_mask & NSFF_OUT_STATIC) && desc->use_viewdir && !pts->dirs) return NSFF_EINVAL;
  if (pts->t_table && ((uintptr_t)pts->t_table & 15u)) return NSFF_EINVAL;
  if ((out_mask & NSFF_OUT_STATIC) && desc->use_viewdir && desc->in_a > 0 &&
      (!pts->a_table || ((uintptr_t)pts->a_table & 15u)))
    return NSFF_EINVAL;
  if ((uintptr_t)save & 15u) return NSFF_EINVAL;
  FwdParams P;
  memset(&P, 0, sizeof(P));
  P.pts = *pts;
  P.x_emb = nullptr;
  return launch_forward(desc, pack, P, out_mask, pts->B, out, out_ld, stream, save);
}

extern "C" int nsff_field_forward_embedded(const nsff_field_desc* desc, const void* pack, const float* x,
                                           int32_t x_ld, int64_t B, int out_mask, float* out, int32_t out_ld,
                                           void* stream) {
  if (!desc || (B > 0 && !x)) return NSFF_EINVAL;
  if (B == 0) return NSFF_OK;
  FwdParams P;
  memset(&P, 0, sizeof(P));
  P.pts.B = B;
  P.pts.S = 1;
  P.x_emb = x;
  P.x_ld = x_ld;
  // column offsets of the embedded row (nerf.py:142-160); sigma_only rows carry no dir/a columns
  if (out_mask & NSFF_OUT_SIGMA_ONLY) {
    P.c_dir = 0;
    P.c_t = desc->in_xyz;
    if (x_ld < desc->in_xyz + ((out_mask & NSFF_OUT_TRANSIENT) ? desc->in_t : 0)) return NSFF_EINVAL;
  } else {
    P.c_dir = desc->in_xyz;
    P.c_a = desc->in_xyz + desc->in_dir;
    P.c_t = desc->in_xyz + desc->in_dir + desc->in_a;
    if (x_ld < desc->in_xyz + desc->in_dir + desc->in_a + ((out_mask & NSFF_OUT_TRANSIENT) ? desc->in_t : 0))
      return NSFF_EINVAL;
  }
  return launch_forward(desc, pack, P, out_mask, B, out, out_ld, stream);
}
