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
//       next layer are ready when half 1 completes: no accumulator-drain bubble between layers.  The K-chunks both halves
//       read first (prologue chunks, channels 0-127 of the previous layer) are issued once, as N = 256 MMAs into the two
//       adjacent accumulators.
//   smem  A operand : only the "extra" chunks written once per tile by the prologue: PE(xyz) (63 -> 64 cols), time code
//                     (48), PE(dir) (27 -> 32) [+ appearance code (48)]  (128 rows x 64 bf16, 128B-swizzled, K-major)
//   smem  B operand : 10-stage ring of weight pieces ((N/2) x 64 bf16 per CTA) streamed from the L2-resident field pack
//                     with cp.async.bulk + mbarrier complete_tx
//   warps 0-15      : epilogue (4 TMEM lane quadrants x 4 column parts)
//   warps 18-19     : prologue: encode the NEXT tile's inputs into the extra chunks as soon as the issuer's commit says
//                     the current tile's last MMA that reads them has completed - concurrently with the epilogues
//   warp  16        : weight + bias producer (one elected lane)
//   warp  17        : leader CTA: tcgen05.mma issuer (one elected lane walks a flat, host-built list of K-chunk ops);
//                     peer CTA: relays "my weight half has landed" to the leader's barriers
//   warp  20        : training kernels only (kSave): ships the staged tile images with cp.async.bulk stores
//
// The same kernel, instantiated with kBwd, runs the BACKWARD chain of a tile (launch_backward_chain, called by
// nsff_field_backward_tc): dZ of a layer is the "hidden activation" kept in tensor memory, W^T blocks are streamed, the
// epilogue masks with the forward's ReLU bits, two N = 64 accumulators produce the trunk-input gradient rows.
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
#include "nsff_field_host.cuh"
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
  L_T1, L_T2, L_T3, L_T4, L_T5, L_T6, L_T7, L_T8, L_TFIN, L_THEAD,
  // inference-only blocks with the activation-free 256x256 "final" layer folded in (see build_layout)
  L_SDIRM, L_SRGBM, L_THEADM, L_COUNT
};
// EPI_MASK_HID / EPI_ROWS: backward chain (field_forward_kernel<.., kBwd = true>): hidden = dZ of a layer, zeroed where
// the forward's ReLU was inactive (bit plane written by the saving forward); rows = 64 fp32 columns per point to global
enum Epi : uint8_t { EPI_RELU_HID = 0, EPI_LIN_HID = 1, EPI_HEAD = 2, EPI_MASK_HID = 3, EPI_ROWS = 4 };
enum HeadAct : uint8_t { HA_NONE = 0, HA_SIGMOID = 1, HA_FLOW = 2 };

constexpr int kMaxChunks = 6;

struct PackLayer {
  bool present;
  uint32_t w_off;      // byte offset of the first weight block
  uint32_t bias_off;   // float index into the bias table
  int n;               // 256 or 16 (backward chain: 64)
  int nchunks;
  uint8_t chunk[kMaxChunks];
  // generalisations used by the backward chain (0 = the forward's defaults)
  uint8_t blk[kMaxChunks];   // block index of K-chunk kc (default: kc)
  bool blk_explicit;
  int block_rows;            // rows of one packed block (default: n)
  int row0;                  // first row of the block this layer uses
};

struct PackLayout {
  PackLayer layer[L_COUNT];
  uint32_t bias_base;   // byte offset of the fp32 bias table
  uint32_t n_bias;      // floats
  size_t bytes;
};
// fp32 scratch behind the image (and the job table) for the folded products: dir layer (256 x 331 + 256), transient
// heads (16 x 256 + 16), static rgb head without view directions (16 x 256 + 16)
constexpr size_t kMergeScratchFloats = 256 * 331 + 256 + 2 * (16 * 256 + 16);


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
  // Folded blocks (inference).  xyz_encoding_final has no activation (nerf.py:84,107), so the layer after it can
  // absorb it exactly:  W2 (Wf h + bf) + b2 = (W2 Wf) h + (b2 + W2 bf).  The static dir layer (or, without view
  // directions, the rgb head) and the transient head block then read the trunk's last hidden layer directly: one
  // 256x256 layer less per branch (65,536 of 630,016 / 583,680 MACs per point).  nsff_field_pack() computes the
  // products in fp32 on the device and rounds them to bf16 once, like every other weight.
  if (d.use_viewdir) {
    if (d.in_a > 0) add(L_SDIRM, 256, {CH_DIR, CH_A, CH_H0, CH_H1, CH_H2, CH_H3});
    else add(L_SDIRM, 256, {CH_DIR, CH_H0, CH_H1, CH_H2, CH_H3});
  } else {
    add(L_SRGBM, 16, {CH_H3, CH_H2, CH_H1, CH_H0});
  }
  if (d.has_transient) add(L_THEADM, 16, {CH_H3, CH_H2, CH_H1, CH_H0});
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
constexpr int kMaxPackJobs = 200;
struct PackJobs {
  PackJob job[kMaxPackJobs];
};

__device__ __forceinline__ uint32_t swz_off(int row, int col /*bf16 column 0..63*/) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((col >> 3) ^ (row & 7)) & 7) << 4) + (col & 7) * 2);
}

__global__ void pack_kernel(const __grid_constant__ PackJobs jobs, uint8_t* __restrict__ pack, uint32_t bias_base) {
  const PackJob& j = jobs.job[blockIdx.x];
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


// out[n, k] = sum_j A[n, j] * B[j, k]   (n < rows, j < 256, k < 256),  out_bias[n] = a_bias[n] + sum_j A[n, j] * b_bias[j]:
// a linear layer (B, b_bias) folded into the rows of the layer after it (A, a_bias); fp32, one block per output row.
struct MergeJob {
  const float* A;        // (rows, lda): the columns that multiply the folded layer's output
  const float* a_bias;   // (rows)
  const float* B;        // (256, 256) row-major (out, in) weight of the folded layer
  const float* b_bias;   // (256)
  float* out;            // (rows, ldo)
  float* out_bias;       // (rows)
  int32_t lda, ldo, rows;
};
constexpr int kMaxMergeJobs = 8;
struct MergeJobs {
  MergeJob job[kMaxMergeJobs];
};
__global__ void __launch_bounds__(256) merge_linear_kernel(const __grid_constant__ MergeJobs JJ) {
  const MergeJob& J = JJ.job[blockIdx.y];
  const int n = blockIdx.x, k = threadIdx.x;
  if (n >= J.rows) return;
  __shared__ float a_row[256];
  __shared__ float red[256];
  a_row[k] = J.A[(int64_t)n * J.lda + k];
  __syncthreads();
  float acc = 0.f;
#pragma unroll 8
  for (int j = 0; j < 256; ++j) acc = fmaf(a_row[j], J.B[j * 256 + k], acc);
  J.out[(int64_t)n * J.ldo + k] = acc;
  red[k] = a_row[k] * J.b_bias[k];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (k < s) red[k] += red[k + s];
    __syncthreads();
  }
  if (k == 0) J.out_bias[n] = J.a_bias[n] + red[0];
}

// ---------------------------------------------------------------------------------------------
// forward kernel
// ---------------------------------------------------------------------------------------------
constexpr int kTileM = 128;
constexpr int kEpiWarps = 16;        // 4 TMEM lane quadrants x 4 column parts
constexpr int kPrologueWarps = 2;       // dedicated input-encoding warps (18 -> 20 warps: same 96-register allocation granule)
constexpr int kSaverWarp = kEpiWarps + 2 + kPrologueWarps;   // training kernels: issues the bulk stores of the saved images
constexpr int kThreads = kSaverWarp * 32;               // inference: 20 warps = 5 per scheduler, 96 registers each
constexpr int kThreadsSave = (kSaverWarp + 1) * 32;     // training kernels: + the saver warp (6 warps on one scheduler: 80 registers)
constexpr uint32_t kChunkBytes = 128 * 128;          // 128 rows x 64 bf16
constexpr uint32_t kSmemBias = 21 * 1024;            // the module's whole fp32 bias table (<= 5200 floats), resident
constexpr uint32_t kSmemBar = 512;
constexpr int kMaxStages = 18;
constexpr int kMaxSubs = 44;                         // sub-steps per tile (main pass, split: 19 x 2 + 3 heads = 41)
constexpr int kMaxOps = 176;                         // ops per tile (main pass, one K-chunk per op: 164 with the appearance chunk)
constexpr int kAccIds = 6;                           // accumulators with their own full/empty barriers: X.h0 X.h1 Y.h0 Y.h1, narrow 0 / 1
constexpr int kHeadAcc = 4;

#ifdef NSFF_TIMELINE
constexpr uint32_t kSmemDbg = 3 * 512 * 8;    // shared-memory event log (tools/timeline.py)
#else
constexpr uint32_t kSmemDbg = 0;
#endif

// kAppear: one more prologue chunk (the 48-wide NeRF-W appearance code feeding the dir layer), paid for with ring space.
// The ring is cut into stages of TWO weight pieces (the two K-chunks of one op): 2 x 16 KB per CTA for N = 256 ops
// 2 x 8 KB per CTA for the N = 128 sub-steps (9 stages; 8 with the appearance chunk).
// kSave (training forward, backward chain): every hidden sub-step also leaves the SM as a 2 x 16 KB tile image.  The
// epilogue warps write their rows into one of two shared-memory staging buffers and a dedicated warp ships each buffer
// with two cp.async.bulk stores (whole 128-byte lines, no LSU work); per-lane 32-byte global stores cost the epilogue
// warps 1600-2100 cycles per sub-step (in-kernel timeline) - more than the drain itself.  The buffers are paid for with
// four ring stages.
template <bool kAppear, bool kSave = false>
struct Cfg {
  static constexpr uint32_t kSmemA = (kAppear ? 4 : 3) * kChunkBytes;   // PE + T + DIR (+ A)
  static constexpr uint32_t kSmemB = ((kAppear ? 8 : 9) - (kSave ? 4 : 0)) * 16384;
  static constexpr uint32_t kSmemSave = kSave ? 2 * 2 * kChunkBytes : 0;
  static constexpr uint32_t kSmemTotal = kSmemA + kSmemB + kSmemSave + kSmemBias + kSmemBar + kSmemDbg + 1024;   // + alignment slack
};

enum : uint8_t { OP_TS = 1, OP_FIRST = 2, OP_LAST = 4, OP_ADV_HID = 8, OP_TWO = 16, OP_EXTRA_DONE = 32 };

// One or two K-chunks (4 tcgen05.mma of K = 16 each) of one sub-step: what the issuer issues (first 16 bytes, one
// constant load) and what the producer copies (second 16 bytes).
struct alignas(16) ChunkOp {
  uint32_t idesc;      // instruction descriptor (M = 256, this sub-step's N)
  uint16_t a;          // OP_TS: TMEM column of K-step 0 of the first chunk (+16 per K-step); else smem offset / 16
  uint16_t a2;         // same for the second chunk (OP_TWO)
  uint16_t d_col;      // TMEM column of the accumulator
  uint16_t b2;         // descriptor units (16 B) between the two chunks' weight pieces inside the stage
  uint8_t flags;       // OP_*
  uint8_t wait_hid;    // 1 + hidden-half barrier to wait for before issuing (0: none; only the first use in a layer)
  uint8_t acc_id;
  uint8_t wait_empty;  // OP_FIRST: accumulators (bit mask) whose drain is not implied by an earlier hidden-half wait
  uint32_t w_src0;     // byte offset in the pack of the leader CTA's rows of the first weight piece (peer: + bytes)
  uint32_t w_src1;     // ... of the second piece (OP_TWO)
  uint32_t bytes;      // bytes per piece per CTA = (N / 2) * 128
  uint32_t pad1;
};
static_assert(sizeof(ChunkOp) == 32, "ChunkOp: 16 bytes for the issuer + 16 bytes for the producer");

// One accumulator's worth of work for the epilogue warps: one N = 128 half of a 256-wide layer, or a head block.
struct alignas(16) SubStep {
  // hot fields, read with one 8-byte constant load
  uint16_t d_col;          // TMEM column of the accumulator (and of the in-place bf16 image)
  uint8_t acc_id;
  uint8_t epi;
  uint16_t bias_off;       // float index of this sub-step's first bias in the (shared-memory resident) bias table
  uint8_t slice0;          // first 64-column slice (= hidden chunk) this sub-step produces: 0 or 2
  uint8_t arrive_empty;    // some op waits for this accumulator to be drained: arrive on its "empty" barrier
  uint8_t head_tbl[16];    // heads: per column dst | act << 6 (0xFF: column not written)
  uint64_t save_off;       // training mode: byte offset of this layer's activation image in the save buffer (kNoSave: none)
  uint64_t bits_off;       // ... and of its ReLU bit plane
  uint16_t n;              // 128 (hidden) or 16 (head)
  uint8_t pad[6];
  // backward chain: byte offsets (from the per-thread base pointer into the forward's bit planes) of this sub-step's
  // ReLU bits and of the bits to prefetch now for the next masked sub-step (kNoSave: none)
  uint64_t mask_off, next_mask_off;
};
static_assert(sizeof(SubStep) == 64, "SubStep layout");

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
  int32_t n_bias;           // floats in the bias table
  int32_t last_extra_sub;   // last sub-step whose K-chunks include PE / time code / PE(dir): once its accumulator is
                            // complete the prologue chunks are dead and the NEXT tile's prologue can be written early
  uint32_t stage_bytes;     // weight-ring stage (two pieces): 32768, or 16384 in split mode
  int32_t n_stages;
  float flow_scale;
  uint8_t* save_base;       // training mode (kSave kernels): activation save buffer, see nsff_field_save.cuh
  // backward chain (kBwd kernels): head pre-activation gradients as bf16 chunk images ([tile][128 rows x 128 B], columns
  // 0..15 used) for the static / transient branch, and the base of the forward's save buffer (ReLU bit planes)
  const uint8_t* dp_img[2];
  const uint8_t* mask_base;
  unsigned long long* dbg;  // NSFF_TIMELINE builds only: (tag, clock64) event log of CTA 0
  SubStep subs[kMaxSubs];
  ChunkOp ops[kMaxOps];
};

// Timeline instrumentation (compiled in only with -DNSFF_TIMELINE; see tools/timeline.py).  Events go to a shared-memory
// log (8 bytes: tag << 40 | clock) and are copied to global memory when the role ends, so that logging neither stalls on
// global stores nor drags them into the release semantics of the next mbarrier arrive.
#ifdef NSFF_TIMELINE
constexpr int kDbgPerRole = 512;
static_assert(kSmemDbg == 3 * kDbgPerRole * 8, "log size");
#define NSFF_TL_DECL(role) const uint32_t tl_base = sDbg + (role) * kDbgPerRole * 8; const int tl_role = (role); int tl_n = 0; const bool tl_on = P.dbg && blockIdx.x == 0
#define NSFF_TL(tag)                                                                                       \
  do {                                                                                                     \
    if (tl_on && tl_n < kDbgPerRole && lane == 0) {                                                        \
      const unsigned long long tl_v = ((unsigned long long)(tag) << 40) | ((unsigned long long)clock64() & 0xFFFFFFFFFFull); \
      asm volatile("st.shared.b64 [%0], %1;" ::"r"(tl_base + 8u * (uint32_t)tl_n), "l"(tl_v) : "memory");   \
      ++tl_n;                                                                                              \
    }                                                                                                      \
  } while (0)
#define NSFF_TL_DUMP()                                                                                     \
  do {                                                                                                     \
    if (tl_on && lane == 0)                                                                                \
      for (int tl_i = 0; tl_i < tl_n; ++tl_i) {                                                            \
        unsigned long long tl_v;                                                                           \
        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(tl_v) : "r"(tl_base + 8u * (uint32_t)tl_i));         \
        P.dbg[tl_role * kDbgPerRole + tl_i] = tl_v;                                                        \
      }                                                                                                    \
  } while (0)
#else
#define NSFF_TL_DECL(role) do {} while (0)
#define NSFF_TL(tag) do {} while (0)
#define NSFF_TL_DUMP() do {} while (0)
#endif

// barrier slots (8 bytes each) inside the barrier region
enum {
  BAR_W_FULL = 0,                           // [kMaxStages] weight stage landed (tx + partner relay)
  BAR_W_EMPTY = BAR_W_FULL + kMaxStages,    // [kMaxStages] MMAs that read the stage completed (commit)
  BAR_HID_FULL = BAR_W_EMPTY + kMaxStages,  // [2]  bf16 image of hidden channels 0-127 / 128-255 written to TMEM (both CTAs)
  BAR_EXTRA_FULL = BAR_HID_FULL + 2,        // [1]  prologue chunks written (prologue warps of both CTAs)
  BAR_EXTRA_EMPTY = BAR_EXTRA_FULL + 1,     // [1]  every MMA that reads the prologue chunks of this tile has completed (commit)
  BAR_ACC_FULL = BAR_EXTRA_EMPTY + 1,       // [kAccIds] accumulator complete (commit)
  BAR_ACC_EMPTY = BAR_ACC_FULL + kAccIds,   // [kAccIds] accumulator drained
  BAR_SAVE_FULL = BAR_ACC_EMPTY + kAccIds,  // [2]  kSave: every epilogue warp of this CTA has written its rows into the staging buffer
  BAR_SAVE_EMPTY = BAR_SAVE_FULL + 2,       // [2]  kSave: the bulk stores have read the staging buffer
  BAR_COUNT = BAR_SAVE_EMPTY + 2
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

// One row of the octave encoding [x, sin(2^0 x), cos(2^0 x), ..., sin(2^(F-1) x), cos(2^(F-1) x)] (nerf.py:17-30) into the
// first 3+6F columns of a swizzled 64-column chunk (remaining columns of the last touched unit zero): one accurate
// sincosf at octave 0, then the doubling recurrence (pe_next_octave, nsff_common.cuh).
template <int kFreqs, int kUnits>
__device__ __forceinline__ void encode_octaves(uint32_t chunk, int row, const float (&x)[3]) {
  float sn[3], cs[3], v[kUnits * 8];
#pragma unroll
  for (int i = 0; i < kUnits * 8; ++i) v[i] = i < 3 ? x[i % 3] : 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) sincosf(x[c], &sn[c], &cs[c]);
#pragma unroll
  for (int f = 0; f < kFreqs; ++f) {
#pragma unroll
    for (int r = 0; r < 6; ++r) v[3 + 6 * f + r] = r < 3 ? sn[r % 3] : cs[r % 3];
    if (f + 1 < kFreqs) pe_next_octave(sn, cs);
  }
#pragma unroll
  for (int u = 0; u < kUnits; ++u) {
    const float w[8] = {v[8 * u], v[8 * u + 1], v[8 * u + 2], v[8 * u + 3], v[8 * u + 4], v[8 * u + 5], v[8 * u + 6], v[8 * u + 7]};
    store_unit(chunk, row, u, w);
  }
}

// 48 code floats (time / appearance) of one row -> units 0..5 of a chunk
__device__ __forceinline__ void encode_code48(uint32_t chunk, int row, const float* src) {
  const float4* tp = reinterpret_cast<const float4*>(src);
#pragma unroll
  for (int u = 0; u < 6; ++u) {
    const float4 a = __ldg(tp + 2 * u), b = __ldg(tp + 2 * u + 1);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    store_unit(chunk, row, u, v);
  }
}

// the position of point p (+ the flow offset of the legacy warp-pass interface): loaded separately so that the prologue
// warps can have the loads of all their rows in flight before they start encoding
__device__ __forceinline__ void load_point(const FwdParams& P, int64_t p, float (&base)[3]) {
  if (P.x_emb != nullptr) return;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v = P.pts.xyz[p * 3 + c];
    if (P.pts.offset) v = __fadd_rn(v, P.pts.offset[p * P.pts.offset_ld + c]);
    base[c] = v;
  }
}

// Prologue of one tile row: encode the inputs of point p into the PE / time-code / PE(dir) [/ appearance] chunks.
__device__ __forceinline__ void encode_tile_row(const FwdParams& P, uint32_t sA, int row, int64_t p, bool valid,
                                                const float (&base)[3]) {
  const uint32_t chPE = sA + 0 * kChunkBytes, chT = sA + 1 * kChunkBytes, chD = sA + 2 * kChunkBytes,
                 chA = sA + 3 * kChunkBytes;
  if (P.x_emb == nullptr) {
    const int64_t ray = p / P.pts.S;
    if (valid && P.pts.xyz_out) {
#pragma unroll
      for (int c = 0; c < 3; ++c) P.pts.xyz_out[p * 3 + c] = base[c];
    }
    encode_octaves<10, 8>(chPE, row, base);                  // 63 columns + the zero pad
    if (P.need_t) {
      int64_t trow = ray;
      if (P.pts.t_index) {
        trow = P.pts.t_index[ray] + P.pts.t_shift;
        trow = trow < 0 ? 0 : (trow > P.pts.max_t ? P.pts.max_t : trow);
      }
      encode_code48(chT, row, P.pts.t_table + trow * 48);
    }
    if (P.need_a) encode_code48(chA, row, P.pts.a_table + (P.pts.a_index ? P.pts.a_index[ray] : ray) * 48);
    if (P.need_dir) {
      const float d[3] = {P.pts.dirs[ray * 3 + 0], P.pts.dirs[ray * 3 + 1], P.pts.dirs[ray * 3 + 2]};
      encode_octaves<4, 4>(chD, row, d);                     // 27 columns, 5 zeros
    }
  } else {
    // embedded mode: x row = [xyz63 | dir27 | a | t48] already encoded by the caller
    const float* xr = P.x_emb + p * P.x_ld;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = u * 8 + i < 63 ? xr[u * 8 + i] : 0.f;
      store_unit(chPE, row, u, v);
    }
    if (P.need_t) {
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = xr[P.c_t + u * 8 + i];
        store_unit(chT, row, u, v);
      }
    }
    if (P.need_a) {
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = xr[P.c_a + u * 8 + i];
        store_unit(chA, row, u, v);
      }
    }
    if (P.need_dir) {
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

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr)
               : "memory");
}

// two fp32 adds in one instruction (add.rn.f32x2, sm_100): same rounding as two FADDs
__device__ __forceinline__ void add2(float a0, float a1, float b0, float b1, float& o0, float& o1) {
  asm("{\n\t.reg .b64 a, b, d;\n\t"
      "mov.b64 a, {%2, %3};\n\t"
      "mov.b64 b, {%4, %5};\n\t"
      "add.rn.f32x2 d, a, b;\n\t"
      "mov.b64 {%0, %1}, d;\n\t}"
      : "=f"(o0), "=f"(o1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

// 16 accumulator columns of one row: + bias, (ReLU,) round to bf16, pack pairs -> 8 words = one K = 16 step of the next
// layer's A operand.  kBits (training forward): also the 16 ReLU bits (bit j = pre-activation j is not negative: one
// funnel shift per column collects the sign bits - the backward's mask; an exact zero counts as active, its gradient
// contribution meets a zero activation in the weight gradient either way).
template <bool kBits = false>
__device__ __forceinline__ void bias_act_pack(const uint32_t (&v)[16], const float4 (&b)[4], bool relu, uint32_t (&w)[8],
                                              uint32_t* bits = nullptr) {
  float f[16];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    add2(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), b[i].x, b[i].y, f[4 * i], f[4 * i + 1]);
    add2(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]), b[i].z, b[i].w, f[4 * i + 2], f[4 * i + 3]);
  }
  if constexpr (kBits) {
    uint32_t neg = 0;
#pragma unroll
    for (int j = 15; j >= 0; --j) neg = __funnelshift_l(__float_as_uint(f[j]), neg, 1);      // (neg << 1) | sign(f[j])
    *bits = ~neg & 0xFFFFu;
  }
  if (relu) {
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = pack_bf16x2_relu(f[2 * i], f[2 * i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
  }
}

// backward chain: 16 accumulator columns of one row -> bf16 pairs, zeroed where the forward's ReLU bit is clear
// (bit pair b1 b0 -> x = (b1 b0) * 0x4080 puts b0 into the sign bit of byte 0 and b1 into the sign bit of byte 1; prmt
// with sign replication expands them to 0x0000FFFF / 0xFFFF0000 halves: 4 integer instructions per bf16 pair)
__device__ __forceinline__ void mask_pack(const uint32_t (&v)[16], uint32_t bits, uint32_t (&w)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t p = pack_bf16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
    uint32_t y, m;
    asm("bfe.u32 %0, %1, %2, 2;" : "=r"(y) : "r"(bits), "r"(2 * i));
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(m) : "r"(y * 0x4080u), "r"(0u), "r"(0x9988u));
    w[i] = p & m;
  }
}

// tanh through one exp: |error| ~ 1e-7 (tanhf's branchy multi-range code costs ~3x the instructions)
__device__ __forceinline__ float tanh_f(float x) {
  const float e = __expf(-2.f * fabsf(x));
  return copysignf(__fdividef(1.f - e, 1.f + e), x);
}

// kBwd: the same machinery runs the BACKWARD chain of a tile (build_chain_program): the "hidden activation" kept in
// tensor memory is dZ of the current layer, the streamed blocks are W^T, the prologue warps copy the head gradients
// instead of encoding inputs, and the epilogue masks with the forward's ReLU bits instead of adding a bias.
template <bool kAppear, bool kSave, bool kBwd = false>
__global__ void __launch_bounds__(kSave ? kThreadsSave : kThreads, 1) field_forward_kernel(const __grid_constant__ FwdParams P) {
  using C = Cfg<kAppear, kSave>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base;                       // extra chunks
  const uint32_t sB = smem_base + C::kSmemA;           // weight ring
  const uint32_t sSave = sB + C::kSmemB;               // kSave: two staging buffers of two chunks
  const uint32_t sBias = sSave + C::kSmemSave;
  const uint32_t sBar = sBias + kSmemBias;
  const uint32_t sTmemPtr = sBar + BAR_COUNT * 8;
#ifdef NSFF_TIMELINE
  const uint32_t sDbg = sBar + kSmemBar;
#endif
  auto bar = [&](int i) { return sBar + (uint32_t)i * 8u; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  // barriers the MMA issuer waits on live in the leader CTA; both CTAs of the pair arrive there
  // (the leader's own warps arrive locally: a shared::cluster arrive takes the DSMEM round trip even to oneself)
  auto arrive_leader = [&](int i) {
    if (leader) mbar_arrive(bar(i));
    else mbar_arrive_cluster(mapa_cluster(bar(i), 0));
  };
  const int group = blockIdx.x >> 1, n_groups = gridDim.x >> 1;
  const int n_units = (P.n_tiles + 1) >> 1;            // tile pairs
  const int n_stages = P.n_stages;
  const uint32_t stage_bytes = P.stage_bytes;

#ifdef NSFF_WATCHDOG
  if (threadIdx.x == 0 && blockIdx.x == 0)
    printf("NSFF_WATCHDOG field kernel (bwd %d): barriers @%u (8 B each: W_FULL[%d] W_EMPTY[%d] HID[2] XF XE ACC_FULL[%d] ACC_EMPTY[%d]), %d subs %d ops\n",
           (int)kBwd, sBar, kMaxStages, kMaxStages, kAccIds, kAccIds, P.n_subs, P.n_ops);
#endif
  if (threadIdx.x == 0) {
    for (int i = 0; i < n_stages; ++i) {
      // leader: own pieces (producer arrive + tx) AND the partner's relay arrive complete a stage
      mbar_init(bar(BAR_W_FULL + i), leader ? 2 : 1);
      mbar_init(bar(BAR_W_EMPTY + i), 1);
    }
    for (int i = 0; i < 2; ++i) mbar_init(bar(BAR_HID_FULL + i), kEpiWarps * 2);
    mbar_init(bar(BAR_EXTRA_FULL), kPrologueWarps * 2);
    mbar_init(bar(BAR_EXTRA_EMPTY), 1);
    for (int i = 0; i < kAccIds; ++i) {
      mbar_init(bar(BAR_ACC_FULL + i), 1);
      mbar_init(bar(BAR_ACC_EMPTY + i), kEpiWarps * 2);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(BAR_SAVE_FULL + i), kEpiWarps);
      mbar_init(bar(BAR_SAVE_EMPTY + i), 1);
    }
    fence_mbar_init();
  }
  if (warp == kEpiWarps + 1) tmem_alloc_pair(sTmemPtr, 512);
  if (warp < kEpiWarps) {
    // the extra chunks are zero-filled once: the prologue only ever writes their used columns (48 of 64 for the time /
    // appearance code, 32 for PE(dir)), while every chunk is issued as four K = 16 steps against zero-padded weights
    for (uint32_t o = threadIdx.x * 16u; o < C::kSmemA; o += kEpiWarps * 32u * 16u)
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sA + o), "r"(0u) : "memory");
    // the module's bias table stays in shared memory for the whole kernel (no per-layer copies, no barriers)
    if constexpr (!kBwd)
      for (int i = threadIdx.x; i < P.n_bias; i += kEpiWarps * 32)
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(sBias + 4u * (uint32_t)i), "f"(__ldg(P.bias + i)) : "memory");
  }
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
      for (int i = 0; i < n_ops; ++i) {
        const uint4 raw = reinterpret_cast<const uint4*>(&P.ops[i])[0];
        const uint4 src = reinterpret_cast<const uint4*>(&P.ops[i])[1];      // w_src0, w_src1, bytes
        const uint32_t flags = raw.w & 0xFFu;
        if (flags & OP_FIRST) ++gsub;
        mbar_wait_sleep(bar(BAR_W_EMPTY + stage), epar);
        NSFF_TL(((gsub - 1) << 8) | 1);                    // producer: stage free, issuing copy
        if (elect_one()) {
          const uint32_t bytes = src.z, two = (flags & OP_TWO) ? 1u : 0u;
          const uint32_t dst = sB + stage * stage_bytes, full = bar(BAR_W_FULL + stage);
          mbar_arrive_expect_tx(full, bytes << two);
          bulk_g2s(dst, P.pack + src.x + cta_rank * bytes, bytes, full);
          if (two) bulk_g2s(dst + bytes, P.pack + src.y + cta_rank * bytes, bytes, full);
        }
        if (++stage == (uint32_t)n_stages) stage = 0, epar ^= 1u;
      }
    }
    NSFF_TL_DUMP();
  } else if (warp == kEpiWarps + 1) {
    if (leader) {
      // ===================== MMA issuer =====================
      // One elected lane walks the flat op list (the other 31 park at the reconvergence point).  The loop body is the
      // serial critical path of the kernel - one instruction stream feeds the tensor pipe - so an op carries up to
      // eight MMAs (two K-chunks), everything that can be precomputed lives in its first 16 bytes, every chunk is four
      // K-steps (no data-dependent branches around the MMAs), and the barrier probes of op i+1 are issued before the
      // MMAs of op i so that their round trip overlaps the issue work.
      const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);          // SBO | version | SWIZZLE_128B
      const uint32_t a_lo0 = ((sA & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t b_lo0 = ((sB & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t stage_units = stage_bytes >> 4;
      auto desc = [&](uint32_t lo) { return ((uint64_t)desc_hi << 32) | (uint64_t)lo; };
      if (elect_one()) {
        uint32_t stage = 0, wpar = 0, hid_gen = 0, unit_iter = 0, gsub = 0;
        uint32_t fill_par = 0;             // bit a: (number of fills of accumulator a so far) & 1
        NSFF_TL_DECL(1);
        for (int unit = group; unit < n_units; unit += n_groups, ++unit_iter) {
          mbar_wait(bar(BAR_EXTRA_FULL), unit_iter & 1u);
          NSFF_TL((gsub << 8) | 15);                           // mma: prologue chunks ready
          // Software pipeline: while the tensor pipe works through the first MMA group of op i, the stage of op i-1 is
          // released, op i+1 is fetched and decoded and its barriers are probed; between the last MMA of op i and the
          // first MMA of op i+1 only the probe results are looked at.
          uint4 op = reinterpret_cast<const uint4*>(&P.ops[0])[0];
          uint32_t wh = (op.w >> 8) & 0xFFu;
          bool hid_ok = wh == 0 || mbar_test(bar(BAR_HID_FULL + wh - 1), (hid_gen - 1u) & 1u);
          bool w_ok = mbar_test(bar(BAR_W_FULL + stage), wpar);
          uint32_t pending_empty = 0;                          // W_EMPTY barrier of the previous op, committed late
          for (int i = 0; i < n_ops; ++i) {
            const uint4 cur = op;
            const uint32_t flags = cur.w & 0xFFu, acc = (cur.w >> 16) & 0xFFu;
            uint32_t we = cur.w >> 24;
            while (we) {
              // accumulators this sub-step overwrites whose drain no earlier wait implies (head accumulators; rare)
              const uint32_t a = (uint32_t)__ffs((int)we) - 1u;
              we &= we - 1u;
              mbar_wait(bar(BAR_ACC_EMPTY + a), ((fill_par >> a) & 1u) ^ 1u);
            }
            if (!hid_ok) mbar_wait(bar(BAR_HID_FULL + wh - 1), (hid_gen - 1u) & 1u);
            if (!w_ok) mbar_wait(bar(BAR_W_FULL + stage), wpar);
            NSFF_TL((gsub << 8) | 4);                          // mma: operands ready -> issue
            const uint32_t b_lo = b_lo0 + stage * stage_units;
            const uint32_t empty_bar = bar(BAR_W_EMPTY + stage);
            if (++stage == (uint32_t)n_stages) stage = 0, wpar ^= 1u;
            hid_gen += (flags & OP_ADV_HID) ? 1u : 0u;
            tc_fence_after_sync();
            const uint32_t idesc = cur.x;
            const uint32_t d_tmem = tmem_base + (cur.z & 0xFFFFu);
            const uint32_t acc0 = (flags & OP_FIRST) ? 0u : 1u;
            const uint32_t b_lo2 = b_lo + (cur.z >> 16);
            const bool ts = flags & OP_TS, two = flags & OP_TWO;
            const uint32_t a1 = (ts ? tmem_base : a_lo0) + (cur.y & 0xFFFFu), a2 = (ts ? tmem_base : a_lo0) + (cur.y >> 16);
            if (ts) {
              mma_bf16_ts_pair(d_tmem, a1, desc(b_lo), idesc, acc0);           // 8 columns (16 bf16) per K-step, 16 apart
              mma_bf16_ts_pair(d_tmem, a1 + 16, desc(b_lo + 2), idesc, 1u);
              mma_bf16_ts_pair(d_tmem, a1 + 32, desc(b_lo + 4), idesc, 1u);
              mma_bf16_ts_pair(d_tmem, a1 + 48, desc(b_lo + 6), idesc, 1u);
            } else {
              mma_bf16_ss_pair(d_tmem, desc(a1), desc(b_lo), idesc, acc0);     // +32 B per K-step inside the 128 B swizzle row
              mma_bf16_ss_pair(d_tmem, desc(a1 + 2), desc(b_lo + 2), idesc, 1u);
              mma_bf16_ss_pair(d_tmem, desc(a1 + 4), desc(b_lo + 4), idesc, 1u);
              mma_bf16_ss_pair(d_tmem, desc(a1 + 6), desc(b_lo + 6), idesc, 1u);
            }
            // the previous op's weight stage (its MMAs precede the ones just issued: released at most four MMAs late)
            if (pending_empty) mma_commit_pair_mc(pending_empty, 3);
            // next op: fetch, probe (the tensor pipe is working through the MMAs just queued)
            const int i_next = i + 1 < n_ops ? i + 1 : i;
            op = reinterpret_cast<const uint4*>(&P.ops[i_next])[0];
            wh = (op.w >> 8) & 0xFFu;
            hid_ok = wh == 0 || mbar_test(bar(BAR_HID_FULL + wh - 1), (hid_gen - 1u) & 1u);
            w_ok = mbar_test(bar(BAR_W_FULL + stage), wpar);
            if (two) {
              if (ts) {
                mma_bf16_ts_pair(d_tmem, a2, desc(b_lo2), idesc, 1u);
                mma_bf16_ts_pair(d_tmem, a2 + 16, desc(b_lo2 + 2), idesc, 1u);
                mma_bf16_ts_pair(d_tmem, a2 + 32, desc(b_lo2 + 4), idesc, 1u);
                mma_bf16_ts_pair(d_tmem, a2 + 48, desc(b_lo2 + 6), idesc, 1u);
              } else {
                mma_bf16_ss_pair(d_tmem, desc(a2), desc(b_lo2), idesc, 1u);
                mma_bf16_ss_pair(d_tmem, desc(a2 + 2), desc(b_lo2 + 2), idesc, 1u);
                mma_bf16_ss_pair(d_tmem, desc(a2 + 4), desc(b_lo2 + 4), idesc, 1u);
                mma_bf16_ss_pair(d_tmem, desc(a2 + 6), desc(b_lo2 + 6), idesc, 1u);
              }
            }
            pending_empty = empty_bar;
            if (flags & OP_EXTRA_DONE) mma_commit_pair_mc(bar(BAR_EXTRA_EMPTY), 3);   // prologue chunks are dead: next tile's may be written
            if (flags & OP_LAST) {
              mma_commit_pair_mc(bar(BAR_ACC_FULL + acc), 3);  // (the epilogue is waiting for this one: not deferred)
              fill_par ^= 1u << acc;
              NSFF_TL((gsub << 8) | 11);                       // mma: sub-step committed
              ++gsub;
            }
          }
          if (pending_empty) mma_commit_pair_mc(pending_empty, 3);
        }
        NSFF_TL_DUMP();
      }
      __syncwarp();
    } else {
      // ===================== peer CTA: relay "my weight pieces have landed" to the leader ==============
      uint32_t stage = 0, wpar = 0;
      const uint32_t leader_w_full = mapa_cluster(bar(BAR_W_FULL), 0);
      for (int unit = group; unit < n_units; unit += n_groups) {
        for (int i = 0; i < n_ops; ++i) {
          mbar_wait_sleep(bar(BAR_W_FULL + stage), wpar);
          if (lane == 0) mbar_arrive_cluster(leader_w_full + stage * 8u);
          if (++stage == (uint32_t)n_stages) stage = 0, wpar ^= 1u;
        }
      }
    }
  } else if (warp == kSaverWarp) {
    // ===================== saved-image store issuer (training kernels) =====================
    if constexpr (kSave) {
      if (elect_one()) {
        uint32_t cnt = 0;
        for (int unit = group; unit < n_units; unit += n_groups) {
          const int tile = unit * 2 + (int)cta_rank;
          for (int s = 0; s < n_subs; ++s) {
            const SubStep& ss = P.subs[s];
            if (ss.epi == EPI_HEAD || ss.epi == EPI_ROWS || ss.save_off == kNoSave) continue;
            const uint32_t buf = cnt & 1u;
            mbar_wait_sleep(bar(BAR_SAVE_FULL + buf), (cnt >> 1) & 1u);
            if (tile < P.n_tiles) {
              uint8_t* g = P.save_base + ss.save_off + (size_t)tile * kSaveImgPerTile + (size_t)ss.slice0 * kChunkBytes;
              bulk_s2g(g, sSave + buf * 2u * kChunkBytes, 2u * kChunkBytes);      // chunks slice0, slice0 + 1 are adjacent
            }
            bulk_commit_group();
            // the previous group has been read out of its buffer once at most this one is still reading
            bulk_wait_group_read<1>();
            if (cnt > 0) mbar_arrive(bar(BAR_SAVE_EMPTY + (buf ^ 1u)));
            ++cnt;
          }
        }
        bulk_wait_group_read<0>();
        if (cnt > 0) mbar_arrive(bar(BAR_SAVE_EMPTY + ((cnt - 1u) & 1u)));
        bulk_wait_group<0>();              // the stores have been performed: the CTA may exit
      }
      __syncwarp();
    }
  } else if (warp >= kEpiWarps + 2) {
    // ===================== prologue warps =====================
    // Encode a tile's inputs into the extra chunks (bf16, swizzled): 128 rows over 64 threads, one whole row at a time.
    // The first tile is encoded at once; tile i+1 as soon as the issuer's commit on EXTRA_EMPTY reports that the last
    // MMA of tile i that reads these chunks has completed - while the epilogue warps carry on with the remaining layers.
    const int t64 = (warp - (kEpiWarps + 2)) * 32 + lane;
    uint32_t it = 0;
    for (int unit = group; unit < n_units; unit += n_groups, ++it) {
      if (it > 0) mbar_wait_sleep(bar(BAR_EXTRA_EMPTY), (it - 1u) & 1u);
      if constexpr (kBwd) {
        // head gradients of this tile: K-step 0 (two 16-byte units per row) of the static / transient chunk images;
        // the other units of the chunks stay zero (zero-filled above, never written)
        const int tile = unit * 2 + (int)cta_rank;
        if (tile < P.n_tiles) {
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            if (!P.dp_img[b]) continue;
            const uint8_t* src = P.dp_img[b] + (size_t)tile * kChunkBytes;
            uint4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int idx = j * 64 + t64, row = idx >> 1, u = idx & 1;
              v[j] = __ldg(reinterpret_cast<const uint4*>(src + (row >> 3) * 1024 + (row & 7) * 128 + (((u ^ row) & 7) << 4)));
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int idx = j * 64 + t64, row = idx >> 1, u = idx & 1;
              st_shared_v4(sA + (uint32_t)b * kChunkBytes + (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + (((u ^ row) & 7) << 4)),
                           v[j].x, v[j].y, v[j].z, v[j].w);
            }
          }
        }
      } else {
        constexpr int kRows = kTileM / (kPrologueWarps * 32);             // 2 rows per thread
        float base[kRows][3];
#pragma unroll
        for (int j = 0; j < kRows; ++j) {                                 // all position loads in flight first
          const int64_t pr = (int64_t)(unit * 2 + (int)cta_rank) * kTileM + j * (kPrologueWarps * 32) + t64;
          load_point(P, pr < P.pts.B ? pr : P.pts.B - 1, base[j]);
        }
        for (int j = 0; j < kRows; ++j) {
          const int row = j * (kPrologueWarps * 32) + t64;
          const int64_t pr = (int64_t)(unit * 2 + (int)cta_rank) * kTileM + row;
          const bool ok = pr < P.pts.B;
          encode_tile_row(P, sA, row, ok ? pr : P.pts.B - 1, ok, base[j]);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) arrive_leader(BAR_EXTRA_FULL);
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3, part = warp >> 2;      // TMEM lane quadrant, 16-column part of every 64-column slice
    const int row = q * 32 + lane;
    const uint32_t lane_field = (uint32_t)(q * 32) << 16;
    const uint32_t row_off = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
    const int rsw = row & 7;
    uint32_t gsub = 0;
    uint32_t drain_par = 0;                        // bit a: (number of drains of accumulator a so far) & 1
    uint32_t save_cnt = 0;                         // kSave: hidden sub-steps staged so far (buffer = count & 1)
#ifdef NSFF_TIMELINE
    const uint32_t tl_base = sDbg + 2 * kDbgPerRole * 8;
    const int tl_role = 2;
    int tl_n = 0;
    const bool tl_on = P.dbg && blockIdx.x == 0 && warp == 0;
#endif
    for (int unit = group; unit < n_units; unit += n_groups) {
      NSFF_TL((gsub << 8) | 13);                             // epi: tile start
      const int tile = unit * 2 + (int)cta_rank;
      const int64_t p_raw = (int64_t)tile * kTileM + row;
      const bool valid = p_raw < P.pts.B;
      const int64_t p = valid ? p_raw : P.pts.B - 1;

      float zrow = 0.f;
      if (P.pts.zs && P.x_emb == nullptr) zrow = P.pts.zs[p];

      // kBwd: the ReLU bits of this warp's 2 x 16 columns of a masked sub-step are fetched one sub-step ahead - their
      // global-memory latency (~700 cycles) would otherwise sit on the epilogue's critical path 38 times per tile.  The
      // plane offsets come precomputed with the sub-step table; the two raw 16-bit loads are kept apart (combining them
      // here would wait for the loads right after issuing them).
      uint32_t mask_lo = 0xFFFFu, mask_hi = 0xFFFFu;
      bool mask_pref_ok = false;
      const uint8_t* mask_tb = nullptr;
      if constexpr (kBwd)
        mask_tb = P.mask_base + (size_t)tile * kSaveBitsPerTile + (size_t)row * 4 + (part & 1) * 2 + (part >> 1) * 512;
      uint8_t* bits_tb = nullptr;                      // training forward: where this thread's ReLU bits of a tile go
      if constexpr (kSave && !kBwd)
        bits_tb = P.save_base + (size_t)tile * kSaveBitsPerTile + (size_t)row * 4 + (part & 1) * 2 + (part >> 1) * 512;
      // ---------------- epilogues, one per sub-step ----------------
      for (int s = 0; s < n_subs; ++s, ++gsub) {
        const SubStep& ss = P.subs[s];
        // everything the hot path needs from the sub-step table comes with one 8-byte constant load
        const uint2 hot = *reinterpret_cast<const uint2*>(&ss);
        const uint32_t d_col = hot.x & 0xFFFFu, acc = (hot.x >> 16) & 0xFFu, epi = hot.x >> 24;
        const uint32_t s_bias = sBias + 4u * (hot.y & 0xFFFFu);
        const uint32_t half = ((hot.y >> 16) & 0xFFu) >> 1;
        const bool arrive_empty = (hot.y >> 24) != 0;
        const uint32_t t_acc = tmem_base + lane_field + d_col;
        const uint32_t acc_full = bar(BAR_ACC_FULL + acc), acc_par = (drain_par >> acc) & 1u;
        drain_par ^= 1u << acc;
        if (epi != EPI_HEAD && epi != EPI_ROWS) {
          // One N = 128 half of a 256-wide layer: two 64-column slices; this warp owns 16 columns of each.
          const bool relu = epi == EPI_RELU_HID;
          // bias of the first slice is fetched before the accumulator is awaited: it is off the critical path
          float4 b[4];
          uint32_t mb0 = 0xFFFFu, mb1 = 0xFFFFu;               // kBwd: ReLU bits of this warp's 16 columns of either slice
          if constexpr (kBwd) {
            if (tile < P.n_tiles) {
              if (epi == EPI_MASK_HID) {
                if (!mask_pref_ok) {                 // (first masked sub-step of a tile only)
                  mask_lo = __ldg(reinterpret_cast<const uint16_t*>(mask_tb + ss.mask_off));
                  mask_hi = __ldg(reinterpret_cast<const uint16_t*>(mask_tb + ss.mask_off + 1024));
                }
                mb0 = mask_lo, mb1 = mask_hi;
              }
              const uint64_t nmo = ss.next_mask_off;
              mask_pref_ok = nmo != kNoSave;
              if (mask_pref_ok) {
                mask_lo = __ldg(reinterpret_cast<const uint16_t*>(mask_tb + nmo));
                mask_hi = __ldg(reinterpret_cast<const uint16_t*>(mask_tb + nmo + 1024));    // chunk slice0 + 1: two word planes on
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) lds_v4(s_bias + (uint32_t)((part * 16 + i * 4) * 4), b[i]);
          }
          NSFF_TL((gsub << 8) | 5);                            // epi: waiting for the accumulator
          mbar_wait(acc_full, acc_par);          // (a non-suspending test_wait spin here was measured 6 % slower: 16 spinning warps take issue slots)
          NSFF_TL((gsub << 8) | 6);                            // epi: accumulator complete
          tc_fence_after_sync();
          uint32_t v0[16], v1[16];
          tmem_ld16(t_acc + (uint32_t)(part * 16), v0);        // both slices in flight together
          tmem_ld16(t_acc + (uint32_t)(64 + part * 16), v1);
          tmem_ld_wait();
          // accumulator drained into registers: its columns may be overwritten by the layer after next (which cannot
          // start before the bf16 image written below has been published and consumed)
          if (arrive_empty) {                                  // (hidden accumulators: normally nobody waits, see build_program)
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) arrive_leader(BAR_ACC_EMPTY + acc);
          }
          uint32_t w0[8], w1[8];
          float4 b1[4];                                        // bias of the second slice: in flight during the first
          if constexpr (kBwd) {
            mask_pack(v0, mb0, w0);
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) lds_v4(s_bias + (uint32_t)((64 + part * 16 + i * 4) * 4), b1[i]);
            bias_act_pack<kSave>(v0, b, relu, w0, &mb0);        // (kSave: mb0 / mb1 receive the ReLU bits)
          }
          // K-step `part` of hidden chunk slice0: 16 bf16 = 8 columns, in place (columns this warp has drained)
          tmem_st8(t_acc + (uint32_t)(part * 16), w0[0], w0[1], w0[2], w0[3], w0[4], w0[5], w0[6], w0[7]);
          if constexpr (kBwd) mask_pack(v1, mb1, w1);
          else bias_act_pack<kSave>(v1, b1, relu, w1, &mb1);
          tmem_st8(t_acc + (uint32_t)(8 + part * 16), w1[0], w1[1], w1[2], w1[3], w1[4], w1[5], w1[6], w1[7]);
          // both chunks of this 128-channel half are written: publish it to the MMA issuer
          tmem_st_wait();
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) arrive_leader(BAR_HID_FULL + half);
          NSFF_TL((gsub << 8) | 7);                            // epi: hidden half published
          if constexpr (kSave) {
            // training mode: the bf16 rows just written are also the backward's tile image (128B-swizzled chunk
            // [tile][chunk][128 rows x 128 B]).  This warp owns one aligned 32-byte sector of its row per chunk: it goes
            // into the staging buffer (same layout), from where the store issuer ships whole chunks; and 16 ReLU bits
            // per chunk (collected by bias_act_pack) go straight to global.
            if (ss.save_off != kNoSave) {
              const uint32_t buf = save_cnt & 1u;
              mbar_wait(bar(BAR_SAVE_EMPTY + buf), ((save_cnt >> 1) & 1u) ^ 1u);
              NSFF_TL((gsub << 8) | 8);                        // epi: staging buffer free
              ++save_cnt;
#pragma unroll
              for (int c = 0; c < 2; ++c) {
                const uint32_t* w = c ? w1 : w0;
                const bool swap = rsw & 1;
                const uint32_t a = sSave + (buf * 2u + (uint32_t)c) * kChunkBytes + row_off + (uint32_t)((((part * 2) ^ rsw) & 6) << 4);
                if (swap) {
                  st_shared_v4(a, w[4], w[5], w[6], w[7]);
                  st_shared_v4(a + 16, w[0], w[1], w[2], w[3]);
                } else {
                  st_shared_v4(a, w[0], w[1], w[2], w[3]);
                  st_shared_v4(a + 16, w[4], w[5], w[6], w[7]);
                }
              }
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar(BAR_SAVE_FULL + buf));
              NSFF_TL((gsub << 8) | 9);                        // epi: rows staged
              if constexpr (!kBwd) {
                if (relu && tile < P.n_tiles) {
                  uint8_t* bp = bits_tb + ss.mask_off;                               // word plane of chunk slice0
                  *reinterpret_cast<uint16_t*>(bp) = (uint16_t)mb0;
                  *reinterpret_cast<uint16_t*>(bp + 1024) = (uint16_t)mb1;         // chunk slice0 + 1: two word planes on
                }
              }
            }
          }
        } else if (kBwd && epi == EPI_ROWS) {
          // 64 fp32 columns per point to global rows (gradient w.r.t. the trunk input columns): 16 columns per warp
          const uint32_t col0 = hot.y & 0xFFFFu;
          const bool beta = ss.head_tbl[0] == 1;
          float* orow = P.out + p * P.out_ld + col0 + part * 16;
          float4 old[4];
          if (beta && valid) {
#pragma unroll
            for (int i = 0; i < 4; ++i) old[i] = *reinterpret_cast<const float4*>(orow + 4 * i);
          }
          mbar_wait(acc_full, acc_par);
          tc_fence_after_sync();
          uint32_t v[16];
          tmem_ld16(t_acc + (uint32_t)(part * 16), v);
          tmem_ld_wait();
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) arrive_leader(BAR_ACC_EMPTY + acc);
          if (valid) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float4 x = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                     __uint_as_float(v[4 * i + 3]));
              if (beta) x.x += old[i].x, x.y += old[i].y, x.z += old[i].z, x.w += old[i].w;
              *reinterpret_cast<float4*>(orow + 4 * i) = x;
            }
          }
        } else {
          // Head block (N = 16): each of the four column-part warps of a lane quadrant owns 4 head columns.
          const uint32_t tbl = reinterpret_cast<const uint32_t*>(ss.head_tbl)[part];   // 4 x (dst | act << 6)
          float4 hb;
          lds_v4(s_bias + (uint32_t)(part * 16), hb);
          mbar_wait(acc_full, acc_par);
          tc_fence_after_sync();
          uint32_t v[4];
          tmem_ld4(t_acc + (uint32_t)(part * 4), v);
          tmem_ld_wait();
          // accumulator and bias row are in registers: hand both back before the activations / global stores
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) arrive_leader(BAR_ACC_EMPTY + acc);
          if (valid) {
            float* orow = P.out + p * P.out_ld;
            const float hbv[4] = {hb.x, hb.y, hb.z, hb.w};
            const bool warp_pts = P.x_emb == nullptr && P.pts.offset == nullptr;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t e = (tbl >> (8 * j)) & 0xFFu;
              if (e == 0xFFu) continue;
              const uint32_t dst = e & 63u, act = e >> 6;
              float x = __uint_as_float(v[j]) + hbv[j];
              if (act == HA_SIGMOID) x = sigmoid_f(x);
              else if (act == HA_FLOW) {
                x = (zrow > kZFar) ? 0.f : P.flow_scale * tanh_f(x);
                // points of the warped passes, written here so those passes do not gather 12 B out of every output
                // row: head columns 4..6 = flow fw, 7..9 = flow bw
                const int hc = part * 4 + j;
                if (warp_pts && hc >= 4 && hc < 10) {
                  float* wp = hc < 7 ? P.pts.xyz_fw_out : P.pts.xyz_bw_out;
                  const int cc = hc < 7 ? hc - 4 : hc - 7;
                  if (wp) wp[p * 3 + cc] = __fadd_rn(P.pts.xyz[p * 3 + cc], x);
                }
              }
              orow[dst] = x;
            }
          }
        }
      }
    }
    NSFF_TL_DUMP();
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
  *bytes = ((L.bytes + 1023) & ~(size_t)1023) + ((sizeof(PackJobs) + 1023) & ~(size_t)1023) + kMergeScratchFloats * sizeof(float);
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
  const size_t scratch_off = jobs_off + ((sizeof(PackJobs) + 1023) & ~(size_t)1023);
  if (pack_bytes < scratch_off + kMergeScratchFloats * sizeof(float)) return NSFF_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* p8 = static_cast<uint8_t*>(pack);
  float* scratch = reinterpret_cast<float*>(p8 + scratch_off);

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
  // ---- folded blocks: products in fp32 scratch first, then packed like any other weight ----
  MergeJobs mj;
  int nm = 0;
  auto merge = [&](const float* A, int lda, int rows, const float* a_bias, const float* B, const float* b_bias, float* out,
                   int ldo, float* out_bias) {
    if (!A || !a_bias || !B || !b_bias) { missing = true; return; }
    MergeJob& m = mj.job[nm++];
    m.A = A, m.lda = lda, m.rows = rows, m.a_bias = a_bias, m.B = B, m.b_bias = b_bias;
    m.out = out, m.ldo = ldo, m.out_bias = out_bias;
  };
  float* m_dir = scratch;                          // (256, 331): [folded 256 | dir 27 | a 48] like static_dir_encoding
  float* m_dir_b = m_dir + 256 * 331;
  float* m_th = m_dir_b + 256;                     // (16, 256): rows 0-2 rgb, 3 sigma, 4-6 fw, 7-9 bw
  float* m_th_b = m_th + 16 * 256;
  float* m_rgb = m_th_b + 16;                      // (16, 256): rows 0-2 static rgb (no view directions)
  float* m_rgb_b = m_rgb + 16 * 256;
  if (desc->use_viewdir) {
    const int ld = 283 + (desc->in_a > 0 ? 48 : 0);
    merge(prm->static_dir_w, ld, 256, prm->static_dir_b, prm->static_final_w, prm->static_final_b, m_dir, 331, m_dir_b);
    const int pre_m[2][2] = {{256, 27}, {283, 48}};
    int kc = 0;
    // the dir / appearance columns are taken from the original weight; the hidden columns from the product
    for (int i = 0; i < (desc->in_a > 0 ? 2 : 1); ++i, ++kc) wjob(L_SDIRM, kc, prm->static_dir_w, ld, 0, 256, pre_m[i][0], pre_m[i][1]);
    for (int h = 0; h < 4; ++h, ++kc) wjob(L_SDIRM, kc, m_dir, 331, 0, 256, h * 64, 64);
    bjob(L_SDIRM, 0, m_dir_b, 256);
  } else {
    merge(prm->static_rgb_w, 256, 3, prm->static_rgb_b, prm->static_final_w, prm->static_final_b, m_rgb, 256, m_rgb_b);
    for (int h = 0; h < 4; ++h) wjob(L_SRGBM, 3 - h, m_rgb, 256, 0, 3, h * 64, 64);
    bjob(L_SRGBM, 0, m_rgb_b, 3);
  }
  if (desc->has_transient) {
    merge(prm->trans_rgb_w, 256, 3, prm->trans_rgb_b, prm->trans_final_w, prm->trans_final_b, m_th, 256, m_th_b);
    merge(prm->trans_sigma_w, 256, 1, prm->trans_sigma_b, prm->trans_final_w, prm->trans_final_b, m_th + 3 * 256, 256, m_th_b + 3);
    if (desc->has_flow) {
      merge(prm->trans_fw_w, 256, 3, prm->trans_fw_b, prm->trans_final_w, prm->trans_final_b, m_th + 4 * 256, 256, m_th_b + 4);
      merge(prm->trans_bw_w, 256, 3, prm->trans_bw_b, prm->trans_final_w, prm->trans_final_b, m_th + 7 * 256, 256, m_th_b + 7);
    }
    const int nrow = desc->has_flow ? 10 : 4;
    for (int h = 0; h < 4; ++h) wjob(L_THEADM, 3 - h, m_th, 256, 0, nrow, h * 64, 64);
    bjob(L_THEADM, 0, m_th_b, nrow);
  }
  if (missing) return NSFF_EINVAL;
  if (nj > kMaxPackJobs || nm > kMaxMergeJobs) return NSFF_EUNSUPPORTED;

  merge_linear_kernel<<<dim3(256, nm), 256, 0, st>>>(mj);
  if (int lrc = after_launch()) return lrc;
  NSFF_CUDA_TRY(cudaMemsetAsync(p8, 0, jobs_off, st));   // zero padding rows/cols and unused bias slots
  // the job table travels as a kernel parameter (no host -> device copy: the whole re-pack can be captured in a CUDA graph)
  pack_kernel<<<nj, 256, 0, st>>>(jobs, p8, L.bias_base);
  return after_launch();
}

// Builds the per-tile program: the sub-step table the epilogue warps walk and the flat K-chunk op list the producer
// and the MMA issuer walk.  Every 256-wide layer becomes two N = 128 sub-steps (see the file header).  Shared by the
// forward (build_program) and the backward chain (build_chain_program).
struct ProgramBuilder {
  FwdParams& P;
  const bool split, fine;
  int n_subs = 0, n_ops = 0;
  int cur = -1;                        // region (0 = X, 1 = Y) that holds the current hidden activation; -1: none yet
  bool overflow = false;
  // Has any op waited for "half h of the current hidden layer published"?  Every publish must be waited for by the
  // issuer before the NEXT publish of the same barrier can complete, or a later parity wait would be satisfied by the
  // stale phase (the issuer may only ever be one phase behind).  The forward reads every hidden layer; in the backward
  // chain dZ of a branch's first layer has no reader, so the layer after it (extras only) takes the waits.
  bool hid_read[2] = {true, true};
  // host-only bookkeeping for the drain analysis in finish()
  struct OpMeta { int acc; bool first, head; int waits_half; bool needs_hi; int acc2; };   // acc2: second accumulator a merged op starts (-1: none)
  bool merge_pre = true;               // issue the K-chunks both halves read first as N = 256 MMAs (see push)
  OpMeta meta[kMaxOps];

  ProgramBuilder(FwdParams& p, bool split_, bool fine_) : P(p), split(split_), fine(fine_) { P.last_extra_sub = 0; }

  uint16_t a_of(int ch) const {                      // A-operand address of K-step 0 of chunk `ch`
    const int half_cols = split ? 128 : 64;          // TMEM columns between the bf16 images of channels 0-127 / 128-255
    if (ch < CH_PE) return (uint16_t)(cur * 256 + (ch >> 1) * half_cols + (ch & 1) * 8);     // TMEM column
    return (uint16_t)((ch - CH_PE) * (int)(kChunkBytes >> 4));                                 // smem, 16-byte units
  }

  // `narrow`: an N <= 64 block (head, fp32 rows) whose accumulator lives in the free upper columns of the region that
  // holds its input - slot 0: above the bf16 image of channels 0-127 (accumulator id kHeadAcc), slot 1: above the image
  // of channels 128-255 (id kHeadAcc + 1)
  SubStep* push(const PackLayer& pl, uint8_t epi, uint64_t save_off = kNoSave, uint64_t bits_off = kNoSave, int slot = 0) {
    const bool head = epi == EPI_HEAD || epi == EPI_ROWS;
    const int parts = (!head && split) ? 2 : 1;
    const int region = head ? cur : (cur < 0 ? 0 : 1 - cur);
    if (head && cur < 0) { overflow = true; return nullptr; }
    if (n_subs + parts > kMaxSubs) { overflow = true; return nullptr; }
    SubStep* first = &P.subs[n_subs];
    ChunkOp half_ops[2][kMaxChunks];
    bool seen_half[2] = {false, false};
    OpMeta half_meta[2][kMaxChunks];
    int n_half = 0;
    for (int h = 0; h < parts; ++h) {
      SubStep& s = P.subs[n_subs + h];
      const int n = head ? pl.n : pl.n / parts;
      s.bias_off = (uint16_t)(pl.bias_off + (uint32_t)(h * n));
      s.arrive_empty = head ? 1 : 0;
      s.n = (uint16_t)n;
      s.epi = epi;
      if (head) {
        // free columns of the region that holds the input: above the in-place bf16 image of channels 0-127 / 128-255
        s.d_col = (uint16_t)(region * 256 + (split ? 64 : 128) + (slot ? 128 : 0));
        s.acc_id = (uint8_t)(kHeadAcc + (slot ? 1 : 0));
        s.slice0 = 0;
      } else {
        s.d_col = (uint16_t)(region * 256 + h * 128);
        s.acc_id = (uint8_t)(region * 2 + h);
        s.slice0 = (uint8_t)(2 * h);
      }
      memset(s.head_tbl, 0xFF, sizeof(s.head_tbl));
      memset(s.pad, 0, sizeof(s.pad));
      s.save_off = save_off;
      s.bits_off = bits_off;
      s.mask_off = s.next_mask_off = kNoSave;
      const int block_rows = pl.block_rows ? pl.block_rows : pl.n;
      const uint32_t block_bytes = (uint32_t)block_rows * 128u;          // whole packed block of one K-chunk
      const uint32_t piece = (uint32_t)n * 128u / 2u;                    // this CTA's N/2 rows of it
      int k = 0;
      for (int kc = 0; kc < pl.nchunks; ++k) {
        ChunkOp& o = half_ops[h][k];
        OpMeta& m = half_meta[h][k];
        memset(&o, 0, sizeof(o));
        const int ch = pl.chunk[kc];
        const bool hidden = ch < CH_PE;
        if (hidden && cur < 0) { overflow = true; return nullptr; }
        // two consecutive K-chunks of the same kind share an op: (H0,H1) (H2,H3) / (H3,H2) (H1,H0) / (PE,T) / (DIR,A)
        const bool two = !(fine && !head) && kc + 1 < pl.nchunks && ((pl.chunk[kc + 1] < CH_PE) == hidden) &&
                         (hidden ? (pl.chunk[kc + 1] >> 1) == (ch >> 1) : pl.chunk[kc + 1] == ch + 1);
        auto src_of = [&](int c) {
          const uint32_t b = pl.blk_explicit ? pl.blk[c] : (uint32_t)c;
          return pl.w_off + b * block_bytes + (uint32_t)(pl.row0 + h * n) * 128u;
        };
        o.idesc = make_idesc_bf16(2 * kTileM, (uint32_t)n);
        o.a = a_of(ch);
        o.d_col = s.d_col;
        o.acc_id = s.acc_id;
        o.bytes = piece;
        o.w_src0 = src_of(kc);
        o.flags = hidden ? OP_TS : 0;
        if (two) {
          o.flags |= OP_TWO;
          o.a2 = a_of(pl.chunk[kc + 1]);
          o.b2 = (uint16_t)(piece >> 4);
          o.w_src1 = src_of(kc + 1);
        }
        m.acc = s.acc_id, m.acc2 = -1, m.first = kc == 0, m.head = head, m.waits_half = -1;
        m.needs_hi = hidden && (ch >> 1) == 1;
        if (hidden) {
          // only the first op of the layer that reads a hidden half waits for it (half 1's ops issue later)
          if (h == 0 && !seen_half[ch >> 1]) o.wait_hid = (uint8_t)(1 + (ch >> 1)), m.waits_half = ch >> 1, seen_half[ch >> 1] = true;
        } else {
          P.last_extra_sub = n_subs + h;
        }
        if (kc == 0) o.flags |= OP_FIRST;
        kc += two ? 2 : 1;
        if (kc == pl.nchunks) o.flags |= OP_LAST;
      }
      n_half = k;
    }
    if (!head && parts == 2 && cur >= 0) {
      for (int h = 0; h < 2; ++h)
        if (!hid_read[h] && !seen_half[h]) {
          // nobody read half h of the previous hidden layer: this layer's half-h ops wait for its publish all the same
          half_ops[h][0].wait_hid = (uint8_t)(1 + h);
          half_meta[h][0].waits_half = h;
          seen_half[h] = true;
        }
    }
    for (int h = 0; h < 2; ++h)
      if (seen_half[h]) hid_read[h] = true;
    // Issue order inside a layer.  The ops that read channels 0-127 of the previous layer run before anything needs
    // channels 128-255, which the epilogue of the previous layer's second half is still producing; and half 0 of THIS
    // layer finishes while three ops of half 1 are still to come, so that its epilogue (whose result the next layer's
    // first ops need) has cover too:      a0 b0 a1 | a2 a3 | b1 b2 b3        (a = half 0, b = half 1, one K-chunk each).
    // With two-chunk ops (fine = false) this degenerates to a01 b01 a23 b23 - and a01 b01 become the merged ab0 ab1 below.
    auto emit = [&](int h, int k) {
      if (n_ops >= kMaxOps) { overflow = true; return; }
      meta[n_ops] = half_meta[h][k];
      P.ops[n_ops++] = half_ops[h][k];
    };
    if (parts == 1) {
      for (int k = 0; k < n_half; ++k) emit(0, k);
    } else {
      int n_pre = 0;
      while (n_pre < n_half && !half_meta[0][n_pre].needs_hi) ++n_pre;
      const int hold = (fine && n_pre < n_half && n_pre > 0) ? 1 : 0;     // ops of half 1 kept back to cover half 0's epilogue
      if (merge_pre && !fine && n_pre > 0 && n_pre < n_half) {
        // The K-chunks that only need channels 0-127 of the previous layer (or the prologue chunks) feed BOTH halves:
        // one N = 256 MMA per K-step into the two adjacent accumulators instead of two N = 128 MMAs (138 vs 2 x 74
        // cycles, profiles/r02a_ubench_mma.txt).  One K-chunk per op: its weight piece (128 rows per CTA) fills a stage.
        for (int k = 0; k < n_pre; ++k) {
          const ChunkOp& a = half_ops[0][k];
          const int pieces = (a.flags & OP_TWO) ? 2 : 1;
          for (int c = 0; c < pieces; ++c) {
            if (n_ops >= kMaxOps) { overflow = true; break; }
            ChunkOp o = a;
            OpMeta m = half_meta[0][k];
            o.idesc = make_idesc_bf16(2 * kTileM, 256u);
            o.flags = (uint8_t)(a.flags & ~(OP_TWO | OP_LAST | (c ? OP_FIRST : 0)));
            o.a = c ? a.a2 : a.a;
            o.a2 = 0, o.b2 = 0;
            o.w_src0 = c ? a.w_src1 : a.w_src0;
            o.w_src1 = 0;
            o.bytes = 256u * 128u / 2u;
            if (c) o.wait_hid = 0, m.waits_half = -1;
            m.first = (o.flags & OP_FIRST) != 0;
            m.acc2 = m.first ? half_meta[1][k].acc : -1;
            meta[n_ops] = m;
            P.ops[n_ops++] = o;
          }
        }
      } else {
        for (int k = 0; k < n_pre; ++k) {
          emit(0, k);
          if (k < n_pre - hold) emit(1, k);
        }
      }
      if (fine) {
        for (int k = n_pre; k < n_half; ++k) emit(0, k);
        for (int k = n_pre - hold; k < n_half; ++k) emit(1, k);
      } else {
        for (int k = n_pre; k < n_half; ++k) emit(0, k), emit(1, k);
      }
    }
    if (overflow) return nullptr;
    if (!head) P.ops[n_ops - 1].flags |= OP_ADV_HID;
    n_subs += parts;
    if (!head) cur = region, hid_read[0] = hid_read[1] = false;
    return first;
  }

  int finish() {
    if (overflow) return NSFF_EUNSUPPORTED;
    // the last hidden layer of the tile, if nobody read it: the first layer of the NEXT tile waits for its publish
    // (first tile: parity of a fresh barrier, passes at once)
    for (int h = 0; h < 2; ++h)
      if (!hid_read[h]) {
        int at = -1;
        for (int i = 0; i < n_ops && at < 0; ++i)
          if (meta[i].first && !meta[i].head && ((meta[i].acc & 1) == h || (meta[i].acc2 >= 0 && (meta[i].acc2 & 1) == h))) at = i;
        if (at < 0 || P.ops[at].wait_hid != 0) return NSFF_EUNSUPPORTED;
        P.ops[at].wait_hid = (uint8_t)(1 + h);
        meta[at].waits_half = h;
      }
    // Which accumulators must the issuer explicitly wait to be drained before overwriting them?  The drain of a hidden
    // accumulator (region r, half h) is implied once any op has waited for "hidden half h of the layer stored there
    // published" (the epilogue reads the accumulator before it publishes), which the next layer always does; what is
    // left are the narrow accumulators (inside a half of a region) and whatever the walk below finds uncovered.
    // The table is the same for every tile, so the walk covers two consecutive tiles and keeps the union.
    {
      int stored[kAccIds], waited[2] = {-1, -1}, layer = 0, cur_layer = -1;
      for (int i = 0; i < kAccIds; ++i) stored[i] = -1;
      bool narrow_pending[2] = {false, false};
      int narrow_region[2] = {-1, -1};
      for (int pass = 0; pass < 2; ++pass)
        for (int i = 0; i < n_ops; ++i) {
          ChunkOp& o = P.ops[i];
          const OpMeta& m = meta[i];
          if (m.waits_half >= 0) waited[m.waits_half] = cur_layer;
          if (m.first) {
            uint8_t mask = 0;
            if (m.head) {
              const int k = m.acc - kHeadAcc;
              if (narrow_pending[k]) mask |= 1u << m.acc;            // previous narrow accumulator (same barrier pair)
              narrow_pending[k] = true;
              narrow_region[k] = (o.d_col >> 8) & 1;
            } else {
              for (int a : {m.acc, m.acc2}) {
                if (a < 0) continue;
                const int h = a & 1, r = a >> 1;
                if (stored[a] >= 0 && waited[h] < stored[a]) mask |= 1u << a;
                if (narrow_pending[h] && r == narrow_region[h]) mask |= 1u << (kHeadAcc + h), narrow_pending[h] = false;
                stored[a] = layer;
              }
            }
            o.wait_empty |= mask;
          }
          if (o.flags & OP_ADV_HID) cur_layer = layer++;
        }
    }
    {
      uint32_t waited_accs = 0;
      for (int i = 0; i < n_ops; ++i) waited_accs |= P.ops[i].wait_empty;
      for (int i = 0; i < n_subs; ++i) P.subs[i].arrive_empty = (waited_accs >> P.subs[i].acc_id) & 1u;
    }
    {
      int last_extra_op = -1;
      for (int i = 0; i < n_ops; ++i)
        if (!(P.ops[i].flags & OP_TS)) last_extra_op = i;
      if (last_extra_op < 0) return NSFF_EUNSUPPORTED;     // every program starts from chunks the prologue warps write
      P.ops[last_extra_op].flags |= OP_EXTRA_DONE;
    }
    P.n_subs = n_subs;
    P.n_ops = n_ops;
    return NSFF_OK;
  }
};

static int build_program(const nsff_field_desc& d, const PackLayout& L, int out_mask, bool split, bool fine, bool fuse,
                         FwdParams& P,
                         int* n_out_cols, const SaveLayout* SL) {
  const bool sigma_only = out_mask & NSFF_OUT_SIGMA_ONLY;
  const bool want_s = out_mask & NSFF_OUT_STATIC, want_t = out_mask & NSFF_OUT_TRANSIENT;
  const bool want_fw = out_mask & NSFF_OUT_FLOW_FW, want_bw = out_mask & NSFF_OUT_FLOW_BW;
  if (!want_s && !want_t) return NSFF_EINVAL;
  if (want_t && !d.has_transient) return NSFF_EINVAL;
  if ((want_fw || want_bw) && (!want_t || !d.has_flow || sigma_only)) return NSFF_EINVAL;
  int col = 0;
  ProgramBuilder pb(P, split, fine);
  if (const char* me = getenv("NSFF_TC_MERGE")) pb.merge_pre = atoi(me) != 0;      // A/B knob
  auto push = [&](int lid, uint8_t epi) -> SubStep* {
    uint64_t save_off = kNoSave, bits_off = kNoSave;
    if (SL) {
      if (lid >= L_S1 && lid <= L_S8) save_off = SL->s_h[lid - L_S1], bits_off = SL->s_hb[lid - L_S1];
      else if (lid == L_SFIN) save_off = SL->s_feat;
      else if (lid == L_SDIR) save_off = SL->s_dir, bits_off = SL->s_dirb;
      else if (lid >= L_T1 && lid <= L_T8) save_off = SL->t_h[lid - L_T1], bits_off = SL->t_hb[lid - L_T1];
      else if (lid == L_TFIN) save_off = SL->t_feat;
    }
    SubStep* sub = pb.push(L.layer[lid], epi, save_off, bits_off);
    if (sub && bits_off != kNoSave && epi == EPI_RELU_HID)
      for (int h = 0; h < 2; ++h) sub[h].mask_off = bits_off + (uint64_t)sub[h].slice0 * 1024u;     // bit-plane offset of chunk slice0
    return sub;
  };
  P.need_dir = 0;
  P.need_t = 0;
  P.need_a = 0;
  auto head_col = [](SubStep* s, int i, int dst, int act) {
    if (s) s->head_tbl[i] = (uint8_t)(dst | (act << 6));
  };
  if (want_s) {
    for (int i = L_S1; i <= L_S8; ++i) push(i, EPI_RELU_HID);
    if (sigma_only) {
      head_col(push(L_SSIG, EPI_HEAD), 0, col, HA_NONE);
      col += 1;
    } else {
      head_col(push(L_SSIG, EPI_HEAD), 0, col + 3, HA_NONE);
      if (!fuse) push(L_SFIN, EPI_LIN_HID);
      if (d.use_viewdir) {
        push(fuse ? L_SDIRM : L_SDIR, EPI_RELU_HID);
        P.need_dir = 1;
        P.need_a = d.in_a > 0;
      }
      SubStep* rg = push((fuse && !d.use_viewdir) ? L_SRGBM : L_SRGB, EPI_HEAD);
      for (int i = 0; i < 3; ++i) head_col(rg, i, col + i, HA_SIGMOID);
      col += 4;
    }
  }
  if (want_t) {
    P.need_t = 1;
    for (int i = L_T1; i <= L_T8; ++i) push(i, EPI_RELU_HID);
    if (!fuse) push(L_TFIN, EPI_LIN_HID);
    SubStep* h = push(fuse ? L_THEADM : L_THEAD, EPI_HEAD);
    if (sigma_only) {
      head_col(h, 3, col, HA_NONE);
      col += 1;
    } else {
      for (int i = 0; i < 3; ++i) head_col(h, i, col + i, HA_SIGMOID);
      head_col(h, 3, col + 3, HA_NONE);
      col += 4;
      if (want_fw) {
        for (int i = 0; i < 3; ++i) head_col(h, 4 + i, col + i, HA_FLOW);
        col += 3;
      }
      if (want_bw) {
        for (int i = 0; i < 3; ++i) head_col(h, 7 + i, col + i, HA_FLOW);
        col += 3;
      }
    }
  }
  *n_out_cols = col;
  return pb.finish();
}

#ifdef NSFF_TIMELINE
// NSFF_TIMELINE builds: every launch with NSFF_TIMELINE_OUT=<prefix> set dumps CTA 0's event log to <prefix>.<n>.bin
static cudaError_t timeline_begin(FwdParams& P, cudaStream_t st, unsigned long long** dev) {
  *dev = nullptr;
  if (!getenv("NSFF_TIMELINE_OUT")) return cudaSuccess;
  const size_t bytes = (size_t)3 * kDbgPerRole * sizeof(unsigned long long);
  cudaError_t e = cudaMalloc(dev, bytes);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(*dev, 0, bytes, st);
  P.dbg = *dev;
  return e;
}
static cudaError_t timeline_end(cudaStream_t st, unsigned long long* dev) {
  if (!dev) return cudaSuccess;
  static int tl_seq = 0;
  const size_t bytes = (size_t)3 * kDbgPerRole * sizeof(unsigned long long);
  unsigned long long* host = (unsigned long long*)malloc(bytes);
  cudaError_t e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = cudaMemcpy(host, dev, bytes, cudaMemcpyDeviceToHost);
  char name[512];
  snprintf(name, sizeof(name), "%s.%d.bin", getenv("NSFF_TIMELINE_OUT"), tl_seq++);
  if (FILE* f = fopen(name, "wb")) {
    fwrite(host, 1, bytes, f);
    fclose(f);
  }
  free(host);
  cudaFree(dev);
  return e;
}
#endif

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
  const bool split = true;       // every 256-wide layer = two N = 128 sub-steps (file header)
  // default: issue order ab0 ab1 a23 b23 (merged N = 256 ops for the chunks both halves read first, then two K-chunks =
  // 8 MMAs per op; NSFF_TC_MERGE=0: a01 b01 a23 b23).  NSFF_TC_FINE=1: one K-chunk per op, order
  // a0 b0 a1 | a2 a3 | b1 b2 b3 (more cover for the first half's epilogue, but twice the per-op issue overhead: measured slower)
  const char* fe = getenv("NSFF_TC_FINE");
  const bool fine = fe ? atoi(fe) != 0 : false;
  // Inference folds the activation-free "final" layers into their successors (build_layout); the training forward keeps
  // them (the backward needs their activations).  NSFF_TC_FUSE=0 disables the folding (A/B, parity experiments).
  const char* fu = getenv("NSFF_TC_FUSE");
  const bool fuse = !save && !(out_mask & NSFF_OUT_EVERY_LAYER) && (fu ? atoi(fu) != 0 : true);
  rc = build_program(*desc, L, out_mask, split, fine, fuse, P, &ncols, save ? &SL : nullptr);
  if (rc != NSFF_OK) return rc;
  if (out_ld < ncols) return NSFF_EINVAL;
  int dev = 0, major = 0, sms = 0;
  NSFF_CUDA_TRY(cudaGetDevice(&dev));
  NSFF_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) return NSFF_EARCH;
  NSFF_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  P.pack = static_cast<const uint8_t*>(pack);
  P.bias = reinterpret_cast<const float*>(P.pack + L.bias_base);
  P.n_bias = (int32_t)L.n_bias;
  if ((uint32_t)P.n_bias * 4u > kSmemBias) return NSFF_EUNSUPPORTED;
  P.out = out;
  P.out_ld = out_ld;
  P.n_tiles = (int)cdiv64(B, kTileM);
  P.flow_scale = desc->flow_scale;
  P.stage_bytes = fine ? 8192u : 16384u;          // one or two 8 KB weight pieces per op
  int max_ctas = sms;
  if (const char* cap = getenv("NSFF_MAX_CTAS")) {          // experiment knob: fewer resident CTAs
    int c = atoi(cap);
    if (c > 0 && c < max_ctas) max_ctas = c;
  }
  cudaStream_t st = (cudaStream_t)stream;
#ifdef NSFF_TIMELINE
  unsigned long long* tl_dev = nullptr;
  NSFF_CUDA_TRY(timeline_begin(P, st, &tl_dev));
#endif
  const bool appear = P.need_a != 0;
  P.n_stages = (int)((save ? (appear ? Cfg<true, true>::kSmemB : Cfg<false, true>::kSmemB)
                           : (appear ? Cfg<true>::kSmemB : Cfg<false>::kSmemB)) / P.stage_bytes);
  auto launch_pair = [&](auto kern, uint32_t smem) -> int {
    NSFF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int n_pairs = (P.n_tiles + 1) / 2;
    int clusters = max_ctas / 2;
    if (clusters < 1) clusters = 1;
    if (n_pairs < clusters) clusters = n_pairs;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(save ? kThreadsSave : kThreads);
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
    lrc = appear ? launch_pair(field_forward_kernel<true, true>, Cfg<true, true>::kSmemTotal)
                 : launch_pair(field_forward_kernel<false, true>, Cfg<false, true>::kSmemTotal);
  } else {
    lrc = appear ? launch_pair(field_forward_kernel<true, false>, Cfg<true>::kSmemTotal)
                 : launch_pair(field_forward_kernel<false, false>, Cfg<false>::kSmemTotal);
  }
  if (lrc != NSFF_OK) return lrc;
#ifdef NSFF_TIMELINE
  NSFF_CUDA_TRY(timeline_end(st, tl_dev));
#endif
  return after_launch();
}

// ---- backward chain launcher (declared in nsff_field_host.cuh; called by nsff_field_backward_tc) ------------------------
namespace nsff {
namespace tc {
int launch_backward_chain(const ChainLayer* layers, int n_layers, const void* pack, const uint8_t* dp_static,
                          const uint8_t* dp_trans, int64_t B, const uint8_t* mask, uint8_t* save, float* out_rows,
                          int32_t out_ld, cudaStream_t st) {
  if (!layers || n_layers < 1 || !pack || B <= 0 || !save) return NSFF_EINVAL;
  static thread_local FwdParams P;          // 8 KB: keep it off the stack of the calling thread
  memset(&P, 0, sizeof(P));
  ProgramBuilder pb(P, /*split=*/true, /*fine=*/false);
  if (const char* me = getenv("NSFF_TC_MERGE")) pb.merge_pre = atoi(me) != 0;
  for (int i = 0; i < n_layers; ++i) {
    const ChainLayer& c = layers[i];
    if (c.nchunks < 1 || c.nchunks > kMaxChunks) return NSFF_EINVAL;
    PackLayer pl;
    memset(&pl, 0, sizeof(pl));
    pl.present = true;
    pl.w_off = c.w_off;
    pl.n = c.n;
    pl.nchunks = c.nchunks;
    pl.blk_explicit = true;
    pl.block_rows = c.block_rows;
    pl.row0 = c.row0;
    for (int k = 0; k < c.nchunks; ++k) {
      pl.chunk[k] = c.chunk[k] < 4 ? c.chunk[k] : (uint8_t)(CH_PE + (c.chunk[k] - CC_DP_STATIC));   // DP chunks: extra slots 0 / 1
      pl.blk[k] = c.blk[k];
      if (c.chunk[k] == CC_DP_STATIC && !dp_static) return NSFF_EINVAL;
      if (c.chunk[k] == CC_DP_TRANS && !dp_trans) return NSFF_EINVAL;
    }
    if (c.epi == CHAIN_ROWS) {
      if (c.n != 64 || !out_rows || (out_ld & 3) || ((uintptr_t)out_rows & 15u) || (c.out_col0 & 3)) return NSFF_EINVAL;
      pl.bias_off = (uint32_t)c.out_col0;            // the sub-step's bias_off field carries the destination column
      SubStep* s = pb.push(pl, EPI_ROWS, kNoSave, kNoSave, c.out_col0 >= 64 ? 1 : 0);
      if (s) s->head_tbl[0] = c.beta ? 1 : 0;
    } else {
      if (c.n != 256) return NSFF_EINVAL;
      if (c.epi == CHAIN_MASK && (!mask || c.bits_off == kNoSave)) return NSFF_EINVAL;
      SubStep* s = pb.push(pl, c.epi == CHAIN_MASK ? EPI_MASK_HID : EPI_LIN_HID, c.save_off, c.bits_off);
      if (s) s[0].bias_off = s[1].bias_off = 0;
    }
  }
  int rc = pb.finish();
  if (rc != NSFF_OK) return rc;
  {
    // ReLU bit planes: [tile][8 word planes][128 rows] uint32; chunk c = planes 2c, 2c+1 (nsff_field_save.cuh)
    uint64_t next = kNoSave;
    for (int i = P.n_subs - 1; i >= 0; --i) {
      SubStep& s = P.subs[i];
      if (s.epi == EPI_ROWS) continue;
      s.next_mask_off = next;
      if (s.epi == EPI_MASK_HID) s.mask_off = s.bits_off + (uint64_t)s.slice0 * 1024u, next = s.mask_off;
      else next = kNoSave;
    }
  }
  int dev = 0, major = 0, sms = 0;
  NSFF_CUDA_TRY(cudaGetDevice(&dev));
  NSFF_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) return NSFF_EARCH;
  NSFF_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  P.pack = static_cast<const uint8_t*>(pack);
  P.pts.B = B;
  P.pts.S = 1;
  P.out = out_rows;
  P.out_ld = out_ld;
  P.n_tiles = (int)cdiv64(B, kTileM);
  P.stage_bytes = 16384u;
  P.n_stages = (int)(Cfg<false, true>::kSmemB / P.stage_bytes);
  P.save_base = save;
  P.mask_base = mask;
  P.dp_img[0] = dp_static;
  P.dp_img[1] = dp_trans;
  auto kern = field_forward_kernel<false, true, true>;
  const uint32_t smem = Cfg<false, true>::kSmemTotal;
  NSFF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int n_pairs = (P.n_tiles + 1) / 2;
  int clusters = sms / 2;
  if (clusters < 1) clusters = 1;
  if (n_pairs < clusters) clusters = n_pairs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(kThreadsSave);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
#ifdef NSFF_TIMELINE
  unsigned long long* tl_dev = nullptr;
  NSFF_CUDA_TRY(timeline_begin(P, st, &tl_dev));
#endif
  NSFF_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, P));
#ifdef NSFF_TIMELINE
  NSFF_CUDA_TRY(timeline_end(st, tl_dev));
#endif
  return after_launch();
}
}  // namespace tc
}  // namespace nsff

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
