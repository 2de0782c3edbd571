// a9-a12: density -> alpha, joint static+dynamic alpha composite, the two flow-warped composites and
// the disocclusion weights (models/rendering.py:202-298 and 98-140) in ONE pass over the per-sample
// MLP outputs.  HBM-bound: every input element is read exactly once, every output written once.
//
// Mapping: one warp per ray.  Lane l owns samples l, l+32, l+64, ... so that each warp-wide load of a
// scalar plane is a contiguous 128 B line (multi-channel rows are 16-56 B apart; the L1 merges the
// sector re-use across the unrolled channel loads).  The exclusive transmittance product is a
// 5-step shuffle scan per 32-sample chunk with a running carry; four scans run side by side
// (joint, static-only, warped-forward, warped-backward).
// Two kernels share the per-sample body: the staged one (bulk copies into per-warp shared-memory stages, below) and
// the direct one (per-lane global loads) for inputs that miss the bulk copy's 16-byte alignment rules.
#include <stdlib.h>

#include "nsff_common.cuh"
#include "nsff_sm100.cuh"

namespace nsff {
using namespace sm100;

__device__ __forceinline__ float warp_excl_prod(float v, float& carry, int lane) {
  // returns carry * prod_{lanes < lane} v ; updates carry *= prod_{all lanes} v
  float inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    float o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc *= o;
  }
  float excl = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane == 0) excl = 1.f;
  float out = carry * excl;
  carry *= __shfl_sync(0xffffffffu, inc, 31);
  return out;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

struct Vec3 {
  float x, y, z;
};
__device__ __forceinline__ Vec3 ld3(const float* p) { return Vec3{p[0], p[1], p[2]}; }
__device__ __forceinline__ void acc3(Vec3& a, float w, Vec3 v) {
  a.x += w * v.x;
  a.y += w * v.y;
  a.z += w * v.z;
}
__device__ __forceinline__ void st3_reduced(float* dst, int64_t ray, Vec3 v, int lane) {
  v.x = warp_sum(v.x);
  v.y = warp_sum(v.y);
  v.z = warp_sum(v.z);
  if (lane == 0) {
    dst[ray * 3 + 0] = v.x;
    dst[ray * 3 + 1] = v.y;
    dst[ray * 3 + 2] = v.z;
  }
}

// Everything one lane carries across the 32-sample chunks of its ray.
struct RayAcc {
  float carry = 1.f, carry_s = 1.f, carry_fw = 1.f, carry_bw = 1.f;
  float depth = 0.f, t_alpha = 0.f, so_depth = 0.f, occ_fw = 0.f, occ_bw = 0.f;
  Vec3 rgb_s{0, 0, 0}, rgb_t{0, 0, 0}, so_rgb{0, 0, 0}, xyz{0, 0, 0}, ffw{0, 0, 0}, fbw{0, 0, 0}, rgbfw{0, 0, 0},
      rgbbw{0, 0, 0};
};

// The row pointers of one sample: global memory (direct kernel) or the warp's shared-memory stage (staged kernel).
struct SampleSrc {
  const float* raw;       // [rgb_s s_s rgb_t s_t fw bw] (or the narrower layouts)
  const float* xyz;       // kFlow
  const float* rw[2];     // kWarp: [rgb sigma flow-back] of the fw / bw pass
  const float* xw[2];     // kWarp: warped points
};

// One sample per lane: alphas, the four transmittance scans, per-sample outputs (index i), per-ray partial sums.
template <bool kTransient, bool kFlow, bool kWarp>
__device__ __forceinline__ void composite_sample(const nsff_composite_args& a, RayAcc& A, const SampleSrc& src, int lane,
                                                 int s, int S, int64_t i, bool ok, float z, float z_next) {
  const int col_sig_s = a.has_static_rgb ? 3 : 0;
  const int col_rgb_t = 4;
  const int col_sig_t = a.has_static_rgb ? 7 : 1;
  const bool last = s >= S - 1;
  const float dz = last ? 0.f : z_next - z;
  const float d_s = last ? kLastDeltaStatic : dz;
  const float d_t = last ? kLastDeltaTrans : dz;
  const float* raw = src.raw;

  float raw_s = raw[col_sig_s];
  if (a.noise_static) raw_s += a.noise_static[i] * a.noise_std;
  const float sig_s = softplus_f(raw_s);
  const float a_s = 1.f - expf(-d_s * sig_s);
  Vec3 c_s{0, 0, 0};
  if (a.has_static_rgb) c_s = ld3(raw);

  float a_t = 0.f, sig_t = 0.f, alpha = a_s;
  Vec3 c_t{0, 0, 0};
  if (kTransient) {
    float raw_t = raw[col_sig_t];
    if (a.vis && a.vis[i] == 0.f) raw_t = -10.f;
    if (a.noise_trans) raw_t += a.noise_trans[i] * a.noise_std;
    sig_t = softplus_f(raw_t);
    a_t = 1.f - expf(-d_t * sig_t);
    alpha = 1.f - (1.f - a_s) * (1.f - a_t);
    if (a.has_static_rgb) c_t = ld3(raw + col_rgb_t);
  }
  // out-of-range lanes must not disturb the scans
  const float om = ok ? 1.f - alpha : 1.f;
  const float T = warp_excl_prod(om, A.carry, lane);
  const float w = alpha * T, w_s = a_s * T, w_t = a_t * T;

  if (ok) {
    if (a.static_sigmas) a.static_sigmas[i] = sig_s;
    if (kTransient) {
      if (a.transient_sigmas) a.transient_sigmas[i] = sig_t;
      if (a.static_weights) a.static_weights[i] = w_s;
      if (a.transient_weights) a.transient_weights[i] = w_t;
      if (a.weights) a.weights[i] = w;
      if (a.static_alphas) a.static_alphas[i] = a_s;
      if (a.transient_alphas) a.transient_alphas[i] = a_t;
    } else {
      if (a.static_weights) a.static_weights[i] = w;     // rendering.py:248
    }
    A.depth += w * z;
    if (kTransient) {
      acc3(A.rgb_s, w_s, c_s);
      acc3(A.rgb_t, w_t, c_t);
      A.t_alpha += w_t;
    } else {
      acc3(A.rgb_s, w, c_s);
    }
  }

  if (kTransient) {
    // static field alone, with its own transmittance (rendering.py:270-278)
    const float T_s = warp_excl_prod(ok ? 1.f - a_s : 1.f, A.carry_s, lane);
    if (ok) {
      const float w_os = a_s * T_s;
      acc3(A.so_rgb, w_os, c_s);
      A.so_depth += w_os * z;
    }
  }

  if (kFlow) {
    if (ok) {
      acc3(A.xyz, w, ld3(src.xyz));
      acc3(A.ffw, w, ld3(raw + 8));
      acc3(A.fbw, w, ld3(raw + 11));
    }
  }

  if (kWarp) {
    // render_transient_warping (rendering.py:98-140): warped dynamic field over the CURRENT static one
#pragma unroll
    for (int dir = 0; dir < 2; ++dir) {
      const float* rw = src.rw[dir];
      const float* nz = dir == 0 ? a.noise_fw : a.noise_bw;
      float raw_w = rw[3];
      if (nz) raw_w += nz[i] * a.noise_std;
      const float a_tw = 1.f - expf(-d_t * softplus_f(raw_w));
      const float a_w = 1.f - (1.f - a_s) * (1.f - a_tw);
      const float T_w = warp_excl_prod(ok ? 1.f - a_w : 1.f, dir == 0 ? A.carry_fw : A.carry_bw, lane);
      if (ok) {
        const float w_tw = a_tw * T_w;
        Vec3& acc = dir == 0 ? A.rgbfw : A.rgbbw;
        acc3(acc, a_s * T_w, c_s);
        acc3(acc, w_tw, ld3(rw));
        const float occ = w_tw - w_t;                     // rendering.py:290-295 (detached)
        float* dis = dir == 0 ? a.disoccs_fw : a.disoccs_bw;
        if (dis) dis[i] = 1.f - fabsf(occ);
        (dir == 0 ? A.occ_fw : A.occ_bw) += occ;
        float* cyc = dir == 0 ? a.xyzs_fw_bw : a.xyzs_bw_fw;     // rendering.py:231-232
        if (cyc) {
          const float* xw = src.xw[dir];
          cyc[i * 3 + 0] = xw[0] + rw[4];
          cyc[i * 3 + 1] = xw[1] + rw[5];
          cyc[i * 3 + 2] = xw[2] + rw[6];
        }
      }
    }
  }
}

// ---- per-ray reductions --------------------------------------------------------------------
template <bool kTransient, bool kFlow, bool kWarp>
__device__ __forceinline__ void composite_finish(const nsff_composite_args& a, RayAcc& A, int64_t ray, int lane) {
  const float depth = warp_sum(A.depth);
  if (a.depth && lane == 0) a.depth[ray] = depth;
  if (kTransient) {
    Vec3 s = A.rgb_s, t = A.rgb_t;
    s.x = warp_sum(s.x); s.y = warp_sum(s.y); s.z = warp_sum(s.z);
    t.x = warp_sum(t.x); t.y = warp_sum(t.y); t.z = warp_sum(t.z);
    const float t_alpha = warp_sum(A.t_alpha);
    if (lane == 0) {
      if (a.rgb) {
        a.rgb[ray * 3 + 0] = s.x + t.x;
        a.rgb[ray * 3 + 1] = s.y + t.y;
        a.rgb[ray * 3 + 2] = s.z + t.z;
      }
      if (a.transient_alpha) a.transient_alpha[ray] = t_alpha;
      if (a.transient_rgb) {
        const float bg = kGreyBg * (1.f - t_alpha);
        a.transient_rgb[ray * 3 + 0] = t.x + bg;
        a.transient_rgb[ray * 3 + 1] = t.y + bg;
        a.transient_rgb[ray * 3 + 2] = t.z + bg;
      }
    }
    if (a.static_only_rgb) st3_reduced(a.static_only_rgb, ray, A.so_rgb, lane);
    const float so_depth = warp_sum(A.so_depth);
    if (a.static_only_depth && lane == 0) a.static_only_depth[ray] = so_depth;
  } else {
    if (a.rgb) st3_reduced(a.rgb, ray, A.rgb_s, lane);
  }
  if (kFlow) {
    Vec3 x = A.xyz, f = A.ffw, b = A.fbw;
    x.x = warp_sum(x.x); x.y = warp_sum(x.y); x.z = warp_sum(x.z);
    f.x = warp_sum(f.x); f.y = warp_sum(f.y); f.z = warp_sum(f.z);
    b.x = warp_sum(b.x); b.y = warp_sum(b.y); b.z = warp_sum(b.z);
    if (lane == 0) {
      const float xv[3] = {x.x, x.y, x.z}, fv[3] = {f.x, f.y, f.z}, bv[3] = {b.x, b.y, b.z};
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (a.xyz_fine) a.xyz_fine[ray * 3 + c] = xv[c];
        if (a.flow_fw) a.flow_fw[ray * 3 + c] = fv[c];
        if (a.xyz_fw_ray) a.xyz_fw_ray[ray * 3 + c] = xv[c] + fv[c];
        if (a.flow_bw) a.flow_bw[ray * 3 + c] = bv[c];
        if (a.xyz_bw_ray) a.xyz_bw_ray[ray * 3 + c] = xv[c] + bv[c];
      }
    }
  }
  if (kWarp) {
    if (a.rgb_fw) st3_reduced(a.rgb_fw, ray, A.rgbfw, lane);
    if (a.rgb_bw) st3_reduced(a.rgb_bw, ray, A.rgbbw, lane);
    const float occ_fw = warp_sum(A.occ_fw), occ_bw = warp_sum(A.occ_bw);
    if (lane == 0) {
      if (a.disocc_fw) a.disocc_fw[ray] = 1.f - fabsf(occ_fw);
      if (a.disocc_bw) a.disocc_bw[ray] = 1.f - fabsf(occ_bw);
    }
  }
}

// ---- direct kernel: every lane loads its sample's rows from global memory ---------------------------------------
// Used when the staged kernel's alignment conditions do not hold (S % 4 != 0, unaligned views).  Latency-bound
// (ncu: long-scoreboard stalls, 35 % issue slots; the 14-float rows cost 56 L1 sector look-ups per warp load); its
// 78 registers allowed only 3 blocks per SM, capping at 64 registers (4 blocks, 80 bytes of spills) measured
// 1.01 -> 0.89 ms on the bench frame (tools/comp_occ.sh).
#ifndef NSFF_COMP_MINB
#define NSFF_COMP_MINB 4
#endif
template <bool kTransient, bool kFlow, bool kWarp>
__global__ void __launch_bounds__(256, NSFF_COMP_MINB) composite_kernel(const nsff_composite_args a) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ray >= a.R) return;
  const int S = a.S;
  const int64_t base = ray * S;
  RayAcc A;
  for (int s0 = 0; s0 < S; s0 += 32) {
    const int s = s0 + lane;
    const bool ok = s < S;
    const int64_t i = base + (ok ? s : S - 1);
    const float z = a.zs[i];
    const float z_next = s >= S - 1 ? z : a.zs[i + 1];
    SampleSrc src;
    src.raw = a.raw + i * a.raw_ld;
    src.xyz = kFlow ? a.xyz + i * 3 : nullptr;
    if (kWarp) {
      src.rw[0] = a.raw_fw + i * a.raw_w_ld;
      src.rw[1] = a.raw_bw + i * a.raw_w_ld;
      src.xw[0] = a.xyz_fw ? a.xyz_fw + i * 3 : nullptr;
      src.xw[1] = a.xyz_bw ? a.xyz_bw + i * 3 : nullptr;
    }
    composite_sample<kTransient, kFlow, kWarp>(a, A, src, lane, s, S, i, ok, z, z_next);
  }
  composite_finish<kTransient, kFlow, kWarp>(a, A, ray, lane);
}

// ---- staged kernel: the bulk-copy engine streams 32-sample chunks into per-warp shared-memory stages --------------
// Persistent; every warp owns kCompStages stages and walks the (ray, chunk) items of its rays.  Lane 0 issues up to
// seven cp.async.bulk copies per item (the chunk's rows of every input plane are contiguous: 32 x 56 B, 32 x 28 B, ...)
// against the stage's mbarrier; the lanes then read their sample's rows from shared memory (row strides of 14 / 7 / 3
// words: at most 2-way bank conflicts).  The loads in flight no longer depend on registers or on the L1's sector
// throughput - kCompStages - 1 chunks (4.9 KB each) per warp are always on their way.  Measured on the bench frame
// (tools/comp_occ.sh): direct 0.956 ms (4.0 TB/s); staged 2 stages x 16 warps 0.753, x 20 warps 0.676 ms (5.7 TB/s =
// 0.87 of the measured HBM peak), x 22 warps 0.806 (register cap 88: spills); 3 stages x 14 warps 0.878 - the kernel
// is bound by warp-level latency (scans, MUFU), so warps beat stages.
#ifndef NSFF_COMP_STAGES
#define NSFF_COMP_STAGES 2
#endif
#ifndef NSFF_COMP_WARPS
#define NSFF_COMP_WARPS 20
#endif
constexpr int kCompStages = NSFF_COMP_STAGES, kCompWarps = NSFF_COMP_WARPS;

struct StageLayout {
  uint32_t raw, xyz, zs, rw[2], xw[2], bytes;
};
__host__ __device__ inline StageLayout stage_layout(int raw_ld, int raw_w_ld, bool flow, bool warp, bool cyc) {
  StageLayout L;
  uint32_t o = 0;
  L.raw = o; o += 32u * raw_ld * 4;
  L.zs = o; o += 36u * 4;
  L.xyz = o; if (flow) o += 32u * 12;
  for (int d = 0; d < 2; ++d) { L.rw[d] = o; if (warp) o += 32u * raw_w_ld * 4; }
  for (int d = 0; d < 2; ++d) { L.xw[d] = o; if (warp && cyc) o += 32u * 12; }
  L.bytes = (o + 15u) & ~15u;
  return L;
}

template <bool kTransient, bool kFlow, bool kWarp>
__global__ void __launch_bounds__(32 * kCompWarps, 1) composite_staged_kernel(const nsff_composite_args a) {
  extern __shared__ __align__(128) uint8_t comp_smem[];
  __shared__ __align__(8) uint64_t bars[kCompWarps * kCompStages];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int S = a.S, nch = (S + 31) >> 5;
  const bool cyc = kWarp && (a.xyzs_fw_bw || a.xyzs_bw_fw);
  const StageLayout L = stage_layout(a.raw_ld, a.raw_w_ld, kFlow, kWarp, cyc);
  uint8_t* my = comp_smem + (size_t)warp * kCompStages * L.bytes;
  const uint32_t bar0 = smem_u32(&bars[warp * kCompStages]);
  if (lane == 0) {
    for (int k = 0; k < kCompStages; ++k) mbar_init(bar0 + 8 * k, 1);
    fence_mbar_init();
  }
  __syncwarp();
  const int64_t n_warps = (int64_t)gridDim.x * kCompWarps, w0 = (int64_t)blockIdx.x * kCompWarps + warp;
  if (w0 >= a.R) return;
  const int64_t n_rays = (a.R - w0 + n_warps - 1) / n_warps, n_items = n_rays * nch;

  auto issue = [&](int64_t item) {      // lane 0 only
    const int st = (int)(item % kCompStages);
    const int64_t ray = w0 + (item / nch) * n_warps;
    const int s0 = (int)(item % nch) * 32, n = S - s0 < 32 ? S - s0 : 32;
    const int64_t i = ray * S + s0;
    const uint32_t dst = smem_u32(my + (size_t)st * L.bytes), bar = bar0 + 8 * st;
    const uint32_t nz = (S - s0 < 36 ? S - s0 : 36) * 4u;
    uint32_t tx = n * a.raw_ld * 4u + nz;
    if (kFlow) tx += n * 12u;
    if (kWarp) tx += 2u * n * a.raw_w_ld * 4u + (cyc ? 2u * n * 12u : 0u);
    mbar_arrive_expect_tx(bar, tx);
    bulk_g2s(dst + L.raw, a.raw + i * a.raw_ld, n * a.raw_ld * 4u, bar);
    bulk_g2s(dst + L.zs, a.zs + i, nz, bar);
    if (kFlow) bulk_g2s(dst + L.xyz, a.xyz + i * 3, n * 12u, bar);
    if (kWarp) {
      bulk_g2s(dst + L.rw[0], a.raw_fw + i * a.raw_w_ld, n * a.raw_w_ld * 4u, bar);
      bulk_g2s(dst + L.rw[1], a.raw_bw + i * a.raw_w_ld, n * a.raw_w_ld * 4u, bar);
      if (cyc) {
        bulk_g2s(dst + L.xw[0], a.xyz_fw + i * 3, n * 12u, bar);
        bulk_g2s(dst + L.xw[1], a.xyz_bw + i * 3, n * 12u, bar);
      }
    }
  };
  if (lane == 0)
    for (int64_t k = 0; k < kCompStages - 1 && k < n_items; ++k) issue(k);

  RayAcc A;
  for (int64_t item = 0; item < n_items; ++item) {
    const int st = (int)(item % kCompStages);
    const int c = (int)(item % nch);
    const int64_t ray = w0 + (item / nch) * n_warps;
    if (lane == 0 && item + kCompStages - 1 < n_items) issue(item + kCompStages - 1);   // its stage was freed below
    mbar_wait(bar0 + 8 * st, (uint32_t)((item / kCompStages) & 1));
    const int s0 = c * 32, s = s0 + lane, n = S - s0 < 32 ? S - s0 : 32;
    const bool ok = lane < n;
    const int r = ok ? lane : n - 1;
    const int64_t i = ray * S + s0 + r;
    const uint8_t* sp = my + (size_t)st * L.bytes;
    const float* zs = reinterpret_cast<const float*>(sp + L.zs);
    const float z = zs[r];
    const float z_next = s0 + r >= S - 1 ? z : zs[r + 1];
    SampleSrc src;
    src.raw = reinterpret_cast<const float*>(sp + L.raw) + r * a.raw_ld;
    src.xyz = reinterpret_cast<const float*>(sp + L.xyz) + r * 3;
    if (kWarp) {
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        src.rw[d] = reinterpret_cast<const float*>(sp + L.rw[d]) + r * a.raw_w_ld;
        src.xw[d] = reinterpret_cast<const float*>(sp + L.xw[d]) + r * 3;
      }
    }
    composite_sample<kTransient, kFlow, kWarp>(a, A, src, lane, s, S, i, ok, z, z_next);
    __syncwarp();                      // every lane is done with this stage before lane 0 refills it (next iteration)
    if (c == nch - 1) {
      composite_finish<kTransient, kFlow, kWarp>(a, A, ray, lane);
      A = RayAcc();
    }
  }
}

}  // namespace nsff

using namespace nsff;

extern "C" int nsff_composite(const nsff_composite_args* args, void* stream) {
  if (!args) return NSFF_EINVAL;
  const nsff_composite_args& a = *args;
  if (a.R < 0 || a.S < 1) return NSFF_EINVAL;
  if (a.R == 0) return NSFF_OK;
  if (!a.zs || !a.raw || a.raw_ld < 1) return NSFF_EINVAL;
  const bool warp = a.raw_fw != nullptr || a.raw_bw != nullptr;
  if (warp && (!a.raw_fw || !a.raw_bw || !a.has_flow || !a.has_transient || !a.has_static_rgb)) return NSFF_EINVAL;
  if (warp && (a.xyzs_fw_bw || a.xyzs_bw_fw) && (!a.xyz_fw || !a.xyz_bw)) return NSFF_EINVAL;
  if (a.has_flow && (!a.has_transient || !a.has_static_rgb || a.raw_ld < 14)) return NSFF_EINVAL;
  if (a.has_flow && !a.xyz) return NSFF_EINVAL;
  if ((a.noise_static || a.noise_trans || a.noise_fw || a.noise_bw) && !(a.noise_std == a.noise_std)) return NSFF_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  // staged kernel: every 32-sample chunk of every plane must be a 16-byte aligned multiple of 16 bytes
  auto al16 = [](const void* p) { return ((uintptr_t)p & 15u) == 0; };
  const char* env_direct = getenv("NSFF_COMP_DIRECT");
  const bool force_direct = env_direct && env_direct[0] == '1';
  const bool cyc = warp && (a.xyzs_fw_bw || a.xyzs_bw_fw);
  bool staged = !force_direct && a.S % 4 == 0 && al16(a.raw) && al16(a.zs) && (!a.has_flow || al16(a.xyz)) &&
                (!warp || (al16(a.raw_fw) && al16(a.raw_bw) && a.raw_w_ld >= 7)) && (!cyc || (al16(a.xyz_fw) && al16(a.xyz_bw)));
  const StageLayout L = stage_layout(a.raw_ld, a.raw_w_ld, a.has_flow, warp, cyc);
  const size_t smem = (size_t)kCompWarps * kCompStages * L.bytes;
  if (smem > 225 * 1024) staged = false;
  if (staged) {
    int dev = 0, sms = 0;
    NSFF_CUDA_TRY(cudaGetDevice(&dev));
    NSFF_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const unsigned grid = (unsigned)(cdiv64(a.R, kCompWarps) < sms ? cdiv64(a.R, kCompWarps) : sms);
    auto go = [&](auto kern) -> int {
      NSFF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<grid, 32 * kCompWarps, smem, st>>>(a);
      return after_launch();
    };
    if (!a.has_transient) return go(composite_staged_kernel<false, false, false>);
    if (!a.has_flow) return go(composite_staged_kernel<true, false, false>);
    if (!warp) return go(composite_staged_kernel<true, true, false>);
    return go(composite_staged_kernel<true, true, true>);
  }
  const int rays_per_block = 8;
  dim3 grid((unsigned)cdiv64(a.R, rays_per_block)), block(32 * rays_per_block);
  if (!a.has_transient)
    composite_kernel<false, false, false><<<grid, block, 0, st>>>(a);
  else if (!a.has_flow)
    composite_kernel<true, false, false><<<grid, block, 0, st>>>(a);
  else if (!warp)
    composite_kernel<true, true, false><<<grid, block, 0, st>>>(a);
  else
    composite_kernel<true, true, true><<<grid, block, 0, st>>>(a);
  return after_launch();
}
