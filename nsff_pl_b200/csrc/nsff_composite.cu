// a9-a12: density -> alpha, joint static+dynamic alpha composite, the two flow-warped composites and
// the disocclusion weights (models/rendering.py:202-298 and 98-140) in ONE pass over the per-sample
// MLP outputs.  HBM-bound: every input element is read exactly once, every output written once.
//
// Mapping: one warp per ray.  Lane l owns samples l, l+32, l+64, ... so that each warp-wide load of a
// scalar plane is a contiguous 128 B line (multi-channel rows are 16-56 B apart; the L1 merges the
// sector re-use across the unrolled channel loads).  The exclusive transmittance product is a
// 5-step shuffle scan per 32-sample chunk with a running carry; four scans run side by side
// (joint, static-only, warped-forward, warped-backward).
// (A variant that staged each 32-sample chunk through shared memory with 128-bit loads measured 1.54 ms vs
// 0.99 ms for this one on the bench frame - fewer loads in flight per warp - and was dropped; DRAM traffic of
// both equals the algorithmic 3.85 GB.)
#include "nsff_common.cuh"

namespace nsff {

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

// The kernel is latency-bound (ncu: long-scoreboard stalls, 35 % issue slots) and its 78 registers allowed only
// 3 blocks = 24 warps per SM; capping at 64 registers (4 blocks, 80 bytes of spills) measured 1.01 -> 0.89 ms on the
// bench frame (3.8 -> 4.3 TB/s); 5 and 6 blocks spill more and are slower (tools/comp_occ.sh).
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
  const int col_sig_s = a.has_static_rgb ? 3 : 0;
  const int col_rgb_t = 4;
  const int col_sig_t = a.has_static_rgb ? 7 : 1;

  float carry = 1.f, carry_s = 1.f, carry_fw = 1.f, carry_bw = 1.f;
  float depth = 0.f, t_alpha = 0.f, so_depth = 0.f, occ_fw = 0.f, occ_bw = 0.f;
  Vec3 rgb_s_acc{0, 0, 0}, rgb_t_acc{0, 0, 0}, so_rgb{0, 0, 0}, xyz_acc{0, 0, 0}, ffw_acc{0, 0, 0},
      fbw_acc{0, 0, 0}, rgbfw_acc{0, 0, 0}, rgbbw_acc{0, 0, 0};

  for (int s0 = 0; s0 < S; s0 += 32) {
    const int s = s0 + lane;
    const bool ok = s < S;
    const int64_t i = base + (ok ? s : S - 1);
    const float z = a.zs[i];
    const bool last = s >= S - 1;
    const float dz = last ? 0.f : a.zs[i + 1] - z;
    const float d_s = last ? kLastDeltaStatic : dz;
    const float d_t = last ? kLastDeltaTrans : dz;
    const float* raw = a.raw + i * a.raw_ld;

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
    const float T = warp_excl_prod(om, carry, lane);
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
      depth += w * z;
      if (kTransient) {
        acc3(rgb_s_acc, w_s, c_s);
        acc3(rgb_t_acc, w_t, c_t);
        t_alpha += w_t;
      } else {
        acc3(rgb_s_acc, w, c_s);
      }
    }

    if (kTransient) {
      // static field alone, with its own transmittance (rendering.py:270-278)
      const float T_s = warp_excl_prod(ok ? 1.f - a_s : 1.f, carry_s, lane);
      if (ok) {
        const float w_os = a_s * T_s;
        acc3(so_rgb, w_os, c_s);
        so_depth += w_os * z;
      }
    }

    if (kFlow) {
      if (ok) {
        acc3(xyz_acc, w, ld3(a.xyz + i * 3));
        acc3(ffw_acc, w, ld3(raw + 8));
        acc3(fbw_acc, w, ld3(raw + 11));
      }
    }

    if (kWarp) {
      // render_transient_warping (rendering.py:98-140): warped dynamic field over the CURRENT static one
#pragma unroll
      for (int dir = 0; dir < 2; ++dir) {
        const float* rw = (dir == 0 ? a.raw_fw : a.raw_bw) + i * a.raw_w_ld;
        const float* nz = dir == 0 ? a.noise_fw : a.noise_bw;
        float raw_w = rw[3];
        if (nz) raw_w += nz[i] * a.noise_std;
        const float a_tw = 1.f - expf(-d_t * softplus_f(raw_w));
        const float a_w = 1.f - (1.f - a_s) * (1.f - a_tw);
        const float T_w = warp_excl_prod(ok ? 1.f - a_w : 1.f, dir == 0 ? carry_fw : carry_bw, lane);
        if (ok) {
          const float w_tw = a_tw * T_w;
          Vec3& acc = dir == 0 ? rgbfw_acc : rgbbw_acc;
          acc3(acc, a_s * T_w, c_s);
          acc3(acc, w_tw, ld3(rw));
          const float occ = w_tw - w_t;                     // rendering.py:290-295 (detached)
          float* dis = dir == 0 ? a.disoccs_fw : a.disoccs_bw;
          if (dis) dis[i] = 1.f - fabsf(occ);
          (dir == 0 ? occ_fw : occ_bw) += occ;
          float* cyc = dir == 0 ? a.xyzs_fw_bw : a.xyzs_bw_fw;     // rendering.py:231-232
          if (cyc) {
            const float* xw = (dir == 0 ? a.xyz_fw : a.xyz_bw) + i * 3;
            cyc[i * 3 + 0] = xw[0] + rw[4];
            cyc[i * 3 + 1] = xw[1] + rw[5];
            cyc[i * 3 + 2] = xw[2] + rw[6];
          }
        }
      }
    }
  }

  // ---- per-ray reductions ------------------------------------------------------------------
  depth = warp_sum(depth);
  if (a.depth && lane == 0) a.depth[ray] = depth;
  if (kTransient) {
    Vec3 s = rgb_s_acc, t = rgb_t_acc;
    s.x = warp_sum(s.x); s.y = warp_sum(s.y); s.z = warp_sum(s.z);
    t.x = warp_sum(t.x); t.y = warp_sum(t.y); t.z = warp_sum(t.z);
    t_alpha = warp_sum(t_alpha);
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
    if (a.static_only_rgb) st3_reduced(a.static_only_rgb, ray, so_rgb, lane);
    so_depth = warp_sum(so_depth);
    if (a.static_only_depth && lane == 0) a.static_only_depth[ray] = so_depth;
  } else {
    if (a.rgb) st3_reduced(a.rgb, ray, rgb_s_acc, lane);
  }
  if (kFlow) {
    Vec3 x = xyz_acc, f = ffw_acc, b = fbw_acc;
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
    if (a.rgb_fw) st3_reduced(a.rgb_fw, ray, rgbfw_acc, lane);
    if (a.rgb_bw) st3_reduced(a.rgb_bw, ray, rgbbw_acc, lane);
    occ_fw = warp_sum(occ_fw);
    occ_bw = warp_sum(occ_bw);
    if (lane == 0) {
      if (a.disocc_fw) a.disocc_fw[ray] = 1.f - fabsf(occ_fw);
      if (a.disocc_bw) a.disocc_bw[ray] = 1.f - fabsf(occ_bw);
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
  const int rays_per_block = 8;
  dim3 grid((unsigned)cdiv64(a.R, rays_per_block)), block(32 * rays_per_block);
  cudaStream_t st = (cudaStream_t)stream;
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
