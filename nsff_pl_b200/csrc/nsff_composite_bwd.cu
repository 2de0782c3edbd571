// Backward of the composite (models/rendering.py:202-298, 98-140 differentiated): gradients of every
// differentiable output of nsff_composite w.r.t. the per-sample MLP outputs (raw, raw_fw, raw_bw) and the warped
// points.  One warp per ray; lane l owns samples l, l+32, ... (<= 8 per lane, S <= 256).  Phase 1 recomputes the
// forward alphas / transmittances (nothing was saved), phase 2 walks the chunks backwards with exclusive
// suffix-sum shuffle scans:   dL/dalpha_j = dw_j T_j - (sum_{i>j} G_i T_i) / (1 - alpha_j),  G_i = dL/dT_i.
// disocc* outputs are detached in the reference (rendering.py:290-291) and get no gradient.
#include "nsff_common.cuh"

namespace nsff {

constexpr int kMaxChunksBwd = 8;

__device__ __forceinline__ float warp_sum_b(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}
// exclusive suffix sum over the warp (lane i gets sum of lanes > i) + running carry from later chunks
__device__ __forceinline__ float warp_excl_suffix(float v, float& carry, int lane) {
  float inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    float o = __shfl_down_sync(0xffffffffu, inc, d);
    if (lane + d < 32) inc += o;
  }
  float excl = __shfl_down_sync(0xffffffffu, inc, 1);
  if (lane == 31) excl = 0.f;
  float out = carry + excl;
  carry += __shfl_sync(0xffffffffu, inc, 0);
  return out;
}
__device__ __forceinline__ float warp_excl_prod_b(float v, float& carry, int lane) {
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
__device__ __forceinline__ float dsoftplus(float x) { return x > 20.f ? 1.f : 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float dot3(const float* g, const float* v) { return g[0] * v[0] + g[1] * v[1] + g[2] * v[2]; }
__device__ __forceinline__ void ld3g(const float* p, int64_t ray, float (&o)[3]) {
  if (p) { o[0] = p[ray * 3]; o[1] = p[ray * 3 + 1]; o[2] = p[ray * 3 + 2]; }
  else { o[0] = o[1] = o[2] = 0.f; }
}
__device__ __forceinline__ float safe_div(float num, float den) { return num / fmaxf(den, 1e-20f); }

template <bool kFlow, bool kWarp>
__global__ void __launch_bounds__(128) composite_bwd_kernel(const nsff_composite_args a, const nsff_composite_grads g) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ray >= a.R) return;
  const int S = a.S;
  const int64_t base = ray * S;
  const int nch = (S + 31) / 32;

  float as_[kMaxChunksBwd], at_[kMaxChunksBwd], T_[kMaxChunksBwd], Ts_[kMaxChunksBwd];
  float xs_[kMaxChunksBwd], xt_[kMaxChunksBwd];                 // pre-softplus values (raw + noise)
  float aw_[2][kMaxChunksBwd], Tw_[2][kMaxChunksBwd], xw_[2][kMaxChunksBwd];
  // ---------------- phase 1: forward recompute ----------------
  {
    float carry = 1.f, carry_s = 1.f, carry_w[2] = {1.f, 1.f};
#pragma unroll
    for (int k = 0; k < kMaxChunksBwd; ++k) {
      if (k >= nch) break;
      const int s = k * 32 + lane;
      const bool ok = s < S;
      const int64_t i = base + (ok ? s : S - 1);
      const bool last = s >= S - 1;
      const float z = a.zs[i];
      const float dz = last ? 0.f : a.zs[i + 1] - z;
      const float d_s = last ? kLastDeltaStatic : dz, d_t = last ? kLastDeltaTrans : dz;
      const float* raw = a.raw + i * a.raw_ld;
      float x_s = raw[3];
      if (a.noise_static) x_s += a.noise_static[i] * a.noise_std;
      float x_t = raw[7];
      if (a.noise_trans) x_t += a.noise_trans[i] * a.noise_std;
      const float a_s = 1.f - expf(-d_s * softplus_f(x_s));
      const float a_t = 1.f - expf(-d_t * softplus_f(x_t));
      const float alpha = 1.f - (1.f - a_s) * (1.f - a_t);
      xs_[k] = x_s; xt_[k] = x_t; as_[k] = a_s; at_[k] = a_t;
      T_[k] = warp_excl_prod_b(ok ? 1.f - alpha : 1.f, carry, lane);
      Ts_[k] = warp_excl_prod_b(ok ? 1.f - a_s : 1.f, carry_s, lane);
      if (kWarp) {
#pragma unroll
        for (int dir = 0; dir < 2; ++dir) {
          const float* rw = (dir == 0 ? a.raw_fw : a.raw_bw) + i * a.raw_w_ld;
          const float* nz = dir == 0 ? a.noise_fw : a.noise_bw;
          float x_w = rw[3];
          if (nz) x_w += nz[i] * a.noise_std;
          const float a_tw = 1.f - expf(-d_t * softplus_f(x_w));
          const float a_w = 1.f - (1.f - a_s) * (1.f - a_tw);
          xw_[dir][k] = x_w; aw_[dir][k] = a_tw;
          Tw_[dir][k] = warp_excl_prod_b(ok ? 1.f - a_w : 1.f, carry_w[dir], lane);
        }
      }
    }
  }
  // ---------------- per-ray upstream gradients ----------------
  const float g_depth = g.depth ? g.depth[ray] : 0.f;
  const float g_talpha = g.transient_alpha ? g.transient_alpha[ray] : 0.f;
  const float g_sodepth = g.static_only_depth ? g.static_only_depth[ray] : 0.f;
  float g_rgb[3], g_trgb[3], g_sorgb[3], g_xyzf[3], g_ffw[3], g_fbw[3], g_xfw[3], g_xbw[3], g_rgbw[2][3];
  ld3g(g.rgb, ray, g_rgb); ld3g(g.transient_rgb, ray, g_trgb); ld3g(g.static_only_rgb, ray, g_sorgb);
  ld3g(g.xyz_fine, ray, g_xyzf); ld3g(g.flow_fw, ray, g_ffw); ld3g(g.flow_bw, ray, g_fbw);
  ld3g(g.xyz_fw_ray, ray, g_xfw); ld3g(g.xyz_bw_ray, ray, g_xbw);
  ld3g(g.rgb_fw, ray, g_rgbw[0]); ld3g(g.rgb_bw, ray, g_rgbw[1]);
  float gx_tot[3], gf_fw[3], gf_bw[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    gx_tot[c] = g_xyzf[c] + g_xfw[c] + g_xbw[c];      // xyz_fw = xyz_fine + flow_fw  (rendering.py:284,287)
    gf_fw[c] = g_ffw[c] + g_xfw[c];
    gf_bw[c] = g_fbw[c] + g_xbw[c];
  }
  const float g_trgb_sum = g_trgb[0] + g_trgb[1] + g_trgb[2];

  // ---------------- phase 2: reverse scans ----------------
  float sfx = 0.f, sfx_s = 0.f, sfx_w[2] = {0.f, 0.f};
#pragma unroll
  for (int kk = 0; kk < kMaxChunksBwd; ++kk) {
    const int k = nch - 1 - kk;
    if (k < 0) break;
    const int s = k * 32 + lane;
    const bool ok = s < S;
    const int64_t i = base + (ok ? s : S - 1);
    const bool last = s >= S - 1;
    const float z = a.zs[i];
    const float dz = last ? 0.f : a.zs[i + 1] - z;
    const float d_s = last ? kLastDeltaStatic : dz, d_t = last ? kLastDeltaTrans : dz;
    const float* raw = a.raw + i * a.raw_ld;
    const float a_s = as_[k], a_t = at_[k], T = T_[k], T_s = Ts_[k];
    const float alpha = 1.f - (1.f - a_s) * (1.f - a_t);
    const float cs[3] = {raw[0], raw[1], raw[2]}, ct[3] = {raw[4], raw[5], raw[6]};
    float ffw[3] = {0, 0, 0}, fbw[3] = {0, 0, 0}, xyz[3] = {0, 0, 0};
    if (kFlow) {
      ffw[0] = raw[8]; ffw[1] = raw[9]; ffw[2] = raw[10];
      fbw[0] = raw[11]; fbw[1] = raw[12]; fbw[2] = raw[13];
      xyz[0] = a.xyz[i * 3]; xyz[1] = a.xyz[i * 3 + 1]; xyz[2] = a.xyz[i * 3 + 2];
    }
    // direct gradients on the three weight planes
    float dw = (g.weights ? g.weights[i] : 0.f) + g_depth * z;
    if (kFlow) dw += dot3(gx_tot, xyz) + dot3(gf_fw, ffw) + dot3(gf_bw, fbw);
    float dws = (g.static_weights ? g.static_weights[i] : 0.f) + dot3(g_rgb, cs);
    float dwt = (g.transient_weights ? g.transient_weights[i] : 0.f) + dot3(g_rgb, ct) + g_talpha + dot3(g_trgb, ct) -
                kGreyBg * g_trgb_sum;
    if (!ok) dw = dws = dwt = 0.f;
    const float G = dw * alpha + dws * a_s + dwt * a_t;              // dL/dT_i
    const float S_after = warp_excl_suffix(ok ? G * T : 0.f, sfx, lane);
    const float dalpha = dw * T - safe_div(S_after, 1.f - alpha);
    float da_s = dalpha * (1.f - a_t) + dws * T;
    float da_t = dalpha * (1.f - a_s) + dwt * T;
    // static field alone
    const float dso = dot3(g_sorgb, cs) + g_sodepth * z;
    const float S_after_s = warp_excl_suffix(ok ? dso * a_s * T_s : 0.f, sfx_s, lane);
    da_s += dso * T_s - safe_div(S_after_s, 1.f - a_s);

    float d_cs[3], d_ct[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      d_cs[c] = g_rgb[c] * a_s * T + g_sorgb[c] * a_s * T_s;
      d_ct[c] = (g_rgb[c] + g_trgb[c]) * a_t * T;
    }
    if (kWarp) {
#pragma unroll
      for (int dir = 0; dir < 2; ++dir) {
        const float* rw = (dir == 0 ? a.raw_fw : a.raw_bw) + i * a.raw_w_ld;
        float* drw = (dir == 0 ? g.d_raw_fw : g.d_raw_bw);
        const float a_tw = aw_[dir][k], T_w = Tw_[dir][k];
        const float a_w = 1.f - (1.f - a_s) * (1.f - a_tw);
        const float cw[3] = {rw[0], rw[1], rw[2]};
        const float u = ok ? dot3(g_rgbw[dir], cs) : 0.f, v = ok ? dot3(g_rgbw[dir], cw) : 0.f;
        const float Gw = u * a_s + v * a_tw;
        const float S_after_w = warp_excl_suffix(ok ? Gw * T_w : 0.f, sfx_w[dir], lane);
        const float dalpha_w = -safe_div(S_after_w, 1.f - a_w);
        da_s += dalpha_w * (1.f - a_tw) + u * T_w;
        const float da_tw = dalpha_w * (1.f - a_s) + v * T_w;
#pragma unroll
        for (int c = 0; c < 3; ++c) d_cs[c] += g_rgbw[dir][c] * a_s * T_w;
        if (ok && drw) {
          float* o = drw + i * 7;
          const float* gc = dir == 0 ? g.xyzs_fw_bw : g.xyzs_bw_fw;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            o[c] = g_rgbw[dir][c] * a_tw * T_w;
            o[4 + c] = gc ? gc[i * 3 + c] : 0.f;            // xyzs_fw_bw = xyzs_fw + flow_back (rendering.py:231)
          }
          o[3] = da_tw * d_t * (1.f - a_tw) * dsoftplus(xw_[dir][k]);
        }
      }
    }
    if (ok && g.d_raw) {
      float* o = g.d_raw + i * a.raw_ld;
      const float gss = g.static_sigmas ? g.static_sigmas[i] : 0.f, gts = g.transient_sigmas ? g.transient_sigmas[i] : 0.f;
      o[0] = d_cs[0]; o[1] = d_cs[1]; o[2] = d_cs[2];
      o[3] = (da_s * d_s * (1.f - a_s) + gss) * dsoftplus(xs_[k]);
      o[4] = d_ct[0]; o[5] = d_ct[1]; o[6] = d_ct[2];
      o[7] = (da_t * d_t * (1.f - a_t) + gts) * dsoftplus(xt_[k]);
      if (kFlow) {
        const float w = alpha * T;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          o[8 + c] = gf_fw[c] * w;
          o[11 + c] = gf_bw[c] * w;
        }
      }
    }
  }
}

// Static-only field (no transient branch): outputs static_sigmas, static_weights (= weights), depth, rgb
// (rendering.py:206-208, 234-248, 256, 298).
__global__ void __launch_bounds__(128) composite_bwd_static_kernel(const nsff_composite_args a,
                                                                   const nsff_composite_grads g) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ray >= a.R) return;
  const int S = a.S;
  const int64_t base = ray * S;
  const int nch = (S + 31) / 32;
  float al_[kMaxChunksBwd], T_[kMaxChunksBwd], x_[kMaxChunksBwd];
  float carry = 1.f;
#pragma unroll
  for (int k = 0; k < kMaxChunksBwd; ++k) {
    if (k >= nch) break;
    const int s = k * 32 + lane;
    const bool ok = s < S;
    const int64_t i = base + (ok ? s : S - 1);
    const bool last = s >= S - 1;
    const float z = a.zs[i];
    const float d_s = last ? kLastDeltaStatic : a.zs[i + 1] - z;
    float x = a.raw[i * a.raw_ld + 3];
    if (a.noise_static) x += a.noise_static[i] * a.noise_std;
    const float al = 1.f - expf(-d_s * softplus_f(x));
    x_[k] = x; al_[k] = al;
    T_[k] = warp_excl_prod_b(ok ? 1.f - al : 1.f, carry, lane);
  }
  const float g_depth = g.depth ? g.depth[ray] : 0.f;
  float g_rgb[3];
  ld3g(g.rgb, ray, g_rgb);
  float sfx = 0.f;
#pragma unroll
  for (int kk = 0; kk < kMaxChunksBwd; ++kk) {
    const int k = nch - 1 - kk;
    if (k < 0) break;
    const int s = k * 32 + lane;
    const bool ok = s < S;
    const int64_t i = base + (ok ? s : S - 1);
    const bool last = s >= S - 1;
    const float z = a.zs[i];
    const float d_s = last ? kLastDeltaStatic : a.zs[i + 1] - z;
    const float* raw = a.raw + i * a.raw_ld;
    const float c[3] = {raw[0], raw[1], raw[2]};
    const float al = al_[k], T = T_[k];
    float dw = (g.static_weights ? g.static_weights[i] : 0.f) + g_depth * z + dot3(g_rgb, c);
    if (!ok) dw = 0.f;
    const float S_after = warp_excl_suffix(ok ? dw * al * T : 0.f, sfx, lane);
    const float dal = dw * T - safe_div(S_after, 1.f - al);
    if (ok) {
      float* o = g.d_raw + i * a.raw_ld;
      const float w = al * T;
      o[0] = g_rgb[0] * w; o[1] = g_rgb[1] * w; o[2] = g_rgb[2] * w;
      o[3] = (dal * d_s * (1.f - al) + (g.static_sigmas ? g.static_sigmas[i] : 0.f)) * dsoftplus(x_[k]);
    }
  }
}

}  // namespace nsff

using namespace nsff;

extern "C" int nsff_composite_backward(const nsff_composite_args* args, const nsff_composite_grads* grads,
                                       void* stream) {
  if (!args || !grads) return NSFF_EINVAL;
  const nsff_composite_args& a = *args;
  if (a.R < 0 || a.S < 1) return NSFF_EINVAL;
  if (a.R == 0) return NSFF_OK;
  if (a.S > 32 * kMaxChunksBwd) return NSFF_EUNSUPPORTED;
  if (!a.zs || !a.raw || !a.has_static_rgb) return NSFF_EUNSUPPORTED;   // training composites only (not sigma-only)
  if (!a.has_transient) {
    if (!grads->d_raw || a.raw_ld < 4) return NSFF_EINVAL;
    composite_bwd_static_kernel<<<(unsigned)cdiv64(a.R, 4), 128, 0, (cudaStream_t)stream>>>(a, *grads);
    return after_launch();
  }
  const bool warp = a.raw_fw != nullptr;
  if (warp && (!a.raw_bw || !a.has_flow || a.raw_w_ld != 7)) return NSFF_EINVAL;
  if (a.has_flow && (!a.xyz || a.raw_ld < 14)) return NSFF_EINVAL;
  if (!grads->d_raw) return NSFF_EINVAL;
  dim3 grid((unsigned)cdiv64(a.R, 4)), block(128);
  cudaStream_t st = (cudaStream_t)stream;
  if (warp) composite_bwd_kernel<true, true><<<grid, block, 0, st>>>(a, *grads);
  else if (a.has_flow) composite_bwd_kernel<true, false><<<grid, block, 0, st>>>(a, *grads);
  else composite_bwd_kernel<false, false><<<grid, block, 0, st>>>(a, *grads);
  return after_launch();
}
