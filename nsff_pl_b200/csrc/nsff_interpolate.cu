// §8f "next" #1: test-time frame interpolation (models/rendering.py:365-460) with the forward "average" splat of
// models/softsplat.py:6-44,303-329.  The reference loops over the S sample planes on the host, launching 2 cupy
// kernels and 4 PCIe copies per plane; here all planes are splatted by ONE launch (projection of the dt-scaled scene
// flow fused in) and composited front to back by a second one.
//   splat buffer: (2 directions, S planes, h, w, 8) fp32 = rgb, alpha | weight-sum, 3 pad: one 32-byte sector per
//   pixel, updated by ONE 16-byte vector reduction (red.global.add.v4.f32) + one scalar; caller-owned scratch
#include "nsff_common.cuh"

namespace nsff {

constexpr int kSplatStride = 8;   // floats per pixel of the splat buffer

__device__ __forceinline__ void ndc2world_d(const float* p, float fx, float fy, float cx, float cy, float (&o)[3]) {
  // datasets/ray_utils.py:143-148 (eps = 1e-6)
  const float rz = 2.f / (p[2] - 1.f - 1e-6f);
  o[0] = -rz * p[0] * cx / fx;
  o[1] = -rz * p[1] * cy / fy;
  o[2] = rz;
}

__global__ void interp_splat_kernel(const nsff_interpolate_args a, float* __restrict__ buf) {
  const int64_t B = a.R * a.S;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2 * B) return;
  const int dir = idx >= B;
  const int64_t p = dir ? idx - B : idx;
  const int64_t r = p / a.S;
  const int s = (int)(p - r * a.S);
  const float* xyz = a.xyzs + p * 3;
  const float* fl = (dir ? a.flow_bw : a.flow_fw) + p * (dir ? a.ld_tp1 : a.ld_t);
  const float scale = dir ? 1.f - a.dt : a.dt;
  float w0[3], w1[3];
  ndc2world_d(xyz, a.fx, a.fy, a.cx, a.cy, w0);
  const float moved[3] = {xyz[0] + fl[0], xyz[1] + fl[1], xyz[2] + fl[2]};
  ndc2world_d(moved, a.fx, a.fy, a.cx, a.cy, w1);
  float q[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) q[c] = w0[c] + scale * (w1[c] - w0[c]);
  const float* P = a.P;
  const float u = P[0] * q[0] + P[1] * q[1] + P[2] * q[2] + P[3];
  const float v = P[4] * q[0] + P[5] * q[1] + P[6] * q[2] + P[7];
  const float d = P[8] * q[0] + P[9] * q[1] + P[10] * q[2] + P[11];
  const float ox = u / d, oy = v / d;          // target pixel = source pixel + (uv - grid)   (rendering.py:405-410)
  const float x0 = floorf(ox), y0 = floorf(oy);
  const float* rgb = (dir ? a.rgb_tp1 : a.rgb_t) + p * (dir ? a.ld_tp1 : a.ld_t);
  const float alpha = (dir ? a.alpha_tp1 : a.alpha_t)[p];
  const float val[4] = {rgb[0], rgb[1], rgb[2], alpha};
  float* plane = buf + ((int64_t)dir * a.S + s) * a.R * kSplatStride;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float xi = x0 + (float)(k & 1), yi = y0 + (float)(k >> 1);
    if (!(xi >= 0.f && xi < (float)a.w && yi >= 0.f && yi < (float)a.h)) continue;
    const float wgt = (1.f - fabsf(ox - xi)) * (1.f - fabsf(oy - yi));
    float* dst = plane + ((int64_t)yi * a.w + (int64_t)xi) * kSplatStride;
    atomicAdd(reinterpret_cast<float4*>(dst), make_float4(val[0] * wgt, val[1] * wgt, val[2] * wgt, val[3] * wgt));
    atomicAdd(dst + 4, wgt);
  }
}

__global__ void interp_composite_kernel(const nsff_interpolate_args a, const float* __restrict__ buf,
                                        float* __restrict__ rgb_out, float* __restrict__ depth_out) {
  int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= a.R) return;
  float acc[3] = {0.f, 0.f, 0.f}, acc_a = 0.f, depth = 0.f;
  for (int s = 0; s < a.S; ++s) {
    const float4 f4 = *reinterpret_cast<const float4*>(buf + ((int64_t)s * a.R + pix) * kSplatStride);
    const float4 b4 = *reinterpret_cast<const float4*>(buf + ((int64_t)(a.S + s) * a.R + pix) * kSplatStride);
    const float fw = buf[((int64_t)s * a.R + pix) * kSplatStride + 4];
    const float bw = buf[((int64_t)(a.S + s) * a.R + pix) * kSplatStride + 4];
    const float f[4] = {f4.x, f4.y, f4.z, f4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
    const float nf = fw == 0.f ? 1.f : fw, nb = bw == 0.f ? 1.f : bw;
    const float fa = f[3] / nf, ba = b[3] / nb;
    const int64_t p = pix * a.S + s;
    const float* srgb = a.static_rgb + p * a.ld_t;
    const float sa = a.static_alpha[p];
    const float ca = 1.f - (1.f - (fa * (1.f - a.dt) + ba * a.dt)) * (1.f - sa);
    const float t = 1.f - acc_a;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float cr = (f[c] / nf) * fa * (1.f - a.dt) + (b[c] / nb) * ba * a.dt + srgb[c] * sa;
      acc[c] += t * cr;
    }
    depth += t * ca * a.zs[p];
    acc_a += t * ca;
  }
  rgb_out[pix * 3 + 0] = acc[0];
  rgb_out[pix * 3 + 1] = acc[1];
  rgb_out[pix * 3 + 2] = acc[2];
  depth_out[pix] = depth;
}

}  // namespace nsff

using namespace nsff;

extern "C" size_t nsff_interpolate_scratch_floats(int64_t R, int S) { return (R < 0 || S < 0) ? 0 : (size_t)2 * S * R * nsff::kSplatStride; }

extern "C" int nsff_interpolate(const nsff_interpolate_args* args, float* scratch, float* rgb, float* depth,
                                void* stream) {
  if (!args || !scratch || !rgb || !depth) return NSFF_EINVAL;
  const nsff_interpolate_args& a = *args;
  if (a.R <= 0 || a.S < 1 || (int64_t)a.w * a.h != a.R) return NSFF_EINVAL;
  if (!a.xyzs || !a.zs || !a.static_rgb || !a.static_alpha || !a.rgb_t || !a.alpha_t || !a.flow_fw || !a.rgb_tp1 ||
      !a.alpha_tp1 || !a.flow_bw)
    return NSFF_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  NSFF_CUDA_TRY(cudaMemsetAsync(scratch, 0, nsff_interpolate_scratch_floats(a.R, a.S) * sizeof(float), st));
  const int64_t n = 2 * a.R * a.S;
  interp_splat_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, st>>>(a, scratch);
  int rc = after_launch();
  if (rc != NSFF_OK) return rc;
  interp_composite_kernel<<<(unsigned)cdiv64(a.R, 128), 128, 0, st>>>(a, scratch, rgb, depth);
  return after_launch();
}
