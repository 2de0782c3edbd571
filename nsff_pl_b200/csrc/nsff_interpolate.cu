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

// One block = 32 consecutive pixels of an image row x 8 sample planes of one direction.  The inputs are ray-major
// ((R,S,..): a ray's 8 consecutive samples are contiguous), the splat buffer is plane-major: the block loads and projects
// with thread = (ray, plane) - contiguous 96 / 448-byte segments per ray - hands the projected samples over through
// shared memory and splats with warp = (plane, 32 neighbouring pixels), so that the reductions of a warp land on ~66
// neighbouring pixels of ONE plane: their 32-byte sectors are hit 2-4 times while they are still in the L2 (the 1.8 GB
// buffer of a 512x288x192 frame does not fit: with one thread per (ray, sample) in ray-major order every reduction
// cost a DRAM round trip - 8.0 ms per frame).
constexpr int kSplatRays = 32, kSplatPlanes = 8;

__global__ void __launch_bounds__(kSplatRays * kSplatPlanes) interp_splat_kernel(const nsff_interpolate_args a,
                                                                                   float* __restrict__ buf) {
  __shared__ float sm[6][kSplatPlanes][kSplatRays + 1];       // ox, oy, rgb * 3, alpha  (NaN ox: nothing to splat)
  const int dir = blockIdx.z;
  const int t = threadIdx.x;
  {
    const int i = t >> 3, j = t & 7;                            // load phase: ray i, plane j
    const int64_t r = (int64_t)blockIdx.x * kSplatRays + i;
    const int sidx = blockIdx.y * kSplatPlanes + j;
    float ox = __int_as_float(0x7fc00000), oy = 0.f, val[4] = {0.f, 0.f, 0.f, 0.f};
    if (r < a.R && sidx < a.S) {
      const int64_t p = r * a.S + sidx;
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
      ox = u / d, oy = v / d;                    // target pixel = source pixel + (uv - grid)   (rendering.py:405-410)
      const float* rgb = (dir ? a.rgb_tp1 : a.rgb_t) + p * (dir ? a.ld_tp1 : a.ld_t);
      val[0] = rgb[0], val[1] = rgb[1], val[2] = rgb[2], val[3] = (dir ? a.alpha_tp1 : a.alpha_t)[p];
    }
    sm[0][j][i] = ox, sm[1][j][i] = oy;
#pragma unroll
    for (int c = 0; c < 4; ++c) sm[2 + c][j][i] = val[c];
  }
  __syncthreads();
  const int j = t >> 5, i = t & 31;                             // splat phase: plane j, ray i
  const int sidx = blockIdx.y * kSplatPlanes + j;
  const float ox = sm[0][j][i], oy = sm[1][j][i];
  if (!(ox == ox) || sidx >= a.S) return;                       // (a NaN projection splats nowhere, as in the reference)
  const float val[4] = {sm[2][j][i], sm[3][j][i], sm[4][j][i], sm[5][j][i]};
  const float x0 = floorf(ox), y0 = floorf(oy);
  float* plane = buf + ((int64_t)dir * a.S + sidx) * a.R * kSplatStride;
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
  const dim3 grid((unsigned)cdiv64(a.R, kSplatRays), (unsigned)cdiv64(a.S, kSplatPlanes), 2);
  if (grid.y > 65535u) return NSFF_EUNSUPPORTED;
  interp_splat_kernel<<<grid, kSplatRays * kSplatPlanes, 0, st>>>(a, scratch);
  int rc = after_launch();
  if (rc != NSFF_OK) return rc;
  interp_composite_kernel<<<(unsigned)cdiv64(a.R, 128), 128, 0, st>>>(a, scratch, rgb, depth);
  return after_launch();
}
