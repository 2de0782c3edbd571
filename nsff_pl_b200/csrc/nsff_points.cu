// Sample generation, positional encoding and camera-visibility kernels (memory-bound, trivial math).
//   a1  PosEmbedding.forward        models/nerf.py:17-30
//   a6  stratified depths + points  models/rendering.py:314-324,332,359
//   a8  camera visibility           models/rendering.py:190-200, datasets/ray_utils.py:127-181
#include "nsff_common.cuh"

namespace nsff {

thread_local cudaError_t g_last_cuda_error = cudaSuccess;
std::atomic<uint64_t> g_launch_count{0};

// ---- a1 ---------------------------------------------------------------------------------------
// One thread per output element so stores are fully coalesced; the three loads per point hit L1.
__global__ void pos_embedding_kernel(const float* __restrict__ x, int64_t n,
                                     const float* __restrict__ freqs, int n_freqs,
                                     float* __restrict__ out) {
  const int C = 3 + 6 * n_freqs;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * C) return;
  int64_t p = idx / C;
  int c = (int)(idx - p * C);
  float v;
  if (c < 3) {
    v = x[p * 3 + c];
  } else {
    int k = c - 3;
    int f = k / 6, r = k - f * 6;
    float arg = freqs[f] * x[p * 3 + (r % 3)];
    v = (r < 3) ? sinf(arg) : cosf(arg);
  }
  out[idx] = v;
}

// ---- a6 ---------------------------------------------------------------------------------------
// torch.linspace(0,1,S) is evaluated symmetrically (start + i*step for the first half,
// end - (S-1-i)*step for the second); reproduce that so zs is bit-identical to the oracle.
__device__ __forceinline__ float linspace01(int i, int S) {
  if (S == 1) return 0.f;
  float step = 1.f / (float)(S - 1);
  return (i < S / 2) ? __fmul_rn((float)i, step) : __fsub_rn(1.f, __fmul_rn((float)(S - 1 - i), step));
}

__global__ void sample_points_kernel(const float* __restrict__ rays, int64_t R, int S, float perturb,
                                     const float* __restrict__ rand, float* __restrict__ zs,
                                     float* __restrict__ xyz) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * S) return;
  int64_t r = idx / S;
  int s = (int)(idx - r * S);
  float z = linspace01(s, S);
  if (perturb > 0.f) {
    float zc = z;
    float lo = (s > 0) ? __fmul_rn(0.5f, __fadd_rn(linspace01(s - 1, S), zc)) : zc;
    float hi = (s < S - 1) ? __fmul_rn(0.5f, __fadd_rn(zc, linspace01(s + 1, S))) : zc;
    float u = __fmul_rn(perturb, rand[idx]);
    z = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), u));
  }
  zs[idx] = z;
  const float* ray = rays + r * 6;
#pragma unroll
  for (int c = 0; c < 3; ++c) xyz[idx * 3 + c] = __fadd_rn(ray[c], __fmul_rn(ray[3 + c], z));
}

__global__ void points_from_depths_kernel(const float* __restrict__ rays, const float* __restrict__ zs,
                                          int64_t R, int S, float* __restrict__ xyz) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * S) return;
  int64_t r = idx / S;
  float z = zs[idx];
  const float* ray = rays + r * 6;
#pragma unroll
  for (int c = 0; c < 3; ++c) xyz[idx * 3 + c] = __fadd_rn(ray[c], __fmul_rn(ray[3 + c], z));
}

// ---- a8 ---------------------------------------------------------------------------------------
struct CamK {
  float fx, fy, cx, cy;
};

__global__ void visibility_kernel(const float* __restrict__ xyz, int64_t B, CamK K,
                                  const float* __restrict__ w2c, int n_cams, float H, float W,
                                  float* __restrict__ vis) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= B) return;
  // ndc2world (ray_utils.py:143-148), eps = 1e-6
  float x = xyz[p * 3 + 0], y = xyz[p * 3 + 1], z = xyz[p * 3 + 2];
  float rz = __fdiv_rn(2.f, __fsub_rn(__fsub_rn(z, 1.f), 1e-6f));
  float wx = __fdiv_rn(__fmul_rn(__fmul_rn(-rz, x), K.cx), K.fx);
  float wy = __fdiv_rn(__fmul_rn(__fmul_rn(-rz, y), K.cy), K.fy);
  float wz = rz;
  float count = 0.f;
  for (int c = 0; c < n_cams; ++c) {
    const float* M = w2c + c * 12;
    float cx_ = M[0] * wx + M[1] * wy + M[2] * wz + M[3];
    float cy_ = M[4] * wx + M[5] * wy + M[6] * wz + M[7];
    float cz_ = M[8] * wx + M[9] * wy + M[10] * wz + M[11];
    if (!(cz_ < 0.f)) continue;             // behind the camera (front is -z)
    cy_ = -cy_;
    cz_ = -cz_;                             // "right down front"
    float u = (K.fx * cx_ + K.cx * cz_) / cz_;
    float v = (K.fy * cy_ + K.cy * cz_) / cz_;
    if (u >= 0.f && u < W && v >= 0.f && v < H) count += 1.f;
  }
  vis[p] = count;
}

}  // namespace nsff

using namespace nsff;

extern "C" {

int nsff_abi_version(void) { return NSFF_ABI_VERSION; }

const char* nsff_error_string(int code) {
  switch (code) {
    case NSFF_OK: return "ok";
    case NSFF_EINVAL: return "invalid argument (shape, null pointer or alignment)";
    case NSFF_EUNSUPPORTED: return "configuration not supported by the sm_100a kernels";
    case NSFF_ECUDA: return "CUDA runtime error";
    case NSFF_EARCH: return "device is not compute capability 10.x (tcgen05 required)";
    default: return "unknown nsff error";
  }
}

const char* nsff_last_cuda_error(void) { return cudaGetErrorString(g_last_cuda_error); }

int nsff_set_device(int device) {
  NSFF_CUDA_TRY(cudaSetDevice(device));
  return NSFF_OK;
}

uint64_t nsff_launch_count(void) { return g_launch_count.load(); }

int nsff_pos_embedding(const float* x, int64_t n, const float* freqs, int n_freqs, float* out, void* stream) {
  if (n < 0 || n_freqs < 0 || (n > 0 && (!x || !out)) || (n_freqs > 0 && !freqs)) return NSFF_EINVAL;
  if (n == 0) return NSFF_OK;
  int64_t total = n * (3 + 6 * n_freqs);
  pos_embedding_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, (cudaStream_t)stream>>>(x, n, freqs, n_freqs, out);
  return after_launch();
}

int nsff_sample_points(const float* rays, int64_t R, int S, float perturb, const float* rand, float* zs,
                       float* xyz, void* stream) {
  if (R < 0 || S < 1 || (R > 0 && (!rays || !zs || !xyz)) || (perturb > 0.f && !rand)) return NSFF_EINVAL;
  if (R == 0) return NSFF_OK;
  sample_points_kernel<<<(unsigned)cdiv64(R * S, 256), 256, 0, (cudaStream_t)stream>>>(rays, R, S, perturb, rand,
                                                                                        zs, xyz);
  return after_launch();
}

int nsff_points_from_depths(const float* rays, const float* zs, int64_t R, int S, float* xyz, void* stream) {
  if (R < 0 || S < 1 || (R > 0 && (!rays || !zs || !xyz))) return NSFF_EINVAL;
  if (R == 0) return NSFF_OK;
  points_from_depths_kernel<<<(unsigned)cdiv64(R * S, 256), 256, 0, (cudaStream_t)stream>>>(rays, zs, R, S, xyz);
  return after_launch();
}

int nsff_visibility(const float* xyz, int64_t B, const float* K, const float* w2c, int n_cams, int H, int W,
                    float* vis, void* stream) {
  if (B < 0 || n_cams < 0 || !K || (B > 0 && (!xyz || !vis)) || (n_cams > 0 && !w2c)) return NSFF_EINVAL;
  if (B == 0) return NSFF_OK;
  CamK k{K[0], K[1], K[2], K[3]};
  visibility_kernel<<<(unsigned)cdiv64(B, 256), 256, 0, (cudaStream_t)stream>>>(xyz, B, k, w2c, n_cams, (float)H,
                                                                                 (float)W, vis);
  return after_launch();
}

}  // extern "C"

// ---- f4: on-device ray generation ----------------------------------------------------------------
// datasets/ray_utils.py:7-106 (get_ray_directions -> get_rays -> get_ndc_rays) as used for a test frame by
// datasets/monocular.py:268-276: pixel (i, j) -> camera direction ((i-cx)/fx, -(j-cy)/fy, -1) -> world ray
// (normalised) from the camera centre -> NDC ray.  One thread per pixel, 24 B written per ray.
namespace nsff {
struct FrameCam {
  float fx, fy, cx, cy;
  float c2w[12];
  float near, shift_near;
};

__device__ __forceinline__ void ndc_ray(const FrameCam& c, const float (&o_in)[3], const float (&d)[3], float* out) {
  // ray_utils.py:85-104
  const float t = -(c.shift_near + o_in[2]) / d[2];
  const float o[3] = {o_in[0] + t * d[0], o_in[1] + t * d[1], o_in[2] + t * d[2]};
  const float ox_oz = o[0] / o[2], oy_oz = o[1] / o[2];
  const float sx = -1.f / (c.cx / c.fx), sy = -1.f / (c.cy / c.fy);
  const float o2 = 1.f + 2.f * c.near / o[2];
  out[0] = sx * ox_oz;
  out[1] = sy * oy_oz;
  out[2] = o2;
  out[3] = sx * (d[0] / d[2] - ox_oz);
  out[4] = sy * (d[1] / d[2] - oy_oz);
  out[5] = 1.f - o2;
}

__global__ void frame_rays_kernel(FrameCam c, int H, int W, float* __restrict__ rays) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)H * W) return;
  const int j = (int)(idx / W), i = (int)(idx - (int64_t)j * W);
  const float dc[3] = {((float)i - c.cx) / c.fx, -((float)j - c.cy) / c.fy, -1.f};
  float d[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) d[r] = dc[0] * c.c2w[r * 4 + 0] + dc[1] * c.c2w[r * 4 + 1] + dc[2] * c.c2w[r * 4 + 2];
  const float inv = 1.f / sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  d[0] *= inv; d[1] *= inv; d[2] *= inv;
  const float o[3] = {c.c2w[3], c.c2w[7], c.c2w[11]};
  ndc_ray(c, o, d, rays + idx * 6);
}

__global__ void ndc_rays_kernel(FrameCam c, const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                int64_t n, float* __restrict__ out_o, float* __restrict__ out_d) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const float o[3] = {rays_o[idx * 3], rays_o[idx * 3 + 1], rays_o[idx * 3 + 2]};
  const float d[3] = {rays_d[idx * 3], rays_d[idx * 3 + 1], rays_d[idx * 3 + 2]};
  float r[6];
  ndc_ray(c, o, d, r);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    out_o[idx * 3 + k] = r[k];
    out_d[idx * 3 + k] = r[3 + k];
  }
}
}  // namespace nsff

extern "C" int nsff_frame_rays(const float* K4, const float* c2w12, float near, float shift_near, int H, int W,
                               float* rays, void* stream) {
  if (!K4 || !c2w12 || !rays || H < 1 || W < 1) return NSFF_EINVAL;
  nsff::FrameCam c;
  c.fx = K4[0]; c.fy = K4[1]; c.cx = K4[2]; c.cy = K4[3];
  for (int i = 0; i < 12; ++i) c.c2w[i] = c2w12[i];
  c.near = near; c.shift_near = shift_near;
  nsff::frame_rays_kernel<<<(unsigned)nsff::cdiv64((int64_t)H * W, 256), 256, 0, (cudaStream_t)stream>>>(c, H, W, rays);
  return nsff::after_launch();
}

extern "C" int nsff_ndc_rays(const float* K4, float near, float shift_near, const float* rays_o, const float* rays_d,
                             int64_t n, float* out_o, float* out_d, void* stream) {
  if (!K4 || n < 0 || (n > 0 && (!rays_o || !rays_d || !out_o || !out_d))) return NSFF_EINVAL;
  if (n == 0) return NSFF_OK;
  nsff::FrameCam c = {};
  c.fx = K4[0]; c.fy = K4[1]; c.cx = K4[2]; c.cy = K4[3];
  c.near = near; c.shift_near = shift_near;
  nsff::ndc_rays_kernel<<<(unsigned)nsff::cdiv64(n, 256), 256, 0, (cudaStream_t)stream>>>(c, rays_o, rays_d, n, out_o, out_d);
  return nsff::after_launch();
}
