// fp32 verification path of NeRF.forward (models/nerf.py:118-213): the same layer schedule as the
// tcgen05 kernel, on CUDA cores, straight from the fp32 nn.Parameters.  It exists so that every
// non-GEMM stage (encoding, column packing, heads, flow zeroing, compositing) can be checked against
// the oracle to ~1e-6 on the GPU, independently of bf16 rounding.  It is NOT a fallback: the
// renderer only takes it when the caller selects precision='fp32'.
#include "nsff_common.cuh"

namespace nsff {

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_SIGMOID = 2, ACT_FLOW = 3 };

// C[m, n] = act(sum_k A[m,k] * W[n,k] + bias[n]);  A (M,K) lda, W (N,K) ldw, C (M,N) ldc.
// 64x64 block tile, 16-deep K slab, 256 threads each owning a 4x4 micro-tile.
template <int BM, int BN, int BK>
__global__ void __launch_bounds__(256) linear_f32_kernel(const float* __restrict__ A, int lda,
                                                         const float* __restrict__ Wt, int ldw,
                                                         const float* __restrict__ bias, float* __restrict__ C,
                                                         int ldc, int64_t M, int N, int K, int act,
                                                         const float* __restrict__ zs, float flow_scale) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int it = 0; it < (BM * BK) / 256; ++it) {
      int e = tid + it * 256;
      int kk = e % BK, mm = e / BK;
      int64_t m = m0 + mm;
      int k = k0 + kk;
      As[kk][mm] = (m < M && k < K) ? A[m * lda + k] : 0.f;
    }
#pragma unroll
    for (int it = 0; it < (BN * BK) / 256; ++it) {
      int e = tid + it * 256;
      int kk = e % BK, nn = e / BK;
      int n = n0 + nn;
      int k = k0 + kk;
      Ws[kk][nn] = (n < N && k < K) ? Wt[(int64_t)n * ldw + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = Ws[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const bool far = (act == ACT_FLOW) && zs != nullptr && zs[m] > kZFar;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (act == ACT_RELU) v = fmaxf(v, 0.f);
      else if (act == ACT_SIGMOID) v = sigmoid_f(v);
      else if (act == ACT_FLOW) v = far ? 0.f : flow_scale * tanhf(v);
      C[m * ldc + n] = v;
    }
  }
}

// dst[m, 0:n] = src[m, 0:n]   (column-block copy between strided matrices)
__global__ void copy_cols_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd,
                                 int64_t M, int n) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * n) return;
  int64_t m = idx / n;
  int c = (int)(idx - m * n);
  dst[m * ldd + c] = src[m * lds + c];
}

// Builds the reference's MLP input row [PE(xyz) | PE(dir) | a | t]  (rendering.py:164-173) from raw
// per-point / per-ray data, one thread per output element.
struct EmbedArgs {
  nsff_points p;
  int in_xyz, in_dir, in_a, in_t;
  float* x;
  int ld;
};

__global__ void embed_points_kernel(EmbedArgs e) {
  const int C = e.in_xyz + e.in_dir + e.in_a + e.in_t;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= e.p.B * C) return;
  int64_t pnt = idx / C;
  int c = (int)(idx - pnt * C);
  int64_t ray = pnt / e.p.S;
  float v;
  if (c < e.in_xyz + e.in_dir) {
    const bool is_dir = c >= e.in_xyz;
    int k = is_dir ? c - e.in_xyz : c;
    int comp = k < 3 ? k : (k - 3) % 3;
    float base;
    if (is_dir) {
      base = e.p.dirs[ray * 3 + comp];
    } else {
      base = e.p.xyz[pnt * 3 + comp];
      if (e.p.offset) base = __fadd_rn(base, e.p.offset[pnt * e.p.offset_ld + comp]);
      if (e.p.xyz_out && k < 3) e.p.xyz_out[pnt * 3 + comp] = base;
    }
    if (k < 3) {
      v = base;
    } else {
      int f = (k - 3) / 6;
      bool is_cos = ((k - 3) % 6) >= 3;
      float arg = __fmul_rn(exp2f((float)f), base);     // freqs = 2^linspace(0, N-1, N)  (nerf.py:13)
      v = is_cos ? cosf(arg) : sinf(arg);
    }
  } else if (c < e.in_xyz + e.in_dir + e.in_a) {
    int k = c - e.in_xyz - e.in_dir;
    int64_t row = e.p.a_index ? e.p.a_index[ray] : ray;
    v = e.p.a_table ? e.p.a_table[row * e.in_a + k] : 0.f;
  } else {
    int k = c - e.in_xyz - e.in_dir - e.in_a;
    int64_t row = ray;
    if (e.p.t_index) {
      row = e.p.t_index[ray] + e.p.t_shift;
      row = row < 0 ? 0 : (row > e.p.max_t ? e.p.max_t : row);
    }
    v = e.p.t_table ? e.p.t_table[row * e.in_t + k] : 0.f;   // unused columns when the dynamic branch is off
  }
  e.x[pnt * e.ld + c] = v;
}

static int linear(const float* A, int lda, const float* W, int ldw, const float* b, float* C, int ldc, int64_t M,
                  int N, int K, int act, const float* zs, float flow_scale, cudaStream_t st) {
  if (!A || !W || !C) return NSFF_EINVAL;
  dim3 grid((unsigned)cdiv64(M, 64), (unsigned)((N + 63) / 64));
  linear_f32_kernel<64, 64, 16><<<grid, 256, 0, st>>>(A, lda, W, ldw, b, C, ldc, M, N, K, act, zs, flow_scale);
  return after_launch();
}

static int copy_cols(const float* src, int lds, float* dst, int ldd, int64_t M, int n, cudaStream_t st) {
  if (n <= 0) return NSFF_OK;
  copy_cols_kernel<<<(unsigned)cdiv64(M * n, 256), 256, 0, st>>>(src, lds, dst, ldd, M, n);
  return after_launch();
}

}  // namespace nsff

using namespace nsff;

#define TRY(x)                \
  do {                        \
    int _rc = (x);            \
    if (_rc != NSFF_OK) return _rc; \
  } while (0)

static int in_cols(const nsff_field_desc* d) { return d->in_xyz + d->in_dir + d->in_a + d->in_t; }
static int hid_ld(const nsff_field_desc* d) { return d->in_xyz + d->in_t + d->W + d->in_dir + d->in_a; }

extern "C" size_t nsff_field_fp32_scratch_floats(const nsff_field_desc* d, int64_t B) {
  if (!d || B < 0) return 0;
  return (size_t)B * (size_t)(in_cols(d) + 3 * hid_ld(d));
}

// One trunk: 8 x (Linear+ReLU) with the input re-read at the skip layer (nerf.py:163-167 / 192-196).
// `hs` holds [skip input | layer-(skip) output] contiguously so the skip layer is a single GEMM.
static int trunk(const nsff_field_desc* d, const float* const* w, const float* const* b, int n_in, float* hs,
                 float* ha, float* hb, int ld, int64_t B, float** h_out, cudaStream_t st) {
  const int W = d->W;
  const float* cur = hs;   // layer-1 input lives in hs[:, 0:n_in]
  int cur_k = n_in;
  float* pp[2] = {ha, hb};
  int flip = 0;
  for (int i = 0; i < d->D; ++i) {
    float* dst;
    if (i + 1 == d->skip) dst = hs + n_in;        // output of the layer right before the skip layer
    else dst = pp[flip], flip ^= 1;
    TRY(linear(cur, ld, w[i], cur_k, b[i], dst, ld, B, W, cur_k, ACT_RELU, nullptr, 0.f, st));
    if (i + 1 == d->skip) cur = hs, cur_k = n_in + W;
    else cur = dst, cur_k = W;
  }
  *h_out = const_cast<float*>(cur);
  return NSFF_OK;
}

extern "C" int nsff_field_forward_fp32(const nsff_field_desc* d, const nsff_field_params* p, const nsff_points* pts,
                                       const float* x_embedded, int32_t x_ld, int64_t B_embedded, int out_mask,
                                       float* out, int32_t out_ld, float* scratch, void* stream) {
  if (!d || !p || !out || !scratch) return NSFF_EINVAL;
  if ((pts == nullptr) == (x_embedded == nullptr)) return NSFF_EINVAL;
  if (d->D < 2 || d->D > 8 || d->skip < 1 || d->skip >= d->D || d->W < 1) return NSFF_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const bool sigma_only = out_mask & NSFF_OUT_SIGMA_ONLY;
  const bool want_s = out_mask & NSFF_OUT_STATIC, want_t = out_mask & NSFF_OUT_TRANSIENT;
  const bool want_fw = out_mask & NSFF_OUT_FLOW_FW, want_bw = out_mask & NSFF_OUT_FLOW_BW;
  if (!want_s && !want_t) return NSFF_EINVAL;
  if (want_t && !d->has_transient) return NSFF_EINVAL;
  if ((want_fw || want_bw) && (!want_t || !d->has_flow)) return NSFF_EINVAL;
  const int W = d->W, ld = hid_ld(d);
  int64_t B;
  const float* X;
  int xld;
  // column offsets inside the embedded row
  int c_dir = d->in_xyz, c_a = c_dir + d->in_dir, c_t = c_a + d->in_a;
  float* Xs = scratch;
  if (pts) {
    B = pts->B;
    if (B < 0 || pts->S < 1 || !pts->xyz) return NSFF_EINVAL;
    if (B == 0) return NSFF_OK;
    if (d->in_dir > 0 && !pts->dirs) return NSFF_EINVAL;
    if (want_t && d->in_t > 0 && !pts->t_table) return NSFF_EINVAL;
    if (want_s && !sigma_only && d->use_viewdir && d->in_a > 0 && !pts->a_table) return NSFF_EINVAL;
    if (d->in_xyz != 63 || (d->in_dir != 27 && d->in_dir != 0)) return NSFF_EUNSUPPORTED;  // PosEmbedding(9,10)/(3,4)
    EmbedArgs e{*pts, d->in_xyz, d->in_dir, d->in_a, d->in_t, Xs, in_cols(d)};
    embed_points_kernel<<<(unsigned)cdiv64(B * in_cols(d), 256), 256, 0, st>>>(e);
    TRY(after_launch());
    X = Xs;
    xld = in_cols(d);
  } else {
    B = B_embedded;
    if (B < 0 || x_ld < in_cols(d)) return NSFF_EINVAL;
    if (B == 0) return NSFF_OK;
    X = x_embedded;
    xld = x_ld;
  }
  float* hs = scratch + (size_t)B * in_cols(d);
  float* ha = hs + (size_t)B * ld;
  float* hb = ha + (size_t)B * ld;
  const float* zs = pts ? pts->zs : nullptr;
  int col = 0;

  if (want_s) {
    TRY(copy_cols(X, xld, hs, ld, B, d->in_xyz, st));
    float* h8;
    TRY(trunk(d, p->static_w, p->static_b, d->in_xyz, hs, ha, hb, ld, B, &h8, st));
    if (sigma_only) {
      TRY(linear(h8, ld, p->static_sigma_w, W, p->static_sigma_b, out + col, out_ld, B, 1, W, ACT_NONE, nullptr, 0, st));
      col += 1;
    } else {
      TRY(linear(h8, ld, p->static_sigma_w, W, p->static_sigma_b, out + col + 3, out_ld, B, 1, W, ACT_NONE, nullptr, 0, st));
      // final (no activation) lands in hs[:, 0:W] so [feat | dir | a] is one contiguous GEMM operand
      float* feat = hs;
      if (h8 == hs) return NSFF_EUNSUPPORTED;   // skip == D would alias; never the case for D=8, skip=4
      TRY(linear(h8, ld, p->static_final_w, W, p->static_final_b, feat, ld, B, W, W, ACT_NONE, nullptr, 0, st));
      const float* rgb_in = feat;
      if (d->use_viewdir) {
        TRY(copy_cols(X + c_dir, xld, feat + W, ld, B, d->in_dir + d->in_a, st));
        float* dst = (h8 == ha) ? hb : ha;
        TRY(linear(feat, ld, p->static_dir_w, W + d->in_dir + d->in_a, p->static_dir_b, dst, ld, B, W,
                   W + d->in_dir + d->in_a, ACT_RELU, nullptr, 0, st));
        rgb_in = dst;
      }
      TRY(linear(rgb_in, ld, p->static_rgb_w, W, p->static_rgb_b, out + col, out_ld, B, 3, W, ACT_SIGMOID, nullptr, 0, st));
      col += 4;
    }
  }
  if (want_t) {
    TRY(copy_cols(X, xld, hs, ld, B, d->in_xyz, st));
    TRY(copy_cols(X + c_t, xld, hs + d->in_xyz, ld, B, d->in_t, st));
    float* h8;
    TRY(trunk(d, p->trans_w, p->trans_b, d->in_xyz + d->in_t, hs, ha, hb, ld, B, &h8, st));
    float* feat = (h8 == ha) ? hb : ha;
    TRY(linear(h8, ld, p->trans_final_w, W, p->trans_final_b, feat, ld, B, W, W, ACT_NONE, nullptr, 0, st));
    if (sigma_only) {
      TRY(linear(feat, ld, p->trans_sigma_w, W, p->trans_sigma_b, out + col, out_ld, B, 1, W, ACT_NONE, nullptr, 0, st));
      col += 1;
    } else {
      TRY(linear(feat, ld, p->trans_rgb_w, W, p->trans_rgb_b, out + col, out_ld, B, 3, W, ACT_SIGMOID, nullptr, 0, st));
      TRY(linear(feat, ld, p->trans_sigma_w, W, p->trans_sigma_b, out + col + 3, out_ld, B, 1, W, ACT_NONE, nullptr, 0, st));
      col += 4;
      if (want_fw) {
        TRY(linear(feat, ld, p->trans_fw_w, W, p->trans_fw_b, out + col, out_ld, B, 3, W, ACT_FLOW, zs, d->flow_scale, st));
        col += 3;
      }
      if (want_bw) {
        TRY(linear(feat, ld, p->trans_bw_w, W, p->trans_bw_b, out + col, out_ld, B, 3, W, ACT_FLOW, zs, d->flow_scale, st));
        col += 3;
      }
    }
  }
  return NSFF_OK;
}
