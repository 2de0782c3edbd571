// §8f "next" #2: NeRFWLoss (losses.py:31-171) as three launches instead of ~300 ATen kernels (forward + autograd):
//   loss_stats_kernel   one block: the global statistics the per-ray terms need - scale of the shift/scale-invariant depth
//                       loss (losses.py:8-29: s = mean|d - median|, and the sums its gradient needs), the means of the
//                       detached disocclusion weights (losses.py:131-145), the number of rays with a valid 2D-3D flow
//                       target (losses.py:118-127).  For R <= kMedianMaxR it also finds the three medians (bitonic
//                       sort in shared memory: exact lower medians, as torch.median); above that the caller supplies them.
//   loss_kernel<false>  one warp per ray: every per-ray term, block-reduced, one atomicAdd per term and block.
//   loss_kernel<true>   same traversal, writes d(sum_k g_k * term_k) / d(input) for every differentiable input.
//   loss_finish_kernel  divides the sums by their denominators (R, or the valid-ray counts).
// Term order (nsff_loss_term): col, disp, entropy, cross_entropy, flow_fw, flow_bw, pho, cyc, reg_temp_sm, reg_min,
// reg_sp_sm.  topk < 1 (hard example mining, losses.py:163-165) is not implemented: the Python wrapper refuses it.
#include "nsff_common.cuh"

namespace nsff {

enum { ST_S_PRED = 0, ST_S_GT, ST_NEQ, ST_A, ST_B, ST_CSGN, ST_S_PRED_C, ST_NEQ_C, ST_A_C, ST_B_C, ST_CSGN_C,
       ST_M_DFW, ST_M_DBW, ST_M_DSFW, ST_M_DSBW, ST_NV_FW, ST_NV_BW, ST_COUNT };

__device__ __forceinline__ float sgnf(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }

__device__ __forceinline__ float block_sum(float v, float* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < nw; ++i) t += sh[i];
  return t;
}

// world point and (optionally) the pull-back of a world-space gradient  (datasets/ray_utils.py:127-150, eps = 1e-6)
struct Cam { float cxfx, cyfy; };
__device__ __forceinline__ void ndc2world3(const float* p, Cam c, float (&o)[3]) {
  const float rz = 2.f / (p[2] - 1.f - 1e-6f);
  o[0] = -rz * p[0] * c.cxfx;
  o[1] = -rz * p[1] * c.cyfy;
  o[2] = rz;
}
__device__ __forceinline__ void ndc2world_pullback(const float* p, Cam c, const float (&g)[3], float (&d)[3]) {
  const float rz = 2.f / (p[2] - 1.f - 1e-6f);
  const float drz = -0.5f * rz * rz;                      // d rz / d z
  d[0] = g[0] * (-rz * c.cxfx);
  d[1] = g[1] * (-rz * c.cyfy);
  d[2] = (g[0] * (-p[0] * c.cxfx) + g[1] * (-p[1] * c.cyfy) + g[2]) * drz;
}

constexpr int kMedianMaxR = 2048;

// lower median (sorted[(n - 1) / 2], torch.median's convention) of n <= kMedianMaxR values: bitonic sort in shared memory
// (padded with +inf to a power of two; 55 compare-exchange stages for 1024 values, ~3 us).  One block.
__device__ void block_median(const float* __restrict__ src, float sign, int n, float* vals, float* out) {
  int m = 1;
  while (m < n) m <<= 1;
  for (int i = threadIdx.x; i < m; i += blockDim.x) vals[i] = i < n ? sign * src[i] : INFINITY;
  __syncthreads();
  for (int k = 2; k <= m; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const int p = i ^ j;
        if (p > i) {
          const float a = vals[i], b = vals[p];
          if ((a > b) == ((i & k) == 0)) vals[i] = b, vals[p] = a;
        }
      }
      __syncthreads();
    }
  if (threadIdx.x == 0) *out = vals[(n - 1) / 2];
  __syncthreads();
}

__global__ void __launch_bounds__(1024) loss_stats_kernel(const nsff_loss_args a) {
  __shared__ float sh[32];
  __shared__ float s_pred, s_gt, s_pred_c;
  __shared__ float med[3];
  __shared__ float vals[kMedianMaxR];
  if (a.R <= kMedianMaxR) {
    block_median(a.depth_fine, 1.f, (int)a.R, vals, &med[0]);
    block_median(a.disps, -1.f, (int)a.R, vals, &med[1]);
    if (a.depth_coarse) block_median(a.depth_coarse, 1.f, (int)a.R, vals, &med[2]);
    if (threadIdx.x < 3 && (threadIdx.x < 2 || a.depth_coarse)) a.medians[threadIdx.x] = med[threadIdx.x];
  } else {
    if (threadIdx.x < 3) med[threadIdx.x] = a.medians[threadIdx.x];
    __syncthreads();
  }
  const float t_pred = med[0], t_gt = med[1], t_pred_c = a.depth_coarse ? med[2] : 0.f;
  const float invR = 1.f / (float)a.R;
  float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // |d-t|, |g-t|, |dc-t|, disocc_fw, disocc_bw, valid_fw, valid_bw, n_eq
  float neq_c = 0.f, csgn = 0.f, csgn_c = 0.f;
  for (int64_t r = threadIdx.x; r < a.R; r += blockDim.x) {
    const float d = a.depth_fine[r], g = -a.disps[r];
    v[0] += fabsf(d - t_pred), v[1] += fabsf(g - t_gt);
    v[7] += d == t_pred ? 1.f : 0.f;
    csgn += sgnf(d - t_pred);
    if (a.depth_coarse) {
      const float dc = a.depth_coarse[r];
      v[2] += fabsf(dc - t_pred_c), neq_c += dc == t_pred_c ? 1.f : 0.f, csgn_c += sgnf(dc - t_pred_c);
    }
    if (a.has_flow) {
      v[3] += a.disocc_fw[r], v[4] += a.disocc_bw[r];
      const int64_t cam = a.cam_ids[r], t = a.ts[r];
      const float* K = a.Ks + cam * 9;
      const Cam c{K[2] / K[0], K[5] / K[4]};
      for (int dir = 0; dir < 2; ++dir) {
        int64_t tt = dir ? t - 1 : t + 1;
        tt = tt < 0 ? 0 : (tt > a.max_t ? a.max_t : tt);
        const float* P = a.Ps + (cam * a.n_frames + tt) * 12;
        float w[3];
        ndc2world3((dir ? a.xyz_bw : a.xyz_fw) + r * 3, c, w);
        const float dep = P[8] * w[0] + P[9] * w[1] + P[10] * w[2] + P[11];
        const bool valid = dep > 0.f && (dir ? t > 0 : t < a.max_t);
        v[5 + dir] += valid ? 1.f : 0.f;
      }
    }
  }
  float tot[8];
  for (int i = 0; i < 8; ++i) tot[i] = block_sum(v[i], sh);
  const float t_neq_c = block_sum(neq_c, sh), t_csgn = block_sum(csgn, sh), t_csgn_c = block_sum(csgn_c, sh);
  float dsf = 0.f, dsb = 0.f;
  if (a.has_flow) {
    const int64_t n = a.R * a.S;
    if ((n & 3) == 0 && (((uintptr_t)a.disoccs_fw | (uintptr_t)a.disoccs_bw) & 15u) == 0) {
      const float4* f4 = reinterpret_cast<const float4*>(a.disoccs_fw);
      const float4* b4 = reinterpret_cast<const float4*>(a.disoccs_bw);
      float f[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
      for (int64_t i = threadIdx.x; i < n / 4; i += blockDim.x) {
        const float4 x = f4[i], y = b4[i];
        f[0] += x.x, f[1] += x.y, f[2] += x.z, f[3] += x.w;
        b[0] += y.x, b[1] += y.y, b[2] += y.z, b[3] += y.w;
      }
      dsf = (f[0] + f[1]) + (f[2] + f[3]), dsb = (b[0] + b[1]) + (b[2] + b[3]);
    } else {
      for (int64_t i = threadIdx.x; i < n; i += blockDim.x) dsf += a.disoccs_fw[i], dsb += a.disoccs_bw[i];
    }
  }
  const float t_dsf = block_sum(dsf, sh), t_dsb = block_sum(dsb, sh);
  if (threadIdx.x == 0) {
    s_pred = tot[0] * invR, s_gt = tot[1] * invR, s_pred_c = tot[2] * invR;
    float* st = a.stats;
    st[ST_S_PRED] = s_pred, st[ST_S_GT] = s_gt, st[ST_NEQ] = tot[7], st[ST_CSGN] = t_csgn;
    st[ST_S_PRED_C] = s_pred_c, st[ST_NEQ_C] = t_neq_c, st[ST_CSGN_C] = t_csgn_c;
    st[ST_M_DFW] = tot[3] * invR, st[ST_M_DBW] = tot[4] * invR;
    st[ST_M_DSFW] = t_dsf * invR / (float)a.S, st[ST_M_DSBW] = t_dsb * invR / (float)a.S;
    st[ST_NV_FW] = tot[5], st[ST_NV_BW] = tot[6];
  }
  __syncthreads();
  // second phase: the sums over u_r = (d_r - t)/s - (g_r - t_g)/s_g that the gradient through s and t needs
  float A = 0.f, Bs = 0.f, Ac = 0.f, Bc = 0.f;
  for (int64_t r = threadIdx.x; r < a.R; r += blockDim.x) {
    const float gn = (-a.disps[r] - t_gt) / s_gt;
    const float w = a.ray_weights ? a.ray_weights[r] : 1.f;
    const float d = a.depth_fine[r], u = (d - t_pred) / s_pred - gn;
    A += w * u, Bs += w * u * (d - t_pred);
    if (a.depth_coarse) {
      const float dc = a.depth_coarse[r], uc = (dc - t_pred_c) / s_pred_c - gn;
      Ac += w * uc, Bc += w * uc * (dc - t_pred_c);
    }
  }
  const float tA = block_sum(A, sh), tB = block_sum(Bs, sh), tAc = block_sum(Ac, sh), tBc = block_sum(Bc, sh);
  if (threadIdx.x == 0) {
    a.stats[ST_A] = tA, a.stats[ST_B] = tB, a.stats[ST_A_C] = tAc, a.stats[ST_B_C] = tBc;
    for (int k = 0; k < NSFF_LOSS_TERMS; ++k) a.term_sums[k] = 0.f;
  }
}

constexpr int kLossWarps = 8;

template <bool kBackward>
__global__ void __launch_bounds__(kLossWarps * 32) loss_kernel(const nsff_loss_args a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * kLossWarps + warp;
  const bool live = r < a.R;
  const int S = a.S;
  const float* st = a.stats;
  float term[NSFF_LOSS_TERMS];
#pragma unroll
  for (int k = 0; k < NSFF_LOSS_TERMS; ++k) term[k] = 0.f;
  if (live) {
    const float wr = a.ray_weights ? a.ray_weights[r] : 1.f;
    const float invR = 1.f / (float)a.R;
    // upstream gradient of every term's MEAN, folded with the ray weight and the 1/denominator
    float gk[NSFF_LOSS_TERMS];
#pragma unroll
    for (int k = 0; k < NSFF_LOSS_TERMS; ++k) gk[k] = kBackward ? a.g_terms[k] * wr * invR : 0.f;
    if (kBackward) {
      gk[NSFF_L_FLOW_FW] = st[ST_NV_FW] > 0.f ? a.g_terms[NSFF_L_FLOW_FW] * wr / st[ST_NV_FW] : 0.f;
      gk[NSFF_L_FLOW_BW] = st[ST_NV_BW] > 0.f ? a.g_terms[NSFF_L_FLOW_BW] * wr / st[ST_NV_BW] : 0.f;
    }
    // ---------------- per-ray terms (lane 0 owns them) ----------------
    if (lane == 0) {
      // colour (losses.py:74-78)
      float col = 0.f;
      for (int c = 0; c < 3; ++c) {
        const float e = a.rgb_fine[r * 3 + c] - a.rgbs[r * 3 + c];
        col += e * e * (1.f / 3.f);
        if (kBackward) a.d_rgb_fine[r * 3 + c] = gk[NSFF_L_COL] * 2.f * e * (1.f / 3.f);
        if (a.rgb_coarse) {
          const float ec = a.rgb_coarse[r * 3 + c] - a.rgbs[r * 3 + c];
          col += 0.1f * ec * ec * (1.f / 3.f);
          if (kBackward) a.d_rgb_coarse[r * 3 + c] = gk[NSFF_L_COL] * 0.2f * ec * (1.f / 3.f);
        }
      }
      term[NSFF_L_COL] = col;
      // shift/scale-invariant depth (losses.py:8-29, 79-83)
      const float gn = (-a.disps[r] - a.medians[1]) / st[ST_S_GT];
      {
        const float t = a.medians[0], s = st[ST_S_PRED], d = a.depth_fine[r], u = (d - t) / s - gn;
        term[NSFF_L_DISP] = a.lambda_geo_d * u * u;
        if (kBackward) {
          const float c2 = 2.f * a.lambda_geo_d * a.g_terms[NSFF_L_DISP] * invR;      // (weights are inside A, B)
          const float dLds = -c2 * st[ST_B] / (s * s);
          const float dLdt = -c2 * st[ST_A] / s + dLds * (-st[ST_CSGN] * invR);
          a.d_depth_fine[r] = c2 * wr * u / s + dLds * sgnf(d - t) * invR + (d == t ? dLdt / st[ST_NEQ] : 0.f);
        }
      }
      if (a.depth_coarse) {
        const float t = a.medians[2], s = st[ST_S_PRED_C], d = a.depth_coarse[r], u = (d - t) / s - gn;
        term[NSFF_L_DISP] += a.lambda_geo_d * u * u;
        if (kBackward) {
          const float c2 = 2.f * a.lambda_geo_d * a.g_terms[NSFF_L_DISP] * invR;
          const float dLds = -c2 * st[ST_B_C] / (s * s);
          const float dLdt = -c2 * st[ST_A_C] / s + dLds * (-st[ST_CSGN_C] * invR);
          a.d_depth_coarse[r] = c2 * wr * u / s + dLds * sgnf(d - t) * invR + (d == t ? dLdt / st[ST_NEQ_C] : 0.f);
        }
      }
    }
    if (a.has_flow) {
      const int64_t cam = a.cam_ids[r], t = a.ts[r];
      const float* K = a.Ks + cam * 9;
      const Cam cm{K[2] / K[0], K[5] / K[4]};
      if (lane == 0) {
        // 2D-3D flow consistency (losses.py:104-127) and photometric warp loss (losses.py:129-137)
        float pho = 0.f;
        for (int dir = 0; dir < 2; ++dir) {
          const float* xyz = (dir ? a.xyz_bw : a.xyz_fw) + r * 3;
          int64_t tt = dir ? t - 1 : t + 1;
          tt = tt < 0 ? 0 : (tt > a.max_t ? a.max_t : tt);
          const float* P = a.Ps + (cam * a.n_frames + tt) * 12;
          float w[3];
          ndc2world3(xyz, cm, w);
          const float ux = P[0] * w[0] + P[1] * w[1] + P[2] * w[2] + P[3];
          const float uy = P[4] * w[0] + P[5] * w[1] + P[6] * w[2] + P[7];
          const float dep = P[8] * w[0] + P[9] * w[1] + P[10] * w[2] + P[11];
          const bool valid = dep > 0.f && (dir ? t > 0 : t < a.max_t);
          const float den = fabsf(dep) + 1e-8f;
          const float* uvt = (dir ? a.uv_bw : a.uv_fw) + r * 2;
          const float ex = ux / den - uvt[0], ey = uy / den - uvt[1];
          const int kt = dir ? NSFF_L_FLOW_BW : NSFF_L_FLOW_FW;
          if (valid) term[kt] = 0.5f * a.lambda_geo_f * 0.5f * (fabsf(ex) + fabsf(ey));
          if (kBackward) {
            float gw[3] = {0.f, 0.f, 0.f}, gx[3];
            if (valid) {
              const float c0 = gk[kt] * 0.5f * a.lambda_geo_f * 0.5f;
              const float gux = c0 * sgnf(ex) / den, guy = c0 * sgnf(ey) / den;
              const float gdep = -(c0 * sgnf(ex) * ux + c0 * sgnf(ey) * uy) / (den * den) * sgnf(dep);
              for (int c = 0; c < 3; ++c) gw[c] = gux * P[c] + guy * P[4 + c] + gdep * P[8 + c];
            }
            ndc2world_pullback(xyz, cm, gw, gx);
            float* dx = (dir ? a.d_xyz_bw : a.d_xyz_fw) + r * 3;
            for (int c = 0; c < 3; ++c) dx[c] = gx[c];
          }
          const float doc = (dir ? a.disocc_bw : a.disocc_fw)[r], mean_doc = st[dir ? ST_M_DBW : ST_M_DFW];
          const float* rw = (dir ? a.rgb_bw : a.rgb_fw) + r * 3;
          for (int c = 0; c < 3; ++c) {
            const float e = rw[c] - a.rgbs[r * 3 + c];
            pho += doc * e * e / mean_doc * (1.f / 3.f);
            if (kBackward) (dir ? a.d_rgb_bw : a.d_rgb_fw)[r * 3 + c] = gk[NSFF_L_PHO] * doc * 2.f * e / mean_doc * (1.f / 3.f);
          }
        }
        term[NSFF_L_PHO] = pho;
      }
      // ---------------- per-sample terms ----------------
      const int N95 = a.n_reg;                                    // int(N * z_far), losses.py:147-150
      const int k = a.thickness > 1 ? a.thickness : 1, padl = (k - 1) / 2;
      const float inv_s3 = 1.f / (3.f * (float)S);
      const float inv_n3 = N95 > 0 ? 1.f / (3.f * (float)N95) : 0.f;
      const float inv_m3 = N95 > 1 ? 1.f / (3.f * (float)(N95 - 1)) : 0.f;
      const float m_dsf = st[ST_M_DSFW], m_dsb = st[ST_M_DSBW];
      float ent = 0.f, cross = 0.f, cyc = 0.f, rtemp = 0.f, rmin = 0.f, rsp = 0.f;
      for (int s = lane; s < S; s += 32) {
        const int64_t i = r * S + s;
        const float wt = a.tw[i], ws = a.sw[i];
        ent += -wt * logf(wt + 1e-8f);                              // losses.py:86-88
        float dil = 0.f;                                            // zero-padded box filter of the detached weights
        for (int j = 0; j < k; ++j) {
          const int q = s - padl + j;
          if (q >= 0 && q < S) dil += a.tw[r * S + q];
        }
        cross += dil * logf(ws + 1e-8f);                            // losses.py:91-97
        if (kBackward) {
          a.d_tw[i] = gk[NSFF_L_ENT] * a.lambda_ent * -(logf(wt + 1e-8f) + wt / (wt + 1e-8f));
          a.d_sw[i] = gk[NSFF_L_CROSS] * a.cross_entropy_w * dil / (ws + 1e-8f);
        }
        // cycle consistency (losses.py:139-145)
        const float dsf = a.disoccs_fw[i], dsb = a.disoccs_bw[i];
        const float* x0 = a.xyzs_fine + i * 3;
        for (int c = 0; c < 3; ++c) {
          const float ef = a.xyzs_fw_bw[i * 3 + c] - x0[c], eb = a.xyzs_bw_fw[i * 3 + c] - x0[c];
          cyc += (dsf * fabsf(ef) / m_dsf + dsb * fabsf(eb) / m_dsb) * inv_s3;
          if (kBackward) {
            a.d_xyzs_fw_bw[i * 3 + c] = gk[NSFF_L_CYC] * dsf * sgnf(ef) / m_dsf * inv_s3;
            a.d_xyzs_bw_fw[i * 3 + c] = gk[NSFF_L_CYC] * dsb * sgnf(eb) / m_dsb * inv_s3;
          }
        }
        // scene-flow regularisers in world space on the first int(0.95 N) samples (losses.py:147-161)
        float gwf[3] = {0.f, 0.f, 0.f}, gwb[3] = {0.f, 0.f, 0.f};
        if (s < N95) {
          float xw[3], xf[3], xb[3];
          ndc2world3(x0, cm, xw);
          ndc2world3(a.xyzs_fw + i * 3, cm, xf);
          ndc2world3(a.xyzs_bw + i * 3, cm, xb);
          for (int c = 0; c < 3; ++c) {
            const float tsm = xf[c] + xb[c] - 2.f * xw[c], mf = xf[c] - xw[c], mb = xb[c] - xw[c];
            rtemp += a.lambda_reg * fabsf(tsm) * inv_n3;
            rmin += a.lambda_reg * (fabsf(mf) + fabsf(mb)) * inv_n3;
            if (kBackward) {
              const float gt = gk[NSFF_L_REG_TEMP] * a.lambda_reg * sgnf(tsm) * inv_n3;
              gwf[c] += gt + gk[NSFF_L_REG_MIN] * a.lambda_reg * sgnf(mf) * inv_n3;
              gwb[c] += gt + gk[NSFF_L_REG_MIN] * a.lambda_reg * sgnf(mb) * inv_n3;
            }
          }
          // spatial smoothness: pairs (s, s+1) and, for the gradient, (s-1, s)
          for (int side = 0; side < 2; ++side) {
            const int s0 = side ? s - 1 : s;                         // the pair is (s0, s0 + 1)
            if (s0 < 0 || s0 + 1 >= N95) continue;
            if (side == 1 && !kBackward) continue;
            const int64_t i0 = r * S + s0, i1 = i0 + 1;
            float w0[3], w1[3], f0[3], f1[3], b0[3], b1[3];
            ndc2world3(a.xyzs_fine + i0 * 3, cm, w0), ndc2world3(a.xyzs_fine + i1 * 3, cm, w1);
            ndc2world3(a.xyzs_fw + i0 * 3, cm, f0), ndc2world3(a.xyzs_fw + i1 * 3, cm, f1);
            ndc2world3(a.xyzs_bw + i0 * 3, cm, b0), ndc2world3(a.xyzs_bw + i1 * 3, cm, b1);
            const float dx = w1[0] - w0[0], dy = w1[1] - w0[1], dz = w1[2] - w0[2];
            const float spw = expf(-2.f * sqrtf(dx * dx + dy * dy + dz * dz));
            for (int c = 0; c < 3; ++c) {
              const float df = (f1[c] - w1[c]) - (f0[c] - w0[c]), db = (b1[c] - w1[c]) - (b0[c] - w0[c]);
              if (side == 0) rsp += a.lambda_reg * (fabsf(df) + fabsf(db)) * spw * inv_m3;
              if (kBackward) {
                const float sg = side ? 1.f : -1.f;                  // this sample is s0+1 (side 1) or s0 (side 0)
                gwf[c] += gk[NSFF_L_REG_SP] * a.lambda_reg * spw * inv_m3 * sgnf(df) * sg;
                gwb[c] += gk[NSFF_L_REG_SP] * a.lambda_reg * spw * inv_m3 * sgnf(db) * sg;
              }
            }
          }
        }
        if (kBackward) {
          float gx[3];
          ndc2world_pullback(a.xyzs_fw + i * 3, cm, gwf, gx);
          for (int c = 0; c < 3; ++c) a.d_xyzs_fw[i * 3 + c] = s < N95 ? gx[c] : 0.f;
          ndc2world_pullback(a.xyzs_bw + i * 3, cm, gwb, gx);
          for (int c = 0; c < 3; ++c) a.d_xyzs_bw[i * 3 + c] = s < N95 ? gx[c] : 0.f;
        }
      }
      float vals[6] = {ent, cross, cyc, rtemp, rmin, rsp};
      for (int j = 0; j < 6; ++j)
        for (int o = 16; o > 0; o >>= 1) vals[j] += __shfl_xor_sync(0xffffffffu, vals[j], o);
      if (lane == 0) {
        term[NSFF_L_ENT] = a.lambda_ent * vals[0];
        term[NSFF_L_CROSS] = a.cross_entropy_w * vals[1];
        term[NSFF_L_CYC] = vals[2];
        term[NSFF_L_REG_TEMP] = vals[3], term[NSFF_L_REG_MIN] = vals[4], term[NSFF_L_REG_SP] = vals[5];
      }
    }
    if (lane == 0)
      for (int k2 = 0; k2 < NSFF_LOSS_TERMS; ++k2) term[k2] *= wr;
  }
  if (!kBackward) {
    __shared__ float sh[NSFF_LOSS_TERMS][kLossWarps];
    if (lane == 0)
      for (int k2 = 0; k2 < NSFF_LOSS_TERMS; ++k2) sh[k2][warp] = live ? term[k2] : 0.f;
    __syncthreads();
    if (threadIdx.x < NSFF_LOSS_TERMS) {
      float t = 0.f;
      for (int w2 = 0; w2 < kLossWarps; ++w2) t += sh[threadIdx.x][w2];
      atomicAdd(a.term_sums + threadIdx.x, t);
    }
  }
}

__global__ void loss_finish_kernel(const nsff_loss_args a) {
  const int k = threadIdx.x;
  if (k >= NSFF_LOSS_TERMS) return;
  float den = (float)a.R;
  if (k == NSFF_L_FLOW_FW) den = a.stats[ST_NV_FW];
  if (k == NSFF_L_FLOW_BW) den = a.stats[ST_NV_BW];
  a.terms[k] = den > 0.f ? a.term_sums[k] / den : 0.f;
}

static int check_args(const nsff_loss_args* a, bool backward) {
  if (!a || a->R <= 0 || a->S <= 0) return NSFF_EINVAL;
  if (!a->rgb_fine || !a->depth_fine || !a->rgbs || !a->disps || !a->medians || !a->stats || !a->term_sums || !a->terms)
    return NSFF_EINVAL;
  if ((a->rgb_coarse == nullptr) != (a->depth_coarse == nullptr)) return NSFF_EINVAL;
  if (a->has_flow) {
    const void* need[] = {a->tw, a->sw, a->xyz_fw, a->xyz_bw, a->disocc_fw, a->disocc_bw, a->rgb_fw, a->rgb_bw,
                          a->disoccs_fw, a->disoccs_bw, a->xyzs_fw_bw, a->xyzs_bw_fw, a->xyzs_fine, a->xyzs_fw,
                          a->xyzs_bw, a->uv_fw, a->uv_bw, a->cam_ids, a->ts, a->Ks, a->Ps};
    for (const void* p : need)
      if (!p) return NSFF_EINVAL;
    if (a->n_frames <= 0 || a->max_t < 0 || a->max_t >= a->n_frames) return NSFF_EINVAL;
  }
  if (backward) {
    if (!a->g_terms || !a->d_rgb_fine || !a->d_depth_fine) return NSFF_EINVAL;
    if (a->rgb_coarse && (!a->d_rgb_coarse || !a->d_depth_coarse)) return NSFF_EINVAL;
    if (a->has_flow) {
      const void* need[] = {a->d_tw, a->d_sw, a->d_xyz_fw, a->d_xyz_bw, a->d_rgb_fw, a->d_rgb_bw, a->d_xyzs_fw_bw,
                            a->d_xyzs_bw_fw, a->d_xyzs_fw, a->d_xyzs_bw};
      for (const void* p : need)
        if (!p) return NSFF_EINVAL;
    }
  }
  return NSFF_OK;
}

}  // namespace nsff

using namespace nsff;

extern "C" int nsff_loss_stats_floats(void) { return ST_COUNT; }

extern "C" int nsff_nerfw_loss_forward(const nsff_loss_args* a, void* stream) {
  int rc = check_args(a, false);
  if (rc != NSFF_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  loss_stats_kernel<<<1, 1024, 0, st>>>(*a);
  loss_kernel<false><<<(unsigned)cdiv64(a->R, kLossWarps), kLossWarps * 32, 0, st>>>(*a);
  loss_finish_kernel<<<1, 32, 0, st>>>(*a);
  return after_launch(3);
}

extern "C" int nsff_nerfw_loss_backward(const nsff_loss_args* a, void* stream) {
  int rc = check_args(a, true);
  if (rc != NSFF_OK) return rc;
  loss_kernel<true><<<(unsigned)cdiv64(a->R, kLossWarps), kLossWarps * 32, 0, (cudaStream_t)stream>>>(*a);
  return after_launch();
}
