// a13: hierarchical resampling - inverse-CDF sampling of the coarse weights and the merge of the
// coarse and the resampled depths (models/rendering.py:10-49, 335-348).  One warp per ray.
#include "nsff_common.cuh"

namespace nsff {

constexpr int kMaxBins = 512;    // S-1 bins per ray held in shared memory (per warp)
constexpr int kWarpsPerBlock = 4;

// bins: (R, nb+1) with row stride bins_rs (0 = one shared row); weights: (R, nb) row stride w_rs.
// u: (R, N) or NULL for the deterministic linspace(0,1,N).   out: (R, N)
__global__ void __launch_bounds__(32 * kWarpsPerBlock)
    sample_pdf_kernel(const float* __restrict__ bins, int64_t bins_rs, const float* __restrict__ weights,
                      int64_t w_rs, int64_t R, int nb, int N, const float* __restrict__ u, float eps,
                      float* __restrict__ out) {
  __shared__ float s_cdf[kWarpsPerBlock][kMaxBins + 1];
  __shared__ float s_bins[kWarpsPerBlock][kMaxBins + 1];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t ray = (int64_t)blockIdx.x * kWarpsPerBlock + wib;
  if (ray >= R) return;
  float* cdf = s_cdf[wib];
  float* bn = s_bins[wib];
  const float* w = weights + ray * w_rs;
  const float* b = bins + ray * bins_rs;
  // sum of (w + eps)
  float part = 0.f;
  for (int i = lane; i < nb; i += 32) part += w[i] + eps;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
  const float total = part;
  // cdf[0] = 0, cdf[i+1] = cumsum(pdf)[i]   (rendering.py:24-26)
  float carry = 0.f;
  for (int i0 = 0; i0 < nb; i0 += 32) {
    int i = i0 + lane;
    float v = (i < nb) ? (w[i] + eps) / total : 0.f;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      float o = __shfl_up_sync(0xffffffffu, v, d);
      if (lane >= d) v += o;
    }
    if (i < nb) cdf[i + 1] = carry + v;
    carry += __shfl_sync(0xffffffffu, v, 31);
  }
  if (lane == 0) cdf[0] = 0.f;
  for (int i = lane; i <= nb; i += 32) bn[i] = b[i];
  __syncwarp();
  for (int j = lane; j < N; j += 32) {
    float uj;
    if (u) uj = u[ray * N + j];
    else if (N == 1) uj = 0.f;
    else {                                  // torch.linspace(0,1,N), symmetric evaluation
      float step = 1.f / (float)(N - 1);
      uj = (j < N / 2) ? __fmul_rn((float)j, step) : __fsub_rn(1.f, __fmul_rn((float)(N - 1 - j), step));
    }
    // searchsorted(cdf, u, right=True): first index with cdf[idx] > u, over nb+1 entries
    int lo = 0, hi = nb + 1;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (cdf[mid] > uj) hi = mid; else lo = mid + 1;
    }
    int below = max(lo - 1, 0), above = min(lo, nb);
    float c0 = cdf[below], c1 = cdf[above];
    float den = c1 - c0;
    if (den < eps) den = 1.f;
    out[ray * N + j] = bn[below] + (uj - c0) / den * (bn[above] - bn[below]);
  }
}

// out (R, total) = ascending sort of the concatenation of up to three rows.  Rank sort: every element
// counts how many precede it (ties broken by position, i.e. stable), O(n^2) but n <= ~512 and the
// values sit in shared memory; produces exactly the multiset order torch.sort returns.
__global__ void __launch_bounds__(128) merge_sorted_kernel(const float* __restrict__ a, int na,
                                                           const float* __restrict__ b, int nb,
                                                           const float* __restrict__ c, int nc, int64_t R,
                                                           float* __restrict__ out) {
  extern __shared__ float vals[];
  const int64_t ray = blockIdx.x;
  const int n = na + nb + nc;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float v;
    if (i < na) v = a[ray * na + i];
    else if (i < na + nb) v = b[ray * nb + (i - na)];
    else v = c[ray * nc + (i - na - nb)];
    vals[i] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = vals[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const float o = vals[j];
      rank += (o < v) || (o == v && j < i);
    }
    out[ray * n + rank] = v;
  }
}

}  // namespace nsff

using namespace nsff;

extern "C" int nsff_sample_pdf(const float* bins, int64_t bins_row_stride, const float* weights,
                               int64_t weights_row_stride, int64_t R, int n_weights, int n_importance,
                               const float* u, float eps, float* out, void* stream) {
  if (R < 0 || n_weights < 1 || n_importance < 1) return NSFF_EINVAL;
  if (R == 0) return NSFF_OK;
  if (!bins || !weights || !out) return NSFF_EINVAL;
  if (n_weights > kMaxBins) return NSFF_EUNSUPPORTED;
  sample_pdf_kernel<<<(unsigned)cdiv64(R, kWarpsPerBlock), 32 * kWarpsPerBlock, 0, (cudaStream_t)stream>>>(
      bins, bins_row_stride, weights, weights_row_stride, R, n_weights, n_importance, u, eps, out);
  return after_launch();
}

extern "C" int nsff_merge_sorted(const float* a, int na, const float* b, int nb, const float* c, int nc, int64_t R,
                                 float* out, void* stream) {
  if (R < 0 || na < 0 || nb < 0 || nc < 0) return NSFF_EINVAL;
  const int n = na + nb + nc;
  if (R == 0 || n == 0) return NSFF_OK;
  if (!out || (na && !a) || (nb && !b) || (nc && !c)) return NSFF_EINVAL;
  if (n > 8192) return NSFF_EUNSUPPORTED;
  merge_sorted_kernel<<<(unsigned)R, 128, n * sizeof(float), (cudaStream_t)stream>>>(a, na, b, nb, c, nc, R, out);
  return after_launch();
}
