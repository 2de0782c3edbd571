/* nsff_b200.h - C ABI of the Blackwell-native NSFF volume-rendering hot path.
 *
 * Drop-in boundary for kwea123/nsff_pl's render_rays() path.  The reference
 * is pure Python/PyTorch, so "the reference's FFI for this path" is the precedent it sets with
 * its only native op: models/softsplat.py:208-301 (a CUDA-only autograd.Function over raw
 * kernels, contiguous fp32 tensors, NotImplementedError on CPU tensors) and
 * third_party/flow/alt_cuda_corr/correlation.cpp:18-54 (CHECK_CUDA / CHECK_CONTIGUOUS, raw
 * pointers handed to launchers).  Each entry point below names the reference lines it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous row-major data unless marked "host";
 *   - floats are fp32, time indices are int64 (torch.long), exactly what the reference passes;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing syncs;
 *   - no allocation happens inside the library: callers own every buffer;
 *   - return value: 0 on success, a negative NSFF_E* code otherwise (nsff_error_string()).
 *   - no global state; the only long-lived object is the caller-owned "field pack" buffer.
 *
 * Shapes: R rays, S samples per ray, B = R*S points, points ordered (ray, sample).
 */
#ifndef NSFF_B200_H_
#define NSFF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NSFF_ABI_VERSION 1

#if defined(__GNUC__)
#define NSFF_API __attribute__((visibility("default")))
#else
#define NSFF_API
#endif

enum {
  NSFF_OK = 0,
  NSFF_EINVAL = -1,      /* bad shape / null pointer / misaligned buffer           */
  NSFF_EUNSUPPORTED = -2,/* configuration this build has no kernel for             */
  NSFF_ECUDA = -3,       /* a CUDA runtime call failed (see nsff_last_cuda_error)  */
  NSFF_EARCH = -4        /* device is not sm_100 (tcgen05 path requested)          */
};

NSFF_API int nsff_abi_version(void);
NSFF_API const char* nsff_error_string(int code);
/* cudaGetErrorString of the last failing runtime call made by this library on this thread. */
NSFF_API const char* nsff_last_cuda_error(void);
/* Select the device subsequent launches target (the library carries its own CUDA runtime). */
NSFF_API int nsff_set_device(int device);
/* Number of kernels this library has launched since load (bench.py's gpu_launches counter). */
NSFF_API uint64_t nsff_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * a1  PosEmbedding.forward                                            models/nerf.py:17-30
 *     out[n] = [x, sin(2^0 x), cos(2^0 x), ..., sin(2^(F-1) x), cos(2^(F-1) x)]   (n, 3+6F)
 *     freqs: F device floats (the module's .freqs, so logscale=False also works).
 */
NSFF_API int nsff_pos_embedding(const float* x, int64_t n, const float* freqs, int n_freqs,
                       float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a6  stratified depths + sample points                      models/rendering.py:314-324,332,359
 *     zs = linspace(0,1,S) (optionally jittered: lower+(upper-lower)*perturb*u, u = `rand` (R,S));
 *     xyz = o + d*z.  `rand` may be NULL iff perturb == 0.
 */
NSFF_API int nsff_sample_points(const float* rays /*(R,6)*/, int64_t R, int S, float perturb,
                       const float* rand /*(R,S) or NULL*/, float* zs /*(R,S)*/,
                       float* xyz /*(R,S,3)*/, void* stream);
/*     xyz = o + d*z for caller-provided depths (the fine pass of the hierarchical mode, :359). */
NSFF_API int nsff_points_from_depths(const float* rays, const float* zs, int64_t R, int S,
                            float* xyz, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a3/a4/a5  NeRF.forward - static and dynamic 8x256 MLPs               models/nerf.py:118-213
 *
 * A "field pack" is the caller-owned device image of one NeRF module's parameters, laid out for
 * the tcgen05 kernel (bf16, 128B-swizzled K-major tiles, one block per 64-wide K chunk) plus an
 * fp32 bias table.  nsff_field_pack_bytes() sizes it, nsff_field_pack() fills it from the fp32
 * nn.Parameters (state_dict names listed in SURVEY.md §8b); repack after every optimizer step.
 */
typedef struct {
  int32_t W;               /* hidden width; the tcgen05 path requires 256                  */
  int32_t D;               /* trunk depth; 8                                               */
  int32_t skip;            /* index of the layer that re-reads the input; 4               */
  int32_t in_xyz;          /* 63                                                           */
  int32_t in_dir;          /* 27 (0 when use_viewdir is false)                             */
  int32_t in_a;            /* 0 or 48                                                      */
  int32_t in_t;            /* 0 or 48 (0 when the module has no transient branch)          */
  int32_t use_viewdir;
  int32_t has_transient;
  int32_t has_flow;        /* fine model with output_flow                                  */
  float   flow_scale;      /* 0.2                                                          */
} nsff_field_desc;

/* fp32 parameter pointers, reference names.  weight (out,in) row-major, bias (out).
 * Unused entries (e.g. transient_* of a static-only module) are NULL. */
typedef struct {
  const float* static_w[8];   const float* static_b[8];    /* static_xyz_encoding_{1..8}.0        */
  const float* static_final_w; const float* static_final_b;/* static_xyz_encoding_final           */
  const float* static_dir_w;  const float* static_dir_b;   /* static_dir_encoding.0               */
  const float* static_sigma_w; const float* static_sigma_b;/* static_sigma                        */
  const float* static_rgb_w;  const float* static_rgb_b;   /* static_rgb.0                        */
  const float* trans_w[8];    const float* trans_b[8];     /* transient_xyz_encoding_{1..8}.0     */
  const float* trans_final_w; const float* trans_final_b;  /* transient_xyz_encoding_final        */
  const float* trans_sigma_w; const float* trans_sigma_b;  /* transient_sigma                     */
  const float* trans_rgb_w;   const float* trans_rgb_b;    /* transient_rgb.0                     */
  const float* trans_fw_w;    const float* trans_fw_b;     /* transient_flow_fw.0                 */
  const float* trans_bw_w;    const float* trans_bw_b;     /* transient_flow_bw.0                 */
} nsff_field_params;

NSFF_API int nsff_field_pack_bytes(const nsff_field_desc* desc /*host*/, size_t* bytes /*host*/);
NSFF_API int nsff_field_pack(const nsff_field_desc* desc /*host*/, const nsff_field_params* params /*host*/,
                    void* pack /*device, 128B aligned*/, size_t pack_bytes, void* stream);

/* What one forward evaluates; mirrors the four call modes of nerf.py:129-141. */
enum {
  NSFF_OUT_STATIC      = 1,   /* static rgb(3) + raw sigma(1)                               */
  NSFF_OUT_TRANSIENT   = 2,   /* transient rgb(3) + raw sigma(1)                            */
  NSFF_OUT_FLOW_FW     = 4,   /* + flow_scale*tanh fw(3)                                    */
  NSFF_OUT_FLOW_BW     = 8,   /* + flow_scale*tanh bw(3)                                    */
  NSFF_OUT_SIGMA_ONLY  = 16,  /* coarse@test: raw sigmas only (static [, transient])        */
  NSFF_OUT_EVERY_LAYER = 32   /* evaluate xyz_encoding_final as its own layer (training forward: its output is an
                                 activation the backward differentiates).  Without it the tensor-core path folds
                                 that activation-free layer into the one after it (same function, one bf16
                                 rounding less); the fp32 path ignores the flag. */
};

/* Per-point inputs for the fused renderer entry (PE computed in-kernel; rendering.py:164-175,
 * 113-120).  Point p belongs to ray p / S. */
typedef struct {
  const float*   xyz;        /* (B,3) sample points                                         */
  const float*   offset;     /* (B,*) optional: points are xyz + offset[.., 0:3]; NULL = none */
  int32_t        offset_ld;  /* row stride (floats) of `offset`                             */
  float*         xyz_out;    /* (B,3) optional: where xyz+offset is written (xyzs_fw/bw)    */
  const float*   dirs;       /* (R,3) view directions (rays_d or kwargs['view_dir'])        */
  const float*   t_table;    /* (N_frames or R, in_t) time codes; NULL if no transient      */
  const int64_t* t_index;    /* (R,) frame index per ray; NULL => row r of t_table is ray r */
  int32_t        t_shift;    /* -1/0/+1: index = clamp(t_index + shift, 0, max_t)           */
  int32_t        max_t;
  const float*   a_table;    /* appearance codes, same addressing (a_index, no shift)       */
  const int64_t* a_index;
  const float*   zs;         /* (B,) depths; flow heads are zeroed where z > 0.95 (:126,187) */
  int64_t        B;
  int32_t        S;
  float*         xyz_fw_out; /* (B,3) optional, passes that evaluate the flow heads: xyz + flow_fw and              */
  float*         xyz_bw_out; /*       xyz + flow_bw = the points of the warped passes (rendering.py:219,225), so   */
                             /*       those passes read 12 B/pt instead of 12 B out of every 56-byte output row    */
} nsff_points;

/* Fused PE + MLP forward on tensor cores (tcgen05/TMEM, bf16 operands, fp32 accumulate).
 * out: (B, C) fp32, C = 4*[static] + 4*[transient] + 3*[fw] + 3*[bw]  (or 1..2 in SIGMA_ONLY),
 * column order as nerf.py:129-141. */
NSFF_API int nsff_field_forward(const nsff_field_desc* desc /*host*/, const void* pack,
                       const nsff_points* pts /*host*/, int out_mask, float* out, int32_t out_ld,
                       void* stream);

/* NeRF.forward drop-in on an already-embedded input x (B, in_xyz+in_dir+in_a+in_t) fp32
 * (column split of nerf.py:142-160).  Same kernel, the prologue converts x instead of encoding. */
NSFF_API int nsff_field_forward_embedded(const nsff_field_desc* desc /*host*/, const void* pack,
                                const float* x, int32_t x_ld, int64_t B, int out_mask,
                                float* out, int32_t out_ld, void* stream);

/* fp32 verification path: the same math on CUDA cores straight from the fp32 parameters
 * (no pack, no tensor cores); `scratch` must hold nsff_field_fp32_scratch_floats(...) floats. */
NSFF_API size_t nsff_field_fp32_scratch_floats(const nsff_field_desc* desc /*host*/, int64_t B);
NSFF_API int nsff_field_forward_fp32(const nsff_field_desc* desc /*host*/, const nsff_field_params* params /*host*/,
                            const nsff_points* pts /*host, or NULL*/,
                            const float* x_embedded /*(B_embedded, x_ld) or NULL*/, int32_t x_ld,
                            int64_t B_embedded, int out_mask, float* out, int32_t out_ld, float* scratch,
                            void* stream);

/* Backward of NeRF.forward (fp32, CUDA cores; the reference relies on torch autograd through nn.Linear):
 * recomputes the forward on `pts` storing every layer, then back-propagates d_out (B, C) through the heads and
 * trunks.  Parameter gradients are ACCUMULATED (+=) into `grads` (same layout as nsff_field_params, fp32, caller
 * zeroes them); d_pos (B,3) (gradient w.r.t. the encoded positions: used by the flow-warped passes), d_t (R,in_t)
 * and d_a (R,in_a) (gradients w.r.t. the per-ray codes) are overwritten when non-NULL.  `out` is the forward
 * output (needed for the sigmoid / tanh derivatives).  pts->t_index / a_index must be NULL (per-ray code rows). */
typedef struct {
  float* static_w[8];   float* static_b[8];
  float* static_final_w; float* static_final_b;
  float* static_dir_w;  float* static_dir_b;
  float* static_sigma_w; float* static_sigma_b;
  float* static_rgb_w;  float* static_rgb_b;
  float* trans_w[8];    float* trans_b[8];
  float* trans_final_w; float* trans_final_b;
  float* trans_sigma_w; float* trans_sigma_b;
  float* trans_rgb_w;   float* trans_rgb_b;
  float* trans_fw_w;    float* trans_fw_b;
  float* trans_bw_w;    float* trans_bw_b;
} nsff_field_grads;

NSFF_API size_t nsff_field_bwd_scratch_floats(const nsff_field_desc* desc /*host*/, int64_t B);
NSFF_API int nsff_field_backward_fp32(const nsff_field_desc* desc /*host*/, const nsff_field_params* params /*host*/,
                                      const nsff_points* pts /*host*/, int out_mask, const float* out,
                                      const float* d_out, int32_t out_ld, const nsff_field_grads* grads /*host*/,
                                      float* d_pos, float* d_t, float* d_a, float* scratch, void* stream);

/* ---------------------------------------------------------------------------------------------
 * b   tensor-core backward of the field           torch autograd of models/nerf.py:118-213 (train.py:178-198)
 *     Same contract as nsff_field_backward_fp32 (gradients ACCUMULATE into `grads`; d_pos / d_t / d_a are overwritten),
 *     but every GEMM (forward recompute, dgrad, wgrad) runs as bf16 tcgen05 MMAs with fp32 accumulation over
 *     activations kept in HBM as 128-byte-swizzled "tile images" ([128-point tile][64-column chunk][128 x 128 B]).
 *     `pack` = nsff_field_bwd_tc_pack of the CURRENT parameters (W and W^T blocks); `params` supplies the fp32 biases.
 */
NSFF_API size_t nsff_field_bwd_tc_pack_bytes(const nsff_field_desc* desc /*host*/);
NSFF_API int nsff_field_bwd_tc_pack(const nsff_field_desc* desc /*host*/, const nsff_field_params* params /*host*/,
                                    void* pack, void* stream);
/*     `saved` (optional) = the buffer nsff_field_forward_save filled for the same (desc, B, out_mask): the hidden
 *     activations of every layer as tile images + ReLU bit planes, written by the forward kernel itself.  With it
 *     nothing is recomputed AND the dgrad GEMMs of all layers run as ONE chained launch per call (dZ of a layer stays in
 *     tensor memory as the A operand of the next), followed by the weight-gradient GEMMs eight to a launch;
 *     `scratch` then holds the dZ image of every layer (nsff_field_bwd_tc_scratch_bytes(.., with_saved = 1)).
 *     Without it (NULL), or with an appearance code, or with NSFF_BWD_CHAIN=0 in the environment, the layers are
 *     re-evaluated / differentiated one GEMM launch at a time.
 */
NSFF_API size_t nsff_field_save_bytes(const nsff_field_desc* desc /*host*/, int64_t B, int out_mask);
NSFF_API int nsff_field_forward_save(const nsff_field_desc* desc /*host*/, const void* pack,
                                     const nsff_points* pts /*host*/, int out_mask, float* out, int32_t out_ld,
                                     void* save /*device, 16 B aligned*/, void* stream);
NSFF_API size_t nsff_field_bwd_tc_scratch_bytes(const nsff_field_desc* desc /*host*/, int64_t B, int with_saved);
NSFF_API int nsff_field_backward_tc(const nsff_field_desc* desc /*host*/, const nsff_field_params* params /*host*/,
                                    const void* pack, const nsff_points* pts /*host*/, int out_mask, const float* out,
                                    const float* d_out, int32_t out_ld, const nsff_field_grads* grads /*host*/,
                                    float* d_pos, float* d_t, float* d_a, const void* saved, void* scratch,
                                    void* stream);

/* The two GEMM kernels of that backward, exposed for hosts that keep their own tile images (and for the unit tests).
 * Image = [ceil(n_rows/128) tiles][chunks][128 rows x 64 bf16, 128B swizzle]; weight block = [N rows x 64 bf16], same
 * swizzle (nsff_img_pack_weight: blk[n][k] = transpose ? src[(k0+k)*ld + n0+n] : src[(n0+n)*ld + k0+k], zero padded). */
NSFF_API size_t nsff_img_bytes(int64_t n_rows, int n_chunks);
NSFF_API int nsff_img_from_rows(const float* src, int32_t ld, int32_t col0, int32_t ncols, int64_t n_rows, void* img,
                                int32_t img_chunks, int32_t chunk, void* stream);
NSFF_API int nsff_img_to_rows(const void* img, int32_t img_chunks, int32_t chunk, float* dst, int32_t ld, int32_t col0,
                              int32_t ncols, int64_t n_rows, void* stream);
NSFF_API int nsff_img_pack_weight(const float* src, int32_t ld, int32_t transpose, int32_t n0, int32_t k0,
                                  int32_t n_valid, int32_t k_valid, int32_t N, void* block, void* stream);
typedef struct {
  const void* a_img; int32_t a_chunks;   /* A: (n_rows, 64*a_chunks), every chunk is a K chunk                    */
  const void* b_blocks; int32_t N;       /* a_chunks weight blocks of N x 64 (N = 64, 128 or 256)                 */
  const float* bias;                     /* N floats or NULL                                                      */
  int32_t epi;                           /* 0 relu->image, 1 linear->image, 2 mask(image)->image, 3 fp32 rows,
                                            4 mask(bit plane)->image                                              */
  void* out_img; const void* mask_img;   /* N/64 chunks per tile; epi 2: out = acc where mask_img > 0             */
  float* out_rows; int32_t ld, ncols, beta; /* epi 3: out_rows[r*ld + c] (+)= acc, c < ncols                       */
  int64_t n_rows;
  uint32_t* relu_bits;                   /* optional (image epilogues): bit plane of (out > 0), per tile
                                            [8 words][128 rows] uint32, bit i of word w = column 32w+i            */
  const uint32_t* mask_bits;             /* epi 4: out = acc where the bit is set                                 */
  float* colsum;                         /* optional (image epilogues): colsum[c] += sum_r out[r, c] (fp32)       */
} nsff_img_gemm_args;
NSFF_API int nsff_img_gemm(const nsff_img_gemm_args* args /*host*/, void* stream);
typedef struct {
  const void* a_img;                     /* (n_rows, 256): 4 chunks                                               */
  const void* b_img; int32_t b_chunks;   /* (n_rows, 64*b_chunks), 1..4 chunks                                    */
  int64_t n_rows;
  float* acc_t;                          /* (64*b_chunks, 256) fp32: acc_t[n*256 + m] += sum_r B[r,n] A[r,m]      */
} nsff_img_wgrad_args;
NSFF_API int nsff_img_wgrad(const nsff_img_wgrad_args* args /*host*/, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a8  eval-only camera visibility                rendering.py:190-200, ray_utils.py:127-181
 *     vis[p] = number of cameras whose image contains the world-space projection of NDC point p.
 *     w2c: (n_cams, 12) row-major 3x4 world-to-camera matrices (inverse of the c2w poses).
 */
NSFF_API int nsff_visibility(const float* xyz /*(B,3)*/, int64_t B, const float* K /*host: fx,fy,cx,cy*/,
                    const float* w2c /*device (n_cams,12)*/, int n_cams, int H, int W,
                    float* vis /*(B,)*/, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a9-a12  density->alpha, joint static+dynamic composite, warped composites, disocclusion
 *         models/rendering.py:202-298 and 98-140.  NULL output pointers are skipped.
 */
typedef struct {
  int64_t R; int32_t S;
  /* inputs */
  const float* zs;          /* (R,S)                                                       */
  const float* xyz;         /* (R,S,3)  (needed only for xyz_fine)                          */
  const float* raw;         /* (R,S,raw_ld) output of nsff_field_forward                    */
  int32_t raw_ld;
  int32_t has_static_rgb;   /* 0 in SIGMA_ONLY mode: raw = [sigma_s(, sigma_t)]             */
  int32_t has_transient;
  int32_t has_flow;         /* raw holds fw,bw flows at columns 8..13                       */
  const float* noise_static;/* (R,S) N(0,1) draws or NULL; multiplied by noise_std          */
  const float* noise_trans;
  float noise_std;
  const float* vis;         /* (R,S) camera-visibility counts or NULL (sigma_t := -10 at 0) */
  /* warped passes (training only): raw_* = (R,S,7) [rgb, sigma, flow-back] at xyz+flow     */
  const float* raw_fw; const float* raw_bw; int32_t raw_w_ld;
  const float* xyz_fw; const float* xyz_bw;      /* (R,S,3) warped points                   */
  const float* noise_fw; const float* noise_bw;  /* (R,S) or NULL                           */
  /* per-sample outputs (R,S[,3]) */
  float* static_sigmas; float* transient_sigmas;
  float* static_weights; float* transient_weights; float* weights;
  float* static_alphas; float* transient_alphas;
  float* xyzs_fw_bw; float* xyzs_bw_fw;          /* (R,S,3)                                 */
  float* disoccs_fw; float* disoccs_bw;          /* (R,S)                                   */
  /* per-ray outputs */
  float* depth; float* rgb; float* transient_alpha; float* transient_rgb;
  float* static_only_rgb; float* static_only_depth;
  float* xyz_fine; float* flow_fw; float* xyz_fw_ray; float* flow_bw; float* xyz_bw_ray;
  float* rgb_fw; float* rgb_bw; float* disocc_fw; float* disocc_bw;
} nsff_composite_args;

NSFF_API int nsff_composite(const nsff_composite_args* args /*host*/, void* stream);

/* Backward of nsff_composite (autograd of rendering.py:202-298 / 98-140; the reference gets it from torch autograd,
 * train.py:178-198).  Upstream gradients of the differentiable outputs (NULL = zero; disocc* are detached in the
 * reference and have none) -> gradients w.r.t. raw (R,S,raw_ld), raw_fw / raw_bw (R,S,7).  The gradient w.r.t. the
 * warped points equals the upstream gradient of xyzs_fw_bw / xyzs_bw_fw and is left to the caller.  Training
 * composites only (static + transient); S <= 256. */
typedef struct {
  const float* static_sigmas; const float* transient_sigmas;
  const float* static_weights; const float* transient_weights; const float* weights;
  const float* xyzs_fw_bw; const float* xyzs_bw_fw;
  const float* depth; const float* rgb; const float* transient_alpha; const float* transient_rgb;
  const float* static_only_rgb; const float* static_only_depth;
  const float* xyz_fine; const float* flow_fw; const float* xyz_fw_ray; const float* flow_bw; const float* xyz_bw_ray;
  const float* rgb_fw; const float* rgb_bw;
  float* d_raw; float* d_raw_fw; float* d_raw_bw;      /* outputs, fully overwritten */
} nsff_composite_grads;

NSFF_API int nsff_composite_backward(const nsff_composite_args* args /*host*/,
                                     const nsff_composite_grads* grads /*host*/, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a13  sample_pdf + merge                                   models/rendering.py:10-49, 335-348
 *     sample_pdf(bins, weights, N_importance, det, eps): bins (R, n_weights+1), weights
 *     (R, n_weights), each with its own row stride in floats (0 = one row shared by all rays, which
 *     is how render_rays passes the un-jittered mid-points, :315,:336); u: (R,N) uniforms, or NULL
 *     for det=True (u = linspace(0,1,N)).  out: (R,N) new depths.
 */
NSFF_API int nsff_sample_pdf(const float* bins, int64_t bins_row_stride, const float* weights,
                    int64_t weights_row_stride, int64_t R, int n_weights, int n_importance,
                    const float* u, float eps, float* out, void* stream);
/*     out (R, total) = sort(cat(parts...)) along the last axis; up to 3 parts (:348).        */
NSFF_API int nsff_merge_sorted(const float* a, int na, const float* b, int nb, const float* c, int nc,
                      int64_t R, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * f1  interpolate + softmax-splat ("average")            models/rendering.py:365-460, softsplat.py:6-44,303-329
 *     Renders time t+dt from two render_rays(test_time=True, flows) results of the same rays: every sample plane of
 *     the dynamic RGBA of frame t is forward-splatted by the projected dt-scaled forward flow, that of frame t+1 by
 *     the (1-dt)-scaled backward flow, both are blended and composited front to back with the static plane.
 *     Per-sample inputs are (R,S,*) with row stride ld_t / ld_tp1 floats (they are views of the field output);
 *     R = w*h rays in pixel order.  P = K @ w2c (rows 1,2 of w2c negated), row-major 3x4.
 */
typedef struct {
  int64_t R; int32_t S; int32_t w, h;
  float dt, fx, fy, cx, cy;
  float P[12];
  const float* xyzs;          /* (R,S,3) contiguous                                   */
  const float* zs;            /* (R,S)                                                */
  const float* static_rgb;    /* (R,S,3) view, stride ld_t                            */
  const float* static_alpha;  /* (R,S)                                                */
  const float* rgb_t;  const float* alpha_t;  const float* flow_fw;   int32_t ld_t;   /* frame t   */
  const float* rgb_tp1; const float* alpha_tp1; const float* flow_bw; int32_t ld_tp1; /* frame t+1 */
} nsff_interpolate_args;

NSFF_API size_t nsff_interpolate_scratch_floats(int64_t R, int S);
NSFF_API int nsff_interpolate(const nsff_interpolate_args* args /*host*/, float* scratch, float* rgb /*(h,w,3)*/,
                              float* depth /*(h,w)*/, void* stream);

/* ---------------------------------------------------------------------------------------------
 * f4  on-device ray generation                    datasets/ray_utils.py:7-106, monocular.py:268-276
 *     nsff_frame_rays: all H*W NDC rays (o,d) of a pinhole frame in pixel order = get_ray_directions ->
 *     get_rays (normalised) -> get_ndc_rays.  K4 = {fx, fy, cx, cy} and c2w12 = row-major 3x4, both HOST.
 *     nsff_ndc_rays: get_ndc_rays alone on caller-provided world rays (near / shift_near scalars).
 */
NSFF_API int nsff_frame_rays(const float* K4 /*host*/, const float* c2w12 /*host*/, float near, float shift_near,
                             int H, int W, float* rays /*(H*W,6)*/, void* stream);
NSFF_API int nsff_ndc_rays(const float* K4 /*host*/, float near, float shift_near, const float* rays_o,
                           const float* rays_d, int64_t n, float* out_o, float* out_d, void* stream);

/* ---------------------------------------------------------------------------------------------
 * f2  NeRFWLoss                                                     losses.py:8-29, 31-171
 *     The training loss on the result dict of render_rays(test_time=False, output_transient_flow=['fw','bw','disocc']):
 *     colour, shift/scale-invariant depth, entropy, static/dynamic cross-entropy, 2D-3D flow consistency, photometric
 *     warp, cycle consistency and the three scene-flow regularisers - every term of losses.py:72-161 - in three
 *     launches forward and one backward (the reference spends ~300 ATen kernels on it).  All tensors fp32 contiguous;
 *     `medians` = {median(depth_fine), median(-disps), median(depth_coarse)} (lower medians as torch.median, device
 *     floats): for R <= 2048 the forward computes and writes them itself, for larger R the caller supplies them;
 *     `terms` receives the NSFF_LOSS_TERMS means in nsff_loss_term order (flow terms: mean over the rays whose
 *     projected depth is positive and that have a neighbouring frame; 0 when there is none).  topk < 1 is not supported.
 */
enum nsff_loss_term {
  NSFF_L_COL = 0, NSFF_L_DISP, NSFF_L_ENT, NSFF_L_CROSS, NSFF_L_FLOW_FW, NSFF_L_FLOW_BW, NSFF_L_PHO, NSFF_L_CYC,
  NSFF_L_REG_TEMP, NSFF_L_REG_MIN, NSFF_L_REG_SP, NSFF_LOSS_TERMS
};
typedef struct {
  int64_t R; int32_t S;
  int32_t has_flow;            /* kwargs['output_transient_flow'] non-empty: all terms; else colour + depth only   */
  int32_t thickness;           /* window of the zero-padded box filter over the transient weights (losses.py:63)   */
  int32_t n_reg;               /* int(S * z_far): samples the regularisers see (losses.py:147)                     */
  int32_t n_frames, max_t;
  float lambda_geo_d, lambda_geo_f, lambda_reg, lambda_ent;
  float cross_entropy_w;       /* lambda_ent/5 * min(epoch/10, 1)  (losses.py:90)                                  */
  /* render_rays outputs */
  const float *rgb_fine, *depth_fine, *rgb_coarse, *depth_coarse;       /* (R,3) (R); coarse pair may be NULL      */
  const float *tw, *sw;                                                  /* transient / static weights (R,S)        */
  const float *xyz_fw, *xyz_bw, *rgb_fw, *rgb_bw;                        /* (R,3)                                   */
  const float *disocc_fw, *disocc_bw;                                    /* (R)    detached                         */
  const float *disoccs_fw, *disoccs_bw;                                  /* (R,S)  detached                         */
  const float *xyzs_fw_bw, *xyzs_bw_fw, *xyzs_fine, *xyzs_fw, *xyzs_bw;  /* (R,S,3)                                 */
  /* targets (the training batch, datasets/monocular.py:240-249) and the buffers train.py:135-138 registers */
  const float *rgbs, *disps, *uv_fw, *uv_bw;
  const int64_t *cam_ids, *ts;
  const float *Ks;             /* (n_cams, 3, 3)                                                                     */
  const float *Ps;             /* (n_cams, n_frames, 3, 4)                                                           */
  const float *ray_weights;    /* kwargs['weights'] (R) or NULL                                                      */
  float *medians;              /* 3 device floats, see above: WRITTEN by the forward when R <= 2048, else supplied   */
  float *stats;                /* nsff_loss_stats_floats() floats, written by the forward, read by the backward      */
  float *term_sums;            /* NSFF_LOSS_TERMS floats of scratch                                                  */
  float *terms;                /* NSFF_LOSS_TERMS floats: the loss dict                                              */
  /* backward: upstream gradient of every term and the gradients of their g-weighted sum (overwritten) */
  const float *g_terms;
  float *d_rgb_fine, *d_depth_fine, *d_rgb_coarse, *d_depth_coarse, *d_tw, *d_sw, *d_xyz_fw, *d_xyz_bw, *d_rgb_fw,
        *d_rgb_bw, *d_xyzs_fw_bw, *d_xyzs_bw_fw, *d_xyzs_fw, *d_xyzs_bw;
} nsff_loss_args;

NSFF_API int nsff_loss_stats_floats(void);
NSFF_API int nsff_nerfw_loss_forward(const nsff_loss_args* args /*host*/, void* stream);
NSFF_API int nsff_nerfw_loss_backward(const nsff_loss_args* args /*host*/, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NSFF_B200_H_ */
