/*
 * bts_b200.h -- C ABI of the B200-native BehindTheScenes render path (libbts_b200.so).
 *
 * The reference (Brummi/BehindTheScenes) is pure Python/PyTorch and has no FFI layer; its seam for this path is the
 * Python object contract of BTSNet / NeRFRenderer / _RenderWrapper.  This header is the boundary a
 * binding would target: plain pointers + sizes + a cudaStream_t, no torch types.  Every entry point cites the
 * reference interface it replaces (paths relative to the reference root).  All pointers are DEVICE pointers unless
 * stated otherwise, all tensors contiguous fp32, all calls asynchronous on `stream` (no internal sync, no host
 * allocation on the call path).  Return value: 0 on success, negative BtsStatus otherwise; the text of the last error
 * of the calling thread is available from bts_last_error().
 *
 * Data layout in HBM (what "encode" leaves resident for the render kernels):
 *   G      (n, H, W, Hd)      "folded feature map": G[p] = W_in[:, :C] * F[:, p] + b_in   (bts_fold_features)
 *   g_empty(Hd)               W_in[:, :C] * empty_feature + b_in                      (learn_empty)
 *   img4   (n, Vc, Hc, Wc, 4) colour views, channels-last padded to 4                (bts_pack_images)
 *   G16    3 x (n*H, ceil((W+6)/8), 2, Hd/8, 8, 8) fp16: G as tcgen05 operand blocks of 8 pixels of a row,
 *          [plane hi|lo][channel group][channel][pixel] (K-major core matrices, hi + lo = fp32-grade); three copies with the
 *          pixels shifted by 0, 3 and 6 so that a 30-pixel span always fits 4 blocks               (bts_pack_g16)
 *   K_*    (n, V, 3, 3), w2c_* (n, V, 4, 4) row-major camera matrices as the reference caches them
 */
#ifndef BTS_B200_H_
#define BTS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BTS_ABI_VERSION 11
#define BTS_MAX_BLOCKS 8   /* ResnetFC blocks supported by the kernels */

typedef struct CUstream_st* bts_stream_t; /* == cudaStream_t */

typedef enum {
  BTS_OK = 0,
  BTS_ERR_INVALID_ARGUMENT = -1,  /* null pointer / bad size                                  */
  BTS_ERR_UNSUPPORTED = -2,       /* configuration the CUDA path does not implement (raises)  */
  BTS_ERR_CUDA = -3               /* a CUDA runtime call or launch failed                      */
} BtsStatus;

/* Scene = what BTSNet.encode caches (models/bts/model/models_bts.py:128-136), in render-ready layout. */
typedef struct {
  int32_t n;            /* batch ("super-batch") size                                  */
  int32_t H, W;         /* feature-map size (scale-0 map is full resolution, :125-126) */
  int32_t Hd;           /* folded channels == mlp_coarse.d_hidden                       */
  const float* G;       /* (n, H, W, Hd)                                                */
  const float* g_empty; /* (Hd) or NULL when learn_empty is off                         */
  const float* K_f;     /* (n, 3, 3)  grid_f_Ks of the single encoder view              */
  const float* w2c_f;   /* (n, 4, 4)  grid_f_poses_w2c                                  */
  int32_t Vc;           /* colour/render views (grid_c_imgs.shape[1])                   */
  int32_t Cc;           /* colour channels per view (3; 4 with auto-masking), <= 4      */
  int32_t Hc, Wc;
  const float* img4;    /* (n, Vc, Hc, Wc, 4), values as cached by encode ([0,1])       */
  const float* K_c;     /* (n, Vc, 3, 3)                                                */
  const float* w2c_c;   /* (n, Vc, 4, 4)                                                */
  int32_t same_view_mask; /* bit v set: colour view v is the encoder view itself (same K, pose, size; eval_depth:
                             ids_render == ids_encoder) -> its projection and flags are reused, not recomputed */
  const void* G16;        /* operand-block copy of G (bts_pack_g16) or NULL: enables the band kernel (bilinear
                             interpolation of G on the tensor core from a TMA-staged pixel band)                     */
  const int32_t* g16_flags; /* (1) device word written by bts_pack_g16: non-zero when a value of G does not fit fp16
                             (the band kernel then hands every ray to the gather kernel)                            */
} BtsScene;

/* MLP = ResnetFC parameters (models/common/model/resnetfc.py:66-184), d_latent = 0, ReLU, d_out = 1.
 * lin_in.weight is (Hd, C + P): its first C columns are consumed by bts_fold_features, the render kernels read the
 * last P columns (the positional-code part, models/bts/model/models_bts.py:184 puts features first, code last). */
typedef struct {
  int32_t C;            /* encoder.latent_size                                          */
  int32_t P;            /* positional code width = d_in*(2*num_freqs) (+ d_in)          */
  int32_t Hd;           /* d_hidden (multiple of 32, <= 128)                            */
  int32_t n_blocks;     /* <= BTS_MAX_BLOCKS                                            */
  const float* w_in;    /* (Hd, C + P) row-major                                        */
  const float* b_in;    /* (Hd)                                                         */
  const float* fc0_w[BTS_MAX_BLOCKS]; /* (Hd, Hd) each */
  const float* fc0_b[BTS_MAX_BLOCKS];
  const float* fc1_w[BTS_MAX_BLOCKS];
  const float* fc1_b[BTS_MAX_BLOCKS];
  const float* w_out;   /* (1, Hd)                                                      */
  const float* b_out;   /* (1)                                                          */
} BtsMlp;

/* Field + renderer configuration: BTSNet.__init__ (models_bts.py:21-40), PositionalEncoding (code.py:12-28),
 * NeRFRenderer.__init__ (models/common/render/nerf.py:66-101). */
typedef struct {
  int32_t n_coarse;        /* K samples per ray                                          */
  int32_t lindisp;         /* nerf.py:117-120                                            */
  int32_t hard_alpha_cap;  /* nerf.py:285-286                                            */
  int32_t white_bkgd;      /* nerf.py:301-304                                            */
  int32_t inv_z;           /* models_bts.py:159-163                                      */
  int32_t code_mode;       /* 0 = "z", 1 = "distance" (models_bts.py:157-171)            */
  int32_t learn_empty;     /* models_bts.py:181-182 (needs scene.g_empty)                */
  int32_t num_freqs;       /* code.py:14                                                 */
  int32_t include_input;   /* code.py:17                                                 */
  float freq_factor;       /* code.py:15                                                 */
  float z_near, z_far;     /* BTSNet.d_min / d_max (for the depth code, not the ray)     */
  /* host-precomputed fp32 constants so the kernel rounds exactly like the reference's python-scalar arithmetic:   */
  float inv_z_far;         /* float(1 / d_max)                       models_bts.py:160   */
  float inv_z_range;       /* float(1 / d_min - 1 / d_max)                               */
  float z_range;           /* float(d_max - d_min)                   models_bts.py:162   */
  const float* z_steps;    /* (K) device: torch.linspace(0, 1 - 1/K, K)   nerf.py:114    */
  float step;              /* float(1.0 / K)                              nerf.py:112    */
  int32_t band_mode;       /* 0 = use the band kernel when scene.G16 and out.scratch are given and the shape allows,
                              1 = never (gather kernel)                                                           */
  const uint64_t* seed_dev; /* optional device word XOR-mixed into `seed` (in-kernel jitter): lets a captured CUDA graph draw
                              new jitter on every replay -- the caller increments the word inside the graph; NULL = unused */
  int32_t precision;       /* 0 = exact tier (tensor-core kernel with two-term operand split where the config
                              allows it, else the general fp32 kernel), 1 = fast tier (single-pass tensor core),
                              2 = exact tier forced onto the general CUDA-core kernel */
} BtsRenderCfg;

/* Optional outputs of NeRFRenderer._format_outputs (nerf.py:377-401).  NULL = not wanted (not computed/written). */
typedef struct {
  float* rgb;        /* (n, R, Vc*Cc)     view-major (models_bts.py:330)                */
  float* depth;      /* (n, R)                                                          */
  float* invalid;    /* (n, R, K, Vc)     0/1                                           */
  float* weights;    /* (n, R, K)                                                       */
  float* alphas;     /* (n, R, K)                                                       */
  float* z_samps;    /* (n, R, K)                                                       */
  float* rgb_samps;  /* (n, R, K, Vc*Cc)                                                */
  /* fused consumers of the per-sample outputs, all optional:                                  */
  float* invalid_mass; /* (n, R, Vc)  sum_k invalid[k,v] * weights[k]  -- the "weight guided" invalid test of
                          models/bts/model/loss.py:105-108 and scripts/inference_setup.py:192 without the (R,K,Vc) tensor */
  float* sigma;        /* (n, R, K)   per-sample density (softplus output, models_bts.py:313): saved by the autograd
                          node so that bts_render_bwd does not have to re-evaluate the MLP for the compositing pass  */
  float* depth_z;      /* (n, R)      z-depth = depth * (unit-ray z in its camera), utils/projection_operations.py:4-16;
                          only with in-kernel ray generation (BtsRayGen)                                            */
  int32_t* scratch;    /* device work list of the band kernel or NULL: >= BTS_BAND_SCRATCH_INTS(n*R) ints.  After the call
                          scratch[0] = number of 8-ray groups the band kernel handed to the gather kernel (debug)    */
  int64_t scratch_ints;
} BtsRenderOut;
#define BTS_BAND_SCRATCH_INTS(rays) (4 + ((rays) + 7) / 8)

/* In-kernel ray generation: replaces ImageRaySampler.sample -> util.gen_rays
 * (models/bts/model/ray_sampler.py:240-262, models/common/util/util.py:113-149, :244-273): one ray per pixel of every
 * view, d = R * normalize([(x-cx)/fx, (y-cy)/fy, 1]) with x,y on linspace(-1,1,.), o = t; ray index = ((b*v + view)*H + y)*W + x. */
typedef struct {
  const float* poses_c2w; /* (n, v, 4, 4) */
  const float* K;         /* (n, v, 3, 3) normalised intrinsics */
  int32_t v, H, W;
  int32_t norm_dir;       /* util.gen_rays(norm_dir=...) */
  float z_near, z_far;    /* ray near / far */
} BtsRayGen;

int bts_abi_version(void);
const char* bts_last_error(void);

/* Encode-side glue, once per encode() / weight update (SURVEY.md §7 "layer-1 linearity"):
 *   G[b, y, x, :] = w_in[:, :C] * feat[b, :, y, x] + b_in;   g_empty = w_in[:, :C] * empty_feature + b_in  (if non-NULL).
 * The lin_in bias is folded in here, so the render kernels never read mlp.b_in.
 * feat is the encoder output exactly as BTSNet.encode caches it: (n, C, H, W) NCHW (models_bts.py:126-128). */
int bts_fold_features(const float* feat_nchw, int32_t n, int32_t C, int32_t H, int32_t W, const float* w_in,
                      int32_t ld_w_in, const float* b_in, int32_t Hd, const float* empty_feature_or_null, float* G_out,
                      float* g_empty_out_or_null, bts_stream_t stream);

/* (n, Vc, Cc, Hc, Wc) NCHW colour views (grid_c_imgs, models_bts.py:133) -> (n, Vc, Hc, Wc, 4) channels-last. */
int bts_pack_images(const float* img_nchw, int32_t nv_total, int32_t Cc, int32_t Hc, int32_t Wc, float* img4_out,
                    bts_stream_t stream);

/* G (n, H, W, Hd) fp32 -> G16 operand blocks (layout above; pixels past W are zero) and the fp16-range flag.
 * Part of the encode-side glue like bts_fold_features: the bilinear feature fetch of models_bts.py:179 then runs as
 * tcgen05 MMAs over a shared-memory band of these blocks.  Hd in {64, 128}. */
int bts_pack_g16(const float* G, int32_t n, int32_t H, int32_t W, int32_t Hd, void* G16_out, int32_t* flags_out,
                 bts_stream_t stream);
/* bytes of the G16 buffer for a (n, H, W, Hd) map */
int64_t bts_g16_bytes(int32_t n, int32_t H, int32_t W, int32_t Hd);

/* Replaces _RenderWrapper.forward -> NeRFRenderer.forward -> sample_coarse + composite -> BTSNet.forward
 * (nerf.py:19-45, :315-375, :103-123, :210-313; models_bts.py:266-338) for the coarse pass.
 * rays (n, R, 8) = [origin, direction, near, far]; u (n*R, K) stratified jitter in [0,1) or NULL to draw it in-kernel
 * from a counter-based generator keyed by (seed, ray, sample) (the reference draws torch.rand_like, nerf.py:116). */
int bts_render_fwd(const BtsScene* scene, const BtsMlp* mlp, const BtsRenderCfg* cfg, const float* rays, int64_t n,
                   int64_t R, const float* u_or_null, uint64_t seed, const BtsRenderOut* out, bts_stream_t stream);

/* Same as bts_render_fwd with the rays generated in the kernel from camera poses (R = gen.v * gen.H * gen.W). */
int bts_render_views_fwd(const BtsScene* scene, const BtsMlp* mlp, const BtsRenderCfg* cfg, const BtsRayGen* gen,
                         int64_t n, const float* u_or_null, uint64_t seed, const BtsRenderOut* out, bts_stream_t stream);

/* Upstream gradients of a scalar loss w.r.t. the differentiable outputs of bts_render_fwd (NULL = zero); shapes as
 * in BtsRenderOut.  invalid / z_samps / rgb_samps do not depend on the trainable tensors. */
typedef struct {
  const float* rgb;      /* (n, R, Vc*Cc) */
  const float* depth;    /* (n, R)        */
  const float* weights;  /* (n, R, K)     */
  const float* alphas;   /* (n, R, K)     */
} BtsRenderGrad;

/* Gradients produced by bts_render_bwd; every buffer is ACCUMULATED into (caller zero-initialises), NULL = skip. */
typedef struct {
  float* d_G;        /* (n, H, W, Hd)  gradient of the folded map (bilinear scatter-add)                 */
  float* d_g_empty;  /* (Hd)                                                                            */
  float* d_w_pe;     /* (Hd, P)        gradient of lin_in.weight[:, C:] (the positional-code columns)     */
  float* d_w_out;    /* (Hd)           gradient of lin_out.weight                                         */
  float* d_b_out;    /* (1)            gradient of lin_out.bias                                           */
  float* d_fc0_w;    /* (Hd, Hd)       blocks.0.fc_0.weight   } only for n_blocks = 1 (RE10K: d_hidden = 32),    */
  float* d_fc0_b;    /* (Hd)           blocks.0.fc_0.bias     } NULL otherwise                                    */
  float* d_fc1_w;    /* (Hd, Hd)       blocks.0.fc_1.weight                                                       */
  float* d_fc1_b;    /* (Hd)           blocks.0.fc_1.bias                                                         */
} BtsParamGrad;

/* Backward of bts_render_fwd (what torch autograd does for the reference's op chain during
 * `loss.backward()`, utils/base_trainer.py:297): recomputes the forward per ray from the same inputs (same `u` or
 * `seed`; with `saved_sigma` (n,R,K) from the forward's BtsRenderOut.sigma the MLP is evaluated once instead of twice), then back-propagates through compositing (nerf.py:283-304), softplus, lin_out and lin_in.  The dense tail
 * of the chain is the caller's: dF = d_G * W_in[:, :C], dW_in[:, :C] = d_G^T * F (+ d_g_empty x empty_feature),
 * db_in = sum(d_G) + d_g_empty.  Supports n_blocks = 0 with d_hidden <= 64 (the KITTI training configs) and
 * n_blocks = 1 with d_hidden = 32 (the RE10K training config). */
int bts_render_bwd(const BtsScene* scene, const BtsMlp* mlp, const BtsRenderCfg* cfg, const float* rays, int64_t n,
                   int64_t R, const float* u_or_null, uint64_t seed, const float* saved_sigma_or_null,
                   const BtsRenderGrad* upstream, const BtsParamGrad* grads, bts_stream_t stream);

/* Replaces a direct BTSNet.forward(xyz, only_density=...) point query (models_bts.py:266-338; callers
 * scripts/inference_setup.py:214, models/bts/evaluator_lidar.py:305).  xyz (n, P, 3).
 * sigma (n, P); invalid_f (n, P) feature-view flag; rgb (n, P, Vc*Cc) and invalid (n, P, Vc) may be NULL
 * (only_density). */
int bts_query_field(const BtsScene* scene, const BtsMlp* mlp, const BtsRenderCfg* cfg, const float* xyz, int64_t n,
                    int64_t P, float* sigma, float* invalid_f, float* rgb_or_null, float* invalid_or_null,
                    bts_stream_t stream);

/* Backward of bts_fold_features (the dense tail of bts_render_bwd): given d_G (n, H, W, Hd) accumulated by
 * bts_render_bwd,  d_feat (n, C, H, W) = W_F^T d_G  (written; NULL when the encoder is frozen),
 * d_w_f (Hd, C) += d_G^T F  and  d_b_in (Hd) += sum d_G  (both ACCUMULATED: zero them first).  This is the lin_in part of
 * what torch autograd does in loss.backward() (utils/base_trainer.py:297; models/common/model/resnetfc.py:88,147).
 * Tensor-core kernel for C in {32, 64}, Hd in {32, 64}, H*W % 4 == 0 and 16-byte aligned tensors; a CUDA-core kernel for
 * every other shape. */
int bts_fold_features_bwd(const float* d_G, const float* feat_nchw, int32_t n, int32_t C, int32_t H, int32_t W,
                          const float* w_in, int32_t ld, int32_t Hd, float* d_feat_or_null, float* d_w_f, float* d_b_in,
                          bts_stream_t stream);

/* Sparse fold, the training form of the encode-side glue.  A training batch (PatchRaySampler: 8x8 pixel patches,
 * models/bts/model/ray_sampler.py:108-195) samples the folded map only along the epipolar lines of its patches, typically
 * 15-25 % of a frame, while F.grid_sample of the reference (models_bts.py:179) reads -- and autograd differentiates -- only
 * the sampled texels.  tile_mask holds one byte per 64 consecutive pixels of a frame (n * BTS_TILE_MASK_PER_FRAME(H, W)
 * bytes; 0 = no ray reads it, 1 = needed, 2 = folded), zeroed by the caller once per encode() / weight update:
 *   bts_mark_tiles            flags the tiles the taps of a ray batch fall into (the renderers' own sampling, projection and
 *                             tap code: exactly the texels bts_render_fwd / bts_render_bwd will touch)
 *   bts_fold_features_sparse  bts_fold_features for the tiles in state 1 only (-> state 2); G elsewhere stays undefined
 *   bts_zero_tiles            zeroes the d_G rows of the tiles in state != 0 (instead of a memset of the whole tensor)
 *   bts_fold_features_bwd_sparse  bts_fold_features_bwd over those tiles; d_feat rows of all other tiles are written as zeros
 * bts_sparse_fold_supported: 1 when the tensor-core fold kernels take the shape (the sparse path has no CUDA-core form). */
#define BTS_TILE_PIXELS 64
#define BTS_TILE_MASK_PER_FRAME(H, W) (((int64_t)(H) * (W) + BTS_TILE_PIXELS - 1) / BTS_TILE_PIXELS)
int bts_sparse_fold_supported(int32_t C, int32_t H, int32_t W, int32_t Hd);
int bts_mark_tiles(const BtsScene* scene, const BtsRenderCfg* cfg, const float* rays, int64_t n, int64_t R,
                   const float* u_or_null, uint64_t seed, uint8_t* tile_mask, bts_stream_t stream);
int bts_fold_features_sparse(const float* feat_nchw, int32_t n, int32_t C, int32_t H, int32_t W, const float* w_in,
                             int32_t ld_w_in, const float* b_in, int32_t Hd, const float* empty_feature_or_null,
                             float* G_out, float* g_empty_out_or_null, uint8_t* tile_mask, bts_stream_t stream);
int bts_zero_tiles(float* d_G, int32_t n, int32_t H, int32_t W, int32_t Hd, const uint8_t* tile_mask, bts_stream_t stream);
int bts_fold_features_bwd_sparse(const float* d_G, const float* feat_nchw, int32_t n, int32_t C, int32_t H, int32_t W,
                                 const float* w_in, int32_t ld, int32_t Hd, float* d_feat_or_null, float* d_w_f,
                                 float* d_b_in, const uint8_t* tile_mask, bts_stream_t stream);

/* Patch photometric loss of the training step: replaces compute_errors_l1ssim (models/bts/model/loss.py:10-18, SSIM of
 * models/common/model/layers.py:79-150 with zero padding, 3x3 gaussian window, comp_mode) and the reduction of
 * ReconstructionLoss.__call__ (loss.py:151-167): per ray 0.85 * mean_c(clamp(1 - SSIM, 0, 1) / 2) + 0.15 * mean_c|x - y|,
 * minimum over the nv rendered views, times (1 - invalid ray).
 *   pred (P, h, w, nv, c): the renderer's rgb output in patch order (PatchRaySampler.reconstruct, ray_sampler.py:164-200)
 *   gt   (P, h, w, c);  invalid_ray (P, h, w) 0/1 or NULL;  alternatively invalid_mass (P, h, w, nv) (BtsRenderOut.invalid_mass)
 *   with the "weight_guided" rule invalid = all_v(mass_v > mass_thresh) (loss.py:105-108) applied in the kernel.
 * Outputs: per_ray (P, h, w) (its mean is the reference's rgb_loss), argmin (P, h, w) view index (needed by the backward;
 * may be NULL), errors (P, h, w, nv) un-reduced criterion or NULL. */
int bts_patch_loss_fwd(const float* pred, const float* gt, const float* invalid_ray_or_null,
                       const float* invalid_mass_or_null, float mass_thresh, int64_t P, int32_t h, int32_t w, int32_t nv,
                       int32_t c, float* per_ray, int32_t* argmin_or_null, float* errors_or_null, bts_stream_t stream);

/* Gradient of per_ray with respect to pred: d_pred (P, h, w, nv, c) is fully written (zero for views that do not attain the
 * minimum and for invalid rays); g_per_ray (P, h, w) is the upstream gradient (1 / numel for the plain mean).  Ties in the
 * view minimum send the gradient to the first minimal view (torch.amin splits it; ties have measure zero). */
int bts_patch_loss_bwd(const float* pred, const float* gt, const float* invalid_ray_or_null,
                       const float* invalid_mass_or_null, float mass_thresh, int64_t P, int32_t h, int32_t w, int32_t nv,
                       int32_t c, const int32_t* argmin, const float* g_per_ray, float* d_pred, bts_stream_t stream);

/* Edge-aware smoothness of the rendered depth: replaces edge_aware_smoothness (models/bts/model/loss.py:21-40) and its
 * invalid-ray mask (loss.py:247-256).  depth (P, h, w) ray distances, gt (P, h, w, c) the target colours of the same rays,
 * masks as in bts_patch_loss_fwd (nv = number of views in invalid_mass).  per_ray (P, h, w): its mean is loss_eas.
 * The backward writes d_depth (P, h, w) from the upstream gradient g_per_ray (P, h, w). */
int bts_patch_smooth_fwd(const float* depth, const float* gt, const float* invalid_ray_or_null,
                         const float* invalid_mass_or_null, float mass_thresh, int64_t P, int32_t h, int32_t w, int32_t nv,
                         int32_t c, float* per_ray, bts_stream_t stream);
int bts_patch_smooth_bwd(const float* depth, const float* gt, const float* invalid_ray_or_null,
                         const float* invalid_mass_or_null, float mass_thresh, int64_t P, int32_t h, int32_t w, int32_t nv,
                         int32_t c, const float* g_per_ray, float* d_depth, bts_stream_t stream);

/* Name of the kernel family the last bts_render_fwd on this thread dispatched to ("general" | "tc" | "band"), for tests. */
const char* bts_last_kernel(void);
/* Same for the last bts_render_bwd: "tc" (tensor-core backward: d_hidden 64, no blocks, saved sigma, K % 32 == 0) | "general". */
const char* bts_last_bwd_kernel(void);

/* Debug/test hook (synchronises): number of bounded mbarrier waits that expired inside the tensor-core kernel since
 * the library was loaded.  Always 0 unless the kernel has a pipeline bug. */
int bts_debug_tc_timeouts(void);

#ifdef __cplusplus
}
#endif
#endif /* BTS_B200_H_ */
