// C ABI of libbts_b200.so (include/bts_b200.h): argument validation, kernel-family dispatch, error reporting.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "common.cuh"
#include "bwd_params.cuh"
#include "patch_loss.cuh"

namespace bts {
cudaError_t launch_general(const KernelParams& p, bool points, cudaStream_t stream);
size_t general_smem_bytes(int Hd, int p_pad, int n_blocks);
cudaError_t launch_mark_tiles(const KernelParams& p, unsigned char* mask, cudaStream_t stream);
cudaError_t launch_zero_tiles(float* dG, int n, int64_t HW, int Hd, const unsigned char* mask, cudaStream_t stream);
bool fold_tc_supported(const float* feat, int C, int64_t HW, int Hd);
cudaError_t launch_fold(const float* feat, int n, int C, int H, int W, const float* w_in, int ld, const float* b_in, int Hd,
                        const float* empty, float* G, float* g_empty, unsigned char* tile_mask, cudaStream_t stream);
cudaError_t launch_bwd(const BwdParams& bp, cudaStream_t stream, bool* used_tc);
int bwd_tc_barrier_timeouts();
size_t bwd_smem_bytes(int Hd, int p_pad, int K, int n_blocks);
bool tc_supported(const KernelParams& p);
cudaError_t launch_tc(const KernelParams& p, bool points, cudaStream_t stream);
int tc_barrier_timeouts();
int fold_barrier_timeouts();
int fold_bwd_barrier_timeouts();
bool fold_bwd_tc_supported(const float* dG, const float* feat, const float* d_feat, int C, int64_t HW, int Hd);
cudaError_t launch_fold_bwd_tc(const float* dG, const float* feat, int n, int C, int64_t HW, const float* w_in, int ld, int Hd,
                               float* d_feat, float* d_wf, float* d_b, const unsigned char* tile_mask, cudaStream_t stream);
cudaError_t launch_patch_loss(const PatchLossParams& p, bool backward, cudaStream_t stream);
bool patch_loss_supported(int h, int w, int nv, int c);
cudaError_t launch_patch_smooth(const PatchLossParams& p, bool backward, cudaStream_t stream);
cudaError_t launch_pack_images(const float* img, int nv_total, int Cc, int Hc, int Wc, float* out, cudaStream_t stream);
bool band_supported(const KernelParams& p);
cudaError_t launch_band(const KernelParams& p, cudaStream_t stream);
cudaError_t launch_pack_g16(const float* G, int n, int H, int W, int Hd, void* out, int32_t* flags, cudaStream_t stream);
int64_t g16_bytes(int n, int H, int W, int Hd);
int band_barrier_timeouts();
cudaError_t launch_fold_bwd_simt(const float* dG, const float* feat, int n, int C, int64_t HW, const float* w_in, int ld, int Hd,
                                 float* d_feat, float* d_wf, float* d_b, cudaStream_t stream);
}  // namespace bts

namespace {
thread_local char g_err[512] = "";
thread_local const char* g_kernel = "none";
const char* volatile g_bwd_kernel = "none";   // process-wide: autograd runs the backward on its own thread

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_common(const BtsScene* s, const BtsMlp* m, const BtsRenderCfg* c) {
  if (!s || !m || !c) return fail(BTS_ERR_INVALID_ARGUMENT, "null scene/mlp/cfg");
  if (!s->G || !s->K_f || !s->w2c_f) return fail(BTS_ERR_INVALID_ARGUMENT, "scene.G / K_f / w2c_f is null");
  if (s->Vc > 0 && (!s->img4 || !s->K_c || !s->w2c_c)) return fail(BTS_ERR_INVALID_ARGUMENT, "scene colour views are null");
  if (s->Vc < 0 || s->Vc > bts::kMaxColorViews) return fail(BTS_ERR_UNSUPPORTED, "Vc=%d not in [0,%d]", s->Vc, bts::kMaxColorViews);
  if (s->Cc < 1 || s->Cc > 4) return fail(BTS_ERR_UNSUPPORTED, "Cc=%d not in [1,4]", s->Cc);
  if (s->n < 1 || s->H < 1 || s->W < 1) return fail(BTS_ERR_INVALID_ARGUMENT, "bad scene sizes");
  if (m->Hd != s->Hd) return fail(BTS_ERR_INVALID_ARGUMENT, "mlp.Hd=%d != scene.Hd=%d", m->Hd, s->Hd);
  if (m->Hd % 32 != 0 || m->Hd < 32 || m->Hd > 128) return fail(BTS_ERR_UNSUPPORTED, "d_hidden=%d must be 32/64/96/128", m->Hd);
  if (m->n_blocks < 0 || m->n_blocks > BTS_MAX_BLOCKS) return fail(BTS_ERR_UNSUPPORTED, "n_blocks=%d > %d", m->n_blocks, BTS_MAX_BLOCKS);
  if (!m->w_in || !m->b_in || !m->w_out || !m->b_out) return fail(BTS_ERR_INVALID_ARGUMENT, "mlp weights are null");
  for (int b = 0; b < m->n_blocks; ++b)
    if (!m->fc0_w[b] || !m->fc0_b[b] || !m->fc1_w[b] || !m->fc1_b[b]) return fail(BTS_ERR_INVALID_ARGUMENT, "block %d weights are null", b);
  const int p_expect = 3 * 2 * c->num_freqs + (c->include_input ? 3 : 0);
  if (m->P != p_expect) return fail(BTS_ERR_INVALID_ARGUMENT, "mlp.P=%d but code width is %d", m->P, p_expect);
  if (c->learn_empty && !s->g_empty) return fail(BTS_ERR_INVALID_ARGUMENT, "learn_empty set but scene.g_empty is null");
  if (c->code_mode != 0 && c->code_mode != 1) return fail(BTS_ERR_UNSUPPORTED, "code_mode=%d", c->code_mode);
  const size_t smem = bts::general_smem_bytes(m->Hd, (m->P + 3) & ~3, m->n_blocks);
  if (smem > 200 * 1024) return fail(BTS_ERR_UNSUPPORTED, "ResnetFC blocks (%d x %d^2) do not fit in shared memory", m->n_blocks, m->Hd);
  return BTS_OK;
}
}  // namespace

extern "C" {

int bts_abi_version(void) { return BTS_ABI_VERSION; }
const char* bts_last_error(void) { return g_err; }
const char* bts_last_kernel(void) { return g_kernel; }
const char* bts_last_bwd_kernel(void) { return g_bwd_kernel; }
int bts_debug_tc_timeouts(void) {
  return bts::tc_barrier_timeouts() + bts::fold_barrier_timeouts() + bts::fold_bwd_barrier_timeouts() + bts::bwd_tc_barrier_timeouts() +
         bts::band_barrier_timeouts();
}

int64_t bts_g16_bytes(int32_t n, int32_t H, int32_t W, int32_t Hd) { return bts::g16_bytes(n, H, W, Hd); }

int bts_pack_g16(const float* G, int32_t n, int32_t H, int32_t W, int32_t Hd, void* G16_out, int32_t* flags_out,
                 bts_stream_t stream) {
  if (!G || !G16_out || !flags_out) return fail(BTS_ERR_INVALID_ARGUMENT, "bts_pack_g16: null pointer");
  if (n < 1 || H < 1 || W < 1) return fail(BTS_ERR_INVALID_ARGUMENT, "bts_pack_g16: bad sizes");
  if (Hd != 64 && Hd != 128) return fail(BTS_ERR_UNSUPPORTED, "bts_pack_g16: d_hidden=%d must be 64 or 128", Hd);
  const cudaError_t e = bts::launch_pack_g16(G, n, H, W, Hd, G16_out, flags_out, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(BTS_ERR_CUDA, "bts_pack_g16: %s", cudaGetErrorString(e));
  return BTS_OK;
}

int bts_fold_features(const float* feat_nchw, int32_t n, int32_t C, int32_t H, int32_t W, const float* w_in,
                      int32_t ld_w_in, const float* b_in, int32_t Hd, const float* empty_feature_or_null, float* G_out,
                      float* g_empty_out_or_null, bts_stream_t stream) {
  if (!feat_nchw || !w_in || !b_in || !G_out) return fail(BTS_ERR_INVALID_ARGUMENT, "bts_fold_features: null pointer");
  if (n < 1 || C < 1 || H < 1 || W < 1 || ld_w_in < C) return fail(BTS_ERR_INVALID_ARGUMENT, "bts_fold_features: bad sizes");
  if (Hd % 32 != 0 || Hd < 32 || Hd > 128) return fail(BTS_ERR_UNSUPPORTED, "d_hidden=%d must be 32/64/96/128", Hd);
  const cudaError_t e = bts::launch_fold(feat_nchw, n, C, H, W, w_in, ld_w_in, b_in, Hd, empty_feature_or_null, G_out,
                                         g_empty_out_or_null, nullptr, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(BTS_ERR_CUDA, "bts_fold_features: %s", cudaGetErrorString(e));
  return BTS_OK;
}

// ---- sparse fold (training) -----------------------------------------------------------------------------------------
int bts_sparse_fold_supported(int32_t C, int32_t H, int32_t W, int32_t Hd) {
  const int64_t HW = (int64_t)H * W;
  return (bts::fold_tc_supported(nullptr, C, HW, Hd) && bts::fold_bwd_tc_supported(nullptr, nullptr, nullptr, C, HW, Hd)) ? 1 : 0;
}

int bts_mark_tiles(const BtsScene* scene, const BtsRenderCfg* cfg, const float* rays, int64_t n, int64_t R,
                   const float* u_or_null, uint64_t seed, uint8_t* tile_mask, bts_stream_t stream) {
  if (!scene || !cfg || !rays || !tile_mask) return fail(BTS_ERR_INVALID_ARGUMENT, "bts_mark_tiles: null pointer");
  if (!scene->K_f || !scene->w2c_f || scene->n < 1 || scene->H < 1 || scene->W < 1)
    return fail(BTS_ERR_INVALID_ARGUMENT, "bts_mark_tiles: scene needs n, H, W, K_f, w2c_f");
  if (n != scene->n) return fail(BTS_ERR_INVALID_ARGUMENT, "rays batch %lld != scene.n %d", (long long)n, scene->n);
  if (cfg->n_coarse < 1 || !cfg->z_steps) return fail(BTS_ERR_INVALID_ARGUMENT, "n_coarse < 1 or z_steps null");
  if (n * R <= 0) return BTS_OK;
  bts::KernelParams p;
  memset(&p, 0, sizeof(p));
  p.scene = *scene; p.cfg = *cfg;
  p.rays = rays; p.u = u_or_null; p.n_items = n * R; p.per_batch = R; p.seed = seed;
  const cudaError_t e = bts::launch_mark_tiles(p, tile_mask, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(BTS_ERR_CUDA, "bts_mark_tiles: %s", cudaGetErrorString(e));
  return BTS_OK;
}

int bts_fold_features_sparse(const float* feat_nchw, int32_t n, int32_t C, int32_t H, int32_t W, const float* w_in,
                             int32_t ld_w_in, const float* b_in, int32_t Hd, const float* empty_feature_or_null,
                             float* G_out, float* g_empty_out_or_null, uint8_t* tile_mask, bts_stream_t stream) {
  if (!feat_nchw || !w_in || !b_in || !G_out || !tile_mask) return fail(BTS_ERR_INVALID_ARGUMENT, "bts_fold_features_sparse: null pointer");
  if (n < 1 || C < 1 || H < 1 || W < 1 || ld_w_in < C) return fail(BTS_ERR_INVALID_ARGUMENT, "bts_fold_features_sparse: bad sizes");
  if (!bts::fold_tc_supported(feat_nchw, C, (int64_t)H * W, Hd))
    return fail(BTS_ERR_UNSUPPORTED, "bts_fold_features_sparse: shape / alignment outside the tensor-core fold (see bts_sparse_fold_supported)");
  const cudaError_t e = bts::launch_fold(feat_nchw, n, C, H, W, w_in, ld_w_in, b_in, Hd, empty_feature_or_null, G_out,
                                         g_empty_out_or_null, tile_mask, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(BTS_ERR_CUDA, "bts_fold_features_sparse: %s", cudaGetErrorString(e));
  return BTS_OK;
}

int bts_zero_tiles(float* d_G, int32_t n, int32_t H, int32_t W, int32_t Hd, const uint8_t* tile_mask, bts_stream_t stream) {
  if (!d_G || !tile_mask) return fail(BTS_ERR_INVALID_ARGUMENT, "bts_zero_tiles: null pointer");
  if (n < 1 || H < 1 || W < 1 || Hd < 4 || Hd % 4 != 0 || (reinterpret_cast<uintptr_t>(d_G) & 15) != 0)
    return fail(BTS_ERR_INVALID_ARGUMENT, "bts_zero_tiles: bad sizes / alignment");
  const cudaError_t e = bts::launch_zero_tiles(d_G, n, (int64_t)H * W, Hd, tile_mask, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(BTS_ERR_CUDA, "bts_zero_tiles: %s", cudaGetErrorString(e));
  return BTS_OK;
}

int bts_fold_features_bwd_sparse(const float* d_G, const float* feat_nchw, int32_t n, int32_t C, int32_t H, int32_t W,
                                 const float* w_in, int32_t ld, int32_t Hd, float* d_feat_or_null, float* d_w_f,
                                 float* d_b_in, const uint8_t* tile_mask, bts_stream_t stream) {
  if (!d_G || !feat_nchw || !w_in || !d_w_f || !d_b_in || !tile_mask) return fail(BTS_ERR_INVALID_ARGUMENT, "bts_fold_features_bwd_sparse: null pointer");
  if (n < 1 || C < 1 || H < 1 || W < 1 || Hd < 1 || ld < C) return fail(BTS_ERR_INVALID_ARGUMENT, "bts_fold_features_bwd_sparse: bad sizes");
  const int64_t HW = (int64_t)H * W;
  if (!bts::fold_bwd_tc_supported(d_G, feat_nchw, d_feat_or_null, C, HW, Hd))
    return fail(BTS_ERR_UNSUPPORTED, "bts_fold_features_bwd_sparse: shape / alignment outside the tensor-core fold (see bts_sparse_fold_supported)");
  const cudaError_t e = bts::launch_fold_bwd_tc(d_G, feat_nchw, n, C, HW, w_in, ld, Hd, d_feat_or_null, d_w_f, d_b_in, tile_mask,
                                                reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(BTS_ERR_CUDA, "bts_fold_features_bwd_sparse: %s", cudaGetErrorString(e));
  return BTS_OK;
}

int bts_pack_images(const float* img_nchw, int32_t nv_total, int32_t Cc, int32_t Hc, int32_t Wc, float* img4_out,
                    bts_stream_t stream) {
  if (!img_nchw || !img4_out) return fail(BTS_ERR_INVALID_ARGUMENT, "bts_pack_images: null pointer");
  if (Cc < 1 || Cc > 4) return fail(BTS_ERR_UNSUPPORTED, "Cc=%d not in [1,4]", Cc);
  const cudaError_t e = bts::launch_pack_images(img_nchw, nv_total, Cc, Hc, Wc, img4_out, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(BTS_ERR_CUDA, "bts_pack_images: %s", cudaGetErrorString(e));
  return BTS_OK;
}

static int render_fwd_common(const char* who, const BtsScene* scene, const BtsMlp* mlp, const BtsRenderCfg* cfg,
                             const float* rays, const BtsRayGen* gen, int64_t n, int64_t R, const float* u_or_null,
                             uint64_t seed, const BtsRenderOut* out, bts_stream_t stream) {
  int rc = check_common(scene, mlp, cfg);
  if (rc != BTS_OK) return rc;
  if (!out) return fail(BTS_ERR_INVALID_ARGUMENT, "%s: out is null", who);
  if (n != scene->n) return fail(BTS_ERR_INVALID_ARGUMENT, "rays batch %lld != scene.n %d", (long long)n, scene->n);
  if (R < 0) return fail(BTS_ERR_INVALID_ARGUMENT, "R < 0");
  if (cfg->n_coarse < 1 || !cfg->z_steps) return fail(BTS_ERR_INVALID_ARGUMENT, "n_coarse < 1 or z_steps null");
  if ((out->rgb || out->invalid || out->rgb_samps || out->invalid_mass) && scene->Vc < 1)
    return fail(BTS_ERR_INVALID_ARGUMENT, "colour outputs requested with Vc=0");
  if (out->depth_z && !gen) return fail(BTS_ERR_INVALID_ARGUMENT, "depth_z needs in-kernel ray generation (bts_render_views_fwd)");
  if (n * R == 0) return BTS_OK;
  bts::KernelParams p;
  memset(&p, 0, sizeof(p));
  p.scene = *scene; p.mlp = *mlp; p.cfg = *cfg; p.out = *out;
  p.rays = rays; p.u = u_or_null; p.n_items = n * R; p.per_batch = R; p.seed = seed;
  if (gen) p.gen = *gen;
  p.p_pad = (mlp->P + 3) & ~3;
  cudaError_t e;
  if (out->scratch && out->scratch_ints >= BTS_BAND_SCRATCH_INTS(n * R)) p.work = out->scratch;
  if (bts::tc_supported(p) && bts::band_supported(p)) {
    // band kernel (feature fetch as tcgen05 MMAs over a TMA-staged band of G16); the 8-ray groups whose footprint does
    // not fit are listed in `work` and rendered by the gather kernel right after it, same outputs
    g_kernel = "band";
    e = bts::launch_band(p, reinterpret_cast<cudaStream_t>(stream));
    if (e == cudaSuccess) {
      p.work_mode = 1;
      e = bts::launch_tc(p, false, reinterpret_cast<cudaStream_t>(stream));
    }
  } else if (cfg->precision != 2 && bts::tc_supported(p)) {
    g_kernel = "tc";
    p.work = nullptr;
    e = bts::launch_tc(p, false, reinterpret_cast<cudaStream_t>(stream));
  } else {
    if (cfg->precision == 1) return fail(BTS_ERR_UNSUPPORTED, "fast tier needs the tensor-core kernel (Hd=64, n_blocks=0, 6-octave code, K%%32==0)");
    g_kernel = "general";
    e = bts::launch_general(p, false, reinterpret_cast<cudaStream_t>(stream));
  }
  if (e != cudaSuccess) return fail(BTS_ERR_CUDA, "%s: %s", who, cudaGetErrorString(e));
  return BTS_OK;
}

int bts_render_fwd(const BtsScene* scene, const BtsMlp* mlp, const BtsRenderCfg* cfg, const float* rays, int64_t n,
                   int64_t R, const float* u_or_null, uint64_t seed, const BtsRenderOut* out, bts_stream_t stream) {
  if (!rays) return fail(BTS_ERR_INVALID_ARGUMENT, "bts_render_fwd: rays is null");
  return render_fwd_common("bts_render_fwd", scene, mlp, cfg, rays, nullptr, n, R, u_or_null, seed, out, stream);
}

int bts_render_views_fwd(const BtsScene* scene, const BtsMlp* mlp, const BtsRenderCfg* cfg, const BtsRayGen* gen,
                         int64_t n, const float* u_or_null, uint64_t seed, const BtsRenderOut* out, bts_stream_t stream) {
  if (!gen || !gen->poses_c2w || !gen->K) return fail(BTS_ERR_INVALID_ARGUMENT, "bts_render_views_fwd: gen / poses / K is null");
  if (gen->v < 1 || gen->H < 1 || gen->W < 1) return fail(BTS_ERR_INVALID_ARGUMENT, "bts_render_views_fwd: bad view sizes");
  return render_fwd_common("bts_render_views_fwd", scene, mlp, cfg, nullptr, gen, n, (int64_t)gen->v * gen->H * gen->W,
                           u_or_null, seed, out, stream);
}

int bts_render_bwd(const BtsScene* scene, const BtsMlp* mlp, const BtsRenderCfg* cfg, const float* rays, int64_t n,
                   int64_t R, const float* u_or_null, uint64_t seed, const float* saved_sigma_or_null,
                   const BtsRenderGrad* upstream, const BtsParamGrad* grads, bts_stream_t stream) {
  int rc = check_common(scene, mlp, cfg);
  if (rc != BTS_OK) return rc;
  if (!rays || !upstream || !grads) return fail(BTS_ERR_INVALID_ARGUMENT, "bts_render_bwd: rays/upstream/grads is null");
  if (n != scene->n) return fail(BTS_ERR_INVALID_ARGUMENT, "rays batch %lld != scene.n %d", (long long)n, scene->n);
  if (cfg->n_coarse < 1 || !cfg->z_steps) return fail(BTS_ERR_INVALID_ARGUMENT, "n_coarse < 1 or z_steps null");
  if (!(mlp->n_blocks == 0 && mlp->Hd <= 128) && !(mlp->n_blocks == 1 && mlp->Hd == 32))
    return fail(BTS_ERR_UNSUPPORTED, "backward implements n_blocks=0 with d_hidden<=128, or n_blocks=1 with d_hidden=32 (got %d, %d)",
                mlp->n_blocks, mlp->Hd);
  if (mlp->n_blocks == 1 && !(grads->d_fc0_w && grads->d_fc0_b && grads->d_fc1_w && grads->d_fc1_b))
    return fail(BTS_ERR_INVALID_ARGUMENT, "block gradients (d_fc0_*, d_fc1_*) are null");
  if (mlp->P > 40) return fail(BTS_ERR_UNSUPPORTED, "backward implements a code width <= 40 (got %d)", mlp->P);
  if (cfg->n_coarse > 256) return fail(BTS_ERR_UNSUPPORTED, "backward implements K <= 256 (got %d)", cfg->n_coarse);
  if (upstream->rgb && scene->Vc < 1) return fail(BTS_ERR_INVALID_ARGUMENT, "rgb gradient with Vc=0");
  if (!grads->d_G) return fail(BTS_ERR_INVALID_ARGUMENT, "grads.d_G is null");
  if (n * R == 0) return BTS_OK;
  bts::BwdParams bp;
  memset(&bp, 0, sizeof(bp));
  bp.k.scene = *scene; bp.k.mlp = *mlp; bp.k.cfg = *cfg;
  bp.k.rays = rays; bp.k.u = u_or_null; bp.k.n_items = n * R; bp.k.per_batch = R; bp.k.seed = seed;
  bp.k.p_pad = (mlp->P + 3) & ~3;
  bp.saved_sigma = saved_sigma_or_null;
  bp.g_rgb = upstream->rgb; bp.g_depth = upstream->depth; bp.g_weights = upstream->weights; bp.g_alphas = upstream->alphas;
  bp.d_G = grads->d_G; bp.d_g_empty = grads->d_g_empty; bp.d_w_pe = grads->d_w_pe; bp.d_w_out = grads->d_w_out;
  bp.d_b_out = grads->d_b_out;
  bp.d_fc0_w = grads->d_fc0_w; bp.d_fc0_b = grads->d_fc0_b; bp.d_fc1_w = grads->d_fc1_w; bp.d_fc1_b = grads->d_fc1_b;
  bool used_tc = false;
  const cudaError_t e = bts::launch_bwd(bp, reinterpret_cast<cudaStream_t>(stream), &used_tc);
  g_bwd_kernel = used_tc ? "tc" : "general";
  if (e != cudaSuccess) return fail(BTS_ERR_CUDA, "bts_render_bwd: %s", cudaGetErrorString(e));
  return BTS_OK;
}

int bts_query_field(const BtsScene* scene, const BtsMlp* mlp, const BtsRenderCfg* cfg, const float* xyz, int64_t n,
                    int64_t P, float* sigma, float* invalid_f, float* rgb_or_null, float* invalid_or_null,
                    bts_stream_t stream) {
  int rc = check_common(scene, mlp, cfg);
  if (rc != BTS_OK) return rc;
  if (!xyz || !sigma || !invalid_f) return fail(BTS_ERR_INVALID_ARGUMENT, "bts_query_field: null pointer");
  if (n != scene->n) return fail(BTS_ERR_INVALID_ARGUMENT, "xyz batch %lld != scene.n %d", (long long)n, scene->n);
  if ((rgb_or_null || invalid_or_null) && scene->Vc < 1) return fail(BTS_ERR_INVALID_ARGUMENT, "colour outputs requested with Vc=0");
  if (n * P == 0) return BTS_OK;
  bts::KernelParams p;
  memset(&p, 0, sizeof(p));
  p.scene = *scene; p.mlp = *mlp; p.cfg = *cfg;
  p.xyz = xyz; p.q_sigma = sigma; p.q_invalid_f = invalid_f; p.q_rgb = rgb_or_null; p.q_invalid = invalid_or_null;
  p.n_items = n * P; p.per_batch = P;
  p.p_pad = (mlp->P + 3) & ~3;
  cudaError_t e;
  if (cfg->precision != 2 && bts::tc_supported(p)) {
    g_kernel = "tc";
    e = bts::launch_tc(p, true, reinterpret_cast<cudaStream_t>(stream));
  } else {
    g_kernel = "general";
    e = bts::launch_general(p, true, reinterpret_cast<cudaStream_t>(stream));
  }
  if (e != cudaSuccess) return fail(BTS_ERR_CUDA, "bts_query_field: %s", cudaGetErrorString(e));
  return BTS_OK;
}

int bts_fold_features_bwd(const float* d_G, const float* feat_nchw, int32_t n, int32_t C, int32_t H, int32_t W,
                          const float* w_in, int32_t ld, int32_t Hd, float* d_feat_or_null, float* d_w_f, float* d_b_in,
                          bts_stream_t stream) {
  if (!d_G || !feat_nchw || !w_in || !d_w_f || !d_b_in) return fail(BTS_ERR_INVALID_ARGUMENT, "bts_fold_features_bwd: null pointer");
  if (n < 1 || C < 1 || H < 1 || W < 1 || Hd < 1 || ld < C) return fail(BTS_ERR_INVALID_ARGUMENT, "bts_fold_features_bwd: bad sizes");
  const int64_t HW = (int64_t)H * W;
  cudaError_t e;
  if (bts::fold_bwd_tc_supported(d_G, feat_nchw, d_feat_or_null, C, HW, Hd))      // TMA-fed tcgen05 kernel
    e = bts::launch_fold_bwd_tc(d_G, feat_nchw, n, C, HW, w_in, ld, Hd, d_feat_or_null, d_w_f, d_b_in, nullptr, reinterpret_cast<cudaStream_t>(stream));
  else                                                                               // any other shape: CUDA-core kernel
    e = bts::launch_fold_bwd_simt(d_G, feat_nchw, n, C, HW, w_in, ld, Hd, d_feat_or_null, d_w_f, d_b_in, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(BTS_ERR_CUDA, "bts_fold_features_bwd: %s", cudaGetErrorString(e));
  return BTS_OK;
}

static int patch_loss_check(const char* who, const float* pred, const float* gt, int64_t P, int32_t h, int32_t w,
                            int32_t nv, int32_t c) {
  if (!pred || !gt) return fail(BTS_ERR_INVALID_ARGUMENT, "%s: null pred / gt", who);
  if (P < 0) return fail(BTS_ERR_INVALID_ARGUMENT, "%s: negative patch count", who);
  if (!bts::patch_loss_supported(h, w, nv, c))
    return fail(BTS_ERR_UNSUPPORTED, "%s: patch %dx%d, %d views, %d channels (limits: side <= 16, views <= 16, channels <= 4)", who, h, w, nv, c);
  return BTS_OK;
}

int bts_patch_loss_fwd(const float* pred, const float* gt, const float* invalid_ray_or_null,
                       const float* invalid_mass_or_null, float mass_thresh, int64_t P, int32_t h, int32_t w, int32_t nv,
                       int32_t c, float* per_ray, int32_t* argmin_or_null, float* errors_or_null, bts_stream_t stream) {
  int rc = patch_loss_check("bts_patch_loss_fwd", pred, gt, P, h, w, nv, c);
  if (rc != BTS_OK) return rc;
  if (!per_ray) return fail(BTS_ERR_INVALID_ARGUMENT, "bts_patch_loss_fwd: per_ray is null");
  if (P == 0) return BTS_OK;
  bts::PatchLossParams p;
  memset(&p, 0, sizeof(p));
  p.pred = pred; p.gt = gt; p.invalid_ray = invalid_ray_or_null; p.invalid_mass = invalid_mass_or_null;
  p.mass_thresh = mass_thresh; p.P = P; p.h = h; p.w = w; p.nv = nv; p.c = c;
  p.per_ray = per_ray; p.argmin = argmin_or_null; p.errors = errors_or_null;
  const cudaError_t e = bts::launch_patch_loss(p, false, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(BTS_ERR_CUDA, "bts_patch_loss_fwd: %s", cudaGetErrorString(e));
  return BTS_OK;
}

int bts_patch_loss_bwd(const float* pred, const float* gt, const float* invalid_ray_or_null,
                       const float* invalid_mass_or_null, float mass_thresh, int64_t P, int32_t h, int32_t w, int32_t nv,
                       int32_t c, const int32_t* argmin, const float* g_per_ray, float* d_pred, bts_stream_t stream) {
  int rc = patch_loss_check("bts_patch_loss_bwd", pred, gt, P, h, w, nv, c);
  if (rc != BTS_OK) return rc;
  if (!argmin || !g_per_ray || !d_pred) return fail(BTS_ERR_INVALID_ARGUMENT, "bts_patch_loss_bwd: null argmin / g_per_ray / d_pred");
  if (P == 0) return BTS_OK;
  bts::PatchLossParams p;
  memset(&p, 0, sizeof(p));
  p.pred = pred; p.gt = gt; p.invalid_ray = invalid_ray_or_null; p.invalid_mass = invalid_mass_or_null;
  p.mass_thresh = mass_thresh; p.P = P; p.h = h; p.w = w; p.nv = nv; p.c = c;
  p.argmin = const_cast<int32_t*>(argmin); p.g_per_ray = g_per_ray; p.d_pred = d_pred;
  const cudaError_t e = bts::launch_patch_loss(p, true, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(BTS_ERR_CUDA, "bts_patch_loss_bwd: %s", cudaGetErrorString(e));
  return BTS_OK;
}

int bts_patch_smooth_fwd(const float* depth, const float* gt, const float* invalid_ray_or_null,
                         const float* invalid_mass_or_null, float mass_thresh, int64_t P, int32_t h, int32_t w, int32_t nv,
                         int32_t c, float* per_ray, bts_stream_t stream) {
  int rc = patch_loss_check("bts_patch_smooth_fwd", depth, gt, P, h, w, nv, c);
  if (rc != BTS_OK) return rc;
  if (!per_ray) return fail(BTS_ERR_INVALID_ARGUMENT, "bts_patch_smooth_fwd: per_ray is null");
  if (P == 0) return BTS_OK;
  bts::PatchLossParams p;
  memset(&p, 0, sizeof(p));
  p.pred = depth; p.gt = gt; p.invalid_ray = invalid_ray_or_null; p.invalid_mass = invalid_mass_or_null;
  p.mass_thresh = mass_thresh; p.P = P; p.h = h; p.w = w; p.nv = nv; p.c = c; p.per_ray = per_ray;
  const cudaError_t e = bts::launch_patch_smooth(p, false, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(BTS_ERR_CUDA, "bts_patch_smooth_fwd: %s", cudaGetErrorString(e));
  return BTS_OK;
}

int bts_patch_smooth_bwd(const float* depth, const float* gt, const float* invalid_ray_or_null,
                         const float* invalid_mass_or_null, float mass_thresh, int64_t P, int32_t h, int32_t w, int32_t nv,
                         int32_t c, const float* g_per_ray, float* d_depth, bts_stream_t stream) {
  int rc = patch_loss_check("bts_patch_smooth_bwd", depth, gt, P, h, w, nv, c);
  if (rc != BTS_OK) return rc;
  if (!g_per_ray || !d_depth) return fail(BTS_ERR_INVALID_ARGUMENT, "bts_patch_smooth_bwd: null g_per_ray / d_depth");
  if (P == 0) return BTS_OK;
  bts::PatchLossParams p;
  memset(&p, 0, sizeof(p));
  p.pred = depth; p.gt = gt; p.invalid_ray = invalid_ray_or_null; p.invalid_mass = invalid_mass_or_null;
  p.mass_thresh = mass_thresh; p.P = P; p.h = h; p.w = w; p.nv = nv; p.c = c; p.g_per_ray = g_per_ray; p.d_pred = d_depth;
  const cudaError_t e = bts::launch_patch_smooth(p, true, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(BTS_ERR_CUDA, "bts_patch_smooth_bwd: %s", cudaGetErrorString(e));
  return BTS_OK;
}

}  // extern "C"
