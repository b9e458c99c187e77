// Encode-side glue kernels (run once per encode()/weight update, not per ray):
//   fold_features : G[b, p, :] = W_in[:, :C] * F[b, :, p] + b_in   -- layer-1 linearity: bilinear
//                   interpolation commutes with the feature part of lin_in, so the render kernels interpolate the
//                   Hd-channel map G instead of the C-channel map F and only contract the 39-wide positional code
//                   per sample.  fp32, k-sequential FMA accumulation.
//   pack_images   : colour views NCHW -> channels-last padded to 4 (one 16-byte load per bilinear tap).
// Reference tensors: grid_f_features / grid_c_imgs as cached by BTSNet.encode (models/bts/model/models_bts.py:126-136);
// weight = mlp_coarse.lin_in.weight[:, :C] (models/common/model/resnetfc.py:88, input order models_bts.py:184).
#include <cstdlib>

#include "common.cuh"

namespace bts {

constexpr int kFoldPix = 128;   // pixels per CTA
constexpr int kFoldCk = 32;     // channels per shared-memory stage

template <int UN>  // units per thread; Hd = 8 * UN
__global__ void __launch_bounds__(256) fold_features_kernel(const float* __restrict__ feat, int C, int64_t HW,
                                                            const float* __restrict__ w_in, int ld, const float* __restrict__ b_in, float* __restrict__ G) {
  constexpr int Hd = 8 * UN;
  __shared__ __align__(16) float Fs[kFoldCk][kFoldPix];
  __shared__ __align__(16) float Ws[kFoldCk][Hd];
  const int b = blockIdx.y;
  const int64_t p0 = (int64_t)blockIdx.x * kFoldPix;
  const int tp = threadIdx.x & 31, tn = threadIdx.x >> 5;
  float acc[4][UN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int n = 0; n < UN; ++n) acc[i][n] = __ldg(b_in + tn * UN + n);   // lin_in bias rides along (resnetfc.py:147)
  const float* fb = feat + (size_t)b * C * HW;
  for (int c0 = 0; c0 < C; c0 += kFoldCk) {
    for (int i = threadIdx.x; i < kFoldCk * kFoldPix; i += 256) {
      const int c = i / kFoldPix, px = i % kFoldPix;
      const int64_t pp = p0 + px;
      Fs[c][px] = (c0 + c < C && pp < HW) ? __ldg(fb + (size_t)(c0 + c) * HW + pp) : 0.0f;
    }
    for (int i = threadIdx.x; i < kFoldCk * Hd; i += 256) {
      const int n = i / kFoldCk, c = i % kFoldCk;
      Ws[c][n] = (c0 + c < C) ? __ldg(w_in + (size_t)n * ld + c0 + c) : 0.0f;
    }
    __syncthreads();
#pragma unroll 4
    for (int c = 0; c < kFoldCk; ++c) {
      const float4 f = *reinterpret_cast<const float4*>(&Fs[c][tp * 4]);
      float w[UN];
#pragma unroll
      for (int n = 0; n < UN; n += 4) {
        const float4 q = *reinterpret_cast<const float4*>(&Ws[c][tn * UN + n]);
        w[n] = q.x; w[n + 1] = q.y; w[n + 2] = q.z; w[n + 3] = q.w;
      }
      const float fv[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int n = 0; n < UN; ++n) acc[i][n] = fmaf(w[n], fv[i], acc[i][n]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t pp = p0 + tp * 4 + i;
    if (pp < HW) {
      float* o = G + ((size_t)b * HW + pp) * Hd + tn * UN;
#pragma unroll
      for (int n = 0; n < UN; n += 4) *reinterpret_cast<float4*>(o + n) = make_float4(acc[i][n], acc[i][n + 1], acc[i][n + 2], acc[i][n + 3]);
    }
  }
}

__global__ void fold_empty_kernel(const float* __restrict__ empty, int C, const float* __restrict__ w_in, int ld, int Hd,
                                  const float* __restrict__ b_in, float* __restrict__ g_empty) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= Hd) return;
  float a = __ldg(b_in + n);
  for (int c = 0; c < C; ++c) a = fmaf(__ldg(w_in + (size_t)n * ld + c), __ldg(empty + c), a);
  g_empty[n] = a;
}

__global__ void pack_images_kernel(const float* __restrict__ img, int Cc, int64_t HW, int64_t total, float4* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t v = i / HW, px = i % HW;
  const float* src = img + (size_t)v * Cc * HW + px;
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
  o.x = __ldg(src);
  if (Cc > 1) o.y = __ldg(src + HW);
  if (Cc > 2) o.z = __ldg(src + 2 * HW);
  if (Cc > 3) o.w = __ldg(src + 3 * HW);
  out[i] = o;
}

bool fold_tc_supported(const float* feat, int C, int64_t HW, int Hd);
cudaError_t launch_fold_tc(const float* feat, int n, int C, int64_t HW, const float* w_in, int ld, const float* b_in, int Hd,
                           float* G, unsigned char* tile_mask, cudaStream_t stream);

// ------------------------------------------------------------------------------------------------------------------
// Sparse fold (training path).  A training batch samples the folded map along the epipolar lines of a few 8x8 patches
// (models/bts/model/ray_sampler.py: PatchRaySampler): ~15-25 % of a frame's 64-pixel tiles.  mark_tiles_kernel runs the
// renderers' sampling + projection (the same device functions, so exactly the same taps) and flags the tiles a ray
// batch reads; the fold kernels then produce / differentiate only those.  Mask byte per 64 pixels of a frame:
// 0 = nothing reads it, 1 = needed, 2 = folded.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mark_tiles_kernel(const KernelParams p, unsigned char* __restrict__ mask, int mask_per_b) {
  const int lane = threadIdx.x & 31, K = p.cfg.n_coarse, H = p.scene.H, W = p.scene.W;
  const bool use_empty = p.cfg.learn_empty && p.scene.g_empty != nullptr;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t ray = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); ray < p.n_items; ray += warps) {
    const int64_t b = p.scene.n == 1 ? 0 : ray / p.per_batch;
    const Ray ry = fetch_ray(p, ray, lane);
    float near_, far_;
    depth_bounds(p, ry, near_, far_);
    View vf;
    load_view(vf, p.scene.K_f + b * 9, p.scene.w2c_f + b * 16);
    unsigned char* m = mask + b * mask_per_b;
    for (int k = lane; k < K; k += 32) {
      const float z = sample_depth(p, ray, k, near_, far_);
      const Proj pf = project(vf, __fadd_rn(ry.o[0], __fmul_rn(z, ry.d[0])), __fadd_rn(ry.o[1], __fmul_rn(z, ry.d[1])),
                              __fadd_rn(ry.o[2], __fmul_rn(z, ry.d[2])));
      if (use_empty && pf.invalid) continue;          // takes the learned empty row: no fetch
      const Taps t = bilinear_taps(pf.u, pf.v, H, W);
      const int p00 = t.y0 * W + t.x0, p10 = p00 + t.dy * W;
      const int tiles[4] = {p00 >> 6, (p00 + t.dx) >> 6, p10 >> 6, (p10 + t.dx) >> 6};
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (m[tiles[i]] == 0) m[tiles[i]] = 1;        // benign race: every writer stores 1
    }
  }
}

cudaError_t launch_mark_tiles(const KernelParams& p, unsigned char* mask, cudaStream_t stream) {
  const int64_t HW = (int64_t)p.scene.H * p.scene.W;
  const int64_t want = (p.n_items + 7) / 8;
  mark_tiles_kernel<<<(unsigned)(want < 1184 ? (want > 0 ? want : 1) : 1184), 256, 0, stream>>>(p, mask, (int)((HW + 63) / 64));
  return cudaGetLastError();
}

// Zero the rows of d_G (n, HW, Hd) that belong to marked tiles: the render backward accumulates into exactly those.
__global__ void __launch_bounds__(256) zero_tiles_kernel(float4* __restrict__ dG, int64_t HW, int Hd, int mask_per_b, int64_t n_tiles,
                                                         const unsigned char* __restrict__ mask) {
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    if (mask[t] == 0) continue;
    const int64_t b = t / mask_per_b, j = t % mask_per_b;
    const int64_t px = HW - j * 64 < 64 ? HW - j * 64 : 64;
    float4* dst = dG + ((b * HW + j * 64) * Hd) / 4;
    for (int i = threadIdx.x; i < px * Hd / 4; i += blockDim.x) dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

cudaError_t launch_zero_tiles(float* dG, int n, int64_t HW, int Hd, const unsigned char* mask, cudaStream_t stream) {
  const int mask_per_b = (int)((HW + 63) / 64);
  const int64_t n_tiles = (int64_t)mask_per_b * n;
  zero_tiles_kernel<<<(unsigned)(n_tiles < 4736 ? n_tiles : 4736), 256, 0, stream>>>(reinterpret_cast<float4*>(dG), HW, Hd, mask_per_b, n_tiles, mask);
  return cudaGetLastError();
}

cudaError_t launch_fold(const float* feat, int n, int C, int H, int W, const float* w_in, int ld, const float* b_in, int Hd,
                        const float* empty, float* G, float* g_empty, unsigned char* tile_mask, cudaStream_t stream) {
  const int64_t HW = (int64_t)H * W;
  // tensor-core path (TMA-fed tcgen05 GEMM, fold_tc.cu) whenever the frame qualifies; BTS_FOLD_SIMT=1 forces the
  // CUDA-core kernel (tests compare the two)
  static const bool force_simt = [] { const char* v = getenv("BTS_FOLD_SIMT"); return v && v[0] == '1'; }();   // read once
  if (!force_simt && fold_tc_supported(feat, C, HW, Hd)) {
    cudaError_t e = launch_fold_tc(feat, n, C, HW, w_in, ld, b_in, Hd, G, tile_mask, stream);
    if (e != cudaSuccess) return e;
    if (empty && g_empty) {
      fold_empty_kernel<<<(Hd + 127) / 128, 128, 0, stream>>>(empty, C, w_in, ld, Hd, b_in, g_empty);
      e = cudaGetLastError();
    }
    return e;
  }
  if (tile_mask) {    // the CUDA-core kernel folds the whole map: every tile is valid afterwards
    cudaError_t em = cudaMemsetAsync(tile_mask, 2, (size_t)((HW + 63) / 64) * n, stream);
    if (em != cudaSuccess) return em;
  }
  dim3 grid((unsigned)((HW + kFoldPix - 1) / kFoldPix), (unsigned)n);
  switch (Hd / 8) {
    case 4: fold_features_kernel<4><<<grid, 256, 0, stream>>>(feat, C, HW, w_in, ld, b_in, G); break;
    case 8: fold_features_kernel<8><<<grid, 256, 0, stream>>>(feat, C, HW, w_in, ld, b_in, G); break;
    case 12: fold_features_kernel<12><<<grid, 256, 0, stream>>>(feat, C, HW, w_in, ld, b_in, G); break;
    case 16: fold_features_kernel<16><<<grid, 256, 0, stream>>>(feat, C, HW, w_in, ld, b_in, G); break;
    default: return cudaErrorInvalidValue;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (empty && g_empty) {
    fold_empty_kernel<<<(Hd + 127) / 128, 128, 0, stream>>>(empty, C, w_in, ld, Hd, b_in, g_empty);
    e = cudaGetLastError();
  }
  return e;
}

// ------------------------------------------------------------------------------------------------------------------
// General CUDA-core backward of the fold (any C, d_hidden, frame size): the shapes fold_bwd_tc.cu does not take.
//   d_feat[b, c, p] = sum_k d_G[b, p, k] W_F[k, c]          (written)
//   d_W_F[k, c]    += sum_{b,p} d_G[b, p, k] F[b, c, p]     (accumulated),   d_b[k] += sum_{b,p} d_G[b, p, k]
// One CTA per tile of 64 pixels: d_G tile and W_F in shared memory; thread = pixel for d_feat (coalesced along pixels),
// thread = (k, c) pairs for the weight gradient (register partial sums over the tile, one atomic per pair and CTA).
// ------------------------------------------------------------------------------------------------------------------
constexpr int kBwdPix = 64;

__global__ void __launch_bounds__(256) fold_bwd_simt_kernel(const float* __restrict__ dG, const float* __restrict__ feat, int C, int64_t HW,
                                                            const float* __restrict__ w_in, int ld, int Hd, float* __restrict__ d_feat,
                                                            float* __restrict__ d_wf, float* __restrict__ d_b) {
  extern __shared__ float sm[];
  float* s_dg = sm;                         // [kBwdPix][Hd + 1]
  float* s_w = s_dg + kBwdPix * (Hd + 1);   // [Hd][C]
  float* s_f = s_w + Hd * C;                // [C][kBwdPix]
  const int tid = threadIdx.x, b = blockIdx.y;
  const int64_t p0 = (int64_t)blockIdx.x * kBwdPix;
  const int valid = (int)((HW - p0) < kBwdPix ? (HW - p0) : kBwdPix);
  for (int i = tid; i < Hd * C; i += 256) s_w[i] = __ldg(w_in + (size_t)(i / C) * ld + (i % C));
  for (int i = tid; i < kBwdPix * Hd; i += 256) {
    const int p = i / Hd, k = i % Hd;
    s_dg[p * (Hd + 1) + k] = p < valid ? __ldg(dG + ((size_t)b * HW + p0 + p) * Hd + k) : 0.0f;
  }
  for (int i = tid; i < C * kBwdPix; i += 256) {
    const int c = i / kBwdPix, p = i % kBwdPix;
    s_f[i] = p < valid ? __ldg(feat + ((size_t)b * C + c) * HW + p0 + p) : 0.0f;
  }
  __syncthreads();
  if (d_feat) {                             // thread = (pixel, channel quarter)
    const int p = tid % kBwdPix, cq = tid / kBwdPix;
    if (p < valid)
      for (int c = cq; c < C; c += 256 / kBwdPix) {
        float a = 0.0f;
        for (int k = 0; k < Hd; ++k) a = fmaf(s_dg[p * (Hd + 1) + k], s_w[k * C + c], a);
        d_feat[((size_t)b * C + c) * HW + p0 + p] = a;
      }
  }
  for (int i = tid; i < Hd * C; i += 256) { // thread = (k, c) pairs
    const int k = i / C, c = i % C;
    float a = 0.0f;
    for (int p = 0; p < kBwdPix; ++p) a = fmaf(s_dg[p * (Hd + 1) + k], s_f[c * kBwdPix + p], a);
    atomicAdd(d_wf + (size_t)k * C + c, a);
  }
  for (int k = tid; k < Hd; k += 256) {
    float a = 0.0f;
    for (int p = 0; p < kBwdPix; ++p) a += s_dg[p * (Hd + 1) + k];
    atomicAdd(d_b + k, a);
  }
}

cudaError_t launch_fold_bwd_simt(const float* dG, const float* feat, int n, int C, int64_t HW, const float* w_in, int ld, int Hd,
                                 float* d_feat, float* d_wf, float* d_b, cudaStream_t stream) {
  const size_t smem = (size_t)(kBwdPix * (Hd + 1) + Hd * C + C * kBwdPix) * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(fold_bwd_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid((unsigned)((HW + kBwdPix - 1) / kBwdPix), (unsigned)n);
  fold_bwd_simt_kernel<<<grid, 256, smem, stream>>>(dG, feat, C, HW, w_in, ld, Hd, d_feat, d_wf, d_b);
  return cudaGetLastError();
}

cudaError_t launch_pack_images(const float* img, int nv_total, int Cc, int Hc, int Wc, float* out, cudaStream_t stream) {
  const int64_t HW = (int64_t)Hc * Wc, total = HW * nv_total;
  if (total == 0) return cudaSuccess;
  pack_images_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(img, Cc, HW, total, reinterpret_cast<float4*>(out));
  return cudaGetLastError();
}

}  // namespace bts
