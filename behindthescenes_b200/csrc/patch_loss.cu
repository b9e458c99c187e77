// Patch photometric loss: 0.85 * SSIM-dissimilarity + 0.15 * L1 per ray, minimum over the render views, invalid rays
// masked -- forward and backward, one CTA per patch (thread = pixel), the 3x3 gaussian SSIM statistics of every channel
// and view taken from one zero-haloed shared-memory copy of the patch.
//
// Reference: compute_errors_l1ssim (models/bts/model/loss.py:10-18) + SSIM(pad_reflection=False, gaussian_average=True,
// comp_mode=True) (models/common/model/layers.py:79-150) + the reduction of ReconstructionLoss.__call__
// (loss.py:151-167, invalid policy "weight_guided" :105-108), which the reference runs as ~40 elementwise / conv2d
// launches over (n*pc*nv, c, h, w) temporaries.
#include "common.cuh"
#include "patch_loss.cuh"

namespace bts {

constexpr int kLossMaxSide = 16;     // patch side (threads = h * w <= 256)
constexpr int kLossMaxC = 4;
constexpr float kG0 = 0.0947f, kG1 = 0.1183f, kG2 = 0.1478f;   // corner / edge / centre of layers.py:82-85
constexpr float kC1 = 0.01f * 0.01f, kC2 = 0.03f * 0.03f;      // layers.py:118-119

struct Stats {
  float mu_x, mu_y, mxx, myy, mxy;
};

// 3x3 gaussian statistics around (y, x) of the zero-haloed planes xs / ys (row stride ws = w + 2).
__device__ __forceinline__ Stats gauss_stats(const float* xs, const float* ys, int y, int x, int ws) {
  Stats s = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const float g = (dy == 1 && dx == 1) ? kG2 : ((dy == 1 || dx == 1) ? kG1 : kG0);
      const float a = xs[(y + dy) * ws + x + dx], b = ys[(y + dy) * ws + x + dx];
      s.mu_x = fmaf(g, a, s.mu_x);
      s.mu_y = fmaf(g, b, s.mu_y);
      s.mxx = fmaf(g, a * a, s.mxx);
      s.myy = fmaf(g, b * b, s.myy);
      s.mxy = fmaf(g, a * b, s.mxy);
    }
  }
  return s;
}

struct SsimTerms {
  float n1, n2, d1, d2, S;
};

__device__ __forceinline__ SsimTerms ssim_terms(const Stats& s) {
  SsimTerms t;
  const float sx = s.mxx - s.mu_x * s.mu_x, sy = s.myy - s.mu_y * s.mu_y, sxy = s.mxy - s.mu_x * s.mu_y;   // layers.py:132-134
  t.n1 = 2.f * s.mu_x * s.mu_y + kC1;
  t.n2 = 2.f * sxy + kC2;
  t.d1 = s.mu_x * s.mu_x + s.mu_y * s.mu_y + kC1;
  t.d2 = sx + sy + kC2;
  t.S = __fdiv_rn(t.n1 * t.n2, t.d1 * t.d2);                                                               // :136-141
  return t;
}

__device__ __forceinline__ bool ray_invalid(const PatchLossParams& p, int64_t ray) {
  if (p.invalid_ray) return p.invalid_ray[ray] > 0.5f;
  if (p.invalid_mass) {
    bool all = true;
    for (int v = 0; v < p.nv; ++v) all = all && (p.invalid_mass[ray * p.nv + v] > p.mass_thresh);
    return all;
  }
  return false;
}

// shared layout: gt planes [c][hs*ws], pred planes [c][hs*ws], (backward) coefficient planes [3][c][hs*ws]
template <bool kBackward>
__global__ void __launch_bounds__(256) patch_loss_kernel(PatchLossParams p) {
  extern __shared__ float sm[];
  const int h = p.h, w = p.w, c = p.c, nv = p.nv;
  const int hs = h + 2, ws = w + 2, plane = hs * ws;
  float* ys = sm;
  float* xs = ys + c * plane;
  float* coef = xs + c * plane;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const bool active = tid < h * w;
  const int y = active ? tid / w : 0, x = active ? tid % w : 0;
  const float inv_c = 1.0f / (float)c;

  for (int64_t patch = blockIdx.x; patch < p.P; patch += gridDim.x) {
    const int64_t ray = patch * h * w + tid;
    __syncthreads();
    for (int i = tid; i < (kBackward ? 5 : 2) * c * plane; i += nthr) sm[i] = 0.f;     // zero halos (ZeroPad2d(1), layers.py:116)
    __syncthreads();
    if (active)
      for (int ch = 0; ch < c; ++ch) ys[ch * plane + (y + 1) * ws + x + 1] = p.gt[ray * c + ch];

    float best = 0.f;
    int best_v = 0;
    if (!kBackward) {
      for (int v = 0; v < nv; ++v) {
        __syncthreads();
        if (active)
          for (int ch = 0; ch < c; ++ch) xs[ch * plane + (y + 1) * ws + x + 1] = p.pred[(ray * nv + v) * c + ch];
        __syncthreads();
        if (active) {
          float ssim = 0.f, l1 = 0.f;
          for (int ch = 0; ch < c; ++ch) {
            const Stats s = gauss_stats(xs + ch * plane, ys + ch * plane, y, x, ws);
            const SsimTerms t = ssim_terms(s);
            ssim += fminf(fmaxf(1.f - t.S, 0.f), 1.f) * 0.5f;                                              // :145
            l1 += fabsf(xs[ch * plane + (y + 1) * ws + x + 1] - ys[ch * plane + (y + 1) * ws + x + 1]);
          }
          const float err = 0.85f * (ssim * inv_c) + 0.15f * (l1 * inv_c);                                  // loss.py:17
          if (p.errors) p.errors[ray * nv + v] = err;
          if (v == 0 || err < best) { best = err; best_v = v; }                                            // loss.py:155
        }
      }
      if (active) {
        const float keep = ray_invalid(p, ray) ? 0.f : 1.f;                                                 // loss.py:161
        p.per_ray[ray] = best * keep;
        if (p.argmin) p.argmin[ray] = best_v;
      }
    } else {
      const int my_v = active ? p.argmin[ray] : -1;
      const float up = active ? (ray_invalid(p, ray) ? 0.f : p.g_per_ray[ray]) : 0.f;
      for (int v = 0; v < nv; ++v) {
        __syncthreads();
        if (active)
          for (int ch = 0; ch < c; ++ch) xs[ch * plane + (y + 1) * ws + x + 1] = p.pred[(ray * nv + v) * c + ch];
        __syncthreads();
        // coefficient fields of this pixel's SSIM term: dL/dmu_x, dL/dm_xx, dL/dm_xy
        const float wp = (my_v == v) ? up : 0.f;
        if (active) {
          for (int ch = 0; ch < c; ++ch) {
            const Stats s = gauss_stats(xs + ch * plane, ys + ch * plane, y, x, ws);
            const SsimTerms t = ssim_terms(s);
            const float one_minus = 1.f - t.S;
            const float de_dS = (one_minus >= 0.f && one_minus <= 1.f) ? -0.5f : 0.f;     // clamp passes gradient on [0,1]
            const float k = wp * 0.85f * inv_c * de_dS;
            const float inv_d = __fdiv_rn(1.f, t.d1 * t.d2);
            const float dS_dmu = 2.f * s.mu_y * (t.n2 - t.n1) * inv_d - t.S * 2.f * s.mu_x * (__fdiv_rn(1.f, t.d1) - __fdiv_rn(1.f, t.d2));
            const float dS_dmxx = -__fdiv_rn(t.S, t.d2);
            const float dS_dmxy = 2.f * t.n1 * inv_d;
            const int o = ch * plane + (y + 1) * ws + x + 1;
            coef[o] = k * dS_dmu;
            coef[c * plane + o] = k * dS_dmxx;
            coef[2 * c * plane + o] = k * dS_dmxy;
          }
        }
        __syncthreads();
        if (active) {
          for (int ch = 0; ch < c; ++ch) {
            const float* A = coef + ch * plane;
            const float* B = A + c * plane;
            const float* Cq = B + c * plane;
            float ga = 0.f, gb = 0.f, gc = 0.f;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
                const float g = (dy == 1 && dx == 1) ? kG2 : ((dy == 1 || dx == 1) ? kG1 : kG0);
                const int o = (y + dy) * ws + x + dx;
                ga = fmaf(g, A[o], ga);
                gb = fmaf(g, B[o], gb);
                gc = fmaf(g, Cq[o], gc);
              }
            }
            const float xq = xs[ch * plane + (y + 1) * ws + x + 1], yq = ys[ch * plane + (y + 1) * ws + x + 1];
            const float diff = xq - yq;
            const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
            p.d_pred[(ray * nv + v) * c + ch] = ga + 2.f * xq * gb + yq * gc + wp * 0.15f * inv_c * sgn;
          }
        }
      }
    }
  }
}

// Edge-aware smoothness of the rendered depth (models/bts/model/loss.py:21-40, masked as in :247-256): per patch
//   d = 1 / clamp(depth, 1e-3, 80);  d /= mean_patch(d)
//   e(y, x) = |d(y, x) - d(y, x+1)| exp(-mean_c |I(y, x) - I(y, x+1)|) + |d(y, x) - d(y+1, x)| exp(-mean_c |I(y, x) - I(y+1, x)|)
// with the terms that would leave the patch dropped; thread = pixel, the patch mean and the backward's sum by block reduction.
__device__ __forceinline__ float block_sum(float v, float* red, int tid, int nthr) {
  v = warp_sum(v);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < (nthr >> 5); ++i) t += red[i];
  return t;
}

template <bool kBackward>
__global__ void __launch_bounds__(256) patch_smooth_kernel(PatchLossParams p) {
  extern __shared__ float sm[];
  const int h = p.h, w = p.w, c = p.c;
  const int n_px = h * w;
  float* s_d = sm;                 // normalised inverse depth
  float* s_i = s_d + n_px;         // image, [c][n_px]
  float* s_gx = s_i + c * n_px;    // backward: upstream * weight * sign of the x / y edge owned by each pixel
  float* s_gy = s_gx + n_px;
  float* red = s_gy + n_px;        // 8 floats
  const int tid = threadIdx.x, nthr = blockDim.x;
  const bool active = tid < n_px;
  const int y = active ? tid / w : 0, x = active ? tid % w : 0;
  const float inv_c = 1.0f / (float)c;

  for (int64_t patch = blockIdx.x; patch < p.P; patch += gridDim.x) {
    const int64_t ray = patch * n_px + tid;
    float depth = 1.f, inv = 0.f;
    if (active) {
      depth = p.pred[ray];                                                   // pred = depth (P, h, w) for this kernel
      inv = __fdiv_rn(1.f, fminf(fmaxf(depth, 1e-3f), 80.f));                // loss.py:24
      for (int ch = 0; ch < c; ++ch) s_i[ch * n_px + tid] = p.gt[ray * c + ch];
    }
    const float mean = block_sum(inv, red, tid, nthr) / (float)n_px;         // :25
    const float d = __fdiv_rn(inv, mean);
    if (active) s_d[tid] = d;
    __syncthreads();
    float wx = 0.f, wy = 0.f, ddx = 0.f, ddy = 0.f;
    if (active) {
      if (x + 1 < w) {
        float a = 0.f;
        for (int ch = 0; ch < c; ++ch) a += fabsf(s_i[ch * n_px + tid] - s_i[ch * n_px + tid + 1]);
        wx = expf(-(a * inv_c));                                             // :32, :35
        ddx = d - s_d[tid + 1];                                              // :29
      }
      if (y + 1 < h) {
        float a = 0.f;
        for (int ch = 0; ch < c; ++ch) a += fabsf(s_i[ch * n_px + tid] - s_i[ch * n_px + tid + w]);
        wy = expf(-(a * inv_c));                                             // :33, :36
        ddy = d - s_d[tid + w];                                              // :30
      }
    }
    const float keep = (active && ray_invalid(p, ray)) ? 0.f : 1.f;          // :252-254
    if (!kBackward) {
      if (active) p.per_ray[ray] = (fabsf(ddx) * wx + fabsf(ddy) * wy) * keep;   // :38
    } else {
      const float up = active ? p.g_per_ray[ray] * keep : 0.f;
      const float gx = up * wx * (ddx > 0.f ? 1.f : (ddx < 0.f ? -1.f : 0.f));
      const float gy = up * wy * (ddy > 0.f ? 1.f : (ddy < 0.f ? -1.f : 0.f));
      if (active) { s_gx[tid] = gx; s_gy[tid] = gy; }
      __syncthreads();
      float gd = gx + gy;                                                    // dE/dd of this pixel: its own edges ...
      if (active && x > 0) gd -= s_gx[tid - 1];                              // ... and the left / upper neighbour's
      if (active && y > 0) gd -= s_gy[tid - w];
      const float dot = block_sum(active ? gd * inv : 0.f, red, tid, nthr);
      if (active) {
        const float g_inv = __fdiv_rn(gd, mean) - __fdiv_rn(dot, mean * mean * (float)n_px);
        const bool pass = depth >= 1e-3f && depth <= 80.f;                   // clamp passes the gradient on [min, max]
        p.d_pred[ray] = pass ? -g_inv * inv * inv : 0.f;
      }
    }
    __syncthreads();
  }
}

cudaError_t launch_patch_smooth(const PatchLossParams& p, bool backward, cudaStream_t stream) {
  const int threads = ((p.h * p.w + 31) / 32) * 32;
  const size_t smem = (size_t)((3 + p.c) * p.h * p.w + 8) * sizeof(float);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)(p.P < (int64_t)sms * 16 ? p.P : (int64_t)sms * 16);
  if (backward)
    patch_smooth_kernel<true><<<grid, threads, smem, stream>>>(p);
  else
    patch_smooth_kernel<false><<<grid, threads, smem, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_patch_loss(const PatchLossParams& p, bool backward, cudaStream_t stream) {
  const int threads = ((p.h * p.w + 31) / 32) * 32;
  const size_t smem = (size_t)(backward ? 5 : 2) * p.c * (p.h + 2) * (p.w + 2) * sizeof(float);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t want = p.P;
  const int grid = (int)(want < (int64_t)sms * 16 ? want : (int64_t)sms * 16);
  if (backward)
    patch_loss_kernel<true><<<grid, threads, smem, stream>>>(p);
  else
    patch_loss_kernel<false><<<grid, threads, smem, stream>>>(p);
  return cudaGetLastError();
}

bool patch_loss_supported(int h, int w, int nv, int c) {
  return h >= 1 && w >= 1 && h <= kLossMaxSide && w <= kLossMaxSide && nv >= 1 && nv <= 16 && c >= 1 && c <= kLossMaxC;
}

}  // namespace bts
