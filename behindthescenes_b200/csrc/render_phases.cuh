// Per-sample phases shared by the render kernel families (lane = sample of one ray chunk):
// alpha compositing with a transmittance scan, colour-view sampling, per-ray finish.
#pragma once
#include "common.cuh"

namespace bts {

// Alpha compositing of one 32-sample chunk (models/common/render/nerf.py:283-299).  Returns this sample's weight,
// updates the ray's running transmittance / depth / weight sum, writes the optional per-sample outputs.
__device__ __forceinline__ float composite_chunk(const KernelParams& p, int lane, int k, int K, bool valid, float z,
                                                 float z_next, float sigma, int64_t sample_id, float& T_carry,
                                                 float& depth_part, float& wsum_part) {
  const float delta = (k + 1 < K) ? __fsub_rn(z_next, z) : 1e10f;              // :225-228
  float alpha = __fsub_rn(1.0f, expf(-(fabsf(delta) * fmaxf(sigma, 0.0f))));    // :283
  if (p.cfg.hard_alpha_cap && k == K - 1) alpha = 1.0f;                         // :285-286
  const float tr = valid ? __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f) : 1.0f;    // :290-292
  float incl = tr;                                                              // inclusive product scan over lanes
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl *= up;
  }
  float excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) excl = 1.0f;
  const float Tk = T_carry * excl;
  T_carry *= __shfl_sync(0xffffffffu, incl, 31);
  const float wgt = valid ? alpha * Tk : 0.0f;                                  // :294
  depth_part = fmaf(wgt, z, depth_part);                                        // :299
  wsum_part += wgt;
  if (valid) {
    if (p.out.weights) p.out.weights[sample_id] = wgt;
    if (p.out.alphas) p.out.alphas[sample_id] = alpha;
    if (p.out.z_samps) p.out.z_samps[sample_id] = z;
    if (p.out.sigma) p.out.sigma[sample_id] = sigma;
  }
  return wgt;
}

// Colour views (models/bts/model/models_bts.py:218-264, :330-334): project the sample into every render view, fetch
// RGB bilinearly (border clamp), emit invalid = invalid_colour | invalid_feature, reduce weight * colour over the warp
// into w_rgb[view*4 + c] (lane 0 accumulates); w_rgb[kMaxColorViews*4 + view] accumulates the invalid mass.
template <bool kPoints>
__device__ __forceinline__ void colour_views(const KernelParams& p, int lane, int64_t bl, float X, float Y, float Z,
                                             bool valid, const Proj& pf, float wgt, int64_t sample_id, float* w_rgb) {
  const bool inv_f = pf.invalid;
  const bool want = kPoints ? (p.q_rgb != nullptr || p.q_invalid != nullptr)
                            : (p.out.rgb != nullptr || p.out.invalid != nullptr || p.out.rgb_samps != nullptr ||
                               p.out.invalid_mass != nullptr);
  if (!want) return;
  const int Vc = p.scene.Vc, Cc = p.scene.Cc, Hc = p.scene.Hc, Wc = p.scene.Wc;
  for (int v = 0; v < Vc; ++v) {
    Proj pc;
    if ((p.scene.same_view_mask >> v) & 1) {
      pc = pf;                                     // this colour view is the encoder view: same projection, same flags
    } else {
      View vc;
      load_view(vc, p.scene.K_c + (bl * Vc + v) * 9, p.scene.w2c_c + (bl * Vc + v) * 16);
      pc = project(vc, X, Y, Z);
    }
    const Taps t = bilinear_taps(pc.u, pc.v, Hc, Wc);
    const float4* base = reinterpret_cast<const float4*>(p.scene.img4) + ((bl * Vc + v) * Hc + t.y0) * (int64_t)Wc + t.x0;
    float col[4] = {0, 0, 0, 0};
    if (valid) {
      const float4 c00 = __ldg(base), c01 = __ldg(base + t.dx);
      const float4 c10 = __ldg(base + t.dy * Wc), c11 = __ldg(base + t.dy * Wc + t.dx);
      const float w00 = t.wx0 * t.wy0, w01 = t.wx1 * t.wy0, w10 = t.wx0 * t.wy1, w11 = t.wx1 * t.wy1;
      col[0] = fmaf(c11.x, w11, fmaf(c10.x, w10, fmaf(c01.x, w01, c00.x * w00)));
      col[1] = fmaf(c11.y, w11, fmaf(c10.y, w10, fmaf(c01.y, w01, c00.y * w00)));
      col[2] = fmaf(c11.z, w11, fmaf(c10.z, w10, fmaf(c01.z, w01, c00.z * w00)));
      col[3] = fmaf(c11.w, w11, fmaf(c10.w, w10, fmaf(c01.w, w01, c00.w * w00)));
    }
    const float inv = (pc.invalid | inv_f) ? 1.0f : 0.0f;   // :333-334
    if (!kPoints) {
      if (valid) {
        if (p.out.invalid) p.out.invalid[sample_id * Vc + v] = inv;
        if (p.out.rgb_samps)
          for (int c = 0; c < Cc; ++c) p.out.rgb_samps[(sample_id * Vc + v) * Cc + c] = col[c];
      }
      if (p.out.rgb || p.out.invalid_mass) {
        // five interleaved butterflies (independent chains): weight * (r, g, b, a) and weight * invalid (loss.py:107)
        float r0 = wgt * col[0], r1 = wgt * col[1], r2 = wgt * col[2], r3 = wgt * col[3], r4 = wgt * inv;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          r0 += __shfl_xor_sync(0xffffffffu, r0, o);   // nerf.py:298
          r1 += __shfl_xor_sync(0xffffffffu, r1, o);
          r2 += __shfl_xor_sync(0xffffffffu, r2, o);
          if (Cc > 3) r3 += __shfl_xor_sync(0xffffffffu, r3, o);
          if (p.out.invalid_mass) r4 += __shfl_xor_sync(0xffffffffu, r4, o);
        }
        if (lane == 0) {
          if (p.out.rgb) {
            w_rgb[v * 4 + 0] += r0; w_rgb[v * 4 + 1] += r1; w_rgb[v * 4 + 2] += r2;
            if (Cc > 3) w_rgb[v * 4 + 3] += r3;
          }
          if (p.out.invalid_mass) w_rgb[kMaxColorViews * 4 + v] += r4;
        }
      }
    } else if (valid) {
      if (p.q_invalid) p.q_invalid[sample_id * Vc + v] = inv;
      if (p.q_rgb)
        for (int c = 0; c < Cc; ++c) p.q_rgb[(sample_id * Vc + v) * Cc + c] = col[c];
    }
  }
}

// Per-ray outputs after the last chunk: depth, rgb (+ white background, nerf.py:301-304).
__device__ __forceinline__ void finish_ray(const KernelParams& p, int lane, int64_t ray, float depth_part, float wsum_part,
                                           const float* w_rgb, float zfac) {
  const int Vc = p.scene.Vc, Cc = p.scene.Cc;
  const float depth = warp_sum(depth_part);
  const float wsum = warp_sum(wsum_part);
  if (lane == 0 && p.out.depth) p.out.depth[ray] = depth;
  if (lane == 0 && p.out.depth_z) p.out.depth_z[ray] = depth * zfac;
  __syncwarp();
  if (p.out.invalid_mass && lane < Vc) p.out.invalid_mass[ray * Vc + lane] = w_rgb[kMaxColorViews * 4 + lane];
  if (p.out.rgb && lane < Vc * Cc) {
    float r = w_rgb[(lane / Cc) * 4 + (lane % Cc)];
    if (p.cfg.white_bkgd) r = r + 1.0f - wsum;
    p.out.rgb[ray * Vc * Cc + lane] = r;
  }
  __syncwarp();
}

}  // namespace bts
