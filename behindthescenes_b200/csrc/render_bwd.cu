// Backward of the fused render (training path, BASELINE config 3): gradients of a scalar loss w.r.t. the folded
// feature map G, the learned empty row, the positional-code columns of lin_in, lin_out -- given the upstream
// gradients of the per-ray / per-sample outputs (rgb, depth, weights, alphas).
//
// Recompute-in-backward: nothing but the inputs is read; one warp owns one ray and makes two sweeps over its chunks.
//   sweep 1 (forward recompute) : depth samples, projection, code, G-interpolation + contraction -> sigma; alpha, T;
//                                  colours -> Wbar_k = dL/dw_k; per-sample scalars parked in shared memory
//   reverse recurrence          : R_j = Wbar_{j+1} a_{j+1} + t_{j+1} R_{j+1}  (division-free suffix term of the
//                                  transmittance product), dL/dalpha_j = Wbar_j T_j + g_alpha_j - T_j R_j,
//                                  dL/dlogit_j = dL/dalpha_j * delta_j (1 - a_j) * sigmoid(logit_j)
//   sweep 2 (MLP backward)      : recompute h per 8-sample block (lanes = hidden units), hbar = lbar * w_out * [h > 0];
//                                  d w_out, d b_out, d W_pe (register accumulators, flushed once per CTA),
//                                  d G (4 taps x Hd red.add per sample), d g_empty
// The dense part of the chain (G = W_F F + b  ->  dF, dW_F, db_in) is two plain GEMMs done by the caller.
//
// Reference forward being differentiated: models/common/render/nerf.py:283-304 (alpha, cumprod, weights, rgb, depth),
// models/bts/model/models_bts.py:138-216, :311-313 (features, softplus), resnetfc.py:132-184 (n_blocks = 0).
// The reference gets these gradients from torch autograd; tests compare against autograd through the oracle.
#include <cstdlib>

#include "common.cuh"
#include "render_phases.cuh"
#include "bwd_params.cuh"

namespace bts {

constexpr int kBwdWarps = 4;
constexpr int kBwdMaxK = 256;

// kBlock: one ResnetFC block (resnetfc.py:53-62) between lin_in and lin_out -- the RE10K configuration (d_hidden = 32,
// NU = 1): h1 = h0 + fc1(relu(fc0(relu(h0)))); lane n owns hidden unit n and row n of the weight gradients.
// This lane's NU consecutive hidden units of one row (one 8-byte access when NU == 2).
template <int NU>
__device__ __forceinline__ void load_units(const float* src, float (&dst)[NU]) {
  if constexpr (NU == 2) {
    const float2 v = *reinterpret_cast<const float2*>(src);
    dst[0] = v.x; dst[1] = v.y;
  } else {
#pragma unroll
    for (int n = 0; n < NU; ++n) dst[n] = src[n];
  }
}

// Bilinear interpolation of the folded feature rows for this lane's units (or the g_empty row for invalid samples).
template <int NU>
__device__ __forceinline__ void gather_units(const KernelParams& p, const float4& a, const float4& w, int flags, int lane,
                                             float (&out)[NU]) {
  constexpr int Hd = NU * 32;
#pragma unroll
  for (int n = 0; n < NU; ++n) out[n] = 0.0f;
  if (flags & 2) {
#pragma unroll
    for (int n = 0; n < NU; ++n) out[n] = __ldg(p.scene.g_empty + lane * NU + n);
  } else if (flags & 1) {
    const float* g00 = p.scene.G + (size_t)__float_as_int(a.x) * Hd + lane * NU;
    const float* g10 = g00 + (size_t)__float_as_int(a.z) * Hd;
    const size_t dxo = (size_t)__float_as_int(a.y) * Hd;
    const float w00 = w.x * w.z, w01 = w.y * w.z, w10 = w.x * w.w, w11 = w.y * w.w;
    if constexpr (NU == 2) {
      const float2 v00 = __ldg(reinterpret_cast<const float2*>(g00)), v01 = __ldg(reinterpret_cast<const float2*>(g00 + dxo));
      const float2 v10 = __ldg(reinterpret_cast<const float2*>(g10)), v11 = __ldg(reinterpret_cast<const float2*>(g10 + dxo));
      out[0] = fmaf(v11.x, w11, fmaf(v10.x, w10, fmaf(v01.x, w01, v00.x * w00)));
      out[1] = fmaf(v11.y, w11, fmaf(v10.y, w10, fmaf(v01.y, w01, v00.y * w00)));
    } else {
#pragma unroll
      for (int n = 0; n < NU; ++n)
        out[n] = fmaf(__ldg(g10 + dxo + n), w11, fmaf(__ldg(g10 + n), w10, fmaf(__ldg(g00 + dxo + n), w01, __ldg(g00 + n) * w00)));
    }
  }
}

template <int NU, bool kBlock>
__global__ void __launch_bounds__(kBwdWarps * 32, kBlock ? 2 : 3) render_bwd_kernel(const BwdParams bp) {
  static_assert(!kBlock || NU == 1, "block backward is implemented for d_hidden = 32");
  extern __shared__ __align__(16) float smem[];
  const KernelParams& p = bp.k;
  constexpr int Hd = NU * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int P = p.mlp.P, Ppad = p.p_pad;
  const int K = p.cfg.n_coarse;
  const int Kpad = (K + 31) & ~31;
  float* s_wpe = smem;                       // [Ppad][Hd]
  float* s_wout = s_wpe + Ppad * Hd;         // [Hd]
  // block weights: [fc0 | fc0^T | fc1 | fc1^T] (Hd x Hd each) + b0, b1
  float* s_w0 = s_wout + Hd;
  float* s_w0t = s_w0 + (kBlock ? Hd * Hd : 0);
  float* s_w1 = s_w0t + (kBlock ? Hd * Hd : 0);
  float* s_w1t = s_w1 + (kBlock ? Hd * Hd : 0);
  float* s_b0 = s_w1t + (kBlock ? Hd * Hd : 0);
  float* s_b1 = s_b0 + (kBlock ? Hd : 0);
  float* wbase = s_b1 + (kBlock ? Hd : 0) + warp * (Ppad * 32 + 32 * 8 + 32 + 5 * Kpad);
  float* w_pe = wbase;                       // [Ppad][32]
  float* w_taps = w_pe + Ppad * 32;          // [32][8]
  float* w_grgb = w_taps + 32 * 8;           // [Vc*Cc <= 32]
  float* sc_alpha = w_grgb + 32;             // [Kpad] each
  float* sc_T = sc_alpha + Kpad;
  float* sc_wbar = sc_T + Kpad;
  float* sc_fac = sc_wbar + Kpad;
  float* sc_lbar = sc_fac + Kpad;

  {
    const int ld = p.mlp.C + P;
    for (int i = threadIdx.x; i < Ppad * Hd; i += blockDim.x) {
      const int j = i / Hd, n = i % Hd;
      s_wpe[i] = j < P ? __ldg(p.mlp.w_in + (size_t)n * ld + p.mlp.C + j) : 0.0f;
    }
    for (int i = threadIdx.x; i < Hd; i += blockDim.x) s_wout[i] = __ldg(p.mlp.w_out + i);
    if constexpr (kBlock) {
      for (int i = threadIdx.x; i < Hd * Hd; i += blockDim.x) {
        const int n = i / Hd, k = i % Hd;
        const float a = __ldg(p.mlp.fc0_w[0] + i), b = __ldg(p.mlp.fc1_w[0] + i);
        s_w0[i] = a; s_w0t[k * Hd + n] = a; s_w1[i] = b; s_w1t[k * Hd + n] = b;
      }
      for (int i = threadIdx.x; i < Hd; i += blockDim.x) { s_b0[i] = __ldg(p.mlp.fc0_b[0] + i); s_b1[i] = __ldg(p.mlp.fc1_b[0] + i); }
    }
  }
  __syncthreads();

  const BtsRenderCfg& cfg = p.cfg;
  const int Vc = p.scene.Vc, Cc = p.scene.Cc;
  const int H = p.scene.H, W = p.scene.W, Hc = p.scene.Hc, Wc = p.scene.Wc;
  const float b_out = __ldg(p.mlp.b_out);
  const bool use_empty = cfg.learn_empty && p.scene.g_empty != nullptr;

  // persistent per-lane gradient accumulators (this lane's NU hidden units)
  float acc_wpe[NU][40];
  float acc_wout[NU], acc_gempty[NU], acc_bout = 0.0f;
  float acc_w0[kBlock ? 32 : 1], acc_w1[kBlock ? 32 : 1], acc_b0 = 0.0f, acc_b1 = 0.0f;   // row `lane` of d fc0 / d fc1
#pragma unroll
  for (int i = 0; i < (kBlock ? 32 : 1); ++i) { acc_w0[i] = 0.0f; acc_w1[i] = 0.0f; }
#pragma unroll
  for (int n = 0; n < NU; ++n) {
    acc_wout[n] = 0.0f; acc_gempty[n] = 0.0f;
#pragma unroll
    for (int j = 0; j < 40; ++j) acc_wpe[n][j] = 0.0f;
  }

  for (int64_t ray = (int64_t)blockIdx.x * kBwdWarps + warp; ray < p.n_items; ray += (int64_t)gridDim.x * kBwdWarps) {
    const int64_t b = p.scene.n == 1 ? 0 : ray / p.per_batch;
    const Ray ry = fetch_ray(p, ray, lane);
    const float ro0 = ry.o[0], ro1 = ry.o[1], ro2 = ry.o[2], rd0 = ry.d[0], rd1 = ry.d[1], rd2 = ry.d[2];
    float near_, far_;
    depth_bounds(p, ry, near_, far_);
    View vf;
    load_view(vf, p.scene.K_f + b * 9, p.scene.w2c_f + b * 16);
    if (lane < Vc * Cc) w_grgb[lane] = bp.g_rgb ? __ldg(bp.g_rgb + ray * Vc * Cc + lane) : 0.0f;
    const float g_depth = bp.g_depth ? __ldg(bp.g_depth + ray) : 0.0f;
    __syncwarp();
    float g_rgb_sum = 0.0f;
    if (cfg.white_bkgd)
      for (int i = 0; i < Vc * Cc; ++i) g_rgb_sum += w_grgb[i];

    // ======================= sweep 1: forward recompute, per-sample scalars =======================
    float T_carry = 1.0f;
    float z_ahead = lane < K ? sample_depth(p, ray, lane, near_, far_) : 0.0f;
    for (int c0 = 0; c0 < K; c0 += 32) {
      const int k = c0 + lane;
      const bool valid = k < K;
      const float z = z_ahead;
      z_ahead = (k + 32 < K) ? sample_depth(p, ray, k + 32, near_, far_) : 0.0f;
      float z_next = __shfl_down_sync(0xffffffffu, z, 1);
      const float z_first_of_next = __shfl_sync(0xffffffffu, z_ahead, 0);
      if (lane == 31) z_next = z_first_of_next;
      const float X = __fadd_rn(ro0, __fmul_rn(z, rd0)), Y = __fadd_rn(ro1, __fmul_rn(z, rd1)), Z = __fadd_rn(ro2, __fmul_rn(z, rd2));
      float logit = 0.0f;
      if (bp.saved_sigma == nullptr) {
      const Proj pf = project(vf, X, Y, Z);
        {
          positional_code(cfg, pf.u, pf.v, depth_code(cfg, pf), w_pe + lane, 32);
          for (int j = P; j < Ppad; ++j) w_pe[j * 32 + lane] = 0.0f;
          const Taps t = bilinear_taps(pf.u, pf.v, H, W);
          float4 a, w;
          a.x = __int_as_float((int)((b * H + t.y0) * (int64_t)W + t.x0));
          a.y = __int_as_float(t.dx);
          a.z = __int_as_float(t.dy * W);
          a.w = __int_as_float((valid ? 1 : 0) | ((use_empty && pf.invalid) ? 2 : 0));
          w.x = t.wx0; w.y = t.wx1; w.z = t.wy0; w.w = t.wy1;
          reinterpret_cast<float4*>(w_taps)[lane * 2 + 0] = a;
          reinterpret_cast<float4*>(w_taps)[lane * 2 + 1] = w;
        }
        __syncwarp();
        const int n_valid = min(32, K - c0);
        for (int s0 = 0; s0 < n_valid; s0 += 8) {
          float acc[8][NU];
#pragma unroll
          for (int ss = 0; ss < 8; ++ss) {
            const float4 a = reinterpret_cast<const float4*>(w_taps)[(s0 + ss) * 2 + 0];
            const float4 w = reinterpret_cast<const float4*>(w_taps)[(s0 + ss) * 2 + 1];
            const int flags = __float_as_int(a.w);
            gather_units<NU>(p, a, w, flags, lane, acc[ss]);
          }
          for (int j = 0; j < P; ++j) {
            float wv[NU];
            load_units<NU>(s_wpe + j * Hd + lane * NU, wv);
            const float4 c0v = *reinterpret_cast<const float4*>(w_pe + j * 32 + s0);
            const float4 c1v = *reinterpret_cast<const float4*>(w_pe + j * 32 + s0 + 4);
            const float cv[8] = {c0v.x, c0v.y, c0v.z, c0v.w, c1v.x, c1v.y, c1v.z, c1v.w};
#pragma unroll
            for (int ss = 0; ss < 8; ++ss)
#pragma unroll
              for (int n = 0; n < NU; ++n) acc[ss][n] = fmaf(wv[n], cv[ss], acc[ss][n]);
          }
          if constexpr (kBlock) {
#pragma unroll
            for (int ss = 0; ss < 8; ++ss) {
              const float a0 = fmaxf(acc[ss][0], 0.0f);
              float t = s_b0[lane];
              for (int i = 0; i < 32; ++i) t = fmaf(s_w0t[i * 32 + lane], __shfl_sync(0xffffffffu, a0, i), t);
              const float a1 = fmaxf(t, 0.0f);
              float d = s_b1[lane];
              for (int i = 0; i < 32; ++i) d = fmaf(s_w1t[i * 32 + lane], __shfl_sync(0xffffffffu, a1, i), d);
              acc[ss][0] += d;
            }
          }
#pragma unroll
          for (int ss = 0; ss < 8; ++ss) {
            float part = 0.0f;
#pragma unroll
            for (int n = 0; n < NU; ++n) part = fmaf(fmaxf(acc[ss][n], 0.0f), s_wout[lane * NU + n], part);
            const float tot = warp_sum(part);
            if (lane == s0 + ss) logit = tot + b_out;
          }
        }
      }
      const float sigma = bp.saved_sigma ? (valid ? __ldg(bp.saved_sigma + ray * K + k) : 0.0f) : softplus(logit);
      __syncwarp();
      // compositing forward (nerf.py:283-294)
      const float delta = (k + 1 < K) ? __fsub_rn(z_next, z) : 1e10f;
      const float e = expf(-(fabsf(delta) * fmaxf(sigma, 0.0f)));
      float alpha = __fsub_rn(1.0f, e);
      const bool capped = cfg.hard_alpha_cap && k == K - 1;
      if (capped) alpha = 1.0f;
      const float tr = valid ? __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f) : 1.0f;
      float incl = tr;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl *= up;
      }
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.0f;
      const float Tk = T_carry * excl;
      T_carry *= __shfl_sync(0xffffffffu, incl, 31);
      // dL/dw_k: rgb (nerf.py:298, :301-304), depth (:299), direct weights gradient
      float wbar = g_depth * z - g_rgb_sum;
      if (bp.g_weights && valid) wbar += __ldg(bp.g_weights + ray * K + k);
      if (bp.g_rgb) {
        for (int v = 0; v < Vc; ++v) {
          View vc;
          load_view(vc, p.scene.K_c + (b * Vc + v) * 9, p.scene.w2c_c + (b * Vc + v) * 16);
          const Proj pc = project(vc, X, Y, Z);
          const Taps t = bilinear_taps(pc.u, pc.v, Hc, Wc);
          const float4* base = reinterpret_cast<const float4*>(p.scene.img4) + ((b * Vc + v) * Hc + t.y0) * (int64_t)Wc + t.x0;
          const float4 c00 = __ldg(base), c01 = __ldg(base + t.dx), c10 = __ldg(base + t.dy * Wc), c11 = __ldg(base + t.dy * Wc + t.dx);
          const float w00 = t.wx0 * t.wy0, w01 = t.wx1 * t.wy0, w10 = t.wx0 * t.wy1, w11 = t.wx1 * t.wy1;
          const float col[4] = {fmaf(c11.x, w11, fmaf(c10.x, w10, fmaf(c01.x, w01, c00.x * w00))),
                                fmaf(c11.y, w11, fmaf(c10.y, w10, fmaf(c01.y, w01, c00.y * w00))),
                                fmaf(c11.z, w11, fmaf(c10.z, w10, fmaf(c01.z, w01, c00.z * w00))),
                                fmaf(c11.w, w11, fmaf(c10.w, w10, fmaf(c01.w, w01, c00.w * w00)))};
          for (int c = 0; c < Cc; ++c) wbar = fmaf(w_grgb[v * Cc + c], col[c], wbar);
        }
      }
      if (valid) {
        sc_alpha[k] = alpha;
        sc_T[k] = Tk;
        sc_wbar[k] = wbar;
        // d alpha / d logit = delta * exp(-delta sigma) * sigmoid(logit);  sigmoid(logit) = 1 - exp(-softplus(logit))
        sc_fac[k] = capped ? 0.0f : fabsf(delta) * e * (1.0f - expf(-sigma));
      }
      __syncwarp();
    }

    // ======================= reverse recurrence (division-free) =======================
    if (lane == 0) {
      float Rj = 0.0f;   // R_{K-1} = 0
      for (int j = K - 1; j >= 0; --j) {
        const float a = sc_alpha[j], T = sc_T[j];
        float abar = fmaf(sc_wbar[j], T, -T * Rj);
        if (bp.g_alphas) abar += __ldg(bp.g_alphas + ray * K + j);
        sc_lbar[j] = abar * sc_fac[j];
        Rj = fmaf(sc_wbar[j], a, ((1.0f - a) + 1e-10f) * Rj);   // becomes R_{j-1}
      }
    }
    __syncwarp();

    // ======================= sweep 2: MLP backward =======================
    z_ahead = lane < K ? sample_depth(p, ray, lane, near_, far_) : 0.0f;
    for (int c0 = 0; c0 < K; c0 += 32) {
      const int k = c0 + lane;
      const bool valid = k < K;
      const float z = z_ahead;
      z_ahead = (k + 32 < K) ? sample_depth(p, ray, k + 32, near_, far_) : 0.0f;
      // samples whose dL/dlogit is exactly zero (rays the loss masked out, the capped last sample, ...) contribute to no
      // gradient: skip whole chunks / 8-sample blocks of them (bit-identical result)
      if (__ballot_sync(0xffffffffu, valid && sc_lbar[k] != 0.0f) == 0u) continue;
      const float X = __fadd_rn(ro0, __fmul_rn(z, rd0)), Y = __fadd_rn(ro1, __fmul_rn(z, rd1)), Z = __fadd_rn(ro2, __fmul_rn(z, rd2));
      const Proj pf = project(vf, X, Y, Z);
      {
        positional_code(cfg, pf.u, pf.v, depth_code(cfg, pf), w_pe + lane, 32);
        for (int j = P; j < Ppad; ++j) w_pe[j * 32 + lane] = 0.0f;
        const Taps t = bilinear_taps(pf.u, pf.v, H, W);
        float4 a, w;
        a.x = __int_as_float((int)((b * H + t.y0) * (int64_t)W + t.x0));
        a.y = __int_as_float(t.dx);
        a.z = __int_as_float(t.dy * W);
        a.w = __int_as_float((valid ? 1 : 0) | ((use_empty && pf.invalid) ? 2 : 0));
        w.x = t.wx0; w.y = t.wx1; w.z = t.wy0; w.w = t.wy1;
        reinterpret_cast<float4*>(w_taps)[lane * 2 + 0] = a;
        reinterpret_cast<float4*>(w_taps)[lane * 2 + 1] = w;
      }
      __syncwarp();
      const int n_valid = min(32, K - c0);
      for (int s0 = 0; s0 < n_valid; s0 += 8) {
        {
          bool any = false;
#pragma unroll
          for (int ss = 0; ss < 8; ++ss) any |= (c0 + s0 + ss < K) && sc_lbar[c0 + s0 + ss] != 0.0f;
          if (!any) continue;
        }
        float acc[8][NU];
#pragma unroll
        for (int ss = 0; ss < 8; ++ss) {
          const float4 a = reinterpret_cast<const float4*>(w_taps)[(s0 + ss) * 2 + 0];
          const float4 w = reinterpret_cast<const float4*>(w_taps)[(s0 + ss) * 2 + 1];
          const int flags = __float_as_int(a.w);
          gather_units<NU>(p, a, w, flags, lane, acc[ss]);
        }
        for (int j = 0; j < P; ++j) {
          float wv[NU];
          load_units<NU>(s_wpe + j * Hd + lane * NU, wv);
          const float4 c0v = *reinterpret_cast<const float4*>(w_pe + j * 32 + s0);
          const float4 c1v = *reinterpret_cast<const float4*>(w_pe + j * 32 + s0 + 4);
          const float cv[8] = {c0v.x, c0v.y, c0v.z, c0v.w, c1v.x, c1v.y, c1v.z, c1v.w};
#pragma unroll
          for (int ss = 0; ss < 8; ++ss)
#pragma unroll
            for (int n = 0; n < NU; ++n) acc[ss][n] = fmaf(wv[n], cv[ss], acc[ss][n]);
        }
        // hbar[ss][n] = lbar_s * w_out[n] * [h > 0];   d w_out += lbar * relu(h);  d b_out += lbar
        float hbar[8][NU];
        if constexpr (kBlock) {
#pragma unroll 1
          for (int ss = 0; ss < 8; ++ss) {   // rolled: keeps the 104 gradient accumulators in registers
            const int ks = c0 + s0 + ss;
            const float lb = ks < K ? sc_lbar[ks] : 0.0f;
            if (lane == 0) acc_bout += lb;
            // recompute the block forward for this sample (lane = unit)
            const float h0 = acc[ss][0], a0 = fmaxf(h0, 0.0f);
            float t = s_b0[lane];
            for (int i = 0; i < 32; ++i) t = fmaf(s_w0t[i * 32 + lane], __shfl_sync(0xffffffffu, a0, i), t);
            const float a1 = fmaxf(t, 0.0f);
            float d = s_b1[lane];
            for (int i = 0; i < 32; ++i) d = fmaf(s_w1t[i * 32 + lane], __shfl_sync(0xffffffffu, a1, i), d);
            const float h1 = h0 + d;
            acc_wout[0] = fmaf(lb, fmaxf(h1, 0.0f), acc_wout[0]);
            const float hb1 = h1 > 0.0f ? lb * s_wout[lane] : 0.0f;          // dL/dh1 (also dL/d d, and the skip path)
            // fc1: d W1[lane][i] += hb1 * a1_i ; d b1 += hb1 ; dL/da1_i = sum_n W1[n][i] hb1_n
            acc_b1 += hb1;
            float ra1 = 0.0f;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              acc_w1[i] = fmaf(hb1, __shfl_sync(0xffffffffu, a1, i), acc_w1[i]);
              ra1 = fmaf(s_w1[i * 32 + lane], __shfl_sync(0xffffffffu, hb1, i), ra1);
            }
            const float tb = t > 0.0f ? ra1 : 0.0f;                            // dL/dt
            // fc0: d W0[lane][i] += tb * a0_i ; d b0 += tb ; dL/da0_i = sum_n W0[n][i] tb_n
            acc_b0 += tb;
            float ra0 = 0.0f;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              acc_w0[i] = fmaf(tb, __shfl_sync(0xffffffffu, a0, i), acc_w0[i]);
              ra0 = fmaf(s_w0[i * 32 + lane], __shfl_sync(0xffffffffu, tb, i), ra0);
            }
            hbar[ss][0] = hb1 + (h0 > 0.0f ? ra0 : 0.0f);                      // dL/dh0
          }
        } else {
#pragma unroll
          for (int ss = 0; ss < 8; ++ss) {
            const int ks = c0 + s0 + ss;
            const float lb = ks < K ? sc_lbar[ks] : 0.0f;
            if (lane == 0) acc_bout += lb;
#pragma unroll
            for (int n = 0; n < NU; ++n) {
              const float h = acc[ss][n];
              acc_wout[n] = fmaf(lb, fmaxf(h, 0.0f), acc_wout[n]);
              hbar[ss][n] = h > 0.0f ? lb * s_wout[lane * NU + n] : 0.0f;
            }
          }
        }
        // d W_pe[n][j] += sum_s hbar[s][n] * code[j][s]
#pragma unroll
        for (int j = 0; j < 40; ++j) {
          if (j < P) {
            const float4 c0v = *reinterpret_cast<const float4*>(w_pe + j * 32 + s0);
            const float4 c1v = *reinterpret_cast<const float4*>(w_pe + j * 32 + s0 + 4);
            const float cv[8] = {c0v.x, c0v.y, c0v.z, c0v.w, c1v.x, c1v.y, c1v.z, c1v.w};
#pragma unroll
            for (int ss = 0; ss < 8; ++ss)
#pragma unroll
              for (int n = 0; n < NU; ++n) acc_wpe[n][j] = fmaf(hbar[ss][n], cv[ss], acc_wpe[n][j]);
          }
        }
        // d G (bilinear scatter) / d g_empty
#pragma unroll
        for (int ss = 0; ss < 8; ++ss) {
          const float4 a = reinterpret_cast<const float4*>(w_taps)[(s0 + ss) * 2 + 0];
          const float4 w = reinterpret_cast<const float4*>(w_taps)[(s0 + ss) * 2 + 1];
          const int flags = __float_as_int(a.w);
          if (flags & 2) {
#pragma unroll
            for (int n = 0; n < NU; ++n) acc_gempty[n] += hbar[ss][n];
          } else if (flags & 1) {
            float* g00 = bp.d_G + (size_t)__float_as_int(a.x) * Hd + lane * NU;
            float* g10 = g00 + (size_t)__float_as_int(a.z) * Hd;
            const size_t dxo = (size_t)__float_as_int(a.y) * Hd;
            const float w00 = w.x * w.z, w01 = w.y * w.z, w10 = w.x * w.w, w11 = w.y * w.w;
            if constexpr (NU == 2) {
              const float2 hb = make_float2(hbar[ss][0], hbar[ss][1]);
              if (hb.x != 0.0f || hb.y != 0.0f) {
                atomicAdd(reinterpret_cast<float2*>(g00), make_float2(hb.x * w00, hb.y * w00));
                atomicAdd(reinterpret_cast<float2*>(g00 + dxo), make_float2(hb.x * w01, hb.y * w01));
                atomicAdd(reinterpret_cast<float2*>(g10), make_float2(hb.x * w10, hb.y * w10));
                atomicAdd(reinterpret_cast<float2*>(g10 + dxo), make_float2(hb.x * w11, hb.y * w11));
              }
            } else {
#pragma unroll
              for (int n = 0; n < NU; ++n) {
                const float hb = hbar[ss][n];
                if (hb != 0.0f) {
                  atomicAdd(g00 + n, hb * w00); atomicAdd(g00 + dxo + n, hb * w01);
                  atomicAdd(g10 + n, hb * w10); atomicAdd(g10 + dxo + n, hb * w11);
                }
              }
            }
          }
        }
      }
      __syncwarp();
    }
  }

  // ---- flush the per-lane accumulators (one atomic per value per warp) ----
#pragma unroll
  for (int n = 0; n < NU; ++n) {
    const int unit = lane * NU + n;
    if (bp.d_w_out) atomicAdd(bp.d_w_out + unit, acc_wout[n]);
    if (bp.d_g_empty) atomicAdd(bp.d_g_empty + unit, acc_gempty[n]);
#pragma unroll
    for (int j = 0; j < 40; ++j)
      if (j < P && bp.d_w_pe) atomicAdd(bp.d_w_pe + (size_t)unit * P + j, acc_wpe[n][j]);
  }
  if (lane == 0 && bp.d_b_out) atomicAdd(bp.d_b_out, acc_bout);
  if constexpr (kBlock) {
    if (bp.d_fc0_w) {
#pragma unroll
      for (int i = 0; i < 32; ++i) { atomicAdd(bp.d_fc0_w + lane * 32 + i, acc_w0[i]); atomicAdd(bp.d_fc1_w + lane * 32 + i, acc_w1[i]); }
      atomicAdd(bp.d_fc0_b + lane, acc_b0);
      atomicAdd(bp.d_fc1_b + lane, acc_b1);
    }
  }
}

size_t bwd_smem_bytes(int Hd, int p_pad, int K, int n_blocks) {
  const int Kpad = (K + 31) & ~31;
  const int blk = n_blocks ? 4 * Hd * Hd + 2 * Hd : 0;
  return (size_t)(p_pad * Hd + Hd + blk + kBwdWarps * (p_pad * 32 + 32 * 8 + 32 + 5 * Kpad)) * sizeof(float);
}

template <int NU, bool kBlock>
static cudaError_t launch_bwd_t(const BwdParams& bp, cudaStream_t stream) {
  const size_t smem = bwd_smem_bytes(NU * 32, bp.k.p_pad, bp.k.cfg.n_coarse, kBlock ? 1 : 0);
  auto kern = render_bwd_kernel<NU, kBlock>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 148, occ = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kBwdWarps * 32, smem);
  if (e != cudaSuccess) return e;
  if (occ < 1) return cudaErrorInvalidConfiguration;
  const int64_t want = (bp.k.n_items + kBwdWarps - 1) / kBwdWarps;
  const int64_t cap = (int64_t)sms * occ;
  const int grid = (int)(want < cap ? (want > 0 ? want : 1) : cap);
  kern<<<grid, kBwdWarps * 32, smem, stream>>>(bp);
  return cudaGetLastError();
}

bool bwd_tc_supported(const BwdParams& bp);
cudaError_t launch_bwd_tc(const BwdParams& bp, cudaStream_t stream);

// *used_tc (optional) reports which kernel family ran.  BTS_BWD_SIMT=1 forces the CUDA-core kernel (tests compare the two).
cudaError_t launch_bwd(const BwdParams& bp, cudaStream_t stream, bool* used_tc) {
  const char* force_simt = getenv("BTS_BWD_SIMT");
  const bool tc = !(force_simt && force_simt[0] == '1') && bwd_tc_supported(bp);
  if (used_tc) *used_tc = tc;
  if (tc) return launch_bwd_tc(bp, stream);
  if (bp.k.mlp.n_blocks == 1 && bp.k.mlp.Hd == 32) return launch_bwd_t<1, true>(bp, stream);
  if (bp.k.mlp.n_blocks != 0) return cudaErrorInvalidValue;
  switch (bp.k.mlp.Hd / 32) {
    case 1: return launch_bwd_t<1, false>(bp, stream);
    case 2: return launch_bwd_t<2, false>(bp, stream);
    case 3: return launch_bwd_t<3, false>(bp, stream);
    case 4: return launch_bwd_t<4, false>(bp, stream);
  }
  return cudaErrorInvalidValue;
}

}  // namespace bts
