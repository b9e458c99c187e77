// General fp32 render / point-query kernel (any K, Hd in {32,64,96,128}, n_blocks >= 0, any Vc).
//
// One warp renders one ray at a time (grid-stride).  A ray's K samples are processed in chunks of 32:
//   A  lane = sample : stratified depth, world point, projection into the encoder view, depth code, positional code
//                      (written to per-warp shared memory, sample-minor), bilinear tap set (shared memory)
//   B  lane = hidden units (NU = Hd/32 per lane): for 8 samples at a time, h = bilinear(G) + W_pe * code  (G carries the lin_in bias),
//      ResnetFC blocks, sigma logit = w_out . relu(h) + b_out  (warp reduction), handed back to lane = sample
//   C  lane = sample : alpha, transmittance scan (warp scan + carry across chunks), weights, depth
//   D  lane = sample : per colour view projection + bilinear RGB fetch, invalid flags, weighted colour reduction
// The (B*R*K, C+P) MLP input of the reference (412 B/sample, models_bts.py:184) is never materialised.
//
// Reference arithmetic: models/common/render/nerf.py:103-123 (sampling), :210-313 (composite);
// models/bts/model/models_bts.py:138-338 (field); models/common/model/resnetfc.py:53-62,132-184 (MLP).
#include "common.cuh"
#include "render_phases.cuh"

namespace bts {

constexpr int kWarpsPerCta = 4;
constexpr int kSB = 8;  // samples per register block in phase B

struct GeneralSmemLayout {
  int wpe, b_in, w_out, blocks, per_warp, total_floats;
  int warp_pe, warp_taps, warp_rgb, warp_floats;
};

__host__ __device__ inline GeneralSmemLayout general_layout(int Hd, int p_pad, int n_blocks) {
  GeneralSmemLayout L;
  int o = 0;
  L.wpe = o;    o += p_pad * Hd;
  L.b_in = o;   o += Hd;
  L.w_out = o;  o += Hd;
  L.blocks = o; o += n_blocks * 2 * (Hd * Hd + Hd);
  L.warp_pe = 0;
  L.warp_taps = p_pad * 32;
  L.warp_rgb = L.warp_taps + 32 * 8;
  L.warp_floats = L.warp_rgb + kMaxColorViews * 4 + kMaxColorViews;
  L.per_warp = o;
  L.total_floats = o + kWarpsPerCta * L.warp_floats;
  return L;
}

template <int NU, bool kPoints>
__global__ void __launch_bounds__(kWarpsPerCta * 32) render_general_kernel(const KernelParams p) {
  extern __shared__ __align__(16) float smem[];
  constexpr int Hd = NU * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int P = p.mlp.P, Ppad = p.p_pad, nb = p.mlp.n_blocks;
  const GeneralSmemLayout L = general_layout(Hd, Ppad, nb);
  float* s_wpe = smem + L.wpe;      // [Ppad][Hd]  (W_in[:, C + j] transposed; rows >= P are zero)
  float* s_wout = smem + L.w_out;
  float* s_blk = smem + L.blocks;   // per block: fc0^T [Hd][Hd], b0 [Hd], fc1^T [Hd][Hd], b1 [Hd]
  float* w_pe = smem + L.per_warp + warp * L.warp_floats + L.warp_pe;     // [Ppad][32]
  float* w_taps = smem + L.per_warp + warp * L.warp_floats + L.warp_taps; // [32][8]
  float* w_rgb = smem + L.per_warp + warp * L.warp_floats + L.warp_rgb;   // [Vc][4]

  // ---- stage the weights once per CTA ----
  {
    const int ld = p.mlp.C + P;
    for (int i = threadIdx.x; i < Ppad * Hd; i += blockDim.x) {
      const int j = i / Hd, n = i % Hd;
      s_wpe[i] = j < P ? __ldg(p.mlp.w_in + (size_t)n * ld + p.mlp.C + j) : 0.0f;
    }
    for (int i = threadIdx.x; i < Hd; i += blockDim.x) {
      s_wout[i] = __ldg(p.mlp.w_out + i);
    }
    for (int b = 0; b < nb; ++b) {
      float* base = s_blk + b * 2 * (Hd * Hd + Hd);
      for (int i = threadIdx.x; i < Hd * Hd; i += blockDim.x) {
        const int n = i / Hd, k = i % Hd;  // coalesced read of W[n][k], transposed store
        base[k * Hd + n] = __ldg(p.mlp.fc0_w[b] + i);
        base[Hd * Hd + Hd + k * Hd + n] = __ldg(p.mlp.fc1_w[b] + i);
      }
      for (int i = threadIdx.x; i < Hd; i += blockDim.x) {
        base[Hd * Hd + i] = __ldg(p.mlp.fc0_b[b] + i);
        base[2 * Hd * Hd + Hd + i] = __ldg(p.mlp.fc1_b[b] + i);
      }
    }
  }
  __syncthreads();

  const BtsRenderCfg& cfg = p.cfg;
  const int K = kPoints ? 32 : cfg.n_coarse;
  const int Vc = p.scene.Vc, Cc = p.scene.Cc;
  const int H = p.scene.H, W = p.scene.W, Hc = p.scene.Hc, Wc = p.scene.Wc;
  const float b_out = __ldg(p.mlp.b_out);
  const bool use_empty = cfg.learn_empty && p.scene.g_empty != nullptr;

  // item = ray (render) or group of 32 points (query)
  const int64_t n_groups = kPoints ? (p.n_items + 31) / 32 : p.n_items;
  for (int64_t item = (int64_t)blockIdx.x * kWarpsPerCta + warp; item < n_groups;
       item += (int64_t)gridDim.x * kWarpsPerCta) {
    // ---- per-item setup ----
    float ro[3] = {0, 0, 0}, rd[3] = {0, 0, 0}, near_ = 0, far_ = 0, zfac = 1.0f;
    int64_t b;
    if (!kPoints) {
      b = p.scene.n == 1 ? 0 : item / p.per_batch;
      const Ray ry = fetch_ray(p, item, lane);
      ro[0] = ry.o[0]; ro[1] = ry.o[1]; ro[2] = ry.o[2]; rd[0] = ry.d[0]; rd[1] = ry.d[1]; rd[2] = ry.d[2];
      depth_bounds(p, ry, near_, far_);
      zfac = ry.zfac;
    } else {
      b = 0;  // per-lane batch index below
    }
    float T_carry = 1.0f, depth_part = 0.0f, wsum_part = 0.0f;
    for (int i = lane; i < kMaxColorViews * 5; i += 32) w_rgb[i] = 0.0f;
    __syncwarp();

    float z_ahead = 0.0f;   // depths are produced one chunk ahead (delta needs z[k+1])
    if (!kPoints && lane < K) z_ahead = sample_depth(p, item, lane, near_, far_);
    for (int c0 = 0; c0 < K; c0 += 32) {
      const int k = c0 + lane;
      int64_t sample_id;  // flat index of this lane's sample/point
      bool valid;
      float z = 0.0f, z_next = 0.0f, X, Y, Z;
      int64_t bl = b;
      if (!kPoints) {
        valid = k < K;
        sample_id = item * K + k;
        z = z_ahead;
        z_ahead = (k + 32 < K) ? sample_depth(p, item, k + 32, near_, far_) : 0.0f;
        z_next = __shfl_down_sync(0xffffffffu, z, 1);
        const float z_first_of_next = __shfl_sync(0xffffffffu, z_ahead, 0);
        if (lane == 31) z_next = z_first_of_next;
        X = __fadd_rn(ro[0], __fmul_rn(z, rd[0]));   // nerf.py:231 (separate mul and add)
        Y = __fadd_rn(ro[1], __fmul_rn(z, rd[1]));
        Z = __fadd_rn(ro[2], __fmul_rn(z, rd[2]));
      } else {
        sample_id = item * 32 + lane;
        valid = sample_id < p.n_items;
        const int64_t sid = valid ? sample_id : 0;
        bl = sid / p.per_batch;
        X = __ldg(p.xyz + sid * 3 + 0); Y = __ldg(p.xyz + sid * 3 + 1); Z = __ldg(p.xyz + sid * 3 + 2);
      }

      // ---- phase A: projection into the encoder view, code, taps ----
      View vf;
      load_view(vf, p.scene.K_f + bl * 9, p.scene.w2c_f + bl * 16);
      const Proj pf = project(vf, X, Y, Z);
      const bool inv_f = pf.invalid;
      {
        const float zh = depth_code(cfg, pf);
        positional_code(cfg, pf.u, pf.v, zh, w_pe + lane, 32);
        for (int j = P; j < Ppad; ++j) w_pe[j * 32 + lane] = 0.0f;
        const Taps t = bilinear_taps(pf.u, pf.v, H, W);
        const int p00 = (int)((bl * H + t.y0) * (int64_t)W + t.x0);
        float4 a, w;
        a.x = __int_as_float(p00);
        a.y = __int_as_float(t.dx);
        a.z = __int_as_float(t.dy * W);
        a.w = __int_as_float((valid ? 1 : 0) | ((use_empty && inv_f) ? 2 : 0));
        w.x = t.wx0; w.y = t.wx1; w.z = t.wy0; w.w = t.wy1;
        reinterpret_cast<float4*>(w_taps)[lane * 2 + 0] = a;
        reinterpret_cast<float4*>(w_taps)[lane * 2 + 1] = w;
      }
      __syncwarp();

      // ---- phase B: MLP, lanes = hidden units ----
      float logit = 0.0f;
      const int n_valid = kPoints ? 32 : min(32, K - c0);
      for (int s0 = 0; s0 < n_valid; s0 += kSB) {
        float acc[kSB][NU];
#pragma unroll
        for (int ss = 0; ss < kSB; ++ss) {
          const float4 a = reinterpret_cast<const float4*>(w_taps)[(s0 + ss) * 2 + 0];
          const float4 w = reinterpret_cast<const float4*>(w_taps)[(s0 + ss) * 2 + 1];
          const int flags = __float_as_int(a.w);
          float g[NU];
          if (flags & 2) {
#pragma unroll
            for (int n = 0; n < NU; ++n) g[n] = __ldg(p.scene.g_empty + lane * NU + n);
          } else if (flags & 1) {
            const float* g00 = p.scene.G + (size_t)__float_as_int(a.x) * Hd + lane * NU;
            const float* g01 = g00 + (size_t)__float_as_int(a.y) * Hd;
            const float* g10 = g00 + (size_t)__float_as_int(a.z) * Hd;
            const float* g11 = g10 + (size_t)__float_as_int(a.y) * Hd;
            float v00[NU], v01[NU], v10[NU], v11[NU];
            if constexpr (NU == 1) {
              v00[0] = __ldg(g00); v01[0] = __ldg(g01); v10[0] = __ldg(g10); v11[0] = __ldg(g11);
            } else if constexpr (NU == 2) {
              const float2 q00 = __ldg(reinterpret_cast<const float2*>(g00)), q01 = __ldg(reinterpret_cast<const float2*>(g01));
              const float2 q10 = __ldg(reinterpret_cast<const float2*>(g10)), q11 = __ldg(reinterpret_cast<const float2*>(g11));
              v00[0] = q00.x; v00[1] = q00.y; v01[0] = q01.x; v01[1] = q01.y;
              v10[0] = q10.x; v10[1] = q10.y; v11[0] = q11.x; v11[1] = q11.y;
            } else if constexpr (NU == 4) {
              const float4 q00 = __ldg(reinterpret_cast<const float4*>(g00)), q01 = __ldg(reinterpret_cast<const float4*>(g01));
              const float4 q10 = __ldg(reinterpret_cast<const float4*>(g10)), q11 = __ldg(reinterpret_cast<const float4*>(g11));
              v00[0] = q00.x; v00[1] = q00.y; v00[2] = q00.z; v00[3] = q00.w;
              v01[0] = q01.x; v01[1] = q01.y; v01[2] = q01.z; v01[3] = q01.w;
              v10[0] = q10.x; v10[1] = q10.y; v10[2] = q10.z; v10[3] = q10.w;
              v11[0] = q11.x; v11[1] = q11.y; v11[2] = q11.z; v11[3] = q11.w;
            } else {
#pragma unroll
              for (int n = 0; n < NU; ++n) { v00[n] = __ldg(g00 + n); v01[n] = __ldg(g01 + n); v10[n] = __ldg(g10 + n); v11[n] = __ldg(g11 + n); }
            }
            const float w00 = w.x * w.z, w01 = w.y * w.z, w10 = w.x * w.w, w11 = w.y * w.w;
#pragma unroll
            for (int n = 0; n < NU; ++n) g[n] = fmaf(v11[n], w11, fmaf(v10[n], w10, fmaf(v01[n], w01, v00[n] * w00)));
          } else {
#pragma unroll
            for (int n = 0; n < NU; ++n) g[n] = 0.0f;
          }
#pragma unroll
          for (int n = 0; n < NU; ++n) acc[ss][n] = g[n];   // lin_in bias is already inside G / g_empty
        }
        // positional-code part of lin_in: acc += W_pe[:, j] * code[j]
        for (int j = 0; j < P; ++j) {
          float wv[NU];
#pragma unroll
          for (int n = 0; n < NU; ++n) wv[n] = s_wpe[j * Hd + lane * NU + n];
          const float4 c0v = *reinterpret_cast<const float4*>(w_pe + j * 32 + s0);
          const float4 c1v = *reinterpret_cast<const float4*>(w_pe + j * 32 + s0 + 4);
          const float cv[kSB] = {c0v.x, c0v.y, c0v.z, c0v.w, c1v.x, c1v.y, c1v.z, c1v.w};
#pragma unroll
          for (int ss = 0; ss < kSB; ++ss)
#pragma unroll
            for (int n = 0; n < NU; ++n) acc[ss][n] = fmaf(wv[n], cv[ss], acc[ss][n]);
        }
        // ResnetFC blocks: h += fc1(relu(fc0(relu(h))))   (resnetfc.py:53-62)
        for (int blk = 0; blk < nb; ++blk) {
          const float* w0t = s_blk + blk * 2 * (Hd * Hd + Hd);
          const float* b0 = w0t + Hd * Hd;
          const float* w1t = b0 + Hd;
          const float* b1 = w1t + Hd * Hd;
          float t[kSB][NU];
#pragma unroll
          for (int ss = 0; ss < kSB; ++ss)
#pragma unroll
            for (int n = 0; n < NU; ++n) t[ss][n] = b0[lane * NU + n];
          for (int il = 0; il < 32; ++il) {
#pragma unroll
            for (int nn = 0; nn < NU; ++nn) {
              const int i = il * NU + nn;
              float wv[NU];
#pragma unroll
              for (int n = 0; n < NU; ++n) wv[n] = w0t[i * Hd + lane * NU + n];
#pragma unroll
              for (int ss = 0; ss < kSB; ++ss) {
                const float x = __shfl_sync(0xffffffffu, fmaxf(acc[ss][nn], 0.0f), il);
#pragma unroll
                for (int n = 0; n < NU; ++n) t[ss][n] = fmaf(wv[n], x, t[ss][n]);
              }
            }
          }
          float d[kSB][NU];
#pragma unroll
          for (int ss = 0; ss < kSB; ++ss)
#pragma unroll
            for (int n = 0; n < NU; ++n) d[ss][n] = b1[lane * NU + n];
          for (int il = 0; il < 32; ++il) {
#pragma unroll
            for (int nn = 0; nn < NU; ++nn) {
              const int i = il * NU + nn;
              float wv[NU];
#pragma unroll
              for (int n = 0; n < NU; ++n) wv[n] = w1t[i * Hd + lane * NU + n];
#pragma unroll
              for (int ss = 0; ss < kSB; ++ss) {
                const float x = __shfl_sync(0xffffffffu, fmaxf(t[ss][nn], 0.0f), il);
#pragma unroll
                for (int n = 0; n < NU; ++n) d[ss][n] = fmaf(wv[n], x, d[ss][n]);
              }
            }
          }
#pragma unroll
          for (int ss = 0; ss < kSB; ++ss)
#pragma unroll
            for (int n = 0; n < NU; ++n) acc[ss][n] += d[ss][n];
        }
        // lin_out on relu(h)
#pragma unroll
        for (int ss = 0; ss < kSB; ++ss) {
          float part = 0.0f;
#pragma unroll
          for (int n = 0; n < NU; ++n) part = fmaf(fmaxf(acc[ss][n], 0.0f), s_wout[lane * NU + n], part);
          const float tot = warp_sum(part);
          if (lane == s0 + ss) logit = tot + b_out;
        }
      }
      const float sigma = softplus(logit);  // models_bts.py:311-313
      __syncwarp();                          // phase B finished reading w_pe / w_taps

      // ---- phase C: alpha compositing (nerf.py:283-299) ----
      float wgt = 0.0f;
      if (!kPoints) {
        wgt = composite_chunk(p, lane, k, K, valid, z, z_next, sigma, sample_id, T_carry, depth_part, wsum_part);
      } else if (valid) {
        p.q_sigma[sample_id] = sigma;
        p.q_invalid_f[sample_id] = inv_f ? 1.0f : 0.0f;
      }

      // ---- phase D: colour views (models_bts.py:218-264, :330-334) ----
      colour_views<kPoints>(p, lane, bl, X, Y, Z, valid, pf, wgt, sample_id, w_rgb);
      __syncwarp();
    }  // chunks

    if (!kPoints) finish_ray(p, lane, item, depth_part, wsum_part, w_rgb, zfac);
  }
}

// ---------------------------------------------------------------- host-side launcher ---------------------------
template <int NU, bool kPoints>
static cudaError_t launch_general_t(const KernelParams& p, cudaStream_t stream) {
  const GeneralSmemLayout L = general_layout(NU * 32, p.p_pad, p.mlp.n_blocks);
  const size_t smem = (size_t)L.total_floats * sizeof(float);
  auto kern = render_general_kernel<NU, kPoints>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 148, occ = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kWarpsPerCta * 32, smem);
  if (e != cudaSuccess) return e;
  if (occ < 1) return cudaErrorInvalidConfiguration;
  const int64_t groups = kPoints ? (p.n_items + 31) / 32 : p.n_items;
  const int64_t want = (groups + kWarpsPerCta - 1) / kWarpsPerCta;
  const int64_t cap = (int64_t)sms * occ * 2;   // two waves of resident CTAs, grid-stride inside
  const int grid = (int)(want < cap ? (want > 0 ? want : 1) : cap);
  kern<<<grid, kWarpsPerCta * 32, smem, stream>>>(p);
  return cudaGetLastError();
}

size_t general_smem_bytes(int Hd, int p_pad, int n_blocks) {
  return (size_t)general_layout(Hd, p_pad, n_blocks).total_floats * sizeof(float);
}

cudaError_t launch_general(const KernelParams& p, bool points, cudaStream_t stream) {
  const int nu = p.mlp.Hd / 32;
  if (points) {
    switch (nu) {
      case 1: return launch_general_t<1, true>(p, stream);
      case 2: return launch_general_t<2, true>(p, stream);
      case 3: return launch_general_t<3, true>(p, stream);
      case 4: return launch_general_t<4, true>(p, stream);
    }
  } else {
    switch (nu) {
      case 1: return launch_general_t<1, false>(p, stream);
      case 2: return launch_general_t<2, false>(p, stream);
      case 3: return launch_general_t<3, false>(p, stream);
      case 4: return launch_general_t<4, false>(p, stream);
    }
  }
  return cudaErrorInvalidValue;
}

}  // namespace bts
