// Tensor-core backward of the fused render for the KITTI training configurations (d_hidden = 64, n_blocks = 0, 6-octave
// code, K a multiple of 32, forward sigma saved): the two 39-wide contractions of the CUDA-core backward
// (render_bwd.cu: recompute of h = G_interp + W_pe code, and dW_pe += hbar^T code) run on tcgen05.
//
// CTA = 4 warps, warp = ray, thread = sample of the ray's current 32-sample chunk (the forward's layout, render_tc.cu):
//   sweep 1  (per warp, CUDA cores): compositing forward from the saved sigma, dL/dw per sample (colour views), the
//            division-free reverse recurrence -> lbar_k = dL/dlogit_k for the whole ray (shared memory)
//   sweep 2  per chunk:
//     P1   positional code -> A operand of MMA 1 in TMEM (fp16 hi/lo + tf32 raw, as the forward), and lbar * code as a
//          THREE-term bf16 split (8 + 8 + 8 mantissa bits, fp32 exponent range: no underflow for tiny lbar) into the
//          K-major (K = sample) operand tile of MMA 2: rows j, 40 + j, 80 + j
//     MMA1 D_h[128 samples x 64] = code W_pe^T                       (same instruction sequence as the forward)
//     P2   bilinear gather of G -> interpolated rows (shared memory)
//     P3   h = D_h + G_interp; relu mask; d w_out, d b_out in registers; hbar -> 16-byte vector reductions into dG
//          (4 taps x 16 per sample); the mask (exact 0/1 in bf16) -> N operand tile of MMA 2 (K = sample)
//     MMA2 D_2[128 x 64 units] += (lbar code)^T_{3 terms} mask        accumulated in TMEM over the whole kernel
//   end      d W_pe[u, j] = w_out[u] * (D_2[j] + D_2[40 + j] + D_2[80 + j])[u]   (one atomic per element per CTA)
// since hbar[s, u] = lbar_s w_out[u] mask[s, u], the outer-product sum factors into a product with an exact 0/1 matrix.
// Operand tiles are K-major / no-swizzle core matrices whose chunk stride (the descriptor's leading-dimension offset) is
// padded by 32 B so the 2-byte scatter of a warp (8 lanes per 16-byte chunk, 4 chunks) touches distinct banks.
// Gradients are identical in structure to render_bwd.cu's; tests compare both against autograd through the oracle.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <type_traits>

#include "common.cuh"
#include "render_phases.cuh"
#include "bwd_params.cuh"
#include "tcgen05.cuh"

namespace bts {

__device__ int g_bwd_tc_barrier_timeouts = 0;

namespace bwdtc {

constexpr int kThreads = 128;
constexpr int HD = 64;
constexpr uint32_t kColD = 0, kColAhi = HD, kColAlo = HD + 24, kColAraw = HD + 48, kColD2 = 128, kTmemCols = 256;
constexpr int kGiStride = HD + 4;
constexpr int kPadLbo = 32;                                   // bytes added to the chunk stride of the MMA-2 operand tiles
constexpr int kStrideA2 = 128 * 16 + kPadLbo;                 // [16 chunks of 8 samples][128 rows][16 B] bf16
constexpr int kStrideMask = HD * 16 + kPadLbo;                // [16 chunks][64 rows][16 B] bf16
constexpr int kOffW16hi = 0;
constexpr int kOffW16lo = kOffW16hi + 6 * HD * 16;
constexpr int kOffWrawA = kOffW16lo + 6 * HD * 16;
constexpr int kOffWrawB = kOffWrawA + 2 * HD * 16;
constexpr int kOffWout = kOffWrawB + 2 * HD * 16;
constexpr int kOffGempty = kOffWout + HD * 4;
constexpr int kOffA2 = kOffGempty + HD * 4;
constexpr int kOffMask = kOffA2 + 16 * kStrideA2;
constexpr int kOffBar = kOffMask + 16 * kStrideMask;          // bar1, bar2, tmem slot
constexpr int kOffWarp = kOffBar + 32;
constexpr int kWarpTaps = 0;                                  // [32][8] floats
constexpr int kWarpGi = 32 * 8 * 4;                           // [32][HD + 4] floats
constexpr int kWarpGrgb = kWarpGi + 32 * kGiStride * 4;       // [32] floats
constexpr int kWarpSc = kWarpGrgb + 32 * 4;                   // 5 x [Kpad] floats
static_assert(kOffA2 % 128 == 0 && kOffMask % 16 == 0 && kOffWarp % 16 == 0, "operand tile alignment");

__host__ __device__ constexpr int warp_bytes(int Kpad) { return kWarpSc + 5 * Kpad * 4; }

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void split_h2(float s, float c, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(s, c);
  const float2 hf = __half22float2(h);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = pack_h2(s - hf.x, c - hf.y);
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// v -> three bf16 terms (v = t0 + t1 + t2 to 2^-24) stored at rows j, 40 + j, 80 + j of this sample's chunk column
__device__ __forceinline__ void store_split3(unsigned char* col, int j, float v) {
  const __nv_bfloat16 t0 = __float2bfloat16_rn(v);
  const float r0 = v - __bfloat162float(t0);
  const __nv_bfloat16 t1 = __float2bfloat16_rn(r0);
  const float r1 = r0 - __bfloat162float(t1);
  const __nv_bfloat16 t2 = __float2bfloat16_rn(r1);
  *reinterpret_cast<__nv_bfloat16*>(col + j * 16) = t0;
  *reinterpret_cast<__nv_bfloat16*>(col + (40 + j) * 16) = t1;
  *reinterpret_cast<__nv_bfloat16*>(col + (80 + j) * 16) = t2;
}

__global__ void __launch_bounds__(kThreads, 2) render_bwd_tc_kernel(const BwdParams bp) {
  extern __shared__ __align__(128) unsigned char smem[];
  const KernelParams& p = bp.k;
  const BtsRenderCfg& cfg = p.cfg;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // the MMA-issuing warp differs between the CTAs that share an SM (warp w runs on scheduler w % 4)
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const int mma_warp = (int)(blockIdx.x / tc::num_sms()) & 3;
  const int K = cfg.n_coarse, Kpad = (K + 31) & ~31;
  const int C = p.mlp.C, P = p.mlp.P, ld = C + P;
  const int Vc = p.scene.Vc, Cc = p.scene.Cc;
  const int H = p.scene.H, W = p.scene.W, Hc = p.scene.Hc, Wc = p.scene.Wc;
  uint64_t* bar1 = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* bar2 = bar1 + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar1 + 2);
  float* s_wout = reinterpret_cast<float*>(smem + kOffWout);
  float* s_gempty = reinterpret_cast<float*>(smem + kOffGempty);
  unsigned char* wbase = smem + kOffWarp + warp * warp_bytes(Kpad);
  float4* w_taps = reinterpret_cast<float4*>(wbase + kWarpTaps);
  float* w_gi = reinterpret_cast<float*>(wbase + kWarpGi);
  float* w_grgb = reinterpret_cast<float*>(wbase + kWarpGrgb);
  float* sc_alpha = reinterpret_cast<float*>(wbase + kWarpSc);
  float* sc_T = sc_alpha + Kpad;
  float* sc_wbar = sc_T + Kpad;
  float* sc_fac = sc_wbar + Kpad;
  float* sc_lbar = sc_fac + Kpad;

  // ---- one-time setup (weights as in render_tc.cu) ----
  for (int i = tid; i < HD * 48; i += kThreads) {
    const int n = i / 48, c = i % 48;
    float w = 0.0f;
    if (c < 36) {   // internal code column c = (d*6 + oct)*2 + sc  <->  reference code index 3 + oct*6 + sc*3 + d
      const int sc = c & 1, di = c >> 1, d = di / 6, oct = di % 6;
      w = __ldg(p.mlp.w_in + (size_t)n * ld + C + 3 + oct * 6 + sc * 3 + d);
    }
    const __half hi = __float2half_rn(w);
    const __half lo = __float2half_rn(w - __half2float(hi));
    const int off = ((c >> 3) * HD + n) * 16 + (c & 7) * 2;
    *reinterpret_cast<__half*>(smem + kOffW16hi + off) = hi;
    *reinterpret_cast<__half*>(smem + kOffW16lo + off) = lo;
  }
  for (int n = tid; n < HD; n += kThreads) {
    float hi[3], lo[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float w = __ldg(p.mlp.w_in + (size_t)n * ld + C + d);
      hi[d] = tf32_hi(w);
      lo[d] = w - hi[d];
    }
    *reinterpret_cast<float4*>(smem + kOffWrawA + (0 * HD + n) * 16) = make_float4(hi[0], hi[1], hi[2], hi[0]);
    *reinterpret_cast<float4*>(smem + kOffWrawA + (1 * HD + n) * 16) = make_float4(hi[1], hi[2], 0.f, 0.f);
    *reinterpret_cast<float4*>(smem + kOffWrawB + (0 * HD + n) * 16) = make_float4(lo[0], lo[1], lo[2], 0.f);
    *reinterpret_cast<float4*>(smem + kOffWrawB + (1 * HD + n) * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
    s_wout[n] = __ldg(p.mlp.w_out + n);
    s_gempty[n] = p.scene.g_empty ? __ldg(p.scene.g_empty + n) : 0.0f;
  }
  for (int i = tid; i < (16 * kStrideA2 + 16 * kStrideMask) / 16; i += kThreads)      // rows 39, 79, 119.. and the padding stay zero
    reinterpret_cast<float4*>(smem + kOffA2)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (warp == 0) tc::tmem_alloc<kTmemCols>(tmem_slot);
  if (tid == 0) {
    tc::mbar_init(bar1, 1);
    tc::mbar_init(bar2, 1);
    tc::fence_mbar_init();
  }
  tc::fence_proxy_async_smem();
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);
  uint32_t phase1 = 0, phase2 = 0;
  bool d2_pending = false, d2_started = false;
  const float b_out = __ldg(p.mlp.b_out);
  const bool use_empty = cfg.learn_empty && p.scene.g_empty != nullptr;
  const uint32_t idesc16 = tc::instr_desc(tc::kFmtF16, 128, HD), idesc32 = tc::instr_desc(tc::kFmtTF32, 128, HD);
  const uint32_t idescbf = tc::instr_desc(tc::kFmtBF16, 128, HD);
  auto wait_bar = [&](uint64_t* bar, uint32_t& phase) {
    tc::mbar_wait_or_trap(bar, phase, &g_bwd_tc_barrier_timeouts);   // bounded: a pipeline bug traps (CUDA error) instead of hanging
    phase ^= 1u;
  };

  float acc_wout[HD];
#pragma unroll
  for (int u = 0; u < HD; ++u) acc_wout[u] = 0.0f;
  float acc_bout = 0.0f;
  // this sample's column inside the MMA-2 operand tiles: chunk = (warp*32 + lane) / 8, position lane % 8
  unsigned char* a2_col = smem + kOffA2 + (warp * 4 + (lane >> 3)) * kStrideA2 + (lane & 7) * 2;
  unsigned char* mask_col = smem + kOffMask + (warp * 4 + (lane >> 3)) * kStrideMask + (lane & 7) * 2;

  const int64_t n_rays = p.n_items;
  const int64_t n_pass = (n_rays + 3) / 4;
  for (int64_t pass = blockIdx.x; pass < n_pass; pass += gridDim.x) {
    const int64_t ray_raw = pass * 4 + warp;
    const bool active = ray_raw < n_rays;
    const int64_t ray = active ? ray_raw : n_rays - 1;
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

    // ======================= sweep 1: compositing forward from the saved sigma (render_bwd.cu) =======================
    {
      float T_carry = 1.0f;
      float z_ahead = sample_depth(p, ray, lane, near_, far_);
      for (int c0 = 0; c0 < K; c0 += 32) {
        const int k = c0 + lane;
        const float z = z_ahead;
        z_ahead = (k + 32 < K) ? sample_depth(p, ray, k + 32, near_, far_) : 0.0f;
        float z_next = __shfl_down_sync(0xffffffffu, z, 1);
        const float z_first_of_next = __shfl_sync(0xffffffffu, z_ahead, 0);
        if (lane == 31) z_next = z_first_of_next;
        const float X = __fadd_rn(ro0, __fmul_rn(z, rd0)), Y = __fadd_rn(ro1, __fmul_rn(z, rd1)), Z = __fadd_rn(ro2, __fmul_rn(z, rd2));
        const float sigma = __ldg(bp.saved_sigma + ray * K + k);
        const float delta = (k + 1 < K) ? __fsub_rn(z_next, z) : 1e10f;
        const float e = expf(-(fabsf(delta) * fmaxf(sigma, 0.0f)));
        float alpha = __fsub_rn(1.0f, e);
        const bool capped = cfg.hard_alpha_cap && k == K - 1;
        if (capped) alpha = 1.0f;
        const float tr = __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f);
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
        float wbar = g_depth * z - g_rgb_sum;
        if (bp.g_weights) wbar += __ldg(bp.g_weights + ray * K + k);
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
        sc_alpha[k] = alpha;
        sc_T[k] = Tk;
        sc_wbar[k] = wbar;
        sc_fac[k] = capped ? 0.0f : fabsf(delta) * e * (1.0f - expf(-sigma));
      }
      __syncwarp();
      // reverse recurrence R_{j-1} = wbar_j a_j + t_j R_j (division-free) as a suffix scan of the affine maps
      // f_j(x) = b_j + t_j x over the lanes of a chunk, chunks visited last to first with the carry R_{c0+31}
      float R_carry = 0.0f;   // R_{K-1} = 0
      for (int c0 = K - 32; c0 >= 0; c0 -= 32) {
        const int k = c0 + lane;
        const float a = sc_alpha[k], Tk = sc_T[k], wb = sc_wbar[k];
        float St = (1.0f - a) + 1e-10f, Sb = wb * a;               // S_l = f_l o f_{l+1} o ... o f_31 after the scan
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float t2 = __shfl_down_sync(0xffffffffu, St, o), b2 = __shfl_down_sync(0xffffffffu, Sb, o);
          if (lane + o < 32) { Sb = fmaf(St, b2, Sb); St *= t2; }
        }
        // R_j = S_{l+1}(R_carry); lane 31: the carry itself
        const float nt = __shfl_down_sync(0xffffffffu, St, 1), nb = __shfl_down_sync(0xffffffffu, Sb, 1);
        const float Rj = lane == 31 ? R_carry : fmaf(nt, R_carry, nb);
        float abar = fmaf(wb, Tk, -Tk * Rj);
        if (bp.g_alphas) abar += __ldg(bp.g_alphas + ray * K + k);
        sc_lbar[k] = active ? abar * sc_fac[k] : 0.0f;
        R_carry = fmaf(__shfl_sync(0xffffffffu, St, 0), R_carry, __shfl_sync(0xffffffffu, Sb, 0));   // R_{c0-1}
      }
      __syncwarp();
    }

    // ======================= sweep 2: MLP backward on the tensor cores =======================
    float z_ahead = sample_depth(p, ray, lane, near_, far_);
    for (int c0 = 0; c0 < K; c0 += 32) {
      const int k = c0 + lane;
      const float z = z_ahead;
      z_ahead = (k + 32 < K) ? sample_depth(p, ray, k + 32, near_, far_) : 0.0f;
      const float lb = sc_lbar[k];
      // a chunk whose 128 samples all have dL/dlogit == 0 contributes nothing (CTA-uniform decision; also orders the
      // previous chunk's accumulator reads before this chunk's MMA)
      if (!__syncthreads_or(lb != 0.0f)) continue;
      const float X = __fadd_rn(ro0, __fmul_rn(z, rd0)), Y = __fadd_rn(ro1, __fmul_rn(z, rd1)), Z = __fadd_rn(ro2, __fmul_rn(z, rd2));
      const Proj pf = project(vf, X, Y, Z);
      const bool inv_f = pf.invalid;
      const float pv[3] = {pf.u, pf.v, depth_code(cfg, pf)};
      bool any_empty;
      {
        const Taps t = bilinear_taps(pf.u, pf.v, H, W);
        const int p00 = (int)((b * H + t.y0) * (int64_t)W + t.x0);
        float4 a, w;
        a.x = __int_as_float(p00 * (HD / 4));
        a.y = __int_as_float(t.dx * (HD / 4));
        a.z = __int_as_float(t.dy * W * (HD / 4));
        a.w = __int_as_float((use_empty && inv_f) ? 1 : 0);
        any_empty = __any_sync(0xffffffffu, use_empty && inv_f);
        w.x = t.wx0 * t.wy0; w.y = t.wx1 * t.wy0; w.z = t.wx0 * t.wy1; w.w = t.wx1 * t.wy1;
        w_taps[lane * 2 + 0] = a;
        w_taps[lane * 2 + 1] = w;
      }
      const bool big = fmaxf(fabsf(pv[0]), fmaxf(fabsf(pv[1]), fabsf(pv[2]))) > kPeFastLimit;
      if (d2_pending) {      // MMA 2 of the previous chunk still reads the operand tiles P1 / P3 rewrite
        wait_bar(bar2, phase2);
        d2_pending = false;
      }
      // ---------------- P1: code -> TMEM A operand (MMA 1) and lbar * code -> smem operand (MMA 2) ----------------
      {
        uint32_t ahi[24], alo[24], araw[8];
#pragma unroll
        for (int i = 18; i < 24; ++i) { ahi[i] = 0u; alo[i] = 0u; }
        const float kHalfPi = 1.57079637050628662109375f;
        if (!big) {
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            float s, c;
            sincosf(__fmul_rn(pv[d], cfg.freq_factor), &s, &c);
#pragma unroll
            for (int i = 0; i < 6; ++i) {
              split_h2(s, c, ahi[d * 6 + i], alo[d * 6 + i]);
              store_split3(a2_col, (d * 6 + i) * 2 + 0, lb * s);
              store_split3(a2_col, (d * 6 + i) * 2 + 1, lb * c);
              const float s2 = 2.0f * s * c;
              c = (c - s) * (c + s);
              s = s2;
            }
          }
        } else {
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            float f = cfg.freq_factor;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
              const float s = sinf(__fmul_rn(pv[d], f)), c = sinf(fmaf(pv[d], f, kHalfPi));
              split_h2(s, c, ahi[d * 6 + i], alo[d * 6 + i]);
              store_split3(a2_col, (d * 6 + i) * 2 + 0, lb * s);
              store_split3(a2_col, (d * 6 + i) * 2 + 1, lb * c);
              f *= 2.0f;
            }
          }
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const float h = tf32_hi(pv[d]);
          araw[d] = big ? 0u : __float_as_uint(h);
          araw[3 + d] = big ? 0u : __float_as_uint(pv[d] - h);
          store_split3(a2_col, 36 + d, lb * pv[d]);
        }
        araw[6] = 0u; araw[7] = 0u;
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          uint32_t v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = ahi[g * 8 + e];
          tc::tmem_st_x8(t_lane + kColAhi + g * 8, v);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = alo[g * 8 + e];
          tc::tmem_st_x8(t_lane + kColAlo + g * 8, v);
        }
        tc::tmem_st_x8(t_lane + kColAraw, araw);
        tc::tmem_st_wait();
      }
      tc::tc_fence_before_sync();
      __syncthreads();
      if (warp_u == mma_warp && tc::elect_one()) {   // converged warp, one elected lane: operands stay in uniform registers
        tc::tc_fence_after_sync();
        const uint32_t sW = tc::smem_u32(smem);
#pragma unroll
        for (int ks = 0; ks < 3; ++ks) {
          const uint64_t bhi = tc::smem_desc_kmajor_noswizzle(sW + kOffW16hi + ks * 2 * HD * 16, HD * 16, 128);
          const uint64_t blo = tc::smem_desc_kmajor_noswizzle(sW + kOffW16lo + ks * 2 * HD * 16, HD * 16, 128);
          tc::mma_f16_ts(tmem + kColD, tmem + kColAhi + ks * 8, bhi, idesc16, ks > 0);
          tc::mma_f16_ts(tmem + kColD, tmem + kColAlo + ks * 8, bhi, idesc16, 1);
          tc::mma_f16_ts(tmem + kColD, tmem + kColAhi + ks * 8, blo, idesc16, 1);
        }
        tc::mma_tf32_ts(tmem + kColD, tmem + kColAraw, tc::smem_desc_kmajor_noswizzle(sW + kOffWrawA, HD * 16, 128), idesc32, 1);
        tc::mma_tf32_ts(tmem + kColD, tmem + kColAraw, tc::smem_desc_kmajor_noswizzle(sW + kOffWrawB, HD * 16, 128), idesc32, 1);
        tc::mma_commit(bar1);
      }
      __syncwarp();

      // ---------------- P2: bilinear gather of G (16 lanes per sample) ----------------
      {
        const int hw = lane / 16, q = lane % 16;
        const float4* G4 = reinterpret_cast<const float4*>(p.scene.G) + q;
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
          const int s = 2 * i + hw;
          const float4 a = w_taps[s * 2 + 0];
          const float4 w = w_taps[s * 2 + 1];
          float4 g;
          if (any_empty && __float_as_int(a.w)) {
            g = reinterpret_cast<const float4*>(s_gempty)[q];
          } else {
            const float4* g00 = G4 + (uint32_t)__float_as_int(a.x);
            const float4* g10 = g00 + __float_as_int(a.z);
            const float4 v00 = __ldg(g00), v01 = __ldg(g00 + __float_as_int(a.y));
            const float4 v10 = __ldg(g10), v11 = __ldg(g10 + __float_as_int(a.y));
            g.x = fmaf(v11.x, w.w, fmaf(v10.x, w.z, fmaf(v01.x, w.y, v00.x * w.x)));
            g.y = fmaf(v11.y, w.w, fmaf(v10.y, w.z, fmaf(v01.y, w.y, v00.y * w.x)));
            g.z = fmaf(v11.z, w.w, fmaf(v10.z, w.z, fmaf(v01.z, w.y, v00.z * w.x)));
            g.w = fmaf(v11.w, w.w, fmaf(v10.w, w.z, fmaf(v01.w, w.y, v00.w * w.x)));
          }
          *reinterpret_cast<float4*>(w_gi + s * kGiStride + q * 4) = g;
        }
      }
      __syncwarp();

      // ---------------- P3: h -> relu mask, d w_out, hbar scatter into dG, mask operand ----------------
      wait_bar(bar1, phase1);
      tc::tc_fence_after_sync();
      acc_bout += lb;
      {
        const float4 ta = w_taps[lane * 2 + 0];
        const float4 tw = w_taps[lane * 2 + 1];
        const bool is_empty = __float_as_int(ta.w) != 0;
        float4* d00 = reinterpret_cast<float4*>(bp.d_G) + (uint32_t)__float_as_int(ta.x);
        float4* d10 = d00 + __float_as_int(ta.z);
        const int dxo = __float_as_int(ta.y);
        const float4* grow = reinterpret_cast<const float4*>(w_gi + lane * kGiStride);
        const float4* wo4 = reinterpret_cast<const float4*>(s_wout);
        const float4* wrh = reinterpret_cast<const float4*>(smem + kOffWrawA);
        const float4* wrl = reinterpret_cast<const float4*>(smem + kOffWrawB);
        const __nv_bfloat16 one = __float2bfloat16_rn(1.0f), zero = __float2bfloat16_rn(0.0f);
        auto epilogue = [&](auto raw_tag) {
          constexpr bool kRaw = decltype(raw_tag)::value;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t dv[16];
            tc::tmem_ld_x16(t_lane + kColD + c * 16, dv);
            tc::tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float4 g = grow[c * 4 + e];
              const float4 wv = wo4[c * 4 + e];
              const float gv[4] = {g.x, g.y, g.z, g.w}, wo[4] = {wv.x, wv.y, wv.z, wv.w};
              float hb[4];
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const int u = c * 16 + e * 4 + t;
                float h = __uint_as_float(dv[e * 4 + t]) + gv[t];
                if (kRaw) {      // raw-coordinate part of lin_in in fp32 for far-outside samples (as the forward)
                  const float4 wh = wrh[u], wl = wrl[u];
                  if (big) h = fmaf(wh.z + wl.z, pv[2], fmaf(wh.y + wl.y, pv[1], fmaf(wh.x + wl.x, pv[0], h)));
                }
                const bool on = h > 0.0f;
                acc_wout[u] = fmaf(lb, fmaxf(h, 0.0f), acc_wout[u]);
                hb[t] = on ? lb * wo[t] : 0.0f;
                *reinterpret_cast<__nv_bfloat16*>(mask_col + u * 16) = on ? one : zero;
              }
              if (hb[0] != 0.0f || hb[1] != 0.0f || hb[2] != 0.0f || hb[3] != 0.0f) {
                const int q = c * 4 + e;
                if (is_empty) {
#pragma unroll
                  for (int t = 0; t < 4; ++t)
                    if (hb[t] != 0.0f) atomicAdd(bp.d_g_empty + q * 4 + t, hb[t]);
                } else {
                  atomicAdd(d00 + q, make_float4(hb[0] * tw.x, hb[1] * tw.x, hb[2] * tw.x, hb[3] * tw.x));
                  atomicAdd(d00 + dxo + q, make_float4(hb[0] * tw.y, hb[1] * tw.y, hb[2] * tw.y, hb[3] * tw.y));
                  atomicAdd(d10 + q, make_float4(hb[0] * tw.z, hb[1] * tw.z, hb[2] * tw.z, hb[3] * tw.z));
                  atomicAdd(d10 + dxo + q, make_float4(hb[0] * tw.w, hb[1] * tw.w, hb[2] * tw.w, hb[3] * tw.w));
                }
              }
            }
          }
        };
        if (__any_sync(0xffffffffu, big)) epilogue(std::true_type{});
        else epilogue(std::false_type{});
      }
      tc::fence_proxy_async_smem();        // operand tiles (P1: lbar*code, P3: mask) -> async proxy
      tc::tc_fence_before_sync();
      __syncthreads();
      if (warp_u == mma_warp && tc::elect_one()) {   // converged warp, one elected lane: operands stay in uniform registers
        tc::tc_fence_after_sync();
        const uint32_t sW = tc::smem_u32(smem);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t a2 = tc::smem_desc_kmajor_noswizzle(sW + kOffA2 + ks * 2 * kStrideA2, kStrideA2, 128);
          const uint64_t mk = tc::smem_desc_kmajor_noswizzle(sW + kOffMask + ks * 2 * kStrideMask, kStrideMask, 128);
          tc::mma_f16_ss(tmem + kColD2, a2, mk, idescbf, (d2_started || ks > 0) ? 1u : 0u);
        }
        tc::mma_commit(bar2);
      }
      d2_pending = true;
      d2_started = true;
      __syncwarp();
    }
    __syncwarp();
  }

  // ---- flush: d w_out, d b_out (warp reductions), d W_pe from the TMEM accumulator ----
  if (d2_pending) wait_bar(bar2, phase2);
  tc::tc_fence_after_sync();
#pragma unroll
  for (int u = 0; u < HD; ++u) {
    const float v = warp_sum(acc_wout[u]);
    if (lane == 0 && bp.d_w_out) atomicAdd(bp.d_w_out + u, v);
  }
  {
    const float v = warp_sum(acc_bout);
    if (lane == 0 && bp.d_b_out) atomicAdd(bp.d_b_out, v);
  }
  if (d2_started && bp.d_w_pe) {
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t dv[16];
      tc::tmem_ld_x16(t_lane + kColD2 + c * 16, dv);
      tc::tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 16; ++e) w_gi[lane * kGiStride + c * 16 + e] = __uint_as_float(dv[e]);
    }
    __syncthreads();
    auto row = [&](int r) { return reinterpret_cast<const float*>(smem + kOffWarp + (r >> 5) * warp_bytes(Kpad) + kWarpGi) + (r & 31) * kGiStride; };
    for (int i = tid; i < 39 * HD; i += kThreads) {
      const int j = i / HD, u = i % HD;
      const float v = row(j)[u] + row(40 + j)[u] + row(80 + j)[u];
      int ref;
      if (j < 36) {
        const int sc = j & 1, di = j >> 1, d = di / 6, oct = di % 6;
        ref = 3 + oct * 6 + sc * 3 + d;
      } else {
        ref = j - 36;
      }
      atomicAdd(bp.d_w_pe + (size_t)u * P + ref, v * s_wout[u]);
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<kTmemCols>(tmem);
}

}  // namespace bwdtc

bool bwd_tc_supported(const BwdParams& bp) {
  const KernelParams& p = bp.k;
  const BtsRenderCfg& c = p.cfg;
  const int64_t pixels = (int64_t)p.scene.n * p.scene.H * p.scene.W;
  return bp.saved_sigma != nullptr && p.mlp.Hd == 64 && p.mlp.n_blocks == 0 && p.mlp.P == 39 && c.num_freqs == 6 && c.include_input &&
         c.n_coarse % 32 == 0 && c.n_coarse <= 256 && p.scene.Vc <= kMaxColorViews && p.scene.Vc * p.scene.Cc <= 32 &&
         pixels * 16 < (int64_t)0x7fffffff && (reinterpret_cast<uintptr_t>(bp.d_G) & 15) == 0 && p.rays != nullptr;
}

cudaError_t launch_bwd_tc(const BwdParams& bp, cudaStream_t stream) {
  const int Kpad = (bp.k.cfg.n_coarse + 31) & ~31;
  const size_t smem = (size_t)bwdtc::kOffWarp + 4 * (size_t)bwdtc::warp_bytes(Kpad);
  auto kern = bwdtc::render_bwd_tc_kernel;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // two CTAs per SM (256 TMEM columns each) when two footprints fit the 228 KB of an SM; the occupancy query would
  // answer for the default carve-out, not the one requested above
  const int occ = 2 * (smem + 1024) <= 233472 ? 2 : 1;
  const int64_t want = (bp.k.n_items + 3) / 4;
  const int64_t cap = (int64_t)sms * occ;
  const int grid = (int)(want < cap ? (want > 0 ? want : 1) : cap);
  kern<<<grid, bwdtc::kThreads, smem, stream>>>(bp);
  return cudaGetLastError();
}

int bwd_tc_barrier_timeouts() {
  int v = 0;
  cudaMemcpyFromSymbol(&v, g_bwd_tc_barrier_timeouts, sizeof(int));
  return v;
}

}  // namespace bts
