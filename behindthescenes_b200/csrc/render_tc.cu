// Tensor-core render kernel for the BTS configurations without ResnetFC blocks (d_hidden = 64 -- the KITTI configs --
// or 128 -- the BASELINE stress shape; n_blocks = 0, 6-octave code with include_input, K a multiple of 32):
// the positional-code part of lin_in runs on the 5th-gen tensor cores
// (tcgen05.mma, accumulator in TMEM), everything else on the CUDA cores, all fused per ray.
//
// CTA = 4 warps; each warp owns one ray per pass; a pass is processed in K/32 sub-tiles of 128 samples
// (row r of the MMA tile = lane r%32 of warp r/32 = sample 32*j + lane of that warp's ray):
//   P1  lane = sample : stratified depth, world point, projection into the encoder view, depth code, positional code
//                       -> A operand written straight into TMEM (tcgen05.st): 36 sin/cos values as fp16 hi + fp16 lo
//                       (two-term split = fp32-grade products), the 3 raw coordinates as tf32 hi/lo (unbounded range)
//   MMA one thread    : D[128x64] = A16hi*W16hi + A16lo*W16hi + A16hi*W16lo (kind::f16, K=48)
//                                   + Araw*WrawA + Araw*WrawB            (kind::tf32, K=8), fp32 accumulate in TMEM
//   P2  half-warp = sample : coalesced bilinear gather of the folded map G (4 taps x 64 channels = 4 x 256 B per
//                       sample, one 16-byte load per lane and tap) -> interpolated row in shared memory
//   P3  lane = sample : h = D (tcgen05.ld) + G-row ; sigma = softplus(w_out . relu(h) + b_out)
//   C/D lane = sample : alpha compositing scan, colour views (render_phases.cuh)
// The fast tier (cfg.precision = 1) drops the lo terms (single fp16 / tf32 pass).
//
// Reference arithmetic: models/bts/model/models_bts.py:138-216 (sample_features), :266-338 (forward),
// models/common/model/code.py:30-42, resnetfc.py:132-184 (n_blocks = 0), models/common/render/nerf.py:103-123, :210-313.
#include <cuda_fp16.h>

#include "common.cuh"
#include "render_phases.cuh"
#include "tcgen05.cuh"

namespace bts {

__device__ int g_tc_barrier_timeouts = 0;   // bug guard: bounded mbarrier waits that expired (read by tests)

namespace tcr {

constexpr int kThreads = 128;

// Per-shape layout: TMEM columns [D: HD][A: code operand hi 24 | lo 24 | raw 8; reused as the ResnetFC block operand hi HD/2 |
// lo HD/2], shared-memory map (bytes).  NB = number of ResnetFC blocks run on the tensor cores (0 | 1).
template <int HD, int NB>
struct L {
  static_assert(HD == 32 || HD == 64 || HD == 128, "tensor-core kernel: d_hidden 32, 64 or 128");
  static_assert(NB == 0 || (NB == 1 && HD <= 64), "ResnetFC block on the tensor cores: one block, d_hidden <= 64");
  static constexpr uint32_t kColD = 0, kColAhi = HD, kColAlo = HD + 24, kColAraw = HD + 48;
  static constexpr uint32_t kColBhi = HD, kColBlo = HD + HD / 2;            // block operand (relu'd activations, fp16 hi / lo)
  static constexpr uint32_t kColsUsed = HD + (NB && HD > 56 ? HD : 56);
  static constexpr uint32_t kTmemCols = kColsUsed <= 128 ? 128 : 256;
  // 512 TMEM columns per SM; the 64-wide block variant keeps 64 activations per thread in registers -> 2 CTAs (255 registers)
  static constexpr int kCtasPerSm = (HD == 128 || (HD == 64 && NB == 1)) ? 2 : 4;
  static constexpr int kLanesPerSample = HD / 4;               // one float4 (4 channels) per lane in the gather
  static constexpr int kSamplesPerIter = 32 / kLanesPerSample;
  static constexpr int kGiStride = HD + 4;                     // floats per interpolated row (+4: conflict-free both ways)
  static constexpr int kOffW16hi = 0;                          // [6 chunks][HD rows][16 B]  fp16, K-major no-swizzle
  static constexpr int kOffW16lo = kOffW16hi + 6 * HD * 16;
  static constexpr int kOffWrawA = kOffW16lo + 6 * HD * 16;    // [2][HD][16 B] tf32: [hi0 hi1 hi2 hi0 | hi1 hi2 0 0]
  static constexpr int kOffWrawB = kOffWrawA + 2 * HD * 16;    //               tf32: [lo0 lo1 lo2 0 | 0 0 0 0]
  static constexpr int kOffWout = kOffWrawB + 2 * HD * 16;     // HD floats
  static constexpr int kOffGempty = kOffWout + HD * 4;         // HD floats
  static constexpr int kBlkPlane = 2 * HD * HD;                // one fp16 plane of a HD x HD block matrix: [HD/8 chunks][HD rows][16 B]
  static constexpr int kOffBlk = kOffGempty + HD * 4;          // [fc_0 | fc_1][hi | lo] planes, then the two bias vectors
  static constexpr int kOffBlkBias = kOffBlk + NB * 4 * kBlkPlane;
  static constexpr int kOffBar = kOffBlkBias + NB * 2 * HD * 4;   // mbarrier (8 B) + tmem base (4 B) + pad
  static constexpr int kOffWarp = kOffBar + 16;
  static constexpr int kWarpTaps = 0;                          // [32][8] floats
  static constexpr int kWarpGi = 32 * 8 * 4;                   // [32][HD + 4] floats
  static constexpr int kWarpRgb = kWarpGi + 32 * kGiStride * 4;
  static constexpr int kWarpBytes = kWarpRgb + (kMaxColorViews * 4 + kMaxColorViews) * 4;
  static constexpr int kSmemBytes = kOffWarp + 4 * kWarpBytes;
  static_assert(kCtasPerSm * (kSmemBytes + 1024) <= 233472, "resident CTAs must fit the 228 KB shared memory of an SM");
};

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// (s, c) -> fp16 pair hi and the fp16 pair of the residuals
__device__ __forceinline__ void split_h2(float s, float c, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(s, c);
  const float2 hf = __half22float2(h);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = pack_h2(s - hf.x, c - hf.y);
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// kPoints: point-query entry (BTSNet.forward on explicit xyz, models_bts.py:266-338): a pass is 128 points (lane = point),
// the same P1/MMA/P2/P3, no compositing.
// NB = 1: one ResnetFC block (models/common/model/resnetfc.py:53-62: h + fc_1(relu(fc_0(relu(h))))) as two more MMA rounds
// on the accumulator: relu'd activations -> fp16 hi/lo -> TMEM operand (the code operand's columns) -> D; the RealEstate10K
// field (configs/exp_re10k.yaml:43-44: d_hidden 32, n_blocks 1).
template <int HD, bool kExact, bool kPoints, int NB>
__global__ void __launch_bounds__(kThreads, L<HD, NB>::kCtasPerSm) render_tc_kernel(const KernelParams p) {
  using Lt = L<HD, NB>;
  constexpr uint32_t kColD = Lt::kColD, kColAhi = Lt::kColAhi, kColAlo = Lt::kColAlo, kColAraw = Lt::kColAraw, kTmemCols = Lt::kTmemCols;
  constexpr int kOffW16hi = Lt::kOffW16hi, kOffW16lo = Lt::kOffW16lo, kOffWrawA = Lt::kOffWrawA, kOffWrawB = Lt::kOffWrawB;
  constexpr int kOffWout = Lt::kOffWout, kOffGempty = Lt::kOffGempty, kOffBar = Lt::kOffBar, kOffWarp = Lt::kOffWarp;
  constexpr int kWarpTaps = Lt::kWarpTaps, kWarpGi = Lt::kWarpGi, kWarpRgb = Lt::kWarpRgb, kWarpBytes = Lt::kWarpBytes;
  constexpr int kGiStride = Lt::kGiStride, kLPS = Lt::kLanesPerSample, kSPI = Lt::kSamplesPerIter;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // the MMA-issuing warp differs between the CTAs that share an SM (warp w runs on scheduler w % 4)
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const int mma_warp = (int)(blockIdx.x / tc::num_sms()) & 3;
  // work-list mode: render only the 8-ray groups the band kernel (render_band.cu) handed over; one group = 2 passes
  const int work_groups = (!kPoints && p.work_mode) ? *reinterpret_cast<const volatile int32_t*>(p.work) : 0;
  if (!kPoints && p.work_mode && (int)blockIdx.x >= 2 * work_groups) return;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + 8);
  float* s_wout = reinterpret_cast<float*>(smem + kOffWout);
  float* s_gempty = reinterpret_cast<float*>(smem + kOffGempty);
  unsigned char* wbase = smem + kOffWarp + warp * kWarpBytes;
  float4* w_taps = reinterpret_cast<float4*>(wbase + kWarpTaps);
  float* w_gi = reinterpret_cast<float*>(wbase + kWarpGi);
  float* w_rgb = reinterpret_cast<float*>(wbase + kWarpRgb);

  const BtsRenderCfg& cfg = p.cfg;
  const int C = p.mlp.C, ld = C + p.mlp.P;

  // ---- one-time setup: weights -> B operands (fp16 hi/lo, tf32 hi/lo), TMEM allocation, mbarrier ----
  for (int i = tid; i < HD * 48; i += kThreads) {
    const int n = i / 48, c = i % 48;
    float w = 0.0f;
    if (c < 36) {   // A column c = (d*6 + i)*2 + sc  <->  reference code index 3 + i*6 + sc*3 + d (code.py:39-41)
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
  if constexpr (NB == 1) {   // fc_0 / fc_1 (nn.Linear weight [out][in]) as B operands: row n = output unit, K index = input unit
    for (int i = tid; i < 2 * HD * HD; i += kThreads) {
      const int m = i / (HD * HD), n = (i / HD) % HD, c = i % HD;
      const float w = __ldg((m == 0 ? p.mlp.fc0_w[0] : p.mlp.fc1_w[0]) + (size_t)n * HD + c);
      const __half hi = __float2half_rn(w);
      const __half lo = __float2half_rn(w - __half2float(hi));
      const int off = ((c >> 3) * HD + n) * 16 + (c & 7) * 2;
      *reinterpret_cast<__half*>(smem + Lt::kOffBlk + (2 * m + 0) * Lt::kBlkPlane + off) = hi;
      *reinterpret_cast<__half*>(smem + Lt::kOffBlk + (2 * m + 1) * Lt::kBlkPlane + off) = lo;
    }
    for (int i = tid; i < 2 * HD; i += kThreads)
      reinterpret_cast<float*>(smem + Lt::kOffBlkBias)[i] = __ldg((i < HD ? p.mlp.fc0_b[0] : p.mlp.fc1_b[0]) + (i % HD));
  }
  if (warp == 0) tc::tmem_alloc<kTmemCols>(tmem_slot);
  if (tid == 0) {
    tc::mbar_init(bar, 1);
    tc::fence_mbar_init();
  }
  tc::fence_proxy_async_smem();          // weight tiles are read by the tensor core through the async proxy
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);   // this warp's 32 TMEM lanes
  uint32_t bar_phase = 0;

  const int K = cfg.n_coarse;
  const int H = p.scene.H, W = p.scene.W;
  const float b_out = __ldg(p.mlp.b_out);
  const bool use_empty = cfg.learn_empty && p.scene.g_empty != nullptr;
  const uint32_t idesc16 = tc::instr_desc(tc::kFmtF16, 128, HD), idesc32 = tc::instr_desc(tc::kFmtTF32, 128, HD);

  const int64_t n_rays = p.n_items;                       // rays, or points in the query variant
  const int64_t n_pass = kPoints ? (n_rays + 127) / 128 : ((!kPoints && p.work_mode) ? 2 * (int64_t)work_groups : (n_rays + 3) / 4);
  // first ray of a pass (4 consecutive rays)
  auto pass_base = [&](int64_t pass) -> int64_t {
    if (!kPoints && p.work_mode) return (int64_t)__ldg(p.work + 4 + (pass >> 1)) * 8 + (pass & 1) * 4;
    return pass * 4;
  };
  View vf;
  load_view(vf, p.scene.K_f, p.scene.w2c_f);          // batch element 0; reloaded per ray only for batched scenes
  // the ray tensor row of the next pass is loaded one pass ahead (its latency hides behind a whole pass)
  auto ray_of = [&](int64_t pass) { const int64_t r = pass_base(pass) + warp; return r < n_rays ? r : n_rays - 1; };
  float rv_next = (!kPoints && p.rays && lane < 8 && (int64_t)blockIdx.x < n_pass) ? __ldg(p.rays + ray_of(blockIdx.x) * 8 + lane) : 0.0f;
  for (int64_t pass = blockIdx.x; pass < n_pass; pass += gridDim.x) {
    // ---- per-pass setup: one ray per warp (render) or one point per lane (query) ----
    bool active;
    int64_t ray = 0, b = 0;
    float ro0 = 0, ro1 = 0, ro2 = 0, rd0 = 0, rd1 = 0, rd2 = 0, near_ = 0, far_ = 0, zfac = 1.0f;
    float PX = 0, PY = 0, PZ = 0;
    if constexpr (!kPoints) {
      const int64_t ray_raw = pass_base(pass) + warp;
      active = ray_raw < n_rays;
      ray = active ? ray_raw : n_rays - 1;
      b = p.scene.n == 1 ? 0 : ray / p.per_batch;   // 64-bit division only for batched scenes
      const float rv = rv_next;
      if (p.rays && lane < 8 && pass + gridDim.x < n_pass) rv_next = __ldg(p.rays + ray_of(pass + gridDim.x) * 8 + lane);
      const Ray ry = p.rays ? ray_from_lanes(rv) : fetch_ray(p, ray, lane);
      ro0 = ry.o[0]; ro1 = ry.o[1]; ro2 = ry.o[2]; rd0 = ry.d[0]; rd1 = ry.d[1]; rd2 = ry.d[2];
      zfac = ry.zfac;
      depth_bounds(p, ry, near_, far_);
    } else {
      const int64_t idx_raw = pass * 128 + warp * 32 + lane;
      active = idx_raw < n_rays;
      ray = active ? idx_raw : n_rays - 1;          // "ray" = flat point index in this variant
      b = p.scene.n == 1 ? 0 : ray / p.per_batch;
      PX = __ldg(p.xyz + ray * 3 + 0); PY = __ldg(p.xyz + ray * 3 + 1); PZ = __ldg(p.xyz + ray * 3 + 2);
    }
    float T_carry = 1.0f, depth_part = 0.0f, wsum_part = 0.0f;
    for (int i = lane; i < kMaxColorViews * 5; i += 32) w_rgb[i] = 0.0f;
    if (p.scene.n != 1) load_view(vf, p.scene.K_f + b * 9, p.scene.w2c_f + b * 16);

    float z_ahead = kPoints ? 0.0f : sample_depth(p, ray, (lane < K ? lane : K - 1), near_, far_);   // depths are produced one chunk ahead
    const int Kloop = kPoints ? 32 : K;
    for (int c0 = 0; c0 < Kloop; c0 += 32) {
      const int k_raw = c0 + lane;
      const bool in_k = kPoints || k_raw < K;            // ragged last chunk (K % 32 != 0): lanes past K idle along
      const int k = in_k ? k_raw : K - 1;
      const int64_t sample_id = kPoints ? ray : ray * K + k;
      // ---------------- P1: per-sample scalar work, A operand -> TMEM ----------------
      float z = 0.0f, z_next = 0.0f, X, Y, Z;
      if constexpr (!kPoints) {
        z = z_ahead;
        z_ahead = (c0 + 32 < K) ? sample_depth(p, ray, (k_raw + 32 < K ? k_raw + 32 : K - 1), near_, far_) : 0.0f;
        z_next = __shfl_down_sync(0xffffffffu, z, 1);
        const float z_first_of_next = __shfl_sync(0xffffffffu, z_ahead, 0);
        if (lane == 31) z_next = z_first_of_next;
        X = __fadd_rn(ro0, __fmul_rn(z, rd0));   // nerf.py:231
        Y = __fadd_rn(ro1, __fmul_rn(z, rd1));
        Z = __fadd_rn(ro2, __fmul_rn(z, rd2));
      } else {
        X = PX; Y = PY; Z = PZ;
      }
      const Proj pf = project(vf, X, Y, Z);
      const bool inv_f = pf.invalid;
      const float pv[3] = {pf.u, pf.v, depth_code(cfg, pf)};
      bool any_empty, same_block;
      {
        const Taps t = bilinear_taps(pf.u, pf.v, H, W);
        const int p00 = (int)((b * H + t.y0) * (int64_t)W + t.x0);
        float4 a, w;
        a.x = __int_as_float(p00 * (HD / 4));           // in float4 units of an HD-channel pixel
        a.y = __int_as_float(t.dx * (HD / 4));
        a.z = __int_as_float(t.dy * W * (HD / 4));
        a.w = __int_as_float((use_empty && inv_f) ? 1 : 0);
        any_empty = __any_sync(0xffffffffu, use_empty && inv_f);
        // all 32 samples interpolate inside the same 2x2 pixel block (rays through the encoder camera centre: every
        // sample of the ray projects to the same pixel, only the ulp-level weights differ) -> fetch the block once
        const int key = p00 * 4 + t.dx * 2 + t.dy;
        same_block = __all_sync(0xffffffffu, key == __shfl_sync(0xffffffffu, key, 0));
        w.x = t.wx0 * t.wy0; w.y = t.wx1 * t.wy0; w.z = t.wx0 * t.wy1; w.w = t.wx1 * t.wy1;
        w_taps[lane * 2 + 0] = a;
        w_taps[lane * 2 + 1] = w;
      }
      const bool big = fmaxf(fabsf(pv[0]), fmaxf(fabsf(pv[1]), fabsf(pv[2]))) > kPeFastLimit;
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
              if (kExact) split_h2(s, c, ahi[d * 6 + i], alo[d * 6 + i]);
              else ahi[d * 6 + i] = pack_h2(s, c);
              const float s2 = 2.0f * s * c;
              c = (c - s) * (c + s);
              s = s2;
            }
          }
        } else {   // far outside the frustum: evaluate every term on the reference's own rounded argument
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            float f = cfg.freq_factor;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
              const float s = sinf(__fmul_rn(pv[d], f)), c = sinf(fmaf(pv[d], f, kHalfPi));
              if (kExact) split_h2(s, c, ahi[d * 6 + i], alo[d * 6 + i]);
              else ahi[d * 6 + i] = pack_h2(s, c);
              f *= 2.0f;
            }
          }
        }
        // raw coordinates: tf32 hi/lo through the tensor core when they are O(1); far-outside samples (|p| in the
        // thousands, clamped z) would lose absolute accuracy in a two-term split, they are added in fp32 in P3 instead
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const float h = tf32_hi(pv[d]);
          araw[d] = big ? 0u : __float_as_uint(h);
          araw[3 + d] = (kExact && !big) ? __float_as_uint(pv[d] - h) : 0u;
        }
        araw[6] = 0u; araw[7] = 0u;
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          uint32_t v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = ahi[g * 8 + e];
          tc::tmem_st_x8(t_lane + kColAhi + g * 8, v);
          if (kExact) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = alo[g * 8 + e];
            tc::tmem_st_x8(t_lane + kColAlo + g * 8, v);
          }
        }
        tc::tmem_st_x8(t_lane + kColAraw, araw);
        tc::tmem_st_wait();
      }
      tc::tc_fence_before_sync();
      __syncthreads();                     // all 128 rows of the A tile are in TMEM; D of the previous tile is consumed

      // ---------------- MMA issue (one thread) ----------------
      if (warp_u == mma_warp && tc::elect_one()) {   // converged warp, one elected lane: operands stay in uniform registers
        tc::tc_fence_after_sync();
        const uint32_t sW = tc::smem_u32(smem);
#pragma unroll
        for (int ks = 0; ks < 3; ++ks) {
          const uint64_t bhi = tc::smem_desc_kmajor_noswizzle(sW + kOffW16hi + ks * 2 * HD * 16, HD * 16, 128);
          tc::mma_f16_ts(tmem + kColD, tmem + kColAhi + ks * 8, bhi, idesc16, ks > 0);
          if (kExact) {
            const uint64_t blo = tc::smem_desc_kmajor_noswizzle(sW + kOffW16lo + ks * 2 * HD * 16, HD * 16, 128);
            tc::mma_f16_ts(tmem + kColD, tmem + kColAlo + ks * 8, bhi, idesc16, 1);
            tc::mma_f16_ts(tmem + kColD, tmem + kColAhi + ks * 8, blo, idesc16, 1);
          }
        }
        tc::mma_tf32_ts(tmem + kColD, tmem + kColAraw, tc::smem_desc_kmajor_noswizzle(sW + kOffWrawA, HD * 16, 128), idesc32, 1);
        if (kExact)
          tc::mma_tf32_ts(tmem + kColD, tmem + kColAraw, tc::smem_desc_kmajor_noswizzle(sW + kOffWrawB, HD * 16, 128), idesc32, 1);
        tc::mma_commit(bar);
      }
      __syncwarp();

      // ---------------- P2: coalesced bilinear gather of G: HD/4 lanes per sample, one 16-byte load per lane and tap ----
      {
        const int hw = lane / kLPS, q = lane % kLPS;
        const float4* G4 = reinterpret_cast<const float4*>(p.scene.G) + q;
        constexpr int kIters = 32 / kSPI;
        if (same_block && !any_empty) {
          const float4 a = w_taps[0];
          const float4* g00 = G4 + (uint32_t)__float_as_int(a.x);
          const float4* g10 = g00 + __float_as_int(a.z);
          const float4 v00 = __ldg(g00), v01 = __ldg(g00 + __float_as_int(a.y));
          const float4 v10 = __ldg(g10), v11 = __ldg(g10 + __float_as_int(a.y));
#pragma unroll 4
          for (int i = 0; i < kIters; ++i) {
            const int s = kSPI * i + hw;
            const float4 w = w_taps[s * 2 + 1];
            float4 g;
            g.x = fmaf(v11.x, w.w, fmaf(v10.x, w.z, fmaf(v01.x, w.y, v00.x * w.x)));
            g.y = fmaf(v11.y, w.w, fmaf(v10.y, w.z, fmaf(v01.y, w.y, v00.y * w.x)));
            g.z = fmaf(v11.z, w.w, fmaf(v10.z, w.z, fmaf(v01.z, w.y, v00.z * w.x)));
            g.w = fmaf(v11.w, w.w, fmaf(v10.w, w.z, fmaf(v01.w, w.y, v00.w * w.x)));
            *reinterpret_cast<float4*>(w_gi + s * kGiStride + q * 4) = g;
          }
        } else if (!any_empty) {
#pragma unroll 4
          for (int i = 0; i < kIters; ++i) {
            const int s = kSPI * i + hw;
            const float4 a = w_taps[s * 2 + 0];
            const float4 w = w_taps[s * 2 + 1];
            const float4* g00 = G4 + (uint32_t)__float_as_int(a.x);
            const float4* g10 = g00 + __float_as_int(a.z);
            const float4 v00 = __ldg(g00), v01 = __ldg(g00 + __float_as_int(a.y));
            const float4 v10 = __ldg(g10), v11 = __ldg(g10 + __float_as_int(a.y));
            float4 g;
            g.x = fmaf(v11.x, w.w, fmaf(v10.x, w.z, fmaf(v01.x, w.y, v00.x * w.x)));
            g.y = fmaf(v11.y, w.w, fmaf(v10.y, w.z, fmaf(v01.y, w.y, v00.y * w.x)));
            g.z = fmaf(v11.z, w.w, fmaf(v10.z, w.z, fmaf(v01.z, w.y, v00.z * w.x)));
            g.w = fmaf(v11.w, w.w, fmaf(v10.w, w.z, fmaf(v01.w, w.y, v00.w * w.x)));
            *reinterpret_cast<float4*>(w_gi + s * kGiStride + q * 4) = g;
          }
        } else {   // some samples of this chunk left the encoder frustum: they take the learned empty row
#pragma unroll 2
          for (int i = 0; i < kIters; ++i) {
            const int s = kSPI * i + hw;
            const float4 a = w_taps[s * 2 + 0];
            const float4 w = w_taps[s * 2 + 1];
            float4 g;
            if (__float_as_int(a.w)) {
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
      }
      __syncwarp();

      // ---------------- P3: accumulator + interpolated row -> sigma ----------------
      // bounded wait (bug guard): on expiry the kernel traps -- the launch fails with a CUDA error instead of hanging
      // the GPU or reading an unfinished accumulator
      tc::mbar_wait_or_trap(bar, bar_phase, &g_tc_barrier_timeouts);
      bar_phase ^= 1u;
      tc::tc_fence_after_sync();
      float logit = b_out;
      if constexpr (NB == 0 && HD >= 64) {
  #pragma unroll
        for (int grp = 0; grp < HD / 64; ++grp) {      // 64 accumulator columns at a time (register budget)
          uint32_t dv[4][16];
  #pragma unroll
          for (int c = 0; c < 4; ++c) tc::tmem_ld_x16(t_lane + kColD + grp * 64 + c * 16, dv[c]);
          tc::tmem_ld_wait();
          const float4* grow = reinterpret_cast<const float4*>(w_gi + lane * kGiStride + grp * 64);
          const float4* wo4 = reinterpret_cast<const float4*>(s_wout + grp * 64);
          if (!big) {
            float l0 = 0.0f, l1 = 0.0f, l2 = 0.0f, l3 = 0.0f;   // four independent chains (ILP)
  #pragma unroll
            for (int c = 0; c < 4; ++c) {
  #pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float4 g = grow[c * 4 + e];
                const float4 wv = wo4[c * 4 + e];
                l0 = fmaf(fmaxf(__uint_as_float(dv[c][e * 4 + 0]) + g.x, 0.0f), wv.x, l0);
                l1 = fmaf(fmaxf(__uint_as_float(dv[c][e * 4 + 1]) + g.y, 0.0f), wv.y, l1);
                l2 = fmaf(fmaxf(__uint_as_float(dv[c][e * 4 + 2]) + g.z, 0.0f), wv.z, l2);
                l3 = fmaf(fmaxf(__uint_as_float(dv[c][e * 4 + 3]) + g.w, 0.0f), wv.w, l3);
              }
            }
            logit += (l0 + l1) + (l2 + l3);
          } else {   // rare: raw-coordinate part of lin_in in fp32
            const float4* wrh = reinterpret_cast<const float4*>(smem + kOffWrawA) + grp * 64;   // chunk 0: [hi0 hi1 hi2 hi0]
            const float4* wrl = reinterpret_cast<const float4*>(smem + kOffWrawB) + grp * 64;   // chunk 0: [lo0 lo1 lo2 0]; hi+lo = fp32 weight
  #pragma unroll
            for (int c = 0; c < 4; ++c) {
  #pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float4 g = grow[c * 4 + e];
                const float4 wv = wo4[c * 4 + e];
                const float gv[4] = {g.x, g.y, g.z, g.w}, wo[4] = {wv.x, wv.y, wv.z, wv.w};
  #pragma unroll
                for (int t = 0; t < 4; ++t) {
                  const float4 wh = wrh[c * 16 + e * 4 + t], wl = wrl[c * 16 + e * 4 + t];
                  float h = __uint_as_float(dv[c][e * 4 + t]) + gv[t];
                  h = fmaf(wh.z + wl.z, pv[2], fmaf(wh.y + wl.y, pv[1], fmaf(wh.x + wl.x, pv[0], h)));
                  logit = fmaf(fmaxf(h, 0.0f), wo[t], logit);
                }
              }
            }
          }
        }
      } else {
        // d_hidden 32 and / or one ResnetFC block: the whole pre-activation row in registers
        float h[HD];
        {
          const float4* wrh = reinterpret_cast<const float4*>(smem + kOffWrawA);   // chunk 0: [hi0 hi1 hi2 hi0]
          const float4* wrl = reinterpret_cast<const float4*>(smem + kOffWrawB);   // chunk 0: [lo0 lo1 lo2 0]
#pragma unroll
          for (int c = 0; c < HD / 16; ++c) {
            uint32_t dv[16];
            tc::tmem_ld_x16(t_lane + kColD + c * 16, dv);
            tc::tmem_ld_wait();
            const float4* grow = reinterpret_cast<const float4*>(w_gi + lane * kGiStride + c * 16);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float4 g = grow[e];
              h[c * 16 + e * 4 + 0] = __uint_as_float(dv[e * 4 + 0]) + g.x;
              h[c * 16 + e * 4 + 1] = __uint_as_float(dv[e * 4 + 1]) + g.y;
              h[c * 16 + e * 4 + 2] = __uint_as_float(dv[e * 4 + 2]) + g.z;
              h[c * 16 + e * 4 + 3] = __uint_as_float(dv[e * 4 + 3]) + g.w;
            }
            if (big) {   // rare: raw-coordinate part of lin_in in fp32
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const float4 wh = wrh[c * 16 + e], wl = wrl[c * 16 + e];
                h[c * 16 + e] = fmaf(wh.z + wl.z, pv[2], fmaf(wh.y + wl.y, pv[1], fmaf(wh.x + wl.x, pv[0], h[c * 16 + e])));
              }
            }
          }
        }
        if constexpr (NB == 1) {
          const float* s_bias = reinterpret_cast<const float*>(smem + Lt::kOffBlkBias);
          const uint32_t idesc_blk = tc::instr_desc(tc::kFmtF16, 128, HD);
          // x -> relu(x) as fp16 hi/lo pairs -> the block operand in TMEM (over the code operand: its MMAs are complete)
          auto put_operand = [&](const float (&x)[HD]) {
#pragma unroll
            for (int g = 0; g < HD / 16; ++g) {
              uint32_t vh[8], vl[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float a0 = fmaxf(x[g * 16 + 2 * e], 0.0f), a1 = fmaxf(x[g * 16 + 2 * e + 1], 0.0f);
                vl[e] = 0u;
                if (kExact) split_h2(a0, a1, vh[e], vl[e]);
                else vh[e] = pack_h2(a0, a1);
              }
              tc::tmem_st_x8(t_lane + Lt::kColBhi + g * 8, vh);
              if (kExact) tc::tmem_st_x8(t_lane + Lt::kColBlo + g * 8, vl);
            }
            tc::tmem_st_wait();
          };
          // D[128 x HD] = operand x W^T of block matrix m (0: fc_0, 1: fc_1); every thread has finished with D before
          auto block_mma = [&](int m) {
            tc::tc_fence_before_sync();
            __syncthreads();
            if (warp_u == mma_warp && tc::elect_one()) {   // converged warp, one elected lane: operands stay in uniform registers
              tc::tc_fence_after_sync();
              const uint32_t sB = tc::smem_u32(smem) + Lt::kOffBlk + 2 * m * Lt::kBlkPlane;
#pragma unroll
              for (int ks = 0; ks < HD / 16; ++ks) {
                const uint64_t bhi = tc::smem_desc_kmajor_noswizzle(sB + ks * 2 * HD * 16, HD * 16, 128);
                tc::mma_f16_ts(tmem + kColD, tmem + Lt::kColBhi + ks * 8, bhi, idesc_blk, ks > 0);
                if (kExact) {
                  const uint64_t blo = tc::smem_desc_kmajor_noswizzle(sB + Lt::kBlkPlane + ks * 2 * HD * 16, HD * 16, 128);
                  tc::mma_f16_ts(tmem + kColD, tmem + Lt::kColBlo + ks * 8, bhi, idesc_blk, 1);
                  tc::mma_f16_ts(tmem + kColD, tmem + Lt::kColBhi + ks * 8, blo, idesc_blk, 1);
                }
              }
              tc::mma_commit(bar);
            }
            __syncwarp();
            tc::mbar_wait_or_trap(bar, bar_phase, &g_tc_barrier_timeouts);
            bar_phase ^= 1u;
            tc::tc_fence_after_sync();
          };
          put_operand(h);
          block_mma(0);                              // net = fc_0(relu(h))
          {
            float t[HD];
#pragma unroll
            for (int c = 0; c < HD / 16; ++c) {
              uint32_t dv[16];
              tc::tmem_ld_x16(t_lane + kColD + c * 16, dv);
              tc::tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 16; ++e) t[c * 16 + e] = __uint_as_float(dv[e]) + s_bias[c * 16 + e];
            }
            put_operand(t);
          }
          block_mma(1);                              // dx = fc_1(relu(net))
#pragma unroll
          for (int c = 0; c < HD / 16; ++c) {
            uint32_t dv[16];
            tc::tmem_ld_x16(t_lane + kColD + c * 16, dv);
            tc::tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) h[c * 16 + e] += __uint_as_float(dv[e]) + s_bias[HD + c * 16 + e];   // h + dx (identity shortcut)
          }
        }
        float l0 = 0.0f, l1 = 0.0f, l2 = 0.0f, l3 = 0.0f;
#pragma unroll
        for (int j = 0; j < HD; j += 4) {
          const float4 wv = *reinterpret_cast<const float4*>(s_wout + j);
          l0 = fmaf(fmaxf(h[j + 0], 0.0f), wv.x, l0);
          l1 = fmaf(fmaxf(h[j + 1], 0.0f), wv.y, l1);
          l2 = fmaf(fmaxf(h[j + 2], 0.0f), wv.z, l2);
          l3 = fmaf(fmaxf(h[j + 3], 0.0f), wv.w, l3);
        }
        logit += (l0 + l1) + (l2 + l3);
      }
      tc::tc_fence_before_sync();          // my TMEM reads are ordered before the next tile's barrier / MMA
      const float sigma = softplus(logit);
      __syncwarp();                        // w_taps / w_gi are rewritten by the next sub-tile

      // ---------------- C, D ----------------
      if constexpr (!kPoints) {
        const bool valid = active && in_k;
        const float wgt = composite_chunk(p, lane, k_raw, K, valid, z, z_next, sigma, sample_id, T_carry, depth_part, wsum_part);
        colour_views<false>(p, lane, b, X, Y, Z, valid, pf, wgt, sample_id, w_rgb);
      } else {
        if (active) {
          p.q_sigma[sample_id] = sigma;
          p.q_invalid_f[sample_id] = inv_f ? 1.0f : 0.0f;
        }
        colour_views<true>(p, lane, b, X, Y, Z, active, pf, 0.0f, sample_id, w_rgb);
      }
      __syncwarp();
    }
    if constexpr (!kPoints) {
      if (active) finish_ray(p, lane, ray, depth_part, wsum_part, w_rgb, zfac);
    }
    __syncwarp();
  }

  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<kTmemCols>(tmem);
}

}  // namespace tcr

bool tc_supported(const KernelParams& p) {
  const BtsRenderCfg& c = p.cfg;
  const int64_t pixels = (int64_t)p.scene.n * p.scene.H * p.scene.W;
  const bool shape = (p.mlp.n_blocks == 0 && (p.mlp.Hd == 32 || p.mlp.Hd == 64 || p.mlp.Hd == 128)) ||
                     (p.mlp.n_blocks == 1 && (p.mlp.Hd == 32 || p.mlp.Hd == 64));
  return shape && p.mlp.P == 39 && c.num_freqs == 6 && c.include_input && c.n_coarse >= 1 && p.scene.Vc <= kMaxColorViews &&
         pixels * (p.mlp.Hd / 4) < (int64_t)0x7fffffff;   // tap offsets are 32-bit (float4 units)
}

template <int HD, bool kExact, bool kPoints, int NB>
static cudaError_t launch_tc_t(const KernelParams& p, cudaStream_t stream) {
  using Lt = tcr::L<HD, NB>;
  auto kern = tcr::render_tc_kernel<HD, kExact, kPoints, NB>;
  static bool configured[64] = {};
  static int sm_count[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !configured[dev]) {   // once per device: function attributes and the SM count
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Lt::kSmemBytes);
    if (e != cudaSuccess) return e;
    // kCtasPerSm CTAs/SM by design (registers, shared memory, TMEM columns): ask for the full shared-memory carve-out,
    // otherwise the default split leaves room for a single CTA.
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
    configured[dev] = true;
  }
  const int sms = dev < 64 && sm_count[dev] > 0 ? sm_count[dev] : 148;
  const int64_t n_pass = kPoints ? (p.n_items + 127) / 128 : (p.n_items + 3) / 4;   // work-list mode: upper bound
  const int64_t cap = (int64_t)sms * Lt::kCtasPerSm;      // persistent: every CTA resident, strided over passes
  const int grid = (int)(n_pass < cap ? n_pass : cap);
  kern<<<grid, tcr::kThreads, Lt::kSmemBytes, stream>>>(p);
  return cudaGetLastError();
}

template <int HD, int NB>
static cudaError_t launch_tc_hd(const KernelParams& p, bool points, cudaStream_t stream) {
  const bool exact = p.cfg.precision == 0;
  if (points) return launch_tc_t<HD, true, true, NB>(p, stream);   // query: exact tier only (occupancy thresholds / BEV profiles)
  return exact ? launch_tc_t<HD, true, false, NB>(p, stream) : launch_tc_t<HD, false, false, NB>(p, stream);
}

cudaError_t launch_tc(const KernelParams& p, bool points, cudaStream_t stream) {
  if (p.mlp.n_blocks == 1) return p.mlp.Hd == 32 ? launch_tc_hd<32, 1>(p, points, stream) : launch_tc_hd<64, 1>(p, points, stream);
  if (p.mlp.Hd == 32) return launch_tc_hd<32, 0>(p, points, stream);
  if (p.mlp.Hd == 64) return launch_tc_hd<64, 0>(p, points, stream);
  return launch_tc_hd<128, 0>(p, points, stream);
}

int tc_barrier_timeouts() {
  int v = 0;
  cudaMemcpyFromSymbol(&v, g_tc_barrier_timeouts, sizeof(int));
  return v;
}

}  // namespace bts
