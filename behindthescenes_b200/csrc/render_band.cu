// Band render kernel: the whole field evaluation of a sample tile on the tensor cores, the feature fetch staged by TMA.
//
// The gather kernel (render_tc.cu) interpolates the folded map G with CUDA cores: 4 taps x Hd channels per sample from
// L1/L2 (1 KiB per sample), ~20 % of its instructions and a third of its stalls.  But the samples of neighbouring rays
// do not scatter over the map: they walk along epipolar lines, and with the reference's disparity-linear sampling
// (models/common/render/nerf.py:117-120) consecutive samples land about a pixel apart.  A tile of 8 consecutive rays x
// 16 consecutive samples therefore touches a thin band of G -- 2 rows x <= 32 pixels for the stereo / same-view
// geometry of the KITTI configs -- and the bilinear fetch of models/bts/model/models_bts.py:179 is a product of a
// sparse weight matrix with that band:
//
//      h[128 samples x Hd] = PE[128 x 48]·W_pe  +  A_w[128 x 16·steps] · Band[16·steps x Hd]
//
//   * Band: G is kept a second time as fp16 hi + lo operand blocks (bts_pack_g16: one 2·Hd·16-byte block per 8 pixels of
//     a row, K-major core matrices; three copies with the pixels shifted by 0, 3 and 6 so that any 30-pixel span fits 4
//     blocks).  One cp.async.bulk per band row (TMA engine, mbarrier complete_tx) brings the blocks the tile needs into
//     shared memory, ready to be the B operand: no per-lane loads, no address arithmetic.
//   * A_w: every sample scatters its 4 bilinear weights (fp16 hi + lo) into its row of a zeroed K-major operand tile
//     and clears them again after the MMA; A_hi·B_hi + A_lo·B_hi + A_hi·B_lo keeps the product fp32-grade (2^-22).
//   * Positional code: operands straight into TMEM (tcgen05.st): 36 sin/cos values as fp16 hi + lo, the 3 raw
//     coordinates as a three-term fp16 split in the spare columns of the same K = 48 operand; 9 MMAs.
// One accumulator, one tcgen05.commit per tile; the epilogue (ReLU, lin_out, softplus), compositing and colour sampling
// stay on the CUDA cores, lane = sample, 16 lanes per ray.
// CTA = 8 warps = two independent tile groups of 4 warps (8 rays each; own accumulator, band, weight tile, named barrier)
// sharing the weight operands; persistent grid, 2 CTAs per SM = 16 warps per SM (Hd 64).
//
// Tiles whose footprint needs more than kTmax k-steps (steep epipolar lines) are not rendered here: the 8-ray group
// is appended to a work list and the gather kernel renders it afterwards (abi.cu) -- same outputs either way.
//
// Reference arithmetic: models/bts/model/models_bts.py:138-216, :218-264, :266-338; models/common/model/code.py:30-42;
// models/common/model/resnetfc.py:132-184 (n_blocks = 0); models/common/render/nerf.py:103-123, :210-313.
#include <climits>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace bts {

__device__ int g_band_barrier_timeouts = 0;

namespace band {

constexpr int kGroups = 2;             // tile groups per CTA
constexpr int kGroupThreads = 128;
constexpr int kThreads = kGroups * kGroupThreads;
constexpr int kTmax = 4;               // band k-steps per tile; one k-step = 2 rows x 8 pixels (K = 16)
constexpr int kMaxK = 256;             // samples per ray (z_steps table in shared memory)
constexpr int kViews = 2;              // colour views (register accumulators)
constexpr int kChunkA = 128 * 16;      // one K-chunk (8 pixels) of the weight operand: 128 rows x 16 B
constexpr uint32_t kFull = 0xffffffffu;

template <int HD>
struct L {
  static_assert(HD == 64, "band kernel: d_hidden 64");
  static constexpr uint32_t kColD = 0, kColAhi = HD, kColAlo = HD + 24, kGroupCols = 128;   // per group
  static constexpr uint32_t kTmemCols = kGroups * kGroupCols;
  static constexpr int kCtasPerSm = 2;
  static constexpr int kBlk = 32 * HD;                          // G16 block of 8 pixels: [hi|lo][HD/8][8 ch][8 px] fp16
  static constexpr int kPlaneB = 16 * HD;                       // lo plane offset inside a block
  static constexpr int kOffW16hi = 0;                           // [6 chunks][HD rows][16 B] fp16, K-major no-swizzle
  static constexpr int kOffW16lo = kOffW16hi + 6 * HD * 16;
  static constexpr int kOffWraw = kOffW16lo + 6 * HD * 16;      // [HD] float4: raw-coordinate columns of lin_in in fp32
  static constexpr int kOffWout = kOffWraw + HD * 16;           // HD floats
  static constexpr int kOffGempty = kOffWout + HD * 4;          // HD floats
  static constexpr int kOffZsteps = kOffGempty + HD * 4;        // kMaxK floats
  static constexpr int kOffViews = kOffZsteps + kMaxK * 4;      // (1 + kViews) x 24 floats: K (9) + w2c rows (12); [0] = encoder view
  static constexpr int kOffBar = kOffViews + (1 + kViews) * 24 * 4;   // per group 128 B: mma, tma, footprint bars, 4 x bbox; then tmem slot
  static constexpr int kOffRays = ((kOffBar + kGroups * 128 + 16 + 15) / 16) * 16;   // per group 8 rays x 32 B (prefetched rows)
  static constexpr int kOffGroup = ((kOffRays + kGroups * 256 + 127) / 128) * 128;
  static constexpr int kBandBytes = 2 * kTmax * kBlk;           // [<= 2 kTmax blocks]
  static constexpr int kPlaneA = 2 * kTmax * kChunkA;           // weight operand: [hi|lo][2 kTmax chunks][128 rows][16 B]
  static constexpr int kGroupBytes = kBandBytes + 2 * kPlaneA;
  static constexpr int kSmemBytes = kOffGroup + kGroups * kGroupBytes;
  static_assert(kCtasPerSm * (kSmemBytes + 1024) <= 233472, "resident CTAs must fit the 228 KB shared memory of an SM");
};

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
__device__ __forceinline__ float seg16_sum(float v) {   // sum over the 16 lanes of a ray (xor offsets < 16 stay inside)
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
// packed fp32 pairs (Blackwell FFMA2: one issue slot for two fused multiply-adds)
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void ffma2(uint64_t& acc, uint64_t a, uint64_t b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("sub.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float2 unpack2(uint64_t v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ void group_sync(int group) {   // barrier over the 4 warps of a tile group
  asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(kGroupThreads) : "memory");
}
// F.softplus(beta=1, threshold=20) (models_bts.py:313) = max(x, 0) + log1p(exp(-|x|)): the exponential is <= 1, so
// ex2 / lg2 on the SFU keep the absolute error at the 1e-7 level (the argument 1 + e is rounded once)
__device__ __forceinline__ float ex2_sfu(float x) {   // 2^x, x <= 0 here: no overflow; results below 2^-126 flush to 0
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float lg2_sfu(float x) {   // log2(x), x in [1, 2] here
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float softplus_sfu(float x) {
  const float e = ex2_sfu(-fabsf(x) * 1.44269504088896340736f);
  return fmaxf(x, 0.0f) + 0.69314718055994530942f * lg2_sfu(1.0f + e);
}

// Software pipeline of a tile group (tile i = chunk i of the group's 8 rays), one group barrier per tile:
//     group_sync                                  operands of tile i complete (TMEM A, weight tile), accumulator free
//     thread 0: 9 code MMAs, wait band(i), band MMAs, commit
//     P0(i+1): depths, projection, taps of the next chunk; footprint published (mbarrier)   } while the MMAs run
//     wait MMA(i); clear own weights(i)
//     footprint(i+1) -> band geometry; thread 0: bulk copies of band(i+1)                     } band(i+1) lands during
//     P1(i+1): positional code -> TMEM A, bilinear weights -> weight tile                     } the epilogue of tile i
//     P3(i), C(i), D(i): epilogue, compositing, colour
// sin and cos of a moderate argument (|x| <= 12 here: |coordinate| <= 8, base frequency 1.5): three-constant
// Cody-Waite reduction by pi/2 and the usual degree-7 / degree-8 minimax polynomials on [-pi/4, pi/4]; ~1 ulp, no
// large-argument path (that one lives in code_far_to_tmem).
__device__ __forceinline__ void sincos_moderate(float x, float& s, float& c) {
  const float q = rintf(x * 0.636619772f);
  const int iq = (int)q;
  float r = fmaf(q, -1.57079601e+00f, x);
  r = fmaf(q, -3.13916473e-07f, r);
  r = fmaf(q, -5.39030253e-15f, r);
  const float r2 = r * r;
  float sp = fmaf(r2, -1.95152959e-04f, 8.33216087e-03f);
  sp = fmaf(sp, r2, -1.66666546e-01f);
  const float sr = fmaf(sp * r2, r, r);
  float cp = fmaf(r2, 2.44331571e-05f, -1.38873163e-03f);
  cp = fmaf(cp, r2, 4.16666418e-02f);
  cp = fmaf(cp, r2, -0.5f);
  const float cr = fmaf(cp, r2, 1.0f);
  const float a = (iq & 1) ? cr : sr, b = (iq & 1) ? sr : cr;     // quadrant: (s,c), (c,-s), (-s,-c), (-c,s)
  s = (iq & 2) ? -a : a;
  c = ((iq + 1) & 2) ? -b : b;
}

// Positional code of the samples of a warp that holds a sample far outside the frustum (|coordinate| > 8: clamped z,
// points behind the camera; called by the whole warp -- the TMEM store is warp-collective): every
// term is evaluated on the reference's own rounded argument sin(fma(x, f_i, phase)) (code.py:38) with the full-range
// sinf, and stored to the TMEM operand directly.  Rare, kept out of line so that its large-argument reduction does not
// cost the common path registers or instruction cache.
template <bool kExact>
__device__ __noinline__ void code_far_to_tmem(float p0, float p1, float p2, float f0, uint32_t t_hi, uint32_t t_lo) {
  const float kHalfPi = 1.57079637050628662109375f;
  const float pv[3] = {p0, p1, p2};
  uint32_t ahi[24], alo[24];
#pragma unroll
  for (int i = 18; i < 24; ++i) { ahi[i] = 0u; alo[i] = 0u; }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    float f = f0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const float s = sinf(__fmul_rn(pv[d], f)), c = sinf(fmaf(pv[d], f, kHalfPi));
      alo[d * 6 + i] = 0u;
      if (kExact) split_h2(s, c, ahi[d * 6 + i], alo[d * 6 + i]);
      else ahi[d * 6 + i] = pack_h2(s, c);
      f *= 2.0f;
    }
  }
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    uint32_t v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = ahi[g * 8 + e];
    tc::tmem_st_x8(t_hi + g * 8, v);
    if (kExact) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = alo[g * 8 + e];
      tc::tmem_st_x8(t_lo + g * 8, v);
    }
  }
}

template <int HD, bool kExact, int VC>
__global__ void __launch_bounds__(kThreads, L<HD>::kCtasPerSm) render_band_kernel(const KernelParams p) {
  using Lt = L<HD>;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, hl = lane & 15;
  const int group = tid >> 7, gt = tid & 127, warp = gt >> 5;   // tile group, thread / warp inside the group
  unsigned char* gbar = smem + Lt::kOffBar + group * 128;
  uint64_t* bar_mma = reinterpret_cast<uint64_t*>(gbar);
  uint64_t* bar_tma = bar_mma + 1;
  uint64_t* bar_box = bar_mma + 2;
  int4* s_bbox = reinterpret_cast<int4*>(gbar + 32);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Lt::kOffBar + kGroups * 128);
  float4* s_ray = reinterpret_cast<float4*>(smem + Lt::kOffRays + group * 256);   // [8 rays][2] prefetched ray rows
  unsigned char* s_band = smem + Lt::kOffGroup + group * Lt::kGroupBytes;
  unsigned char* s_aw = s_band + Lt::kBandBytes;
  const float* s_wout = reinterpret_cast<const float*>(smem + Lt::kOffWout);
  const float* s_gempty = reinterpret_cast<const float*>(smem + Lt::kOffGempty);
  const float* s_zsteps = reinterpret_cast<const float*>(smem + Lt::kOffZsteps);
  const float* s_views = reinterpret_cast<const float*>(smem + Lt::kOffViews);

  const BtsRenderCfg& cfg = p.cfg;
  const int C = p.mlp.C, ld = C + p.mlp.P;
  const int K = cfg.n_coarse, Vc = p.scene.Vc, Cc = p.scene.Cc;
  const int H = p.scene.H, W = p.scene.W, Wo = (W + 6 + 7) >> 3;   // blocks per row in every copy of G16

  // ---- one-time setup: weights -> B operands, tables, zeroed operand tiles, TMEM, mbarriers ----
  for (int i = tid; i < HD * 48; i += kThreads) {
    const int n = i / 48, c = i % 48;
    float w = 0.0f;
    if (c < 36) {   // A column c = (d*6 + i)*2 + sc  <->  reference code index 3 + i*6 + sc*3 + d (code.py:39-41)
      const int sc = c & 1, di = c >> 1, d = di / 6, oct = di % 6;
      w = __ldg(p.mlp.w_in + (size_t)n * ld + C + 3 + oct * 6 + sc * 3 + d);
    } else if (c < 45) {   // three columns per raw coordinate (its fp16 hi, lo, lo2 terms): reference code index d
      w = __ldg(p.mlp.w_in + (size_t)n * ld + C + (c - 36) / 3);
    }
    const __half hi = __float2half_rn(w);
    const __half lo = __float2half_rn(w - __half2float(hi));
    const int off = ((c >> 3) * HD + n) * 16 + (c & 7) * 2;
    *reinterpret_cast<__half*>(smem + Lt::kOffW16hi + off) = hi;
    *reinterpret_cast<__half*>(smem + Lt::kOffW16lo + off) = lo;
  }
  for (int n = tid; n < HD; n += kThreads) {
    const float* wr = p.mlp.w_in + (size_t)n * ld + C;
    reinterpret_cast<float4*>(smem + Lt::kOffWraw)[n] = make_float4(__ldg(wr), __ldg(wr + 1), __ldg(wr + 2), 0.0f);
    reinterpret_cast<float*>(smem + Lt::kOffWout)[n] = __ldg(p.mlp.w_out + n);
    reinterpret_cast<float*>(smem + Lt::kOffGempty)[n] = p.scene.g_empty ? __ldg(p.scene.g_empty + n) : 0.0f;
  }
  for (int i = tid; i < K; i += kThreads) reinterpret_cast<float*>(smem + Lt::kOffZsteps)[i] = __ldg(cfg.z_steps + i);
  for (int i = tid; i < (1 + Vc) * 21; i += kThreads) {
    const int v = i / 21, e = i % 21;
    const float* Kv = v == 0 ? p.scene.K_f : p.scene.K_c + (v - 1) * 9;
    const float* Mv = v == 0 ? p.scene.w2c_f : p.scene.w2c_c + (v - 1) * 16;
    reinterpret_cast<float*>(smem + Lt::kOffViews)[v * 24 + e] = e < 9 ? __ldg(Kv + e) : __ldg(Mv + (e - 9));
  }
  // bands (finite data under zero weights) and the weight operands start as zeros
  for (int i = tid; i < kGroups * Lt::kGroupBytes / 16; i += kThreads)
    reinterpret_cast<int4*>(smem + Lt::kOffGroup)[i] = make_int4(0, 0, 0, 0);
  if (tid < 32) tc::tmem_alloc<Lt::kTmemCols>(tmem_slot);
  if (gt == 0) {
    tc::mbar_init(bar_mma, 1);
    tc::mbar_init(bar_tma, 1);
    tc::mbar_init(bar_box, 4);             // one arrival per warp of the group
    tc::fence_mbar_init();
  }
  tc::fence_proxy_async_smem();
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot + group * Lt::kGroupCols;
  const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);
  const uint32_t tD = tmem + Lt::kColD, tAhi = tmem + Lt::kColAhi, tAlo = tmem + Lt::kColAlo;
  uint32_t mma_phase = 0, tma_phase = 0, box_phase = 0;

  const float b_out = __ldg(p.mlp.b_out);
  const bool use_empty = cfg.learn_empty && p.scene.g_empty != nullptr;
  const uint32_t idesc16 = tc::instr_desc(tc::kFmtF16, 128, HD);
  // Side jobs of a tile group (MMA issue, band copies) go to different warps in different groups: warp w of every CTA runs on
  // scheduler w % 4, so fixed roles would pile all four resident groups' issue work onto scheduler 0.  CTAs of the second half
  // of the grid are the ones that share an SM with the first half (2 CTAs per SM, grid = 2 x SMs).
  const int role = group + (blockIdx.x >= (gridDim.x >> 1) ? 2 : 0);
  const int mma_warp = role & 3, tma_warp = (role + 2) & 3;
  const bool g16_bad = p.scene.g16_flags != nullptr && __ldg(p.scene.g16_flags) != 0;
  const uint32_t sW = tc::smem_u32(smem), sBand = tc::smem_u32(s_band), sAw = tc::smem_u32(s_aw);
  // operand descriptor of the first positional-code weight chunk pair; the others differ only in the start address
  // (descriptor bits 0-13, units of 16 bytes)
  const uint64_t d_w = tc::smem_desc_kmajor_noswizzle(sW + Lt::kOffW16hi, HD * 16, 128);
  constexpr uint64_t kStepW = 2 * HD, kLoW = (Lt::kOffW16lo - Lt::kOffW16hi) / 16;
  const size_t copy_bytes = (size_t)H * Wo * Lt::kBlk;     // one copy of G16 (scene.n == 1)

  const uint64_t seed_eff = jitter_seed(p);
  const uint32_t seed_lo = (uint32_t)seed_eff, seed_mix = (uint32_t)(seed_eff >> 32) * 0x27D4EB2Fu;   // hash_uniform32
  const int n_rays = (int)p.n_items;                        // < 2^31 (launcher)
  const int n_pass = (n_rays + 7) / 8;
  const int pass0 = (int)blockIdx.x * kGroups + group, pass_stride = (int)gridDim.x * kGroups;
  auto view_of = [&](int slot) {                       // camera of a view from shared memory (short live range)
    View vw;
    const float* sv = s_views + slot * 24;
#pragma unroll
    for (int i = 0; i < 9; ++i) vw.K[i] = sv[i];
#pragma unroll
    for (int i = 0; i < 12; ++i) vw.M[i] = sv[9 + i];
    return vw;
  };
  const float4* rays4 = reinterpret_cast<const float4*>(p.rays);
  // ray rows of a pass -> shared memory, asynchronously (lanes 0,1 of each ray's 16 lanes copy its two 16-byte halves)
  auto prefetch_rays = [&](int pass) {
    if (rays4 && pass < n_pass && hl < 2) {
      int r = pass * 8 + warp * 2 + (lane >> 4);
      r = r < n_rays ? r : n_rays - 1;
      tc::cp_async16(s_ray + (warp * 2 + (lane >> 4)) * 2 + hl, rays4 + (size_t)r * 2 + hl);
    }
  };
  prefetch_rays(pass0);

  // tap state of a sample, packed: x0 | y0 << 12 | dx << 24 | dy << 25 | invalid << 26 | empty << 27 | far (warp) << 28
  struct Tap { int bits; float wx1, wy1; };
  struct Geom { int xs, ob, no, yb, nsteps; };   // band: copy shift (0 | 4 px), first block, blocks per row, first row, k-steps
  struct Slots { uint32_t a00, a01, a10, a11; };

  for (int pass = pass0; pass < n_pass; pass += pass_stride) {
    // ---- per-pass setup: this lane's ray (16 lanes share a ray) ----
    const int ray_raw = pass * 8 + warp * 2 + (lane >> 4);
    const bool active = ray_raw < n_rays;
    const int ray = active ? ray_raw : n_rays - 1;
    Ray ry;
    if (rays4) {
      tc::cp_async_wait_all();
      __syncwarp();
      const float4 ra = s_ray[(warp * 2 + (lane >> 4)) * 2], rb = s_ray[(warp * 2 + (lane >> 4)) * 2 + 1];
      ry.o[0] = ra.x; ry.o[1] = ra.y; ry.o[2] = ra.z;
      ry.d[0] = ra.w; ry.d[1] = rb.x; ry.d[2] = rb.y;
      ry.near_ = rb.z; ry.far_ = rb.w; ry.zfac = 1.0f;
      __syncwarp();
      prefetch_rays(pass + pass_stride);  // next pass's rays: the copy lands while this pass renders
    } else {
      ry = fetch_ray(p, ray, lane);       // generated from the pixel's camera (no shuffles on this path)
    }
    if (__builtin_expect(g16_bad, 0)) {   // G does not fit fp16: every group goes to the gather kernel
      if (gt == 0) p.work[4 + atomicAdd(p.work, 1)] = pass;
      continue;
    }
    float near_, far_;
    depth_bounds(p, ry, near_, far_);

    // P0 of a chunk: depth (nerf.py:103-123, rounded like the reference's separate elementwise ops), world point,
    // projection into the encoder view, depth code, bilinear taps; the warp's tap footprint is published
    auto stage_p0 = [&](int k, float& z, float& pu, float& pv_, float& pcode) -> Tap {
      const float uu = p.u ? __ldg(p.u + (int64_t)ray * K + k) : hash_uniform32(seed_lo, seed_mix, (uint32_t)ray, (uint32_t)k);
      const float s = __fadd_rn(s_zsteps[k], __fmul_rn(uu, cfg.step));
      const float one_m = __fsub_rn(1.0f, s);
      const float zz = __fadd_rn(__fmul_rn(near_, one_m), __fmul_rn(far_, s));
      z = cfg.lindisp ? __frcp_rn(zz) : zz;
      const float X = __fadd_rn(ry.o[0], __fmul_rn(z, ry.d[0]));   // nerf.py:231
      const float Y = __fadd_rn(ry.o[1], __fmul_rn(z, ry.d[1]));
      const float Z = __fadd_rn(ry.o[2], __fmul_rn(z, ry.d[2]));
      const Proj pf = project(view_of(0), X, Y, Z);      // encoder view (scene.n == 1)
      pu = pf.u; pv_ = pf.v; pcode = depth_code(cfg, pf);
      const Taps t = bilinear_taps(pf.u, pf.v, H, W);
      const bool empty_me = use_empty && pf.invalid;      // takes the learned empty row instead of a fetch
      // warp-uniform (the TMEM stores / loads of the two code paths are warp-collective): one far sample sends its warp
      const bool big = __any_sync(kFull, fmaxf(fabsf(pu), fmaxf(fabsf(pv_), fabsf(pcode))) > kPeFastLimit);
      int xl = empty_me ? INT_MAX : t.x0, xh = empty_me ? -1 : t.x0 + t.dx;
      int yl = empty_me ? INT_MAX : t.y0, yh = empty_me ? -1 : t.y0 + t.dy;
      xl = __reduce_min_sync(kFull, xl); xh = __reduce_max_sync(kFull, xh);
      yl = __reduce_min_sync(kFull, yl); yh = __reduce_max_sync(kFull, yh);
      if (lane == 0) {
        s_bbox[warp] = make_int4(xl, xh, yl, yh);
        tc::mbar_arrive(bar_box);
      }
      Tap tp;
      tp.bits = t.x0 | (t.y0 << 12) | (t.dx << 24) | (t.dy << 25) | ((int)pf.invalid << 26) | ((int)empty_me << 27) | ((int)big << 28);
      tp.wx1 = t.wx1; tp.wy1 = t.wy1;
      return tp;
    };
    // footprints of the 4 warps -> band geometry (uniform over the group)
    auto stage_geom = [&]() -> Geom {
      tc::mbar_wait_or_trap(bar_box, box_phase, &g_band_barrier_timeouts);
      box_phase ^= 1u;
      const int4 b0 = s_bbox[0], b1 = s_bbox[1], b2 = s_bbox[2], b3 = s_bbox[3];
      const int xl = min(min(b0.x, b1.x), min(b2.x, b3.x)), xh = max(max(b0.y, b1.y), max(b2.y, b3.y));
      const int yl = min(min(b0.z, b1.z), min(b2.z, b3.z)), yh = max(max(b0.w, b1.w), max(b2.w, b3.w));
      Geom g;
      // three copies of G16 whose blocks start at 0, 5 and 2 mod 8 (pixels shifted by 0, 3, 6): the one that leaves the span
      // at most 2 pixels into its first block, so any span <= 30 pixels fits kTmax = 4 blocks
      const int r = xl & 7;
      g.xs = r < 2 ? 0 : (r < 5 ? 6 : 3);
      g.ob = (xl + g.xs) >> 3; g.no = ((xh + g.xs) >> 3) - g.ob + 1; g.yb = yl;
      g.nsteps = xh >= xl ? ((yh - yl + 2) >> 1) * g.no : 0;       // row pairs x blocks
      return g;
    };
    auto issue_band = [&](const Geom& g) {   // one bulk copy per band row: `no` consecutive blocks of G16
      if (__shfl_sync(kFull, warp, 0) == tma_warp) {       // whole warp (uniform operands), one elected lane issues
        const int nsteps_u = __shfl_sync(kFull, g.nsteps, 0), no_u = __shfl_sync(kFull, g.no, 0);
        const int yb_u = __shfl_sync(kFull, g.yb, 0), ob_u = __shfl_sync(kFull, g.ob, 0), xs_u = __shfl_sync(kFull, g.xs, 0);
        if (nsteps_u > 0 && tc::elect_one()) {
          const uint32_t row_bytes = (uint32_t)no_u * Lt::kBlk;
          const int rows_all = 2 * (nsteps_u / no_u);
          const int nrows = rows_all < H - yb_u ? rows_all : H - yb_u;
          tc::mbar_expect_tx(bar_tma, row_bytes * (uint32_t)nrows);
          const unsigned char* src = reinterpret_cast<const unsigned char*>(p.scene.G16) + (size_t)(xs_u / 3) * copy_bytes +
                                     ((size_t)yb_u * Wo + ob_u) * Lt::kBlk;
#pragma unroll 1
          for (int r = 0; r < nrows; ++r)
            tc::bulk_load(s_band + r * row_bytes, src + (size_t)r * Wo * Lt::kBlk, row_bytes, bar_tma);
        }
      }
    };
    // P1a: positional code of (u, v, code) -> A operand in TMEM
    auto stage_code = [&](float pu, float pv_, float pcode, bool big) {
      if (__builtin_expect(big, 0)) {   // warp-uniform flag (stage_p0): the whole warp evaluates term by term
        code_far_to_tmem<kExact>(pu, pv_, pcode, cfg.freq_factor, t_lane + Lt::kColAhi, t_lane + Lt::kColAlo);
        return;
      }
      const float pv[3] = {pu, pv_, pcode};
      uint32_t ahi[24], alo[24];
#pragma unroll
      for (int i = 18; i < 24; ++i) { ahi[i] = 0u; alo[i] = 0u; }
      // one accurate sincos per coordinate, then exact double-angle steps per octave; the recurrences of u and v run as
      // packed fp32 pairs (FADD2 / FMUL2: same IEEE operations, half the issue slots), the depth code's as scalars
      float s0, c0, s1, c1, s2, c2;
      sincos_moderate(__fmul_rn(pv[0], cfg.freq_factor), s0, c0);
      sincos_moderate(__fmul_rn(pv[1], cfg.freq_factor), s1, c1);
      sincos_moderate(__fmul_rn(pv[2], cfg.freq_factor), s2, c2);
      uint64_t S = pack2(s0, s1), C = pack2(c0, c1);
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const float2 sv = unpack2(S), cv = unpack2(C);
        if (kExact) {
          split_h2(sv.x, cv.x, ahi[0 * 6 + i], alo[0 * 6 + i]);
          split_h2(sv.y, cv.y, ahi[1 * 6 + i], alo[1 * 6 + i]);
          split_h2(s2, c2, ahi[2 * 6 + i], alo[2 * 6 + i]);
        } else {
          ahi[0 * 6 + i] = pack_h2(sv.x, cv.x); ahi[1 * 6 + i] = pack_h2(sv.y, cv.y); ahi[2 * 6 + i] = pack_h2(s2, c2);
        }
        const uint64_t S2 = mul2(add2(S, S), C);
        C = mul2(sub2(C, S), add2(C, S));
        S = S2;
        const float t2 = (s2 + s2) * c2;
        c2 = (c2 - s2) * (c2 + s2);
        s2 = t2;
      }
      // raw coordinates (|p| <= 8): exact three-term fp16 split h + l + l2 in columns 36..44 of the hi operand
      __half r[9];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const __half h0 = __float2half_rn(pv[d]);
        const float e1 = pv[d] - __half2float(h0);
        const __half h1 = __float2half_rn(e1);
        r[3 * d + 0] = h0; r[3 * d + 1] = h1; r[3 * d + 2] = __float2half_rn(e1 - __half2float(h1));
      }
      auto pk = [](__half a, __half b) { return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16); };
      ahi[18] = pk(r[0], r[1]); ahi[19] = pk(r[2], r[3]); ahi[20] = pk(r[4], r[5]); ahi[21] = pk(r[6], r[7]);
      ahi[22] = pk(r[8], __float2half_rn(0.0f));
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        uint32_t v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = ahi[g * 8 + e];
        tc::tmem_st_x8(t_lane + Lt::kColAhi + g * 8, v);
        if (kExact) {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = alo[g * 8 + e];
          tc::tmem_st_x8(t_lane + Lt::kColAlo + g * 8, v);
        }
      }
    };
    // P1b: the 4 bilinear weights of a sample -> its row of the weight operand.  Slot of a tap: k-step = (row pair,
    // block), chunk = row parity, element = pixel in block
    unsigned char* aw_row = s_aw + gt * 16;
    auto stage_scatter = [&](const Tap& tp, const Geom& g) -> Slots {
      Slots sl = {0u, 0u, 0u, 0u};
      if (!((tp.bits >> 27) & 1)) {
        const int x0 = tp.bits & 0xfff, y0 = (tp.bits >> 12) & 0xfff, dx = (tp.bits >> 24) & 1, dy = (tp.bits >> 25) & 1;
        const int oy0 = y0 - g.yb, oy1 = oy0 + dy, x0s = x0 + g.xs - 8 * g.ob, x1s = x0s + dx;
        const uint32_t r0 = (uint32_t)((2 * (oy0 >> 1) * g.no + (oy0 & 1)) * kChunkA), r1 = (uint32_t)((2 * (oy1 >> 1) * g.no + (oy1 & 1)) * kChunkA);
        const uint32_t q0 = (uint32_t)((x0s >> 3) * 2 * kChunkA + ((x0s & 7) << 1)), q1 = (uint32_t)((x1s >> 3) * 2 * kChunkA + ((x1s & 7) << 1));
        sl.a00 = r0 + q0; sl.a01 = r0 + q1; sl.a10 = r1 + q0; sl.a11 = r1 + q1;
        const float wx0 = 1.0f - tp.wx1, wy0 = 1.0f - tp.wy1;   // == (floor + 1) - ix of the reference (both are exact)
        auto put = [&](uint32_t a, float w) {
          const __half hi = __float2half_rn(w);
          *reinterpret_cast<__half*>(aw_row + a) = hi;
          if (kExact) *reinterpret_cast<__half*>(aw_row + Lt::kPlaneA + a) = __float2half_rn(w - __half2float(hi));
        };
        // clamped neighbours coincide with their base tap and carry weight exactly 0: base taps are written last
        put(sl.a11, tp.wx1 * tp.wy1); put(sl.a10, wx0 * tp.wy1); put(sl.a01, tp.wx1 * wy0); put(sl.a00, wx0 * wy0);
      }
      return sl;
    };
    auto fail_group = [&]() {              // a tile does not fit the band: the gather kernel renders this 8-ray group
      if (gt == 0) p.work[4 + atomicAdd(p.work, 1)] = pass;
      group_sync(group);                   // everyone has read the footprint before the next pass overwrites it
    };

    float T_carry = 1.0f, depth_part = 0.0f, wsum_part = 0.0f;
    float acc[VC][4], mass[VC];
#pragma unroll
    for (int v = 0; v < VC; ++v) {
      mass[v] = 0.0f;
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[v][c] = 0.0f;
    }
    // One loop body serves the whole pass: iteration -1 only prepares the operands of the first tile (the "next tile" half of
    // the body), iterations 0 .. K/16 - 1 consume a tile and prepare the following one -- a single copy of the sampling / code /
    // scatter code in the instruction stream.
    float z = 0.0f, pu = 0.0f, pv_ = 0.0f, pcode = 0.0f;
    Tap tap = {0, 0.0f, 0.0f};
    Geom geom = {0, 0, 1, 0, 0};
    Slots slots = {0u, 0u, 0u, 0u};
    bool failed = false;

    for (int c0 = -16; c0 < K; c0 += 16) {
      const bool cur = c0 >= 0;            // a tile to consume (warp-uniform)
      const int k = c0 + hl;
      const uint32_t sample_id = (uint32_t)ray * (uint32_t)K + (uint32_t)k;   // < 2^31 elements per output (launcher)
      if (cur) group_sync(group);          // operands of this tile complete; the previous tile is fully consumed
      // ---------------- MMA issue (one thread of the group) ----------------
      if (cur && __shfl_sync(kFull, warp, 0) == mma_warp) {   // whole warp (uniform operands), one elected lane issues
        const int nsteps_u = __shfl_sync(kFull, geom.nsteps, 0), no_u = __shfl_sync(kFull, geom.no, 0);
        const bool leader = tc::elect_one();
        tc::tc_fence_after_sync();
        if (leader) {
#pragma unroll
          for (int ks = 0; ks < 3; ++ks) {
            tc::mma_f16_ts(tD, tAhi + ks * 8, d_w + ks * kStepW, idesc16, ks > 0);
            if (kExact) {
              tc::mma_f16_ts(tD, tAlo + ks * 8, d_w + ks * kStepW, idesc16, 1);
              tc::mma_f16_ts(tD, tAhi + ks * 8, d_w + ks * kStepW + kLoW, idesc16, 1);
            }
          }
        }
        if (nsteps_u > 0) {
          tc::mbar_wait_or_trap(bar_tma, tma_phase, &g_band_barrier_timeouts);   // the band has landed
          if (leader) {
            const uint32_t row_bytes = (uint32_t)no_u * Lt::kBlk;
            uint64_t a_hi = tc::smem_desc_kmajor_noswizzle(sAw, kChunkA, 128);
            uint64_t b_hi = tc::smem_desc_kmajor_noswizzle(sBand, row_bytes, 128);
            int oct = 0;
#pragma unroll 1
            for (int ts = 0; ts < nsteps_u; ++ts) {
              tc::mma_f16_ss(tD, a_hi, b_hi, idesc16, 1);
              if (kExact) {
                tc::mma_f16_ss(tD, a_hi + Lt::kPlaneA / 16, b_hi, idesc16, 1);
                tc::mma_f16_ss(tD, a_hi, b_hi + Lt::kPlaneB / 16, idesc16, 1);
              }
              a_hi += 2 * kChunkA / 16;
              b_hi += Lt::kBlk / 16;
              if (++oct == no_u) { oct = 0; b_hi += row_bytes / 16; }   // next row pair
            }
          }
        }
        if (leader) tc::mma_commit(bar_mma);
      }
      if (cur && geom.nsteps > 0) tma_phase ^= 1u;
      __syncwarp();

      // ---------------- P0 of the next chunk, while the MMAs run ----------------
      const bool has_next = c0 + 16 < K;
      float z_n = 0.0f;
      Tap tap_n = {0, 0.0f, 0.0f};
      if (has_next) tap_n = stage_p0(k + 16, z_n, pu, pv_, pcode);
      float z_next = __shfl_down_sync(kFull, z, 1, 16);
      const float z_first_of_next = __shfl_sync(kFull, z_n, 0, 16);
      if (hl == 15) z_next = z_first_of_next;

      if (cur) {
        tc::mbar_wait_or_trap(bar_mma, mma_phase, &g_band_barrier_timeouts);
        mma_phase ^= 1u;
        tc::tc_fence_after_sync();
        if (!((tap.bits >> 27) & 1)) {       // clear my weights again: the operand tile is all zeros between tiles
          const __half zero = __float2half_rn(0.0f);
          *reinterpret_cast<__half*>(aw_row + slots.a11) = zero; *reinterpret_cast<__half*>(aw_row + slots.a10) = zero;
          *reinterpret_cast<__half*>(aw_row + slots.a01) = zero; *reinterpret_cast<__half*>(aw_row + slots.a00) = zero;
          if (kExact) {
            *reinterpret_cast<__half*>(aw_row + Lt::kPlaneA + slots.a11) = zero; *reinterpret_cast<__half*>(aw_row + Lt::kPlaneA + slots.a10) = zero;
            *reinterpret_cast<__half*>(aw_row + Lt::kPlaneA + slots.a01) = zero; *reinterpret_cast<__half*>(aw_row + Lt::kPlaneA + slots.a00) = zero;
          }
        }
      }
      // ---------------- operands of the next tile (its band lands during the epilogue below) ----------------
      Geom geom_n = {0, 0, 1, 0, 0};
      Slots slots_n = {0u, 0u, 0u, 0u};
      if (has_next) {
        geom_n = stage_geom();
        if (__builtin_expect(geom_n.nsteps > kTmax, 0)) { failed = true; break; }     // uniform over the group
        issue_band(geom_n);
        stage_code(pu, pv_, pcode, (tap_n.bits >> 28) & 1);
        slots_n = stage_scatter(tap_n, geom_n);
      }

      if (cur) {
        // ---------------- P3: accumulator -> sigma ----------------
        const bool inv_f = (tap.bits >> 26) & 1, empty_me = (tap.bits >> 27) & 1, big = (tap.bits >> 28) & 1;
        float logit = b_out;
        const bool any_big = __any_sync(kFull, big), any_empty = __any_sync(kFull, empty_me);   // warp-uniform: tcgen05.ld is warp-collective
        if (__builtin_expect(!any_big, 1)) {
          // samples outside the encoder frustum take the learned empty row instead of the band (their weights are zero):
          // the warp that holds one adds m * g_empty to every accumulator column, m = 1 on those lanes
          auto fast_logit = [&](auto with_empty) {
            constexpr bool kE = decltype(with_empty)::value;
            const uint64_t m2 = pack2(empty_me ? 1.0f : 0.0f, empty_me ? 1.0f : 0.0f);
            uint64_t l01 = pack2(0.0f, 0.0f), l23 = l01, l45 = l01, l67 = l01;   // 8 chains in 4 packed accumulators
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {                    // 32 accumulator columns at a time (register budget)
              uint32_t dv[2][16];
              tc::tmem_ld_x16(t_lane + Lt::kColD + hh * 32, dv[0]);
              tc::tmem_ld_x16(t_lane + Lt::kColD + hh * 32 + 16, dv[1]);
              tc::tmem_ld_wait();
              const float4* wo4 = reinterpret_cast<const float4*>(s_wout) + hh * 8;
              const float4* ge4 = reinterpret_cast<const float4*>(s_gempty) + hh * 8;
#pragma unroll
              for (int c = 0; c < 2; ++c) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float4 wa = wo4[c * 4 + e];
                  uint64_t h01 = pack2(__uint_as_float(dv[c][e * 4 + 0]), __uint_as_float(dv[c][e * 4 + 1]));
                  uint64_t h23 = pack2(__uint_as_float(dv[c][e * 4 + 2]), __uint_as_float(dv[c][e * 4 + 3]));
                  if (kE) {
                    const float4 ga = ge4[c * 4 + e];
                    ffma2(h01, m2, pack2(ga.x, ga.y));
                    ffma2(h23, m2, pack2(ga.z, ga.w));
                  }
                  const float2 a = unpack2(h01), bq = unpack2(h23);
                  ffma2((e & 1) ? l45 : l01, pack2(fmaxf(a.x, 0.0f), fmaxf(a.y, 0.0f)), pack2(wa.x, wa.y));
                  ffma2((e & 1) ? l67 : l23, pack2(fmaxf(bq.x, 0.0f), fmaxf(bq.y, 0.0f)), pack2(wa.z, wa.w));
                }
              }
            }
            const float2 s01 = unpack2(l01), s23 = unpack2(l23), s45 = unpack2(l45), s67 = unpack2(l67);
            const float l0 = s01.x + s45.x, l1 = s01.y + s45.y, l2 = s23.x + s67.x, l3 = s23.y + s67.y;
            logit += (l0 + l1) + (l2 + l3);
          };
          if (any_empty) fast_logit(std::true_type{});
          else fast_logit(std::false_type{});
        } else {   // rare (a sample far outside the frustum): the raw-coordinate part of lin_in added in fp32 (code recomputed)
          float r0 = 0.0f, r1 = 0.0f, r2 = 0.0f;
          if (big) {
            const Proj pq = project(view_of(0), __fadd_rn(ry.o[0], __fmul_rn(z, ry.d[0])), __fadd_rn(ry.o[1], __fmul_rn(z, ry.d[1])),
                                    __fadd_rn(ry.o[2], __fmul_rn(z, ry.d[2])));
            r0 = pq.u; r1 = pq.v; r2 = depth_code(cfg, pq);
          }
          const float4* wr4 = reinterpret_cast<const float4*>(smem + Lt::kOffWraw);
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t dv[16];
            tc::tmem_ld_x16(t_lane + Lt::kColD + c * 16, dv);
            tc::tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int j = c * 16 + e;
              const float4 wr = wr4[j];
              float h = __uint_as_float(dv[e]) + (empty_me ? s_gempty[j] : 0.0f);
              h = fmaf(wr.z, r2, fmaf(wr.y, r1, fmaf(wr.x, r0, h)));
              logit = fmaf(fmaxf(h, 0.0f), s_wout[j], logit);
            }
          }
        }
        const float sigma = softplus_sfu(logit);

        // ---------------- C: alpha compositing over the 16 samples of the chunk (nerf.py:283-299) ----------------
        const float delta = (k + 1 < K) ? __fsub_rn(z_next, z) : 1e10f;
        // exp(-x), x >= 0, as 2^(-x log2 e) on the SFU: absolute error ~1e-7
        float alpha = __fsub_rn(1.0f, ex2_sfu(-(fabsf(delta) * fmaxf(sigma, 0.0f)) * 1.44269504088896340736f));
        if (cfg.hard_alpha_cap && k == K - 1) alpha = 1.0f;
        const float tr = active ? __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f) : 1.0f;
        float incl = tr;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
          const float up = __shfl_up_sync(kFull, incl, o, 16);
          if (hl >= o) incl *= up;
        }
        float excl = __shfl_up_sync(kFull, incl, 1, 16);
        if (hl == 0) excl = 1.0f;
        const float Tk = T_carry * excl;
        T_carry *= __shfl_sync(kFull, incl, 15, 16);
        const float wgt = active ? alpha * Tk : 0.0f;
        depth_part = fmaf(wgt, z, depth_part);
        wsum_part += wgt;
        if (active) {
          if (p.out.weights) p.out.weights[sample_id] = wgt;
          if (p.out.alphas) p.out.alphas[sample_id] = alpha;
          if (p.out.z_samps) p.out.z_samps[sample_id] = z;
          if (p.out.sigma) p.out.sigma[sample_id] = sigma;
        }

        // ---------------- D: colour views (models_bts.py:218-264, :330-334) ----------------
        if (p.out.rgb || p.out.invalid || p.out.rgb_samps || p.out.invalid_mass) {
          const int Hc = p.scene.Hc, Wc = p.scene.Wc;
#pragma unroll
          for (int v = 0; v < VC; ++v) {
            if (v < Vc) {
              int x0, y0, dx, dy;
              float wx1, wy1;
              bool inv_c;
              if ((p.scene.same_view_mask >> v) & 1) {   // this colour view is the encoder view: same projection, taps, flags
                x0 = tap.bits & 0xfff; y0 = (tap.bits >> 12) & 0xfff; dx = (tap.bits >> 24) & 1; dy = (tap.bits >> 25) & 1;
                wx1 = tap.wx1; wy1 = tap.wy1; inv_c = inv_f;
              } else {
                const Proj pc = project(view_of(1 + v), __fadd_rn(ry.o[0], __fmul_rn(z, ry.d[0])), __fadd_rn(ry.o[1], __fmul_rn(z, ry.d[1])),
                                        __fadd_rn(ry.o[2], __fmul_rn(z, ry.d[2])));
                const Taps tv = bilinear_taps(pc.u, pc.v, Hc, Wc);
                x0 = tv.x0; y0 = tv.y0; dx = tv.dx; dy = tv.dy; wx1 = tv.wx1; wy1 = tv.wy1; inv_c = pc.invalid;
              }
              float col[4] = {0, 0, 0, 0};
              if (active) {
                const float4* base = reinterpret_cast<const float4*>(p.scene.img4) + (uint32_t)((v * Hc + y0) * Wc + x0);
                const float4 c00 = __ldg(base), c01 = __ldg(base + dx);
                const float4 c10 = __ldg(base + dy * Wc), c11 = __ldg(base + dy * Wc + dx);
                const float wx0 = 1.0f - wx1, wy0 = 1.0f - wy1;
                const float w00 = wx0 * wy0, w01 = wx1 * wy0, w10 = wx0 * wy1, w11 = wx1 * wy1;
                col[0] = fmaf(c11.x, w11, fmaf(c10.x, w10, fmaf(c01.x, w01, c00.x * w00)));
                col[1] = fmaf(c11.y, w11, fmaf(c10.y, w10, fmaf(c01.y, w01, c00.y * w00)));
                col[2] = fmaf(c11.z, w11, fmaf(c10.z, w10, fmaf(c01.z, w01, c00.z * w00)));
                if (Cc > 3) col[3] = fmaf(c11.w, w11, fmaf(c10.w, w10, fmaf(c01.w, w01, c00.w * w00)));
              }
              const float inv = (inv_c | inv_f) ? 1.0f : 0.0f;   // :333-334
              if (active) {
                if (p.out.invalid) p.out.invalid[sample_id * (uint32_t)Vc + v] = inv;
                if (p.out.rgb_samps) {
#pragma unroll
                  for (int c = 0; c < 4; ++c)
                    if (c < Cc) p.out.rgb_samps[(sample_id * (uint32_t)Vc + v) * (uint32_t)Cc + c] = col[c];
                }
              }
#pragma unroll
              for (int c = 0; c < 4; ++c) acc[v][c] = fmaf(wgt, col[c], acc[v][c]);   // nerf.py:298, reduced once per ray
              mass[v] = fmaf(wgt, inv, mass[v]);                                      // loss.py:107
            }
          }
        }
      }
      // next tile becomes current
      z = z_n; tap = tap_n; geom = geom_n; slots = slots_n;
      tc::tmem_st_wait();
      tc::fence_proxy_async_smem();        // weight-operand stores -> async proxy (tensor core)
      tc::tc_fence_before_sync();          // TMEM reads / writes of this iteration are ordered before the group barrier
    }
    if (failed) { fail_group(); continue; }
    // ---- per-ray outputs (nerf.py:299-304) ----
    {
      const float depth = seg16_sum(depth_part), wsum = seg16_sum(wsum_part);
      if (hl == 0 && active) {
        if (p.out.depth) p.out.depth[ray] = depth;
        if (p.out.depth_z) p.out.depth_z[ray] = depth * ry.zfac;
      }
#pragma unroll
      for (int v = 0; v < VC; ++v) {
        if (v < Vc) {
          const float m = seg16_sum(mass[v]);
          float r[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) r[c] = seg16_sum(acc[v][c]);
          if (hl == 0 && active) {
            if (p.out.invalid_mass) p.out.invalid_mass[(int64_t)ray * Vc + v] = m;
            if (p.out.rgb) {
#pragma unroll
              for (int c = 0; c < 4; ++c)
                if (c < Cc) p.out.rgb[((int64_t)ray * Vc + v) * Cc + c] = cfg.white_bkgd ? r[c] + 1.0f - wsum : r[c];
            }
          }
        }
      }
    }
  }

  tc::cp_async_wait_all();
  tc::tc_fence_before_sync();
  __syncthreads();
  if (tid < 32) tc::tmem_dealloc<Lt::kTmemCols>(*tmem_slot);
}

// G (n, H, W, HD) fp32 -> operand blocks, three copies (pixels shifted right by 0, 3 and 6: block o of copy s covers pixels
// 8 o - 3 s .. 8 o - 3 s + 7): thread = (copy, row, block, channel): 8 strided loads, two 16-byte stores.
template <int HD>
__global__ void pack_g16_kernel(const float* __restrict__ G, int64_t rows, int W, int Wo, unsigned char* __restrict__ out,
                                int32_t* __restrict__ flags) {
  const int64_t per_copy = rows * Wo * HD, total = 3 * per_copy;
  bool bad = false;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int copy = (int)(i / per_copy), shift = 3 * copy;
    const int64_t ii = i - copy * per_copy;
    const int c = (int)(ii % HD);
    const int64_t oi = ii / HD;                // row * Wo + block
    const int o = (int)(oi % Wo);
    const int64_t row = oi / Wo;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      float v[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int x = o * 8 - shift + j + e;
        v[e] = (x >= 0 && x < W) ? __ldg(G + ((row * W + x) * HD + c)) : 0.0f;
        bad |= !(fabsf(v[e]) <= 60000.0f);      // also catches NaN / inf
      }
      const __half2 h = __floats2half2_rn(v[0], v[1]);
      const float2 hf = __half22float2(h);
      const __half2 l = __floats2half2_rn(v[0] - hf.x, v[1] - hf.y);
      hi[j >> 1] = *reinterpret_cast<const uint32_t*>(&h);
      lo[j >> 1] = *reinterpret_cast<const uint32_t*>(&l);
    }
    unsigned char* blk = out + (copy * per_copy + oi * HD) * 32 + (c >> 3) * 128 + (c & 7) * 16;
    *reinterpret_cast<uint4*>(blk) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(blk + 16 * HD) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(flags, 1);
}

}  // namespace band

bool band_supported(const KernelParams& p) {
  const BtsRenderCfg& c = p.cfg;
  return p.scene.G16 != nullptr && p.work != nullptr && p.scene.n == 1 && p.mlp.Hd == 64 && p.mlp.n_blocks == 0 &&
         p.mlp.P == 39 && c.num_freqs == 6 && c.include_input && c.n_coarse % 16 == 0 && c.n_coarse <= band::kMaxK &&
         p.scene.Vc <= band::kViews && p.scene.H <= 4096 && p.scene.W <= 4096 &&
         p.n_items * c.n_coarse * (p.scene.Vc > 0 ? p.scene.Vc * p.scene.Cc : 1) < (int64_t)0x7fffffff &&   // 32-bit output indices
         (int64_t)p.scene.Vc * p.scene.Hc * p.scene.Wc < (int64_t)0x7fffffff &&
         c.band_mode == 0 && c.precision == 0 && p.xyz == nullptr &&
         (reinterpret_cast<uintptr_t>(p.rays) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.scene.G16) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(p.scene.img4) & 15) == 0;
}

cudaError_t launch_band(const KernelParams& p, cudaStream_t stream) {
  using Lt = band::L<64>;
  const bool two_views = p.scene.Vc > 1;
  auto kern = two_views ? band::render_band_kernel<64, true, 2> : band::render_band_kernel<64, true, 1>;
  static bool configured[64][2] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  static int sms[64] = {};
  if (dev < 64 && !configured[dev][two_views]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Lt::kSmemBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
    configured[dev][two_views] = true;
  }
  const int n_sm = dev < 64 && sms[dev] > 0 ? sms[dev] : 148;
  cudaError_t e = cudaMemsetAsync(p.work, 0, 16, stream);     // work[0] = counter of handed-over groups
  if (e != cudaSuccess) return e;
  const int64_t n_pass = (p.n_items + 7) / 8;
  const int64_t n_cta = (n_pass + band::kGroups - 1) / band::kGroups;
  const int64_t cap = (int64_t)n_sm * Lt::kCtasPerSm;
  const int grid = (int)(n_cta < cap ? n_cta : cap);
  kern<<<grid, band::kThreads, Lt::kSmemBytes, stream>>>(p);
  return cudaGetLastError();
}

// three copies (pixels shifted by 0, 3, 6), each (n*H rows) x ceil((W+6)/8) blocks of 32*Hd bytes
int64_t g16_bytes(int n, int H, int W, int Hd) { return 3 * (int64_t)n * H * ((W + 6 + 7) / 8) * 32 * Hd; }

cudaError_t launch_pack_g16(const float* G, int n, int H, int W, int Hd, void* out, int32_t* flags, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(flags, 0, sizeof(int32_t), stream);
  if (e != cudaSuccess) return e;
  const int Wo = (W + 6 + 7) / 8;
  const int64_t rows = (int64_t)n * H, total = 3 * rows * Wo * Hd;
  const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  if (Hd == 64) band::pack_g16_kernel<64><<<grid, 256, 0, stream>>>(G, rows, W, Wo, static_cast<unsigned char*>(out), flags);
  else band::pack_g16_kernel<128><<<grid, 256, 0, stream>>>(G, rows, W, Wo, static_cast<unsigned char*>(out), flags);
  return cudaGetLastError();
}

int band_barrier_timeouts() {
  int v = 0;
  cudaMemcpyFromSymbol(&v, g_band_barrier_timeouts, sizeof(int));
  return v;
}

}  // namespace bts
