// Shared device helpers for the BTS render kernels (sm_100a).
// Every helper cites the reference lines whose arithmetic it reproduces (paths relative to the reference root).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bts_b200.h"

namespace bts {

constexpr float kEps = 1e-3f;            // models/bts/model/models_bts.py:14
constexpr int kMaxColorViews = 8;
constexpr float kPeFastLimit = 8.0f;     // |coordinate| above this -> per-term sinf (reference-exact argument)

struct KernelParams {
  BtsScene scene;
  BtsMlp mlp;
  BtsRenderCfg cfg;
  BtsRenderOut out;
  const float* rays;     // (n*R, 8)            [render]; null -> generated from `gen`
  BtsRayGen gen;         // in-kernel ray generation (rays == null)
  const float* u;        // (n*R, K) or null    [render]
  const float* xyz;      // (n*P, 3)            [query]
  float* q_sigma;        // (n*P)               [query]
  float* q_invalid_f;    // (n*P)               [query]
  float* q_rgb;          // (n*P, Vc*Cc)|null   [query]
  float* q_invalid;      // (n*P, Vc)|null      [query]
  int64_t n_items;       // n*R rays or n*P points
  int64_t per_batch;     // R or P
  uint64_t seed;
  int p_pad;             // P rounded up to a multiple of 4
  int32_t* work;         // band kernel: work[0] = number of 8-ray groups handed over, work[4 + i] = their indices
  int work_mode;         // gather kernel: 1 = render only the 8-ray groups listed in `work` (after the band kernel)
};

// ------------------------------------------------------------------------------------------------------------------
// Counter-based generator for the stratified jitter when the caller does not inject `u`
// (the reference draws torch.rand_like, models/common/render/nerf.py:116).  Stateless: the murmur3 32-bit finaliser
// over a word that mixes (seed, ray, sample) with odd multipliers -- every (ray, sample) gets an independent-looking
// uniform, reproducible per seed, ~10 instructions.  uniform = top 24 bits * 2^-24 in [0,1) like torch's float uniform.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}
// the same generator for ray indices below 2^32, with the seed words pre-mixed once per kernel:
// seed_lo = (uint32_t)seed, seed_mix = (uint32_t)(seed >> 32) * 0x27D4EB2F
__device__ __forceinline__ float hash_uniform32(uint32_t seed_lo, uint32_t seed_mix, uint32_t ray, uint32_t k) {
  uint32_t h = ray * 0x9E3779B1u + k * 0x85EBCA77u + seed_lo;
  h = fmix32(h ^ seed_mix ^ 0x165667B1u);
  return (float)(h >> 8) * (1.0f / 16777216.0f);
}
__device__ __forceinline__ float hash_uniform(uint64_t seed, uint64_t ray, uint32_t k) {
  uint32_t h = (uint32_t)ray * 0x9E3779B1u + k * 0x85EBCA77u + (uint32_t)seed;
  h ^= (uint32_t)(ray >> 32) * 0xC2B2AE3Du + (uint32_t)(seed >> 32) * 0x27D4EB2Fu;
  h = fmix32(h ^ 0x165667B1u);
  return (float)(h >> 8) * (1.0f / 16777216.0f);
}

// seed of a launch: the host word, mixed with a device word when the caller keeps one (cfg.seed_dev: the jitter then
// changes from replay to replay of a captured CUDA graph -- the caller bumps the device word inside the graph)
__device__ __forceinline__ uint64_t jitter_seed(const KernelParams& p) {
  return p.cfg.seed_dev ? p.seed ^ (__ldg(p.cfg.seed_dev) * 0x9E3779B97F4A7C15ull) : p.seed;
}

// ------------------------------------------------------------------------------------------------------------------
// One ray = [origin, direction, near, far] (models/common/util/util.py:244-273), either read from the caller's tensor
// (8 lanes load, shuffle-broadcast) or generated from the camera of its pixel.  All lanes return the same values.
// zfac = z-component of the unit ray in its camera = the distance->z-depth factor of projection_operations.py:4-16.
// ------------------------------------------------------------------------------------------------------------------
struct Ray {
  float o[3], d[3], near_, far_, zfac;
};

__device__ __forceinline__ float linspace_m1_1(int i, int steps) {   // torch.linspace(-1, 1, steps)[i], float32
  if (steps == 1) return -1.0f;
  const float step = 2.0f / (float)(steps - 1);
  return i < steps / 2 ? -1.0f + step * (float)i : 1.0f - step * (float)(steps - 1 - i);
}

// lanes 0..7 hold the 8 floats of a ray tensor row (prefetched by the caller) -> broadcast
__device__ __forceinline__ Ray ray_from_lanes(float rv) {
  Ray r;
  r.o[0] = __shfl_sync(0xffffffffu, rv, 0); r.o[1] = __shfl_sync(0xffffffffu, rv, 1); r.o[2] = __shfl_sync(0xffffffffu, rv, 2);
  r.d[0] = __shfl_sync(0xffffffffu, rv, 3); r.d[1] = __shfl_sync(0xffffffffu, rv, 4); r.d[2] = __shfl_sync(0xffffffffu, rv, 5);
  r.near_ = __shfl_sync(0xffffffffu, rv, 6); r.far_ = __shfl_sync(0xffffffffu, rv, 7);
  r.zfac = 1.0f;
  return r;
}

__device__ __forceinline__ Ray fetch_ray(const KernelParams& p, int64_t ray, int lane) {
  Ray r;
  if (p.rays) {
    const float rv = lane < 8 ? __ldg(p.rays + ray * 8 + lane) : 0.0f;
    r.o[0] = __shfl_sync(0xffffffffu, rv, 0); r.o[1] = __shfl_sync(0xffffffffu, rv, 1); r.o[2] = __shfl_sync(0xffffffffu, rv, 2);
    r.d[0] = __shfl_sync(0xffffffffu, rv, 3); r.d[1] = __shfl_sync(0xffffffffu, rv, 4); r.d[2] = __shfl_sync(0xffffffffu, rv, 5);
    r.near_ = __shfl_sync(0xffffffffu, rv, 6); r.far_ = __shfl_sync(0xffffffffu, rv, 7);
    r.zfac = 1.0f;
    return r;
  }
  const BtsRayGen& g = p.gen;
  const int64_t hw = (int64_t)g.H * g.W;
  const int64_t cam = ray / hw;                 // b * v + view
  const int pix = (int)(ray - cam * hw);
  const int y = pix / g.W, x = pix - y * g.W;
  const float* Kc = g.K + cam * 9;
  const float* M = g.poses_c2w + cam * 16;
  const float cx = __fdiv_rn(__fsub_rn(linspace_m1_1(x, g.W), __ldg(Kc + 2)), __ldg(Kc + 0));   // util.py:140-144
  const float cy = __fdiv_rn(__fsub_rn(linspace_m1_1(y, g.H), __ldg(Kc + 5)), __ldg(Kc + 4));
  const float nrm = sqrtf(fmaf(cy, cy, cx * cx) + 1.0f);
  float dx = cx, dy = cy, dz = 1.0f;
  if (g.norm_dir) { dx = __fdiv_rn(cx, nrm); dy = __fdiv_rn(cy, nrm); dz = __fdiv_rn(1.0f, nrm); }   // util.py:147-148
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    r.d[i] = fmaf(__ldg(M + i * 4 + 2), dz, fmaf(__ldg(M + i * 4 + 1), dy, __ldg(M + i * 4 + 0) * dx));   // util.py:256-258
    r.o[i] = __ldg(M + i * 4 + 3);
  }
  r.near_ = g.z_near; r.far_ = g.z_far;
  r.zfac = g.norm_dir ? __fdiv_rn(1.0f, nrm) : 1.0f;
  return r;
}

// ------------------------------------------------------------------------------------------------------------------
// Stratified depth of sample k on a ray (nerf.py:103-123), rounding exactly like the reference's separate
// elementwise ops (no FMA contraction): steps = linspace[k] + u*step;  lindisp: 1/(1/near*(1-s) + 1/far*s).
// ------------------------------------------------------------------------------------------------------------------
// `near_` / `far_` are 1/near, 1/far (rounded once per ray, like the reference's `1 / near` tensors) when lindisp.
__device__ __forceinline__ float sample_depth(const KernelParams& p, int64_t ray, int k, float near_, float far_) {
  const float uu = p.u ? __ldg(p.u + ray * p.cfg.n_coarse + k) : hash_uniform(jitter_seed(p), (uint64_t)ray, (uint32_t)k);
  const float s = __fadd_rn(__ldg(p.cfg.z_steps + k), __fmul_rn(uu, p.cfg.step));
  const float one_m = __fsub_rn(1.0f, s);
  if (p.cfg.lindisp) return __fdiv_rn(1.0f, __fadd_rn(__fmul_rn(near_, one_m), __fmul_rn(far_, s)));
  return __fadd_rn(__fmul_rn(near_, one_m), __fmul_rn(far_, s));
}
// per-ray bounds in the form sample_depth consumes
__device__ __forceinline__ void depth_bounds(const KernelParams& p, const Ray& r, float& a, float& b) {
  a = p.cfg.lindisp ? __fdiv_rn(1.0f, r.near_) : r.near_;
  b = p.cfg.lindisp ? __fdiv_rn(1.0f, r.far_) : r.far_;
}

// ------------------------------------------------------------------------------------------------------------------
// Pinhole projection of a world point into one view (models_bts.py:145-155 / :221-232):
//   cam = M[:3,:] @ [x;1];  dist = |cam|;  (a,b,c) = K @ cam;  uv = (a,b)/max(c,EPS);  invalid = c<=EPS | |u|>1 | |v|>1
// ------------------------------------------------------------------------------------------------------------------
struct View {
  float K[9];
  float M[12];
};

__device__ __forceinline__ void load_view(View& v, const float* __restrict__ K33, const float* __restrict__ M44) {
#pragma unroll
  for (int i = 0; i < 9; ++i) v.K[i] = __ldg(K33 + i);
#pragma unroll
  for (int i = 0; i < 12; ++i) v.M[i] = __ldg(M44 + i);
}

struct Proj {
  float u, v, c, dist;
  bool invalid;
};

__device__ __forceinline__ float dot4(const float* m, float x, float y, float z) {
  float a = m[0] * x;
  a = fmaf(m[1], y, a);
  a = fmaf(m[2], z, a);
  return a + m[3];
}
__device__ __forceinline__ float dot3(const float* m, float x, float y, float z) {
  float a = m[0] * x;
  a = fmaf(m[1], y, a);
  return fmaf(m[2], z, a);
}

__device__ __forceinline__ Proj project(const View& vw, float x, float y, float z) {
  Proj r;
  const float cx = dot4(vw.M + 0, x, y, z), cy = dot4(vw.M + 4, x, y, z), cz = dot4(vw.M + 8, x, y, z);
  r.dist = sqrtf(fmaf(cz, cz, fmaf(cy, cy, cx * cx)));
  const float a = dot3(vw.K + 0, cx, cy, cz), b = dot3(vw.K + 3, cx, cy, cz);
  r.c = dot3(vw.K + 6, cx, cy, cz);
  const float den = fmaxf(r.c, kEps);
  r.u = __fdiv_rn(a, den);
  r.v = __fdiv_rn(b, den);
  r.invalid = (r.c <= kEps) | (r.u < -1.0f) | (r.u > 1.0f) | (r.v < -1.0f) | (r.v > 1.0f);
  return r;
}

// Depth code in [-1,1] (models_bts.py:157-171); python-scalar constants pre-rounded on the host.
__device__ __forceinline__ float depth_code(const BtsRenderCfg& cfg, const Proj& pr) {
  const float zeta = cfg.code_mode == 0 ? pr.c : pr.dist;
  float t;
  if (cfg.inv_z) {
    t = __fdiv_rn(__fsub_rn(__fdiv_rn(1.0f, fmaxf(zeta, kEps)), cfg.inv_z_far), cfg.inv_z_range);
  } else {
    t = __fdiv_rn(__fsub_rn(zeta, cfg.z_near), cfg.z_range);
  }
  return __fsub_rn(__fmul_rn(2.0f, t), 1.0f);
}

// ------------------------------------------------------------------------------------------------------------------
// Bilinear tap set of F.grid_sample(mode="bilinear", padding_mode="border", align_corners=False)
// (models_bts.py:179, :234):  ix = ((u+1)*W-1)/2 clipped to [0,W-1]; taps floor(ix), +1 (clamped; its weight is 0 there).
// ------------------------------------------------------------------------------------------------------------------
struct Taps {
  int x0, y0, dx, dy;          // dx in {0,1}, dy in {0,1}: clamped neighbour offsets
  float wx0, wx1, wy0, wy1;
};

__device__ __forceinline__ void axis_taps(float coord, int size, int& i0, int& d, float& w0, float& w1) {
  float ix = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(coord, 1.0f), (float)size), 1.0f), 0.5f);
  ix = fminf(fmaxf(ix, 0.0f), (float)(size - 1));
  const float f = floorf(ix);
  w1 = ix - f;
  w0 = (f + 1.0f) - ix;
  i0 = (int)f;
  d = (i0 + 1 < size) ? 1 : 0;
}

__device__ __forceinline__ Taps bilinear_taps(float u, float v, int H, int W) {
  Taps t;
  axis_taps(u, W, t.x0, t.dx, t.wx0, t.wx1);
  axis_taps(v, H, t.y0, t.dy, t.wy0, t.wy1);
  return t;
}

// ------------------------------------------------------------------------------------------------------------------
// Positional code of (u, v, zhat) (models/common/model/code.py:30-42): out = [x, sin(f0 x), sin(f0 x + pi/2), ...],
// octave-major, then (sin,cos), then dim; the reference evaluates sin(fma(x, f_i, phase)) (torch.addcmul is a fused
// multiply-add on CPU and GPU).  Fast path: one accurate sincosf at the base octave, then exact double-angle steps
// (error grows ~2x per octave, like the reference's own argument rounding).  Coordinates beyond kPeFastLimit
// (samples far outside the frustum / behind the camera, clamped z) take the per-term path so the huge arguments are
// rounded exactly like the reference's.  `dst[j * stride]` receives code entry j.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void positional_code(const BtsRenderCfg& cfg, float p0, float p1, float p2, float* dst,
                                                int stride) {
  const float pv[3] = {p0, p1, p2};
  int off = 0;
  if (cfg.include_input) {
#pragma unroll
    for (int d = 0; d < 3; ++d) dst[d * stride] = pv[d];
    off = 3;
  }
  const float kHalfPi = 1.57079637050628662109375f;  // float32(np.pi * 0.5), code.py:27
  const bool big = fmaxf(fabsf(p0), fmaxf(fabsf(p1), fabsf(p2))) > kPeFastLimit;
  if (!big) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      float s, c;
      sincosf(__fmul_rn(pv[d], cfg.freq_factor), &s, &c);
      for (int i = 0; i < cfg.num_freqs; ++i) {
        dst[(off + i * 6 + d) * stride] = s;
        dst[(off + i * 6 + 3 + d) * stride] = c;
        const float s2 = 2.0f * s * c;
        c = (c - s) * (c + s);
        s = s2;
      }
    }
  } else {
    float f = cfg.freq_factor;
    for (int i = 0; i < cfg.num_freqs; ++i) {
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        dst[(off + i * 6 + d) * stride] = sinf(__fmul_rn(pv[d], f));
        dst[(off + i * 6 + 3 + d) * stride] = sinf(fmaf(pv[d], f, kHalfPi));
      }
      f *= 2.0f;
    }
  }
}

__device__ __forceinline__ float softplus(float x) {  // F.softplus beta=1 threshold=20 (models_bts.py:313)
  return x > 20.0f ? x : log1pf(expf(x));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace bts
