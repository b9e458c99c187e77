// Backward of the layer-1 fold G[b, p, :] = W_F * F[b, :, p] + b_in as one persistent TMA-fed tcgen05 kernel:
//   dF[b, :, p]  = W_F^T dG[b, p, :]            (what flows back into the encoder)
//   dW_F[k, c]  += sum_{b,p} dG[b, p, k] F[b, c, p]
//   db_in[k]    += sum_{b,p} dG[b, p, k]
// dG (n, HW, Hd) and F (n, C, HW) are each read from HBM exactly once and dF written once (the torch formulation this
// replaces ran three SIMT sgemms plus two permute copies of the 500 MB tensors).
//
// Per 64-pixel tile, all operands K-major / no-swizzle core matrices in shared memory, 3xTF32 by hi/lo splitting with
// the two halves of the SPLIT operand stacked along M so every product is an M = 128 MMA:
//   GEMM 1  D1[128 x 64 px]   = [W_hi^T ; W_lo^T](128 x Hd) * dG_hi^T(Hd x 64) + (same) * dG_lo^T      rows c and 64 + c add up to dF[c, px]
//   GEMM 2  D2[128 x C + 16] += [dG_hi^T ; dG_lo^T](128 x 64 px) * [F_hi ; 1](64 px x C + 16) + (same) * [F_lo ; 0]
//                                                                      rows k and 64 + k add up to dW_F[k, :], column C to db_in[k]
// (all four hi/lo cross terms are present, so the result is fp32-grade).  D2 stays in TMEM for the whole kernel and is
// reduced into global memory with one atomicAdd per element per CTA at the end.
// Reference: this is the part of loss.backward() (utils/base_trainer.py:297) that differentiates ResnetFC.lin_in
// (models/common/model/resnetfc.py:88,147) with respect to the sampled encoder features and its weights.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tcgen05.cuh"
#include "tensor_map.cuh"

namespace bts {

__device__ int g_fold_bwd_barrier_timeouts = 0;

namespace foldbwd {

constexpr int kThreads = 256;
constexpr int kTile = 64;            // pixels per tile
constexpr int kPad = 4;              // floats of row padding in the row-exchange buffer
// The raw staging tiles are dense (one bulk copy for the contiguous dG tile, one tensor-map copy for the F tile); the
// transposes below walk them diagonally (chunk index offset by the row index) so that every quarter-warp touches eight
// distinct 16-byte bank groups on both the read and the write side.

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc::smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

template <int HD, int C>
struct L {
  static_assert(HD == 32 || HD == 64, "stacked hi/lo rows need Hd <= 64");
  static_assert(C == 32 || C == 64, "stacked hi/lo rows need C <= 64");
  static constexpr int kN2 = C + 16;                                // GEMM 2 columns: C features + the ones column (+ padding to 16)
  static constexpr int kRawGRow = HD * 4;                           // bytes per staged dG row (dense)
  static constexpr int kRawFRow = kTile * 4;                        // bytes per staged F row (dense)
  static constexpr int kRawG = kTile * kRawGRow;
  static constexpr int kRawF = C * kRawFRow;
  static constexpr int kB1 = (HD / 4) * kTile * 16;                 // dG^T, K = hidden unit: [Hd/4 chunks][64 px rows][16 B]
  static constexpr int kA2 = (kTile / 4) * 128 * 16;                // dG,   K = pixel:       [16 chunks][128 rows][16 B]
  static constexpr int kB2 = (kTile / 4) * kN2 * 16;                // F,    K = pixel:       [16 chunks][C + 16 rows][16 B]
  static constexpr int kA1 = (HD / 4) * 128 * 16;                   // W_F^T stacked:         [Hd/4 chunks][128 rows][16 B]
  static constexpr int kExRow = (kTile + kPad) * 4;
  static constexpr int kEx = 64 * kExRow;                           // row exchange (>= C + 1 floats per row)
  static constexpr int kOffRawG = 0;                                // x2
  static constexpr int kOffRawF = 2 * kRawG;                        // x2
  static constexpr int kOffB1hi = kOffRawF + 2 * kRawF;
  static constexpr int kOffB1lo = kOffB1hi + kB1;
  static constexpr int kOffA2 = kOffB1lo + kB1;
  static constexpr int kOffB2hi = kOffA2 + kA2;
  static constexpr int kOffB2lo = kOffB2hi + kB2;
  static constexpr int kOffA1 = kOffB2lo + kB2;
  static constexpr int kOffEx = kOffA1 + kA1;
  static constexpr int kOffBar = kOffEx + kEx;                      // full[2], mma: 3 mbarriers + tmem slot
  static constexpr int kSmemBytes = kOffBar + 64;
  static constexpr uint32_t kTmemCols = 256;                        // D1: 64 columns at 0, D2: C + 16 columns at 64
  static_assert(kTile + kPad >= C + 1, "exchange rows hold a D2 row");
  static_assert(kSmemBytes <= 227 * 1024, "backward fold tile does not fit shared memory");
};

template <int HD, int C>
__global__ void __launch_bounds__(kThreads, 1) fold_bwd_tc_kernel(const float* __restrict__ dG, const float* __restrict__ feat,
                                                                   int64_t HW, int n_batches, const float* __restrict__ w_in,
                                                                   int ld, float* __restrict__ d_feat, float* __restrict__ d_wf,
                                                                   float* __restrict__ d_b,
                                                                   const __grid_constant__ CUtensorMap feat_map, int use_map,
                                                                   const unsigned char* __restrict__ tile_mask) {
  using Lt = L<HD, C>;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Lt::kOffBar);        // [2]
  uint64_t* mma_bar = full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full + 3);
  const bool want_df = d_feat != nullptr;

  // ---- constant operands ----
  for (int i = tid; i < (Lt::kA2 + 2 * Lt::kB2 + Lt::kA1) / 16; i += kThreads)            // rows no transform ever writes stay zero
    reinterpret_cast<float4*>(smem + Lt::kOffA2)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  for (int i = tid; i < HD * C; i += kThreads) {                                             // A1 row c (hi) / 64 + c (lo), K = k
    const int k = i / C, c = i % C;
    const float w = __ldg(w_in + (size_t)k * ld + c);
    const float hi = tf32_hi(w);
    unsigned char* base = smem + Lt::kOffA1 + ((k >> 2) * 128) * 16 + (k & 3) * 4;
    *reinterpret_cast<float*>(base + c * 16) = hi;
    *reinterpret_cast<float*>(base + (64 + c) * 16) = w - hi;
  }
  for (int i = tid; i < kTile / 4; i += kThreads)                                           // the ones row of [F ; 1] (hi half only)
    *reinterpret_cast<float4*>(smem + Lt::kOffB2hi + (i * Lt::kN2 + C) * 16) = make_float4(1.f, 1.f, 1.f, 1.f);
  if (warp == 0) tc::tmem_alloc<Lt::kTmemCols>(tmem_slot);
  if (tid == 0) {
    tc::mbar_init(&full[0], 1);
    tc::mbar_init(&full[1], 1);
    tc::mbar_init(mma_bar, 1);
    tc::fence_mbar_init();
  }
  tc::fence_proxy_async_smem();
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);   // TMEM lanes of this warp
  const int trow = (warp & 3) * 32 + lane, thalf = warp >> 2;              // accumulator row / column half this thread reads
  const uint32_t idesc1 = tc::instr_desc(tc::kFmtTF32, 128, kTile);
  const uint32_t idesc2 = tc::instr_desc(tc::kFmtTF32, 128, Lt::kN2);

  const int64_t tiles_per_b = (HW + kTile - 1) / kTile;
  const int64_t n_tiles = tiles_per_b * n_batches;
  auto issue_load = [&](int64_t tile, int buf) {          // one elected thread: the dG tile is contiguous, the F tile one box
    const int64_t b = tile / tiles_per_b, p0 = (tile % tiles_per_b) * kTile;
    const int valid = (int)((HW - p0) < kTile ? (HW - p0) : kTile);
    unsigned char* gdst = smem + Lt::kOffRawG + buf * Lt::kRawG;
    unsigned char* fdst = smem + Lt::kOffRawF + buf * Lt::kRawF;
    const uint32_t g_bytes = (uint32_t)valid * HD * 4u;
    if (use_map) {
      mbar_expect_tx(&full[buf], g_bytes + (uint32_t)Lt::kRawF);
      bulk_load(gdst, dG + ((size_t)b * HW + p0) * HD, g_bytes, &full[buf]);
      tmap::load_tile_3d(fdst, &feat_map, (int)p0, 0, (int)b, &full[buf]);
    } else {
      mbar_expect_tx(&full[buf], g_bytes + (uint32_t)C * valid * 4u);
      bulk_load(gdst, dG + ((size_t)b * HW + p0) * HD, g_bytes, &full[buf]);
      const float* fsrc = feat + ((size_t)b * C) * HW + p0;
      for (int c = 0; c < C; ++c) bulk_load(fdst + c * Lt::kRawFRow, fsrc + (size_t)c * HW, (uint32_t)valid * 4u, &full[buf]);
    }
  };

  uint32_t full_phase[2] = {0, 0}, mma_phase = 0;
  // Sparse fold (training): tile_mask has one byte per 64-pixel tile, 0 = no ray of the batch sampled it -- its d_G rows were
  // never written (and not even zeroed): the tile contributes nothing to dW_F / db_in and its dF rows are zeros.
  auto zero_tile = [&](int64_t t) {
    if (!want_df) return;
    const int64_t b = t / tiles_per_b, p0 = (t % tiles_per_b) * kTile;
    const int valid = (int)((HW - p0) < kTile ? (HW - p0) : kTile);
    float* dst = d_feat + ((size_t)b * C) * HW + p0;
    for (int i = tid; i < C * (kTile / 4); i += kThreads) {
      const int c = i / (kTile / 4), seg = i % (kTile / 4);
      if (seg * 4 < valid) *reinterpret_cast<float4*>(dst + (size_t)c * HW + seg * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto advance = [&](int64_t t) {                          // next tile of this CTA that carries gradient; the others get zeros
    while (t < n_tiles && tile_mask && tile_mask[t] == 0) { zero_tile(t); t += gridDim.x; }
    return t;
  };
  int64_t tile = advance(blockIdx.x);
  if (tid == 0 && use_map) tmap::prefetch_map(&feat_map);
  if (tid == 0 && tile < n_tiles) issue_load(tile, 0);
  int buf = 0;
  bool first = true;
  float* ex = reinterpret_cast<float*>(smem + Lt::kOffEx);
  constexpr int kExF = kTile + kPad;
  int64_t next;
  for (; tile < n_tiles; tile = next, buf ^= 1) {
    next = advance(tile + gridDim.x);
    if (tid == 0 && next < n_tiles) issue_load(next, buf ^ 1);             // raw[buf^1] was consumed one iteration ago
    const int64_t b = tile / tiles_per_b, p0 = (tile % tiles_per_b) * kTile;
    const int valid = (int)((HW - p0) < kTile ? (HW - p0) : kTile);

    tc::mbar_wait_or_trap(&full[buf], full_phase[buf], &g_fold_bwd_barrier_timeouts);   // bounded: a pipeline bug traps instead of hanging or running on
    full_phase[buf] ^= 1u;
    const unsigned char* rawg = smem + Lt::kOffRawG + buf * Lt::kRawG;
    const unsigned char* rawf = smem + Lt::kOffRawF + buf * Lt::kRawF;
    // ---- (a) dG^T with K = hidden unit (GEMM 1's B): thread = pixel row, 16-byte chunks of its dG row ----
    if (want_df) {
      const int p = tid % kTile, r = tid / kTile;
      constexpr int kChunks = HD / 4, kRows = kThreads / kTile;
#pragma unroll
      for (int it = 0; it < kChunks / kRows; ++it) {
        const int kc = (p + kRows * it + r) % kChunks;                        // diagonal walk
        float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p < valid) f = *reinterpret_cast<const float4*>(rawg + p * Lt::kRawGRow + kc * 16);
        const float4 h = make_float4(tf32_hi(f.x), tf32_hi(f.y), tf32_hi(f.z), tf32_hi(f.w));
        *reinterpret_cast<float4*>(smem + Lt::kOffB1hi + (kc * kTile + p) * 16) = h;
        *reinterpret_cast<float4*>(smem + Lt::kOffB1lo + (kc * kTile + p) * 16) = make_float4(f.x - h.x, f.y - h.y, f.z - h.z, f.w - h.w);
      }
    }
    // ---- (b) dG with K = pixel (GEMM 2's A): thread = hidden unit, rows k (hi) and 64 + k (lo) ----
    {
      const int k = tid % HD;
#pragma unroll 4
      for (int pc = tid / HD; pc < kTile / 4; pc += kThreads / HD) {
        float f[4], h[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int p = pc * 4 + j;
          f[j] = p < valid ? *reinterpret_cast<const float*>(rawg + p * Lt::kRawGRow + k * 4) : 0.f;
          h[j] = tf32_hi(f[j]);
        }
        *reinterpret_cast<float4*>(smem + Lt::kOffA2 + (pc * 128 + k) * 16) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(smem + Lt::kOffA2 + (pc * 128 + 64 + k) * 16) = make_float4(f[0] - h[0], f[1] - h[1], f[2] - h[2], f[3] - h[3]);
      }
    }
    // ---- (c) F with K = pixel (GEMM 2's B): thread = channel row, 16-byte chunks of its F row ----
    {
      const int c = tid % C, r = tid / C;
      constexpr int kChunks = kTile / 4, kRows = kThreads / C;
#pragma unroll
      for (int it = 0; it < kChunks / kRows; ++it) {
        const int pc = (c + kRows * it + r) % kChunks;                        // diagonal walk
        float4 f = *reinterpret_cast<const float4*>(rawf + c * Lt::kRawFRow + pc * 16);
        if (pc * 4 + 0 >= valid) f.x = 0.f;
        if (pc * 4 + 1 >= valid) f.y = 0.f;
        if (pc * 4 + 2 >= valid) f.z = 0.f;
        if (pc * 4 + 3 >= valid) f.w = 0.f;
        const float4 h = make_float4(tf32_hi(f.x), tf32_hi(f.y), tf32_hi(f.z), tf32_hi(f.w));
        *reinterpret_cast<float4*>(smem + Lt::kOffB2hi + (pc * Lt::kN2 + c) * 16) = h;
        *reinterpret_cast<float4*>(smem + Lt::kOffB2lo + (pc * Lt::kN2 + c) * 16) = make_float4(f.x - h.x, f.y - h.y, f.z - h.z, f.w - h.w);
      }
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();                       // operand tiles complete; the previous tile's D1 has been read
    if (tid == 0) {
      tc::tc_fence_after_sync();
      const uint32_t sb = tc::smem_u32(smem);
      if (want_df) {
#pragma unroll
        for (int ks = 0; ks < HD / 8; ++ks) {
          const uint64_t a1 = tc::smem_desc_kmajor_noswizzle(sb + Lt::kOffA1 + ks * 2 * 128 * 16, 128 * 16, 128);
          const uint64_t bhi = tc::smem_desc_kmajor_noswizzle(sb + Lt::kOffB1hi + ks * 2 * kTile * 16, kTile * 16, 128);
          const uint64_t blo = tc::smem_desc_kmajor_noswizzle(sb + Lt::kOffB1lo + ks * 2 * kTile * 16, kTile * 16, 128);
          tc::mma_tf32_ss(tmem, a1, bhi, idesc1, ks > 0);
          tc::mma_tf32_ss(tmem, a1, blo, idesc1, 1);
        }
      }
#pragma unroll
      for (int ks = 0; ks < kTile / 8; ++ks) {
        const uint64_t a2 = tc::smem_desc_kmajor_noswizzle(sb + Lt::kOffA2 + ks * 2 * 128 * 16, 128 * 16, 128);
        const uint64_t bhi = tc::smem_desc_kmajor_noswizzle(sb + Lt::kOffB2hi + ks * 2 * Lt::kN2 * 16, Lt::kN2 * 16, 128);
        const uint64_t blo = tc::smem_desc_kmajor_noswizzle(sb + Lt::kOffB2lo + ks * 2 * Lt::kN2 * 16, Lt::kN2 * 16, 128);
        tc::mma_tf32_ss(tmem + 64, a2, bhi, idesc2, (!first || ks > 0) ? 1u : 0u);
        tc::mma_tf32_ss(tmem + 64, a2, blo, idesc2, 1);
      }
      tc::mma_commit(mma_bar);
    }
    first = false;
    __syncwarp();
    tc::mbar_wait_or_trap(mma_bar, mma_phase, &g_fold_bwd_barrier_timeouts);   // bounded: a pipeline bug traps instead of hanging or running on
    mma_phase ^= 1u;
    tc::tc_fence_after_sync();
    // ---- epilogue: dF[c, px] = D1[c, px] + D1[64 + c, px], stored as rows of the NCHW gradient ----
    if (want_df) {
      constexpr int kHalf = kTile / 2;                       // warps 0-3 read columns [0, 32), warps 4-7 columns [32, 64)
      float row[kHalf];
#pragma unroll
      for (int c0 = 0; c0 < kHalf; c0 += 16) {
        uint32_t v[16];
        tc::tmem_ld_x16(t_lane + thalf * kHalf + c0, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) row[c0 + e] = __uint_as_float(v[e]);
      }
      float* exr = ex + (trow & 63) * kExF + thalf * kHalf;
      if (trow >= 64 && trow - 64 < C) {
#pragma unroll
        for (int e = 0; e < kHalf; e += 4) *reinterpret_cast<float4*>(exr + e) = make_float4(row[e], row[e + 1], row[e + 2], row[e + 3]);
      }
      __syncthreads();
      if (trow < C) {
#pragma unroll
        for (int e = 0; e < kHalf; e += 4) {
          float4 o = *reinterpret_cast<float4*>(exr + e);
          o.x += row[e]; o.y += row[e + 1]; o.z += row[e + 2]; o.w += row[e + 3];
          *reinterpret_cast<float4*>(exr + e) = o;
        }
      }
      __syncthreads();
      float* dst = d_feat + ((size_t)b * C) * HW + p0;
#pragma unroll 4
      for (int i = tid; i < C * (kTile / 4); i += kThreads) {
        const int c = i / (kTile / 4), seg = i % (kTile / 4);
        if (seg * 4 < valid) *reinterpret_cast<float4*>(dst + (size_t)c * HW + seg * 4) = *reinterpret_cast<const float4*>(ex + c * kExF + seg * 4);
      }
    }
    tc::tc_fence_before_sync();
  }
  // ---- final: dW_F[k, :] and db_in[k] from D2 rows k and 64 + k ----
  __syncthreads();
  tc::tc_fence_after_sync();
  if (!first) {                            // a CTA whose tiles were all skipped never wrote its accumulator
    float row[Lt::kN2];
    if (warp < 4) {
#pragma unroll
      for (int c0 = 0; c0 < Lt::kN2; c0 += 16) {
        uint32_t v[16];
        tc::tmem_ld_x16(t_lane + 64 + c0, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) row[c0 + e] = __uint_as_float(v[e]);
      }
      if (tid >= 64 && tid - 64 < HD) {
#pragma unroll
        for (int e = 0; e <= C; ++e) ex[(tid - 64) * kExF + e] = row[e];
      }
    }
    __syncthreads();
    if (tid < HD) {
#pragma unroll
      for (int e = 0; e <= C; ++e) ex[tid * kExF + e] += row[e];
    }
    __syncthreads();
    for (int i = tid; i < HD * C; i += kThreads) atomicAdd(d_wf + i, ex[(i / C) * kExF + (i % C)]);
    for (int k = tid; k < HD; k += kThreads) atomicAdd(d_b + k, ex[k * kExF + C]);
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<Lt::kTmemCols>(tmem);
}

}  // namespace foldbwd

bool fold_bwd_tc_supported(const float* dG, const float* feat, const float* d_feat, int C, int64_t HW, int Hd) {
  const uintptr_t align = reinterpret_cast<uintptr_t>(dG) | reinterpret_cast<uintptr_t>(feat) | reinterpret_cast<uintptr_t>(d_feat);
  return (C == 32 || C == 64) && (Hd == 32 || Hd == 64) && HW % 4 == 0 && (align & 15) == 0;
}

template <int HD, int C>
static cudaError_t launch_fold_bwd_t(const float* dG, const float* feat, int n, int64_t HW, const float* w_in, int ld,
                                     float* d_feat, float* d_wf, float* d_b, const unsigned char* tile_mask, cudaStream_t stream) {
  using Lt = foldbwd::L<HD, C>;
  auto kern = foldbwd::fold_bwd_tc_kernel<HD, C>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Lt::kSmemBytes);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t n_tiles = ((HW + foldbwd::kTile - 1) / foldbwd::kTile) * n;
  const int grid = (int)(n_tiles < sms ? n_tiles : sms);
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  const char* rows = getenv("BTS_FOLD_TMA_ROWS");            // tests: force the row-copy loads
  const int use_map = !(rows && rows[0] == '1') && tmap::make_feature_map(&map, feat, n, C, HW, foldbwd::kTile) ? 1 : 0;
  kern<<<grid, foldbwd::kThreads, Lt::kSmemBytes, stream>>>(dG, feat, HW, n, w_in, ld, d_feat, d_wf, d_b, map, use_map, tile_mask);
  return cudaGetLastError();
}

cudaError_t launch_fold_bwd_tc(const float* dG, const float* feat, int n, int C, int64_t HW, const float* w_in, int ld, int Hd,
                               float* d_feat, float* d_wf, float* d_b, const unsigned char* tile_mask, cudaStream_t stream) {
  if (C == 64 && Hd == 64) return launch_fold_bwd_t<64, 64>(dG, feat, n, HW, w_in, ld, d_feat, d_wf, d_b, tile_mask, stream);
  if (C == 64 && Hd == 32) return launch_fold_bwd_t<32, 64>(dG, feat, n, HW, w_in, ld, d_feat, d_wf, d_b, tile_mask, stream);
  if (C == 32 && Hd == 64) return launch_fold_bwd_t<64, 32>(dG, feat, n, HW, w_in, ld, d_feat, d_wf, d_b, tile_mask, stream);
  if (C == 32 && Hd == 32) return launch_fold_bwd_t<32, 32>(dG, feat, n, HW, w_in, ld, d_feat, d_wf, d_b, tile_mask, stream);
  return cudaErrorInvalidValue;
}

int fold_bwd_barrier_timeouts() {
  int v = 0;
  cudaMemcpyFromSymbol(&v, g_fold_bwd_barrier_timeouts, sizeof(int));
  return v;
}

}  // namespace bts
