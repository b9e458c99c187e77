// Tensor-core fold: G[b, p, :] = W_in[:, :C] * F[b, :, p] + b_in as a persistent TMA-fed tcgen05 GEMM.
//
// Per 128-pixel tile (M = 128 pixels, N = Hd, K = C):
//   * the raw feature tile F[b, 0:C, p0:p0+128] (C rows of 512 B, NCHW) is fetched by the TMA engine with one
//     cp.async.bulk.tensor box (tensor_map.cuh; one bulk copy per channel row when the driver's tensor-map encoder is
//     unavailable) into a double-buffered staging area (mbarrier complete_tx), so the loads of tile i+1 overlap the work
//     on tile i;
//   * the 128 threads transpose it into the K-major core-matrix layout the MMA reads ([chunk of 4 channels][pixel][16 B])
//     and split every value into tf32 hi + lo on the way (hi = low 13 mantissa bits cleared, lo = exact residual);
//   * one thread issues D[128 x Hd] = A_hi W_hi + A_lo W_hi + A_hi W_lo (kind::tf32; the dropped lo*lo term is ~2^-22
//     relative: fp32-grade, the parity tier is unchanged), accumulator in TMEM;
//   * epilogue: tcgen05.ld of the pixel's row, + bias, 16-byte stores of the 4*Hd contiguous bytes of G.
// Bound: HBM (C*4 B read + Hd*4 B written per pixel); the CUDA-core kernel in fold.cu stays as the general path
// (any C, unaligned frames).  Reference tensors: see fold.cu.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tcgen05.cuh"
#include "tensor_map.cuh"

namespace bts {

__device__ int g_fold_barrier_timeouts = 0;

namespace foldtc {

constexpr int kTile = 128;   // pixels per tile (one tile group = 128 threads)

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc::smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}

template <int HD, int C>
struct L {
  static constexpr int kRaw = C * kTile * 4;                  // one raw tile: [C][128] floats
  static constexpr int kA = (C / 4) * kTile * 16;             // one operand tile: [C/4 chunks][128 rows][16 B]
  static constexpr int kW = (C / 4) * HD * 16;                // weights: [C/4 chunks][HD rows][16 B]
  static constexpr int kOffWhi = 0;
  static constexpr int kOffWlo = kOffWhi + kW;
  static constexpr int kOffBias = kOffWlo + kW;               // HD floats
  static constexpr int kOffBar = kOffBias + HD * 4;           // per group: full, mma (2 mbarriers); then the tmem slot
  static constexpr int kOffGroup = ((kOffBar + 64 + 127) / 128) * 128;
  static constexpr int kGroupBytes = kRaw + 2 * kA;           // raw tile, operand hi, operand lo
  // two independent tile pipelines per CTA when they fit: one group's transpose overlaps the other's MMA + epilogue
  static constexpr int kGroups = (kOffGroup + 2 * kGroupBytes <= 227 * 1024) ? 2 : 1;
  static constexpr int kThreads = kGroups * 128;
  static constexpr int kSmemBytes = kOffGroup + kGroups * kGroupBytes;
  // epilogue staging: the finished operand tiles hold the 128 output rows (padded: conflict-free 16-byte stores), each
  // thread then hands its row to the copy engine (one 4*HD-byte bulk store: full lines instead of 16-byte partial writes)
  static constexpr int kStageRow = HD * 4 + 16;
  static constexpr bool kStage = 2 * kA >= kTile * kStageRow;
  static constexpr uint32_t kGroupCols = HD <= 32 ? 32 : (HD <= 64 ? 64 : 128);
  static constexpr uint32_t kTmemCols = kGroupCols * kGroups;
  static_assert(kSmemBytes <= 227 * 1024, "fold tile does not fit shared memory");
};

template <int HD, int C>
__global__ void __launch_bounds__(L<HD, C>::kThreads, 1) fold_tc_kernel(const float* __restrict__ feat, int64_t HW, int n_batches,
                                                                        const float* __restrict__ w_in, int ld,
                                                                        const float* __restrict__ b_in, float* __restrict__ G,
                                                                        const __grid_constant__ CUtensorMap feat_map, int use_map,
                                                                        unsigned char* __restrict__ tile_mask, int mask_per_b) {
  using Lt = L<HD, C>;
  constexpr int kThreads = Lt::kThreads;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, group = tid >> 7, gt = tid & 127, warp = gt >> 5;   // tile group, thread / warp inside it
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Lt::kOffBar) + 2 * group;
  uint64_t* mma_bar = full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Lt::kOffBar + 32);
  float* s_bias = reinterpret_cast<float*>(smem + Lt::kOffBias);
  unsigned char* g_raw = smem + Lt::kOffGroup + group * Lt::kGroupBytes;
  unsigned char* g_ahi = g_raw + Lt::kRaw;
  unsigned char* g_alo = g_ahi + Lt::kA;

  // ---- weights: tf32 hi/lo, K-major no-swizzle [chunk][n][4 floats]; bias ----
  for (int i = tid; i < HD * C; i += kThreads) {
    const int n = i / C, c = i % C;
    const float w = __ldg(w_in + (size_t)n * ld + c);
    const float hi = __uint_as_float(__float_as_uint(w) & 0xffffe000u);
    const int off = ((c >> 2) * HD + n) * 16 + (c & 3) * 4;
    *reinterpret_cast<float*>(smem + Lt::kOffWhi + off) = hi;
    *reinterpret_cast<float*>(smem + Lt::kOffWlo + off) = w - hi;
  }
  for (int i = tid; i < HD; i += kThreads) s_bias[i] = __ldg(b_in + i);
  if (tid < 32) tc::tmem_alloc<Lt::kTmemCols>(tmem_slot);
  if (gt == 0) {
    tc::mbar_init(full, 1);
    tc::mbar_init(mma_bar, 1);
    tc::fence_mbar_init();
  }
  tc::fence_proxy_async_smem();
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot + group * Lt::kGroupCols;
  const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);
  const uint32_t idesc = tc::instr_desc(tc::kFmtTF32, 128, HD);
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  auto group_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory"); };

  const int64_t tiles_per_b = (HW + kTile - 1) / kTile;
  const int64_t n_tiles = tiles_per_b * n_batches;
  auto issue_load = [&](int64_t tile) {                   // one elected thread: one tensor-map box (or C row copies)
    const int64_t b = tile / tiles_per_b, p0 = (tile % tiles_per_b) * kTile;
    const int valid = (int)((HW - p0) < kTile ? (HW - p0) : kTile);
    if (use_map) {
      mbar_expect_tx(full, (uint32_t)Lt::kRaw);
      tmap::load_tile_3d(g_raw, &feat_map, (int)p0, 0, (int)b, full);
      return;
    }
    const uint32_t row_bytes = (uint32_t)valid * 4u;
    mbar_expect_tx(full, row_bytes * C);
    const float* src = feat + ((size_t)b * C) * HW + p0;
    for (int c = 0; c < C; ++c) bulk_load(g_raw + c * kTile * 4, src + (size_t)c * HW, row_bytes, full);
  };

  uint32_t full_phase = 0, mma_phase = 0;
  const int64_t stride = (int64_t)gridDim.x * Lt::kGroups;
  // Sparse fold (training): tile_mask has one byte per 64 pixels of a frame -- 0 nothing samples it, 1 needed, 2 folded
  // already; a 128-pixel tile is produced when one of its halves is needed, and both halves are then marked folded.
  auto mask_of = [&](int64_t t) { return tile_mask + (t / tiles_per_b) * mask_per_b + (t % tiles_per_b) * 2; };
  auto needed = [&](int64_t t) {
    if (!tile_mask) return true;
    const unsigned char* m = mask_of(t);
    return m[0] == 1 || ((t % tiles_per_b) * 2 + 1 < mask_per_b && m[1] == 1);
  };
  auto advance = [&](int64_t t) {                         // first tile of this group at or after t that has to be folded
    while (t < n_tiles && !needed(t)) t += stride;
    return t;
  };
  int64_t tile = advance((int64_t)blockIdx.x * Lt::kGroups + group);
  if (tid == 0 && use_map) tmap::prefetch_map(&feat_map);
  if (gt == 0 && tile < n_tiles) issue_load(tile);
  int64_t next;
  for (; tile < n_tiles; tile = next) {
    next = advance(tile + stride);
    const int64_t b = tile / tiles_per_b, p0 = (tile % tiles_per_b) * kTile;
    const int valid = (int)((HW - p0) < kTile ? (HW - p0) : kTile);

    tc::mbar_wait_or_trap(full, full_phase, &g_fold_barrier_timeouts);   // bounded: a pipeline bug traps instead of hanging or running on
    full_phase ^= 1u;
    if (Lt::kStage) {                      // the previous tile's row stores have read the staging rows (= the operand tiles)
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      group_sync();
    }
    // ---- transpose + split: thread = pixel row ----
    {
      const float* raw = reinterpret_cast<const float*>(g_raw);
#pragma unroll 4
      for (int kc = 0; kc < C / 4; ++kc) {
        float f[4], h[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          f[j] = gt < valid ? raw[(kc * 4 + j) * kTile + gt] : 0.0f;
          h[j] = __uint_as_float(__float_as_uint(f[j]) & 0xffffe000u);
        }
        *reinterpret_cast<float4*>(g_ahi + (kc * kTile + gt) * 16) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(g_alo + (kc * kTile + gt) * 16) = make_float4(f[0] - h[0], f[1] - h[1], f[2] - h[2], f[3] - h[3]);
      }
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    group_sync();                          // operand tiles complete, the raw tile is consumed; the previous accumulator has been read
    if (warp_u == 0 && tc::elect_one()) {  // converged warp, one elected lane: operands stay in uniform registers
      if (next < n_tiles) issue_load(next);                              // lands under the MMAs and the epilogue
      tc::tc_fence_after_sync();
      const uint32_t sb = tc::smem_u32(smem), sa = tc::smem_u32(g_ahi), sl = tc::smem_u32(g_alo);
#pragma unroll
      for (int ks = 0; ks < C / 8; ++ks) {
        const uint64_t ahi = tc::smem_desc_kmajor_noswizzle(sa + ks * 2 * kTile * 16, kTile * 16, 128);
        const uint64_t alo = tc::smem_desc_kmajor_noswizzle(sl + ks * 2 * kTile * 16, kTile * 16, 128);
        const uint64_t whi = tc::smem_desc_kmajor_noswizzle(sb + Lt::kOffWhi + ks * 2 * HD * 16, HD * 16, 128);
        const uint64_t wlo = tc::smem_desc_kmajor_noswizzle(sb + Lt::kOffWlo + ks * 2 * HD * 16, HD * 16, 128);
        tc::mma_tf32_ss(tmem, ahi, whi, idesc, ks > 0);
        tc::mma_tf32_ss(tmem, alo, whi, idesc, 1);
        tc::mma_tf32_ss(tmem, ahi, wlo, idesc, 1);
      }
      tc::mma_commit(mma_bar);
    }
    __syncwarp();
    tc::mbar_wait_or_trap(mma_bar, mma_phase, &g_fold_barrier_timeouts);   // bounded: a pipeline bug traps instead of hanging or running on
    mma_phase ^= 1u;
    tc::tc_fence_after_sync();
    // ---- epilogue: accumulator row + bias -> G ----
    {
      float* dst = G + ((size_t)b * HW + p0 + gt) * HD;
      unsigned char* stage = g_ahi + gt * Lt::kStageRow;
#pragma unroll
      for (int c0 = 0; c0 < HD; c0 += 16) {
        uint32_t v[16];
        tc::tmem_ld_x16(t_lane + c0, v);
        tc::tmem_ld_wait();
        if (gt < valid) {
#pragma unroll
          for (int e = 0; e < 16; e += 4) {
            const float4 o = make_float4(__uint_as_float(v[e]) + s_bias[c0 + e], __uint_as_float(v[e + 1]) + s_bias[c0 + e + 1],
                                         __uint_as_float(v[e + 2]) + s_bias[c0 + e + 2], __uint_as_float(v[e + 3]) + s_bias[c0 + e + 3]);
            if (Lt::kStage) *reinterpret_cast<float4*>(stage + (c0 + e) * 4) = o;
            else *reinterpret_cast<float4*>(dst + c0 + e) = o;
          }
        }
      }
      if (Lt::kStage) {
        tc::fence_proxy_async_smem();      // my row -> async proxy
        if (gt < valid)
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(tc::smem_u32(stage)), "n"(HD * 4) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      if (tile_mask && gt == 0) {
        unsigned char* m = mask_of(tile);
        m[0] = 2;
        if ((tile % tiles_per_b) * 2 + 1 < mask_per_b) m[1] = 2;
      }
    }
    tc::tc_fence_before_sync();
  }
  if (Lt::kStage) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // my row stores are complete before the CTA's shared memory goes away
  __syncthreads();
  if (tid < 32) tc::tmem_dealloc<Lt::kTmemCols>(*tmem_slot);
}

}  // namespace foldtc

bool fold_tc_supported(const float* feat, int C, int64_t HW, int Hd) {
  return (C == 32 || C == 64) && (Hd == 32 || Hd == 64 || Hd == 128) && HW % 4 == 0 && (reinterpret_cast<uintptr_t>(feat) & 15) == 0;
}

template <int HD, int C>
static cudaError_t launch_fold_tc_t(const float* feat, int n, int64_t HW, const float* w_in, int ld, const float* b_in, float* G,
                                    unsigned char* tile_mask, cudaStream_t stream) {
  using Lt = foldtc::L<HD, C>;
  auto kern = foldtc::fold_tc_kernel<HD, C>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Lt::kSmemBytes);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t n_tiles = ((HW + foldtc::kTile - 1) / foldtc::kTile) * n;
  const int64_t want = (n_tiles + Lt::kGroups - 1) / Lt::kGroups;
  const int grid = (int)(want < sms ? want : sms);
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  const char* rows = getenv("BTS_FOLD_TMA_ROWS");            // tests: force the row-copy loads
  const int use_map = !(rows && rows[0] == '1') && tmap::make_feature_map(&map, feat, n, C, HW, foldtc::kTile) ? 1 : 0;
  kern<<<grid, Lt::kThreads, Lt::kSmemBytes, stream>>>(feat, HW, n, w_in, ld, b_in, G, map, use_map, tile_mask, (int)((HW + 63) / 64));
  return cudaGetLastError();
}

cudaError_t launch_fold_tc(const float* feat, int n, int C, int64_t HW, const float* w_in, int ld, const float* b_in, int Hd,
                           float* G, unsigned char* tile_mask, cudaStream_t stream) {
  if (C == 64) {
    if (Hd == 64) return launch_fold_tc_t<64, 64>(feat, n, HW, w_in, ld, b_in, G, tile_mask, stream);
    if (Hd == 128) return launch_fold_tc_t<128, 64>(feat, n, HW, w_in, ld, b_in, G, tile_mask, stream);
    if (Hd == 32) return launch_fold_tc_t<32, 64>(feat, n, HW, w_in, ld, b_in, G, tile_mask, stream);
  } else if (C == 32) {
    if (Hd == 32) return launch_fold_tc_t<32, 32>(feat, n, HW, w_in, ld, b_in, G, tile_mask, stream);
    if (Hd == 64) return launch_fold_tc_t<64, 32>(feat, n, HW, w_in, ld, b_in, G, tile_mask, stream);
    if (Hd == 128) return launch_fold_tc_t<128, 32>(feat, n, HW, w_in, ld, b_in, G, tile_mask, stream);
  }
  return cudaErrorInvalidValue;
}

int fold_barrier_timeouts() {
  int v = 0;
  cudaMemcpyFromSymbol(&v, g_fold_barrier_timeouts, sizeof(int));
  return v;
}

}  // namespace bts
