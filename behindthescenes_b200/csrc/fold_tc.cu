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

constexpr int kThreads = 128;
constexpr int kTile = 128;   // pixels per tile

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
  static constexpr int kOffRaw = 0;                           // x2
  static constexpr int kOffAhi = 2 * kRaw;
  static constexpr int kOffAlo = kOffAhi + kA;
  static constexpr int kOffWhi = kOffAlo + kA;
  static constexpr int kOffWlo = kOffWhi + kW;
  static constexpr int kOffBias = kOffWlo + kW;               // HD floats
  static constexpr int kOffBar = kOffBias + HD * 4;           // full[2], mma : 3 mbarriers + tmem slot
  static constexpr int kSmemBytes = kOffBar + 64;
  static constexpr uint32_t kTmemCols = HD <= 32 ? 32 : (HD <= 64 ? 64 : 128);
  static_assert(kSmemBytes <= 227 * 1024, "fold tile does not fit shared memory");
};

template <int HD, int C>
__global__ void __launch_bounds__(kThreads, 1) fold_tc_kernel(const float* __restrict__ feat, int64_t HW, int n_batches,
                                                               const float* __restrict__ w_in, int ld,
                                                               const float* __restrict__ b_in, float* __restrict__ G,
                                                               const __grid_constant__ CUtensorMap feat_map, int use_map) {
  using Lt = L<HD, C>;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Lt::kOffBar);        // [2]
  uint64_t* mma_bar = full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full + 3);
  float* s_bias = reinterpret_cast<float*>(smem + Lt::kOffBias);

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
  const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);
  const uint32_t idesc = tc::instr_desc(tc::kFmtTF32, 128, HD);

  const int64_t tiles_per_b = (HW + kTile - 1) / kTile;
  const int64_t n_tiles = tiles_per_b * n_batches;
  auto issue_load = [&](int64_t tile, int buf) {          // one elected thread: one tensor-map box (or C row copies)
    const int64_t b = tile / tiles_per_b, p0 = (tile % tiles_per_b) * kTile;
    const int valid = (int)((HW - p0) < kTile ? (HW - p0) : kTile);
    unsigned char* dst = smem + Lt::kOffRaw + buf * Lt::kRaw;
    if (use_map) {
      mbar_expect_tx(&full[buf], (uint32_t)Lt::kRaw);
      tmap::load_tile_3d(dst, &feat_map, (int)p0, 0, (int)b, &full[buf]);
      return;
    }
    const uint32_t row_bytes = (uint32_t)valid * 4u;
    mbar_expect_tx(&full[buf], row_bytes * C);
    const float* src = feat + ((size_t)b * C) * HW + p0;
    for (int c = 0; c < C; ++c) bulk_load(dst + c * kTile * 4, src + (size_t)c * HW, row_bytes, &full[buf]);
  };

  uint32_t full_phase[2] = {0, 0}, mma_phase = 0;
  int64_t tile = blockIdx.x;
  if (tid == 0 && use_map) tmap::prefetch_map(&feat_map);
  if (tid == 0 && tile < n_tiles) issue_load(tile, 0);
  int buf = 0;
  for (; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
    const int64_t next = tile + gridDim.x;
    if (tid == 0 && next < n_tiles) issue_load(next, buf ^ 1);            // raw[buf^1] was consumed two iterations ago
    const int64_t b = tile / tiles_per_b, p0 = (tile % tiles_per_b) * kTile;
    const int valid = (int)((HW - p0) < kTile ? (HW - p0) : kTile);

    tc::mbar_wait_or_trap(&full[buf], full_phase[buf], &g_fold_barrier_timeouts);   // bounded: a pipeline bug traps instead of hanging or running on
    full_phase[buf] ^= 1u;
    // ---- transpose + split: thread = pixel row ----
    {
      const float* raw = reinterpret_cast<const float*>(smem + Lt::kOffRaw + buf * Lt::kRaw);
#pragma unroll 4
      for (int kc = 0; kc < C / 4; ++kc) {
        float f[4], h[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          f[j] = tid < valid ? raw[(kc * 4 + j) * kTile + tid] : 0.0f;
          h[j] = __uint_as_float(__float_as_uint(f[j]) & 0xffffe000u);
        }
        *reinterpret_cast<float4*>(smem + Lt::kOffAhi + (kc * kTile + tid) * 16) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(smem + Lt::kOffAlo + (kc * kTile + tid) * 16) = make_float4(f[0] - h[0], f[1] - h[1], f[2] - h[2], f[3] - h[3]);
      }
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();                       // operand tiles complete; the previous tile's accumulator has been read
    if (tid == 0) {
      tc::tc_fence_after_sync();
      const uint32_t sb = tc::smem_u32(smem);
#pragma unroll
      for (int ks = 0; ks < C / 8; ++ks) {
        const uint64_t ahi = tc::smem_desc_kmajor_noswizzle(sb + Lt::kOffAhi + ks * 2 * kTile * 16, kTile * 16, 128);
        const uint64_t alo = tc::smem_desc_kmajor_noswizzle(sb + Lt::kOffAlo + ks * 2 * kTile * 16, kTile * 16, 128);
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
      float* dst = G + ((size_t)b * HW + p0 + tid) * HD;
#pragma unroll
      for (int c0 = 0; c0 < HD; c0 += 16) {
        uint32_t v[16];
        tc::tmem_ld_x16(t_lane + c0, v);
        tc::tmem_ld_wait();
        if (tid < valid) {
#pragma unroll
          for (int e = 0; e < 16; e += 4)
            *reinterpret_cast<float4*>(dst + c0 + e) =
                make_float4(__uint_as_float(v[e]) + s_bias[c0 + e], __uint_as_float(v[e + 1]) + s_bias[c0 + e + 1],
                            __uint_as_float(v[e + 2]) + s_bias[c0 + e + 2], __uint_as_float(v[e + 3]) + s_bias[c0 + e + 3]);
        }
      }
    }
    tc::tc_fence_before_sync();
  }
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<Lt::kTmemCols>(tmem);
}

}  // namespace foldtc

bool fold_tc_supported(const float* feat, int C, int64_t HW, int Hd) {
  return (C == 32 || C == 64) && (Hd == 32 || Hd == 64 || Hd == 128) && HW % 4 == 0 && (reinterpret_cast<uintptr_t>(feat) & 15) == 0;
}

template <int HD, int C>
static cudaError_t launch_fold_tc_t(const float* feat, int n, int64_t HW, const float* w_in, int ld, const float* b_in, float* G,
                                    cudaStream_t stream) {
  using Lt = foldtc::L<HD, C>;
  auto kern = foldtc::fold_tc_kernel<HD, C>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Lt::kSmemBytes);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t n_tiles = ((HW + foldtc::kTile - 1) / foldtc::kTile) * n;
  const int grid = (int)(n_tiles < sms ? n_tiles : sms);
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  const char* rows = getenv("BTS_FOLD_TMA_ROWS");            // tests: force the row-copy loads
  const int use_map = !(rows && rows[0] == '1') && tmap::make_feature_map(&map, feat, n, C, HW, foldtc::kTile) ? 1 : 0;
  kern<<<grid, foldtc::kThreads, Lt::kSmemBytes, stream>>>(feat, HW, n, w_in, ld, b_in, G, map, use_map);
  return cudaGetLastError();
}

cudaError_t launch_fold_tc(const float* feat, int n, int C, int64_t HW, const float* w_in, int ld, const float* b_in, int Hd,
                           float* G, cudaStream_t stream) {
  if (C == 64) {
    if (Hd == 64) return launch_fold_tc_t<64, 64>(feat, n, HW, w_in, ld, b_in, G, stream);
    if (Hd == 128) return launch_fold_tc_t<128, 64>(feat, n, HW, w_in, ld, b_in, G, stream);
    if (Hd == 32) return launch_fold_tc_t<32, 64>(feat, n, HW, w_in, ld, b_in, G, stream);
  } else if (C == 32) {
    if (Hd == 32) return launch_fold_tc_t<32, 32>(feat, n, HW, w_in, ld, b_in, G, stream);
    if (Hd == 64) return launch_fold_tc_t<64, 32>(feat, n, HW, w_in, ld, b_in, G, stream);
    if (Hd == 128) return launch_fold_tc_t<128, 32>(feat, n, HW, w_in, ld, b_in, G, stream);
  }
  return cudaErrorInvalidValue;
}

int fold_barrier_timeouts() {
  int v = 0;
  cudaMemcpyFromSymbol(&v, g_fold_barrier_timeouts, sizeof(int));
  return v;
}

}  // namespace bts
