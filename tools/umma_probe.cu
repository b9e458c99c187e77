// Probe: D[128x64] = A[128xK] * B[64xK]^T with tcgen05.mma (cta_group::1), four operand paths.
//   0: SS kind::tf32 (A,B in smem, K-major no-swizzle)   1: SS kind::f16
//   2: TS kind::tf32 (A in TMEM)                          3: TS kind::f16
// Inputs are small dyadic rationals, exactly representable in fp16/tf32, so a correct layout gives a bit-exact D.
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include <vector>
#include "../behindthescenes_b200/csrc/tcgen05.cuh"
using namespace bts::tc;

__global__ void __launch_bounds__(128) probe(const float* A, const float* B, float* D, int K, int variant, int* status) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool mixed = variant >= 4;
  const bool f16 = mixed ? true : (variant & 1), ts = mixed ? (variant == 5) : (variant >= 2);
  if (mixed) K -= 8;                       // first K-8 columns go through kind::f16, last 8 through kind::tf32
  const int epc = f16 ? 8 : 4;            // elements per 16-byte chunk
  const int chunks = K / epc;
  unsigned char* sA = smem;                               // [chunk][128 rows][16 B]
  unsigned char* sB = smem + chunks * 128 * 16;            // [chunk][64 rows][16 B]
  unsigned char* sA2 = sB + chunks * 64 * 16;              // mixed: tf32 tail of A  [2][128][16 B]
  unsigned char* sB2 = sA2 + 2 * 128 * 16;                 // mixed: tf32 tail of B  [2][64][16 B]
  const int ldk = mixed ? K + 8 : K;
  if (warp == 0) tmem_alloc<128>(&tmem_base_s);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase = tmem_base_s;
  const uint32_t d_t = tbase, a_t = tbase + 64;
  // ---- operands ----
  for (int c = 0; c < chunks; ++c) {
    if (!ts) {
      if (f16) {
        __half h[8];
        for (int e = 0; e < 8; ++e) h[e] = __float2half(A[tid * ldk + c * 8 + e]);
        *reinterpret_cast<uint4*>(sA + (c * 128 + tid) * 16) = *reinterpret_cast<uint4*>(h);
      } else {
        *reinterpret_cast<float4*>(sA + (c * 128 + tid) * 16) = *reinterpret_cast<const float4*>(A + tid * ldk + c * 4);
      }
    }
    if (tid < 64) {
      if (f16) {
        __half h[8];
        for (int e = 0; e < 8; ++e) h[e] = __float2half(B[tid * ldk + c * 8 + e]);
        *reinterpret_cast<uint4*>(sB + (c * 64 + tid) * 16) = *reinterpret_cast<uint4*>(h);
      } else {
        *reinterpret_cast<float4*>(sB + (c * 64 + tid) * 16) = *reinterpret_cast<const float4*>(B + tid * ldk + c * 4);
      }
    }
  }
  if (ts) {
    const int cols = f16 ? K / 2 : K;
    const uint32_t lane_addr = a_t + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < cols; c0 += 8) {
      uint32_t v[8];
      for (int e = 0; e < 8; ++e) {
        if (f16) {
          __half2 h = __floats2half2_rn(A[tid * ldk + (c0 + e) * 2], A[tid * ldk + (c0 + e) * 2 + 1]);
          v[e] = *reinterpret_cast<uint32_t*>(&h);
        } else {
          v[e] = __float_as_uint(A[tid * ldk + c0 + e]);
        }
      }
      tmem_st_x8(lane_addr + c0, v);
    }
    if (mixed) {   // tf32 tail in TMEM columns right after the f16 part
      uint32_t v[8];
      for (int e = 0; e < 8; ++e) v[e] = __float_as_uint(A[tid * ldk + K + e]);
      tmem_st_x8(lane_addr + cols, v);
    }
    tmem_st_wait();
  }
  if (mixed) {
    for (int c = 0; c < 2; ++c) {
      *reinterpret_cast<float4*>(sA2 + (c * 128 + tid) * 16) = *reinterpret_cast<const float4*>(A + tid * ldk + K + c * 4);
      if (tid < 64) *reinterpret_cast<float4*>(sB2 + (c * 64 + tid) * 16) = *reinterpret_cast<const float4*>(B + tid * ldk + K + c * 4);
    }
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after_sync();
    const uint32_t idesc = instr_desc(f16 ? kFmtF16 : kFmtTF32, 128, 64);
    const int ksteps = K / (f16 ? 16 : 8);
    for (int ks = 0; ks < ksteps; ++ks) {
      const uint64_t bd = smem_desc_kmajor_noswizzle(smem_u32(sB) + ks * 2 * 64 * 16, 64 * 16, 128);
      if (ts) {
        if (f16) mma_f16_ts(d_t, a_t + ks * 8, bd, idesc, ks > 0);
        else mma_tf32_ts(d_t, a_t + ks * 8, bd, idesc, ks > 0);
      } else {
        const uint64_t ad = smem_desc_kmajor_noswizzle(smem_u32(sA) + ks * 2 * 128 * 16, 128 * 16, 128);
        if (f16) mma_f16_ss(d_t, ad, bd, idesc, ks > 0);
        else mma_tf32_ss(d_t, ad, bd, idesc, ks > 0);
      }
    }
    if (mixed) {
      const uint32_t idesc2 = instr_desc(kFmtTF32, 128, 64);
      const uint64_t bd2 = smem_desc_kmajor_noswizzle(smem_u32(sB2), 64 * 16, 128);
      if (ts) mma_tf32_ts(d_t, a_t + K / 2, bd2, idesc2, 1);
      else mma_tf32_ss(d_t, smem_desc_kmajor_noswizzle(smem_u32(sA2), 128 * 16, 128), bd2, idesc2, 1);
    }
    mma_commit(&bar);
  }
  __syncwarp();
  const bool ok = mbar_wait(&bar, 0, 1u << 22);
  if (!ok) { if (tid == 0) *status = 1; }
  tc_fence_after_sync();
  if (ok) {
    for (int c0 = 0; c0 < 64; c0 += 16) {
      uint32_t v[16];
      tmem_ld_x16(d_t + ((uint32_t)(warp * 32) << 16) + c0, v);
      tmem_ld_wait();
      for (int e = 0; e < 16; ++e) D[tid * 64 + c0 + e] = __uint_as_float(v[e]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<128>(tbase);
}

int main() {
  const char* names[6] = {"SS tf32 K=40", "SS f16 K=48", "TS tf32 K=40", "TS f16 K=48", "SS f16 K=48 + tf32 K=8 mixed", "TS f16 K=48 + tf32 K=8 mixed"};
  int fails = 0;
  for (int variant = 0; variant < 6; ++variant) {
    const int K = variant >= 4 ? 56 : ((variant & 1) ? 48 : 40);
    std::vector<float> A(128 * K), B(64 * K), D(128 * 64, -777.f), R(128 * 64);
    srand(123 + variant);
    for (auto& x : A) x = (float)((rand() % 33) - 16) / 8.f;
    for (auto& x : B) x = (float)((rand() % 33) - 16) / 8.f;
    for (int i = 0; i < 128; ++i) for (int j = 0; j < 64; ++j) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[i * K + k] * B[j * K + k]; R[i * 64 + j] = (float)s; }
    float *dA, *dB, *dD; int* dS; int st = 0;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dS, 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dS, &st, 4, cudaMemcpyHostToDevice);
    const size_t smem = variant >= 4 ? (size_t)6 * 192 * 16 + 2 * 192 * 16 : (size_t)(K / ((variant & 1) ? 8 : 4)) * (128 + 64) * 16;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<<<1, 128, smem>>>(dA, dB, dD, K, variant, dS);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (size_t i = 0; i < D.size(); ++i) { double d = fabs((double)D[i] - R[i]); if (d > maxerr) maxerr = d; if (d > 1e-3) ++bad; }
    printf("variant %d (%s): cuda=%s barrier_timeout=%d max_err=%g bad=%d/%zu  D[0..3]=%g %g %g %g  R=%g %g %g %g\n", variant, names[variant],
           cudaGetErrorString(e), st, maxerr, bad, D.size(), D[0], D[1], D[2], D[3], R[0], R[1], R[2], R[3]);
    if (e != cudaSuccess) { printf("aborting after CUDA error\n"); return 2; }
    if (bad || st) ++fails;
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS);
  }
  printf("umma_probe: %d failing variants\n", fails);
  return fails ? 1 : 0;
}
