// TMA tensor maps for the feature tensor F (n, C, H*W) fp32: one cp.async.bulk.tensor per [C x pixels] tile instead of one
// bulk copy per channel row.  The driver's encoder is looked up through the runtime (no -lcuda link dependency); when it is
// unavailable the kernels fall back to row copies (same shared-memory layout).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace bts {
namespace tmap {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encoder() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// Map of feat (n, C, HW) with a [1][C][box_px] box (dense [C][box_px] fp32 in shared memory, pixels past HW read as zero).
inline bool make_feature_map(CUtensorMap* m, const float* feat, int n, int C, int64_t HW, int box_px) {
  EncodeTiledFn fn = encoder();
  if (!fn) return false;
  const cuuint64_t gdim[3] = {(cuuint64_t)HW, (cuuint64_t)C, (cuuint64_t)n};
  const cuuint64_t gstride[2] = {(cuuint64_t)HW * 4, (cuuint64_t)C * HW * 4};
  const cuuint32_t box[3] = {(cuuint32_t)box_px, (cuuint32_t)C, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(feat), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

#ifdef __CUDACC__
// One tile load: box origin (pixel p0, channel 0, frame b); completes `bytes of the whole box` on the mbarrier.
__device__ __forceinline__ void load_tile_3d(void* smem_dst, const CUtensorMap* map, int p0, int c0, int b, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(p0), "r"(c0), "r"(b), "r"((uint32_t)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void prefetch_map(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
#endif

}  // namespace tmap
}  // namespace bts
