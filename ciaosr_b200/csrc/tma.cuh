// Tensor-map (TMA) helpers shared by the kernels that load operand tiles with cp.async.bulk.tensor:
// host-side descriptor encoding through the driver entry point (no -lcuda link dependency) and the
// device-side tile loads.
#pragma once
#include <cuda.h>
#include "tc_common.cuh"

namespace ciaosr {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn tma_get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

#ifdef CIAOSR_SPLIT_BF16
constexpr CUtensorMapDataType TMA_SPLIT_DTYPE = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
#else
constexpr CUtensorMapDataType TMA_SPLIT_DTYPE = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
#endif

// 2-D map over a row-major 16-bit matrix [rows, cols] (cols % 8 == 0): box = 64 columns x 128 rows, 128B swizzle =
// exactly one [128 x 64] SW128 K-major operand slab; rows / columns past the matrix read as zero.
inline int tma_make_map_2d(CUtensorMap* m, void* base, long long rows, int cols) {
  EncodeTiledFn enc = tma_get_encode();
  CIAOSR_REQUIRE(enc != nullptr, CIAOSR_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)cols * 2};
  const cuuint32_t box[2] = {64, 128};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(m, TMA_SPLIT_DTYPE, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CIAOSR_REQUIRE(r == CUDA_SUCCESS, CIAOSR_E_CUDA, "cuTensorMapEncodeTiled (2-D) failed (%d)", (int)r);
  return CIAOSR_OK;
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}
// Un-swizzled 2-D map over a byte blob seen as rows of 64 16-bit elements (128 B): a box of [64 x 64 rows] is a plain
// 8 KB linear copy.  Used to load PRE-SWIZZLED weight slabs through the tensor path, whose cta_group::2 form may signal
// the mbarrier of the peer (leader) CTA -- the plain bulk copy cannot.
inline int tma_make_map_linear_rows(CUtensorMap* m, const void* base, long long rows) {
  EncodeTiledFn enc = tma_get_encode();
  CIAOSR_REQUIRE(enc != nullptr, CIAOSR_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t gdim[2] = {64, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {128};
  const cuuint32_t box[2] = {64, 64};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CIAOSR_REQUIRE(r == CUDA_SUCCESS, CIAOSR_E_CUDA, "cuTensorMapEncodeTiled (linear rows) failed (%d)", (int)r);
  return CIAOSR_OK;
}
// tile load into THIS CTA's shared memory whose completion bytes are counted on the mbarrier at the same offset in the
// LEADER CTA of the pair (peer bit of the shared-window address cleared, as CUTLASS' SM100_TMA_2SM_LOAD does)
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  const uint32_t leader_bar = bar & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(leader_bar)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

}  // namespace ciaosr
