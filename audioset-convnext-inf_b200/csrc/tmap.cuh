// Host-side TMA tensor-map construction (cuTensorMapEncodeTiled resolved through the runtime so the
// library links against cudart only and still loads on a box without libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace acx {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline PFN_encodeTiled get_encode_tiled() {
  static std::atomic<PFN_encodeTiled> fn{nullptr};        // idempotent lookup; racing threads store the same pointer
  PFN_encodeTiled f = fn.load(std::memory_order_acquire);
  if (!f) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      f = reinterpret_cast<PFN_encodeTiled>(p);
      fn.store(f, std::memory_order_release);
    }
  }
  return f;
}

// Generic rank-R bf16 map, 128B swizzle, OOB reads return zero.
// dims[0] is the contiguous dimension; strides_bytes[i] is the byte stride of dims[i+1].
static inline int make_tmap_bf16(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims,
                                 const cuuint64_t* strides_bytes, const cuuint32_t* box,
                                 CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  PFN_encodeTiled enc = get_encode_tiled();
  ACX_CHECK(enc != nullptr, ACX_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims,
                   strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ACX_CHECK(r == CUDA_SUCCESS, ACX_ERR_CUDA,
            "cuTensorMapEncodeTiled failed (CUresult %d; rank %d dims %llu x %llu, stride %llu B, box %u x %u)", (int)r,
            rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)strides_bytes[0],
            box[0], box[1]);
  return ACX_OK;
}

// Row-major (rows, cols) bf16 matrix, box = box_cols x box_rows.
static inline int make_tmap_2d_bf16(CUtensorMap* tm, const void* base, uint64_t cols, uint64_t rows,
                                    uint64_t row_stride_bytes, uint32_t box_cols, uint32_t box_rows,
                                    CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  return make_tmap_bf16(tm, base, 2, dims, strides, box, swizzle);
}

// Group-planar activation tensor [groups][rows_pad][8] bf16 (16-byte channel groups as planes; the plane stride is the
// row count rounded up to 128, gp_rows_pad).  A plane is contiguous, so the map views it as rows of 256 elements (32
// tensor rows x 16 B): (256, rows_pad / 32, groups), box {256, box_rows / 32, box_groups}, no swizzle -> smem
// [group][row][16 B], the un-swizzled canonical K-major operand layout.  (First version: inner box = one 16-byte piece --
// the TMA unit then works piece by piece, ~14 B/clk/SM for loads and ~9 for stores, tools/ubench/tma_gather.cu, and the
// fused MLP lost 25 us per layer to it.)  Tensor rows >= rows only exist as padding: loads return whatever is there,
// stores land in the padding.
static inline uint64_t gp_rows_pad(uint64_t rows) { return (rows + 127) / 128 * 128; }
static inline int make_tmap_gp_bf16(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t groups, uint32_t box_rows,
                                    uint32_t box_groups) {
  const uint64_t rp = gp_rows_pad(rows);
  cuuint64_t dims[3] = {256, rp / 32, groups};
  cuuint64_t strides[2] = {512, rp * 16};
  cuuint32_t box[3] = {256, box_rows / 32, box_groups};
  return make_tmap_bf16(tm, base, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
}

}  // namespace acx
