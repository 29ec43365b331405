// Host-side TMA tensor-map encoding for 2-D row-major fp32 tensors.  cuTensorMapEncodeTiled is a host-only driver
// function: resolved through the runtime (no link dependency on libcuda).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200ppo {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Tensor map over T[rows][cols] (row pitch cols * 4 bytes, a multiple of 16) with a box_cols x box_rows box; elements
// outside the tensor read as zero.  swizzle128: the box lands in the 128-byte swizzle pattern with 32-byte atoms
// (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, what an MN-major tf32 UMMA operand needs; box_cols * 4 <= 128), else densely.  Returns 0, or a CUDA / B200PPO error code.
inline int encode_tiled_2d(CUtensorMap* tm, const float* base, int cols, int rows, int box_cols, int box_rows,
                           bool swizzle128) {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    const cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess) return static_cast<int>(e);
    if (q != cudaDriverEntryPointSuccess || p == nullptr) return -2;   // B200PPO_ELIMIT
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t gstr[1] = {static_cast<cuuint64_t>(cols) * 4u};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -1;                                   // B200PPO_EINVAL
}

// The same row-major tensor T[rows][cols] seen as {32 columns, rows, ceil(cols / 32) column groups} with strides
// {4, cols * 4, 128} bytes, box {32, box_rows, box_atoms}, 32-byte-atom 128-byte swizzle: ONE TMA instruction lands
// box_atoms MN-major operand atoms (tc.cuh) of 32 columns x box_rows rows back to back.  Column groups beyond the
// tensor read as zero; the tail of a last, partial group reads the first columns of the next row (finite values the
// caller must not use: they only reach output rows / columns beyond the operand's width).
inline int encode_mn_atoms_3d(CUtensorMap* tm, const float* base, int cols, int rows, int box_rows, int box_atoms) {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    const cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess) return static_cast<int>(e);
    if (q != cudaDriverEntryPointSuccess || p == nullptr) return -2;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  const cuuint64_t gdim[3] = {32u, static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>((cols + 31) / 32)};
  const cuuint64_t gstr[2] = {static_cast<cuuint64_t>(cols) * 4u, 128u};
  const cuuint32_t box[3] = {32u, static_cast<cuuint32_t>(box_rows), static_cast<cuuint32_t>(box_atoms)};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -1;
}

}  // namespace b200ppo
