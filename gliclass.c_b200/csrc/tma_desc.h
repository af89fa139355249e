// Host-side TMA tensor-map construction.  cuTensorMapEncodeTiled is resolved through
// cudaGetDriverEntryPoint so the library has no link-time dependency on libcuda.so (the build
// container has no driver).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <stdexcept>
#include <string>

namespace glc {

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p)
      throw std::runtime_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 16-bit-element (fp16 / bf16) tensor of rank 2 or 3, innermost dim contiguous; box inner extent must be 64 elements
// (128 bytes; data is moved verbatim, so one UINT16 map serves fp16 and bf16) for the 128-byte swizzle used by every tcgen05 operand in this engine.
// dims/box are innermost-first; strides_bytes has rank-1 entries (dims 1..rank-1).
inline CUtensorMap make_tmap_16b(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                                  const uint32_t* box, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  CUtensorMap m;
  cuuint64_t gdims[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { gdims[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i < rank - 1; ++i) gstr[i] = strides_bytes[i];
  CUresult r = get_encode_tiled()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT16, (cuuint32_t)rank, const_cast<void*>(base), gdims,
                                  gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  swizzle,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed, CUresult=" + std::to_string((int)r));
  return m;
}

// Byte-element tensor (e4m3 operands and outputs): same conventions, strides in bytes = elements.
inline CUtensorMap make_tmap_8b(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                                 const uint32_t* box, CUtensorMapSwizzle swizzle) {
  CUtensorMap m;
  cuuint64_t gdims[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { gdims[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i < rank - 1; ++i) gstr[i] = strides_bytes[i];
  CUresult r = get_encode_tiled()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstr,
                                  bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled (uint8) failed, CUresult=" + std::to_string((int)r));
  return m;
}

}  // namespace glc
