// Host-side TMA tensor-map construction without linking libcuda: cuTensorMapEncodeTiled is fetched through
// cudaGetDriverEntryPoint so the shared library also loads (and exports its symbols) on machines with no driver.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

typedef CUresult (*indm_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                         const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                         CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline indm_encode_tiled_fn indm_get_encode_tiled() {
  static indm_encode_tiled_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
      cudaGetLastError();
      return nullptr;
    }
    fn = (indm_encode_tiled_fn)p;
  }
  return fn;
}

// rank-R map over a tensor whose fastest dimension is contiguous; dims/box listed fastest-first;
// strides_bytes[i] = byte stride of dim i+1 (R-1 entries).  128-byte swizzle, zero fill out of bounds.
// elem_strides (optional): traversal stride per dimension; a box of extent box[i] then delivers ceil(box[i] / stride) elements
// (how the stride-2 convolution of the input pyramid samples every other pixel without a gather kernel).
static inline int indm_make_tmap(CUtensorMap* out, CUtensorMapDataType dt, int rank, const void* base,
                                 const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                                 const char* what, const uint32_t* elem_strides = nullptr,
                                 CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  indm_encode_tiled_fn fn = indm_get_encode_tiled();
  if (!fn) {
    indm_set_error("%s: cuTensorMapEncodeTiled unavailable (no CUDA driver?)", what);
    return INDM_ERR_CUDA;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides ? elem_strides[i] : 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  if (((uintptr_t)base & 15) != 0) {
    indm_set_error("%s: tensor base %p not 16-byte aligned", what, base);
    return INDM_ERR_ARG;
  }
  for (int i = 0; i + 1 < rank; ++i) {
    if (gstr[i] % 16 != 0) {
      indm_set_error("%s: stride[%d]=%llu bytes is not a multiple of 16", what, i, (unsigned long long)gstr[i]);
      return INDM_ERR_ARG;
    }
  }
  CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    indm_set_error("%s: cuTensorMapEncodeTiled failed with CUresult %d (dims %llu,%llu,%llu,%llu box %u,%u,%u,%u)", what,
                   (int)r, (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0),
                   (unsigned long long)(rank > 2 ? gdim[2] : 0), (unsigned long long)(rank > 3 ? gdim[3] : 0), bx[0],
                   rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0, rank > 3 ? bx[3] : 0);
    return INDM_ERR_CUDA;
  }
  return INDM_OK;
}
