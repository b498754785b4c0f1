// tmap.h — host-side construction of TMA tensor maps (cuTensorMapEncodeTiled).
//
// The driver entry point is resolved at run time through the CUDA runtime
// (cudaGetDriverEntryPoint), so the library links against libcudart only.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace carel {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) !=
            cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

inline void ensure_context() {
  typedef CUresult (*GetCurFn)(CUcontext*);
  static GetCurFn get_cur = nullptr;
  if (!get_cur) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuCtxGetCurrent", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      get_cur = reinterpret_cast<GetCurFn>(p);
  }
  CUcontext ctx = nullptr;
  if (get_cur == nullptr || get_cur(&ctx) != CUDA_SUCCESS || ctx == nullptr) cudaFree(nullptr);
}

// Row-major matrix of 2-byte elements (bf16 / fp16) [rows, cols] with row pitch `ld_elems` elements; TMA box =
// `box_cols` x `box_rows`; swizzle 128 B (box_cols = 64) or 64 B (box_cols = 32).  Out-of-bounds elements are
// zero-filled on loads and dropped on stores.
inline int make_tmap_2b(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                        uint32_t box_cols, uint32_t box_rows, bool fp16) {
  EncodeTiledFn fn = get_encode_tiled();
  if (!fn) return -1;
  // cuTensorMapEncodeTiled is a driver-API call and needs a current context; a thread that has not
  // made a runtime call yet (e.g. autograd's backward thread) has none bound: bind the primary one.
  // (Checked first so that nothing but the encode itself happens while a stream is being captured.)
  ensure_context();
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : int(r);
}

// bf16, 64-column boxes, 128-byte swizzle (the operand layout of every UMMA in this library)
inline int make_tmap_bf16_sw128(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols,
                                uint64_t ld_elems, uint32_t box_rows) {
  return make_tmap_2b(m, base, rows, cols, ld_elems, 64, box_rows, false);
}
inline int make_tmap_f16_sw128(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols,
                               uint64_t ld_elems, uint32_t box_rows) {
  return make_tmap_2b(m, base, rows, cols, ld_elems, 64, box_rows, true);
}

}  // namespace carel
