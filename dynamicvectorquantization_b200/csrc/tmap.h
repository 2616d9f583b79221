// Host-side TMA descriptor (CUtensorMap) construction without linking libcuda:
// cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint at first use.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2 {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) !=
            cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// bf16 tensor, `rank` dims (innermost first), strides in ELEMENTS for dims 1..rank-1
// (dim 0 is contiguous), box in elements, 128 B swizzle, zero fill out of bounds.
// Returns 0 on success, a negative code otherwise.
inline int make_tmap_elem(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_elems, const uint32_t* box, bool f32);
inline int make_tmap_bf16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_elems, const uint32_t* box) {
  return make_tmap_elem(tm, base, rank, dims, strides_elems, box, false);
}
// Same for bf16 (f32 = false) or fp32 (f32 = true) elements; the inner box extent must span 128 B.
inline int make_tmap_elem(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_elems, const uint32_t* box, bool f32) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return -100;
  // cuTensorMapEncodeTiled is a driver call and needs a current context; a thread that has not made
  // a runtime call yet (e.g. PyTorch's autograd worker entering our backward first) has none bound.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaFree(0);
    ctx_bound = true;
  }
  cuuint64_t gdims[5];
  cuuint64_t gstr[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_elems[i] * (f32 ? 4 : 2);  // bytes
  }
  CUresult r = enc(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base),
                   gdims, gstr, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(200 + (int)r);
}

}  // namespace b2
