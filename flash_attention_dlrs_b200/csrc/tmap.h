// tmap.h — host-side TMA tensor-map encoding without linking libcuda:
// cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint at first use.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <mutex>

namespace fa {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
  });
  return fn;
}

// 16-bit element tensor (B, H, N, D) with element strides {sB, sH, sN, 1}; box = 64 x box_rows (x1x1),
// SWIZZLE_128B, out-of-bounds rows read as zero.  Returns 0 on success.
inline int make_tmap_bhnd_16bit(CUtensorMap* out, const void* base, int is_bf16, int B, int H, int N, int D,
                                int64_t sB, int64_t sH, int64_t sN, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return -100;
  cuuint64_t dims[4] = {(cuuint64_t)D, (cuuint64_t)N, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)sN * 2, (cuuint64_t)sH * 2, (cuuint64_t)sB * 2};
  // size-1 dims may carry arbitrary strides in torch; TMA wants multiples of 16 bytes
  if (H == 1) strides[1] = strides[0] * (cuuint64_t)N;
  if (B == 1) strides[2] = strides[1] * (cuuint64_t)H;
  cuuint32_t box[4] = {64u, (cuuint32_t)box_rows, 1u, 1u};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = enc(out, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(200 + (int)r);
}

// 8-bit element tensor (B, H, N, D) with element (= byte) strides {sB, sH, sN, 1}; box = 128 x box_rows (x1x1) bytes,
// SWIZZLE_128B, out-of-bounds rows read as zero.  Used by the FP8 forward (D = 128: one box per 128-row tile).
inline int make_tmap_bhnd_8bit(CUtensorMap* out, const void* base, int B, int H, int N, int D, int64_t sB, int64_t sH,
                               int64_t sN, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return -100;
  cuuint64_t dims[4] = {(cuuint64_t)D, (cuuint64_t)N, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)sN, (cuuint64_t)sH, (cuuint64_t)sB};
  if (H == 1) strides[1] = strides[0] * (cuuint64_t)N;
  if (B == 1) strides[2] = strides[1] * (cuuint64_t)H;
  cuuint32_t box[4] = {128u, (cuuint32_t)box_rows, 1u, 1u};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(200 + (int)r);
}

// Encoded tensor maps are pure functions of (address, element type, shape, strides, box): keep the last few per thread so
// that a training loop launching on the same buffers does not pay cuTensorMapEncodeTiled (~1 us each, 3-4 per launch)
// on every call.  Thread-local: no locks on the launch path; an entry holds no reference to the memory it describes.
struct TmapKey {
  const void* base;
  int kind;   // 0 f16, 1 bf16, 2 8-bit
  int B, H, N, D, box_rows;
  int64_t sB, sH, sN;
  bool operator==(const TmapKey& o) const {
    return base == o.base && kind == o.kind && B == o.B && H == o.H && N == o.N && D == o.D && box_rows == o.box_rows &&
           sB == o.sB && sH == o.sH && sN == o.sN;
  }
};
inline int cached_tmap_bhnd(CUtensorMap* out, const void* base, int kind, int B, int H, int N, int D, int64_t sB,
                            int64_t sH, int64_t sN, int box_rows) {
  constexpr int kEntries = 16;
  struct Entry {
    TmapKey key;
    CUtensorMap map;
    bool valid;
  };
  thread_local Entry cache[kEntries] = {};
  thread_local int next = 0;
  const TmapKey key{base, kind, B, H, N, D, box_rows, sB, sH, sN};
  for (int i = 0; i < kEntries; ++i)
    if (cache[i].valid && cache[i].key == key) {
      *out = cache[i].map;
      return 0;
    }
  const int r = kind == 2 ? make_tmap_bhnd_8bit(out, base, B, H, N, D, sB, sH, sN, box_rows)
                          : make_tmap_bhnd_16bit(out, base, kind, B, H, N, D, sB, sH, sN, box_rows);
  if (r == 0) {
    cache[next] = Entry{key, *out, true};
    next = (next + 1) % kEntries;
  }
  return r;
}

}  // namespace fa
