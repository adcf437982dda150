// fa_preprocess.cuh — backward preprocess: delta[b,h,i] = sum_d O[b,h,i,d] * dO[b,h,i,d].
//
// Role of the reference's bwd_D_kernel (flash_attention_kernels.py:120-166), with two deliberate
// differences: the product is accumulated in fp32 (the reference multiplies in the input dtype,
// :165) and delta is stored as contiguous fp32 (B,H,N) instead of the input dtype.
//
// HBM-bound: algorithmic bytes = 2*B*H*N*D*sizeof(elt) + 4*B*H*N.  Each thread owns one 16-byte
// chunk of a row, kRows rows in flight per thread; a row's D/kVec threads are adjacent lanes and
// reduce with shuffles.
#pragma once

#include "sm100_ptx.cuh"

namespace fa {

struct PreParams {
  const void* o;
  const void* dout;
  float* delta;  // (B,H,N) contiguous
  int B, H, N, D;
  int64_t o_sB, o_sH, o_sN;
  int64_t do_sB, do_sH, do_sN;
  long long total_rows;  // B*H*N
};

__device__ __forceinline__ uint4 ld_nc_16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// kElt: 0 = f16, 1 = bf16, 2 = f32.  kTPR = threads per row = D / elements-per-16-bytes (power of two <= 32).
template <int kElt, int kTPR>
__global__ void __launch_bounds__(256) fa_bwd_preprocess_kernel(const PreParams p) {
  constexpr int kVec = (kElt == 2) ? 4 : 8;      // elements per 16-byte chunk
  constexpr int kEltBytes = (kElt == 2) ? 4 : 2;
  constexpr int kRowsPerPass = 256 / kTPR;       // rows one CTA covers per pass
  constexpr int kUnroll = 4;                     // passes in flight per thread
  const int sub = threadIdx.x % kTPR;            // chunk inside the row
  const int rloc = threadIdx.x / kTPR;
  const long long stride_rows = (long long)gridDim.x * kRowsPerPass * kUnroll;

  for (long long base = (long long)blockIdx.x * kRowsPerPass * kUnroll; base < p.total_rows; base += stride_rows) {
    uint4 a[kUnroll], g[kUnroll];
    long long row[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      row[u] = base + (long long)u * kRowsPerPass + rloc;
      a[u] = make_uint4(0, 0, 0, 0);
      g[u] = make_uint4(0, 0, 0, 0);
      if (row[u] < p.total_rows) {
        const long long i = row[u] % p.N;
        const long long bh = row[u] / p.N;
        const long long h = bh % p.H, b = bh / p.H;
        const char* po = static_cast<const char*>(p.o) +
                         (b * p.o_sB + h * p.o_sH + i * p.o_sN + (long long)sub * kVec) * kEltBytes;
        const char* pg = static_cast<const char*>(p.dout) +
                         (b * p.do_sB + h * p.do_sH + i * p.do_sN + (long long)sub * kVec) * kEltBytes;
        a[u] = ld_nc_16(po);
        g[u] = ld_nc_16(pg);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      float acc;
      if constexpr (kElt == 2) {
        acc = __uint_as_float(a[u].x) * __uint_as_float(g[u].x);
        acc = fmaf(__uint_as_float(a[u].y), __uint_as_float(g[u].y), acc);
        acc = fmaf(__uint_as_float(a[u].z), __uint_as_float(g[u].z), acc);
        acc = fmaf(__uint_as_float(a[u].w), __uint_as_float(g[u].w), acc);
      } else {
        const uint32_t av[4] = {a[u].x, a[u].y, a[u].z, a[u].w};
        const uint32_t gv[4] = {g[u].x, g[u].y, g[u].z, g[u].w};
        acc = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 x = unpack2<kElt == 1>(av[e]);
          const float2 y = unpack2<kElt == 1>(gv[e]);
          acc = fmaf(x.x, y.x, acc);
          acc = fmaf(x.y, y.y, acc);
        }
      }
#pragma unroll
      for (int off = kTPR / 2; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
      if (sub == 0 && row[u] < p.total_rows) p.delta[row[u]] = acc;
    }
  }
}

}  // namespace fa
