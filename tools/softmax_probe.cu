// softmax_probe.cu — stand-alone timing of the forward kernel's exp phase (the part of the softmax stage that follows the
// row max): 128 fp32 scores per thread in registers -> p = exp2(s * sl2 - m) (a share on the FMA pipe) -> row sum ->
// packed bf16 pairs -> TMEM.  No tensor-core work runs beside it; W warps per SM (4 = one per sub-partition, 8 = two).
// Variants of the SAME arithmetic (identical results) that only differ in how the work is laid out for the scheduler:
//   0  the kernel's loop: four 32-column chunks, each chunk's 16 pairs unrolled, one TMEM store per chunk
//   1  two 64-column chunks (two TMEM stores of 32 registers)
//   2  per chunk: all scale FMAs first, then all exponentials, then sums and packs (three passes over the chunk)
//   3  one pass over all 128 columns, single TMEM store sequence at the end
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/softmax_probe tools/softmax_probe.cu
#include "../flash_attention_dlrs_b200/csrc/sm100_ptx.cuh"

#include <cstdio>
#include <cstdlib>

using namespace fa;

#ifndef POLY_MASK
#define POLY_MASK 0x92
#endif

template <int kVariant>
__device__ __forceinline__ float exp_phase(uint32_t (&sr)[128], uint32_t tS, float sl2, float neg_ms) {
  const uint64_t sl2_2 = f32x2_pack(sl2, sl2), nm2 = f32x2_pack(neg_ms, neg_ms);
  uint64_t ls[4] = {0ull, 0ull, 0ull, 0ull};
  auto one_pair = [&](int e, uint32_t& out, int i) {
    const uint64_t x2 = f32x2_fma(f32x2_pack_bits(sr[e], sr[e + 1]), sl2_2, nm2);
    float x0, x1, p0, p1;
    f32x2_unpack(x2, x0, x1);
    if ((POLY_MASK >> (i & 7)) & 1) ex2_poly_x2(x0, x1, p0, p1);
    else p0 = ex2_approx(x0), p1 = ex2_approx(x1);
    ls[i & 3] = f32x2_add(ls[i & 3], f32x2_pack(p0, p1));
    out = pack2<true>(p0, p1);
  };
  if constexpr (kVariant == 0) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) one_pair(c * 32 + 2 * i, pk[i], i);
      tmem_st_x16(tS + c * 16, pk);
    }
  } else if constexpr (kVariant == 1) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) one_pair(c * 64 + 2 * i, pk[i], i);
      tmem_st_x32(tS + c * 32, pk);
    }
  } else if constexpr (kVariant == 2) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t pk[16];
      float xs[32], ps[32];
#pragma unroll
      for (int i = 0; i < 16; ++i)
        f32x2_unpack(f32x2_fma(f32x2_pack_bits(sr[c * 32 + 2 * i], sr[c * 32 + 2 * i + 1]), sl2_2, nm2), xs[2 * i], xs[2 * i + 1]);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if ((POLY_MASK >> (i & 7)) & 1) ex2_poly_x2(xs[2 * i], xs[2 * i + 1], ps[2 * i], ps[2 * i + 1]);
        else ps[2 * i] = ex2_approx(xs[2 * i]), ps[2 * i + 1] = ex2_approx(xs[2 * i + 1]);
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        ls[i & 3] = f32x2_add(ls[i & 3], f32x2_pack(ps[2 * i], ps[2 * i + 1]));
        pk[i] = pack2<true>(ps[2 * i], ps[2 * i + 1]);
      }
      tmem_st_x16(tS + c * 16, pk);
    }
  } else {
    uint32_t pk[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) one_pair(2 * i, pk[i], i);
    tmem_st_x32(tS, *reinterpret_cast<uint32_t(*)[32]>(&pk[0]));
    tmem_st_x32(tS + 32, *reinterpret_cast<uint32_t(*)[32]>(&pk[32]));
  }
  float la, lb, lc, ld;
  f32x2_unpack(f32x2_add(ls[0], ls[1]), la, lb);
  f32x2_unpack(f32x2_add(ls[2], ls[3]), lc, ld);
  return (la + lb) + (lc + ld);
}

template <int kVariant>
__global__ void __launch_bounds__(384, 1) probe(int iters, float sl2, long long* out, float* sink) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tS = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  // seed the score columns
  {
    uint32_t z[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) z[i] = __float_as_uint(-0.01f * (float)((threadIdx.x * 7 + i * 13) & 255));
    for (int c = 0; c < 4; ++c) tmem_st_x32(tS + c * 32, z);
    tc_wait_st();
  }
  __syncthreads();
  float l = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t sr[128];
#pragma unroll
    for (int c = 0; c < 4; ++c) tmem_ld_x32(tS + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&sr[c * 32]));
    tc_wait_ld();
    float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
    for (int c = 0; c < 128; c += 4) {
      mx0 = fmaxf(mx0, __uint_as_float(sr[c]));
      mx1 = fmaxf(mx1, __uint_as_float(sr[c + 1]));
      mx2 = fmaxf(mx2, __uint_as_float(sr[c + 2]));
      mx3 = fmaxf(mx3, __uint_as_float(sr[c + 3]));
    }
    const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
    if constexpr (kVariant == 9) {   // no exp phase at all: what the load, the max and the restore cost
      l += mx;
    } else {
      l += exp_phase<kVariant>(sr, tS, sl2, -mx * sl2);
    }
    tc_wait_st();
    // restore scores for the next round (not timed separately: the same for every variant)
#pragma unroll
    for (int c = 0; c < 4; ++c) tmem_st_x32(tS + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&sr[c * 32]));
    tc_wait_st();
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = l;
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

template <int kVariant>
static void run(int warps) {
  long long* out;
  float* sink;
  cudaMalloc(&out, 8);
  cudaMalloc(&sink, 148 * 384 * 4);
  const int iters = 2000;
  probe<kVariant><<<148, warps * 32>>>(iters, 0.1275f, out, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  long long c;
  cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
  float s0;
  cudaMemcpy(&s0, sink, 4, cudaMemcpyDeviceToHost);
  printf("variant %d  warps %2d : %7.1f clk per 128-column row block (incl. TMEM load, max, restore)   checksum %.6e\n", kVariant, warps,
         (double)c / iters, (double)s0);
  cudaFree(out), cudaFree(sink);
}

int main() {
  for (int w : {4, 8}) {
    run<0>(w);
    run<1>(w);
    run<2>(w);
    run<3>(w);
    run<9>(w);
  }
  return 0;
}
