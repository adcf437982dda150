// tmem_bw_probe.cu — how fast can warps read TMEM (tcgen05.ld 32x32b.x32) on one SM, and what do ex2 / cvt cost?
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/tmem_bw_probe tools/tmem_bw_probe.cu
#include "../flash_attention_dlrs_b200/csrc/fa_bwd_fused_sm100.cuh"
#include <cstdio>
using namespace fa;

__global__ void __launch_bounds__(512, 1) probe(int mode, int iters, long long* out, float* sink) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  if (mode == 0) {          // TMEM loads only: 2 x32 in flight
    for (int i = 0; i < iters; ++i) {
      uint32_t a[32], b[32];
      tmem_ld_x32(t + ((i * 64) & 255), a);
      tmem_ld_x32(t + ((i * 64 + 32) & 255), b);
      tc_wait_ld();
      acc += __uint_as_float(a[0] ^ b[31] ^ a[17] ^ b[5]);
    }
  } else if (mode == 1) {   // ex2 only: 64 per iteration
    float x = threadIdx.x * 1e-3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 64; ++k) acc += ex2_approx(x + k * 0.01f + acc * 1e-9f);
    }
  } else if (mode == 3 || mode == 4) {   // P^T stage of the fused backward (mode 4: + the dS arithmetic)
    __shared__ __align__(16) float stat[256];
    stat[threadIdx.x & 255] = -1.0f * (threadIdx.x & 7);
    __syncthreads();
    const uint64_t sl2 = f32x2_pack(0.127f, 0.127f);
    for (int i = 0; i < iters; ++i) {
      uint32_t pf[64];
      fused_p_stage<true, false>(t, smem_u32(stat), sl2, threadIdx.x & 127, 0, pf);
      tc_wait_st();
      if (mode == 4) {
        uint32_t pd[32];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t dr[32];
          tmem_ld_x32(t + 128 + c * 32, dr);
          tc_wait_ld();
#pragma unroll
          for (int g4 = 0; g4 < 8; ++g4) {
            uint64_t nd4[2];
            lds_f32x2x2(smem_u32(stat) + (128 + c * 32 + g4 * 4) * 4, nd4[0], nd4[1]);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int e = c * 32 + g4 * 4 + u * 2;
              float d0, d1;
              f32x2_unpack(f32x2_mul(f32x2_pack_bits(pf[e], pf[e + 1]),
                                     f32x2_add(f32x2_pack_bits(dr[g4 * 4 + u * 2], dr[g4 * 4 + u * 2 + 1]), nd4[u])), d0, d1);
              pd[e >> 1] = pack2<true>(d0, d1);
            }
          }
        }
        tmem_st_x32(t + 128, pd);
        tc_wait_st();
      }
      acc += __uint_as_float(pf[3]);
    }
  } else if (mode == 2) {   // TMEM stores x32
    uint32_t a[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) a[k] = threadIdx.x + k;
    for (int i = 0; i < iters; ++i) {
      tmem_st_x32(t + ((i * 32) & 255), a);
      tmem_st_x32(t + ((i * 32 + 128) & 255), a);
      tc_wait_st();
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  sink[threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base_s);
}

int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 8); cudaMalloc(&sink, 4096);
  const int iters = 2000;
  for (int mode = 0; mode < 5; ++mode)
    for (int warps : {1, 4, 8, 12, 16}) {
      probe<<<1, warps * 32, 0>>>(mode, iters, out, sink);
      long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      if (mode == 0) printf("tmem ld x32x2: warps %2d  %.1f clk/iter  -> %.1f B/clk/SM\n", warps, (double)c / iters, warps * 8192.0 * iters / c);
      if (mode == 1) printf("ex2 x64      : warps %2d  %.1f clk/iter  -> %.2f ex2/clk/SM\n", warps, (double)c / iters, warps * 32 * 64.0 * iters / c);
      if (mode == 3) printf("P stage      : warps %2d  %.1f clk/iter\n", warps, (double)c / iters);
      if (mode == 4) printf("P + dS stage : warps %2d  %.1f clk/iter\n", warps, (double)c / iters);
      if (mode == 2) printf("tmem st x32x2: warps %2d  %.1f clk/iter  -> %.1f B/clk/SM\n", warps, (double)c / iters, warps * 8192.0 * iters / c);
    }
  return 0;
}
