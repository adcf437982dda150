// umma8_probe.cu — bring-up check of tcgen05.mma kind::f8f6f4 with E4M3 / E5M2 operands as the FP8 forward uses them:
//   case SS : D[128 x 128] = A[128 x 128] (K-major smem) * B[128 x 128]^T (K-major smem)          (S = Q K^T)
//   case TS : D[128 x 128] = A[128 x 128] (TMEM, 4 values per 32-bit column) * B (MN-major smem: [K rows x N bytes])  (O = P V)
// Each is checked against a host fp32 reference on the exactly representable FP8 inputs.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/umma8_probe tools/umma8_probe.cu
#include "../flash_attention_dlrs_b200/csrc/sm100_ptx.cuh"
#include "../flash_attention_dlrs_b200/csrc/tmap.h"

#include <cuda_fp8.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace fa;

struct P8 {
  int a_tmem, b_mn, is_e5m2;
};

__global__ void __launch_bounds__(128, 1)
probe8(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const uint8_t* A_gmem,
       float* D_out, P8 p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;           // 128 rows x 128 B
  uint8_t* sB = smem + 16384;   // 128 rows x 128 B
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<256>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s, tmem_D = tmem, tmem_A = tmem + 128;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar_load, p.a_tmem ? 16384 : 32768);
    if (!p.a_tmem) tma_load_4d(sA, &tmA, &bar_load, 0, 0, 0, 0);
    tma_load_4d(sB, &tmB, &bar_load, 0, 0, 0, 0);
  }
  if (p.a_tmem) {   // thread r owns row r of A: 4 consecutive K values per 32-bit column
    const uint32_t* arow = reinterpret_cast<const uint32_t*>(A_gmem + (size_t)threadIdx.x * 128);
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = arow[i];
    tmem_st_x32(tmem_A + ((uint32_t)(warp * 32) << 16), v);
    tc_wait_st();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    mbar_wait(&bar_load, 0);
    tc_fence_after();
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_f8(p.is_e5m2, 128, 128, 0, p.b_mn);
      const uint32_t a_lo = umma_lo_kmajor(smem_u32(sA));
      const uint32_t b_lo = p.b_mn ? umma_lo_mnmajor(smem_u32(sB), 16384) : umma_lo_kmajor(smem_u32(sB));
      for (int k = 0; k < 4; ++k) {   // 32 elements of K per instruction
        const uint32_t boff = p.b_mn ? (uint32_t)(k * 4096) >> 4 : (uint32_t)(k * 32) >> 4;
        if (p.a_tmem)
          umma8_ts_off<0, 0>(tmem_D, tmem_A + k * 8, b_lo + boff, idesc, k > 0);
        else
          umma8_ss_off<0, 0>(tmem_D, a_lo + ((uint32_t)(k * 32) >> 4), b_lo + boff, idesc, k > 0);
      }
      tc_commit(&bar_mma);
    }
    __syncwarp();
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t v[32];
    tmem_ld_x32(tmem_D + ((uint32_t)(warp * 32) << 16) + c0, v);
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i) D_out[(size_t)threadIdx.x * 128 + c0 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

template <typename F8>
static int run(const P8& p, const char* name) {
  const int M = 128, N = 128, K = 128;
  std::vector<uint8_t> hA(M * K), hB(N * K);
  std::vector<float> fA(M * K), fB(N * K);
  srand(77 + p.a_tmem * 3 + p.b_mn * 5 + p.is_e5m2);
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) {
      F8 q(((rand() % 2001) - 1000) / 400.0f);
      hA[m * K + k] = *reinterpret_cast<uint8_t*>(&q);
      fA[m * K + k] = float(q);
    }
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      F8 q(((rand() % 2001) - 1000) / 400.0f);
      hB[p.b_mn ? k * N + n : n * K + k] = *reinterpret_cast<uint8_t*>(&q);
      fB[n * K + k] = float(q);
    }
  uint8_t *dA, *dB;
  float* dD;
  cudaMalloc(&dA, hA.size());
  cudaMalloc(&dB, hB.size());
  cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, M * N * 4);
  CUtensorMap tmA, tmB;
  int r1 = make_tmap_bhnd_8bit(&tmA, dA, 1, 1, 128, 128, 128 * 128, 128 * 128, 128, 128);
  int r2 = make_tmap_bhnd_8bit(&tmB, dB, 1, 1, 128, 128, 128 * 128, 128 * 128, 128, 128);
  if (r1 || r2) { printf("%s: tensor map encode failed %d %d\n", name, r1, r2); return 1; }
  cudaFuncSetAttribute(probe8, cudaFuncAttributeMaxDynamicSharedMemorySize, 34 * 1024);
  probe8<<<1, 128, 34 * 1024>>>(tmA, tmB, dA, dD, p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: kernel failed: %s\n", name, cudaGetErrorString(e)); exit(3); }
  std::vector<float> hD(M * N);
  cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = 0;
      for (int k = 0; k < K; ++k) acc += (double)fA[m * K + k] * fB[n * K + k];
      double d = fabs(acc - (double)hD[m * N + n]);
      if (!(d <= 1e30)) d = 1e30;
      maxerr = d > maxerr ? d : maxerr;
      maxref = fabs(acc) > maxref ? fabs(acc) : maxref;
    }
  printf("%-52s max|err| = %.3e (max|ref| %.1f)  %s\n", name, maxerr, maxref, maxerr < 1e-2 ? "OK" : "MISMATCH");
  return maxerr < 1e-2 ? 0 : 1;
}

int main() {
  int fails = 0;
  for (int e5 = 0; e5 < 2; ++e5) {
    P8 p{};
    p.is_e5m2 = e5;
    p.a_tmem = 0, p.b_mn = 0;
    fails += e5 ? run<__nv_fp8_e5m2>(p, "E5M2  SS  A K-major smem, B K-major") : run<__nv_fp8_e4m3>(p, "E4M3  SS  A K-major smem, B K-major");
    p.a_tmem = 1, p.b_mn = 1;
    fails += e5 ? run<__nv_fp8_e5m2>(p, "E5M2  TS  A TMEM (4 per column), B MN-major") : run<__nv_fp8_e4m3>(p, "E4M3  TS  A TMEM (4 per column), B MN-major");
    p.a_tmem = 0, p.b_mn = 1;
    fails += e5 ? run<__nv_fp8_e5m2>(p, "E5M2  SS  A K-major smem, B MN-major") : run<__nv_fp8_e4m3>(p, "E4M3  SS  A K-major smem, B MN-major");
  }
  printf("probe8: %d case(s) failed\n", fails);
  return fails ? 1 : 0;
}
