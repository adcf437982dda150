// umma_probe.cu — stand-alone bring-up check of the sm_100a building blocks used by the attention kernels:
// TMA SWIZZLE_128B loads, tcgen05.mma with K-major / MN-major shared-memory operands, A operand read from
// TMEM (packed 16-bit pairs written with tcgen05.st), tcgen05.ld of the fp32 accumulator.
// Each case computes D[128 x N] = A[128 x K] * B and is checked against a host fp32 reference.
//
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_probe tools/umma_probe.cu
//   run  : ./umma_probe            (exit code 0 = every mandatory case matched)
#include "../flash_attention_dlrs_b200/csrc/sm100_ptx.cuh"
#include "../flash_attention_dlrs_b200/csrc/tmap.h"

#include <cmath>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace fa;

struct ProbeParams {
  int a_tmem;    // 0: A from smem (K-major), 1: A from TMEM
  int a_mn;      // A from smem stored [K x 128] (MN-major): the dS operand of the fused backward's dQ product
  int b_mn;      // 0: B is [N x K] (K-major), 1: B is [K x N] (MN-major)
  int N, K;      // N in {64,128}, K in {64,128}
  int is_bf16;
  uint32_t lbo, sbo;  // MN-major B descriptor fields (bytes)
  uint32_t kstep;     // MN-major B: start-address advance per 16 of K (bytes)
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const uint16_t* __restrict__ A_gmem, float* __restrict__ D_out, ProbeParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;               // up to 2 boxes x 16 KB
  uint8_t* sB = smem + 32768;       // up to 2 boxes x 16 KB
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
  const uint32_t tmem = tmem_base_s;
  const uint32_t tmem_D = tmem;         // columns [0, N)
  const uint32_t tmem_A = tmem + 128;   // columns [128, 128 + K/2)

  const int a_boxes = p.K / 64;
  const int b_rows = p.b_mn ? p.K : p.N;        // rows of one B box
  const int b_boxes = p.b_mn ? p.N / 64 : p.K / 64;

  if (threadIdx.x == 0) {
    uint32_t bytes = b_boxes * b_rows * 128;
    if (!p.a_tmem) bytes += p.K * 128 * 2;
    mbar_arrive_expect_tx(&bar_load, bytes);
    if (!p.a_tmem) {
      if (p.a_mn)   // [K rows x 128 columns]: two boxes of K rows x 64 columns
        for (int b = 0; b < 2; ++b) tma_load_4d(sA + b * p.K * 128, &tmA, &bar_load, b * 64, 0, 0, 0);
      else
        for (int b = 0; b < a_boxes; ++b) tma_load_4d(sA + b * 16384, &tmA, &bar_load, b * 64, 0, 0, 0);
    }
    for (int b = 0; b < b_boxes; ++b) tma_load_4d(sB + b * b_rows * 128, &tmB, &bar_load, b * 64, 0, 0, 0);
  }
  if (p.a_tmem) {
    // thread r owns row r of A: pack pairs (k, k+1) into one 32-bit TMEM column
    const uint16_t* arow = A_gmem + (size_t)threadIdx.x * p.K;
    for (int c0 = 0; c0 < p.K / 2; c0 += 16) {
      uint32_t v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        uint32_t lo = arow[2 * (c0 + i)], hi = arow[2 * (c0 + i) + 1];
        v[i] = lo | (hi << 16);
      }
      tmem_st_x16(tmem_A + ((uint32_t)(warp * 32) << 16) + c0, v);
    }
    tc_wait_st();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    mbar_wait(&bar_load, 0);
    tc_fence_after();
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_f16(p.is_bf16, 128, p.N, p.a_mn, p.b_mn);
      for (int k = 0; k < p.K / 16; ++k) {
        uint64_t db;
        if (p.b_mn)
          db = umma_smem_desc_sw128(smem_u32(sB) + k * p.kstep, p.lbo, p.sbo);
        else
          db = umma_desc_kmajor(smem_u32(sB) + (k / 4) * (p.N * 128), k % 4);
        if (p.a_tmem) {
          umma_ts(tmem_D, tmem_A + k * 8, db, idesc, k > 0);
        } else {
          uint64_t da = p.a_mn ? umma_desc_mnmajor(smem_u32(sA), (uint32_t)p.K * 128u, k)
                               : umma_desc_kmajor(smem_u32(sA) + (k / 4) * 16384, k % 4);
          umma_ss(tmem_D, da, db, idesc, k > 0);
        }
      }
      tc_commit(&bar_mma);
    }
    __syncwarp();
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  {
    const int row = threadIdx.x;
    for (int c0 = 0; c0 < p.N; c0 += 32) {
      uint32_t v[32];
      tmem_ld_x32(tmem_D + ((uint32_t)(warp * 32) << 16) + c0, v);
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) D_out[(size_t)row * p.N + c0 + i] = __uint_as_float(v[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

static uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static float bf2f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e_ = (x);                                                            \
    if (e_ != cudaSuccess) {                                                         \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                       \
    }                                                                                \
  } while (0)

static double run_case(const ProbeParams& p, const char* name) {
  const int M = 128, N = p.N, K = p.K;
  std::vector<uint16_t> hA(M * K), hB(N * K);
  std::vector<float> fA(M * K), fB(N * K);
  srand(1234 + N * 7 + K * 3 + p.a_tmem * 11 + p.b_mn * 5);
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) {
      float v = (float)(rand() % 2001 - 1000) / 500.0f;
      uint16_t h = f2bf(v);
      hA[p.a_mn ? k * M + m : m * K + k] = h;
      fA[m * K + k] = bf2f(h);
    }
  // B logical: Bmat[n][k]; storage is [N x K] (K-major) or [K x N] (MN-major)
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      float v = (float)(rand() % 2001 - 1000) / 500.0f;
      uint16_t h = f2bf(v);
      if (p.b_mn) {
        hB[k * N + n] = h;
      } else {
        hB[n * K + k] = h;
      }
      fB[n * K + k] = bf2f(h);
    }
  uint16_t *dA, *dB;
  float* dD;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dD, M * N * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, M * N * 4));
  CUtensorMap tmA, tmB;
  int r1 = p.a_mn ? make_tmap_bhnd_16bit(&tmA, dA, 1, 1, 1, K, M, (int64_t)M * K, (int64_t)M * K, M, K)
                  : make_tmap_bhnd_16bit(&tmA, dA, 1, 1, 1, M, K, (int64_t)M * K, (int64_t)M * K, K, 128);
  int r2 = p.b_mn ? make_tmap_bhnd_16bit(&tmB, dB, 1, 1, 1, K, N, (int64_t)K * N, (int64_t)K * N, N, K)
                  : make_tmap_bhnd_16bit(&tmB, dB, 1, 1, 1, N, K, (int64_t)N * K, (int64_t)N * K, K, N);
  if (r1 || r2) {
    printf("%s: tensor map encode failed %d %d\n", name, r1, r2);
    return 1e30;
  }
  const int smem_bytes = 65536 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  probe_kernel<<<1, 128, smem_bytes>>>(tmA, tmB, dA, dD, p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%s: kernel failed: %s\n", name, cudaGetErrorString(e));
    exit(3);  // context is dead after a trap
  }
  std::vector<float> hD(M * N);
  CK(cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = 0;
      for (int k = 0; k < K; ++k) acc += (double)fA[m * K + k] * fB[n * K + k];
      double d = fabs(acc - (double)hD[m * N + n]);
      if (!(d <= 1e30)) d = 1e30;
      if (d > maxerr) maxerr = d;
    }
  printf("%-44s N=%3d K=%3d lbo=%5u sbo=%5u kstep=%5u  max|err| = %.3e  %s\n", name, N, K, p.lbo, p.sbo,
         p.kstep, maxerr, maxerr < 1e-2 ? "OK" : "MISMATCH");
  cudaFree(dA);
  cudaFree(dB);
  cudaFree(dD);
  return maxerr;
}

int main() {
  int fails = 0;
  for (int N : {128, 64})
    for (int K : {128, 64}) {
      ProbeParams p{};
      p.is_bf16 = 1;
      p.N = N;
      p.K = K;
      // K-major SS
      p.a_tmem = 0; p.b_mn = 0;
      fails += run_case(p, "SS  A K-major smem, B K-major") > 1e-2;
      // K-major B with A from TMEM
      p.a_tmem = 1; p.b_mn = 0;
      fails += run_case(p, "TS  A TMEM,         B K-major") > 1e-2;
      // MN-major B: expected encoding (LBO = box stride, SBO = 1024, +2048 B per 16 of K)
      p.b_mn = 1; p.lbo = (uint32_t)K * 128; p.sbo = 1024; p.kstep = 2048;
      p.a_tmem = 0;
      double e1 = run_case(p, "SS  A K-major smem, B MN-major (expected)");
      p.a_tmem = 1;
      double e2 = run_case(p, "TS  A TMEM,         B MN-major (expected)");
      fails += (e1 > 1e-2) + (e2 > 1e-2);
      // MN-major A from shared memory (dQ = dS K in the fused backward: both operands indexed by key row)
      p.a_tmem = 0; p.a_mn = 1;
      fails += run_case(p, "SS  A MN-major smem, B MN-major") > 1e-2;
      p.a_mn = 0;
      if (e1 > 1e-2) {
        // diagnostics only: alternative readings of the LBO/SBO fields
        p.a_tmem = 0;
        p.lbo = 1024; p.sbo = (uint32_t)K * 128;
        run_case(p, "SS  B MN-major alt: LBO/SBO swapped");
        p.lbo = (uint32_t)K * 128; p.sbo = 2048;
        run_case(p, "SS  B MN-major alt: SBO=2048");
        p.lbo = 16; p.sbo = 1024;
        run_case(p, "SS  B MN-major alt: LBO=16");
      }
    }
  printf("probe: %d mandatory case(s) failed\n", fails);
  return fails ? 1 : 0;
}
