// umma2_probe.cu — bring-up check and micro-benchmark of CTA-pair (cta_group::2) tcgen05 MMAs on sm_100a.
//
// Part 1 (semantics): D[256 x 128] = A[256 x 128] * B^T computed by ONE tcgen05.mma.cta_group::2 stream issued by the
// leader CTA of a 2-CTA cluster.  CTA c holds rows [128c, 128c+128) of A and HALF of B:
//   case S  (scores-like)  : A from shared memory (K-major), B = [N x K] K-major, CTA c holds B rows n in [64c, 64c+64)
//   case PV (P.V-like)     : A from TMEM (packed 16-bit pairs), B = [K x N] MN-major, CTA c holds B columns n in [64c, 64c+64)
// Each CTA reads its own 128 TMEM lanes; the host checks against an fp32 reference and, on a mismatch, prints which
// reference row / column every output row / column matches best (to read the actual operand split off the hardware).
// TMA loads use the .cta_group::2 form that signals the LEADER's mbarrier; tcgen05.commit multicasts to both CTAs.
//
// Part 2 (throughput): clocks per MMA instruction (M128 N128 K16 per CTA = 64 clk at the tensor peak) for back-to-back
// streams on every SM at once: SS (both operands in shared memory: 128 B/clk of operand reads), TS (A in TMEM: 64 B/clk),
// the same with a concurrent TMA fill stream into other shared-memory buffers, and the cta_group::2 SS stream
// (B halved: 96 B/clk per CTA).  This is the measurement behind "the attention kernels are bound by shared-memory
// bandwidth" (DESIGN.md).
//
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/umma2_probe tools/umma2_probe.cu
#include "../flash_attention_dlrs_b200/csrc/sm100_ptx.cuh"
#include "../flash_attention_dlrs_b200/csrc/tmap.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace fa;

// ------------------------------------------------------------------------------------------------ 2-CTA MMA (plain descriptors)
// (cluster helpers, cta_group::2 TMEM allocation, the leader-barrier TMA form and the multicast commit: sm100_ptx.cuh)
__device__ __forceinline__ void umma2_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma2_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
// ------------------------------------------------------------------------------------------------ part 1: semantics
struct P1 {
  int pv;   // 0: case S, 1: case PV
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB64,
            const __grid_constant__ CUtensorMap tmB128, const uint16_t* __restrict__ A_gmem, float* __restrict__ D_out,
            P1 p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;           // 128 rows x 128 K: two boxes of 128 x 64 (16 KiB each)
  uint8_t* sB = smem + 32768;   // case S: two boxes of 64 rows x 64 K (8 KiB each); case PV: one box of 128 K-rows x 64 N
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();

  if (threadIdx.x == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc2<256>(&tmem_base_s);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tmem_D = tmem, tmem_A = tmem + 128;

  const uint32_t leader_bar = mapa_u32(smem_u32(&bar_load), 0);
  if (threadIdx.x == 0) {
    // the leader arms its barrier with the bytes of BOTH CTAs
    const uint32_t per_cta = 16384u + (p.pv ? 0u : 32768u);
    if (rank == 0) mbar_arrive_expect_tx(&bar_load, 2 * per_cta);
    if (!p.pv) {
      for (int b = 0; b < 2; ++b) tma_load_4d_2sm(sA + b * 16384, &tmA, leader_bar, b * 64, 128 * rank, 0, 0);
      for (int b = 0; b < 2; ++b) tma_load_4d_2sm(sB + b * 8192, &tmB64, leader_bar, b * 64, 64 * rank, 0, 0);
    } else {
      tma_load_4d_2sm(sB, &tmB128, leader_bar, 64 * rank, 0, 0, 0);   // V columns [64 rank, +64), all 128 K rows
    }
  }
  if (p.pv) {   // P rows of this CTA -> TMEM (thread = row, pairs (k, k+1) -> column k/2)
    const uint16_t* arow = A_gmem + (size_t)(128 * rank + threadIdx.x) * 128;
    for (int c0 = 0; c0 < 64; c0 += 16) {
      uint32_t v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = (uint32_t)arow[2 * (c0 + i)] | ((uint32_t)arow[2 * (c0 + i) + 1] << 16);
      tmem_st_x16(tmem_A + ((uint32_t)(warp * 32) << 16) + c0, v);
    }
    tc_wait_st();
    tc_fence_before();
  }
  cluster_sync_all();   // both CTAs' TMEM operands are written
  tc_fence_after();

  if (rank == 0 && warp == 0) {
    mbar_wait(&bar_load, 0);
    tc_fence_after();
    if (elect_one()) {
      if (!p.pv) {
        const uint32_t idesc = umma_idesc_f16(1, 256, 128, 0, 0);
        for (int k = 0; k < 8; ++k) {
          const uint64_t da = umma_desc_kmajor(smem_u32(sA) + (k / 4) * 16384, k % 4);
          const uint64_t db = umma_desc_kmajor(smem_u32(sB) + (k / 4) * 8192, k % 4);
          umma2_ss(tmem_D, da, db, idesc, k > 0);
        }
      } else {
        const uint32_t idesc = umma_idesc_f16(1, 256, 128, 0, 1);
        for (int k = 0; k < 8; ++k) {
          const uint64_t db = umma_desc_mnmajor(smem_u32(sB), 16384u, k);
          umma2_ts(tmem_D, tmem_A + k * 8, db, idesc, k > 0);
        }
      }
      tc_commit2(&bar_mma, 3);
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
    for (int i = 0; i < 32; ++i) D_out[(size_t)(128 * rank + threadIdx.x) * 128 + c0 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc2<256>(tmem);
}

// ------------------------------------------------------------------------------------------------ part 2: throughput
// kMode 0: 1-CTA SS   1: 1-CTA TS   2: 2-CTA SS   3: 2-CTA TS   (K-major B, the score product)
//       4: 1-CTA TS with MN-major B   5: 2-CTA TS with MN-major B   (the P.V product);   fill: concurrent TMA stream.
// The issue loop is what the attention kernels use: descriptor low words in uniform registers + compile-time offsets
// (umma*_off), no branches — a loop with run-time branches per MMA is issue-bound (~100 clk per instruction) and says
// nothing about the operand path.
template <int kMode, int kN = 128>
__global__ void __launch_bounds__(128, 1)
rate_kernel(const __grid_constant__ CUtensorMap tmFill, int fill, int iters, long long* out) {
  constexpr bool kPair = kMode == 2 || kMode == 3 || kMode == 5;
  constexpr bool kTs = kMode == 1 || kMode == 3 || kMode >= 4;
  constexpr bool kBmn = kMode >= 4;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;             // 32 KiB
  uint8_t* sB = smem + 32768;     // 32 KiB
  uint8_t* sF = smem + 65536;     // fill ring: 4 x 16 KiB
  __shared__ uint64_t bar_mma, bar_fill[4];
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = kPair ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar_mma, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&bar_fill[i], 1);
    fence_mbar_init();
    stop = 0;
  }
  for (int i = threadIdx.x; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // finite operands
  fence_proxy_async_smem();
  if (warp == 0) {
    if (kPair) tmem_alloc2<512>(&tmem_base_s); else tmem_alloc<512>(&tmem_base_s);
  }
  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0 && rank == 0) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(1, kPair ? 256 : 128, kN, 0, kBmn ? 1 : 0);
      constexpr int kBBox = kPair ? 8192 : 16384;
      const uint32_t a_lo = umma_lo_kmajor(smem_u32(sA));
      const uint32_t b_lo = kBmn ? umma_lo_mnmajor(smem_u32(sB), 16384) : umma_lo_kmajor(smem_u32(sB));
      const uint32_t tA = tmem + 256;
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        const uint32_t d = tmem + (it & 1) * 128;
        static_for<0, 8>([&](auto kc) {
          constexpr int k = decltype(kc)::value;
          constexpr uint32_t offA = umma_koff_kmajor(k, 16384);
          constexpr uint32_t offB = kBmn ? umma_koff_mnmajor(k) : umma_koff_kmajor(k, kBBox);
          if constexpr (kPair && kTs) umma2_ts_off<k * 8, offB>(d, tA, b_lo, idesc, k > 0);
          else if constexpr (kPair) umma2_ss_off<offA, offB>(d, a_lo, b_lo, idesc, k > 0);
          else if constexpr (kTs) umma_ts_off<k * 8, offB>(d, tA, b_lo, idesc, k > 0);
          else umma_ss_off<offA, offB>(d, a_lo, b_lo, idesc, k > 0);
        });
      }
      if (kPair) tc_commit2(&bar_mma, 3); else tc_commit(&bar_mma);
      mbar_wait(&bar_mma, 0);
      const long long t1 = clock64();
      stop = 1;
      if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = (long long)iters * 8; }
    }
    __syncwarp();
  } else if (kPair && warp == 0 && rank == 1) {
    if (elect_one()) {   // the leader's commit is multicast to this CTA's barrier as well
      mbar_wait(&bar_mma, 0);
      stop = 1;
    }
    __syncwarp();
  } else if (warp == 1 && fill) {
    if (elect_one()) {   // keep 4 x 16 KiB box loads in flight until the MMA stream is done
      int n = 0;
      uint32_t ph[4] = {0, 0, 0, 0};
      for (int i = 0; i < 4; ++i) {
        mbar_arrive_expect_tx(&bar_fill[i], 16384);
        tma_load_4d(sF + i * 16384, &tmFill, &bar_fill[i], 0, (i * 128 + blockIdx.x * 512) & 8191, 0, 0);
      }
      while (!stop) {
        const int i = n & 3;
        mbar_wait(&bar_fill[i], ph[i]);
        ph[i] ^= 1;
        mbar_arrive_expect_tx(&bar_fill[i], 16384);
        tma_load_4d(sF + i * 16384, &tmFill, &bar_fill[i], 0, ((n + 4) * 128 + blockIdx.x * 512) & 8191, 0, 0);
        ++n;
      }
      for (int i = 0; i < 4; ++i) mbar_wait(&bar_fill[(n + i) & 3], ph[(n + i) & 3]);
      if (blockIdx.x == 0) out[2] = (long long)(n + 4) * 16384;
    }
    __syncwarp();
  }
  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();
  if (warp == 0) {
    if (kPair) tmem_dealloc2<512>(tmem); else tmem_dealloc<512>(tmem);
  }
}

// ------------------------------------------------------------------------------------------------ host
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
#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

static int run_pair(int pv) {
  const int M = 256, N = 128, K = 128;
  std::vector<uint16_t> hA(M * K), hB(N * K);
  std::vector<float> fA(M * K), fB(N * K);   // fB[n][k]
  srand(777 + pv);
  for (int i = 0; i < M * K; ++i) {
    hA[i] = f2bf((float)(rand() % 2001 - 1000) / 500.0f);
    fA[i] = bf2f(hA[i]);
  }
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      const uint16_t h = f2bf((float)(rand() % 2001 - 1000) / 500.0f);
      hB[pv ? k * N + n : n * K + k] = h;   // S: [N x K] K-major; PV: [K x N] MN-major
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
  CUtensorMap tmA, tmB64, tmB128;
  // (B,H,N,D) maps: "N" = rows, "D" = contiguous columns
  int r = make_tmap_bhnd_16bit(&tmA, dA, 1, 1, 1, M, K, (int64_t)M * K, (int64_t)M * K, K, 128);
  r |= make_tmap_bhnd_16bit(&tmB64, dB, 1, 1, 1, N, K, (int64_t)N * K, (int64_t)N * K, K, 64);     // S: box 64 x 64 rows
  r |= make_tmap_bhnd_16bit(&tmB128, dB, 1, 1, 1, K, N, (int64_t)K * N, (int64_t)K * N, N, 128);   // PV: box 64 x 128 rows
  if (r) { printf("tensor map encode failed\n"); return 1; }
  const int smem_bytes = 65536 + 1024;
  CK(cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  P1 p{pv};
  pair_kernel<<<2, 128, smem_bytes>>>(tmA, tmB64, tmB128, dA, dD, p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("pair case %s: kernel failed: %s\n", pv ? "PV" : "S", cudaGetErrorString(e)); exit(3); }
  std::vector<float> hD(M * N), ref(M * N);
  CK(cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = 0;
      for (int k = 0; k < K; ++k) acc += (double)fA[m * K + k] * fB[n * K + k];
      ref[m * N + n] = (float)acc;
      maxerr = std::max(maxerr, std::fabs(acc - hD[m * N + n]));
    }
  printf("pair case %-2s: max |D - ref| = %.3e  %s\n", pv ? "PV" : "S", maxerr, maxerr < 1e-2 ? "OK" : "MISMATCH");
  if (maxerr >= 1e-2) {
    // which reference column does every output column match (using rows that match some reference row)?
    printf("  output column -> best reference column (first 8 of each 64): ");
    for (int n = 0; n < N; ++n) {
      int best = -1; double be = 1e30;
      for (int n2 = 0; n2 < N; ++n2) {
        double err = 0;
        for (int m = 0; m < 16; ++m) err += std::fabs(hD[m * N + n] - ref[m * N + n2]);
        if (err < be) { be = err; best = n2; }
      }
      if ((n & 63) < 8) printf("%d->%d(%.1e) ", n, best, be);
    }
    printf("\n  output row -> best reference row (rows 0,1,64,65,128,129,192,193): ");
    const int rows[8] = {0, 1, 64, 65, 128, 129, 192, 193};
    for (int ri = 0; ri < 8; ++ri) {
      int best = -1; double be = 1e30;
      for (int m2 = 0; m2 < M; ++m2) {
        double err = 0;
        for (int n = 0; n < N; ++n) err += std::fabs(hD[rows[ri] * N + n] - ref[m2 * N + n]);
        if (err < be) { be = err; best = m2; }
      }
      printf("%d->%d(%.1e) ", rows[ri], best, be);
    }
    printf("\n");
  }
  cudaFree(dA), cudaFree(dB), cudaFree(dD);
  return maxerr < 1e-2 ? 0 : 1;
}

template <int kMode, int kN = 128>
static void run_rate(int fill, const char* name, int grid) {
  uint16_t* dF;
  long long* dOut;
  CK(cudaMalloc(&dF, 8192 * 64 * 2));
  CK(cudaMemset(dF, 0, 8192 * 64 * 2));
  CK(cudaMalloc(&dOut, 64));
  CK(cudaMemset(dOut, 0, 64));
  CUtensorMap tmF;
  if (make_tmap_bhnd_16bit(&tmF, dF, 1, 1, 1, 8192, 64, 8192 * 64, 8192 * 64, 64, 128)) { printf("tmap failed\n"); return; }
  const int smem_bytes = 65536 + 65536 + 1024;
  constexpr bool pair = kMode == 2 || kMode == 3 || kMode == 5;
  const int iters = 2000;
  if (pair && grid < 2) grid = 2;
  CK(cudaFuncSetAttribute(rate_kernel<kMode, kN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(pair ? (grid & ~1) : grid), cfg.blockDim = dim3(128), cfg.dynamicSmemBytes = smem_bytes;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = pair ? 2 : 1, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
  cfg.attrs = at, cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, rate_kernel<kMode, kN>, tmF, fill, iters, dOut));
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: kernel failed: %s\n", name, cudaGetErrorString(e)); exit(3); }
  long long h[3];
  CK(cudaMemcpy(h, dOut, 24, cudaMemcpyDeviceToHost));
  const double clk = (double)h[0] / (double)h[1];
  printf("%-30s%s grid %3d : %6.1f clk per MMA instruction (tensor peak %d) -> %3.0f %% of peak", name, fill ? " + TMA fill" : "           ",
         grid, clk, kN / 2, 50.0 * kN / clk);
  if (h[2]) printf("   fill %.1f B/clk", (double)h[2] / (double)h[0]);
  printf("\n");
  cudaFree(dF), cudaFree(dOut);
}

int main() {
  int bad = 0;
  bad += run_pair(0);
  bad += run_pair(1);
  for (int grid : {1, 148})
    for (int fill : {0, 1}) {
      run_rate<0>(fill, "1-CTA SS (A, B smem)", grid);
      run_rate<1>(fill, "1-CTA TS (A tmem)", grid);
      run_rate<2>(fill, "2-CTA SS (B halved)", grid);
      run_rate<3>(fill, "2-CTA TS (B halved)", grid);
      run_rate<4>(fill, "1-CTA TS, MN-major B (P.V)", grid);
      run_rate<5>(fill, "2-CTA TS, MN-major B (P.V)", grid);
    }
  // smaller N per instruction: is the single issuing thread fast enough to keep the pipe full? (the backward kernels issue
  // their score products as N = 64 halves)
  run_rate<0, 64>(0, "1-CTA SS N=64", 148);
  run_rate<1, 64>(0, "1-CTA TS N=64", 148);
  run_rate<0, 32>(0, "1-CTA SS N=32", 148);
  run_rate<1, 32>(0, "1-CTA TS N=32", 148);
  run_rate<2, 64>(0, "2-CTA SS N=64", 148);
  run_rate<3, 64>(0, "2-CTA TS N=64", 148);
  return bad;
}
