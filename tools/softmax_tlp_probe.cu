// softmax_tlp_probe.cu — how the exp phase of the forward softmax scales with warps per SM sub-partition.
// Every warp owns kCols score columns of 32 rows in TMEM and repeats: load -> row max -> p = exp2(s * sl2 - m) (a share on
// the FMA pipe) -> row sum -> packed bf16 -> TMEM (the loop of fa_fwd_sm100.cuh).  Reported: clocks per round and the
// clocks the SM needs per 128 x 128 score tile at that rate (= 16384 elements).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/softmax_tlp_probe tools/softmax_tlp_probe.cu
#include "../flash_attention_dlrs_b200/csrc/sm100_ptx.cuh"

#include <cstdio>
#include <cstdlib>

using namespace fa;

#ifndef POLY_MASK
#define POLY_MASK 0x92
#endif

template <int kCols, int kMask>
__device__ __forceinline__ float exp_phase(uint32_t (&sr)[kCols], uint32_t tS, float sl2, float neg_ms) {
  const uint64_t sl2_2 = f32x2_pack(sl2, sl2), nm2 = f32x2_pack(neg_ms, neg_ms);
  uint64_t ls[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
  for (int c = 0; c < kCols / 32; ++c) {
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int e = c * 32 + 2 * i;
      const uint64_t x2 = f32x2_fma(f32x2_pack_bits(sr[e], sr[e + 1]), sl2_2, nm2);
      float x0, x1, p0, p1;
      f32x2_unpack(x2, x0, x1);
      if ((kMask >> (i & 7)) & 1) ex2_poly_x2(x0, x1, p0, p1);
      else p0 = ex2_approx(x0), p1 = ex2_approx(x1);
      ls[i & 3] = f32x2_add(ls[i & 3], f32x2_pack(p0, p1));
      pk[i] = pack2<true>(p0, p1);
    }
    tmem_st_x16(tS + c * 16, pk);
  }
  float la, lb, lc, ld;
  f32x2_unpack(f32x2_add(ls[0], ls[1]), la, lb);
  f32x2_unpack(f32x2_add(ls[2], ls[3]), lc, ld);
  return (la + lb) + (lc + ld);
}

// kBg: a ninth / seventeenth warp keeps the tensor pipe saturated meanwhile (1: both operands in shared memory, like the
// score products; 2: A from TMEM, like P.V) with accumulators in TMEM columns 256-383 — does the softmax slow down?
template <int kCols, int kWarps, int kMask, bool kExp, int kBg, int kWork = 0, int kBgN = 128, int kSame = 0>
__global__ void __launch_bounds__(kWarps * 32 + (kBg ? 32 : 0), 1) probe(int iters, float sl2, long long* out, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint32_t tmem_base_s;
  __shared__ uint64_t bar_mma[2];
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar_mma[0], 1);
    mbar_init(&bar_mma[1], 1);
    fence_mbar_init();
    stop = 0;
  }
  if constexpr (kBg != 0) {
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // finite operands
    fence_proxy_async_smem();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (kBg != 0 && warp == kWarps) {
    if (elect_one()) {
      uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
      constexpr uint32_t idesc = umma_idesc_f16(1, 128, kBgN, 0, kBg == 2 ? 1 : 0);
      const uint32_t a_lo = umma_lo_kmajor(smem_u32(smem));
      const uint32_t b_lo = kBg == 2 ? umma_lo_mnmajor(smem_u32(smem + 32768), 16384) : umma_lo_kmajor(smem_u32(smem + 32768));
      const uint32_t d = tmem + 256, tA = tmem + 384;
      long long n = 0;
      const long long t0 = clock64();
      for (int g = 0; !stop; ++g) {
        // at most two groups of 8 in flight (kWork / 10 odd: no wait at all, the issue queue's back-pressure paces the thread)
        if (((kWork / 10) & 1) == 0 && g >= 2) mbar_wait(&bar_mma[g & 1], ((g - 2) >> 1) & 1);
        static_for<0, 8>([&](auto kc) {
          constexpr int k = decltype(kc)::value;
          constexpr int kk = kSame ? 0 : k;   // kSame: the same operand addresses for every instruction (no descriptor arithmetic)
          if constexpr (kBg == 2) umma_ts_off<kk * 8, umma_koff_mnmajor(kk)>(d, tA, b_lo, idesc, k > 0);
          else umma_ss_off<umma_koff_kmajor(kk, 16384), umma_koff_kmajor(kk, 16384)>(d, a_lo, b_lo, idesc, k > 0);
        });
        if (((kWork / 10) & 1) == 0) tc_commit(&bar_mma[g & 1]);
        n += 8;
      }
      const long long t1 = clock64();
      if (blockIdx.x == 0) { out[1] = t1 - t0; out[2] = n; }
    }
    __syncwarp();
    __syncthreads();
    return;
  }
  const uint32_t tS = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * kCols;
  {
    uint32_t z[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) z[i] = __float_as_uint(-0.01f * (float)((threadIdx.x * 7 + i * 13) & 255));
    for (int c = 0; c < kCols / 32; ++c) tmem_st_x32(tS + c * 32, z);
    tc_wait_st();
  }
  named_bar_sync(1, kWarps * 32);   // (not __syncthreads: the background warp is already in its loop)
  float l = 0.f;
  const long long t0 = clock64();
  if (kWork % 10 == 3 || (kWork >= 20 && (warp & 3) == 0)) {   // control: the softmax warps sleep (kWork >= 20: those of sub-partition 0 only)
    for (int it = 0; it < iters; ++it) __nanosleep(500);
  } else if constexpr (kWork % 10 == 2) {   // arithmetic only: no TMEM access inside the loop
    uint32_t sr[kCols];
#pragma unroll
    for (int c = 0; c < kCols / 32; ++c) tmem_ld_x32(tS + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&sr[c * 32]));
    tc_wait_ld();
    for (int it = 0; it < iters; ++it) {
      const uint64_t sl2_2 = f32x2_pack(sl2, sl2), nm2 = f32x2_pack(l * 1e-9f, l * 1e-9f);
      uint64_t ls[4] = {0ull, 0ull, 0ull, 0ull};
      uint32_t acc = 0;
#pragma unroll
      for (int i = 0; i < kCols / 2; ++i) {
        const uint64_t x2 = f32x2_fma(f32x2_pack_bits(sr[2 * i], sr[2 * i + 1]), sl2_2, nm2);
        float x0, x1, p0, p1;
        f32x2_unpack(x2, x0, x1);
        if ((kMask >> (i & 7)) & 1) ex2_poly_x2(x0, x1, p0, p1);
        else p0 = ex2_approx(x0), p1 = ex2_approx(x1);
        ls[i & 3] = f32x2_add(ls[i & 3], f32x2_pack(p0, p1));
        acc ^= pack2<true>(p0, p1);
      }
      float la, lb, lc, ld;
      f32x2_unpack(f32x2_add(ls[0], ls[1]), la, lb);
      f32x2_unpack(f32x2_add(ls[2], ls[3]), lc, ld);
      l += (la + lb) + (lc + ld) + __uint_as_float(acc & 0x3f803f80u);
    }
  } else
  for (int it = 0; it < iters; ++it) {
    uint32_t sr[kCols];
#pragma unroll
    for (int c = 0; c < kCols / 32; ++c) tmem_ld_x32(tS + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&sr[c * 32]));
    tc_wait_ld();
    float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
    for (int c = 0; c < kCols; c += 4) {
      mx0 = fmaxf(mx0, __uint_as_float(sr[c]));
      mx1 = fmaxf(mx1, __uint_as_float(sr[c + 1]));
      mx2 = fmaxf(mx2, __uint_as_float(sr[c + 2]));
      mx3 = fmaxf(mx3, __uint_as_float(sr[c + 3]));
    }
    const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
    if constexpr (!kExp || kWork % 10 == 1) l += mx;
    else l += exp_phase<kCols, kMask>(sr, tS, sl2, -mx * sl2);
    tc_wait_st();
#pragma unroll
    for (int c = 0; c < kCols / 32; ++c) tmem_st_x32(tS + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&sr[c * 32]));
    tc_wait_st();
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = l;
  named_bar_sync(1, kWarps * 32);
  if (threadIdx.x == 0) stop = 1;
  __syncthreads();
  if (warp == 0) {
    // (the background stream's last groups may still be running: let them drain before the columns go away)
    if (kBg != 0) __nanosleep(20000);
    tmem_dealloc<512>(tmem);
  }
}

template <int kCols, int kWarps, int kMask, int kBg = 0, int kWork = 0, int kBgN = 128, int kSame = 0>
static void run() {
  long long* out;
  float* sink;
  cudaMalloc(&out, 24);
  cudaMemset(out, 0, 24);
  cudaMalloc(&sink, 148 * 544 * 4);
  const int iters = 2000;
  double clk[2];
  long long bg[3] = {0, 0, 0};
  const int threads = kWarps * 32 + (kBg ? 32 : 0), smem = kBg ? 65536 + 1024 : 0;
  for (int e = 0; e < 2; ++e) {
    if (e) {
      auto k = probe<kCols, kWarps, kMask, true, kBg, kWork, kBgN, kSame>;
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      k<<<148, threads, smem>>>(iters, 0.1275f, out, sink);
    } else {
      auto k = probe<kCols, kWarps, kMask, false, kBg, kWork, kBgN, kSame>;
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      k<<<148, threads, smem>>>(iters, 0.1275f, out, sink);
    }
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("error %s\n", cudaGetErrorString(err)); exit(1); }
    cudaMemcpy(bg, out, 24, cudaMemcpyDeviceToHost);
    clk[e] = (double)bg[0] / iters;
  }
  printf("bg %d work %d (0 all, 1 TMEM traffic only, 2 arithmetic only, 3 asleep): ", kBg, kWork);
  if (kBg) printf("[MMA stream beside it, N = %d, %s operands%s: %.1f clk per MMA] ", kBgN, kBg == 2 ? "TMEM A" : "shared-memory", kSame ? ", same addresses" : "", (double)bg[1] / (double)bg[2]);
  const double elems = (double)kWarps * 32 * kCols;   // per round and SM
  printf("cols/warp %3d  warps %2d (%d per sub-partition)  poly mask 0x%02x : round %7.1f clk, exp phase %7.1f clk -> %6.0f clk per 128x128 tile "
         "(exp phase only %6.0f; MUFU floor %4.0f)\n", kCols, kWarps, kWarps / 4, kMask, clk[1], clk[1] - clk[0], clk[1] * 16384.0 / elems,
         (clk[1] - clk[0]) * 16384.0 / elems, 16384.0 / 16.0 * (8 - __builtin_popcount(kMask)) / 8.0);
  cudaFree(out), cudaFree(sink);
}

int main(int argc, char** argv) {
  setvbuf(stdout, nullptr, _IONBF, 0);
  if (argc > 1) {   // only the runs with a busy tensor pipe
    run<128, 4, POLY_MASK, 1>();
    run<128, 8, POLY_MASK, 1>();
    run<128, 8, POLY_MASK, 2>();
    run<64, 8, POLY_MASK, 1>();
    run<64, 8, POLY_MASK, 2>();
    run<128, 8, POLY_MASK, 1, 3>();
    run<128, 8, POLY_MASK, 2, 3>();
    run<128, 8, POLY_MASK, 1, 1>();
    run<128, 8, POLY_MASK, 2, 1>();
    run<128, 8, POLY_MASK, 1, 2>();
    run<128, 8, POLY_MASK, 2, 2>();
    run<128, 8, 0x00, 1, 2>();
    run<128, 8, 0xff, 1, 2>();
    run<128, 4, POLY_MASK, 1, 2>();
    printf("-- no wait inside the issue loop\n");
    run<128, 8, POLY_MASK, 1, 13>();
    run<128, 8, POLY_MASK, 2, 13>();
    run<128, 8, POLY_MASK, 1, 12>();
    run<128, 8, POLY_MASK, 2, 12>();
    run<128, 8, POLY_MASK, 1, 10>();
    run<128, 8, POLY_MASK, 2, 10>();
    printf("-- the issuing warp's sub-partition free of softmax work (no wait inside the issue loop)\n");
    run<128, 8, POLY_MASK, 1, 32>();
    run<128, 8, POLY_MASK, 2, 32>();
    run<128, 8, POLY_MASK, 2, 30>();
    printf("-- N = 64 instructions (tensor time 32 clk; shared-memory operands 48): arithmetic beside it (12), asleep (13), issuing sub-partition free (32)\n");
    run<128, 8, POLY_MASK, 1, 13, 64>();
    run<128, 8, POLY_MASK, 2, 13, 64>();
    run<128, 8, POLY_MASK, 1, 12, 64>();
    run<128, 8, POLY_MASK, 2, 12, 64>();
    run<128, 8, POLY_MASK, 1, 32, 64>();
    run<128, 8, POLY_MASK, 2, 32, 64>();
    run<128, 8, POLY_MASK, 1, 12, 64, 1>();
    run<128, 8, POLY_MASK, 2, 12, 64, 1>();
    run<128, 8, POLY_MASK, 2, 13, 64, 1>();
    run<128, 4, POLY_MASK, 2, 12, 64>();
    run<128, 8, POLY_MASK, 2, 10, 64>();
    run<128, 8, POLY_MASK, 2, 11, 64>();
    printf("-- N = 32\n");
    run<128, 8, POLY_MASK, 2, 13, 32>();
    run<128, 8, POLY_MASK, 2, 12, 32>();
    return 0;
  }
  run<128, 4, POLY_MASK>();
  run<128, 8, POLY_MASK>();
  run<64, 4, POLY_MASK>();
  run<64, 8, POLY_MASK>();
  run<64, 12, POLY_MASK>();
  run<64, 16, POLY_MASK>();
  run<32, 16, POLY_MASK>();
  run<64, 16, 0x00>();
  run<64, 16, 0x88>();
  run<64, 16, 0xaa>();
  run<64, 16, 0xda>();
  run<64, 8, 0x00>();
  run<64, 8, 0xaa>();
  run<128, 8, 0x00>();
  run<128, 8, 0xaa>();
  return 0;
}
