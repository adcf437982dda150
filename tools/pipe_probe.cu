// pipe_probe.cu — issue / pipe throughput per SM sub-partition of the instructions the softmax stages are made of, alone and
// mixed, for 1 / 2 / 4 warps per sub-partition (independent dependency chains, 8 per thread).  Reported: clocks per warp
// instruction and sub-partition (1.0 = one instruction issued every clock).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/pipe_probe tools/pipe_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define EX2(r) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(r))
#define FMA2(r, a, b) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r) : "l"(a), "l"(b))
#define ADD2(r, a) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(r) : "l"(a))
#define FMA1(r, a, b) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(r) : "f"(a), "f"(b))
#define CVT(o, x, y) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o) : "f"(x), "f"(y))
#define IMAD(r, a) asm volatile("mad.lo.s32 %0, %1, 8388608, %0;" : "+r"(r) : "r"(a))
#define MAX3(r, a, b) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(r) : "f"(a), "f"(b))

template <int kMode>
__global__ void __launch_bounds__(512, 1) probe(int iters, long long* out, float* sink, float seed) {
  float f[8];
  uint64_t d[8];
  uint32_t u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    f[i] = seed * (float)(threadIdx.x + i);
    d[i] = ((uint64_t)__float_as_uint(f[i]) << 32) | __float_as_uint(f[i] * 0.5f);
    u[i] = threadIdx.x + i;
  }
  const uint64_t ca = ((uint64_t)__float_as_uint(0.999f) << 32) | __float_as_uint(1.001f);
  const uint64_t cb = ((uint64_t)__float_as_uint(1e-3f) << 32) | __float_as_uint(-1e-3f);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if constexpr (kMode == 0) EX2(f[i]);
        if constexpr (kMode == 1) FMA2(d[i], ca, cb);
        if constexpr (kMode == 2) ADD2(d[i], cb);
        if constexpr (kMode == 3) FMA1(f[i], 0.999f, 1e-3f);
        if constexpr (kMode == 4) { CVT(u[i], f[i], f[(i + 1) & 7]); }
        if constexpr (kMode == 5) IMAD(u[i], u[(i + 1) & 7]);
        if constexpr (kMode == 6) MAX3(f[i], f[(i + 1) & 7], f[(i + 2) & 7]);
        if constexpr (kMode == 10) { EX2(f[i]); FMA2(d[i], ca, cb); }                                        // 1 MUFU : 1 FFMA2
        if constexpr (kMode == 11) { EX2(f[i]); FMA2(d[i], ca, cb); FMA2(d[(i + 4) & 7], ca, cb); ADD2(d[(i + 2) & 7], cb); ADD2(d[(i + 6) & 7], cb); }   // 1 : 4 packed
        if constexpr (kMode == 12) { EX2(f[i]); FMA1(f[(i + 4) & 7], 0.999f, 1e-3f); FMA1(f[(i + 5) & 7], 0.999f, 1e-3f); FMA1(f[(i + 6) & 7], 0.999f, 1e-3f); FMA1(f[(i + 7) & 7], 0.999f, 1e-3f); }   // 1 : 4 scalar FFMA
        if constexpr (kMode == 13) { EX2(f[i]); CVT(u[i], f[(i + 3) & 7], f[(i + 4) & 7]); }                 // 1 MUFU : 1 F2FP
        if constexpr (kMode == 14) { EX2(f[i]); IMAD(u[i], u[(i + 1) & 7]); IMAD(u[(i + 2) & 7], u[(i + 3) & 7]); }   // 1 MUFU : 2 IMAD
        if constexpr (kMode == 15) { FMA2(d[i], ca, cb); IMAD(u[i], u[(i + 1) & 7]); }                        // FFMA2 + IMAD (same pipe?)
        if constexpr (kMode == 16) { FMA2(d[i], ca, cb); CVT(u[i], f[(i + 3) & 7], f[(i + 4) & 7]); }         // FFMA2 + F2FP
        if constexpr (kMode == 17) { FMA2(d[i], ca, cb); MAX3(f[i], f[(i + 1) & 7], f[(i + 2) & 7]); }        // FFMA2 + FMNMX3
        if constexpr (kMode == 18) { FMA1(f[i], 0.999f, 1e-3f); IMAD(u[i], u[(i + 1) & 7]); }                 // FFMA + IMAD
      }
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += f[i] + __uint_as_float((uint32_t)d[i]) + __uint_as_float((uint32_t)(d[i] >> 32)) + (float)u[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int kMode>
static void run(const char* name, int per_iter) {
  long long* out;
  float* sink;
  cudaMalloc(&out, 8);
  cudaMalloc(&sink, 148 * 512 * 4);
  const int iters = 4000;
  printf("%-44s", name);
  for (int warps : {4, 8, 16}) {
    probe<kMode><<<148, warps * 32>>>(iters, out, sink, 1e-3f);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    long long c;
    cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
    // instructions per sub-partition = iters * 32 * per_iter * (warps / 4)
    printf("  %d/SMSP: %6.2f clk/instr", warps / 4, (double)c / ((double)iters * 32 * per_iter * (warps / 4)));
  }
  printf("\n");
  cudaFree(out), cudaFree(sink);
}

int main() {
  run<0>("MUFU.EX2", 1);
  run<1>("FFMA2", 1);
  run<2>("FADD2", 1);
  run<3>("FFMA", 1);
  run<4>("F2FP.BF16.PACK_AB", 1);
  run<5>("IMAD", 1);
  run<6>("FMNMX3", 1);
  run<10>("1 MUFU + 1 FFMA2", 2);
  run<11>("1 MUFU + 2 FFMA2 + 2 FADD2", 5);
  run<12>("1 MUFU + 4 FFMA", 5);
  run<13>("1 MUFU + 1 F2FP", 2);
  run<14>("1 MUFU + 2 IMAD", 3);
  run<15>("1 FFMA2 + 1 IMAD", 2);
  run<16>("1 FFMA2 + 1 F2FP", 2);
  run<17>("1 FFMA2 + 1 FMNMX3", 2);
  run<18>("1 FFMA + 1 IMAD", 2);
  return 0;
}
