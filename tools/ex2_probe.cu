// ex2_probe.cu — MUFU throughput of ex2.approx: f32 (one value per lane-op) against f16x2 / bf16x2 (two values per op?).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/ex2_probe tools/ex2_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int kMode>
__global__ void __launch_bounds__(1024) probe(int iters, long long* out, uint32_t* sink) {
  uint32_t v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = (kMode == 0) ? __float_as_uint(-0.5f - 0.01f * (threadIdx.x + i)) : (kMode == 1 ? 0xB800B900u : 0xBF00BF20u) + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (kMode == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(v[i]));
      if (kMode == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(v[i]));
      if (kMode == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(v[i]));
      if (kMode == 3) asm volatile("{.reg .f16 h; mov.b32 {h, _}, %0; ex2.approx.f16 h, h; mov.b32 %0, {h, h};}" : "+r"(v[i]));
    }
  }
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s ^= v[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}
template <int kMode>
void run(const char* name, int warps, int per_op) {
  long long* out; uint32_t* sink;
  cudaMalloc(&out, 8); cudaMalloc(&sink, 148 * 1024 * 4);
  const int iters = 4000;
  probe<kMode><<<148, warps * 32>>>(iters, out, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
  printf("%-26s warps %2d : %.2f instr/clk/SM -> %.1f exponentials/clk/SM\n", name, warps, 16.0 * iters * warps * 32 / c / 32.0 * 32 / 32 * 1.0, 16.0 * iters * warps * 32 * per_op / c);
  cudaFree(out); cudaFree(sink);
}
int main() {
  for (int w : {8, 16, 32}) {
    run<0>("ex2.approx.ftz.f32", w, 1);
    run<1>("ex2.approx.ftz.f16x2", w, 2);
    run<2>("ex2.approx.ftz.bf16x2", w, 2);
    run<3>("ex2.approx.ftz.f16", w, 1);
  }
  return 0;
}
