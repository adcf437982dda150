// l2_reduce_probe.cu — throughput of fp32 reductions into L2-resident global memory on B200:
//   mode 0: TMA bulk reduce-add (cp.reduce.async.bulk .add.f32) of 32 KiB from shared memory
//   mode 1: red.global.add.v4.f32 issued by 128 threads (32 KiB per round)
//   mode 2: TMA bulk store (no reduction) of 32 KiB
//   mode 3: ld.global.cg.v4 + add + st.global.v4 (exclusive read-modify-write), 128 threads
// Each CTA cycles over `tiles` private 32 KiB tiles (so everything stays in L2); grid = number of SMs under test.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/l2_reduce_probe tools/l2_reduce_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(128, 1) probe(int mode, int rounds, int tiles, float* g, long long* out) {
  extern __shared__ __align__(1024) float s[];
  for (int i = threadIdx.x; i < 8192; i += 128) s[i] = 1.0f;
  asm volatile("fence.proxy.async;" ::: "memory");
  __syncthreads();
  float* base = g + (size_t)blockIdx.x * tiles * 8192;
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(s);
  long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    float* tile = base + (size_t)(r % tiles) * 8192;
    if (mode == 0 || mode == 2) {
      if (threadIdx.x == 0) {
        if (mode == 0)
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(tile), "r"(sa), "r"(32768) : "memory");
        else
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(tile), "r"(sa), "r"(32768) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      }
    } else if (mode == 1) {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        float* p = tile + (k * 128 + threadIdx.x) * 4;
        asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(p), "f"(1.0f) : "memory");
      }
    } else {
      float4 v[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = __ldcg(reinterpret_cast<const float4*>(tile) + k * 128 + threadIdx.x);
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        v[k].x += 1.f, v[k].y += 1.f, v[k].z += 1.f, v[k].w += 1.f;
        reinterpret_cast<float4*>(tile)[k * 128 + threadIdx.x] = v[k];
      }
    }
  }
  if ((mode == 0 || mode == 2) && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __threadfence();
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

int main() {
  const int tiles = 8, rounds = 400;
  float* g; long long* out;
  cudaMalloc(&g, (size_t)148 * tiles * 32768);
  cudaMemset(g, 0, (size_t)148 * tiles * 32768);
  cudaMalloc(&out, 148 * 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char* names[4] = {"bulk reduce.add.f32", "red.v4.f32 (threads)", "bulk store", "ld.cg + add + st (threads)"};
  for (int mode = 0; mode < 4; ++mode)
    for (int grid : {1, 8, 37, 74, 148}) {
      probe<<<grid, 128, 200 * 1024>>>(mode, rounds, tiles, g, out);   // 200 KiB of smem: one CTA per SM
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[148]; cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
      double bpc = 32768.0 * rounds / mx;
      printf("%-28s grid %3d : %.1f B/clk/SM  -> %.2f TB/s chip-wide at 1.965 GHz\n", names[mode], grid, bpc, bpc * grid * 1.965e9 / 1e12);
    }
  return 0;
}
