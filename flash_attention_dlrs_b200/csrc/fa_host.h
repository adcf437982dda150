// fa_host.h — host-side declarations shared by the translation units of libfa_b200.so.
//
// The kernels are instantiated in several .cu files (forward, dK/dV, dQ; everything else with the C ABI in fa_api.cu) so
// that nvcc compiles them in parallel (_lib.build); fa_api.cu calls the per-kernel dispatchers declared here.
#pragma once

#include <atomic>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

namespace fa {
struct FwdParams;
struct BwdParams;
struct BwdMaps;
}  // namespace fa

namespace fa_host {

// thread-local error text behind fa_last_error() (defined in fa_api.cu)
int fail(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

template <typename K>
int set_smem(K kernel, int bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
  return 0;
}
// The same, once per kernel instantiation and device instead of on every launch: `done` is a bit mask of the devices
// the attribute has been set on, owned by the caller (one static per instantiation — the function-pointer TYPE is shared
// by all instantiations, so it cannot live here).  Thread safe: a lost race only repeats an idempotent call.
template <typename K>
int set_smem_once(K kernel, int bytes, std::atomic<uint64_t>& done) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return set_smem(kernel, bytes);
  const uint64_t bit = 1ull << dev;
  if (done.load(std::memory_order_acquire) & bit) return 0;
  if (int r = set_smem(kernel, bytes)) return r;
  done.fetch_or(bit, std::memory_order_release);
  return 0;
}

// tcgen05 forward (fa_launch_fwd.cu).  elt: 0 f16, 1 bf16, 3 e4m3, 4 e5m2; dropout / mask variants chosen from p.
int launch_fwd16(int elt, int D, bool causal, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                 const fa::FwdParams& p, int H, int B, cudaStream_t st);
// CTA-pair forward (fa_fwd2_sm100.cuh): true when the problem runs on it (16-bit, D = 128, no dropout / mask)
bool fwd_pair_eligible(int elt, int D, const fa::FwdParams& p);
// tk64: K with a 64-row box (each CTA of a pair loads half of every key block)
int launch_fwd16_pair(int elt, int D, bool causal, const CUtensorMap& tq, const CUtensorMap& tk64, const CUtensorMap& tv,
                      const fa::FwdParams& p, int H, int B, cudaStream_t st);
// tcgen05 backward, one kernel each (fa_launch_dkdv.cu, fa_launch_dq.cu)
int launch_bwd16_dkdv(bool bf16, int D, bool causal, const fa::BwdMaps& m, const fa::BwdParams& p, cudaStream_t st);
int launch_bwd16_dq(bool bf16, int D, bool causal, const fa::BwdMaps& m, const fa::BwdParams& p, cudaStream_t st);

}  // namespace fa_host
