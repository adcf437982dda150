// fa_launch_dkdv.cu — instantiations and launch dispatch of the tcgen05 dK/dV kernel (own translation unit: see fa_host.h).
#include "../../include/fa_b200.h"
#include "fa_bwd_sm100.cuh"
#include "fa_host.h"

#if FA_TRACE
// debug builds only: every translation unit has its own copy of the trace pointer (sm100_ptx.cuh)
extern "C" int fa_debug_set_trace_dkdv(void* dev_buf, int capacity_events) {
  long long* p = static_cast<long long*>(dev_buf);
  cudaMemcpyToSymbol(fa::g_fa_trace, &p, sizeof(p));
  cudaMemcpyToSymbol(fa::g_fa_trace_cap, &capacity_events, sizeof(int));
  return 0;
}
#endif

namespace {

template <bool kBf16, int kD, bool kCausal, bool kDrop = false, bool kAmask = false>
int launch(const fa::BwdMaps& m, const fa::BwdParams& p, cudaStream_t st) {
  if constexpr (!kDrop && !kAmask) {
    const bool masked = p.amask != nullptr || p.band != 0;
    if (p.drop.thresh && masked) return launch<kBf16, kD, kCausal, true, true>(m, p, st);
    if (p.drop.thresh) return launch<kBf16, kD, kCausal, true, false>(m, p, st);
    if (masked) return launch<kBf16, kD, kCausal, false, true>(m, p, st);
  }
  auto kern = fa::fa_bwd_dkdv_kernel<kBf16, kD, kCausal, kDrop, kAmask>;
  static std::atomic<uint64_t> smem_set{0};   // per instantiation: devices whose attribute is set
  if (int r = fa_host::set_smem_once(kern, fa::BwdCfg<kD>::kSmemDkdv, smem_set)) return r;
  dim3 grid(((p.Nk > 0 ? p.Nk : p.N) + 127) / 128, p.H, p.B);   // one CTA per key block
  kern<<<grid, fa::BwdCfg<kD>::kThreads, fa::BwdCfg<kD>::kSmemDkdv, st>>>(m.q, m.k, m.v, m.dout, p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : fa_host::cuda_fail(e, "fa_bwd(dK/dV) launch");
}

}  // namespace

int fa_host::launch_bwd16_dkdv(bool bf16, int D, bool causal, const fa::BwdMaps& m, const fa::BwdParams& p, cudaStream_t st) {
#define FA_BWD_CASE(BF, DD, C) \
  if (bf16 == BF && D == DD && causal == C) return launch<BF, DD, C>(m, p, st);
  FA_BWD_CASE(true, 128, true)
  FA_BWD_CASE(true, 128, false)
  FA_BWD_CASE(true, 64, true)
  FA_BWD_CASE(true, 64, false)
  FA_BWD_CASE(false, 128, true)
  FA_BWD_CASE(false, 128, false)
  FA_BWD_CASE(false, 64, true)
  FA_BWD_CASE(false, 64, false)
#undef FA_BWD_CASE
  return fa_host::fail(-3, "fa_bwd: no kernel for D %d", D);
}
