// fa_launch_fwd.cu — instantiations and launch dispatch of the tcgen05 forward kernel (own translation unit: see fa_host.h).
#include "../../include/fa_b200.h"
#include "fa_fwd_sm100.cuh"
#include "fa_host.h"

namespace {

template <int kElt, int kD, bool kCausal, bool kDrop = false, bool kAmask = false>
int launch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const fa::FwdParams& p, int H, int B,
           cudaStream_t st) {
  using Cfg = fa::FwdCfg<kD, kElt>;
  if constexpr (kElt < 3 && !kDrop && !kAmask) {
    const bool masked = p.amask != nullptr || p.band != 0;
    if (p.drop.thresh && masked) return launch<kElt, kD, kCausal, true, true>(tq, tk, tv, p, H, B, st);
    if (p.drop.thresh) return launch<kElt, kD, kCausal, true, false>(tq, tk, tv, p, H, B, st);
    if (masked) return launch<kElt, kD, kCausal, false, true>(tq, tk, tv, p, H, B, st);
  }
  auto kern = fa::fa_fwd_kernel<kElt, kD, kCausal, kDrop, kAmask>;
  static std::atomic<uint64_t> smem_set{0};   // per instantiation: devices whose attribute is set
  if (int r = fa_host::set_smem_once(kern, Cfg::kSmemBytes, smem_set)) return r;
  dim3 grid(p.q_blocks, H, B);
  kern<<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(tq, tk, tv, p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : fa_host::cuda_fail(e, "fa_fwd launch");
}

}  // namespace

int fa_host::launch_fwd16(int elt, int D, bool causal, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                          const fa::FwdParams& p, int H, int B, cudaStream_t st) {
#define FA_FWD_CASE(E, DD, C) \
  if (elt == E && D == DD && causal == C) return launch<E, DD, C>(tq, tk, tv, p, H, B, st);
  FA_FWD_CASE(FA_DTYPE_BF16, 128, true)
  FA_FWD_CASE(FA_DTYPE_BF16, 128, false)
  FA_FWD_CASE(FA_DTYPE_BF16, 64, true)
  FA_FWD_CASE(FA_DTYPE_BF16, 64, false)
  FA_FWD_CASE(FA_DTYPE_F16, 128, true)
  FA_FWD_CASE(FA_DTYPE_F16, 128, false)
  FA_FWD_CASE(FA_DTYPE_F16, 64, true)
  FA_FWD_CASE(FA_DTYPE_F16, 64, false)
  FA_FWD_CASE(FA_DTYPE_F8E4M3, 128, true)
  FA_FWD_CASE(FA_DTYPE_F8E4M3, 128, false)
  FA_FWD_CASE(FA_DTYPE_F8E5M2, 128, true)
  FA_FWD_CASE(FA_DTYPE_F8E5M2, 128, false)
#undef FA_FWD_CASE
  return fa_host::fail(-3, "fa_fwd: no kernel for dtype %d D %d", elt, D);
}
