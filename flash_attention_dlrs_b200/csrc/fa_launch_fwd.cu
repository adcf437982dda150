// fa_launch_fwd.cu — instantiations and launch dispatch of the tcgen05 forward kernel (own translation unit: see fa_host.h).
#include "../../include/fa_b200.h"
// FA_EXPERIMENTAL_FWD=1 also compiles three measured-and-rejected forward variants, selectable at run time with
// FA_FWD_PAIR=1 / FA_FWD_W16=1 / FA_FWD_DUO=1: the CTA-pair kernel (bit-identical, 35 % slower:
// longer hand-over chains, and shared memory was never the limit), the sixteen-softmax-warp kernel (10 % slower) and the
// kernel whose two softmax warps per sub-partition share one tile (6 % slower).
#ifndef FA_EXPERIMENTAL_FWD
#define FA_EXPERIMENTAL_FWD 0
#endif
#if FA_EXPERIMENTAL_FWD
#include "fa_fwd2_sm100.cuh"
#include "fa_fwd_w16_sm100.cuh"
#include "fa_fwd_duo_sm100.cuh"
#endif
#include "fa_fwd_sm100.cuh"

#include <cstdlib>
#include "fa_host.h"

#ifndef FA_FWD_W16_DEFAULT
#define FA_FWD_W16_DEFAULT 0
#endif

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
#if FA_EXPERIMENTAL_FWD
  if constexpr (kElt < 3 && !kDrop && !kAmask) {
    static const int duo = [] {
      const char* e = std::getenv("FA_FWD_DUO");
      return e ? std::atoi(e) : 0;
    }();
    if (duo) {
      using CfgD = fa::FwdDuoCfg<kD>;
      auto kernd = fa::fa_fwd_duo_kernel<kElt == 1, kD, kCausal>;
      static std::atomic<uint64_t> smem_setd{0};
      if (int r = fa_host::set_smem_once(kernd, CfgD::kSmemBytes, smem_setd)) return r;
      dim3 gridd(p.q_blocks, H, B);
      kernd<<<gridd, CfgD::kThreads, CfgD::kSmemBytes, st>>>(tq, tk, tv, p);
      cudaError_t ed = cudaGetLastError();
      return ed == cudaSuccess ? 0 : fa_host::cuda_fail(ed, "fa_fwd (duo) launch");
    }
  }
  if constexpr (kElt < 3 && !kDrop && !kAmask) {
    // experiment (slower, see above): sixteen softmax warps (fa_fwd_w16_sm100.cuh); FA_FWD_W16=0 keeps the
    // eight-warp kernel (A/B measurements, and what the feature variants still run on)
    static const int w16 = [] {
      const char* e = std::getenv("FA_FWD_W16");
      return e ? std::atoi(e) : FA_FWD_W16_DEFAULT;
    }();
    if (w16) {
      using Cfg16 = fa::FwdW16Cfg<kD>;
      auto kern16 = fa::fa_fwd_w16_kernel<kElt == 1, kD, kCausal>;
      static std::atomic<uint64_t> smem_set16{0};
      if (int r = fa_host::set_smem_once(kern16, Cfg16::kSmemBytes, smem_set16)) return r;
      dim3 grid16(p.q_blocks, H, B);
      kern16<<<grid16, Cfg16::kThreads, Cfg16::kSmemBytes, st>>>(tq, tk, tv, p);
      cudaError_t e16 = cudaGetLastError();
      return e16 == cudaSuccess ? 0 : fa_host::cuda_fail(e16, "fa_fwd (16 softmax warps) launch");
    }
  }
#endif
  auto kern = fa::fa_fwd_kernel<kElt, kD, kCausal, kDrop, kAmask>;
  static std::atomic<uint64_t> smem_set{0};   // per instantiation: devices whose attribute is set
  if (int r = fa_host::set_smem_once(kern, Cfg::kSmemBytes, smem_set)) return r;
  dim3 grid(p.q_blocks, H, B);
  kern<<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(tq, tk, tv, p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : fa_host::cuda_fail(e, "fa_fwd launch");
}

#if FA_EXPERIMENTAL_FWD
template <bool kBf16, int kD, bool kCausal>
int launch_pair(const CUtensorMap& tq, const CUtensorMap& tk64, const CUtensorMap& tv, fa::FwdParams p, int H, int B,
                cudaStream_t st) {
  using Cfg = fa::Fwd2Cfg<kD>;
  auto kern = fa::fa_fwd2_kernel<kBf16, kD, kCausal>;
  static std::atomic<uint64_t> smem_set{0};
  if (int r = fa_host::set_smem_once(kern, Cfg::kSmemBytes, smem_set)) return r;
  p.q_blocks = (p.N + 511) / 512;   // 512-row quads: one per CTA pair (the cluster dimension is compiled in)
  dim3 grid(2 * p.q_blocks, H, B);
  kern<<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(tq, tk64, tv, p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : fa_host::cuda_fail(e, "fa_fwd (CTA pairs) launch");
}
#endif

}  // namespace

#if FA_TRACE
// debug builds only: every translation unit has its own copy of the trace pointer (sm100_ptx.cuh)
extern "C" int fa_debug_set_trace_fwd(void* dev_buf, int capacity_events) {
  long long* p = static_cast<long long*>(dev_buf);
  cudaMemcpyToSymbol(fa::g_fa_trace, &p, sizeof(p));
  cudaMemcpyToSymbol(fa::g_fa_trace_cap, &capacity_events, sizeof(int));
  return 0;
}
#endif

// FA_FWD_PAIR = 1 selects the CTA-pair forward in FA_EXPERIMENTAL_FWD builds (A/B measurements); never in the product build.
bool fa_host::fwd_pair_eligible(int elt, int D, const fa::FwdParams& p) {
#if FA_EXPERIMENTAL_FWD
  static const int mode = [] {
    const char* e = std::getenv("FA_FWD_PAIR");
    return e ? std::atoi(e) : 0;
  }();
  return mode != 0 && (elt == FA_DTYPE_F16 || elt == FA_DTYPE_BF16) && D == 128 && !p.drop.thresh && !p.amask && !p.band;
#else
  (void)elt, (void)D, (void)p;
  return false;
#endif
}

int fa_host::launch_fwd16_pair(int elt, int D, bool causal, const CUtensorMap& tq, const CUtensorMap& tk64,
                               const CUtensorMap& tv, const fa::FwdParams& p, int H, int B, cudaStream_t st) {
#if FA_EXPERIMENTAL_FWD
  const bool bf = elt == FA_DTYPE_BF16;
  if (D == 128) {
    if (bf && causal) return launch_pair<true, 128, true>(tq, tk64, tv, p, H, B, st);
    if (bf && !causal) return launch_pair<true, 128, false>(tq, tk64, tv, p, H, B, st);
    if (!bf && causal) return launch_pair<false, 128, true>(tq, tk64, tv, p, H, B, st);
    return launch_pair<false, 128, false>(tq, tk64, tv, p, H, B, st);
  }
#endif
  (void)elt, (void)causal, (void)tq, (void)tk64, (void)tv, (void)p, (void)H, (void)B, (void)st;
  return fa_host::fail(-3, "fa_fwd: no CTA-pair kernel for D %d in this build", D);
}

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
