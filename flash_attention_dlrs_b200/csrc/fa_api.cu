// fa_api.cu — C ABI of libfa_b200.so (declared in include/fa_b200.h): argument checking, TMA tensor-map
// encoding and kernel launches.  No torch types, no device allocation, no synchronisation.
// The tcgen05 forward / dK,dV / dQ kernels are instantiated in fa_launch_*.cu (parallel compilation, fa_host.h).
#include "../../include/fa_b200.h"

#include "fa_bwd_fused_sm100.cuh"
#include "fa_bwd_sm100.cuh"
#include "fa_fwd_sm100.cuh"
#include "fa_host.h"
#include "fa_merge.cuh"
#include "fa_preprocess.cuh"
#include "fa_simt_f32.cuh"
#include "tmap.h"

#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>

namespace {
thread_local char g_err[512] = "";
}

int fa_host::fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
int fa_host::cuda_fail(cudaError_t e, const char* what) {
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return (int)e;
}

namespace {
using fa_host::cuda_fail;
using fa_host::fail;

constexpr float kLog2e = 1.4426950408889634f;

int check_common(const char* fn, int B, int H, int N, int D, int dtype, float scale) {
  if (B <= 0 || H <= 0 || N <= 0 || D <= 0) return fail(-1, "%s: B, H, N, D must be positive (got %d,%d,%d,%d)", fn, B, H, N, D);
  if (dtype != FA_DTYPE_F16 && dtype != FA_DTYPE_BF16 && dtype != FA_DTYPE_F32 && dtype != FA_DTYPE_F8E4M3 &&
      dtype != FA_DTYPE_F8E5M2)
    return fail(-2, "%s: dtype %d not supported (0 = f16, 1 = bf16, 2 = f32, 3 = f8e4m3, 4 = f8e5m2)", fn, dtype);
  if (dtype == FA_DTYPE_F8E4M3 || dtype == FA_DTYPE_F8E5M2) {
    if (D != 128) return fail(-3, "%s: FP8 head size must be 128 (got %d); pad in the caller", fn, D);
  } else if (dtype == FA_DTYPE_F32) {
    if (!(D == 16 || D == 32 || D == 64 || D == 128))
      return fail(-3, "%s: float32 head size must be 16, 32, 64 or 128 (got %d); pad in the caller", fn, D);
  } else if (!(D == 64 || D == 128)) {
    return fail(-3, "%s: 16-bit head size must be 64 or 128 (got %d); pad in the caller", fn, D);
  }
  if (!(scale > 0.f) || !std::isfinite(scale)) return fail(-4, "%s: softmax_scale must be positive and finite", fn);
  if (H > 65535 || B > 65535) return fail(-5, "%s: B and H must be <= 65535", fn);
  return 0;
}
int check_tensor(const char* fn, const char* name, const void* p, const int64_t s[4], int dtype, int B, int H) {
  if (!p) return fail(-6, "%s: %s is null", fn, name);
  if (s[3] != 1) return fail(-7, "%s: %s last-dim stride must be 1 (got %lld)", fn, name, (long long)s[3]);
  const int64_t gran = (dtype == FA_DTYPE_F32) ? 4 : (dtype >= FA_DTYPE_F8E4M3 ? 16 : 8);  // 16 bytes
  if ((reinterpret_cast<uintptr_t>(p) & 15u) != 0) return fail(-8, "%s: %s must be 16-byte aligned", fn, name);
  if (s[2] % gran != 0 || (H > 1 && s[1] % gran != 0) || (B > 1 && s[0] % gran != 0))
    return fail(-9, "%s: %s strides must be multiples of 16 bytes", fn, name);
  return 0;
}

// dropout_p -> DropParams (fa_dropout.cuh): the probability is quantised to 1/256; 0 (or anything that rounds to 0) = off
int make_drop(const char* fn, float dropout_p, uint64_t seed, fa::DropParams* d) {
  *d = fa::DropParams{0u, 0u, 0u, 1.0f};
  if (!(dropout_p >= 0.f) || !(dropout_p < 1.f)) return fail(-13, "%s: dropout_p must be in [0, 1) (got %g)", fn, (double)dropout_p);
  long t = lroundf(dropout_p * 256.0f);
  if (t > 255) t = 255;
  d->thresh = (uint32_t)t;
  d->seed_lo = (uint32_t)seed, d->seed_hi = (uint32_t)(seed >> 32);
  d->rp = 256.0f / (float)(256 - t);
  return 0;
}

// attention mask: one bit per entry [.., row, col / 8], 1 = attend; strides {sB, sH, sRow} in bytes (sB / sH may be 0 = broadcast)
int check_amask(const char* fn, const char* name, const uint8_t* m, const int64_t s[3], int N) {
  if (!m) return 0;
  if (!s) return fail(-14, "%s: %s given without strides", fn, name);
  const int64_t pitch = ((int64_t)N + 127) / 128 * 16;   // 16 bytes of bits per 128-entry block
  if ((reinterpret_cast<uintptr_t>(m) & 15u) != 0) return fail(-14, "%s: %s must be 16-byte aligned", fn, name);
  if (s[2] < pitch || s[2] % 16 != 0)
    return fail(-14, "%s: %s row pitch must be a multiple of 16 bytes and >= 16 bytes per 128 entries (%lld), got %lld", fn,
                name, (long long)pitch, (long long)s[2]);
  if (s[0] < 0 || s[1] < 0 || s[0] % 16 != 0 || s[1] % 16 != 0)
    return fail(-14, "%s: %s batch / head strides must be non-negative multiples of 16 bytes (0 = broadcast)", fn, name);
  return 0;
}

// 128 x 128 block summary of a mask: bytes [.., query block, key block], strides {sB, sH, sI} in bytes
int check_ablock(const char* fn, const fa_attn_mask* m) {
  if (m && !m->rows && (m->window_left < 0 || m->window_right < 0))
    return fail(-14, "%s: attn_mask without rows is a band mask and needs window_left, window_right >= 0", fn);
  if (!m || !m->blocks) return 0;
  const int64_t* s = m->blocks_strides;
  if (s[0] < 0 || s[1] < 0 || s[2] <= 0) return fail(-14, "%s: attn_mask->blocks strides must be non-negative (row stride positive)", fn);
  return 0;
}

void fill3(int64_t (&dst)[3], const int64_t s[4]) {
  dst[0] = s[0], dst[1] = s[1], dst[2] = s[2];
}

// ---------------------------------------------------------------------------------------------- forward (float32)
template <int kD>
int launch_fwd32(const fa::SimtParams& p, cudaStream_t st) {
  auto kern = fa::fa_fwd_f32_kernel<kD>;
  const int bytes = fa::SimtSmem<kD>::fwd_floats * 4;
  static std::atomic<uint64_t> smem_set{0};
  if (int r = fa_host::set_smem_once(kern, bytes, smem_set)) return r;
  dim3 grid((p.N + 63) / 64, p.H, p.B);
  kern<<<grid, 256, bytes, st>>>(p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "fa_fwd(f32) launch");
}

template <int kD>
int launch_bwd32(const fa::SimtParams& p, int which, cudaStream_t st) {
  if (which & FA_BWD_DKDV) {
    auto kern = fa::fa_bwd_dkdv_f32_kernel<kD>;
    const int bytes = fa::SimtSmem<kD>::dkdv_floats * 4;
    static std::atomic<uint64_t> smem_set{0};
    if (int r = fa_host::set_smem_once(kern, bytes, smem_set)) return r;
    dim3 grid((p.N + 63) / 64, p.H, p.B);
    kern<<<grid, 256, bytes, st>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "fa_bwd(f32 dK/dV) launch");
  }
  if (which & FA_BWD_DQ) {
    auto kern = fa::fa_bwd_dq_f32_kernel<kD>;
    const int bytes = fa::SimtSmem<kD>::dq_floats * 4;
    static std::atomic<uint64_t> smem_set{0};
    if (int r = fa_host::set_smem_once(kern, bytes, smem_set)) return r;
    dim3 grid((p.N + 63) / 64, p.H, p.B);
    kern<<<grid, 256, bytes, st>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "fa_bwd(f32 dQ) launch");
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------- backward (16-bit)
// Workspace of the fused backward: [ticket | pad to 256 B][turn counters 2 x tiles x chunks, padded to 256 B][fp32 dQ tiles A][B]
struct FusedLayout {
  size_t ctrl_bytes, acc_floats, total_bytes;
  int n_blocks;
  long long tiles;
};
FusedLayout fused_layout(int B, int H, int N, int D, int causal) {
  FusedLayout L{};
  L.n_blocks = (N + 127) / 128;
  L.tiles = (long long)B * H * L.n_blocks;
  L.ctrl_bytes = 256 + (((size_t)L.tiles * 2 * fa::kSemPerTile * sizeof(int) + 255) & ~(size_t)255);
  L.acc_floats = (size_t)L.tiles * 128 * D;
  L.total_bytes = L.ctrl_bytes + L.acc_floats * sizeof(float) * (causal ? 1 : 2);
  return L;
}

template <bool kBf16, int kD, bool kCausal>
int launch_bwd16_fused(const fa::BwdMaps& m, const fa::BwdParams& p, void* workspace, cudaStream_t st) {
  const FusedLayout L = fused_layout(p.B, p.H, p.N, kD, kCausal);
  fa::FusedParams fp{};
  fp.base = p;
  char* ws = static_cast<char*>(workspace);
  fp.ticket = reinterpret_cast<int*>(ws);
  fp.sem = reinterpret_cast<int*>(ws + 256);
  fp.dq_acc = reinterpret_cast<float*>(ws + L.ctrl_bytes);
  fp.acc_b_off = kCausal ? 0 : (long long)L.acc_floats;
  fp.n_blocks = L.n_blocks;
#if FA_FUSED_TMA && FA_FUSED_UNORDERED   // timing experiments: every contribution is a reduction into a zeroed workspace
  cudaError_t e = cudaMemsetAsync(ws, 0, L.total_bytes, st);
#else
  cudaError_t e = cudaMemsetAsync(ws, 0, L.ctrl_bytes, st);
#endif
  if (e != cudaSuccess) return cuda_fail(e, "fa_bwd(fused) cudaMemsetAsync");
  auto kern = fa::fa_bwd_fused_kernel<kBf16, kD, kCausal>;
  static std::atomic<uint64_t> smem_set{0};
  if (int r = fa_host::set_smem_once(kern, fa::FusedCfg<kD>::kSmemBytes, smem_set)) return r;
  kern<<<(unsigned)L.tiles, fa::FusedCfg<kD>::kThreads, fa::FusedCfg<kD>::kSmemBytes, st>>>(m.q, m.k, m.v, m.dout, fp);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "fa_bwd(fused) launch");
  dim3 grid(L.n_blocks, p.H, p.B);
  fa::fa_bwd_dq_convert_kernel<kBf16, kD><<<grid, 256, 0, st>>>(fp.dq_acc, fp.acc_b_off, p.dq, p.dq_s[0], p.dq_s[1],
                                                               p.dq_s[2], p.H, p.N, L.n_blocks, p.scale);
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "fa_bwd(dQ convert) launch");
}

}  // namespace

#if FA_TRACE
// debug builds only (not declared in include/fa_b200.h)
extern "C" int fa_debug_set_trace(void* dev_buf, int capacity_events) {
  long long* p = static_cast<long long*>(dev_buf);
  cudaMemcpyToSymbol(fa::g_fa_trace, &p, sizeof(p));
  cudaMemcpyToSymbol(fa::g_fa_trace_cap, &capacity_events, sizeof(int));
  return 0;
}
#endif

extern "C" {

int fa_version(void) { return 9; }

const char* fa_last_error(void) { return g_err; }

int fa_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int N, int D,
           const int64_t q_strides[4], const int64_t k_strides[4], const int64_t v_strides[4],
           const int64_t o_strides[4], int dtype, float softmax_scale, int causal, void* stream) {
  return fa_fwd_peers(q, k, v, o, lse, B, H, N, D, q_strides, k_strides, v_strides, o_strides, dtype, softmax_scale,
                      causal, 0, nullptr, nullptr, 0.f, 0, nullptr, stream);
}

// Rectangular problems (Nk key / value rows != N query rows): 16-bit tcgen05 kernels, non-causal, no seqlens / dropout / mask.
static int check_rect(const char* fn, int N, int Nk, int dtype, int causal, const int32_t* seqlens, float dropout_p,
                      const fa_attn_mask* mask) {
  if (Nk == N) return 0;
  if (Nk <= 0) return fail(-1, "%s: Nk must be positive (got %d)", fn, Nk);
  if (dtype != FA_DTYPE_F16 && dtype != FA_DTYPE_BF16)
    return fail(-2, "%s: a key length different from the query length needs float16 / bfloat16 (dtype %d)", fn, dtype);
  if (causal) return fail(-15, "%s: the causal mask needs Nk == N (got %d, %d)", fn, Nk, N);
  if (seqlens || mask || dropout_p != 0.f)
    return fail(-15, "%s: seqlens, dropout and attention masks need Nk == N", fn);
  return 0;
}

static int fwd_impl(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int N, int Nk, int D,
                    const int64_t q_strides[4], const int64_t k_strides[4], const int64_t v_strides[4],
                    const int64_t o_strides[4], int dtype, float softmax_scale, int causal, int n_peers,
                    void* const* peer_o, const int32_t* seqlens, float dropout_p, uint64_t dropout_seed,
                    const fa_attn_mask* mask, void* stream);

int fa_fwd_peers(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int N, int D,
                 const int64_t q_strides[4], const int64_t k_strides[4], const int64_t v_strides[4],
                 const int64_t o_strides[4], int dtype, float softmax_scale, int causal, int n_peers,
                 void* const* peer_o, const int32_t* seqlens, float dropout_p, uint64_t dropout_seed,
                 const fa_attn_mask* mask, void* stream) {
  return fwd_impl(q, k, v, o, lse, B, H, N, N, D, q_strides, k_strides, v_strides, o_strides, dtype, softmax_scale, causal,
                  n_peers, peer_o, seqlens, dropout_p, dropout_seed, mask, stream);
}

int fa_fwd_rect(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Nq, int Nk, int D,
                const int64_t q_strides[4], const int64_t k_strides[4], const int64_t v_strides[4],
                const int64_t o_strides[4], int dtype, float softmax_scale, void* stream) {
  return fwd_impl(q, k, v, o, lse, B, H, Nq, Nk, D, q_strides, k_strides, v_strides, o_strides, dtype, softmax_scale, 0, 0,
                  nullptr, nullptr, 0.f, 0, nullptr, stream);
}

static int fwd_impl(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int N, int Nk, int D,
                    const int64_t q_strides[4], const int64_t k_strides[4], const int64_t v_strides[4],
                    const int64_t o_strides[4], int dtype, float softmax_scale, int causal, int n_peers,
                    void* const* peer_o, const int32_t* seqlens, float dropout_p, uint64_t dropout_seed,
                    const fa_attn_mask* mask, void* stream) {
  g_err[0] = 0;
  if (int r = check_rect("fa_fwd", N, Nk, dtype, causal, seqlens, dropout_p, mask)) return r;
  const uint8_t* attn_mask = mask ? mask->rows : nullptr;
  const int64_t* attn_mask_strides = mask ? mask->rows_strides : nullptr;
  if (int r = check_amask("fa_fwd_peers", "attn_mask->rows", attn_mask, attn_mask_strides, N)) return r;
  if (int r = check_ablock("fa_fwd_peers", mask)) return r;
  const bool band = mask && !mask->rows;
  if (mask && (dtype == FA_DTYPE_F8E4M3 || dtype == FA_DTYPE_F8E5M2))
    return fail(-14, "fa_fwd_peers: attention masks are not implemented for the FP8 forward");
  fa::DropParams drop;
  if (int r = make_drop("fa_fwd_peers", dropout_p, dropout_seed, &drop)) return r;
  if (drop.thresh && (dtype == FA_DTYPE_F8E4M3 || dtype == FA_DTYPE_F8E5M2))
    return fail(-13, "fa_fwd_peers: dropout is not implemented for the FP8 forward");
  if (n_peers < 0 || n_peers > 7 || (n_peers > 0 && !peer_o)) return fail(-12, "fa_fwd_peers: 0 <= n_peers <= 7 and peer_o non-null");
  if (n_peers > 0 && dtype == FA_DTYPE_F32) return fail(-12, "fa_fwd_peers: peer copies are implemented for the 16-bit and FP8 kernels");
  for (int i = 0; i < n_peers; ++i)
    if (!peer_o[i] || (reinterpret_cast<uintptr_t>(peer_o[i]) & 15u)) return fail(-8, "fa_fwd_peers: peer_o[%d] must be a 16-byte aligned device pointer", i);
  if (int r = check_common("fa_fwd", B, H, N, D, dtype, softmax_scale)) return r;
  if (!q_strides || !k_strides || !v_strides || !o_strides) return fail(-6, "fa_fwd: null stride array");
  if (int r = check_tensor("fa_fwd", "q", q, q_strides, dtype, B, H)) return r;
  if (int r = check_tensor("fa_fwd", "k", k, k_strides, dtype, B, H)) return r;
  if (int r = check_tensor("fa_fwd", "v", v, v_strides, dtype, B, H)) return r;
  if (int r = check_tensor("fa_fwd", "o", o, o_strides, dtype, B, H)) return r;
  if (!lse) return fail(-6, "fa_fwd: lse is null");
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  if (dtype == FA_DTYPE_F32) {
    fa::SimtParams p{};
    p.q = static_cast<const float*>(q), p.k = static_cast<const float*>(k), p.v = static_cast<const float*>(v);
    p.out_o = static_cast<float*>(o), p.out_lse = lse;
    p.B = B, p.H = H, p.N = N, p.D = D;
    fill3(p.q_s, q_strides), fill3(p.k_s, k_strides), fill3(p.v_s, v_strides), fill3(p.o_s, o_strides);
    p.scale = softmax_scale, p.scale_log2 = softmax_scale * kLog2e, p.causal = causal ? 1 : 0;
    p.seqlens = seqlens;
    p.drop = drop;
    p.amask = attn_mask;
    if (attn_mask) p.am_s[0] = attn_mask_strides[0], p.am_s[1] = attn_mask_strides[1], p.am_s[2] = attn_mask_strides[2];
    if (band) p.band = 1, p.win_left = mask->window_left, p.win_right = mask->window_right;
    if (mask && mask->blocks) {
      p.ablock = mask->blocks;
      for (int i = 0; i < 3; ++i) p.ab_s[i] = mask->blocks_strides[i];
    }
    switch (D) {
      case 16: return launch_fwd32<16>(p, st);
      case 32: return launch_fwd32<32>(p, st);
      case 64: return launch_fwd32<64>(p, st);
      default: return launch_fwd32<128>(p, st);
    }
  }

  const int bf = dtype == FA_DTYPE_BF16;
  const bool f8 = dtype == FA_DTYPE_F8E4M3 || dtype == FA_DTYPE_F8E5M2;
  CUtensorMap tq, tk, tv;
  auto encode = [&](CUtensorMap* m, const void* ptr, const int64_t* st4, int rows) {
    return fa::cached_tmap_bhnd(m, ptr, f8 ? 2 : bf, B, H, rows, D, st4[0], st4[1], st4[2], 128);
  };
  if (int r = encode(&tq, q, q_strides, N)) return fail(r, "fa_fwd: cuTensorMapEncodeTiled(q) failed (%d)", r);
  if (int r = encode(&tk, k, k_strides, Nk)) return fail(r, "fa_fwd: cuTensorMapEncodeTiled(k) failed (%d)", r);
  if (int r = encode(&tv, v, v_strides, Nk)) return fail(r, "fa_fwd: cuTensorMapEncodeTiled(v) failed (%d)", r);
  fa::FwdParams p{};
  p.o = o, p.lse = lse, p.B = B, p.H = H, p.N = N;
  p.Nk = Nk != N ? Nk : 0;
  p.o_sB = o_strides[0], p.o_sH = o_strides[1], p.o_sN = o_strides[2];
  p.scale_log2 = softmax_scale * kLog2e;
  p.q_blocks = (N + 255) / 256;
  p.n_peer = n_peers;
  for (int i = 0; i < n_peers; ++i) p.o_peer[i] = peer_o[i];
  p.seqlens = seqlens;
  p.drop = drop;
  p.amask = attn_mask;
  if (attn_mask) p.am_sB = attn_mask_strides[0], p.am_sH = attn_mask_strides[1], p.am_sN = attn_mask_strides[2];
  if (band) p.band = 1, p.win_left = mask->window_left, p.win_right = mask->window_right;
  if (mask && mask->blocks)
    p.ablock = mask->blocks, p.ab_sB = mask->blocks_strides[0], p.ab_sH = mask->blocks_strides[1], p.ab_sI = mask->blocks_strides[2];
  if (Nk == N && fa_host::fwd_pair_eligible(dtype, D, p)) {
    CUtensorMap tk64;
    if (int r = fa::cached_tmap_bhnd(&tk64, k, bf, B, H, N, D, k_strides[0], k_strides[1], k_strides[2], 64))
      return fail(r, "fa_fwd: cuTensorMapEncodeTiled(k, 64-row box) failed (%d)", r);
    return fa_host::launch_fwd16_pair(dtype, D, causal != 0, tq, tk64, tv, p, H, B, st);
  }
  return fa_host::launch_fwd16(dtype, D, causal != 0, tq, tk, tv, p, H, B, st);
}

int fa_bwd_preprocess(const void* o, const void* dout, float* delta, int B, int H, int N, int D,
                      const int64_t o_strides[4], const int64_t do_strides[4], int dtype, void* stream) {
  g_err[0] = 0;
  if (dtype == FA_DTYPE_F8E4M3 || dtype == FA_DTYPE_F8E5M2)
    return fail(-2, "fa_bwd_preprocess: FP8 is forward-only (dtype %d)", dtype);
  if (int r = check_common("fa_bwd_preprocess", B, H, N, D, dtype, 1.0f)) return r;
  if (!o_strides || !do_strides) return fail(-6, "fa_bwd_preprocess: null stride array");
  if (int r = check_tensor("fa_bwd_preprocess", "o", o, o_strides, dtype, B, H)) return r;
  if (int r = check_tensor("fa_bwd_preprocess", "dout", dout, do_strides, dtype, B, H)) return r;
  if (!delta) return fail(-6, "fa_bwd_preprocess: delta is null");
  fa::PreParams p{};
  p.o = o, p.dout = dout, p.delta = delta, p.B = B, p.H = H, p.N = N, p.D = D;
  p.o_sB = o_strides[0], p.o_sH = o_strides[1], p.o_sN = o_strides[2];
  p.do_sB = do_strides[0], p.do_sH = do_strides[1], p.do_sN = do_strides[2];
  p.total_rows = (long long)B * H * N;
  const int vec = dtype == FA_DTYPE_F32 ? 4 : 8;
  const int tpr = D / vec;
  const long long rows_per_cta = (256 / tpr) * 4;
  long long ctas = (p.total_rows + rows_per_cta - 1) / rows_per_cta;
  // grid-stride beyond kPreCtasPerSm CTAs per SM (FA_PRE_CTAS_PER_SM overrides it for measurements)
  static const long long per_sm = [] {
    const char* e = std::getenv("FA_PRE_CTAS_PER_SM");
    const long long v = e ? std::atoll(e) : 0;
    return v > 0 ? v : 32LL;
  }();
  if (ctas > 148 * per_sm) ctas = 148 * per_sm;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define FA_PRE_CASE(E, T)                                                     \
  if (dtype == E && tpr == T) {                                               \
    fa::fa_bwd_preprocess_kernel<E, T><<<(unsigned)ctas, 256, 0, st>>>(p);    \
    cudaError_t e = cudaGetLastError();                                       \
    return e == cudaSuccess ? 0 : cuda_fail(e, "fa_bwd_preprocess launch");   \
  }
  FA_PRE_CASE(0, 8) FA_PRE_CASE(0, 16) FA_PRE_CASE(1, 8) FA_PRE_CASE(1, 16)
  FA_PRE_CASE(2, 4) FA_PRE_CASE(2, 8) FA_PRE_CASE(2, 16) FA_PRE_CASE(2, 32)
#undef FA_PRE_CASE
  return fail(-3, "fa_bwd_preprocess: no kernel for dtype %d D %d", dtype, D);
}

size_t fa_bwd_workspace_bytes(int B, int H, int N, int D, int dtype, int causal, int which) {
  if (B <= 0 || H <= 0 || N <= 0 || D <= 0) return 0;
  // the two-kernel backward needs no scratch: every gradient tile has a single owner CTA
  if (which != FA_BWD_FUSED || dtype == FA_DTYPE_F32) return 0;
  return fused_layout(B, H, N, D, causal).total_bytes;
}

int fa_bwd(const void* q, const void* k, const void* v, const void* dout, const float* lse, const float* delta,
           void* dq, void* dk, void* dv, void* workspace, size_t workspace_bytes, int B, int H, int N, int D,
           const int64_t q_strides[4], const int64_t k_strides[4], const int64_t v_strides[4],
           const int64_t do_strides[4], const int64_t dq_strides[4], const int64_t dk_strides[4],
           const int64_t dv_strides[4], int dtype, float softmax_scale, int causal, void* stream) {
  // The two-kernel path (every gradient tile has one owner CTA, no workspace) is the default: on B200 it is faster than
  // the single-pass kernel, whose dQ reduction is bound by the SM -> L2 path.
  return fa_bwd_partial(q, k, v, dout, lse, delta, dq, dk, dv, workspace, workspace_bytes, B, H, N, D, q_strides,
                        k_strides, v_strides, do_strides, dq_strides, dk_strides, dv_strides, dtype, softmax_scale,
                        causal, FA_BWD_DKDV | FA_BWD_DQ, nullptr, 0.f, 0, nullptr, stream);
}

static int bwd_impl(const void* q, const void* k, const void* v, const void* dout, const float* lse,
                    const float* delta, void* dq, void* dk, void* dv, void* workspace, size_t workspace_bytes, int B,
                    int H, int N, int Nk, int D, const int64_t q_strides[4], const int64_t k_strides[4],
                    const int64_t v_strides[4], const int64_t do_strides[4], const int64_t dq_strides[4],
                    const int64_t dk_strides[4], const int64_t dv_strides[4], int dtype, float softmax_scale,
                    int causal, int which, const int32_t* seqlens, float dropout_p, uint64_t dropout_seed,
                    const fa_attn_mask* mask, void* stream);

int fa_bwd_partial(const void* q, const void* k, const void* v, const void* dout, const float* lse,
                   const float* delta, void* dq, void* dk, void* dv, void* workspace, size_t workspace_bytes, int B,
                   int H, int N, int D, const int64_t q_strides[4], const int64_t k_strides[4],
                   const int64_t v_strides[4], const int64_t do_strides[4], const int64_t dq_strides[4],
                   const int64_t dk_strides[4], const int64_t dv_strides[4], int dtype, float softmax_scale,
                   int causal, int which, const int32_t* seqlens, float dropout_p, uint64_t dropout_seed,
                   const fa_attn_mask* mask, void* stream) {
  return bwd_impl(q, k, v, dout, lse, delta, dq, dk, dv, workspace, workspace_bytes, B, H, N, N, D, q_strides, k_strides,
                  v_strides, do_strides, dq_strides, dk_strides, dv_strides, dtype, softmax_scale, causal, which, seqlens,
                  dropout_p, dropout_seed, mask, stream);
}

int fa_bwd_rect(const void* q, const void* k, const void* v, const void* dout, const float* lse, const float* delta,
                void* dq, void* dk, void* dv, int B, int H, int Nq, int Nk, int D, const int64_t q_strides[4],
                const int64_t k_strides[4], const int64_t v_strides[4], const int64_t do_strides[4],
                const int64_t dq_strides[4], const int64_t dk_strides[4], const int64_t dv_strides[4], int dtype,
                float softmax_scale, void* stream) {
  return bwd_impl(q, k, v, dout, lse, delta, dq, dk, dv, nullptr, 0, B, H, Nq, Nk, D, q_strides, k_strides, v_strides,
                  do_strides, dq_strides, dk_strides, dv_strides, dtype, softmax_scale, 0, FA_BWD_DKDV | FA_BWD_DQ, nullptr,
                  0.f, 0, nullptr, stream);
}

static int bwd_impl(const void* q, const void* k, const void* v, const void* dout, const float* lse,
                    const float* delta, void* dq, void* dk, void* dv, void* workspace, size_t workspace_bytes, int B,
                    int H, int N, int Nk, int D, const int64_t q_strides[4], const int64_t k_strides[4],
                    const int64_t v_strides[4], const int64_t do_strides[4], const int64_t dq_strides[4],
                    const int64_t dk_strides[4], const int64_t dv_strides[4], int dtype, float softmax_scale,
                    int causal, int which, const int32_t* seqlens, float dropout_p, uint64_t dropout_seed,
                    const fa_attn_mask* mask, void* stream) {
  g_err[0] = 0;
  if (int r = check_rect("fa_bwd", N, Nk, dtype, causal, seqlens, dropout_p, mask)) return r;
  if (Nk != N && which == FA_BWD_FUSED) return fail(-15, "fa_bwd: FA_BWD_FUSED needs Nk == N");
  const uint8_t* attn_mask = mask ? mask->rows : nullptr;
  const uint8_t* attn_mask_t = mask ? mask->cols : nullptr;
  const int64_t* attn_mask_strides = mask ? mask->rows_strides : nullptr;
  const int64_t* attn_mask_t_strides = mask ? mask->cols_strides : nullptr;
  if (int r = check_amask("fa_bwd_partial", "attn_mask->rows", attn_mask, attn_mask_strides, N)) return r;
  if (int r = check_amask("fa_bwd_partial", "attn_mask->cols", attn_mask_t, attn_mask_t_strides, N)) return r;
  if (int r = check_ablock("fa_bwd_partial", mask)) return r;
  const bool band = mask && !mask->rows;
  if (mask && which == FA_BWD_FUSED)
    return fail(-14, "fa_bwd_partial: FA_BWD_FUSED takes no attention mask; use the two-kernel path");
  if (attn_mask && !attn_mask_t && dtype != FA_DTYPE_F32)
    return fail(-14, "fa_bwd_partial: the 16-bit kernels need the transposed mask (attn_mask->cols) as well");
  fa::DropParams drop;
  if (int r = make_drop("fa_bwd_partial", dropout_p, dropout_seed, &drop)) return r;
  if (drop.thresh && which == FA_BWD_FUSED)
    return fail(-13, "fa_bwd_partial: FA_BWD_FUSED has no dropout; use the two-kernel path");
  if (which != FA_BWD_FUSED && ((which & (FA_BWD_DKDV | FA_BWD_DQ)) == 0 || (which & ~(FA_BWD_DKDV | FA_BWD_DQ))))
    return fail(-10, "fa_bwd_partial: which must be FA_BWD_FUSED or a non-empty subset of FA_BWD_DKDV | FA_BWD_DQ");
  if (which == FA_BWD_FUSED) {
    if (seqlens) return fail(-10, "fa_bwd_partial: FA_BWD_FUSED does not take per-batch sequence lengths; use the two-kernel path");
    if (dtype == FA_DTYPE_F32) return fail(-10, "fa_bwd_partial: FA_BWD_FUSED is a 16-bit kernel; float32 uses FA_BWD_DKDV | FA_BWD_DQ");
    const size_t need = fa_bwd_workspace_bytes(B, H, N, D, dtype, causal, FA_BWD_FUSED);
    if (!workspace || workspace_bytes < need)
      return fail(-11, "fa_bwd: workspace of %zu bytes required (got %zu); size it with fa_bwd_workspace_bytes", need,
                  workspace ? workspace_bytes : (size_t)0);
    if ((reinterpret_cast<uintptr_t>(workspace) & 255u) != 0) return fail(-8, "fa_bwd: workspace must be 256-byte aligned");
  }
  if (dtype == FA_DTYPE_F8E4M3 || dtype == FA_DTYPE_F8E5M2) return fail(-2, "fa_bwd: FP8 is forward-only (dtype %d)", dtype);
  if (int r = check_common("fa_bwd", B, H, N, D, dtype, softmax_scale)) return r;
  if (!q_strides || !k_strides || !v_strides || !do_strides || !dq_strides || !dk_strides || !dv_strides)
    return fail(-6, "fa_bwd: null stride array");
  if (int r = check_tensor("fa_bwd", "q", q, q_strides, dtype, B, H)) return r;
  if (int r = check_tensor("fa_bwd", "k", k, k_strides, dtype, B, H)) return r;
  if (int r = check_tensor("fa_bwd", "v", v, v_strides, dtype, B, H)) return r;
  if (int r = check_tensor("fa_bwd", "dout", dout, do_strides, dtype, B, H)) return r;
  if (int r = check_tensor("fa_bwd", "dq", dq, dq_strides, dtype, B, H)) return r;
  if (int r = check_tensor("fa_bwd", "dk", dk, dk_strides, dtype, B, H)) return r;
  if (int r = check_tensor("fa_bwd", "dv", dv, dv_strides, dtype, B, H)) return r;
  if (!lse || !delta) return fail(-6, "fa_bwd: lse / delta is null");
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  if (dtype == FA_DTYPE_F32) {
    fa::SimtParams p{};
    p.q = static_cast<const float*>(q), p.k = static_cast<const float*>(k), p.v = static_cast<const float*>(v);
    p.dout = static_cast<const float*>(dout), p.lse = lse, p.delta = delta;
    p.dq = static_cast<float*>(dq), p.dk = static_cast<float*>(dk), p.dv = static_cast<float*>(dv);
    p.B = B, p.H = H, p.N = N, p.D = D;
    fill3(p.q_s, q_strides), fill3(p.k_s, k_strides), fill3(p.v_s, v_strides), fill3(p.do_s, do_strides);
    fill3(p.dq_s, dq_strides), fill3(p.dk_s, dk_strides), fill3(p.dv_s, dv_strides);
    p.scale = softmax_scale, p.scale_log2 = softmax_scale * kLog2e, p.causal = causal ? 1 : 0;
    p.seqlens = seqlens;
    p.drop = drop;
    p.amask = attn_mask;
    if (attn_mask) p.am_s[0] = attn_mask_strides[0], p.am_s[1] = attn_mask_strides[1], p.am_s[2] = attn_mask_strides[2];
    if (band) p.band = 1, p.win_left = mask->window_left, p.win_right = mask->window_right;
    if (mask && mask->blocks) {
      p.ablock = mask->blocks;
      for (int i = 0; i < 3; ++i) p.ab_s[i] = mask->blocks_strides[i];
    }
    switch (D) {
      case 16: return launch_bwd32<16>(p, which, st);
      case 32: return launch_bwd32<32>(p, which, st);
      case 64: return launch_bwd32<64>(p, which, st);
      default: return launch_bwd32<128>(p, which, st);
    }
  }

  const int bf = dtype == FA_DTYPE_BF16;
  fa::BwdMaps m;
  if (int r = fa::cached_tmap_bhnd(&m.q, q, bf, B, H, N, D, q_strides[0], q_strides[1], q_strides[2], 128))
    return fail(r, "fa_bwd: cuTensorMapEncodeTiled(q) failed (%d)", r);
  if (int r = fa::cached_tmap_bhnd(&m.k, k, bf, B, H, Nk, D, k_strides[0], k_strides[1], k_strides[2], 128))
    return fail(r, "fa_bwd: cuTensorMapEncodeTiled(k) failed (%d)", r);
  if (int r = fa::cached_tmap_bhnd(&m.v, v, bf, B, H, Nk, D, v_strides[0], v_strides[1], v_strides[2], 128))
    return fail(r, "fa_bwd: cuTensorMapEncodeTiled(v) failed (%d)", r);
  if (int r = fa::cached_tmap_bhnd(&m.dout, dout, bf, B, H, N, D, do_strides[0], do_strides[1], do_strides[2], 128))
    return fail(r, "fa_bwd: cuTensorMapEncodeTiled(dout) failed (%d)", r);
  fa::BwdParams p{};
  p.lse = lse, p.delta = delta, p.dq = dq, p.dk = dk, p.dv = dv;
  p.B = B, p.H = H, p.N = N;
  p.Nk = Nk != N ? Nk : 0;
  fill3(p.dq_s, dq_strides), fill3(p.dk_s, dk_strides), fill3(p.dv_s, dv_strides);
  p.scale = softmax_scale, p.scale_log2 = softmax_scale * kLog2e;
  p.seqlens = seqlens;
  p.drop = drop;
  p.amask = attn_mask, p.amask_t = attn_mask_t;
  if (attn_mask) {
    for (int i = 0; i < 3; ++i) p.am_s[i] = attn_mask_strides[i], p.amt_s[i] = attn_mask_t_strides[i];
  }
  if (band) p.band = 1, p.win_left = mask->window_left, p.win_right = mask->window_right;
  if (mask && mask->blocks) {
    p.ablock = mask->blocks;
    for (int i = 0; i < 3; ++i) p.ab_s[i] = mask->blocks_strides[i];
  }
  if (which != FA_BWD_FUSED) {
    if (which & FA_BWD_DKDV)
      if (int r = fa_host::launch_bwd16_dkdv(bf != 0, D, causal != 0, m, p, st)) return r;
    if (which & FA_BWD_DQ)
      if (int r = fa_host::launch_bwd16_dq(bf != 0, D, causal != 0, m, p, st)) return r;
    return 0;
  }
#define FA_BWD_CASE(BF, DD, C) \
  if (bf == BF && D == DD && (causal != 0) == C) return launch_bwd16_fused<BF, DD, C>(m, p, workspace, st);
  FA_BWD_CASE(true, 128, true)
  FA_BWD_CASE(true, 128, false)
  FA_BWD_CASE(true, 64, true)
  FA_BWD_CASE(true, 64, false)
  FA_BWD_CASE(false, 128, true)
  FA_BWD_CASE(false, 128, false)
  FA_BWD_CASE(false, 64, true)
  FA_BWD_CASE(false, 64, false)
#undef FA_BWD_CASE
  return fail(-3, "fa_bwd: no kernel for dtype %d D %d", dtype, D);
}

// ---------------------------------------------------------------------------------------------- ring attention helpers
static int check_rows(const char* fn, long long rows, int D, int dtype) {
  if (rows <= 0 || D <= 0 || D % 8 != 0 || D > 256) return fail(-1, "%s: rows > 0 and D a multiple of 8, <= 256 (got %lld, %d)", fn, rows, D);
  if (dtype != FA_DTYPE_F16 && dtype != FA_DTYPE_BF16) return fail(-2, "%s: 16-bit partials only (dtype %d)", fn, dtype);
  if (256 % (D / 8) != 0 || D / 8 > 32) return fail(-3, "%s: D / 8 must divide 256 and be <= 32 (D = %d)", fn, D);
  return 0;
}
static unsigned grid_for(long long work_items) {
  long long g = (work_items + 255) / 256;
  return (unsigned)(g < 1 ? 1 : (g > 148LL * 16 ? 148LL * 16 : g));
}

int fa_merge_partial(float* o_acc, float* lse_acc, const void* o_part, const float* lse_part, long long rows, int D,
                     int dtype, int first, void* stream) {
  g_err[0] = 0;
  if (int r = check_rows("fa_merge_partial", rows, D, dtype)) return r;
  if (!o_acc || !lse_acc || !o_part || !lse_part) return fail(-6, "fa_merge_partial: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int tpr = D / 8;
  const unsigned grid = grid_for(rows * tpr);
  const uint16_t* part = static_cast<const uint16_t*>(o_part);
#define FA_MERGE_CASE(T)                                                                                              \
  if (tpr == T) {                                                                                                     \
    if (dtype == FA_DTYPE_BF16)                                                                                       \
      fa::fa_merge_partial_kernel<true, T><<<grid, 256, 0, st>>>(o_acc, lse_acc, part, lse_part, rows, first);        \
    else                                                                                                              \
      fa::fa_merge_partial_kernel<false, T><<<grid, 256, 0, st>>>(o_acc, lse_acc, part, lse_part, rows, first);       \
    cudaError_t e = cudaGetLastError();                                                                               \
    return e == cudaSuccess ? 0 : cuda_fail(e, "fa_merge_partial launch");                                            \
  }
  FA_MERGE_CASE(1) FA_MERGE_CASE(2) FA_MERGE_CASE(4) FA_MERGE_CASE(8) FA_MERGE_CASE(16) FA_MERGE_CASE(32)
#undef FA_MERGE_CASE
  return fail(-3, "fa_merge_partial: no kernel for D %d", D);
}

int fa_accumulate(float* acc, const void* part, long long n, int dtype, int first, void* stream) {
  g_err[0] = 0;
  if (n <= 0 || n % 8 != 0) return fail(-1, "fa_accumulate: n must be a positive multiple of 8 (got %lld)", n);
  if (dtype != FA_DTYPE_F16 && dtype != FA_DTYPE_BF16) return fail(-2, "fa_accumulate: 16-bit partials only (dtype %d)", dtype);
  if (!acc || !part) return fail(-6, "fa_accumulate: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint16_t* p16 = static_cast<const uint16_t*>(part);
  if (dtype == FA_DTYPE_BF16)
    fa::fa_accumulate_kernel<true><<<grid_for(n / 8), 256, 0, st>>>(acc, p16, n / 8, first);
  else
    fa::fa_accumulate_kernel<false><<<grid_for(n / 8), 256, 0, st>>>(acc, p16, n / 8, first);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "fa_accumulate launch");
}

int fa_round_rows(void* out, const float* in, long long n, int dtype, void* stream) {
  g_err[0] = 0;
  if (n <= 0 || n % 8 != 0) return fail(-1, "fa_round_rows: n must be a positive multiple of 8 (got %lld)", n);
  if (dtype != FA_DTYPE_F16 && dtype != FA_DTYPE_BF16) return fail(-2, "fa_round_rows: 16-bit output only (dtype %d)", dtype);
  if (!out || !in) return fail(-6, "fa_round_rows: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint16_t* o16 = static_cast<uint16_t*>(out);
  if (dtype == FA_DTYPE_BF16)
    fa::fa_round_kernel<true><<<grid_for(n / 8), 256, 0, st>>>(o16, in, n / 8);
  else
    fa::fa_round_kernel<false><<<grid_for(n / 8), 256, 0, st>>>(o16, in, n / 8);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "fa_round_rows launch");
}

}  // extern "C"
