// fa_bwd_sm100.cuh — deterministic FlashAttention-2 backward for sm_100a (16-bit inputs, fp32 accumulate).
//
// Math follows the reference backward (flash_attention_kernels.py:275-329, generalised with the softmax scale
// and causal mask of flash_attention_openai_tutorial.py:239,292,389):
//   P = exp2(S * scale*log2e - L) ; dV += P~^T dO ; dP = dO V^T ; dS = P o (dP - delta) ;
//   dK += scale * dS~^T Q ; dQ += scale * dS~ K          (P~, dS~ rounded to the input dtype, RTNE)
//
// The reference accumulates dQ across key blocks through a global spin lock (flash_attention_kernels.py:305-320)
// and is non-deterministic / broken (README.md:45-53).  Here every gradient tile has exactly one owner CTA and a
// fixed accumulation order, so results are bit-identical run to run, with no atomics and no inter-CTA waiting:
//   fa_bwd_dkdv_kernel : one CTA per 128-row key block j, loops over query blocks i, dK_j / dV_j live in TMEM;
//   fa_bwd_dq_kernel   : one CTA per 128-row query block i, loops over key blocks j, dQ_i lives in TMEM
//                        (S and dP are recomputed, precedent flash_attention_openai_tutorial.py:393-435).
//
// Both kernels share one structure.  Per (i, j) pair the two 128x128 "score" products are issued as two
// 64-column halves; softmax warpgroup a (warps 0-3) owns half a, warpgroup b (warps 4-7) half b, thread = one
// TMEM lane.  The elementwise results go back into TMEM as packed 16-bit A operands of the gradient MMAs, so the
// only shared-memory traffic is TMA -> smem -> tensor core.  The MMA warp interleaves
//   [grad a](t)  [score a](t+1)  [grad b](t)  [score b](t+1)
// so the tensor core works on one half while the other half is in the elementwise stage.
//
//   warps 0-7  elementwise (P, dS) + epilogue      warp 8  TMA producer      warp 9  MMA issuer
//   warp 10    dK/dV kernel: row statistics (-lse, -delta) of the query blocks, a ring of their own filled blocks ahead
// D = 64 variations (all bit-identical to the plain form): split elementwise stage (FA_BWD_EW_SPLIT_MASK), K_j / V_j as
// TMEM A operands of the dK/dV kernel's score products (FA_DKDV_KV_TMEM) or three rotating score slots (FA_BWD_SLOTS3),
// dS handed over through TMEM in the dQ kernel (FA_DQ_DS_TMEM).
#pragma once

#include "fa_dropout.cuh"
#include "sm100_ptx.cuh"

// Share of the exponentials computed on the FMA pipe instead of MUFU (bit i = column pair i of every 8).  Off: the
// backward kernels are bound by the TMEM round trips of their elementwise stage, not by MUFU (tools/bwd_probe.py: no gain
// at 1/4, slower at 1/2).
#ifndef FA_BWD_POLY_MASK
#define FA_BWD_POLY_MASK 0x00
#endif
// D = 64 halves the tensor work per exponential: there a quarter of them on the FMA pipe pays (round 2, tools/kernel_times.py
// on config 2: dK/dV 0.676 -> 0.660 ms, dQ 0.515 -> 0.492 ms; at D = 128 the same share costs 1-4 %)
#ifndef FA_BWD_POLY_MASK_D64
#define FA_BWD_POLY_MASK_D64 0x88
#endif

namespace fa {

struct BwdMaps {
  CUtensorMap q, k, v, dout;
};

struct BwdParams {
  const float* lse;    // (B,H,N) log2 units
  const float* delta;  // (B,H,N)
  void *dq, *dk, *dv;  // (B,H,N,D) 16-bit
  int B, H, N;
  int Nk;              // key / value rows when they differ from the N query rows (rectangular attention); 0 = N
  int64_t dq_s[3], dk_s[3], dv_s[3];  // {sB,sH,sN}
  float scale, scale_log2;
  const int* seqlens;  // per-batch valid length (key-padding mask), nullptr = N; see FwdParams::seqlens
  DropParams drop;     // dropout of the attention probabilities (kDrop instantiations only), fa_dropout.cuh
  // Arbitrary attention mask (kAmask instantiations only), one bit per entry, 1 = attend; see FwdParams::amask.  The dQ
  // kernel walks query rows of `amask` [.., query, key]; the dK/dV kernel walks key rows of `amask_t` [.., key, query],
  // the same mask transposed (made once by the caller), so both read 8 contiguous bytes per thread and half block.
  const uint8_t *amask, *amask_t;
  int64_t am_s[3], amt_s[3];  // {sB, sH, sRow} in bytes
  // optional block summary [.., query block, key block] (see FwdParams::ablock): both kernels loop over the blocks
  // flagged non-zero only (a compacted list built at CTA start), so a skipped block costs nothing; blocks flagged 2
  // (fully visible) do not load their mask bytes
  const uint8_t* ablock;
  int64_t ab_s[3];            // {sB, sH, sI} in bytes
  // band mask (kAmask instantiations with amask == nullptr): query i sees keys j with -win_left <= j - i <= win_right
  int band, win_left, win_right;
};

// bit e (0..63) of a thread's 64 mask bits held as 2 words
__device__ __forceinline__ bool amask_byte(const uint32_t (&mk)[2], int e) {
  return (mk[e >> 5] & (1u << (e & 31))) != 0u;
}
__device__ __forceinline__ void amask_load64(uint32_t (&mk)[2], const uint8_t* src, bool full = false) {
  if (full) {   // block summary says every entry is visible
    mk[0] = mk[1] = 0xffffffffu;
    return;
  }
  const uint2 v = __ldg(reinterpret_cast<const uint2*>(src));
  mk[0] = v.x, mk[1] = v.y;
}

// FA_BWD_SLOTS3: dK/dV kernel at D = 64 (with the split elementwise stage): the score halves rotate through THREE TMEM
// slots of [S^T half 64 | dP^T half 64] columns instead of two (D = 64 leaves 128 of the 512 columns free), so the score
// MMAs of half k + 3 are issued right after the gradient MMAs of half k and are long complete when the elementwise warps
// get there: the chain "elementwise(k) -> gradient MMAs(k) + score MMAs(k + 2) -> elementwise(k + 2)" of the two-slot
// pipeline is gone, the elementwise warps never wait for scores.  The Q / dO ring gets a third stage (scores run two
// blocks ahead of the gradients).  Same arithmetic in the same order: bit-identical results.
#ifndef FA_BWD_SLOTS3
#define FA_BWD_SLOTS3 1
#endif
template <int kD>
__host__ __device__ constexpr bool bwd_slots3();

template <int kD>
struct BwdCfg {
  static constexpr int kStages = (kD == 64 && FA_BWD_SLOTS3) ? 3 : 2;
  static constexpr int kTileBytes = 128 * kD * 2;
  static constexpr int kBoxBytes = 128 * 128;
  static constexpr int kBoxes = kD / 64;
  static constexpr int kStatBytes = 2 * 128 * 4;  // -lse and delta of one query block
  // dK/dV kernel: the row statistics have a ring of their own, filled by an otherwise idle warp several blocks ahead
  // (their global loads are a dependent round trip of ~1 us; behind the Q / dO ring's release they arrived late)
  static constexpr int kStatStages = 4;
  static constexpr int kThreads = 384;  // 8 elementwise warps + producer + MMA + 2 idle (complete the 3rd warpgroup)
  // stationary pair (2 tiles) + streamed pair ring (2 tiles per stage) + alignment slack
  static constexpr int kSmemDkdv = 2 * kTileBytes + kStages * 2 * kTileBytes + kStatStages * kStatBytes + 1024;
  // dQ kernel: Q_i/dO_i staging (2 tiles) + 3-stage K ring (K_j is held from its score MMAs until its dQ MMAs, one
  // block later) + 2-stage V ring
  static constexpr int kStagesK = 3, kStagesV = 2;
  static constexpr int kSmemDq = 2 * kTileBytes + (kStagesK + kStagesV) * kTileBytes + 1024;
  // TMEM columns
  static constexpr uint32_t kTmemS = 0, kTmemDP = 128, kTmemAcc0 = 256, kTmemAcc1 = 256 + kD;
  // dQ kernel only: Q_i and dO_i as TMEM-resident A operands (packed 16-bit pairs, D/2 columns each)
  static constexpr uint32_t kTmemQA = 256 + kD, kTmemDOA = 256 + kD + kD / 2;
  // dQ kernel at D = 64 (FA_DQ_DS_TMEM): packed dS of the current block, 32 columns per half, as the TMEM A operand of
  // the dQ MMAs (columns [384, 448); D = 128 has no columns left and keeps dS in shared memory)
  static constexpr uint32_t kTmemDS = 256 + 2 * kD;
};

// TMEM accumulator rows -> 16-bit global rows: thread = one row, `ncols` fp32 columns starting at taddr.
// `valid` = false: the accumulator was never written (every block of this CTA was skipped) -> zeros.
template <bool kBf16>
__device__ __forceinline__ void store_acc_rows(uint32_t taddr, int ncols, float mul, uint16_t* dst, bool in_range,
                                               bool valid = true) {
  for (int c = 0; c < ncols; c += 32) {
    uint32_t r[32];
    tmem_ld_x32(taddr + c, r);
    tc_wait_ld();
    if (!valid) {
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = 0u;
    }
    if (in_range) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 v;
        v.x = pack2<kBf16>(__uint_as_float(r[8 * i + 0]) * mul, __uint_as_float(r[8 * i + 1]) * mul);
        v.y = pack2<kBf16>(__uint_as_float(r[8 * i + 2]) * mul, __uint_as_float(r[8 * i + 3]) * mul);
        v.z = pack2<kBf16>(__uint_as_float(r[8 * i + 4]) * mul, __uint_as_float(r[8 * i + 5]) * mul);
        v.w = pack2<kBf16>(__uint_as_float(r[8 * i + 6]) * mul, __uint_as_float(r[8 * i + 7]) * mul);
        *reinterpret_cast<uint4*>(dst + c + i * 8) = v;
      }
    }
  }
}

// Elementwise stage of one 64-column half (thread = one TMEM lane): P = exp2(S*sl2 - L), dS = P o (dP - delta),
// both rounded to 16 bits and written back over S / dP as packed TMEM A operands.  Two fp32 lanes per issue slot
// (FFMA2 / FADD2 / FMUL2).  kColStats: -L and -delta vary along the columns and come from shared memory (dK/dV
// kernel, transposed scores); otherwise they are per-thread constants (dQ kernel).  kMask: causal diagonal block.
// kStoreP: the dK/dV kernel needs P^T (for dV); the dQ kernel only needs dS, stored over S.
// kDrop: dropout — the stored P is keep o P (its 1 / (1 - p) goes into the dV epilogue) and dS = P o (keep * rp * dP - delta);
// `drop_word` = key + word index of this thread's first element pair, `drop_shift` / kDropSecond as in fa_dropout.cuh,
// consecutive pairs are kDropStep words apart.
// kAmask: `mk` = this thread's 64 attention-mask bytes for the half block (non-zero = attend).
// kBand (with kAmask): no mask bytes, the visible elements of this thread's 64 are [band_lo, band_hi].
template <bool kBf16, bool kColStats, bool kMask, bool kTransposed, bool kStoreP, bool kDrop = false, bool kAmask = false,
          bool kBand = false, int kPoly = FA_BWD_POLY_MASK>
__device__ __forceinline__ void bwd_elementwise_half(uint32_t tS, uint32_t tDP, uint32_t st_saddr, uint64_t nl_c,
                                                     uint64_t nd_c, float sl2, int row, int col0,
                                                     uint32_t drop_word, uint32_t drop_shift, uint32_t drop_thresh,
                                                     float drop_rp, const uint32_t (&mk)[2], int band_lo = 0,
                                                     int band_hi = 0) {
  constexpr uint32_t kDropStep = kTransposed ? (1u << 15) : 1u;
  constexpr int kDropSecond = kTransposed ? 16 : 8;
#if FA_ABLATE == 3
  return;
#endif
  uint32_t sr[64], dr[64];
  tmem_ld_x32(tS, *reinterpret_cast<uint32_t(*)[32]>(&sr[0]));
  tmem_ld_x32(tDP, *reinterpret_cast<uint32_t(*)[32]>(&dr[0]));
  tmem_ld_x32(tS + 32, *reinterpret_cast<uint32_t(*)[32]>(&sr[32]));
  tmem_ld_x32(tDP + 32, *reinterpret_cast<uint32_t(*)[32]>(&dr[32]));
  tc_wait_ld();
  const uint64_t sl2_2 = f32x2_pack(sl2, sl2);
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t pp[16], pd[16];
#pragma unroll
    for (int g4 = 0; g4 < 8; ++g4) {
      uint64_t nl4[2] = {nl_c, nl_c}, nd4[2] = {nd_c, nd_c};
      if constexpr (kColStats) {
        lds_f32x2x2(st_saddr + (c * 32 + g4 * 4) * 4, nl4[0], nl4[1]);
        lds_f32x2x2(st_saddr + (128 + c * 32 + g4 * 4) * 4, nd4[0], nd4[1]);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int g = g4 * 2 + u;
        const int e = c * 32 + g * 2;
#if FA_ABLATE == 2
        if constexpr (kStoreP) pp[g] = sr[e] ^ sr[e + 1];
        pd[g] = dr[e] ^ dr[e + 1];
        continue;
#endif
        float x0, x1;
        f32x2_unpack(f32x2_fma(f32x2_pack_bits(sr[e], sr[e + 1]), sl2_2, nl4[u]), x0, x1);
        float p0, p1;
        if ((kPoly >> (g & 7)) & 1) {   // FMA-pipe exp2 for a share of the pairs (see fa_fwd_sm100.cuh)
          ex2_poly_x2(x0, x1, p0, p1);
        } else {
          p0 = ex2_approx(x0), p1 = ex2_approx(x1);
        }
        if constexpr (kMask) {
          // keep key <= query.  Transposed scores: row = key, column = query; otherwise row = query, column = key.
          const int c0 = col0 + e;
          if (kTransposed ? (row > c0) : (c0 > row)) p0 = 0.f;
          if (kTransposed ? (row > c0 + 1) : (c0 + 1 > row)) p1 = 0.f;
        }
        if constexpr (kAmask && kBand) {
          if (e < band_lo || e > band_hi) p0 = 0.f;
          if (e + 1 < band_lo || e + 1 > band_hi) p1 = 0.f;
        } else if constexpr (kAmask) {
          if (!amask_byte(mk, e)) p0 = 0.f;
          if (!amask_byte(mk, e + 1)) p1 = 0.f;
        }
        float d0, d1;
        if constexpr (kDrop) {
          bool keep0, keep1;
          drop_keep_pair<kDropSecond>(drop_word + (uint32_t)(e >> 1) * kDropStep, drop_shift, drop_thresh, keep0, keep1);
          const uint64_t f2 = f32x2_pack(keep0 ? drop_rp : 0.f, keep1 ? drop_rp : 0.f);
          f32x2_unpack(f32x2_mul(f32x2_pack(p0, p1), f32x2_fma(f32x2_pack_bits(dr[e], dr[e + 1]), f2, nd4[u])), d0, d1);
          if constexpr (kStoreP) pp[g] = pack2<kBf16>(keep0 ? p0 : 0.f, keep1 ? p1 : 0.f);
        } else {
          f32x2_unpack(f32x2_mul(f32x2_pack(p0, p1), f32x2_add(f32x2_pack_bits(dr[e], dr[e + 1]), nd4[u])), d0, d1);
          if constexpr (kStoreP) pp[g] = pack2<kBf16>(p0, p1);
        }
        pd[g] = pack2<kBf16>(d0, d1);
      }
    }
    if constexpr (kStoreP) {
      tmem_st_x16(tS + c * 16, pp);
      tmem_st_x16(tDP + c * 16, pd);
    } else {
      tmem_st_x16(tS + c * 16, pd);
    }
  }
}

// dQ kernel flavour: per-thread statistics, dS only.  The two score accumulators are copied to registers first and
// released to the MMA warp (`sc_free`) before any arithmetic, so the next block's score MMAs overlap this stage.
// Output: 64 values of this thread's row as 32 packed 16-bit pairs in `pd`.
template <bool kBf16, bool kMask, bool kDrop = false, bool kAmask = false, bool kBand = false, int kPoly = FA_BWD_POLY_MASK>
__device__ __forceinline__ void dq_elementwise_half(uint32_t tS, uint32_t tDP, uint64_t* sc_free_bar, uint64_t nl,
                                                    uint64_t nd, float sl2, int row, int col0, uint32_t (&pd)[32],
                                                    uint32_t drop_word, uint32_t drop_shift, uint32_t drop_thresh,
                                                    float drop_rp, const uint32_t (&mk)[2], int band_lo = 0,
                                                    int band_hi = 0) {
  uint32_t sr[64], dr[64];
  tmem_ld_x32(tS, *reinterpret_cast<uint32_t(*)[32]>(&sr[0]));
  tmem_ld_x32(tDP, *reinterpret_cast<uint32_t(*)[32]>(&dr[0]));
  tmem_ld_x32(tS + 32, *reinterpret_cast<uint32_t(*)[32]>(&sr[32]));
  tmem_ld_x32(tDP + 32, *reinterpret_cast<uint32_t(*)[32]>(&dr[32]));
  tc_wait_ld();
  tc_fence_before();
  mbar_arrive(sc_free_bar);
#if FA_ABLATE == 3
#pragma unroll
  for (int g = 0; g < 32; ++g) pd[g] = 0;
  return;
#endif
  const uint64_t sl2_2 = f32x2_pack(sl2, sl2);
#pragma unroll
  for (int g = 0; g < 32; ++g) {
    const int e = g * 2;
#if FA_ABLATE == 2
    pd[g] = sr[e] ^ dr[e + 1];
    continue;
#endif
    float x0, x1;
    f32x2_unpack(f32x2_fma(f32x2_pack_bits(sr[e], sr[e + 1]), sl2_2, nl), x0, x1);
    float p0, p1;
    if ((kPoly >> (g & 7)) & 1) {
      ex2_poly_x2(x0, x1, p0, p1);
    } else {
      p0 = ex2_approx(x0), p1 = ex2_approx(x1);
    }
    if constexpr (kMask) {   // causal diagonal block: keep key <= query (row = query, column = key)
      const int c0 = col0 + e;
      if (c0 > row) p0 = 0.f;
      if (c0 + 1 > row) p1 = 0.f;
    }
    if constexpr (kAmask && kBand) {
      if (e < band_lo || e > band_hi) p0 = 0.f;
      if (e + 1 < band_lo || e + 1 > band_hi) p1 = 0.f;
    } else if constexpr (kAmask) {
      if (!amask_byte(mk, e)) p0 = 0.f;
      if (!amask_byte(mk, e + 1)) p1 = 0.f;
    }
    float d0, d1;
    if constexpr (kDrop) {   // dS = P o (keep * rp * dP - delta); row walk, one word per key pair
      bool keep0, keep1;
      drop_keep_pair<8>(drop_word + (uint32_t)g, drop_shift, drop_thresh, keep0, keep1);
      const uint64_t f2 = f32x2_pack(keep0 ? drop_rp : 0.f, keep1 ? drop_rp : 0.f);
      f32x2_unpack(f32x2_mul(f32x2_pack(p0, p1), f32x2_fma(f32x2_pack_bits(dr[e], dr[e + 1]), f2, nd)), d0, d1);
    } else {
      f32x2_unpack(f32x2_mul(f32x2_pack(p0, p1), f32x2_add(f32x2_pack_bits(dr[e], dr[e + 1]), nd)), d0, d1);
    }
    pd[g] = pack2<kBf16>(d0, d1);
  }
}

// ---- two-stage elementwise of the dK/dV kernel (FA_DKDV_TWO_STAGE, one warpgroup per half, no dropout; off) ----
// The P stage needs only S^T, so it can start as soon as the S^T MMAs of the half have completed (their own commit) with
// the dP^T MMAs running under it; the dS stage then picks dP^T up.  Same arithmetic as bwd_elementwise_half, P kept in
// fp32 registers between the stages: bit-identical results (all GPU tests green with it).  Measured on config 3
// (profiles/r02_dkdv_two_stage.txt): 1.70-1.72 ms against 1.66-1.68 — taking the dP^T MMAs (8 x 48 clk) off a half's
// chain buys less than the second TMEM round trip and the lost MUFU / FMA overlap inside a thread cost.  Default 0.
#ifndef FA_DKDV_TWO_STAGE
#define FA_DKDV_TWO_STAGE 0
#endif
template <bool kBf16, bool kMask, bool kAmask, bool kBand, int kPoly>
__device__ __forceinline__ void dkdv_p_stage(uint32_t tS, uint32_t st_saddr, float sl2, int row, int col0,
                                             const uint32_t (&mk)[2], int band_lo, int band_hi, uint32_t (&pf)[64]) {
  tmem_ld_x32(tS, *reinterpret_cast<uint32_t(*)[32]>(&pf[0]));
  tmem_ld_x32(tS + 32, *reinterpret_cast<uint32_t(*)[32]>(&pf[32]));
  tc_wait_ld();
  const uint64_t sl2_2 = f32x2_pack(sl2, sl2);
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t pp[16];
#pragma unroll
    for (int g4 = 0; g4 < 8; ++g4) {
      uint64_t nl4[2];
      lds_f32x2x2(st_saddr + (c * 32 + g4 * 4) * 4, nl4[0], nl4[1]);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int g = g4 * 2 + u;
        const int e = c * 32 + g * 2;
        float x0, x1;
        f32x2_unpack(f32x2_fma(f32x2_pack_bits(pf[e], pf[e + 1]), sl2_2, nl4[u]), x0, x1);
        float p0, p1;
        if ((kPoly >> (g & 7)) & 1) {
          ex2_poly_x2(x0, x1, p0, p1);
        } else {
          p0 = ex2_approx(x0), p1 = ex2_approx(x1);
        }
        if constexpr (kMask) {   // transposed scores: row = key, column = query; keep key <= query
          const int c0 = col0 + e;
          if (row > c0) p0 = 0.f;
          if (row > c0 + 1) p1 = 0.f;
        }
        if constexpr (kAmask && kBand) {
          if (e < band_lo || e > band_hi) p0 = 0.f;
          if (e + 1 < band_lo || e + 1 > band_hi) p1 = 0.f;
        } else if constexpr (kAmask) {
          if (!amask_byte(mk, e)) p0 = 0.f;
          if (!amask_byte(mk, e + 1)) p1 = 0.f;
        }
        pf[e] = __float_as_uint(p0);
        pf[e + 1] = __float_as_uint(p1);
        pp[g] = pack2<kBf16>(p0, p1);
      }
    }
    tmem_st_x16(tS + c * 16, pp);
  }
}
template <bool kBf16>
__device__ __forceinline__ void dkdv_ds_stage(uint32_t tDP, uint32_t st_saddr, const uint32_t (&pf)[64]) {
  uint32_t dr[64];
  tmem_ld_x32(tDP, *reinterpret_cast<uint32_t(*)[32]>(&dr[0]));
  tmem_ld_x32(tDP + 32, *reinterpret_cast<uint32_t(*)[32]>(&dr[32]));
  tc_wait_ld();
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t pd[16];
#pragma unroll
    for (int g4 = 0; g4 < 8; ++g4) {
      uint64_t nd4[2];
      lds_f32x2x2(st_saddr + (128 + c * 32 + g4 * 4) * 4, nd4[0], nd4[1]);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int g = g4 * 2 + u;
        const int e = c * 32 + g * 2;
        float d0, d1;
        f32x2_unpack(f32x2_mul(f32x2_pack_bits(pf[e], pf[e + 1]), f32x2_add(f32x2_pack_bits(dr[e], dr[e + 1]), nd4[u])),
                     d0, d1);
        pd[g] = pack2<kBf16>(d0, d1);
      }
    }
    tmem_st_x16(tDP + c * 16, pd);
  }
}

// ---- split elementwise stage: all eight warps work on ONE 64-column half at a time ----
// With the two warps of a sub-partition sharing a half (32 columns each: warp w and w + 4 hold the same TMEM lanes) the
// elementwise link of a half's chain — score MMAs -> elementwise -> gradient MMAs — gets shorter, and the other half's MMAs
// fill the tensor pipe meanwhile.  The arithmetic per element is unchanged, so the results are bit-identical to the
// one-warpgroup-per-half layout.  Measured (tools/kernel_times.py, profiles/r02_bwd_split.txt): D = 64, where the tensor
// work per exponential is half and the kernels wait on the elementwise stage, gains 4-5 % (dK/dV 0.683 -> 0.657 ms, dQ
// 0.548 -> 0.523 on config 2); D = 128 does not (dK/dV 1.68 -> 1.67, dQ 1.32 -> 1.36-1.41 ms on config 3: there a chunk of
// 32 columns takes two warps as long, ~930 clk, as 64 columns take one warp alone, and shared-memory bandwidth — 2800 of
// the 3340 clk of a dK/dV block — bounds the MMAs).  Hence: D = 64 only (bit k of FA_BWD_EW_SPLIT_MASK = head size 64 << k).
#ifndef FA_BWD_EW_SPLIT_MASK
#define FA_BWD_EW_SPLIT_MASK 0x1
#endif
template <int kD>
__host__ __device__ constexpr bool bwd_ew_split() { return ((FA_BWD_EW_SPLIT_MASK >> (kD == 64 ? 0 : 1)) & 1) != 0; }
// FA_DQ_DS_TMEM: dQ kernel at D = 64 with the split elementwise stage: dS goes back to TMEM (as in the dK/dV kernel)
// instead of shared memory: no st.shared + fence.proxy.async in the hand-over (16 % of the elementwise warps' samples
// in ncu), and dQ += dS K_j becomes a TS product (an SS product with N = 64 is shared-memory bound: 48 clk instead of 32).
#ifndef FA_DQ_DS_TMEM
#define FA_DQ_DS_TMEM 1
#endif
// FA_DQ_FULL_SCORE (dQ kernel at D = 128, experiment, off): the score products issued as full 128-column SS products (an
// SS product with N = 128 runs at the tensor peak, so Q_i / dO_i need no TMEM copies) — that frees 128 TMEM columns, 64 of
// which take the packed dS of the block: dQ += dS K_j becomes a TS product and the hand-over loses its st.shared +
// fence.proxy.async (10 % of the elementwise warps' ncu samples at D = 128).  Bit-identical (all GPU tests green), but
// 1.48-1.58 ms against 1.19-1.22 on config 3 (profiles/r02_dq_full_score.txt): with one 1024-clock score product per block
// both warpgroups get their scores at the same moment and the tensor pipe has no other half to work on meanwhile.
#ifndef FA_DQ_FULL_SCORE
#define FA_DQ_FULL_SCORE 0
#endif
template <int kD>
__host__ __device__ constexpr bool dq_full_score() { return kD == 128 && FA_DQ_FULL_SCORE != 0 && !bwd_ew_split<kD>(); }
template <int kD>
__host__ __device__ constexpr bool dq_ds_tmem() { return kD == 64 && FA_DQ_DS_TMEM != 0 && bwd_ew_split<kD>(); }
// FA_DKDV_KV_TMEM (takes the place of the third slot: 2 x 128 + 128 + 64 columns): dK/dV kernel at D = 64 with K_j / V_j
// copied once into TMEM (columns 384-447) as the A operands of the score products — TS instead of shared-memory-bound SS
// products with N = 64 (32 clk per instruction instead of 48), as the dQ kernel does with Q_i / dO_i.  With the statistics
// ring in place (tools/kernel_times.py, config 2, same box, profiles/r02_dkdv_d64_kv_tmem.txt): two slots 0.599 ms, three
// slots 0.511, two slots + K / V in TMEM 0.497 — both remove about the same wait, the TS products also halve the kernel's
// shared-memory operand traffic; all three do not fit (576 columns).  Bit-identical.
#ifndef FA_DKDV_KV_TMEM
#define FA_DKDV_KV_TMEM 1
#endif
template <int kD>
__host__ __device__ constexpr bool dkdv_kv_tmem() { return kD == 64 && FA_DKDV_KV_TMEM != 0 && bwd_ew_split<kD>(); }
template <int kD>
__host__ __device__ constexpr bool bwd_slots3() {
  return kD == 64 && FA_BWD_SLOTS3 != 0 && bwd_ew_split<kD>() && !dkdv_kv_tmem<kD>();
}

// One 32-column chunk of a half (dK/dV kernel; thread = one TMEM lane).  Same arithmetic as bwd_elementwise_half.
// tS / tDP: the chunk's 32 fp32 columns; the packed results go to the first 16 columns of the chunk's OWN score columns
// (the partner warp's columns are never written, so the two need no synchronisation).  `cbase` = first column of the
// chunk inside the half (0 / 32): statistics, causal / band limits, dropout words and mask bits are indexed with it.
template <bool kBf16, bool kMask, bool kDrop, bool kAmask, bool kBand, int kPoly>
__device__ __forceinline__ void dkdv_elementwise_chunk(uint32_t tS, uint32_t tDP, uint32_t st_saddr, float sl2, int row,
                                                       int col0, int cbase, uint32_t drop_word, uint32_t drop_shift,
                                                       uint32_t drop_thresh, float drop_rp, uint32_t mkw, int band_lo,
                                                       int band_hi) {
  uint32_t sr[32], dr[32];
  tmem_ld_x32(tS, sr);
  tmem_ld_x32(tDP, dr);
  tc_wait_ld();
  const uint64_t sl2_2 = f32x2_pack(sl2, sl2);
  uint32_t pp[16], pd[16];
#pragma unroll
  for (int g4 = 0; g4 < 8; ++g4) {
    uint64_t nl4[2], nd4[2];
    lds_f32x2x2(st_saddr + (cbase + g4 * 4) * 4, nl4[0], nl4[1]);
    lds_f32x2x2(st_saddr + (128 + cbase + g4 * 4) * 4, nd4[0], nd4[1]);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int g = g4 * 2 + u;
      const int e = g * 2;              // column inside the chunk
      const int eh = cbase + e;         // column inside the half
      float x0, x1;
      f32x2_unpack(f32x2_fma(f32x2_pack_bits(sr[e], sr[e + 1]), sl2_2, nl4[u]), x0, x1);
      float p0, p1;
      if ((kPoly >> (g & 7)) & 1) {
        ex2_poly_x2(x0, x1, p0, p1);
      } else {
        p0 = ex2_approx(x0), p1 = ex2_approx(x1);
      }
      if constexpr (kMask) {   // transposed scores: row = key, column = query; keep key <= query
        const int c0 = col0 + eh;
        if (row > c0) p0 = 0.f;
        if (row > c0 + 1) p1 = 0.f;
      }
      if constexpr (kAmask && kBand) {
        if (eh < band_lo || eh > band_hi) p0 = 0.f;
        if (eh + 1 < band_lo || eh + 1 > band_hi) p1 = 0.f;
      } else if constexpr (kAmask) {
        if (!(mkw & (1u << e))) p0 = 0.f;
        if (!(mkw & (2u << e))) p1 = 0.f;
      }
      float d0, d1;
      if constexpr (kDrop) {
        bool keep0, keep1;
        drop_keep_pair<16>(drop_word + (uint32_t)(eh >> 1) * (1u << 15), drop_shift, drop_thresh, keep0, keep1);
        const uint64_t f2 = f32x2_pack(keep0 ? drop_rp : 0.f, keep1 ? drop_rp : 0.f);
        f32x2_unpack(f32x2_mul(f32x2_pack(p0, p1), f32x2_fma(f32x2_pack_bits(dr[e], dr[e + 1]), f2, nd4[u])), d0, d1);
        pp[g] = pack2<kBf16>(keep0 ? p0 : 0.f, keep1 ? p1 : 0.f);
      } else {
        f32x2_unpack(f32x2_mul(f32x2_pack(p0, p1), f32x2_add(f32x2_pack_bits(dr[e], dr[e + 1]), nd4[u])), d0, d1);
        pp[g] = pack2<kBf16>(p0, p1);
      }
      pd[g] = pack2<kBf16>(d0, d1);
    }
  }
  tmem_st_x16(tS, pp);
  tmem_st_x16(tDP, pd);
}

// One 32-column chunk of a half (dQ kernel): per-thread statistics, dS only, 16 packed pairs in `pd`.
template <bool kBf16, bool kMask, bool kDrop, bool kAmask, bool kBand, int kPoly>
__device__ __forceinline__ void dq_elementwise_chunk(uint32_t tS, uint32_t tDP, uint64_t* sc_free_bar, uint64_t nl,
                                                     uint64_t nd, float sl2, int row, int col0, int cbase,
                                                     uint32_t (&pd)[16], uint32_t drop_word, uint32_t drop_shift,
                                                     uint32_t drop_thresh, float drop_rp, uint32_t mkw, int band_lo,
                                                     int band_hi) {
  uint32_t sr[32], dr[32];
  tmem_ld_x32(tS, sr);
  tmem_ld_x32(tDP, dr);
  tc_wait_ld();
  tc_fence_before();
  mbar_arrive(sc_free_bar);
  const uint64_t sl2_2 = f32x2_pack(sl2, sl2);
#pragma unroll
  for (int g = 0; g < 16; ++g) {
    const int e = g * 2, eh = cbase + e;
    float x0, x1;
    f32x2_unpack(f32x2_fma(f32x2_pack_bits(sr[e], sr[e + 1]), sl2_2, nl), x0, x1);
    float p0, p1;
    if ((kPoly >> (g & 7)) & 1) {
      ex2_poly_x2(x0, x1, p0, p1);
    } else {
      p0 = ex2_approx(x0), p1 = ex2_approx(x1);
    }
    if constexpr (kMask) {   // keep key <= `row` (row = query, column = key)
      const int c0 = col0 + eh;
      if (c0 > row) p0 = 0.f;
      if (c0 + 1 > row) p1 = 0.f;
    }
    if constexpr (kAmask && kBand) {
      if (eh < band_lo || eh > band_hi) p0 = 0.f;
      if (eh + 1 < band_lo || eh + 1 > band_hi) p1 = 0.f;
    } else if constexpr (kAmask) {
      if (!(mkw & (1u << e))) p0 = 0.f;
      if (!(mkw & (2u << e))) p1 = 0.f;
    }
    float d0, d1;
    if constexpr (kDrop) {
      bool keep0, keep1;
      drop_keep_pair<8>(drop_word + (uint32_t)(eh >> 1), drop_shift, drop_thresh, keep0, keep1);
      const uint64_t f2 = f32x2_pack(keep0 ? drop_rp : 0.f, keep1 ? drop_rp : 0.f);
      f32x2_unpack(f32x2_mul(f32x2_pack(p0, p1), f32x2_fma(f32x2_pack_bits(dr[e], dr[e + 1]), f2, nd)), d0, d1);
    } else {
      f32x2_unpack(f32x2_mul(f32x2_pack(p0, p1), f32x2_add(f32x2_pack_bits(dr[e], dr[e + 1]), nd)), d0, d1);
    }
    pd[g] = pack2<kBf16>(d0, d1);
  }
}

// ================================================================================================ dK / dV
template <bool kBf16, int kD, bool kCausal, bool kDrop = false, bool kAmask = false>
__global__ void __launch_bounds__(384, 1)
fa_bwd_dkdv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                   const BwdParams p) {
  using Cfg = BwdCfg<kD>;
  constexpr int NS = Cfg::kStages;
  constexpr bool kSplit = bwd_ew_split<kD>();
  constexpr bool kSlots3 = bwd_slots3<kD>();
  constexpr int kSlots = kSlots3 ? 3 : 2;
  constexpr bool kTwoStage = FA_DKDV_TWO_STAGE != 0 && !kSplit && !kDrop;
  constexpr bool kKvTmem = dkdv_kv_tmem<kD>();
  constexpr uint32_t kTmemKA = 384, kTmemVA = 416;   // kKvTmem: K_j / V_j as packed TMEM A operands (32 columns each)
  // TMEM: two slots = halves a / b at S [0,128) and dP [128,256); three slots = [S half | dP half] at 0 / 128 / 256
  constexpr uint32_t kAccV = kSlots3 ? 384 : Cfg::kTmemAcc0, kAccK = kSlots3 ? 448 : Cfg::kTmemAcc1;
  auto slot_s = [](int slot) -> uint32_t { return kSlots3 ? slot * 128 : Cfg::kTmemS + slot * 64; };
  auto slot_dp = [](int slot) -> uint32_t { return kSlots3 ? slot * 128 + 64 : Cfg::kTmemDP + slot * 64; };
  // (attention masks need exact zeros: the polynomial clamps at 2^-125, so masked variants keep MUFU)
  constexpr int kPolyMask = kAmask ? 0 : (kD == 64 ? FA_BWD_POLY_MASK_D64 : FA_BWD_POLY_MASK);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;                                   // stationary K_j
  uint8_t* sV = sK + Cfg::kTileBytes;                   // stationary V_j
  uint8_t* sQ = sV + Cfg::kTileBytes;                   // [NS] streamed Q_i
  uint8_t* sDO = sQ + NS * Cfg::kTileBytes;             // [NS] streamed dO_i
  constexpr int NT = Cfg::kStatStages;
  float* sStat = reinterpret_cast<float*>(sDO + NS * Cfg::kTileBytes);  // [NT][2][128]: -lse, -delta

  __shared__ uint64_t kv_full, acc_full;
  __shared__ uint64_t in_full[NS], in_empty[NS], stat_full[NT], stat_empty[NT];
  __shared__ uint64_t sc_full[kSlots], p_full[kSlots];
  __shared__ uint64_t s_full[2];   // two-stage elementwise: the S^T MMAs of a half have completed
  __shared__ uint64_t kv_tmem;     // kKvTmem: K_j / V_j are in TMEM
  __shared__ uint32_t tmem_base_s;
  // kAmask with a block summary: the query blocks with something visible for this key block, in order
  __shared__ uint16_t s_list[kAmask ? 512 : 2];
  __shared__ int s_nact;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int jb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int k0 = jb * 128;
  const int nv = p.seqlens ? min(max(p.seqlens[b], 0), p.N) : p.N;   // valid length of this batch element (queries)
  const int nk = p.Nk > 0 ? p.Nk : nv;                               // valid key rows
  if (k0 >= nk) return;   // padded key block: dK / dV rows stay as the caller initialised them
  // Padded keys inside the last block need no mask here: rows = keys, and a padded key only feeds its own dK / dV rows,
  // which are never stored.  Padded queries get -L = -inf below, i.e. P = 0.
  const int n_q_total = (nv + 127) >> 7;
  const int i_begin = kCausal ? jb : 0;
  const int n_it = n_q_total - i_begin;
  const bool use_list = kAmask && p.ablock != nullptr && n_it <= 512;
  if constexpr (kAmask) {
    if (use_list && warp == 0) {   // one warp compacts the flags (ballot + prefix count), 32 blocks per step
      const uint8_t* ab = p.ablock + (int64_t)b * p.ab_s[0] + (int64_t)h * p.ab_s[1] + jb;
      int c = 0;
      for (int base = 0; base < n_it; base += 32) {
        const int it = base + lane;
        const bool on = it < n_it && ab[(int64_t)(i_begin + it) * p.ab_s[2]] != 0;
        const uint32_t m = __ballot_sync(0xffffffffu, on);
        if (on) s_list[c + __popc(m & ((1u << lane) - 1u))] = (uint16_t)it;
        c += __popc(m);
      }
      if (lane == 0) s_nact = c;
    }
  }

  if (threadIdx.x == 0) {
    mbar_init(&kv_full, 1);
    mbar_init(&acc_full, 1);
    for (int s = 0; s < NS; ++s) {
      mbar_init(&in_full[s], 1);
      mbar_init(&in_empty[s], 1);
    }
    for (int s = 0; s < NT; ++s) {
      mbar_init(&stat_full[s], 32);
      mbar_init(&stat_empty[s], 8);   // one arrival per elementwise warp
    }
    for (int t = 0; t < kSlots; ++t) {
      mbar_init(&sc_full[t], 1);
      mbar_init(&p_full[t], kSplit ? 256 : 128);
    }
    for (int t = 0; t < 2; ++t) mbar_init(&s_full[t], 1);
    mbar_init(&kv_tmem, 256);
    fence_mbar_init();
  }
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmDO);
  }
  if (warp == 9) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  // every role walks the same list: step k handles query block i_begin + block_of(k); ring stages and barrier phases
  // are functions of k, so the pipeline protocol is exactly the dense one over a shorter sequence
  const int n_loop = use_list ? s_nact : n_it;
  auto block_of = [&](int k) -> int { return use_list ? (int)s_list[k] : k; };

  if (warp >= 8) {
  setmaxnreg_dec<72>();
  if (warp == 8) {
    // ------------------------------------------------------------------ producer: TMA + row statistics
    if (lane == 0) {
      mbar_arrive_expect_tx(&kv_full, 2 * Cfg::kTileBytes);
      for (int bx = 0; bx < Cfg::kBoxes; ++bx) {
        tma_load_4d(sK + bx * Cfg::kBoxBytes, &tmK, &kv_full, bx * 64, k0, h, b);
        tma_load_4d(sV + bx * Cfg::kBoxBytes, &tmV, &kv_full, bx * 64, k0, h, b);
      }
    }
    for (int k = 0; k < n_loop; ++k) {
      const int it = kAmask ? block_of(k) : k;
      const int s = k % NS;
      const uint32_t ph = (k / NS) & 1;
      const int q0 = (i_begin + it) * 128;
      mbar_wait(&in_empty[s], ph ^ 1);
      if (lane == 0) {
        mbar_arrive_expect_tx(&in_full[s], 2 * Cfg::kTileBytes);
        for (int bx = 0; bx < Cfg::kBoxes; ++bx) {
          tma_load_4d(sQ + s * Cfg::kTileBytes + bx * Cfg::kBoxBytes, &tmQ, &in_full[s], bx * 64, q0, h, b);
          tma_load_4d(sDO + s * Cfg::kTileBytes + bx * Cfg::kBoxBytes, &tmDO, &in_full[s], bx * 64, q0, h, b);
        }
      }
    }
  } else if (warp == 10) {
    // ------------------------------------------------------------------ row statistics of the query blocks
    // -lse and -delta of query block k into stage k % NT, up to NT blocks ahead of the elementwise warps; all eight
    // loads of a lane are issued before the first store (a store to shared memory through a generic pointer keeps
    // ptxas from moving later loads above it: four dependent round trips per block instead of one)
    const float* lsep = p.lse + ((int64_t)b * p.H + h) * p.N;
    const float* dlp = p.delta + ((int64_t)b * p.H + h) * p.N;
    for (int k = 0; k < n_loop; ++k) {
      const int it = kAmask ? block_of(k) : k;
      const int s = k % NT;
      const int q0 = (i_begin + it) * 128;
      float nl[4], nd[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = q0 + lane * 4 + e;
        const bool ok = r < nv;
        nl[e] = ok ? -__ldg(lsep + r) : -INFINITY;   // rows past the valid length: P = exp2(-inf) = 0
        nd[e] = ok ? -__ldg(dlp + r) : 0.f;
        if (kAmask && nl[e] == INFINITY) nl[e] = -INFINITY;   // a query that saw no key (L = -inf): P = 0 as well
      }
      mbar_wait(&stat_empty[s], ((k / NT) & 1) ^ 1);
      float4* st = reinterpret_cast<float4*>(sStat + s * 256);
      st[lane] = make_float4(nl[0], nl[1], nl[2], nl[3]);
      st[32 + lane] = make_float4(nd[0], nd[1], nd[2], nd[3]);
      mbar_arrive(&stat_full[s]);
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_sc = umma_idesc_f16(kBf16, 128, 64, 0, 0);   // [128 kv] x [64 q], K = D
      constexpr uint32_t idesc_gr = umma_idesc_f16(kBf16, 128, kD, 0, 1);   // [128 kv] x [D],   K = 64 q
      constexpr uint32_t kTileLo = Cfg::kTileBytes >> 4, kHalfLo = 8192u >> 4;
      const uint32_t k_lo = umma_lo_kmajor(smem_u32(sK)), v_lo = umma_lo_kmajor(smem_u32(sV));
      const uint32_t q_lo = umma_lo_kmajor(smem_u32(sQ)), do_lo = umma_lo_kmajor(smem_u32(sDO));
      const uint32_t q_mn = umma_lo_mnmajor(smem_u32(sQ), Cfg::kBoxBytes);
      const uint32_t do_mn = umma_lo_mnmajor(smem_u32(sDO), Cfg::kBoxBytes);

      // S^T half = K_j Q_i[half]^T ; dP^T half = V_j dO_i[half]^T  (into TMEM slot `slot`; two slots: slot == half)
      auto issue_score = [&](int half, int s, int slot) {
        const uint32_t bq = q_lo + s * kTileLo + half * kHalfLo, bdo = do_lo + s * kTileLo + half * kHalfLo;
        const uint32_t dS = tmem + slot_s(slot), dDP = tmem + slot_dp(slot);
        static_for<0, kD / 16>([&](auto kc) {
          constexpr int k = decltype(kc)::value;
          constexpr uint32_t off = umma_koff_kmajor(k, Cfg::kBoxBytes);
          if constexpr (kKvTmem) umma_ts_off<k * 8, off>(dS, tmem + kTmemKA, bq, idesc_sc, k > 0);
          else umma_ss_off<off, off>(dS, k_lo, bq, idesc_sc, k > 0);
        });
        if constexpr (kTwoStage) tc_commit(&s_full[slot]);
        static_for<0, kD / 16>([&](auto kc) {
          constexpr int k = decltype(kc)::value;
          constexpr uint32_t off = umma_koff_kmajor(k, Cfg::kBoxBytes);
          if constexpr (kKvTmem) umma_ts_off<k * 8, off>(dDP, tmem + kTmemVA, bdo, idesc_sc, k > 0);
          else umma_ss_off<off, off>(dDP, v_lo, bdo, idesc_sc, k > 0);
        });
        tc_commit(&sc_full[slot]);
      };
      // dV += P^T[half] dO_i[half] ; dK += dS^T[half] Q_i[half]
      auto issue_grad = [&](int half, int s, int slot, bool first) {
        const uint32_t bdo = do_mn + s * kTileLo + half * umma_koff_mnmajor(4);
        const uint32_t bq = q_mn + s * kTileLo + half * umma_koff_mnmajor(4);
        const uint32_t aP = tmem + slot_s(slot), aDS = tmem + slot_dp(slot);
        const uint32_t dV_t = tmem + kAccV, dK_t = tmem + kAccK;
        // (split elementwise stage: the packed pairs of queries 32-63 of the half start at column 32, not 16)
        static_for<0, 4>([&](auto kc) {
          constexpr int k = decltype(kc)::value;
          constexpr uint32_t acol = kSplit ? (k * 8 + (k >= 2 ? 16 : 0)) : k * 8;
          umma_ts_off<acol, umma_koff_mnmajor(k)>(dV_t, aP, bdo, idesc_gr, !(first && k == 0));
        });
        static_for<0, 4>([&](auto kc) {
          constexpr int k = decltype(kc)::value;
          constexpr uint32_t acol = kSplit ? (k * 8 + (k >= 2 ? 16 : 0)) : k * 8;
          umma_ts_off<acol, umma_koff_mnmajor(k)>(dK_t, aDS, bq, idesc_gr, !(first && k == 0));
        });
      };

      mbar_wait(&kv_full, 0);
      if constexpr (kKvTmem) mbar_wait(&kv_tmem, 0);
      if constexpr (kSlots3) {
        // halves k = 2 * step + half rotate through slots k % 3; the scores run three halves ahead of the gradients
        const int n_half = 2 * n_loop;
        if (n_loop > 0) {
          mbar_wait(&in_full[0], 0);
          tc_fence_after();
          issue_score(0, 0, 0);
          issue_score(1, 0, 1);
        }
        if (n_loop > 1) {
          mbar_wait(&in_full[1], 0);
          tc_fence_after();
          issue_score(0, 1, 2);
        }
        int slot = 0;
        uint32_t ph = 0;
        for (int k = 0; k < n_half; ++k) {
          const int it = k >> 1, hf = k & 1, s = it % NS;
          if (hf == 0) fa_trace(0, it, 0);
          mbar_wait(&p_full[slot], ph);
          fa_trace(0, it, 1 + 2 * hf);
          tc_fence_after();
          issue_grad(hf, s, slot, k == 0);
          if (hf == 1) tc_commit(&in_empty[s]);
          const int kn = k + 3;
          if (kn < n_half) {
            const int itn = kn >> 1, hfn = kn & 1, sn = itn % NS;
            if (hfn == 0) {   // first touch of that block's Q / dO stage
              mbar_wait(&in_full[sn], (itn / NS) & 1);
              tc_fence_after();
            }
            issue_score(hfn, sn, slot);
          }
          fa_trace(0, it, 2 + 2 * hf);
          if (++slot == 3) slot = 0, ph ^= 1;
        }
      } else {
      if (!kAmask || n_loop > 0) {
        mbar_wait(&in_full[0], 0);
        tc_fence_after();
        issue_score(0, 0, 0);
        issue_score(1, 0, 1);
      }
      for (int it = 0; it < n_loop; ++it) {   // (`it` counts list steps here: only stages and phases depend on it)
        const int s = it % NS, sn = (it + 1) % NS;
        const bool more = it + 1 < n_loop;
        fa_trace(0, it, 0);
        mbar_wait(&p_full[0], it & 1);
        fa_trace(0, it, 1);
        tc_fence_after();
        issue_grad(0, s, 0, it == 0);
        if (more) {
          mbar_wait(&in_full[sn], ((it + 1) / NS) & 1);
          tc_fence_after();
          issue_score(0, sn, 0);
        }
        fa_trace(0, it, 2);
        mbar_wait(&p_full[1], it & 1);
        fa_trace(0, it, 3);
        tc_fence_after();
        issue_grad(1, s, 1, false);
        tc_commit(&in_empty[s]);
        if (more) issue_score(1, sn, 1);
        fa_trace(0, it, 4);
      }
      }
      tc_commit(&acc_full);
    }
    __syncwarp();
  }
  } else {
    setmaxnreg_inc<216>();
    // ------------------------------------------------------------------ elementwise: P^T, dS^T  (warps 0-7)
    const int half = warp >> 2;               // (split: epilogue role only; in the loop the warp serves both halves)
    const int row = (warp & 3) * 32 + lane;   // key row inside the block == TMEM lane
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const float sl2 = p.scale_log2;
    if constexpr (kKvTmem) {
      // stationary operands: warpgroup a moves K_j, warpgroup b moves V_j from the swizzled TMA tile into TMEM,
      // row r -> lane r, elements (2c, 2c + 1) -> column c (the dQ kernel does the same with Q_i / dO_i)
      mbar_wait(&kv_full, 0);
      const uint32_t src = smem_u32(half == 0 ? sK : sV);
      uint32_t v[32];
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const uint32_t a = src + sw128_offset(row, ch);
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(v[ch * 4]), "=r"(v[ch * 4 + 1]), "=r"(v[ch * 4 + 2]), "=r"(v[ch * 4 + 3])
                     : "r"(a));
      }
      tmem_st_x32(tmem + (half == 0 ? kTmemKA : kTmemVA) + lane_base, v);
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&kv_tmem);
    }
    if constexpr (kSplit) {
    const int cbase = (warp >> 2) * 32;       // my 32 columns of every half
    uint32_t drop_col = 0, drop_shift = 0;
    if constexpr (kDrop) {
      drop_col = drop_key(p.drop, b * p.H + h) + drop_word_index(0, k0 + row);
      drop_shift = 8u * ((k0 + row) & 1);
    }
    const uint8_t* amt_row = nullptr;   // this key's row of the transposed attention mask
    if constexpr (kAmask) {
      if (p.amask_t)
        amt_row = p.amask_t + (int64_t)b * p.amt_s[0] + (int64_t)h * p.amt_s[1] + (int64_t)min(k0 + row, p.N - 1) * p.amt_s[2];
    }
    int slot3 = 0;          // three-slot rotation: slot and barrier phase of the current half
    uint32_t ph3 = 0;
    for (int k = 0; k < n_loop; ++k) {
      const int it = kAmask ? block_of(k) : k;
      const int s = k % NS;
      bool full = false;
      if constexpr (kAmask)
        full = use_list && p.ablock[(int64_t)b * p.ab_s[0] + (int64_t)h * p.ab_s[1] + (int64_t)(i_begin + it) * p.ab_s[2] + jb] == 2;
      const bool band = kAmask && !amt_row && !full;   // cut by a band mask (no mask bytes: the visible range is computed)
      const int ss = k % NT;   // statistics stage
      mbar_wait(&stat_full[ss], (k / NT) & 1);
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int slot = kSlots3 ? slot3 : hf;
        const uint32_t sc_ph = kSlots3 ? ph3 : (uint32_t)(k & 1);
        uint32_t mkw = 0xffffffffu;
        if constexpr (kAmask) {
          if (amt_row && !full) mkw = __ldg(reinterpret_cast<const uint32_t*>(amt_row + (i_begin + it) * 16 + hf * 8 + (cbase >> 3)));
        }
        // band: key j = k0 + row sees the queries i with j - win_right <= i <= j + win_left
        const int band_base = (i_begin + it) * 128 + hf * 64;
        const int band_lo = k0 + row - p.win_right - band_base, band_hi = k0 + row + p.win_left - band_base;
        if ((threadIdx.x & 127) == 0) fa_trace(1 + (warp >> 2), k, 3 * hf);
        mbar_wait(&sc_full[slot], sc_ph);
        if ((threadIdx.x & 127) == 0) fa_trace(1 + (warp >> 2), k, 3 * hf + 1);
        tc_fence_after();
        const uint32_t tS = tmem + slot_s(slot) + cbase + lane_base;
        const uint32_t tDP = tmem + slot_dp(slot) + cbase + lane_base;
        const uint32_t st = smem_u32(sStat + ss * 256 + hf * 64);
        const uint32_t dw = drop_col + drop_word_index((i_begin + it) * 128 + hf * 64, 0);
        if (kAmask && band) {
          if (kCausal && it == 0)
            dkdv_elementwise_chunk<kBf16, true, kDrop, kAmask, kAmask, kPolyMask>(tS, tDP, st, sl2, row, hf * 64, cbase, dw, drop_shift,
                                                                                  p.drop.thresh, p.drop.rp, mkw, band_lo, band_hi);
          else
            dkdv_elementwise_chunk<kBf16, false, kDrop, kAmask, kAmask, kPolyMask>(tS, tDP, st, sl2, row, hf * 64, cbase, dw, drop_shift,
                                                                                   p.drop.thresh, p.drop.rp, mkw, band_lo, band_hi);
        } else if (kCausal && it == 0)   // query block == key block: the only block that needs the causal mask
          dkdv_elementwise_chunk<kBf16, true, kDrop, kAmask, false, kPolyMask>(tS, tDP, st, sl2, row, hf * 64, cbase, dw, drop_shift,
                                                                               p.drop.thresh, p.drop.rp, mkw, 0, 0);
        else
          dkdv_elementwise_chunk<kBf16, false, kDrop, kAmask, false, kPolyMask>(tS, tDP, st, sl2, row, hf * 64, cbase, dw, drop_shift,
                                                                                p.drop.thresh, p.drop.rp, mkw, 0, 0);
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(&p_full[slot]);
        if ((threadIdx.x & 127) == 0) fa_trace(1 + (warp >> 2), k, 3 * hf + 2);
        if constexpr (kSlots3) {
          if (++slot3 == 3) slot3 = 0, ph3 ^= 1;
        }
      }
      __syncwarp();   // every lane has read its statistics of this block
      if (lane == 0) mbar_arrive(&stat_empty[ss]);
    }
    } else {
    const uint32_t tS = tmem + Cfg::kTmemS + half * 64 + lane_base;
    const uint32_t tDP = tmem + Cfg::kTmemDP + half * 64 + lane_base;
    // dropout: this thread walks column (key) k0 + row of the mask, one hash per pair of queries (fa_dropout.cuh)
    uint32_t drop_col = 0, drop_shift = 0;
    if constexpr (kDrop) {
      drop_col = drop_key(p.drop, b * p.H + h) + drop_word_index(half * 64, k0 + row);
      drop_shift = 8u * ((k0 + row) & 1);
    }

    const uint8_t* amt_row = nullptr;   // this key's row of the transposed attention mask
    if constexpr (kAmask) {
      if (p.amask_t)
        amt_row = p.amask_t + (int64_t)b * p.amt_s[0] + (int64_t)h * p.amt_s[1] + (int64_t)min(k0 + row, p.N - 1) * p.amt_s[2];
    }

    for (int k = 0; k < n_loop; ++k) {
      const int it = kAmask ? block_of(k) : k;
      const int s = k % NS;
      uint32_t mk[2] = {};
      bool band = false;   // kAmask: this block is cut by a band mask (no mask bytes: the visible range is computed)
      if constexpr (kAmask) {
        const bool full = use_list && p.ablock[(int64_t)b * p.ab_s[0] + (int64_t)h * p.ab_s[1] + (int64_t)(i_begin + it) * p.ab_s[2] + jb] == 2;
        band = !amt_row && !full;
        amask_load64(mk, amt_row + (i_begin + it) * 16 + half * 8, full || !amt_row);
      }
      // band: key j = k0 + row sees the queries i with j - win_right <= i <= j + win_left
      const int band_base = (i_begin + it) * 128 + half * 64;
      const int band_lo = k0 + row - p.win_right - band_base, band_hi = k0 + row + p.win_left - band_base;
      if ((threadIdx.x & 127) == 0) fa_trace(1 + half, k, 0);
      const int ss = k % NT;   // statistics stage
      mbar_wait(&stat_full[ss], (k / NT) & 1);
      const uint32_t st = smem_u32(sStat + ss * 256 + half * 64);
      if constexpr (kTwoStage) {
        mbar_wait(&s_full[half], k & 1);
        if ((threadIdx.x & 127) == 0) fa_trace(1 + half, k, 1);
        tc_fence_after();
        uint32_t pf[64];
        const bool diag = kCausal && it == 0;   // query block == key block: the only block that needs the causal mask
        if (kAmask && band) {
          if (diag) dkdv_p_stage<kBf16, true, kAmask, kAmask, kPolyMask>(tS, st, sl2, row, half * 64, mk, band_lo, band_hi, pf);
          else dkdv_p_stage<kBf16, false, kAmask, kAmask, kPolyMask>(tS, st, sl2, row, half * 64, mk, band_lo, band_hi, pf);
        } else if (diag) {
          dkdv_p_stage<kBf16, true, kAmask, false, kPolyMask>(tS, st, sl2, row, half * 64, mk, 0, 0, pf);
        } else {
          dkdv_p_stage<kBf16, false, kAmask, false, kPolyMask>(tS, st, sl2, row, half * 64, mk, 0, 0, pf);
        }
        mbar_wait(&sc_full[half], k & 1);
        if ((threadIdx.x & 127) == 0) fa_trace(1 + half, k, 3);
        tc_fence_after();
        dkdv_ds_stage<kBf16>(tDP, st, pf);
      } else {
      mbar_wait(&sc_full[half], k & 1);
      if ((threadIdx.x & 127) == 0) fa_trace(1 + half, k, 1);
      tc_fence_after();
      const uint32_t dw = drop_col + drop_word_index((i_begin + it) * 128, 0);
      if (kAmask && band) {
        if (kCausal && it == 0)
          bwd_elementwise_half<kBf16, true, true, true, true, kDrop, kAmask, kAmask, kPolyMask>(
              tS, tDP, st, 0ull, 0ull, sl2, row, half * 64, dw, drop_shift, p.drop.thresh, p.drop.rp, mk, band_lo, band_hi);
        else
          bwd_elementwise_half<kBf16, true, false, true, true, kDrop, kAmask, kAmask, kPolyMask>(
              tS, tDP, st, 0ull, 0ull, sl2, row, half * 64, dw, drop_shift, p.drop.thresh, p.drop.rp, mk, band_lo, band_hi);
      } else if (kCausal && it == 0)   // query block == key block: the only block that needs the causal mask
        bwd_elementwise_half<kBf16, true, true, true, true, kDrop, kAmask, false, kPolyMask>(tS, tDP, st, 0ull, 0ull, sl2, row, half * 64, dw,
                                                                           drop_shift, p.drop.thresh, p.drop.rp, mk);
      else
        bwd_elementwise_half<kBf16, true, false, true, true, kDrop, kAmask, false, kPolyMask>(tS, tDP, st, 0ull, 0ull, sl2, row, half * 64,
                                                                            dw, drop_shift, p.drop.thresh, p.drop.rp, mk);
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&p_full[half]);
      if ((threadIdx.x & 127) == 0) fa_trace(1 + half, k, 2);
      __syncwarp();   // every lane has read its statistics of this block
      if (lane == 0) mbar_arrive(&stat_empty[ss]);
    }
    }

    // epilogue: warpgroup a stores dV, warpgroup b stores scale * dK
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    const int kv_row = k0 + row;
    const bool in_range = kv_row < nk;
    if (half == 0) {
      uint16_t* dst = reinterpret_cast<uint16_t*>(p.dv) + b * p.dv_s[0] + h * p.dv_s[1] + (int64_t)kv_row * p.dv_s[2];
      store_acc_rows<kBf16>(tmem + kAccV + lane_base, kD, kDrop ? p.drop.rp : 1.0f, dst, in_range,
                            !kAmask || n_loop > 0);
    } else {
      uint16_t* dst = reinterpret_cast<uint16_t*>(p.dk) + b * p.dk_s[0] + h * p.dk_s[1] + (int64_t)kv_row * p.dk_s[2];
      store_acc_rows<kBf16>(tmem + kAccK + lane_base, kD, p.scale, dst, in_range, !kAmask || n_loop > 0);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem);
}

// ================================================================================================ dQ
template <bool kBf16, int kD, bool kCausal, bool kDrop = false, bool kAmask = false>
__global__ void __launch_bounds__(384, 1)
fa_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                 const BwdParams p) {
  using Cfg = BwdCfg<kD>;
  constexpr int NK = Cfg::kStagesK, NV = Cfg::kStagesV;
  constexpr bool kSplit = bwd_ew_split<kD>();
  constexpr bool kDsTmem = dq_ds_tmem<kD>();
  constexpr bool kFullScore = dq_full_score<kD>();
  constexpr uint32_t kTmemDSFull = 384;   // kFullScore: packed dS, 32 columns per half
  constexpr int kPolyMask = kAmask ? 0 : (kD == 64 ? FA_BWD_POLY_MASK_D64 : FA_BWD_POLY_MASK);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // stationary Q_i
  uint8_t* sDO = sQ + Cfg::kTileBytes;                  // stationary dO_i
  uint8_t* sK = sDO + Cfg::kTileBytes;                  // [NK] streamed K_j
  uint8_t* sV = sK + NK * Cfg::kTileBytes;              // [NV] streamed V_j

  __shared__ uint64_t qdo_full, qdo_tmem, acc_full;
  __shared__ uint64_t k_full[NK], k_empty[NK], v_full[NV], v_empty[NV];
  __shared__ uint64_t sc_full[2], sc_free[2], p_full[2], ds_free[2];
  __shared__ uint32_t tmem_base_s;
  // kAmask with a block summary: the key blocks with something visible for this query block, in order
  __shared__ uint16_t s_list[kAmask ? 512 : 2];
  __shared__ int s_nact;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_blocks = (p.N + 127) >> 7;
  const int ib = n_blocks - 1 - (int)blockIdx.x;        // heaviest (most key blocks) first under the causal mask
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = ib * 128;
  const int nv = p.seqlens ? min(max(p.seqlens[b], 0), p.N) : p.N;   // valid length of this batch element
  if (q0 >= nv) return;   // padded query block
  const int nk = p.Nk > 0 ? p.Nk : nv;   // valid key rows
  const int n_kv_valid = (nk + 127) >> 7;
  const int n_it = kCausal ? ib + 1 : n_kv_valid;
  // Non-causal: keys >= nv in the last key block are masked explicitly (padded K rows are real data, not TMA zero fill);
  // causal: the diagonal mask of the last block already removes them (valid rows are < nv).
  const bool tail_mask = !kCausal && (nk & 127) != 0;
  const bool use_list = kAmask && p.ablock != nullptr && n_it <= 512;
  if constexpr (kAmask) {
    if (use_list && warp == 0) {   // one warp compacts the flags (ballot + prefix count), 32 blocks per step
      const uint8_t* ab = p.ablock + (int64_t)b * p.ab_s[0] + (int64_t)h * p.ab_s[1] + (int64_t)ib * p.ab_s[2];
      int c = 0;
      for (int base = 0; base < n_it; base += 32) {
        const int it = base + lane;
        const bool on = it < n_it && ab[it] != 0;
        const uint32_t m = __ballot_sync(0xffffffffu, on);
        if (on) s_list[c + __popc(m & ((1u << lane) - 1u))] = (uint16_t)it;
        c += __popc(m);
      }
      if (lane == 0) s_nact = c;
    }
  }

  if (threadIdx.x == 0) {
    mbar_init(&qdo_full, 1);
    mbar_init(&qdo_tmem, 256);
    mbar_init(&acc_full, 1);
    for (int s = 0; s < NK; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
    }
    for (int s = 0; s < NV; ++s) {
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&sc_full[t], 1);
      mbar_init(&sc_free[t], kSplit ? 256 : 128);
      mbar_init(&p_full[t], kSplit ? 256 : 128);
      mbar_init(&ds_free[t], 1);
    }
    fence_mbar_init();
  }
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmDO);
  }
  if (warp == 9) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  // every role walks the same list: step k handles key block block_of(k); ring stages and barrier phases are functions
  // of k, so the pipeline protocol is exactly the dense one over a shorter sequence
  const int n_loop = use_list ? s_nact : n_it;
  auto block_of = [&](int k) -> int { return use_list ? (int)s_list[k] : k; };

  if (warp >= 8) {
  setmaxnreg_dec<72>();
  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(&qdo_full, 2 * Cfg::kTileBytes);
      for (int bx = 0; bx < Cfg::kBoxes; ++bx) {
        tma_load_4d(sQ + bx * Cfg::kBoxBytes, &tmQ, &qdo_full, bx * 64, q0, h, b);
        tma_load_4d(sDO + bx * Cfg::kBoxBytes, &tmDO, &qdo_full, bx * 64, q0, h, b);
      }
      for (int k = 0; k < n_loop; ++k) {
        const int kv0 = (kAmask ? block_of(k) : k) * 128;
        const int sk = k % NK, sv = k % NV;
        mbar_wait(&k_empty[sk], ((k / NK) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[sk], Cfg::kTileBytes);
        for (int bx = 0; bx < Cfg::kBoxes; ++bx)
          tma_load_4d(sK + sk * Cfg::kTileBytes + bx * Cfg::kBoxBytes, &tmK, &k_full[sk], bx * 64, kv0, h, b);
        mbar_wait(&v_empty[sv], ((k / NV) & 1) ^ 1);
        mbar_arrive_expect_tx(&v_full[sv], Cfg::kTileBytes);
        for (int bx = 0; bx < Cfg::kBoxes; ++bx)
          tma_load_4d(sV + sv * Cfg::kTileBytes + bx * Cfg::kBoxBytes, &tmV, &v_full[sv], bx * 64, kv0, h, b);
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_sc = umma_idesc_f16(kBf16, 128, 64, 0, 0);   // [128 q] x [64 kv], K = D
      constexpr uint32_t idesc_gr = umma_idesc_f16(kBf16, 128, kD, 0, 1);   // [128 q] x [D],    K = 64 kv
      constexpr uint32_t kTileLo = Cfg::kTileBytes >> 4, kHalfLo = 8192u >> 4;
      const uint32_t k_lo = umma_lo_kmajor(smem_u32(sK)), v_lo = umma_lo_kmajor(smem_u32(sV));
      const uint32_t k_mn = umma_lo_mnmajor(smem_u32(sK), Cfg::kBoxBytes);

      // S half = Q_i K_j[half]^T ; dP half = dO_i V_j[half]^T      (A operands resident in TMEM)
      const uint32_t q_lo = umma_lo_kmajor(smem_u32(sQ)), do_lo = umma_lo_kmajor(smem_u32(sDO));
      constexpr uint32_t idesc_full = umma_idesc_f16(kBf16, 128, 128, 0, 0);   // [128 q] x [128 kv], K = D
      auto issue_score = [&](int half, int sk, int sv) {
        if constexpr (kFullScore) {
          // both halves at once, operands from shared memory; each half's barrier is committed
          if (half != 0) return;
          const uint32_t bk = k_lo + sk * kTileLo, bv = v_lo + sv * kTileLo;
          static_for<0, kD / 16>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            constexpr uint32_t off = umma_koff_kmajor(k, Cfg::kBoxBytes);
            umma_ss_off<off, off>(tmem + Cfg::kTmemS, q_lo, bk, idesc_full, k > 0);
          });
          static_for<0, kD / 16>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            constexpr uint32_t off = umma_koff_kmajor(k, Cfg::kBoxBytes);
            umma_ss_off<off, off>(tmem + Cfg::kTmemDP, do_lo, bv, idesc_full, k > 0);
          });
          tc_commit(&sc_full[0]);
          tc_commit(&sc_full[1]);
          return;
        }
        const uint32_t bk = k_lo + sk * kTileLo + half * kHalfLo, bv = v_lo + sv * kTileLo + half * kHalfLo;
        const uint32_t dS = tmem + Cfg::kTmemS + half * 64, dDP = tmem + Cfg::kTmemDP + half * 64;
        const uint32_t aQ = tmem + Cfg::kTmemQA, aDO = tmem + Cfg::kTmemDOA;
        static_for<0, kD / 16>([&](auto kc) {
          constexpr int k = decltype(kc)::value;
          umma_ts_off<k * 8, umma_koff_kmajor(k, Cfg::kBoxBytes)>(dS, aQ, bk, idesc_sc, k > 0);
        });
        static_for<0, kD / 16>([&](auto kc) {
          constexpr int k = decltype(kc)::value;
          umma_ts_off<k * 8, umma_koff_kmajor(k, Cfg::kBoxBytes)>(dDP, aDO, bv, idesc_sc, k > 0);
        });
        tc_commit(&sc_full[half]);
      };
      // dQ += dS[half] K_j[half]
      // dS half lives in shared memory (one 128 x 64 swizzled box per half, K-major A operand): sdS[a] reuses the
      // first box of the Q_i staging tile, sdS[b] the first box of the dO_i staging tile (both dead after the TMEM copy).
      const uint32_t ds_lo[2] = {umma_lo_kmajor(smem_u32(sQ)), umma_lo_kmajor(smem_u32(sDO))};
      constexpr uint32_t idesc_gs = umma_idesc_f16(kBf16, 128, kD, 0, 1);
      auto issue_grad = [&](int half, int s, bool first) {
        const uint32_t bk = k_mn + s * kTileLo + half * umma_koff_mnmajor(4);
        const uint32_t dQ_t = tmem + Cfg::kTmemAcc0, aDS = ds_lo[half];
        static_for<0, 4>([&](auto kc) {
          constexpr int k = decltype(kc)::value;
          if constexpr (kDsTmem)
            umma_ts_off<k * 8, umma_koff_mnmajor(k)>(dQ_t, tmem + Cfg::kTmemDS + half * 32, bk, idesc_gs,
                                                     !(first && k == 0));
          else if constexpr (kFullScore)
            umma_ts_off<k * 8, umma_koff_mnmajor(k)>(dQ_t, tmem + kTmemDSFull + half * 32, bk, idesc_gs,
                                                     !(first && k == 0));
          else
            umma_ss_off<umma_koff_kmajor(k, Cfg::kBoxBytes), umma_koff_mnmajor(k)>(dQ_t, aDS, bk, idesc_gs,
                                                                                  !(first && k == 0));
        });
        tc_commit(&ds_free[half]);
      };

      if constexpr (kFullScore) mbar_wait(&qdo_full, 0);
      else mbar_wait(&qdo_tmem, 0);
      if (!kAmask || n_loop > 0) {
        mbar_wait(&k_full[0], 0);
        mbar_wait(&v_full[0], 0);
        tc_fence_after();
        issue_score(0, 0, 0);
        issue_score(1, 0, 0);
        tc_commit(&v_empty[0]);
      }
      for (int it = 0; it < n_loop; ++it) {   // (`it` counts list steps here: only stages and phases depend on it)
        const int sk = it % NK, skn = (it + 1) % NK, svn = (it + 1) % NV;
        const bool more = it + 1 < n_loop;
        // the score accumulators are free as soon as the elementwise warps hold them in registers: the next
        // scores run on the tensor core while dS is being computed
        fa_trace(0, it, 0);
        if (more) {
          mbar_wait(&sc_free[0], it & 1);
          if constexpr (kFullScore) mbar_wait(&sc_free[1], it & 1);
          mbar_wait(&k_full[skn], ((it + 1) / NK) & 1);
          mbar_wait(&v_full[svn], ((it + 1) / NV) & 1);
          tc_fence_after();
          issue_score(0, skn, svn);
          fa_trace(0, it, 1);
          mbar_wait(&sc_free[1], it & 1);
          tc_fence_after();
          issue_score(1, skn, svn);
          tc_commit(&v_empty[svn]);
        }
        fa_trace(0, it, 2);
        mbar_wait(&p_full[0], it & 1);
        fa_trace(0, it, 3);
        tc_fence_after();
        issue_grad(0, sk, it == 0);
        mbar_wait(&p_full[1], it & 1);
        fa_trace(0, it, 4);
        tc_fence_after();
        issue_grad(1, sk, false);
        tc_commit(&k_empty[sk]);
        fa_trace(0, it, 5);
      }
      tc_commit(&acc_full);
    }
    __syncwarp();
  }
  } else {
    setmaxnreg_inc<216>();
    // ------------------------------------------------------------------ elementwise: dS  (warps 0-7)
    const int half = warp >> 2;
    const int row = (warp & 3) * 32 + lane;   // query row inside the block == TMEM lane
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + Cfg::kTmemS + half * 64 + lane_base;
    const uint32_t tDP = tmem + Cfg::kTmemDP + half * 64 + lane_base;
    const float sl2 = p.scale_log2;
    const int q_row = q0 + row;
    const bool in_range = q_row < nv;
    const int64_t stat_idx = ((int64_t)b * p.H + h) * p.N + q_row;
    float neg_lse = in_range ? -p.lse[stat_idx] : -INFINITY;
    if (kAmask && neg_lse == INFINITY) neg_lse = -INFINITY;   // a query that saw no key (L = -inf): P = 0
    const uint8_t* am_row = nullptr;   // this query's row of the attention mask
    if constexpr (kAmask) {
      if (p.amask) am_row = p.amask + (int64_t)b * p.am_s[0] + (int64_t)h * p.am_s[1] + (int64_t)min(q_row, p.N - 1) * p.am_s[2];
    }
    const float neg_dl = in_range ? -p.delta[stat_idx] : 0.f;
    const uint64_t nl2 = f32x2_pack(neg_lse, neg_lse), nd2 = f32x2_pack(neg_dl, neg_dl);
    // dropout: this thread walks row q_row of the mask, one hash per pair of keys (fa_dropout.cuh)
    uint32_t drop_row = 0, drop_shift = 0;
    if constexpr (kDrop) {
      drop_row = drop_key(p.drop, b * p.H + h) + drop_word_index(q_row, half * 64);
      drop_shift = 16u * (q_row & 1);
    }

    // Stationary operands: warpgroup a moves Q_i, warpgroup b moves dO_i from the (swizzled) TMA tile into TMEM,
    // row r -> lane r, elements (2c, 2c+1) -> column c.  As TMEM A operands they cost no shared-memory bandwidth
    // in the score MMAs (an SS MMA with N = 64 needs 192 B/clk of smem reads, above the 128 B/clk an SM has).
    if constexpr (!kFullScore) {
      mbar_wait(&qdo_full, 0);
      const uint32_t src = smem_u32(half == 0 ? sQ : sDO);
      const uint32_t dstA = tmem + (half == 0 ? Cfg::kTmemQA : Cfg::kTmemDOA) + lane_base;
#pragma unroll
      for (int bx = 0; bx < Cfg::kBoxes; ++bx) {
        uint32_t v[32];
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const uint32_t a = src + bx * Cfg::kBoxBytes + sw128_offset(row, ch);
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(v[ch * 4]), "=r"(v[ch * 4 + 1]), "=r"(v[ch * 4 + 2]), "=r"(v[ch * 4 + 3])
                       : "r"(a));
        }
        tmem_st_x32(dstA + bx * 32, v);
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&qdo_tmem);
    }

    if constexpr (kSplit) {
    // all eight warps serve half a, then half b: warp w and w + 4 split the half's 64 key columns (32 each)
    const int cbase = (warp >> 2) * 32;
    uint32_t drop_row0 = 0;
    if constexpr (kDrop) drop_row0 = drop_key(p.drop, b * p.H + h) + drop_word_index(q_row, 0);
    for (int k = 0; k < n_loop; ++k) {
      const int it = kAmask ? block_of(k) : k;   // key block of this step
      bool full = false;
      if constexpr (kAmask)
        full = use_list && p.ablock[(int64_t)b * p.ab_s[0] + (int64_t)h * p.ab_s[1] + (int64_t)ib * p.ab_s[2] + it] == 2;
      const bool band = kAmask && !am_row && !full;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        if ((threadIdx.x & 127) == 0) fa_trace(1 + (warp >> 2), it, 3 * hf);   // about to wait for scores
        uint32_t mkw = 0xffffffffu;
        if constexpr (kAmask) {
          if (am_row && !full) mkw = __ldg(reinterpret_cast<const uint32_t*>(am_row + it * 16 + hf * 8 + (cbase >> 3)));
        }
        // band: query i = q_row sees the keys j with i - win_left <= j <= i + win_right
        const int band_base = it * 128 + hf * 64;
        const int band_lo = q_row - p.win_left - band_base, band_hi = q_row + p.win_right - band_base;
        mbar_wait(&sc_full[hf], k & 1);
        if ((threadIdx.x & 127) == 0) fa_trace(1 + (warp >> 2), it, 3 * hf + 1);   // scores ready
        tc_fence_after();
        const uint32_t tSc = tmem + Cfg::kTmemS + hf * 64 + cbase + lane_base;
        const uint32_t tDPc = tmem + Cfg::kTmemDP + hf * 64 + cbase + lane_base;
        uint32_t pd[16];
        const uint32_t dw = drop_row0 + (uint32_t)(it * 64 + hf * 32);
        if (kAmask && band) {
          if (kCausal && it == n_it - 1)
            dq_elementwise_chunk<kBf16, true, kDrop, kAmask, kAmask, kPolyMask>(tSc, tDPc, &sc_free[hf], nl2, nd2, sl2, row, hf * 64, cbase, pd,
                                                                               dw, drop_shift, p.drop.thresh, p.drop.rp, mkw, band_lo, band_hi);
          else if (tail_mask && it == n_it - 1)
            dq_elementwise_chunk<kBf16, true, kDrop, kAmask, kAmask, kPolyMask>(tSc, tDPc, &sc_free[hf], nl2, nd2, sl2, nk - it * 128 - 1,
                                                                               hf * 64, cbase, pd, dw, drop_shift, p.drop.thresh, p.drop.rp, mkw,
                                                                               band_lo, band_hi);
          else
            dq_elementwise_chunk<kBf16, false, kDrop, kAmask, kAmask, kPolyMask>(tSc, tDPc, &sc_free[hf], nl2, nd2, sl2, row, hf * 64, cbase, pd,
                                                                                dw, drop_shift, p.drop.thresh, p.drop.rp, mkw, band_lo, band_hi);
        } else if (kCausal && it == n_it - 1)   // key block == query block: keep key <= query
          dq_elementwise_chunk<kBf16, true, kDrop, kAmask, false, kPolyMask>(tSc, tDPc, &sc_free[hf], nl2, nd2, sl2, row, hf * 64, cbase, pd, dw,
                                                                            drop_shift, p.drop.thresh, p.drop.rp, mkw, 0, 0);
        else if (tail_mask && it == n_it - 1)   // last key block: keep key < nv
          dq_elementwise_chunk<kBf16, true, kDrop, kAmask, false, kPolyMask>(tSc, tDPc, &sc_free[hf], nl2, nd2, sl2, nk - it * 128 - 1,
                                                                            hf * 64, cbase, pd, dw, drop_shift, p.drop.thresh, p.drop.rp, mkw, 0, 0);
        else
          dq_elementwise_chunk<kBf16, false, kDrop, kAmask, false, kPolyMask>(tSc, tDPc, &sc_free[hf], nl2, nd2, sl2, row, hf * 64, cbase, pd, dw,
                                                                             drop_shift, p.drop.thresh, p.drop.rp, mkw, 0, 0);
        if (k > 0) mbar_wait(&ds_free[hf], (k - 1) & 1);          // dQ MMAs of the previous block have read dS
        if constexpr (kDsTmem) {
          tc_fence_after();
          tmem_st_x16(tmem + Cfg::kTmemDS + hf * 32 + (cbase >> 1) + lane_base, pd);
          tc_wait_st();
          tc_fence_before();
        } else {
          const uint32_t sds = smem_u32(hf == 0 ? sQ : sDO);       // this half's dS box (see the MMA warp)
#pragma unroll
          for (int ch = 0; ch < 4; ++ch)
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sds + sw128_offset(row, (cbase >> 3) + ch)), "r"(pd[ch * 4]),
                         "r"(pd[ch * 4 + 1]), "r"(pd[ch * 4 + 2]), "r"(pd[ch * 4 + 3])
                         : "memory");
          fence_proxy_async_smem();
        }
        mbar_arrive(&p_full[hf]);
        if ((threadIdx.x & 127) == 0) fa_trace(1 + (warp >> 2), it, 3 * hf + 2);   // dS handed over
      }
    }
    } else {
    const uint32_t sds = smem_u32(half == 0 ? sQ : sDO);   // this half's dS box (see the MMA warp)
    for (int k = 0; k < n_loop; ++k) {
      const int it = kAmask ? block_of(k) : k;   // key block of this step
      if (threadIdx.x == half * 128) fa_trace(1 + half, it, 0);   // about to wait for scores
      uint32_t mk[2] = {};
      bool band = false;   // kAmask: this block is cut by a band mask (no mask bytes: the visible range is computed)
      if constexpr (kAmask) {
        const bool full = use_list && p.ablock[(int64_t)b * p.ab_s[0] + (int64_t)h * p.ab_s[1] + (int64_t)ib * p.ab_s[2] + it] == 2;
        band = !am_row && !full;
        amask_load64(mk, am_row + it * 16 + half * 8, full || !am_row);
      }
      // band: query i = q_row sees the keys j with i - win_left <= j <= i + win_right
      const int band_base = it * 128 + half * 64;
      const int band_lo = q_row - p.win_left - band_base, band_hi = q_row + p.win_right - band_base;
      mbar_wait(&sc_full[half], k & 1);
      if (threadIdx.x == half * 128) fa_trace(1 + half, it, 1);   // scores ready
      tc_fence_after();
      uint32_t pd[32];
      const uint32_t dw = drop_row + (uint32_t)(it * 64);
      if (kAmask && band) {
        if (kCausal && it == n_it - 1)
          dq_elementwise_half<kBf16, true, kDrop, kAmask, kAmask, kPolyMask>(tS, tDP, &sc_free[half], nl2, nd2, sl2, row, half * 64, pd,
                                                                  dw, drop_shift, p.drop.thresh, p.drop.rp, mk, band_lo, band_hi);
        else if (tail_mask && it == n_it - 1)
          dq_elementwise_half<kBf16, true, kDrop, kAmask, kAmask, kPolyMask>(tS, tDP, &sc_free[half], nl2, nd2, sl2, nk - it * 128 - 1,
                                                                  half * 64, pd, dw, drop_shift, p.drop.thresh, p.drop.rp, mk,
                                                                  band_lo, band_hi);
        else
          dq_elementwise_half<kBf16, false, kDrop, kAmask, kAmask, kPolyMask>(tS, tDP, &sc_free[half], nl2, nd2, sl2, row, half * 64, pd,
                                                                   dw, drop_shift, p.drop.thresh, p.drop.rp, mk, band_lo, band_hi);
      } else if (kCausal && it == n_it - 1)   // key block == query block: keep key <= query
        dq_elementwise_half<kBf16, true, kDrop, kAmask, false, kPolyMask>(tS, tDP, &sc_free[half], nl2, nd2, sl2, row, half * 64, pd, dw,
                                                        drop_shift, p.drop.thresh, p.drop.rp, mk);
      else if (tail_mask && it == n_it - 1)   // last key block: keep key < nv
        dq_elementwise_half<kBf16, true, kDrop, kAmask, false, kPolyMask>(tS, tDP, &sc_free[half], nl2, nd2, sl2, nk - it * 128 - 1,
                                                        half * 64, pd, dw, drop_shift, p.drop.thresh, p.drop.rp, mk);
      else
        dq_elementwise_half<kBf16, false, kDrop, kAmask, false, kPolyMask>(tS, tDP, &sc_free[half], nl2, nd2, sl2, row, half * 64, pd, dw,
                                                         drop_shift, p.drop.thresh, p.drop.rp, mk);
      if (k > 0) mbar_wait(&ds_free[half], (k - 1) & 1);          // dQ MMAs of the previous block have read dS
      if constexpr (kFullScore) {
        tc_fence_after();
        tmem_st_x32(tmem + kTmemDSFull + half * 32 + lane_base, pd);
        tc_wait_st();
        tc_fence_before();
      } else {
#pragma unroll
        for (int ch = 0; ch < 8; ++ch)
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sds + sw128_offset(row, ch)), "r"(pd[ch * 4]),
                       "r"(pd[ch * 4 + 1]), "r"(pd[ch * 4 + 2]), "r"(pd[ch * 4 + 3])
                       : "memory");
        fence_proxy_async_smem();
      }
      mbar_arrive(&p_full[half]);
      if (threadIdx.x == half * 128) fa_trace(1 + half, it, 2);   // dS handed over
    }
    }

    // epilogue: each warpgroup stores half of the D columns of scale * dQ
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    uint16_t* dst = reinterpret_cast<uint16_t*>(p.dq) + b * p.dq_s[0] + h * p.dq_s[1] + (int64_t)q_row * p.dq_s[2] +
                    half * (kD / 2);
    store_acc_rows<kBf16>(tmem + Cfg::kTmemAcc0 + lane_base + half * (kD / 2), kD / 2, p.scale, dst, in_range,
                          !kAmask || n_loop > 0);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem);
}

}  // namespace fa
