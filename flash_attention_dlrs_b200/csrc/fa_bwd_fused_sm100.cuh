// fa_bwd_fused_sm100.cuh — single-pass deterministic FlashAttention-2 backward for sm_100a (16-bit inputs).
//
// Same math as fa_bwd_sm100.cuh (flash_attention_kernels.py:275-329 with the scale / causal mask of
// flash_attention_openai_tutorial.py:239,292,389), but S and dP are computed ONCE per (query block, key block) pair:
// five matmuls per pair instead of the seven of the two-kernel backward.  The reference does the same five and adds
// each key block's dQ contribution into global memory under a CAS spin lock in whatever order the scheduler produces
// (flash_attention_kernels.py:305-320) — non-deterministic, and broken (README.md:45-53).  Here dQ goes through an
// ORDERED reduction:
//
//   * one CTA per key block j; it visits query blocks i = j, j+1, ... (causal: up to the last block; non-causal:
//     wrapping around to j-1), so that at any time the CTAs of one head work on different query blocks;
//   * the dQ_i partial of every visit is added to an fp32 tile in the workspace with `red.global.add.v4.f32` (the
//     additions happen in L2, nothing is read back), but only when the per-tile turn counter says it is this CTA's turn: contributions arrive in the fixed order
//     j = i, i-1, ..., 0 (buffer A) and j = n-1, n-2, ..., i+1 (buffer B, non-causal only).  The first contributor
//     stores instead of adding, so the workspace needs no zero fill.  A CTA only ever waits for the CTA that took the
//     ticket just before it (key block j+1 of the same head), which by construction reached the same query block
//     one visit earlier: no deadlock (the predecessor is already resident or finished) and, in steady state, no stall;
//   * fa_bwd_dq_convert_kernel finally writes dQ = scale * (A [+ B]) in the 16-bit output layout.
//   Floating-point addition order is therefore fixed: results are bit-identical run to run.
//
// CTA layout (512 threads):
//   warps 0-7   elementwise: P^T = exp2(S^T*scale*log2e - L), dS^T = P^T o (dP^T - delta); thread = key row (TMEM lane),
//               warpgroup a = query columns 0-63, warpgroup b = 64-127.  P^T goes back to TMEM as the packed 16-bit A
//               operand of dV; dS^T goes to shared memory, where it is the A operand of dK (K-major) and of dQ (MN-major).
//   warps 8-11  dQ reducer: TMEM -> registers -> ordered vector reductions into the workspace
//   warp 12     TMA producer (K_j, V_j once; Q_i, dO_i, -L_i, -delta_i through a 2-stage ring)
//   warp 13     MMA issuer (one elected thread), TMEM owner
// TMEM columns: S^T/P^T [0,128) | dP^T/dQ [128,256) | dV [256,256+D) | dK [256+D,256+2D).
// Tensor-pipe order per visit v:  dV(v)  S(v+1)  dQ(v)  dK(v)  dP(v+1).  dP(v+1) is issued as early as the drain of
// dQ(v) allows, so it is complete when the elementwise warps finish P(v+1): they never idle between the two stages
// (they are the critical resource: 128 x 128 exp2 per visit is 1024 clk of MUFU alone).
#pragma once

#include "fa_bwd_sm100.cuh"

// FA_FUSED_ABLATE (timing experiments only, results are wrong or unordered): 1 = no turn-taking, 2 = reducer only
// drains TMEM (no global traffic, no turn-taking).
#ifndef FA_FUSED_ABLATE
#define FA_FUSED_ABLATE 0
#endif
#ifndef FA_FUSED_PACE
#define FA_FUSED_PACE 40
#endif
// FA_FUSED_ABLATE 3 (timing experiment, wrong results): unordered and only half of every dQ partial leaves the SM.
// 5 / 6 (timing experiments, order not guaranteed): turn published with a relaxed store instead of the release /
// turns published but never waited for — which half of the hand-over costs what.
// FA_FUSED_TURN_WARPS: 1 = the turn-taking of the ordered dQ reduction runs on two otherwise idle warps (warp 14 polls
// the turn counter, warp 15 publishes the next turn) instead of on the reducer warps, so that neither the poll's L2 round
// trip nor the release's wait for the drain of ~64 KiB of reductions ever blocks the warps that feed the SM -> L2 path;
// 0 = the round-1 protocol (every reducer thread fences, thread 0 polls and releases).
#ifndef FA_FUSED_TURN_WARPS
#define FA_FUSED_TURN_WARPS 1
#endif
#ifndef FA_FUSED_HALF
#define FA_FUSED_HALF (FA_FUSED_ABLATE == 3 || FA_FUSED_ABLATE == 4)
#endif
#define FA_FUSED_UNORDERED (FA_FUSED_ABLATE == 1 || FA_FUSED_ABLATE == 2 || FA_FUSED_ABLATE == 3)
// FA_FUSED_TMA: 1 = the dQ partial leaves the SM through the TMA engine: the reducer warps stage it in shared memory in
// 16 KiB chunks (32 columns x 128 rows, the workspace's own [D/4][128][4] order, conflict-free 16-byte stores) and warp 15
// issues one `cp.reduce.async.bulk ... add.f32` per chunk (a plain bulk store for the tile's first contributor); the
// staging buffers take the place of the second dO stage, and the fixed order is kept per CHUNK (one turn counter per
// tile and chunk, polled ahead by warp 14).  Bit-identical to the register path and to itself (tests green), but
// measured slower (profiles/r02_fused_tma_egress.txt, config 3 causal): no egress 2.33 ms (the single dO stage alone
// costs 0.25 ms against 2.07), unordered 4.16, ordered 6.91 — against 2.97 / 3.85 for the register path and 2.9 for the
// two kernels.  Why no egress mechanism can win at D = 128: per visit the kernel moves 64 KiB of Q / dO in and 64 KiB of
// fp32 partial out, and a reduction costs the L2 slices twice what a load does (21 B/clk per SM for reductions, ~42 for
// loads, chip-wide) — 192 KiB-equivalents at ~42.6 B/clk per SM = 4600 clk per visit against 4030 of compute: the single
// pass is bound by L2 slice throughput at ~2.35 ms (cuDNN's kernel of the same design: 2.30), whichever unit issues the
// traffic; the TMA engine adds its in-order queue (Q / dO loads wait behind 16 KiB reductions).  Default 0.
#ifndef FA_FUSED_TMA
#define FA_FUSED_TMA 0
#endif

namespace fa {

struct FusedParams {
  BwdParams base;
  float* dq_acc;         // buffer A: (B*H*n_blocks) tiles of 128 x D fp32, tile layout [D/4][128 rows][4]
  long long acc_b_off;   // floats from buffer A to buffer B (0 when causal)
  int* ticket;           // CTA ticket counter (zeroed before the launch)
  int* sem;              // turn counters: [2][B*H*n_blocks][kSemPerTile] (zeroed before the launch)
  int n_blocks;
};

constexpr int kSemPerTile = 4;   // one turn counter per 32-column chunk of a tile (FA_FUSED_TMA), D <= 128

template <int kD>
struct FusedCfg {
  static constexpr int kTileBytes = 128 * kD * 2;
  static constexpr int kBoxBytes = 128 * 128;
  static constexpr int kBoxes = kD / 64;
  static constexpr int kDsBytes = 2 * kBoxBytes;   // dS^T: 128 key rows x 128 queries (two 64-query boxes)
  static constexpr int kStatFloats = 256;          // -lse and -delta of one query block
  static constexpr int kThreads = 512;
  static constexpr int kChunks = kD / 32;           // dQ egress chunks of 32 columns (FA_FUSED_TMA)
  static constexpr int kStageBytes = 128 * 32 * 4;  // one staged chunk: [8 vectors][128 rows][16 bytes]
#if FA_FUSED_TMA
  static constexpr int kTiles = 5 * kTileBytes + kDsBytes + 2 * kStageBytes;   // K, V, Q[2], dO, dS, staging[2]
#else
  static constexpr int kTiles = 6 * kTileBytes + kDsBytes;   // K, V, Q[2], dO[2], dS
#endif
  static constexpr int kCtrlBytes = 2 * kStatFloats * 4 + 256;           // statistics ring + barriers / scalars
  static constexpr int kAlignSlack = 512;   // the dynamic window starts 1024-aligned in practice (checked in-kernel)
  static constexpr int kSmemBytes = kTiles + kCtrlBytes + kAlignSlack;
  static constexpr uint32_t kTmemS = 0, kTmemDP = 128, kTmemDV = 256, kTmemDK = 256 + kD;
};

// FA_TRACE == 2 (debug builds): per-CTA wall-clock milestones {entry, set-up done, first P ready, main loop done, exit,
// SM id, visits} at g_fa_trace[ticket * 8 ...] (globaltimer, ns) to reconstruct the schedule of the whole grid.
#if FA_TRACE == 2
__device__ __forceinline__ void fa_cta_trace(int ticket, int k, long long v) {
  if (g_fa_trace != nullptr && ticket * 8 + k < g_fa_trace_cap) g_fa_trace[ticket * 8 + k] = v;
}
__device__ __forceinline__ long long fa_globaltimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define FA_CTA_TRACE(k) fa_cta_trace(ticket, k, fa_globaltimer())
#else
#define FA_CTA_TRACE(k)
#endif

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_add_f32x4(float* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ void st_f32x4(float* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// P^T stage of one thread (key row `row`, 64 query columns starting at `col0`): S^T from TMEM, P = exp2(S*sl2 - L[col])
// kept in fp32 in `pf` (dS needs it) and written back over S^T as packed 16-bit pairs.  kMask: causal diagonal block.
// The two 32-column chunks are software-pipelined: the TMEM load of the second and the TMEM store of the first overlap
// the exponentials (a TMEM round trip costs ~250 clk while the tensor core is busy).
// kPoly: which of every 8 column pairs take exp2 on the FMA pipe — the same pairs as in the two-kernel path
// (fa_bwd_sm100.cuh), so that dK / dV stay bit-identical between the two implementations.
template <bool kBf16, bool kMask, int kPoly>
__device__ __forceinline__ void fused_p_chunk(uint32_t st_saddr, uint64_t sl2_2, int row, int col0, uint32_t* pf,
                                              uint32_t (&pp)[16]) {
#pragma unroll
  for (int g4 = 0; g4 < 8; ++g4) {
    uint64_t nl4[2];
    lds_f32x2x2(st_saddr + g4 * 16, nl4[0], nl4[1]);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int e = g4 * 4 + u * 2;
      float x0, x1;
      f32x2_unpack(f32x2_fma(f32x2_pack_bits(pf[e], pf[e + 1]), sl2_2, nl4[u]), x0, x1);
      float p0, p1;
      if ((kPoly >> ((g4 * 2 + u) & 7)) & 1) {
        ex2_poly_x2(x0, x1, p0, p1);
      } else {
        p0 = ex2_approx(x0), p1 = ex2_approx(x1);
      }
      if constexpr (kMask) {   // keep key <= query: row = key, column = query
        const int c0 = col0 + e;
        if (row > c0) p0 = 0.f;
        if (row > c0 + 1) p1 = 0.f;
      }
      pf[e] = __float_as_uint(p0);
      pf[e + 1] = __float_as_uint(p1);
      pp[e >> 1] = pack2<kBf16>(p0, p1);
    }
  }
}
template <bool kBf16, bool kMask, int kPoly>
__device__ __forceinline__ void fused_p_stage(uint32_t tS, uint32_t st_saddr, uint64_t sl2_2, int row, int col0,
                                              uint32_t (&pf)[64]) {
  tmem_ld_x32(tS, *reinterpret_cast<uint32_t(*)[32]>(&pf[0]));
  tc_wait_ld();
  tmem_ld_x32(tS + 32, *reinterpret_cast<uint32_t(*)[32]>(&pf[32]));
  uint32_t pp[16];
  fused_p_chunk<kBf16, kMask, kPoly>(st_saddr, sl2_2, row, col0, &pf[0], pp);
  tc_wait_ld();
  tmem_st_x16(tS, pp);
  fused_p_chunk<kBf16, kMask, kPoly>(st_saddr + 128, sl2_2, row, col0 + 32, &pf[32], pp);
  tmem_st_x16(tS + 16, pp);
}

// dS^T stage: dS = P o (dP - delta[col]) for the thread's 64 columns, packed to 16 bits in `pd`.
template <bool kBf16>
__device__ __forceinline__ void fused_ds_chunk(uint32_t nd_saddr, const uint32_t* pf, const uint32_t (&dr)[32],
                                               uint32_t* pd) {
#pragma unroll
  for (int g4 = 0; g4 < 8; ++g4) {
    uint64_t nd4[2];
    lds_f32x2x2(nd_saddr + g4 * 16, nd4[0], nd4[1]);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int e = g4 * 4 + u * 2;
      float d0, d1;
      f32x2_unpack(f32x2_mul(f32x2_pack_bits(pf[e], pf[e + 1]), f32x2_add(f32x2_pack_bits(dr[e], dr[e + 1]), nd4[u])),
                   d0, d1);
      pd[e >> 1] = pack2<kBf16>(d0, d1);
    }
  }
}
template <bool kBf16>
__device__ __forceinline__ void fused_ds_stage(uint32_t tDP, uint32_t nd_saddr, const uint32_t (&pf)[64],
                                               uint32_t (&pd)[32]) {
  uint32_t d0[32], d1[32];
  tmem_ld_x32(tDP, d0);
  tc_wait_ld();
  tmem_ld_x32(tDP + 32, d1);
  fused_ds_chunk<kBf16>(nd_saddr, &pf[0], d0, &pd[0]);
  tc_wait_ld();
  fused_ds_chunk<kBf16>(nd_saddr + 128, &pf[32], d1, &pd[16]);
}

template <bool kBf16, int kD, bool kCausal>
__global__ void __launch_bounds__(512, 1)
fa_bwd_fused_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                    const FusedParams fp) {
  using Cfg = FusedCfg<kD>;
  const BwdParams& p = fp.base;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;                                   // stationary K_j
  uint8_t* sV = sK + Cfg::kTileBytes;                   // stationary V_j
  uint8_t* sQ = sV + Cfg::kTileBytes;                   // [2] streamed Q_i
#if FA_FUSED_TMA
  uint8_t* sDO = sQ + 2 * Cfg::kTileBytes;              // streamed dO_i (one stage)
  uint8_t* sDS = sDO + Cfg::kTileBytes;                 // dS^T of the current visit
  uint8_t* sStage = sDS + Cfg::kDsBytes;                // [2] staged dQ chunks
  float* sStat = reinterpret_cast<float*>(sStage + 2 * Cfg::kStageBytes);   // [2][2][128]: -lse, -delta
#else
  uint8_t* sDO = sQ + 2 * Cfg::kTileBytes;              // [2] streamed dO_i
  uint8_t* sDS = sDO + 2 * Cfg::kTileBytes;             // dS^T of the current visit
  float* sStat = reinterpret_cast<float*>(sDS + Cfg::kDsBytes);   // [2][2][128]: -lse, -delta
#endif
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStat + 2 * Cfg::kStatFloats);
  uint64_t* kv_full = bars + 0;
  uint64_t* acc_full = bars + 1;
  uint64_t* q_full = bars + 2;       // [2]
  uint64_t* stat_full = bars + 4;    // [2]
  uint64_t* q_empty = bars + 6;      // [2]
  uint64_t* s_full = bars + 8;
  uint64_t* p_ready = bars + 9;
  uint64_t* dp_full = bars + 10;
  uint64_t* ds_ready = bars + 11;
  uint64_t* ds_free = bars + 12;
  uint64_t* dq_full = bars + 13;
  uint64_t* dq_free = bars + 14;
#if FA_FUSED_TMA
  uint64_t* do_full = bars + 15;      // producer -> MMA: dO_v has landed
  uint64_t* do_empty = bars + 16;     // MMA -> producer: dV(v), the last reader of dO_v, has completed
  uint64_t* staged = bars + 17;       // [2] reducer -> warp 15: a chunk is in staging buffer g & 1
  uint64_t* stage_free = bars + 19;   // [2] warp 15 -> reducer: the TMA engine has read that buffer
  volatile int* turn_seen = reinterpret_cast<volatile int*>(bars + 21);   // warp 14 -> warp 15: chunks whose turn has come
#else
  uint64_t* turn_ok = bars + 15;      // [2] warp 14 -> reducer: it is this CTA's turn on the tile of visit v
  uint64_t* reds_out = bars + 17;     // [2] reducer -> warps 14 / 15: the reductions of visit v have been issued
  uint64_t* rel_done = bars + 19;     // [2] warp 15 -> reducer: the turn of visit v has been passed on
#endif
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + 22);
  int* ticket_s = reinterpret_cast<int*>(bars + 23);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
#if FA_TRACE == 2
  const long long t_entry = fa_globaltimer();
#endif

  if (threadIdx.x == 0) {
    if (static_cast<int>(smem - smem_raw) > Cfg::kAlignSlack) {
      printf("[fa_bwd_fused] dynamic shared memory base is not aligned as assumed (%d bytes of padding)\n",
             static_cast<int>(smem - smem_raw));
      __trap();
    }
    mbar_init(kv_full, 1);
    mbar_init(acc_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&stat_full[s], 32);
      mbar_init(&q_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 256);
    mbar_init(dp_full, 1);
    mbar_init(ds_ready, 256);
    mbar_init(ds_free, 1);
    mbar_init(dq_full, 1);
    mbar_init(dq_free, 128);
#if FA_FUSED_TMA
    mbar_init(do_full, 1);
    mbar_init(do_empty, 1);
    for (int t = 0; t < 2; ++t) {
      mbar_init(&staged[t], 128);
      mbar_init(&stage_free[t], 1);
    }
    *turn_seen = 0;
#else
    for (int t = 0; t < 2; ++t) {
      mbar_init(&turn_ok[t], 1);
      mbar_init(&reds_out[t], 128);
      mbar_init(&rel_done[t], 1);
    }
#endif
    fence_mbar_init();
    // Tickets are handed out in launch order: a CTA's predecessor in the dQ reduction always holds a smaller ticket.
    *ticket_s = atomicAdd(fp.ticket, 1);
  }
  if (warp == 12 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmDO);
  }
  if (warp == 13) tmem_alloc<512>(tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_base_s;

  const int n = fp.n_blocks;
  const int ticket = *ticket_s;
  const int bh = ticket / n;
  const int jb = n - 1 - (ticket - bh * n);   // lightest key block of a head first (its successors wait on it)
  const int b = bh / p.H, h = bh - b * p.H;
  const int k0 = jb * 128;
  const int n_vis = kCausal ? n - jb : n;
#if FA_TRACE == 2
  if (threadIdx.x == 0) {
    fa_cta_trace(ticket, 0, t_entry);
    FA_CTA_TRACE(1);
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    fa_cta_trace(ticket, 5, smid);
    fa_cta_trace(ticket, 6, n_vis);
  }
#endif
  auto q_block = [&](int v) {   // query block of visit v
    int i = jb + v;
    return i >= n ? i - n : i;
  };

  if (warp >= 12) {
    setmaxnreg_dec<64>();
    if (warp == 12) {
      // ------------------------------------------------------------------ producer: TMA + row statistics
      if (lane == 0) {
        mbar_arrive_expect_tx(kv_full, 2 * Cfg::kTileBytes);
        for (int bx = 0; bx < Cfg::kBoxes; ++bx) {
          tma_load_4d(sK + bx * Cfg::kBoxBytes, &tmK, kv_full, bx * 64, k0, h, b);
          tma_load_4d(sV + bx * Cfg::kBoxBytes, &tmV, kv_full, bx * 64, k0, h, b);
        }
      }
      const float* lsep = p.lse + (int64_t)bh * p.N;
      const float* dlp = p.delta + (int64_t)bh * p.N;
      for (int v = 0; v < n_vis; ++v) {
        const int s = v & 1;
        const int q0 = q_block(v) * 128;
        // row statistics of the query block: all eight loads of a lane in flight at once and before the ring wait (behind
        // generic stores to shared memory ptxas keeps them in program order: four dependent round trips per visit)
        float nl[4], nd[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int r = q0 + lane * 4 + e;
          const bool ok = r < p.N;
          nl[e] = ok ? -__ldg(lsep + r) : -INFINITY;   // query rows past N: P = exp2(-inf) = 0
          nd[e] = ok ? -__ldg(dlp + r) : 0.f;
        }
        mbar_wait(&q_empty[s], ((v >> 1) & 1) ^ 1);          // Q_{v-2}, dO_{v-2} consumed (dK(v-2) is their last reader)
#if FA_FUSED_TMA
        if (lane == 0) {
          mbar_arrive_expect_tx(&q_full[s], Cfg::kTileBytes);
          for (int bx = 0; bx < Cfg::kBoxes; ++bx)
            tma_load_4d(sQ + s * Cfg::kTileBytes + bx * Cfg::kBoxBytes, &tmQ, &q_full[s], bx * 64, q0, h, b);
        }
        mbar_wait(do_empty, (v & 1) ^ 1);                    // dV(v-1) has read dO_{v-1}
        if (lane == 0) {
          mbar_arrive_expect_tx(do_full, Cfg::kTileBytes);
          for (int bx = 0; bx < Cfg::kBoxes; ++bx)
            tma_load_4d(sDO + bx * Cfg::kBoxBytes, &tmDO, do_full, bx * 64, q0, h, b);
        }
#else
        if (lane == 0) {
          mbar_arrive_expect_tx(&q_full[s], 2 * Cfg::kTileBytes);
          for (int bx = 0; bx < Cfg::kBoxes; ++bx) {
            tma_load_4d(sQ + s * Cfg::kTileBytes + bx * Cfg::kBoxBytes, &tmQ, &q_full[s], bx * 64, q0, h, b);
            tma_load_4d(sDO + s * Cfg::kTileBytes + bx * Cfg::kBoxBytes, &tmDO, &q_full[s], bx * 64, q0, h, b);
          }
        }
#endif
        float4* st = reinterpret_cast<float4*>(sStat + s * Cfg::kStatFloats);
        st[lane] = make_float4(nl[0], nl[1], nl[2], nl[3]);
        st[32 + lane] = make_float4(nd[0], nd[1], nd[2], nd[3]);
        mbar_arrive(&stat_full[s]);
      }
    } else if (warp == 13) {
      // ------------------------------------------------------------------ MMA issuer
      if (elect_one()) {
        constexpr uint32_t idesc_sc = umma_idesc_f16(kBf16, 128, 128, 0, 0);  // [128 keys] x [128 queries], K = D
        constexpr uint32_t idesc_dv = umma_idesc_f16(kBf16, 128, kD, 0, 1);   // [128 keys] x [D], K = 128 queries
        constexpr uint32_t idesc_dq = umma_idesc_f16(kBf16, 128, kD, 1, 1);   // [128 queries] x [D], K = 128 keys
        constexpr uint32_t kTileLo = Cfg::kTileBytes >> 4;
        constexpr uint32_t kDoLo = FA_FUSED_TMA ? 0 : kTileLo;   // dO: one stage with the TMA egress
        const uint32_t k_lo = umma_lo_kmajor(smem_u32(sK)), v_lo = umma_lo_kmajor(smem_u32(sV));
        const uint32_t q_lo = umma_lo_kmajor(smem_u32(sQ)), do_lo = umma_lo_kmajor(smem_u32(sDO));
        const uint32_t q_mn = umma_lo_mnmajor(smem_u32(sQ), Cfg::kBoxBytes);
        const uint32_t do_mn = umma_lo_mnmajor(smem_u32(sDO), Cfg::kBoxBytes);
        const uint32_t k_mn = umma_lo_mnmajor(smem_u32(sK), Cfg::kBoxBytes);
        const uint32_t ds_mn = umma_lo_mnmajor(smem_u32(sDS), Cfg::kBoxBytes);
        const uint32_t ds_lo = umma_lo_kmajor(smem_u32(sDS));
        const uint32_t tS = tmem + Cfg::kTmemS, tDP = tmem + Cfg::kTmemDP;
        const uint32_t tDV = tmem + Cfg::kTmemDV, tDK = tmem + Cfg::kTmemDK;

        // S^T = K_j Q_i^T
        auto issue_s = [&](int s) {
          const uint32_t bq = q_lo + s * kTileLo;
          static_for<0, kD / 16>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            constexpr uint32_t off = umma_koff_kmajor(k, Cfg::kBoxBytes);
            umma_ss_off<off, off>(tS, k_lo, bq, idesc_sc, k > 0);
          });
          tc_commit(s_full);
        };
        // dP^T = V_j dO_i^T
        auto issue_dp = [&](int s) {
          const uint32_t bdo = do_lo + s * kDoLo;
          static_for<0, kD / 16>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            constexpr uint32_t off = umma_koff_kmajor(k, Cfg::kBoxBytes);
            umma_ss_off<off, off>(tDP, v_lo, bdo, idesc_sc, k > 0);
          });
          tc_commit(dp_full);
        };
        // dV += P^T dO_i : packed 16-bit P^T from TMEM, queries 0-63 in columns [0,32) of the S region, 64-127 in [64,96)
        auto issue_dv = [&](int s, bool first) {
          const uint32_t bdo = do_mn + s * kDoLo;
          static_for<0, 8>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            umma_ts_off<(k & 3) * 8 + (k >> 2) * 64, umma_koff_mnmajor(k)>(tDV, tS, bdo, idesc_dv, !(first && k == 0));
          });
        };
        // dQ_i partial = dS K_j  (both operands indexed by key row in shared memory: MN-major A and B); it overwrites
        // the dP region, which the elementwise warps have finished reading
        auto issue_dq = [&]() {
          static_for<0, 8>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            umma_ss_off<umma_koff_mnmajor(k), umma_koff_mnmajor(k)>(tDP, ds_mn, k_mn, idesc_dq, k > 0);
          });
          tc_commit(dq_full);
        };
        // dK += dS^T Q_i : dS^T from shared memory as the K-major A operand (rows = keys, 64 queries per box)
        auto issue_dk = [&](int s, bool first) {
          const uint32_t bq = q_mn + s * kTileLo;
          static_for<0, 8>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            umma_ss_off<umma_koff_kmajor(k, Cfg::kBoxBytes), umma_koff_mnmajor(k)>(tDK, ds_lo, bq, idesc_dv,
                                                                                  !(first && k == 0));
          });
        };

        mbar_wait(kv_full, 0);
        mbar_wait(&q_full[0], 0);
#if FA_FUSED_TMA
        mbar_wait(do_full, 0);
#endif
        tc_fence_after();
        issue_s(0);
        issue_dp(0);
        for (int v = 0; v < n_vis; ++v) {
          const int s = v & 1;
          const bool more = v + 1 < n_vis;
          fa_trace(0, v, 0);
          mbar_wait(p_ready, v & 1);
          fa_trace(0, v, 1);
          if (v == 0) { FA_CTA_TRACE(2); }
          tc_fence_after();
          issue_dv(s, v == 0);
#if FA_FUSED_TMA
          tc_commit(do_empty);
#endif
          if (more) {
            mbar_wait(&q_full[s ^ 1], ((v + 1) >> 1) & 1);
            tc_fence_after();
            issue_s(s ^ 1);
          }
          fa_trace(0, v, 2);
          mbar_wait(ds_ready, v & 1);
          fa_trace(0, v, 3);
          tc_fence_after();
          issue_dq();
          issue_dk(s, v == 0);
          tc_commit(&q_empty[s]);
          tc_commit(ds_free);
          fa_trace(0, v, 4);
          if (more) {
            mbar_wait(dq_free, v & 1);   // the reducer has copied dQ(v) out of TMEM
            fa_trace(0, v, 5);
#if FA_FUSED_TMA
            mbar_wait(do_full, (v + 1) & 1);
#endif
            tc_fence_after();
            issue_dp(s ^ 1);             // Q_{v+1} / dO_{v+1} arrived before S(v+1) was issued
          }
        }
        tc_commit(acc_full);
        FA_CTA_TRACE(3);
      }
      __syncwarp();
    }
#if FA_FUSED_TMA
    else if (lane == 0) {
      // ------------------------------------------------------------------ egress warps (14: turn poll, 15: TMA issue)
      const int64_t tiles_bh = (int64_t)bh * n;
      constexpr int C = Cfg::kChunks;
      if (warp == 14) {
#if !FA_FUSED_UNORDERED
        // Polls the turn counters chunk by chunk, ahead of warp 15, and publishes how far the turns have come.
        int g = 0;
        for (int v = 0; v < n_vis; ++v) {
          const int i = q_block(v);
          const bool buf_b = !kCausal && i < jb;
          const int rank = buf_b ? n - 1 - jb : v;
          const int* sem = fp.sem + ((buf_b ? (int64_t)p.B * p.H * n : 0) + tiles_bh + i) * kSemPerTile;
          for (int c = 0; c < C; ++c, ++g) {
            if (rank != 0) {
#if FA_WATCHDOG
              long long t0 = clock64();
#endif
              while (ld_acquire_gpu(sem + c) != rank) {
#if FA_WATCHDOG
                if (clock64() - t0 > FA_WATCHDOG_CYCLES) {
                  printf("[fa watchdog] dQ turn wait timed out: ticket %d bh %d j %d i %d chunk %d rank %d sem %d\n", ticket,
                         bh, jb, i, c, rank, ld_acquire_gpu(sem + c));
                  __trap();
                }
#endif
              }
            }
            __threadfence_block();
            *turn_seen = g + 1;
          }
        }
#endif
      } else {
        // Issues one bulk reduction per staged chunk, hands the staging buffer back when the engine has read it, and
        // passes the chunk's turn on when the reduction has completed.  Up to two chunks stay in flight; whenever the next
        // chunk is not staged yet everything outstanding is completed and released first.
        int g = 0, freed = 0, n_pend = 0;
        int* pend_sem[3];
        int pend_rank[3];
        auto flush = [&](auto keep) {   // complete all but the newest `keep` groups, release their turns in order
          constexpr int kKeep = decltype(keep)::value;
          tma_store_wait<kKeep>();
#if !FA_FUSED_UNORDERED
          while (n_pend > kKeep) {
            st_release_gpu(pend_sem[0], pend_rank[0] + 1);
            pend_sem[0] = pend_sem[1], pend_rank[0] = pend_rank[1];
            pend_sem[1] = pend_sem[2], pend_rank[1] = pend_rank[2];
            --n_pend;
          }
#endif
        };
        for (int v = 0; v < (FA_FUSED_ABLATE == 2 ? 0 : n_vis); ++v) {
          const int i = q_block(v);
          const bool buf_b = !kCausal && i < jb;
          const int rank = buf_b ? n - 1 - jb : v;
          const int64_t t = (buf_b ? (int64_t)p.B * p.H * n : 0) + tiles_bh + i;
          int* sem = fp.sem + t * kSemPerTile;
          float* tile = fp.dq_acc + (buf_b ? fp.acc_b_off : 0) + (tiles_bh + i) * (int64_t)(128 * kD);
          for (int c = 0; c < C; ++c, ++g) {
            const int buf = g & 1;
            if (!mbar_try_wait(&staged[buf], (g >> 1) & 1)) {
              flush(std::integral_constant<int, 0>{});
              while (freed < g) { mbar_arrive(&stage_free[freed & 1]); ++freed; }
              mbar_wait(&staged[buf], (g >> 1) & 1);
            }
#if !FA_FUSED_UNORDERED
            while (*turn_seen <= g) {}
            __threadfence_block();
#endif
            asm volatile("fence.proxy.async;" ::: "memory");
            const uint32_t src = smem_u32(sStage + buf * Cfg::kStageBytes);
            float* dst = tile + c * (Cfg::kStageBytes / 4);
#if FA_FUSED_UNORDERED
            const bool plain = false;   // the host zero-fills the workspace
#else
            const bool plain = rank == 0;
#endif
            if (plain)
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src),
                           "n"(Cfg::kStageBytes)
                           : "memory");
            else
              asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst),
                           "r"(src), "n"(Cfg::kStageBytes)
                           : "memory");
            tma_store_commit();
#if !FA_FUSED_UNORDERED
            pend_sem[n_pend] = sem + c, pend_rank[n_pend] = rank, ++n_pend;
#endif
            tma_store_wait_read<1>();     // the engine has read chunk g - 1's buffer
            while (freed < g) { mbar_arrive(&stage_free[freed & 1]); ++freed; }
#if !FA_FUSED_UNORDERED
            if (n_pend == 3) flush(std::integral_constant<int, 2>{});
#endif
          }
        }
        flush(std::integral_constant<int, 0>{});
      }
    }
#elif FA_FUSED_TURN_WARPS && !FA_FUSED_UNORDERED
    else if (lane == 0) {
      // ------------------------------------------------------------------ turn warps (14: acquire, 15: release)
      const int64_t tiles_bh = (int64_t)bh * n;
      for (int v = 0; v < n_vis; ++v) {
        const int i = q_block(v);
        const bool buf_b = !kCausal && i < jb;
        const int rank = buf_b ? n - 1 - jb : v;
        int* sem = fp.sem + (buf_b ? (int64_t)p.B * p.H * n : 0) + tiles_bh + i;
        if (warp == 14) {
          // slot v & 1 of turn_ok is free once the reducer has passed visit v - 2 (it waited on it before its reductions)
          if (v >= 2) mbar_wait(&reds_out[v & 1], ((v - 2) >> 1) & 1);
#if FA_FUSED_ABLATE == 6   // timing experiment (order NOT guaranteed): turns are published but nobody waits for them
          if (false) {
#else
          if (rank != 0) {
#endif
#if FA_WATCHDOG
            long long t0 = clock64();
#endif
            while (ld_acquire_gpu(sem) != rank) {
#if FA_WATCHDOG
              if (clock64() - t0 > FA_WATCHDOG_CYCLES) {
                printf("[fa watchdog] dQ turn wait timed out: ticket %d bh %d j %d i %d rank %d sem %d\n", ticket, bh, jb,
                       i, rank, ld_acquire_gpu(sem));
                __trap();
              }
#endif
            }
          }
          mbar_arrive(&turn_ok[v & 1]);
        } else {
          // every reducer thread arrived (release, CTA scope) after issuing its reductions; the gpu-scope release below
          // is cumulative over them: the next contributor's acquire of `rank + 1` orders its reductions after ours
          mbar_wait(&reds_out[v & 1], (v >> 1) & 1);
#if FA_FUSED_ABLATE == 5   // timing experiment (order NOT guaranteed): publish the turn without the release fence
          asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(sem), "r"(rank + 1) : "memory");
#else
          st_release_gpu(sem, rank + 1);
#endif
          mbar_arrive(&rel_done[v & 1]);
        }
      }
    }
#endif
  } else if (warp >= 8) {
    setmaxnreg_inc<160>();
    // ------------------------------------------------------------------ dQ reducer (warps 8-11)
    // The SM -> L2 path moves ~25 B/clk, so the 64 KiB partial of one visit takes about as long to leave the SM as the
    // visit takes to compute.  The whole tile is therefore pulled into registers at once (TMEM is handed back to the
    // tensor core after ~300 clk) and trickles out as fire-and-forget vector reductions while the next visit runs.
    const int rt = threadIdx.x - 256;
    const int row = (warp & 3) * 32 + lane;   // query row inside the block == TMEM lane
    const uint32_t tDQ = tmem + Cfg::kTmemDP + ((uint32_t)((warp & 3) * 32) << 16);
    const int64_t tiles_bh = (int64_t)bh * n;
    for (int v = 0; v < n_vis; ++v) {
      const int i = q_block(v);
      const bool buf_b = !kCausal && i < jb;
      const int rank = buf_b ? n - 1 - jb : v;           // position of this CTA in the tile's fixed order
      const bool first = rank == 0;
      float* tile = fp.dq_acc + (buf_b ? fp.acc_b_off : 0) + (tiles_bh + i) * (int64_t)(128 * kD) + row * 4;
      int* sem = fp.sem + (buf_b ? (int64_t)p.B * p.H * n : 0) + tiles_bh + i;
      if (rt == 0) fa_trace(3, v, 0);
      mbar_wait(dq_full, v & 1);
      if (rt == 0) fa_trace(3, v, 1);
      tc_fence_after();
      uint32_t r[kD];
#pragma unroll
      for (int c = 0; c < kD / 32; ++c) tmem_ld_x32(tDQ + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&r[c * 32]));
      tc_wait_ld();
      tc_fence_before();
      mbar_arrive(dq_free);
      if (rt == 0) fa_trace(3, v, 2);
#if FA_FUSED_TMA
#if FA_FUSED_ABLATE != 2
#pragma unroll
      for (int c = 0; c < Cfg::kChunks; ++c) {
        const int g = v * Cfg::kChunks + c, buf = g & 1;
        if (g >= 2) mbar_wait(&stage_free[buf], ((g >> 1) - 1) & 1);
        const uint32_t dst = smem_u32(sStage + buf * Cfg::kStageBytes) + row * 16;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst + k * 2048), "r"(r[c * 32 + 4 * k]),
                       "r"(r[c * 32 + 4 * k + 1]), "r"(r[c * 32 + 4 * k + 2]), "r"(r[c * 32 + 4 * k + 3])
                       : "memory");
        fence_proxy_async_smem();
        mbar_arrive(&staged[buf]);
      }
#endif
      if (rt == 0) fa_trace(3, v, 3);
#else   // !FA_FUSED_TMA: reductions issued by the reducer threads themselves
#if FA_FUSED_TURN_WARPS && !FA_FUSED_UNORDERED
      mbar_wait(&turn_ok[v & 1], (v >> 1) & 1);
#elif !FA_FUSED_UNORDERED
      if (!first) {
        if (rt == 0) {
#if FA_WATCHDOG
          long long t0 = clock64();
#endif
          while (ld_acquire_gpu(sem) != rank) {
#if FA_WATCHDOG
            if (clock64() - t0 > FA_WATCHDOG_CYCLES) {
              printf("[fa watchdog] dQ turn wait timed out: ticket %d bh %d j %d i %d rank %d sem %d\n", ticket, bh, jb,
                     i, rank, ld_acquire_gpu(sem));
              __trap();
            }
#endif
          }
        }
        named_bar_sync(1, 128);
      }
#endif
#if FA_FUSED_ABLATE != 2
      constexpr int kOutVec = FA_FUSED_HALF ? kD / 8 : kD / 4;   // 16-byte vectors of the row that leave the SM
      // Paced: a burst of 64 KiB would monopolise the SM's path to L2 (~25 B/clk) and hold up the TMA requests of the
      // next Q / dO tiles; one 512-byte warp instruction every FA_FUSED_PACE ns per warp keeps the path shared.
      if (first) {
#pragma unroll
        for (int k = 0; k < kOutVec; ++k) {
          st_f32x4(tile + k * 512, r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
          if (FA_FUSED_PACE) __nanosleep(FA_FUSED_PACE);
        }
      } else {
#pragma unroll
        for (int k = 0; k < kOutVec; ++k) {
          red_add_f32x4(tile + k * 512, r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
          if (FA_FUSED_PACE) __nanosleep(FA_FUSED_PACE);
        }
      }
#endif
#if FA_FUSED_TURN_WARPS && !FA_FUSED_UNORDERED
      // the turn is passed on by warp 15 once all 128 reducer threads have issued their reductions; slot v & 1 of
      // reds_out is reusable because warp 15 has consumed visit v - 2 (rel_done)
      if (v >= 2) mbar_wait(&rel_done[v & 1], ((v - 2) >> 1) & 1);
      mbar_arrive(&reds_out[v & 1]);
#elif !FA_FUSED_UNORDERED
      __threadfence();
      named_bar_sync(1, 128);
      if (rt == 0) st_release_gpu(sem, rank + 1);
#endif
      if (rt == 0) fa_trace(3, v, 3);
#endif  // FA_FUSED_TMA
    }
  } else {
    setmaxnreg_inc<144>();
    // ------------------------------------------------------------------ elementwise: P^T, dS^T  (warps 0-7)
    const int half = warp >> 2;
    const int row = (warp & 3) * 32 + lane;   // key row inside the block == TMEM lane
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + Cfg::kTmemS + half * 64 + lane_base;
    const uint32_t tDP = tmem + Cfg::kTmemDP + half * 64 + lane_base;
    const uint32_t sds = smem_u32(sDS) + half * Cfg::kBoxBytes;
    const float sl2 = p.scale_log2;
    const uint64_t sl2_2 = f32x2_pack(sl2, sl2);

    for (int v = 0; v < n_vis; ++v) {
      const int s = v & 1;
      const uint32_t st = smem_u32(sStat + s * Cfg::kStatFloats + half * 64);
      const bool diag = kCausal && v == 0;   // query block == key block: the only block that needs the causal mask
      const bool tr = (threadIdx.x & 127) == 0;
      if (tr) fa_trace(1 + half, v, 0);
      mbar_wait(&stat_full[s], (v >> 1) & 1);
      mbar_wait(s_full, v & 1);
      if (tr) fa_trace(1 + half, v, 1);
      tc_fence_after();
      // ---- P^T = exp2(S^T * scale*log2e - L[query]) : fp32 copy kept in registers for dS, 16-bit copy to TMEM
      uint32_t pf[64];
      constexpr int kPolyMask = kD == 64 ? FA_BWD_POLY_MASK_D64 : FA_BWD_POLY_MASK;
      if (diag)
        fused_p_stage<kBf16, true, kPolyMask>(tS, st, sl2_2, row, half * 64, pf);
      else
        fused_p_stage<kBf16, false, kPolyMask>(tS, st, sl2_2, row, half * 64, pf);
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(p_ready);
      if (tr) fa_trace(1 + half, v, 2);

      // ---- dS^T = P^T o (dP^T - delta[query]) -> shared memory (row = key, 64 queries = 128 swizzled bytes)
      mbar_wait(dp_full, v & 1);
      if (tr) fa_trace(1 + half, v, 3);
      tc_fence_after();
      uint32_t pd[32];
      fused_ds_stage<kBf16>(tDP, st + 512, pf, pd);
      tc_fence_before();
      if (tr) fa_trace(1 + half, v, 4);
      if (v > 0) mbar_wait(ds_free, (v - 1) & 1);             // dQ(v-1) and dK(v-1) have read the previous dS^T
#pragma unroll
      for (int ch = 0; ch < 8; ++ch)
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sds + sw128_offset(row, ch)), "r"(pd[ch * 4]),
                     "r"(pd[ch * 4 + 1]), "r"(pd[ch * 4 + 2]), "r"(pd[ch * 4 + 3])
                     : "memory");
      fence_proxy_async_smem();
      mbar_arrive(ds_ready);
      if (tr) fa_trace(1 + half, v, 5);
    }

    // epilogue: warpgroup a stores dV, warpgroup b stores scale * dK
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int kv_row = k0 + row;
    const bool in_range = kv_row < p.N;
    if (half == 0) {
      uint16_t* dst = reinterpret_cast<uint16_t*>(p.dv) + b * p.dv_s[0] + h * p.dv_s[1] + (int64_t)kv_row * p.dv_s[2];
      store_acc_rows<kBf16>(tmem + Cfg::kTmemDV + lane_base, kD, 1.0f, dst, in_range);
    } else {
      uint16_t* dst = reinterpret_cast<uint16_t*>(p.dk) + b * p.dk_s[0] + h * p.dk_s[1] + (int64_t)kv_row * p.dk_s[2];
      store_acc_rows<kBf16>(tmem + Cfg::kTmemDK + lane_base, kD, p.scale, dst, in_range);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 13) tmem_dealloc<512>(tmem);
  if (threadIdx.x == 0) { FA_CTA_TRACE(4); }
}

// dQ = scale * (A [+ B]) : fp32 workspace tiles ([D/4][128 rows][4]) -> 16-bit (B,H,N,D).  HBM-bound.
template <bool kBf16, int kD>
__global__ void __launch_bounds__(256)
fa_bwd_dq_convert_kernel(const float* __restrict__ acc, long long acc_b_off, void* dq, int64_t sB, int64_t sH,
                         int64_t sN, int H, int N, int n_blocks, float scale) {
  constexpr int kRowHalfs = kD + 4;   // 8-byte aligned rows, conflict-free 8-byte column-order stores
  __shared__ __align__(16) uint16_t tile16[128 * kRowHalfs];
  const int i = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int64_t t = ((int64_t)b * H + h) * n_blocks + i;
  const float4* a = reinterpret_cast<const float4*>(acc + t * (128 * kD));
  const bool has_b = acc_b_off != 0 && i < n_blocks - 1;   // the last query block has no key block after it
  const float4* bb = reinterpret_cast<const float4*>(acc + acc_b_off + t * (128 * kD));
#pragma unroll 4
  for (int idx = threadIdx.x; idx < 128 * kD / 4; idx += 256) {
    const int c = idx >> 7, r = idx & 127;
    float4 x = __ldcs(a + idx);
    if (has_b) {
      const float4 y = __ldcs(bb + idx);
      x.x += y.x, x.y += y.y, x.z += y.z, x.w += y.w;
    }
    uint2 o;
    o.x = pack2<kBf16>(x.x * scale, x.y * scale);
    o.y = pack2<kBf16>(x.z * scale, x.w * scale);
    *reinterpret_cast<uint2*>(&tile16[r * kRowHalfs + c * 4]) = o;
  }
  __syncthreads();
  uint16_t* dst = reinterpret_cast<uint16_t*>(dq) + b * sB + h * sH;
  constexpr int kChunks = kD / 4;   // 8-byte chunks per row
  for (int idx = threadIdx.x; idx < 128 * kChunks; idx += 256) {
    const int r = idx / kChunks, ch = idx - r * kChunks;
    const int q = i * 128 + r;
    if (q < N)
      *reinterpret_cast<uint2*>(dst + (int64_t)q * sN + ch * 4) =
          *reinterpret_cast<const uint2*>(&tile16[r * kRowHalfs + ch * 4]);
  }
}

}  // namespace fa
