// fa_fwd2_sm100.cuh — FlashAttention-2 forward for sm_100a on CTA PAIRS (tcgen05 cta_group::2), D = 128, 16-bit inputs.
//
// Same math and the same warp roles as fa_fwd_sm100.cuh (reference semantics: flash_attention_kernels.py:88-108 with the
// scale / causal mask of flash_attention_openai_tutorial.py:50,160-161).  What changes is who feeds the tensor cores.
// The single-CTA kernel is bound by shared-memory bandwidth, not by its softmax stage: per key block it reads
// 2 x (Q_t 32 KiB + K_j 32 KiB) for the score MMAs and 2 x V_j 32 KiB for the P.V MMAs and receives 64 KiB of K_j / V_j
// from TMA — 256 KiB per 2048 tensor clocks = 125 B/clk of the 128 B/clk an SM has (with every elementwise stage
// compiled out it still runs at 0.77 of the tensor peak; tools/umma2_probe.cu measures the operand streams alone).
// Here two CTAs of a cluster (one 512-row query block of one head, four 128-row tiles) issue ONE M = 256 MMA stream:
// each CTA supplies its own Q_t / P_t rows (A operand) and only HALF of K_j (64 key rows, N split of the score MMA) and
// HALF of V_j (64 of the D columns, N split of the P.V MMA); the hardware shares the halves across the pair.  Per CTA
// and key block: 2 x (32 + 16) + 2 x 16 + 32 KiB of fill = 160 KiB, 78 B/clk, and the K / V ring is four stages deep
// in the same 128 KiB.
//
//   CTA c (cluster rank) of quad p owns tiles g = 4p + 2t + c, t = 0, 1 (slot t pairs tiles 4p+2t and 4p+2t+1, so the two
//   halves of one MMA differ by at most one key block under the causal mask)
//   warps 0-3 / 4-7  softmax of slot 0 / 1 (thread = query row = TMEM lane), epilogue
//   warp 8           TMA producer: own Q tiles, own halves of K_j / V_j; completion bytes go to the LEADER's barriers
//   warp 9           leader (rank 0): issues every MMA for both CTAs; both: TMEM allocation (cta_group::2)
//   MMA -> softmax / producer hand-overs are tcgen05.commit multicasts to both CTAs; softmax -> MMA hand-overs are
//   arrivals of one elected lane per warp on the leader's barrier (through the cluster address space).
#pragma once

#include "fa_fwd_sm100.cuh"

namespace fa {

// FA_TRACE builds: the heaviest pair of head (0, 0) records clock64 per role / block / event (role = 3 * rank + {0: MMA
// thread, 1: softmax slot 0, 2: softmax slot 1}); clocks of the two SMs are not comparable with each other.
#if FA_TRACE
__device__ __forceinline__ void fa_trace2(int role, int it, int k) {
  if (g_fa_trace != nullptr && (blockIdx.x >> 1) == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
    const int slot = (role + 3 * (int)(blockIdx.x & 1)) * 8192 + it * 8 + k;
    if (slot < g_fa_trace_cap) g_fa_trace[slot] = clock64();
  }
}
#else
__device__ __forceinline__ void fa_trace2(int, int, int) {}
#endif

template <int kD>
struct Fwd2Cfg {
  static_assert(kD == 128, "the CTA-pair forward is instantiated for D = 128");
  static constexpr int kStages = 4;
  static constexpr int kTileBytes = 128 * kD * 2;        // one 128-row Q tile
  static constexpr int kBoxBytes = 128 * 128;            // one 64-column box of it
  static constexpr int kKHalfBytes = 64 * kD * 2;        // 64 key rows x D: two boxes of 64 rows
  static constexpr int kKBoxBytes = 64 * 128;
  static constexpr int kVHalfBytes = 128 * 64 * 2;       // 128 key rows x 64 of the D columns: one box
  static constexpr int kSmemQ = 2 * kTileBytes;
  static constexpr int kSmemBytes = kSmemQ + kStages * (kKHalfBytes + kVHalfBytes) + 1024 /*alignment slack*/;
  static constexpr int kThreads = 384;
  static constexpr uint32_t kTmemS0 = 0, kTmemS1 = 128, kTmemO0 = 256, kTmemO1 = 256 + kD;
};

struct Fwd2Maps {
  CUtensorMap q;     // box 64 x 128 rows
  CUtensorMap k64;   // box 64 x 64 rows
  CUtensorMap v;     // box 64 x 128 rows
};

template <bool kBf16, int kD, bool kCausal>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
fa_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK64,
               const __grid_constant__ CUtensorMap tmV, const FwdParams p) {
  using Cfg = Fwd2Cfg<kD>;
  constexpr int NS = Cfg::kStages;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                  // [2][tile]
  uint8_t* sK = smem + Cfg::kSmemQ;                    // [NS][64 key rows x D]
  uint8_t* sV = sK + NS * Cfg::kKHalfBytes;            // [NS][128 key rows x 64 columns]

  __shared__ uint64_t q_full[2], s_full[2], p_full[2][2], o_full[2];
  __shared__ uint64_t k_full[NS], k_empty[NS], v_full[NS], v_empty[NS];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();             // 0 = leader

  // heaviest (largest q index) quads first so the causal triangle load-balances
  const int quad = p.q_blocks - 1 - (int)(blockIdx.x >> 1);   // q_blocks = number of 512-row quads here
  const int h = blockIdx.y, b = blockIdx.z;
  const int nv = p.seqlens ? min(max(p.seqlens[b], 0), p.N) : p.N;   // valid length of this batch element
  if (quad * 512 >= nv) return;                        // the whole cluster is padding (same decision in both CTAs)
  const int n_kv_total = (nv + 127) >> 7;
  // tiles: g(c, t) = 4 quad + 2 t + c.  own = this CTA's tile of slot t; slot = what the pair's MMA covers
  int g_own[2], nkv_own[2], nkv_slot[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    g_own[t] = 4 * quad + 2 * t + (int)rank;
    const int g_hi = 4 * quad + 2 * t + 1, g_lo = 4 * quad + 2 * t;
    auto blocks_of = [&](int g) { return g * 128 >= nv ? 0 : (kCausal ? min(n_kv_total, g + 1) : n_kv_total); };
    nkv_own[t] = blocks_of(g_own[t]);
    nkv_slot[t] = max(blocks_of(g_lo), blocks_of(g_hi));
  }
  const int nslots = nkv_slot[1] > 0 ? 2 : 1;
  const int nkv_max = max(nkv_slot[0], nkv_slot[1]);

  if (threadIdx.x == 0) {
    for (int t = 0; t < 2; ++t) {
      mbar_init(&q_full[t], 1);
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t][0], 8);    // one elected lane per softmax warp, 4 warps x 2 CTAs
      mbar_init(&p_full[t][1], 8);
      mbar_init(&o_full[t], 1);
    }
    for (int s = 0; s < NS; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK64);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 9) tmem_alloc2<512>(&tmem_base_s);
  tc_fence_before();
  cluster_sync_all();     // barriers of both CTAs initialised, both TMEM allocations done
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp >= 8) {
  setmaxnreg_dec<72>();   // third warpgroup: producer, MMA issuer, two idle warps
  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one()) {
      for (int t = 0; t < nslots; ++t) {
        if (rank == 0) mbar_arrive_expect_tx(&q_full[t], 2 * Cfg::kTileBytes);   // both CTAs' tiles
        const uint32_t bar = mapa_u32(smem_u32(&q_full[t]), 0);
        for (int bx = 0; bx < 2; ++bx)
          tma_load_4d_2sm(sQ + t * Cfg::kTileBytes + bx * Cfg::kBoxBytes, &tmQ, bar, bx * 64, g_own[t] * 128, h, b);
      }
      for (int j = 0; j < nkv_max; ++j) {
        const int s = j % NS;
        const uint32_t ph = (j / NS) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        if (rank == 0) mbar_arrive_expect_tx(&k_full[s], 2 * Cfg::kKHalfBytes);
        const uint32_t kbar = mapa_u32(smem_u32(&k_full[s]), 0);
        for (int bx = 0; bx < 2; ++bx)   // key rows [128 j + 64 rank, +64): this CTA's half of the score MMA's N
          tma_load_4d_2sm(sK + s * Cfg::kKHalfBytes + bx * Cfg::kKBoxBytes, &tmK64, kbar, bx * 64, j * 128 + 64 * (int)rank, h, b);
        mbar_wait(&v_empty[s], ph ^ 1);
        if (rank == 0) mbar_arrive_expect_tx(&v_full[s], 2 * Cfg::kVHalfBytes);
        const uint32_t vbar = mapa_u32(smem_u32(&v_full[s]), 0);
        // columns [64 rank, +64) of V_j: this CTA's half of the P.V MMA's N
        tma_load_4d_2sm(sV + s * Cfg::kVHalfBytes, &tmV, vbar, 64 * (int)rank, j * 128, h, b);
      }
    }
    __syncwarp();
  } else if (warp == 9 && rank == 0) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA, one thread)
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_f16(kBf16, 256, 128, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_f16(kBf16, 256, kD, 0, 1);
      const uint32_t qlo = umma_lo_kmajor(smem_u32(sQ)), klo = umma_lo_kmajor(smem_u32(sK));
      const uint32_t vlo = umma_lo_mnmajor(smem_u32(sV), Cfg::kBoxBytes);   // one box per CTA: LBO unused
      constexpr uint32_t kTileLo = Cfg::kTileBytes >> 4, kKLo = Cfg::kKHalfBytes >> 4, kVLo = Cfg::kVHalfBytes >> 4;
      auto tS = [&](int t) { return tmem + (t ? Cfg::kTmemS1 : Cfg::kTmemS0); };
      auto tO = [&](int t) { return tmem + (t ? Cfg::kTmemO1 : Cfg::kTmemO0); };

      auto issue_s = [&](int t, int j) {   // S of slot t (both CTAs' tiles) for key block j
        const int s = j % NS;
        const uint32_t a0 = qlo + t * kTileLo, b0 = klo + s * kKLo, d0 = tS(t);
        mbar_wait(&k_full[s], (j / NS) & 1);
        tc_fence_after();
        static_for<0, kD / 16>([&](auto kc) {
          constexpr int k = decltype(kc)::value;
          umma2_ss_off<umma_koff_kmajor(k, Cfg::kBoxBytes), umma_koff_kmajor(k, Cfg::kKBoxBytes)>(d0, a0, b0, idesc_s, k > 0);
        });
        tc_commit2(&s_full[t]);
        const bool last_user = (t == 1) || (nkv_slot[1] <= j);   // last slot that reads K block j releases the stage
        if (last_user) tc_commit2(&k_empty[s]);
      };

      for (int t = 0; t < nslots; ++t) {
        mbar_wait(&q_full[t], 0);
        issue_s(t, 0);
      }
      for (int j = 0; j < nkv_max; ++j) {
        const int s = j % NS;
        for (int t = 0; t < nslots; ++t) {
          if (j >= nkv_slot[t]) continue;
          mbar_wait(&v_full[s], (j / NS) & 1);
          const uint32_t dO_t = tO(t), aP = tS(t), bV = vlo + s * kVLo;
          const bool acc0 = j > 0;
          fa_trace2(0, j, 4 * t);
          mbar_wait(&p_full[t][0], j & 1);   // P arrives in two 64-key halves (both CTAs)
          fa_trace2(0, j, 4 * t + 1);
          tc_fence_after();
          static_for<0, 4>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            umma2_ts_off<k * 8, umma_koff_mnmajor(k)>(dO_t, aP, bV, idesc_o, acc0 || (k > 0));
          });
          mbar_wait(&p_full[t][1], j & 1);
          fa_trace2(0, j, 4 * t + 2);
          tc_fence_after();
          static_for<4, 8>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            umma2_ts_off<k * 8, umma_koff_mnmajor(k)>(dO_t, aP, bV, idesc_o, 1u);
          });
          tc_commit2(&o_full[t]);
          const bool last_user = (t == 1) || (nkv_slot[1] <= j);
          if (last_user) tc_commit2(&v_empty[s]);
          if (j + 1 < nkv_slot[t]) issue_s(t, j + 1);
          fa_trace2(0, j, 4 * t + 3);
        }
      }
    }
    __syncwarp();
  }
  } else {
    setmaxnreg_inc<216>();
    // ------------------------------------------------------------------ softmax + epilogue (warps 0-7, both CTAs)
    const int t = warp >> 2;
    const int row = (warp & 3) * 32 + lane;           // row inside the tile == TMEM lane
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + (t ? Cfg::kTmemS1 : Cfg::kTmemS0) + lane_base;
    const uint32_t tO = tmem + (t ? Cfg::kTmemO1 : Cfg::kTmemO0) + lane_base;
    const int my_own = nkv_own[t], my_slot = nkv_slot[t];
    const int q0t = g_own[t] * 128;                   // first query row of this tile
    const int q_row = q0t + row;                      // global query index
    const float sl2 = p.scale_log2;
    // the leader's P barriers, through the cluster address space (rank 0: its own)
    const uint32_t pbar0 = mapa_u32(smem_u32(&p_full[t][0]), 0), pbar1 = mapa_u32(smem_u32(&p_full[t][1]), 0);
    auto p_arrive = [&](uint32_t bar) {   // this warp's P columns are in TMEM: one arrival per warp
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(bar);
    };

    float m_used = -INFINITY, l = 0.f;
    const bool tr = (threadIdx.x & 127) == 0;
    for (int j = 0; j < my_slot; ++j) {
      if (tr) fa_trace2(1 + t, j, 0);
      mbar_wait(&s_full[t], j & 1);
      if (tr) fa_trace2(1 + t, j, 1);
      tc_fence_after();
      if (j >= my_own) {
        // The pair's other tile still needs this key block; for this tile it is entirely above the diagonal (or the tile
        // is padding): P = 0, so that the pair's P.V MMA adds nothing to this CTA's rows.
        uint32_t z[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) z[i] = 0u;
        tmem_st_x32(tS, z);
        p_arrive(pbar0);
        tmem_st_x32(tS + 32, z);
        p_arrive(pbar1);
        continue;
      }
      uint32_t sr[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_x32(tS + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&sr[c * 32]));
      tc_wait_ld();

      const int kv0 = j * 128;
      const bool diag = kCausal && (kv0 + 127 > q0t);   // block touches the diagonal
      const bool ragged = (kv0 + 128 > nv);
      if (diag || ragged) {
        int limit = nv - kv0;                          // first invalid column (ragged / padded keys)
        if (kCausal) limit = min(limit, q_row - kv0 + 1);
#pragma unroll
        for (int c = 0; c < 128; ++c)
          if (c >= limit) sr[c] = 0xff800000u;   // -inf
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int c = 0; c < 128; c += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(sr[c]));
        mx1 = fmaxf(mx1, __uint_as_float(sr[c + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(sr[c + 2]));
        mx3 = fmaxf(mx3, __uint_as_float(sr[c + 3]));
      }
      const float m_new = fmaxf(fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)), m_used);
      if (j == 0) {
        m_used = m_new;
      } else {
        const bool need = (m_new - m_used) * sl2 > 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          const float alpha = need ? ex2_approx((m_used - m_new) * sl2) : 1.0f;
          if (need) m_used = m_new;
          l *= alpha;
          // O_t is stable once P.V of block j-1 has completed
          mbar_wait(&o_full[t], (j - 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < kD / 32; ++c) {
            uint32_t orr[32];
            tmem_ld_x32(tO + c * 32, orr);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * alpha);
            tmem_st_x32(tO + c * 32, orr);
          }
        }
      }
      // (a row of a partially valid tile may be padding with every key masked: m = -inf; subtract 0 instead, P = 0)
      const float neg_ms = (m_used == -INFINITY) ? 0.f : -m_used * sl2;
      const uint64_t sl2_2 = f32x2_pack(sl2, sl2), nm2 = f32x2_pack(neg_ms, neg_ms);
      uint64_t ls[4] = {0ull, 0ull, 0ull, 0ull};   // four packed partial row sums (8 fp32 chains)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint64_t x2 = f32x2_fma(f32x2_pack_bits(sr[c * 32 + 2 * i], sr[c * 32 + 2 * i + 1]), sl2_2, nm2);
          float x0, x1;
          f32x2_unpack(x2, x0, x1);
          float p0, p1;
          if ((FA_FWD_POLY_MASK_D128 >> (i & 7)) & 1) {   // FMA-pipe exp2 for a share of the pairs (fa_fwd_sm100.cuh)
            ex2_poly_x2(x0, x1, p0, p1);
          } else {
            p0 = ex2_approx(x0), p1 = ex2_approx(x1);
          }
          ls[i & 3] = f32x2_add(ls[i & 3], f32x2_pack(p0, p1));
          pk[i] = pack2<kBf16>(p0, p1);
        }
        tmem_st_x16(tS + c * 16, pk);
        if (c == 1) {
          p_arrive(pbar0);
          if (tr) fa_trace2(1 + t, j, 2);
        }
      }
      float la, lb, lc, ld;
      f32x2_unpack(f32x2_add(ls[0], ls[1]), la, lb);
      f32x2_unpack(f32x2_add(ls[2], ls[3]), lc, ld);
      l += (la + lb) + (lc + ld);
      p_arrive(pbar1);
      if (tr) fa_trace2(1 + t, j, 3);
    }

    if (my_slot > 0) {
      // every MMA of this slot has completed (also the other CTA's reads of nothing of ours: A operands are private)
      mbar_wait(&o_full[t], (my_slot - 1) & 1);
      tc_fence_after();
    }
    if (my_own > 0) {
      const float inv_l = 1.0f / l;
      const bool in_range = q_row < nv;
      // Epilogue as in fa_fwd_sm100.cuh: O_t / l -> output dtype -> this tile's Q staging buffer (dead: the last score
      // MMA of the slot completed before o_full) -> global, 512 contiguous bytes per warp instruction.
      constexpr int kRowBytes = kD * 2, kRowChunks = kRowBytes / 16;
      const uint32_t stage = smem_u32(sQ + t * Cfg::kTileBytes);
#pragma unroll
      for (int c = 0; c < kD / 32; ++c) {
        uint32_t orr[32];
        tmem_ld_x32(tO + c * 32, orr);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t a = pack2<kBf16>(__uint_as_float(orr[8 * i + 0]) * inv_l, __uint_as_float(orr[8 * i + 1]) * inv_l);
          const uint32_t bq = pack2<kBf16>(__uint_as_float(orr[8 * i + 2]) * inv_l, __uint_as_float(orr[8 * i + 3]) * inv_l);
          const uint32_t cq = pack2<kBf16>(__uint_as_float(orr[8 * i + 4]) * inv_l, __uint_as_float(orr[8 * i + 5]) * inv_l);
          const uint32_t dq = pack2<kBf16>(__uint_as_float(orr[8 * i + 6]) * inv_l, __uint_as_float(orr[8 * i + 7]) * inv_l);
          const uint32_t ch = c * 4 + i;   // 16-byte chunk of the row
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(stage + row * kRowBytes + ((ch ^ (row & 7)) << 4)),
                       "r"(a), "r"(bq), "r"(cq), "r"(dq)
                       : "memory");
        }
      }
      named_bar_sync(2 + t, 128);   // the 128 threads of this tile
      {
        const int tid = threadIdx.x & 127;
        const int64_t tile_off = ((int64_t)b * p.o_sB + (int64_t)h * p.o_sH) * 2;
        const int64_t row_pitch = p.o_sN * 2;
#pragma unroll 4
        for (int it = 0; it < kRowChunks; ++it) {
          const int idx = it * 128 + tid;
          const int r = idx / kRowChunks, ch = idx - r * kRowChunks;
          if (q0t + r < nv) {
            uint4 v;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "r"(stage + r * kRowBytes + ((ch ^ (r & 7)) << 4)));
            const int64_t off = tile_off + (int64_t)(q0t + r) * row_pitch + ch * 16;
            *reinterpret_cast<uint4*>(static_cast<uint8_t*>(p.o) + off) = v;
            for (int gp = 0; gp < p.n_peer; ++gp) *reinterpret_cast<uint4*>(static_cast<uint8_t*>(p.o_peer[gp]) + off) = v;
          }
        }
      }
      if (in_range) p.lse[((int64_t)b * p.H + h) * p.N + q_row] = m_used * sl2 + log2f(l);
    }
  }

  tc_fence_before();
  cluster_sync_all();   // nobody leaves while the pair's MMAs, multicast commits or remote arrivals can still touch it
  if (warp == 9) tmem_dealloc2<512>(tmem);
}

}  // namespace fa
