// fa_fwd_duo_sm100.cuh — FlashAttention-2 forward for sm_100a in which the two softmax warps of an SM sub-partition work
// on the SAME query tile at the same time (16-bit inputs, no dropout / attention mask: the hot path).
//
// Same math, same TMA ring, same MMA issue order and the same TMEM layout as fa_fwd_sm100.cuh (reference semantics:
// flash_attention_kernels.py:88-108, scale / causal mask of flash_attention_openai_tutorial.py:50,160-161).  What
// changes is who computes the softmax of a tile.  In fa_fwd_sm100.cuh warps 0-3 own tile 0 and warps 4-7 own tile 1;
// the timeline (profiles/r02_trace_fwd_single_cta.txt) shows each tile's chain S_t(j) ready -> softmax (2100 clk, one
// warp per sub-partition walking 128 columns per row) -> P.V_t(j), S_t(j+1) (1024 tensor clocks + hand-overs) to be
// serial, so a key block costs ~3270 clk while the tensor core has 2 x 1024 to do and the sub-partition ~1700 of issue.
// Here the eight warps serve BOTH tiles alternately: warp w and warp w + 4 (same sub-partition, same TMEM lanes) split
// the 128 columns of a row between them in 32-column chunks (w: chunks 0 and 2, w + 4: chunks 1 and 3), so a tile's
// softmax takes about half as long and the other tile's scores are produced meanwhile.  The two exchange their block
// maxima through shared memory (a 64-thread named barrier per sub-partition and block).  The first half of P (keys
// 0-63 = chunks 0 and 1) is complete when both have finished their first chunk, so P.V still starts mid-softmax.
// Unlike the sixteen-warp experiment (fa_fwd_w16_sm100.cuh) this keeps 10 warps and all the registers.
// Results equal fa_fwd_sm100.cuh except for the order in which a row's l is summed (two partial sums of 64).
//
//   warps 0-7    softmax of tiles 0 and 1: warp w -> TMEM lanes 32 (w % 4) .., column chunks (w / 4), (w / 4) + 2
//   warp 8       TMA producer        warp 9  MMA issuer (one elected thread) + TMEM owner
#pragma once

#include "fa_fwd_sm100.cuh"

namespace fa {

template <int kD>
struct FwdDuoCfg {
  static constexpr int kRowBytes = kD * 2;
  static constexpr int kStages = (kD == 128) ? 2 : 4;
  static constexpr int kTileBytes = 128 * kRowBytes;
  static constexpr int kBoxBytes = 128 * 128;
  static constexpr int kBoxes = kRowBytes / 128;
  static constexpr int kSteps = kD / 16;
  static constexpr int kSmemQ = 2 * kTileBytes;
  static constexpr int kSmemKV = kStages * 2 * kTileBytes;
  static constexpr int kSmemBytes = kSmemQ + kSmemKV + 1024 /*alignment slack*/;
  static constexpr int kThreads = 320;   // 8 softmax warps + producer + MMA issuer
  static constexpr uint32_t kTmemS0 = 0, kTmemS1 = 128, kTmemO0 = 256, kTmemO1 = 256 + kD;
};

template <bool kBf16, int kD, bool kCausal>
__global__ void __launch_bounds__(320, 1)
fa_fwd_duo_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const FwdParams p) {
  using Cfg = FwdDuoCfg<kD>;
  constexpr int NS = Cfg::kStages;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                  // [2][tile]
  uint8_t* sK = smem + Cfg::kSmemQ;                    // [NS][tile]
  uint8_t* sV = sK + NS * Cfg::kTileBytes;             // [NS][tile]

  __shared__ uint64_t q_full[2], s_full[2], p_full[2][2], o_full[2];
  __shared__ uint64_t k_full[NS], k_empty[NS], v_full[NS], v_empty[NS];
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_xch[2][2][2][128];   // [block parity][tile][warp of the pair][row]: block maxima, finally the row sums

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // heaviest (largest q index) blocks first so the causal triangle load-balances
  const int qb = p.q_blocks - 1 - (int)blockIdx.x;
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = qb * 256;
  const int nv = p.seqlens ? min(max(p.seqlens[b], 0), p.N) : p.N;   // valid length of this batch element
  if (q0 >= nv) return;                                              // whole CTA is padding (uniform, before any set-up)
  const int n_kv_total = (nv + 127) >> 7;
  const int ntiles = (nv - q0 > 128) ? 2 : 1;
  int nkv[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int n = kCausal ? min(n_kv_total, ((q0 + 128 * t) >> 7) + 1) : n_kv_total;
    nkv[t] = (t < ntiles) ? n : 0;
  }
  const int nkv_max = max(nkv[0], nkv[1]);

  if (threadIdx.x == 0) {
    for (int t = 0; t < 2; ++t) {
      mbar_init(&q_full[t], 1);
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t][0], 256);   // keys 0-63: every softmax warp has stored its first chunk
      mbar_init(&p_full[t][1], 256);   // keys 64-127: ... its second chunk
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
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 9) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      for (int t = 0; t < ntiles; ++t) {
        mbar_arrive_expect_tx(&q_full[t], Cfg::kTileBytes);
        for (int bx = 0; bx < Cfg::kBoxes; ++bx)
          tma_load_4d(sQ + t * Cfg::kTileBytes + bx * Cfg::kBoxBytes, &tmQ, &q_full[t], bx * 64, q0 + 128 * t, h, b);
      }
      for (int j = 0; j < nkv_max; ++j) {
        const int s = j % NS;
        const uint32_t ph = (j / NS) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[s], Cfg::kTileBytes);
        for (int bx = 0; bx < Cfg::kBoxes; ++bx)
          tma_load_4d(sK + s * Cfg::kTileBytes + bx * Cfg::kBoxBytes, &tmK, &k_full[s], bx * 64, j * 128, h, b);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[s], Cfg::kTileBytes);
        for (int bx = 0; bx < Cfg::kBoxes; ++bx)
          tma_load_4d(sV + s * Cfg::kTileBytes + bx * Cfg::kBoxBytes, &tmV, &v_full[s], bx * 64, j * 128, h, b);
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer (as in fa_fwd_sm100.cuh)
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_f16(kBf16, 128, 128, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_f16(kBf16, 128, kD, 0, 1);
      const uint32_t qlo = umma_lo_kmajor(smem_u32(sQ)), klo = umma_lo_kmajor(smem_u32(sK));
      const uint32_t vlo = umma_lo_mnmajor(smem_u32(sV), Cfg::kBoxBytes);
      constexpr uint32_t kTileLo = Cfg::kTileBytes >> 4;
      auto tS = [&](int t) { return tmem + (t ? Cfg::kTmemS1 : Cfg::kTmemS0); };
      auto tO = [&](int t) { return tmem + (t ? Cfg::kTmemO1 : Cfg::kTmemO0); };

      auto issue_s = [&](int t, int j) {   // S of tile t for key block j
        const int s = j % NS;
        const uint32_t a0 = qlo + t * kTileLo, b0 = klo + s * kTileLo, d0 = tS(t);
        mbar_wait(&k_full[s], (j / NS) & 1);
        tc_fence_after();
        static_for<0, Cfg::kSteps>([&](auto kc) {
          constexpr int k = decltype(kc)::value;
          constexpr uint32_t off = umma_koff_kmajor(k, Cfg::kBoxBytes);
          umma_ss_off<off, off>(d0, a0, b0, idesc_s, k > 0);
        });
        tc_commit(&s_full[t]);
        const bool last_user = (t == 1) || (nkv[1] <= j);   // last tile that reads K block j releases the stage
        if (last_user) tc_commit(&k_empty[s]);
      };

      for (int t = 0; t < ntiles; ++t) {
        mbar_wait(&q_full[t], 0);
        issue_s(t, 0);
      }
      for (int j = 0; j < nkv_max; ++j) {
        const int s = j % NS;
        for (int t = 0; t < ntiles; ++t) {
          if (j >= nkv[t]) continue;
          mbar_wait(&v_full[s], (j / NS) & 1);
          fa_trace(0, j, 4 * t);
          const uint32_t dO_t = tO(t), aP = tS(t), bV = vlo + s * kTileLo;
          const bool acc0 = j > 0;
          mbar_wait(&p_full[t][0], j & 1);   // keys 0-63 of the block
          tc_fence_after();
          fa_trace(0, j, 4 * t + 1);
          static_for<0, 4>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            umma_ts_off<k * 8, umma_koff_mnmajor(k)>(dO_t, aP, bV, idesc_o, acc0 || (k > 0));
          });
          mbar_wait(&p_full[t][1], j & 1);   // keys 64-127
          tc_fence_after();
          fa_trace(0, j, 4 * t + 2);
          static_for<4, 8>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            umma_ts_off<k * 8, umma_koff_mnmajor(k)>(dO_t, aP, bV, idesc_o, 1u);
          });
          tc_commit(&o_full[t]);
          const bool last_user = (t == 1) || (nkv[1] <= j);
          if (last_user) tc_commit(&v_empty[s]);
          if (j + 1 < nkv[t]) issue_s(t, j + 1);
          fa_trace(0, j, 4 * t + 3);
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ softmax + epilogue (warps 0-7, both tiles)
    const int hf = warp >> 2;                         // my column chunks of every block: hf and hf + 2
    const int row = (warp & 3) * 32 + lane;           // row inside either tile == TMEM lane
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const float sl2 = p.scale_log2;
    const uint32_t pair_bar = 1 + (warp & 3);         // named barrier of the two warps of this sub-partition
    constexpr int kPolyMask = (kD == 64) ? FA_FWD_POLY_MASK_D64 : FA_FWD_POLY_MASK_D128;

    float m_used[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};   // l: the row sum over MY columns only
    for (int j = 0; j < nkv_max; ++j) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (j >= nkv[t]) continue;
        const uint32_t tS = tmem + (t ? Cfg::kTmemS1 : Cfg::kTmemS0) + lane_base;
        const int q_row = q0 + 128 * t + row;         // global query index
        if ((threadIdx.x & 127) == 0) fa_trace(1 + 2 * hf + t, j, 0);
        mbar_wait(&s_full[t], j & 1);
        tc_fence_after();
        if ((threadIdx.x & 127) == 0) fa_trace(1 + 2 * hf + t, j, 1);
        uint32_t sr[64];                              // chunk hf in sr[0..32), chunk hf + 2 in sr[32..64)
        tmem_ld_x32(tS + hf * 32, *reinterpret_cast<uint32_t(*)[32]>(&sr[0]));
        tmem_ld_x32(tS + hf * 32 + 64, *reinterpret_cast<uint32_t(*)[32]>(&sr[32]));
        tc_wait_ld();

        const int kv0 = j * 128 + hf * 32;            // first key of my first chunk (the second starts 64 keys later)
        const bool diag = kCausal && (kv0 + 95 > q0 + 128 * t);   // my columns touch the diagonal
        const bool ragged = (kv0 + 96 > nv);
        if (diag || ragged) {
          int limit = nv - kv0;                        // first invalid column (ragged / padded keys)
          if (kCausal) limit = min(limit, q_row - kv0 + 1);
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            if (c >= limit) sr[c] = 0xff800000u;   // -inf
            if (c + 64 >= limit) sr[32 + c] = 0xff800000u;
          }
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int c = 0; c < 64; c += 4) {
          mx0 = fmaxf(mx0, __uint_as_float(sr[c]));
          mx1 = fmaxf(mx1, __uint_as_float(sr[c + 1]));
          mx2 = fmaxf(mx2, __uint_as_float(sr[c + 2]));
          mx3 = fmaxf(mx3, __uint_as_float(sr[c + 3]));
        }
        const float m_loc = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        // both warps of a row must scale by the same maximum: exchange the block maxima (double-buffered by block parity,
        // so the partner's next write cannot overtake this read).  The barrier also orders the partner's TMEM loads of its
        // score columns before my P stores, which land in columns it reads (P of chunk c goes to columns 16 c .. 16 c + 15).
        s_xch[j & 1][t][hf][row] = m_loc;
        if ((threadIdx.x & 127) == 0) fa_trace(1 + 2 * hf + t, j, 2);
        named_bar_sync(pair_bar, 64);
        if ((threadIdx.x & 127) == 0) fa_trace(1 + 2 * hf + t, j, 3);
        const float m_new = fmaxf(fmaxf(m_loc, s_xch[j & 1][t][hf ^ 1][row]), m_used[t]);
        if (j == 0) {
          m_used[t] = m_new;
        } else {
          const bool need = (m_new - m_used[t]) * sl2 > 8.0f;
          if (__any_sync(0xffffffffu, need)) {   // (the partner warp holds the same rows and takes the same decision)
            const float alpha = need ? ex2_approx((m_used[t] - m_new) * sl2) : 1.0f;
            if (need) m_used[t] = m_new;
            l[t] *= alpha;
            // O_t is stable once P.V of block j-1 has completed; each warp of the pair rescales its own D/2 columns
            const uint32_t tO = tmem + (t ? Cfg::kTmemO1 : Cfg::kTmemO0) + lane_base + hf * (kD / 2);
            mbar_wait(&o_full[t], (j - 1) & 1);
            tc_fence_after();
            // (16 columns at a time: the 64 score registers stay live across this rare path)
#pragma unroll 1
            for (int c = 0; c < kD / 32; ++c) {
              uint32_t orr[16];
              tmem_ld_x16(tO + c * 16, orr);
              tc_wait_ld();
#pragma unroll
              for (int i = 0; i < 16; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * alpha);
              tmem_st_x16(tO + c * 16, orr);
            }
          }
        }
        const float neg_ms = -m_used[t] * sl2;
        const uint64_t sl2_2 = f32x2_pack(sl2, sl2), nm2 = f32x2_pack(neg_ms, neg_ms);
        uint64_t ls[4] = {0ull, 0ull, 0ull, 0ull};   // four packed partial row sums (8 fp32 chains)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint64_t x2 = f32x2_fma(f32x2_pack_bits(sr[c * 32 + 2 * i], sr[c * 32 + 2 * i + 1]), sl2_2, nm2);
            float x0, x1;
            f32x2_unpack(x2, x0, x1);
            float p0, p1;
            if ((kPolyMask >> (i & 7)) & 1) {   // FMA-pipe exp2
              ex2_poly_x2(x0, x1, p0, p1);
            } else {
              p0 = ex2_approx(x0), p1 = ex2_approx(x1);
            }
            ls[i & 3] = f32x2_add(ls[i & 3], f32x2_pack(p0, p1));
            pk[i] = pack2<kBf16>(p0, p1);
          }
          tmem_st_x16(tS + (hf + 2 * c) * 16, pk);   // packed P of chunk hf + 2 c
          tc_wait_st();
          tc_fence_before();
          mbar_arrive(&p_full[t][c]);
          if ((threadIdx.x & 127) == 0) fa_trace(1 + 2 * hf + t, j, 4 + c);
        }
        float la, lb, lc, ld;
        f32x2_unpack(f32x2_add(ls[0], ls[1]), la, lb);
        f32x2_unpack(f32x2_add(ls[2], ls[3]), lc, ld);
        l[t] += (la + lb) + (lc + ld);
      }
    }

#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int my_nkv = nkv[t];
      if (my_nkv == 0) continue;
      const uint32_t tO = tmem + (t ? Cfg::kTmemO1 : Cfg::kTmemO0) + lane_base + hf * (kD / 2);   // my half of O
      const int q_row = q0 + 128 * t + row;
      mbar_wait(&o_full[t], (my_nkv - 1) & 1);
      tc_fence_after();
      // the row sum is the sum of the two warps' partial sums
      s_xch[my_nkv & 1][t][hf][row] = l[t];
      named_bar_sync(pair_bar, 64);
      const float l_row = s_xch[my_nkv & 1][t][0][row] + s_xch[my_nkv & 1][t][1][row];
      const float inv_l = 1.0f / l_row;
      const bool in_range = q_row < nv;
      // Epilogue: my D/2 columns of O_t / l -> output dtype -> this tile's Q staging buffer (dead since its last S MMA)
      // -> global with 512 contiguous bytes per warp instruction (also to the peer windows of the fused all-gather).
      constexpr int kRowChunks = Cfg::kRowBytes / 16;
      const uint32_t stage = smem_u32(sQ + t * Cfg::kTileBytes);
#pragma unroll
      for (int c = 0; c < kD / 64; ++c) {
        uint32_t orr[32];
        tmem_ld_x32(tO + c * 32, orr);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t a = pack2<kBf16>(__uint_as_float(orr[8 * i + 0]) * inv_l, __uint_as_float(orr[8 * i + 1]) * inv_l);
          const uint32_t bq = pack2<kBf16>(__uint_as_float(orr[8 * i + 2]) * inv_l, __uint_as_float(orr[8 * i + 3]) * inv_l);
          const uint32_t cq = pack2<kBf16>(__uint_as_float(orr[8 * i + 4]) * inv_l, __uint_as_float(orr[8 * i + 5]) * inv_l);
          const uint32_t dq = pack2<kBf16>(__uint_as_float(orr[8 * i + 6]) * inv_l, __uint_as_float(orr[8 * i + 7]) * inv_l);
          const uint32_t ch = hf * (kRowChunks / 2) + c * 4 + i;   // 16-byte chunk of the row
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(stage + row * Cfg::kRowBytes + ((ch ^ (row & 7)) << 4)),
                       "r"(a), "r"(bq), "r"(cq), "r"(dq)
                       : "memory");
        }
      }
      named_bar_sync(5, 256);   // all eight softmax warps: the tile is staged
      {
        const int tid = threadIdx.x & 255;
        const int64_t tile_off = ((int64_t)b * p.o_sB + (int64_t)h * p.o_sH) * 2;
        const int64_t row_pitch = p.o_sN * 2;
        const int row0 = q0 + 128 * t;
#pragma unroll 4
        for (int it = 0; it < kRowChunks / 2; ++it) {
          const int idx = it * 256 + tid;
          const int r = idx / kRowChunks, ch = idx - r * kRowChunks;
          if (row0 + r < nv) {
            uint4 v;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "r"(stage + r * Cfg::kRowBytes + ((ch ^ (r & 7)) << 4)));
            const int64_t off = tile_off + (int64_t)(row0 + r) * row_pitch + ch * 16;
            *reinterpret_cast<uint4*>(static_cast<uint8_t*>(p.o) + off) = v;
            for (int g = 0; g < p.n_peer; ++g) *reinterpret_cast<uint4*>(static_cast<uint8_t*>(p.o_peer[g]) + off) = v;
          }
        }
      }
      if (hf == 0 && in_range) p.lse[((int64_t)b * p.H + h) * p.N + q_row] = m_used[t] * sl2 + log2f(l_row);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem);
}

}  // namespace fa
