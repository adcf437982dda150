// fa_fwd_w16_sm100.cuh — FlashAttention-2 forward for sm_100a with SIXTEEN softmax warps (16-bit inputs, no dropout /
// attention mask: the hot path; the feature variants stay on fa_fwd_sm100.cuh).
//
// Same math, same TMA ring, same MMA issue order and the same TMEM layout as fa_fwd_sm100.cuh (reference semantics:
// flash_attention_kernels.py:88-108, scale / causal mask of flash_attention_openai_tutorial.py:50,160-161).  What
// changes is the softmax stage.  The timeline of the 8-warp kernel (profiles/r02_trace_fwd_single_cta.txt) shows why it
// stops at 62 % tensor activity: the two query tiles of a CTA ping-pong perfectly, but each tile's own chain is serial —
// S_t(j) ready -> softmax_t (2100 clk: one warp per SM sub-partition walks 128 columns per row, in order) -> P.V_t(j)
// and S_t(j+1) (1024 tensor clocks + hand-overs) -> S_t(j+1) ready — so a block costs 2100 + 1024 + ~150 clk per tile
// while the tensor core only has 2 x 1024 to do.  Shared memory is not the limit (tools/umma2_probe.cu: every operand
// form runs at the tensor peak) and neither is MUFU (47 % busy).  Here every tile gets TWO warps per sub-partition, each
// owning 64 of the 128 key columns of its rows: half the serial work per thread, twice the warps to hide TMEM / MUFU
// latency.  The two halves of a row exchange their block maxima through shared memory (one 256-thread named barrier
// per block) so that both use the same running maximum — results are bit-identical to the 8-warp kernel except for the
// order in which a row's l is summed (two partial sums of 64 instead of one of 128).
//
//   warps 0-7    softmax of tile 0: warp w -> TMEM lanes 32 (w % 4) .., key columns 64 (w / 4) .. of every block
//   warps 8-15   softmax of tile 1
//   warp 16      TMA producer        warp 17  MMA issuer (one elected thread) + TMEM owner
#pragma once

#include "fa_fwd_sm100.cuh"

namespace fa {

template <int kD>
struct FwdW16Cfg {
  static constexpr int kRowBytes = kD * 2;
  static constexpr int kStages = (kD == 128) ? 2 : 4;
  static constexpr int kTileBytes = 128 * kRowBytes;
  static constexpr int kBoxBytes = 128 * 128;
  static constexpr int kBoxes = kRowBytes / 128;
  static constexpr int kSteps = kD / 16;
  static constexpr int kSmemQ = 2 * kTileBytes;
  static constexpr int kSmemKV = kStages * 2 * kTileBytes;
  static constexpr int kSmemBytes = kSmemQ + kSmemKV + 1024 /*alignment slack*/;
  static constexpr int kThreads = 576;   // 16 softmax warps + producer + MMA issuer
  static constexpr uint32_t kTmemS0 = 0, kTmemS1 = 128, kTmemO0 = 256, kTmemO1 = 256 + kD;
};

template <bool kBf16, int kD, bool kCausal>
__global__ void __launch_bounds__(576, 1)
fa_fwd_w16_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const FwdParams p) {
  using Cfg = FwdW16Cfg<kD>;
  constexpr int NS = Cfg::kStages;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                  // [2][tile]
  uint8_t* sK = smem + Cfg::kSmemQ;                    // [NS][tile]
  uint8_t* sV = sK + NS * Cfg::kTileBytes;             // [NS][tile]

  __shared__ uint64_t q_full[2], s_full[2], p_full[2][2], o_full[2];
  __shared__ uint64_t k_full[NS], k_empty[NS], v_full[NS], v_empty[NS];
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_xch[2][2][2][128];   // [block parity][tile][column half][row]: block maxima, finally the row sums

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
      mbar_init(&p_full[t][0], 128);   // the four warps that own key columns 0-63 of the tile
      mbar_init(&p_full[t][1], 128);   // ... 64-127
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
  if (warp == 16 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 17) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 16) {
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
  } else if (warp == 17) {
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
          const uint32_t dO_t = tO(t), aP = tS(t), bV = vlo + s * kTileLo;
          const bool acc0 = j > 0;
          mbar_wait(&p_full[t][0], j & 1);   // keys 0-63 of the block
          tc_fence_after();
          static_for<0, 4>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            umma_ts_off<k * 8, umma_koff_mnmajor(k)>(dO_t, aP, bV, idesc_o, acc0 || (k > 0));
          });
          mbar_wait(&p_full[t][1], j & 1);   // keys 64-127
          tc_fence_after();
          static_for<4, 8>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            umma_ts_off<k * 8, umma_koff_mnmajor(k)>(dO_t, aP, bV, idesc_o, 1u);
          });
          tc_commit(&o_full[t]);
          const bool last_user = (t == 1) || (nkv[1] <= j);
          if (last_user) tc_commit(&v_empty[s]);
          if (j + 1 < nkv[t]) issue_s(t, j + 1);
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ softmax + epilogue (warps 0-15)
    const int t = warp >> 3;                          // query tile
    const int hf = (warp >> 2) & 1;                   // which 64 key columns of every block
    const int row = (warp & 3) * 32 + lane;           // row inside the tile == TMEM lane
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + (t ? Cfg::kTmemS1 : Cfg::kTmemS0) + lane_base + hf * 64;   // my 64 score columns
    const uint32_t tP = tmem + (t ? Cfg::kTmemS1 : Cfg::kTmemS0) + lane_base + hf * 32;   // my 32 packed P columns
    const uint32_t tO = tmem + (t ? Cfg::kTmemO1 : Cfg::kTmemO0) + lane_base + hf * (kD / 2);   // my half of O
    const int my_nkv = nkv[t];
    const int q_row = q0 + 128 * t + row;             // global query index
    const float sl2 = p.scale_log2;
    const uint32_t bar_id = 1 + t;                    // named barrier of this tile's 256 threads

    float m_used = -INFINITY, l = 0.f;                // l: the row sum over MY columns only
    for (int j = 0; j < my_nkv; ++j) {
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      uint32_t sr[64];
      tmem_ld_x32(tS, *reinterpret_cast<uint32_t(*)[32]>(&sr[0]));
      tmem_ld_x32(tS + 32, *reinterpret_cast<uint32_t(*)[32]>(&sr[32]));
      tc_wait_ld();

      const int kv0 = j * 128 + hf * 64;              // first key of my columns
      const bool diag = kCausal && (kv0 + 63 > q0 + 128 * t);   // my columns touch the diagonal
      const bool ragged = (kv0 + 64 > nv);
      if (diag || ragged) {
        int limit = nv - kv0;                          // first invalid column (ragged / padded keys)
        if (kCausal) limit = min(limit, q_row - kv0 + 1);
#pragma unroll
        for (int c = 0; c < 64; ++c)
          if (c >= limit) sr[c] = 0xff800000u;   // -inf
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
      // both halves of the row must scale by the same maximum: exchange the block maxima (double-buffered by block parity,
      // so the partner's next write cannot overtake this read).  The barrier also orders the partner's TMEM loads of its
      // score columns before my P stores, which land in columns it reads (P of keys 64-127 goes to columns 32-63).
      s_xch[j & 1][t][hf][row] = m_loc;
      named_bar_sync(bar_id, 256);
      const float m_new = fmaxf(fmaxf(m_loc, s_xch[j & 1][t][hf ^ 1][row]), m_used);
      if (j == 0) {
        m_used = m_new;
      } else {
        const bool need = (m_new - m_used) * sl2 > 8.0f;
        if (__any_sync(0xffffffffu, need)) {   // (the partner warp holds the same rows and takes the same decision)
          const float alpha = need ? ex2_approx((m_used - m_new) * sl2) : 1.0f;
          if (need) m_used = m_new;
          l *= alpha;
          // O_t is stable once P.V of block j-1 has completed; each half rescales its own D/2 columns
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
      const float neg_ms = -m_used * sl2;
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
          if (((kD == 64 ? FA_FWD_POLY_MASK_D64 : FA_FWD_POLY_MASK_D128) >> (i & 7)) & 1) {   // FMA-pipe exp2
            ex2_poly_x2(x0, x1, p0, p1);
          } else {
            p0 = ex2_approx(x0), p1 = ex2_approx(x1);
          }
          ls[i & 3] = f32x2_add(ls[i & 3], f32x2_pack(p0, p1));
          pk[i] = pack2<kBf16>(p0, p1);
        }
        tmem_st_x16(tP + c * 16, pk);
      }
      float la, lb, lc, ld;
      f32x2_unpack(f32x2_add(ls[0], ls[1]), la, lb);
      f32x2_unpack(f32x2_add(ls[2], ls[3]), lc, ld);
      l += (la + lb) + (lc + ld);
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&p_full[t][hf]);
    }

    if (my_nkv > 0) {
      mbar_wait(&o_full[t], (my_nkv - 1) & 1);
      tc_fence_after();
      // the row sum is the sum of the two halves' partial sums
      s_xch[my_nkv & 1][t][hf][row] = l;
      named_bar_sync(bar_id, 256);
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
      named_bar_sync(bar_id, 256);   // the 256 threads of this tile
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
      if (hf == 0 && in_range) p.lse[((int64_t)b * p.H + h) * p.N + q_row] = m_used * sl2 + log2f(l_row);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 17) tmem_dealloc<512>(tmem);
}

}  // namespace fa
