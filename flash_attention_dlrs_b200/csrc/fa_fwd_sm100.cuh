// fa_fwd_sm100.cuh — FlashAttention-2 forward for sm_100a: TMA-fed K/V ring, tcgen05 MMAs with TMEM
// accumulators, warp-specialised online softmax.
//
// Semantics follow the reference forward kernel (flash_attention_kernels.py:88-108, generalised with the
// softmax scale and causal mask of flash_attention_openai_tutorial.py:50,160-161):
//   S = Q K^T ; S2 = S * (scale*log2 e) ; mask j>i ; online max / exp2 / row-sum in fp32 ;
//   P cast to the input dtype (RTNE) before P·V ; O = acc / l ; L = m + log2(l)  (log2 units, fp32).
// (An fp16 P against bf16 V would cut the P rounding error 8x, but tcgen05.mma kind::f16 with A = f16, B = bf16
// raises an illegal-instruction fault on B200 — measured — so P follows the input dtype like the reference.)
//
// One CTA owns a 256-row query block as two 128-row tiles that ping-pong on the tensor core:
//   warps 0-3  softmax for tile 0 (thread = one query row, `tcgen05.ld 32x32b`)
//   warps 4-7  softmax for tile 1
//   warp  8    TMA producer (Q once; K and V through an mbarrier ring)
//   warp  9    MMA issuer (one elected thread) + TMEM owner
//   warps 10-11 idle: they complete the third warpgroup so that `setmaxnreg` can move its registers to the
//              softmax warpgroups (224 registers per softmax thread, no spills; 48 for the third warpgroup)
// TMEM columns: S0 [0,128) S1 [128,256) O0 [256,256+D) O1 [256+D,256+2D).  P (16-bit) overwrites the first 64
// columns of its S tile and is the TMEM A-operand of the P·V MMA; V is the MN-major shared-memory B operand.
// The O rescale is lazy: a warp rewrites its O rows only when a row max grew by more than 2^8.
#pragma once

#include "fa_dropout.cuh"
#include "sm100_ptx.cuh"

// Which of every 8 consecutive column pairs compute exp2 with the FMA-pipe polynomial instead of MUFU (bit i = pair i).
// Measured on B200 (tools/fwd_probe.py): D = 128 is bound by the S -> softmax -> P.V dependency chain, a quarter of the
// exponentials off MUFU gives +2 %; D = 64 (half the tensor work per exponential) gains 10 % with three in eight;
// one in two is slower everywhere (issue-slot bound).
#ifndef FA_FWD_POLY_MASK_D128
#define FA_FWD_POLY_MASK_D128 0x92   // round 2 (watchdog-free build): 0.869 ms on config 3 against 0.883 (0x88), 0.909 (none), 0.966 (0xAA)
#endif
#ifndef FA_FWD_POLY_MASK_D64
#define FA_FWD_POLY_MASK_D64 0x92
#endif
#ifndef FA_FWD_POLY_MASK_F8
#define FA_FWD_POLY_MASK_F8 0x92
#endif

namespace fa {

struct FwdParams {
  void* o;       // (B,H,N,D) 16-bit
  float* lse;    // (B,H,N) fp32, log2 units
  int B, H, N;
  int Nk;   // key / value rows when they differ from the N query rows (rectangular attention: non-causal, no seqlens / masks); 0 = N
  int64_t o_sB, o_sH, o_sN;  // element strides of O (last dim contiguous)
  float scale_log2;          // softmax_scale * log2(e)
  int q_blocks;              // ceil(N / 256)
  // Fused all-gather epilogue (head-sharded multi-GPU): every O row is also stored, with the same strides, into the
  // peer-mapped windows of the other GPUs' gathered output (NVLink P2P stores issued by the epilogue warps, so the
  // transfer overlaps the remaining tiles).  With NVLS the caller passes a multicast address as `o` instead and the
  // switch replicates each store; n_peer is then 0.
  void* o_peer[7];
  int n_peer;
  // Key-padding mask ("masking" on the reference's roadmap, README.md:35-37): seqlens[b] = number of valid tokens of
  // batch element b (nullptr = all N).  Keys >= seqlens[b] are masked out; query rows >= seqlens[b] are not computed
  // (their O / L are left untouched: the caller zero-fills).
  const int* seqlens;
  // Dropout of the attention probabilities (kDrop instantiations only; 16-bit dtypes): see fa_dropout.cuh.
  DropParams drop;
  // Arbitrary attention mask (kAmask instantiations only; 16-bit dtypes): one BIT per (query, key), 1 = attend (key j of
  // a row is bit j & 7 of byte j >> 3), combined (AND) with the causal flag and seqlens.  Row pitch am_sN is a multiple
  // of 16 bytes >= N rounded up to 128, / 8 (a key block is one 16-byte load per row: four registers, held while the
  // thread waits for S); am_sB / am_sH may be 0 (broadcast).  A row with no visible key gives O = 0, L = -inf.
  const uint8_t* amask;
  int64_t am_sB, am_sH, am_sN;
  // Optional block summary of the mask: ablock[.., i, j] != 0 iff some (query, key) of the 128 x 128 block (i, j) is
  // visible; 2 = every entry of the block is visible (its mask bytes are then neither loaded nor applied).  Blocks flagged 0 are skipped: every role of the CTA walks the list of key blocks that at least one of its
  // two query tiles needs (K / V ring stages and phases count list steps), and a tile touches its S / P / O barriers only
  // for its own visible blocks (phases count those) — a skipped block costs nothing.  nullptr, or more than 512 key
  // blocks: no skipping.
  const uint8_t* ablock;
  int64_t ab_sB, ab_sH, ab_sI;
  // Band mask (kAmask instantiations with amask == nullptr): no mask bytes at all, query i sees the keys j with
  // -win_left <= j - i <= win_right (sliding-window / local attention); the block summary is computed by the caller.
  int band, win_left, win_right;
};

// kElt: element type of Q, K, V, P and O — 0 = float16, 1 = bfloat16 (tcgen05 kind::f16), 3 = FP8 E4M3, 4 = FP8 E5M2
// (kind::f8f6f4; forward only — the reference's dtype map lists float8_e5m2, flash_attention_torch.py:15-16, and the
// tutorial's fp8 path casts P to the FP8 type of V before P.V, flash_attention_openai_tutorial.py:66-67).
template <int kD, int kElt = 1>
struct FwdCfg {
  static constexpr bool kF8 = kElt >= 3;
  static constexpr int kEltBytes = kF8 ? 1 : 2;
  static constexpr int kRowBytes = kD * kEltBytes;         // one operand row
  static constexpr int kStages = (kRowBytes == 256) ? 2 : 4;
  static constexpr int kTileBytes = 128 * kRowBytes;       // one 128-row operand tile
  static constexpr int kBoxBytes = 128 * 128;              // one 128-byte-wide box of it
  static constexpr int kBoxes = kRowBytes / 128;
  static constexpr int kBoxElems = 128 / kEltBytes;        // TMA x-coordinate step between boxes
  static constexpr int kSteps = kRowBytes / 32;            // one MMA contracts 32 bytes of K
  static constexpr int kSmemQ = 2 * kTileBytes;
  static constexpr int kSmemKV = kStages * 2 * kTileBytes;
  static constexpr int kSmemBytes = kSmemQ + kSmemKV + 1024 /*alignment slack*/;
  static constexpr int kThreads = 384;
  static constexpr uint32_t kTmemS0 = 0, kTmemS1 = 128, kTmemO0 = 256, kTmemO1 = 256 + kD;
};

template <int kElt, int kD, bool kCausal, bool kDrop = false, bool kAmask = false>
__global__ void __launch_bounds__(384, 1)
fa_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
              const __grid_constant__ CUtensorMap tmV, const FwdParams p) {
  using Cfg = FwdCfg<kD, kElt>;
  constexpr bool kF8 = Cfg::kF8, kBf16 = kElt == 1, kE5M2 = kElt == 4;
  static_assert(!kF8 || kD == 128, "the FP8 forward runs at D = 128 (one 128-byte box per row); pad in the caller");
  static_assert(!(kF8 && (kDrop || kAmask)), "dropout and attention masks are implemented for the 16-bit kernels");
  constexpr int NS = Cfg::kStages;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                  // [2][tile]
  uint8_t* sK = smem + Cfg::kSmemQ;                    // [NS][tile]
  uint8_t* sV = sK + NS * Cfg::kTileBytes;             // [NS][tile]

  __shared__ uint64_t q_full[2], s_full[2], p_full[2][2], o_full[2];
  __shared__ uint64_t k_full[NS], k_empty[NS], v_full[NS], v_empty[NS];
  __shared__ uint32_t tmem_base_s;
  __shared__ uint8_t s_act[2][kAmask ? 512 : 4];   // kAmask: block (tile t, key block j) has a visible entry
  __shared__ uint16_t s_list[kAmask ? 512 : 2];    // kAmask: the key blocks some tile of this CTA needs, in order
  __shared__ int s_nlist;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // heaviest (largest q index) blocks first so the causal triangle load-balances
  const int qb = p.q_blocks - 1 - (int)blockIdx.x;
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = qb * 256;
  const int nv = p.seqlens ? min(max(p.seqlens[b], 0), p.N) : p.N;   // valid length of this batch element
  if (q0 >= nv) return;                                              // whole CTA is padding (uniform, before any set-up)
  const int nk = p.Nk > 0 ? p.Nk : nv;                               // valid key rows
  const int n_kv_total = (nk + 127) >> 7;
  const int ntiles = (nv - q0 > 128) ? 2 : 1;
  int nkv[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    int n = kCausal ? min(n_kv_total, ((q0 + 128 * t) >> 7) + 1) : n_kv_total;
    nkv[t] = (t < ntiles) ? n : 0;
  }
  const int nkv_max = max(nkv[0], nkv[1]);
  const bool use_act = kAmask && p.ablock != nullptr && n_kv_total <= 512;
  if constexpr (kAmask) {
    if (use_act) {
      for (int t = 0; t < 2; ++t) {
        const uint8_t* ab = p.ablock + (int64_t)b * p.ab_sB + (int64_t)h * p.ab_sH + (int64_t)(2 * qb + t) * p.ab_sI;
        for (int j = threadIdx.x; j < nkv[t]; j += blockDim.x) s_act[t][j] = ab[j];
      }
      __syncthreads();   // (use_act is uniform over the CTA)
      if (warp == 0) {   // one warp compacts the union of the two tiles' flags (ballot + prefix count), 32 blocks per step
        int c = 0;
        for (int base = 0; base < nkv_max; base += 32) {
          const int j = base + lane;
          const bool on = (j < nkv[0] && s_act[0][j]) || (j < nkv[1] && s_act[1][j]);
          const uint32_t m = __ballot_sync(0xffffffffu, on);
          if (on) s_list[c + __popc(m & ((1u << lane) - 1u))] = (uint16_t)j;
          c += __popc(m);
        }
        if (lane == 0) s_nlist = c;
      }
    }
  }
  // (published by the __syncthreads below)
  auto active = [&](int t, int j) -> bool { return !use_act || s_act[t][j] != 0; };

  if (threadIdx.x == 0) {
    for (int t = 0; t < 2; ++t) {
      mbar_init(&q_full[t], 1);
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t][0], 128);
      mbar_init(&p_full[t][1], 128);
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
  // list step u handles key block block_of(u); without a block summary the list is 0, 1, 2, ...
  const int n_list = use_act ? s_nlist : nkv_max;
  auto block_of = [&](int u) -> int { return use_act ? (int)s_list[u] : u; };

  if (warp >= 8) {
  setmaxnreg_dec<72>();   // third warpgroup: producer, MMA issuer, two idle warps
  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      for (int t = 0; t < ntiles; ++t) {
        mbar_arrive_expect_tx(&q_full[t], Cfg::kTileBytes);
        for (int bx = 0; bx < Cfg::kBoxes; ++bx)
          tma_load_4d(sQ + t * Cfg::kTileBytes + bx * Cfg::kBoxBytes, &tmQ, &q_full[t], bx * Cfg::kBoxElems, q0 + 128 * t, h,
                      b);
      }
      for (int u = 0; u < n_list; ++u) {
        const int j = kAmask ? block_of(u) : u;
        const int s = u % NS;
        const uint32_t ph = (u / NS) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[s], Cfg::kTileBytes);
        for (int bx = 0; bx < Cfg::kBoxes; ++bx)
          tma_load_4d(sK + s * Cfg::kTileBytes + bx * Cfg::kBoxBytes, &tmK, &k_full[s], bx * Cfg::kBoxElems, j * 128, h, b);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[s], Cfg::kTileBytes);
        for (int bx = 0; bx < Cfg::kBoxes; ++bx)
          tma_load_4d(sV + s * Cfg::kTileBytes + bx * Cfg::kBoxBytes, &tmV, &v_full[s], bx * Cfg::kBoxElems, j * 128, h, b);
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_s = kF8 ? umma_idesc_f8(kE5M2, 128, 128, 0, 0) : umma_idesc_f16(kBf16, 128, 128, 0, 0);
      constexpr uint32_t idesc_o = kF8 ? umma_idesc_f8(kE5M2, 128, kD, 0, 1) : umma_idesc_f16(kBf16, 128, kD, 0, 1);
      const uint32_t qlo = umma_lo_kmajor(smem_u32(sQ)), klo = umma_lo_kmajor(smem_u32(sK));
      const uint32_t vlo = umma_lo_mnmajor(smem_u32(sV), Cfg::kBoxBytes);
      constexpr uint32_t kTileLo = Cfg::kTileBytes >> 4;
      auto tS = [&](int t) { return tmem + (t ? Cfg::kTmemS1 : Cfg::kTmemS0); };
      auto tO = [&](int t) { return tmem + (t ? Cfg::kTmemO1 : Cfg::kTmemO0); };

      auto issue_s = [&](int t, int u) {   // S of tile t for list step u (key block j)
        const int j = kAmask ? block_of(u) : u;
        const int s = u % NS;
        const uint32_t a0 = qlo + t * kTileLo, b0 = klo + s * kTileLo, d0 = tS(t);
        if (!kAmask || active(t, j)) {
          mbar_wait(&k_full[s], (u / NS) & 1);
          tc_fence_after();
          static_for<0, Cfg::kSteps>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            constexpr uint32_t off = umma_koff_kmajor(k, Cfg::kBoxBytes);
            if constexpr (kF8)
              umma8_ss_off<off, off>(d0, a0, b0, idesc_s, k > 0);
            else
              umma_ss_off<off, off>(d0, a0, b0, idesc_s, k > 0);
          });
          tc_commit(&s_full[t]);
        }
        // last tile that reads K block j releases the stage
        const bool last_user = (t == 1) || (nkv[1] <= j);
        if (last_user) tc_commit(&k_empty[s]);
      };

      for (int t = 0; t < ntiles; ++t) {
        mbar_wait(&q_full[t], 0);
        if (!kAmask || (n_list > 0 && block_of(0) < nkv[t])) issue_s(t, 0);
      }
      int n_pv[2] = {0, 0};   // kAmask: P.V products issued per tile so far (barrier phases; the first one initialises O)
      for (int u = 0; u < n_list; ++u) {
        const int j = kAmask ? block_of(u) : u;
        const int jn = (kAmask && u + 1 < n_list) ? block_of(u + 1) : j + 1;   // key block of the next list step
        const int s = u % NS;
        for (int t = 0; t < ntiles; ++t) {
          if (j >= nkv[t]) continue;
          if (!kAmask || active(t, j)) {
            mbar_wait(&v_full[s], (u / NS) & 1);
            // P arrives in two 64-key halves so the first half of P·V overlaps the second half of the exponentials
            const uint32_t dO_t = tO(t), aP = tS(t), bV = vlo + s * kTileLo;
            // 16-bit P: 16 keys = 8 TMEM columns and 2048 bytes of V per MMA; FP8 P: 32 keys = 8 columns and 4096 bytes
            constexpr int kPSteps = kF8 ? 4 : 8, kVStep = kF8 ? 2 : 1;
            const int c = kAmask ? n_pv[t] : j;   // this tile's count of visible blocks so far
            const bool acc0 = c > 0;
            fa_trace(0, u, 4 * t);
            mbar_wait(&p_full[t][0], c & 1);
            fa_trace(0, u, 4 * t + 1);
            tc_fence_after();
            static_for<0, kPSteps / 2>([&](auto kc) {
              constexpr int k = decltype(kc)::value;
              if constexpr (kF8)
                umma8_ts_off<k * 8, umma_koff_mnmajor(k * kVStep)>(dO_t, aP, bV, idesc_o, acc0 || (k > 0));
              else
                umma_ts_off<k * 8, umma_koff_mnmajor(k * kVStep)>(dO_t, aP, bV, idesc_o, acc0 || (k > 0));
            });
            mbar_wait(&p_full[t][1], c & 1);
            fa_trace(0, u, 4 * t + 2);
            tc_fence_after();
            static_for<kPSteps / 2, kPSteps>([&](auto kc) {
              constexpr int k = decltype(kc)::value;
              if constexpr (kF8)
                umma8_ts_off<k * 8, umma_koff_mnmajor(k * kVStep)>(dO_t, aP, bV, idesc_o, 1u);
              else
                umma_ts_off<k * 8, umma_koff_mnmajor(k * kVStep)>(dO_t, aP, bV, idesc_o, 1u);
            });
            n_pv[t] = c + 1;
            tc_commit(&o_full[t]);
          }
          const bool last_user = (t == 1) || (nkv[1] <= j);
          if (last_user) tc_commit(&v_empty[s]);
          if (kAmask ? (u + 1 < n_list && jn < nkv[t]) : (j + 1 < nkv[t])) issue_s(t, u + 1);
          fa_trace(0, u, 4 * t + 3);
        }
      }
    }
    __syncwarp();
  }
  } else {
    setmaxnreg_inc<216>();
    // ------------------------------------------------------------------ softmax + epilogue (warps 0-7)
    const int t = warp >> 2;
    const int row = (warp & 3) * 32 + lane;           // row inside the tile == TMEM lane
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + (t ? Cfg::kTmemS1 : Cfg::kTmemS0) + lane_base;
    const uint32_t tO = tmem + (t ? Cfg::kTmemO1 : Cfg::kTmemO0) + lane_base;
    const int my_nkv = nkv[t];
    const int q_row = q0 + 128 * t + row;             // global query index
    const float sl2 = p.scale_log2;
    // dropout: this thread walks row q_row of the mask, one hash per pair of keys (fa_dropout.cuh)
    uint32_t drop_row = 0, drop_shift = 0;
    if constexpr (kDrop) {
      drop_row = drop_key(p.drop, b * p.H + h) + drop_word_index(q_row, 0);
      drop_shift = 16u * (q_row & 1);
    }

    const uint8_t* am_row = nullptr;
    if constexpr (kAmask) {
      if (p.amask) am_row = p.amask + (int64_t)b * p.am_sB + (int64_t)h * p.am_sH + (int64_t)min(q_row, p.N - 1) * p.am_sN;
    }

    float m_used = -INFINITY, l = 0.f;
    int n_seen = 0;   // kAmask: key blocks of this tile that were not skipped so far (barrier phases count these)
    const int n_mine = kAmask ? n_list : my_nkv;
    for (int u = 0; u < n_mine; ++u) {
      const int j = kAmask ? block_of(u) : u;
      if (kAmask && j >= my_nkv) break;
      uint4 mk = make_uint4(0u, 0u, 0u, 0u);   // this row's 128 mask bits of key block j, requested before the wait for S
      if constexpr (kAmask) {
        if (!active(t, j)) continue;   // skipped block: nobody touches this tile's barriers for it
      }
      const bool partial = kAmask && !(use_act && s_act[t][j] == 2);   // some entries of the block are masked out
      if constexpr (kAmask) {
        if (partial && am_row) mk = __ldg(reinterpret_cast<const uint4*>(am_row + j * 16));
      }
      const int c = kAmask ? n_seen : j;
      if ((threadIdx.x & 127) == 0) fa_trace(1 + t, u, 0);
      mbar_wait(&s_full[t], c & 1);
      if ((threadIdx.x & 127) == 0) fa_trace(1 + t, u, 1);
      tc_fence_after();
#if FA_ABLATE == 3
      tc_fence_before();
      mbar_arrive(&p_full[t][0]);
      mbar_arrive(&p_full[t][1]);
      l = 1.f;
      m_used = 0.f;
      continue;
#endif
      uint32_t sr[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_x32(tS + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&sr[c * 32]));
      tc_wait_ld();

      const int kv0 = j * 128;
      const bool diag = kCausal && (kv0 + 127 > q0 + 128 * t);   // block touches the diagonal
      const bool ragged = (kv0 + 128 > nk);
      const bool band_cut = kAmask && partial && !am_row;   // band mask (no bytes): the band's edge crosses this block
      if (diag || ragged || band_cut) {
        int limit = nk - kv0;                          // first invalid column (ragged / padded keys)
        if (kCausal) limit = min(limit, q_row - kv0 + 1);
        int lo = 0;                                    // first visible column (band masks only)
        if constexpr (kAmask) {
          if (band_cut) {
            limit = min(limit, q_row + p.win_right - kv0 + 1);
            lo = q_row - p.win_left - kv0;
          }
        }
#pragma unroll
        for (int c = 0; c < 128; ++c)
          if (c >= limit || (kAmask && c < lo)) sr[c] = 0xff800000u;   // -inf
      }
      if (kAmask && partial && am_row) {
#pragma unroll
        for (int c = 0; c < 128; ++c) {
          const uint32_t w = (c >> 5) == 0 ? mk.x : (c >> 5) == 1 ? mk.y : (c >> 5) == 2 ? mk.z : mk.w;
          if (!(w & (1u << (c & 31)))) sr[c] = 0xff800000u;   // -inf
        }
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
      n_seen = c + 1;
      if (c == 0) {
        m_used = m_new;
      } else {
        const bool need = (m_new - m_used) * sl2 > 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          const float alpha = need ? ex2_approx((m_used - m_new) * sl2) : 1.0f;
          if (need) m_used = m_new;
          l *= alpha;
          // O_t is stable once P·V of block j-1 has completed
          mbar_wait(&o_full[t], (c - 1) & 1);
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
      // (with an attention mask a row may have seen no key yet: m = -inf; use 0 so that exp2(-inf - 0) = 0, not NaN)
      const float neg_ms = (kAmask && m_used == -INFINITY) ? 0.f : -m_used * sl2;
      const uint64_t sl2_2 = f32x2_pack(sl2, sl2), nm2 = f32x2_pack(neg_ms, neg_ms);
      uint64_t ls[4] = {0ull, 0ull, 0ull, 0ull};   // four packed partial row sums (8 fp32 chains)
      if constexpr (!kF8) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
#if FA_ABLATE == 2
            pk[i] = sr[c * 32 + 2 * i] ^ sr[c * 32 + 2 * i + 1];
            continue;
#endif
            const uint64_t x2 = f32x2_fma(f32x2_pack_bits(sr[c * 32 + 2 * i], sr[c * 32 + 2 * i + 1]), sl2_2, nm2);
            float x0, x1;
            f32x2_unpack(x2, x0, x1);
            float p0, p1;
            // (the polynomial clamps at 2^-125 instead of 0: masked-out rows must sum to exactly 0, so kAmask keeps MUFU)
            if (!kAmask && (((kD == 64 ? FA_FWD_POLY_MASK_D64 : FA_FWD_POLY_MASK_D128) >> (i & 7)) & 1)) {   // FMA-pipe exp2
              ex2_poly_x2(x0, x1, p0, p1);
            } else {
              p0 = ex2_approx(x0), p1 = ex2_approx(x1);
            }
            ls[i & 3] = f32x2_add(ls[i & 3], f32x2_pack(p0, p1));   // the row sum is that of the undropped P
            if constexpr (kDrop) {
              bool keep0, keep1;
              drop_keep_pair<8>(drop_row + (uint32_t)(kv0 >> 1) + (uint32_t)(c * 16 + i), drop_shift, p.drop.thresh,
                                keep0, keep1);
              p0 = keep0 ? p0 : 0.f;
              p1 = keep1 ? p1 : 0.f;
            }
            pk[i] = pack2<kBf16>(p0, p1);
          }
          tmem_st_x16(tS + c * 16, pk);
          if (c == 1) {
            tc_wait_st();
            tc_fence_before();
            mbar_arrive(&p_full[t][0]);
            if ((threadIdx.x & 127) == 0) fa_trace(1 + t, u, 2);
          }
        }
      } else {
        // FP8 P: four keys per 32-bit TMEM column, 64 keys = 16 columns per store
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int e = c * 64 + 4 * i;
            float x0, x1, x2, x3;
            f32x2_unpack(f32x2_fma(f32x2_pack_bits(sr[e], sr[e + 1]), sl2_2, nm2), x0, x1);
            f32x2_unpack(f32x2_fma(f32x2_pack_bits(sr[e + 2], sr[e + 3]), sl2_2, nm2), x2, x3);
            float p0, p1, p2, p3;
            if ((FA_FWD_POLY_MASK_F8 >> ((2 * i) & 7)) & 1) {
              ex2_poly_x2(x0, x1, p0, p1);
            } else {
              p0 = ex2_approx(x0), p1 = ex2_approx(x1);
            }
            if ((FA_FWD_POLY_MASK_F8 >> ((2 * i + 1) & 7)) & 1) {
              ex2_poly_x2(x2, x3, p2, p3);
            } else {
              p2 = ex2_approx(x2), p3 = ex2_approx(x3);
            }
            ls[i & 3] = f32x2_add(ls[i & 3], f32x2_add(f32x2_pack(p0, p1), f32x2_pack(p2, p3)));
            pk[i] = pack4_f8<kE5M2>(p0, p1, p2, p3);
          }
          tmem_st_x16(tS + c * 16, pk);
          if (c == 0) {
            tc_wait_st();
            tc_fence_before();
            mbar_arrive(&p_full[t][0]);
          }
        }
      }
      float la, lb, lc, ld;
      f32x2_unpack(f32x2_add(ls[0], ls[1]), la, lb);
      f32x2_unpack(f32x2_add(ls[2], ls[3]), lc, ld);
      l += (la + lb) + (lc + ld);
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&p_full[t][1]);
      if ((threadIdx.x & 127) == 0) fa_trace(1 + t, u, 3);
    }

    if (my_nkv > 0) {
      const int n_done = kAmask ? n_seen : my_nkv;
      if (n_done > 0) mbar_wait(&o_full[t], (n_done - 1) & 1);
      // (a tile with every block skipped never waited for anything: its Q load must have landed before the staging
      // buffer below, which is the Q tile, is overwritten)
      if (kAmask && n_done == 0) mbar_wait(&q_full[t], 0);
      tc_fence_after();
      float inv_l = (kDrop ? p.drop.rp : 1.0f) / l;   // kept probabilities are scaled by 1 / (1 - p_drop)
      if (kAmask && !(l > 0.f)) inv_l = 0.f;          // no visible key: O = 0 (and L = -inf below)
      const bool in_range = q_row < nv;
      // Epilogue: O_t / l -> output dtype -> this tile's Q staging buffer (dead since its last S MMA; same size as the O
      // tile) -> global.  Going through shared memory turns "thread = row" into "warp = two full rows": every warp
      // store covers 512 contiguous bytes, which is what makes the peer / multicast copies of the fused all-gather
      // travel as full NVLink packets instead of one packet per 16 bytes.
      constexpr int kRowChunks = Cfg::kRowBytes / 16;
      const uint32_t stage = smem_u32(sQ + t * Cfg::kTileBytes);
#pragma unroll
      for (int c = 0; c < kD / 32; ++c) {
        uint32_t orr[32];
        tmem_ld_x32(tO + c * 32, orr);
        tc_wait_ld();
        if (kAmask && n_seen == 0) {   // every key block of this tile was skipped: O was never written
#pragma unroll
          for (int i = 0; i < 32; ++i) orr[i] = 0u;
        }
        if constexpr (!kF8) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t a = pack2<kBf16>(__uint_as_float(orr[8 * i + 0]) * inv_l, __uint_as_float(orr[8 * i + 1]) * inv_l);
            const uint32_t bq = pack2<kBf16>(__uint_as_float(orr[8 * i + 2]) * inv_l, __uint_as_float(orr[8 * i + 3]) * inv_l);
            const uint32_t cq = pack2<kBf16>(__uint_as_float(orr[8 * i + 4]) * inv_l, __uint_as_float(orr[8 * i + 5]) * inv_l);
            const uint32_t dq = pack2<kBf16>(__uint_as_float(orr[8 * i + 6]) * inv_l, __uint_as_float(orr[8 * i + 7]) * inv_l);
            const uint32_t ch = c * 4 + i;   // 16-byte chunk of the row
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(stage + row * Cfg::kRowBytes +
                                                                            ((ch ^ (row & 7)) << 4)),
                         "r"(a), "r"(bq), "r"(cq), "r"(dq)
                         : "memory");
          }
        } else {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            uint32_t w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int e = 16 * i + 4 * u;
              w[u] = pack4_f8<kE5M2>(__uint_as_float(orr[e]) * inv_l, __uint_as_float(orr[e + 1]) * inv_l,
                                     __uint_as_float(orr[e + 2]) * inv_l, __uint_as_float(orr[e + 3]) * inv_l);
            }
            const uint32_t ch = c * 2 + i;
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(stage + row * Cfg::kRowBytes +
                                                                            ((ch ^ (row & 7)) << 4)),
                         "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                         : "memory");
          }
        }
      }
      named_bar_sync(2 + t, 128);   // the 128 threads of this tile
      {
        const int tid = threadIdx.x & 127;
        const int64_t tile_off = ((int64_t)b * p.o_sB + (int64_t)h * p.o_sH) * Cfg::kEltBytes;
        const int64_t row_pitch = p.o_sN * Cfg::kEltBytes;
        const int row0 = q0 + 128 * t;
#pragma unroll 4
        for (int it = 0; it < kRowChunks; ++it) {
          const int idx = it * 128 + tid;
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
      if (in_range) p.lse[((int64_t)b * p.H + h) * p.N + q_row] = m_used * sl2 + log2f(l);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem);
}

}  // namespace fa
