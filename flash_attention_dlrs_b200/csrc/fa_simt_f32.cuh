// fa_simt_f32.cuh — float32 attention forward / backward on the FP32 pipe (FFMA), sm_100a.
//
// tcgen05 has no fp32-input MMA (kind::tf32 keeps 10 mantissa bits, ~1e-2 absolute error on logits at the
// reference's scale=1), and the reference's own fp32 path is IEEE FFMA (DOT_PRECISION="ieee",
// flash_attention_kernels.py:6), so the fp32 dtype is served by a tiled SIMT kernel.  Same math and the same
// log2-domain conventions as the 16-bit tcgen05 kernels (flash_attention_kernels.py:88-108, 275-329):
//   fwd : S2 = (Q K^T) * scale*log2e ; online max / exp2 / rowsum ; O = acc / l ; L = m + log2 l
//   bwd : P = exp2(S2 - L) ; dV += P^T dO ; dP = dO V^T ; dS = P o (dP - delta) ; dK += s dS^T Q ; dQ += s dS K
// The backward is two kernels (dK/dV per key block, dQ per query block): every output element has exactly one
// owner thread and a fixed summation order, so results are bit-identical run to run.
//
// 64 x 64 score tiles, 256 threads as a 16 x 16 grid, each thread a 4 x 4 micro-tile of S and a 4 x (D/16)
// micro-tile of the D-wide accumulators.  Rows past N are loaded as zeros and never stored.
#pragma once

#include "fa_dropout.cuh"
#include "sm100_ptx.cuh"

namespace fa {

struct SimtParams {
  const float *q, *k, *v, *o, *dout;
  const float *lse, *delta;  // (B,H,N) contiguous, log2 units / fp32
  float *out_o, *out_lse;    // forward outputs
  float *dq, *dk, *dv;       // backward outputs
  int B, H, N, D;
  int64_t q_s[3], k_s[3], v_s[3], o_s[3], do_s[3], dq_s[3], dk_s[3], dv_s[3];  // {sB,sH,sN} in elements
  float scale;       // softmax scale
  float scale_log2;  // scale * log2(e)
  int causal;
  const int* seqlens;  // per-batch valid length (key-padding mask), nullptr = N; rows past it are loaded as zeros
  DropParams drop;     // dropout of the attention probabilities (thresh = 0: off), fa_dropout.cuh
  // arbitrary attention mask, one bit per entry [.., query, key / 8], 1 = attend (nullptr: none); strides in bytes, sB / sH may be 0
  const uint8_t* amask;
  int64_t am_s[3];
  const uint8_t* ablock;   // optional 128 x 128 block summary [.., query block, key block]: 0 = nothing visible, skip
  int64_t ab_s[3];
  int band, win_left, win_right;   // band mask (amask == nullptr, band != 0): -win_left <= key - query <= win_right
};

// mask of one (b, h): bytes and/or band
struct SimtMask {
  const uint8_t* bytes;
  int64_t sN;
  int band, wl, wr;
  __device__ __forceinline__ bool any() const { return bytes != nullptr || band != 0; }
  __device__ __forceinline__ bool visible(int row, int col) const {
    if (band && (col - row < -wl || col - row > wr)) return false;
    return bytes == nullptr || ((__ldg(bytes + (int64_t)row * sN + (col >> 3)) >> (col & 7)) & 1) != 0;
  }
};
__device__ __forceinline__ SimtMask simt_mask(const SimtParams& p, int b, int h) {
  return SimtMask{p.amask ? p.amask + b * p.am_s[0] + h * p.am_s[1] : nullptr, p.am_s[2], p.band, p.win_left, p.win_right};
}

constexpr int kSimtTile = 64;
// columns of a D-wide accumulator one thread owns: 4 per 64-wide column group it participates in
template <int kD>
constexpr int kSimtAccCols = (kD >= 64) ? kD / 16 : 4;
constexpr int kSimtLdT = 68;  // leading dim of the transposed [d][row] and of the [row][col] score tiles

template <int kD>
struct SimtSmem {
  static constexpr int kLdR = kD + 4;  // leading dim of row-major [row][d] tiles
  static constexpr int fwd_floats = kD * kSimtLdT /*Qt*/ + kD * kSimtLdT /*Kt*/ + 64 * kD /*V*/ + 64 * kSimtLdT /*P*/;
  static constexpr int dkdv_floats = 2 * kD * kSimtLdT /*Kt,Vt*/ + 2 * 64 * kLdR /*Q,dO*/ + 2 * 64 * kSimtLdT /*P,dS*/;
  static constexpr int dq_floats =
      2 * kD * kSimtLdT /*Kt,Vt*/ + 64 * kD /*K*/ + 2 * 64 * kLdR /*Q,dO*/ + 64 * kSimtLdT /*dS*/;
};

// global [64 rows x kD] (row stride `ld` elements) -> smem transposed dst[d * 68 + r]; rows >= rows_valid are zero.
template <int kD>
__device__ __forceinline__ void simt_load_transposed(float* dst, const float* src, int64_t ld, int rows_valid) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = (warp & 1) * 32 + lane;
  for (int c = warp >> 1; c < kD / 4; c += 4) {
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < rows_valid) x = *reinterpret_cast<const float4*>(src + (int64_t)r * ld + c * 4);
    dst[(c * 4 + 0) * kSimtLdT + r] = x.x;
    dst[(c * 4 + 1) * kSimtLdT + r] = x.y;
    dst[(c * 4 + 2) * kSimtLdT + r] = x.z;
    dst[(c * 4 + 3) * kSimtLdT + r] = x.w;
  }
}
// global [64 rows x kD] -> smem row-major dst[r * ldd + d]; rows >= rows_valid are zero.
template <int kD>
__device__ __forceinline__ void simt_load_rowmajor(float* dst, int ldd, const float* src, int64_t ld, int rows_valid) {
  for (int idx = threadIdx.x; idx < 64 * (kD / 4); idx += 256) {
    const int r = idx / (kD / 4), c = idx % (kD / 4);
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < rows_valid) x = *reinterpret_cast<const float4*>(src + (int64_t)r * ld + c * 4);
    *reinterpret_cast<float4*>(dst + r * ldd + c * 4) = x;
  }
}

// acc[i][j] = sum_d At[d][ty*4+i] * Bt[d][tx*4+j]          (both operands stored transposed, ld = 68)
template <int kD>
__device__ __forceinline__ void simt_scores_tt(float (&acc)[4][4], const float* At, const float* Bt, int ty, int tx) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 8
  for (int d = 0; d < kD; ++d) {
    const float4 a = *reinterpret_cast<const float4*>(At + d * kSimtLdT + ty * 4);
    const float4 b = *reinterpret_cast<const float4*>(Bt + d * kSimtLdT + tx * 4);
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}
// acc[i][j] = sum_d A[ty*4+i][d] * Bt[d][tx*4+j]           (A row-major with leading dim lda)
template <int kD>
__device__ __forceinline__ void simt_scores_rt(float (&acc)[4][4], const float* A, int lda, const float* Bt, int ty,
                                               int tx) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 2
  for (int d = 0; d < kD; d += 4) {
    float av[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 a = *reinterpret_cast<const float4*>(A + (ty * 4 + i) * lda + d);
      av[i][0] = a.x, av[i][1] = a.y, av[i][2] = a.z, av[i][3] = a.w;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float4 b = *reinterpret_cast<const float4*>(Bt + (d + e) * kSimtLdT + tx * 4);
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i][e], bv[j], acc[i][j]);
    }
  }
}
// acc[i][cc*4+j] += sum_k W[(ty*4+i) , k] * X[k][cc*64 + tx*4 + j]   with W given as Wrow[r * 68 + k] (kTransW = false)
// or as its transpose Wt[k * 68 + r] (kTransW = true); X row-major with leading dim ldx; k runs over 64.
template <int kD, bool kTransW>
__device__ __forceinline__ void simt_accum_wide(float (&acc)[4][kSimtAccCols<kD>], const float* W, const float* X, int ldx, int ty,
                                                int tx) {
#pragma unroll 4
  for (int k = 0; k < 64; ++k) {
    float w[4];
    if constexpr (kTransW) {
      const float4 t = *reinterpret_cast<const float4*>(W + k * kSimtLdT + ty * 4);
      w[0] = t.x, w[1] = t.y, w[2] = t.z, w[3] = t.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) w[i] = W[(ty * 4 + i) * kSimtLdT + k];
    }
#pragma unroll
    for (int cc = 0; cc < kD / 64 + (kD < 64 ? 1 : 0); ++cc) {
      if (cc * 64 + tx * 4 < kD) {
        const float4 x = *reinterpret_cast<const float4*>(X + k * ldx + cc * 64 + tx * 4);
        const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][cc * 4 + j] = fmaf(w[i], xv[j], acc[i][cc * 4 + j]);
      }
    }
  }
}

__device__ __forceinline__ float half16_max(float v) {
#pragma unroll
  for (int off = 8; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;
}
__device__ __forceinline__ float half16_sum(float v) {
#pragma unroll
  for (int off = 8; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// ------------------------------------------------------------------------------------------------ forward
template <int kD>
__global__ void __launch_bounds__(256) fa_fwd_f32_kernel(const SimtParams p) {
  extern __shared__ __align__(16) float smem_f[];
  float* Qt = smem_f;
  float* Kt = Qt + kD * kSimtLdT;
  float* Vs = Kt + kD * kSimtLdT;
  float* Ps = Vs + 64 * kD;
  constexpr int kAcc = kSimtAccCols<kD>;

  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int qb = gridDim.x - 1 - blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qb * 64;
  const int nv = p.seqlens ? min(max(p.seqlens[b], 0), p.N) : p.N;   // valid length of this batch element
  if (q0 >= nv) return;
  const float* qp = p.q + b * p.q_s[0] + h * p.q_s[1];
  const float* kp = p.k + b * p.k_s[0] + h * p.k_s[1];
  const float* vp = p.v + b * p.v_s[0] + h * p.v_s[1];

  simt_load_transposed<kD>(Qt, qp + (int64_t)q0 * p.q_s[2], p.q_s[2], nv - q0);

  float m[4], l[4], acc[4][kAcc];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[i] = -INFINITY;
    l[i] = 0.f;
#pragma unroll
    for (int j = 0; j < kAcc; ++j) acc[i][j] = 0.f;
  }
  const uint32_t drop_thresh = p.drop.thresh;
  const uint32_t dkey = drop_thresh ? drop_key(p.drop, b * p.H + h) : 0u;
  const SimtMask am = simt_mask(p, b, h);
  const int n_kv = p.causal ? min((nv + 63) / 64, qb + 1) : (nv + 63) / 64;
  const uint8_t* ab_q = p.ablock ? p.ablock + b * p.ab_s[0] + h * p.ab_s[1] + (int64_t)(q0 >> 7) * p.ab_s[2] : nullptr;
  for (int jb = 0; jb < n_kv; ++jb) {
    const int k0 = jb * 64;
    if (ab_q && !ab_q[k0 >> 7]) continue;   // nothing visible in this block (uniform over the CTA)
    __syncthreads();  // previous iteration's readers of Kt / Vs / Ps are done (also covers the Qt fill)
    simt_load_transposed<kD>(Kt, kp + (int64_t)k0 * p.k_s[2], p.k_s[2], nv - k0);
    simt_load_rowmajor<kD>(Vs, kD, vp + (int64_t)k0 * p.v_s[2], p.v_s[2], nv - k0);
    __syncthreads();
    float s[4][4];
    simt_scores_tt<kD>(s, Qt, Kt, ty, tx);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = q0 + ty * 4 + i;
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = k0 + tx * 4 + j;
        const bool dead = (col >= nv) || (p.causal && col > row) ||
                          (row < nv && !am.visible(row, col));
        s[i][j] = dead ? -INFINITY : s[i][j] * p.scale_log2;
        mx = fmaxf(mx, s[i][j]);
      }
      mx = half16_max(mx);
      const float m_new = fmaxf(m[i], mx);
      const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = exp2f(m[i] - m_safe);  // m[i] = -inf on the first block -> 0
      float rs = 0.f;
      float4 pr;
      float* prv = reinterpret_cast<float*>(&pr);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        prv[j] = exp2f(s[i][j] - m_safe);
        rs += prv[j];   // the row sum is that of the undropped P
        if (drop_thresh && !drop_keep(dkey, row, k0 + tx * 4 + j, drop_thresh)) prv[j] = 0.f;
      }
      rs = half16_sum(rs);
      l[i] = l[i] * alpha + rs;
      m[i] = m_new;
#pragma unroll
      for (int j = 0; j < kAcc; ++j) acc[i][j] *= alpha;
      *reinterpret_cast<float4*>(Ps + (ty * 4 + i) * kSimtLdT + tx * 4) = pr;
    }
    __syncthreads();
    simt_accum_wide<kD, false>(acc, Ps, Vs, kD, ty, tx);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = q0 + ty * 4 + i;
    if (row >= nv) continue;
    const float inv = l[i] > 0.f ? (drop_thresh ? p.drop.rp : 1.0f) / l[i] : 0.f;   // no visible key (mask): O = 0, L = -inf
    float* orow = p.out_o + b * p.o_s[0] + h * p.o_s[1] + (int64_t)row * p.o_s[2];
#pragma unroll
    for (int cc = 0; cc < kAcc / 4; ++cc) {
      if (cc * 64 + tx * 4 < kD) {
        float4 ov = make_float4(acc[i][cc * 4 + 0] * inv, acc[i][cc * 4 + 1] * inv, acc[i][cc * 4 + 2] * inv,
                                acc[i][cc * 4 + 3] * inv);
        *reinterpret_cast<float4*>(orow + cc * 64 + tx * 4) = ov;
      }
    }
    if (tx == 0) p.out_lse[((int64_t)b * p.H + h) * p.N + row] = m[i] + log2f(l[i]);
  }
}

// P and dS for one 64x64 tile from raw S = Q K^T and dP = dO V^T (shared by both backward kernels).
// s <- P, dp <- dS (unscaled: the softmax scale is applied once to the finished dQ / dK accumulators).
// Dropout (thresh != 0): s <- keep o P (the 1 / (1 - p) goes into the dV epilogue), dS = P o (keep * rp * dP - delta).
__device__ __forceinline__ void simt_p_ds(float (&s)[4][4], float (&dp)[4][4], const float (&lse)[4],
                                          const float (&dl)[4], int q0, int k0, int ty, int tx, int N, int causal,
                                          float scale_log2, uint32_t dkey, uint32_t drop_thresh, float drop_rp,
                                          const SimtMask& am) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = q0 + ty * 4 + i;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = k0 + tx * 4 + j;
      bool dead = (col >= N) || (row >= N) || (causal && col > row);
      if (!dead && am.any()) dead = !am.visible(row, col) || lse[i] == -INFINITY;
      const float pv = dead ? 0.f : exp2f(fmaf(s[i][j], scale_log2, -lse[i]));
      if (drop_thresh) {
        const bool keep = drop_keep(dkey, row, col, drop_thresh);
        s[i][j] = keep ? pv : 0.f;
        dp[i][j] = pv * ((keep ? drop_rp * dp[i][j] : 0.f) - dl[i]);
      } else {
        s[i][j] = pv;
        dp[i][j] = pv * (dp[i][j] - dl[i]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ dK / dV
template <int kD>
__global__ void __launch_bounds__(256) fa_bwd_dkdv_f32_kernel(const SimtParams p) {
  using SS = SimtSmem<kD>;
  extern __shared__ __align__(16) float smem_f[];
  float* Kt = smem_f;
  float* Vt = Kt + kD * kSimtLdT;
  float* Qs = Vt + kD * kSimtLdT;
  float* dOs = Qs + 64 * SS::kLdR;
  float* Ps = dOs + 64 * SS::kLdR;
  float* dSs = Ps + 64 * kSimtLdT;
  constexpr int kAcc = kSimtAccCols<kD>;

  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int jb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int k0 = jb * 64;
  const int nv = p.seqlens ? min(max(p.seqlens[b], 0), p.N) : p.N;   // valid length of this batch element
  if (k0 >= nv) return;
  const float* qp = p.q + b * p.q_s[0] + h * p.q_s[1];
  const float* dop = p.dout + b * p.do_s[0] + h * p.do_s[1];
  const float* lsep = p.lse + ((int64_t)b * p.H + h) * p.N;
  const float* dlp = p.delta + ((int64_t)b * p.H + h) * p.N;

  simt_load_transposed<kD>(Kt, p.k + b * p.k_s[0] + h * p.k_s[1] + (int64_t)k0 * p.k_s[2], p.k_s[2], nv - k0);
  simt_load_transposed<kD>(Vt, p.v + b * p.v_s[0] + h * p.v_s[1] + (int64_t)k0 * p.v_s[2], p.v_s[2], nv - k0);

  const uint32_t dkey = p.drop.thresh ? drop_key(p.drop, b * p.H + h) : 0u;
  const float dv_mul = p.drop.thresh ? p.drop.rp : 1.0f;
  const SimtMask am = simt_mask(p, b, h);
  float dk[4][kAcc], dv[4][kAcc];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < kAcc; ++j) dk[i][j] = 0.f, dv[i][j] = 0.f;

  const int n_q = (nv + 63) / 64;
  const uint8_t* ab_k = p.ablock ? p.ablock + b * p.ab_s[0] + h * p.ab_s[1] + (k0 >> 7) : nullptr;
  for (int ib = p.causal ? jb : 0; ib < n_q; ++ib) {
    const int q0 = ib * 64;
    if (ab_k && !ab_k[(int64_t)(q0 >> 7) * p.ab_s[2]]) continue;   // nothing visible in this block
    __syncthreads();
    simt_load_rowmajor<kD>(Qs, SS::kLdR, qp + (int64_t)q0 * p.q_s[2], p.q_s[2], nv - q0);
    simt_load_rowmajor<kD>(dOs, SS::kLdR, dop + (int64_t)q0 * p.do_s[2], p.do_s[2], nv - q0);
    __syncthreads();
    float s[4][4], dp[4][4], lse[4], dl[4];
    simt_scores_rt<kD>(s, Qs, SS::kLdR, Kt, ty, tx);
    simt_scores_rt<kD>(dp, dOs, SS::kLdR, Vt, ty, tx);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = q0 + ty * 4 + i;
      lse[i] = row < nv ? lsep[row] : 0.f;
      dl[i] = row < nv ? dlp[row] : 0.f;
    }
    simt_p_ds(s, dp, lse, dl, q0, k0, ty, tx, nv, p.causal, p.scale_log2, dkey, p.drop.thresh, p.drop.rp, am);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      *reinterpret_cast<float4*>(Ps + (ty * 4 + i) * kSimtLdT + tx * 4) = make_float4(s[i][0], s[i][1], s[i][2], s[i][3]);
      *reinterpret_cast<float4*>(dSs + (ty * 4 + i) * kSimtLdT + tx * 4) =
          make_float4(dp[i][0], dp[i][1], dp[i][2], dp[i][3]);
    }
    __syncthreads();
    // dV[c][:] += sum_r P[r][c] dO[r][:] ; dK[c][:] += sum_r dS[r][c] Q[r][:]   (Ps / dSs read as their transposes)
    simt_accum_wide<kD, true>(dv, Ps, dOs, SS::kLdR, ty, tx);
    simt_accum_wide<kD, true>(dk, dSs, Qs, SS::kLdR, ty, tx);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = k0 + ty * 4 + i;
    if (row >= nv) continue;
    float* dkrow = p.dk + b * p.dk_s[0] + h * p.dk_s[1] + (int64_t)row * p.dk_s[2];
    float* dvrow = p.dv + b * p.dv_s[0] + h * p.dv_s[1] + (int64_t)row * p.dv_s[2];
#pragma unroll
    for (int cc = 0; cc < kAcc / 4; ++cc) {
      if (cc * 64 + tx * 4 < kD) {
        *reinterpret_cast<float4*>(dkrow + cc * 64 + tx * 4) =
            make_float4(dk[i][cc * 4 + 0] * p.scale, dk[i][cc * 4 + 1] * p.scale, dk[i][cc * 4 + 2] * p.scale,
                        dk[i][cc * 4 + 3] * p.scale);
        *reinterpret_cast<float4*>(dvrow + cc * 64 + tx * 4) =
            make_float4(dv[i][cc * 4 + 0] * dv_mul, dv[i][cc * 4 + 1] * dv_mul, dv[i][cc * 4 + 2] * dv_mul,
                        dv[i][cc * 4 + 3] * dv_mul);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ dQ
template <int kD>
__global__ void __launch_bounds__(256) fa_bwd_dq_f32_kernel(const SimtParams p) {
  using SS = SimtSmem<kD>;
  extern __shared__ __align__(16) float smem_f[];
  float* Kt = smem_f;
  float* Vt = Kt + kD * kSimtLdT;
  float* Ks = Vt + kD * kSimtLdT;
  float* Qs = Ks + 64 * kD;
  float* dOs = Qs + 64 * SS::kLdR;
  float* dSs = dOs + 64 * SS::kLdR;
  constexpr int kAcc = kSimtAccCols<kD>;

  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int ib = gridDim.x - 1 - blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = ib * 64;
  const int nv = p.seqlens ? min(max(p.seqlens[b], 0), p.N) : p.N;   // valid length of this batch element
  if (q0 >= nv) return;
  const float* kp = p.k + b * p.k_s[0] + h * p.k_s[1];
  const float* vp = p.v + b * p.v_s[0] + h * p.v_s[1];

  simt_load_rowmajor<kD>(Qs, SS::kLdR, p.q + b * p.q_s[0] + h * p.q_s[1] + (int64_t)q0 * p.q_s[2], p.q_s[2], nv - q0);
  simt_load_rowmajor<kD>(dOs, SS::kLdR, p.dout + b * p.do_s[0] + h * p.do_s[1] + (int64_t)q0 * p.do_s[2], p.do_s[2],
                         nv - q0);
  const uint32_t dkey = p.drop.thresh ? drop_key(p.drop, b * p.H + h) : 0u;
  const SimtMask am = simt_mask(p, b, h);
  float lse[4], dl[4], dq[4][kAcc];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = q0 + ty * 4 + i;
    lse[i] = row < nv ? p.lse[((int64_t)b * p.H + h) * p.N + row] : 0.f;
    dl[i] = row < nv ? p.delta[((int64_t)b * p.H + h) * p.N + row] : 0.f;
#pragma unroll
    for (int j = 0; j < kAcc; ++j) dq[i][j] = 0.f;
  }
  const int n_kv = p.causal ? min((nv + 63) / 64, ib + 1) : (nv + 63) / 64;
  const uint8_t* ab_q = p.ablock ? p.ablock + b * p.ab_s[0] + h * p.ab_s[1] + (int64_t)(q0 >> 7) * p.ab_s[2] : nullptr;
  for (int jb = 0; jb < n_kv; ++jb) {
    const int k0 = jb * 64;
    if (ab_q && !ab_q[k0 >> 7]) continue;   // nothing visible in this block
    __syncthreads();
    simt_load_transposed<kD>(Kt, kp + (int64_t)k0 * p.k_s[2], p.k_s[2], nv - k0);
    simt_load_transposed<kD>(Vt, vp + (int64_t)k0 * p.v_s[2], p.v_s[2], nv - k0);
    simt_load_rowmajor<kD>(Ks, kD, kp + (int64_t)k0 * p.k_s[2], p.k_s[2], nv - k0);
    __syncthreads();
    float s[4][4], dp[4][4];
    simt_scores_rt<kD>(s, Qs, SS::kLdR, Kt, ty, tx);
    simt_scores_rt<kD>(dp, dOs, SS::kLdR, Vt, ty, tx);
    simt_p_ds(s, dp, lse, dl, q0, k0, ty, tx, nv, p.causal, p.scale_log2, dkey, p.drop.thresh, p.drop.rp, am);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4*>(dSs + (ty * 4 + i) * kSimtLdT + tx * 4) =
          make_float4(dp[i][0], dp[i][1], dp[i][2], dp[i][3]);
    __syncthreads();
    simt_accum_wide<kD, false>(dq, dSs, Ks, kD, ty, tx);  // dQ[r][:] += sum_c dS[r][c] K[c][:]
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = q0 + ty * 4 + i;
    if (row >= nv) continue;
    float* dqrow = p.dq + b * p.dq_s[0] + h * p.dq_s[1] + (int64_t)row * p.dq_s[2];
#pragma unroll
    for (int cc = 0; cc < kAcc / 4; ++cc) {
      if (cc * 64 + tx * 4 < kD)
        *reinterpret_cast<float4*>(dqrow + cc * 64 + tx * 4) =
            make_float4(dq[i][cc * 4 + 0] * p.scale, dq[i][cc * 4 + 1] * p.scale, dq[i][cc * 4 + 2] * p.scale,
                        dq[i][cc * 4 + 3] * p.scale);
    }
  }
}

}  // namespace fa
