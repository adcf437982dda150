// fa_merge.cuh — HBM-bound helpers of the sequence-parallel (ring) attention path (the reference has no
// multi-GPU path).  A rank that holds a query shard sees the keys / values one shard at a time; every visit yields a
// normalised partial output O_s and its base-2 logsumexp L_s (fa_fwd).  Two partials combine exactly:
//     L = log2(2^La + 2^Lp),   O = O_a * 2^(La - L) + O_p * 2^(Lp - L)
// which is the same log2-domain bookkeeping the forward kernel does between key blocks (flash_attention_kernels.py:
// 93-97,105-106), one level up.  The running O stays in fp32 until fa_round_rows writes the 16-bit result.
// fa_accumulate adds 16-bit gradient partials into fp32 accumulators (dQ stays local, dK / dV travel with their K / V).
#pragma once

#include "sm100_ptx.cuh"

namespace fa {

// rows x D, contiguous.  kTPR threads per row (adjacent lanes of one warp), 8 elements (16 bytes of the 16-bit partial)
// per thread.  The loop is uniform per warp, so the __syncwarp between reading and rewriting a row's L is convergent.
template <bool kBf16, int kTPR>
__global__ void __launch_bounds__(256)
fa_merge_partial_kernel(float* __restrict__ o_acc, float* __restrict__ l_acc, const uint16_t* __restrict__ o_part,
                        const float* __restrict__ l_part, long long rows, int first) {
  constexpr int kRowsPerWarp = 32 / kTPR, kRowsPerCta = 8 * kRowsPerWarp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane % kTPR, rloc = lane / kTPR;
  for (long long base = (long long)blockIdx.x * kRowsPerCta + warp * kRowsPerWarp; base < rows;
       base += (long long)gridDim.x * kRowsPerCta) {
    const long long r = base + rloc;
    const bool valid = r < rows;
    float p[8], la = -INFINITY, lp = -INFINITY;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    float4* acc = reinterpret_cast<float4*>(o_acc + (r * kTPR + sub) * 8);
    if (valid) {
      const uint4 pv = *reinterpret_cast<const uint4*>(o_part + (r * kTPR + sub) * 8);
      const uint32_t pw[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 x = unpack2<kBf16>(pw[e]);
        p[2 * e] = x.x, p[2 * e + 1] = x.y;
      }
      lp = l_part[r];
      if (!first) {
        la = l_acc[r];
        a0 = acc[0], a1 = acc[1];
      }
    }
    __syncwarp();   // every lane of a row has read l_acc[r] before its first lane rewrites it
    if (!valid) continue;
    const float m = fmaxf(la, lp);
    // a shard that contributed nothing to this row has L = -inf (fully masked): weight 0, not NaN
    const float wa = (la == -INFINITY) ? 0.f : exp2f(la - m), wp = (lp == -INFINITY) ? 0.f : exp2f(lp - m);
    const float den = wa + wp;
    const float ca = den > 0.f ? wa / den : 0.f, cp = den > 0.f ? wp / den : 0.f;
    a0.x = a0.x * ca + p[0] * cp, a0.y = a0.y * ca + p[1] * cp, a0.z = a0.z * ca + p[2] * cp, a0.w = a0.w * ca + p[3] * cp;
    a1.x = a1.x * ca + p[4] * cp, a1.y = a1.y * ca + p[5] * cp, a1.z = a1.z * ca + p[6] * cp, a1.w = a1.w * ca + p[7] * cp;
    acc[0] = a0, acc[1] = a1;
    if (sub == 0) l_acc[r] = den > 0.f ? m + log2f(den) : -INFINITY;
  }
}

// acc (fp32) (+)= part (16-bit), n elements (multiple of 8).
template <bool kBf16>
__global__ void __launch_bounds__(256)
fa_accumulate_kernel(float* __restrict__ acc, const uint16_t* __restrict__ part, long long n8, int first) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n8; i += (long long)gridDim.x * 256) {
    const uint4 pv = *reinterpret_cast<const uint4*>(part + i * 8);
    const uint32_t pw[4] = {pv.x, pv.y, pv.z, pv.w};
    float4* a = reinterpret_cast<float4*>(acc + i * 8);
    float4 a0 = first ? make_float4(0.f, 0.f, 0.f, 0.f) : a[0], a1 = first ? make_float4(0.f, 0.f, 0.f, 0.f) : a[1];
    const float2 x0 = unpack2<kBf16>(pw[0]), x1 = unpack2<kBf16>(pw[1]), x2 = unpack2<kBf16>(pw[2]),
                 x3 = unpack2<kBf16>(pw[3]);
    a0.x += x0.x, a0.y += x0.y, a0.z += x1.x, a0.w += x1.y;
    a1.x += x2.x, a1.y += x2.y, a1.z += x3.x, a1.w += x3.y;
    a[0] = a0, a[1] = a1;
  }
}

// out (16-bit, RTNE) = in (fp32), n elements (multiple of 8).
template <bool kBf16>
__global__ void __launch_bounds__(256)
fa_round_kernel(uint16_t* __restrict__ out, const float* __restrict__ in, long long n8) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n8; i += (long long)gridDim.x * 256) {
    const float4 a0 = reinterpret_cast<const float4*>(in + i * 8)[0], a1 = reinterpret_cast<const float4*>(in + i * 8)[1];
    uint4 v;
    v.x = pack2<kBf16>(a0.x, a0.y), v.y = pack2<kBf16>(a0.z, a0.w);
    v.z = pack2<kBf16>(a1.x, a1.y), v.w = pack2<kBf16>(a1.z, a1.w);
    *reinterpret_cast<uint4*>(out + i * 8) = v;
  }
}

}  // namespace fa
