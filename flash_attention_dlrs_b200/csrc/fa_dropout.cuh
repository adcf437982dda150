// fa_dropout.cuh — counter-based dropout mask of the attention probabilities, shared by every kernel of the path.
//
// "dropout ... fused in the kernel" is on the reference's roadmap (README.md:35-37); the reference itself has none, so
// the convention is ours and is restated bit for bit by the oracle (oracle/attention_oracle.py: dropout_keep_mask):
//   key(b,h)      = mix(seed_lo ^ mix(seed_hi + (b*H + h) * 0x9E3779B9))
//   word(i, j)    = mix(key + ((i >> 1) << 15) + (j >> 1))           one 32-bit word per 2 x 2 patch of (query i, key j)
//   byte(i, j)    = (word >> 8 * (2 * (i & 1) + (j & 1))) & 0xff
//   keep(i, j)    = byte >= thresh,   thresh = round(256 * p) in [1, 255]      (drop probability thresh / 256)
//   mix           = the "lowbias32" integer finaliser (xorshift 16, * 0x7feb352d, xorshift 15, * 0x846ca68b, xorshift 16)
// The mask depends only on (seed, b, h, i, j): forward, the dK/dV kernel (which walks the scores transposed) and the dQ
// kernel regenerate it instead of storing it.  The 2 x 2 patch gives both walks two mask bytes per hash.
// Kept probabilities are scaled by 1 / (1 - thresh / 256); the softmax statistics (L) are those of the undropped scores.
// Word indices are distinct for N <= 65536; beyond that rows / columns 65536 apart share words.
#pragma once

#include <cstdint>

namespace fa {

struct DropParams {
  uint32_t thresh;            // 0 = dropout off
  uint32_t seed_lo, seed_hi;
  float rp;                   // 1 / (1 - thresh / 256)
};

__host__ __device__ __forceinline__ uint32_t drop_mix(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}

__host__ __device__ __forceinline__ uint32_t drop_key(const DropParams& d, int bh) {
  return drop_mix(d.seed_lo ^ drop_mix(d.seed_hi + (uint32_t)bh * 0x9E3779B9u));
}

// word index of the 2 x 2 patch holding (query i, key j), relative to key(b,h)
__host__ __device__ __forceinline__ uint32_t drop_word_index(int i, int j) {
  return ((uint32_t)(i >> 1) << 15) + (uint32_t)(j >> 1);
}

// scalar form (SIMT kernels)
__host__ __device__ __forceinline__ bool drop_keep(uint32_t key, int i, int j, uint32_t thresh) {
  const uint32_t w = drop_mix(key + drop_word_index(i, j));
  return ((w >> (8 * (2 * (i & 1) + (j & 1)))) & 0xffu) >= thresh;
}

// Pair form (tcgen05 kernels): a thread walks two neighbouring elements of one patch row / column per hash.
//   row walk  (thread = query i, elements = keys j, j + 1, j even):   shift = 16 * (i & 1), kSecond = 8
//   col walk  (thread = key j, elements = queries i, i + 1, i even):  shift = 8 * (j & 1),  kSecond = 16
template <int kSecond>
__device__ __forceinline__ void drop_keep_pair(uint32_t word_index_with_key, uint32_t shift, uint32_t thresh, bool& k0,
                                               bool& k1) {
  const uint32_t w = drop_mix(word_index_with_key) >> shift;
  k0 = (w & 0xffu) >= thresh;
  k1 = ((w >> kSecond) & 0xffu) >= thresh;
}

}  // namespace fa
