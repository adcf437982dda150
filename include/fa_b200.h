/* fa_b200.h — C ABI of the B200-native FlashAttention-2 hot path (libfa_b200.so).
 *
 * Drop-in boundary for the attention path of 17ex/flash_attention_dlrs.  Each entry point replaces one
 * Triton kernel launch the reference's torch boundary performs; the reference call site it stands in for is
 * cited beside it (paths relative to the reference's src/).
 *
 * Conventions
 *   - Tensors are (B, H, N, D) with ELEMENT strides {sB, sH, sN, sD}; sD must be 1.  For 16-bit dtypes the
 *     other strides must be multiples of 8 elements and base pointers 16-byte aligned (TMA requirement).
 *   - dtype: 0 = float16, 1 = bfloat16, 2 = float32, 3 / 4 = FP8 E4M3 / E5M2 (fa_fwd only, D = 128, byte strides that are
 *     multiples of 16).   D in {64, 128} for 16-bit, D in {16, 32, 64, 128} for
 *     float32 (the Python boundary pads other head sizes, as flash_attention_torch.py:38-47 does).
 *   - lse / delta are contiguous fp32 (B, H, N); lse is in LOG2 units: lse = log2(e) * logsumexp_j(scale*S_ij)
 *     (flash_attention_kernels.py:106).
 *   - Every call is asynchronous on `stream` (a cudaStream_t), never synchronises, allocates nothing on the
 *     device and keeps no pointer after returning.  The caller owns all buffers.
 *   - Return value: 0 = OK, < 0 = argument / unsupported-shape error, > 0 = cudaError_t.  A description of the
 *     last error on the calling thread is available from fa_last_error().
 *   - Thread safe and re-entrant; CUDA-graph capturable.
 */
#ifndef FA_B200_H_
#define FA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FA_DTYPE_F16 0
#define FA_DTYPE_BF16 1
#define FA_DTYPE_F32 2
#define FA_DTYPE_F8E4M3 3 /* forward only, D = 128 */
#define FA_DTYPE_F8E5M2 4 /* forward only, D = 128; the FP8 type of the reference's dtype map, flash_attention_torch.py:15-16 */

/* An arbitrary attention mask (the general form of the "masking" on the reference's roadmap, README.md:35-37), ANDed with
 * `causal` and `seqlens`.  All pointers are device pointers; strides are {sB, sH, sRow} in bytes, sB / sH may be 0 (one
 * mask for all batch elements / heads).
 *   rows   [.., query, key / 8]: one BIT per entry, 1 = attend; key j of a row is bit (j & 7) of byte (j >> 3).  Read by
 *          the forward and the dQ kernel.  Row pitch a multiple of 16 bytes and at least 16 bytes per 128 keys (a key
 *          block is one 16-byte load per row), base 16-byte aligned.
 *   cols   [.., key, query]: the same mask transposed, same layout rules; read by the dK/dV kernel (16-bit dtypes;
 *          float32 reads `rows` only and accepts NULL).  Not used by the forward.
 *   blocks [.., query block, key block] (128 x 128 blocks), one BYTE per block, optional (NULL = none): 0 = no visible entry in the block
 *          (such blocks are skipped: not loaded, no MMAs, no softmax work), 2 = every entry visible (`rows` / `cols` are
 *          not read for it), 1 = mixed.  Up to 512 blocks per row / column.
 * Band mask (sliding-window / local attention): rows == NULL (cols ignored) and window_left, window_right >= 0 — query i
 * sees the keys j with -window_left <= j - i <= window_right; no mask bytes are read at all, `blocks` (computed by the
 * caller from the two numbers) gives the skipping.
 * A query with no visible key gets O = 0, lse = -inf and contributes nothing to the gradients. */
typedef struct fa_attn_mask {
  const uint8_t* rows;
  int64_t rows_strides[3];
  const uint8_t* cols;
  int64_t cols_strides[3];
  const uint8_t* blocks;
  int64_t blocks_strides[3];
  int32_t window_left, window_right; /* used when rows == NULL */
} fa_attn_mask;

/* ABI version of this header (bumped on any signature change). */
int fa_version(void);

/* Thread-local text of the last error returned on this thread ("" if none). */
const char* fa_last_error(void);

/* Forward pass: O = softmax(scale * Q K^T [+ causal mask]) V and lse.
 * Replaces fwd_kernel[grid](...) at flash_attention_torch.py:61-74 and flash_attention_wrappers.py:46-61. */
int fa_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int N, int D,
           const int64_t q_strides[4], const int64_t k_strides[4], const int64_t v_strides[4],
           const int64_t o_strides[4], int dtype, float softmax_scale, int causal, void* stream);

/* fa_fwd with a fused all-gather epilogue for head-sharded multi-GPU runs (16-bit and FP8 dtypes): every row of O is
 * stored to `o` AND, with the same strides, to the `n_peers` (<= 7) peer-mapped windows `peer_o[i]` of the other GPUs'
 * gathered output buffers — NVLink peer-to-peer stores issued by the epilogue warps, overlapped with the remaining tiles.
 * With NVLS, pass a multicast address as `o` and n_peers = 0: the switch replicates each store.  The caller synchronises
 * the ranks (e.g. a barrier on the stream) before anybody reads the gathered buffer.  The reference has no multi-GPU path;
 * this stands where a caller would otherwise follow fa_fwd by an NCCL all-gather of O along H.
 * `seqlens` (device pointer to B int32, or NULL) is a key-padding mask — "masking" on the reference's roadmap
 * (README.md:35-37): batch element b has seqlens[b] valid tokens; keys beyond are masked out, query rows beyond are not
 * computed and their O / lse are left untouched (the caller zero-fills).  All dtypes.
 * `dropout_p` in [0, 1) drops attention probabilities inside the kernel — "dropout ... fused in the kernel" on the same
 * roadmap (README.md:35-37); float16 / bfloat16 / float32.  The probability is quantised to round(256 p) / 256 (0 = off)
 * and kept entries are scaled by 1 / (1 - that); lse stays the logsumexp of the undropped scores.  The keep mask is a
 * pure function of (dropout_seed, b, h, query, key), restated by the oracle (oracle/attention_oracle.py:
 * dropout_keep_mask); fa_bwd_partial regenerates it from the same (dropout_p, dropout_seed).
 * `attn_mask` (NULL = none): an arbitrary attention mask, see fa_attn_mask above.  float16 / bfloat16 / float32. */
int fa_fwd_peers(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int N, int D,
                 const int64_t q_strides[4], const int64_t k_strides[4], const int64_t v_strides[4],
                 const int64_t o_strides[4], int dtype, float softmax_scale, int causal, int n_peers,
                 void* const* peer_o, const int32_t* seqlens, float dropout_p, uint64_t dropout_seed,
                 const fa_attn_mask* attn_mask, void* stream);

/* Backward preprocess: delta[b,h,i] = sum_d O[b,h,i,d] * dO[b,h,i,d]  (fp32 accumulate).
 * Replaces bwd_D_kernel[grid](...) at flash_attention_torch.py:125-133 and flash_attention_wrappers.py:110-118. */
int fa_bwd_preprocess(const void* o, const void* dout, float* delta, int B, int H, int N, int D,
                      const int64_t o_strides[4], const int64_t do_strides[4], int dtype, void* stream);

/* Bytes of 256-byte aligned scratch fa_bwd_partial needs for this problem and choice of kernels (`which`, below):
 * 0 for the two-kernel path and for float32; for FA_BWD_FUSED the fp32 dQ partial sums and turn counters of its
 * ordered dQ reduction (the library zeroes what needs zeroing).  The reference allocates its backward scratch — the dQ
 * lock / flag buffers — in the same place: flash_attention_torch.py:107-109,247, flash_attention_wrappers.py:104-108,123. */
size_t fa_bwd_workspace_bytes(int B, int H, int N, int D, int dtype, int causal, int which);

/* Backward pass: dQ, dK, dV from Q, K, V, dO, lse, delta.  Deterministic: bit-identical across runs.
 * Runs the two-kernel path (FA_BWD_DKDV | FA_BWD_DQ below); `workspace` may be NULL.
 * Replaces bwd_kernel / bwd_deterministic_kernel launches at flash_attention_torch.py:136-154,274-292 and
 * flash_attention_wrappers.py:122-174. */
int fa_bwd(const void* q, const void* k, const void* v, const void* dout, const float* lse, const float* delta,
           void* dq, void* dk, void* dv, void* workspace, size_t workspace_bytes, int B, int H, int N, int D,
           const int64_t q_strides[4], const int64_t k_strides[4], const int64_t v_strides[4],
           const int64_t do_strides[4], const int64_t dq_strides[4], const int64_t dk_strides[4],
           const int64_t dv_strides[4], int dtype, float softmax_scale, int causal, void* stream);

/* Backward with an explicit choice of kernels.  `which` = FA_BWD_DKDV (the dK/dV kernel, owner = key block), FA_BWD_DQ
 * (the dQ kernel, owner = query block, recomputes S and dP) or both — what fa_bwd runs; or FA_BWD_FUSED (16-bit dtypes
 * only): one pass like the reference's bwd_kernel, five matmuls per block pair instead of seven, with dQ accumulated
 * across key blocks by an ORDERED reduction in `workspace` (a turn counter per tile instead of the reference's spin lock,
 * flash_attention_kernels.py:305-320), deterministic as well.  `seqlens` as in fa_fwd_peers (two-kernel path only): gradient
 * rows beyond seqlens[b] are left untouched (the caller zero-fills).  The reference has a single backward launch (flash_attention_torch.py:136-154)
 * whose dQ part is the spin-locked read-modify-write of flash_attention_kernels.py:305-320; here the two parts are
 * separate kernels, exposed for callers that need only some gradients and for per-kernel timing.  Outputs not
 * selected are left untouched (their pointers must still be valid).  `dropout_p`, `dropout_seed`: the values the forward
 * ran with (two-kernel path only; `delta` must come from the dropped-out O, as fa_bwd_preprocess gives).
 * `attn_mask` as in fa_fwd_peers (its `cols` member is required for 16-bit dtypes). */
#define FA_BWD_DKDV 1
#define FA_BWD_DQ 2
#define FA_BWD_FUSED 4
int fa_bwd_partial(const void* q, const void* k, const void* v, const void* dout, const float* lse,
                   const float* delta, void* dq, void* dk, void* dv, void* workspace, size_t workspace_bytes, int B,
                   int H, int N, int D, const int64_t q_strides[4], const int64_t k_strides[4],
                   const int64_t v_strides[4], const int64_t do_strides[4], const int64_t dq_strides[4],
                   const int64_t dk_strides[4], const int64_t dv_strides[4], int dtype, float softmax_scale,
                   int causal, int which, const int32_t* seqlens, float dropout_p, uint64_t dropout_seed,
                   const fa_attn_mask* attn_mask, void* stream);

/* Rectangular attention: Nq query rows against Nk key / value rows (Nk != Nq allowed) — cross-attention, and the step of the
 * sequence-parallel path in which one query chunk meets a whole visiting key / value shard (or a whole query shard one key
 * chunk).  The same kernels as fa_fwd / fa_bwd with the key loops bounded by Nk; float16 / bfloat16, D in {64, 128},
 * non-causal, no seqlens / dropout / mask.  o, dq: (B, H, Nq, D); k, v, dk, dv: (B, H, Nk, D); lse, delta: (B, H, Nq).
 * The reference has square problems only (flash_attention_torch.py:27-32 requires equal shapes). */
int fa_fwd_rect(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Nq, int Nk, int D,
                const int64_t q_strides[4], const int64_t k_strides[4], const int64_t v_strides[4],
                const int64_t o_strides[4], int dtype, float softmax_scale, void* stream);
int fa_bwd_rect(const void* q, const void* k, const void* v, const void* dout, const float* lse, const float* delta,
                void* dq, void* dk, void* dv, int B, int H, int Nq, int Nk, int D, const int64_t q_strides[4],
                const int64_t k_strides[4], const int64_t v_strides[4], const int64_t do_strides[4],
                const int64_t dq_strides[4], const int64_t dk_strides[4], const int64_t dv_strides[4], int dtype,
                float softmax_scale, void* stream);

/* Sequence-parallel (ring) attention helpers — no counterpart in the reference (single GPU); they sit where a caller that
 * shards the SEQUENCE across GPUs combines what fa_fwd / fa_bwd return for one key / value shard at a time.
 * All buffers contiguous; 16-bit partials (dtype 0 / 1), fp32 accumulators.
 *   fa_merge_partial: (o_acc, lse_acc) <- exact combination with the partial (o_part, lse_part) of another key shard:
 *     L = log2(2^La + 2^Lp), O = Oa 2^(La-L) + Op 2^(Lp-L); rows x D, lse per row (log2 units, -inf = nothing seen);
 *     first != 0 initialises the accumulators from the partial.
 *   fa_accumulate:    acc (+)= part  (n elements, n % 8 == 0); first != 0 overwrites.
 *   fa_round_rows:    out (16-bit, round to nearest even) = in (fp32). */
int fa_merge_partial(float* o_acc, float* lse_acc, const void* o_part, const float* lse_part, long long rows, int D,
                     int dtype, int first, void* stream);
int fa_accumulate(float* acc, const void* part, long long n, int dtype, int first, void* stream);
int fa_round_rows(void* out, const float* in, long long n, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FA_B200_H_ */
