// sm100_ptx.cuh — hand-written inline-PTX building blocks for sm_100a (B200).
//
// Everything the attention kernels need from the Blackwell programming model:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld / st /
// fences), UMMA shared-memory + instruction descriptors.  No CUTLASS/CuTe types.
//
// Descriptor bit layouts follow the PTX ISA "tcgen05 matrix / instruction
// descriptor" tables (same fields CUTLASS names in cute/arch/mma_sm100_desc.hpp).
#pragma once

#include <cstdint>
#include <cstdio>
#include <type_traits>
#include <cuda.h>          // CUtensorMap (type only; the encode function is fetched at run time)
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace fa {

// ----------------------------------------------------------------------------------------------
// Debug watchdog (debug builds only: -DFA_WATCHDOG=1, `FA_B200_DEBUG=1 python -c "import __graft_entry__ as g; g.build()"`):
// every mbarrier wait gives up after ~2 s of GPU time and traps, so a protocol bug surfaces as a launch
// failure instead of a hung box.  Off in the production library: the slow path of every wait is then a bare
// try_wait loop (no clock64, no printf, fewer live registers in the hot loops).
// ----------------------------------------------------------------------------------------------
#ifndef FA_WATCHDOG
#define FA_WATCHDOG 0
#endif
#ifndef FA_WATCHDOG_CYCLES
#define FA_WATCHDOG_CYCLES 4000000000LL
#endif

// FA_TRACE (debug builds only): CTA (0,0,0) of the kernels appends {event id, clock64} pairs to a device buffer set with
// fa_debug_set_trace(); used to reconstruct the warp-role timeline.  Compiled out of the production library.
#ifndef FA_TRACE
#define FA_TRACE 0
#endif
#ifndef FA_TRACE_BLOCK
#define FA_TRACE_BLOCK 0
#endif
#if FA_TRACE
__device__ long long* g_fa_trace = nullptr;
__device__ int g_fa_trace_cap = 0;
// slot = role * 8192 + iteration * 8 + k : plain store, no atomics (an atomic's round trip would distort the timeline)
__device__ __forceinline__ void fa_trace(int role, int it, int k) {
  if (g_fa_trace != nullptr && blockIdx.x == FA_TRACE_BLOCK && blockIdx.y == 0 && blockIdx.z == 0) {
    const int slot = role * 8192 + it * 8 + k;
    if (slot < g_fa_trace_cap) g_fa_trace[slot] = clock64();
  }
}
#else
__device__ __forceinline__ void fa_trace(int, int, int) {}
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "elect.sync _|P1, %1;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(pred)
      : "r"(0xffffffffu));
  return pred != 0;
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// generic-proxy writes to smem -> visible to the async proxy (TMA store, UMMA operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
#if FA_WATCHDOG
  long long t0 = clock64();
#endif
  while (!mbar_try_wait(bar, parity)) {
#if FA_WATCHDOG
    if (clock64() - t0 > FA_WATCHDOG_CYCLES) {
      printf("[fa watchdog] mbarrier wait timed out: block (%d,%d,%d) thread %d bar smem 0x%x parity %u\n",
             blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
#endif
  }
}

// ----------------------------------------------------------------------------------------------
// TMA
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 4-D tiled load global -> shared, completion on an mbarrier (bytes).
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "r"(c2), "r"(c3)
      : "memory");
}
// 4-D tiled store shared -> global (bulk-group completion).
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2,
                                             int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
      :
      : "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, fences, commit
// ----------------------------------------------------------------------------------------------
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// All previously issued tcgen05.mma of this thread arrive on `bar` when they complete
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ----------------------------------------------------------------------------------------------
// CTA pairs (thread-block cluster of 2, tcgen05 cta_group::2): one MMA stream issued by the leader CTA (cluster rank 0)
// computes M = 256 rows, 128 per CTA, each CTA supplying its own A rows and HALF of the B operand (N split) — the pair
// reads every B tile from shared memory once instead of twice.  Validated on hardware by tools/umma2_probe.cu.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {   // every thread of both CTAs
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
// arrive (release, cluster scope) on an mbarrier given by its shared::cluster address (possibly in the other CTA)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result) {  // one full warp, the same warp index in both CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// TMA load into THIS CTA's shared memory; the completion bytes are signalled on the mbarrier at cluster address
// `mbar_cluster` (the leader's barrier: the MMA issuer waits for both CTAs' halves on one barrier)
__device__ __forceinline__ void tma_load_4d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t mbar_cluster, int c0,
                                                int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar_cluster), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3)
      : "memory");
}
// all MMAs issued so far by this thread -> one arrival on the barrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void tc_commit2(uint64_t* bar, uint16_t mask = 3) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

// ----------------------------------------------------------------------------------------------
// UMMA descriptors
// ----------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_128B, sm_100 version bits.
//   bits [ 0,14) start address >> 4        bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4   bits [46,48) version = 1
//   bits [49,52) base offset = 0 (tiles are 1024-B aligned)   bits [61,64) layout type (2 = 128B swizzle)
__device__ __forceinline__ uint64_t umma_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes,
                                                        uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// Tiles in this project are stored as TMA SWIZZLE_128B boxes: `rows` x 128 bytes (64 x 16-bit),
// row r at byte r*128, 16-byte chunk c of row r at chunk position c ^ (r & 7).
//
// K-major operand (the contraction index runs along the 128-byte row):
//   8-row group stride (SBO) = 1024 B; LBO ignored for swizzled K-major (CUTLASS writes 1).
//   Advancing 16 elements of K = +32 B on the start address (within one box),
//   the next 64 elements of K live in the next box.
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t box_addr, uint32_t k16_in_box) {
  return umma_smem_desc_sw128(box_addr + k16_in_box * 32u, 16u, 1024u);
}
// MN-major operand (the contraction index runs down the rows, M/N runs along the 128-byte row):
//   64-element column groups live in successive boxes -> LBO = box stride in bytes,
//   8 K-rows = 1024 B -> SBO = 1024. Advancing 16 of K = +16 rows = +2048 B.
__device__ __forceinline__ uint64_t umma_desc_mnmajor(uint32_t box0_addr, uint32_t box_stride_bytes,
                                                     uint32_t k16) {
  return umma_smem_desc_sw128(box0_addr + k16 * 2048u, box_stride_bytes, 1024u);
}

// Cheap per-MMA descriptor arithmetic for the issue loops.  With SWIZZLE_128B, SBO = 1024 B and version 1 the high
// word of every descriptor used here is the same constant; the low word is (address >> 4) | (LBO >> 4) << 16, and
// stepping through K (or selecting a ring stage / a 64-row half) only adds a constant to the low word (shared
// memory addresses stay below 2^18, so the 14-bit address field never carries).
constexpr uint32_t kUmmaDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t umma_lo_kmajor(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint32_t umma_lo_mnmajor(uint32_t saddr, uint32_t box_stride_bytes) {
  return ((saddr & 0x3FFFFu) >> 4) | ((box_stride_bytes >> 4) << 16);
}
// low-word offset of K step k (16 elements) inside a K-major tile made of 64-column boxes of `box_bytes`
__host__ __device__ constexpr uint32_t umma_koff_kmajor(int k, int box_bytes) {
  return (uint32_t)(((k / 4) * box_bytes + (k % 4) * 32) >> 4);
}
// low-word offset of K step k16 (16 rows of 128 B) inside an MN-major tile
__host__ __device__ constexpr uint32_t umma_koff_mnmajor(int k16) { return (uint32_t)((k16 * 2048) >> 4); }

__device__ __forceinline__ void umma_ss_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kUmmaDescHi)
      : "memory");
}
__device__ __forceinline__ void umma_ts_lo(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kUmmaDescHi)
      : "memory");
}

// Issue-loop variants: operand = base low word (kept in one uniform register) + COMPILE-TIME offset, added inside the
// asm block.  If the additions are visible to the compiler it hoists all of them out of the block loop, runs out of
// uniform registers and re-materialises every descriptor through R2UR / spills (~40 clk per MMA, measured with
// tools/trace_dq.py), which made the single issuing thread the bottleneck of the backward kernels.
// FA_ABLATE_SAMEDESC=1 (timing ablation, wrong results): every MMA of a group uses the group's first operand addresses,
// so no descriptor arithmetic is left between the instructions — the upper bound of what a cheaper issue loop could give.
#ifndef FA_ABLATE_SAMEDESC
#define FA_ABLATE_SAMEDESC 0
#endif
template <uint32_t kOffA_, uint32_t kOffB_>
__device__ __forceinline__ void umma_ss_off(uint32_t tmem_d, uint32_t a_base, uint32_t b_base, uint32_t idesc,
                                            uint32_t accumulate) {
  constexpr uint32_t kOffA = FA_ABLATE_SAMEDESC ? 0u : kOffA_, kOffB = FA_ABLATE_SAMEDESC ? 0u : kOffB_;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b32 al, bl;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "add.u32 al, %1, %6;\n\t"
      "add.u32 bl, %2, %7;\n\t"
      "mov.b64 da, {al, %5};\n\t"
      "mov.b64 db, {bl, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "r"(a_base), "r"(b_base), "r"(idesc), "r"(accumulate), "r"(kUmmaDescHi), "n"(kOffA), "n"(kOffB)
      : "memory");
}
template <uint32_t kOffA_, uint32_t kOffB_>
__device__ __forceinline__ void umma_ts_off(uint32_t tmem_d, uint32_t tmem_a_base, uint32_t b_base, uint32_t idesc,
                                            uint32_t accumulate) {
  constexpr uint32_t kOffA = FA_ABLATE_SAMEDESC ? 0u : kOffA_, kOffB = FA_ABLATE_SAMEDESC ? 0u : kOffB_;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b32 ta, bl;\n\t"
      ".reg .b64 db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "add.u32 ta, %1, %6;\n\t"
      "add.u32 bl, %2, %7;\n\t"
      "mov.b64 db, {bl, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "r"(tmem_a_base), "r"(b_base), "r"(idesc), "r"(accumulate), "r"(kUmmaDescHi), "n"(kOffA),
        "n"(kOffB)
      : "memory");
}
// cta_group::2 variants (issued by the leader CTA only; descriptors / TMEM addresses are the same offsets in both CTAs)
template <uint32_t kOffA, uint32_t kOffB>
__device__ __forceinline__ void umma2_ss_off(uint32_t tmem_d, uint32_t a_base, uint32_t b_base, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b32 al, bl;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "add.u32 al, %1, %6;\n\t"
      "add.u32 bl, %2, %7;\n\t"
      "mov.b64 da, {al, %5};\n\t"
      "mov.b64 db, {bl, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "r"(a_base), "r"(b_base), "r"(idesc), "r"(accumulate), "r"(kUmmaDescHi), "n"(kOffA), "n"(kOffB)
      : "memory");
}
template <uint32_t kOffA, uint32_t kOffB>
__device__ __forceinline__ void umma2_ts_off(uint32_t tmem_d, uint32_t tmem_a_base, uint32_t b_base, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b32 ta, bl;\n\t"
      ".reg .b64 db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "add.u32 ta, %1, %6;\n\t"
      "add.u32 bl, %2, %7;\n\t"
      "mov.b64 db, {bl, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [ta], db, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "r"(tmem_a_base), "r"(b_base), "r"(idesc), "r"(accumulate), "r"(kUmmaDescHi), "n"(kOffA),
        "n"(kOffB)
      : "memory");
}
// compile-time loop: f(std::integral_constant<int, I>) for I in [0, N)
template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

// Instruction descriptor for kind::f16 (fp16/bf16 inputs, fp32 accumulate).
//   [4,6) D format (1 = f32)  [7,10) A format  [10,13) B format (0 = f16, 1 = bf16)
//   [15] A major  [16] B major (0 = K-major, 1 = MN-major)  [17,23) N>>3  [24,29) M>>4
// A and B formats are separate fields, but mixing f16 with bf16 faults on B200 (illegal instruction): keep them equal.
__host__ __device__ constexpr uint32_t umma_idesc_f16_ab(int a_bf16, int b_bf16, int M, int N, int a_mn_major,
                                                        int b_mn_major) {
  return (1u << 4) | (static_cast<uint32_t>(a_bf16) << 7) | (static_cast<uint32_t>(b_bf16) << 10) |
         (static_cast<uint32_t>(a_mn_major) << 15) | (static_cast<uint32_t>(b_mn_major) << 16) |
         (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t umma_idesc_f16(int is_bf16, int M, int N, int a_mn_major,
                                                     int b_mn_major) {
  return umma_idesc_f16_ab(is_bf16, is_bf16, M, N, a_mn_major, b_mn_major);
}

// Instruction descriptor for kind::f8f6f4 with 8-bit floating-point inputs (fp32 accumulate): same fields as
// kind::f16; A / B format 0 = E4M3, 1 = E5M2.  One instruction contracts 32 elements (32 bytes) of K.  MN-major
// operands are legal for the 8-bit formats.
__host__ __device__ constexpr uint32_t umma_idesc_f8(int is_e5m2, int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (static_cast<uint32_t>(is_e5m2) << 7) | (static_cast<uint32_t>(is_e5m2) << 10) |
         (static_cast<uint32_t>(a_mn_major) << 15) | (static_cast<uint32_t>(b_mn_major) << 16) |
         (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
// FP8 issue-loop variants of umma_ss_off / umma_ts_off (compile-time offsets added to uniform base low words).
template <uint32_t kOffA, uint32_t kOffB>
__device__ __forceinline__ void umma8_ss_off(uint32_t tmem_d, uint32_t a_base, uint32_t b_base, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b32 al, bl;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "add.u32 al, %1, %6;\n\t"
      "add.u32 bl, %2, %7;\n\t"
      "mov.b64 da, {al, %5};\n\t"
      "mov.b64 db, {bl, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], da, db, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "r"(a_base), "r"(b_base), "r"(idesc), "r"(accumulate), "r"(kUmmaDescHi), "n"(kOffA), "n"(kOffB)
      : "memory");
}
template <uint32_t kOffA, uint32_t kOffB>
__device__ __forceinline__ void umma8_ts_off(uint32_t tmem_d, uint32_t tmem_a_base, uint32_t b_base, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b32 ta, bl;\n\t"
      ".reg .b64 db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "add.u32 ta, %1, %6;\n\t"
      "add.u32 bl, %2, %7;\n\t"
      "mov.b64 db, {bl, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [ta], db, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "r"(tmem_a_base), "r"(b_base), "r"(idesc), "r"(accumulate), "r"(kUmmaDescHi), "n"(kOffA),
        "n"(kOffB)
      : "memory");
}
// four fp32 -> four packed FP8 values (RN, saturating to the largest finite value); `a` goes to bits [0,8).
template <bool kE5M2>
__device__ __forceinline__ uint32_t pack4_f8(float a, float b, float c, float d) {
  uint16_t lo, hi;
  if constexpr (kE5M2) {
    asm("cvt.rn.satfinite.e5m2x2.f32 %0, %1, %2;" : "=h"(lo) : "f"(b), "f"(a));
    asm("cvt.rn.satfinite.e5m2x2.f32 %0, %1, %2;" : "=h"(hi) : "f"(d), "f"(c));
  } else {
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(lo) : "f"(b), "f"(a));
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(hi) : "f"(d), "f"(c));
  }
  return static_cast<uint32_t>(lo) | (static_cast<uint32_t>(hi) << 16);
}

// D[tmem] (+)= A[smem] * B[smem]     (single thread issues)
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ----------------------------------------------------------------------------------------------
// TMEM <-> registers.  32x32b shape: thread i of warp w touches lane 32*(w%4)+i, consecutive columns.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
        "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]),
        "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// ----------------------------------------------------------------------------------------------
// small math / packing helpers
// ----------------------------------------------------------------------------------------------
// FA_ABLATE (timing experiments only, results are wrong): 1 = no MUFU (exp2 -> identity),
// 2 = elementwise stages only move data TMEM -> registers -> TMEM, 3 = elementwise stages skipped.
#ifndef FA_ABLATE
#define FA_ABLATE 0
#endif
__device__ __forceinline__ float ex2_approx(float x) {
#if FA_ABLATE == 1
  return x;
#else
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#endif
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// ---- packed fp32 pairs (sm_100: FFMA2 / FADD2 / FMUL2 process two fp32 lanes per issue slot)
__device__ __forceinline__ uint64_t f32x2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t f32x2_pack_bits(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void f32x2_unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f32x2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t f32x2_add(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t f32x2_mul(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// exp2 of two fp32 values on the FMA pipe (no MUFU): Cody-Waite split x = n + f, |f| <= 0.5, 2^f by a degree-3 minimax
// polynomial (max relative error 7.5e-5, below the half-ulp of the 16-bit P it feeds), 2^n by adding n to the exponent
// field.  The MUFU unit does 4 ex2 per clock per SM sub-partition and is the tightest unit of the softmax stage;
// moving a share of the exponentials here balances it against the issue slots.  Valid for x <= 126; x is clamped at
// -125 (masked scores are -inf).
__device__ __forceinline__ void ex2_poly_x2(float x0, float x1, float& p0, float& p1) {
  x0 = fmaxf(x0, -125.f);
  x1 = fmaxf(x1, -125.f);
  const uint64_t x = f32x2_pack(x0, x1);
  const uint64_t t = f32x2_add(x, f32x2_pack(12582912.f, 12582912.f));          // 1.5 * 2^23: low mantissa bits = n
  const uint64_t n = f32x2_add(t, f32x2_pack(-12582912.f, -12582912.f));
  const uint64_t f = f32x2_fma(n, f32x2_pack(-1.f, -1.f), x);
  uint64_t q = f32x2_fma(f32x2_pack(0.05517164617776871f, 0.05517164617776871f), f,
                         f32x2_pack(0.2426111251115799f, 0.2426111251115799f));
  q = f32x2_fma(q, f, f32x2_pack(0.6932609677314758f, 0.6932609677314758f));
  q = f32x2_fma(q, f, f32x2_pack(0.9999280571937561f, 0.9999280571937561f));
  float q0, q1, t0, t1;
  f32x2_unpack(q, q0, q1);
  f32x2_unpack(t, t0, t1);
  p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(t0) << 23));
  p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(t1) << 23));
}
// 16-byte shared-memory load as two packed fp32 pairs (explicit .shared so the compiler does not fall back to
// generic loads when the pointer's address space is not provable)
__device__ __forceinline__ void lds_f32x2x2(uint32_t saddr, uint64_t& a, uint64_t& b) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(saddr));
}
// per-warpgroup register re-budgeting (all four warps of the warpgroup execute it)
template <int kRegs>
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs));
}
template <int kRegs>
__device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs));
}

// two fp32 -> packed 16-bit pair, round-to-nearest-even; `lo` goes to bits [0,16).
template <bool kBf16>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t r;
  if constexpr (kBf16) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  } else {
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  }
  return r;
}
template <bool kBf16>
__device__ __forceinline__ float2 unpack2(uint32_t v) {
  if constexpr (kBf16) {
    return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
  } else {
    __half2 h = *reinterpret_cast<__half2*>(&v);
    return __half22float2(h);
  }
}

// named barrier among `nthreads` threads (id 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// byte offset of (row, 16-byte chunk) inside a 128-byte-swizzled box whose base is 1024-B aligned
__device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t chunk16) {
  return row * 128u + ((chunk16 ^ (row & 7u)) << 4);
}

}  // namespace fa
