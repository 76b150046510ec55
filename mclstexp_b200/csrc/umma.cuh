// Blackwell (sm_100a) primitives as thin inline-PTX wrappers: mbarrier, bulk-copy TMA
// (cp.async.bulk, SASS UBLKCP), tcgen05 MMA / TMEM alloc / TMEM load, UMMA descriptors.
// Encodings follow the PTX ISA; the bit layouts were cross-checked against CUTLASS'
// cute/arch/mma_sm100_desc.hpp (InstrDescriptor, SmemDescriptor).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mclst {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint64_t globaltimer() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must not hang the GPU box -- trap after ~4 s.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = globaltimer();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ffu) == 0 && globaltimer() - t0 > 4000000000ull) __trap();
  }
}

// ---------------------------------------------------------------- bulk copy (TMA engine)
// global -> shared::cta, completion signalled on an mbarrier (complete_tx::bytes).
// size multiple of 16, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void* smem_dst, const void* gsrc, uint32_t bytes,
                                              uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

// Multicast form: the bytes land at the same CTA-relative offset in every CTA of ctaMask and
// complete_tx is signalled on the mbarrier at the same offset in each of them.
__device__ __forceinline__ void bulk_g2s_mcast(void* smem_dst, const void* gsrc, uint32_t bytes,
                                               uint64_t* bar, uint16_t cta_mask, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      ".L2::cache_hint [%0], [%1], %2, [%3], %4, %5;"
      ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask),
        "l"(policy)
      : "memory");
}

// ---------------------------------------------------------------- clusters
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {     // every thread of every CTA
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// 1 in exactly one lane of a converged warp.  The MMA-issuing warps run their loops in uniform
// control flow (all 32 lanes) and only issue under this predicate: with `if (lane == 0)` around the
// whole loop the descriptor arithmetic is per-thread code with R2UR moves in front of every
// tcgen05.mma (~100 cycles per instruction, measured), which outruns the tensor pipe for tiles of
// 128 columns or fewer.
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t r;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(r));
  return r;
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, kind::f16 (fp16/bf16 inputs, fp32 accumulate); one thread.
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives when every previously issued tcgen05.mma of this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// Same, arriving on the mbarrier at this CTA-relative offset in every CTA of cta_mask.
__device__ __forceinline__ void mma_commit_mcast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns: thread i gets row (lane base + i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------- descriptors
// Instruction descriptor, kind::f16: fp32 accumulator, A/B both K-major.
//   [4,6) c_format (1 = f32)  [7,10) a_format  [10,13) b_format (0 = f16, 1 = bf16)
//   [15] a_major  [16] b_major (0 = K-major)  [17,23) N>>3  [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, bool bf16) {
  return (1u << 4) | ((bf16 ? 1u : 0u) << 7) | ((bf16 ? 1u : 0u) << 10) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// Shared-memory matrix descriptor for a K-major operand tile stored as consecutive
// 8-row x 128-byte swizzle atoms (SWIZZLE_128B): row pitch 128 B, atom pitch (SBO) 1024 B.
//   [0,14) start address >> 4   [16,30) LBO >> 4 (unused for swizzled K-major, set to 1)
//   [32,46) SBO >> 4            [46,48) version = 1 (sm_100)   [61,64) layout = 2 (128B swizzle)
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) |
         (2ull << 61);
}

// Same atoms read MN-major (the operand TRANSPOSED): the 128-byte rows run along M/N (64 halves),
// each of the 8 rows of an atom is one K index.  Canonical layout (cute/atom/mma_traits_sm100.hpp,
// make_umma_desc<Major::MN>, SWIZZLE_128B, in 16-byte units): ((8,n),(8,k)):((1,LBO),(8,SBO)) --
// LBO = byte distance between consecutive 64-element M/N blocks, SBO = between 8-row K groups.
__device__ __forceinline__ uint64_t make_smem_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes,
                                                            uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
         ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t idesc_a_mn_major(uint32_t idesc) { return idesc | (1u << 15); }

}  // namespace ptx

// ---------------------------------------------------------------- packed operand layout
// A row-major [rows, K] matrix is re-laid out as fp16 "TilePack": 128-row blocks, each
// holding K/64 slices of (128 rows x 64 halves) = 16 swizzle atoms = 16 KiB, byte-identical
// to the shared-memory image tcgen05.mma reads (SWIZZLE_128B, K-major).  One cp.async.bulk
// of 16 KiB therefore moves a ready-to-use operand slice; no tensor map is needed.
//   byte offset of element (r, k):
//     (((r/128) * nkb + k/64) * 16 + (r%128)/8) * 1024 + (r%8) * 128
//       + ((((k%64)/8) ^ (r%8)) * 16) + (k%8) * 2
constexpr int TP_ROWS = 128;      // rows per block
constexpr int TP_K = 64;          // halves per K slice (128 bytes)
constexpr int TP_SLICE_BYTES = TP_ROWS * TP_K * 2;   // 16384

__host__ __device__ inline size_t tilepack_bytes(int64_t rows_padded, int k_padded) {
  return (size_t)rows_padded * (size_t)k_padded * 2;
}
__host__ __device__ inline size_t tilepack_chunk_offset(int64_t r, int chunk /* k/8 */, int nkb) {
  const int64_t rb = r >> 7;
  const int kb = chunk >> 3;
  const int c = chunk & 7;
  const int rr = (int)(r & 127);
  return ((size_t)((rb * nkb + kb) * 16 + (rr >> 3)) << 10) + (size_t)((rr & 7) << 7) +
         (size_t)(((c ^ (rr & 7)) << 4));
}

}  // namespace mclst
