// Thin inline-PTX wrappers for the sm_100a features the GEMM-class kernels use:
// mbarrier, tcgen05 (alloc / mma / commit / ld / fences), bulk async copies (TMA engine,
// no tensor map) and the shared-memory matrix descriptors of the 5th-gen tensor cores.
//
// Everything here is written for sm_100a only (no other arch, no fallbacks).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace fmm {

// ----------------------------------------------------------------------------------------
// Watchdog: every mbarrier wait is bounded. A protocol bug then becomes a trapped kernel
// (CUDA error on the host) instead of a hung GPU box.
// ----------------------------------------------------------------------------------------
#ifndef FMM_WAIT_LIMIT_CYCLES
#define FMM_WAIT_LIMIT_CYCLES (4000000000ll)  // ~2 s at 1.9 GHz
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() {
  uint32_t l;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
  return l;
}

// ----------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(bar)
               : "memory");
}
// try_wait with an explicit suspend-time hint: the thread sleeps IN HARDWARE until the phase completes (woken by the
// arrive, ~60 cycles) or the hint expires, instead of returning after the short default window. Without the hint the
// wait loops of the idle roles (epilogue, loaders, producers ahead of the tensor core) were a third of all instructions the
// graph-conv kernel issued (ncu r02: 1.0 M iterations of one wait site per launch), taken from the CUDA-core producers.
#ifndef FMM_TRYWAIT_HINT_NS
#define FMM_TRYWAIT_HINT_NS 100000u
#endif
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(FMM_TRYWAIT_HINT_NS)
      : "memory");
  return ok;
}
// Optional wait-time profile (dev aid): cycles spent blocked per wait site `tag`, summed over threads.
extern __device__ int g_wait_prof_enable;
extern __device__ unsigned long long g_wait_prof[32];

// Bounded wait. `tag` identifies the wait site in the error word.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, unsigned* err, unsigned tag) {
  // note: try_wait itself suspends the thread for a bounded time, so the clock starts before it
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > FMM_WAIT_LIMIT_CYCLES) {
      if (err) atomicCAS(err, 0u, 0x80000000u | (tag << 16) | (blockIdx.x & 0xffffu));
      __threadfence_system();
      asm volatile("trap;");
    }
  }
  if (g_wait_prof_enable && (threadIdx.x & 31) == 0) atomicAdd(&g_wait_prof[tag & 31], static_cast<unsigned long long>(clock64() - t0));
}

// Wait for roles that are idle most of the time (epilogue, loaders): a failed try is followed by a short sleep, and the
// watchdog clock is read only every 64 tries. A tight spin costs ~9 issue slots per ~20 cycles PER WAITING WARP - measured
// 26 % of all instructions of the graph-conv kernel - which the CUDA-core producer warps of the same SM need.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity, unsigned* err, unsigned tag, unsigned sleep_ns = 128) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  unsigned it = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(sleep_ns);
    if ((++it & 63u) == 0 && clock64() - t0 > FMM_WAIT_LIMIT_CYCLES) {
      if (err) atomicCAS(err, 0u, 0x80000000u | (tag << 16) | (blockIdx.x & 0xffffu));
      __threadfence_system();
      asm volatile("trap;");
    }
  }
  if (g_wait_prof_enable && (threadIdx.x & 31) == 0) atomicAdd(&g_wait_prof[tag & 31], static_cast<unsigned long long>(clock64() - t0));
}

// ----------------------------------------------------------------------------------------
// Thread-block clusters / CTA pairs: barriers of the peer CTA are reached through shared::cluster addresses.
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {  // every thread of every CTA of the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(FMM_TRYWAIT_HINT_NS)
      : "memory");
  return ok;
}
// Bounded wait on a barrier that threads of the peer CTA arrive on (acquire at cluster scope).
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, unsigned* err, unsigned tag) {
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > FMM_WAIT_LIMIT_CYCLES) {
      if (err) atomicCAS(err, 0u, 0x80000000u | (tag << 16) | (blockIdx.x & 0xffffu));
      __threadfence_system();
      asm volatile("trap;");
    }
  }
  if (g_wait_prof_enable && (threadIdx.x & 31) == 0) atomicAdd(&g_wait_prof[tag & 31], static_cast<unsigned long long>(clock64() - t0));
}

// generic-proxy smem writes -> visible to the async proxy (tensor core / TMA reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ----------------------------------------------------------------------------------------
// Bulk async copy global -> shared (TMA engine, 1-D, no tensor map). Completion is
// signalled as transaction bytes on an mbarrier. Size and both addresses: multiples of 16.
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes,
                                         uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

// ----------------------------------------------------------------------------------------
// tcgen05: tensor memory + MMA
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {  // whole warp
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// CTA-pair (cta_group::2) variants: one warp of EACH CTA of the pair allocates / frees
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {  // whole warp
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same MMA with the descriptors passed as (lo, hi) 32-bit halves: the hi half (SBO, version, swizzle
// mode) is loop invariant and the lo half (start address >> 4 | LBO << 16) advances by plain 32-bit
// adds, which keeps the single issuing thread's instruction stream short.
__device__ __forceinline__ void umma_bf16_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                             uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// CTA-pair MMA: M = 256 rows split over the two CTAs (each reads its own A rows and its own half of the B rows from its
// own shared memory at the descriptors' offsets; each accumulates its 128 rows x N columns in its own TMEM). Issued by ONE
// thread of the leader CTA (cluster rank 0).
__device__ __forceinline__ void umma2_bf16_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                              uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr & 0x3ffffu) >> 4) | (((lbo_bytes >> 4) & 0x3fffu) << 16);
}
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes) {
  return ((sbo_bytes >> 4) & 0x3fffu) | (1u << 14) | (2u << 29);  // version 1, SWIZZLE_128B
}
// One lane of a converged warp (the idiom nvcc recognises: the guarded region needs no divergence handling).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "@px mov.s32 %0, 1;\n\t}"
      : "+r"(pred));
  return pred != 0;
}

// mbarrier arrive once all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
// CTA-pair commit: arrives on the barrier at this offset in BOTH CTAs of the pair once the leader's MMAs have completed.
__device__ __forceinline__ void umma2_commit(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask)
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i = lane base+i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ----------------------------------------------------------------------------------------
// Descriptors.
//
// Shared-memory matrix descriptor (64 bit), sm_100 format:
//   [0,14)  start address >> 4        [16,30) leading-dim byte offset >> 4 (LBO)
//   [32,46) stride-dim byte offset >> 4 (SBO)   [46,48) version = 1
//   [49,52) base offset = 0           [61,64) layout: 0 none, 2 = SWIZZLE_128B
//
// Canonical SWIZZLE_128B images used everywhere in this library (bf16, 16-byte units):
//   a "row" is 128 bytes = 64 bf16; 8 rows form a 1024-byte atom; inside an atom the
//   16-byte chunk c of row r sits at chunk position c ^ (r & 7). Atom bases are 1024-aligned.
//   * K-major operand  (rows = M or N index, 64 K-elements per row):
//       SBO = byte distance between consecutive 8-row groups; LBO unused.
//       A K=16 slice inside the 64-wide row is selected by adding 32 bytes to the start.
//   * MN-major operand (rows = K index, 64 M/N-elements per row):
//       SBO = byte distance between consecutive 8-row (K) groups,
//       LBO = byte distance between consecutive 64-element M/N atoms.
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3ffffu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;  // SWIZZLE_128B
  return d;
}

// Instruction descriptor for kind::f16 with bf16 A/B and fp32 D, M = 128.
//   [4,6) D fmt (1 = f32)  [7,10) A fmt (1 = bf16)  [10,13) B fmt  [15] A major  [16] B major
//   (0 = K-major, 1 = MN-major)  [17,23) N >> 3  [24,29) M >> 4
__host__ __device__ __forceinline__ uint32_t make_idesc_bf16(int n, int a_mn_major, int b_mn_major, int m = 128) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 1u << 7;
  d |= 1u << 10;
  d |= static_cast<uint32_t>(a_mn_major & 1) << 15;
  d |= static_cast<uint32_t>(b_mn_major & 1) << 16;
  d |= static_cast<uint32_t>(n >> 3) << 17;
  d |= static_cast<uint32_t>(m >> 4) << 24;  // 256: CTA pair (cta_group::2)
  return d;
}

// byte offset of 16-byte chunk `c` (0..7) of row `r` inside a SWIZZLE_128B image whose rows are
// 128 bytes and whose 8-row atoms are contiguous (row r at r*128).
__host__ __device__ __forceinline__ uint32_t sw128_off(uint32_t r, uint32_t c) {
  return r * 128u + ((c ^ (r & 7u)) << 4);
}

}  // namespace fmm
