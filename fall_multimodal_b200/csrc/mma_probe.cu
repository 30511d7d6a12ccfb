// Dev probe: issue rate of back-to-back tcgen05.mma (M=128, K=16, bf16) from one thread, operands in
// shared memory (contents irrelevant). Reports cycles per MMA per CTA; used to calibrate the GEMM
// engines against the tensor-pipe floor (128*N/256 cycles per MMA).
#include "common.cuh"
#include "ptx.cuh"

namespace fmm {

__global__ void __launch_bounds__(128, 1) mma_probe_kernel(int N, int iters, int a_mn, int b_mn, int distinct_acc,
                                                           unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar, bar2, bar3;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_init(smem_u32(&bar2), 1);
    mbar_init(smem_u32(&bar3), 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_slot), 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const bool converged = (distinct_acc & 16) != 0;
  const bool leader = (threadIdx.x & 31) == 0;
  if ((distinct_acc & 32) && warp == 1) {
    const uint32_t idesc = make_idesc_bf16(N, a_mn, b_mn);
    const uint32_t a_lo0 = a_mn ? desc_lo(base, 8192) : desc_lo(base, 16);
    const uint32_t b_lo0 = b_mn ? desc_lo(base + 16384, 8192) : desc_lo(base + 16384, 16), hi = desc_hi(1024);
    const uint32_t a_st = a_mn ? (2048u >> 4) : 2u, b_st = b_mn ? (2048u >> 4) : 2u;
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 4) {
      if (distinct_acc & 8) (void)mbar_try_wait(smem_u32(&bar2), 1);
      if (distinct_acc & 2) tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (uint32_t kk = 0; kk < 4; ++kk) umma_bf16_lh(tmem_base, a_lo0 + kk * a_st, hi, b_lo0 + kk * b_st, hi, idesc, (i | kk) ? 1u : 0u);
        if (distinct_acc & 4) umma_commit(smem_u32(&bar3));
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (elect_one()) umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0, nullptr, 31);
    const long long t2 = clock64();
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) {
      out[0] = static_cast<unsigned long long>(t1 - t0);
      out[1] = static_cast<unsigned long long>(t2 - t0);
    }
  } else if ((distinct_acc & 32) == 0 && (converged ? (warp == 1) : (threadIdx.x == 32))) {
    const uint32_t idesc = make_idesc_bf16(N, a_mn, b_mn);
    const uint32_t a0 = base, b0 = base + 16384;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t kk = i & 3;
      if (kk == 0) {
        // distinct_acc bits: 2 = fence::after_thread_sync per group, 4 = commit per group (to a
        // second barrier nobody waits on), 8 = try_wait on an already completed barrier per group
        if (distinct_acc & 8) (void)mbar_try_wait(smem_u32(&bar2), 1);
        if (distinct_acc & 2) tc_fence_after();
      }
      const uint32_t d = tmem_base + ((distinct_acc & 1) ? static_cast<uint32_t>((i & 1) * 256) : 0u);
      const uint64_t ad = a_mn ? make_smem_desc(a0 + kk * 2048u, 8192, 1024) : make_smem_desc(a0 + kk * 32u, 16, 1024);
      const uint64_t bd = b_mn ? make_smem_desc(b0 + kk * 2048u, 8192, 1024) : make_smem_desc(b0 + kk * 32u, 16, 1024);
      if (leader) umma_bf16(d, ad, bd, idesc, i > 1 ? 1u : 0u);
      if (kk == 3 && (distinct_acc & 4) && leader) umma_commit(smem_u32(&bar3));
    }
    const long long t1 = clock64();
    if (leader) umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0, nullptr, 31);
    const long long t2 = clock64();
    if (blockIdx.x == 0 && leader) {
      out[0] = static_cast<unsigned long long>(t1 - t0);
      out[1] = static_cast<unsigned long long>(t2 - t0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace fmm

extern "C" int fmm_debug_mma_probe(int N, int iters, int a_mn, int b_mn, int distinct_acc, int ctas,
                                   unsigned long long* out2_dev, cudaStream_t stream) {
  cudaFuncSetAttribute(fmm::mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  fmm::mma_probe_kernel<<<ctas, 128, 64 * 1024, stream>>>(N, iters, a_mn, b_mn, distinct_acc, out2_dev);
  FMM_CHECK_LAUNCH("mma_probe");
  return FMM_OK;
}
