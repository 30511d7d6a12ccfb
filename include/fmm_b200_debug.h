/* fmm_b200_debug.h — measurement aids of libfmm_b200.so. NOT part of the product interface (include/fmm_b200.h): nothing on
 * the train / inference path calls them; scripts/wait_profile*.py and scripts/mma_probe.py do. */
#ifndef FMM_B200_DEBUG_H
#define FMM_B200_DEBUG_H

#include <cuda_runtime_api.h>

#ifdef __cplusplus
extern "C" {
#endif

/* per-wait-site blocked cycles of the GEMM kernels' mbarrier waits (enable >= 0: reset + set; out32: 32 counters) */
int fmm_debug_wait_profile(int enable, unsigned long long* out32);
/* cycles for `iters` back-to-back tcgen05.mma (M=128, K=16) per CTA: out2 = {issue cycles, issue + drain cycles} */
int fmm_debug_mma_probe(int N, int iters, int a_mn, int b_mn, int distinct_acc, int ctas, unsigned long long* out2_dev,
                        cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif
