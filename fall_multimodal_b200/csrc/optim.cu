// Multi-tensor RMSprop step (the reference's optimizer: F2/optimizer.py:20-21 torch.optim.RMSprop(lr=1e-3), defaults
// alpha=0.99, eps=1e-8, momentum=0, centered=False; MF3/main.py:103-113 adds GradScaler unscaling and clip_grad_norm_).
//
// One launch updates EVERY parameter tensor of the model (torch's foreach path is ~25 launches per step for the 190-tensor
// trunks); a second, optional launch before it reduces the global gradient norm / non-finite flag that clipping and loss
// scaling need. All scalars the step depends on (lr, the clip coefficient, the inverse loss scale, the skip flag) are read
// from device memory, so the step can live inside a CUDA graph while a scheduler changes lr between replays.
//
//   g   = grad * inv_scale * clip_coef        (clip_coef = min(1, max_norm / (||grad*inv_scale|| + 1e-6)); skip if non-finite)
//   g  += weight_decay * p
//   sq  = alpha * sq + (1 - alpha) * g^2
//   p  -= lr * g / (sqrt(sq) + eps)
#include "common.cuh"

namespace fmm {

struct OptTensor {
  float* p;
  const float* g;
  float* sq;
  long long n;
};

constexpr int kOptChunk = 4096;   // elements per block

// chunk -> (tensor, offset) table: chunk_tensor[c], chunk_off[c]
__global__ void rmsprop_kernel(const OptTensor* __restrict__ tensors, const int* __restrict__ chunk_tensor,
                               const long long* __restrict__ chunk_off, const float* __restrict__ lr, float alpha, float eps,
                               float weight_decay, const float* __restrict__ norm_sq, float max_norm,
                               const float* __restrict__ inv_scale) {
  const OptTensor t = tensors[chunk_tensor[blockIdx.x]];
  const long long off = chunk_off[blockIdx.x];
  float coef = inv_scale ? inv_scale[0] : 1.f;
  if (norm_sq) {
    const float nsq = norm_sq[0];
    if (!(nsq == nsq) || nsq > 3.0e38f) return;   // non-finite gradients: skip the step (GradScaler semantics)
    if (max_norm > 0.f) {
      const float nrm = sqrtf(nsq) * coef;
      coef *= fminf(1.f, max_norm / (nrm + 1e-6f));
    }
  }
  const float step = lr[0];
  const long long end = off + kOptChunk < t.n ? off + kOptChunk : t.n;
  for (long long i = off + threadIdx.x; i < end; i += blockDim.x) {
    float p = t.p[i];
    float g = t.g[i] * coef;
    if (weight_decay != 0.f) g = fmaf(weight_decay, p, g);
    const float sq = fmaf(alpha, t.sq[i], (1.f - alpha) * g * g);
    t.sq[i] = sq;
    t.p[i] = p - step * g / (sqrtf(sq) + eps);
  }
}

// sum of squares of all gradients (fp32 atomics of per-block partial sums; non-finite values propagate into the sum)
__global__ void gradnorm_kernel(const OptTensor* __restrict__ tensors, const int* __restrict__ chunk_tensor,
                                const long long* __restrict__ chunk_off, float* __restrict__ norm_sq) {
  __shared__ float red[8];
  const OptTensor t = tensors[chunk_tensor[blockIdx.x]];
  const long long off = chunk_off[blockIdx.x];
  const long long end = off + kOptChunk < t.n ? off + kOptChunk : t.n;
  float s = 0.f;
  for (long long i = off + threadIdx.x; i < end; i += blockDim.x) {
    const float g = t.g[i];
    s = fmaf(g, g, s);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(norm_sq, v);
  }
}

}  // namespace fmm

using namespace fmm;

extern "C" {

int fmm_opt_chunk(void) { return kOptChunk; }

/* tensors: device array of {param, grad, state, numel} (fp32); chunk_tensor / chunk_off: device chunk table (kOptChunk elements
 * per chunk). lr, norm_sq (nullable), inv_scale (nullable): device scalars. */
int fmm_rmsprop_step(const void* tensors, const int* chunk_tensor, const long long* chunk_off, int nchunks, const float* lr,
                     float alpha, float eps, float weight_decay, const float* norm_sq, float max_norm, const float* inv_scale,
                     cudaStream_t stream) {
  FMM_CHECK_ARG(tensors && chunk_tensor && chunk_off && lr && nchunks > 0, "rmsprop_step: bad arguments");
  rmsprop_kernel<<<nchunks, 256, 0, stream>>>(reinterpret_cast<const OptTensor*>(tensors), chunk_tensor, chunk_off, lr, alpha, eps,
                                               weight_decay, norm_sq, max_norm, inv_scale);
  FMM_CHECK_LAUNCH("rmsprop_step");
  return FMM_OK;
}

int fmm_grad_norm_sq(const void* tensors, const int* chunk_tensor, const long long* chunk_off, int nchunks, float* norm_sq,
                     cudaStream_t stream) {
  FMM_CHECK_ARG(tensors && chunk_tensor && chunk_off && norm_sq && nchunks > 0, "grad_norm_sq: bad arguments");
  gradnorm_kernel<<<nchunks, 256, 0, stream>>>(reinterpret_cast<const OptTensor*>(tensors), chunk_tensor, chunk_off, norm_sq);
  FMM_CHECK_LAUNCH("grad_norm_sq");
  return FMM_OK;
}

}  // extern "C"
