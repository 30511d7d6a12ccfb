// TMA-staged row streaming for the HBM-bound kernels of the block glue (csrc/elementwise.cu).
//
// A register-only streaming kernel keeps (threads per SM) x (loads in flight per thread) x 16 bytes on the wire: with the
// 96-155 registers these kernels need that is ~64 KB per SM, and at the ~2.5 us loaded HBM latency of a B200 it caps the
// READ rate at ~3.8 TB/s (measured: every reduce / apply kernel, whatever its arithmetic). Here one producer lane issues bulk
// async copies (`cp.async.bulk`, the TMA engine) of whole row chunks of every input tensor into a 4-stage shared-memory ring
// (128 KB in flight per SM, no registers); 8 consumer warps read the chunks with conflict-free 16-byte shared loads.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace fmm {

constexpr int kStConsumers = 256;             // 8 consumer warps
constexpr int kStThreads = kStConsumers + 32;  // + the producer warp
constexpr int kStStages = 4;

// A block's balanced share of the flattened (clip, frame * joint) row space, cut into chunks of <= chunk_rows rows.
// unit_rows rows are never split (1, or V for kernels whose threads own (joint, channel) pairs); with per_clip a chunk never
// straddles two clips (per-clip coefficients / sums). Producer and consumers walk the same sequence.
struct ChunkIter {
  long long cur, end;
  int rows_per_n, chunk_rows;
  bool per_clip;
  __device__ ChunkIter(long long total_rows, int unit_rows, int rows_per_n_, int chunk_rows_, bool per_clip_)
      : rows_per_n(rows_per_n_), chunk_rows(chunk_rows_), per_clip(per_clip_) {
    const long long units = total_rows / unit_rows;
    cur = units * blockIdx.x / gridDim.x * unit_rows;
    end = units * (blockIdx.x + 1) / gridDim.x * unit_rows;
  }
  __device__ bool next(long long& row0, int& nrows, int& n) {
    if (cur >= end) return false;
    long long lim = end;
    n = static_cast<int>(cur / rows_per_n);
    if (per_clip) {
      const long long e = static_cast<long long>(n + 1) * rows_per_n;
      lim = e < lim ? e : lim;
    }
    const long long left = lim - cur;
    nrows = static_cast<int>(left < chunk_rows ? left : chunk_rows);
    row0 = cur;
    cur += nrows;
    return true;
  }
};

template <int NT>
struct StreamPipe {
  uint32_t data0, full0, empty0, chunk_bytes;
  // shared memory: [stage][tensor][chunk_bytes] | full[kStStages] | empty[kStStages]
  __device__ void init(uint8_t* smem_raw, uint32_t chunk_bytes_) {
    data0 = (smem_u32(smem_raw) + 127u) & ~127u;
    chunk_bytes = chunk_bytes_;
    full0 = data0 + kStStages * NT * chunk_bytes;
    empty0 = full0 + 8u * kStStages;
    if (threadIdx.x == 0) {
      for (int s = 0; s < kStStages; ++s) {
        mbar_init(full0 + 8u * s, 1);
        mbar_init(empty0 + 8u * s, kStConsumers / 32);
      }
      mbar_fence_init();
    }
    __syncthreads();
  }
  static size_t bytes(uint32_t chunk_bytes_) { return 128 + static_cast<size_t>(kStStages) * NT * chunk_bytes_ + 16 * kStStages; }
  __device__ uint32_t tensor(int s, int k) const { return data0 + (static_cast<uint32_t>(s) * NT + k) * chunk_bytes; }

  // producer side (ONE lane): queue chunk i = rows [row0, row0 + nrows) of every tensor (row_bytes each)
  __device__ void produce(int i, const void* const (&src)[NT], long long row0, int nrows, uint32_t row_bytes, unsigned* err) {
    const int s = i % kStStages;
    const uint32_t ph = static_cast<uint32_t>(i / kStStages) & 1u;
    mbar_wait_relaxed(empty0 + 8u * s, ph ^ 1u, err, 20, 32);
    const uint32_t nb = static_cast<uint32_t>(nrows) * row_bytes;
    uint32_t live = 0;   // optional tensors (null) are skipped
#pragma unroll
    for (int k = 0; k < NT; ++k) live += src[k] ? 1u : 0u;
    mbar_arrive_expect_tx(full0 + 8u * s, live * nb);
#pragma unroll
    for (int k = 0; k < NT; ++k)
      if (src[k])
        bulk_g2s(tensor(s, k), reinterpret_cast<const uint8_t*>(src[k]) + static_cast<size_t>(row0) * row_bytes, nb, full0 + 8u * s);
  }
  // consumer side (all consumer threads): wait for chunk i, returns its stage
  __device__ int acquire(int i, unsigned* err) const {
    const int s = i % kStStages;
    mbar_wait(full0 + 8u * s, static_cast<uint32_t>(i / kStStages) & 1u, err, 21);
    return s;
  }
  __device__ void release(int s) const {  // every consumer warp, converged
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(empty0 + 8u * s);
  }
};

// 8 activations at a shared-memory address (16 bytes bf16 / 32 bytes fp32) as fp32
__device__ __forceinline__ void lds8(uint32_t addr, float (&f)[8], const __nv_bfloat16*) {
  uint32_t w[4];
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(addr));
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void lds8(uint32_t addr, float (&f)[8], const float*) {
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]) : "r"(addr));
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(f[4]), "=f"(f[5]), "=f"(f[6]), "=f"(f[7]) : "r"(addr + 16u));
}

// rows per chunk for a row of `row_bytes` bytes: as many whole units as fit `budget` bytes (at least one unit)
static inline int stream_chunk_rows(size_t budget, size_t row_bytes, int unit_rows) {
  long long r = static_cast<long long>(budget / row_bytes) / unit_rows * unit_rows;
  return static_cast<int>(r < unit_rows ? unit_rows : r);
}

}  // namespace fmm
