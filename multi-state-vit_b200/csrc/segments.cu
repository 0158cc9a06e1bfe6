// Hierarchy bookkeeping around the per-segment kernels.
//
// Reference: model/clustering/modeling_spectral.py:80-84 (tokens of parent p are gathered with a boolean
// mask), :91-94 (child id = children of earlier parents + local id) and the consumer
// model/multistate_encoder/modeling_msvitencoder.py:491-499 (children of a parent form a contiguous id range,
// ranges ordered by parent).  The reference pools each parent across the whole batch; the per-image variant
// (:260-279) is the one built here: a segment is (image, parent).
#include <cuda_bf16.h>

#include "common.cuh"

namespace msvit {
namespace seg {

constexpr int kThreads = 256;

// one CTA per image: stable counting sort of the tokens by parent id
__global__ void __launch_bounds__(kThreads) build_segments_kernel(const int64_t* __restrict__ parent,
                                                                   int32_t* __restrict__ perm,
                                                                   int32_t* __restrict__ seg_off,
                                                                   int64_t* __restrict__ a_off, int B, int N, int P,
                                                                   long long a_stride) {
  extern __shared__ int sm[];
  int* par = sm;          // [N]
  int* start = par + N;   // [P + 1]
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < N; i += kThreads) {
    long long p = parent[static_cast<long long>(b) * N + i];
    p = p < 0 ? 0 : (p >= P ? P - 1 : p);
    par[i] = static_cast<int>(p);
  }
  for (int c = threadIdx.x; c <= P; c += kThreads) start[c] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += kThreads) atomicAdd(&start[par[i] + 1], 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    long long a = static_cast<long long>(b) * a_stride;
    int run = 0;
    for (int p = 0; p < P; ++p) {
      const int n = start[p + 1];
      seg_off[b * P + p] = b * N + run;
      a_off[b * P + p] = a;
      a += static_cast<long long>(n) * lda_of(n);
      start[p + 1] = run + n;
      run += n;
    }
    if (b == B - 1) {
      seg_off[B * P] = B * N;
      a_off[B * P] = static_cast<long long>(B) * a_stride;
    }
  }
  __syncthreads();
  // stable placement by one warp, 32 tokens a step: `match.any` ranks a lane among the lanes with the same parent,
  // start[p] doubles as the running cursor of parent p
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    const unsigned below = (1u << lane) - 1u;
    for (int base = 0; base < N; base += 32) {
      const int i = base + lane;
      const int p = i < N ? par[i] : -1;
      const unsigned peers = __match_any_sync(0xffffffffu, p);
      const int at = p >= 0 ? start[p] : 0;
      __syncwarp();
      if (p >= 0) {
        perm[b * N + at + __popc(peers & below)] = b * N + i;
        if ((peers & below) == 0) start[p] = at + __popc(peers);
      }
      __syncwarp();
    }
  }
}

// xs[j, :] = x[perm[j], :]   (16-byte chunks when the row size allows)
__global__ void gather_rows_kernel(const uint8_t* __restrict__ x, const int32_t* __restrict__ perm,
                                   uint8_t* __restrict__ xs, long long rows, int row_bytes, int vec) {
  const long long chunks = row_bytes / vec;
  const long long total = rows * chunks;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long j = e / chunks, c = e - j * chunks;
    const uint8_t* src = x + static_cast<long long>(perm[j]) * row_bytes + c * vec;
    uint8_t* dst = xs + j * row_bytes + c * vec;
    if (vec == 16) *reinterpret_cast<uint4*>(dst) = __ldg(reinterpret_cast<const uint4*>(src));
    else if (vec == 4) *reinterpret_cast<uint32_t*>(dst) = __ldg(reinterpret_cast<const uint32_t*>(src));
    else *reinterpret_cast<uint16_t*>(dst) = __ldg(reinterpret_cast<const uint16_t*>(src));
  }
}

// one CTA per image: child id = children of earlier parents + local id, scattered back through perm
__global__ void __launch_bounds__(kThreads) compose_labels_kernel(const int32_t* __restrict__ labels_sorted,
                                                                   const int32_t* __restrict__ n_child,
                                                                   const int32_t* __restrict__ perm,
                                                                   const int32_t* __restrict__ seg_off,
                                                                   int64_t* __restrict__ child, int N, int P) {
  extern __shared__ int sm[];
  int* coff = sm;  // [P] exclusive scan of the children counts of this image
  const int b = blockIdx.x;
  if (threadIdx.x == 0) {
    int run = 0;
    for (int p = 0; p < P; ++p) {
      coff[p] = run;
      run += n_child[b * P + p];
    }
  }
  __syncthreads();
  for (int p = 0; p < P; ++p) {
    const int j0 = seg_off ? seg_off[b * P + p] : b * N;
    const int j1 = seg_off ? seg_off[b * P + p + 1] : (b + 1) * N;
    for (int j = j0 + threadIdx.x; j < j1; j += kThreads) {
      const int dst = perm ? perm[j] : j;
      child[dst] = static_cast<int64_t>(coff[p] + labels_sorted[j]);
    }
  }
}

}  // namespace seg
}  // namespace msvit

extern "C" int msvit_build_segments(const int64_t* parent_indices, int32_t* perm, int32_t* seg_off, int64_t* a_off,
                                    int B, int N, int P, msvit_stream_t stream_) {
  using namespace msvit;
  if (!parent_indices || !perm || !seg_off || !a_off) return MSVIT_ERR_NULL;
  if (B < 0 || N <= 0 || P <= 0 || static_cast<int64_t>(B) * N > 0x7fffffffLL ||
      static_cast<int64_t>(B) * P >= 0x7fffffffLL)
    return MSVIT_ERR_SHAPE;
  if (B == 0) return MSVIT_OK;
  const size_t smem = sizeof(int) * (static_cast<size_t>(N) + P + 1);
  if (smem > 200 * 1024) return MSVIT_ERR_SHAPE;
  cudaError_t e = cudaFuncSetAttribute(seg::build_segments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return cuda_status(e);
  // per-image affinity capacity: sum n_p * lda(n_p) <= N * (N + 3), kept a multiple of 4
  const long long a_stride = (static_cast<long long>(N) * (N + 3) + 3) & ~3LL;
  seg::build_segments_kernel<<<B, seg::kThreads, smem, static_cast<cudaStream_t>(stream_)>>>(
      parent_indices, perm, seg_off, a_off, B, N, P, a_stride);
  return cuda_status(cudaGetLastError());
}

extern "C" int msvit_gather_rows(const void* x, int x_dtype, const int32_t* perm, void* xs, int64_t total_rows, int D,
                                 msvit_stream_t stream_) {
  using namespace msvit;
  if (!x || !perm || !xs) return MSVIT_ERR_NULL;
  if (x_dtype != MSVIT_F32 && x_dtype != MSVIT_BF16) return MSVIT_ERR_MODE;
  if (total_rows < 0 || D <= 0) return MSVIT_ERR_SHAPE;
  if (total_rows == 0) return MSVIT_OK;
  const int esz = x_dtype == MSVIT_F32 ? 4 : 2;
  const long long row_bytes = static_cast<long long>(D) * esz;
  if (row_bytes > 0x7fffffffLL) return MSVIT_ERR_SHAPE;
  const bool a16 = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(xs)) & 15) == 0 && row_bytes % 16 == 0;
  const int vec = a16 ? 16 : esz;
  const long long total = total_rows * (row_bytes / vec);
  long long blocks = (total + 255) / 256;
  const long long cap = 32LL * sm_count();
  if (blocks > cap) blocks = cap;
  seg::gather_rows_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const uint8_t*>(x), perm, static_cast<uint8_t*>(xs), total_rows, static_cast<int>(row_bytes), vec);
  return cuda_status(cudaGetLastError());
}

extern "C" int msvit_compose_labels(const int32_t* labels_sorted, const int32_t* n_child, const int32_t* perm,
                                    const int32_t* seg_off, int64_t* child, int B, int N, int P,
                                    msvit_stream_t stream_) {
  using namespace msvit;
  if (!labels_sorted || !n_child || !child) return MSVIT_ERR_NULL;
  if (B < 0 || N <= 0 || P <= 0) return MSVIT_ERR_SHAPE;
  if (!seg_off && P != 1) return MSVIT_ERR_SHAPE;
  if (B == 0) return MSVIT_OK;
  const size_t smem = sizeof(int) * static_cast<size_t>(P);
  if (smem > 48 * 1024) return MSVIT_ERR_SHAPE;
  seg::compose_labels_kernel<<<B, seg::kThreads, smem, static_cast<cudaStream_t>(stream_)>>>(
      labels_sorted, n_child, perm, seg_off, child, N, P);
  return cuda_status(cudaGetLastError());
}
