// Cluster-mean pooling of ViT tokens into multi-state tokens (HBM-bound streaming reduction).
//
// Reference: per-label mean loops, model/clustering/modeling_spectral.py:125-127 and :271-273
// (`mean(features[labels == c], dim=0)` for every c: K boolean-mask passes over x).
// Here x is read exactly once: a CTA owns (image, 16-byte column slab); it counting-sorts the image's
// tokens by label in shared memory (stable, so the summation order is fixed and the result is
// bit-reproducible), then every thread streams its 128-bit column through the clusters' token lists,
// accumulating in registers -- no atomics, coalesced 128-bit loads, one store per (cluster, column).
#include <cuda_bf16.h>

#include <cstdlib>

#include "common.cuh"

namespace msvit {
namespace pool {

template <typename T, int VEC>
struct Loader;

template <>
struct Loader<float, 4> {
  static __device__ __forceinline__ void add(const float* p, float (&acc)[4]) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(p));
    acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
  }
};
template <>
struct Loader<float, 1> {
  static __device__ __forceinline__ void add(const float* p, float (&acc)[1]) { acc[0] += __ldcs(p); }
};
template <>
struct Loader<__nv_bfloat16, 8> {
  static __device__ __forceinline__ void add(const __nv_bfloat16* p, float (&acc)[8]) {
    const uint4 v = __ldcs(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc[2 * i] += __uint_as_float(w[i] << 16);
      acc[2 * i + 1] += __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
};
template <>
struct Loader<__nv_bfloat16, 1> {
  static __device__ __forceinline__ void add(const __nv_bfloat16* p, float (&acc)[1]) {
    acc[0] += __bfloat162float(*p);
  }
};

// grid (image, column slab, cluster range): CTA z owns clusters [z K / Z, (z+1) K / Z) so that a small batch of long
// images still fills the machine; every CTA of an image repeats the (cheap) grouping.
template <typename T, int VEC>
__global__ void pool_kernel(const T* __restrict__ x, const int64_t* __restrict__ labels, float* __restrict__ pooled,
                            int32_t* __restrict__ counts, int N, int D, int K) {
  extern __shared__ int sm[];
  int* lab = sm;            // [N]   label or -1
  int* order = lab + N;     // [N]   tokens grouped by label, ascending token id inside a group
  int* start = order + N;   // [K+1] group offsets
  int* cursor = start + K + 1;  // [K]
  const int b = blockIdx.x;
  group_tokens(labels + static_cast<long long>(b) * N, lab, order, start, cursor, N, K);
  if (blockIdx.y == 0 && blockIdx.z == 0)
    for (int c = threadIdx.x; c < K; c += blockDim.x)
      counts[static_cast<long long>(b) * K + c] = start[c + 1] - start[c];

  const int col = (blockIdx.y * blockDim.x + threadIdx.x) * VEC;
  if (col >= D) return;
  const int c0 = static_cast<int>(static_cast<long long>(blockIdx.z) * K / gridDim.z);
  const int c1 = static_cast<int>(static_cast<long long>(blockIdx.z + 1) * K / gridDim.z);
  const T* xb = x + static_cast<long long>(b) * N * D + col;
  float* pb = pooled + static_cast<long long>(b) * K * D + col;
  for (int c = c0; c < c1; ++c) {
    const int s0 = start[c], s1 = start[c + 1];
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
    int i = s0;
    for (; i + 8 <= s1; i += 8) {
      int t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) t[u] = order[i + u];
#pragma unroll
      for (int u = 0; u < 8; ++u) Loader<T, VEC>::add(xb + static_cast<long long>(t[u]) * D, acc);
    }
    for (; i < s1; ++i) Loader<T, VEC>::add(xb + static_cast<long long>(order[i]) * D, acc);
    const float inv = s1 > s0 ? 1.0f / static_cast<float>(s1 - s0) : 0.f;
    if constexpr (VEC >= 4) {
#pragma unroll
      for (int v = 0; v < VEC; v += 4)
        *reinterpret_cast<float4*>(pb + static_cast<long long>(c) * D + v) =
            make_float4(acc[v] * inv, acc[v + 1] * inv, acc[v + 2] * inv, acc[v + 3] * inv);
    } else {
      pb[static_cast<long long>(c) * D] = acc[0] * inv;
    }
  }
}

template <typename T, int VEC>
static int launch(const void* x, const int64_t* labels, float* pooled, int32_t* counts, int B, int N, int D, int K,
                  cudaStream_t stream) {
  const int cols = ceil_div(D, VEC);
  int threads = round_up(cols, 32);
  if (threads > 256) threads = 256;
  const int ysplit = ceil_div(cols, threads);
  const size_t smem = sizeof(int) * (2 * static_cast<size_t>(N) + 2 * static_cast<size_t>(K) + 1);
  if (smem > 200 * 1024) return MSVIT_ERR_SHAPE;
  // fewer than 1024 threads per SM: split the clusters over more CTAs, at most 8 ranges (each CTA repeats the grouping)
  const long long total_threads = static_cast<long long>(B) * ysplit * threads;
  const long long want_threads = 1024LL * sm_count();
  long long zsplit = (want_threads + total_threads - 1) / total_threads;
  if (zsplit > 8) zsplit = 8;
  if (zsplit > K) zsplit = K;
  if (zsplit < 1) zsplit = 1;
  cudaError_t e = cudaFuncSetAttribute(pool_kernel<T, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return cuda_status(e);
  pool_kernel<T, VEC><<<dim3(B, ysplit, static_cast<unsigned>(zsplit)), threads, smem, stream>>>(
      static_cast<const T*>(x), labels, pooled, counts, N, D, K);
  return cuda_status(cudaGetLastError());
}

}  // namespace pool
}  // namespace msvit

extern "C" int msvit_pool(const void* x, int x_dtype, const int64_t* labels, float* pooled, int32_t* counts, int B,
                          int N, int D, int K, msvit_stream_t stream_) {
  using namespace msvit;
  if (!x || !labels || !pooled || !counts) return MSVIT_ERR_NULL;
  if (x_dtype != MSVIT_F32 && x_dtype != MSVIT_BF16) return MSVIT_ERR_MODE;
  if (B < 0 || N <= 0 || D <= 0 || K <= 0 || B > 65535 * 1024) return MSVIT_ERR_SHAPE;
  if (B == 0) return MSVIT_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool a16 = (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(pooled) & 15) == 0;
  if (x_dtype == MSVIT_F32) {
    if (a16 && D % 4 == 0) return pool::launch<float, 4>(x, labels, pooled, counts, B, N, D, K, stream);
    return pool::launch<float, 1>(x, labels, pooled, counts, B, N, D, K, stream);
  }
  if (a16 && D % 8 == 0) return pool::launch<__nv_bfloat16, 8>(x, labels, pooled, counts, B, N, D, K, stream);
  return pool::launch<__nv_bfloat16, 1>(x, labels, pooled, counts, B, N, D, K, stream);
}
