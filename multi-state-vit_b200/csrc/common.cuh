// Shared host/device helpers for the msvit kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/msvit.h"

namespace msvit {

constexpr float kLog2e = 1.4426950408889634f;

__host__ __device__ __forceinline__ int round_up(int a, int b) { return (a + b - 1) / b * b; }
__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ int lda_of(int n) { return (n + 3) & ~3; }

// Geometry of segment s (see msvit.h): first row, length, affinity offset / leading dimension.
struct Seg {
  int row0;
  int n;
  int lda;
  long long a0;
};

__device__ __forceinline__ Seg seg_info(int s, int N, const int32_t* __restrict__ seg_off,
                                        const int64_t* __restrict__ a_off) {
  Seg g;
  if (seg_off) {
    g.row0 = seg_off[s];
    g.n = seg_off[s + 1] - g.row0;
  } else {
    g.row0 = s * N;
    g.n = N;
  }
  g.lda = lda_of(g.n);
  g.a0 = a_off ? static_cast<long long>(a_off[s]) : static_cast<long long>(s) * N * lda_of(N);
  return g;
}

inline int cuda_status(cudaError_t e) { return e == cudaSuccess ? MSVIT_OK : static_cast<int>(e); }

inline int sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  return n;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Stable grouping of one image's tokens by label, shared by every CTA of the image: histogram, one-warp scan, then one
// warp walks the tokens 32 at a time -- `match.any` gives each lane its rank among the lanes of the same label and a
// per-label cursor carries the count between steps, so a group lists its tokens in ascending order (fixed summation
// order => bit-reproducible means) at N/32 steps instead of a rank loop over all earlier tokens.
__device__ __forceinline__ void group_tokens(const int64_t* __restrict__ lb, int* lab, int* order, int* start,
                                             int* cursor, int N, int K) {
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const long long l = lb[i];
    lab[i] = (l >= 0 && l < K) ? static_cast<int>(l) : -1;
  }
  for (int c = threadIdx.x; c <= K; c += blockDim.x) start[c] = 0;
  for (int c = threadIdx.x; c < K; c += blockDim.x) cursor[c] = 0;
  __syncthreads();
  // histogram (integer shared-memory atomics: order independent, hence deterministic)
  for (int i = threadIdx.x; i < N; i += blockDim.x)
    if (lab[i] >= 0) atomicAdd(&start[lab[i] + 1], 1);
  __syncthreads();
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    int carry = 0;
    for (int base = 1; base <= K; base += 32) {
      const int idx = base + lane;
      int v = idx <= K ? start[idx] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
      }
      if (idx <= K) start[idx] = v + carry;
      carry += __shfl_sync(0xffffffffu, v, 31);
    }
    __syncwarp();
    const unsigned below = (1u << lane) - 1u;
    for (int base = 0; base < N; base += 32) {
      const int i = base + lane;
      const int l = i < N ? lab[i] : -1;
      const unsigned peers = __match_any_sync(0xffffffffu, l);
      const int seen = l >= 0 ? cursor[l] : 0;
      __syncwarp();
      if (l >= 0) {
        order[start[l] + seen + __popc(peers & below)] = i;
        if ((peers & below) == 0) cursor[l] = seen + __popc(peers);
      }
      __syncwarp();
    }
  }
  __syncthreads();
}

}  // namespace msvit
