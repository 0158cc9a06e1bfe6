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

}  // namespace msvit
