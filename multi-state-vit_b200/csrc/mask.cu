// Cluster-restricted attention mask of the multi-state encoder.
//
// Reference: MultiStateViTEncoderBackbone._construct_attention_mask(_indices)
// (model/multistate_encoder/modeling_msvitencoder.py:426-467).  The sequence the encoder attends over is
//     [T_0, R_0, T_1, R_1, ..., T_{C-1}, R_{C-1}, token_0 .. token_{N-1}],   C = max clusters over the batch,
// and mask[b, q, k] is True for: two tokens of the same cluster; transmitter T_c -> tokens of cluster c; token of
// cluster c -> receiver R_c; receiver R_r -> transmitter T_t for r, t < (clusters of image b).
// The reference builds it with three N x N / C x N equality tensors, torch.where and index scatters; here every
// output byte is a closed-form predicate of (q, k) and the labels, written once with 32-bit stores: HBM-write bound,
// B * L * L bytes, L = 2C + N.
#include "common.cuh"

namespace msvit {
namespace mask {

__device__ __forceinline__ uint32_t predicate(int q, int k, int C2, int nb, const int* __restrict__ lab) {
  const int qt = q - C2, kt = k - C2;
  bool v;
  if (qt >= 0) {
    if (kt >= 0) v = lab[qt] == lab[kt];            // token -> token of the same cluster
    else v = (k & 1) && lab[qt] == (k >> 1);        // token -> its receiver
  } else {
    if (kt >= 0) v = !(q & 1) && lab[kt] == (q >> 1);                     // transmitter -> its tokens
    else v = (q & 1) && !(k & 1) && (q >> 1) < nb && (k >> 1) < nb;        // receiver -> transmitter
  }
  return v ? 1u : 0u;
}

__global__ void __launch_bounds__(256) attention_mask_kernel(const int64_t* __restrict__ cluster_indices,
                                                             uint8_t* __restrict__ mask, int N, int C) {
  extern __shared__ int lab[];  // [N] labels of this image
  __shared__ int wmax[8];
  const int b = blockIdx.y;
  const int C2 = 2 * C, L = C2 + N;
  int mx = -1;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const int v = static_cast<int>(cluster_indices[static_cast<size_t>(b) * N + i]);
    lab[i] = v;
    mx = max(mx, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = mx;
  __syncthreads();
  int nb = -1;
  for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) nb = max(nb, wmax[w]);
  nb += 1;

  const size_t total = static_cast<size_t>(L) * L;
  uint8_t* out = mask + static_cast<size_t>(b) * total;
  // 4 output bytes per thread where the image block allows aligned 32-bit stores
  const bool vec = (total & 3) == 0;
  if (vec) {
    const size_t words = total >> 2;
    for (size_t w = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; w < words;
         w += static_cast<size_t>(gridDim.x) * blockDim.x) {
      const size_t e = w << 2;
      int q = static_cast<int>(e / L), k = static_cast<int>(e - static_cast<size_t>(q) * L);
      uint32_t word = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        word |= predicate(q, k, C2, nb, lab) << (8 * j);
        if (++k == L) { k = 0; ++q; }
      }
      reinterpret_cast<uint32_t*>(out)[w] = word;
    }
  } else {
    for (size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<size_t>(gridDim.x) * blockDim.x) {
      const int q = static_cast<int>(e / L), k = static_cast<int>(e - static_cast<size_t>(q) * L);
      out[e] = static_cast<uint8_t>(predicate(q, k, C2, nb, lab));
    }
  }
}

}  // namespace mask
}  // namespace msvit

extern "C" int msvit_attention_mask(const int64_t* cluster_indices, uint8_t* mask, int B, int N, int C,
                                    msvit_stream_t stream_) {
  using namespace msvit;
  if (!cluster_indices || !mask) return MSVIT_ERR_NULL;
  if (B < 0 || N <= 0 || C <= 0 || N > 8192 || C > 4096 || B > 65535) return MSVIT_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(mask) & 3) != 0) return MSVIT_ERR_ALIGN;
  if (B == 0) return MSVIT_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t L = static_cast<size_t>(2) * C + N;
  const size_t per_block = 256 * 4 * 32;  // ~32 words per thread: the per-block label load and reduction amortise
  int gx = static_cast<int>((L * L + per_block - 1) / per_block);
  if (gx < 1) gx = 1;
  if (gx > 1024) gx = 1024;
  mask::attention_mask_kernel<<<dim3(gx, B), 256, static_cast<size_t>(N) * sizeof(int), stream>>>(cluster_indices, mask, N, C);
  return cuda_status(cudaGetLastError());
}

// ----------------------------------------------------------------------------- cluster-compressed attention
// Transmitter statistics of compress_tokens_with_cluster_indices
// (model/multistate_encoder/modeling_msvitencoder.py:182-186):
//     out[b, h, q, c] = sum_{k : cluster(b, k) = c} attn[b, h, q, k]
// (the reference multiplies by a one-hot [N, C] mask through a 5-D broadcast and sums).  One warp per attention row,
// 128-bit loads along k, C predicated accumulators per lane, fixed-order warp reduction: the attention tensor is
// read exactly once (HBM-read bound, B*H*N*N*4 bytes).  The receiver statistics of the same function (:187-190,
// mean over the QUERY tokens of a cluster) are msvit_pool applied to the [B*H, N, N] view.
namespace msvit {
namespace mask {

template <int CMAX>
__global__ void __launch_bounds__(256) key_sums_kernel(const float* __restrict__ attn,
                                                       const int64_t* __restrict__ cluster_indices,
                                                       float* __restrict__ out, int H, int N, int C) {
  extern __shared__ int lab[];  // [N rounded up to 4] labels of this image (pad = -1)
  const int bh = blockIdx.y, b = bh / H;
  const int N4 = (N + 3) & ~3;
  for (int i = threadIdx.x; i < N4; i += blockDim.x)
    lab[i] = i < N ? static_cast<int>(cluster_indices[static_cast<size_t>(b) * N + i]) : -1;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const bool vec = (N & 3) == 0;
  for (int q = blockIdx.x * nwarps + warp; q < N; q += gridDim.x * nwarps) {
    const float* row = attn + (static_cast<size_t>(bh) * N + q) * N;
    float acc[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) acc[c] = 0.f;
    if (vec) {
      for (int k = 4 * lane; k < N; k += 128) {
        const float4 a = *reinterpret_cast<const float4*>(row + k);
        const int4 l = *reinterpret_cast<const int4*>(lab + k);
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
          acc[c] += (l.x == c ? a.x : 0.f) + (l.y == c ? a.y : 0.f) + (l.z == c ? a.z : 0.f) + (l.w == c ? a.w : 0.f);
      }
    } else {
      for (int k = lane; k < N; k += 32) {
        const float a = row[k];
        const int l = lab[k];
#pragma unroll
        for (int c = 0; c < CMAX; ++c) acc[c] += l == c ? a : 0.f;
      }
    }
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
    }
    float* o = out + (static_cast<size_t>(bh) * N + q) * C;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (lane == (c & 31) && c < C) o[c] = acc[c];   // (lane l writes clusters l, l + 32)
  }
}


// Both statistics from ONE pass over the attention tensor.  grid (cluster range z, image x head); the CTA groups the
// image's queries by cluster (stable, group_tokens) and walks its clusters: a warp takes every 8th query row of the
// cluster, reduces it over the keys of each cluster (transmitter row, as key_sums_kernel) and adds the same registers
// into its running column sums; the eight warps' column sums are combined in warp order and divided by the cluster
// size (receiver row).  Every attention element is loaded once: B*H*N*N*4 bytes read, (2 C / N) of that written.
constexpr int kStatsChunks = 8;   // per-lane register chunks along the key axis: N <= 32 * 8 * VEC

// Key sums for many clusters (C > 16): the predicated accumulators above cost C operations per attention element.
// Here the image's keys are grouped by cluster once per CTA (stable, group_tokens); a warp stages one attention row in
// shared memory (coalesced 128-bit loads) and lane c, c + 32, .. sums the row over the key list of its clusters in list
// order -- N shared-memory reads per row whatever C is, fixed summation order (bit-reproducible).
__global__ void __launch_bounds__(256) key_sums_sorted_kernel(const float* __restrict__ attn,
                                                              const int64_t* __restrict__ cluster_indices,
                                                              float* __restrict__ out, int H, int N, int C) {
  extern __shared__ __align__(16) int sm[];
  const int N4 = (N + 3) & ~3;
  float* rows = reinterpret_cast<float*>(sm);                            // [warps][N4], 16-byte aligned rows
  int* lab = sm + static_cast<size_t>(blockDim.x >> 5) * N4;             // [N4]
  int* order = lab + N4;            // [N]   keys grouped by cluster
  int* start = order + N;           // [C+1]
  int* cursor = start + C + 1;      // [C]
  const int bh = blockIdx.y, b = bh / H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  group_tokens(cluster_indices + static_cast<size_t>(b) * N, lab, order, start, cursor, N, C);
  float* mine = rows + static_cast<size_t>(warp) * N4;
  const bool vec = (N & 3) == 0;
  for (int q = blockIdx.x * nwarps + warp; q < N; q += gridDim.x * nwarps) {
    const float* row = attn + (static_cast<size_t>(bh) * N + q) * N;
    if (vec) {
      for (int k = 4 * lane; k < N; k += 128)
        *reinterpret_cast<float4*>(mine + k) = __ldcs(reinterpret_cast<const float4*>(row + k));
    } else {
      for (int k = lane; k < N; k += 32) mine[k] = __ldcs(row + k);
    }
    __syncwarp();
    float* o = out + (static_cast<size_t>(bh) * N + q) * C;
    for (int c = lane; c < C; c += 32) {
      const int s0 = start[c], s1 = start[c + 1];
      float sum = 0.f;
      int p = s0;
      for (; p + 4 <= s1; p += 4) {
        const float a0 = mine[order[p]], a1 = mine[order[p + 1]], a2 = mine[order[p + 2]], a3 = mine[order[p + 3]];
        sum += a0; sum += a1; sum += a2; sum += a3;
      }
      for (; p < s1; ++p) sum += mine[order[p]];
      o[c] = sum;
    }
    __syncwarp();
  }
}

// Sum CMAX per-lane values over the warp with a transposing butterfly: at offsets 16, 8, ... a lane keeps one half of
// its live values and hands the other half to its partner, so the 5 steps cost CMAX - 1 (+ the steps left once a
// single value remains) shuffles instead of 5 CMAX.  On return acc[i], i < R, of lane l holds the total of column
// (l >> (5 - H)) * R + i, with H = min(5, log2 CMAX) and R = CMAX >> H.
template <int CMAX>
struct ColumnReduce {
  static constexpr int H = CMAX >= 32 ? 5 : (CMAX == 16 ? 4 : (CMAX == 8 ? 3 : (CMAX == 4 ? 2 : 1)));
  static constexpr int R = CMAX >> H;
  static __device__ __forceinline__ int column(int lane, int i) { return (lane >> (5 - H)) * R + i; }
  static __device__ __forceinline__ bool writer(int lane) { return (lane & ((1 << (5 - H)) - 1)) == 0; }
  template <int N, int O>
  static __device__ __forceinline__ void step(float (&acc)[CMAX], int lane) {
    if constexpr (O >= 1) {
      if constexpr (N > R) {
        const bool up = (lane & O) != 0;
#pragma unroll
        for (int i = 0; i < N / 2; ++i) {
          const float keep = up ? acc[i + N / 2] : acc[i];
          const float give = up ? acc[i] : acc[i + N / 2];
          acc[i] = keep + __shfl_xor_sync(0xffffffffu, give, O);
        }
        step<N / 2, O / 2>(acc, lane);
      } else {
#pragma unroll
        for (int i = 0; i < N; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], O);
        step<N, O / 2>(acc, lane);
      }
    }
  }
  static __device__ __forceinline__ void run(float (&acc)[CMAX], int lane) { step<CMAX, 16>(acc, lane); }
};

// One attention row in registers: lane l holds the VEC keys starting at (32 j + l) VEC, j < NCH (0 beyond N).
template <int VEC, int NCH>
__device__ __forceinline__ void load_row(const float* __restrict__ row, int N, int lane, float (&v)[NCH][VEC]) {
#pragma unroll
  for (int j = 0; j < NCH; ++j) {
    const int k = (j * 32 + lane) * VEC;
    if constexpr (VEC == 4) {
      const float4 a = k < N ? __ldcs(reinterpret_cast<const float4*>(row + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
      v[j][0] = a.x; v[j][1] = a.y; v[j][2] = a.z; v[j][3] = a.w;
    } else {
      v[j][0] = k < N ? __ldcs(row + k) : 0.f;
    }
  }
}

// key sums of the row per cluster (reduced over the warp, see ColumnReduce) and the row added to the column sums
template <int CMAX, int VEC, int NCH>
__device__ __forceinline__ void stats_row(const float (&v)[NCH][VEC], const int* lab, int N, int lane,
                                          float (&acc)[CMAX], float (&col)[NCH][VEC]) {
#pragma unroll
  for (int c = 0; c < CMAX; ++c) acc[c] = 0.f;
#pragma unroll
  for (int j = 0; j < NCH; ++j) {
    const int k = (j * 32 + lane) * VEC;
    if (k < N) {
      if constexpr (VEC == 4) {
        const int4 l = *reinterpret_cast<const int4*>(lab + k);
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
          acc[c] += (l.x == c ? v[j][0] : 0.f) + (l.y == c ? v[j][1] : 0.f) + (l.z == c ? v[j][2] : 0.f) +
                    (l.w == c ? v[j][3] : 0.f);
      } else {
        const int l = lab[k];
#pragma unroll
        for (int c = 0; c < CMAX; ++c) acc[c] += l == c ? v[j][0] : 0.f;
      }
#pragma unroll
      for (int e = 0; e < VEC; ++e) col[j][e] += v[j][e];
    }
  }
  ColumnReduce<CMAX>::run(acc, lane);
}

// the transmitter row of query q from the reduced accumulators (see ColumnReduce for who holds what)
template <int CMAX>
__device__ __forceinline__ void store_key_sums(float* __restrict__ o, const float (&acc)[CMAX], int lane, int C) {
  using CR = ColumnReduce<CMAX>;
  if (CR::writer(lane)) {
#pragma unroll
    for (int i = 0; i < CR::R; ++i) {
      const int c = CR::column(lane, i);
      if (c < C) o[c] = acc[i];
    }
  }
}

// NCH = register chunks per row: 2 (rows of at most 64 VEC keys, the usual 196 / 256-token case; the next row of the
// warp is requested before the current one is reduced, two rows in flight per warp) or 8 (no prefetch).
template <int CMAX, int VEC, int NCH>
__global__ void __launch_bounds__(256) attention_stats_kernel(const float* __restrict__ attn,
                                                              const int64_t* __restrict__ cluster_indices,
                                                              float* __restrict__ tr, float* __restrict__ rc, int H,
                                                              int N, int C) {
  extern __shared__ int sm[];
  const int N4 = (N + 3) & ~3;
  int* lab = sm;                    // [N4]   label or -1 (pad = -1)
  int* order = lab + N4;            // [N]    queries grouped by cluster
  int* start = order + N;           // [C+1]
  int* cursor = start + C + 1;      // [C]
  float* part = reinterpret_cast<float*>(cursor + C);   // [8][N4]
  const int bh = blockIdx.y, b = bh / H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  if (threadIdx.x < N4 - N) lab[N + threadIdx.x] = -1;
  group_tokens(cluster_indices + static_cast<size_t>(b) * N, lab, order, start, cursor, N, C);

  const int c0 = static_cast<int>(static_cast<long long>(blockIdx.x) * C / gridDim.x);
  const int c1 = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * C / gridDim.x);
  const float* base = attn + static_cast<size_t>(bh) * N * N;
  constexpr bool kPrefetch = NCH <= 2;
  float acc[CMAX];
  float col[NCH][VEC];
  float cur[NCH][VEC];
  float nxt[kPrefetch ? NCH : 1][VEC];
  for (int c = c0; c < c1; ++c) {
    const int s0 = start[c], s1 = start[c + 1];
#pragma unroll
    for (int j = 0; j < NCH; ++j)
#pragma unroll
      for (int v = 0; v < VEC; ++v) col[j][v] = 0.f;
    int i = s0 + warp;
    if (kPrefetch && i < s1) load_row<VEC, NCH>(base + static_cast<size_t>(order[i]) * N, N, lane, cur);
    for (; i < s1; i += nwarps) {
      const int q = order[i];
      if constexpr (kPrefetch) {
        if (i + nwarps < s1) load_row<VEC, NCH>(base + static_cast<size_t>(order[i + nwarps]) * N, N, lane, nxt);
      } else {
        load_row<VEC, NCH>(base + static_cast<size_t>(q) * N, N, lane, cur);
      }
      stats_row<CMAX, VEC, NCH>(cur, lab, N, lane, acc, col);
      store_key_sums<CMAX>(tr + (static_cast<size_t>(bh) * N + q) * C, acc, lane, C);
      if constexpr (kPrefetch) {
#pragma unroll
        for (int j = 0; j < NCH; ++j)
#pragma unroll
          for (int v = 0; v < VEC; ++v) cur[j][v] = nxt[j][v];
      }
    }
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const int k = (j * 32 + lane) * VEC;
      if (k < N) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) part[warp * N4 + k + v] = col[j][v];
      }
    }
    __syncthreads();
    const float inv = s1 > s0 ? 1.0f / static_cast<float>(s1 - s0) : 0.f;
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
      float s = 0.f;
      for (int w = 0; w < nwarps; ++w) s += part[w * N4 + k];
      rc[(static_cast<size_t>(bh) * C + c) * N + k] = s * inv;
    }
    __syncthreads();
  }
  // queries whose label is outside [0, C) belong to no receiver row but still have a transmitter row
  if (blockIdx.x == 0) {
    for (int q = warp; q < N; q += nwarps) {
      if (lab[q] >= 0) continue;
      load_row<VEC, NCH>(base + static_cast<size_t>(q) * N, N, lane, cur);
      stats_row<CMAX, VEC, NCH>(cur, lab, N, lane, acc, col);
      store_key_sums<CMAX>(tr + (static_cast<size_t>(bh) * N + q) * C, acc, lane, C);
    }
  }
}

template <int CMAX, int VEC, int NCH>
static int launch_stats(const float* attn, const int64_t* ci, float* tr, float* rc, int B, int H, int N, int C,
                        cudaStream_t stream) {
  const int N4 = (N + 3) & ~3;
  const size_t ints = static_cast<size_t>(N4) + N + 2 * C + 1;
  const size_t smem = sizeof(int) * ints + sizeof(float) * 8 * static_cast<size_t>(N4);
  long long z = (4LL * sm_count() + static_cast<long long>(B) * H - 1) / (static_cast<long long>(B) * H);
  if (z > 8) z = 8;
  if (z > C) z = C;
  if (z < 1) z = 1;
  cudaError_t e = cudaFuncSetAttribute(attention_stats_kernel<CMAX, VEC, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return cuda_status(e);
  attention_stats_kernel<CMAX, VEC, NCH><<<dim3(static_cast<unsigned>(z), B * H), 256, smem, stream>>>(attn, ci, tr, rc, H,
                                                                                                N, C);
  return cuda_status(cudaGetLastError());
}

template <int VEC, int NCH>
static int dispatch_stats(const float* attn, const int64_t* ci, float* tr, float* rc, int B, int H, int N, int C,
                          cudaStream_t stream) {
  if (C <= 8) return launch_stats<8, VEC, NCH>(attn, ci, tr, rc, B, H, N, C, stream);
  return launch_stats<16, VEC, NCH>(attn, ci, tr, rc, B, H, N, C, stream);
}

}  // namespace mask
}  // namespace msvit

extern "C" int msvit_cluster_key_sums(const float* attn, const int64_t* cluster_indices, float* out, int B, int H, int N,
                                      int C, msvit_stream_t stream_) {
  using namespace msvit;
  if (!attn || !cluster_indices || !out) return MSVIT_ERR_NULL;
  if (B < 0 || H <= 0 || N <= 0 || C <= 0 || C > 64 || N > 8192 || static_cast<long long>(B) * H > 65535) return MSVIT_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(attn) & 15) != 0) return MSVIT_ERR_ALIGN;
  if (B == 0) return MSVIT_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int gx = (N + 63) / 64;  // 8 warps x 8 rows per CTA
  const dim3 grid(gx, B * H);
  const size_t smem = static_cast<size_t>((N + 3) & ~3) * sizeof(int);
  if (C <= 8) mask::key_sums_kernel<8><<<grid, 256, smem, stream>>>(attn, cluster_indices, out, H, N, C);
  else if (C <= 16) mask::key_sums_kernel<16><<<grid, 256, smem, stream>>>(attn, cluster_indices, out, H, N, C);
  else {
    // one accumulator per cluster does not scale: keys grouped by cluster, rows staged in shared memory
    const size_t N4 = static_cast<size_t>((N + 3) & ~3);
    int warps = 8;
    while (warps > 1 && warps * N4 * sizeof(float) > 40 * 1024) warps >>= 1;
    const size_t smem2 = sizeof(int) * (N4 + N + 2 * static_cast<size_t>(C) + 1) + sizeof(float) * warps * N4;
    cudaError_t e = cudaFuncSetAttribute(mask::key_sums_sorted_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem2));
    if (e != cudaSuccess) return cuda_status(e);
    const dim3 grid2((N + 8 * warps - 1) / (8 * warps), B * H);   // 8 rows per warp
    mask::key_sums_sorted_kernel<<<grid2, 32 * warps, smem2, stream>>>(attn, cluster_indices, out, H, N, C);
  }
  return cuda_status(cudaGetLastError());
}

// transmitter [B, H, N, C] and receiver [B, H, C, N] statistics in one pass over attn [B, H, N, N]
// (modeling_msvitencoder.py:182-190).  The one-pass kernel keeps a row and one accumulator per cluster in registers:
// it is the faster route for N <= 256 and C <= 16 (measured, tools/microbench/next_rows_time.py) and returns
// MSVIT_ERR_SHAPE beyond that, where msvit_cluster_key_sums + msvit_pool (two passes) win.
extern "C" int msvit_cluster_attention_stats(const float* attn, const int64_t* cluster_indices, float* transmitter,
                                             float* receiver, int B, int H, int N, int C, msvit_stream_t stream_) {
  using namespace msvit;
  if (!attn || !cluster_indices || !transmitter || !receiver) return MSVIT_ERR_NULL;
  if (B < 0 || H <= 0 || N <= 0 || C <= 0 || C > 16 || N > 256 || static_cast<long long>(B) * H > 65535) return MSVIT_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(attn) & 15) != 0) return MSVIT_ERR_ALIGN;
  if (B == 0) return MSVIT_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if ((N & 3) == 0) return mask::dispatch_stats<4, 2>(attn, cluster_indices, transmitter, receiver, B, H, N, C, stream);
  return mask::dispatch_stats<1, mask::kStatsChunks>(attn, cluster_indices, transmitter, receiver, B, H, N, C, stream);
}
