// Dataset-level ("DeepCluster-style") k-means over a row shard of a feature matrix (sm_100a).
//
// Reference: the flattened-batch clustering of model/clustering/modeling_spectral.py:254-256 (all B*N tokens at
// once) and the KMeans(n_clusters).fit_predict call sites (:90, :130-133); BASELINE.json configs[4] runs it on
// 1M x 768 CLS features with k = 1000, rows sharded over the GPUs of a box.
//
// One Lloyd iteration on a shard is four launches, all deterministic (no floating-point atomics):
//   msvit_gkm_assign      label_i = argmin_c |c|^2 - 2 x_i.c : the n x k x D contraction runs on tcgen05 (operands
//                         TMA-staged in 128B-swizzled shared memory, accumulators in TMEM), the squared centroid
//                         norms are taken from the staged tiles, and the epilogue reduces every accumulator row to
//                         a running (min, argmin) -- the n x k score matrix never exists in memory.
//   msvit_gkm_sort        stable counting sort of the row ids by label (per-block histograms, scan, ranked scatter)
//   msvit_gkm_accumulate  per-centroid sums of the member rows in sorted order: one CTA per centroid, 128-bit
//                         loads, fixed reduction order; emits the packed [k, D + 1] buffer (sums | count) that the
//                         host all-reduces over NCCL when the rows are sharded
//   msvit_gkm_finalize    centroid = sum / count (an empty cluster keeps its centre), fp32 master + operand copy
#include "tile_ops.cuh"

namespace msvit {
namespace gkm {

constexpr int kThreads = 320;
constexpr int kEpiThreads = 256;
constexpr int kTile = 256;         // rows of x / centroids per job
constexpr int kSliceBytes = 128;   // bytes of one row of a k-slice (64 bf16 or 32 fp32) = swizzle span
constexpr int kTileBytes = kTile * kSliceBytes;  // 32 KB
constexpr int kStages = 3;
constexpr int kStageBytes = 2 * kTileBytes;      // centroid tile + feature tile
constexpr int kTmemCols = 512;

struct Params {
  int32_t* labels;
  float* best;     // may be NULL
  int n, k;
  int n_kslices;   // ceil(D * elsize / 128)
  int k_step;      // elements per k-slice (64 bf16 / 32 fp32)
};

struct Shared {
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t conv[kStages];  // fp32 input: the staged tiles have been rounded to TF32
  uint64_t tmem_full;
  uint64_t tmem_empty;
  uint32_t tmem_base;
  uint32_t pad;
  float colq[kTile];       // |c|^2 of the centroid rows of the current job, as the tensor core sees them
  float bests[2][kTile];
  int besti[2][kTile];
};

// Item = 256 rows of x against all centroids, walked in column blocks of 256 centroids.  Job (item, nt): a 256 x 256
// score tile as two M=128 UMMA tiles in the two halves of TMEM.
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 = centroid norms
// (and TF32 rounding) while the k-loop runs, then the (min, argmin) epilogue.
template <bool TF32>
__global__ void __launch_bounds__(kThreads, 1)
assign_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_c, const Params P) {
  extern __shared__ uint8_t smem_raw[];
  // (pointer arithmetic on the shared array itself keeps the shared state space: LDS / STS instead of generic accesses)
  uint8_t* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  Shared& sh = *reinterpret_cast<Shared*>(tiles + static_cast<size_t>(kStages) * kStageBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&sh.full[i], 1);
      mbar_init(&sh.empty[i], 1 + kEpiThreads / 32);
      mbar_init(&sh.conv[i], kEpiThreads / 32);
    }
    mbar_init(&sh.tmem_full, 1);
    mbar_init(&sh.tmem_empty, kEpiThreads / 32);
    fence_mbar_init();
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_c);
  }
  if (warp == 1) {
    tmem_alloc(&sh.tmem_base, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh.tmem_base;

  const int n_items = ceil_div(P.n, kTile);
  const int n_nt = ceil_div(P.k, kTile);
  uint32_t it = 0;   // k-slice counter (ring position)
  uint32_t job = 0;  // job counter (TMEM phase)

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int rows_here = min(kTile, P.n - item * kTile);
        for (int nt = 0; nt < n_nt; ++nt) {
          const int n_umma = round_up(min(kTile, P.k - nt * kTile), 16);
          for (int ks = 0; ks < P.n_kslices; ++ks, ++it) {
            const int st = it % kStages;
            const uint32_t ph = (it / kStages) & 1;
            mbar_wait(&sh.empty[st], ph ^ 1);
            uint8_t* bt = tiles + static_cast<size_t>(st) * kStageBytes;
            // whole 128-row boxes; rows past the end of either matrix are zero-filled by the TMA unit
            const int cbox = n_umma > 128 ? 2 : 1, xbox = rows_here > 128 ? 2 : 1;
            mbar_arrive_expect_tx(&sh.full[st], static_cast<uint32_t>(cbox + xbox) * 128 * kSliceBytes);
            for (int j = 0; j < cbox; ++j)
              tma_load_2d(bt + j * 128 * kSliceBytes, &tm_c, &sh.full[st], ks * P.k_step, nt * kTile + j * 128);
            for (int j = 0; j < xbox; ++j)
              tma_load_2d(bt + kTileBytes + j * 128 * kSliceBytes, &tm_x, &sh.full[st], ks * P.k_step,
                          item * kTile + j * 128);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (lane 0 issues and commits)
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int rows_here = min(kTile, P.n - item * kTile);
      const bool m2 = rows_here > 128;
      for (int nt = 0; nt < n_nt; ++nt, ++job) {
        const int n_umma = round_up(min(kTile, P.k - nt * kTile), 16);
        const uint32_t idesc = make_idesc(TF32 ? 2u : 1u, 128u, static_cast<uint32_t>(n_umma));
        mbar_wait(&sh.tmem_empty, (job & 1) ^ 1);
        tc_fence_after();
        for (int ks = 0; ks < P.n_kslices; ++ks, ++it) {
          const int st = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(TF32 ? &sh.conv[st] : &sh.full[st], ph);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t b_addr = smem_u32(tiles + static_cast<size_t>(st) * kStageBytes);
            const uint32_t a_addr = b_addr + kTileBytes;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {  // 4 x 32 bytes of K per 128-byte slice
              const uint64_t bd = make_kmajor_sw128_desc(b_addr + kk * 32);
              const uint32_t acc = (ks | kk) != 0 ? 1u : 0u;
              umma_ss<TF32>(tmem_base, make_kmajor_sw128_desc(a_addr + kk * 32), bd, idesc, acc);
              if (m2)
                umma_ss<TF32>(tmem_base + 256, make_kmajor_sw128_desc(a_addr + 128 * kSliceBytes + kk * 32), bd,
                              idesc, acc);
            }
            tc_commit(&sh.empty[st]);
            if (ks == P.n_kslices - 1) tc_commit(&sh.tmem_full);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ norms + epilogue (warps 2..9)
    const int e = warp - 2;          // 0..7
    const int t = e * 32 + lane;     // 0..255: the tile row this thread owns in the staged tiles
    const int q = warp & 3;          // TMEM lane quadrant this warp may read
    const int cg = e >> 2;           // which half of the 16-column chunks this warp takes
    const int r0 = q * 32 + lane;    // accumulator row in M tile 0 (M tile 1: r0 + 128)
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int rows_here = min(kTile, P.n - item * kTile);
      float best0 = 3.0e38f, best1 = 3.0e38f;
      int idx0 = 0x7fffffff, idx1 = 0x7fffffff;
      for (int nt = 0; nt < n_nt; ++nt, ++job) {
        const int cols_here = min(kTile, P.k - nt * kTile);
        const int n_umma = round_up(cols_here, 16);
        float ssb = 0.f;
        for (int ks = 0; ks < P.n_kslices; ++ks, ++it) {
          const int st = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(&sh.full[st], ph);
          uint8_t* bt = tiles + static_cast<size_t>(st) * kStageBytes;
          if (t < n_umma) ssb += row_sumsq<TF32>(bt + t * kSliceBytes, lane);
          if constexpr (TF32) {
            if (t < rows_here) row_sumsq<TF32>(bt + kTileBytes + t * kSliceBytes, lane);  // rounds the feature row
            fence_proxy_async_smem();  // the rounded tiles must be visible to the tensor core
          }
          __syncwarp();
          if (lane == 0) {
            if constexpr (TF32) mbar_arrive(&sh.conv[st]);
            mbar_arrive(&sh.empty[st]);
          }
        }
        sh.colq[t] = ssb;
        named_bar_sync(1, kEpiThreads);

        mbar_wait(&sh.tmem_full, job & 1);
        tc_fence_after();
        for (int c = cg; c * 16 < n_umma; c += 2) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            if (half * 128 + q * 32 >= rows_here) continue;  // warp-uniform
            float v[16];
            tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + half * 256 + c * 16, v);
            float b = half ? best1 : best0;
            int bi = half ? idx1 : idx0;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int col = c * 16 + i;
              const float s = fmaf(-2.0f, v[i], sh.colq[col]);
              if (col < cols_here && s < b) { b = s; bi = nt * kTile + col; }  // ascending scan: ties keep the lowest id
            }
            if (half) { best1 = b; idx1 = bi; } else { best0 = b; idx0 = bi; }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.tmem_empty);
        named_bar_sync(1, kEpiThreads);  // colq may be overwritten by the next job
      }
      sh.bests[cg][r0] = best0; sh.besti[cg][r0] = idx0;
      sh.bests[cg][r0 + 128] = best1; sh.besti[cg][r0 + 128] = idx1;
      named_bar_sync(1, kEpiThreads);
      if (t < rows_here) {
        float b = sh.bests[0][t];
        int bi = sh.besti[0][t];
        const float b2 = sh.bests[1][t];
        const int bi2 = sh.besti[1][t];
        if (b2 < b || (b2 == b && bi2 < bi)) { b = b2; bi = bi2; }
        P.labels[static_cast<size_t>(item) * kTile + t] = bi;
        if (P.best) P.best[static_cast<size_t>(item) * kTile + t] = b;
      }
      named_bar_sync(1, kEpiThreads);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// ----------------------------------------------------------------------------- stable counting sort by label
constexpr int kSortRows = 2048;     // rows per sort block (one warp each)
constexpr int kAccParts = 8;        // runs the member rows of a centroid are cut into (accumulate_bulk_kernel)

// hist[b][c] = number of rows of block b with label c (labels outside [0, k) are dropped)
__global__ void __launch_bounds__(128) hist_kernel(const int32_t* __restrict__ labels, int n, int k,
                                                   int32_t* __restrict__ hist) {
  extern __shared__ int32_t cnt[];
  for (int c = threadIdx.x; c < k; c += blockDim.x) cnt[c] = 0;
  __syncthreads();
  const int b = blockIdx.x;
  const int r0 = b * kSortRows, r1 = min(n, r0 + kSortRows);
  for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
    const int c = labels[r];
    if (c >= 0 && c < k) atomicAdd(&cnt[c], 1);  // integer counts: order does not matter
  }
  __syncthreads();
  for (int c = threadIdx.x; c < k; c += blockDim.x) hist[static_cast<size_t>(b) * k + c] = cnt[c];
}

// Offsets of the blocks inside each label: hist[b][c] <- number of rows of label c in the blocks before b, tot[c] = the
// label's row count (the workspace holds the totals behind the histograms).  A CTA owns 32 labels (lane = label, so
// every load is one 128-byte line) and cuts the blocks into one chunk per warp: chunk sums, a scan of the 16 sums per
// label in shared memory, then the running offsets -- two short passes instead of one walk over all the blocks.
constexpr int kOffWarps = 16;
__global__ void __launch_bounds__(32 * kOffWarps) block_offsets_kernel(int32_t* __restrict__ hist, int nb, int k,
                                                                       int32_t* __restrict__ tot) {
  __shared__ int32_t part[kOffWarps][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const int chunk = (nb + kOffWarps - 1) / kOffWarps;
  const int b0 = min(nb, warp * chunk), b1 = min(nb, b0 + chunk);
  int sum = 0;
  if (c < k)
    for (int b = b0; b < b1; ++b) sum += hist[static_cast<size_t>(b) * k + c];
  part[warp][lane] = sum;
  __syncthreads();
  int run = 0;
  for (int w = 0; w < warp; ++w) run += part[w][lane];
  if (c < k) {
    int b = b0;
    for (; b + 8 <= b1; b += 8) {
      int v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = hist[static_cast<size_t>(b + i) * k + c];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        hist[static_cast<size_t>(b + i) * k + c] = run;
        run += v[i];
      }
    }
    for (; b < b1; ++b) {
      const int v = hist[static_cast<size_t>(b) * k + c];
      hist[static_cast<size_t>(b) * k + c] = run;
      run += v;
    }
    if (warp == kOffWarps - 1) tot[c] = run;
  }
}

// seg_off = exclusive scan of the label totals (one CTA, 1024 labels per pass)
__global__ void __launch_bounds__(1024) label_scan_kernel(const int32_t* __restrict__ tot, int k,
                                                          int32_t* __restrict__ seg_off) {
  __shared__ int32_t warp_tot[32];
  __shared__ int32_t carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int c0 = 0; c0 < k; c0 += 1024) {
    const int c = c0 + threadIdx.x;
    const int v = c < k ? tot[c] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const int w = warp_tot[lane];
      int wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      warp_tot[lane] = wi - w;   // exclusive prefix of the warp totals
    }
    __syncthreads();
    const int base = carry + warp_tot[warp];
    if (c < k) seg_off[c] = base + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = base + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) seg_off[k] = carry;
}

// perm[seg_off[c] + hist[b][c] + rank of the row among the block's rows with label c] = row   (one warp per block).
// The block's labels are fetched into shared memory in one go (all loads in flight together); the 64 ordered steps of
// the stable placement then run at shared-memory latency instead of one global round trip each.
__global__ void __launch_bounds__(32) scatter_kernel(const int32_t* __restrict__ labels, int n, int k,
                                                     const int32_t* __restrict__ hist,
                                                     const int32_t* __restrict__ seg_off, int32_t* __restrict__ perm) {
  extern __shared__ int32_t cur[];          // [k] running output position per label, then [kSortRows] labels
  int32_t* lab = cur + k;
  const int b = blockIdx.x, lane = threadIdx.x;
  const int r0 = b * kSortRows, r1 = min(n, r0 + kSortRows);
  for (int i = lane; i < r1 - r0; i += 32) {
    const int c = labels[r0 + i];
    lab[i] = (c < 0 || c >= k) ? -1 : c;
  }
  for (int c = lane; c < k; c += 32) cur[c] = seg_off[c] + hist[static_cast<size_t>(b) * k + c];
  __syncwarp();
  for (int base = r0; base < r1; base += 32) {
    const int r = base + lane;
    const int c = r < r1 ? lab[r - r0] : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, c);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    int pos = 0;
    if (c >= 0) pos = cur[c] + rank;
    __syncwarp();
    if (c >= 0 && rank == 0) cur[c] += __popc(peers);  // one leader per label value
    __syncwarp();
    if (c >= 0) perm[pos] = r;
  }
}

// ----------------------------------------------------------------------------- centroid sums
// One CTA per centroid: RL row lanes x (D / V) column threads; row lane j sums the member rows j, j+RL, ... in sorted
// order (two rows in flight), the RL partial sums are combined in a fixed order.
// packed[c][0..D) = sum, packed[c][D] = count.
template <typename T>
__device__ __forceinline__ void add_row(float* acc, const uint4& a) {
  const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if constexpr (sizeof(T) == 4) {
      acc[i] += __uint_as_float(w[i]);
    } else {
      acc[2 * i] += __uint_as_float(w[i] << 16);
      acc[2 * i + 1] += __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(512) accumulate_kernel(const T* __restrict__ x, const int32_t* __restrict__ perm,
                                                         const int32_t* __restrict__ seg_off, int D, int RL,
                                                         float* __restrict__ packed) {
  constexpr int V = 16 / sizeof(T);  // elements per 128-bit load
  extern __shared__ float part[];    // [RL][D]
  const int c = blockIdx.x;
  const int cols = D / V;            // column threads per row lane
  const int j = threadIdx.x / cols, ct = threadIdx.x - j * cols;
  const int s0 = seg_off[c], s1 = seg_off[c + 1];
  float acc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc[i] = 0.f;
  // four member rows in flight per thread; the row ids of a batch are loaded first (they are the addresses of the
  // feature loads), so that one round trip to memory covers four rows instead of one
  int r = s0 + j;
  for (; r + 3 * RL < s1; r += 4 * RL) {
    const int i0 = perm[r], i1 = perm[r + RL], i2 = perm[r + 2 * RL], i3 = perm[r + 3 * RL];
    const uint4 a = *reinterpret_cast<const uint4*>(x + static_cast<size_t>(i0) * D + ct * V);
    const uint4 b = *reinterpret_cast<const uint4*>(x + static_cast<size_t>(i1) * D + ct * V);
    const uint4 c4 = *reinterpret_cast<const uint4*>(x + static_cast<size_t>(i2) * D + ct * V);
    const uint4 d4 = *reinterpret_cast<const uint4*>(x + static_cast<size_t>(i3) * D + ct * V);
    add_row<T>(acc, a);
    add_row<T>(acc, b);
    add_row<T>(acc, c4);
    add_row<T>(acc, d4);
  }
  for (; r < s1; r += RL) add_row<T>(acc, *reinterpret_cast<const uint4*>(x + static_cast<size_t>(perm[r]) * D + ct * V));
#pragma unroll
  for (int i = 0; i < V; ++i) part[j * D + ct * V + i] = acc[i];
  __syncthreads();
  float* out = packed + static_cast<size_t>(c) * (D + 1);
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float s = part[d];
    for (int l = 1; l < RL; ++l) s += part[l * D + d];
    out[d] = s;
  }
  if (threadIdx.x == 0) out[D] = static_cast<float>(s1 - s0);
}

// Same sums with the member rows fetched by asynchronous bulk copies (cp.async.bulk global -> shared, mbarrier
// completion): the row gather is pure memory latency (ncu: long-scoreboard stalls 41 per issue, 1.5 TB/s with
// register loads), so every CTA keeps two batches of R whole rows in flight without holding a register for them.
// One CTA per centroid, one thread per 16-byte column chunk; a thread adds its chunk of the rows in sorted order
// (one accumulator per column: fixed order, no atomics, no cross-thread reduction).
template <typename T>
__global__ void __launch_bounds__(256) accumulate_bulk_kernel(const T* __restrict__ x, const int32_t* __restrict__ perm,
                                                              const int32_t* __restrict__ seg_off, int D, int R,
                                                              int parts, float* __restrict__ packed,
                                                              float* __restrict__ partial) {
  constexpr int V = 16 / sizeof(T);
  extern __shared__ __align__(128) uint8_t rows_sm[];   // [2][R][row bytes]
  __shared__ uint64_t bar[2];
  const int rowbytes = D * static_cast<int>(sizeof(T));
  const int chunks = rowbytes >> 4;
  // CTA = (centroid c, part p): the member rows of a centroid are cut into `parts` equal runs (cluster sizes are very
  // unequal after a few Lloyd steps: one CTA per centroid leaves the largest cluster on the critical path); the
  // partial sums go to `partial` and are combined in part order by combine_parts_kernel
  const int c = blockIdx.x / parts, part = blockIdx.x - c * parts;
  const int c0 = seg_off[c], c1 = seg_off[c + 1];
  const int per = (c1 - c0 + parts - 1) / parts;
  const int s0 = min(c1, c0 + part * per), s1 = min(c1, s0 + per);
  const int nb = (s1 - s0 + R - 1) / R;
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  __syncthreads();
  auto issue = [&](int b) {   // warp 0: batch b -> buffer b & 1
    const int r0 = s0 + b * R;
    const int cnt = min(R, s1 - r0);
    uint8_t* buf = rows_sm + static_cast<size_t>(b & 1) * R * rowbytes;
    if (threadIdx.x == 0) mbar_arrive_expect_tx(&bar[b & 1], static_cast<uint32_t>(cnt) * rowbytes);
    __syncwarp();
    for (int i = threadIdx.x; i < cnt; i += 32)
      bulk_load_1d(buf + static_cast<size_t>(i) * rowbytes, x + static_cast<size_t>(perm[r0 + i]) * D, rowbytes, &bar[b & 1]);
  };
  if (threadIdx.x < 32 && nb > 0) issue(0);
  float acc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc[i] = 0.f;
  for (int b = 0; b < nb; ++b) {
    if (threadIdx.x < 32 && b + 1 < nb) issue(b + 1);   // its buffer was released by the barrier that ended batch b - 1
    mbar_wait(&bar[b & 1], (b >> 1) & 1);
    const int cnt = min(R, s1 - (s0 + b * R));
    const uint8_t* buf = rows_sm + static_cast<size_t>(b & 1) * R * rowbytes;
    if (threadIdx.x < chunks) {
      int i = 0;
      for (; i + 4 <= cnt; i += 4) {
        const uint4 a0 = *reinterpret_cast<const uint4*>(buf + static_cast<size_t>(i) * rowbytes + threadIdx.x * 16);
        const uint4 a1 = *reinterpret_cast<const uint4*>(buf + static_cast<size_t>(i + 1) * rowbytes + threadIdx.x * 16);
        const uint4 a2 = *reinterpret_cast<const uint4*>(buf + static_cast<size_t>(i + 2) * rowbytes + threadIdx.x * 16);
        const uint4 a3 = *reinterpret_cast<const uint4*>(buf + static_cast<size_t>(i + 3) * rowbytes + threadIdx.x * 16);
        add_row<T>(acc, a0); add_row<T>(acc, a1); add_row<T>(acc, a2); add_row<T>(acc, a3);
      }
      for (; i < cnt; ++i)
        add_row<T>(acc, *reinterpret_cast<const uint4*>(buf + static_cast<size_t>(i) * rowbytes + threadIdx.x * 16));
    }
    __syncthreads();
  }
  float* out = parts > 1 ? partial + static_cast<size_t>(blockIdx.x) * D : packed + static_cast<size_t>(c) * (D + 1);
  if (threadIdx.x < chunks) {
#pragma unroll
    for (int i = 0; i < V; ++i) out[threadIdx.x * V + i] = acc[i];
  }
  if (parts == 1 && threadIdx.x == 0) out[D] = static_cast<float>(c1 - c0);
}

// packed[c][d] = sum over the parts (in order) of partial[c][part][d]; packed[c][D] = member count
__global__ void __launch_bounds__(256) combine_parts_kernel(const float* __restrict__ partial,
                                                            const int32_t* __restrict__ seg_off, int D, int parts,
                                                            float* __restrict__ packed) {
  const int c = blockIdx.x;
  const float* in = partial + static_cast<size_t>(c) * parts * D;
  float* out = packed + static_cast<size_t>(c) * (D + 1);
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float v = in[d];
    for (int p = 1; p < parts; ++p) v += in[static_cast<size_t>(p) * D + d];
    out[d] = v;
  }
  if (threadIdx.x == 0) out[D] = static_cast<float>(seg_off[c + 1] - seg_off[c]);
}

template <typename T>
__global__ void __launch_bounds__(256) finalize_kernel(const float* __restrict__ packed, float* __restrict__ centroids,
                                                       T* __restrict__ centroids_op, int k, int D) {
  const size_t total = static_cast<size_t>(k) * D;
  for (size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
       e += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(e / D), d = static_cast<int>(e - static_cast<size_t>(c) * D);
    const float cnt = packed[static_cast<size_t>(c) * (D + 1) + D];
    float v = centroids[e];
    if (cnt > 0.f) v = packed[static_cast<size_t>(c) * (D + 1) + d] / cnt;
    centroids[e] = v;
    if (centroids_op) {
      if constexpr (sizeof(T) == 4) {
        reinterpret_cast<float*>(centroids_op)[e] = v;
      } else {
        // round to nearest even bf16
        uint32_t u = __float_as_uint(v);
        u += 0x7FFFu + ((u >> 16) & 1u);
        reinterpret_cast<uint16_t*>(centroids_op)[e] = static_cast<uint16_t>(u >> 16);
      }
    }
  }
}

static inline int sort_blocks(int64_t n) { return static_cast<int>((n + kSortRows - 1) / kSortRows); }

}  // namespace gkm
}  // namespace msvit

extern "C" int msvit_gkm_assign(const void* x, int x_dtype, const void* centroids_op, int32_t* labels, float* best,
                                int64_t n, int k, int D, msvit_stream_t stream_) {
  using namespace msvit;
  using namespace msvit::gkm;
  if (!x || !centroids_op || !labels) return MSVIT_ERR_NULL;
  if (x_dtype != MSVIT_F32 && x_dtype != MSVIT_BF16) return MSVIT_ERR_MODE;
  if (n < 0 || k <= 0 || D <= 0 || n > 0x7fffff00LL) return MSVIT_ERR_SHAPE;
  const bool f32 = x_dtype == MSVIT_F32;
  const int esz = f32 ? 4 : 2;
  if ((static_cast<int64_t>(D) * esz) % 16 != 0 || (reinterpret_cast<uintptr_t>(x) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(centroids_op) & 15) != 0)
    return MSVIT_ERR_ALIGN;
  if (n == 0) return MSVIT_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EncodeTiledFn enc = encode_fn();
  if (!enc) return MSVIT_ERR_DRIVER;
  Params P;
  P.labels = labels; P.best = best;
  P.n = static_cast<int>(n); P.k = k;
  P.k_step = f32 ? 32 : 64;
  P.n_kslices = ceil_div(D, P.k_step);
  CUtensorMap tm_x, tm_c;
  int rc = make_map(enc, &tm_x, x, f32, n, D, 128);
  if (rc != MSVIT_OK) return rc;
  rc = make_map(enc, &tm_c, centroids_op, f32, k, D, 128);
  if (rc != MSVIT_OK) return rc;
  const size_t smem = 1024 + static_cast<size_t>(kStages) * kStageBytes + sizeof(Shared);
  const int n_items = ceil_div(P.n, kTile);
  const int grid = n_items < sm_count() ? n_items : sm_count();
  cudaError_t e;
  if (f32) {
    e = cudaFuncSetAttribute(assign_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return cuda_status(e);
    assign_kernel<true><<<grid, kThreads, smem, stream>>>(tm_x, tm_c, P);
  } else {
    e = cudaFuncSetAttribute(assign_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return cuda_status(e);
    assign_kernel<false><<<grid, kThreads, smem, stream>>>(tm_x, tm_c, P);
  }
  return cuda_status(cudaGetLastError());
}

extern "C" size_t msvit_gkm_workspace_bytes(int64_t n, int k) {
  using namespace msvit::gkm;
  if (n < 0 || k <= 0) return 0;
  return (static_cast<size_t>(sort_blocks(n) > 0 ? sort_blocks(n) : 1) + 1) * k * sizeof(int32_t);   // histograms + label totals
}

extern "C" int msvit_gkm_sort(const int32_t* labels, int64_t n, int k, int32_t* perm, int32_t* seg_off, void* workspace,
                              size_t workspace_bytes, msvit_stream_t stream_) {
  using namespace msvit;
  using namespace msvit::gkm;
  if (!labels || !perm || !seg_off || !workspace) return MSVIT_ERR_NULL;
  if (n < 0 || k <= 0 || n > 0x7fffff00LL || k > 10000) return MSVIT_ERR_SHAPE;  // k counters + one block of labels live in shared memory (48 KB)
  if (workspace_bytes < msvit_gkm_workspace_bytes(n, k)) return MSVIT_ERR_WORKSPACE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int32_t* hist = static_cast<int32_t*>(workspace);
  const int nb = sort_blocks(n);
  const size_t smem = static_cast<size_t>(k) * sizeof(int32_t);
  if (nb > 0) hist_kernel<<<nb, 128, smem, stream>>>(labels, static_cast<int>(n), k, hist);
  int32_t* tot = hist + static_cast<size_t>(nb > 0 ? nb : 1) * k;
  block_offsets_kernel<<<ceil_div(k, 32), 32 * kOffWarps, 0, stream>>>(hist, nb, k, tot);
  label_scan_kernel<<<1, 1024, 0, stream>>>(tot, k, seg_off);
  if (nb > 0)
    scatter_kernel<<<nb, 32, smem + kSortRows * sizeof(int32_t), stream>>>(labels, static_cast<int>(n), k, hist, seg_off, perm);
  return cuda_status(cudaGetLastError());
}

extern "C" size_t msvit_gkm_accumulate_workspace_bytes(int k, int D) {
  if (k <= 0 || D <= 0) return 0;
  return static_cast<size_t>(k) * msvit::gkm::kAccParts * D * sizeof(float);
}

extern "C" int msvit_gkm_accumulate(const void* x, int x_dtype, const int32_t* perm, const int32_t* seg_off,
                                    float* packed, int64_t n, int k, int D, void* workspace, size_t workspace_bytes,
                                    msvit_stream_t stream_) {
  using namespace msvit;
  using namespace msvit::gkm;
  if (!x || !perm || !seg_off || !packed) return MSVIT_ERR_NULL;
  if (x_dtype != MSVIT_F32 && x_dtype != MSVIT_BF16) return MSVIT_ERR_MODE;
  const int V = x_dtype == MSVIT_F32 ? 4 : 8;
  if (n < 0 || k <= 0 || D <= 0 || D % V != 0 || D / V > 512) return MSVIT_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) return MSVIT_ERR_ALIGN;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int cols = D / V;
  if (cols <= 256) {
    // rows fetched by bulk copies: two batches of R rows per CTA in shared memory (about 48 KB, four CTAs per SM)
    const int rowbytes = cols * 16;
    int R = 24576 / rowbytes;
    R = R > 16 ? 16 : (R < 4 ? 4 : R);
    const size_t smem = 2u * static_cast<size_t>(R) * rowbytes;
    const int threads = round_up(cols < 32 ? 32 : cols, 32);
    // with a workspace the rows of every centroid are cut into kAccParts runs (load balance), else one CTA per centroid
    const int parts = (workspace && workspace_bytes >= msvit_gkm_accumulate_workspace_bytes(k, D)) ? kAccParts : 1;
    float* partial = static_cast<float*>(workspace);
    cudaError_t e;
    if (x_dtype == MSVIT_F32) {
      e = cudaFuncSetAttribute(accumulate_bulk_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e != cudaSuccess) return cuda_status(e);
      accumulate_bulk_kernel<float><<<k * parts, threads, smem, stream>>>(static_cast<const float*>(x), perm, seg_off, D, R,
                                                                          parts, packed, partial);
    } else {
      e = cudaFuncSetAttribute(accumulate_bulk_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e != cudaSuccess) return cuda_status(e);
      accumulate_bulk_kernel<uint16_t><<<k * parts, threads, smem, stream>>>(static_cast<const uint16_t*>(x), perm, seg_off, D,
                                                                             R, parts, packed, partial);
    }
    if (parts > 1) combine_parts_kernel<<<k, 256, 0, stream>>>(partial, seg_off, D, parts, packed);
    return cuda_status(cudaGetLastError());
  }
  const int RL = cols <= 128 ? 4 : (cols <= 256 ? 2 : 1);
  const int threads = RL * cols;
  const size_t smem = static_cast<size_t>(RL) * D * sizeof(float);
  if (x_dtype == MSVIT_F32)
    accumulate_kernel<float><<<k, threads, smem, stream>>>(static_cast<const float*>(x), perm, seg_off, D, RL, packed);
  else
    accumulate_kernel<uint16_t><<<k, threads, smem, stream>>>(static_cast<const uint16_t*>(x), perm, seg_off, D, RL,
                                                              packed);
  return cuda_status(cudaGetLastError());
}

extern "C" int msvit_gkm_finalize(const float* packed, float* centroids, void* centroids_op, int op_dtype, int k, int D,
                                  msvit_stream_t stream_) {
  using namespace msvit;
  using namespace msvit::gkm;
  if (!packed || !centroids) return MSVIT_ERR_NULL;
  if (op_dtype != MSVIT_F32 && op_dtype != MSVIT_BF16) return MSVIT_ERR_MODE;
  if (k <= 0 || D <= 0) return MSVIT_ERR_SHAPE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t total = static_cast<size_t>(k) * D;
  const int grid = static_cast<int>((total + 255) / 256 < 2048 ? (total + 255) / 256 : 2048);
  if (op_dtype == MSVIT_F32)
    finalize_kernel<float><<<grid, 256, 0, stream>>>(packed, centroids, static_cast<float*>(centroids_op), k, D);
  else
    finalize_kernel<uint16_t><<<grid, 256, 0, stream>>>(packed, centroids, static_cast<uint16_t*>(centroids_op), k, D);
  return cuda_status(cudaGetLastError());
}
