// Pairwise token affinity + NCut degree on the 5th-generation tensor cores (sm_100a).
//
//   G = X X^T                      tcgen05.mma, operands TMA-staged in 128B-swizzled shared memory,
//                                  accumulators in TMEM (bf16 input -> kind::f16, fp32 input -> kind::tf32)
//   d_ij from G and the row norms  fused epilogue (rbf / cosine / normprod distance)
//   A_ij = exp(-d_ij / gamma)      ex2.approx in the same epilogue
//   deg_i = sum_j A_ij             accumulated in registers while A streams out
//
// Reference math: sandbox/test.py:108-114, sandbox/ncut_euclidean.py:19,23-29,
// model/clustering/modeling_spectral.py:54-61 (gamma, distance modes).
//
// Work decomposition.  An item is (segment s, row block p): 256 rows of the segment against all of
// its columns.  A persistent CTA walks items round-robin; inside an item it walks column blocks nt of
// 256 columns.  One job (s, p, nt) accumulates a 256 x 256 tile as two M=128 UMMA tiles into the two
// halves of TMEM (columns [0,256) and [256,512)).  Because both operands are rows of X, a diagonal job
// (nt == p) loads ONE k-slice tile and uses it as the A operand (two 128-row views) and as the B operand.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..9 = row-norm accumulation while the k-loop runs, then the epilogue.
#include "tile_ops.cuh"

namespace msvit {

namespace aff {

constexpr int kThreads = 320;
constexpr int kEpiThreads = 256;
constexpr int kTile = 256;         // rows / columns per job
constexpr int kSliceBytes = 128;   // bytes of one row of a k-slice (64 bf16 or 32 fp32) = swizzle span
constexpr int kTileBytes = kTile * kSliceBytes;  // 32 KB: one 256-row k-slice tile
constexpr int kMaxStages = 6;
constexpr int kTmemCols = 512;

struct Params {
  float* A;
  float* deg;
  const int32_t* seg_off;
  const int64_t* a_off;
  int S, N;
  int mode;
  float c2;        // rbf / normprod: log2(e) / (gamma * scale);  cosine: log2(e) / gamma
  int n_kslices;   // ceil(D * elsize / 128)
  int k_step;      // elements per k-slice (64 bf16 / 32 fp32)
  int stages;
  int stage_bytes; // 32 KB when every job is diagonal (N <= 256), else 64 KB (B tile + A tile)
  int pb;          // row blocks per segment = ceil(N / 256)
  int tail_rows;   // rows of the small TMA box (0 = none)
};

struct Shared {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t conv[kMaxStages];  // fp32 input: the staged tile has been rounded to TF32 by the norm warps
  uint64_t tmem_full;
  uint64_t tmem_empty;
  uint32_t tmem_base;
  uint32_t pad;
  float rowq[kTile];
  float colq[kTile];
  float degp[2][kTile];
  // per-warp transposition tile of the epilogue: 32 accumulator rows x 16 columns (row stride 20 floats), so that
  // the affinity is stored as 64-byte row segments (8 rows per store instruction) instead of one 16-byte piece
  // per lane in 32 different rows
  alignas(16) float stage[kEpiThreads / 32][32 * 20];
};

// Per-row quantity cached for the epilogue.
__device__ __forceinline__ float row_quantity(int mode, float sumsq, float c2) {
  if (mode == MSVIT_DIST_RBF) return 0.5f * sumsq * c2;
  if (mode == MSVIT_DIST_COSINE) return rsqrtf(fmaxf(sumsq, 1e-30f));
  return sqrtf(sumsq);
}

// log2 of the affinity from the Gram entry and the two cached row quantities.
__device__ __forceinline__ float affinity_log2(int mode, float g, float rq, float cq, float c2) {
  float t;
  if (mode == MSVIT_DIST_RBF) {
    t = fmaf(g, c2, -cq) - rq;
  } else if (mode == MSVIT_DIST_COSINE) {
    t = fmaf(g * rq, cq, -1.0f) * c2;
  } else {
    t = (g - rq * cq) * c2;
  }
  return fminf(t, 0.0f);
}

template <bool TF32>
__global__ void __launch_bounds__(kThreads, 1)
affinity_kernel(const __grid_constant__ CUtensorMap tm_full, const __grid_constant__ CUtensorMap tm_tail,
                const Params P) {
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled tiles need 1024-byte alignment
  // (pointer arithmetic on the shared array itself keeps the shared state space: LDS / STS instead of generic accesses)
  uint8_t* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  Shared& sh = *reinterpret_cast<Shared*>(tiles + static_cast<size_t>(P.stages) * P.stage_bytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < P.stages; ++i) {
      mbar_init(&sh.full[i], 1);
      mbar_init(&sh.empty[i], 1 + kEpiThreads / 32);
      mbar_init(&sh.conv[i], kEpiThreads / 32);
    }
    mbar_init(&sh.tmem_full, 1);
    mbar_init(&sh.tmem_empty, kEpiThreads / 32);
    fence_mbar_init();
    tma_prefetch_desc(&tm_full);
    tma_prefetch_desc(&tm_tail);
  }
  if (warp == 1) {
    tmem_alloc(&sh.tmem_base, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh.tmem_base;

  const int n_items = P.S * P.pb;
  uint32_t it = 0;   // k-slice counter (ring position)
  uint32_t job = 0;  // job counter (TMEM phase)

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int s = item / P.pb, p = item % P.pb;
        const Seg g = seg_info(s, P.N, P.seg_off, P.a_off);
        if (p * kTile >= g.n) continue;
        const int rows_here = min(kTile, g.n - p * kTile);
        const int n_nt = ceil_div(g.n, kTile);
        for (int nt = 0; nt < n_nt; ++nt) {
          const int n_umma = round_up(min(kTile, g.n - nt * kTile), 16);
          const bool diag = (nt == p);
          for (int ks = 0; ks < P.n_kslices; ++ks, ++it) {
            const int st = it % P.stages;
            const uint32_t ph = (it / P.stages) & 1;
            mbar_wait(&sh.empty[st], ph ^ 1);
            uint8_t* bt = tiles + static_cast<size_t>(st) * P.stage_bytes;
            // decide the boxes first so that the expected byte count is armed before any copy lands
            int nbox = 0, box_rows[4], box_row0[4];
            uint8_t* box_dst[4];
            for (int j = 0; j * 128 < n_umma; ++j) {
              const int need = min(128, n_umma - j * 128);
              box_rows[nbox] = (P.tail_rows > 0 && need <= P.tail_rows) ? P.tail_rows : 128;
              box_row0[nbox] = g.row0 + nt * kTile + j * 128;
              box_dst[nbox] = bt + j * 128 * kSliceBytes;
              ++nbox;
            }
            if (!diag) {
              for (int j = 0; j * 128 < rows_here; ++j) {
                const int need = round_up(min(128, rows_here - j * 128), 16);
                box_rows[nbox] = (P.tail_rows > 0 && need <= P.tail_rows) ? P.tail_rows : 128;
                box_row0[nbox] = g.row0 + p * kTile + j * 128;
                box_dst[nbox] = bt + kTileBytes + j * 128 * kSliceBytes;
                ++nbox;
              }
            }
            uint32_t bytes = 0;
            for (int j = 0; j < nbox; ++j) bytes += box_rows[j] * kSliceBytes;
            mbar_arrive_expect_tx(&sh.full[st], bytes);
            for (int j = 0; j < nbox; ++j)
              tma_load_2d(box_dst[j], box_rows[j] == 128 ? &tm_full : &tm_tail, &sh.full[st], ks * P.k_step,
                          box_row0[j]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (lane 0 issues and commits)
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int s = item / P.pb, p = item % P.pb;
      const Seg g = seg_info(s, P.N, P.seg_off, P.a_off);
      if (p * kTile >= g.n) continue;
      const int rows_here = min(kTile, g.n - p * kTile);
      const bool m2 = rows_here > 128;
      const int n_nt = ceil_div(g.n, kTile);
      for (int nt = 0; nt < n_nt; ++nt, ++job) {
        const int n_umma = round_up(min(kTile, g.n - nt * kTile), 16);
        const bool diag = (nt == p);
        const uint32_t idesc = make_idesc(TF32 ? 2u : 1u, 128u, static_cast<uint32_t>(n_umma));
        mbar_wait(&sh.tmem_empty, (job & 1) ^ 1);
        tc_fence_after();
        for (int ks = 0; ks < P.n_kslices; ++ks, ++it) {
          const int st = it % P.stages;
          const uint32_t ph = (it / P.stages) & 1;
          mbar_wait(TF32 ? &sh.conv[st] : &sh.full[st], ph);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t b_addr = smem_u32(tiles + static_cast<size_t>(st) * P.stage_bytes);
            const uint32_t a_addr = diag ? b_addr : b_addr + kTileBytes;
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
    const int t = e * 32 + lane;     // 0..255: the tile row whose norm this thread accumulates
    const int q = warp & 3;          // TMEM lane quadrant this warp may read
    const int cg = e >> 2;           // which half of the 16-column chunks this warp takes
    const int r0 = q * 32 + lane;    // accumulator row in M tile 0 (M tile 1: r0 + 128)
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int s = item / P.pb, p = item % P.pb;
      const Seg g = seg_info(s, P.N, P.seg_off, P.a_off);
      if (p * kTile >= g.n) continue;
      const int rows_here = min(kTile, g.n - p * kTile);
      const int n_nt = ceil_div(g.n, kTile);
      float rowsum0 = 0.f, rowsum1 = 0.f;
      for (int nt = 0; nt < n_nt; ++nt, ++job) {
        const int cols_here = min(kTile, g.n - nt * kTile);
        const int n_umma = round_up(cols_here, 16);
        const bool diag = (nt == p);
        float ssb = 0.f, ssa = 0.f;
        for (int ks = 0; ks < P.n_kslices; ++ks, ++it) {
          const int st = it % P.stages;
          const uint32_t ph = (it / P.stages) & 1;
          mbar_wait(&sh.full[st], ph);
          uint8_t* bt = tiles + static_cast<size_t>(st) * P.stage_bytes;
          if (t < n_umma) ssb += row_sumsq<TF32>(bt + t * kSliceBytes, lane);
          if (!diag && t < rows_here) ssa += row_sumsq<TF32>(bt + kTileBytes + t * kSliceBytes, lane);
          if constexpr (TF32) fence_proxy_async_smem();  // the rounded tile must be visible to the tensor core
          __syncwarp();
          if (lane == 0) {
            if constexpr (TF32) mbar_arrive(&sh.conv[st]);
            mbar_arrive(&sh.empty[st]);
          }
        }
        sh.colq[t] = row_quantity(P.mode, ssb, P.c2);
        sh.rowq[t] = row_quantity(P.mode, diag ? ssb : ssa, P.c2);
        named_bar_sync(1, kEpiThreads);

        mbar_wait(&sh.tmem_full, job & 1);
        tc_fence_after();
        const float rq0 = sh.rowq[r0], rq1 = sh.rowq[r0 + 128];
        const bool v0 = r0 < rows_here, v1 = r0 + 128 < rows_here;
        float* __restrict__ arow0 = nullptr;
        float* __restrict__ arow1 = nullptr;
        if (P.A) {
          arow0 = P.A + g.a0 + static_cast<long long>(p * kTile + r0) * g.lda + nt * kTile;
          arow1 = arow0 + static_cast<long long>(128) * g.lda;
        }
        const int lda_here = g.lda - nt * kTile;  // columns of this block that exist in storage
        for (int c = cg; c * 16 < n_umma; c += 2) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            // warp-uniform: skip 32-row groups that lie wholly outside the segment
            if (half * 128 + q * 32 >= rows_here) continue;
            float v[16];
            tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + half * 256 + c * 16, v);
            const float rq = half ? rq1 : rq0;
            const bool valid = half ? v1 : v0;
            float* __restrict__ abase = half ? arow1 : arow0;  // this lane's own row (used for the row offset only)
            float* st = sh.stage[e];
            float rs = 0.f;
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              float a[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int col = c * 16 + i4 * 4 + i;
                const float l2 = affinity_log2(P.mode, v[i4 * 4 + i], rq, sh.colq[col], P.c2);
                a[i] = col < cols_here ? fast_exp2(l2) : 0.f;
                rs += a[i];
              }
              if (abase) *reinterpret_cast<float4*>(st + lane * 20 + i4 * 4) = make_float4(a[0], a[1], a[2], a[3]);
            }
            if (half) rowsum1 += rs; else rowsum0 += rs;
            if (abase) {
              __syncwarp();
              // lane l stores columns 4 (l % 4) .. +3 of the rows l / 4 + 8 pass of this 32-row group
              const int cc = c * 16 + (lane & 3) * 4;
              const int rbase = half * 128 + q * 32;  // first tile row of the group
              float* __restrict__ g0 = P.A + g.a0 + static_cast<long long>(p * kTile + rbase) * g.lda + nt * kTile + cc;
#pragma unroll
              for (int pass = 0; pass < 4; ++pass) {
                const int rr = (lane >> 2) + 8 * pass;
                if (rbase + rr < rows_here && cc < lda_here)
                  *reinterpret_cast<float4*>(g0 + static_cast<long long>(rr) * g.lda) =
                      *reinterpret_cast<const float4*>(st + rr * 20 + (lane & 3) * 4);
              }
              __syncwarp();
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.tmem_empty);
        named_bar_sync(1, kEpiThreads);  // rowq / colq may be overwritten by the next job
      }
      sh.degp[cg][r0] = rowsum0;
      sh.degp[cg][r0 + 128] = rowsum1;
      named_bar_sync(1, kEpiThreads);
      if (t < rows_here) P.deg[g.row0 + p * kTile + t] = sh.degp[0][t] + sh.degp[1][t];
      named_bar_sync(1, kEpiThreads);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace aff
}  // namespace msvit

extern "C" int msvit_affinity_degree(const void* x, int x_dtype, float* A, float* deg, int64_t total_rows, int S,
                                     int N, int D, int mode, float gamma, float scale, const int32_t* seg_off,
                                     const int64_t* a_off, msvit_stream_t stream_) {
  using namespace msvit;
  using namespace msvit::aff;
  if (!x || !deg) return MSVIT_ERR_NULL;
  if (x_dtype != MSVIT_F32 && x_dtype != MSVIT_BF16) return MSVIT_ERR_MODE;
  if (mode < MSVIT_DIST_RBF || mode > MSVIT_DIST_NORMPROD) return MSVIT_ERR_MODE;
  if (S < 0 || N <= 0 || D <= 0 || total_rows < 0 || !(gamma > 0.f) || !(scale > 0.f)) return MSVIT_ERR_SHAPE;
  if (!seg_off && total_rows != static_cast<int64_t>(S) * N) return MSVIT_ERR_SHAPE;
  if (total_rows > 0x7fffffffLL || static_cast<int64_t>(S) * ceil_div(N, kTile) > 0x7fffffffLL) return MSVIT_ERR_SHAPE;
  const bool f32 = x_dtype == MSVIT_F32;
  const int esz = f32 ? 4 : 2;
  if ((static_cast<int64_t>(D) * esz) % 16 != 0 || (reinterpret_cast<uintptr_t>(x) & 15) != 0) return MSVIT_ERR_ALIGN;
  if (A && (reinterpret_cast<uintptr_t>(A) & 15) != 0) return MSVIT_ERR_ALIGN;
  if (S == 0 || total_rows == 0) return MSVIT_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);

  EncodeTiledFn enc = encode_fn();
  if (!enc) return MSVIT_ERR_DRIVER;

  Params P;
  P.A = A;
  P.deg = deg;
  P.seg_off = seg_off;
  P.a_off = a_off;
  P.S = S;
  P.N = N;
  P.mode = mode;
  P.c2 = mode == MSVIT_DIST_COSINE ? kLog2e / gamma : kLog2e / (gamma * scale);
  P.k_step = f32 ? 32 : 64;
  P.n_kslices = ceil_div(D, P.k_step);
  P.pb = ceil_div(N, kTile);
  const bool all_diag = N <= kTile;
  P.stage_bytes = all_diag ? kTileBytes : 2 * kTileBytes;
  P.stages = all_diag ? 6 : 3;
  // small TMA box for the last 128-row group of a uniform segment (skips rows that would be discarded)
  const int last = (round_up(N, 16) - 1) % 128 + 1;
  P.tail_rows = (!seg_off && last < 128) ? last : 0;

  CUtensorMap tm_full, tm_tail;
  int rc = make_map(enc, &tm_full, x, f32, total_rows, D, 128);
  if (rc != MSVIT_OK) return rc;
  rc = make_map(enc, &tm_tail, x, f32, total_rows, D, P.tail_rows > 0 ? P.tail_rows : 128);
  if (rc != MSVIT_OK) return rc;

  const size_t smem = 1024 + static_cast<size_t>(P.stages) * P.stage_bytes + sizeof(Shared);
  const int n_items = S * P.pb;
  const int grid = n_items < sm_count() ? n_items : sm_count();
  cudaError_t e;
  if (f32) {
    e = cudaFuncSetAttribute(affinity_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return cuda_status(e);
    affinity_kernel<true><<<grid, kThreads, smem, stream>>>(tm_full, tm_tail, P);
  } else {
    e = cudaFuncSetAttribute(affinity_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return cuda_status(e);
    affinity_kernel<false><<<grid, kThreads, smem, stream>>>(tm_full, tm_tail, P);
  }
  return cuda_status(cudaGetLastError());
}
