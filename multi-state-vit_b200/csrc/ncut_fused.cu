// Fused affinity + NCut subspace iteration for whole images of up to 224 tokens (ViT-B/16: 196): the affinity matrix
// is produced in tensor memory and never leaves the SM.
//
//   G = X X^T                        tcgen05.mma (SS form), k-slices of the tokens TMA-staged in 128B-swizzled shared
//                                    memory, fp32 accumulators in TMEM (two M=128 tiles, T = round16(N) columns each)
//   A = exp(-d(G) / gamma)           epilogue IN PLACE in TMEM: every 16 fp32 accumulator columns become 8 columns of
//                                    packed fp16 pairs A_hi followed by 8 columns A_lo (A = A_hi + A_lo to ~2^-22);
//                                    deg = row sums of the stored values
//   Y = D^-1 A U                     tcgen05.mma (TS form: A operand read from TMEM, kind::f16): A_hi [U_hi | U_lo] and
//                                    A_lo U_hi accumulate into 32 more TMEM columns per tile; the block U (16 columns,
//                                    fp16 hi/lo pair scaled by 2^10) is the K-major shared-memory operand
//   G2 = Y^T D Y, H = U^T D Y        mma.sync TF32x3 on the transposed blocks in shared memory (eig_core.cuh)
//   U = Y L^-T,  L L^T = G2          Cholesky QR in the D inner product: register Cholesky on one warp, forward
//                                    substitution per token (thread = token, u and y live in registers)
//   stop                             when the leading columns span an invariant subspace to the tolerance (or at the
//                                    iteration cap); the Rayleigh-Ritz rotation of the result is left to the next
//                                    kernel (kmeans.cu: ritz_kmeans_kernel), where hundreds of segments overlap
//
// Reference math: sandbox/test.py:108-118 (affinity, degree, normalised operator, leading eigenvectors),
// model/clustering/modeling_spectral.py:54-61 (gamma, distance modes).  Same iteration as ncut_eig.cu, which streams
// the affinity from L2 and serves every other shape.
//
// One persistent CTA per SM, 320 threads: warp 0 = TMA producer (runs ahead into the next image while the
// eigensolver works), warp 1 = TMEM allocator + Gram MMA issuer, warps 2..9 = row norms / TF32 rounding during the
// k-loop, then epilogue and eigensolver (thread = token row; warps 2 and 6 also issue the TS-form MMAs of their tile).
#include <cuda_fp16.h>

#include <cstdlib>

#include "eig_core.cuh"
#include "tile_ops.cuh"

namespace msvit {
namespace fused {

#ifdef FUSED_PROFILE
// Development instrumentation: cycles the first compute thread of every CTA spends in each phase.
enum { FP_GRAM, FP_EPI, FP_INIT, FP_UOP, FP_PRODUCT, FP_DRAIN, FP_GRAMS, FP_TRIGGER, FP_CHOL, FP_SUBST, FP_OUTPUT, FP_COUNT };
__device__ unsigned long long g_fused_cycles[FP_COUNT];
#define FPHASE_BEGIN() long long fp_t0 = clock64()
#define FPHASE_END(ph)                                                                     \
  do {                                                                                     \
    const long long fp_t1 = clock64();                                                     \
    if (ct == 0) atomicAdd(&g_fused_cycles[ph], (unsigned long long)(fp_t1 - fp_t0));      \
    fp_t0 = fp_t1;                                                                         \
  } while (0)
#else
#define FPHASE_BEGIN()
#define FPHASE_END(ph)
#endif

constexpr int kThreads = 320;
constexpr int kComputeBase = 64;
constexpr int kCompute = 256;
constexpr int kSliceBytes = 128;
constexpr int kMaxStages = 6;
constexpr int kTmemCols = 512;
constexpr int kMB = 16;               // subspace block width (columns of U)
constexpr int kMaxT = 224;            // 2 * 224 + 64 accumulator columns fill the 512 of TMEM
constexpr int kTileCols = 224;        // TMEM column stride of the two affinity tiles (a multiple of 32)
constexpr float kUScale = 1024.f;     // |u| <= 1 for a D-orthonormal block (deg >= 1): fp16 operands never overflow
constexpr int kUopBytes = 4 * 32 * kSliceBytes;   // 4 k-slices of 64 tokens x 32 operand rows (U_hi | U_lo)
using G = ThreadGroup<kComputeBase, kCompute, 1>;

struct Params {
  float* deg;       // [rows]
  float* U;         // [rows, 16] D-orthonormal basis of the final subspace
  float* H;         // [S, 16, 16] projected operator U^T D (D^-1 A) U
  int32_t* iters;   // [S]
  int32_t* info;    // [S] 1 = the leading block met the tolerance, 0 = stopped at the iteration cap
  int S, N, T;
  int mode;
  float c2;
  int n_kslices, k_step, stages, stage_bytes, tail_rows;
  int m, kconv, max_iter, fast_iters;
  float tol, lam_floor;
  int debug;        // development: stop the segment early (0 = off)
};

struct Shared {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t conv[kMaxStages];
  uint64_t tmem_full;
  uint64_t tmem_empty;
  uint64_t ybar;
  uint32_t tmem_base;
  uint32_t pad;
  float rq[256];    // per-token quantity of the distance (row and column side are the same tokens)
};

// float offsets of the eigensolver's shared arrays
struct EigLayout {
  int Ut, Yt, dg, Gs, Hs, LT, pinv, misc, colred, gscr, total;
  int ldt;
};
__host__ __device__ inline EigLayout make_eig_layout(int N) {
  EigLayout L;
  L.ldt = eig::ldt_of(N);
  const int mm = round_up(kMB * (kMB + 1), 4);
  int o = 0;
  L.Ut = o;      o += kMB * L.ldt;
  L.Yt = o;      o += kMB * L.ldt;
  L.dg = o;      o += 256;
  L.Gs = o;      o += mm;
  L.Hs = o;      o += mm;
  L.LT = o;      o += 256;
  L.pinv = o;    o += 16;
  L.misc = o;    o += 8 + 3 * MSVIT_MAX_EIG_BLOCK;
  L.colred = o;  o += (kCompute / 32) * MSVIT_MAX_EIG_BLOCK;
  L.gscr = o;    o += (kCompute / 32) * 128;     // partial Gram tiles (weighted_grams splits the tokens over warps)
  L.total = o;
  return L;
}

__device__ __forceinline__ float row_quantity(int mode, float sumsq, float c2) {
  if (mode == MSVIT_DIST_RBF) return 0.5f * sumsq * c2;
  if (mode == MSVIT_DIST_COSINE) return rsqrtf(fmaxf(sumsq, 1e-30f));
  return sqrtf(sumsq);
}
__device__ __forceinline__ float affinity_log2(int mode, float g, float rq, float cq, float c2) {
  float t;
  if (mode == MSVIT_DIST_RBF) t = fmaf(g, c2, -cq) - rq;
  else if (mode == MSVIT_DIST_COSINE) t = fmaf(g * rq, cq, -1.0f) * c2;
  else t = (g - rq * cq) * c2;
  return fminf(t, 0.0f);
}

// tcgen05.mma, A from TMEM, kind::f16, issued by one elected lane of a converged warp (the loop around it stays
// warp-uniform, which keeps the operands in uniform registers).
__device__ __forceinline__ void umma_ts_f16_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}

// two floats -> one register of two fp16 (round to nearest), `lo` in the low half
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

template <bool TF32>
__global__ void __launch_bounds__(kThreads, 1)
ncut_fused_kernel(const __grid_constant__ CUtensorMap tm_full, const __grid_constant__ CUtensorMap tm_tail,
                  const Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* uop = tiles + static_cast<size_t>(P.stages) * P.stage_bytes;   // 1024-aligned (stage_bytes % 1024 == 0)
  Shared& sh = *reinterpret_cast<Shared*>(uop + kUopBytes);
  float* ef = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(&sh) + ((sizeof(Shared) + 15) & ~size_t(15)));
  const EigLayout L = make_eig_layout(P.N);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n = P.N, T = P.T;
  const int n_tiles = T > 128 ? 2 : 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < P.stages; ++i) {
      mbar_init(&sh.full[i], 1);
      mbar_init(&sh.empty[i], 1 + kCompute / 32);
      mbar_init(&sh.conv[i], kCompute / 32);
    }
    mbar_init(&sh.tmem_full, 1);
    mbar_init(&sh.tmem_empty, kCompute / 32);
    mbar_init(&sh.ybar, n_tiles);
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

  uint32_t it_ring = 0;  // k-slice counter (ring position)
  uint32_t job = 0;      // segment counter of this CTA (TMEM phase)

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const int rows0 = T > 128 ? 128 : T;
      const int rows1 = T - rows0;  // = P.tail_rows when > 0
      for (int s = blockIdx.x; s < P.S; s += gridDim.x) {
        const int row0 = s * n;
        for (int ks = 0; ks < P.n_kslices; ++ks, ++it_ring) {
          const int st = it_ring % P.stages;
          const uint32_t ph = (it_ring / P.stages) & 1;
          mbar_wait(&sh.empty[st], ph ^ 1);
          uint8_t* bt = tiles + static_cast<size_t>(st) * P.stage_bytes;
          mbar_arrive_expect_tx(&sh.full[st], static_cast<uint32_t>(T) * kSliceBytes);
          tma_load_2d(bt, rows0 == 128 ? &tm_full : &tm_tail, &sh.full[st], ks * P.k_step, row0);
          if (rows1 > 0) tma_load_2d(bt + 128 * kSliceBytes, &tm_tail, &sh.full[st], ks * P.k_step, row0 + 128);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ Gram MMA issuer (lane 0 issues and commits)
    const uint32_t idesc = make_idesc(TF32 ? 2u : 1u, 128u, static_cast<uint32_t>(T));
    for (int s = blockIdx.x; s < P.S; s += gridDim.x, ++job) {
      mbar_wait(&sh.tmem_empty, (job & 1) ^ 1);
      tc_fence_after();
      for (int ks = 0; ks < P.n_kslices; ++ks, ++it_ring) {
        const int st = it_ring % P.stages;
        const uint32_t ph = (it_ring / P.stages) & 1;
        mbar_wait(TF32 ? &sh.conv[st] : &sh.full[st], ph);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t b_addr = smem_u32(tiles + static_cast<size_t>(st) * P.stage_bytes);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {  // 4 x 32 bytes of K per 128-byte slice
            const uint64_t bd = make_kmajor_sw128_desc(b_addr + kk * 32);
            const uint32_t acc = (ks | kk) != 0 ? 1u : 0u;
            umma_ss<TF32>(tmem_base, bd, bd, idesc, acc);
            if (n_tiles > 1)
              umma_ss<TF32>(tmem_base + kTileCols, make_kmajor_sw128_desc(b_addr + 128 * kSliceBytes + kk * 32), bd, idesc,
                            acc);
          }
          tc_commit(&sh.empty[st]);
          if (ks == P.n_kslices - 1) tc_commit(&sh.tmem_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ norms, epilogue, eigensolver (warps 2..9)
    const int ct = G::tid();            // 0..255
    const int cw = ct >> 5;             // 0..7
    const int q = warp & 3;             // TMEM lane quadrant this warp may access
    const int tile = cw >> 2;           // M tile of this warp
    const int row = tile * 128 + q * 32 + lane;   // token row of this thread inside the segment
    const bool valid = row < n;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t a_col = tile * kTileCols;    // this tile's affinity columns
    const uint32_t y_col = 2 * kTileCols + 32 * tile;   // this tile's product accumulator
    float* Ut = ef + L.Ut;
    float* Yt = ef + L.Yt;
    float* dg = ef + L.dg;
    float* Gs = ef + L.Gs;
    float* Hs = ef + L.Hs;
    float* LT = ef + L.LT;
    float* pinv = ef + L.pinv;
    float* misc = ef + L.misc;          // [0] scalar broadcast, [2] trigger flag
    float* res = misc + 8 + MSVIT_MAX_EIG_BLOCK;
    float* colred = ef + L.colred;
    float* gscr = ef + L.gscr;
    const int ldt = L.ldt;
    const int m = P.m, ld = m + 1;
    const int me = m;                   // n > m is required by the launcher
    const int kk = P.kconv < me ? P.kconv : me;
    const int npad = round_up(n, 16);
    const int nks = npad >> 4;          // K = 16 steps of the product
    const float tol2 = P.tol * P.tol;
    const uint32_t idesc32 = make_idesc(0u, 128u, 32u), idesc16 = make_idesc(0u, 128u, 16u);
    const uint32_t uop_addr = smem_u32(uop);
    uint32_t yphase = 0;

    for (int s = blockIdx.x; s < P.S; s += gridDim.x, ++job) {
      const int row0 = s * n;
      FPHASE_BEGIN();
      // ---- row norms (and TF32 rounding of the staged tile) while the Gram k-loop runs
      float ss = 0.f;
      for (int ks = 0; ks < P.n_kslices; ++ks, ++it_ring) {
        const int st = it_ring % P.stages;
        const uint32_t ph = (it_ring / P.stages) & 1;
        mbar_wait(&sh.full[st], ph);
        uint8_t* bt = tiles + static_cast<size_t>(st) * P.stage_bytes;
        if (P.debug & 32) {   // timing experiment: no in-place rounding
          if (ct < T) ss += row_sumsq<false>(bt + ct * kSliceBytes, lane);
        } else {
          if (ct < T) ss += row_sumsq<TF32>(bt + ct * kSliceBytes, lane);
          if constexpr (TF32) fence_proxy_async_smem();  // the rounded tile must be visible to the tensor core
        }
        __syncwarp();
        if (lane == 0) {
          if constexpr (TF32) mbar_arrive(&sh.conv[st]);
          mbar_arrive(&sh.empty[st]);
        }
      }
      sh.rq[ct] = row_quantity(P.mode, ss, P.c2);
      G::sync();

      // ---- epilogue, in place in TMEM: 16 fp32 Gram columns -> 8 columns of fp16 pairs A_hi + 8 columns A_lo
      mbar_wait(&sh.tmem_full, job & 1);
      FPHASE_END(FP_GRAM);
      tc_fence_after();
      const float rq = sh.rq[row];
      float rowsum = 0.f;
      if (tile < n_tiles) {
        for (int c = 0; c * 16 < T; ++c) {
          float v[16];
          tmem_ld16(lane_addr + a_col + c * 16, v);
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int col = c * 16 + 2 * i;
            const float2 cq = *reinterpret_cast<const float2*>(&sh.rq[col]);
            const float a0 = col < n ? fast_exp2(affinity_log2(P.mode, v[2 * i], rq, cq.x, P.c2)) : 0.f;
            const float a1 = col + 1 < n ? fast_exp2(affinity_log2(P.mode, v[2 * i + 1], rq, cq.y, P.c2)) : 0.f;
            // hi = the value with its low 13 mantissa bits cleared (exactly an fp16 number for a >= 2^-14; below that
            // the conversion rounds by < 2^-25), lo = the exact remainder rounded to fp16: two packed conversions
            const float h0 = __uint_as_float(__float_as_uint(a0) & 0xffffe000u);
            const float h1 = __uint_as_float(__float_as_uint(a1) & 0xffffe000u);
            rowsum += a0 + a1;
            w[i] = pack_f16x2(h0, h1);
            w[8 + i] = pack_f16x2(a0 - h0, a1 - h1);
          }
          tmem_st16(lane_addr + a_col + c * 16, w);
        }
        tmem_wait_st();
      }
      if (valid) P.deg[row0 + row] = rowsum;
      const float d = valid ? rowsum : 0.f;
      const float dinv_s = (valid && rowsum > 0.f) ? (1.0f / kUScale) / rowsum : 0.f;  // 1 / (deg * operand scale)
      dg[row] = d;    // every row 0..255 is owned by exactly one thread; pad tokens get 0
      FPHASE_END(FP_EPI);

      // ---- start block (same as ncut_eig.cu): column 0 constant, the rest pseudo-random; pad tokens zero
      float u[16], y[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        u[c] = (valid && c < me) ? (c == 0 ? 1.f : eig::hash_unit(static_cast<uint32_t>(row), static_cast<uint32_t>(c))) : 0.f;
        y[c] = 0.f;
      }
      if (row < npad) {
#pragma unroll
        for (int c = 0; c < 16; ++c) Ut[c * ldt + row] = u[c];
      }

      int it = 0;
      bool fired = false;
      FPHASE_END(FP_INIT);
      while ((P.debug & 15) != 1) {
        ++it;
        const bool full = !(it <= P.fast_iters && it < P.max_iter);
        // ---- shared-memory operand: rows 0..15 = fp16(2^10 u), rows 16..31 = the fp16 remainder; K-major, 128B swizzle
        if (row < npad) {
          uint8_t* ob = uop + (row >> 6) * (32 * kSliceBytes);
          const int kc = row & 63;
          const int within = (kc & 7) * 2;
#pragma unroll
          for (int c = 0; c < 16; c += 2) {
            const float s0 = u[c] * kUScale, s1 = u[c + 1] * kUScale;
            const float h0 = __uint_as_float(__float_as_uint(s0) & 0xffffe000u);
            const float h1 = __uint_as_float(__float_as_uint(s1) & 0xffffe000u);
            const uint32_t hh = pack_f16x2(h0, h1);
            const int off0 = ((((kc >> 3) ^ (c & 7))) << 4) + within;
            const int off1 = ((((kc >> 3) ^ ((c + 1) & 7))) << 4) + within;
            *reinterpret_cast<uint16_t*>(ob + c * kSliceBytes + off0) = static_cast<uint16_t>(hh & 0xffffu);
            *reinterpret_cast<uint16_t*>(ob + (c + 1) * kSliceBytes + off1) = static_cast<uint16_t>(hh >> 16);
            if (full) {
              const uint32_t ll = pack_f16x2(s0 - h0, s1 - h1);
              *reinterpret_cast<uint16_t*>(ob + (16 + c) * kSliceBytes + off0) = static_cast<uint16_t>(ll & 0xffffu);
              *reinterpret_cast<uint16_t*>(ob + (17 + c) * kSliceBytes + off1) = static_cast<uint16_t>(ll >> 16);
            }
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        G::sync();
        if ((P.debug & 15) == 2) break;
        FPHASE_END(FP_UOP);
        // ---- Y = A U on the tensor cores (warps 2 and 6: one M tile each)
        if ((cw & 3) == 0 && tile < n_tiles) {
          tc_fence_after();
          const uint32_t a0 = tmem_base + a_col, d0 = tmem_base + y_col;
          for (int ks = 0; ks < nks; ++ks) {
            const uint64_t bd = make_kmajor_sw128_desc(uop_addr + (ks >> 2) * (32 * kSliceBytes) + (ks & 3) * 32);
            umma_ts_f16_elect(d0, a0 + 16 * ks, bd, full ? idesc32 : idesc16, ks ? 1u : 0u);
          }
          if (full) {
            for (int ks = 0; ks < nks; ++ks) {
              const uint64_t bd = make_kmajor_sw128_desc(uop_addr + (ks >> 2) * (32 * kSliceBytes) + (ks & 3) * 32);
              umma_ts_f16_elect(d0, a0 + 16 * ks + 8, bd, idesc16, 1u);
            }
          }
          tc_commit_elect(&sh.ybar);
        }
        mbar_wait(&sh.ybar, yphase);
        yphase ^= 1;
        tc_fence_after();
        FPHASE_END(FP_PRODUCT);
        if (tile < n_tiles) {
          float ya[16];
          tmem_ld16(lane_addr + y_col, ya);
          if (full) {
            float yb[16];
            tmem_ld16(lane_addr + y_col + 16, yb);
#pragma unroll
            for (int c = 0; c < 16; ++c) ya[c] += yb[c];
          }
#pragma unroll
          for (int c = 0; c < 16; ++c) y[c] = valid ? ya[c] * dinv_s : 0.f;
        }
        if (row < npad) {
#pragma unroll
          for (int c = 0; c < 16; ++c) Yt[c * ldt + row] = y[c];
        }
        tc_fence_before();
        G::sync();
        FPHASE_END(FP_DRAIN);
        if ((P.debug & 15) == 3) {
#pragma unroll
          for (int c = 0; c < 16; ++c) u[c] = y[c];
          break;
        }

        const bool last = it >= P.max_iter;
        // ---- G2 = Y^T D Y, H = U^T D Y
        eig::weighted_grams<G>(Ut, Yt, dg, n, m, ldt, kMB, Gs, Hs, true, gscr);
        FPHASE_END(FP_GRAMS);
        if ((P.debug & 15) == 4) break;
        bool test = !last && it >= 2 && it > P.fast_iters;
        if (test) {
          // cheap screen (see ncut_eig.cu): G_cc - sum_{a < kk} H_ac^2 far above the tolerance means not converged
          if (cw == 0) {
            float v = 0.f;
            for (int c = lane; c < kk; c += 32) {
              const float gcc = Gs[c * ld + c];
              float e = gcc;
              for (int a = 0; a < kk; ++a) e = fmaf(-Hs[a * ld + c], Hs[a * ld + c], e);
              e -= 64.f * tol2 + 8e-6f * gcc;
              if (Hs[c * ld + c] >= P.lam_floor) v = fmaxf(v, e);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
            if (lane == 0) misc[2] = v > 0.f ? 0.f : 1.f;
          }
          G::sync();
          test = misc[2] != 0.f;
          G::sync();
        }
        if (test) {
          // |y_c - U h_c|_D^2 + coupling to the trailing columns, summed over the wanted columns: bounds the residual
          // of every Ritz pair of the leading block
          eig::span_residuals<1, G>(Hs, m, Ut, Yt, dg, kk, npad, ldt, kMB, colred, res);
          if (cw == 0) {
            float tot = 0.f;
            for (int c = lane; c < kk; c += 32) {
              float v = res[c];
              for (int a = kk; a < me; ++a) v = fmaf(Hs[a * ld + c], Hs[a * ld + c], v);
              if (Hs[c * ld + c] >= P.lam_floor) tot = (P.debug & 16) ? fmaxf(tot, v) : tot + v;
            }
            if (P.debug & 16) {
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) tot = fmaxf(tot, __shfl_xor_sync(0xffffffffu, tot, o));
            } else {
              tot = warp_sum(tot);
            }
            if (lane == 0) misc[2] = tot <= tol2 ? 1.f : 0.f;
          }
          G::sync();
          fired = misc[2] != 0.f;
        }
        FPHASE_END(FP_TRIGGER);
        if (fired || last) break;

        // ---- U = orth_D(Y): Cholesky of G2 on one warp, forward substitution per token
        float piv = eig::cholesky_lt16<G>(Gs, m, me, LT, pinv, misc);
        FPHASE_END(FP_CHOL);
        eig::forward_subst16(y, LT, pinv, u);
        if ((P.debug & 15) == 5) break;
        if (piv < EIG_REORTH) {
          // ill-conditioned block (early iterations): orthonormalise once more
          G::sync();   // every thread has read LT / pinv
          if (row < npad) {
#pragma unroll
            for (int c = 0; c < 16; ++c) Ut[c * ldt + row] = u[c];
          }
          G::sync();
          eig::weighted_grams<G>(Ut, Ut, dg, n, m, ldt, kMB, Gs, Hs, false, gscr);
          piv = eig::cholesky_lt16<G>(Gs, m, me, LT, pinv, misc);
          float u2[16];
          eig::forward_subst16(u, LT, pinv, u2);
#pragma unroll
          for (int c = 0; c < 16; ++c) u[c] = u2[c];
        }
        if (row < npad) {
#pragma unroll
          for (int c = 0; c < 16; ++c) Ut[c * ldt + row] = u[c];
        }
        // (the barrier before the next product orders these writes before the next Gram step)
        FPHASE_END(FP_SUBST);
      }

      // ---- output: the basis, the projected operator, the verdict
      if (valid) {
        float4* uo = reinterpret_cast<float4*>(P.U + static_cast<size_t>(row0 + row) * kMB);
        uo[0] = make_float4(u[0], u[1], u[2], u[3]);
        uo[1] = make_float4(u[4], u[5], u[6], u[7]);
        uo[2] = make_float4(u[8], u[9], u[10], u[11]);
        uo[3] = make_float4(u[12], u[13], u[14], u[15]);
      }
      {
        const int a = ct >> 4, c = ct & 15;
        P.H[static_cast<size_t>(s) * 256 + ct] = (a < m && c < m) ? Hs[a * ld + c] : 0.f;
      }
      if (ct == 0) {
        P.iters[s] = it;
        P.info[s] = fired ? 1 : 0;
      }
      // the affinity in TMEM is dead: the Gram of the next image may overwrite it
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh.tmem_empty);
      G::sync();   // rq / dg / Hs may be overwritten by the next segment
      FPHASE_END(FP_OUTPUT);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace fused
}  // namespace msvit

#ifdef FUSED_PROFILE
extern "C" int msvit_fused_profile(unsigned long long* host_out, int reset) {
  using namespace msvit::fused;
  cudaError_t e = cudaMemcpyFromSymbol(host_out, g_fused_cycles, sizeof(unsigned long long) * FP_COUNT);
  if (e != cudaSuccess) return static_cast<int>(e);
  if (reset) {
    unsigned long long z[FP_COUNT] = {};
    e = cudaMemcpyToSymbol(g_fused_cycles, z, sizeof(z));
  }
  return static_cast<int>(e);
}
#endif

extern "C" int msvit_ncut_fused(const void* x, int x_dtype, float* deg, float* U, float* H, int32_t* iters,
                                int32_t* info, int64_t total_rows, int S, int N, int D, int mode, float gamma,
                                float scale, int block, int max_iter, float tol, float lam_floor, int n_converge,
                                msvit_stream_t stream_) {
  using namespace msvit;
  using namespace msvit::fused;
  if (!x || !deg || !U || !H || !iters || !info) return MSVIT_ERR_NULL;
  if (x_dtype != MSVIT_F32 && x_dtype != MSVIT_BF16) return MSVIT_ERR_MODE;
  if (mode < MSVIT_DIST_RBF || mode > MSVIT_DIST_NORMPROD) return MSVIT_ERR_MODE;
  if (S < 0 || N <= 0 || D <= 0 || total_rows < 0 || !(gamma > 0.f) || !(scale > 0.f)) return MSVIT_ERR_SHAPE;
  if (total_rows != static_cast<int64_t>(S) * N || total_rows > 0x7fffffffLL) return MSVIT_ERR_SHAPE;
  if (block != kMB || N <= kMB || round_up(N, 16) > kMaxT) return MSVIT_ERR_SHAPE;
  if (max_iter <= 0 || !(tol > 0.f) || n_converge < 0 || n_converge > kMB) return MSVIT_ERR_SHAPE;
  const bool f32 = x_dtype == MSVIT_F32;
  const int esz = f32 ? 4 : 2;
  if ((static_cast<int64_t>(D) * esz) % 16 != 0 || (reinterpret_cast<uintptr_t>(x) & 15) != 0) return MSVIT_ERR_ALIGN;
  if ((reinterpret_cast<uintptr_t>(U) & 15) != 0) return MSVIT_ERR_ALIGN;
  if (S == 0) return MSVIT_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);

  EncodeTiledFn enc = encode_fn();
  if (!enc) return MSVIT_ERR_DRIVER;

  Params P;
  P.deg = deg; P.U = U; P.H = H; P.iters = iters; P.info = info;
  P.S = S; P.N = N; P.T = round_up(N, 16);
  P.mode = mode;
  P.c2 = mode == MSVIT_DIST_COSINE ? kLog2e / gamma : kLog2e / (gamma * scale);
  P.k_step = f32 ? 32 : 64;
  P.n_kslices = ceil_div(D, P.k_step);
  P.stage_bytes = round_up(P.T * kSliceBytes, 1024);
  P.tail_rows = P.T > 128 ? P.T - 128 : P.T;
  P.m = block;
  P.kconv = n_converge > 0 ? n_converge : block;
  P.max_iter = max_iter;
  P.fast_iters = EIG_FAST_ITERS;
  P.tol = tol;
  P.lam_floor = lam_floor;
  {
    const char* dbg = getenv("MSVIT_FUSED_DEBUG");
    P.debug = dbg ? atoi(dbg) : 0;
  }

  const size_t fixed = 1024 + kUopBytes + ((sizeof(Shared) + 15) & ~size_t(15)) +
                       static_cast<size_t>(make_eig_layout(N).total) * sizeof(float);
  const size_t kMaxSmem = 227 * 1024;
  int stages = static_cast<int>((kMaxSmem - fixed) / P.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return MSVIT_ERR_SHAPE;
  P.stages = stages;
  const size_t smem = fixed + static_cast<size_t>(stages) * P.stage_bytes;

  CUtensorMap tm_full, tm_tail;
  int rc = make_map(enc, &tm_full, x, f32, total_rows, D, 128);
  if (rc != MSVIT_OK) return rc;
  rc = make_map(enc, &tm_tail, x, f32, total_rows, D, P.tail_rows);
  if (rc != MSVIT_OK) return rc;

  const int grid = S < sm_count() ? S : sm_count();
  cudaError_t e;
  if (f32) {
    e = cudaFuncSetAttribute(ncut_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return cuda_status(e);
    ncut_fused_kernel<true><<<grid, kThreads, smem, stream>>>(tm_full, tm_tail, P);
  } else {
    e = cudaFuncSetAttribute(ncut_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return cuda_status(e);
    ncut_fused_kernel<false><<<grid, kThreads, smem, stream>>>(tm_full, tm_tail, P);
  }
  return cuda_status(cudaGetLastError());
}
