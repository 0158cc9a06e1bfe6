// Shared device pieces of the batched NCut eigensolvers (ncut_eig.cu: affinity streamed from L2; ncut_fused.cu:
// affinity resident in tensor memory).  Every routine is run by a thread group G (a whole CTA or a subset of its
// warps with a named barrier): G::tid() in [0, G::kThreads), G::sync(), G::sync_or().
#pragma once
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace msvit {

// Threads [BASE, BASE + NTHREADS) of the CTA (BASE a multiple of 32) with hardware barrier BAR.  BASE = 0, BAR = 0 and
// NTHREADS = the CTA size is the whole CTA.
template <int BASE, int NTHREADS, int BAR>
struct ThreadGroup {
  static constexpr int kThreads = NTHREADS;
  static __device__ __forceinline__ int tid() { return static_cast<int>(threadIdx.x) - BASE; }
  static __device__ __forceinline__ void sync() {
    if constexpr (BAR == 0) __syncthreads();
    else named_bar_sync(BAR, NTHREADS);
  }
  static __device__ __forceinline__ bool sync_or(bool pred) {
    if constexpr (BAR == 0) {
      return __syncthreads_or(pred ? 1 : 0) != 0;
    } else {
      uint32_t r;
      asm volatile(
          "{\n\t.reg .pred p, q;\n\t"
          "setp.ne.u32 q, %3, 0;\n\t"
          "barrier.cta.red.or.pred p, %1, %2, q;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(r)
          : "r"(BAR), "r"(NTHREADS), "r"(pred ? 1u : 0u)
          : "memory");
      return r != 0;
    }
  }
};

}  // namespace msvit

// Tuning knobs (compile-time; the defaults are what ships, the others exist for tools/microbench/eig_variants.py)
#ifndef EIG_T          // 8-row tiles of A per matvec pass when m <= 16
#define EIG_T 4
#endif
#ifndef EIG_PF         // 16-column blocks of A loaded together per lane in the matvec (one L2 round trip per batch)
#define EIG_PF 3
#endif
#ifndef EIG_BATCH      // 1: 128-thread CTAs load EIG_PF blocks together, then multiply them (see matvec_pass); 0: rolling window
#define EIG_BATCH 1
#endif
#ifndef EIG_MINB       // CTAs per SM the register budget is sized for, at 128 threads
#define EIG_MINB 4
#endif
#ifndef EIG_RR_EVERY   // scheduled whole-block Rayleigh-Ritz period
#define EIG_RR_EVERY 4
#endif
#ifndef EIG_SWEEPS     // Jacobi sweeps of a scheduled (not final) Rayleigh-Ritz step
#define EIG_SWEEPS 2
#endif
// Cholesky QR loses orthogonality like eps / (smallest scaled pivot of the Gram matrix); it is repeated below this
// pivot.  Early iterates (far from the Ritz basis) are ill conditioned but only need a well-conditioned basis, not an
// orthonormal one: 1e-3 keeps |U^T D U - I| <~ 1e-4 there, and near convergence Y is almost D-orthogonal (pivot ~ 1).
#ifndef EIG_REORTH
#define EIG_REORTH 1e-3f
#endif
#ifndef EIG_FAST_ITERS // leading products done in a single TF32 pass
#define EIG_FAST_ITERS 3
#endif

namespace msvit {
namespace eig {

#ifdef EIG_PROFILE
// Development instrumentation: cycles thread 0 of every CTA spends in each phase (tools/microbench/eig_variants.py).
enum { PH_INIT, PH_MATVEC, PH_GRAMS, PH_TRIGGER, PH_JACOBI, PH_ROTATE, PH_CHOL, PH_ORTH, PH_OUTPUT, PH_FACT, PH_INV,
       PH_COUNT };
__device__ unsigned long long g_phase_cycles[PH_COUNT];
#define PHASE_BEGIN() long long ph_t0 = clock64()
#define PHASE_END(ph)                                                                          \
  do {                                                                                         \
    const long long ph_t1 = clock64();                                                         \
    if (threadIdx.x == 0) atomicAdd(&g_phase_cycles[ph], (unsigned long long)(ph_t1 - ph_t0)); \
    ph_t0 = ph_t1;                                                                             \
  } while (0)
#else
#define PHASE_BEGIN()
#define PHASE_END(ph)
#endif

struct Params {
  const float* A;
  const float* deg;
  float* V;
  float* lam;
  int32_t* iters;
  const int32_t* seg_off;
  const int64_t* a_off;
  int S, N;
  int k, m;
  int kconv;        // leading pairs that must meet the tolerance (<= k)
  int max_iter, rr_every;
  float tol;
  float lam_floor;  // wanted pairs whose Ritz value is below this are exempt from the residual test
  int fast_iters;   // the first fast_iters products use a single TF32 pass (far from convergence)
};

struct Layout {
  // offsets in floats from the dynamic shared memory base
  int Ut, Yt, Uf, dg, dinv, Gs, Hs, Ss, Ws, pinv, misc, colred, rot, total;
  int ldt, rows;
};

__host__ __device__ inline int ldt_of(int N) {
  const int np = round_up(N, 16);
  return (np % 32 == 16) ? np : np + 16;
}

// MT: 16-row tiles of the transposed blocks (m <= 16 * MT)
__host__ __device__ inline Layout make_layout(int N, int m, int MT, int nwarps) {
  Layout L;
  L.ldt = ldt_of(N);
  L.rows = round_up(m, 8);
  const int mm = round_up(m * (m + 1), 4);
  int o = 0;
  L.Ut = o;      o += L.rows * L.ldt;
  L.Yt = o;      o += L.rows * L.ldt;
  L.Uf = o;      o += (round_up(N, 16) / 16) * MT * 256;   // U in MMA fragment order (see matvec)
  L.dg = o;      o += round_up(N, 16);
  L.dinv = o;    o += round_up(N, 16);
  L.Gs = o;      o += mm;
  L.Hs = o;      o += mm;
  L.Ss = o;      o += mm;
  L.Ws = o;      o += mm;
  L.pinv = o;    o += MSVIT_MAX_EIG_BLOCK + 128;          // reciprocal pivots, then 2 x 64 floats of Cholesky scratch
  L.misc = o;    o += 8 + 3 * MSVIT_MAX_EIG_BLOCK;
  L.colred = o;  o += nwarps * MSVIT_MAX_EIG_BLOCK;
  L.rot = o;     o += 4 * (MSVIT_MAX_EIG_BLOCK / 2);
  L.total = o;
  return L;
}

__device__ __forceinline__ float hash_unit(uint32_t i, uint32_t c) {
  uint32_t h = i * 0x9E3779B1u + c * 0x85EBCA77u + 0x165667B1u;
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
  return static_cast<float>(static_cast<int32_t>(h)) * (1.0f / 2147483648.0f);
}

// ----------------------------------------------------------------------------- tensor-core pieces
// D += A(16x8, row) * B(8x8, col), TF32 operands (low 13 mantissa bits ignored), fp32 accumulate.
// Fragments (g = lane / 4, t = lane % 4):  a0 (g, t)  a1 (g+8, t)  a2 (g, t+4)  a3 (g+8, t+4);
// b0 (k = t, n = g)  b1 (k = t+4, n = g);  c0 (g, 2t)  c1 (g, 2t+1)  c2 (g+8, 2t)  c3 (g+8, 2t+1).
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// x = hi + lo exactly: hi = x truncated to TF32 (what the tensor core reads from x's bits), lo = the remainder.
__device__ __forceinline__ uint32_t hi_bits(float x) { return __float_as_uint(x) & 0xffffe000u; }
__device__ __forceinline__ uint32_t lo_bits(float x) { return __float_as_uint(x - __uint_as_float(hi_bits(x))); }
__device__ __forceinline__ void split4(const float4& v, uint32_t (&hi)[4], uint32_t (&lo)[4]) {
  hi[0] = hi_bits(v.x); lo[0] = lo_bits(v.x);
  hi[1] = hi_bits(v.y); lo[1] = lo_bits(v.y);
  hi[2] = hi_bits(v.z); lo[2] = lo_bits(v.z);
  hi[3] = hi_bits(v.w); lo[3] = lo_bits(v.w);
}

// Fragment-order copy of U^T ("Uf"): for the 16-token block kb, 16-row tile mt and k-step ks, lane (g, t) finds its
// left-operand fragment as one float4 at  ((kb * MT + mt) * 2 + ks) * 128 + lane * 4:
//     { U^T[16mt+g][j], U^T[16mt+g+8][j], U^T[16mt+g][j+1], U^T[16mt+g+8][j+1] },   j = 16 kb + 4 t + 2 ks
// i.e. inside a block the MMA's k index is permuted: k = t <-> token 4t + 2ks, k = t+4 <-> token 4t + 2ks + 1.
// The matching right operand of 8 rows of A is one 128-bit load per lane: x = A[i0+g][16kb+4t .. +3],
// (b0, b1) = (x.x, x.y) for ks = 0 and (x.z, x.w) for ks = 1.

// position of U^T[c][j] in the fragment-order copy
template <int MT>
__device__ __forceinline__ int uf_index(int c, int j) {
  const int kb = j >> 4, tt = (j >> 2) & 3, ks = (j >> 1) & 1, kh = j & 1;
  const int mt = c >> 4, half = (c >> 3) & 1, cg = c & 7;
  return ((kb * MT + mt) * 2 + ks) * 128 + (cg * 4 + tt) * 4 + kh * 2 + half;
}

// One 16-column block of the product for TC row tiles: acc[q] += Ufrag(kb) * x[q]^T.  The warp issues in order, so
// the MMAs are emitted round-robin over the TC independent accumulators (a dependent MMA would stall the issue slot
// for its whole latency).
template <int MT, int TC, bool FULL>
__device__ __forceinline__ void matvec_block(float (&acc)[TC][MT][4], const float4 (&x)[TC],
                                             const float* __restrict__ uf, int kb) {
  uint32_t xl[TC][4];
  if constexpr (FULL) {
#pragma unroll
    for (int q = 0; q < TC; ++q) {
      xl[q][0] = lo_bits(x[q].x); xl[q][1] = lo_bits(x[q].y); xl[q][2] = lo_bits(x[q].z); xl[q][3] = lo_bits(x[q].w);
    }
  }
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const float4 ua = *reinterpret_cast<const float4*>(uf + ((kb * MT + mt) * 2) * 128);
    const float4 ub = *reinterpret_cast<const float4*>(uf + ((kb * MT + mt) * 2 + 1) * 128);
    uint32_t uah[4], ual[4], ubh[4], ubl[4];
    split4(ua, uah, ual);
    split4(ub, ubh, ubl);
    if constexpr (FULL) {
#pragma unroll
      for (int q = 0; q < TC; ++q) mma_tf32(acc[q][mt], ual, __float_as_uint(x[q].x), __float_as_uint(x[q].y));
#pragma unroll
      for (int q = 0; q < TC; ++q) mma_tf32(acc[q][mt], uah, xl[q][0], xl[q][1]);
#pragma unroll
      for (int q = 0; q < TC; ++q) mma_tf32(acc[q][mt], ubl, __float_as_uint(x[q].z), __float_as_uint(x[q].w));
#pragma unroll
      for (int q = 0; q < TC; ++q) mma_tf32(acc[q][mt], ubh, xl[q][2], xl[q][3]);
    }
#pragma unroll
    for (int q = 0; q < TC; ++q) mma_tf32(acc[q][mt], uah, __float_as_uint(x[q].x), __float_as_uint(x[q].y));
#pragma unroll
    for (int q = 0; q < TC; ++q) mma_tf32(acc[q][mt], ubh, __float_as_uint(x[q].z), __float_as_uint(x[q].w));
  }
}

// TC consecutive 8-row tiles of A starting at tile0, all 16-column blocks: EIG_PF blocks in flight per lane (a
// rolling register window, refilled right after a block's MMAs are issued), rows past the end are clamped
// (their products are scaled by dinv = 0).
template <int MT, int TC, bool FULL, bool BATCH>
__device__ __forceinline__ void matvec_pass(const float* __restrict__ Ag, int lda, int n, int KB, bool last_ok,
                                            const float* __restrict__ uf, float* __restrict__ Yt,
                                            const float* __restrict__ dinv, int ldt, int rows, int tile0, int g,
                                            int t) {
  constexpr int PF = EIG_PF;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  uint32_t off[TC];
  float acc[TC][MT][4];
  float4 x[PF][TC];
#pragma unroll
  for (int q = 0; q < TC; ++q) {
    const int r = min(8 * (tile0 + q) + g, n - 1);
    off[q] = static_cast<uint32_t>(r) * static_cast<uint32_t>(lda) + 4u * t;
#ifdef EIG_FAKE_A  // experiment: every lane streams row 0 (always cached)
    off[q] = 4u * t;
#endif
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[q][mt][e] = 0.f;
  }
  if constexpr (BATCH) {
  // batch mode: all global loads of a warp complete on one hardware scoreboard, so a rolling prefetch window
  // degenerates to one exposed L2 round trip per block.  Instead PF blocks are loaded together, then multiplied:
  // one round trip per PF blocks.
  for (int kb0 = 0; kb0 < KB; kb0 += PF) {
#pragma unroll
    for (int s = 0; s < PF; ++s) {
      const int kb = kb0 + s;
      const bool ok = kb + 1 < KB || (kb < KB && last_ok);
      const float* ak = Ag + 16 * kb;
#pragma unroll
      for (int q = 0; q < TC; ++q) x[s][q] = ok ? __ldcg(reinterpret_cast<const float4*>(ak + off[q])) : zero4;
    }
#pragma unroll
    for (int s = 0; s < PF; ++s)
      if (kb0 + s < KB) matvec_block<MT, TC, FULL>(acc, x[s], uf, kb0 + s);
  }
  } else {
  // rolling window (CTAs with many warps hide the round trips across warps; measured better at 512 threads)
#pragma unroll
  for (int s = 0; s < PF; ++s) {
    const bool ok = s + 1 < KB || (s < KB && last_ok);
    const float* ak = Ag + 16 * s;
#pragma unroll
    for (int q = 0; q < TC; ++q) x[s][q] = ok ? __ldcg(reinterpret_cast<const float4*>(ak + off[q])) : zero4;
  }
  for (int kb0 = 0; kb0 < KB; kb0 += PF) {
#pragma unroll
    for (int s = 0; s < PF; ++s) {
      const int kb = kb0 + s;
      if (kb < KB) {  // warp-uniform
        matvec_block<MT, TC, FULL>(acc, x[s], uf, kb);
        const int nk = kb + PF;
        if (nk < KB) {
          const bool ok = nk + 1 < KB || last_ok;
          const float* ak = Ag + 16 * nk;
#pragma unroll
          for (int q = 0; q < TC; ++q) x[s][q] = ok ? __ldcg(reinterpret_cast<const float4*>(ak + off[q])) : zero4;
        }
      }
    }
  }
  }
#pragma unroll
  for (int q = 0; q < TC; ++q) {
    const int i = 8 * (tile0 + q) + 2 * t;
    const float2 dv = *reinterpret_cast<const float2*>(dinv + i);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int r0 = 16 * mt + g, r1 = r0 + 8;
      if (r0 < rows)
        *reinterpret_cast<float2*>(Yt + r0 * ldt + i) = make_float2(acc[q][mt][0] * dv.x, acc[q][mt][1] * dv.y);
      if (r1 < rows)
        *reinterpret_cast<float2*>(Yt + r1 * ldt + i) = make_float2(acc[q][mt][2] * dv.x, acc[q][mt][3] * dv.y);
    }
  }
}

// Y^T[c][i] = dinv[i] * sum_j U^T[c][j] A[i][j]   (A symmetric).  Each warp owns a contiguous range of 8-row tiles of
// A and works on up to T of them at a time so that one U fragment load feeds T MMAs; A comes straight from global
// memory (L2).
template <int MT, int T, bool FULL, class G>
__device__ __forceinline__ void matvec(const float* __restrict__ Ag, int lda, int n, const float* __restrict__ Uf,
                                       float* __restrict__ Yt, const float* __restrict__ dinv, int ldt, int rows) {
  const int warp = G::tid() >> 5, lane = G::tid() & 31;
  const int g = lane >> 2, t = lane & 3;
  const int KB = (n + 15) >> 4;
  const int ntile = (n + 7) >> 3;
  const int per = (ntile + G::kThreads / 32 - 1) / (G::kThreads / 32);
  const int tbeg = warp * per;
  const int tend = min(ntile, tbeg + per);
  // every 16-column block but the last lies inside the row; the last one is cut at lda (a multiple of 4)
  const bool last_ok = 16 * (KB - 1) + 4 * t < lda;
  const float* uf = Uf + lane * 4;
  constexpr bool BATCH = EIG_BATCH != 0 && (G::kThreads / 32 <= 4 || MT == 2);  // measured per configuration (profiles/r1c_summary.md)
  int tile0 = tbeg;
  for (; tile0 + T <= tend; tile0 += T)
    matvec_pass<MT, T, FULL, BATCH>(Ag, lda, n, KB, last_ok, uf, Yt, dinv, ldt, rows, tile0, g, t);
  const int left = tend - tile0;
  if constexpr (T > 3) {
    if (left == 3) matvec_pass<MT, 3, FULL, BATCH>(Ag, lda, n, KB, last_ok, uf, Yt, dinv, ldt, rows, tile0, g, t);
  }
  if constexpr (T > 2) {
    if (left == 2) matvec_pass<MT, 2, FULL, BATCH>(Ag, lda, n, KB, last_ok, uf, Yt, dinv, ldt, rows, tile0, g, t);
  }
  if (left == 1) matvec_pass<MT, 1, FULL, BATCH>(Ag, lda, n, KB, last_ok, uf, Yt, dinv, ldt, rows, tile0, g, t);
}

// G[a][c] = sum_i dg[i] Q^T[a][i] Q^T[c][i]   and   H[a][c] = sum_i dg[i] P^T[a][i] Q^T[c][i]   (m x m, row
// stride m + 1).  One warp per 16 x 8 output tile; when the group has at least twice as many warps as tiles and a
// scratch buffer is given (ksplit * tiles * 128 floats), the token range of a tile is split over ksplit warps and the
// partial tiles are added in a fixed order.  Operands are 128-bit loads of 4 consecutive tokens per lane (same k
// permutation on both sides).
template <class G>
__device__ __forceinline__ void weighted_grams(const float* __restrict__ Pt, const float* __restrict__ Qt,
                                               const float* __restrict__ dg, int n, int m, int ldt, int rows,
                                               float* __restrict__ Gs, float* __restrict__ Hs, bool want_h,
                                               float* __restrict__ scratch = nullptr) {
  constexpr int NW = G::kThreads / 32;
  const int warp = G::tid() >> 5, lane = G::tid() & 31;
  const int g = lane >> 2, t = lane & 3;
  const int KB = (n + 15) >> 4;
  const int MTg = (m + 15) >> 4, NTg = (m + 7) >> 3;
  const int per = MTg * NTg;
  const int ntiles = want_h ? 2 * per : per;
  const int ksplit = (scratch != nullptr && NW >= 2 * ntiles) ? NW / ntiles : 1;
  const int kb_per = ksplit > 1 ? (((KB + ksplit - 1) / ksplit + 1) & ~1) : KB;   // even: the loop takes two blocks a step
  const int ld = m + 1;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int item = warp; item < ntiles * ksplit; item += NW) {
    const int part = item / ntiles, tile = item - part * ntiles;
    const int kb_lo = part * kb_per, kb_hi = min(KB, kb_lo + kb_per);
    const int mat = tile / per, rem = tile - mat * per;
    const int mt = rem / NTg, nt = rem - mt * NTg;
    const bool v1 = 16 * mt + g + 8 < rows;
    const float* at = (mat == 0 ? Qt : Pt) + (16 * mt + g) * ldt + 4 * t;
    const float* bt = Qt + (8 * nt + g) * ldt + 4 * t;
    const float* dp = dg + 4 * t;
    // six independent accumulators (k-block parity x {lo*hi, hi*lo, hi*hi}): consecutive MMAs never depend
    float ac[2][3][4];
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int e = 0; e < 4; ++e) ac[p][c][e] = 0.f;
    for (int kb0 = kb_lo; kb0 < kb_hi; kb0 += 2) {
      uint32_t ah0[2][4], al0[2][4], ah1[2][4], al1[2][4], bh[2][4], bl[2][4];
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const int kb = kb0 + p;
        const bool in = kb < kb_hi;  // warp-uniform
        const float4 a0 = in ? *reinterpret_cast<const float4*>(at + 16 * kb) : zero4;
        const float4 a1 = (in && v1) ? *reinterpret_cast<const float4*>(at + 8 * ldt + 16 * kb) : zero4;
        float4 b = in ? *reinterpret_cast<const float4*>(bt + 16 * kb) : zero4;
        const float4 d = in ? *reinterpret_cast<const float4*>(dp + 16 * kb) : zero4;
        b.x *= d.x; b.y *= d.y; b.z *= d.z; b.w *= d.w;
        ah0[p][0] = hi_bits(a0.x); ah0[p][1] = hi_bits(a1.x); ah0[p][2] = hi_bits(a0.y); ah0[p][3] = hi_bits(a1.y);
        al0[p][0] = lo_bits(a0.x); al0[p][1] = lo_bits(a1.x); al0[p][2] = lo_bits(a0.y); al0[p][3] = lo_bits(a1.y);
        ah1[p][0] = hi_bits(a0.z); ah1[p][1] = hi_bits(a1.z); ah1[p][2] = hi_bits(a0.w); ah1[p][3] = hi_bits(a1.w);
        al1[p][0] = lo_bits(a0.z); al1[p][1] = lo_bits(a1.z); al1[p][2] = lo_bits(a0.w); al1[p][3] = lo_bits(a1.w);
        bh[p][0] = __float_as_uint(b.x); bh[p][1] = __float_as_uint(b.y);
        bh[p][2] = __float_as_uint(b.z); bh[p][3] = __float_as_uint(b.w);
        bl[p][0] = lo_bits(b.x); bl[p][1] = lo_bits(b.y); bl[p][2] = lo_bits(b.z); bl[p][3] = lo_bits(b.w);
      }
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        mma_tf32(ac[p][0], al0[p], bh[p][0], bh[p][1]);
        mma_tf32(ac[p][1], ah0[p], bl[p][0], bl[p][1]);
        mma_tf32(ac[p][2], ah0[p], bh[p][0], bh[p][1]);
      }
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        mma_tf32(ac[p][0], al1[p], bh[p][2], bh[p][3]);
        mma_tf32(ac[p][1], ah1[p], bl[p][2], bl[p][3]);
        mma_tf32(ac[p][2], ah1[p], bh[p][2], bh[p][3]);
      }
    }
    float val[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float lo = (ac[0][0][e] + ac[1][0][e]) + (ac[0][1][e] + ac[1][1][e]);
      const float hi = ac[0][2][e] + ac[1][2][e];
      val[e] = hi + lo;
    }
    if (ksplit > 1) {
      *reinterpret_cast<float4*>(scratch + (part * ntiles + tile) * 128 + lane * 4) = make_float4(val[0], val[1], val[2], val[3]);
    } else {
      float* out = mat == 0 ? Gs : Hs;
      const int r0 = 16 * mt + g, r1 = r0 + 8, c0 = 8 * nt + 2 * t, c1 = c0 + 1;
      if (r0 < m && c0 < m) out[r0 * ld + c0] = val[0];
      if (r0 < m && c1 < m) out[r0 * ld + c1] = val[1];
      if (r1 < m && c0 < m) out[r1 * ld + c0] = val[2];
      if (r1 < m && c1 < m) out[r1 * ld + c1] = val[3];
    }
  }
  G::sync();
  if (ksplit > 1) {
    for (int idx = G::tid(); idx < ntiles * 128; idx += G::kThreads) {
      const int tile = idx >> 7, l = (idx >> 2) & 31, e = idx & 3;
      float v = 0.f;
      for (int part = 0; part < ksplit; ++part) v += scratch[(part * ntiles + tile) * 128 + l * 4 + e];
      const int mat = tile / per, rem = tile - mat * per;
      const int mt = rem / NTg, nt = rem - mt * NTg;
      const int r = 16 * mt + (l >> 2) + 8 * (e >> 1), c = 8 * nt + 2 * (l & 3) + (e & 1);
      if (r < m && c < m) (mat == 0 ? Gs : Hs)[r * ld + c] = v;
    }
    G::sync();
  }
}

// Left operand W (m x m in shared memory, row stride ld) of the small products below, split once per warp.
template <int MT>
struct WFrag {
  uint32_t hi[MT][2 * MT][4];
  uint32_t lo[MT][2 * MT][4];
};

// w = W (or W^T) restricted to rows < mr and columns < mc
template <int MT>
__device__ __forceinline__ void load_wfrag(WFrag<MT>& w, const float* __restrict__ W, int ld, int mr, int mc,
                                           bool transpose) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int ks = 0; ks < 2 * MT; ++ks) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = 16 * mt + g + 8 * (e & 1), c = 8 * ks + t + 4 * (e >> 1);
        float x = 0.f;
        if (r < mr && c < mc) x = transpose ? W[c * ld + r] : W[r * ld + c];
        w.hi[mt][ks][e] = hi_bits(x);
        w.lo[mt][ks][e] = lo_bits(x);
      }
    }
  }
}

// acc[u][mt] = fragment of (W X^T)[16mt .. 16mt+15][i0[u] .. i0[u]+7] for NU token tiles at once.  MMAs are issued
// round-robin over 3 * NU * MT independent accumulators (lo*hi, hi*lo, hi*hi per tile).
template <int MT, int NU>
__device__ __forceinline__ void tile_product(const WFrag<MT>& w, const float* __restrict__ Xt, int ldt, int rows,
                                             const int (&i0)[NU], float (&acc)[NU][MT][4]) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  float l1[NU][MT][4], l2[NU][MT][4];
#pragma unroll
  for (int u = 0; u < NU; ++u)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[u][mt][e] = l1[u][mt][e] = l2[u][mt][e] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 2 * MT; ++ks) {
    const int ra = 8 * ks + t, rb = ra + 4;
    float x0[NU], x1[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      x0[u] = ra < rows ? Xt[ra * ldt + i0[u] + g] : 0.f;
      x1[u] = rb < rows ? Xt[rb * ldt + i0[u] + g] : 0.f;
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
      for (int u = 0; u < NU; ++u) mma_tf32(l1[u][mt], w.lo[mt][ks], __float_as_uint(x0[u]), __float_as_uint(x1[u]));
#pragma unroll
      for (int u = 0; u < NU; ++u) mma_tf32(l2[u][mt], w.hi[mt][ks], lo_bits(x0[u]), lo_bits(x1[u]));
#pragma unroll
      for (int u = 0; u < NU; ++u) mma_tf32(acc[u][mt], w.hi[mt][ks], __float_as_uint(x0[u]), __float_as_uint(x1[u]));
    }
  }
#pragma unroll
  for (int u = 0; u < NU; ++u)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[u][mt][e] += l1[u][mt][e] + l2[u][mt][e];
}

// Per-row sums held as fragment partials rs[mt][h] (row 16mt + 8h + g, summed over this lane's tokens):
// out[r] = total over the CTA for r < nrows.  Shuffle over the 4 lanes of a row, then a fixed-order sum over warps.
template <int MT, class G>
__device__ __forceinline__ void reduce_rows(const float (&rs)[MT][2], int nrows, float* __restrict__ colred,
                                            float* __restrict__ out) {
  const int warp = G::tid() >> 5, lane = G::tid() & 31;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float v = rs[mt][h];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      if (t == 0) colred[warp * MSVIT_MAX_EIG_BLOCK + 16 * mt + 8 * h + g] = v;
    }
  G::sync();
  if (G::tid() < nrows) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < G::kThreads / 32; ++w) v += colred[w * MSVIT_MAX_EIG_BLOCK + G::tid()];
    out[G::tid()] = v;
  }
  G::sync();
}

// U^T <- W X^T  (W = L^-1: Cholesky QR) for all tokens, written both as the plain transposed block and in MMA
// fragment order.  X^T may alias U^T: a warp reads its 8-token tile completely before it writes it.
template <int MT, class G>
__device__ __forceinline__ void orthonormalise(const float* __restrict__ W, int m, const float* Xt, float* Ut,
                                               float* __restrict__ Uf, int npad, int ldt, int rows) {
  const int warp = G::tid() >> 5, lane = G::tid() & 31;
  const int g = lane >> 2, t = lane & 3;
  WFrag<MT> w;
  load_wfrag<MT>(w, W, m + 1, m, m, false);
  // two 8-token tiles per step (npad is a multiple of 16): tile pairs are dealt round-robin to the warps
  for (int pair = warp; pair < (npad >> 4); pair += G::kThreads / 32) {
    const int i0[2] = {16 * pair, 16 * pair + 8};
    float acc[2][MT][4];
    tile_product<MT, 2>(w, Xt, ldt, rows, i0, acc);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int j = i0[u] + 2 * t;
      const int kb = j >> 4, tt = (j >> 2) & 3, ks = (j >> 1) & 1;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int r0 = 16 * mt + g, r1 = r0 + 8;
        if (r0 < rows) *reinterpret_cast<float2*>(Ut + r0 * ldt + j) = make_float2(acc[u][mt][0], acc[u][mt][1]);
        if (r1 < rows) *reinterpret_cast<float2*>(Ut + r1 * ldt + j) = make_float2(acc[u][mt][2], acc[u][mt][3]);
        *reinterpret_cast<float4*>(Uf + ((kb * MT + mt) * 2 + ks) * 128 + (g * 4 + tt) * 4) =
            make_float4(acc[u][mt][0], acc[u][mt][2], acc[u][mt][1], acc[u][mt][3]);
      }
    }
  }
  G::sync();
}

// [U^T; Y^T] <- W [U^T; Y^T]  (Rayleigh-Ritz rotation) and res[c] = |y_c - theta_c u_c|_D^2 for c < kk.
template <int MT, class G>
__device__ __forceinline__ void rotate(const float* __restrict__ W, int m, float* Ut, float* Yt,
                                       const float* __restrict__ dg, const float* __restrict__ theta, int kk,
                                       int npad, int ldt, int rows, float* __restrict__ colred,
                                       float* __restrict__ res) {
  const int warp = G::tid() >> 5, lane = G::tid() & 31;
  const int g = lane >> 2, t = lane & 3;
  WFrag<MT> w;
  load_wfrag<MT>(w, W, m + 1, m, m, false);
  float rs[MT][2], th[MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      rs[mt][h] = 0.f;
      const int r = 16 * mt + 8 * h + g;
      th[mt][h] = r < m ? theta[r] : 0.f;
    }
  for (int tile = warp; tile < (npad >> 3); tile += G::kThreads / 32) {
    const int i0 = 8 * tile;
    const int i1[1] = {i0};
    float au1[1][MT][4], ay1[1][MT][4];
    tile_product<MT, 1>(w, Ut, ldt, rows, i1, au1);
    tile_product<MT, 1>(w, Yt, ldt, rows, i1, ay1);
    float (&au)[MT][4] = au1[0];
    float (&ay)[MT][4] = ay1[0];
    const int j = i0 + 2 * t;
    const float2 d = *reinterpret_cast<const float2*>(dg + j);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int r0 = 16 * mt + g, r1 = r0 + 8;
      if (r0 < rows) {
        *reinterpret_cast<float2*>(Ut + r0 * ldt + j) = make_float2(au[mt][0], au[mt][1]);
        *reinterpret_cast<float2*>(Yt + r0 * ldt + j) = make_float2(ay[mt][0], ay[mt][1]);
      }
      if (r1 < rows) {
        *reinterpret_cast<float2*>(Ut + r1 * ldt + j) = make_float2(au[mt][2], au[mt][3]);
        *reinterpret_cast<float2*>(Yt + r1 * ldt + j) = make_float2(ay[mt][2], ay[mt][3]);
      }
      const float e0 = ay[mt][0] - th[mt][0] * au[mt][0], e1 = ay[mt][1] - th[mt][0] * au[mt][1];
      const float e2 = ay[mt][2] - th[mt][1] * au[mt][2], e3 = ay[mt][3] - th[mt][1] * au[mt][3];
      rs[mt][0] = fmaf(d.x * e0, e0, fmaf(d.y * e1, e1, rs[mt][0]));
      rs[mt][1] = fmaf(d.x * e2, e2, fmaf(d.y * e3, e3, rs[mt][1]));
    }
  }
  reduce_rows<MT, G>(rs, kk, colred, res);
}

// res[c] = |y_c - U h_c|_D^2 for c < kk  (h_c = column c of H): how far the leading columns are from span(U).
template <int MT, class G>
__device__ __forceinline__ void span_residuals(const float* __restrict__ Hs, int m, const float* __restrict__ Ut,
                                               const float* __restrict__ Yt, const float* __restrict__ dg, int kk,
                                               int npad, int ldt, int rows, float* __restrict__ colred,
                                               float* __restrict__ res) {
  const int warp = G::tid() >> 5, lane = G::tid() & 31;
  const int g = lane >> 2, t = lane & 3;
  WFrag<MT> w;
  load_wfrag<MT>(w, Hs, m + 1, kk, m, true);  // rows c < kk of H^T
  float rs[MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) rs[mt][0] = rs[mt][1] = 0.f;
  for (int tile = warp; tile < (npad >> 3); tile += G::kThreads / 32) {
    const int i0 = 8 * tile;
    const int i1[1] = {i0};
    float acc1[1][MT][4];
    tile_product<MT, 1>(w, Ut, ldt, rows, i1, acc1);
    float (&acc)[MT][4] = acc1[0];
    const int j = i0 + 2 * t;
    const float2 d = *reinterpret_cast<const float2*>(dg + j);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int r0 = 16 * mt + g, r1 = r0 + 8;
      if (r0 < kk) {
        const float2 y = *reinterpret_cast<const float2*>(Yt + r0 * ldt + j);
        const float e0 = y.x - acc[mt][0], e1 = y.y - acc[mt][1];
        rs[mt][0] = fmaf(d.x * e0, e0, fmaf(d.y * e1, e1, rs[mt][0]));
      }
      if (r1 < kk) {
        const float2 y = *reinterpret_cast<const float2*>(Yt + r1 * ldt + j);
        const float e2 = y.x - acc[mt][2], e3 = y.y - acc[mt][3];
        rs[mt][1] = fmaf(d.x * e2, e2, fmaf(d.y * e3, e3, rs[mt][1]));
      }
    }
  }
  reduce_rows<MT, G>(rs, kk, colred, res);
}

// ----------------------------------------------------------------------------- small dense pieces
// Cholesky G = L L^T of the leading me x me block (row stride m + 1) and W = L^-1, both by warp 0.
//   factorisation: lane i keeps row i in registers, row j is broadcast by shuffles (left-looking); L goes back to
//                  the lower triangle of G, the reciprocal pivots to pinv (0 for a dropped, rank-deficient column);
//   inverse:       lane j owns column j of W (forward substitution, L read back as warp-uniform loads).
// Returns (to every thread) the smallest pivot of the unit-diagonal-scaled matrix, i.e. a conditioning
// estimate that ignores column scaling.
template <int MB, class G>
__device__ __forceinline__ float cholesky_inverse(float* __restrict__ Gm, int m, int me, float* __restrict__ pinv,
                                                  float* __restrict__ W, float* __restrict__ misc) {
  const int ld = m + 1;
  if (G::tid() < 32) {
    const int lane = G::tid();
    PHASE_BEGIN();
    {
      // Right-looking factorisation.  Lane i keeps the not-yet-eliminated part of row i in registers, shifted so that
      // g[0] is the current column.  One step: every lane publishes g[0] (= column j of the Schur complement) in
      // shared memory, reads the pivot and the 15 entries below it as broadcast loads (a chain of warp shuffles
      // each feeding one multiply-add costs ~60 cycles per pair on the in-order pipe), scales, and applies the
      // rank-1 update fused with the shift.  L[i][j] = s_ij / sqrt(s_jj);  s'_ik = s_ik - s_ij s_kj / s_jj.
      float* colbuf = pinv + MSVIT_MAX_EIG_BLOCK;  // 2 x 32 floats of scratch (reserved behind pinv, see Layout)
      const int row = lane < me ? lane : 0;
      float g[MB];
#pragma unroll
      for (int c = 0; c < MB; ++c) g[c] = (c < me && lane < me) ? Gm[row * ld + c] : 0.f;
      const float dorig = (lane < me) ? Gm[row * ld + row] : 0.f;
      float minpiv = 1.0f;
#pragma unroll 1
      for (int j = 0; j < me; ++j) {
        float* col = colbuf + (j & 1) * 64;
        col[lane] = g[0];
        col[32 + lane] = 0.f;  // reads past lane 31 (j + c > 31) see zeros
        __syncwarp();
        const float piv = col[j];
        const float gjj = __shfl_sync(0xffffffffu, dorig, j);
        float below[MB];
#pragma unroll
        for (int c = 1; c < MB; ++c) below[c] = col[j + c];
        const float rel = gjj > 0.f ? __fdividef(piv, gjj) : 0.f;
        const bool ok = rel > 1e-6f && piv > 0.f;
        minpiv = fminf(minpiv, ok ? rel : 1.0f);
        float inv = 0.f, ljj = 0.f;
        if (ok) {
          inv = rsqrtf(piv);
          inv = inv * (1.5f - 0.5f * piv * inv * inv);  // one Newton step: fp32-accurate 1/sqrt
          ljj = piv * inv;
        }
        const bool mine = lane > j && lane < me;
        const float l = lane == j ? ljj : (mine ? g[0] * inv : 0.f);  // L[lane][j]
        if (lane >= j && lane < me) Gm[lane * ld + j] = l;
        if (lane == j) pinv[j] = inv;
        const float f = mine ? g[0] * inv * inv : 0.f;  // s_ij / s_jj
#pragma unroll
        for (int c = 1; c < MB; ++c) g[c - 1] = fmaf(-f, below[c], g[c]);
        g[MB - 1] = 0.f;
      }
      if (lane == 0) misc[0] = minpiv;
    }
    __syncwarp();
    PHASE_END(PH_FACT);
    {
      // w_ij = pinv_i * (delta_ij - sum_{c<i} L[i][c] w_cj)
      float w[MB];
#pragma unroll
      for (int i = 0; i < MB; ++i) {
        float v = 0.f;
        if (i < me) {  // warp-uniform
          float s0 = (i == lane) ? 1.f : 0.f, s1 = 0.f;
#pragma unroll
          for (int c = 0; c + 1 < i; c += 2) {
            s0 = fmaf(-Gm[i * ld + c], w[c], s0);
            s1 = fmaf(-Gm[i * ld + c + 1], w[c + 1], s1);
          }
          if (i & 1) s0 = fmaf(-Gm[i * ld + i - 1], w[i - 1], s0);
          v = (s0 + s1) * pinv[i];
        }
        w[i] = v;
      }
      if (lane < m) {
#pragma unroll
        for (int i = 0; i < MB; ++i)
          if (i < m) W[i * ld + lane] = (i < me && lane <= i) ? w[i] : 0.f;
      }
    }
    PHASE_END(PH_INV);
  }
  G::sync();
  return misc[0];
}

// Cholesky G = L L^T of the leading me x me block (row stride m + 1), by warp 0 of the group, entirely in registers:
// lane i keeps row i of the Schur complement, step j broadcasts row j by warp shuffles (the j loop is fully
// unrolled, so every register index is static).  Output: LT[a * 16 + c] = L[c][a] for c >= a (row a of LT = column a
// of L, 16 floats apart so that forward substitutions read it with 128-bit broadcast loads), pinv[a] = 1 / L[a][a]
// (0 for a dropped, rank-deficient column and for a >= me).  Returns (to every thread) the smallest pivot of the
// unit-diagonal-scaled matrix.  MB = 16 only.
template <class G>
__device__ __forceinline__ float cholesky_lt16(const float* __restrict__ Gm, int m, int me, float* __restrict__ LT,
                                               float* __restrict__ pinv, float* __restrict__ misc) {
  const int ld = m + 1;
  if (G::tid() < 32) {
    const int lane = G::tid();
    const bool act = lane < me;
    const int row = act ? lane : 0;
    float g[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) g[c] = (act && c < me) ? Gm[row * ld + c] : 0.f;
    const float dorig = act ? Gm[row * ld + row] : 0.f;
    float minpiv = 1.0f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float piv = __shfl_sync(0xffffffffu, g[j], j);
      const float gjj = __shfl_sync(0xffffffffu, dorig, j);
      float rowj[16];
#pragma unroll
      for (int c = j + 1; c < 16; ++c) rowj[c] = __shfl_sync(0xffffffffu, g[c], j);
      const float rel = gjj > 0.f ? __fdividef(piv, gjj) : 0.f;
      const bool ok = j < me && rel > 1e-6f && piv > 0.f;
      if (j < me) minpiv = fminf(minpiv, ok ? rel : 1.0f);
      float inv = 0.f;
      if (ok) {
        inv = rsqrtf(piv);
        inv = inv * (1.5f - 0.5f * piv * inv * inv);  // one Newton step: fp32-accurate 1/sqrt
      }
      const float lij = lane >= j ? g[j] * inv : 0.f;     // L[lane][j]  (lane == j: piv * inv = sqrt(piv))
      if (lane < 16) LT[j * 16 + lane] = act ? lij : 0.f;
      if (lane == j) pinv[j] = inv;
      const float f = lane > j ? lij * inv : (lane == j ? 1.f : 0.f);  // s_ij / s_jj; row j itself is eliminated
#pragma unroll
      for (int c = j + 1; c < 16; ++c) g[c] = fmaf(-f, rowj[c], g[c]);
    }
    if (lane == 0) misc[0] = minpiv;
  }
  G::sync();
  return misc[0];
}

__device__ __forceinline__ float rsqrt_approx(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Same factorisation for a matrix that already sits in registers: lane i (< 16) holds row i in g[], dorig = G[i][i].
// Run by one converged warp; the caller orders the LT / pinv / misc[0] writes before other warps read them.
// Step j: lane j publishes its row of the Schur complement in shared memory (rowbuf: 2 x 16 floats, 16-byte aligned)
// and every lane reads it back with 128-bit broadcast loads -- measured 2.2 k cycles per factorisation against 4.3 k
// with warp shuffles (tools/microbench/chol_time.cu); the factor is kept in registers and stored after the loop.
// Nothing but the reciprocal root sits on the pivot-to-pivot path (the dropped-column test is a comparison,
// rsqrt.approx is accurate to 2 ulp, which the orthonormalisation does not notice).
__device__ __forceinline__ void cholesky_lt16_regs(float (&g)[16], float dorig, int me, float* __restrict__ LT,
                                                   float* __restrict__ pinv, float* __restrict__ misc,
                                                   float* rowbuf) {   // not __restrict__: other lanes write it
  const int lane = threadIdx.x & 31;
  const bool act = lane < me;
  float minpiv = 1.0f;
  float lcol[16], invs[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float* rb = rowbuf + (j & 1) * 16;
    if (lane == j) {
#pragma unroll
      for (int q4 = (j >> 2); q4 < 4; ++q4)
        *reinterpret_cast<float4*>(rb + 4 * q4) = make_float4(g[4 * q4], g[4 * q4 + 1], g[4 * q4 + 2], g[4 * q4 + 3]);
    }
    const float gjj = __shfl_sync(0xffffffffu, dorig, j);
    __syncwarp();
    float rowj[16];
#pragma unroll
    for (int q4 = (j >> 2); q4 < 4; ++q4) {
      const float4 v = *reinterpret_cast<const float4*>(rb + 4 * q4);
      rowj[4 * q4] = v.x; rowj[4 * q4 + 1] = v.y; rowj[4 * q4 + 2] = v.z; rowj[4 * q4 + 3] = v.w;
    }
    const float piv = rowj[j];
    const bool ok = j < me && piv > 1e-6f * gjj && piv > 0.f;
    float inv = rsqrt_approx(piv);
    inv = ok ? inv : 0.f;
    invs[j] = inv;
    if (j < me) minpiv = fminf(minpiv, ok ? __fdividef(piv, gjj) : 1.0f);   // off the critical path
    const float lij = (lane >= j && act) ? g[j] * inv : 0.f;     // L[lane][j]
    lcol[j] = lij;
    const float f = lane > j ? lij * inv : (lane == j ? 1.f : 0.f);  // s_ij / s_jj; row j itself is eliminated
#pragma unroll
    for (int c = j + 1; c < 16; ++c) g[c] = fmaf(-f, rowj[c], g[c]);
  }
  if (lane < 16) {
#pragma unroll
    for (int j = 0; j < 16; ++j) LT[j * 16 + lane] = lcol[j];
  }
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < 16; ++j) pinv[j] = invs[j];
    misc[0] = minpiv;
  }
}

// u = L^-1 y by forward substitution (right-looking: as soon as u[a] is known it is eliminated from all later
// entries), L^T and the reciprocal pivots read from shared memory as broadcast loads.  Dropped columns give u = 0.
__device__ __forceinline__ void forward_subst16(const float (&y)[16], const float* __restrict__ LT,
                                                const float* __restrict__ pinv, float (&u)[16]) {
  float acc[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = y[c];
  float pv[16];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 p4 = *reinterpret_cast<const float4*>(pinv + 4 * q);
    pv[4 * q] = p4.x; pv[4 * q + 1] = p4.y; pv[4 * q + 2] = p4.z; pv[4 * q + 3] = p4.w;
  }
#pragma unroll
  for (int a = 0; a < 16; ++a) {
    u[a] = acc[a] * pv[a];
#pragma unroll
    for (int q = (a + 1) >> 2; q < 4; ++q) {
      const float4 l4 = *reinterpret_cast<const float4*>(LT + a * 16 + 4 * q);
      const float l[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (4 * q + e > a) acc[4 * q + e] = fmaf(-l[e], u[a], acc[4 * q + e]);
    }
  }
}

// Jacobi rotation that annihilates the (p, q) entry: J = [[c, s], [-s, c]] on (p, q), H' = J^T H J.
__device__ __forceinline__ void jacobi_rot(float app, float aqq, float apq, float& c, float& s, bool& big) {
  c = 1.f;
  s = 0.f;
  const float pq = fabsf(app * aqq);
  const float scale = pq * rsqrtf(fmaxf(pq, 1e-37f));  // sqrt(|app aqq|), only used in the thresholds below
  const float aabs = fabsf(apq);
  if (aabs > 1e-30f && aabs > 1e-9f * scale) {
    const float delta = 0.5f * (aqq - app);
    const float r2 = fmaf(delta, delta, apq * apq);
    const float r = r2 * rsqrtf(r2);                    // r2 > 0 here; a 2-ulp root only perturbs the angle
    const float t = __fdividef(delta >= 0.f ? apq : -apq, fabsf(delta) + r);
    c = rsqrtf(fmaf(t, t, 1.f));
    c = c * (1.5f - 0.5f * fmaf(t, t, 1.f) * c * c);  // one Newton step: c^2 + s^2 = 1 to fp32 accuracy
    s = t * c;
  }
  big = big || aabs > fmaxf(1e-4f * scale, 3e-8f);
}

// Symmetric eigen-decomposition of the leading md x md block of H (row stride ld, md even): parallel-order
// two-sided Jacobi.  Every round the md/2 rotations of a round-robin pairing are computed once (one thread each),
// then every 2 x 2 block of J^T H J and every row pair of S J is updated in place by its own thread.  On exit H's
// diagonal holds the eigenvalues and Sm (same stride) the eigenvectors (columns).  A sweep whose rotations were
// all below 1e-4 (relative) ends the iteration: convergence is quadratic.
// WARP = true: run by one warp (barriers are __syncwarp); false: by the whole CTA.
template <bool WARP, class G>
__device__ __forceinline__ void jacobi_impl(float* __restrict__ H, float* __restrict__ Sm, int ld, int md,
                                            int max_sweeps, float* __restrict__ rot) {
  const int nthreads = WARP ? 32 : G::kThreads;
  const int tid = WARP ? (G::tid() & 31) : G::tid();
  auto sync = [&]() {
    if constexpr (WARP) __syncwarp();
    else G::sync();
  };
  for (int e = tid; e < md * md; e += nthreads) {
    const int a = e / md, b = e - a * md;
    Sm[a * ld + b] = a == b ? 1.f : 0.f;
    if (a < b) {
      const float v = 0.5f * (H[a * ld + b] + H[b * ld + a]);
      H[a * ld + b] = v;
      H[b * ld + a] = v;
    }
  }
  sync();
  const int half = md >> 1;
  const int nb = half * half, ns = md * half;
  // item -> (row pair / row, column pair): shifts when md / 2 is a power of two (the usual 16 and 8), else divisions
  const bool pow2 = (half & (half - 1)) == 0;
  const int hshift = 31 - __clz(half);
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    bool big = false;
    for (int r = 0; r < md - 1; ++r) {
      if (tid < half) {
        const int tq = tid;
        int p, q;
        if (tq == 0) { p = r; q = md - 1; }
        else {
          p = r + tq; if (p >= md - 1) p -= md - 1;
          q = r - tq; if (q < 0) q += md - 1;
        }
        if (p > q) { const int x = p; p = q; q = x; }
        float c, s;
        jacobi_rot(H[p * ld + p], H[q * ld + q], H[p * ld + q], c, s, big);
        *reinterpret_cast<float4*>(rot + 4 * tq) = make_float4(c, s, __int_as_float(p), __int_as_float(q));
      }
      sync();
      for (int item = tid; item < nb + ns; item += nthreads) {
        if (item < nb) {
          // H[P1][P2] <- J1^T H[P1][P2] J2
          const int t1 = pow2 ? item >> hshift : item / half, t2 = item - t1 * half;
          const float4 r1 = *reinterpret_cast<const float4*>(rot + 4 * t1);
          const float4 r2 = *reinterpret_cast<const float4*>(rot + 4 * t2);
          const int p1 = __float_as_int(r1.z), q1 = __float_as_int(r1.w);
          const int p2 = __float_as_int(r2.z), q2 = __float_as_int(r2.w);
          const float hpp = H[p1 * ld + p2], hpq = H[p1 * ld + q2], hqp = H[q1 * ld + p2], hqq = H[q1 * ld + q2];
          const float rpp = r1.x * hpp - r1.y * hqp, rpq = r1.x * hpq - r1.y * hqq;
          const float rqp = r1.y * hpp + r1.x * hqp, rqq = r1.y * hpq + r1.x * hqq;
          float npp = r2.x * rpp - r2.y * rpq, npq = r2.y * rpp + r2.x * rpq;
          float nqp = r2.x * rqp - r2.y * rqq, nqq = r2.y * rqp + r2.x * rqq;
          if (t1 == t2) { npq = 0.f; nqp = 0.f; }
          H[p1 * ld + p2] = npp; H[p1 * ld + q2] = npq; H[q1 * ld + p2] = nqp; H[q1 * ld + q2] = nqq;
        } else {
          // S[:, P2] <- S[:, P2] J2
          const int e = item - nb;
          const int a = pow2 ? e >> hshift : e / half, t2 = e - a * half;
          const float4 r2 = *reinterpret_cast<const float4*>(rot + 4 * t2);
          const int p2 = __float_as_int(r2.z), q2 = __float_as_int(r2.w);
          const float sp = Sm[a * ld + p2], sq = Sm[a * ld + q2];
          Sm[a * ld + p2] = r2.x * sp - r2.y * sq;
          Sm[a * ld + q2] = r2.y * sp + r2.x * sq;
        }
      }
      sync();
    }
    if constexpr (WARP) {
      if (!__any_sync(0xffffffffu, big)) break;
    } else {
      if (!G::sync_or(big)) break;
    }
  }
}

// Blocks of at most 8 x 8 (the leading block of the final Rayleigh-Ritz step) by ONE warp with no rotation exchange:
// lane (a, t) = (lane / half, lane % half) owns the 2 x 2 block (pair a, pair t) of H (lanes below half^2) and the
// row-a, pair-t slice of S, and computes the two rotations it needs itself from H's diagonal blocks -- a round is
// load -> rotate -> __syncwarp -> store -> __syncwarp, about half the latency of the broadcast-through-shared form.
template <class G>
__device__ __forceinline__ void jacobi_warp8(float* H, float* Sm, int ld, int md, int max_sweeps) {
  const int lane = G::tid() & 31;
  for (int e = lane; e < md * md; e += 32) {
    const int a = e / md, b = e - a * md;
    Sm[a * ld + b] = a == b ? 1.f : 0.f;
    if (a < b) {
      const float v = 0.5f * (H[a * ld + b] + H[b * ld + a]);
      H[a * ld + b] = v;
      H[b * ld + a] = v;
    }
  }
  __syncwarp();
  const int half = md >> 1;
  const int nb = half * half, ns = md * half;   // <= 16, <= 32
  const int ta = lane / half, tb = lane - ta * half;
  auto pair_of = [&](int t, int r, int& p, int& q) {
    if (t == 0) { p = r; q = md - 1; }
    else {
      p = r + t; if (p >= md - 1) p -= md - 1;
      q = r - t; if (q < 0) q += md - 1;
    }
    if (p > q) { const int x = p; p = q; q = x; }
  };
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    bool big = false;
    for (int r = 0; r < md - 1; ++r) {
      // every lane runs both rotations and both updates on valid addresses (lanes past the item counts repeat an
      // earlier item and drop the result): no divergence, the two dependent chains overlap
      const int tr = ta % half;                       // row pair of the 2 x 2 block
      const int sa = lane < ns ? ta : 0;              // row of the S slice
      int p1, q1, p2, q2;
      pair_of(tb, r, p2, q2);
      pair_of(tr, r, p1, q1);
      float c1, s1, c2, s2;
      bool unused = false;
      jacobi_rot(H[p2 * ld + p2], H[q2 * ld + q2], H[p2 * ld + q2], c2, s2, big);
      jacobi_rot(H[p1 * ld + p1], H[q1 * ld + q1], H[p1 * ld + q1], c1, s1, unused);
      const float sp = Sm[sa * ld + p2], sq = Sm[sa * ld + q2];
      const float nsp = c2 * sp - s2 * sq, nsq = s2 * sp + c2 * sq;
      const float hpp = H[p1 * ld + p2], hpq = H[p1 * ld + q2], hqp = H[q1 * ld + p2], hqq = H[q1 * ld + q2];
      const float rpp = c1 * hpp - s1 * hqp, rpq = c1 * hpq - s1 * hqq;
      const float rqp = s1 * hpp + c1 * hqp, rqq = s1 * hpq + c1 * hqq;
      float npp = c2 * rpp - s2 * rpq, npq = s2 * rpp + c2 * rpq;
      float nqp = c2 * rqp - s2 * rqq, nqq = s2 * rqp + c2 * rqq;
      if (tr == tb) { npq = 0.f; nqp = 0.f; }
      __syncwarp();
      if (lane < ns) { Sm[ta * ld + p2] = nsp; Sm[ta * ld + q2] = nsq; }
      if (lane < nb) {
        H[p1 * ld + p2] = npp; H[p1 * ld + q2] = npq; H[q1 * ld + p2] = nqp; H[q1 * ld + q2] = nqq;
      }
      __syncwarp();
    }
    if (!__any_sync(0xffffffffu, big)) break;
  }
}

// Small blocks are diagonalised by warp 0 alone (no CTA barrier per round, the other warps wait once).
template <class G>
__device__ __forceinline__ void jacobi(float* __restrict__ H, float* __restrict__ Sm, int ld, int md, int max_sweeps,
                                       float* __restrict__ rot) {
  if (md <= 8) {
    if (G::tid() < 32) jacobi_warp8<G>(H, Sm, ld, md, max_sweeps);
    G::sync();
  } else {
    jacobi_impl<false, G>(H, Sm, ld, md, max_sweeps, rot);
  }
}

}  // namespace eig
}  // namespace msvit
