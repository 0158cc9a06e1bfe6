// Batched top-k eigensolver of the normalised affinity  Abar = D^-1/2 A D^-1/2  (one CTA per segment).
//
// Reference math: sandbox/test.py:114-118 (exact eigh of I - Abar, leading k); the reference's own path
// (ncut_pytorch NCUT.fit_transform, call site model/clustering/modeling_spectral.py:86) uses a randomised
// low-rank solver.  Here: block subspace (orthogonal) iteration with Rayleigh-Ritz, run to a residual
// tolerance, deterministic start.
//
// The iteration runs in the "random walk" coordinates u = D^-1/2 v:
//     Abar v = lam v   <=>   D^-1 A u = lam u,      v-orthonormal  <=>  u^T D u = I
// so the affinity is used exactly as stored (no scaled copy).  The n x m blocks U and Y live in shared
// memory TRANSPOSED (row c = column c of the block, row stride ldt with ldt % 32 == 16, so that the
// 128-bit fragment loads below are bank-conflict free); the affinity block itself is streamed from
// global memory / L2 once per iteration, which keeps the CTA small enough (~37 KB, 128 threads at
// N = 196) for several segments to be in flight per SM: the serial m x m steps of one segment
// (Cholesky, Jacobi) overlap the tensor-core products of the others.
//
// Per iteration:
//     Y = D^-1 (A U)             warp-level tensor-core MMA (mma.sync m16n8k8, TF32 operands, fp32 accumulate)
//                                with the 3-product split  a*b ~ a_hi*b_hi + a_hi*b_lo + a_lo*b_hi  (a_hi = a
//                                truncated to TF32, a_lo = a - a_hi exactly), i.e. fp32-level accuracy.
//                                A fragments are 128-bit loads of 4 consecutive columns per lane: the k index
//                                inside a 16-column block is permuted identically for both operands.
//     G = Y^T D Y,  H = U^T D Y  same MMA scheme, one warp per 16 x 8 output tile
//     trigger                    column residuals |y_j - U h_j|_D of the k leading columns plus their coupling
//                                to the trailing ones (ordered iteration): fires the Rayleigh-Ritz step
//     Rayleigh-Ritz (on trigger / every rr_every-th iteration): block-wide parallel-order Jacobi on H (one
//                                matrix entry per thread, one barrier per round), rotate U and Y, true
//                                residuals |Abar v - theta v| of the k wanted pairs
//     U = Y L^-T,  L L^T = G     Cholesky QR in the D inner product (register Cholesky on one warp, rows
//                                exchanged by shuffles; repeated if ill conditioned)
// Nothing is atomically accumulated and every reduction has a fixed order: results are bit-reproducible.
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace msvit {
namespace eig {

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
  int max_iter, rr_every;
  float tol;
  float lam_floor;  // wanted pairs whose Ritz value is below this are exempt from the residual test
  int fast_iters;   // the first fast_iters products use a single TF32 pass (far from convergence)
};

struct Layout {
  // offsets in floats from the dynamic shared memory base
  int Ut, Yt, dg, dinv, Gs, Hs, Ss, pinv, misc, colred, jac, ptab, total;
  int ldt;
};

__host__ __device__ inline int ldt_of(int N) {
  const int np = round_up(N, 16);
  return (np % 32 == 16) ? np : np + 16;
}

// rows: 16 or 32 rows of the transposed blocks (>= m, multiple of 16)
__host__ __device__ inline Layout make_layout(int N, int m, int rows, int nwarps) {
  Layout L;
  L.ldt = ldt_of(N);
  int o = 0;
  L.Ut = o;      o += rows * L.ldt;
  L.Yt = o;      o += rows * L.ldt;
  L.dg = o;      o += round_up(N, 16);
  L.dinv = o;    o += round_up(N, 16);
  L.Gs = o;      o += round_up(m * (m + 1), 4);
  L.Hs = o;      o += round_up(m * (m + 1), 4);
  L.Ss = o;      o += round_up(m * (m + 1), 4);
  L.pinv = o;    o += MSVIT_MAX_EIG_BLOCK;
  L.misc = o;    o += 8 + 3 * MSVIT_MAX_EIG_BLOCK;
  L.colred = o;  o += nwarps * MSVIT_MAX_EIG_BLOCK;
  L.jac = o;     o += 4 * m * m;                      // Jacobi ping-pong: H[2][m*m], S[2][m*m]
  L.ptab = o;    o += (m * m + 3) / 4;                // round-robin partner table, (m-1) x m bytes
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
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// x = hi + lo exactly: hi = x truncated to TF32 (what the tensor core reads from x's bits), lo = the remainder.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void split4(const float4& v, uint32_t (&hi)[4], uint32_t (&lo)[4]) {
  split_tf32(v.x, hi[0], lo[0]);
  split_tf32(v.y, hi[1], lo[1]);
  split_tf32(v.z, hi[2], lo[2]);
  split_tf32(v.w, hi[3], lo[3]);
}

// One 16-column block of  acc += [e; f] * b^T : e, f = 4 consecutive columns of rows g, g+8 of the left operand,
// b = the same 4 columns of row g of the (transposed) right operand.  Inside the block the MMA's k index is
// permuted (lane t covers columns 4t..4t+3 as k = t, t+4 of two k-steps), identically for both operands.
template <bool FULL>
__device__ __forceinline__ void mma_block(float (&hi)[4], float (&lo)[4], const uint32_t (&eh)[4],
                                          const uint32_t (&el)[4], const uint32_t (&fh)[4], const uint32_t (&fl)[4],
                                          const uint32_t (&bh)[4], const uint32_t (&bl)[4]) {
  if constexpr (FULL) {
    mma_tf32(lo, el[0], fl[0], el[1], fl[1], bh[0], bh[1]);
    mma_tf32(lo, eh[0], fh[0], eh[1], fh[1], bl[0], bl[1]);
    mma_tf32(lo, el[2], fl[2], el[3], fl[3], bh[2], bh[3]);
    mma_tf32(lo, eh[2], fh[2], eh[3], fh[3], bl[2], bl[3]);
  }
  mma_tf32(hi, eh[0], fh[0], eh[1], fh[1], bh[0], bh[1]);
  mma_tf32(hi, eh[2], fh[2], eh[3], fh[3], bh[2], bh[3]);
}

// Y^T[c][i] = dinv[i] * sum_j A[i][j] U^T[c][j].  One warp per 16 rows of A; the affinity block is read straight
// from global memory (L2), two 16-column blocks in flight per lane.
template <int NT, bool FULL>
__device__ __forceinline__ void matvec(const float* __restrict__ Ag, int lda, int n, const float* __restrict__ Ut,
                                       float* __restrict__ Yt, const float* __restrict__ dinv, int ldt, int nwarps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int KB = (n + 15) >> 4;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int tile = warp; tile < KB; tile += nwarps) {
    const int r0 = 16 * tile + g, r1 = r0 + 8;
    const bool v0 = r0 < n, v1 = r1 < n;
    const float* p0 = Ag + static_cast<size_t>(v0 ? r0 : 0) * lda + 4 * t;
    const float* p1 = Ag + static_cast<size_t>(v1 ? r1 : 0) * lda + 4 * t;
    const float* up = Ut + g * ldt + 4 * t;
    float hi[NT][4], lo[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int q = 0; q < 4; ++q) hi[nt][q] = lo[nt][q] = 0.f;
    const int cmax = lda - 4 * t;  // column block kb is in range iff 16 * kb < cmax
    float4 e0 = (v0 && 0 < cmax) ? __ldcg(reinterpret_cast<const float4*>(p0)) : zero4;
    float4 f0 = (v1 && 0 < cmax) ? __ldcg(reinterpret_cast<const float4*>(p1)) : zero4;
    float4 e1 = (v0 && 16 < cmax) ? __ldcg(reinterpret_cast<const float4*>(p0 + 16)) : zero4;
    float4 f1 = (v1 && 16 < cmax) ? __ldcg(reinterpret_cast<const float4*>(p1 + 16)) : zero4;
    for (int kb = 0; kb < KB; kb += 2) {
      const float4 ce0 = e0, cf0 = f0, ce1 = e1, cf1 = f1;
      const int c2 = 16 * (kb + 2), c3 = c2 + 16;
      e0 = (v0 && c2 < cmax) ? __ldcg(reinterpret_cast<const float4*>(p0 + c2)) : zero4;
      f0 = (v1 && c2 < cmax) ? __ldcg(reinterpret_cast<const float4*>(p1 + c2)) : zero4;
      e1 = (v0 && c3 < cmax) ? __ldcg(reinterpret_cast<const float4*>(p0 + c3)) : zero4;
      f1 = (v1 && c3 < cmax) ? __ldcg(reinterpret_cast<const float4*>(p1 + c3)) : zero4;
      {
        uint32_t eh[4], el[4], fh[4], fl[4];
        split4(ce0, eh, el);
        split4(cf0, fh, fl);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const float4 u = *reinterpret_cast<const float4*>(up + 8 * nt * ldt + 16 * kb);
          uint32_t uh[4], ul[4];
          split4(u, uh, ul);
          mma_block<FULL>(hi[nt], lo[nt], eh, el, fh, fl, uh, ul);
        }
      }
      if (kb + 1 < KB) {
        uint32_t eh[4], el[4], fh[4], fl[4];
        split4(ce1, eh, el);
        split4(cf1, fh, fl);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const float4 u = *reinterpret_cast<const float4*>(up + 8 * nt * ldt + 16 * (kb + 1));
          uint32_t uh[4], ul[4];
          split4(u, uh, ul);
          mma_block<FULL>(hi[nt], lo[nt], eh, el, fh, fl, uh, ul);
        }
      }
    }
    const float d0 = v0 ? dinv[r0] : 0.f, d1 = v1 ? dinv[r1] : 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      float* y = Yt + (8 * nt + 2 * t) * ldt;
      y[r0] = (hi[nt][0] + lo[nt][0]) * d0;
      y[ldt + r0] = (hi[nt][1] + lo[nt][1]) * d0;
      y[r1] = (hi[nt][2] + lo[nt][2]) * d1;
      y[ldt + r1] = (hi[nt][3] + lo[nt][3]) * d1;
    }
  }
}

// G[a][c] = sum_i dg[i] Q^T[a][i] Q^T[c][i]   and   H[a][c] = sum_i dg[i] P^T[a][i] Q^T[c][i]   (m x m, row
// stride m + 1).  One warp per 16 x 8 output tile, the whole token range per warp: no cross-warp reduction.
__device__ __forceinline__ void weighted_grams(const float* __restrict__ Pt, const float* __restrict__ Qt,
                                               const float* __restrict__ dg, int n, int m, int ldt,
                                               float* __restrict__ Gs, float* __restrict__ Hs, bool want_h,
                                               int nwarps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int KB = (n + 15) >> 4;
  const int MT = (m + 15) >> 4, NTg = (m + 7) >> 3;
  const int per = MT * NTg;
  const int ntiles = want_h ? 2 * per : per;
  const int ld = m + 1;
  for (int tile = warp; tile < ntiles; tile += nwarps) {
    const int mat = tile / per, rem = tile - mat * per;
    const int mt = rem / NTg, nt = rem - mt * NTg;
    const float* at = (mat == 0 ? Qt : Pt) + (16 * mt + g) * ldt + 4 * t;
    const float* bt = Qt + (8 * nt + g) * ldt + 4 * t;
    const float* dp = dg + 4 * t;
    float hi[4] = {0.f, 0.f, 0.f, 0.f}, lo[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
    for (int kb = 0; kb < KB; ++kb) {
      const float4 a0 = *reinterpret_cast<const float4*>(at + 16 * kb);
      const float4 a1 = *reinterpret_cast<const float4*>(at + 8 * ldt + 16 * kb);
      float4 b = *reinterpret_cast<const float4*>(bt + 16 * kb);
      const float4 d = *reinterpret_cast<const float4*>(dp + 16 * kb);
      b.x *= d.x; b.y *= d.y; b.z *= d.z; b.w *= d.w;
      uint32_t eh[4], el[4], fh[4], fl[4], bh[4], bl[4];
      split4(a0, eh, el);
      split4(a1, fh, fl);
      split4(b, bh, bl);
      mma_block<true>(hi, lo, eh, el, fh, fl, bh, bl);
    }
    float* out = mat == 0 ? Gs : Hs;
    const int r0 = 16 * mt + g, r1 = r0 + 8, c0 = 8 * nt + 2 * t, c1 = c0 + 1;
    if (r0 < m && c0 < m) out[r0 * ld + c0] = hi[0] + lo[0];
    if (r0 < m && c1 < m) out[r0 * ld + c1] = hi[1] + lo[1];
    if (r1 < m && c0 < m) out[r1 * ld + c0] = hi[2] + lo[2];
    if (r1 < m && c1 < m) out[r1 * ld + c1] = hi[3] + lo[3];
  }
  __syncthreads();
}

// ----------------------------------------------------------------------------- small dense pieces
// In-place Cholesky of the leading me x me block of G (row stride m + 1) by warp 0: lane i keeps row i in
// registers, row j is broadcast by shuffles (left-looking), L is written back to the lower triangle and the
// reciprocal pivots to pinv (0 for a dropped column).  Returns (to every thread) the smallest pivot of the
// unit-diagonal-scaled matrix, i.e. a conditioning estimate that ignores column scaling.
template <int MB>
__device__ __forceinline__ float cholesky(float* __restrict__ G, int m, int me, float* __restrict__ pinv,
                                          float* __restrict__ misc) {
  const int ld = m + 1;
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    const int row = lane < me ? lane : 0;
    float g[MB];
#pragma unroll
    for (int c = 0; c < MB; ++c) g[c] = (c < me) ? G[row * ld + c] : 0.f;
    float minpiv = 1.0f;
#pragma unroll
    for (int j = 0; j < MB; ++j) {
      if (j < me) {  // warp-uniform
        // s_i = G[i][j] - sum_{c<j} L[i][c] L[j][c]; lane j's value is the pivot
        float s0 = g[j], s1 = 0.f;
#pragma unroll
        for (int c = 0; c + 1 < j; c += 2) {
          s0 = fmaf(-g[c], __shfl_sync(0xffffffffu, g[c], j), s0);
          s1 = fmaf(-g[c + 1], __shfl_sync(0xffffffffu, g[c + 1], j), s1);
        }
        if (j & 1) s0 = fmaf(-g[j - 1], __shfl_sync(0xffffffffu, g[j - 1], j), s0);
        const float s = s0 + s1;
        const float piv = __shfl_sync(0xffffffffu, s, j);
        const float gjj = __shfl_sync(0xffffffffu, g[j], j);
        const float rel = gjj > 0.f ? piv / gjj : 0.f;
        const bool ok = rel > 1e-6f && piv > 0.f;
        minpiv = fminf(minpiv, ok ? rel : 1.0f);
        const float ljj = ok ? sqrtf(piv) : 0.f;
        const float inv = ok ? 1.0f / ljj : 0.f;
        g[j] = lane == j ? ljj : (lane > j ? s * inv : 0.f);
        if (lane == j) pinv[j] = inv;
      }
    }
    if (lane < me) {
#pragma unroll
      for (int c = 0; c < MB; ++c)
        if (c <= lane && c < me) G[lane * ld + c] = g[c];
    }
    if (lane == 0) misc[0] = minpiv;
  }
  __syncthreads();
  return misc[0];
}

// X^T <- (Y L^-T)^T token by token (forward substitution), L from `cholesky`.  Columns with a zero pivot and
// the pad tokens n .. npad-1 become 0.
template <int MB>
__device__ __forceinline__ void trisolve(float* __restrict__ Xt, const float* __restrict__ Yt, int n, int npad,
                                         int me, int ldt, const float* __restrict__ L, int ld,
                                         const float* __restrict__ pinv) {
  for (int i = threadIdx.x; i < npad; i += blockDim.x) {
    float y[MB];
#pragma unroll
    for (int c = 0; c < MB; ++c) y[c] = (c < me && i < n) ? Yt[c * ldt + i] : 0.f;
#pragma unroll
    for (int c = 0; c < MB; ++c) {
      if (c < me) {
        float s0 = y[c], s1 = 0.f;
#pragma unroll
        for (int a = 0; a + 1 < c; a += 2) {
          s0 = fmaf(-y[a], L[c * ld + a], s0);
          s1 = fmaf(-y[a + 1], L[c * ld + a + 1], s1);
        }
        if (c & 1) s0 = fmaf(-y[c - 1], L[c * ld + c - 1], s0);
        y[c] = (s0 + s1) * pinv[c];
      }
    }
#pragma unroll
    for (int c = 0; c < MB; ++c)
      if (c < me) Xt[c * ldt + i] = y[c];
  }
  __syncthreads();
}

// Round-robin tournament: partner of player i in round r (m even players, m - 1 rounds).
__device__ __forceinline__ int rr_partner(int i, int r, int m) {
  if (i == m - 1) return r;
  int j = 2 * r - i;
  if (j < 0) j += m - 1;
  if (j >= m - 1) j -= m - 1;
  return j == i ? m - 1 : j;
}

// Jacobi rotation that annihilates the (p, q) entry: J = [[c, s], [-s, c]] on (p, q), H' = J^T H J.
__device__ __forceinline__ void jacobi_rot(float app, float aqq, float apq, float& c, float& s, bool& big) {
  c = 1.f;
  s = 0.f;
  const float scale = sqrtf(fabsf(app * aqq));
  const float aabs = fabsf(apq);
  if (aabs > 1e-30f && aabs > 1e-9f * scale) {
    const float delta = 0.5f * (aqq - app);
    const float r = sqrtf(fmaf(delta, delta, apq * apq));
    const float t = (delta >= 0.f ? apq : -apq) / (fabsf(delta) + r);
    c = rsqrtf(fmaf(t, t, 1.f));
    c = c * (1.5f - 0.5f * fmaf(t, t, 1.f) * c * c);  // one Newton step: c^2 + s^2 = 1 to fp32 accuracy
    s = t * c;
  }
  big = big || aabs > fmaxf(1e-4f * scale, 3e-8f);
}

// Symmetric eigen-decomposition of H (m x m, row stride m + 1, m even) by the whole CTA: parallel-order
// two-sided Jacobi with one matrix entry per thread and one barrier per round (ping-pong buffers).
// On exit H's diagonal holds the eigenvalues and Sm (same stride) the eigenvectors (columns).
// A sweep whose rotations were all below 1e-4 (relative) ends the iteration: convergence is quadratic.
__device__ __forceinline__ void jacobi(float* __restrict__ H, float* __restrict__ Sm, int m, float* __restrict__ jac,
                                       const uint8_t* __restrict__ ptab) {
  const int ld = m + 1, mm = m * m;
  const int nthreads = blockDim.x;
  float* JH = jac;           // [2][mm]
  float* JS = jac + 2 * mm;  // [2][mm]
  for (int e = threadIdx.x; e < mm; e += nthreads) {
    const int a = e / m, b = e - a * m;
    JH[e] = 0.5f * (H[a * ld + b] + H[b * ld + a]);
    JS[e] = a == b ? 1.f : 0.f;
  }
  __syncthreads();
  int cur = 0;
  for (int sweep = 0; sweep < 10; ++sweep) {
    bool big = false;
    for (int r = 0; r < m - 1; ++r) {
      const float* __restrict__ hin = JH + cur * mm;
      const float* __restrict__ sin_ = JS + cur * mm;
      float* __restrict__ hout = JH + (cur ^ 1) * mm;
      float* __restrict__ sout = JS + (cur ^ 1) * mm;
      for (int e = threadIdx.x; e < mm; e += nthreads) {
        const int a = e / m, b = e - a * m;
        const int pa = ptab[r * m + a], pb = ptab[r * m + b];
        const int p1 = min(a, pa), q1 = max(a, pa), p2 = min(b, pb), q2 = max(b, pb);
        float c1, s1, c2, s2;
        bool dummy = false;
        jacobi_rot(hin[p1 * m + p1], hin[q1 * m + q1], hin[p1 * m + q1], c1, s1, dummy);
        jacobi_rot(hin[p2 * m + p2], hin[q2 * m + q2], hin[p2 * m + q2], c2, s2, big);
        const float g1 = a < pa ? -s1 : s1;  // coefficient of the partner row
        const float g2 = b < pb ? -s2 : s2;  // coefficient of the partner column
        const float v = c1 * fmaf(g2, hin[a * m + pb], c2 * hin[e]) + g1 * fmaf(g2, hin[pa * m + pb], c2 * hin[pa * m + b]);
        hout[e] = pa == b ? 0.f : v;
        sout[e] = fmaf(g2, sin_[a * m + pb], c2 * sin_[e]);
      }
      cur ^= 1;
      __syncthreads();
    }
    if (!__syncthreads_or(big ? 1 : 0)) break;
  }
  const float* __restrict__ hf = JH + cur * mm;
  const float* __restrict__ sf = JS + cur * mm;
  for (int e = threadIdx.x; e < mm; e += nthreads) {
    const int a = e / m, b = e - a * m;
    H[a * ld + b] = hf[e];
    Sm[a * ld + b] = sf[e];
  }
  __syncthreads();
}

// Per-column sums of per-thread partials: out[c] = sum over threads of rloc[c].  Warp shuffle reduction, then a
// fixed-order sum over warps (deterministic).
template <int MB>
__device__ __forceinline__ void reduce_columns(const float (&rloc)[MB], int ncols, float* __restrict__ colred,
                                               float* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
#pragma unroll
  for (int c = 0; c < MB; ++c) {
    if (c < ncols) {
      const float v = warp_sum(rloc[c]);
      if (lane == 0) colred[warp * MSVIT_MAX_EIG_BLOCK + c] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x < ncols) {
    float v = 0.f;
    for (int w = 0; w < nwarps; ++w) v += colred[w * MSVIT_MAX_EIG_BLOCK + threadIdx.x];
    out[threadIdx.x] = v;
  }
  __syncthreads();
}

// NT: 8-column tiles of the block (m <= 8 * NT).  THREADS: CTA size.
template <int NT, int THREADS>
__global__ void __launch_bounds__(THREADS, (512 / THREADS) > 0 ? (512 / THREADS) : 1) ncut_eig_kernel(const Params P) {
  constexpr int MB = 8 * NT;
  constexpr int ROWS = NT > 2 ? 32 : 16;
  constexpr int NWARPS = THREADS / 32;
  extern __shared__ __align__(16) float smem[];
  const Layout L = make_layout(P.N, P.m, ROWS, NWARPS);
  float* Ut = smem + L.Ut;
  float* Yt = smem + L.Yt;
  float* dg = smem + L.dg;
  float* dinv = smem + L.dinv;
  float* Gs = smem + L.Gs;
  float* Hs = smem + L.Hs;
  float* Ss = smem + L.Ss;
  float* pinv = smem + L.pinv;
  float* misc = smem + L.misc;           // [0] = scalar broadcast, [1] = worst residual, [2] = trigger flag
  float* theta = misc + 8;               // [m] Ritz values in sorted order
  float* res = theta + MSVIT_MAX_EIG_BLOCK;   // [m] squared residuals / column signs
  int* order = reinterpret_cast<int*>(res + MSVIT_MAX_EIG_BLOCK);  // [m] sorted position -> Jacobi column
  float* colred = smem + L.colred;
  float* jac = smem + L.jac;
  uint8_t* ptab = reinterpret_cast<uint8_t*>(smem + L.ptab);
  const int ldt = L.ldt;

  const int m = P.m, k = P.k;
  const int ld = m + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int e = threadIdx.x; e < (m - 1) * m; e += THREADS) ptab[e] = static_cast<uint8_t>(rr_partner(e % m, e / m, m));

  for (int s = blockIdx.x; s < P.S; s += gridDim.x) {
    const Seg g = seg_info(s, P.N, P.seg_off, P.a_off);
    const int n = g.n;
    float* __restrict__ Vout = P.V + static_cast<long long>(g.row0) * k;
    float* __restrict__ lout = P.lam + static_cast<long long>(s) * k;
    if (n <= 0) {
      for (int c = threadIdx.x; c < k; c += THREADS) lout[c] = 0.f;
      if (P.iters && threadIdx.x == 0) P.iters[s] = 0;
      continue;
    }
    const float* Ag = P.A + g.a0;
    const int lda = g.lda;
    const int me = n < m ? n : m;    // effective block width
    const int kk = k < me ? k : me;  // wanted pairs that exist
    const int npad = ((n + 15) >> 4) << 4;

    // ---- load: degree, start block (pad tokens and pad columns are zero)
    __syncthreads();  // the previous segment's readers of the shared arrays are done
    for (int i = threadIdx.x; i < npad; i += THREADS) {
      const float d = i < n ? P.deg[g.row0 + i] : 0.f;
      dg[i] = d;
      dinv[i] = d > 0.f ? 1.0f / d : 0.f;
    }
    for (int e = threadIdx.x; e < ROWS * npad; e += THREADS) {
      const int c = e / npad, i = e - c * npad;
      float v = 0.f;
      if (c < me && i < n) {
        if (n <= m) v = (i == c) ? 1.f : 0.f;
        else v = (c == 0) ? 1.f : hash_unit(i, c);
      }
      Ut[c * ldt + i] = v;
      Yt[c * ldt + i] = 0.f;
    }
    __syncthreads();

    // ---- D-orthonormalise the start block
    weighted_grams(Ut, Ut, dg, n, m, ldt, Gs, Hs, false, NWARPS);
    cholesky<MB>(Gs, m, me, pinv, misc);
    trisolve<MB>(Ut, Ut, n, npad, me, ldt, Gs, ld, pinv);

    int it = 0;
    const float tol2 = P.tol * P.tol;
    while (true) {
      ++it;
      if (it <= P.fast_iters) matvec<NT, false>(Ag, lda, n, Ut, Yt, dinv, ldt, NWARPS);
      else matvec<NT, true>(Ag, lda, n, Ut, Yt, dinv, ldt, NWARPS);
      __syncthreads();
      const bool last = it >= P.max_iter || n <= m;  // n <= m: span(U) is the whole space, one step is exact
      // ---- G = Y^T D Y, H = U^T D Y (U is D-orthonormal) and the trigger: for each wanted column j
      //        |y_j - U h_j|_D^2 + sum_{a >= kk} H[a][j]^2   =   residual of the Ritz problem on the leading columns
      weighted_grams(Ut, Yt, dg, n, m, ldt, Gs, Hs, true, NWARPS);
      bool do_rr = last || (it % P.rr_every) == 0;
      if (!do_rr && it >= 2 && it > P.fast_iters) {
        float rloc[MB];
#pragma unroll
        for (int c = 0; c < MB; ++c) rloc[c] = 0.f;
        for (int i = threadIdx.x; i < n; i += THREADS) {
          float u[MB];
#pragma unroll
          for (int a = 0; a < MB; ++a) u[a] = a < m ? Ut[a * ldt + i] : 0.f;
          const float d = dg[i];
#pragma unroll
          for (int c = 0; c < MB; ++c) {
            if (c < kk) {
              float r = Yt[c * ldt + i];
#pragma unroll
              for (int a = 0; a < MB; ++a)
                if (a < m) r = fmaf(-u[a], Hs[a * ld + c], r);
              rloc[c] = fmaf(d * r, r, rloc[c]);
            }
          }
        }
        reduce_columns<MB>(rloc, kk, colred, res);
        if (threadIdx.x == 0) {
          float worst = 0.f;
          for (int c = 0; c < kk; ++c) {
            float v = res[c];
            for (int a = kk; a < me; ++a) v = fmaf(Hs[a * ld + c], Hs[a * ld + c], v);
            if (Hs[c * ld + c] >= P.lam_floor) worst = fmaxf(worst, v);
          }
          misc[2] = worst <= tol2 ? 1.f : 0.f;
        }
        __syncthreads();
        do_rr = misc[2] != 0.f;
      }
      bool rotated = false;
      if (do_rr) {
        // ---- Rayleigh-Ritz on span(U)
        jacobi(Hs, Ss, m, jac, ptab);
        if (threadIdx.x < m) {
          const int a = threadIdx.x;
          const float ta = Hs[a * ld + a];
          int rank = 0;
          for (int b = 0; b < m; ++b) {
            const float tb = Hs[b * ld + b];
            rank += (tb > ta || (tb == ta && b < a)) ? 1 : 0;
          }
          order[rank] = a;
          theta[rank] = ta;
        }
        __syncthreads();
        // rotate U and Y into the Ritz basis, accumulate weighted residuals of the wanted pairs
        float rloc[MB];
#pragma unroll
        for (int c = 0; c < MB; ++c) rloc[c] = 0.f;
        for (int i = threadIdx.x; i < n; i += THREADS) {
          float u[MB], y[MB];
#pragma unroll
          for (int c = 0; c < MB; ++c) {
            u[c] = c < m ? Ut[c * ldt + i] : 0.f;
            y[c] = c < m ? Yt[c * ldt + i] : 0.f;
          }
          const float d = dg[i];
#pragma unroll
          for (int c = 0; c < MB; ++c) {
            if (c < m) {
              const int col = order[c];
              float nu = 0.f, ny = 0.f;
#pragma unroll
              for (int a = 0; a < MB; ++a) {
                if (a < m) {
                  const float sv = Ss[a * ld + col];
                  nu = fmaf(u[a], sv, nu);
                  ny = fmaf(y[a], sv, ny);
                }
              }
              Ut[c * ldt + i] = nu;
              Yt[c * ldt + i] = ny;
              const float rr = ny - theta[c] * nu;
              rloc[c] = fmaf(d * rr, rr, rloc[c]);
            }
          }
        }
        reduce_columns<MB>(rloc, kk, colred, res);
        if (threadIdx.x == 0) {
          float worst = 0.f;
          for (int c = 0; c < kk; ++c)
            if (theta[c] >= P.lam_floor) worst = fmaxf(worst, res[c]);
          misc[1] = worst;
        }
        __syncthreads();
        if (last || misc[1] <= tol2) break;
        rotated = true;
      }
      // ---- U = orth_D(Y)
      if (rotated) weighted_grams(Ut, Yt, dg, n, m, ldt, Gs, Hs, false, NWARPS);  // G of the rotated Y
      const float piv = cholesky<MB>(Gs, m, me, pinv, misc);
      trisolve<MB>(Ut, Yt, n, npad, me, ldt, Gs, ld, pinv);
      if (piv < 0.05f) {
        weighted_grams(Ut, Ut, dg, n, m, ldt, Gs, Hs, false, NWARPS);
        cholesky<MB>(Gs, m, me, pinv, misc);
        trisolve<MB>(Ut, Ut, n, npad, me, ldt, Gs, ld, pinv);
      }
    }

    // ---- output: v = sqrt(d) * u, canonical sign, eigenvalues
    for (int c = warp; c < k; c += NWARPS) {
      float best = -1.f;
      int bidx = 0x7fffffff;
      float bval = 0.f;
      if (c < me) {
        for (int i = lane; i < n; i += 32) {
          const float v = Ut[c * ldt + i] * sqrtf(dg[i]);
          const float av = fabsf(v);
          if (av > best) { best = av; bidx = i; bval = v; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ob = __shfl_xor_sync(0xffffffffu, best, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
          const float ov = __shfl_xor_sync(0xffffffffu, bval, o);
          if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; bval = ov; }
        }
      }
      if (lane == 0) {
        res[c] = (c < me && bval < 0.f) ? -1.f : 1.f;  // res[] now holds the column signs
        lout[c] = c < me ? theta[c] : 0.f;
      }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < n * k; e += THREADS) {
      const int i = e / k, c = e - i * k;
      Vout[e] = c < me ? Ut[c * ldt + i] * sqrtf(dg[i]) * res[c] : 0.f;
    }
    if (P.iters && threadIdx.x == 0) P.iters[s] = it;
  }
}

template <int NT, int THREADS>
static int launch(const Params& P, cudaStream_t stream) {
  constexpr int ROWS = NT > 2 ? 32 : 16;
  const size_t smem = static_cast<size_t>(make_layout(P.N, P.m, ROWS, THREADS / 32).total) * 4;
  const size_t kMaxSmem = 227 * 1024;
  if (smem > kMaxSmem) return MSVIT_ERR_SHAPE;
  cudaError_t e = cudaFuncSetAttribute(ncut_eig_kernel<NT, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return cuda_status(e);
  int per_sm = 1;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ncut_eig_kernel<NT, THREADS>, THREADS, smem);
  if (e != cudaSuccess) return cuda_status(e);
  if (per_sm < 1) per_sm = 1;
  // one CTA per segment while that is at most a few waves, else a persistent grid-stride loop
  const long long cap = 16LL * per_sm * sm_count();
  const int grid = static_cast<int>(P.S < cap ? P.S : cap);
  ncut_eig_kernel<NT, THREADS><<<grid, THREADS, smem, stream>>>(P);
  return cuda_status(cudaGetLastError());
}

template <int NT>
static int launch_threads(const Params& P, cudaStream_t stream) {
  if (P.N <= 256) return launch<NT, 128>(P, stream);
  if (P.N <= 512) return launch<NT, 256>(P, stream);
  return launch<NT, 512>(P, stream);
}

}  // namespace eig
}  // namespace msvit

extern "C" int msvit_ncut_eig(const float* A, const float* deg, float* V, float* lam, int32_t* iters,
                              int64_t total_rows, int S, int N, int k, int block, int max_iter, float tol,
                              float lam_floor, const int32_t* seg_off, const int64_t* a_off, msvit_stream_t stream_) {
  using namespace msvit;
  using namespace msvit::eig;
  if (!A || !deg || !V || !lam) return MSVIT_ERR_NULL;
  if (S < 0 || N <= 0 || k <= 0 || total_rows < 0 || max_iter <= 0 || !(tol > 0.f)) return MSVIT_ERR_SHAPE;
  if (block < k || block > MSVIT_MAX_EIG_BLOCK || (block & 3) != 0) return MSVIT_ERR_SHAPE;
  if (!seg_off && total_rows != static_cast<int64_t>(S) * N) return MSVIT_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(A) & 15) != 0) return MSVIT_ERR_ALIGN;
  if (S == 0 || total_rows == 0) return MSVIT_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);

  Params P;
  P.A = A; P.deg = deg; P.V = V; P.lam = lam; P.iters = iters;
  P.seg_off = seg_off; P.a_off = a_off;
  P.S = S; P.N = N; P.k = k; P.m = block;
  P.max_iter = max_iter; P.rr_every = 3; P.tol = tol; P.lam_floor = lam_floor;
  P.fast_iters = 0;
  if (block <= 16) return launch_threads<2>(P, stream);
  if (block <= 24) return launch_threads<3>(P, stream);
  return launch_threads<4>(P, stream);
}
