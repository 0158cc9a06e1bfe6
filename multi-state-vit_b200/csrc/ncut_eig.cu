// Batched top-k eigensolver of the normalised affinity  Abar = D^-1/2 A D^-1/2  (one CTA per segment).
//
// Reference math: sandbox/test.py:114-118 (exact eigh of I - Abar, leading k); the reference's own path
// (ncut_pytorch NCUT.fit_transform, call site model/clustering/modeling_spectral.py:86) uses a randomised
// low-rank solver.  Here: block subspace (orthogonal) iteration with Rayleigh-Ritz, run to a residual
// tolerance, deterministic start.
//
// The iteration runs in the "random walk" coordinates u = D^-1/2 v:
//     Abar v = lam v   <=>   D^-1 A u = lam u,      v-orthonormal  <=>  u^T D u = I
// so the affinity is used exactly as stored (no scaled copy) and only two n x m blocks (U, Y) live in
// shared memory next to A.  Per iteration:
//     Y = D^-1 (A U)                          4x4 register tiles, A read through its symmetric column
//     H = U^T D Y                             and the cheap trigger: column residuals |y_j - U h_j|_D of the k leading
//                                             columns plus their coupling to the trailing ones (ordered iteration)
//     when the trigger fires (or every rr_every-th iteration): Rayleigh-Ritz -- block-wide parallel-order Jacobi
//                                             on H (one matrix entry per thread, one barrier per round), rotate U
//                                             and Y, true residuals |Abar v - theta v| of the k wanted pairs
//     U = Y L^-T,  L L^T = Y^T D Y            Cholesky QR in the D inner product (register Cholesky on one warp,
//                                             rows exchanged by shuffles; repeated if ill conditioned)
// Dot products are reduced with warp shuffles; nothing is atomically accumulated, so results are
// bit-reproducible run to run.
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace msvit {
namespace eig {

#ifndef EIG_THREADS
#define EIG_THREADS 256
#endif
#ifndef EIG_MINBLOCKS
#define EIG_MINBLOCKS 1
#endif
constexpr int kThreads = EIG_THREADS;
constexpr int kWarps = kThreads / 32;
constexpr int kScratchFloats = 4096;  // 16 KB partial-sum scratch for the m x m Gram reductions

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
  int a_resident;  // 1: the affinity block is copied into shared memory once (TMA bulk copy)
};

struct Layout {
  // offsets in floats from the dynamic shared memory base
  int As, Us, Ys, Gs, Ss, dg, scratch, misc, jac, ptab, total;
};

__host__ __device__ inline Layout make_layout(int N, int m, bool resident) {
  Layout L;
  int o = 0;
  L.As = o;       o += resident ? N * lda_of(N) : 0;
  L.Us = o;       o += round_up(N, 4) * m;
  L.Ys = o;       o += round_up(N, 4) * m;
  L.Gs = o;       o += m * (m + 1);
  L.Ss = o;       o += m * (m + 1);
  L.dg = o;       o += round_up(N, 4);
  L.scratch = o;  o += kScratchFloats;
  L.misc = o;     o += 6 * MSVIT_MAX_EIG_BLOCK;
  L.jac = o;      o += 4 * m * m;                      // Jacobi ping-pong: H[2][m*m], S[2][m*m]
  L.ptab = o;     o += (m * m + 3) / 4;                // round-robin partner table, (m-1) x m bytes
  L.total = o;
  return L;
}

__device__ __forceinline__ float hash_unit(uint32_t i, uint32_t c) {
  uint32_t h = i * 0x9E3779B1u + c * 0x85EBCA77u + 0x165667B1u;
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
  return static_cast<float>(static_cast<int32_t>(h)) * (1.0f / 2147483648.0f);
}

// Y = D^-1 (A U).  A symmetric: the 4 output rows 4rg..4rg+3 read A[j][4rg..4rg+3] (one 16-byte load per j).
template <bool RESIDENT>
__device__ __forceinline__ void matvec(const float* __restrict__ A, int lda, int n, int m,
                                       const float* __restrict__ Us, float* __restrict__ Ys,
                                       const float* __restrict__ dg) {
  const int CG = m >> 2;
  const int tiles = ((n + 3) >> 2) * CG;
  for (int tile = threadIdx.x; tile < tiles; tile += kThreads) {
    const int rg = tile / CG, cg = tile - rg * CG;
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
    const float* ap = A + 4 * rg;
    const float* up = Us + 4 * cg;
#pragma unroll 4
    for (int j = 0; j < n; ++j) {
      float4 a;
      if constexpr (RESIDENT) a = *reinterpret_cast<const float4*>(ap + static_cast<size_t>(j) * lda);
      else a = __ldg(reinterpret_cast<const float4*>(ap + static_cast<size_t>(j) * lda));
      const float4 u = *reinterpret_cast<const float4*>(up + j * m);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float uv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(av[r], uv[c], acc[r][c]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int row = 4 * rg + r;
      if (row < n) {
        const float inv = 1.0f / dg[row];
        *reinterpret_cast<float4*>(Ys + row * m + 4 * cg) =
            make_float4(acc[r][0] * inv, acc[r][1] * inv, acc[r][2] * inv, acc[r][3] * inv);
      }
    }
  }
}

// Out[a][b] = sum_i dg[i] * P[i][a] * Q[i][b]   (m x m, row stride m + 1).
// 4x4 blocks of the result x row slices; partials go through `scratch` and are summed in a fixed order.
__device__ __forceinline__ void weighted_gram(const float* __restrict__ Ps, const float* __restrict__ Qs,
                                              const float* __restrict__ dg, int n, int m, float* __restrict__ Out,
                                              float* __restrict__ scratch) {
  const int CG = m >> 2;
  const int blocks = CG * CG;
  int slices = kThreads / blocks;
  if (slices < 1) slices = 1;
  if (slices * m * m > kScratchFloats) slices = kScratchFloats / (m * m);
  for (int w = threadIdx.x; w < blocks * slices; w += kThreads) {
    const int blk = w % blocks, sl = w / blocks;
    const int a0 = (blk / CG) * 4, b0 = (blk % CG) * 4;
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
    for (int i = sl; i < n; i += slices) {
      const float d = dg[i];
      const float4 p = *reinterpret_cast<const float4*>(Ps + i * m + a0);
      const float4 q = *reinterpret_cast<const float4*>(Qs + i * m + b0);
      const float pv[4] = {p.x * d, p.y * d, p.z * d, p.w * d};
      const float qv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(pv[r], qv[c], acc[r][c]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) scratch[sl * m * m + (a0 + r) * m + b0 + c] = acc[r][c];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < m * m; e += kThreads) {
    float s = 0.f;
    for (int sl = 0; sl < slices; ++sl) s += scratch[sl * m * m + e];
    Out[(e / m) * (m + 1) + (e % m)] = s;
  }
  __syncthreads();
}

// In-place Cholesky of the leading me x me block of G (row stride m + 1) by warp 0: lane i keeps row i in
// registers, row j is broadcast by shuffles (left-looking), L is written back to the lower triangle.
// Returns (to every thread) the smallest pivot of the unit-diagonal-scaled matrix, i.e. a conditioning
// estimate that ignores column scaling.  Non-positive pivots zero the column (rank deficiency).
template <int MB>
__device__ __forceinline__ float cholesky(float* __restrict__ G, int m, int me, float* __restrict__ misc) {
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

// X <- X L^-T  row by row (forward substitution), L from `cholesky`.  Columns with a zero pivot become 0.
template <int MB>
__device__ __forceinline__ void trisolve_rows(float* __restrict__ Xs, const float* __restrict__ Ys, int n, int m,
                                              int me, const float* __restrict__ L) {
  const int ld = m + 1;
  for (int i = threadIdx.x; i < n; i += kThreads) {
    float y[MB];
#pragma unroll
    for (int c = 0; c < MB; ++c) y[c] = c < me ? Ys[i * m + c] : 0.f;
#pragma unroll
    for (int c = 0; c < MB; ++c) {
      if (c < me) {
        float t = y[c];
#pragma unroll
        for (int k2 = 0; k2 < c; ++k2) t = fmaf(-y[k2], L[c * ld + k2], t);
        const float piv = L[c * ld + c];
        y[c] = piv > 0.f ? t / piv : 0.f;
      }
    }
#pragma unroll
    for (int c = 0; c < MB; ++c)
      if (c < m) Xs[i * m + c] = y[c];
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
  float* JH = jac;           // [2][mm]
  float* JS = jac + 2 * mm;  // [2][mm]
  for (int e = threadIdx.x; e < mm; e += kThreads) {
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
      for (int e = threadIdx.x; e < mm; e += kThreads) {
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
      if (r < m - 2) __syncthreads();
    }
    if (!__syncthreads_or(big ? 1 : 0)) break;
  }
  const float* __restrict__ hf = JH + cur * mm;
  const float* __restrict__ sf = JS + cur * mm;
  for (int e = threadIdx.x; e < mm; e += kThreads) {
    const int a = e / m, b = e - a * m;
    H[a * ld + b] = hf[e];
    Sm[a * ld + b] = sf[e];
  }
  __syncthreads();
}

// Residuals of the leading columns against span(U) or against their Ritz values, reduced per column:
// out[c] = sum_i dg[i] * r_ic^2.  Warp shuffle reduction, then a fixed-order sum over warps (deterministic).
template <int MB>
__device__ __forceinline__ void reduce_columns(const float (&rloc)[MB], int ncols, float* __restrict__ scratch,
                                               float* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < MB; ++c) {
    if (c < ncols) {
      const float v = warp_sum(rloc[c]);
      if (lane == 0) scratch[warp * MSVIT_MAX_EIG_BLOCK + c] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x < ncols) {
    float v = 0.f;
    for (int w = 0; w < kWarps; ++w) v += scratch[w * MSVIT_MAX_EIG_BLOCK + threadIdx.x];
    out[threadIdx.x] = v;
  }
  __syncthreads();
}

template <int MB, bool RESIDENT>
__global__ void __launch_bounds__(kThreads, EIG_MINBLOCKS) ncut_eig_kernel(const Params P) {
  extern __shared__ __align__(16) float smem[];
  __shared__ __align__(8) uint64_t load_bar;
  const Layout L = make_layout(P.N, P.m, RESIDENT);
  float* As = smem + L.As;
  float* Us = smem + L.Us;
  float* Ys = smem + L.Ys;
  float* Gs = smem + L.Gs;
  float* Ss = smem + L.Ss;
  float* dg = smem + L.dg;
  float* scratch = smem + L.scratch;
  float* misc = smem + L.misc;           // [0] = scalar broadcast, [1] = worst residual, [2] = trigger flag
  float* theta = misc + 8;               // [m] Ritz values in sorted order
  float* res = theta + MSVIT_MAX_EIG_BLOCK;   // [m] squared residuals / column signs
  int* order = reinterpret_cast<int*>(res + MSVIT_MAX_EIG_BLOCK);  // [m] sorted position -> Jacobi column
  float* jac = smem + L.jac;
  uint8_t* ptab = reinterpret_cast<uint8_t*>(smem + L.ptab);

  const int m = P.m, k = P.k;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&load_bar, 1);
    fence_mbar_init();
  }
  for (int e = threadIdx.x; e < (m - 1) * m; e += kThreads) ptab[e] = static_cast<uint8_t>(rr_partner(e % m, e / m, m));
  __syncthreads();
  uint32_t load_phase = 0;

  for (int s = blockIdx.x; s < P.S; s += gridDim.x) {
    const Seg g = seg_info(s, P.N, P.seg_off, P.a_off);
    const int n = g.n;
    float* __restrict__ Vout = P.V + static_cast<long long>(g.row0) * k;
    float* __restrict__ lout = P.lam + static_cast<long long>(s) * k;
    if (n <= 0) {
      for (int c = threadIdx.x; c < k; c += kThreads) lout[c] = 0.f;
      if (P.iters && threadIdx.x == 0) P.iters[s] = 0;
      continue;
    }
    const float* Ag = P.A + g.a0;
    const int lda = g.lda;
    const int me = n < m ? n : m;  // effective block width
    const int kk = k < me ? k : me;  // wanted pairs that exist

    // ---- load: affinity block (bulk async copy), degree, start block
    if constexpr (RESIDENT) {
      if (threadIdx.x == 0) {
        fence_proxy_async_smem();  // earlier generic-proxy reads of As are ordered before the async writes
        const uint32_t total = static_cast<uint32_t>(n) * lda * 4u;
        mbar_arrive_expect_tx(&load_bar, total);
        for (uint32_t off = 0; off < total; off += 32768u) {
          const uint32_t len = total - off < 32768u ? total - off : 32768u;
          bulk_load_1d(reinterpret_cast<uint8_t*>(As) + off, reinterpret_cast<const uint8_t*>(Ag) + off, len,
                       &load_bar);
        }
      }
    }
    for (int i = threadIdx.x; i < n; i += kThreads) dg[i] = P.deg[g.row0 + i];
    for (int e = threadIdx.x; e < n * m; e += kThreads) {
      const int i = e / m, c = e - i * m;
      float v;
      if (n <= m) v = (i == c) ? 1.f : 0.f;
      else v = (c == 0) ? 1.f : hash_unit(i, c);
      Us[e] = c < me ? v : 0.f;
    }
    __syncthreads();

    // ---- D-orthonormalise the start block
    weighted_gram(Us, Us, dg, n, m, Gs, scratch);
    cholesky<MB>(Gs, m, me, misc);
    trisolve_rows<MB>(Us, Us, n, m, me, Gs);

    if constexpr (RESIDENT) {
      mbar_wait(&load_bar, load_phase);
      load_phase ^= 1;
    }
    const float* Amat = RESIDENT ? As : Ag;

    int it = 0;
    const float tol2 = P.tol * P.tol;
    while (true) {
      ++it;
      matvec<RESIDENT>(Amat, lda, n, m, Us, Ys, dg);
      __syncthreads();
      const bool last = it >= P.max_iter || n <= m;  // n <= m: span(U) is the whole space, one step is exact
      // ---- H = U^T D Y (U is D-orthonormal) and the trigger: for each wanted column j
      //        |y_j - U h_j|_D^2 + sum_{a >= kk} H[a][j]^2   =   residual of the Ritz problem on the leading columns
      weighted_gram(Us, Ys, dg, n, m, Gs, scratch);
      bool do_rr = last || (it % P.rr_every) == 0;
      if (!do_rr && it >= 2) {
        float rloc[MB];
#pragma unroll
        for (int c = 0; c < MB; ++c) rloc[c] = 0.f;
        for (int i = threadIdx.x; i < n; i += kThreads) {
          float u[MB];
#pragma unroll
          for (int a = 0; a < MB; ++a) u[a] = a < m ? Us[i * m + a] : 0.f;
          const float d = dg[i];
#pragma unroll
          for (int c = 0; c < MB; ++c) {
            if (c < kk) {
              float r = Ys[i * m + c];
#pragma unroll
              for (int a = 0; a < MB; ++a)
                if (a < m) r = fmaf(-u[a], Gs[a * (m + 1) + c], r);
              rloc[c] = fmaf(d * r, r, rloc[c]);
            }
          }
        }
        reduce_columns<MB>(rloc, kk, scratch, res);
        if (threadIdx.x == 0) {
          float worst = 0.f;
          for (int c = 0; c < kk; ++c) {
            float v = res[c];
            for (int a = kk; a < me; ++a) v = fmaf(Gs[a * (m + 1) + c], Gs[a * (m + 1) + c], v);
            if (Gs[c * (m + 1) + c] >= P.lam_floor) worst = fmaxf(worst, v);
          }
          misc[2] = worst <= tol2 ? 1.f : 0.f;
        }
        __syncthreads();
        do_rr = misc[2] != 0.f;
      }
      if (do_rr) {
        // ---- Rayleigh-Ritz on span(U)
        jacobi(Gs, Ss, m, jac, ptab);
        if (threadIdx.x < m) {
          const int a = threadIdx.x;
          const float ta = Gs[a * (m + 1) + a];
          int rank = 0;
          for (int b = 0; b < m; ++b) {
            const float tb = Gs[b * (m + 1) + b];
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
        for (int i = threadIdx.x; i < n; i += kThreads) {
          float u[MB], y[MB];
#pragma unroll
          for (int c = 0; c < MB; ++c) {
            u[c] = c < m ? Us[i * m + c] : 0.f;
            y[c] = c < m ? Ys[i * m + c] : 0.f;
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
                  const float sv = Ss[a * (m + 1) + col];
                  nu = fmaf(u[a], sv, nu);
                  ny = fmaf(y[a], sv, ny);
                }
              }
              Us[i * m + c] = nu;
              Ys[i * m + c] = ny;
              const float rr = ny - theta[c] * nu;
              rloc[c] = fmaf(d * rr, rr, rloc[c]);
            }
          }
        }
        reduce_columns<MB>(rloc, kk, scratch, res);
        if (threadIdx.x == 0) {
          float worst = 0.f;
          for (int c = 0; c < kk; ++c)
            if (theta[c] >= P.lam_floor) worst = fmaxf(worst, res[c]);
          misc[1] = worst;
        }
        __syncthreads();
        if (last || misc[1] <= tol2) break;
      }
      // ---- U = orth_D(Y)
      weighted_gram(Ys, Ys, dg, n, m, Gs, scratch);
      const float piv = cholesky<MB>(Gs, m, me, misc);
      trisolve_rows<MB>(Us, Ys, n, m, me, Gs);
      if (piv < 0.05f) {
        weighted_gram(Us, Us, dg, n, m, Gs, scratch);
        cholesky<MB>(Gs, m, me, misc);
        trisolve_rows<MB>(Us, Us, n, m, me, Gs);
      }
    }

    // ---- output: v = sqrt(d) * u, canonical sign, eigenvalues
    for (int c = warp; c < k; c += kWarps) {
      float best = -1.f;
      int bidx = 0x7fffffff;
      float bval = 0.f;
      if (c < me) {
        for (int i = lane; i < n; i += 32) {
          const float v = Us[i * m + c] * sqrtf(dg[i]);
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
    for (int e = threadIdx.x; e < n * k; e += kThreads) {
      const int i = e / k, c = e - i * k;
      Vout[e] = c < me ? Us[i * m + c] * sqrtf(dg[i]) * res[c] : 0.f;
    }
    if (P.iters && threadIdx.x == 0) P.iters[s] = it;
    __syncthreads();
  }
}

template <int MB>
static int launch(const Params& P, int grid, size_t smem, cudaStream_t stream) {
  cudaError_t e;
  if (P.a_resident) {
    e = cudaFuncSetAttribute(ncut_eig_kernel<MB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(smem));
    if (e != cudaSuccess) return cuda_status(e);
    ncut_eig_kernel<MB, true><<<grid, kThreads, smem, stream>>>(P);
  } else {
    e = cudaFuncSetAttribute(ncut_eig_kernel<MB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(smem));
    if (e != cudaSuccess) return cuda_status(e);
    ncut_eig_kernel<MB, false><<<grid, kThreads, smem, stream>>>(P);
  }
  return cuda_status(cudaGetLastError());
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

  const size_t kMaxSmem = 227 * 1024 - 64;  // static mbarrier lives in the same budget
  Params P;
  P.A = A; P.deg = deg; P.V = V; P.lam = lam; P.iters = iters;
  P.seg_off = seg_off; P.a_off = a_off;
  P.S = S; P.N = N; P.k = k; P.m = block;
  P.max_iter = max_iter; P.rr_every = 3; P.tol = tol; P.lam_floor = lam_floor;
  size_t smem = static_cast<size_t>(make_layout(N, block, true).total) * 4;
  P.a_resident = smem <= kMaxSmem ? 1 : 0;
#ifdef EIG_FORCE_STREAM
  P.a_resident = 0;
#endif
  if (!P.a_resident) smem = static_cast<size_t>(make_layout(N, block, false).total) * 4;
  if (smem > kMaxSmem) return MSVIT_ERR_SHAPE;
  const int grid = S < 8 * sm_count() ? S : 8 * sm_count();
  if (block <= 16) return launch<16>(P, grid, smem, stream);
  return launch<32>(P, grid, smem, stream);
}
