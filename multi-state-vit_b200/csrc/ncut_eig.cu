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
//     every rr_every-th iteration:            H = U^T D Y, parallel-order Jacobi on one warp, rotate U and Y,
//                                             residuals |Abar v - theta v| for the k wanted pairs
//     U = Y L^-T,  L L^T = Y^T D Y            Cholesky QR in the D inner product (repeated if ill conditioned)
// Dot products are reduced with warp shuffles; nothing is atomically accumulated, so results are
// bit-reproducible run to run.
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace msvit {
namespace eig {

constexpr int kThreads = 256;
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
  int As, Us, Ys, Gs, Ss, dg, scratch, misc, total;
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

// In-place Cholesky of the leading me x me block of G (row stride m + 1) by warp 0; L is left in the lower
// triangle.  Returns (to every thread) the smallest pivot of the unit-diagonal-scaled matrix, i.e. a
// conditioning estimate that ignores column scaling.  Non-positive pivots zero the column (rank deficiency).
__device__ __forceinline__ float cholesky(float* __restrict__ G, int m, int me, float* __restrict__ misc) {
  const int ld = m + 1;
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    float minpiv = 1.0f;
    for (int j = 0; j < me; ++j) {
      // column j: L[i][j] = (G[i][j] - sum_{k<j} L[i][k] L[j][k]) / L[j][j]
      float gjj = G[j * ld + j];
      float s = gjj;
      for (int k2 = 0; k2 < j; ++k2) s = fmaf(-G[j * ld + k2], G[j * ld + k2], s);
      const float rel = gjj > 0.f ? s / gjj : 0.f;
      const bool ok = rel > 1e-6f && s > 0.f;
      minpiv = fminf(minpiv, ok ? rel : 1.0f);
      const float ljj = ok ? sqrtf(s) : 0.f;
      const float inv = ok ? 1.0f / ljj : 0.f;
      __syncwarp();
      for (int i = j + 1 + lane; i < me; i += 32) {
        float t = G[i * ld + j];
        for (int k2 = 0; k2 < j; ++k2) t = fmaf(-G[i * ld + k2], G[j * ld + k2], t);
        G[i * ld + j] = t * inv;
      }
      if (lane == 0) G[j * ld + j] = ljj;
      __syncwarp();
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

// Symmetric eigen-decomposition of H (m x m, row stride m + 1, m even) by warp 0: parallel-order
// two-sided Jacobi.  On exit H's diagonal holds the eigenvalues and Sm (same stride) the eigenvectors
// (columns).  Every round applies m/2 disjoint rotations; the 2x2 blocks of J^T H J are independent.
__device__ __forceinline__ void jacobi(float* __restrict__ H, float* __restrict__ Sm, int m,
                                       float* __restrict__ cs) {
  const int ld = m + 1;
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    const int half = m >> 1;
    for (int e = lane; e < m * m; e += 32) Sm[(e / m) * ld + (e % m)] = (e / m == e % m) ? 1.f : 0.f;
    // symmetrise
    for (int e = lane; e < m * m; e += 32) {
      const int a = e / m, b = e % m;
      if (a < b) {
        const float v = 0.5f * (H[a * ld + b] + H[b * ld + a]);
        H[a * ld + b] = v;
        H[b * ld + a] = v;
      }
    }
    __syncwarp();
    for (int sweep = 0; sweep < 12; ++sweep) {
      float off = 0.f, dia = 0.f;
      for (int r = 0; r < m - 1; ++r) {
        // rotation angles of this round's pairs
        for (int t = lane; t < half; t += 32) {
          int p, q;
          if (t == 0) { p = r; q = m - 1; }
          else { p = (r + t) % (m - 1); q = (r - t + (m - 1)) % (m - 1); }
          if (p > q) { const int x = p; p = q; q = x; }
          const float app = H[p * ld + p], aqq = H[q * ld + q], apq = H[p * ld + q];
          float c = 1.f, s = 0.f;
          off = fmaf(apq, apq, off);
          if (fabsf(apq) > 1e-30f && fabsf(apq) > 1e-9f * sqrtf(fabsf(app * aqq))) {
            const float tau = (aqq - app) / (2.f * apq);
            const float tt = (tau >= 0.f ? 1.f : -1.f) / (fabsf(tau) + sqrtf(1.f + tau * tau));
            c = rsqrtf(1.f + tt * tt);
            s = tt * c;
          }
          cs[4 * t + 0] = c;
          cs[4 * t + 1] = s;
          cs[4 * t + 2] = __int_as_float(p);
          cs[4 * t + 3] = __int_as_float(q);
        }
        __syncwarp();
        // H[P][P'] <- J_P^T H[P][P'] J_P'   with J = [[c, s], [-s, c]] on (p, q)
        for (int b2 = lane; b2 < half * half; b2 += 32) {
          const int t1 = b2 / half, t2 = b2 % half;
          const float c1 = cs[4 * t1], s1 = cs[4 * t1 + 1];
          const int p1 = __float_as_int(cs[4 * t1 + 2]), q1 = __float_as_int(cs[4 * t1 + 3]);
          const float c2 = cs[4 * t2], s2 = cs[4 * t2 + 1];
          const int p2 = __float_as_int(cs[4 * t2 + 2]), q2 = __float_as_int(cs[4 * t2 + 3]);
          const float hpp = H[p1 * ld + p2], hpq = H[p1 * ld + q2], hqp = H[q1 * ld + p2], hqq = H[q1 * ld + q2];
          // rows: J1^T
          const float rpp = c1 * hpp - s1 * hqp, rpq = c1 * hpq - s1 * hqq;
          const float rqp = s1 * hpp + c1 * hqp, rqq = s1 * hpq + c1 * hqq;
          // columns: J2
          float npp = c2 * rpp - s2 * rpq, npq = s2 * rpp + c2 * rpq;
          float nqp = c2 * rqp - s2 * rqq, nqq = s2 * rqp + c2 * rqq;
          if (t1 == t2) { npq = 0.f; nqp = 0.f; }
          H[p1 * ld + p2] = npp; H[p1 * ld + q2] = npq; H[q1 * ld + p2] = nqp; H[q1 * ld + q2] = nqq;
        }
        // S[:, P'] <- S[:, P'] J_P'
        for (int e = lane; e < m * half; e += 32) {
          const int a = e / half, t2 = e % half;
          const float c2 = cs[4 * t2], s2 = cs[4 * t2 + 1];
          const int p2 = __float_as_int(cs[4 * t2 + 2]), q2 = __float_as_int(cs[4 * t2 + 3]);
          const float sp = Sm[a * ld + p2], sq = Sm[a * ld + q2];
          Sm[a * ld + p2] = c2 * sp - s2 * sq;
          Sm[a * ld + q2] = s2 * sp + c2 * sq;
        }
        __syncwarp();
      }
      for (int a = lane; a < m; a += 32) dia = fmaf(H[a * ld + a], H[a * ld + a], dia);
      off = warp_sum(off);
      dia = warp_sum(dia);
      if (off <= 1e-14f * dia) break;
    }
  }
  __syncthreads();
}

template <int MB, bool RESIDENT>
__global__ void __launch_bounds__(kThreads, 1) ncut_eig_kernel(const Params P) {
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
  float* misc = smem + L.misc;           // [0] = scalar broadcast
  float* theta = misc + 8;               // [m] Ritz values in sorted order
  float* res = theta + MSVIT_MAX_EIG_BLOCK;   // [m] squared residuals
  int* order = reinterpret_cast<int*>(res + MSVIT_MAX_EIG_BLOCK);  // [m] sorted position -> Jacobi column
  float* cs = misc + 8 + 3 * MSVIT_MAX_EIG_BLOCK;  // 4 * m/2 rotation records

  const int m = P.m, k = P.k;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&load_bar, 1);
    fence_mbar_init();
  }
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
    cholesky(Gs, m, me, misc);
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
      if (last || (it % P.rr_every) == 0) {
        // ---- Rayleigh-Ritz on span(U)
        weighted_gram(Us, Ys, dg, n, m, Gs, scratch);
        jacobi(Gs, Ss, m, cs);
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
          res[a] = 0.f;
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
        // per-column residual: warp shuffle reduction, then a fixed-order sum over warps
#pragma unroll
        for (int c = 0; c < MB; ++c) {
          if (c < k) {
            const float v = warp_sum(rloc[c]);
            if (lane == 0) scratch[warp * MSVIT_MAX_EIG_BLOCK + c] = v;
          }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
          float worst = 0.f;
          const int kk = k < me ? k : me;
          for (int c = 0; c < kk; ++c) {
            float v = 0.f;
            for (int w = 0; w < kWarps; ++w) v += scratch[w * MSVIT_MAX_EIG_BLOCK + c];
            if (theta[c] >= P.lam_floor) worst = fmaxf(worst, v);
          }
          misc[1] = worst;
        }
        __syncthreads();
        if (last || misc[1] <= tol2) break;
      }
      // ---- U = orth_D(Y)
      weighted_gram(Ys, Ys, dg, n, m, Gs, scratch);
      const float piv = cholesky(Gs, m, me, misc);
      trisolve_rows<MB>(Us, Ys, n, m, me, Gs);
      if (piv < 0.05f) {
        weighted_gram(Us, Us, dg, n, m, Gs, scratch);
        cholesky(Gs, m, me, misc);
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
  if (!P.a_resident) smem = static_cast<size_t>(make_layout(N, block, false).total) * 4;
  if (smem > kMaxSmem) return MSVIT_ERR_SHAPE;
  const int grid = S < 4 * sm_count() ? S : 4 * sm_count();
  if (block <= 16) return launch<16>(P, grid, smem, stream);
  return launch<32>(P, grid, smem, stream);
}
