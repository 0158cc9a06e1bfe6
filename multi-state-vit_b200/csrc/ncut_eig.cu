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
// memory TRANSPOSED (row c = column c of the block, row stride ldt with ldt % 32 == 16, so that 128-bit
// fragment loads are bank-conflict free), U additionally in MMA fragment order; the affinity block itself
// is streamed from global memory / L2 once per iteration, which keeps the CTA small (~50 KB, 128 threads at
// N = 196) so that several segments are in flight per SM: the serial m x m steps of one segment (Cholesky,
// Jacobi) overlap the tensor-core products of the others.
//
// All n-sized products run on the tensor cores (mma.sync m16n8k8, TF32 operands, fp32 accumulate) with the
// 3-product split  a*b ~ a_hi*b_hi + a_hi*b_lo + a_lo*b_hi  (a_hi = a truncated to TF32 -- what the tensor
// core reads from a's bits --, a_lo = a - a_hi exactly), i.e. fp32-level accuracy.  Per iteration:
//     Y^T = (U^T A) D^-1         U^T fragments are the 16 x 8 left operand (read from the fragment-order copy
//                                with 128-bit loads), 8 rows of the symmetric A are the right operand: each lane
//                                loads 4 consecutive columns (one 128-bit load, two k-steps); the k index inside
//                                a 16-column block is permuted identically for both operands.
//     G = Y^T D Y,  H = U^T D Y  one warp per 16 x 8 output tile
//     trigger                    column residuals |y_j - U h_j|_D of the k leading columns plus their coupling
//                                to the trailing ones (ordered iteration): fires the Rayleigh-Ritz step
//     Rayleigh-Ritz              parallel-order Jacobi (rotations computed once per round, 2 x 2 blocks updated
//                                in place), then [U; Y] <- W [U; Y] and the true residuals |Abar v - theta v|.
//                                Scheduled every rr_every-th (4th) iteration on the whole m x m H with a few sweeps
//                                (the span does not change, only its basis); when the trigger fires, on the
//                                leading block only, to full accuracy.
//     U^T = L^-1 Y^T, L L^T = G  Cholesky QR in the D inner product (register Cholesky and explicit L^-1 on
//                                one warp; repeated if ill conditioned)
// Nothing is atomically accumulated and every reduction has a fixed order: results are bit-reproducible.
#include "eig_core.cuh"

namespace msvit {
namespace eig {

// MT: 16-row tiles of the block (m <= 16 * MT).  THREADS: CTA size.
template <int MT, int THREADS>
__global__ void __launch_bounds__(THREADS, (128 * EIG_MINB / THREADS) > 0 ? (128 * EIG_MINB / THREADS) : 1)
    ncut_eig_kernel(const Params P) {
  constexpr int MB = 16 * MT;
  constexpr int NWARPS = THREADS / 32;
  using G = ThreadGroup<0, THREADS, 0>;
  constexpr int T = MT == 1 ? EIG_T : 2;
  extern __shared__ __align__(16) float smem[];
  const Layout L = make_layout(P.N, P.m, MT, NWARPS);
  float* Ut = smem + L.Ut;
  float* Yt = smem + L.Yt;
  float* Uf = smem + L.Uf;
  float* dg = smem + L.dg;
  float* dinv = smem + L.dinv;
  float* Gs = smem + L.Gs;
  float* Hs = smem + L.Hs;
  float* Ss = smem + L.Ss;
  float* Ws = smem + L.Ws;
  float* pinv = smem + L.pinv;
  float* misc = smem + L.misc;           // [0] = scalar broadcast, [1] = worst residual, [2] = trigger flag
  float* theta = misc + 8;               // [m] Ritz values in sorted order
  float* res = theta + MSVIT_MAX_EIG_BLOCK;   // [m] squared residuals / column signs
  int* order = reinterpret_cast<int*>(res + MSVIT_MAX_EIG_BLOCK);  // [m] sorted position -> Jacobi column
  float* colred = smem + L.colred;
  float* rot = smem + L.rot;
  const int ldt = L.ldt, rows = L.rows;

  const int m = P.m, k = P.k;
  const int ld = m + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

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
    const int kk = P.kconv < me ? P.kconv : me;  // pairs that must converge (and exist)
    const int npad = ((n + 15) >> 4) << 4;

    // ---- load: degree, start block (pad tokens and pad rows are zero)
    __syncthreads();  // the previous segment's readers of the shared arrays are done
    PHASE_BEGIN();
    for (int i = threadIdx.x; i < npad; i += THREADS) {
      const float d = i < n ? P.deg[g.row0 + i] : 0.f;
      dg[i] = d;
      dinv[i] = d > 0.f ? 1.0f / d : 0.f;
    }
    // one token per thread and step, all block rows in an inner loop (no index divisions); rows of the 16-row tiles
    // beyond `rows` exist only in the fragment-order copy
    for (int i = threadIdx.x; i < npad; i += THREADS) {
#pragma unroll 4
      for (int c = 0; c < MB; ++c) {
        float v = 0.f;
        if (c < me && i < n) {
          if (n <= m) v = (i == c) ? 1.f : 0.f;
          else v = (c == 0) ? 1.f : hash_unit(i, c);
        }
        if (c < rows) {
          Ut[c * ldt + i] = v;
          Yt[c * ldt + i] = 0.f;
        }
        Uf[uf_index<MT>(c, i)] = n > m ? v : 0.f;
      }
    }
    __syncthreads();

    // ---- D-orthonormalise the start block: only when its Rayleigh quotients are used right away (a tiny segment,
    //      one exact step); otherwise the first product runs on the raw block and is orthonormalised after it
    const bool ortho_start = n <= m || P.max_iter <= 1;
    if (ortho_start) {
      weighted_grams<G>(Ut, Ut, dg, n, m, ldt, rows, Gs, Hs, false);
      cholesky_inverse<MB, G>(Gs, m, me, pinv, Ws, misc);
      orthonormalise<MT, G>(Ws, m, Ut, Ut, Uf, npad, ldt, rows);
    }
    PHASE_END(PH_INIT);

    int it = 0;
    const float tol2 = P.tol * P.tol;
    while (true) {
      ++it;
      // single-pass products only where later full-precision steps follow (not for the one exact step of a tiny segment)
      if (it <= P.fast_iters && n > m && it < P.max_iter) matvec<MT, T, false, G>(Ag, lda, n, Uf, Yt, dinv, ldt, rows);
      else matvec<MT, T, true, G>(Ag, lda, n, Uf, Yt, dinv, ldt, rows);
      __syncthreads();
      PHASE_END(PH_MATVEC);
      const bool last = it >= P.max_iter || n <= m;  // n <= m: span(U) is the whole space, one step is exact
      // ---- G = Y^T D Y, H = U^T D Y (U is D-orthonormal) and the trigger: for each wanted column j
      //        |y_j - U h_j|_D^2 + sum_{a >= kk} H[a][j]^2   =   residual of the Ritz problem on the leading columns
      weighted_grams<G>(Ut, Yt, dg, n, m, ldt, rows, Gs, Hs, true);
      PHASE_END(PH_GRAMS);
      const bool scheduled = last || (it % P.rr_every) == 0;
      bool fired = false;
      // the trigger is evaluated on scheduled iterations too: if the leading block has already converged, the exact
      // leading-block step below finishes the segment instead of another approximate whole-block step
      bool test = !last && it >= 2 && it > P.fast_iters;
      if (test) {
        // cheap screen: |y_c - U h_c|_D^2 + coupling = G_cc - sum_{a < kk} H_ac^2 up to rounding (U is D-orthonormal);
        // far above the tolerance (and the rounding floor) means not converged -- skip the explicit residuals
        if (warp == 0) {
          float v = 0.f;
          for (int c = lane; c < kk; c += 32) {
            const float gcc = Gs[c * ld + c];
            float e = gcc;
            for (int a = 0; a < kk; ++a) e = fmaf(-Hs[a * ld + c], Hs[a * ld + c], e);
            // exempt only if even the upper bound theta + |r| of the eigenvalue is below the floor (Ritz values
            // approach it from below: an early estimate alone must not excuse a pair from converging)
            const bool wanted = Hs[c * ld + c] + sqrtf(fmaxf(e, 0.f)) >= P.lam_floor;
            e -= 64.f * tol2 + 8e-6f * gcc;
            if (wanted) v = fmaxf(v, e);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
          if (lane == 0) misc[2] = v > 0.f ? 0.f : 1.f;
        }
        __syncthreads();
        test = misc[2] != 0.f;
        __syncthreads();
      }
      if (test) {
        span_residuals<MT, G>(Hs, m, Ut, Yt, dg, kk, npad, ldt, rows, colred, res);
        if (warp == 0) {  // one wanted column per lane, then a warp maximum
          float worst = 0.f;
          for (int c = lane; c < kk; c += 32) {
            float v = res[c];
            for (int a = kk; a < me; ++a) v = fmaf(Hs[a * ld + c], Hs[a * ld + c], v);
            if (Hs[c * ld + c] + sqrtf(fmaxf(v, 0.f)) >= P.lam_floor) worst = fmaxf(worst, v);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) worst = fmaxf(worst, __shfl_xor_sync(0xffffffffu, worst, o));
          if (lane == 0) misc[2] = worst <= tol2 ? 1.f : 0.f;
        }
        __syncthreads();
        fired = misc[2] != 0.f;
        PHASE_END(PH_TRIGGER);
      }
      bool rotated = false;
      if (scheduled || fired) {
        // ---- Rayleigh-Ritz: the whole block (scheduled; a few sweeps unless it is the last step), or only the
        //      leading columns once they span an invariant subspace (trigger; to full accuracy)
        int md = m, sweeps = last ? 12 : EIG_SWEEPS;
        if (fired) {
          const int mdb = (kk + 1) & ~1;
          if (mdb <= me) { md = mdb; sweeps = 12; }
        }
        jacobi<G>(Hs, Ss, ld, md, sweeps, rot);
        PHASE_END(PH_JACOBI);
        if (threadIdx.x < m) {
          const int a = threadIdx.x;
          const float ta = Hs[a * ld + a];
          if (a < md) {
            int rank = 0;
            for (int b = 0; b < md; ++b) {
              const float tb = Hs[b * ld + b];
              rank += (tb > ta || (tb == ta && b < a)) ? 1 : 0;
            }
            order[rank] = a;
            theta[rank] = ta;
          } else {
            order[a] = a;
            theta[a] = ta;
          }
        }
        __syncthreads();
        for (int e = threadIdx.x; e < m * m; e += THREADS) {
          const int c = (m & (m - 1)) == 0 ? e >> (31 - __clz(m)) : e / m, a = e - c * m;
          Ws[c * ld + a] = (c < md && a < md) ? Ss[a * ld + order[c]] : (a == c ? 1.f : 0.f);
        }
        __syncthreads();
        rotate<MT, G>(Ws, m, Ut, Yt, dg, theta, kk, npad, ldt, rows, colred, res);
        if (warp == 0) {
          float worst = 0.f;
          for (int c = lane; c < kk; c += 32)
            if (theta[c] + sqrtf(fmaxf(res[c], 0.f)) >= P.lam_floor) worst = fmaxf(worst, res[c]);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) worst = fmaxf(worst, __shfl_xor_sync(0xffffffffu, worst, o));
          if (lane == 0) misc[1] = worst;
        }
        __syncthreads();
        PHASE_END(PH_ROTATE);
        if (last || misc[1] <= tol2) break;
        rotated = true;
      }
      // ---- U = orth_D(Y)
      if (rotated) {
        weighted_grams<G>(Ut, Yt, dg, n, m, ldt, rows, Gs, Hs, false);  // G of the rotated Y
        PHASE_END(PH_GRAMS);
      }
      const float piv = cholesky_inverse<MB, G>(Gs, m, me, pinv, Ws, misc);
      PHASE_END(PH_CHOL);
      orthonormalise<MT, G>(Ws, m, Yt, Ut, Uf, npad, ldt, rows);
      PHASE_END(PH_ORTH);
      if (piv < EIG_REORTH) {
        weighted_grams<G>(Ut, Ut, dg, n, m, ldt, rows, Gs, Hs, false);
        cholesky_inverse<MB, G>(Gs, m, me, pinv, Ws, misc);
        orthonormalise<MT, G>(Ws, m, Ut, Ut, Uf, npad, ldt, rows);
        PHASE_END(PH_ORTH);
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
    if ((k & 3) == 0 && (reinterpret_cast<uintptr_t>(P.V) & 15) == 0) {
      // 4 consecutive columns of one token per thread: one sqrt and one 128-bit store (k % 4 == 0 keeps them aligned)
      const int kq = k >> 2;
      for (int e = threadIdx.x; e < n * kq; e += THREADS) {
        const int i = e / kq, c = (e - i * kq) << 2;
        const float sd = sqrtf(dg[i]);
        float4 v;
        v.x = c + 0 < me ? Ut[(c + 0) * ldt + i] * sd * res[c + 0] : 0.f;
        v.y = c + 1 < me ? Ut[(c + 1) * ldt + i] * sd * res[c + 1] : 0.f;
        v.z = c + 2 < me ? Ut[(c + 2) * ldt + i] * sd * res[c + 2] : 0.f;
        v.w = c + 3 < me ? Ut[(c + 3) * ldt + i] * sd * res[c + 3] : 0.f;
        *reinterpret_cast<float4*>(Vout + static_cast<size_t>(i) * k + c) = v;
      }
    } else {
      for (int e = threadIdx.x; e < n * k; e += THREADS) {
        const int i = e / k, c = e - i * k;
        Vout[e] = c < me ? Ut[c * ldt + i] * sqrtf(dg[i]) * res[c] : 0.f;
      }
    }
    if (P.iters && threadIdx.x == 0) P.iters[s] = it;
    PHASE_END(PH_OUTPUT);
  }
}

template <int MT, int THREADS>
static int launch(const Params& P, cudaStream_t stream) {
  const size_t smem = static_cast<size_t>(make_layout(P.N, P.m, MT, THREADS / 32).total) * 4;
  const size_t kMaxSmem = 227 * 1024;
  if (smem > kMaxSmem) return MSVIT_ERR_SHAPE;
  cudaError_t e = cudaFuncSetAttribute(ncut_eig_kernel<MT, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return cuda_status(e);
  int per_sm = 1;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ncut_eig_kernel<MT, THREADS>, THREADS, smem);
  if (e != cudaSuccess) return cuda_status(e);
  if (per_sm < 1) per_sm = 1;
  // one CTA per segment while that is at most a few waves, else a persistent grid-stride loop
  const long long cap = 16LL * per_sm * sm_count();
  int grid = static_cast<int>(P.S < cap ? P.S : cap);
#ifdef EIG_MAX_GRID
  if (grid > EIG_MAX_GRID) grid = EIG_MAX_GRID;
#endif
  ncut_eig_kernel<MT, THREADS><<<grid, THREADS, smem, stream>>>(P);
  return cuda_status(cudaGetLastError());
}

template <int MT>
static int launch_threads(const Params& P, cudaStream_t stream) {
#ifdef EIG_FORCE_THREADS
  return launch<MT, EIG_FORCE_THREADS>(P, stream);
#endif
  if (P.N <= 256) return launch<MT, 128>(P, stream);
  if (P.N <= 512) return launch<MT, 256>(P, stream);
  return launch<MT, 512>(P, stream);
}

}  // namespace eig
}  // namespace msvit

#ifdef EIG_PROFILE
extern "C" int msvit_eig_profile(unsigned long long* host_out, int reset) {
  using namespace msvit::eig;
  cudaError_t e = cudaMemcpyFromSymbol(host_out, g_phase_cycles, sizeof(unsigned long long) * PH_COUNT);
  if (e != cudaSuccess) return static_cast<int>(e);
  if (reset) {
    unsigned long long z[PH_COUNT] = {};
    e = cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z));
  }
  return static_cast<int>(e);
}
#endif

extern "C" int msvit_ncut_eig(const float* A, const float* deg, float* V, float* lam, int32_t* iters,
                              int64_t total_rows, int S, int N, int k, int block, int max_iter, float tol,
                              float lam_floor, int n_converge, const int32_t* seg_off, const int64_t* a_off,
                              msvit_stream_t stream_) {
  using namespace msvit;
  using namespace msvit::eig;
  if (!A || !deg || !V || !lam) return MSVIT_ERR_NULL;
  if (S < 0 || N <= 0 || k <= 0 || total_rows < 0 || max_iter <= 0 || !(tol > 0.f)) return MSVIT_ERR_SHAPE;
  if (block < k || block > MSVIT_MAX_EIG_BLOCK || (block & 3) != 0) return MSVIT_ERR_SHAPE;
  if (n_converge < 0 || n_converge > k) return MSVIT_ERR_SHAPE;
  if (!seg_off && total_rows != static_cast<int64_t>(S) * N) return MSVIT_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(A) & 15) != 0) return MSVIT_ERR_ALIGN;
  if (S == 0 || total_rows == 0) return MSVIT_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);

  Params P;
  P.A = A; P.deg = deg; P.V = V; P.lam = lam; P.iters = iters;
  P.seg_off = seg_off; P.a_off = a_off;
  P.S = S; P.N = N; P.k = k; P.m = block;
  P.kconv = n_converge > 0 ? n_converge : k;
  P.max_iter = max_iter; P.rr_every = EIG_RR_EVERY; P.tol = tol; P.lam_floor = lam_floor;
  P.fast_iters = EIG_FAST_ITERS;
  if (block <= 16) return launch_threads<1>(P, stream);
  return launch_threads<2>(P, stream);
}
