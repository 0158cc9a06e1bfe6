// Microbenchmark: latency of the 16 x 16 register Cholesky variants on one warp (cycles per factorisation) + check.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I../../multi-state-vit_b200/csrc -o chol_time chol_time.cu
#include <cstdio>
#include <cmath>
#include "eig_core.cuh"
using namespace msvit;

// variant 1: no division on the critical path, plain rsqrt
__device__ __forceinline__ void chol_v1(float (&g)[16], float dorig, int me, float* LT, float* pinv, float* misc) {
  const int lane = threadIdx.x & 31;
  const bool act = lane < me;
  float minpiv = 1.0f;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float piv = __shfl_sync(0xffffffffu, g[j], j);
    const float gjj = __shfl_sync(0xffffffffu, dorig, j);
    float rowj[16];
#pragma unroll
    for (int c = j + 1; c < 16; ++c) rowj[c] = __shfl_sync(0xffffffffu, g[c], j);
    const bool ok = j < me && piv > 1e-6f * gjj && piv > 0.f;
    float inv = rsqrtf(piv);
    inv = ok ? inv : 0.f;
    if (j < me) minpiv = fminf(minpiv, ok ? piv * __frcp_rn(gjj) : 1.0f);
    const float lij = (lane >= j && act) ? g[j] * inv : 0.f;
    if (lane < 16) LT[j * 16 + lane] = lij;
    if (lane == j) pinv[j] = inv;
    const float f = lane > j ? lij * inv : (lane == j ? 1.f : 0.f);
#pragma unroll
    for (int c = j + 1; c < 16; ++c) g[c] = fmaf(-f, rowj[c], g[c]);
  }
  if (lane == 0) misc[0] = minpiv;
}

// variant 2: two columns per step (2 x 2 pivot block in closed form: both reciprocal roots start together)
__device__ __forceinline__ void chol_v2(float (&g)[16], float dorig, int me, float* LT, float* pinv, float* misc) {
  const int lane = threadIdx.x & 31;
  const bool act = lane < me;
  float minpiv = 1.0f;
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    // rows j and j+1 of the current Schur complement
    float r0[16], r1[16];
#pragma unroll
    for (int c = j; c < 16; ++c) {
      r0[c] = __shfl_sync(0xffffffffu, g[c], j);
      r1[c] = __shfl_sync(0xffffffffu, g[c], j + 1);
    }
    const float g0 = __shfl_sync(0xffffffffu, dorig, j), g1 = __shfl_sync(0xffffffffu, dorig, j + 1);
    const float a = r0[j], b = r0[j + 1], cc = r1[j + 1];
    const bool ok0 = j < me && a > 1e-6f * g0 && a > 0.f;
    const float ia = ok0 ? __frcp_rn(a) : 0.f;          // 1 / a
    const float s11 = fmaf(-b * ia, b, cc);              // Schur pivot of column j+1
    const bool ok1 = j + 1 < me && s11 > 1e-6f * g1 && s11 > 0.f;
    float inv0 = ok0 ? rsqrtf(a) : 0.f;
    float inv1 = ok1 ? rsqrtf(s11) : 0.f;
    if (j < me) minpiv = fminf(minpiv, ok0 ? a * __frcp_rn(g0) : 1.0f);
    if (j + 1 < me) minpiv = fminf(minpiv, ok1 ? s11 * __frcp_rn(g1) : 1.0f);
    // this lane's entries: L[i][j] = s_ij inv0 ; L[i][j+1] = (s_i,j+1 - L[i][j] L[j+1][j]) inv1
    const float l10 = b * inv0;                          // L[j+1][j]
    const float li0 = (lane >= j && act) ? g[j] * inv0 : 0.f;
    float li1 = (lane >= j + 1 && act) ? fmaf(-li0, l10, g[j + 1]) * inv1 : 0.f;
    if (lane < 16) { LT[j * 16 + lane] = li0; LT[(j + 1) * 16 + lane] = li1; }
    if (lane == j) pinv[j] = inv0;
    if (lane == j + 1) pinv[j + 1] = inv1;
    // Schur update with both columns: s_ic -= L[i][j] L[c][j] + L[i][j+1] L[c][j+1], L[c][j] = r0[c] inv0,
    // L[c][j+1] = (r1[c] - l10 r0[c] inv0) inv1
    const float f0 = lane > j + 1 ? li0 * inv0 : 0.f;
    const float f1 = lane > j + 1 ? li1 * inv1 : 0.f;
    const float e = l10 * inv0;
#pragma unroll
    for (int c = j + 2; c < 16; ++c) {
      const float t1 = fmaf(-e, r0[c], r1[c]);          // L[c][j+1] / inv1
      g[c] = fmaf(-f1, t1, fmaf(-f0, r0[c], g[c]));
    }
    if (lane == j || lane == j + 1) {
#pragma unroll
      for (int c = j + 2; c < 16; ++c) g[c] = 0.f;
    }
  }
  if (lane == 0) misc[0] = minpiv;
}


// variant 3: every lane factorises the whole matrix redundantly in registers (no shuffles at all); lane a then writes
// row a of L^T.  The matrix comes from shared memory (row stride 16) as broadcast loads.
__device__ __forceinline__ void chol_v3(const float* __restrict__ Gs16, const float* __restrict__ dorig16, int me,
                                        float* LT, float* pinv, float* misc) {
  const int lane = threadIdx.x & 31;
  float L[16][16];   // lower triangle: L[i][j], j <= i (fully unrolled: registers)
#pragma unroll
  for (int i = 0; i < 16; ++i) {
#pragma unroll
    for (int q4 = 0; q4 <= (i >> 2); ++q4) {
      const float4 v = *reinterpret_cast<const float4*>(Gs16 + i * 16 + 4 * q4);
      L[i][4 * q4] = v.x;
      if (4 * q4 + 1 <= i) L[i][4 * q4 + 1] = v.y;
      if (4 * q4 + 2 <= i) L[i][4 * q4 + 2] = v.z;
      if (4 * q4 + 3 <= i) L[i][4 * q4 + 3] = v.w;
    }
  }
  float minpiv = 1.0f;
  float inv[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float piv = L[j][j];
    const float gjj = dorig16[j];
    const bool ok = j < me && piv > 1e-6f * gjj && piv > 0.f;
    const float r = ok ? rsqrtf(piv) : 0.f;
    inv[j] = r;
    if (j < me) minpiv = fminf(minpiv, ok ? piv * __frcp_rn(gjj) : 1.0f);
    L[j][j] = piv * r;
#pragma unroll
    for (int i = j + 1; i < 16; ++i) L[i][j] *= r;
#pragma unroll
    for (int i = j + 1; i < 16; ++i)
#pragma unroll
      for (int c = j + 1; c <= i; ++c) L[i][c] = fmaf(-L[i][j], L[c][j], L[i][c]);
  }
  // lane a writes row a of L^T: LT[a][c] = L[c][a], c >= a
#pragma unroll
  for (int a = 0; a < 16; ++a) {
    if (lane == a) {
#pragma unroll
      for (int c = 0; c < 16; ++c) LT[a * 16 + c] = c >= a ? L[c][a] : 0.f;
      pinv[a] = inv[a];
    }
  }
  if (lane == 0) misc[0] = minpiv;
}

// variant 4: the 136 entries of the lower triangle are dealt to the 32 lanes (entry e = lane + 32 t); per step the
// raw column j is published in shared memory, every lane derives 1 / pivot itself and updates its own entries:
// s_ic -= s_ij s_cj / s_jj.  No shuffles, no reciprocal root on the critical path.
__device__ __forceinline__ void chol_v4(const float* __restrict__ Gs16, const float* __restrict__ dor, int me,
                                        float* LT, float* pinv, float* misc, float* col) {
  const int lane = threadIdx.x & 31;
  int ei[5], ec[5];
  float s[5];
#pragma unroll
  for (int t = 0; t < 5; ++t) {
    const int e = lane + 32 * t;
    int i = 0;
    while ((i + 1) * (i + 2) / 2 <= e) ++i;      // row of entry e (e < 136 -> i < 16)
    const int c = e - i * (i + 1) / 2;
    ei[t] = e < 136 ? i : 16;
    ec[t] = e < 136 ? c : -1;                     // never matches a column
    s[t] = e < 136 ? Gs16[i * 16 + c] : 0.f;
  }
  float minpiv = 1.0f;
#pragma unroll 1
  for (int j = 0; j < 16; ++j) {
    float* cb = col + (j & 1) * 16;
#pragma unroll
    for (int t = 0; t < 5; ++t)
      if (ec[t] == j) cb[ei[t]] = s[t];
    __syncwarp();
    const float piv = cb[j];
    const float gjj = dor[j];
    const bool ok = j < me && piv > 1e-6f * gjj && piv > 0.f;
    const float inv2 = ok ? __frcp_rn(piv) : 0.f;
    const float r = ok ? rsqrtf(piv) : 0.f;
    if (j < me) minpiv = fminf(minpiv, ok ? piv * __frcp_rn(gjj) : 1.0f);
#pragma unroll
    for (int t = 0; t < 5; ++t) {
      if (ec[t] > j) s[t] = fmaf(-cb[ei[t]] * cb[ec[t]], inv2, s[t]);
      else if (ec[t] == j) LT[j * 16 + ei[t]] = s[t] * r;
    }
    if (lane == 0) pinv[j] = r;
  }
  if (lane == 0) misc[0] = minpiv;
}

// variant 5: as variant 1, all stores after the loop (L[lane][j] kept in registers)
__device__ __forceinline__ void chol_v5(float (&g)[16], float dorig, int me, float* LT, float* pinv, float* misc) {
  const int lane = threadIdx.x & 31;
  const bool act = lane < me;
  float minpiv = 1.0f;
  float lcol[16], invs[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float piv = __shfl_sync(0xffffffffu, g[j], j);
    const float gjj = __shfl_sync(0xffffffffu, dorig, j);
    float rowj[16];
#pragma unroll
    for (int c = j + 1; c < 16; ++c) rowj[c] = __shfl_sync(0xffffffffu, g[c], j);
    const bool ok = j < me && piv > 1e-6f * gjj && piv > 0.f;
    float inv = rsqrtf(piv);
    inv = ok ? inv : 0.f;
    invs[j] = inv;
    if (j < me) minpiv = fminf(minpiv, ok ? piv * __frcp_rn(gjj) : 1.0f);
    const float lij = (lane >= j && act) ? g[j] * inv : 0.f;
    lcol[j] = lij;
    const float f = lane > j ? lij * inv : (lane == j ? 1.f : 0.f);
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
// variant 6: as variant 5, the pivot row is published in shared memory and read back as four 128-bit broadcast loads
__device__ __forceinline__ void chol_v6(float (&g)[16], float dorig, int me, float* LT, float* pinv, float* misc, float* rowbuf) {
  const int lane = threadIdx.x & 31;
  const bool act = lane < me;
  float minpiv = 1.0f;
  float lcol[16], invs[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float* rb = rowbuf + (j & 1) * 16;
    if (lane == j) {
#pragma unroll
      for (int q4 = (j >> 2); q4 < 4; ++q4) *reinterpret_cast<float4*>(rb + 4 * q4) = make_float4(g[4 * q4], g[4 * q4 + 1], g[4 * q4 + 2], g[4 * q4 + 3]);
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
    float inv = rsqrtf(piv);
    inv = ok ? inv : 0.f;
    invs[j] = inv;
    if (j < me) minpiv = fminf(minpiv, ok ? piv * __frcp_rn(gjj) : 1.0f);
    const float lij = (lane >= j && act) ? g[j] * inv : 0.f;
    lcol[j] = lij;
    const float f = lane > j ? lij * inv : (lane == j ? 1.f : 0.f);
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

template <int V>
__global__ void k(const float* Gin, float* LTout, float* pout, long long* cyc, int reps) {
  __shared__ __align__(16) float LT[256];
  __shared__ __align__(16) float pinv[16];
  __shared__ float misc[8];
  __shared__ __align__(16) float Gs16[256];
  __shared__ float dor[16];
  __shared__ float colb[32];
  __shared__ __align__(16) float rowbuf[32];
  for (int e = threadIdx.x; e < 256; e += 32) Gs16[e] = Gin[e];
  if (threadIdx.x < 16) dor[threadIdx.x] = Gin[threadIdx.x * 17];
  __syncwarp();
  const int lane = threadIdx.x;
  long long best = 1ll << 60;
  for (int r = 0; r < reps; ++r) {
    float g[16];
    for (int c = 0; c < 16; ++c) g[c] = lane < 16 ? Gin[lane * 16 + c] : 0.f;
    const float dorig = lane < 16 ? Gin[lane * 17] : 0.f;
    __syncwarp();
    const long long t0 = clock64();
    if (V == 0) eig::cholesky_lt16_regs(g, dorig, 16, LT, pinv, misc, rowbuf);
    if (V == 1) chol_v1(g, dorig, 16, LT, pinv, misc);
    if (V == 2) chol_v2(g, dorig, 16, LT, pinv, misc);
    if (V == 3) chol_v3(Gs16, dor, 16, LT, pinv, misc);
    if (V == 4) chol_v4(Gs16, dor, 16, LT, pinv, misc, colb);
    if (V == 5) chol_v5(g, dorig, 16, LT, pinv, misc);
    if (V == 6) chol_v6(g, dorig, 16, LT, pinv, misc, rowbuf);
    __syncwarp();
    const long long t1 = clock64();
    if (t1 - t0 < best) best = t1 - t0;
    if (r == 0 && lane == 0) cyc[1] = t1 - t0;
  }
  for (int e = lane; e < 256; e += 32) LTout[e] = LT[e];
  if (lane < 16) pout[lane] = pinv[lane];
  if (lane == 0) cyc[0] = best;
}

int main() {
  float G[256], Y[64][16];
  srand(1);
  for (int i = 0; i < 64; ++i) for (int c = 0; c < 16; ++c) Y[i][c] = (rand() / (float)RAND_MAX - 0.5f) * powf(0.7f, c);
  for (int a = 0; a < 16; ++a) for (int b = 0; b < 16; ++b) { double s = 0; for (int i = 0; i < 64; ++i) s += (double)Y[i][a] * Y[i][b]; G[a * 16 + b] = (float)s; }
  float *dG, *dL, *dp; long long* dc;
  cudaMalloc(&dG, sizeof(G)); cudaMalloc(&dL, 1024); cudaMalloc(&dp, 64); cudaMalloc(&dc, 16);
  cudaMemcpy(dG, G, sizeof(G), cudaMemcpyHostToDevice);
  for (int v = 0; v < 7; ++v) {
    if (v == 0) k<0><<<1, 32>>>(dG, dL, dp, dc, 20);
    if (v == 1) k<1><<<1, 32>>>(dG, dL, dp, dc, 20);
    if (v == 2) k<2><<<1, 32>>>(dG, dL, dp, dc, 20);
    if (v == 3) k<3><<<1, 32>>>(dG, dL, dp, dc, 20);
    if (v == 4) k<4><<<1, 32>>>(dG, dL, dp, dc, 20);
    if (v == 5) k<5><<<1, 32>>>(dG, dL, dp, dc, 20);
    if (v == 6) k<6><<<1, 32>>>(dG, dL, dp, dc, 20);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    float L[256]; long long c2[2]; 
    cudaMemcpy(L, dL, 1024, cudaMemcpyDeviceToHost); cudaMemcpy(c2, dc, 16, cudaMemcpyDeviceToHost); const long long c = c2[0];
    // check L L^T = G  (LT[a*16+c] = L[c][a])
    double maxerr = 0;
    for (int i = 0; i < 16; ++i) for (int j = 0; j <= i; ++j) {
      double s = 0; for (int a = 0; a <= j; ++a) s += (double)L[a * 16 + i] * L[a * 16 + j];
      maxerr = fmax(maxerr, fabs(s - G[i * 16 + j]) / sqrt((double)G[i * 17] * G[j * 17]));
    }
    printf("variant %d: %lld cycles (first call %lld), max scaled |L L^T - G| = %.2e\n", v, c, c2[1], maxerr);
  }
  return 0;
}
