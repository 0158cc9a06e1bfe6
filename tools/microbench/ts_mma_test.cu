// Experiment: tcgen05.mma with the A operand in tensor memory (TS form), kind::tf32 and kind::f16 (packed halves).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../multi-state-vit_b200/csrc -o ts_mma_test ts_mma_test.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include "common.cuh"
#include "sm100_ptx.cuh"
using namespace msvit;

__host__ __device__ inline float aval(int r, int c) { return c < 196 && r < 196 ? float(((r * 5 + c * 3) % 13) - 6) * 0.125f : 0.f; }
__host__ __device__ inline float uval(int comp, int j) { return j < 196 ? float(((comp * 7 + j * 11) % 9) - 4) * 0.25f : 0.f; }

struct Sh {
  alignas(1024) float u32[7][16][32];      // tf32 B operand: 7 k-slices of [16 rows x 128 B], 128B swizzle
  alignas(1024) __half u16[4][16][64];     // f16 B operand: 4 k-slices of [16 rows x 128 B]
  uint64_t bar;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(128, 1) ts_test(float* yout, long long* cyc, int reps) {
  extern __shared__ uint8_t raw[];
  Sh& sh = *reinterpret_cast<Sh*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&sh.bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&sh.tmem_base, 512); tmem_relinquish(); }
  // B operands (swizzled: 16-byte chunk c of row r sits at chunk c ^ (r & 7))
  for (int e = threadIdx.x; e < 7 * 16 * 32; e += 128) {
    const int sl = e / 512, r = (e / 32) % 16, k = e % 32;
    const int chunk = (k >> 2) ^ (r & 7);
    sh.u32[sl][r][chunk * 4 + (k & 3)] = uval(r, sl * 32 + k);
  }
  for (int e = threadIdx.x; e < 4 * 16 * 64; e += 128) {
    const int sl = e / 1024, r = (e / 64) % 16, k = e % 64;
    const int chunk = (k >> 3) ^ (r & 7);
    sh.u16[sl][r][chunk * 8 + (k & 7)] = __float2half(uval(r, sl * 64 + k));
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = sh.tmem_base;
  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
  uint32_t phase = 0;

  // ---------------- test 1: tf32, A tiles at columns 0 and 208, D at 416 + 16 t
  for (int t = 0; t < 2; ++t)
    for (int c0 = 0; c0 < 208; c0 += 16) {
      uint32_t v[16];
      for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(aval(t * 128 + warp * 32 + lane, c0 + i));
      tmem_st16(tb + lane_base + t * 208 + c0, v);
    }
  tmem_wait_st();
  tc_fence_before();
  __syncthreads();
  long long t0 = 0, t1 = 0, t2 = 0;
  for (int rep = 0; rep < reps; ++rep) {
    if (threadIdx.x == 0) {
      tc_fence_after();
      t0 = clock64();
      const uint32_t idesc = make_idesc(2u, 128u, 16u);
      for (int t = 0; t < 2; ++t)
        for (int ks = 0; ks < 26; ++ks) {
          const uint64_t bd = make_kmajor_sw128_desc(smem_u32(&sh.u32[ks >> 2][0][0]) + (ks & 3) * 32);
          umma_ts<true>(tb + 416 + 16 * t, tb + t * 208 + 8 * ks, bd, idesc, ks ? 1u : 0u);
        }
      tc_commit(&sh.bar);
      t1 = clock64();
    }
    mbar_wait(&sh.bar, phase);
    phase ^= 1;
    tc_fence_after();
    if (threadIdx.x == 0) t2 = clock64();
  }
  for (int t = 0; t < 2; ++t) {
    float v[16];
    tmem_ld16(tb + lane_base + 416 + 16 * t, v);
    for (int i = 0; i < 16; ++i) yout[(0 * 256 + t * 128 + warp * 32 + lane) * 16 + i] = v[i];
  }
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t0; }
  tc_fence_before();
  __syncthreads();

  // ---------------- test 2 / 3: f16 packed, A tiles at columns 0 and 104; packing variant pv: 0 = even k in the low half
  for (int pv = 0; pv < 2; ++pv) {
    for (int t = 0; t < 2; ++t)
      for (int c0 = 0; c0 < 112; c0 += 16) {
        uint32_t v[16];
        for (int i = 0; i < 16; ++i) {
          const int q = c0 + i;
          const int r = t * 128 + warp * 32 + lane;
          const __half lo = __float2half(aval(r, 2 * q + pv)), hi = __float2half(aval(r, 2 * q + 1 - pv));
          v[i] = static_cast<uint32_t>(__half_as_ushort(lo)) | (static_cast<uint32_t>(__half_as_ushort(hi)) << 16);
        }
        if (c0 + 16 <= 104 || true) tmem_st16(tb + lane_base + t * 112 + c0, v);
      }
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
      t0 = clock64();
      const uint32_t idesc = make_idesc(0u, 128u, 16u);
      for (int t = 0; t < 2; ++t)
        for (int ks = 0; ks < 13; ++ks) {
          const uint64_t bd = make_kmajor_sw128_desc(smem_u32(&sh.u16[ks >> 2][0][0]) + (ks & 3) * 32);
          umma_ts<false>(tb + 448 + 16 * t, tb + t * 112 + 8 * ks, bd, idesc, ks ? 1u : 0u);
        }
      tc_commit(&sh.bar);
    }
    mbar_wait(&sh.bar, phase);
    phase ^= 1;
    tc_fence_after();
    if (threadIdx.x == 0) { t2 = clock64(); cyc[2 + pv] = t2 - t0; }
    for (int t = 0; t < 2; ++t) {
      float v[16];
      tmem_ld16(tb + lane_base + 448 + 16 * t, v);
      for (int i = 0; i < 16; ++i) yout[((1 + pv) * 256 + t * 128 + warp * 32 + lane) * 16 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
  }
  if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
  float* y; long long* cyc;
  cudaMalloc(&y, 3 * 256 * 16 * sizeof(float));
  cudaMalloc(&cyc, 8 * sizeof(long long));
  cudaMemset(y, 0, 3 * 256 * 16 * sizeof(float));
  const size_t smem = sizeof(Sh) + 1024;
  cudaFuncSetAttribute(ts_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int reps : {1, 8}) {
    ts_test<<<1, 128, smem>>>(y, cyc, reps);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    long long hc[4];
    cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
    printf("reps %d: tf32 issue %lld cyc, issue+complete %lld cyc; f16 pv0 %lld pv1 %lld\n", reps, hc[0], hc[1], hc[2], hc[3]);
  }
  static float h[3 * 256 * 16];
  cudaMemcpy(h, y, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[3] = {"tf32 TS", "f16 TS (even k low)", "f16 TS (odd k low)"};
  for (int v = 0; v < 3; ++v) {
    int bad = 0; double maxerr = 0;
    for (int r = 0; r < 196; ++r)
      for (int c = 0; c < 16; ++c) {
        double ref = 0;
        for (int j = 0; j < 196; ++j) ref += (double)aval(r, j) * uval(c, j);
        const double err = fabs(ref - h[(v * 256 + r) * 16 + c]);
        if (err > 1e-3) ++bad;
        if (err > maxerr) maxerr = err;
      }
    printf("%-22s mismatches %d / %d, max err %.3g   sample y[3][2] = %g\n", names[v], bad, 196 * 16, maxerr, h[(v * 256 + 3) * 16 + 2]);
  }
  return 0;
}
