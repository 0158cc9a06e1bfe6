// Experiment: cycles per tcgen05.mma for small N, SS vs TS operand forms (timing only; operand values are zero).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I../../multi-state-vit_b200/csrc \
//        -o mma_shape_time mma_shape_time.cu
#include <cstdio>
#include "common.cuh"
#include "sm100_ptx.cuh"
using namespace msvit;

struct Sh {
  alignas(1024) uint8_t a[128 * 128];   // A operand k-slice: 128 rows x 128 B
  alignas(1024) uint8_t b[256 * 128];   // B operand k-slice: up to 256 rows x 128 B
  uint64_t bar;
  uint32_t tmem_base;
};

template <bool TF32, bool TS>
__device__ void run(Sh& sh, uint32_t tb, int N, int count, long long* out, uint32_t& phase) {
  long long t0 = 0, t1 = 0, t2 = 0;
  if (threadIdx.x == 0) {
    tc_fence_after();
    t0 = clock64();
    const uint32_t idesc = make_idesc(TF32 ? 2u : 0u, 128u, static_cast<uint32_t>(N));
    const uint32_t a_addr = smem_u32(sh.a), b_addr = smem_u32(sh.b);
    for (int i = 0; i < count; ++i) {
      const uint64_t bd = make_kmajor_sw128_desc(b_addr + (i & 3) * 32);
      if constexpr (TS) umma_ts<TF32>(tb + 256, tb + 8 * (i & 7), bd, idesc, i ? 1u : 0u);
      else umma_ss<TF32>(tb + 256, make_kmajor_sw128_desc(a_addr + (i & 3) * 32), bd, idesc, i ? 1u : 0u);
    }
    tc_commit(&sh.bar);
    t1 = clock64();
  }
  mbar_wait(&sh.bar, phase);
  phase ^= 1;
  tc_fence_after();
  if (threadIdx.x == 0) { t2 = clock64(); out[0] = t1 - t0; out[1] = t2 - t0; }
  tc_fence_before();
  __syncthreads();
}

__global__ void __launch_bounds__(128, 1) k(long long* out) {
  extern __shared__ uint8_t raw[];
  Sh& sh = *reinterpret_cast<Sh*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&sh.bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&sh.tmem_base, 512); tmem_relinquish(); }
  for (int e = threadIdx.x; e < (128 + 256) * 128 / 4; e += 128) reinterpret_cast<uint32_t*>(sh.a)[e] = 0;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = sh.tmem_base;
  // zero the TMEM A region (columns 0..63) so that TS operands are finite
  {
    uint32_t z[16] = {};
    for (int c = 0; c < 64; c += 16) tmem_st16(tb + (static_cast<uint32_t>(warp * 32) << 16) + c, z);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  uint32_t phase = 0;
  const int Ns[5] = {16, 32, 64, 128, 256};
  for (int rep = 0; rep < 2; ++rep) {
    int slot = 0;
    for (int ni = 0; ni < 5; ++ni) {
      for (int count : {1, 26, 52}) {
        run<true, true>(sh, tb, Ns[ni], count, out + 2 * slot++, phase);
        run<true, false>(sh, tb, Ns[ni], count, out + 2 * slot++, phase);
        run<false, true>(sh, tb, Ns[ni], count, out + 2 * slot++, phase);
        run<false, false>(sh, tb, Ns[ni], count, out + 2 * slot++, phase);
      }
    }
  }
  if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
  long long* out;
  cudaMalloc(&out, 2 * 60 * sizeof(long long));
  const size_t smem = sizeof(Sh) + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<<<1, 128, smem>>>(out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  long long h[120];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  const int Ns[5] = {16, 32, 64, 128, 256};
  const char* kinds[4] = {"tf32 TS", "tf32 SS", "f16 TS", "f16 SS"};
  int slot = 0;
  for (int ni = 0; ni < 5; ++ni)
    for (int count : {1, 26, 52})
      for (int kd = 0; kd < 4; ++kd, ++slot)
        printf("N=%3d count=%2d %-8s issue %6lld  complete %6lld  (%.1f cyc/mma)\n", Ns[ni], count, kinds[kd], h[2 * slot],
               h[2 * slot + 1], (double)h[2 * slot + 1] / count);
  return 0;
}
