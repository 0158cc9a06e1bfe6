// Experiment: how the tcgen05.mma issue loop is written vs cycles per MMA (small N, TS form, kind::f16).
//   V0: one thread (threadIdx.x == 0) runs the loop          V1: whole warp runs the loop, `if (elect_one())` around each MMA
//   V2: whole warp runs the loop, elect.sync inside the asm   V3: as V2, two warps issue to different accumulators
#include <cstdio>
#include "common.cuh"
#include "sm100_ptx.cuh"
using namespace msvit;

struct Sh {
  alignas(1024) uint8_t b[256 * 128];
  uint64_t bar[2];
  uint32_t tmem_base;
};

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

template <int V>
__device__ void run(Sh& sh, uint32_t tb, int N, int count, long long* out, uint32_t& phase) {
  long long t0 = 0, t1 = 0, t2 = 0;
  const int warp = threadIdx.x >> 5;
  const uint32_t idesc = make_idesc(0u, 128u, static_cast<uint32_t>(N));
  const uint32_t b_addr = smem_u32(sh.b);
  if (V == 0) {
    if (threadIdx.x == 0) {
      tc_fence_after();
      t0 = clock64();
      for (int i = 0; i < count; ++i)
        umma_ts<false>(tb + 256, tb + 8 * (i & 7), make_kmajor_sw128_desc(b_addr + (i & 3) * 32), idesc, i ? 1u : 0u);
      tc_commit(&sh.bar[0]);
      t1 = clock64();
    }
  } else if (V == 1) {
    if (warp == 0) {
      tc_fence_after();
      t0 = clock64();
      for (int i = 0; i < count; ++i)
        if (elect_one())
          umma_ts<false>(tb + 256, tb + 8 * (i & 7), make_kmajor_sw128_desc(b_addr + (i & 3) * 32), idesc, i ? 1u : 0u);
      if (elect_one()) tc_commit(&sh.bar[0]);
      t1 = clock64();
    }
  } else {
    if (warp == 0 || (V == 3 && warp == 1)) {
      tc_fence_after();
      t0 = clock64();
      const int n = V == 3 ? count / 2 : count;
      for (int i = 0; i < n; ++i)
        umma_ts_f16_elect(tb + 256 + 64 * warp, tb + 8 * (i & 7), make_kmajor_sw128_desc(b_addr + (i & 3) * 32), idesc,
                          i ? 1u : 0u);
      tc_commit_elect(&sh.bar[warp]);
      t1 = clock64();
    }
  }
  mbar_wait(&sh.bar[0], phase);
  if (V == 3) mbar_wait(&sh.bar[1], phase);
  phase ^= 1;
  tc_fence_after();
  if (threadIdx.x == 0) { t2 = clock64(); out[0] = t1 - t0; out[1] = t2 - t0; }
  tc_fence_before();
  __syncthreads();
  if (V != 3) {  // keep bar[1]'s phase in step
    if (threadIdx.x == 0) mbar_arrive(&sh.bar[1]);
    __syncthreads();
  }
}

__global__ void __launch_bounds__(128, 1) k(long long* out) {
  extern __shared__ uint8_t raw[];
  Sh& sh = *reinterpret_cast<Sh*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&sh.bar[0], 1); mbar_init(&sh.bar[1], 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&sh.tmem_base, 512); tmem_relinquish(); }
  for (int e = threadIdx.x; e < 256 * 128 / 4; e += 128) reinterpret_cast<uint32_t*>(sh.b)[e] = 0;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = sh.tmem_base;
  {
    uint32_t z[16] = {};
    for (int c = 0; c < 64; c += 16) tmem_st16(tb + (static_cast<uint32_t>(warp * 32) << 16) + c, z);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  uint32_t phase = 0;
  for (int rep = 0; rep < 2; ++rep) {
    int slot = 0;
    for (int N : {16, 32}) {
      run<0>(sh, tb, N, 52, out + 2 * slot++, phase);
      run<1>(sh, tb, N, 52, out + 2 * slot++, phase);
      run<2>(sh, tb, N, 52, out + 2 * slot++, phase);
      run<3>(sh, tb, N, 52, out + 2 * slot++, phase);
    }
  }
  if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
  long long* out;
  cudaMalloc(&out, 32 * sizeof(long long));
  const size_t smem = sizeof(Sh) + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<<<1, 128, smem>>>(out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  long long h[32];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  int slot = 0;
  for (int N : {16, 32})
    for (int v = 0; v < 4; ++v, ++slot)
      printf("N=%2d V%d: issue %5lld complete %5lld (%.1f cyc/mma over 52)\n", N, v, h[2 * slot], h[2 * slot + 1], h[2 * slot + 1] / 52.0);
  return 0;
}
