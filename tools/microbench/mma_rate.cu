// Microbenchmark: issue rate of legacy mma.sync (tf32 m16n8k8, f16 m16n8k16) and FFMA on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_f16(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int KIND, int NACC>
__global__ void bench(float* out, int iters, long long* cyc) {
  unsigned a[4] = {threadIdx.x + 1u, threadIdx.x * 3u, 7u, 9u}, b[2] = {threadIdx.x, 5u};
  float c[NACC][4];
  for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  float f[NACC * 4];
  for (int i = 0; i < NACC * 4; ++i) f[i] = threadIdx.x * 0.001f + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (KIND == 0) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) mma_tf32(c[i], a, b);
    } else if (KIND == 1) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) mma_f16(c[i], a, b);
    } else {
#pragma unroll
      for (int i = 0; i < NACC * 4; ++i) f[i] = fmaf(f[i], 1.0001f + f[(i + 1) % (NACC * 4)] * 1e-9f, 0.5f);
    }
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  for (int i = 0; i < NACC * 4; ++i) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int KIND, int NACC>
void run(const char* name, int warps, double work_per_instr) {
  float* out; long long* cyc;
  const int grid = 148, iters = 4096;
  cudaMalloc(&out, grid * warps * 32 * sizeof(float));
  cudaMalloc(&cyc, grid * sizeof(long long));
  bench<KIND, NACC><<<grid, warps * 32>>>(out, 16, cyc);
  bench<KIND, NACC><<<grid, warps * 32>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < grid; ++i) c += h[i]; c /= grid;
  const double n_instr = (double)iters * NACC * (KIND == 2 ? 4 : 1) * warps;  // warp-instructions per SM
  printf("%-10s warps/SM=%2d  acc=%d  cycles=%.0f  cycles per warp-instr (SM-wide)=%.3f  %s/clk/SM=%.1f\n", name, warps, NACC,
         c, c / n_instr, KIND == 2 ? "FMA" : "MAC", n_instr * work_per_instr / c);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {1, 4, 8, 16}) {
    if (w == 1) { run<0, 8>("tf32 k8", 1, 1024); run<1, 8>("f16 k16", 1, 2048); run<2, 8>("ffma", 1, 32); }
    if (w == 4) { run<0, 8>("tf32 k8", 4, 1024); run<1, 8>("f16 k16", 4, 2048); run<2, 8>("ffma", 4, 32); }
    if (w == 8) { run<0, 8>("tf32 k8", 8, 1024); run<1, 8>("f16 k16", 8, 2048); run<2, 8>("ffma", 8, 32); }
    if (w == 16) { run<0, 8>("tf32 k8", 16, 1024); run<1, 8>("f16 k16", 16, 2048); run<2, 8>("ffma", 16, 32); }
  }
  run<0, 2>("tf32 k8", 4, 1024);
  run<0, 4>("tf32 k8", 4, 1024);
  return 0;
}
