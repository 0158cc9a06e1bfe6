// Pieces shared by the tcgen05 kernels (affinity.cu, ncut_fused.cu, gkmeans.cu): the per-row pass over a staged
// 128-byte k-slice row (sum of squares, in-place TF32 rounding) and the host-side TMA tensor-map encoding.
#pragma once
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace msvit {

// Sum of squares of one 128-byte row of a k-slice as the tensor core sees it.
// fp32 rows are first rounded to TF32 (round to nearest, in place): the tensor core would otherwise TRUNCATE
// the low 13 mantissa bits, which shrinks every distance by ~1e-3 relative (a bias, not noise).
template <bool TF32>
__device__ __forceinline__ float row_sumsq(uint8_t* row, int lane) {
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    // rotate the 16-byte chunk order by lane so that the 8 lanes of a phase hit distinct bank groups
    uint4* qp = reinterpret_cast<uint4*>(row + (((c + lane) & 7) << 4));
    const uint4 q = *qp;
    uint32_t w[4] = {q.x, q.y, q.z, q.w};
    if constexpr (TF32) {
#pragma unroll
      for (int i = 0; i < 4; ++i) w[i] = (w[i] + 0x1000u) & 0xFFFFE000u;
      *qp = make_uint4(w[0], w[1], w[2], w[3]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if constexpr (TF32) {
        const float v = __uint_as_float(w[i]);
        acc = fmaf(v, v, acc);
      } else {
        const float lo = __uint_as_float(w[i] << 16);
        const float hi = __uint_as_float(w[i] & 0xFFFF0000u);
        acc = fmaf(lo, lo, acc);
        acc = fmaf(hi, hi, acc);
      }
    }
  }
  return acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Looked up once per process (the driver entry point does not change); nullptr if the driver lacks it.
inline EncodeTiledFn encode_fn() {
  static const EncodeTiledFn cached = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      fn = nullptr;
    return reinterpret_cast<EncodeTiledFn>(fn);
  }();
  return cached;
}

// 2-D map over x [rows, D] (row-major, 16-bit or fp32): box = one 128-byte k-slice x box_rows rows, 128B swizzle.
inline int make_map(EncodeTiledFn enc, CUtensorMap* m, const void* x, bool f32, int64_t rows, int D, int box_rows) {
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(D), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(D) * (f32 ? 4 : 2)};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(f32 ? 32 : 64), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                         const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MSVIT_OK : MSVIT_ERR_DRIVER;
}

}  // namespace msvit
