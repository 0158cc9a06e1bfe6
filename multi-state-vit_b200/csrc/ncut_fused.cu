// Fused affinity + NCut subspace iteration for whole images of up to 224 tokens (ViT-B/16: 196): the affinity matrix
// is produced in tensor memory and never leaves the SM.
//
//   G = X X^T                        tcgen05.mma (SS form), k-slices of the tokens TMA-staged in 128B-swizzled shared
//                                    memory, fp32 accumulators in TMEM (two M=128 tiles, T = round16(N) columns each)
//   A = exp(-d(G) / gamma)           epilogue IN PLACE in TMEM: every 16 fp32 accumulator columns become 8 columns of
//                                    packed fp16 pairs A_hi followed by 8 columns A_lo (A = A_hi + A_lo to ~2^-22);
//                                    deg = row sums of the stored values
//   Y = D^-1 A U                     tcgen05.mma (TS form: A operand read from TMEM, kind::f16): A_hi [U_hi | U_lo] and
//                                    A_lo U_hi accumulate into 32 more TMEM columns per tile; the block U (16 columns,
//                                    fp16 hi/lo pair scaled by 2^10) is the K-major shared-memory operand
//   G2 = Y^T D Y, H = U^T D Y        one more tcgen05.mma chain: [Y_hi; Y_lo; U_hi; U_lo] (64 operand rows, K = tokens)
//                                    times [DY_hi | DY_lo]; warp quadrant 0 reads G2, quadrant 1 reads H from TMEM
//   U = Y L^-T,  L L^T = G2          Cholesky QR in the D inner product: the warp that read G2 factorises it in
//                                    registers (lane = row), forward substitution per token (thread = token, u and y
//                                    live in registers); meanwhile the warp that read H screens for convergence
//   stop                             when the leading columns span an invariant subspace to the tolerance (or at the
//                                    iteration cap); the Rayleigh-Ritz rotation of the result is left to the next
//                                    kernel (kmeans.cu: ritz_kmeans_kernel), where hundreds of segments overlap
//
// Reference math: sandbox/test.py:108-118 (affinity, degree, normalised operator, leading eigenvectors),
// model/clustering/modeling_spectral.py:54-61 (gamma, distance modes).  Same iteration as ncut_eig.cu, which streams
// the affinity from L2 and serves every other shape.
//
// One persistent CTA per SM, 320 threads: warp 0 = TMA producer (runs ahead into the next image while the
// eigensolver works), warp 1 = TMEM allocator + Gram MMA issuer, warps 2..9 = row norms / TF32 rounding during the
// k-loop, then epilogue and eigensolver (thread = token row; warps 2 and 6 also issue the TS-form MMAs of their tile).
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>

#include "eig_core.cuh"
#include "tile_ops.cuh"

namespace msvit {
namespace fused {

#ifdef FUSED_PROFILE
// Development instrumentation: cycles the first compute thread of every CTA spends in each phase.
enum { FP_GRAM, FP_EPI, FP_INIT, FP_UOP, FP_PRODUCT, FP_DRAIN, FP_GRAMS, FP_TRIGGER, FP_CHOL, FP_SUBST, FP_OUTPUT, FP_X1, FP_X2, FP_COUNT };
__device__ unsigned long long g_fused_cycles[FP_COUNT];
#define FPHASE_BEGIN() long long fp_t0 = clock64()
#define FPHASE_END(ph)                                                                     \
  do {                                                                                     \
    const long long fp_t1 = clock64();                                                     \
    if (ct == 0) atomicAdd(&g_fused_cycles[ph], (unsigned long long)(fp_t1 - fp_t0));      \
    fp_t0 = fp_t1;                                                                         \
  } while (0)
#else
#define FPHASE_BEGIN()
#define FPHASE_END(ph)
#endif

constexpr int kThreads = 320;
constexpr int kRitzFirst = 10;     // first iteration at which a segment that has not converged rotates to its Ritz basis
constexpr int kRitzPeriod = 5;     // ... and how often after that
constexpr int kRitzSweeps = 3;     // Jacobi sweeps of such a rotation (any orthogonal rotation keeps the invariants)
constexpr int kComputeBase = 64;
constexpr int kCompute = 256;
constexpr int kSliceBytes = 128;
constexpr int kMaxStages = 6;
constexpr int kTmemCols = 512;
constexpr int kMB = 16;               // subspace block width (columns of U)
static_assert(kCompute == kMB * kMB, "the Ritz rotation fills the kMB x kMB matrix with one entry per compute thread");
constexpr int kMaxT = 208;            // 2 * 208 affinity columns + 96 accumulator columns fill the 512 of TMEM
constexpr int kTileCols = 208;        // TMEM column stride of the two affinity tiles
constexpr float kUScale = 1024.f;     // |u| <= 1 for a D-orthonormal block (deg >= 1): fp16 operands never overflow
constexpr float kDScale = 64.f;       // |deg * y| <= deg <= N before the first orthonormalisation, <= sqrt(deg) after
constexpr int kOpARows = 64;          // operand block A per 64-token k-slice: rows [Y_hi; Y_lo; U_hi; U_lo] x 128 bytes
constexpr int kOpBRows = 32;          // operand block B: rows [DY_hi; DY_lo]
constexpr int kOpABytes = 4 * kOpARows * kSliceBytes;
constexpr int kOpBBytes = 4 * kOpBRows * kSliceBytes;
constexpr int kUopBytes = kOpABytes + kOpBBytes;
using G = ThreadGroup<kComputeBase, kCompute, 1>;

struct Params {
  float* deg;       // [rows]
  float* U;         // [rows, 16] D-orthonormal basis of the final subspace
  float* H;         // [S, 16, 16] projected operator U^T D (D^-1 A) U
  int32_t* iters;   // [S]
  int32_t* info;    // [S] 1 = the leading block met the tolerance, 0 = stopped at the iteration cap
  int S, N, T;
  int mode;
  float c2;
  int n_kslices, k_step, stages, stage_bytes, tail_rows;
  float in_scale;   // fp16 Gram: the tokens are multiplied by this power of two before the conversion
  int m, kconv, max_iter, fast_iters;
  float tol, lam_floor;
  int debug;        // development: stop the segment early (0 = off)
};

struct Shared {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t conv[kMaxStages];
  uint64_t tmem_full;
  uint64_t tmem_empty;
  uint64_t ybar;        // product accumulators complete (one commit per M tile)
  uint64_t gbar;        // Gram accumulator complete
  uint32_t tmem_base;
  uint32_t pad;
  float rq[256];    // per-token quantity of the distance (row and column side are the same tokens)
};

// float offsets of the eigensolver's small shared arrays
struct EigLayout {
  int Gs, Hs, LT, pinv, misc, red, rowbuf, total;
};
__host__ __device__ inline EigLayout make_eig_layout() {
  EigLayout L;
  int o = 0;
  L.Gs = o;      o += kMB * kMB;   // row stride 16: rows are read as 128-bit broadcast loads
  L.Hs = o;      o += kMB * kMB;
  L.LT = o;      o += 256;
  L.pinv = o;    o += 16;
  L.misc = o;    o += 8;
  L.red = o;     o += (kCompute / 32) * 16;   // per-warp partial sums of the span residuals
  L.rowbuf = o;  o += 32;                     // pivot rows of the Cholesky (16-byte aligned)
  L.total = o;
  return L;
}

__device__ __forceinline__ float row_quantity(int mode, float sumsq, float c2) {
  if (mode == MSVIT_DIST_RBF) return 0.5f * sumsq * c2;
  if (mode == MSVIT_DIST_COSINE) return rsqrtf(fmaxf(sumsq, 1e-30f));
  return sqrtf(sumsq);
}
__device__ __forceinline__ float affinity_log2(int mode, float g, float rq, float cq, float c2) {
  float t;
  if (mode == MSVIT_DIST_RBF) t = fmaf(g, c2, -cq) - rq;
  else if (mode == MSVIT_DIST_COSINE) t = fmaf(g * rq, cq, -1.0f) * c2;
  else t = (g - rq * cq) * c2;
  return fminf(t, 0.0f);
}

// tcgen05.mma, A from TMEM, kind::f16, issued by one elected lane of a converged warp (the loop around it stays
// warp-uniform, which keeps the operands in uniform registers).
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
__device__ __forceinline__ void umma_ss_f16_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}

// two floats -> one register of two fp16 (round to nearest), `lo` in the low half
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// Token input of the Gram contraction.
//   kInBf16: bf16 tokens, kind::f16 straight from the TMA stages.
//   kInTf32: fp32 tokens rounded to TF32 in place by the compute warps, kind::tf32 (any distance mode).
//   kInFp16: fp32 tokens scaled by a power of two, rounded to the same 11-bit significand and converted IN PLACE to fp16
//            by the compute warps (the 128-byte row of 32 fp32 becomes 64 bytes of fp16 in its first four 16-byte chunks);
//            kind::f16 then needs half the instructions and reads half the bytes per k-slice.  rbf distance only: the
//            caller's distance scale fixes the magnitude of the tokens that matter, which chooses the power of two.
constexpr int kInBf16 = 0, kInTf32 = 1, kInFp16 = 2;

// Row of an fp32 k-slice (128 bytes, 128B-swizzled by TMA; `row` = its index in the tile) -> fp16 in logical chunks 0..3
// of the same row (round to nearest even, saturating; the token norms come from the diagonal of the Gram matrix, i.e.
// from the converted values themselves).
__device__ __forceinline__ void convert_row_fp16(uint8_t* rowp, int row, float scale) {
  const int sw = row & 7;                      // physical 16-byte chunk = logical chunk ^ sw
  const bool odd = (row & 1) != 0;
  uint4 q[8];
#pragma unroll
  for (int c = 0; c < 8; ++c)                  // physical chunk (c + row) & 7: 8 consecutive rows hit 8 bank groups
    q[c] = *reinterpret_cast<const uint4*>(rowp + (((c + row) & 7) << 4));
#pragma unroll
  for (int c = 0; c < 8; c += 2) {
    // q[c] always holds the even logical chunk of its pair (physical (c + row) & 7, logical = physical ^ sw, and sw has
    // the parity of row); its partner is the next physical chunk for an even row, the previous one for an odd row
    const uint4 a4 = q[c];
    const uint4 b4 = odd ? q[(c + 7) & 7] : q[c + 1];
    const uint32_t w[8] = {a4.x, a4.y, a4.z, a4.w, b4.x, b4.y, b4.z, b4.w};
    uint32_t h[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h[e])
          : "f"(__uint_as_float(w[2 * e + 1]) * scale), "f"(__uint_as_float(w[2 * e]) * scale));
    const int j = ((((c + row) & 6)) ^ sw) >> 1;             // logical pair = logical 16-byte chunk of the fp16 row
    *reinterpret_cast<uint4*>(rowp + ((j ^ sw) << 4)) = make_uint4(h[0], h[1], h[2], h[3]);
  }
}

template <int IN>
__global__ void __launch_bounds__(kThreads, 1)
ncut_fused_kernel(const __grid_constant__ CUtensorMap tm_full, const __grid_constant__ CUtensorMap tm_tail,
                  const Params P) {
  extern __shared__ uint8_t smem_raw[];
  // (pointer arithmetic on the shared array itself, not on an integer: the compiler keeps the shared address space and
  // emits LDS / STS instead of generic loads and stores for everything derived from it)
  uint8_t* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr bool TF32 = IN == kInTf32;
  constexpr bool CONV = IN != kInBf16;   // the compute warps rewrite the staged tile before the tensor core reads it
  uint8_t* uop = tiles + static_cast<size_t>(P.stages) * P.stage_bytes;   // 1024-aligned (stage_bytes % 1024 == 0)
  Shared& sh = *reinterpret_cast<Shared*>(uop + kUopBytes);
  float* ef = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(&sh) + ((sizeof(Shared) + 15) & ~size_t(15)));
  const EigLayout L = make_eig_layout();

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n = P.N, T = P.T;
  const int n_tiles = T > 128 ? 2 : 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < P.stages; ++i) {
      mbar_init(&sh.full[i], 1);
      mbar_init(&sh.empty[i], 1 + kCompute / 32);
      mbar_init(&sh.conv[i], kCompute / 32);
    }
    mbar_init(&sh.tmem_full, 1);
    mbar_init(&sh.tmem_empty, kCompute / 32);
    mbar_init(&sh.ybar, n_tiles);
    mbar_init(&sh.gbar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tm_full);
    tma_prefetch_desc(&tm_tail);
  }
  if (warp == 1) {
    tmem_alloc(&sh.tmem_base, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh.tmem_base;

  uint32_t it_ring = 0;  // k-slice counter (ring position)
  uint32_t job = 0;      // segment counter of this CTA (TMEM phase)

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const int rows0 = T > 128 ? 128 : T;
      const int rows1 = T - rows0;  // = P.tail_rows when > 0
      for (int s = blockIdx.x; s < P.S; s += gridDim.x) {
        const int row0 = s * n;
        for (int ks = 0; ks < P.n_kslices; ++ks, ++it_ring) {
          const int st = it_ring % P.stages;
          const uint32_t ph = (it_ring / P.stages) & 1;
          mbar_wait(&sh.empty[st], ph ^ 1);
          uint8_t* bt = tiles + static_cast<size_t>(st) * P.stage_bytes;
          mbar_arrive_expect_tx(&sh.full[st], static_cast<uint32_t>(T) * kSliceBytes);
          tma_load_2d(bt, rows0 == 128 ? &tm_full : &tm_tail, &sh.full[st], ks * P.k_step, row0);
          if (rows1 > 0) tma_load_2d(bt + 128 * kSliceBytes, &tm_tail, &sh.full[st], ks * P.k_step, row0 + 128);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ Gram MMA issuer (lane 0 issues and commits)
    const uint32_t idesc = make_idesc(TF32 ? 2u : (IN == kInFp16 ? 0u : 1u), 128u, static_cast<uint32_t>(T));
    constexpr int kSteps = IN == kInFp16 ? 2 : 4;   // MMA k-steps per staged row: 64 bytes of fp16, else the whole 128 bytes
    for (int s = blockIdx.x; s < P.S; s += gridDim.x, ++job) {
      mbar_wait(&sh.tmem_empty, (job & 1) ^ 1);
      tc_fence_after();
      for (int ks = 0; ks < P.n_kslices; ++ks, ++it_ring) {
        const int st = it_ring % P.stages;
        const uint32_t ph = (it_ring / P.stages) & 1;
        mbar_wait(CONV ? &sh.conv[st] : &sh.full[st], ph);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t b_addr = smem_u32(tiles + static_cast<size_t>(st) * P.stage_bytes);
#pragma unroll
          for (int kk = 0; kk < kSteps; ++kk) {  // 32 bytes of K per step
            const uint64_t bd = make_kmajor_sw128_desc(b_addr + kk * 32);
            const uint32_t acc = (ks | kk) != 0 ? 1u : 0u;
            umma_ss<TF32>(tmem_base, bd, bd, idesc, acc);
            if (n_tiles > 1)
              umma_ss<TF32>(tmem_base + kTileCols, make_kmajor_sw128_desc(b_addr + 128 * kSliceBytes + kk * 32), bd, idesc,
                            acc);
          }
          tc_commit(&sh.empty[st]);
          if (ks == P.n_kslices - 1) tc_commit(&sh.tmem_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ norms, epilogue, eigensolver (warps 2..9)
    const int ct = G::tid();            // 0..255
    const int cw = ct >> 5;             // 0..7
    const int q = warp & 3;             // TMEM lane quadrant this warp may access
    const int tile = cw >> 2;           // M tile of this warp
    const int row = tile * 128 + q * 32 + lane;   // token row of this thread inside the segment
    const bool valid = row < n;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t a_col = tile * kTileCols;    // this tile's affinity columns
    const uint32_t y_col = 2 * kTileCols + 32 * tile;   // this tile's product accumulator (32 columns: x_hi | x_lo)
    float* Gs = ef + L.Gs;
    float* Hs = ef + L.Hs;
    float* LT = ef + L.LT;
    float* pinv = ef + L.pinv;
    float* misc = ef + L.misc;          // [0] smallest scaled pivot, [2] trigger flag
    float* red = ef + L.red;
    float* rowbuf = ef + L.rowbuf;
    constexpr int m = kMB, ld = kMB;    // the launcher only accepts a block width of 16 ...
    constexpr int me = kMB;             // ... and more tokens than that
    const int kk = P.kconv < me ? P.kconv : me;
    const int npad = round_up(n, 16);
    const int nks = npad >> 4;          // K = 16 steps of every tensor-core contraction over the tokens
    const float tol2 = P.tol * P.tol;
    const uint32_t idesc32 = make_idesc(0u, 128u, 32u), idesc16 = make_idesc(0u, 128u, 16u);
    const uint32_t opa_addr = smem_u32(uop), opb_addr = opa_addr + kOpABytes;
    const uint32_t g_col = 2 * kTileCols + 64;   // Gram accumulator (32 columns), in flight together with the product
    uint32_t yphase = 0, gphase = 0;
    // the warps that issue the products sit on different schedulers (warp % 4) than each other and than the Cholesky
    // warp (4): warp 2 for M tile 0, warp 9 for M tile 1
    const bool prod_issuer = (cw == 0) || (cw == 7 && n_tiles > 1);
    // this thread's token inside the operand blocks: k-slice row >> 6, 16-byte chunk (kc >> 3) ^ (r & 7), 2 bytes at kc & 7
    uint8_t* const opa_tok = uop + (row >> 6) * (kOpARows * kSliceBytes);
    uint8_t* const opb_tok = uop + kOpABytes + (row >> 6) * (kOpBRows * kSliceBytes);
    const int kc = row & 63;
    // rows r0 .. r0+15 of an operand block <- fp16(scale * v), rows r0+16 .. r0+31 <- the fp16 remainder
    auto store_rows = [&](uint8_t* blk, int r0, const float (&v)[16], float scale, bool with_lo) {
      const int within = (kc & 7) * 2;
#pragma unroll
      for (int c = 0; c < 16; c += 2) {
        const float s0 = v[c] * scale, s1 = v[c + 1] * scale;
        const float h0 = __uint_as_float(__float_as_uint(s0) & 0xffffe000u);
        const float h1 = __uint_as_float(__float_as_uint(s1) & 0xffffe000u);
        const uint32_t hh = pack_f16x2(h0, h1);
        const int off0 = (((kc >> 3) ^ (c & 7)) << 4) + within;
        const int off1 = (((kc >> 3) ^ ((c + 1) & 7)) << 4) + within;
        *reinterpret_cast<uint16_t*>(blk + (r0 + c) * kSliceBytes + off0) = static_cast<uint16_t>(hh & 0xffffu);
        *reinterpret_cast<uint16_t*>(blk + (r0 + c + 1) * kSliceBytes + off1) = static_cast<uint16_t>(hh >> 16);
        if (with_lo) {
          const uint32_t ll = pack_f16x2(s0 - h0, s1 - h1);
          *reinterpret_cast<uint16_t*>(blk + (r0 + 16 + c) * kSliceBytes + off0) = static_cast<uint16_t>(ll & 0xffffu);
          *reinterpret_cast<uint16_t*>(blk + (r0 + 17 + c) * kSliceBytes + off1) = static_cast<uint16_t>(ll >> 16);
        }
      }
    };
    // [X_hi; X_lo]^T D [X'_hi | X'_lo] over the tokens: operand rows a_row0 .. a_row0+31 of block A against block B,
    // 32 accumulator columns at g_col; issued by one warp
    auto issue_gram = [&](int a_row0) {
      tc_fence_after();
      for (int ks = 0; ks < nks; ++ks) {
        const uint32_t koff = (ks >> 2) * (kOpARows * kSliceBytes) + (ks & 3) * 32;
        const uint32_t boff = (ks >> 2) * (kOpBRows * kSliceBytes) + (ks & 3) * 32;
        umma_ss_f16_elect(tmem_base + g_col, make_kmajor_sw128_desc(opa_addr + a_row0 * kSliceBytes + koff),
                          make_kmajor_sw128_desc(opb_addr + boff), idesc32, ks ? 1u : 0u);
      }
      tc_commit_elect(&sh.gbar);
    };
    // Z = A X on the tensor cores for this warp's M tile: X = the operand rows b_row0 .. b_row0+15 (hi) and +16 .. +31
    // (lo) of block A; A_hi [X_hi | X_lo] and A_lo X_hi (single pass A_hi X_hi unless `full`) -> the tile's 32 columns
    auto issue_product = [&](int b_row0, bool full) {
      tc_fence_after();
      const uint32_t a0 = tmem_base + a_col, d0 = tmem_base + y_col;
      const uint32_t xb = opa_addr + b_row0 * kSliceBytes;
      for (int ks = 0; ks < nks; ++ks) {
        const uint64_t bd = make_kmajor_sw128_desc(xb + (ks >> 2) * (kOpARows * kSliceBytes) + (ks & 3) * 32);
        umma_ts_f16_elect(d0, a0 + 16 * ks, bd, full ? idesc32 : idesc16, ks ? 1u : 0u);   // A_hi [x_hi | x_lo]
      }
      if (full) {
        for (int ks = 0; ks < nks; ++ks) {
          const uint64_t bd = make_kmajor_sw128_desc(xb + (ks >> 2) * (kOpARows * kSliceBytes) + (ks & 3) * 32);
          umma_ts_f16_elect(d0, a0 + 16 * ks + 8, bd, idesc16, 1u);                          // A_lo x_hi
        }
      }
      tc_commit_elect(&sh.ybar);
    };
    // this thread's row of the finished product, scaled to D^-1 A x
    auto read_product = [&](float (&z)[16], float dinv_s, bool full) {
      if (tile < n_tiles) {
        float za[16];
        tmem_ld16(lane_addr + y_col, za);
        if (full) {
          float zb[16];
          tmem_ld16(lane_addr + y_col + 16, zb);
#pragma unroll
          for (int c = 0; c < 16; ++c) za[c] += zb[c];
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) z[c] = valid ? za[c] * dinv_s : 0.f;
      }
    };
    // Accumulator rows 32 q .. 32 q + 31 (q = this warp's lane quadrant) hold [X_hi; X_lo] . [DX'_hi | DX'_lo]:
    // lanes 0..15 return row `lane` of X^T D X' = hi.hi + hi.lo + lo.hi, unscaled
    auto read_gram_rows = [&](float (&g)[16], float inv_scale) {
      float v0[16], v1[16];
      tmem_ld16(lane_addr + g_col, v0);
      tmem_ld16(lane_addr + g_col + 16, v1);
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const float mine = lane < 16 ? v0[c] + v1[c] : v0[c];
        g[c] = (mine + __shfl_down_sync(0xffffffffu, mine, 16)) * inv_scale;
      }
    };

    for (int s = blockIdx.x; s < P.S; s += gridDim.x, ++job) {
      const int row0 = s * n;
      FPHASE_BEGIN();
      // ---- row norms (and TF32 rounding of the staged tile) while the Gram k-loop runs
      float ss = 0.f;
      for (int ks = 0; ks < P.n_kslices; ++ks, ++it_ring) {
        const int st = it_ring % P.stages;
        const uint32_t ph = (it_ring / P.stages) & 1;
        mbar_wait(&sh.full[st], ph);
        uint8_t* bt = tiles + static_cast<size_t>(st) * P.stage_bytes;
        if constexpr (IN == kInFp16) {
          if (ct < T) convert_row_fp16(bt + ct * kSliceBytes, ct, P.in_scale);
        } else {
          if (ct < T) ss += row_sumsq<TF32>(bt + ct * kSliceBytes, lane);
        }
        if constexpr (CONV) fence_proxy_async_smem();  // the rewritten tile must be visible to the tensor core
        __syncwarp();
        if (lane == 0) {
          if constexpr (CONV) mbar_arrive(&sh.conv[st]);
          mbar_arrive(&sh.empty[st]);
        }
      }
      if constexpr (IN != kInFp16) {
        sh.rq[ct] = row_quantity(P.mode, ss, P.c2);
        G::sync();
      }

      // ---- epilogue, in place in TMEM: 16 fp32 Gram columns -> 8 columns of fp16 pairs A_hi + 8 columns A_lo
      mbar_wait(&sh.tmem_full, job & 1);
      FPHASE_END(FP_GRAM);
      tc_fence_after();
      if constexpr (IN == kInFp16) {
        // fp16 Gram: the squared norm of a token is the diagonal of the Gram matrix (the same products, fp32
        // accumulation), so the conversion loop above carries no per-element norm arithmetic.  A warp's 32 diagonal
        // entries sit in two 16-column chunks of its own lanes.
        const int rbase = tile * 128 + q * 32;
        if (tile < n_tiles && rbase < T) {
          float v0[16], v1[16];
          tmem_ld16(lane_addr + a_col + rbase, v0);
          if (rbase + 16 < T) tmem_ld16(lane_addr + a_col + rbase + 16, v1);
          else {
#pragma unroll
            for (int e = 0; e < 16; ++e) v1[e] = 0.f;
          }
          float g = 0.f;
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            g = lane == e ? v0[e] : g;
            g = lane == 16 + e ? v1[e] : g;
          }
          sh.rq[row] = row_quantity(P.mode, g, P.c2);
        }
        G::sync();
      }
      const float rq = sh.rq[row];
      float rowsum = 0.f;
      if (tile < n_tiles) {
        for (int c = 0; c * 16 < T; ++c) {
          float v[16];
          tmem_ld16(lane_addr + a_col + c * 16, v);
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int col = c * 16 + 2 * i;
            const float2 cq = *reinterpret_cast<const float2*>(&sh.rq[col]);
            const float a0 = col < n ? fast_exp2(affinity_log2(P.mode, v[2 * i], rq, cq.x, P.c2)) : 0.f;
            const float a1 = col + 1 < n ? fast_exp2(affinity_log2(P.mode, v[2 * i + 1], rq, cq.y, P.c2)) : 0.f;
            // hi = the value with its low 13 mantissa bits cleared (exactly an fp16 number for a >= 2^-14; below that
            // the conversion rounds by < 2^-25), lo = the exact remainder rounded to fp16: two packed conversions
            const float h0 = __uint_as_float(__float_as_uint(a0) & 0xffffe000u);
            const float h1 = __uint_as_float(__float_as_uint(a1) & 0xffffe000u);
            rowsum += a0 + a1;
            w[i] = pack_f16x2(h0, h1);
            w[8 + i] = pack_f16x2(a0 - h0, a1 - h1);
          }
          tmem_st16(lane_addr + a_col + c * 16, w);
        }
        tmem_wait_st();
      }
      if (valid) P.deg[row0 + row] = rowsum;
      const float d = valid ? rowsum : 0.f;
      const float dinv_s = (valid && rowsum > 0.f) ? (1.0f / kUScale) / rowsum : 0.f;  // 1 / (deg * operand scale)
      FPHASE_END(FP_EPI);

      // ---- start block (same as ncut_eig.cu): column 0 constant, the rest pseudo-random; pad tokens zero
      float u[16], y[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        u[c] = (valid && c < me) ? (c == 0 ? 1.f : eig::hash_unit(static_cast<uint32_t>(row), static_cast<uint32_t>(c))) : 0.f;
        y[c] = 0.f;
      }

      int it = 0;
      bool fired = false;
      float z[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) z[c] = 0.f;
      // ---- first product on the raw start block: y = D^-1 A u  (operand rows [U_hi; U_lo])
      if (row < npad) store_rows(opa_tok, 32, u, kUScale, true);
      fence_proxy_async_smem();
      tc_fence_before();
      G::sync();
      const bool full0 = (P.debug & 128) || P.fast_iters < 1;
      if (prod_issuer) issue_product(32, full0);
      mbar_wait(&sh.ybar, yphase);
      yphase ^= 1;
      tc_fence_after();
      read_product(y, dinv_s, full0);
      FPHASE_END(FP_INIT);
      // Invariant at the top of every iteration: y = D^-1 A u.  The iteration orthonormalises y (Cholesky QR in the D
      // inner product) and, IN THE SHADOW of the Gram + Cholesky step, already multiplies the un-orthonormalised y
      // by the operator: z = D^-1 A y, so that the next iterate's image is L^-1 z (the factor is triangular: the
      // leading columns never see the small trailing ones).
      for (;;) {
      bool need_rotate = false;
      while ((P.debug & 15) != 1) {
        ++it;
        const bool last = it >= P.max_iter;
        const bool test = !last && it >= 2 && it > P.fast_iters;
        // ---- operand rows: [Y_hi; Y_lo] and [DY_hi; DY_lo]; [U_hi; U_lo] when H = U^T D Y is looked at
        if (row < npad) {
          store_rows(opa_tok, 0, y, kUScale, true);
          float dy[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) dy[c] = d * y[c];
          store_rows(opb_tok, 0, dy, kDScale, true);
          if (test || last) store_rows(opa_tok, 32, u, kUScale, true);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        G::sync();
        FPHASE_END(FP_UOP);
        // ---- tensor cores: G2 = Y^T D Y and H = U^T D Y (warp 3), z = A y (warps 2 and 6: one M tile each)
        if (cw == 1) issue_gram(0);
        const bool fulln = (P.debug & 128) || it + 1 > P.fast_iters;   // precision of the product that feeds iteration it + 1
        if ((P.debug & 64) && prod_issuer) issue_product(0, fulln);   // development: issue before the Gram completes
        // (only the warps that need a result wait on its mbarrier; the others block in the hardware barrier below and
        // leave their scheduler's issue slots to the warps that work)
        float grow[16];
        if (q < 2 && tile == 0) {   // warp 4 (quadrant 0): G2, warp 5 (quadrant 1): H
          mbar_wait(&sh.gbar, gphase);
          tc_fence_after();
          read_gram_rows(grow, 1.0f / (kUScale * kDScale));
          if (lane < 16) {
            float4* dst = reinterpret_cast<float4*>((q == 0 ? Gs : Hs) + lane * ld);
            dst[0] = make_float4(grow[0], grow[1], grow[2], grow[3]);
            dst[1] = make_float4(grow[4], grow[5], grow[6], grow[7]);
            dst[2] = make_float4(grow[8], grow[9], grow[10], grow[11]);
            dst[3] = make_float4(grow[12], grow[13], grow[14], grow[15]);
          }
        }
        gphase ^= 1;
        tc_fence_before();
        G::sync();
        FPHASE_END(FP_GRAMS);
        // (issued only now: product MMAs queued ahead of the Gram chain would delay the Cholesky by their length)
        if (!(P.debug & (64 | 2048)) && prod_issuer) issue_product(0, fulln);
        // ---- quadrant-0 warp: Cholesky of G2 in registers; quadrant-1 warp: cheap convergence screen on (G2, H)
        if (q == 0 && tile == 0) {
          float dorig = 0.f;
#pragma unroll
          for (int c = 0; c < 16; ++c) dorig = lane == c ? grow[c] : dorig;
#ifdef FUSED_PROFILE
          const long long xc0 = clock64();
#endif
          eig::cholesky_lt16_regs(grow, dorig, me, LT, pinv, misc, rowbuf);
#ifdef FUSED_PROFILE
          if (lane == 0) atomicAdd(&g_fused_cycles[FP_X1], (unsigned long long)(clock64() - xc0));
#endif
        } else if (q == 1 && tile == 0 && test) {
          // G_cc - sum_{a < kk} H_ac^2 far above the tolerance means not converged (see ncut_eig.cu)
          float v = 0.f;
          for (int c = lane; c < kk; c += 32) {
            const float gcc = Gs[c * ld + c];
            float e = gcc;
            for (int a = 0; a < kk; ++a) e = fmaf(-Hs[a * ld + c], Hs[a * ld + c], e);
            // exempt only if even the upper bound theta + |r| of the eigenvalue is below the floor
            const bool wanted = Hs[c * ld + c] + sqrtf(fmaxf(e, 0.f)) >= P.lam_floor;
            e -= 64.f * tol2 + 8e-6f * gcc;
            if (wanted) v = fmaxf(v, e);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
          if (lane == 0) misc[2] = v > 0.f ? 0.f : 1.f;
        }
        FPHASE_END(FP_CHOL);
#ifdef FUSED_PROFILE
        const long long xp0 = clock64();
#endif
        if (P.debug & 2048) {   // development: product only after the Cholesky
          G::sync();
          if (prod_issuer) issue_product(0, fulln);
        }
        if (prod_issuer) mbar_wait(&sh.ybar, yphase);
#ifdef FUSED_PROFILE
        if (ct == 0) atomicAdd(&g_fused_cycles[FP_X2], (unsigned long long)(clock64() - xp0));
#endif
        yphase ^= 1;
        tc_fence_before();
        G::sync();
        // ---- everyone: this thread's row of z
        tc_fence_after();
        read_product(z, dinv_s, fulln);
        FPHASE_END(FP_PRODUCT);
        if (test && misc[2] != 0.f) {
          // |y_c - U h_c|_D^2 + coupling to the trailing columns, summed over the wanted columns: bounds the residual
          // of every Ritz pair of the leading block.  Thread = token: r_c = y_c - sum_a u_a H[a][c]
          // (columns c >= kk are not needed: with kk <= 8 only the first two 128-bit pieces of every H row are read)
          float r[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) r[c] = y[c];
          const int nq = (kk + 3) >> 2;   // 128-bit pieces of an H row that hold wanted columns
#pragma unroll
          for (int a = 0; a < 16; ++a) {
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
              if (qq < nq) {   // uniform
                const float4 h = *reinterpret_cast<const float4*>(Hs + a * ld + 4 * qq);
                r[4 * qq] = fmaf(-u[a], h.x, r[4 * qq]);
                r[4 * qq + 1] = fmaf(-u[a], h.y, r[4 * qq + 1]);
                r[4 * qq + 2] = fmaf(-u[a], h.z, r[4 * qq + 2]);
                r[4 * qq + 3] = fmaf(-u[a], h.w, r[4 * qq + 3]);
              }
            }
          }
          if (P.lam_floor <= 0.f) {
            float tot = 0.f;
#pragma unroll
            for (int c = 0; c < 16; ++c)
              if (c < kk) tot = fmaf(d * r[c], r[c], tot);
            if (ct < kk)   // the coupling terms of column ct are added once
              for (int a = kk; a < me; ++a) tot = fmaf(Hs[a * ld + ct], Hs[a * ld + ct], tot);
            tot = warp_sum(tot);
            if (lane == 0) red[cw * 16] = tot;
            G::sync();
            float all = 0.f;
#pragma unroll
            for (int w8 = 0; w8 < kCompute / 32; ++w8) all += red[w8 * 16];
            fired = all <= tol2;
          } else {
            // eigenvalue-threshold mode: a column whose eigenvalue cannot reach the floor (theta + |r| < floor) is
            // exempt, which needs the residual column by column
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              const float e = warp_sum(c < kk ? d * r[c] * r[c] : 0.f);
              if (lane == 0) red[cw * 16 + c] = e;
            }
            G::sync();
            float all = 0.f;
            for (int c = 0; c < kk; ++c) {
              float e = 0.f;
#pragma unroll
              for (int w8 = 0; w8 < kCompute / 32; ++w8) e += red[w8 * 16 + c];
              for (int a = kk; a < me; ++a) e = fmaf(Hs[a * ld + c], Hs[a * ld + c], e);
              if (Hs[c * ld + c] + sqrtf(e) >= P.lam_floor) all += e;
            }
            fired = all <= tol2;
          }
        }
        FPHASE_END(FP_TRIGGER);
        if (fired || last) break;
        // ---- slow segments only: Rayleigh-Ritz on the whole block.  Plain orthogonal iteration drives the leading kk
        // columns to their invariant subspace at the rate lambda_{kk+1} / lambda_kk, which is what the test above
        // measures; a spectrum without a gap after kk (non-planted tokens) makes that slow although the 16-column
        // block already holds the wanted vectors to (lambda_17 / lambda_kk)^it.  Rotating the basis to the (approximate)
        // Ritz vectors of the block, eigenvalues descending, moves the wanted directions into the leading columns; the
        // rotation is orthogonal, so u stays D-orthonormal and y = D^-1 A u holds for the rotated pair.  Segments with a
        // gap (the planted workloads: 5-7 iterations) never get here.
        if (test && it >= kRitzFirst && (it - kRitzFirst) % kRitzPeriod == 0) {
          need_rotate = true;   // leave the hot loop: the rotation code sits behind it
          break;
        }
        // ---- next iterate: u = L^-1 y (D-orthonormal), y = L^-1 z = D^-1 A u
        float piv = misc[0];
        {
          float t[16];
          eig::forward_subst16(y, LT, pinv, t);
#pragma unroll
          for (int c = 0; c < 16; ++c) u[c] = t[c];
          eig::forward_subst16(z, LT, pinv, t);
#pragma unroll
          for (int c = 0; c < 16; ++c) y[c] = t[c];
        }
        if (piv < EIG_REORTH && !(P.debug & 1024)) {
          // ill-conditioned block (early iterations): orthonormalise once more.  G3 = U^T D U from the operand rows
          // [U_hi; U_lo] of block A against [DU_hi; DU_lo] in block B (the product that read block A has completed).
          if (row < npad) {
            store_rows(opa_tok, 32, u, kUScale, true);
            float du[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) du[c] = d * u[c];
            store_rows(opb_tok, 0, du, kDScale, true);
          }
          fence_proxy_async_smem();
          tc_fence_before();
          G::sync();   // also: every thread has read LT / pinv
          if (cw == 1) issue_gram(32);
          if (q == 0 && tile == 0) {
            mbar_wait(&sh.gbar, gphase);
            tc_fence_after();
            float g3[16];
            read_gram_rows(g3, 1.0f / (kUScale * kDScale));
            float dorig = 0.f;
#pragma unroll
            for (int c = 0; c < 16; ++c) dorig = lane == c ? g3[c] : dorig;
            eig::cholesky_lt16_regs(g3, dorig, me, LT, pinv, misc, rowbuf);
          }
          gphase ^= 1;
          tc_fence_before();
          G::sync();
          float t[16];
          eig::forward_subst16(u, LT, pinv, t);
#pragma unroll
          for (int c = 0; c < 16; ++c) u[c] = t[c];
          eig::forward_subst16(y, LT, pinv, t);
#pragma unroll
          for (int c = 0; c < 16; ++c) y[c] = t[c];
        }
        FPHASE_END(FP_SUBST);
        if ((P.debug & 15) == 6 && it == P.max_iter - 1) break;            // development: output u after `it` updates
        if ((P.debug & 15) == 7 && it == P.max_iter - 1) {                 // ... or its image y
#pragma unroll
          for (int c = 0; c < 16; ++c) u[c] = y[c];
          break;
        }
        // (the barrier after the next operand write orders the reads of LT / pinv before the next factorisation)
      }
      if (!need_rotate) break;
      {
        float* Sm = LT;                              // the factor is not applied in this iteration
        int* ord = reinterpret_cast<int*>(red);      // [16]
        G::sync();                                   // everyone has read Hs / LT
        eig::jacobi_impl<false, G>(Hs, Sm, ld, kMB, kRitzSweeps, rowbuf);
        if (ct < kMB) {
          const float ta = Hs[ct * ld + ct];
          int rank = 0;
          for (int b2 = 0; b2 < kMB; ++b2) {
            const float tb = Hs[b2 * ld + b2];
            rank += (tb > ta || (tb == ta && b2 < ct)) ? 1 : 0;
          }
          ord[rank] = ct;
        }
        G::sync();
        Gs[ct] = Sm[(ct >> 4) * ld + ord[ct & 15]];  // W[a][c], columns in eigenvalue order (kCompute = 256 entries)
        G::sync();
        auto rotate = [&](float (&v)[16]) {
          float t[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) t[c] = 0.f;
#pragma unroll
          for (int a = 0; a < 16; ++a) {
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
              const float4 w = *reinterpret_cast<const float4*>(Gs + a * ld + 4 * qq);
              t[4 * qq] = fmaf(v[a], w.x, t[4 * qq]);
              t[4 * qq + 1] = fmaf(v[a], w.y, t[4 * qq + 1]);
              t[4 * qq + 2] = fmaf(v[a], w.z, t[4 * qq + 2]);
              t[4 * qq + 3] = fmaf(v[a], w.w, t[4 * qq + 3]);
            }
          }
#pragma unroll
          for (int c = 0; c < 16; ++c) v[c] = t[c];
        };
        rotate(u);
        rotate(y);
        G::sync();                                   // Gs / red / LT are free again
      }
      }

      // ---- output: the basis, the projected operator, the verdict
      if (valid) {
        float4* uo = reinterpret_cast<float4*>(P.U + static_cast<size_t>(row0 + row) * kMB);
        uo[0] = make_float4(u[0], u[1], u[2], u[3]);
        uo[1] = make_float4(u[4], u[5], u[6], u[7]);
        uo[2] = make_float4(u[8], u[9], u[10], u[11]);
        uo[3] = make_float4(u[12], u[13], u[14], u[15]);
      }
      {
        const int a = ct >> 4, c = ct & 15;
        P.H[static_cast<size_t>(s) * 256 + ct] = (a < m && c < m) ? ((P.debug & 512) ? LT : ((P.debug & 256) ? Gs : Hs))[a * ld + c] : 0.f;
      }
      if (ct == 0) {
        P.iters[s] = it;
        P.info[s] = fired ? 1 : 0;
      }
      // the affinity in TMEM is dead: the Gram of the next image may overwrite it
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh.tmem_empty);
      G::sync();   // rq / Hs may be overwritten by the next segment
      FPHASE_END(FP_OUTPUT);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace fused
}  // namespace msvit

#ifdef FUSED_PROFILE
extern "C" int msvit_fused_profile(unsigned long long* host_out, int reset) {
  using namespace msvit::fused;
  cudaError_t e = cudaMemcpyFromSymbol(host_out, g_fused_cycles, sizeof(unsigned long long) * FP_COUNT);
  if (e != cudaSuccess) return static_cast<int>(e);
  if (reset) {
    unsigned long long z[FP_COUNT] = {};
    e = cudaMemcpyToSymbol(g_fused_cycles, z, sizeof(z));
  }
  return static_cast<int>(e);
}
#endif

extern "C" int msvit_ncut_fused(const void* x, int x_dtype, float* deg, float* U, float* H, int32_t* iters,
                                int32_t* info, int64_t total_rows, int S, int N, int D, int mode, float gamma,
                                float scale, int block, int max_iter, float tol, float lam_floor, int n_converge,
                                msvit_stream_t stream_) {
  using namespace msvit;
  using namespace msvit::fused;
  if (!x || !deg || !U || !H || !iters || !info) return MSVIT_ERR_NULL;
  if (x_dtype != MSVIT_F32 && x_dtype != MSVIT_BF16) return MSVIT_ERR_MODE;
  if (mode < MSVIT_DIST_RBF || mode > MSVIT_DIST_NORMPROD) return MSVIT_ERR_MODE;
  if (S < 0 || N <= 0 || D <= 0 || total_rows < 0 || !(gamma > 0.f) || !(scale > 0.f)) return MSVIT_ERR_SHAPE;
  if (total_rows != static_cast<int64_t>(S) * N || total_rows > 0x7fffffffLL) return MSVIT_ERR_SHAPE;
  if (block != kMB || N <= kMB || round_up(N, 16) > kMaxT) return MSVIT_ERR_SHAPE;
  if (max_iter <= 0 || !(tol > 0.f) || n_converge < 0 || n_converge > kMB) return MSVIT_ERR_SHAPE;
  const bool f32 = x_dtype == MSVIT_F32;
  const int esz = f32 ? 4 : 2;
  if ((static_cast<int64_t>(D) * esz) % 16 != 0 || (reinterpret_cast<uintptr_t>(x) & 15) != 0) return MSVIT_ERR_ALIGN;
  if ((reinterpret_cast<uintptr_t>(U) & 15) != 0) return MSVIT_ERR_ALIGN;
  if (S == 0) return MSVIT_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);

  EncodeTiledFn enc = encode_fn();
  if (!enc) return MSVIT_ERR_DRIVER;

  Params P;
  P.deg = deg; P.U = U; P.H = H; P.iters = iters; P.info = info;
  P.S = S; P.N = N; P.T = round_up(N, 16);
  P.mode = mode;
  P.c2 = mode == MSVIT_DIST_COSINE ? kLog2e / gamma : kLog2e / (gamma * scale);
  P.k_step = f32 ? 32 : 64;
  P.n_kslices = ceil_div(D, P.k_step);
  P.stage_bytes = round_up(P.T * kSliceBytes, 1024);
  P.tail_rows = P.T > 128 ? P.T - 128 : P.T;
  P.m = block;
  P.kconv = n_converge > 0 ? n_converge : block;
  P.max_iter = max_iter;
  P.fast_iters = EIG_FAST_ITERS;
  P.tol = tol;
  P.lam_floor = lam_floor;
  {
    const char* dbg = getenv("MSVIT_FUSED_DEBUG");
    P.debug = dbg ? atoi(dbg) : 0;
  }

  // fp32 tokens under the rbf distance take the fp16 Gram: |xi - xj|^2 / scale is what the affinity depends on, so the
  // tokens that matter have elements of the order sqrt(scale / D); they are multiplied by the power of two that brings
  // that magnitude to about 16 (fp16 then spans 2^-18 .. 2^12 times it; larger elements saturate, which only happens for
  // tokens whose affinity to everything else underflows anyway).  MSVIT_FUSED_TF32=1 keeps kind::tf32.
  int in_mode = f32 ? kInTf32 : kInBf16;
  P.in_scale = 1.0f;
  if (f32 && mode == MSVIT_DIST_RBF && !getenv("MSVIT_FUSED_TF32")) {
    const float mag = sqrtf(scale / static_cast<float>(D));
    if (mag > 1e-30f && mag < 1e30f) {
      in_mode = kInFp16;
      P.in_scale = exp2f(rintf(log2f(16.0f / mag)));
      P.c2 /= P.in_scale * P.in_scale;   // the Gram entries and the row norms carry in_scale^2
    }
  }

  const size_t fixed = 1024 + kUopBytes + ((sizeof(Shared) + 15) & ~size_t(15)) +
                       static_cast<size_t>(make_eig_layout().total) * sizeof(float);
  const size_t kMaxSmem = 227 * 1024;
  int stages = static_cast<int>((kMaxSmem - fixed) / P.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return MSVIT_ERR_SHAPE;
  P.stages = stages;
  const size_t smem = fixed + static_cast<size_t>(stages) * P.stage_bytes;

  CUtensorMap tm_full, tm_tail;
  int rc = make_map(enc, &tm_full, x, f32, total_rows, D, 128);
  if (rc != MSVIT_OK) return rc;
  rc = make_map(enc, &tm_tail, x, f32, total_rows, D, P.tail_rows);
  if (rc != MSVIT_OK) return rc;

  const int grid = S < sm_count() ? S : sm_count();
  cudaError_t e;
  if (in_mode == kInFp16) {
    e = cudaFuncSetAttribute(ncut_fused_kernel<kInFp16>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return cuda_status(e);
    ncut_fused_kernel<kInFp16><<<grid, kThreads, smem, stream>>>(tm_full, tm_tail, P);
  } else if (in_mode == kInTf32) {
    e = cudaFuncSetAttribute(ncut_fused_kernel<kInTf32>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return cuda_status(e);
    ncut_fused_kernel<kInTf32><<<grid, kThreads, smem, stream>>>(tm_full, tm_tail, P);
  } else {
    e = cudaFuncSetAttribute(ncut_fused_kernel<kInBf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return cuda_status(e);
    ncut_fused_kernel<kInBf16><<<grid, kThreads, smem, stream>>>(tm_full, tm_tail, P);
  }
  return cuda_status(cudaGetLastError());
}
