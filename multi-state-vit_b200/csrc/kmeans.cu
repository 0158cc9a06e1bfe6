// Lloyd k-means on the spectral embedding, one CTA per segment, everything on-chip.
//
// Reference: cuml KMeans(n_clusters=n_child).fit_predict(ncut_x[:, :n_child]) at
// model/clustering/modeling_spectral.py:90, n_child = sum(eigenvalues > threshold) at :87,92-93,
// assignment = argmin cdist at :129, centre = label mean at :125-127, centroid-seeded variants at
// :130-133 and :277-278.  Deterministic restatement: farthest-point seeding, lowest-index tie breaks,
// empty cluster keeps its centre, first-occurrence relabelling (oracle/ncut_oracle.py:kmeans).
//
// The segment's embedding is staged once into shared memory (transposed: coordinate j of all points is contiguous,
// so a warp reads it without bank conflicts).  Distance + argmin run in one pass per point; the centre update is a
// segmented, atomic-free sum in a fixed order (thread (c, d, h) adds the members of cluster c in ascending token
// order over half h of the tokens, the halves are combined in a fixed order).
#include "common.cuh"

namespace msvit {
namespace km {

constexpr int kThreads = 128;
constexpr int kMaxK = MSVIT_MAX_EIG_BLOCK;

struct Params {
  const float* V;
  const float* lam;
  const float* weight;
  const float* init;
  int32_t* labels;
  int32_t* n_child;
  float* centres;
  const int32_t* seg_off;
  int S, N, ldv, n_clusters, Kmax, max_iter;
  float thr;
};

// block-wide argmax of (value, index) with ties -> lowest index; result broadcast to all threads
__device__ __forceinline__ int block_argmax(float v, int idx, float* sval, int* sidx) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
  if (lane == 0) { sval[warp] = v; sidx[warp] = idx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kThreads / 32; ++w)
      if (sval[w] > v || (sval[w] == v && sidx[w] < idx)) { v = sval[w]; idx = sidx[w]; }
    sidx[0] = idx;
  }
  __syncthreads();
  const int r = sidx[0];
  __syncthreads();
  return r;
}

// squared distance of point i (column i of the transposed embedding, row stride ldp) to centre c
__device__ __forceinline__ float sqdist(const float* __restrict__ pt, int ldp, int i, const float* __restrict__ c, int K) {
  float d = 0.f;
  for (int j = 0; j < K; ++j) {
    const float t = pt[j * ldp + i] - c[j];
    d = fmaf(t, t, d);
  }
  return d;
}

__global__ void __launch_bounds__(kThreads) kmeans_kernel(const Params P) {
  extern __shared__ __align__(16) float smem[];
  float* cen = smem;                                   // [kMaxK][kMaxK + 1]
  float* mind = cen + kMaxK * (kMaxK + 1);             // [N]
  int* lab = reinterpret_cast<int*>(mind + P.N);       // [N]
  int* map = lab + P.N;                                // [kMaxK]
  float* part = reinterpret_cast<float*>(map + kMaxK); // [2][kMaxK * kMaxK] partial centre sums
  int* pcnt = reinterpret_cast<int*>(part + 2 * kMaxK * kMaxK);  // [2][kMaxK] partial counts
  float* pts = reinterpret_cast<float*>(pcnt + 2 * kMaxK);       // [Kcap][ldp] transposed embedding
  const int ldp = P.N | 1;                             // odd row stride
  __shared__ float sval[kThreads / 32];
  __shared__ int sidx[kThreads / 32];
  __shared__ int s_changed;
  constexpr int LDC = kMaxK + 1;

  for (int s = blockIdx.x; s < P.S; s += gridDim.x) {
    const int row0 = P.seg_off ? P.seg_off[s] : s * P.N;
    const int n = P.seg_off ? P.seg_off[s + 1] - row0 : P.N;
    if (n <= 0) {
      if (threadIdx.x == 0) P.n_child[s] = 0;
      continue;
    }
    const float* __restrict__ Pt = P.V + static_cast<long long>(row0) * P.ldv;
    int K;
    if (P.n_clusters > 0) {
      K = P.n_clusters;
    } else {
      K = 0;
      for (int j = 0; j < P.ldv; ++j) K += P.lam[static_cast<long long>(s) * P.ldv + j] > P.thr ? 1 : 0;
      K = K < 1 ? 1 : K;
    }
    K = min(K, min(n, min(P.ldv, kMaxK)));
    // stage the K leading coordinates of the segment's points (coalesced global reads, transposed writes)
    __syncthreads();
    for (int e = threadIdx.x; e < n * P.ldv; e += kThreads) {
      const int i = e / P.ldv, j = e - i * P.ldv;
      if (j < K) pts[j * ldp + i] = Pt[e];
    }
    __syncthreads();

    // ---- seeding
    if (P.init) {
      for (int e = threadIdx.x; e < K * K; e += kThreads)
        cen[(e / K) * LDC + (e % K)] = P.init[(static_cast<long long>(s) * P.Kmax + e / K) * P.Kmax + (e % K)];
      __syncthreads();
    } else {
      int first = 0;
      if (P.weight) {
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int i = threadIdx.x; i < n; i += kThreads) {
          const float w = P.weight[row0 + i];
          if (w > bv) { bv = w; bi = i; }
        }
        first = block_argmax(bv, bi, sval, sidx);
        if (first < 0 || first >= n) first = 0;
      }
      for (int j = threadIdx.x; j < K; j += kThreads) cen[j] = pts[j * ldp + first];
      __syncthreads();
      for (int i = threadIdx.x; i < n; i += kThreads) mind[i] = sqdist(pts, ldp, i, cen, K);
      for (int c = 1; c < K; ++c) {
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int i = threadIdx.x; i < n; i += kThreads)
          if (mind[i] > bv) { bv = mind[i]; bi = i; }
        const int nxt = block_argmax(bv, bi, sval, sidx);
        for (int j = threadIdx.x; j < K; j += kThreads) cen[c * LDC + j] = pts[j * ldp + nxt];
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += kThreads) mind[i] = fminf(mind[i], sqdist(pts, ldp, i, cen + c * LDC, K));
      }
    }
    for (int i = threadIdx.x; i < n; i += kThreads) lab[i] = -1;
    __syncthreads();

    // ---- Lloyd
    for (int it = 0; it < P.max_iter; ++it) {
      if (threadIdx.x == 0) s_changed = 0;
      __syncthreads();
      int changed = 0;
      for (int i = threadIdx.x; i < n; i += kThreads) {
        float bd = INFINITY;
        int bc = 0;
        for (int c = 0; c < K; ++c) {
          const float d = sqdist(pts, ldp, i, cen + c * LDC, K);
          if (d < bd) { bd = d; bc = c; }
        }
        if (lab[i] != bc) { lab[i] = bc; changed = 1; }
      }
      if (changed) s_changed = 1;
      __syncthreads();
      if (!s_changed) break;
      // partial sums over the two halves of the tokens (ascending order inside a half), then a fixed-order combine
      const int nh = (n + 1) >> 1;
      for (int e = threadIdx.x; e < 2 * K * K; e += kThreads) {
        const int h = e / (K * K), r = e - h * K * K;
        const int c = r / K, d = r - c * K;
        const int i0 = h * nh, i1 = min(n, i0 + nh);
        const float* pd = pts + d * ldp;
        float sum = 0.f;
        int cnt = 0;
        for (int i = i0; i < i1; ++i)
          if (lab[i] == c) { sum += pd[i]; ++cnt; }
        part[h * kMaxK * kMaxK + r] = sum;
        if (d == 0) pcnt[h * kMaxK + c] = cnt;
      }
      __syncthreads();
      for (int r = threadIdx.x; r < K * K; r += kThreads) {
        const int c = r / K, d = r - c * K;
        const int cnt = pcnt[c] + pcnt[kMaxK + c];
        if (cnt > 0) cen[c * LDC + d] = (part[r] + part[kMaxK * kMaxK + r]) / static_cast<float>(cnt);
      }
      __syncthreads();
    }

    // ---- canonical ids: clusters renamed in order of first occurrence
    if (threadIdx.x == 0) {
      for (int c = 0; c < K; ++c) map[c] = -1;
      int next = 0;
      for (int i = 0; i < n && next < K; ++i)
        if (map[lab[i]] < 0) map[lab[i]] = next++;
      P.n_child[s] = next;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += kThreads) P.labels[row0 + i] = map[lab[i]];
    if (P.centres) {
      float* co = P.centres + static_cast<long long>(s) * P.Kmax * P.Kmax;
      for (int e = threadIdx.x; e < P.Kmax * P.Kmax; e += kThreads) co[e] = 0.f;
      __syncthreads();
      for (int e = threadIdx.x; e < K * K; e += kThreads) {
        const int c = e / K, d = e % K;
        if (map[c] >= 0) co[map[c] * P.Kmax + d] = cen[c * LDC + d];
      }
    }
    __syncthreads();
  }
}

}  // namespace km
}  // namespace msvit

extern "C" int msvit_kmeans(const float* V, const float* lam, const float* weight, const float* init,
                            int32_t* labels, int32_t* n_child, float* centres, int64_t total_rows, int S, int N,
                            int ldv, int n_clusters, float eig_threshold, int max_iter, const int32_t* seg_off,
                            msvit_stream_t stream_) {
  using namespace msvit;
  using namespace msvit::km;
  if (!V || !labels || !n_child) return MSVIT_ERR_NULL;
  if (n_clusters <= 0 && !lam) return MSVIT_ERR_NULL;
  if (S < 0 || N <= 0 || ldv <= 0 || total_rows < 0 || max_iter <= 0) return MSVIT_ERR_SHAPE;
  if (n_clusters > kMaxK || n_clusters > ldv || (n_clusters <= 0 && ldv > kMaxK)) return MSVIT_ERR_SHAPE;
  if (!seg_off && total_rows != static_cast<int64_t>(S) * N) return MSVIT_ERR_SHAPE;
  if (S == 0 || total_rows == 0) return MSVIT_OK;
  Params P;
  P.V = V; P.lam = lam; P.weight = weight; P.init = init; P.labels = labels; P.n_child = n_child;
  P.centres = centres; P.seg_off = seg_off;
  P.S = S; P.N = N; P.ldv = ldv; P.n_clusters = n_clusters; P.Kmax = n_clusters > 0 ? n_clusters : ldv;
  P.max_iter = max_iter; P.thr = eig_threshold;
  const int kcap = P.Kmax < kMaxK ? P.Kmax : kMaxK;   // coordinates staged per point
  const size_t smem = sizeof(float) * (kMaxK * (kMaxK + 1) + N + 2 * kMaxK * kMaxK + static_cast<size_t>(kcap) * (N | 1)) +
                      sizeof(int) * (N + kMaxK + 2 * kMaxK);
  if (smem > 200 * 1024) return MSVIT_ERR_SHAPE;
  cudaError_t e = cudaFuncSetAttribute(kmeans_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return cuda_status(e);
  const int grid = S < 16 * sm_count() ? S : 16 * sm_count();
  kmeans_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(stream_)>>>(P);
  return cuda_status(cudaGetLastError());
}
