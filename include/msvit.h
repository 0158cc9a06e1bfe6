/* msvit.h -- C ABI of the B200 (sm_100a) token-grouping hot path.
 *
 * Drop-in boundary for the clustering plugin layer of JophiArcana/multi-state-ViT.
 * Each entry point names the reference interface it replaces (paths relative to the
 * reference checkout).  The reference reaches this arithmetic through third-party
 * Python packages (ncut-pytorch==1.7.9, cuml~=24.10, fast_pytorch_kmeans); a maintainer
 * binds this library with ctypes behind ClusteringModule.forward.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is DEVICE memory unless stated otherwise;
 *   - the caller owns every buffer (inputs, outputs, workspace); nothing is allocated, freed
 *     or cached by the library, there is no global state, calls are re-entrant;
 *   - all work is enqueued on `stream` (a cudaStream_t) of the current device and no entry
 *     point synchronises with the host;
 *   - return value: 0 = ok, < 0 = argument check failed (MSVIT_ERR_*), > 0 = cudaError_t;
 *   - "segment" = one (image, parent cluster) group of tokens: rows seg_off[s] .. seg_off[s+1]-1
 *     of the flattened token matrix.  seg_off == NULL means S equal segments of N rows.
 *   - affinity storage: segment s is an n_s x lda_s row-major block, lda_s = (n_s + 3) & ~3, starting at
 *     element a_off[s] (a multiple of 4).  a_off == NULL means a_off[s] = s * N * ((N + 3) & ~3).
 */
#ifndef MSVIT_H_
#define MSVIT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSVIT_OK 0
#define MSVIT_ERR_NULL (-1)        /* a required pointer is NULL */
#define MSVIT_ERR_SHAPE (-2)       /* a size is out of the supported range */
#define MSVIT_ERR_ALIGN (-3)       /* pointer / row stride alignment requirement not met */
#define MSVIT_ERR_MODE (-4)        /* unknown dtype / distance mode */
#define MSVIT_ERR_WORKSPACE (-5)   /* workspace too small */
#define MSVIT_ERR_DRIVER (-6)      /* CUDA driver entry point (cuTensorMapEncodeTiled) unavailable */

#define MSVIT_F32 0
#define MSVIT_BF16 1

#define MSVIT_DIST_RBF 0       /* d = 1/2 |xi-xj|^2 / scale       sandbox/ncut_euclidean.py:19,23-29 */
#define MSVIT_DIST_COSINE 1    /* d = 1 - cos(xi,xj)              model/clustering/modeling_spectral.py:62-69 */
#define MSVIT_DIST_NORMPROD 2  /* d = (|xi||xj| - xi.xj) / scale  sandbox/test.py:108-110 */

#define MSVIT_DISC_KMEANS 0      /* Lloyd k-means on the embedding      model/clustering/modeling_spectral.py:90 */
#define MSVIT_DISC_AXIS_ALIGN 1  /* axis-aligned rotation (kway_ncut)   model/clustering/modeling_spectral.py:136-138 */

#define MSVIT_MAX_EIG_BLOCK 32 /* subspace block (ncut_dim + oversampling) upper bound */

typedef void* msvit_stream_t; /* cudaStream_t */

int msvit_version(void);
const char* msvit_error_string(int code);

/* Pairwise affinity + NCut degree.
 * Replaces the affinity stage inside NCUT.fit_transform (call sites
 * model/clustering/modeling_spectral.py:86,109,185,255; closed form sandbox/test.py:108-114):
 *   A_ij = exp(-d_ij / gamma),  deg_i = sum_j A_ij,   per segment.
 * x        [total_rows, D] row-major, MSVIT_F32 (tensor cores read it as TF32) or MSVIT_BF16.
 *          Row stride D*elsize must be a multiple of 16 bytes, x 16-byte aligned.
 * A        affinity blocks (layout above), may be NULL (only deg is produced).
 * deg      [total_rows]
 * N        max segment length (every n_s <= N). */
int msvit_affinity_degree(const void* x, int x_dtype, float* A, float* deg, int64_t total_rows, int S, int N, int D,
                          int mode, float gamma, float scale, const int32_t* seg_off, const int64_t* a_off,
                          msvit_stream_t stream);

/* Top-k eigenpairs of the normalised affinity D^-1/2 A D^-1/2 per segment.
 * Replaces the eigensolve inside NCUT.fit_transform (default torch.svd_lowrank there;
 * exact closed form sandbox/test.py:114-118).  Block subspace iteration with Rayleigh-Ritz,
 * eigenvalues descending, eigenvector sign canonical (largest-|entry| positive, ties -> lowest row).
 * V        [total_rows, k] (row i = embedding of token i inside its segment)
 * lam      [S, k]
 * iters    [S] iterations used (may be NULL)
 * block    subspace width m, k <= m <= MSVIT_MAX_EIG_BLOCK, m % 4 == 0
 * tol      residual tolerance |Abar v - lam v| <= tol for the k wanted pairs
 * lam_floor wanted pairs whose eigenvalue estimate is below lam_floor are exempt from the residual test
 *          (they are never clustered on when the eigenvalue threshold selects the children); 0 = none.
 * n_converge  only the leading n_converge pairs must meet the tolerance (0 = all k): with a fixed number of
 *          clusters K < k the k-means step reads V[:, :K] only (modeling_spectral.py:90), the remaining pairs
 *          are returned as the current Ritz estimates. */
int msvit_ncut_eig(const float* A, const float* deg, float* V, float* lam, int32_t* iters, int64_t total_rows, int S,
                   int N, int k, int block, int max_iter, float tol, float lam_floor, int n_converge,
                   const int32_t* seg_off, const int64_t* a_off, msvit_stream_t stream);

/* Fused affinity + NCut subspace iteration for whole images of N <= 208 tokens (uniform segments, N > block):
 * the affinity is produced and consumed in tensor memory and never written to global memory.
 * Operands of the Gram contraction: bf16 tokens as they are; fp32 tokens rounded to an 11-bit significand on the way to
 * the tensor core -- under MSVIT_DIST_RBF as fp16 after a power-of-two scaling derived from `scale` (elements more than
 * 2^12 times sqrt(scale / D) saturate: such a token is far from every other one either way), otherwise as TF32; fp32
 * accumulation and fp32 arithmetic everywhere else.  The environment variable MSVIT_FUSED_TF32=1 forces TF32.
 * Replaces NCUT.fit_transform as msvit_affinity_degree + msvit_ncut_eig do (same call sites, sandbox/test.py:108-118);
 * the Rayleigh-Ritz rotation of the result is done by msvit_ritz_kmeans.
 * x [S*N, D] tokens (MSVIT_F32 | MSVIT_BF16), deg [S*N] NCut degree (out),
 * U [S*N, 16] D-orthonormal basis of the converged subspace (out, 16-byte aligned),
 * H [S, 16, 16] projected operator U^T A U (out), iters [S], info [S] (1 = the leading n_converge columns met `tol`,
 * 0 = stopped at max_iter).  block must be 16. */
int msvit_ncut_fused(const void* x, int x_dtype, float* deg, float* U, float* H, int32_t* iters, int32_t* info,
                     int64_t total_rows, int S, int N, int D, int mode, float gamma, float scale, int block,
                     int max_iter, float tol, float lam_floor, int n_converge, msvit_stream_t stream);

/* Rayleigh-Ritz finish of msvit_ncut_fused + the k-means of msvit_kmeans in one kernel (uniform segments):
 * V [S*N, k] eigenvectors (eigenvalues descending, sign canonical), lam [S, k]; then Lloyd k-means on V[:, :K] with
 * K = n_clusters > 0 ? n_clusters : max(1, #{lam > eig_threshold}) (modeling_spectral.py:87-93), seeded from the
 * row of largest degree (method = MSVIT_DISC_KMEANS) or the axis-aligned rotation (MSVIT_DISC_AXIS_ALIGN).
 * labels [S*N] int32 and / or child [S*N] int64 (canonical ids; either may be NULL), n_child [S]. */
int msvit_ritz_kmeans(const float* U, const float* H, const int32_t* info, const float* deg, float* V, float* lam,
                      int32_t* labels, int64_t* child, int32_t* n_child, int64_t total_rows, int S, int N, int k,
                      int block, int n_converge, int n_clusters, float eig_threshold, int max_iter, int method,
                      msvit_stream_t stream);

/* Lloyd k-means on the leading columns of the spectral embedding, per segment.
 * Replaces cuml KMeans(n_clusters).fit_predict(ncut_x[:, :n_child]) (modeling_spectral.py:90), the
 * seeded variants (:130-133, :277-278) and n_child = sum(eigenvalues > threshold) (:87,92-93).
 * V          [total_rows, ldv] embedding, the first K_s columns are clustered
 * lam        [S, ldv] eigenvalues, used when n_clusters <= 0:  K_s = max(1, #{lam > eig_threshold})
 * weight     [total_rows] seeding weight (NCut degree): first centre = row argmax weight; NULL -> row 0
 * init       [S, Kmax, Kmax] caller-supplied initial centres (row c = centre c), NULL -> farthest-point seeding
 * labels     [total_rows] int32, canonical (first-occurrence order) ids local to the segment
 * n_child    [S] number of non-empty clusters
 * centres    [S, Kmax, Kmax] final centres in canonical order (may be NULL); Kmax = n_clusters>0 ? n_clusters : ldv */
int msvit_kmeans(const float* V, const float* lam, const float* weight, const float* init, int32_t* labels,
                 int32_t* n_child, float* centres, int64_t total_rows, int S, int N, int ldv, int n_clusters,
                 float eig_threshold, int max_iter, const int32_t* seg_off, msvit_stream_t stream);

/* Same interface with the discretisation method as an argument: MSVIT_DISC_KMEANS (= msvit_kmeans) or
 * MSVIT_DISC_AXIS_ALIGN, the axis-aligned rotation of the embedding followed by argmax
 * (replaces ncut_pytorch.kway_ncut, call sites modeling_spectral.py:136-138, modeling_axisalign.py:35-36; Yu & Shi
 * 2003; at most 16 columns, no `init` / `centres`). */
int msvit_discretise(const float* V, const float* lam, const float* weight, const float* init, int32_t* labels,
                     int32_t* n_child, float* centres, int64_t total_rows, int S, int N, int ldv, int n_clusters,
                     float eig_threshold, int max_iter, int method, const int32_t* seg_off, msvit_stream_t stream);

/* Cluster-mean pooling of tokens into multi-state tokens.
 * Replaces the per-label mean loops (modeling_spectral.py:125-127, :271-273).
 * x [B, N, D] (MSVIT_F32 or MSVIT_BF16), labels [B, N] int64 (ids outside [0,K) are ignored)
 * pooled [B, K, D] fp32 (empty cluster -> 0), counts [B, K] int32. */
int msvit_pool(const void* x, int x_dtype, const int64_t* labels, float* pooled, int32_t* counts, int B, int N, int D,
               int K, msvit_stream_t stream);

/* Hierarchy bookkeeping (modeling_spectral.py:80-84,91-94; caller msvitencoder.py:491-499).
 * msvit_build_segments: parent_indices [B, N] int64 with values in [0, P) ->
 *   perm [B*N] int32 (flat source row of sorted row j: tokens grouped by (image, parent), stable),
 *   seg_off [B*P + 1] int32 (segment b*P + p), a_off [B*P + 1] int64.
 * msvit_gather_rows: xs[j, :] = x[perm[j], :]  (dtype preserved).
 * msvit_compose_labels: child[perm[j]] = (sum of n_child of earlier parents of the image) + labels_sorted[j].
 *   perm == NULL means identity (single parent). */
int msvit_build_segments(const int64_t* parent_indices, int32_t* perm, int32_t* seg_off, int64_t* a_off, int B, int N,
                         int P, msvit_stream_t stream);
int msvit_gather_rows(const void* x, int x_dtype, const int32_t* perm, void* xs, int64_t total_rows, int D,
                      msvit_stream_t stream);
int msvit_compose_labels(const int32_t* labels_sorted, const int32_t* n_child, const int32_t* perm,
                         const int32_t* seg_off, int64_t* child, int B, int N, int P, msvit_stream_t stream);

/* Dataset-level k-means over a row shard of a feature matrix (one Lloyd iteration = assign, sort, accumulate,
 * [all-reduce of `packed` by the caller when rows are sharded over GPUs], finalize).
 * Replaces the flattened-batch clustering of model/clustering/modeling_spectral.py:254-256 and the
 * KMeans(n_clusters).fit_predict call sites (:90, :130-133) at dataset scale (BASELINE.json configs[4]).
 *
 * msvit_gkm_assign: labels[i] = argmin_c |c|^2 - 2 x_i.c (ties -> lowest c), on the tensor cores.
 *   x [n, D] and centroids_op [k, D] share x_dtype (MSVIT_F32 is read as TF32, rounded to nearest);
 *   labels [n] int32; best [n] = the minimal score (may be NULL).  D*elsize % 16 == 0.
 * msvit_gkm_sort: stable counting sort of row ids by label: perm [n] (rows grouped by label, ascending row id
 *   inside a label), seg_off [k+1].  workspace: msvit_gkm_workspace_bytes(n, k) bytes.  k <= 10000.
 * msvit_gkm_accumulate: packed [k, D+1] fp32: columns [0, D) = sum of the member rows (fixed order, no atomics),
 *   column D = member count.  workspace (msvit_gkm_accumulate_workspace_bytes(k, D) bytes, may be NULL): partial sums
 *   of the runs every centroid's rows are cut into, so that very unequal clusters do not serialise on one CTA.
 * msvit_gkm_finalize: centroids [k, D] fp32 (in/out) = sum / count where count > 0 (an empty cluster keeps its
 *   centre); centroids_op [k, D] in op_dtype (may be NULL) = the tensor-core operand copy. */
int msvit_gkm_assign(const void* x, int x_dtype, const void* centroids_op, int32_t* labels, float* best, int64_t n,
                     int k, int D, msvit_stream_t stream);
size_t msvit_gkm_workspace_bytes(int64_t n, int k);
int msvit_gkm_sort(const int32_t* labels, int64_t n, int k, int32_t* perm, int32_t* seg_off, void* workspace,
                   size_t workspace_bytes, msvit_stream_t stream);
size_t msvit_gkm_accumulate_workspace_bytes(int k, int D);
int msvit_gkm_accumulate(const void* x, int x_dtype, const int32_t* perm, const int32_t* seg_off, float* packed,
                         int64_t n, int k, int D, void* workspace, size_t workspace_bytes, msvit_stream_t stream);
int msvit_gkm_finalize(const float* packed, float* centroids, void* centroids_op, int op_dtype, int k, int D,
                       msvit_stream_t stream);

/* Cluster-restricted attention mask of the multi-state encoder.
 * Replaces MultiStateViTEncoderBackbone._construct_attention_mask
 * (model/multistate_encoder/modeling_msvitencoder.py:426-467).
 * cluster_indices [B, N] int64 (per-image contiguous ids, as produced by the clustering module), C = the number of
 * transmitter/receiver pairs = max cluster count over the batch (the reference reads it on the host, :451).
 * mask [B, L, L] bytes (0/1; torch.bool layout), L = 2C + N, sequence order [T_0, R_0, .., T_{C-1}, R_{C-1}, tokens];
 * 4-byte aligned. */
int msvit_attention_mask(const int64_t* cluster_indices, uint8_t* mask, int B, int N, int C, msvit_stream_t stream);

/* Cluster-compressed attention statistics: the transmitter sums of compress_tokens_with_cluster_indices
 * (model/multistate_encoder/modeling_msvitencoder.py:182-186):
 *   out[b, h, q, c] = sum over the keys k of cluster c of attn[b, h, q, k].
 * attn [B, H, N, N] fp32 (16-byte aligned), cluster_indices [B, N] int64, out [B, H, N, C] fp32, C <= 64.
 * The receiver means of the same function (:187-190) are msvit_pool on the [B*H, N, N] view of attn. */
int msvit_cluster_key_sums(const float* attn, const int64_t* cluster_indices, float* out, int B, int H, int N, int C,
                           msvit_stream_t stream);

/* Both statistics of compress_tokens_with_cluster_indices (modeling_msvitencoder.py:182-190) in ONE pass over attn:
 *   transmitter[b, h, q, c] = sum over the keys k of cluster c of attn[b, h, q, k]            [B, H, N, C]
 *   receiver[b, h, c, k]    = mean over the queries q of cluster c of attn[b, h, q, k]        [B, H, C, N]
 * (an empty cluster gives a receiver row of 0 where the reference's 0/0 gives NaN).  attn fp32, 16-byte aligned;
 * N <= 256 and C <= 16 (the range where one pass beats two) -- MSVIT_ERR_SHAPE beyond that: use
 * msvit_cluster_key_sums + msvit_pool there (C <= 64, any N). */
int msvit_cluster_attention_stats(const float* attn, const int64_t* cluster_indices, float* transmitter,
                                  float* receiver, int B, int H, int N, int C, msvit_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MSVIT_H_ */
