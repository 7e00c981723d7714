/*
 * b2vs — C ABI of the B200-native vector-search hot path (libb2vs.so).
 *
 * This is the drop-in boundary for the per-GPU ANN calls the reference makes into
 * un-vendored cuVS 25.06 / FAISS.  Every entry point takes plain pointers
 * and sizes (device pointers unless the name ends in _host), an explicit device ordinal and
 * a cudaStream_t passed as void*; none relies on the ambient CUDA device.  All functions
 * return 0 on success or a negative B2VS_E* code; b2vs_last_error() gives the message
 * (thread-local).  There is no CPU fallback anywhere behind this header.
 *
 * Reference interfaces replaced (file:line under the reference repo):
 *   cuvs.neighbors.ivf_flat.build   Attempt_1/index_building_coordinator.py:392-396,
 *                                   Latest/cuVS-2-gpu/improved_multi_gpu_rag.py:126-130
 *   cuvs.neighbors.ivf_pq.build     Attempt_1/index_building_coordinator.py:398-404,
 *                                   Latest/cuVS-2-gpu/improved_multi_gpu_rag.py:132-138
 *   cuvs.neighbors.*.search         Latest/cuVS-2-gpu/improved_multi_gpu_rag.py:225-233,
 *                                   Latest/cuVS-2-gpu/old/cuvs-2gpu-main.ipynb:L1801
 *   faiss.IndexFlatIP/L2.search     Latest/faiss-main.ipynb cells 9-10 (exact search)
 *   host merge np.argsort(...)[:k]  Latest/cuVS-2-gpu/improved_multi_gpu_rag.py:266-275,
 *                                   Latest/cuVS-2-gpu/old/cuvs-2gpu-main.ipynb:L1806-1832
 *   id fix-up  ids + shard_start    Latest/cuVS-2-gpu/old/cuvs-2gpu-main.ipynb:L1803
 *                                   (correct offset: embedding_distribution_manager.py:25)
 */
#ifndef B2VS_H_
#define B2VS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2VS_VERSION 100

/* status codes */
#define B2VS_OK 0
#define B2VS_EINVAL (-1)  /* bad argument */
#define B2VS_ECUDA (-2)   /* CUDA runtime / driver error */
#define B2VS_ENOMEM (-3)  /* device allocation failed */
#define B2VS_EUNSUP (-4)  /* unsupported combination (e.g. k too large for the fused path) */

/* metric: scores are returned best-first. L2 = squared euclidean, ascending;
 * IP = inner product, descending (FAISS IndexFlatIP convention). */
#define B2VS_METRIC_L2 0
#define B2VS_METRIC_IP 1

/* element type of a database / query matrix (row-major [n, dim]) */
#define B2VS_F32 0
#define B2VS_F16 1
#define B2VS_BF16 2

/* index kinds */
#define B2VS_KIND_FLAT 0
#define B2VS_KIND_IVF_FLAT 1
#define B2VS_KIND_IVF_PQ 2

typedef struct b2vs_index b2vs_index;

typedef struct b2vs_ivf_params {
  int32_t n_lists;        /* coarse clusters (cuVS IndexParams.n_lists) */
  int32_t kmeans_iters;   /* Lloyd iterations, default 20 (cuVS kmeans_n_iters) */
  float train_fraction;   /* fraction of rows used to train, default 0.5 (cuVS kmeans_trainset_fraction) */
  int32_t pq_dim;         /* IVF-PQ: number of sub-quantizers M (cuVS pq_dim); 0 for IVF-Flat */
  int32_t pq_bits;        /* IVF-PQ: bits per code, only 8 is supported */
  uint64_t seed;          /* sampling seed */
} b2vs_ivf_params;

typedef struct b2vs_search_params {
  int32_t n_probes;       /* IVF: lists scanned per query (cuVS SearchParams.n_probes, default 20);
                             clamped to min(n_lists, 2048) */
  int32_t refine_ratio;   /* IVF-PQ: exact re-rank of min(2048, refine_ratio*k) ADC candidates (0/1 = off;
                             min(128, ..) on shapes without the grouped scan) */
  int32_t n_splits;       /* flat: force the number of db splits (0 = heuristic) */
  int32_t flags;          /* bit 0: time the dominant kernel with CUDA events (see stats.kernel_ms) */
} b2vs_search_params;
#define B2VS_FLAG_TIME_KERNEL 1
#define B2VS_FLAG_TC_SINGLE 2   /* flat: force the single-CTA (cta_group::1) kernel */
#define B2VS_FLAG_TC_PAIR 4     /* flat: force the CTA-pair (cta_group::2) kernel */
#define B2VS_FLAG_GRAPH 8       /* IVF, nq <= 64, k (and k*refine_ratio) <= 128: replay the call as one
                                   CUDA graph (captured on the second call of a signature
                                   (nq, k, dtype, n_probes, refine_ratio); env B2VS_GRAPH=1/0 forces it
                                   on / off for every call) */

typedef struct b2vs_index_info {
  int32_t kind, device, metric, dtype, dim, n_lists, pq_dim, pq_bits;
  int64_t n_rows, id_offset;
  int64_t device_bytes;   /* memory owned by the index (excludes borrowed database rows) */
} b2vs_index_info;

/* timing / accounting of the most recent b2vs_search on this index (filled on request) */
typedef struct b2vs_search_stats {
  int32_t launches;        /* kernels launched by the call */
  int32_t n_splits;        /* flat: db splits used */
  int32_t grid;            /* CTAs of the dominant kernel */
  int32_t mean_candidates; /* IVF-Flat grouped scan: appended candidates per query (mean), else 0 */
  double algo_flops;       /* 2*Q*N*D for the distance contraction (flat / coarse) */
  double algo_bytes;       /* IVF: sum of probed list bytes actually scanned */
  double kernel_ms;        /* device time of the dominant kernel (fused distance / list scan) when
                              B2VS_FLAG_TIME_KERNEL was set, else 0 */
} b2vs_search_stats;

const char* b2vs_last_error(void);
int b2vs_version(void);
int b2vs_device_count(int* count);

/* Exact (brute-force) index over `db` ([n, dim], `dtype`) resident on device `dev`.
 * 16-bit databases with dim % 8 == 0 are BORROWED (the caller keeps them alive); fp32
 * databases are re-encoded into an owned bf16 hi/lo split so the tensor cores reproduce the
 * fp32 contraction to ~1e-5 relative.  Row norms are computed here.  Returned ids are
 * `row + id_offset` (the shard's start_index). */
int b2vs_bf_create(int dev, int metric, int dtype, int dim, const void* db, int64_t n,
                   int64_t id_offset, void* stream, b2vs_index** out);

/* IVF-Flat: GPU k-means coarse quantizer + lists holding the raw vectors (16-bit storage). */
int b2vs_ivfflat_build(int dev, int metric, int dtype, int dim, const void* db, int64_t n,
                       int64_t id_offset, const b2vs_ivf_params* params, void* stream,
                       b2vs_index** out);

/* IVF-PQ: coarse quantizer + per-subspace 256-entry codebooks on residuals, 8-bit codes.
 * `db` is additionally BORROWED for searches with refine_ratio > 1 (exact re-rank against the
 * original rows); callers that never refine may free it after the build. */
int b2vs_ivfpq_build(int dev, int metric, int dtype, int dim, const void* db, int64_t n,
                     int64_t id_offset, const b2vs_ivf_params* params, void* stream,
                     b2vs_index** out);

/* k nearest rows for each of the nq queries ([nq, dim], `q_dtype`, device memory).
 * k <= 128 on every index kind (fused in-kernel top-k); flat, IVF-Flat and (grouped-scan shapes:
 * dsub 2/4/8, dim % 64 == 0) IVF-PQ indexes also serve 128 < k <= 2048 (append + radix-select
 * paths, which synchronise `stream`).
 * out_d [nq, k] float32 and out_i [nq, k] int64 are caller-owned DEVICE buffers; missing
 * results are (inf | -inf, -1).  Asynchronous on `stream`.  `params` may be NULL. */
int b2vs_search(b2vs_index* index, const void* queries, int q_dtype, int nq, int k,
                const b2vs_search_params* params, float* out_d, int64_t* out_i, void* stream);

/* Same call with HOST buffers: copies queries H2D, searches, copies results D2H and
 * synchronises `stream` before returning (the reference's cuVS calls return host arrays
 * through pylibraft's copy_to_host hook, improved_multi_gpu_rag.py:114). */
int b2vs_search_host(b2vs_index* index, const void* queries_host, int q_dtype, int nq, int k,
                     const b2vs_search_params* params, float* out_d_host, int64_t* out_i_host,
                     void* stream);

/* Global top-k over per-shard results: d_all / i_all are [n_parts, nq, k_in] (each part sorted
 * best-first, ids already global); writes the best k_out per query.  Ties keep the lower part
 * first (stable), matching the reference's concatenate + argsort merge.  k_out <= 2048; for
 * k_out > 128 at most 16384 candidates per query (n_parts * k_in). */
int b2vs_merge_topk(int dev, const float* d_all, const int64_t* i_all, int n_parts, int nq,
                    int k_in, int k_out, int descending, float* out_d, int64_t* out_i,
                    void* stream);

/* Lloyd k-means on device data (used by the IVF builders; exposed for tests).
 * centroids [n_clusters, dim] float32 device (output), labels [n] int32 device (output, may be
 * NULL).  Assignment runs on the tensor cores (fused GEMM + arg-min). */
int b2vs_kmeans_fit(int dev, int dtype, int dim, const void* x, int64_t n, int n_clusters,
                    int iters, uint64_t seed, float* centroids, int32_t* labels, void* stream);

int b2vs_index_info_get(const b2vs_index* index, b2vs_index_info* info);
int b2vs_index_last_stats(const b2vs_index* index, b2vs_search_stats* stats);
/* IVF introspection for tests: list sizes [n_lists] int32 (device or host pointer = host) */
int b2vs_ivf_list_sizes_host(const b2vs_index* index, int32_t* sizes_host);
int b2vs_ivf_centroids_host(const b2vs_index* index, float* centroids_host);
int b2vs_index_destroy(b2vs_index* index);

/* Persistence of trained IVF-Flat / IVF-PQ indexes (the reference re-trains on every run and only
 * stores raw embeddings: cuvs-2gpu-main.ipynb cells 10/12).  One little-endian file per shard.
 * b2vs_index_load: `rows_for_refine` (may be NULL) is the shard's original [n, dim] device matrix,
 * borrowed for IVF-PQ refine; `id_offset` < 0 keeps the offset stored in the file. */
int b2vs_index_save(const b2vs_index* index, const char* path);
int b2vs_index_load(int dev, const char* path, const void* rows_for_refine, int64_t id_offset,
                    void* stream, b2vs_index** out);

/* Encoder hand-off: pooled (and L2-normalised) query rows from a bi-encoder's hidden states,
 * on the device, in the dtype the index wants -- replaces last_token_pool + F.normalize +
 * `.cpu().numpy()` of Latest/cuVS-2-gpu/old/generate_embeddings.py:11-21, :100-105 (and the
 * `.to(device)` before the search in cuvs-2gpu-main.ipynb cell 16).
 * hidden [batch, seq_len, dim] `dtype` row-major device; attention_mask [batch, seq_len] int64 device
 * or NULL (= all ones); out [batch, dim] `out_dtype` device.  pooling: LAST_TOKEN follows the
 * reference (left-padding test on the mask's last column, else row sum - 1 with python's negative
 * index wrap); MEAN is the sentence-transformers masked mean (prepare_dataset.py:149).
 * normalize != 0: x / max(||x||_2, 1e-12).  Arithmetic in fp32.  Asynchronous on `stream`; the
 * scratch is per calling thread, so one call at a time per thread. */
#define B2VS_POOL_LAST_TOKEN 0
#define B2VS_POOL_MEAN 1
int b2vs_pool_normalize(int dev, int dtype, const void* hidden, int batch, int seq_len, int dim,
                        const int64_t* attention_mask, int pooling, int normalize, int out_dtype,
                        void* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B2VS_H_ */
