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

#define B2VS_VERSION 200

/* status codes */
#define B2VS_OK 0
#define B2VS_EINVAL (-1)  /* bad argument */
#define B2VS_ECUDA (-2)   /* CUDA runtime / driver error */
#define B2VS_ENOMEM (-3)  /* device allocation failed */
#define B2VS_EUNSUP (-4)  /* unsupported combination (e.g. k too large for the fused path) */

/* metric: scores are returned best-first. L2 = squared euclidean, ascending;
 * IP = inner product, descending (FAISS IndexFlatIP convention);
 * COSINE = cosine DISTANCE 1 - cos(q, x), ascending - scikit-learn's metric='cosine', the
 * reference's CPU baseline (Attempt_1/VectorSearch_QuestionRetrieval.ipynb:L878).  A cosine index
 * owns an L2-normalised copy of the rows (x / max(||x||, 1e-12), rounded once into the row
 * dtype) and normalises every query batch the same way: it is the IP engine on unit vectors.
 * Callers whose rows are already unit-norm (the encoder hand-off) should ask for IP instead and
 * keep their rows borrowed. */
#define B2VS_METRIC_L2 0
#define B2VS_METRIC_IP 1
#define B2VS_METRIC_COSINE 2

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
                             min(128, ..) on shapes without the grouped scan).  Needs the borrowed source
                             rows: B2VS_EINVAL when the index was loaded without rows_for_refine */
  int32_t n_splits;       /* flat: force the number of db splits (0 = heuristic) */
  int32_t flags;          /* bit 0: time the dominant kernel with CUDA events (see stats.kernel_ms) */
} b2vs_search_params;
#define B2VS_FLAG_TIME_KERNEL 1
#define B2VS_FLAG_TC_SINGLE 2   /* flat: force the single-CTA (cta_group::1) kernel */
#define B2VS_FLAG_TC_PAIR 4     /* flat: force the CTA-pair (cta_group::2) kernel */
#define B2VS_FLAG_EPI2 16       /* flat, k <= 128: two epilogue warp groups (measurement switch; slower in every case measured) */
#define B2VS_FLAG_GRAPH 8       /* IVF, nq <= 64, k (and k*refine_ratio) <= 128: replay the call as one
                                   CUDA graph (captured on the second call of a signature
                                   (nq, k, dtype, n_probes, refine_ratio); env B2VS_GRAPH=1/0 forces it
                                   on / off for every call).  IVF batches of 2048 queries and more are
                                   replayed as graphs WITHOUT the flag (unless B2VS_GRAPH=0 or
                                   B2VS_FLAG_TIME_KERNEL is set): queries and results pass through
                                   library-owned staging buffers, any pointers may be given.  Large IVF
                                   batches also use an internal side stream (joined before the call
                                   returns to the caller's stream order) */

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
  double kernel_ms;        /* device time of the dominant kernel (the fused distance kernel's full
                              pass / the list-scan kernel alone) when B2VS_FLAG_TIME_KERNEL was set,
                              else 0 */
  double distinct_bytes;   /* IVF grouped scans: bytes of the DISTINCT probed lists - what the batch
                              has to read from HBM once (algo_bytes counts a list once per probing
                              query); 0 on the per-(query, probe) paths */
} b2vs_search_stats;

const char* b2vs_last_error(void);
int b2vs_version(void);
int b2vs_device_count(int* count);
/* Re-reads the B2VS_* environment switches (A/B knobs; they are otherwise read once, at first
 * use, so nothing on the search path calls getenv).  For tests and measurement tools. */
int b2vs_reload_env(void);
/* Own bounds checking (compute-sanitizer is unavailable on the GPU pool): in a process started
 * with B2VS_CANARY=1 every device buffer the library allocates sits between two 256-byte guard
 * zones; this call reads all of them back: *n_buffers = live buffers checked, *n_corrupt = buffers
 * whose guards were overwritten (b2vs_last_error names one).  B2VS_EUNSUP when the mode is off. */
int b2vs_debug_check_canaries(int* n_buffers, int* n_corrupt);

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

/* k nearest rows for each of the nq queries ([nq, dim], `q_dtype`, device memory).  `dim` must
 * equal the index dimension (B2VS_EINVAL otherwise: a narrower matrix would be read out of
 * bounds, a wider one would silently give wrong neighbours).
 * Searches on ONE index must be serialised by the caller: workspaces, captured graphs and the
 * statistics live in the index (different indexes / devices may be searched concurrently).
 * k <= 128 on every index kind (fused in-kernel top-k); flat, IVF-Flat and (grouped-scan shapes:
 * dsub 2/4/8, dim % 64 == 0) IVF-PQ indexes also serve 128 < k <= 2048 (append + radix-select
 * paths, which synchronise `stream`).
 * out_d [nq, k] float32 and out_i [nq, k] int64 are caller-owned DEVICE buffers; missing
 * results are (inf | -inf, -1).  Asynchronous on `stream`.  `params` may be NULL. */
int b2vs_search(b2vs_index* index, const void* queries, int q_dtype, int nq, int dim, int k,
                const b2vs_search_params* params, float* out_d, int64_t* out_i, void* stream);

/* Same call with HOST buffers: copies queries H2D, searches, copies results D2H and
 * synchronises `stream` before returning (the reference's cuVS calls return host arrays
 * through pylibraft's copy_to_host hook, improved_multi_gpu_rag.py:114). */
int b2vs_search_host(b2vs_index* index, const void* queries_host, int q_dtype, int nq, int dim,
                     int k, const b2vs_search_params* params, float* out_d_host,
                     int64_t* out_i_host, void* stream);

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

/* ---------------------------------------------------------------------------------------------
 * Cross-shard exchange: the NCCL communicator and the collectives of the
 * sharded search live behind the C ABI, so a caller of this header gets the exchange step too
 * (the reference concatenates per-GPU results on the host: improved_multi_gpu_rag.py:251-277,
 * cuvs-2gpu-main.ipynb:L1806-1832).  NCCL is loaded at run time (dlopen "libnccl.so.2"; the copy
 * already loaded by the host process - e.g. PyTorch's - is reused); B2VS_EUNSUP if it is absent.
 *
 * One b2vs_comm per rank (= per GPU).  All collective calls below must be made by every rank of
 * the communicator in the same order; they are asynchronous on `stream`. */
typedef struct b2vs_comm b2vs_comm;
#define B2VS_UNIQUE_ID_BYTES 128
/* rank 0: fill a 128-byte id and hand it to the other ranks out of band (torch.distributed
 * broadcast, MPI, a file ...); then every rank calls b2vs_comm_init_rank with the same bytes. */
int b2vs_comm_unique_id(void* id128);
int b2vs_comm_init_rank(int dev, int n_ranks, int rank, const void* id128, b2vs_comm** out);
/* single process driving n GPUs (the reference's thread-per-GPU mode): comms[i] is the
 * communicator of rank i on device devs[i] */
int b2vs_comm_init_all(int n, const int* devs, b2vs_comm** comms);
int b2vs_comm_info(const b2vs_comm* comm, int* n_ranks, int* rank, int* dev);
int b2vs_comm_destroy(b2vs_comm* comm);

/* Row range [*begin, *end) of rank `rank` when `n` items are split into `n_parts` contiguous
 * near-equal parts (remainder to the first parts): the reference's 'even' strategy
 * (gpu_resource_manager.py:190-202), used for database rows AND for the query slices below. */
int b2vs_partition_even(int64_t n, int n_parts, int rank, int64_t* begin, int64_t* end);

/* Query all-gather: rank r holds its slice q_local = rows partition_even(nq_total, n_ranks)[r] of
 * the batch; afterwards q_all [nq_total, dim] is complete on every rank (each rank uploads 1/G of
 * the batch over PCIe and NVLink carries the rest). */
int b2vs_allgather_queries(b2vs_comm* comm, const void* q_local, int q_dtype, int nq_total, int dim,
                           void* q_all, void* stream);
/* Per-shard top-k all-gather: d_local / i_local [nq, k] -> d_all / i_all [n_ranks, nq, k] on every
 * rank, both buffers in ONE NCCL group (one collective launch). */
int b2vs_allgather_topk(b2vs_comm* comm, const float* d_local, const int64_t* i_local, int nq, int k,
                        float* d_all, int64_t* i_all, void* stream);
/* All-gather + b2vs_merge_topk: every rank ends with the global top-k_out of all nq queries
 * (out_d / out_i [nq, k_out]).  k <= 128. */
int b2vs_allgather_merge_topk(b2vs_comm* comm, const float* d_local, const int64_t* i_local, int nq,
                              int k, int k_out, int descending, float* out_d, int64_t* out_i,
                              void* stream);
/* All-to-all + merge: rank r receives from every rank the lists of ITS query slice
 * partition_even(nq, n_ranks)[r] and merges only those: out_d / out_i [slice rows, k_out].  Per
 * rank this moves and merges 1/G of what the all-gather variant does; the batch's answer is then
 * distributed like its queries were (rank r answers the queries rank r uploaded). */
int b2vs_exchange_merge_topk(b2vs_comm* comm, const float* d_local, const int64_t* i_local, int nq,
                             int k, int k_out, int descending, float* out_d, int64_t* out_i,
                             void* stream);
/* In-place element-wise MIN all-reduce of a float vector (per-query threshold exchange). */
int b2vs_allreduce_min_f32(b2vs_comm* comm, float* values, int64_t n, void* stream);

/* Collective, once per (communicator, index) after the build: the ranks agree on the size of the
 * smallest shard, from which sharded flat searches derive ONE pass schedule for every rank (the
 * threshold exchange sits between the passes).  Synchronises `stream`.  Without it
 * b2vs_search_sharded still works, with private per-shard thresholds. */
int b2vs_comm_register_index(b2vs_comm* comm, b2vs_index* index, void* stream);

/* The whole sharded step behind one call: all-gather of the query slices, local search of this
 * rank's shard - flat indexes pool the samples of their sampled passes (all-gather of each rank's k
 * best sampled scores, k-th best of the union = every shard's threshold; B2VS_SAMPLE_UNION=0: MIN
 * all-reduce of the per-shard k-th scores), so every shard runs its full pass against the GLOBAL
 * k-th bound - then all-to-all + merge.
 * q_local: this rank's query slice (device); out_d / out_i [slice rows, k] device.
 * The local per-shard result is NOT the shard's full top-k in this mode (rows that cannot be in
 * the global top-k are skipped), which is why it is not returned. */
int b2vs_search_sharded(b2vs_comm* comm, b2vs_index* index, const void* q_local, int q_dtype,
                        int nq_total, int dim, int k, const b2vs_search_params* params, float* out_d,
                        int64_t* out_i, void* stream);
/* Same with HOST buffers: H2D of the slice, the step above, D2H of the slice's answer, stream
 * synchronised before returning. */
int b2vs_search_sharded_host(b2vs_comm* comm, b2vs_index* index, const void* q_local_host, int q_dtype,
                             int nq_total, int dim, int k, const b2vs_search_params* params,
                             float* out_d_host, int64_t* out_i_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B2VS_H_ */
