// Internal definitions shared by the b2vs translation units (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/b2vs.h"

namespace b2vs {

using u64 = unsigned long long;

void set_error(const char* fmt, ...);

#define B2VS_CUDA(call)                                                                   \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      ::b2vs::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return (e__ == cudaErrorMemoryAllocation) ? B2VS_ENOMEM : B2VS_ECUDA;               \
    }                                                                                     \
  } while (0)

#define B2VS_CHECK(cond, code, ...)  \
  do {                               \
    if (!(cond)) {                   \
      ::b2vs::set_error(__VA_ARGS__); \
      return (code);                 \
    }                                \
  } while (0)

#define B2VS_TRY(expr)          \
  do {                          \
    int rc__ = (expr);          \
    if (rc__ != B2VS_OK) return rc__; \
  } while (0)

// RAII device switch: every entry point runs on the device it was given and restores the
// caller's ambient device (the reference relies on `with torch.cuda.device(g)`).
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// Bumped whenever a DevBuf frees or replaces its allocation: a captured CUDA graph of a search
// holds workspace pointers and is only replayed while this value is the one it was captured at.
uint64_t realloc_generation();
void note_realloc();

// Bounds checking of our own (compute-sanitizer is not available on the GPU pool): with
// B2VS_CANARY=1 every DevBuf allocation is wrapped in two 256-byte guard zones filled with a
// pattern, all live allocations are registered, and b2vs_debug_check_canaries() reads the guards
// back - any kernel that wrote just before or past one of the library's device buffers is caught
// (tests/test_gpu_canary.py runs every search path under it).
constexpr size_t kCanaryBytes = 256;
bool canary_enabled();
void canary_register(void* user_ptr, size_t bytes);
void canary_unregister(void* user_ptr);

// Growable device buffer (grow-only; steady-state searches allocate nothing).
struct DevBuf {
  void* ptr = nullptr;
  size_t bytes = 0;
  int reserve(size_t need) {
    if (need <= bytes) return B2VS_OK;
    note_realloc();
    release_raw();
    const size_t guard = canary_enabled() ? kCanaryBytes : 0;
    size_t want = (need + need / 8 + 255) & ~static_cast<size_t>(255);
    void* raw = nullptr;
    cudaError_t e = cudaMalloc(&raw, want + 2 * guard);
    if (e != cudaSuccess) {
      cudaGetLastError();
      want = (need + 255) & ~static_cast<size_t>(255);
      e = cudaMalloc(&raw, want + 2 * guard);
    }
    if (e != cudaSuccess) {
      cudaGetLastError();
      set_error("cudaMalloc(%zu bytes) failed: %s", need, cudaGetErrorString(e));
      ptr = nullptr;
      return B2VS_ENOMEM;
    }
    ptr = static_cast<char*>(raw) + guard;
    bytes = want;
    if (guard) {
      cudaMemset(raw, 0xA5, guard);
      cudaMemset(static_cast<char*>(ptr) + want, 0xA5, guard);
      canary_register(ptr, want);
    }
    return B2VS_OK;
  }
  void release_raw() {
    if (!ptr) return;
    const size_t guard = canary_enabled() ? kCanaryBytes : 0;
    if (guard) canary_unregister(ptr);
    cudaFree(static_cast<char*>(ptr) - guard);
    ptr = nullptr;
    bytes = 0;
  }
  void release() {
    if (ptr) note_realloc();
    release_raw();
  }
  template <class T> T* as() const { return reinterpret_cast<T*>(ptr); }
};

inline int elem_bytes(int dtype) { return dtype == B2VS_F32 ? 4 : 2; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

int sm_count(int dev);

// Every B2VS_* environment switch (A/B knobs for measurements and tests), read ONCE into this
// struct at first use: nothing on the search path calls getenv.  b2vs_reload_env() re-reads them
// (tests flip switches inside one process).
struct EnvConfig {
  int tc_group = 0;            // B2VS_TC_GROUP=1|2: force the single-CTA / CTA-pair flat kernel
  int pass_n = 0;              // B2VS_PASSES="256,16,1": tile strides of the flat passes
  int pass_s[3] = {0, 0, 0};
  bool no_qpad = false;        // B2VS_NO_QPAD: stage ragged query blocks through TMA's OOB fill
  bool ivf_no_rank = false;    // B2VS_IVF_NO_RANK: keep lists in id order (no size ranking)
  int grouped_cap = 0;         // B2VS_IVF_GROUPED_CAP: candidate-buffer capacity (power of two, 32..4096)
  int seed_rows = 0;           // B2VS_IVF_SEED_ROWS: seed-sample length (>= 32)
  int ivf_grouped = -1;        // B2VS_IVF_GROUPED=0|1: force per-item / grouped scans
  int work_chunk_tiles = 0;    // B2VS_WORK_CHUNK_TILES: row-range size of grouped work items
  bool debug_split = false;    // B2VS_DEBUG_SPLIT: print the work split
  int coarse_scan = -1;        // B2VS_COARSE_SCAN=0: always probe on the tensor cores
  int coarse_scan_maxq = -1;   // B2VS_COARSE_SCAN_MAXQ
  int coarse_scan_ctas = -1;   // B2VS_COARSE_SCAN_CTAS
  bool no_item_sort = false;   // B2VS_NO_ITEM_SORT: per-item scan without the list ordering
  int graph = -1;              // B2VS_GRAPH=0|1: never / always replay small IVF batches as a graph
  int tail_boxes = -1;         // B2VS_TAIL_BOXES=0: always load whole 256-row list tiles in the grouped IVF-Flat scan
  int plan_overlap = -1;       // B2VS_PLAN_OVERLAP=0: plan the main IVF pass in line instead of on the side stream
  int graph_maxq = 0;          // B2VS_GRAPH_MAXQ: largest batch replayed as a graph (default 64)
  int seed_mode = -1;          // B2VS_IVF_SEED=0|1: legacy seed kernels / seeds from the tensor-core pass
  int epi_groups = 0;          // B2VS_EPI_GROUPS=1|2: epilogue warp groups of the flat kernel (0 = heuristic)
  int sample_union = -1;       // B2VS_SAMPLE_UNION=0: sharded searches exchange k-th scores (MIN) instead of sampled top-k lists
  int a_quarters = -1;         // B2VS_A_QUARTERS=0: always load whole 128-row query blocks in the grouped IVF-Flat scan
  int raw_emit = -1;           // B2VS_RAW_EMIT=0: the fused kernel's items sort and emit their own top-k lists
  int k0_debug = 0;            // B2VS_K0_DEBUG: 1 = the fused kernel's items skip their final sort + emission (timing only)
  int pq_debug = 0;            // B2VS_PQ_DEBUG: role-skipping bits of pq_tc_kernel (timing experiments only)
  int work_epi = 0;            // B2VS_WORK_EPI=1|2: epilogue groups of every work-table launch (0 = per call site)
  int seed_lists = 0;          // B2VS_IVF_SEED_LISTS: lists per query scored by the seed pass (1..16)
  int seed_tile = 0;           // B2VS_IVF_SEED_TILE: rows of each seed list that are scored (32..256, multiple of 32)
  int two_pass = -1;           // B2VS_TWO_PASS=0|1: never / whenever possible the two-pass selection of the flat engine
  int two_pass_chunk_mb = 0;   // B2VS_TWO_PASS_CHUNK_MB: candidate-buffer bytes per sub-batch of queries
  bool canary = false;         // B2VS_CANARY=1: guard zones around every device buffer (read ONCE, at first use)
};
const EnvConfig& env();

// ------------------------------------------------------------------------------------------
// The exact-search engine: a [n, kdim] 16-bit K-major matrix + per-row additive term.
// Used directly by flat indexes and re-used for the coarse quantizer / k-means assignment.
struct BfTcParams;  // kernel parameters (bf_tc.cuh)

// Hook called between the passes of a flat search with the per-query thresholds of the sampled
// pass (device, n floats): the sharded search all-reduces them (MIN) across ranks (comm.cu).
struct TauExchange {
  int (*fn)(void* ctx, float* tau, int64_t n, cudaStream_t st);
  void* ctx;
  int64_t schedule_rows;   // rows of the SMALLEST shard: the pass schedule (= number of exchanges)
                           // is derived from it, so every rank makes the same collective calls
  // Union mode (world > 1 and union_fn set): a sampled pass hands over its k best raw SCORES per
  // query ([n][k], ascending, +inf padded); union_fn all-gathers them and writes the k-th best of
  // the union (+ one ulp) to tau.  The sample of the job is then `world` times larger than a rank's
  // own, so the sampled passes run at `world` times the stride.
  int world;
  int (*union_fn)(void* ctx, const float* scores, int64_t n, int k, float* tau, cudaStream_t st);
};
// True when a flat search over at least `min_rows` rows with this k runs a sampled pass before
// the full one, i.e. has thresholds to exchange (must evaluate identically on every rank).
// stride_mult = the factor applied to the sampled passes' strides (union mode: the world size).
bool flat_exchanges_tau(int64_t min_rows, int k, int stride_mult = 1);

struct FlatEngine {
  int dev = 0;
  int metric = B2VS_METRIC_L2;
  int src_dtype = B2VS_BF16;  // dtype the caller gave us
  int dim = 0;                // logical dimension
  int kdim = 0;               // GEMM K extent in 16-bit elements (dim padded to 8, x3 for fp32)
  int ab_format = 1;          // 0 = fp16, 1 = bf16 (tensor-core operand format)
  bool split3 = false;        // fp32 source re-encoded as bf16 [hi | hi | lo]
  int64_t n = 0;
  const void* mat = nullptr;  // operand matrix used by TMA (borrowed or == owned.ptr)
  DevBuf owned;               // owned re-encoding, if any
  DevBuf beta;                // [tiles*256] float: ||x||^2 (L2) or 0 (IP); +inf on padding
  CUtensorMap tm_x;       // db map, box = 256 rows (single-CTA kernel)
  CUtensorMap tm_x_half;  // db map, box = 128 rows (CTA-pair kernel: each CTA stages half a tile)
  // workspaces (grow-only)
  DevBuf ws_cand, ws_keys, ws_q, ws_qnorm, ws_tau, ws_big, ws_bigcnt, ws_chunk, ws_work, ws_rawcnt, ws_samp;
  b2vs_search_stats stats{};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;  // optional timing of the dominant kernel
  bool timing_pending = false;

  // force_fmt >= 0 with an fp32 source: plain conversion to that 16-bit operand format instead
  // of the hi/lo split (used for coarse-quantizer centroids, which live in fp32).
  int init(int dev, int metric, int dtype, int dim, const void* db, int64_t n, cudaStream_t st,
           int force_fmt = -1);
  // Top-k of queries vs this matrix. out_keys layout is internal; results are written as
  // (dist fp32, id int64 = row + id_offset) when out_d/out_i are given, or as int32 labels
  // (k == 1) when out_label is given.
  int search(const void* q, int q_dtype, int nq, int k, int force_splits, int64_t id_offset,
             float* out_d, int64_t* out_i, int32_t* out_label, cudaStream_t st, int flags = 0,
             const TauExchange* tau_exchange = nullptr);
  int launch_fused(int group, int grid, const CUtensorMap& tm_q, const BfTcParams& p,
                   cudaStream_t st, int epi_groups = 1) const;
  void resolve_timing();  // fills stats.kernel_ms once the timed launch has finished
  // 128 < k <= 2048 (bigk.cu): append-mode passes + per-query radix select
  int search_bigk(const void* q_mat, int nq, int q_pad, int group, int k, int64_t id_offset,
                  float* out_d, int64_t* out_i, cudaStream_t st, int* launches);
  // small database, large k (the coarse probes of the IVF indexes): two tensor-core passes,
  // per-chunk minima -> exact (score, chunk) threshold -> the <= 32 k survivors per query
  bool two_pass_applies(int nq, int k, int group_forced, int flags) const;
  int search_two_pass(const void* q_mat, int64_t q_rows, int nq, int k, int64_t id_offset, float* out_d,
                      int64_t* out_i, int32_t* out_label, cudaStream_t st, bool timed, int* launches);
  size_t owned_bytes() const { return owned.bytes + beta.bytes; }
  void destroy();
};

int encode_tmap_2d(CUtensorMap* tm, const void* base, int ab_format, int64_t rows, int64_t cols,
                   int box_rows);

// Grouped list scan on the tensor cores (flat.cu; used by the IVF-Flat batch search): work
// item = (block of 128 gathered query rows, row range of one inverted list), append mode.
struct GroupedScanArgs {
  const void* q_mat;      // gathered query operand [q_rows, kdim] 16-bit, q_rows % 128 == 0
  int64_t q_rows;
  const void* x_mat;      // list rows [x_rows, kdim] 16-bit
  int64_t x_rows;
  int kdim, ab_format;
  int q_split;            // 1: q_mat is [hi | lo], each half padded to a multiple of 64 columns
  const float* beta;      // [x_rows + 256] additive term per list row (+inf on padding slots)
  float alpha;
  const void* work;       // int4 [max_work] {query block, first row, end row, -}
  const int* n_work;      // device scalar
  int max_work;
  const int* row_query;   // [q_rows] query of each gathered row, -1 on padding
  const float* tau;       // [nq] per-query threshold
  u64* cand;              // [nq][cap]
  int* count;             // [nq]
  int cap;
  const int* row_slot;    // seed pass (seed_all): which seed list of its query a gathered row probes
  int seed_all;           // 1: every score of the (single-tile) items goes to its fixed slot
                          // 2: only per-chunk minima are stored (chunk_min)
  float* chunk_min;       // seed_all == 2: [nq][chunk_ld] minimum of every 32-row chunk
  int chunk_ld;
  const int* tau_chunk;   // optional [nq]: thresholds are (score, chunk) pairs (see BfTcParams)
  int rows_in_work;       // 1: work[i].w = query rows of the block, placed in its leading ceil(w / 32) quarters
                          // (packed or contiguous dealing): the kernel loads only those quarters
  int x_box_rows;         // 0 / 256: whole 256-row list tiles; 64 or 128: only that many rows of each (single-tile)
                          // item are loaded and multiplied - the seed pass reads just the head of every list
  int epi_groups;         // 0/1: four epilogue warps; 2: eight (two groups on alternate tiles) for
                          // launches whose epilogue, not HBM, sets the pace
};
int launch_grouped_scan(int dev, const GroupedScanArgs& a, cudaStream_t st);

// Grouped IVF-PQ scan (pq_tc.cu): lists decoded once per batch into bf16 tiles for the tensor cores.
struct PqGroupedScanArgs {
  const void* q_mat;      // gathered residual queries [q_rows, dim] bf16
  int64_t q_rows;
  int dim, pq_dim, dsub;
  const void* codes;      // PQ codes [n_groups][pq_dim][32] (sub-space major inside 32-row groups)
  uint32_t n_groups;      // 32-row groups in `codes`
  const void* cb16t;      // bf16 codebooks, code-major [256][pq_dim][dsub]
  const float* beta;      // [n_slots + 256] ||r^||^2 (L2) / 0 (IP), +inf on padding slots
  float alpha;
  const void* work;       // int4 [max_work]
  const int* n_work;
  int max_work;
  const int* row_query;   // [q_rows]
  const float* row_bias;  // [q_rows]
  const float* tau;       // [nq]
  u64* cand;
  int* count;
  int cap;
  const int* row_slot;    // seed pass, as in GroupedScanArgs
  int seed_all;
};
bool pq_grouped_supported(int dim, int dsub);
int launch_pq_grouped_scan(int dev, const PqGroupedScanArgs& a, cudaStream_t st);

// merge.cu
// `remap` (optional) translates key ids (list slots) to shard-local row ids before id_offset.
int launch_merge_splits(const u64* keys, int n_splits, int q_pad, int nq, int k, int metric,
                        const float* qnorm, int64_t id_offset, float* out_d, int64_t* out_i,
                        int32_t* out_label, cudaStream_t st, const uint32_t* remap = nullptr,
                        float* out_tau = nullptr, float* out_scores = nullptr);

// Raw emission of the fused kernel: top-k (or the k-th score) per query from the unsorted per-item lists.
int launch_merge_raw(const u64* cand, const int* count, int n_splits, int n_qblocks, int group,
                     int epi_groups, int nq, int k, int metric, const float* qnorm, int64_t id_offset,
                     float* out_d, int64_t* out_i, int32_t* out_label, float* out_tau, cudaStream_t st,
                     float* out_scores = nullptr);
// k-th best (+ one ulp) of the union of n_ranks sorted raw-score lists per query (sharded sampled pass)
int launch_union_kth(const float* all_scores, int n_ranks, int nq, int k, float* tau, cudaStream_t st);
// Two-pass selection (flat.cu: search_two_pass).  chunk_tau: per query the k-th smallest
// (chunk minimum, chunk index) pair -> tau / tau_chunk, and count[q] = 0.  cand_select: the k best
// of each query's appended candidates, straight to answer rows.
int launch_chunk_tau(const float* chunk_min, int chunk_ld, int n_chunks, int nq, int k, float* tau,
                     int* tau_chunk, int* count, cudaStream_t st);
int launch_cand_select(const u64* cand, const int* count, int cap, int nq, int k, int metric,
                       const float* qnorm, int64_t id_offset, float* out_d, int64_t* out_i,
                       int32_t* out_label, cudaStream_t st);

// cosine.cu
int launch_unit_rows(const void* src, void* dst, int dtype, int64_t n, int dim, cudaStream_t st);
int launch_cosine_fixup(float* d, int64_t total, cudaStream_t st);

// bigk.cu
constexpr int kMaxBigK = 2048;
int launch_bigk_select(const u64* cand, const int* counts, int cap, int nq, int k, int final_pass,
                       int metric, const float* qnorm, int64_t id_offset, float* out_tau,
                       float* out_d, int64_t* out_i, int* overflow, cudaStream_t st,
                       const uint32_t* remap = nullptr);
int launch_merge_parts_big(const float* d_all, const int64_t* i_all, int n_parts, int nq, int k_in,
                           int k_out, int descending, float* out_d, int64_t* out_i, cudaStream_t st);

}  // namespace b2vs

struct b2vs_index {
  int kind = B2VS_KIND_FLAT;
  int dev = 0;
  int metric = 0;
  int dtype = 0;
  int dim = 0;
  int64_t n = 0;
  int64_t id_offset = 0;
  b2vs::FlatEngine flat;  // flat index, or the coarse quantizer of an IVF index
  void* ivf = nullptr;    // b2vs::IvfData* (ivf_internal.cuh)
  // B2VS_METRIC_COSINE: the index runs as IP (metric == B2VS_METRIC_IP) over an owned unit-norm
  // copy of the rows; queries are normalised into cos_q and distances leave as 1 - similarity
  bool cosine = false;
  b2vs::DevBuf cos_rows, cos_q;
  // rows of the smallest shard of the sharded job this index belongs to, agreed by
  // b2vs_comm_register_index (0 = not registered: sharded searches keep private thresholds)
  int64_t sharded_min_rows = 0;
};
