// IVF-Flat / IVF-PQ internals shared by the IVF translation units:
//   kmeans.cu      GPU k-means (K1 assign on the tensor cores, K2 segmented update) + label grouping (K3)
//   ivf_build.cu   list construction, PQ training / encoding, the two build entry points
//   ivf_plan.cu    grouping of (query, probe) items by list, work tables of the grouped scans
//   ivf_scan.cu    list-scan kernels (K5 / K7 families), seed / rescue / select / refine
//   ivf_search.cu  one search batch: coarse probe -> seed -> grouped scan -> select (-> refine)
//   ivf_graph.cu   small batches replayed as one CUDA graph; ivf_search entry
//   persist.cu     b2vs_index_save / b2vs_index_load
//
// Replaces cuvs.neighbors.ivf_flat / ivf_pq build + search at the reference call sites
// (index_building_coordinator.py:392-404, improved_multi_gpu_rag.py:126-138 and :225-233).
// Semantics restated from the published algorithms (cuVS/FAISS sources are not vendored):
//   IVF-Flat = Lloyd k-means(n_lists) on a strided subsample, every row assigned to its nearest
//              centroid, search probes the n_probes best centroids and scans those lists exactly.
//   IVF-PQ   = the same coarse quantizer, residual PQ with pq_dim sub-quantizers x 256 codes,
//              ADC with a per-(query, list) look-up table.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <utility>
#include <vector>

#include "ivf.h"
#include "warp_select.cuh"

namespace b2vs {

constexpr uint32_t kNoRow = 0xFFFFFFFFu;
constexpr int kScanThreads = 256;
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kGroupRows = 128;  // query rows per grouped-scan work item (= the UMMA M extent)
#ifndef B2VS_SKIP_PAD_ROWS
#define B2VS_SKIP_PAD_ROWS 1
#endif
constexpr bool kSkipPaddingRows = B2VS_SKIP_PAD_ROWS != 0;   // gather kernels leave group-padding rows unwritten
constexpr int kSeedSortMinQueries = 2048;  // below this the seed pass skips its ordering sort
constexpr int kTcSeedMinQueries = 256;     // from here on thresholds come from a tensor-core seed pass
constexpr int kSeedTileRows = 256;         // rows of every seed list that the seed pass scores
constexpr int kMaxProbes = 2048;  // coarse probe = exact top-n_probes (large-k path above 128)
constexpr int kNormSlack = 256;  // slot_norm floats past the last slot (whole-tile beta loads)

// A small-batch search replayed as one CUDA graph (the Q <= 64 path is ~15-20 tiny kernels and
// launch-bound).  The graph is captured against library-owned staging buffers, so a replay is
// copy-in, graph launch, two copies out, whatever pointers the caller passes.
struct SearchGraph {
  int nq = 0, k = 0, q_dtype = 0, n_probes = 0, refine_ratio = 0, flags = 0;
  int seen = 0;              // direct (un-captured) calls with this signature so far
  bool failed = false;       // capture was refused once: stay on the direct path
  uint64_t generation = 0;   // realloc_generation() at capture time
  uint64_t last_use = 0;
  cudaGraphExec_t exec = nullptr;
  char* io = nullptr;        // staging: queries | distances | ids
  size_t q_bytes = 0, d_off = 0, i_off = 0;
  b2vs_search_stats stats{};
  void destroy() {
    if (exec) cudaGraphExecDestroy(exec);
    if (io) cudaFree(io);
    exec = nullptr;
    io = nullptr;
  }
};
constexpr int kGraphMaxQueries = 64;   // opt-in (B2VS_FLAG_GRAPH) replay of launch-bound small batches
constexpr int kGraphAutoMinQueries = 2048;    // large batches are replayed as graphs by default ...
constexpr int kGraphAutoMaxQueries = 1 << 20; // ... up to here
constexpr int kGraphMaxEntries = 16;

struct IvfData {
  int n_lists = 0, pq_dim = 0, pq_bits = 0, dsub = 0, mp = 0;  // mp = pq_dim padded to 16
  int dp = 0;       // dim padded to 8 (16-bit storage pitch)
  float max_norm2 = 0.f;  // IVF-Flat: largest ||x||^2 over the rows (rounding cushion of the seed threshold)
  int fmt = 1;      // storage format of list vectors: 0 fp16, 1 bf16
  int64_t n = 0;
  int64_t n_slots = 0;
  DevBuf centroids;   // f32 [n_lists, dim]
  DevBuf offsets;     // u32 [n_lists + 1] slot offsets
  DevBuf sizes;       // i32 [n_lists]
  DevBuf row_ids;     // u32 [n_slots] shard-local row of each slot (kNoRow on padding)
  DevBuf data;        // IVF-Flat: u16 [n_slots, dp]
  DevBuf slot_norm;   // IVF-Flat: f32 [n_slots + kNormSlack] ||x||^2 (L2) or 0 (IP); +inf on padding
  DevBuf codebooks;   // IVF-PQ: f32 [pq_dim, 256, dsub]
  DevBuf codes;       // IVF-PQ: u8 [n_slots / 32][mp][32]: 32-row groups, sub-space major inside a
                      // group (see pq_code_offset): the decoder warps of the grouped scan read the
                      // 32 codes of one (group, sub-space) as ONE 32-byte piece per lane
  // IVF-PQ grouped tensor-core scan (derived from codebooks + codes at build / load time)
  DevBuf cb16;        // bf16 [pq_dim, 256, dsub]: the codebooks as the MMA sees them
  DevBuf cb16t;       // bf16 [256, pq_dim, dsub]: the same, code-major (decoder look-ups, bank = sub-space)
  DevBuf cbn;         // f32 [pq_dim, 256] squared norm of each (rounded) codebook entry
  DevBuf pq_norm;     // f32 [n_slots + kNormSlack] ||decoded residual||^2 (L2) / 0 (IP); +inf on padding
  float max_rhat2 = 0.f;
  bool pq_tc_ready = false;
  DevBuf ws_probe_d, ws_probe_i, ws_keys, ws_qf, ws_qnorm, ws_counter, ws_ref_d, ws_ref_i;
  DevBuf ws_item_lab, ws_item_cnt, ws_item_off, ws_item_perm, ws_item_slot;  // list-ordered scan items
  DevBuf ws_g_work, ws_g_q, ws_g_rowq, ws_g_tau, ws_g_cand, ws_g_cnt, ws_g_bias;  // grouped scan
  DevBuf ws_over, ws_rescue;   // overflow queue of the select kernel; sliced rescue lists
  DevBuf ws_g_rowslot;  // [gathered rows] probe rank of each row inside its query
  DevBuf ws_seed_ids;   // [nq, m] the nearest probes of every query (tensor-core seed pass)
  // Second set of the planning buffers (item sort, work table, gathered operand): the main pass of a
  // large batch is planned and gathered on `side` while the seed pass runs on the caller's stream.
  // swap_plan_ws() exchanges the two sets; launches see whichever is current when they are issued.
  DevBuf alt_item_lab, alt_item_cnt, alt_item_off, alt_item_perm, alt_item_slot;
  DevBuf alt_g_work, alt_g_q, alt_g_rowq, alt_g_rowslot, alt_g_bias;
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  void swap_plan_ws() {
    std::swap(ws_item_lab, alt_item_lab); std::swap(ws_item_cnt, alt_item_cnt);
    std::swap(ws_item_off, alt_item_off); std::swap(ws_item_perm, alt_item_perm);
    std::swap(ws_item_slot, alt_item_slot); std::swap(ws_g_work, alt_g_work);
    std::swap(ws_g_q, alt_g_q); std::swap(ws_g_rowq, alt_g_rowq);
    std::swap(ws_g_rowslot, alt_g_rowslot); std::swap(ws_g_bias, alt_g_bias);
  }
  const void* src_rows = nullptr;  // IVF-PQ: the caller's [n, dim] rows, BORROWED for refine
  std::vector<int32_t> h_sizes;
  DevBuf rank_of_list, list_of_rank;  // int [n_lists]: lists in descending-size order (scan scheduling)
  int max_list_rows = 0;              // longest list, padded to 32 slots
  b2vs_search_stats stats{};
  bool counter_pending = false;
  int last_nq = 0;
  int row_bytes = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timing_pending = false;
  // coarse probe of very small batches on the CUDA cores (coarse_probe_scan): the centroid table
  // seen as 256-row pseudo-lists
  DevBuf cq_offsets, cq_probe, ws_cq_keys;
  bool cq_ready = false;
  std::vector<SearchGraph> graphs;      // captured small-batch searches (see ivf_search)
  cudaStream_t cap_stream = nullptr;    // capture happens here: the caller's stream may be the legacy one
  uint64_t graph_clock = 0;
  size_t owned_bytes() const {
    return centroids.bytes + offsets.bytes + sizes.bytes + row_ids.bytes + data.bytes +
           slot_norm.bytes + codebooks.bytes + codes.bytes + cb16.bytes + cb16t.bytes + cbn.bytes + pq_norm.bytes +
           rank_of_list.bytes + list_of_rank.bytes;
  }
  void destroy() {
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    ev0 = ev1 = nullptr;
    for (SearchGraph& g : graphs) g.destroy();
    graphs.clear();
    if (cap_stream) cudaStreamDestroy(cap_stream);
    cap_stream = nullptr;
    if (side) cudaStreamDestroy(side);
    side = nullptr;
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    ev_fork = ev_join = nullptr;
    for (DevBuf* b : {&centroids, &offsets, &sizes, &row_ids, &data, &slot_norm, &codebooks, &codes,
                      &rank_of_list, &list_of_rank,
                      &ws_probe_d, &ws_probe_i, &ws_keys, &ws_qf, &ws_qnorm, &ws_counter,
                      &ws_ref_d, &ws_ref_i, &ws_item_lab, &ws_item_cnt, &ws_item_off, &ws_item_perm,
                      &ws_item_slot, &ws_g_work, &ws_g_q, &ws_g_rowq, &ws_g_tau, &ws_g_cand, &ws_g_cnt,
                      &ws_g_bias, &ws_g_rowslot, &ws_seed_ids, &alt_item_lab, &alt_item_cnt, &alt_item_off, &alt_item_perm,
                      &alt_item_slot, &alt_g_work, &alt_g_q, &alt_g_rowq, &alt_g_rowslot, &alt_g_bias, &ws_over, &ws_rescue, &cb16, &cb16t, &cbn, &pq_norm, &cq_offsets, &cq_probe, &ws_cq_keys})
      b->release();
  }
};

// ------------------------------------------------------------------------------------------
// Four consecutive elements (p aligned to 4 elements) as floats: one 8- or 16-byte load.
template <typename T> __device__ __forceinline__ float4 ld4_f32(const T* p);
template <> __device__ __forceinline__ float4 ld4_f32<float>(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
template <> __device__ __forceinline__ float4 ld4_f32<__half>(const __half* p) {
  const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
template <> __device__ __forceinline__ float4 ld4_f32<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
  return make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xFFFF0000u),
                     __uint_as_float(v.y << 16), __uint_as_float(v.y & 0xFFFF0000u));
}
template <typename T> __device__ __forceinline__ float ld_f32(const T* p);
template <> __device__ __forceinline__ float ld_f32<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_f32<__half>(const __half* p) { return __half2float(*p); }
template <> __device__ __forceinline__ float ld_f32<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

#define DISPATCH_DTYPE(dtype, T, ...)                              \
  switch (dtype) {                                                 \
    case B2VS_F32: { using T = float; __VA_ARGS__; break; }        \
    case B2VS_F16: { using T = __half; __VA_ARGS__; break; }       \
    default: { using T = __nv_bfloat16; __VA_ARGS__; break; }      \
  }

__device__ __forceinline__ uint16_t to_op16(float v, int fmt, float* back) {
  if (fmt == 0) {
    const __half h = __float2half_rn(v);
    *back = __half2float(h);
    return __half_as_ushort(h);
  }
  const __nv_bfloat16 b = __float2bfloat16_rn(v);
  *back = __bfloat162float(b);
  return __bfloat16_as_ushort(b);
}

// byte offset of the code of sub-space m of list slot `slot` (mp = pq_dim padded to 16)
__host__ __device__ __forceinline__ size_t pq_code_offset(size_t slot, int m, int mp) {
  return ((slot >> 5) * static_cast<size_t>(mp) + static_cast<size_t>(m)) * 32 + (slot & 31);
}

// Position of the i-th of the c queries of a list's group inside the gathered operand.  A 128-row
// block is four 32-row quarters (TMEM lane quarter = epilogue warp); the queries of a block are
// dealt round-robin over as FEW quarters as hold them (ceil(rows in block / 32)), so
//  * a partly filled block - the common case: ~40 of 128 rows at C4 - loads its active epilogue
//    warps evenly instead of leaving all hits to the warp of rows 0-31, and
//  * quarters left empty are skipped by their epilogue warps altogether (bf_tc.cuh / pq_tc.cuh).
// c = 0 (size unknown to the caller): deal over all four quarters.
__host__ __device__ __forceinline__ uint32_t group_row_pos(uint32_t i, uint32_t c) {
  const uint32_t blk = i >> 7, j = i & 127u;
  uint32_t in_block = 128u;
  if (c != 0u && blk == ((c - 1u) >> 7)) in_block = ((c - 1u) & 127u) + 1u;   // the group's last block
  const uint32_t nq = c == 0u ? 4u : (in_block + 31u) >> 5;                  // quarters in use: 1..4
  return (blk << 7) | ((j % nq) << 5) | (j / nq);
}

// ---- kmeans.cu ------------------------------------------------------------------------------
// Temporaries of one k-means fit.  A caller that fits many small problems in a row (the 64+ PQ
// sub-codebooks) passes the same workspace to every fit, so device memory is allocated once
// instead of being malloc'ed and freed (= device-synchronised) per fit.
struct KmWorkspace {
  DevBuf sums, counts, labels, donors, seg_off, seg_cur, seg_rows, seg_slot, order, donor_scratch;
  FlatEngine eng;
  void release() {
    for (DevBuf* b : {&sums, &counts, &labels, &donors, &seg_off, &seg_cur, &seg_rows, &seg_slot,
                      &order, &donor_scratch})
      b->release();
    eng.destroy();
  }
};
int kmeans_fit_impl(int dev, int dtype, int dim, const void* x, int64_t n, int ncl, int iters,
                    uint64_t seed, float* cent, int32_t* labels_out, cudaStream_t st,
                    KmWorkspace* shared_ws = nullptr);
// Label grouping (K3; also the counting sort of (query, probe) items): histogram of labels,
// exclusive scan of the sizes rounded up to `pad`, scatter of the rows into their groups.
int launch_histogram(const int* labels, int64_t n, int* sizes, int blocks, cudaStream_t st);
int launch_scan_sizes(const int* sizes, int n_lists, int pad, uint32_t* offsets, cudaStream_t st);
int launch_scatter_rows(const int* labels, int64_t n, const uint32_t* offsets, int* cursor,
                        uint32_t* row_ids, uint32_t* slot_of_row, int blocks, cudaStream_t st,
                        const int* deal_sizes = nullptr, int deal_four = 0);
int launch_strided_rows(const void* src, void* dst, int dtype, int64_t n_out, int64_t stride, int dim,
                        cudaStream_t st);

// ---- ivf_build.cu ---------------------------------------------------------------------------
int launch_fill_f32(float* p, size_t n, float v, cudaStream_t st);
int build_list_ranks(IvfData* d);
int pq_prepare_grouped(b2vs_index* index, IvfData* d, cudaStream_t st);

// ---- ivf_plan.cu ----------------------------------------------------------------------------
size_t sorted_rows_cap(const IvfData* d, int items, int group_pad);
int reserve_item_sort(IvfData* d, int items, int group_pad);
// deal: how the queries of a group are placed inside their 128-row blocks (group_row_pos):
// kDealPacked = over as few 32-row quarters as hold them (empty quarters are skipped by their
// epilogue warps: fewest instructions - the HBM-bound IVF-Flat scan), kDealFour = over all four
// quarters, kDealNone = in arrival order, contiguous from row 0 (the PQ scan: its queries are the
// COLUMNS of the accumulator, only the leading 16-column units that hold queries are read back)
constexpr int kDealPacked = 0, kDealFour = 1, kDealNone = 2;
int sort_items_by_list(IvfData* d, const long long* probe_ids, int items, int stride, int group_pad,
                       cudaStream_t st, int deal = kDealPacked);
// row_limit > 0: only the first row_limit rows of every list become work (the seed pass)
int plan_grouped_work(IvfData* d, const long long* probe_ids, int items, int chunk_rows, int slots,
                      int4* work, int* n_work, unsigned long long* counter, cudaStream_t st,
                      int row_limit = 0, int deal = kDealPacked);
void choose_work_split(const b2vs_index* index, const IvfData* d, int items, int* chunk_rows, int* slots);
bool plan_is_small(const IvfData* d, int items);

// ---- ivf_scan.cu (launchers of the scan-side kernels) ---------------------------------------
int launch_queries_to_f32(const void* q, int q_dtype, int nq, int dim, int dp, int fmt, int round16,
                          float* qf, float* qnorm, cudaStream_t st);
// per-(query, pseudo-)list scan over any 16-bit row matrix (K5; also the coarse probe K4b)
int launch_flat_item_scan(int fmt, int dp, int grid, const uint16_t* data, const float* slot_norm,
                          const uint32_t* offsets, const long long* probe_ids, const float* qf,
                          int n_probes, int q_pad, int k, float alpha, u64* out_keys,
                          unsigned long long* scanned_rows, const uint32_t* item_perm, cudaStream_t st);
int launch_flat_seed_tau(const b2vs_index* index, IvfData* d, const long long* probe_ids, int n_probes,
                         int nq, int k, uint32_t seed_rows, float extra_eps, const uint32_t* q_perm,
                         cudaStream_t st);
int launch_flat_rescue(const b2vs_index* index, IvfData* d, const long long* probe_ids, int n_probes,
                       int nq, int k, int cap, cudaStream_t st);
int launch_group_select(IvfData* d, int nq, int cap, int k, unsigned long long* total_cand, cudaStream_t st);
// thresholds of the full pass from the candidates the seed pass appended (k-th best per query)
int launch_seed_select(IvfData* d, int nq, int n_keys, int cap, int k, cudaStream_t st);
// items_sorted > 0: gather by item (the counting-sort kernels ran); 0: walk the rows (one-CTA planner)
int launch_gather_group_queries(IvfData* d, int64_t rows_cap, int n_probes, int q_split, int items_sorted,
                                cudaStream_t st);
int launch_gather_group_residuals(const b2vs_index* index, IvfData* d, int64_t rows_cap,
                                  const long long* probe_ids, int n_probes, int items_sorted, cudaStream_t st);
// mode 0 = seed thresholds (ws_g_tau), mode 1 = rescue of overflowed queries (ws_keys)
int launch_pq_lut_scan(int mode, const b2vs_index* index, IvfData* d, const long long* probe_ids,
                       int n_probes, int nq, int k, int cap, uint32_t seed_rows, const uint32_t* q_perm,
                       cudaStream_t st);
// look-up-table scans (shapes without the grouped scan / B2VS_IVF_GROUPED=0): *single_list tells
// whether ws_keys holds one sorted list per query or one per (probe, query)
int launch_pq_table_scan(const b2vs_index* index, IvfData* d, const long long* probe_ids, int n_probes,
                         int nq, int q_pad, int k, unsigned long long* counter, bool* single_list,
                         cudaStream_t st);
int launch_refine(const b2vs_index* index, IvfData* d, const long long* cand, int nq, int k_in, int k_out,
                  float* out_d, int64_t* out_i, cudaStream_t st);
int launch_refine_big(const b2vs_index* index, IvfData* d, const long long* cand, int nq, int k_in,
                      int k_out, float* out_d, int64_t* out_i, cudaStream_t st);

// ---- ivf_search.cu --------------------------------------------------------------------------
bool uses_bigk_path(const b2vs_index* index, const IvfData* d, int k, const b2vs_search_params& sp);
int ivf_search_direct(b2vs_index* index, const void* q, int q_dtype, int nq, int k,
                      const b2vs_search_params& sp, float* out_d, int64_t* out_i, cudaStream_t st);

}  // namespace b2vs
