// IVF-Flat / IVF-PQ: GPU k-means (K1/K2), list construction (K3), coarse probe (K4 = the fused
// tensor-core engine with k = n_probes), list scan (K5), PQ encode (K6) and LUT scan (K7).
//
// Replaces cuvs.neighbors.ivf_flat / ivf_pq build + search at the reference call sites
// (index_building_coordinator.py:392-404, improved_multi_gpu_rag.py:126-138 and :225-233).
// Semantics restated from the published algorithms (cuVS/FAISS sources are not vendored):
//   IVF-Flat = Lloyd k-means(n_lists) on a strided subsample, every row assigned to its nearest
//              centroid, search probes the n_probes best centroids and scans those lists exactly.
//   IVF-PQ   = the same coarse quantizer, residual PQ with pq_dim sub-quantizers x 256 codes,
//              ADC with a per-(query, list) look-up table.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "ivf.h"
#include "topk.cuh"

namespace b2vs {

constexpr uint32_t kNoRow = 0xFFFFFFFFu;
constexpr int kScanThreads = 256;
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kListE = 4;  // warp-resident sorted list: 32 * 4 = 128 keys
constexpr int kGroupRows = 128;  // query rows per grouped-scan work item (= the UMMA M extent)
#ifndef B2VS_SKIP_PAD_ROWS
#define B2VS_SKIP_PAD_ROWS 1
#endif
constexpr bool kSkipPaddingRows = B2VS_SKIP_PAD_ROWS != 0;   // gather kernels leave group-padding rows unwritten
constexpr int kSeedSortMinQueries = 2048;  // below this the seed pass skips its ordering sort
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
constexpr int kGraphMaxQueries = 64;   // larger batches are no longer launch-bound
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
  DevBuf codes;       // IVF-PQ: u8, 32-row groups interleaved by 16-byte chunks
  // IVF-PQ grouped tensor-core scan (derived from codebooks + codes at build / load time)
  DevBuf cb16;        // bf16 [pq_dim, 256, dsub]: the codebooks as the MMA sees them
  DevBuf cbn;         // f32 [pq_dim, 256] squared norm of each (rounded) codebook entry
  DevBuf pq_norm;     // f32 [n_slots + kNormSlack] ||decoded residual||^2 (L2) / 0 (IP); +inf on padding
  float max_rhat2 = 0.f;
  bool pq_tc_ready = false;
  DevBuf ws_probe_d, ws_probe_i, ws_keys, ws_qf, ws_qnorm, ws_counter, ws_ref_d, ws_ref_i;
  DevBuf ws_item_lab, ws_item_cnt, ws_item_off, ws_item_perm, ws_item_slot;  // list-ordered scan items
  DevBuf ws_g_work, ws_g_q, ws_g_rowq, ws_g_tau, ws_g_cand, ws_g_cnt, ws_g_bias;  // grouped scan
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
           slot_norm.bytes + codebooks.bytes + codes.bytes + cb16.bytes + cbn.bytes + pq_norm.bytes +
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
    for (DevBuf* b : {&centroids, &offsets, &sizes, &row_ids, &data, &slot_norm, &codebooks, &codes,
                      &rank_of_list, &list_of_rank,
                      &ws_probe_d, &ws_probe_i, &ws_keys, &ws_qf, &ws_qnorm, &ws_counter,
                      &ws_ref_d, &ws_ref_i, &ws_item_lab, &ws_item_cnt, &ws_item_off, &ws_item_perm,
                      &ws_item_slot, &ws_g_work, &ws_g_q, &ws_g_rowq, &ws_g_tau, &ws_g_cand, &ws_g_cnt,
                      &ws_g_bias, &cb16, &cbn, &pq_norm, &cq_offsets, &cq_probe, &ws_cq_keys})
      b->release();
  }
};

// ------------------------------------------------------------------------------------------
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

// ---- K-means pieces -----------------------------------------------------------------------
template <typename T>
__global__ void strided_rows_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t n_out,
                                    int64_t stride, int dim) {
  const int64_t total = n_out * dim;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / dim;
    const int j = static_cast<int>(i - r * dim);
    dst[i] = src[r * stride * dim + j];
  }
}

template <typename T>
__global__ void seed_centroids_kernel(const T* __restrict__ x, int64_t n, int dim, int ncl,
                                      uint64_t seed, float* __restrict__ cent) {
  const int c = blockIdx.x;
  // one distinct stratum per centroid, random offset inside it
  const int64_t lo = static_cast<int64_t>((static_cast<double>(c) * n) / ncl);
  const int64_t hi = static_cast<int64_t>((static_cast<double>(c + 1) * n) / ncl);
  const int64_t span = hi > lo ? hi - lo : 1;
  const int64_t row = min(n - 1, lo + static_cast<int64_t>(mix64(seed ^ (0x51ull * (c + 1))) % span));
  for (int j = threadIdx.x; j < dim; j += blockDim.x)
    cent[static_cast<size_t>(c) * dim + j] = ld_f32<T>(x + row * dim + j);
}

// K2 update, first half: SEGMENTED reduction.  Rows are grouped by label with the same
// histogram -> scan -> scatter kernels that build the IVF lists (K3); then one CTA column per
// cluster sums its segment, thread j owning dimension j (coalesced row reads, no atomics).
__global__ void histogram_kernel(const int* __restrict__ labels, int64_t n, int* __restrict__ sizes);
__global__ void scan_sizes_kernel(const int* __restrict__ sizes, int n_lists, int pad,
                                  uint32_t* __restrict__ offsets);
__global__ void scatter_rows_kernel(const int* __restrict__ labels, int64_t n,
                                    const uint32_t* __restrict__ offsets, int* __restrict__ cursor,
                                    uint32_t* __restrict__ row_ids, uint32_t* __restrict__ slot_of_row);

template <typename T>
__global__ void segment_sum_kernel(const T* __restrict__ x, const uint32_t* __restrict__ offsets,
                                   const uint32_t* __restrict__ row_ids, int dim,
                                   float* __restrict__ sums) {
  const int c = blockIdx.x;
  const int j = blockIdx.y * blockDim.x + threadIdx.x;
  if (j >= dim) return;
  const uint32_t begin = offsets[c], end = offsets[c + 1];
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  uint32_t i = begin;
  for (; i + 4 <= end; i += 4) {
    const uint32_t r0 = row_ids[i], r1 = row_ids[i + 1], r2 = row_ids[i + 2], r3 = row_ids[i + 3];
    a0 += ld_f32<T>(x + static_cast<size_t>(r0) * dim + j);
    a1 += ld_f32<T>(x + static_cast<size_t>(r1) * dim + j);
    a2 += ld_f32<T>(x + static_cast<size_t>(r2) * dim + j);
    a3 += ld_f32<T>(x + static_cast<size_t>(r3) * dim + j);
  }
  for (; i < end; ++i) a0 += ld_f32<T>(x + static_cast<size_t>(row_ids[i]) * dim + j);
  sums[static_cast<size_t>(c) * dim + j] = (a0 + a1) + (a2 + a3);
}

// K2 update, second half: mean of each cluster.  Balancing (the role cuVS's balanced k-means
// "adjust centers" step plays): `donor_of[c] >= 0` tells cluster c - empty, or far below the
// average size - to restart on a data row of the over-full cluster donor_of[c] (the pairing is
// computed on the host from the cluster sizes, see balance_pairs), so Lloyd does not leave a few
// giant lists next to starved ones.  Empty clusters without a donor restart on a random row.
template <typename T>
__global__ void finalize_centroids_kernel(const T* __restrict__ x, const int* __restrict__ labels,
                                          int64_t n, int dim, const float* __restrict__ sums,
                                          const int* __restrict__ counts,
                                          const int* __restrict__ donor_of, uint64_t seed,
                                          float* __restrict__ cent) {
  const int c = blockIdx.x;
  const int cnt = counts[c];
  const int want = donor_of ? donor_of[c] : -1;
  if (cnt > 0 && want < 0) {
    const float inv = 1.f / static_cast<float>(cnt);
    for (int j = threadIdx.x; j < dim; j += blockDim.x)
      cent[static_cast<size_t>(c) * dim + j] = sums[static_cast<size_t>(c) * dim + j] * inv;
    return;
  }
  __shared__ long long donor_row;
  if (threadIdx.x == 0) {
    long long row = 0;
    // rejection-sample a row of the donor cluster (it is over-full, so this ends quickly)
    for (int attempt = 0; attempt < 8192; ++attempt) {
      row = static_cast<long long>(mix64(seed ^ (0xA5ull * (c + 1)) ^ (0x9E3779B9ull * attempt)) %
                                   static_cast<uint64_t>(n));
      if (want < 0 || labels[row] == want) break;
    }
    donor_row = row;
  }
  __syncthreads();
  const long long row = donor_row;
  for (int j = threadIdx.x; j < dim; j += blockDim.x)
    cent[static_cast<size_t>(c) * dim + j] = ld_f32<T>(x + row * dim + j);
}

// Host side of the balancing step: clusters above 1.5x the average size want floor(size/avg) - 1
// extra centroids; they are taken from the smallest clusters below 0.5x the average.
static void balance_pairs(const std::vector<int>& counts, int64_t n, std::vector<int>* donor_of) {
  const int ncl = static_cast<int>(counts.size());
  donor_of->assign(ncl, -1);
  const double avg = static_cast<double>(n) / ncl;
  std::vector<int> order(ncl);
  for (int i = 0; i < ncl; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int a, int b) { return counts[a] < counts[b]; });
  int lo = 0, hi = ncl - 1;
  while (lo < hi) {
    const int big = order[hi];
    if (counts[big] <= 1.5 * avg) break;
    int quota = static_cast<int>(counts[big] / avg) - 1;
    if (quota < 1) quota = 1;
    while (quota > 0 && lo < hi && counts[order[lo]] < 0.5 * avg) {
      (*donor_of)[order[lo]] = big;
      ++lo;
      --quota;
    }
    if (quota > 0) break;  // no small clusters left to move
    --hi;
  }
}

// Temporaries of one k-means fit.  A caller that fits many small problems in a row (the 64+ PQ
// sub-codebooks) passes the same workspace to every fit, so device memory is allocated once
// instead of being malloc'ed and freed (= device-synchronised) per fit.
struct KmWorkspace {
  DevBuf sums, counts, labels, donors, seg_off, seg_cur, seg_rows, seg_slot;
  FlatEngine eng;
  void release() {
    for (DevBuf* b : {&sums, &counts, &labels, &donors, &seg_off, &seg_cur, &seg_rows, &seg_slot}) b->release();
    eng.destroy();
  }
};

static int kmeans_fit_impl(int dev, int dtype, int dim, const void* x, int64_t n, int ncl, int iters,
                           uint64_t seed, float* cent, int32_t* labels_out, cudaStream_t st,
                           KmWorkspace* shared_ws = nullptr) {
  B2VS_CHECK(n >= 1 && ncl >= 1 && ncl <= n, B2VS_EINVAL,
             "k-means needs 1 <= n_clusters <= n (n_clusters=%d, n=%lld)", ncl,
             static_cast<long long>(n));
  B2VS_CHECK(n < (1ll << 31), B2VS_EINVAL, "k-means input too large (n=%lld)", static_cast<long long>(n));
  KmWorkspace local_ws;
  KmWorkspace& w = shared_ws ? *shared_ws : local_ws;
  DevBuf &sums = w.sums, &counts = w.counts, &labels = w.labels, &donors = w.donors,
         &seg_off = w.seg_off, &seg_cur = w.seg_cur, &seg_rows = w.seg_rows, &seg_slot = w.seg_slot;
  FlatEngine& eng = w.eng;
  std::vector<int> h_counts, h_donor;
  int rc = B2VS_OK;
  auto cleanup = [&]() {
    if (!shared_ws) local_ws.release();   // a shared workspace is released by its owner
  };
#define KM_TRY(expr) do { rc = (expr); if (rc != B2VS_OK) { cleanup(); return rc; } } while (0)
#define KM_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); cleanup(); return B2VS_ECUDA; } } while (0)
  KM_TRY(sums.reserve(static_cast<size_t>(ncl) * dim * sizeof(float)));
  KM_TRY(counts.reserve(static_cast<size_t>(ncl) * sizeof(int)));
  KM_TRY(donors.reserve(static_cast<size_t>(ncl) * sizeof(int)));
  KM_TRY(seg_off.reserve(static_cast<size_t>(ncl + 1) * sizeof(uint32_t)));
  KM_TRY(seg_cur.reserve(static_cast<size_t>(ncl) * sizeof(int)));
  KM_TRY(seg_rows.reserve(static_cast<size_t>(n) * sizeof(uint32_t)));
  KM_TRY(seg_slot.reserve(static_cast<size_t>(n) * sizeof(uint32_t)));
  int32_t* lab = labels_out;
  if (!lab) {
    KM_TRY(labels.reserve(static_cast<size_t>(n) * sizeof(int32_t)));
    lab = labels.as<int32_t>();
  }
  const int fmt = (dtype == B2VS_F16) ? 0 : 1;
  const int force = (dtype == B2VS_F32) ? -1 : fmt;
  DISPATCH_DTYPE(dtype, T, (seed_centroids_kernel<T><<<ncl, 128, 0, st>>>(
                               static_cast<const T*>(x), n, dim, ncl, seed, cent)));
  KM_CUDA(cudaGetLastError());
  const int acc_blocks = static_cast<int>(std::min<int64_t>(ceil_div(n, 8), 148 * 16));
  for (int it = 0; it < iters; ++it) {
    KM_TRY(eng.init(dev, B2VS_METRIC_L2, B2VS_F32, dim, cent, ncl, st, force));
    KM_TRY(eng.search(x, dtype, static_cast<int>(n), 1, 0, 0, nullptr, nullptr, lab, st));
    KM_CUDA(cudaMemsetAsync(counts.ptr, 0, static_cast<size_t>(ncl) * sizeof(int), st));
    KM_CUDA(cudaMemsetAsync(seg_cur.ptr, 0, static_cast<size_t>(ncl) * sizeof(int), st));
    histogram_kernel<<<acc_blocks, 256, 0, st>>>(lab, n, counts.as<int>());
    scan_sizes_kernel<<<1, 1024, 0, st>>>(counts.as<int>(), ncl, 1, seg_off.as<uint32_t>());
    scatter_rows_kernel<<<acc_blocks, 256, 0, st>>>(lab, n, seg_off.as<uint32_t>(), seg_cur.as<int>(),
                                                    seg_rows.as<uint32_t>(), seg_slot.as<uint32_t>());
    KM_CUDA(cudaGetLastError());
    {
      const int tpb = dim >= 256 ? 256 : ((dim + 31) / 32) * 32;
      const dim3 grid(ncl, static_cast<unsigned>(ceil_div(dim, tpb)));
      DISPATCH_DTYPE(dtype, T, (segment_sum_kernel<T><<<grid, tpb, 0, st>>>(
                                   static_cast<const T*>(x), seg_off.as<uint32_t>(),
                                   seg_rows.as<uint32_t>(), dim, sums.as<float>())));
    }
    KM_CUDA(cudaGetLastError());
    const int* donor_ptr = nullptr;
    if (it + 2 < iters && ncl > 1) {  // the last two iterations are plain Lloyd
      h_counts.resize(ncl);
      KM_CUDA(cudaMemcpyAsync(h_counts.data(), counts.ptr, static_cast<size_t>(ncl) * sizeof(int),
                              cudaMemcpyDeviceToHost, st));
      KM_CUDA(cudaStreamSynchronize(st));
      balance_pairs(h_counts, n, &h_donor);
      KM_CUDA(cudaMemcpyAsync(donors.ptr, h_donor.data(), static_cast<size_t>(ncl) * sizeof(int),
                              cudaMemcpyHostToDevice, st));
      donor_ptr = donors.as<int>();
    }
    DISPATCH_DTYPE(dtype, T, (finalize_centroids_kernel<T><<<ncl, 128, 0, st>>>(
                                 static_cast<const T*>(x), lab, n, dim, sums.as<float>(),
                                 counts.as<int>(), donor_ptr, seed + 977ull * (it + 1), cent)));
    KM_CUDA(cudaGetLastError());
  }
  if (labels_out) {
    KM_TRY(eng.init(dev, B2VS_METRIC_L2, B2VS_F32, dim, cent, ncl, st, force));
    KM_TRY(eng.search(x, dtype, static_cast<int>(n), 1, 0, 0, nullptr, nullptr, labels_out, st));
  }
  KM_CUDA(cudaStreamSynchronize(st));  // temporaries below are freed; make sure nothing is in flight
  cleanup();
#undef KM_TRY
#undef KM_CUDA
  return B2VS_OK;
}

// ---- K3 list construction -----------------------------------------------------------------
__global__ void histogram_kernel(const int* __restrict__ labels, int64_t n, int* __restrict__ sizes) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = labels[i];
    if (c >= 0) atomicAdd(sizes + c, 1);
  }
}

// Exclusive scan of list sizes rounded up to `pad`; single block.
__global__ void scan_sizes_kernel(const int* __restrict__ sizes, int n_lists, int pad,
                                  uint32_t* __restrict__ offsets) {
  __shared__ uint32_t part[1024];
  const int t = threadIdx.x;
  const int per = (n_lists + blockDim.x - 1) / blockDim.x;
  const int lo = t * per, hi = min(n_lists, lo + per);
  uint32_t s = 0;
  for (int i = lo; i < hi; ++i) s += static_cast<uint32_t>((sizes[i] + pad - 1) / pad * pad);
  part[t] = s;
  __syncthreads();
  if (t == 0) {
    uint32_t run = 0;
    for (int i = 0; i < blockDim.x; ++i) { const uint32_t v = part[i]; part[i] = run; run += v; }
    offsets[n_lists] = run;
  }
  __syncthreads();
  uint32_t run = part[t];
  for (int i = lo; i < hi; ++i) {
    offsets[i] = run;
    run += static_cast<uint32_t>((sizes[i] + pad - 1) / pad * pad);
  }
}

__global__ void scatter_rows_kernel(const int* __restrict__ labels, int64_t n,
                                    const uint32_t* __restrict__ offsets, int* __restrict__ cursor,
                                    uint32_t* __restrict__ row_ids, uint32_t* __restrict__ slot_of_row) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = labels[i];
    if (c < 0) { slot_of_row[i] = kNoRow; continue; }
    const uint32_t slot = offsets[c] + static_cast<uint32_t>(atomicAdd(cursor + c, 1));
    row_ids[slot] = static_cast<uint32_t>(i);
    slot_of_row[i] = slot;
  }
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

// one warp per row: copy the row into its list slot (16-bit storage) and record ||x||^2
__global__ void fill_f32_kernel(float* __restrict__ p, size_t n, float v) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

template <typename T>
__global__ void fill_flat_lists_kernel(const T* __restrict__ x, int64_t n, int dim, int dp, int fmt,
                                       const uint32_t* __restrict__ slot_of_row,
                                       uint16_t* __restrict__ data, float* __restrict__ slot_norm,
                                       int want_norm, unsigned int* __restrict__ max_norm_bits) {
  const int lane = threadIdx.x & 31;
  float warp_max = 0.f;
  const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = warp0; r < n; r += nwarps) {
    const uint32_t slot = slot_of_row[r];
    if (slot == kNoRow) continue;
    const T* row = x + r * dim;
    uint16_t* orow = data + static_cast<size_t>(slot) * dp;
    float acc = 0.f;
    for (int j = lane; j < dp; j += 32) {
      const float v = j < dim ? ld_f32<T>(row + j) : 0.f;
      float back;
      orow[j] = to_op16(v, fmt, &back);
      acc = fmaf(back, back, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) slot_norm[slot] = want_norm ? acc : 0.f;  // inner product: no additive term
    warp_max = fmaxf(warp_max, acc);
  }
  // largest ||x||^2 of the index (non-negative floats order like their bit patterns)
  if (lane == 0 && warp_max > 0.f) atomicMax(max_norm_bits, __float_as_uint(warp_max));
}

// ---- warp-resident sorted top-k list -------------------------------------------------------
struct WarpTopK {
  u64 acc[kListE];
  float tau;
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int e = 0; e < kListE; ++e) acc[e] = kKeyInf;
    tau = __int_as_float(0x7f800000);
  }
  // Each lane offers at most one candidate key (kKeyInf = none). Warp-collective.
  __device__ __forceinline__ void offer(u64 ck, int k, int lane) {
    if (!__any_sync(0xffffffffu, ck != kKeyInf)) return;
    u64 c1[1] = {ck};
    warp_bitonic_sort<1>(c1, lane);
#pragma unroll
    for (int e = 0; e < kListE; ++e) {
      const int i = lane * kListE + e;                       // list element index
      const u64 r = shfl_u64(c1[0], (32 * kListE - 1 - i) & 31);  // reversed candidate run
      if (i >= 32 * kListE - 32) acc[e] = acc[e] < r ? acc[e] : r;
    }
    warp_bitonic_merge<kListE>(acc, lane);
    u64 kth = kKeyInf;
#pragma unroll
    for (int e = 0; e < kListE; ++e)
      if (lane * kListE + e == k - 1) kth = acc[e];
    kth = shfl_u64(kth, (k - 1) / kListE);
    tau = (kth == kKeyInf) ? __int_as_float(0x7f800000) : key_score(kth);
  }
};

// Block epilogue shared by both scans: warps publish their lists, warp 0 folds them and writes
// the item's k sorted keys.
__device__ __forceinline__ void block_merge_and_store(WarpTopK& tk, u64 (*lists)[32 * kListE],
                                                      int k, int warp, int lane, u64* out) {
#pragma unroll
  for (int e = 0; e < kListE; ++e) lists[warp][lane * kListE + e] = tk.acc[e];
  __syncthreads();
  if (warp == 0) {
    for (int w = 1; w < kScanWarps; ++w) {
#pragma unroll
      for (int e = 0; e < kListE; ++e) {
        const int src = 32 * kListE - 1 - (lane * kListE + e);
        const u64 b = lists[w][src];
        tk.acc[e] = tk.acc[e] < b ? tk.acc[e] : b;
      }
      warp_bitonic_merge<kListE>(tk.acc, lane);
    }
#pragma unroll
    for (int e = 0; e < kListE; ++e) {
      const int i = lane * kListE + e;
      if (i < k) out[i] = tk.acc[e];
    }
  }
}

// ---- K5 IVF-Flat list scan -----------------------------------------------------------------
// One CTA per (query, probe).  A warp streams 32 consecutive list rows per batch, 4 rows at a
// time; each lane owns the same 16-byte chunks of every row, so its slice of the query stays in
// registers (J chunks of 8 elements).  128-bit loads, fp32 accumulate.
template <int FMT>
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (FMT == 1) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    } else {
      const __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
      const float2 t = __half22float2(h);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
}

// Loads lane's slice of query q into registers (J chunks of 8 elements).
template <int J>
__device__ __forceinline__ void load_query_regs(const float* __restrict__ qf, int q, int dp, int lane,
                                                float (&qr)[J][8]) {
  const int n_chunks = dp >> 3;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int c = lane + 32 * j;
#pragma unroll
    for (int e = 0; e < 8; ++e) qr[j][e] = (c < n_chunks) ? qf[static_cast<size_t>(q) * dp + c * 8 + e] : 0.f;
  }
}

// Streams list rows [begin, end) through the CTA's warps and offers score = alpha*dot + slot_norm
// to each warp's top-k list.  Padding slots carry slot_norm = +inf and never qualify.
template <int FMT, int J>
__device__ __forceinline__ void scan_list_rows(const uint4* __restrict__ data4,
                                               const float* __restrict__ slot_norm, uint32_t begin,
                                               uint32_t end, const float (&qr)[J][8], int n_chunks,
                                               float alpha, WarpTopK& tk, int k, int warp, int lane) {
  for (uint32_t b0 = begin + warp * 32; b0 < end; b0 += kScanWarps * 32) {
    u64 ck = kKeyInf;
#pragma unroll 1
    for (int it = 0; it < 8; ++it) {
      const uint32_t r0 = b0 + it * 4;
      if (r0 >= end) break;
      uint4 v[4][J];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const uint32_t row = r0 + r;
#pragma unroll
        for (int j = 0; j < J; ++j) {
          const int c = lane + 32 * j;
          v[r][j] = make_uint4(0, 0, 0, 0);
          if (row < end && c < n_chunks) v[r][j] = __ldg(data4 + static_cast<size_t>(row) * n_chunks + c);
        }
      }
      float dot[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < J; ++j) {
          float f[8];
          unpack8<FMT>(v[r][j], f);
#pragma unroll
          for (int e = 0; e < 8; ++e) a = fmaf(f[e], qr[j][e], a);
        }
        dot[r] = a;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int r = 0; r < 4; ++r) dot[r] += __shfl_xor_sync(0xffffffffu, dot[r], o);
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const uint32_t row = r0 + r;
        if (lane == it * 4 + r && row < end) {
          const float sc = fmaf(alpha, dot[r], slot_norm[row]);
          if (sc < tk.tau) ck = pack_key(sc, row);
        }
      }
    }
    tk.offer(ck, k, lane);
  }
}

template <int FMT, int J>
__global__ void __launch_bounds__(kScanThreads, 2)
ivf_flat_scan_kernel(const uint16_t* __restrict__ data, const float* __restrict__ slot_norm,
                     const uint32_t* __restrict__ offsets, const long long* __restrict__ probe_ids,
                     const float* __restrict__ qf, int dp, int n_probes, int q_pad, int k,
                     float alpha, u64* __restrict__ out_keys,
                     unsigned long long* __restrict__ scanned_rows,
                     const uint32_t* __restrict__ item_perm) {
  __shared__ u64 lists[kScanWarps][32 * kListE];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // item_perm (optional) orders the items by list, so the CTAs resident at one time read the
  // same few lists and all but the first touch of a row is served by L2
  const int item = item_perm ? static_cast<int>(item_perm[blockIdx.x]) : blockIdx.x;
  const int q = item / n_probes, p = item - q * n_probes;
  const long long list = probe_ids[item];
  uint32_t begin = 0, end = 0;
  if (list >= 0) { begin = offsets[list]; end = offsets[list + 1]; }
  if (threadIdx.x == 0 && scanned_rows) atomicAdd(scanned_rows, static_cast<unsigned long long>(end - begin));
  float qr[J][8];
  load_query_regs<J>(qf, q, dp, lane, qr);
  WarpTopK tk;
  tk.init();
  scan_list_rows<FMT, J>(reinterpret_cast<const uint4*>(data), slot_norm, begin, end, qr, dp >> 3,
                         alpha, tk, k, warp, lane);
  block_merge_and_store(tk, lists, k, warp, lane,
                        out_keys + (static_cast<size_t>(p) * q_pad + q) * k);
}

// ---- K5b grouped IVF-Flat scan (large batches) ----------------------------------------------
// When a batch holds many queries per list, the (query, probe) items are grouped by list and each
// list is multiplied against the block of queries that probe it on the tensor cores
// (bf_tc_kernel<1, true>, work-table mode).  The pieces around that kernel:
//   seed    per query: k-th best score of the first rows of its NEAREST list = a valid upper
//           bound of its final k-th score; every candidate below it is appended to the query's
//           buffer by the tensor-core kernel
//   work    one item per (list, 128-row slice of its query group)
//   gather  the 16-bit query operand, rows in group order
//   select  per query: sort the appended candidates, keep k
//   rescue  queries whose buffer overflowed (threshold too loose) are rescanned exactly
template <int FMT, int J>
__global__ void __launch_bounds__(kScanThreads, 2)
ivf_seed_tau_kernel(const uint16_t* __restrict__ data, const float* __restrict__ slot_norm,
                    const uint32_t* __restrict__ offsets, const long long* __restrict__ probe_ids,
                    const float* __restrict__ qf, int dp, int n_probes, int k, float alpha,
                    uint32_t row_limit, float max_norm2, int l2, float extra_eps,
                    const uint32_t* __restrict__ q_perm, float* __restrict__ tau) {
  __shared__ u64 lists[kScanWarps][32 * kListE];
  __shared__ u64 top[kMaxFusedK];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // large batches: queries ordered by nearest list (L2 reuse); small ones skip the sort
  const int q = q_perm ? static_cast<int>(q_perm[blockIdx.x]) : static_cast<int>(blockIdx.x);
  const long long list = probe_ids[static_cast<size_t>(q) * n_probes];
  uint32_t begin = 0, end = 0;
  if (list >= 0) { begin = offsets[list]; end = min(offsets[list + 1], begin + row_limit); }
  float qr[J][8];
  load_query_regs<J>(qf, q, dp, lane, qr);
  WarpTopK tk;
  tk.init();
  scan_list_rows<FMT, J>(reinterpret_cast<const uint4*>(data), slot_norm, begin, end, qr, dp >> 3,
                         alpha, tk, k, warp, lane);
  block_merge_and_store(tk, lists, k, warp, lane, top);
  float qn = 0.f;
#pragma unroll
  for (int j = 0; j < J; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) qn = fmaf(qr[j][e], qr[j][e], qn);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) qn += __shfl_xor_sync(0xffffffffu, qn, o);
  __syncthreads();
  if (threadIdx.x == 0) {
    // The tensor-core kernel recomputes these scores with a different summation order: the
    // threshold gets a cushion that bounds the fp32 rounding difference of two length-dp dot
    // products (|q.x| <= ||q|| ||x||_max), so no row of the true top-k can fall outside it.
    const u64 kth = top[k - 1];
    float t = INFINITY;
    if (kth != kKeyInf) {
      const float sc = key_score(kth);
      const float eps = static_cast<float>(dp) * 1.2e-7f + 1e-6f + extra_eps;
      t = sc + fabsf(alpha) * eps * sqrtf(qn * max_norm2) + 4e-7f * (fabsf(sc) + (l2 ? max_norm2 : 0.f));
    }
    tau[q] = t;
  }
}

// One thread per list: emits the list's work items.  group_off = exclusive scan (in size-rank
// order) of the per-list query counts rounded up to 128, in gathered-row units.
__global__ void build_group_work_kernel(const uint32_t* __restrict__ group_off,
                                        const uint32_t* __restrict__ offsets,
                                        const int* __restrict__ group_cnt,
                                        const int* __restrict__ list_of_rank, int n_lists,
                                        int chunk_rows, int slots,
                                        int4* __restrict__ work, int* __restrict__ n_work,
                                        unsigned long long* __restrict__ scanned_rows) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;   // size rank: items come out longest first
  if (r == 0) *n_work = static_cast<int>(group_off[n_lists] >> 7) * slots;
  if (r >= n_lists) return;
  const int l = list_of_rank[r];
  const int b0 = static_cast<int>(group_off[r] >> 7), b1 = static_cast<int>(group_off[r + 1] >> 7);
  const int begin = static_cast<int>(offsets[l]), end = static_cast<int>(offsets[l + 1]);
  // Small batches have fewer (list, query block) pairs than SMs: a list is then cut into `slots`
  // row ranges of chunk_rows (a multiple of the 256-row tile), one work item each - append mode
  // keeps no per-item state, so the pieces are independent.  Ranges past the list end are empty.
  for (int b = b0; b < b1; ++b)
    for (int c = 0; c < slots; ++c) {
      const int rb = min(end, begin + c * chunk_rows);
      const int re = c + 1 == slots ? end : min(end, rb + chunk_rows);
      work[b * slots + c] = make_int4(b, rb, re, 0);
    }
  if (scanned_rows && group_cnt[r] > 0)   // algorithmic work: every probing query sees every row
    atomicAdd(scanned_rows, static_cast<unsigned long long>(group_cnt[r]) * static_cast<unsigned>(end - begin));
}

// One warp per gathered row: row_item[v] = (query, probe) item or kNoRow on group padding.
__global__ void gather_group_queries_kernel(const uint32_t* __restrict__ row_item,
                                            const uint32_t* __restrict__ group_off, int n_lists,
                                            const float* __restrict__ qf, int dp, int n_probes,
                                            int fmt, int split, uint16_t* __restrict__ out,
                                            int* __restrict__ row_query) {
  const int lane = threadIdx.x & 31;
  const int64_t v = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (v >= static_cast<int64_t>(group_off[n_lists])) return;
  const uint32_t item = row_item[v];
  const int q = item == kNoRow ? -1 : static_cast<int>(item / static_cast<uint32_t>(n_probes));
  if (lane == 0) row_query[v] = q;
  if (kSkipPaddingRows && q < 0) return;   // never qualifies (threshold -inf): bytes are don't-care
  if (!split) {
    uint16_t* orow = out + static_cast<size_t>(v) * dp;
    for (int j = lane; j < dp; j += 32) {
      float back;
      orow[j] = q < 0 ? uint16_t(0) : to_op16(qf[static_cast<size_t>(q) * dp + j], fmt, &back);
    }
    return;
  }
  // fp32 queries: [hi | lo] bf16 halves, each padded to a multiple of 64 columns
  const int half = (dp + 63) & ~63;
  uint16_t* orow = out + static_cast<size_t>(v) * 2 * half;
  for (int j = lane; j < half; j += 32) {
    uint16_t hi = 0, lo = 0;
    if (q >= 0 && j < dp) {
      const float x = qf[static_cast<size_t>(q) * dp + j];
      float hb, lb;
      hi = to_op16(x, 1, &hb);
      lo = to_op16(x - hb, 1, &lb);
    }
    orow[j] = hi;
    orow[half + j] = lo;
  }
}

// One CTA per query: bitonic sort of the appended candidates in shared memory, first k kept.
constexpr int kSelectThreads = 256;
__global__ void __launch_bounds__(kSelectThreads)
ivf_group_select_kernel(const u64* __restrict__ cand, const int* __restrict__ count, int cap, int k,
                        u64* __restrict__ out_keys, unsigned long long* __restrict__ total_cand) {
  extern __shared__ u64 sk[];
  const int q = blockIdx.x;
  const int n = count[q];
  if (threadIdx.x == 0 && total_cand) atomicAdd(total_cand, static_cast<unsigned long long>(n));
  if (n > cap) return;  // overflow: left to the rescue kernel
  int P = 32;
  while (P < n) P <<= 1;
  const u64* src = cand + static_cast<size_t>(q) * cap;
  for (int i = threadIdx.x; i < P; i += blockDim.x) sk[i] = i < n ? __ldcg(src + i) : kKeyInf;
  __syncthreads();
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool up = (lo & size) == 0;
        const u64 a = sk[lo], b = sk[hi];
        if ((a > b) == up) { sk[lo] = b; sk[hi] = a; }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < k; i += blockDim.x) out_keys[static_cast<size_t>(q) * k + i] = i < P ? sk[i] : kKeyInf;
}

// One CTA per query whose candidate buffer overflowed: exact scan of all its probes.
template <int FMT, int J>
__global__ void __launch_bounds__(kScanThreads, 2)
ivf_flat_rescue_kernel(const uint16_t* __restrict__ data, const float* __restrict__ slot_norm,
                       const uint32_t* __restrict__ offsets, const long long* __restrict__ probe_ids,
                       const float* __restrict__ qf, int dp, int n_probes, int k, float alpha,
                       const int* __restrict__ count, int cap, u64* __restrict__ out_keys) {
  __shared__ u64 lists[kScanWarps][32 * kListE];
  const int q = blockIdx.x;
  if (count[q] <= cap) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float qr[J][8];
  load_query_regs<J>(qf, q, dp, lane, qr);
  WarpTopK tk;
  tk.init();
  for (int p = 0; p < n_probes; ++p) {
    const long long list = probe_ids[static_cast<size_t>(q) * n_probes + p];
    if (list < 0) continue;
    scan_list_rows<FMT, J>(reinterpret_cast<const uint4*>(data), slot_norm, offsets[list],
                           offsets[list + 1], qr, dp >> 3, alpha, tk, k, warp, lane);
  }
  block_merge_and_store(tk, lists, k, warp, lane, out_keys + static_cast<size_t>(q) * k);
}

// ---- K6 / K7 IVF-PQ -----------------------------------------------------------------------
// residual sub-vectors of the training rows, laid out [pq_dim][n_train][dsub] fp32
template <typename T>
__global__ void pq_train_slices_kernel(const T* __restrict__ x, const int* __restrict__ labels,
                                       const float* __restrict__ cent, int64_t n_train,
                                       int64_t stride, int dim, int dsub, float* __restrict__ out) {
  const int64_t total = n_train * dim;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t t = i / dim;
    const int j = static_cast<int>(i - t * dim);
    const int64_t r = t * stride;
    const int c = labels[r];
    const float res = ld_f32<T>(x + r * dim + j) - cent[static_cast<size_t>(c) * dim + j];
    const int m = j / dsub, d = j - m * dsub;
    out[(static_cast<size_t>(m) * n_train + t) * dsub + d] = res;
  }
}

// grid (row blocks, pq_dim): each block holds one sub-codebook in smem; thread = row
template <typename T>
__global__ void __launch_bounds__(256)
pq_encode_kernel(const T* __restrict__ x, const int* __restrict__ labels,
                 const float* __restrict__ cent, const float* __restrict__ codebooks,
                 const uint32_t* __restrict__ slot_of_row, int64_t n, int dim, int dsub, int mp,
                 uint8_t* __restrict__ codes) {
  extern __shared__ float cb[];  // [256][dsub]
  const int m = blockIdx.y;
  for (int i = threadIdx.x; i < 256 * dsub; i += blockDim.x)
    cb[i] = codebooks[static_cast<size_t>(m) * 256 * dsub + i];
  __syncthreads();
  const int n_chunks = mp >> 4;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < n;
       r += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint32_t slot = slot_of_row[r];
    if (slot == kNoRow) continue;
    const int c = labels[r];
    float res[16];
    for (int d = 0; d < dsub; ++d)
      res[d] = ld_f32<T>(x + r * dim + m * dsub + d) - cent[static_cast<size_t>(c) * dim + m * dsub + d];
    float best = __int_as_float(0x7f800000);
    int best_j = 0;
    for (int j = 0; j < 256; ++j) {
      float s = 0.f;
      for (int d = 0; d < dsub; ++d) {
        const float t = res[d] - cb[j * dsub + d];
        s = fmaf(t, t, s);
      }
      if (s < best) { best = s; best_j = j; }
    }
    const uint32_t g = slot >> 5, l = slot & 31;
    codes[((static_cast<size_t>(g) * n_chunks + (m >> 4)) * 32 + l) * 16 + (m & 15)] =
        static_cast<uint8_t>(best_j);
  }
}

// One CTA per (query, probe): build the [pq_dim][256] LUT in smem, then every lane scores one
// row of a 32-row group per step (16-byte coalesced code loads from the interleaved layout).
__global__ void __launch_bounds__(kScanThreads)
ivf_pq_scan_kernel(const uint8_t* __restrict__ codes, const uint32_t* __restrict__ row_ids,
                   const uint32_t* __restrict__ offsets, const long long* __restrict__ probe_ids,
                   const float* __restrict__ qf, const float* __restrict__ cent,
                   const float* __restrict__ codebooks, int dim, int dp, int pq_dim, int mp, int dsub,
                   int n_probes, int q_pad, int k, int metric, u64* __restrict__ out_keys,
                   unsigned long long* __restrict__ scanned_rows) {
  extern __shared__ float smem_f[];
  float* lut = smem_f;                 // [mp * 256]
  float* rq = smem_f + mp * 256;       // [dim]
  __shared__ u64 lists[kScanWarps][32 * kListE];
  __shared__ float bias_part[kScanWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x;
  const int q = item / n_probes, p = item - q * n_probes;
  const long long list = probe_ids[item];
  uint32_t begin = 0, end = 0;
  if (list >= 0) { begin = offsets[list]; end = offsets[list + 1]; }
  if (threadIdx.x == 0 && scanned_rows) atomicAdd(scanned_rows, static_cast<unsigned long long>(end - begin));
  const float* c = cent + static_cast<size_t>(list < 0 ? 0 : list) * dim;
  float bpart = 0.f;
  for (int d = threadIdx.x; d < dim; d += blockDim.x) {
    const float qv = qf[static_cast<size_t>(q) * dp + d];
    if (metric == B2VS_METRIC_L2) rq[d] = qv - c[d];
    else { rq[d] = qv; bpart = fmaf(qv, c[d], bpart); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) bpart += __shfl_xor_sync(0xffffffffu, bpart, o);
  if (lane == 0) bias_part[warp] = bpart;
  __syncthreads();
  float bias = 0.f;
#pragma unroll
  for (int w = 0; w < kScanWarps; ++w) bias += bias_part[w];
  bias = -bias;  // IP score = -(q.c + sum q.cb)
  for (int idx = threadIdx.x; idx < mp * 256; idx += blockDim.x) {
    const int m = idx >> 8, j = idx & 255;
    float s = 0.f;
    if (m < pq_dim) {
      const float* cbp = codebooks + (static_cast<size_t>(m) * 256 + j) * dsub;
      for (int d = 0; d < dsub; ++d) {
        if (metric == B2VS_METRIC_L2) {
          const float t = rq[m * dsub + d] - cbp[d];
          s = fmaf(t, t, s);
        } else {
          s = fmaf(-rq[m * dsub + d], cbp[d], s);
        }
      }
    }
    lut[idx] = s;
  }
  __syncthreads();

  WarpTopK tk;
  tk.init();
  const int n_chunks = mp >> 4;
  const uint4* codes4 = reinterpret_cast<const uint4*>(codes);
  for (uint32_t g0 = (begin >> 5) + warp; g0 < (end >> 5); g0 += kScanWarps) {
    const uint32_t slot = (g0 << 5) + lane;
    float s = bias;
    for (int ch = 0; ch < n_chunks; ++ch) {
      const uint4 v = __ldg(codes4 + (static_cast<size_t>(g0) * n_chunks + ch) * 32 + lane);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int m = ch * 16 + i * 4 + b;
          s += lut[m * 256 + ((w[i] >> (8 * b)) & 0xFFu)];
        }
      }
    }
    u64 ck = kKeyInf;
    if (row_ids[slot] != kNoRow && s < tk.tau) ck = pack_key(s, slot);
    tk.offer(ck, k, lane);
  }
  block_merge_and_store(tk, lists, k, warp, lane,
                        out_keys + (static_cast<size_t>(p) * q_pad + q) * k);
}

// Query-major persistent variant of the PQ scan for indexes whose codebooks fit in shared memory
// (pq_dim * 256 * dsub floats <= 128 KB, e.g. C4: M = 64, dsub = 2).  One CTA per SM keeps the
// codebooks resident and owns whole queries: it walks the query's probes, rebuilding the LUT from
// smem for each list, while every warp's sorted top-k list and threshold persist ACROSS the
// probes.  Compared with one CTA per (query, probe) this removes the 128 KB L2 read per LUT, the
// threshold warm-up of every list and all but one block-level merge per query.
constexpr int kPqPersistThreads = 512;
constexpr int kPqPersistWarps = kPqPersistThreads / 32;

template <int NCH, int DSUB>  // code chunks per row (mp / 16) and sub-vector length; 0 = runtime
__global__ void __launch_bounds__(kPqPersistThreads, 1)
ivf_pq_scan_query_kernel(const uint8_t* __restrict__ codes, const uint32_t* __restrict__ row_ids,
                         const uint32_t* __restrict__ offsets, const long long* __restrict__ probe_ids,
                         const float* __restrict__ qf, const float* __restrict__ cent,
                         const float* __restrict__ codebooks, int dim, int dp, int pq_dim, int mp,
                         int dsub, int n_probes, int nq, int k, int metric, u64* __restrict__ out_keys,
                         unsigned long long* __restrict__ scanned_rows) {
  extern __shared__ float smem_f[];
  float* cb = smem_f;                                // [pq_dim * 256 * dsub]
  float* lut = cb + pq_dim * 256 * dsub;             // [mp * 256]
  float* rq = lut + mp * 256;                        // [dim] residual query of the current probe
  float* sq = rq + dim;                              // [dim] the query
  __shared__ u64 lists[kPqPersistWarps][32 * kListE];
  __shared__ float bias_part[kPqPersistWarps];
  __shared__ int s_list[kMaxFusedK];
  __shared__ uint32_t s_begin[kMaxFusedK], s_end[kMaxFusedK];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < pq_dim * 256 * dsub; i += blockDim.x) cb[i] = codebooks[i];
  for (int i = pq_dim * 256 + threadIdx.x; i < mp * 256; i += blockDim.x) lut[i] = 0.f;  // padding rows
  unsigned long long rows_seen = 0;
  constexpr int kV = NCH > 0 ? NCH : 1;
  const int n_chunks = NCH > 0 ? NCH : (mp >> 4);
  const uint4* codes4 = reinterpret_cast<const uint4*>(codes);
  for (int q = blockIdx.x; q < nq; q += gridDim.x) {
    __syncthreads();  // previous query fully retired (lists, s_*, sq)
    // the query's probe lists and their extents, fetched once in parallel
    if (threadIdx.x < n_probes) {
      const long long l = probe_ids[static_cast<size_t>(q) * n_probes + threadIdx.x];
      s_list[threadIdx.x] = static_cast<int>(l);
      s_begin[threadIdx.x] = l >= 0 ? offsets[l] : 0u;
      s_end[threadIdx.x] = l >= 0 ? offsets[l + 1] : 0u;
    }
    for (int d = threadIdx.x; d < dim; d += blockDim.x) sq[d] = qf[static_cast<size_t>(q) * dp + d];
    __syncthreads();
    WarpTopK tk;
    tk.init();
    // centroid values of the NEXT probe travel in registers while the current probe is scanned
    float c_next[2] = {0.f, 0.f};
    {
      const int l0 = s_list[0];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int d = threadIdx.x + t * kPqPersistThreads;
        if (d < dim && l0 >= 0) c_next[t] = cent[static_cast<size_t>(l0) * dim + d];
      }
    }
    for (int p = 0; p < n_probes; ++p) {
      const int list = s_list[p];
      const uint32_t begin = s_begin[p], end = s_end[p];
      rows_seen += (threadIdx.x == 0) ? (end - begin) : 0u;
      // issue this warp's first code loads now: they complete during the rq / LUT phases
      const uint32_t g_first = (begin >> 5) + warp;
      uint4 v[kV];
      uint32_t rid = kNoRow;
      if (NCH > 0 && g_first < (end >> 5)) {
#pragma unroll
        for (int ch = 0; ch < kV; ++ch)
          v[ch] = __ldg(codes4 + (static_cast<size_t>(g_first) * kV + ch) * 32 + lane);
        rid = __ldg(row_ids + (g_first << 5) + lane);
      }
      float bpart = 0.f;
      const float c_cur[2] = {c_next[0], c_next[1]};
      if (p + 1 < n_probes) {
        const int ln = s_list[p + 1];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int d = threadIdx.x + t * kPqPersistThreads;
          if (d < dim && ln >= 0) c_next[t] = cent[static_cast<size_t>(ln) * dim + d];
        }
      }
      __syncthreads();  // the previous probe's LUT / rq are no longer read
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int d = threadIdx.x + t * kPqPersistThreads;
        if (d < dim) {
          const float qv = sq[d];
          if (metric == B2VS_METRIC_L2) rq[d] = qv - c_cur[t];
          else { rq[d] = qv; bpart = fmaf(qv, c_cur[t], bpart); }
        }
      }
      for (int d = threadIdx.x + 2 * kPqPersistThreads; d < dim; d += kPqPersistThreads) {  // dim > 1024
        const float qv = sq[d];
        const float cv = list >= 0 ? cent[static_cast<size_t>(list) * dim + d] : 0.f;
        if (metric == B2VS_METRIC_L2) rq[d] = qv - cv;
        else { rq[d] = qv; bpart = fmaf(qv, cv, bpart); }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) bpart += __shfl_xor_sync(0xffffffffu, bpart, o);
      if (lane == 0) bias_part[warp] = bpart;
      __syncthreads();
      float bias = 0.f;
#pragma unroll
      for (int w = 0; w < kPqPersistWarps; ++w) bias += bias_part[w];
      bias = -bias;
      if (DSUB > 0) {
        // specialised LUT build: sub-vector length known at compile time, metric hoisted
        if (metric == B2VS_METRIC_L2) {
          for (int idx = threadIdx.x; idx < pq_dim * 256; idx += kPqPersistThreads) {
            const float* cbp = cb + idx * DSUB;
            const float* rqm = rq + (idx >> 8) * DSUB;
            float sacc = 0.f;
#pragma unroll
            for (int d = 0; d < DSUB; ++d) {
              const float t = rqm[d] - cbp[d];
              sacc = fmaf(t, t, sacc);
            }
            lut[idx] = sacc;
          }
        } else {
          for (int idx = threadIdx.x; idx < pq_dim * 256; idx += kPqPersistThreads) {
            const float* cbp = cb + idx * DSUB;
            const float* rqm = rq + (idx >> 8) * DSUB;
            float sacc = 0.f;
#pragma unroll
            for (int d = 0; d < DSUB; ++d) sacc = fmaf(-rqm[d], cbp[d], sacc);
            lut[idx] = sacc;
          }
        }
      } else {
        for (int idx = threadIdx.x; idx < pq_dim * 256; idx += blockDim.x) {
          const int m = idx >> 8;
          const float* cbp = cb + static_cast<size_t>(idx) * dsub;
          float sacc = 0.f;
          for (int d = 0; d < dsub; ++d) {
            if (metric == B2VS_METRIC_L2) {
              const float t = rq[m * dsub + d] - cbp[d];
              sacc = fmaf(t, t, sacc);
            } else {
              sacc = fmaf(-rq[m * dsub + d], cbp[d], sacc);
            }
          }
          lut[idx] = sacc;
        }
      }
      __syncthreads();
      for (uint32_t g0 = g_first; g0 < (end >> 5); g0 += kPqPersistWarps) {
        const uint32_t slot = (g0 << 5) + lane;
        if (NCH == 0 || g0 != g_first) {
          if (NCH > 0) {
#pragma unroll
            for (int ch = 0; ch < kV; ++ch)
              v[ch] = __ldg(codes4 + (static_cast<size_t>(g0) * kV + ch) * 32 + lane);
          }
          rid = __ldg(row_ids + slot);
        }
        // four independent partial sums keep the LDS -> FADD chains short
        float part[4] = {bias, 0.f, 0.f, 0.f};
        if (NCH > 0) {
#pragma unroll
          for (int cc = 0; cc < kV; ++cc) {
            const uint32_t w[4] = {v[cc].x, v[cc].y, v[cc].z, v[cc].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
              for (int b = 0; b < 4; ++b)
                part[b] += lut[(cc * 16 + i * 4 + b) * 256 + ((w[i] >> (8 * b)) & 0xFFu)];
            }
          }
        } else {
          for (int cc = 0; cc < n_chunks; ++cc) {
            const uint4 vv = __ldg(codes4 + (static_cast<size_t>(g0) * n_chunks + cc) * 32 + lane);
            const uint32_t w[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
              for (int b = 0; b < 4; ++b)
                part[b] += lut[(cc * 16 + i * 4 + b) * 256 + ((w[i] >> (8 * b)) & 0xFFu)];
            }
          }
        }
        const float sacc = (part[0] + part[1]) + (part[2] + part[3]);
        u64 ck = kKeyInf;
        if (rid != kNoRow && sacc < tk.tau) ck = pack_key(sacc, slot);
        tk.offer(ck, k, lane);
      }
    }
    // one fold of the warps' lists per query
#pragma unroll
    for (int e = 0; e < kListE; ++e) lists[warp][lane * kListE + e] = tk.acc[e];
    __syncthreads();
    if (warp == 0) {
      for (int w = 1; w < kPqPersistWarps; ++w) {
#pragma unroll
        for (int e = 0; e < kListE; ++e) {
          const int src = 32 * kListE - 1 - (lane * kListE + e);
          const u64 b = lists[w][src];
          tk.acc[e] = tk.acc[e] < b ? tk.acc[e] : b;
        }
        warp_bitonic_merge<kListE>(tk.acc, lane);
      }
      u64* out = out_keys + static_cast<size_t>(q) * k;
#pragma unroll
      for (int e = 0; e < kListE; ++e) {
        const int i = lane * kListE + e;
        if (i < k) out[i] = tk.acc[e];
      }
    }
  }
  if (threadIdx.x == 0 && scanned_rows) atomicAdd(scanned_rows, rows_seen);
}

// Refine (cuVS `refine` / FAISS IndexRefineFlat semantics): exact re-rank of the k' ADC candidates
// of each query against the original rows.  One warp per query: lanes split the dimensions,
// candidate j's exact score lands in lane j % 32, a 128-key warp sort orders them.
constexpr int kRefineAhead = 8;   // candidate rows in flight per warp (divides 32)
template <typename T>
__global__ void refine_kernel(const T* __restrict__ rows, int dim, const float* __restrict__ qf,
                              int dp, const long long* __restrict__ cand, int nq, int k_in, int k_out,
                              int metric, long long id_offset, float* __restrict__ out_d,
                              long long* __restrict__ out_i) {
  const int lane = threadIdx.x & 31;
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= nq) return;
  u64 key[kListE];
#pragma unroll
  for (int e = 0; e < kListE; ++e) key[e] = kKeyInf;
  const float* qv = qf + static_cast<size_t>(q) * dp;
  // candidate ids: lane l holds candidates l, l + 32, ... (k_in <= 128), broadcast by shuffle, so
  // the row reads do not wait for a dependent id load; kRefineAhead rows are in flight at a time
  // (a single query used to be a chain of k_in dependent HBM round trips: 92 us at k_in = 80)
  long long my_cand[kListE];
#pragma unroll
  for (int e = 0; e < kListE; ++e) {
    const int j = lane + 32 * e;
    my_cand[e] = j < k_in ? cand[static_cast<size_t>(q) * k_in + j] : -1ll;  // shard-local row, -1 = none
  }
  for (int j0 = 0; j0 < k_in; j0 += kRefineAhead) {
    long long row[kRefineAhead];
    float acc[kRefineAhead];
#pragma unroll
    for (int b = 0; b < kRefineAhead; ++b) {
      const int j = j0 + b;   // j0 is a multiple of kRefineAhead (which divides 32): same register for all b
      long long r = -1ll;
#pragma unroll
      for (int e = 0; e < kListE; ++e)
        if (e == (j0 >> 5)) r = __shfl_sync(0xffffffffu, my_cand[e], j & 31);
      row[b] = j < k_in ? r : -1ll;
      acc[b] = 0.f;
    }
    for (int t = lane; t < dim; t += 32) {
      const float qt = qv[t];
      float xv[kRefineAhead];
#pragma unroll
      for (int b = 0; b < kRefineAhead; ++b)
        xv[b] = row[b] >= 0 ? ld_f32<T>(rows + static_cast<size_t>(row[b]) * dim + t) : 0.f;
#pragma unroll
      for (int b = 0; b < kRefineAhead; ++b) {
        if (row[b] < 0) continue;
        if (metric == B2VS_METRIC_L2) { const float df = qt - xv[b]; acc[b] = fmaf(df, df, acc[b]); }
        else acc[b] = fmaf(-qt, xv[b], acc[b]);
      }
    }
#pragma unroll
    for (int b = 0; b < kRefineAhead; ++b) {
      const int j = j0 + b;
      float a = acc[b];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      // blocked layout of the warp sort: element j lives in lane j / kListE, register j % kListE
      if (row[b] >= 0 && lane == j / kListE) {
#pragma unroll
        for (int e = 0; e < kListE; ++e)
          if (e == j % kListE) key[e] = pack_key(a, static_cast<uint32_t>(row[b]));
      }
    }
  }
  warp_bitonic_sort<kListE>(key, lane);
#pragma unroll
  for (int e = 0; e < kListE; ++e) {
    const int i = lane * kListE + e;
    if (i >= k_out) continue;
    const size_t o = static_cast<size_t>(q) * k_out + i;
    if (key[e] == kKeyInf) {
      out_d[o] = metric == B2VS_METRIC_L2 ? INFINITY : -INFINITY;
      out_i[o] = -1;
    } else {
      const float sc = key_score(key[e]);
      out_d[o] = metric == B2VS_METRIC_L2 ? sc : -sc;
      out_i[o] = static_cast<long long>(key_id(key[e])) + id_offset;
    }
  }
}

// queries -> fp32 [nq, dp] (+ ||q||^2)
template <typename T>
__global__ void queries_to_f32_kernel(const T* __restrict__ q, int nq, int dim, int dp, int fmt,
                                      int round16, float* __restrict__ qf, float* __restrict__ qnorm) {
  const int lane = threadIdx.x & 31;
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= nq) return;
  float acc = 0.f;
  for (int j = lane; j < dp; j += 32) {
    float v = j < dim ? ld_f32<T>(q + static_cast<size_t>(r) * dim + j) : 0.f;
    if (round16) { float back; to_op16(v, fmt, &back); v = back; }
    qf[static_cast<size_t>(r) * dp + j] = v;
    acc = fmaf(v, v, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) qnorm[r] = acc;
}

// ------------------------------------------------------------------------------------------
// Sort label of item i = size rank of the list it probes (rank 0 = longest list), so that work
// derived from the sorted order starts with the longest lists.
__global__ void probe_labels_kernel(const long long* __restrict__ probe_ids, int items, int stride,
                                    const int* __restrict__ rank_of_list, int* __restrict__ labels) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < items; i += gridDim.x * blockDim.x) {
    const long long l = probe_ids[static_cast<size_t>(i) * stride];
    labels[i] = rank_of_list[l < 0 ? 0 : l];  // empty items still run
  }
}

// Lists in descending-size order (host; sizes are known after the build / load).
static int build_list_ranks(IvfData* d) {
  std::vector<int> order(d->n_lists), rank(d->n_lists);
  for (int i = 0; i < d->n_lists; ++i) order[i] = i;
  if (std::getenv("B2VS_IVF_NO_RANK") == nullptr)  // A/B knob: keep list-id order
    std::stable_sort(order.begin(), order.end(),
                     [&](int a, int b) { return d->h_sizes[a] > d->h_sizes[b]; });
  for (int r = 0; r < d->n_lists; ++r) rank[order[r]] = r;
  d->max_list_rows = 0;
  for (int v : d->h_sizes) d->max_list_rows = std::max(d->max_list_rows, static_cast<int>(round_up(v, 32)));
  const size_t bytes = static_cast<size_t>(d->n_lists) * sizeof(int);
  B2VS_TRY(d->rank_of_list.reserve(bytes));
  B2VS_TRY(d->list_of_rank.reserve(bytes));
  B2VS_CUDA(cudaMemcpy(d->rank_of_list.ptr, rank.data(), bytes, cudaMemcpyHostToDevice));
  B2VS_CUDA(cudaMemcpy(d->list_of_rank.ptr, order.data(), bytes, cudaMemcpyHostToDevice));
  return B2VS_OK;
}

// ---- K7b grouped IVF-PQ scan: pieces around pq_tc_kernel (pq_tc.cuh) ------------------------
// bf16 copy of the codebooks + squared norm of every rounded entry
__global__ void pq_cb16_kernel(const float* __restrict__ codebooks, int entries, int dsub,
                               uint16_t* __restrict__ cb16, float* __restrict__ cbn) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= entries) return;
  float n2 = 0.f;
  for (int d = 0; d < dsub; ++d) {
    float back;
    cb16[static_cast<size_t>(e) * dsub + d] = to_op16(codebooks[static_cast<size_t>(e) * dsub + d], 1, &back);
    n2 = fmaf(back, back, n2);
  }
  cbn[e] = n2;
}

// ||decoded residual||^2 of every slot (thread = slot); +inf on padding slots and in the slack
__global__ void pq_slot_norms_kernel(const uint4* __restrict__ codes4, const uint32_t* __restrict__ row_ids,
                                     uint32_t n_slots, size_t n_out, int n_chunks,
                                     const float* __restrict__ cbn, int l2, float* __restrict__ out,
                                     unsigned int* __restrict__ max_bits) {
  const size_t slot = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  float n2 = 0.f;
  const bool real = slot < n_slots && row_ids[slot] != kNoRow;
  if (real) {
    const size_t g = slot >> 5;
    const int r = static_cast<int>(slot & 31);
    for (int ch = 0; ch < n_chunks; ++ch) {
      const uint4 v = __ldg(codes4 + (g * n_chunks + ch) * 32 + r);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 16; ++i)
        n2 += __ldg(cbn + (ch * 16 + i) * 256 + ((w[i >> 2] >> (8 * (i & 3))) & 0xFFu));
    }
  }
  if (slot < n_out) out[slot] = real ? (l2 ? n2 : 0.f) : INFINITY;
  const unsigned int wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(n2));
  if ((threadIdx.x & 31) == 0 && wmax != 0u) atomicMax(max_bits, wmax);
}

// LUT scan with the grouped path's operands (bf16-rounded residual query and codebooks), used
// for its two scalar-side jobs.  mode 0 = SEED: query q_perm[blockIdx.x], nearest list only,
// first row_limit rows -> tau[q] (k-th best + rounding cushion).  mode 1 = RESCUE: queries whose
// candidate buffer overflowed are rescanned over all their probes -> out_keys[q][k].
//   L2 score = ||rq||^2 + sum_m (||cb||^2 - 2 rq_m.cb)      IP score = -q.c - sum_m q_m.cb
__global__ void __launch_bounds__(kScanThreads)
ivf_pq_lut_scan_kernel(int mode, const uint8_t* __restrict__ codes, const uint32_t* __restrict__ row_ids,
                       const uint32_t* __restrict__ offsets, const long long* __restrict__ probe_ids,
                       const float* __restrict__ qf, const float* __restrict__ cent,
                       const uint16_t* __restrict__ cb16, const float* __restrict__ cbn, int dim,
                       int dp, int pq_dim, int dsub, int n_probes, int k, int metric,
                       uint32_t row_limit, float max_rhat2, const uint32_t* __restrict__ q_perm,
                       const int* __restrict__ count, int cap, float* __restrict__ tau,
                       u64* __restrict__ out_keys) {
  extern __shared__ float smem_f[];
  float* lut = smem_f;                 // [pq_dim * 256]
  float* rq = smem_f + pq_dim * 256;   // [dim]
  __shared__ u64 lists[kScanWarps][32 * kListE];
  __shared__ u64 top[kMaxFusedK];
  __shared__ float red[kScanWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = (mode == 0 && q_perm) ? static_cast<int>(q_perm[blockIdx.x]) : blockIdx.x;
  if (mode == 1 && count[q] <= cap) return;
  const int l2 = metric == B2VS_METRIC_L2;
  WarpTopK tk;
  tk.init();
  float bias_first = 0.f;
  const int np = mode == 0 ? 1 : n_probes;
  const int n_chunks = pq_dim >> 4;
  const uint4* codes4 = reinterpret_cast<const uint4*>(codes);
  for (int p = 0; p < np; ++p) {
    const long long list = probe_ids[static_cast<size_t>(q) * n_probes + p];
    if (list < 0) continue;
    __syncthreads();   // previous probe's LUT no longer in use
    const float* c = cent + static_cast<size_t>(list) * dim;
    float part_b = 0.f, part_n = 0.f;
    for (int d = threadIdx.x; d < dim; d += blockDim.x) {
      const float qv = qf[static_cast<size_t>(q) * dp + d];
      float back;
      to_op16(l2 ? qv - c[d] : qv, 1, &back);
      rq[d] = back;
      part_n = fmaf(back, back, part_n);
      if (!l2) part_b = fmaf(qv, c[d], part_b);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      part_b += __shfl_xor_sync(0xffffffffu, part_b, o);
      part_n += __shfl_xor_sync(0xffffffffu, part_n, o);
    }
    if (lane == 0) red[warp] = l2 ? part_n : part_b;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < kScanWarps; ++w) tot += red[w];
    const float bias = l2 ? tot : -tot;
    if (p == 0) bias_first = bias;
    // one entry per thread per step; unrolled so several L2 loads are in flight per thread
    if (dsub == 2) {
      const uint32_t* cbw = reinterpret_cast<const uint32_t*>(cb16);
#pragma unroll 8
      for (int idx = threadIdx.x; idx < pq_dim * 256; idx += kScanThreads) {
        const int m = idx >> 8;
        const uint32_t w = __ldg(cbw + idx);
        const float dot = fmaf(rq[2 * m + 1], __uint_as_float(w & 0xFFFF0000u),
                               rq[2 * m] * __uint_as_float(w << 16));
        lut[idx] = l2 ? fmaf(-2.f, dot, __ldg(cbn + idx)) : -dot;
      }
    } else {
#pragma unroll 4
      for (int idx = threadIdx.x; idx < pq_dim * 256; idx += kScanThreads) {
        const int m = idx >> 8;
        float dot = 0.f;
        for (int d = 0; d < dsub; ++d)
          dot = fmaf(rq[m * dsub + d], __uint_as_float(static_cast<uint32_t>(cb16[static_cast<size_t>(idx) * dsub + d]) << 16), dot);
        lut[idx] = l2 ? fmaf(-2.f, dot, __ldg(cbn + idx)) : -dot;
      }
    }
    __syncthreads();
    const uint32_t begin = offsets[list];
    const uint32_t end = mode == 0 ? min(offsets[list + 1], begin + row_limit) : offsets[list + 1];
    for (uint32_t g0 = (begin >> 5) + warp; g0 < (end >> 5); g0 += kScanWarps) {
      const uint32_t slot = (g0 << 5) + lane;
      float sc = bias;
      for (int ch = 0; ch < n_chunks; ++ch) {
        const uint4 v = __ldg(codes4 + (static_cast<size_t>(g0) * n_chunks + ch) * 32 + lane);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 16; ++i)
          sc += lut[(ch * 16 + i) * 256 + ((w[i >> 2] >> (8 * (i & 3))) & 0xFFu)];
      }
      u64 ck = kKeyInf;
      if (row_ids[slot] != kNoRow && sc < tk.tau) ck = pack_key(sc, slot);
      tk.offer(ck, k, lane);
    }
  }
  if (mode == 1) {
    block_merge_and_store(tk, lists, k, warp, lane, out_keys + static_cast<size_t>(q) * k);
    return;
  }
  block_merge_and_store(tk, lists, k, warp, lane, top);
  __syncthreads();
  if (threadIdx.x == 0) {
    // cushion for the different summation order of the tensor-core kernel (same products)
    const u64 kth = top[k - 1];
    float t = INFINITY;
    if (kth != kKeyInf) {
      const float sc = key_score(kth);
      float rqn = 0.f;
      for (int d = 0; d < dim; ++d) rqn = fmaf(rq[d], rq[d], rqn);
      const float eps = static_cast<float>(dim) * 1.2e-7f + 1e-6f;
      t = sc + (l2 ? 2.f : 1.f) * eps * sqrtf(rqn * max_rhat2) +
          4e-7f * (fabsf(sc) + fabsf(bias_first) + rqn + max_rhat2);
    }
    tau[q] = t;
  }
}

// One warp per gathered row: residual query (bf16) of item row_item[v] = (query, probe), its
// additive constant (||rq||^2 or -q.c) and its query id.
__global__ void gather_group_residuals_kernel(const uint32_t* __restrict__ row_item,
                                              const uint32_t* __restrict__ group_off, int n_lists,
                                              const long long* __restrict__ probe_ids,
                                              const float* __restrict__ qf, const float* __restrict__ cent,
                                              int dim, int dp, int n_probes, int l2,
                                              uint16_t* __restrict__ out, int* __restrict__ row_query,
                                              float* __restrict__ row_bias) {
  const int lane = threadIdx.x & 31;
  const int64_t v = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (v >= static_cast<int64_t>(group_off[n_lists])) return;
  const uint32_t item = row_item[v];
  uint16_t* orow = out + static_cast<size_t>(v) * dim;
  if (item == kNoRow) {
    // group padding: the row never qualifies (threshold -inf); its operand bytes are don't-care
    if (!kSkipPaddingRows)
      for (int j = lane; j < dim; j += 32) orow[j] = 0;
    if (lane == 0) { row_query[v] = -1; row_bias[v] = 0.f; }
    return;
  }
  const int q = static_cast<int>(item / static_cast<uint32_t>(n_probes));
  const long long list = probe_ids[item];
  const float* c = cent + static_cast<size_t>(list < 0 ? 0 : list) * dim;
  float acc = 0.f;
  for (int j = lane; j < dim; j += 32) {
    const float qv = qf[static_cast<size_t>(q) * dp + j];
    float back;
    orow[j] = to_op16(l2 ? qv - c[j] : qv, 1, &back);
    acc = l2 ? fmaf(back, back, acc) : fmaf(qv, c[j], acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) { row_query[v] = q; row_bias[v] = l2 ? acc : -acc; }
}

// Derives the grouped scan's operands from codebooks + codes (after a build or a load).
static int pq_prepare_grouped(b2vs_index* index, IvfData* d, cudaStream_t st) {
  d->pq_tc_ready = false;
  if (!pq_grouped_supported(index->dim, d->dsub) || d->mp != d->pq_dim) return B2VS_OK;
  const int entries = d->pq_dim * 256;
  const size_t n_out = static_cast<size_t>(std::max<int64_t>(d->n_slots, 1)) + kNormSlack;
  B2VS_TRY(d->cb16.reserve(static_cast<size_t>(entries) * d->dsub * 2));
  B2VS_TRY(d->cbn.reserve(static_cast<size_t>(entries) * sizeof(float)));
  B2VS_TRY(d->pq_norm.reserve(n_out * sizeof(float)));
  DevBuf cell;
  B2VS_TRY(cell.reserve(sizeof(unsigned int)));
  B2VS_CUDA(cudaMemsetAsync(cell.ptr, 0, sizeof(unsigned int), st));
  pq_cb16_kernel<<<static_cast<unsigned>(ceil_div(entries, 256)), 256, 0, st>>>(
      d->codebooks.as<float>(), entries, d->dsub, d->cb16.as<uint16_t>(), d->cbn.as<float>());
  pq_slot_norms_kernel<<<static_cast<unsigned>(ceil_div(n_out, 256)), 256, 0, st>>>(
      d->codes.as<uint4>(), d->row_ids.as<uint32_t>(), static_cast<uint32_t>(d->n_slots), n_out,
      d->mp >> 4, d->cbn.as<float>(), index->metric == B2VS_METRIC_L2 ? 1 : 0,
      d->pq_norm.as<float>(), cell.as<unsigned int>());
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(&d->max_rhat2, cell.ptr, sizeof(float), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cell.release();
  B2VS_CHECK(e == cudaSuccess, B2VS_ECUDA, "preparing the grouped PQ scan failed: %s", cudaGetErrorString(e));
  d->pq_tc_ready = true;
  return B2VS_OK;
}

// Instantiation table of the <FMT, J> scan kernels: J = 16-byte chunks of a row owned by a lane.
#define FLAT_SCAN_DISPATCH(KERNEL, fmt, j, grid, st, ...)                                    \
  do {                                                                                       \
    if ((fmt) == 0) {                                                                        \
      if ((j) <= 1) KERNEL<0, 1><<<(grid), kScanThreads, 0, (st)>>>(__VA_ARGS__);            \
      else if ((j) == 2) KERNEL<0, 2><<<(grid), kScanThreads, 0, (st)>>>(__VA_ARGS__);       \
      else if ((j) == 3) KERNEL<0, 3><<<(grid), kScanThreads, 0, (st)>>>(__VA_ARGS__);       \
      else if ((j) == 4) KERNEL<0, 4><<<(grid), kScanThreads, 0, (st)>>>(__VA_ARGS__);       \
      else if ((j) <= 6) KERNEL<0, 6><<<(grid), kScanThreads, 0, (st)>>>(__VA_ARGS__);       \
      else KERNEL<0, 8><<<(grid), kScanThreads, 0, (st)>>>(__VA_ARGS__);                     \
    } else {                                                                                 \
      if ((j) <= 1) KERNEL<1, 1><<<(grid), kScanThreads, 0, (st)>>>(__VA_ARGS__);            \
      else if ((j) == 2) KERNEL<1, 2><<<(grid), kScanThreads, 0, (st)>>>(__VA_ARGS__);       \
      else if ((j) == 3) KERNEL<1, 3><<<(grid), kScanThreads, 0, (st)>>>(__VA_ARGS__);       \
      else if ((j) == 4) KERNEL<1, 4><<<(grid), kScanThreads, 0, (st)>>>(__VA_ARGS__);       \
      else if ((j) <= 6) KERNEL<1, 6><<<(grid), kScanThreads, 0, (st)>>>(__VA_ARGS__);       \
      else KERNEL<1, 8><<<(grid), kScanThreads, 0, (st)>>>(__VA_ARGS__);                     \
    }                                                                                        \
  } while (0)

// Counting sort of items by the size rank of their list (the same three kernels that build the
// lists).  Item i probes list probe_ids[i * stride].  group_pad = 1: dense permutation (row_item[i] = i-th item in list order);
// group_pad = 128: every list's group starts on a 128-row boundary, holes hold kNoRow.
static size_t sorted_rows_cap(const IvfData* d, int items, int group_pad) {
  return group_pad == 1 ? static_cast<size_t>(items)
                        : static_cast<size_t>(items) +
                              static_cast<size_t>(group_pad) * std::min(d->n_lists, items);
}

static int reserve_item_sort(IvfData* d, int items, int group_pad) {
  B2VS_TRY(d->ws_item_lab.reserve(static_cast<size_t>(items) * sizeof(int)));
  B2VS_TRY(d->ws_item_cnt.reserve(static_cast<size_t>(d->n_lists) * 2 * sizeof(int)));
  B2VS_TRY(d->ws_item_off.reserve((static_cast<size_t>(d->n_lists) + 1) * sizeof(uint32_t)));
  B2VS_TRY(d->ws_item_perm.reserve(sorted_rows_cap(d, items, group_pad) * sizeof(uint32_t)));
  B2VS_TRY(d->ws_item_slot.reserve(static_cast<size_t>(items) * sizeof(uint32_t)));
  return B2VS_OK;
}

static int sort_items_by_list(IvfData* d, const long long* probe_ids, int items, int stride,
                              int group_pad, cudaStream_t st) {
  const size_t rows_cap = sorted_rows_cap(d, items, group_pad);
  B2VS_TRY(reserve_item_sort(d, items, group_pad));
  const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(ceil_div(items, 256), 2048));
  int* cnt = d->ws_item_cnt.as<int>();
  B2VS_CUDA(cudaMemsetAsync(cnt, 0, static_cast<size_t>(d->n_lists) * 2 * sizeof(int), st));
  if (group_pad > 1) B2VS_CUDA(cudaMemsetAsync(d->ws_item_perm.ptr, 0xFF, rows_cap * sizeof(uint32_t), st));
  probe_labels_kernel<<<blocks, 256, 0, st>>>(probe_ids, items, stride, d->rank_of_list.as<int>(),
                                              d->ws_item_lab.as<int>());
  histogram_kernel<<<blocks, 256, 0, st>>>(d->ws_item_lab.as<int>(), items, cnt);
  scan_sizes_kernel<<<1, 1024, 0, st>>>(cnt, d->n_lists, group_pad, d->ws_item_off.as<uint32_t>());
  scatter_rows_kernel<<<blocks, 256, 0, st>>>(d->ws_item_lab.as<int>(), items,
                                              d->ws_item_off.as<uint32_t>(), cnt + d->n_lists,
                                              d->ws_item_perm.as<uint32_t>(),
                                              d->ws_item_slot.as<uint32_t>());
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}

// Candidate-buffer capacity (a power of two) and seed-sample length of the grouped scan, by k.
// B2VS_IVF_GROUPED_CAP shrinks the buffers so tests can drive the overflow-rescue path.
static int grouped_cap(int k) {
  const char* e = std::getenv("B2VS_IVF_GROUPED_CAP");
  if (e) {
    const int v = std::atoi(e);
    if (v >= 32 && v <= 4096 && (v & (v - 1)) == 0) return v;
  }
  return k <= 32 ? 2048 : 4096;
}
static uint32_t grouped_seed_rows(int k) {
  const char* e = std::getenv("B2VS_IVF_SEED_ROWS");  // A/B knob
  if (e && std::atoi(e) >= 32) return static_cast<uint32_t>(std::atoi(e));
  return static_cast<uint32_t>(std::max(256, 16 * k));
}

// B2VS_IVF_GROUPED=0|1 forces the per-item / grouped IVF-Flat scan (A/B measurements, tests).
static int grouped_override() {
  const char* e = std::getenv("B2VS_IVF_GROUPED");
  return (e && (e[0] == '0' || e[0] == '1')) ? (e[0] - '0') : -1;
}

static int ivf_search_batch(b2vs_index* index, const void* q, int q_dtype, int nq, int k,
                            const b2vs_search_params& sp, float* out_d, int64_t* out_i,
                            cudaStream_t st);
static int ivf_pq_search_bigk(b2vs_index* index, const void* q, int q_dtype, int nq, int k,
                              const b2vs_search_params& sp, float* out_d, int64_t* out_i,
                              cudaStream_t st);

// Small batches are launch-bound (a Q = 1 search is ~20 tiny kernels), so when the item count is
// small one CTA does the whole planning step: rank labels, shared-memory histogram, scan of the
// 128-padded group sizes, scatter of the items into group order and the work table - the job of
// probe_labels / histogram / scan_sizes / scatter_rows / build_group_work (+ two memsets).
constexpr int kPlanThreads = 1024;
constexpr int kPlanMaxItems = 16384;
constexpr int kPlanMaxLists = 16384;
__global__ void __launch_bounds__(kPlanThreads)
ivf_plan_small_kernel(const long long* __restrict__ probe_ids, int items,
                      const int* __restrict__ rank_of_list, const int* __restrict__ list_of_rank,
                      const uint32_t* __restrict__ offsets, int n_lists, int chunk_rows, int slots,
                      uint32_t* __restrict__ row_item, uint32_t* __restrict__ group_off,
                      int4* __restrict__ work, int* __restrict__ n_work,
                      unsigned long long* __restrict__ scanned_rows) {
  extern __shared__ int plan_sm[];
  int* cnt = plan_sm;                                             // [n_lists] by size rank
  uint32_t* off = reinterpret_cast<uint32_t*>(plan_sm + n_lists); // [n_lists + 1]
  __shared__ u64 part[kPlanThreads / 32];
  static_assert(kPlanThreads == 1024, "the block scan assumes 32 full warps");
  const int t = threadIdx.x;
  for (int i = t; i < n_lists; i += kPlanThreads) cnt[i] = 0;
  __syncthreads();
  for (int i = t; i < items; i += kPlanThreads) {
    const long long l = probe_ids[i];
    atomicAdd(&cnt[rank_of_list[l < 0 ? 0 : l]], 1);
  }
  __syncthreads();
  const int per = (n_lists + kPlanThreads - 1) / kPlanThreads;
  const int lo = min(n_lists, t * per), hi = min(n_lists, lo + per);
  // per thread: padded group rows (high half) and work items (low half) of its lists; a list
  // probed by c queries gives ceil(c / 128) query blocks x ceil(rows / chunk_rows) row ranges.
  // Only non-empty ranges become work items, packed densely in rank order (longest list first):
  // the scan kernel strides over them statically, so holes would unbalance its CTAs.
  u64 sum = 0;
  for (int i = lo; i < hi; ++i) {
    const int c = cnt[i];
    if (c == 0) continue;
    const int l = list_of_rank[i];
    const int rows = static_cast<int>(offsets[l + 1] - offsets[l]);
    const u64 blocks = static_cast<u64>((c + kGroupRows - 1) / kGroupRows);
    sum += ((blocks * kGroupRows) << 32) | (blocks * static_cast<u64>((rows + chunk_rows - 1) / chunk_rows));
  }
  // exclusive scan of the per-thread sums: shuffle scan inside each warp, then across the 32
  // warp totals (a one-thread loop over the 1024 partials cost ~10 us of a 160 us Q = 1 search)
  const int lane = t & 31, warp = t >> 5;
  u64 inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u64 v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) part[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const u64 w = part[lane];
    u64 winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u64 v = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += v;
    }
    part[lane] = winc - w;   // totals before this warp
    if (lane == 31) {
      off[n_lists] = static_cast<uint32_t>(winc >> 32);
      group_off[n_lists] = static_cast<uint32_t>(winc >> 32);
      *n_work = static_cast<int>(winc & 0xffffffffu);
    }
  }
  __syncthreads();
  const u64 before = part[warp] + inc - sum;
  uint32_t run = static_cast<uint32_t>(before >> 32);
  uint32_t wrun = static_cast<uint32_t>(before & 0xffffffffu);
  unsigned long long rows_scanned = 0;
  for (int i = lo; i < hi; ++i) {
    off[i] = run;
    group_off[i] = run;
    const int c = cnt[i];
    if (c == 0) continue;
    const int l = list_of_rank[i];
    const int begin = static_cast<int>(offsets[l]), end = static_cast<int>(offsets[l + 1]);
    const int blocks = (c + kGroupRows - 1) / kGroupRows;
    const int b0 = static_cast<int>(run >> 7);
    for (int b = 0; b < blocks; ++b)
      for (int rb = begin; rb < end; rb += chunk_rows)
        work[wrun++] = make_int4(b0 + b, rb, min(end, rb + chunk_rows), 0);
    rows_scanned += static_cast<unsigned long long>(c) * static_cast<unsigned>(end - begin);
    run += static_cast<uint32_t>(blocks * kGroupRows);
  }
  if (scanned_rows && rows_scanned) atomicAdd(scanned_rows, rows_scanned);
  __syncthreads();
  for (int i = t; i < n_lists; i += kPlanThreads) cnt[i] = 0;   // now the scatter cursors
  __syncthreads();
  for (int i = t; i < items; i += kPlanThreads) {
    const long long l = probe_ids[i];
    const int r = rank_of_list[l < 0 ? 0 : l];
    row_item[off[r] + static_cast<uint32_t>(atomicAdd(&cnt[r], 1))] = static_cast<uint32_t>(i);
  }
}

// Sort + work table of the grouped scans: the one-CTA plan for small batches, else the
// counting-sort kernels + build_group_work_kernel.
static int plan_grouped_work(IvfData* d, const long long* probe_ids, int items, int chunk_rows,
                             int slots, int4* work, int* n_work, unsigned long long* counter,
                             cudaStream_t st) {
  if (items <= kPlanMaxItems && d->n_lists <= kPlanMaxLists) {
    B2VS_TRY(reserve_item_sort(d, items, kGroupRows));
    B2VS_CUDA(cudaMemsetAsync(d->ws_item_perm.ptr, 0xFF,
                              sorted_rows_cap(d, items, kGroupRows) * sizeof(uint32_t), st));
    const size_t smem = (2 * static_cast<size_t>(d->n_lists) + 1) * sizeof(int);
    B2VS_CUDA(cudaFuncSetAttribute(ivf_plan_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem)));
    ivf_plan_small_kernel<<<1, kPlanThreads, smem, st>>>(
        probe_ids, items, d->rank_of_list.as<int>(), d->list_of_rank.as<int>(),
        d->offsets.as<uint32_t>(), d->n_lists, chunk_rows, slots, d->ws_item_perm.as<uint32_t>(),
        d->ws_item_off.as<uint32_t>(), work, n_work, counter);
    B2VS_CUDA(cudaGetLastError());
    return B2VS_OK;
  }
  B2VS_TRY(sort_items_by_list(d, probe_ids, items, 1, kGroupRows, st));
  build_group_work_kernel<<<static_cast<unsigned>(ceil_div(d->n_lists, 256)), 256, 0, st>>>(
      d->ws_item_off.as<uint32_t>(), d->offsets.as<uint32_t>(), d->ws_item_cnt.as<int>(),
      d->list_of_rank.as<int>(), d->n_lists, chunk_rows, slots, work, n_work, counter);
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}

// Row-range split of the grouped scan's work items (see build_group_work_kernel): large batches
// aim at two items per SM when the batch alone does not provide them.
static void choose_work_split(const b2vs_index* index, const IvfData* d, int items, int* chunk_rows,
                              int* slots) {
  const int sms = sm_count(index->dev);
  const double mean_rows = std::max(1.0, static_cast<double>(d->n) / std::max(d->n_lists, 1));
  const double est_tiles = static_cast<double>(items) * (mean_rows / 256.0 + 0.5);
  const int max_tiles = std::max(1, static_cast<int>(ceil_div(std::max(d->max_list_rows, 1), 256)));
  int chunk_tiles = static_cast<int>(std::ceil(est_tiles / (2.0 * sms)));
  chunk_tiles = std::max(1, std::min(chunk_tiles, max_tiles));
  // Small batches (the one-CTA planner, which packs the non-empty row ranges densely): measured
  // on B200 (profiles/r1_work_split_sweep_*.jsonl, tools/sweep_work_split.py) one-tile items win
  // up to Q = 8 (IVF-Flat -7..-12 %), four-tile items from Q = 16 to 64 (-4..-19 %), and the
  // per-item cost (query block re-staged per item) only shows from Q = 128 on: aim at ~16 items
  // per SM with at most four tiles each.
  if (items <= kPlanMaxItems && d->n_lists <= kPlanMaxLists) {
    chunk_tiles = static_cast<int>(est_tiles / (16.0 * sms));
    chunk_tiles = std::max(1, std::min(chunk_tiles, std::min(4, max_tiles)));
  }
  // A/B switches (read per call): B2VS_WORK_CHUNK_TILES=n forces the chunk, B2VS_DEBUG_SPLIT prints it
  if (const char* e = std::getenv("B2VS_WORK_CHUNK_TILES")) {
    const int v = std::atoi(e);
    if (v > 0) chunk_tiles = std::min(v, max_tiles);
  }
  *slots = static_cast<int>(ceil_div(max_tiles, chunk_tiles));
  *chunk_rows = chunk_tiles * 256;
  if (std::getenv("B2VS_DEBUG_SPLIT"))
    std::fprintf(stderr, "[b2vs] work split: items=%d mean_rows=%.0f max_tiles=%d chunk_tiles=%d slots=%d\n",
                 items, mean_rows, max_tiles, chunk_tiles, *slots);
}

// Groups the nq * n_probes (query, probe) items by list and runs the tensor-core list scan in
// append mode against the thresholds in ws_g_tau; candidates land in ws_g_cand / ws_g_cnt (the
// caller sizes, zeroes and later selects from them).  probe_ids is [nq, n_probes] dense.
static int run_grouped_flat_scan(b2vs_index* index, IvfData* d, const long long* probe_ids,
                                 int n_probes, int nq, int cap, unsigned long long* counter,
                                 cudaStream_t st) {
  const int items = nq * n_probes;
  const int q_split = index->dtype == B2VS_F32 ? 1 : 0;
  const int q_pitch = q_split ? 2 * static_cast<int>(round_up(d->dp, 64)) : d->dp;
  int chunk_rows = 0, slots = 1;
  choose_work_split(index, d, items, &chunk_rows, &slots);
  const int max_work = (items / kGroupRows + std::min(d->n_lists, items) + 1) * slots;
  const int64_t rows_cap = static_cast<int64_t>(sorted_rows_cap(d, items, kGroupRows));
  B2VS_TRY(d->ws_g_work.reserve(static_cast<size_t>(max_work) * sizeof(int4) + 16));
  B2VS_TRY(d->ws_g_q.reserve(static_cast<size_t>(rows_cap) * q_pitch * 2));
  B2VS_TRY(d->ws_g_rowq.reserve(static_cast<size_t>(rows_cap) * sizeof(int)));
  int* n_work = reinterpret_cast<int*>(d->ws_g_work.as<char>() + static_cast<size_t>(max_work) * sizeof(int4));
  B2VS_TRY(plan_grouped_work(d, probe_ids, items, chunk_rows, slots, d->ws_g_work.as<int4>(), n_work,
                             counter, st));
  gather_group_queries_kernel<<<static_cast<unsigned>(ceil_div(rows_cap, 8)), 256, 0, st>>>(
      d->ws_item_perm.as<uint32_t>(), d->ws_item_off.as<uint32_t>(), d->n_lists,
      d->ws_qf.as<float>(), d->dp, n_probes, d->fmt, q_split, d->ws_g_q.as<uint16_t>(),
      d->ws_g_rowq.as<int>());
  B2VS_CUDA(cudaGetLastError());
  GroupedScanArgs ga{};
  ga.q_mat = d->ws_g_q.ptr; ga.q_rows = rows_cap;
  ga.x_mat = d->data.ptr; ga.x_rows = d->n_slots;
  ga.kdim = d->dp; ga.ab_format = d->fmt; ga.q_split = q_split;
  ga.beta = d->slot_norm.as<float>();
  ga.alpha = index->metric == B2VS_METRIC_L2 ? -2.f : -1.f;
  ga.work = d->ws_g_work.ptr; ga.n_work = n_work; ga.max_work = max_work;
  ga.row_query = d->ws_g_rowq.as<int>(); ga.tau = d->ws_g_tau.as<float>();
  ga.cand = d->ws_g_cand.as<u64>(); ga.count = d->ws_g_cnt.as<int>(); ga.cap = cap;
  return launch_grouped_scan(index->dev, ga, st);
}

// IVF-Flat with 128 < k <= 2048 (the reference's top-2000 retrieval mode on an IVF index,
// improved_multi_gpu_rag.py:37-48 + :247).  Two grouped scans into 64 K-key per-query buffers:
//   1. the m nearest lists of every query with no threshold (m lists hold ~3k rows) -> radix
//      select of the k-th key = a valid threshold computed with the scan's own arithmetic;
//   2. all n_probes lists below that threshold -> radix select + sort of the k best.
static int ivf_flat_search_bigk(b2vs_index* index, const void* q, int q_dtype, int nq, int k,
                                const b2vs_search_params& sp, float* out_d, int64_t* out_i,
                                cudaStream_t st) {
  IvfData* d = static_cast<IvfData*>(index->ivf);
  constexpr int kCapBig = 65536;
  int n_probes = sp.n_probes > 0 ? sp.n_probes : 20;
  n_probes = std::min(n_probes, std::min(d->n_lists, kMaxProbes));
  int max_size = 1;
  for (int v : d->h_sizes) max_size = std::max(max_size, static_cast<int>(round_up(v, 32)));
  const double mean_size = std::max(1.0, static_cast<double>(d->n) / d->n_lists);
  int m = std::min(n_probes, static_cast<int>(std::ceil(3.0 * k / mean_size)) + 1);
  m = std::max(1, std::min(m, kCapBig / max_size));
  B2VS_CHECK(max_size <= kCapBig, B2VS_EUNSUP, "a list of %d rows exceeds the large-k buffer", max_size);
  const int q_pad = static_cast<int>(round_up(nq, 128));
  B2VS_TRY(d->ws_probe_d.reserve(static_cast<size_t>(nq) * n_probes * sizeof(float)));
  B2VS_TRY(d->ws_probe_i.reserve(static_cast<size_t>(nq) * n_probes * sizeof(int64_t)));
  B2VS_TRY(d->ws_ref_i.reserve(static_cast<size_t>(nq) * m * sizeof(int64_t)));   // first m probes
  B2VS_TRY(d->ws_qf.reserve(static_cast<size_t>(nq) * d->dp * sizeof(float)));
  B2VS_TRY(d->ws_qnorm.reserve(static_cast<size_t>(q_pad) * sizeof(float)));
  B2VS_TRY(d->ws_counter.reserve(2 * sizeof(unsigned long long) + sizeof(int)));
  B2VS_TRY(d->ws_g_tau.reserve(static_cast<size_t>(nq) * sizeof(float)));
  B2VS_TRY(d->ws_g_cand.reserve(static_cast<size_t>(nq) * kCapBig * sizeof(u64)));
  B2VS_TRY(d->ws_g_cnt.reserve(static_cast<size_t>(nq) * sizeof(int)));
  B2VS_TRY(reserve_item_sort(d, nq * n_probes, kGroupRows));
  B2VS_TRY(index->flat.search(q, q_dtype, nq, n_probes, 0, 0, d->ws_probe_d.as<float>(),
                              d->ws_probe_i.as<int64_t>(), nullptr, st));
  int launches = index->flat.stats.launches;
  const int round16 = index->dtype != B2VS_F32 ? 1 : 0;
  DISPATCH_DTYPE(q_dtype, T, (queries_to_f32_kernel<T><<<static_cast<unsigned>(ceil_div(nq, 4)), 128, 0, st>>>(
                                 static_cast<const T*>(q), nq, index->dim, d->dp, d->fmt, round16,
                                 d->ws_qf.as<float>(), d->ws_qnorm.as<float>())));
  B2VS_CUDA(cudaGetLastError());
  unsigned long long* counter = d->ws_counter.as<unsigned long long>();
  int* overflow = reinterpret_cast<int*>(counter + 2);
  B2VS_CUDA(cudaMemsetAsync(counter, 0, 2 * sizeof(unsigned long long) + sizeof(int), st));
  const long long* probe_ids = reinterpret_cast<const long long*>(d->ws_probe_i.ptr);
  // ---- pass 1: nearest m lists, no threshold
  B2VS_CUDA(cudaMemcpy2DAsync(d->ws_ref_i.ptr, static_cast<size_t>(m) * sizeof(int64_t), probe_ids,
                              static_cast<size_t>(n_probes) * sizeof(int64_t),
                              static_cast<size_t>(m) * sizeof(int64_t), nq, cudaMemcpyDeviceToDevice, st));
  fill_f32_kernel<<<static_cast<unsigned>(ceil_div(nq, 256)), 256, 0, st>>>(d->ws_g_tau.as<float>(), nq, INFINITY);
  B2VS_CUDA(cudaMemsetAsync(d->ws_g_cnt.ptr, 0, static_cast<size_t>(nq) * sizeof(int), st));
  B2VS_TRY(run_grouped_flat_scan(index, d, reinterpret_cast<const long long*>(d->ws_ref_i.ptr), m, nq,
                                 kCapBig, nullptr, st));
  B2VS_TRY(launch_bigk_select(d->ws_g_cand.as<u64>(), d->ws_g_cnt.as<int>(), kCapBig, nq, k, 0,
                              index->metric, nullptr, 0, d->ws_g_tau.as<float>(), nullptr, nullptr,
                              nullptr, st));
  // ---- pass 2: every probed list below the threshold
  B2VS_CUDA(cudaMemsetAsync(d->ws_g_cnt.ptr, 0, static_cast<size_t>(nq) * sizeof(int), st));
  B2VS_TRY(run_grouped_flat_scan(index, d, probe_ids, n_probes, nq, kCapBig, counter, st));
  B2VS_TRY(launch_bigk_select(d->ws_g_cand.as<u64>(), d->ws_g_cnt.as<int>(), kCapBig, nq, k, 1,
                              index->metric, d->ws_qnorm.as<float>(), index->id_offset, nullptr,
                              out_d, out_i, overflow, st, d->row_ids.as<uint32_t>()));
  int h_over = 0;
  B2VS_CUDA(cudaMemcpyAsync(&h_over, overflow, sizeof(int), cudaMemcpyDeviceToHost, st));
  B2VS_CUDA(cudaStreamSynchronize(st));
  B2VS_CHECK(h_over <= kCapBig, B2VS_EUNSUP,
             "large-k IVF search: %d candidates under the seed threshold exceed the %d-key buffer "
             "(fewer than k rows in the nearest lists); use more lists per query or a flat index",
             h_over, kCapBig);
  launches += 22;
  d->stats = b2vs_search_stats{};
  d->stats.launches = launches;
  d->stats.n_splits = n_probes;
  d->stats.grid = nq * n_probes;
  d->stats.algo_flops = 2.0 * nq * static_cast<double>(d->n_lists) * index->dim;
  d->counter_pending = true;
  d->last_nq = nq;
  d->timing_pending = false;
  return B2VS_OK;
}


// Workspaces scale with nq * n_probes (grouped query operand: up to ~2x that many rows of the
// index dimension), so very large batches run as consecutive sub-batches.
constexpr int64_t kMaxItemsPerBatch = 4 << 20;

// large k: k itself, or (IVF-PQ) the number of ADC candidates kept for the exact re-rank
static bool uses_bigk_path(const b2vs_index* index, const IvfData* d, int k, const b2vs_search_params& sp) {
  const bool pq = index->kind == B2VS_KIND_IVF_PQ;
  const bool pq_refine = pq && sp.refine_ratio > 1 && d->src_rows != nullptr;
  return k > kMaxFusedK ||
         (pq_refine && d->pq_tc_ready && static_cast<int64_t>(k) * sp.refine_ratio > kMaxFusedK);
}

static int ivf_search_direct(b2vs_index* index, const void* q, int q_dtype, int nq, int k,
                             const b2vs_search_params& sp, float* out_d, int64_t* out_i,
                             cudaStream_t st) {
  IvfData* d = static_cast<IvfData*>(index->ivf);
  B2VS_CHECK(d != nullptr, B2VS_EINVAL, "IVF index has no list data");
  int n_probes = sp.n_probes > 0 ? sp.n_probes : 20;
  n_probes = std::max(1, std::min(n_probes, std::min(d->n_lists, kMaxProbes)));
  int chunk = static_cast<int>(std::max<int64_t>(1024, kMaxItemsPerBatch / n_probes));
  const bool pq = index->kind == B2VS_KIND_IVF_PQ;
  const bool bigk = uses_bigk_path(index, d, k, sp);
  if (bigk) {
    B2VS_CHECK(k <= kMaxBigK, B2VS_EUNSUP, "k=%d exceeds the large-k limit %d", k, kMaxBigK);
    chunk = std::min(chunk, 4096);   // 64 K-key candidate buffer per query
  }
  auto run = [&](const void* qq, int n, float* od, int64_t* oi) {
    if (!bigk) return ivf_search_batch(index, qq, q_dtype, n, k, sp, od, oi, st);
    return pq ? ivf_pq_search_bigk(index, qq, q_dtype, n, k, sp, od, oi, st)
              : ivf_flat_search_bigk(index, qq, q_dtype, n, k, sp, od, oi, st);
  };
  if (nq <= chunk) return run(q, nq, out_d, out_i);
  const size_t q_pitch = static_cast<size_t>(index->dim) * elem_bytes(q_dtype);
  int launches = 0;
  for (int q0 = 0; q0 < nq; q0 += chunk) {
    const int nc = std::min(chunk, nq - q0);
    B2VS_TRY(run(static_cast<const char*>(q) + static_cast<size_t>(q0) * q_pitch, nc,
                 out_d + static_cast<size_t>(q0) * k, out_i + static_cast<size_t>(q0) * k));
    launches += d->stats.launches;
  }
  d->stats.launches = launches;   // the other fields describe the last sub-batch
  return B2VS_OK;
}

// PQ counterpart of run_grouped_flat_scan: plan, gather the residual queries, decode + scan on
// the tensor cores.  Thresholds / candidate buffers (ws_g_tau, ws_g_cand, ws_g_cnt) are the caller's.
static int run_grouped_pq_scan(b2vs_index* index, IvfData* d, const long long* probe_ids, int n_probes,
                               int nq, int cap, unsigned long long* counter, cudaStream_t st) {
  const int items = nq * n_probes;
  const int l2 = index->metric == B2VS_METRIC_L2 ? 1 : 0;
  int chunk_rows = 0, slots = 1;
  choose_work_split(index, d, items, &chunk_rows, &slots);
  const int max_work = (items / kGroupRows + std::min(d->n_lists, items) + 1) * slots;
  const int64_t rows_cap = static_cast<int64_t>(sorted_rows_cap(d, items, kGroupRows));
  B2VS_TRY(d->ws_g_work.reserve(static_cast<size_t>(max_work) * sizeof(int4) + 16));
  B2VS_TRY(d->ws_g_q.reserve(static_cast<size_t>(rows_cap) * index->dim * 2));
  B2VS_TRY(d->ws_g_rowq.reserve(static_cast<size_t>(rows_cap) * sizeof(int)));
  B2VS_TRY(d->ws_g_bias.reserve(static_cast<size_t>(rows_cap) * sizeof(float)));
  int* n_work = reinterpret_cast<int*>(d->ws_g_work.as<char>() + static_cast<size_t>(max_work) * sizeof(int4));
  B2VS_TRY(plan_grouped_work(d, probe_ids, items, chunk_rows, slots, d->ws_g_work.as<int4>(), n_work,
                             counter, st));
  gather_group_residuals_kernel<<<static_cast<unsigned>(ceil_div(rows_cap, 8)), 256, 0, st>>>(
      d->ws_item_perm.as<uint32_t>(), d->ws_item_off.as<uint32_t>(), d->n_lists, probe_ids,
      d->ws_qf.as<float>(), d->centroids.as<float>(), index->dim, d->dp, n_probes, l2,
      d->ws_g_q.as<uint16_t>(), d->ws_g_rowq.as<int>(), d->ws_g_bias.as<float>());
  B2VS_CUDA(cudaGetLastError());
  PqGroupedScanArgs ga{};
  ga.q_mat = d->ws_g_q.ptr; ga.q_rows = rows_cap;
  ga.dim = index->dim; ga.pq_dim = d->pq_dim; ga.dsub = d->dsub;
  ga.codes = d->codes.ptr;
  ga.n_groups = static_cast<uint32_t>(std::max<int64_t>(d->n_slots, 32) >> 5);
  ga.cb16 = d->cb16.ptr;
  ga.beta = d->pq_norm.as<float>(); ga.alpha = l2 ? -2.f : -1.f;
  ga.work = d->ws_g_work.ptr; ga.n_work = n_work; ga.max_work = max_work;
  ga.row_query = d->ws_g_rowq.as<int>(); ga.row_bias = d->ws_g_bias.as<float>();
  ga.tau = d->ws_g_tau.as<float>();
  ga.cand = d->ws_g_cand.as<u64>(); ga.count = d->ws_g_cnt.as<int>(); ga.cap = cap;
  return launch_pq_grouped_scan(index->dev, ga, st);
}

// Exact re-rank of up to 2048 candidates per query (one CTA per query): every warp scores
// candidates against the caller's original rows, the CTA sorts them in shared memory.
constexpr int kRefineBigThreads = 256;
template <typename T>
__global__ void __launch_bounds__(kRefineBigThreads)
refine_big_kernel(const T* __restrict__ rows, int dim, const float* __restrict__ qf, int dp,
                  const long long* __restrict__ cand, int k_in, int k_out, int metric,
                  long long id_offset, float* __restrict__ out_d, long long* __restrict__ out_i) {
  __shared__ u64 keys[2048];
  const int q = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int P = 32;
  while (P < k_in) P <<= 1;
  for (int i = threadIdx.x; i < P; i += blockDim.x) keys[i] = kKeyInf;
  __syncthreads();
  const float* qv = qf + static_cast<size_t>(q) * dp;
  for (int j = warp; j < k_in; j += kRefineBigThreads / 32) {
    const long long row = cand[static_cast<size_t>(q) * k_in + j];
    if (row < 0) continue;
    const T* x = rows + static_cast<size_t>(row) * dim;
    float acc = 0.f;
    for (int t = lane; t < dim; t += 32) {
      const float xv = ld_f32<T>(x + t);
      if (metric == B2VS_METRIC_L2) { const float df = qv[t] - xv; acc = fmaf(df, df, acc); }
      else acc = fmaf(-qv[t], xv, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) keys[j] = pack_key(acc, static_cast<uint32_t>(row));
  }
  __syncthreads();
  for (int size = 2; size <= P; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
        const bool up = (lo & size) == 0;
        const u64 a = keys[lo], b = keys[hi];
        if ((a > b) == up) { keys[lo] = b; keys[hi] = a; }
      }
      __syncthreads();
    }
  for (int i = threadIdx.x; i < k_out; i += blockDim.x) {
    const size_t o = static_cast<size_t>(q) * k_out + i;
    const u64 key = i < P ? keys[i] : kKeyInf;
    if (key == kKeyInf) {
      out_d[o] = metric == B2VS_METRIC_L2 ? INFINITY : -INFINITY;
      out_i[o] = -1;
    } else {
      const float sc = key_score(key);
      out_d[o] = metric == B2VS_METRIC_L2 ? sc : -sc;
      out_i[o] = static_cast<long long>(key_id(key)) + id_offset;
    }
  }
}

// IVF-PQ with k (or k * refine_ratio) above 128, up to 2048: the two-pass scheme of
// ivf_flat_search_bigk on the decoded-tile scan, then (optionally) the exact re-rank.
static int ivf_pq_search_bigk(b2vs_index* index, const void* q, int q_dtype, int nq, int k,
                              const b2vs_search_params& sp, float* out_d, int64_t* out_i,
                              cudaStream_t st) {
  IvfData* d = static_cast<IvfData*>(index->ivf);
  B2VS_CHECK(d->pq_tc_ready, B2VS_EUNSUP,
             "k > %d on this IVF-PQ shape needs the grouped scan (dsub 2/4/8, dim %% 64 == 0)", kMaxFusedK);
  constexpr int kCapBig = 65536;
  const bool refine = sp.refine_ratio > 1 && d->src_rows != nullptr;
  const int k_scan = refine ? std::min(kMaxBigK, k * sp.refine_ratio) : k;
  int n_probes = sp.n_probes > 0 ? sp.n_probes : 20;
  n_probes = std::min(n_probes, std::min(d->n_lists, kMaxProbes));
  const int max_size = std::max(32, d->max_list_rows);
  B2VS_CHECK(max_size <= kCapBig, B2VS_EUNSUP, "a list of %d rows exceeds the large-k buffer", max_size);
  const double mean_size = std::max(1.0, static_cast<double>(d->n) / d->n_lists);
  int m = std::min(n_probes, static_cast<int>(std::ceil(3.0 * k_scan / mean_size)) + 1);
  m = std::max(1, std::min(m, kCapBig / max_size));
  const int q_pad = static_cast<int>(round_up(nq, 128));
  B2VS_TRY(d->ws_probe_d.reserve(static_cast<size_t>(nq) * n_probes * sizeof(float)));
  B2VS_TRY(d->ws_probe_i.reserve(static_cast<size_t>(nq) * n_probes * sizeof(int64_t)));
  B2VS_TRY(d->ws_keys.reserve(static_cast<size_t>(nq) * m * sizeof(int64_t)));        // first m probes
  B2VS_TRY(d->ws_qf.reserve(static_cast<size_t>(nq) * d->dp * sizeof(float)));
  B2VS_TRY(d->ws_qnorm.reserve(static_cast<size_t>(q_pad) * sizeof(float)));
  B2VS_TRY(d->ws_counter.reserve(2 * sizeof(unsigned long long) + sizeof(int)));
  B2VS_TRY(d->ws_g_tau.reserve(static_cast<size_t>(nq) * sizeof(float)));
  B2VS_TRY(d->ws_g_cand.reserve(static_cast<size_t>(nq) * kCapBig * sizeof(u64)));
  B2VS_TRY(d->ws_g_cnt.reserve(static_cast<size_t>(nq) * sizeof(int)));
  B2VS_TRY(reserve_item_sort(d, nq * n_probes, kGroupRows));
  B2VS_TRY(index->flat.search(q, q_dtype, nq, n_probes, 0, 0, d->ws_probe_d.as<float>(),
                              d->ws_probe_i.as<int64_t>(), nullptr, st));
  int launches = index->flat.stats.launches;
  DISPATCH_DTYPE(q_dtype, T, (queries_to_f32_kernel<T><<<static_cast<unsigned>(ceil_div(nq, 4)), 128, 0, st>>>(
                                 static_cast<const T*>(q), nq, index->dim, d->dp, d->fmt, 0,
                                 d->ws_qf.as<float>(), d->ws_qnorm.as<float>())));
  B2VS_CUDA(cudaGetLastError());
  unsigned long long* counter = d->ws_counter.as<unsigned long long>();
  int* overflow = reinterpret_cast<int*>(counter + 2);
  B2VS_CUDA(cudaMemsetAsync(counter, 0, 2 * sizeof(unsigned long long) + sizeof(int), st));
  const long long* probe_ids = reinterpret_cast<const long long*>(d->ws_probe_i.ptr);
  B2VS_CUDA(cudaMemcpy2DAsync(d->ws_keys.ptr, static_cast<size_t>(m) * sizeof(int64_t), probe_ids,
                              static_cast<size_t>(n_probes) * sizeof(int64_t),
                              static_cast<size_t>(m) * sizeof(int64_t), nq, cudaMemcpyDeviceToDevice, st));
  fill_f32_kernel<<<static_cast<unsigned>(ceil_div(nq, 256)), 256, 0, st>>>(d->ws_g_tau.as<float>(), nq, INFINITY);
  B2VS_CUDA(cudaMemsetAsync(d->ws_g_cnt.ptr, 0, static_cast<size_t>(nq) * sizeof(int), st));
  B2VS_TRY(run_grouped_pq_scan(index, d, reinterpret_cast<const long long*>(d->ws_keys.ptr), m, nq,
                               kCapBig, nullptr, st));
  B2VS_TRY(launch_bigk_select(d->ws_g_cand.as<u64>(), d->ws_g_cnt.as<int>(), kCapBig, nq, k_scan, 0,
                              index->metric, nullptr, 0, d->ws_g_tau.as<float>(), nullptr, nullptr,
                              nullptr, st));
  B2VS_CUDA(cudaMemsetAsync(d->ws_g_cnt.ptr, 0, static_cast<size_t>(nq) * sizeof(int), st));
  B2VS_TRY(run_grouped_pq_scan(index, d, probe_ids, n_probes, nq, kCapBig, counter, st));
  if (!refine) {
    B2VS_TRY(launch_bigk_select(d->ws_g_cand.as<u64>(), d->ws_g_cnt.as<int>(), kCapBig, nq, k, 1,
                                index->metric, nullptr, index->id_offset, nullptr, out_d, out_i,
                                overflow, st, d->row_ids.as<uint32_t>()));
  } else {
    B2VS_TRY(d->ws_ref_d.reserve(static_cast<size_t>(nq) * k_scan * sizeof(float)));
    B2VS_TRY(d->ws_ref_i.reserve(static_cast<size_t>(nq) * k_scan * sizeof(int64_t)));
    B2VS_TRY(launch_bigk_select(d->ws_g_cand.as<u64>(), d->ws_g_cnt.as<int>(), kCapBig, nq, k_scan, 1,
                                index->metric, nullptr, 0, nullptr, d->ws_ref_d.as<float>(),
                                d->ws_ref_i.as<int64_t>(), overflow, st, d->row_ids.as<uint32_t>()));
    DISPATCH_DTYPE(index->dtype, T, (refine_big_kernel<T><<<nq, kRefineBigThreads, 0, st>>>(
                                        static_cast<const T*>(d->src_rows), index->dim,
                                        d->ws_qf.as<float>(), d->dp,
                                        reinterpret_cast<const long long*>(d->ws_ref_i.ptr), k_scan, k,
                                        index->metric, index->id_offset, out_d,
                                        reinterpret_cast<long long*>(out_i))));
    B2VS_CUDA(cudaGetLastError());
  }
  int h_over = 0;
  B2VS_CUDA(cudaMemcpyAsync(&h_over, overflow, sizeof(int), cudaMemcpyDeviceToHost, st));
  B2VS_CUDA(cudaStreamSynchronize(st));
  B2VS_CHECK(h_over <= kCapBig, B2VS_EUNSUP,
             "large-k IVF-PQ search: %d candidates under the seed threshold exceed the %d-key buffer",
             h_over, kCapBig);
  launches += 24;
  d->stats = b2vs_search_stats{};
  d->stats.launches = launches;
  d->stats.n_splits = n_probes;
  d->stats.grid = nq * n_probes;
  d->stats.algo_flops = 2.0 * nq * static_cast<double>(d->n_lists) * index->dim;
  d->counter_pending = true;
  d->last_nq = nq;
  d->timing_pending = false;
  return B2VS_OK;
}

// ---- K4b coarse probe of very small batches ---------------------------------------------------
// The tensor-core probe works on 128-query blocks: for a handful of queries it is one 256-centroid
// tile per CTA on 16 (4096 lists) to 64 CTAs and takes 30-43 us of a 140-170 us search.  Here (K4b) the
// centroid operand matrix of the flat engine is read as 256-row pseudo-lists by the per-item list
// scan (K5: one CTA per (query, pseudo-list), query slice in registers, warp-resident top-k), and
// the per-chunk lists are folded by merge_splits_kernel - the same two kernels, the same score
// (alpha * q.c + ||c||^2 over the rounded operands) as the tensor-core path.
// Limits measured on B200 (profiles/r1_coarse_scan_ab.txt): against the tensor-core probe the scan
// saves 16-20 % of the whole search from Q = 1 to Q = 32 (IVF-Flat, 4096 lists) / Q = 16 (IVF-PQ,
// 16384 lists = 1024 CTAs); larger batches were not measured and keep the tensor cores.
constexpr int kCoarseScanMaxQueries = 32;   // batch limit (B2VS_COARSE_SCAN_MAXQ overrides, <= 32)
constexpr int kCoarseScanCtasPerSm = 8;     // limit on (queries x pseudo-lists) / SMs (B2VS_COARSE_SCAN_CTAS)
constexpr int kCoarseScanProbeRows = 32;    // rows of the constant pseudo-probe table
constexpr int kCoarseScanRows = 256;

static int env_int(const char* name, int fallback, int lo, int hi) {
  const char* e = std::getenv(name);
  if (!e) return fallback;
  const int v = std::atoi(e);
  return v < lo ? lo : (v > hi ? hi : v);
}

// Number of pseudo-lists, or 0 when the tensor-core probe should run (B2VS_COARSE_SCAN=0 forces that).
static int coarse_scan_chunks(const b2vs_index* index, const IvfData* d, int nq, int n_probes) {
  const char* e = std::getenv("B2VS_COARSE_SCAN");
  if (e && e[0] == '0') return 0;
  if (nq > env_int("B2VS_COARSE_SCAN_MAXQ", kCoarseScanMaxQueries, 1, kCoarseScanProbeRows) ||
      n_probes > kMaxFusedK)
    return 0;
  if (index->flat.split3 || index->flat.kdim != d->dp || (d->dp >> 3) > 32 * 8) return 0;
  const int chunks = static_cast<int>(ceil_div(d->n_lists, kCoarseScanRows));
  const int64_t max_ctas = static_cast<int64_t>(env_int("B2VS_COARSE_SCAN_CTAS", kCoarseScanCtasPerSm, 1, 64)) *
                           sm_count(index->dev);
  if (chunks < 1 || static_cast<int64_t>(nq) * chunks > max_ctas) return 0;
  return chunks;
}

static int coarse_probe_scan(b2vs_index* index, IvfData* d, int nq, int n_probes, int chunks,
                             cudaStream_t st) {
  if (!d->cq_ready) {
    std::vector<uint32_t> offs(static_cast<size_t>(chunks) + 1);
    for (int c = 0; c <= chunks; ++c)
      offs[c] = static_cast<uint32_t>(std::min<int64_t>(d->n_lists, static_cast<int64_t>(c) * kCoarseScanRows));
    std::vector<long long> probes(static_cast<size_t>(kCoarseScanProbeRows) * chunks);
    for (int qi = 0; qi < kCoarseScanProbeRows; ++qi)
      for (int c = 0; c < chunks; ++c) probes[static_cast<size_t>(qi) * chunks + c] = c;
    B2VS_TRY(d->cq_offsets.reserve(offs.size() * sizeof(uint32_t)));
    B2VS_TRY(d->cq_probe.reserve(probes.size() * sizeof(long long)));
    B2VS_CUDA(cudaMemcpy(d->cq_offsets.ptr, offs.data(), offs.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    B2VS_CUDA(cudaMemcpy(d->cq_probe.ptr, probes.data(), probes.size() * sizeof(long long), cudaMemcpyHostToDevice));
    d->cq_ready = true;
  }
  B2VS_TRY(d->ws_cq_keys.reserve(static_cast<size_t>(chunks) * nq * n_probes * sizeof(u64)));
  const int j = static_cast<int>(ceil_div(d->dp / 8, 32));
  const float alpha = index->flat.metric == B2VS_METRIC_L2 ? -2.f : -1.f;
  FLAT_SCAN_DISPATCH(ivf_flat_scan_kernel, index->flat.ab_format, j, nq * chunks, st,
                     static_cast<const uint16_t*>(index->flat.mat), index->flat.beta.as<float>(),
                     d->cq_offsets.as<uint32_t>(), d->cq_probe.as<long long>(), d->ws_qf.as<float>(),
                     d->dp, chunks, nq, n_probes, alpha, d->ws_cq_keys.as<u64>(), nullptr, nullptr);
  B2VS_CUDA(cudaGetLastError());
  return launch_merge_splits(d->ws_cq_keys.as<u64>(), chunks, nq, nq, n_probes, index->flat.metric,
                             d->ws_qnorm.as<float>(), 0, d->ws_probe_d.as<float>(),
                             d->ws_probe_i.as<int64_t>(), nullptr, st);
}

static int ivf_search_batch(b2vs_index* index, const void* q, int q_dtype, int nq, int k,
                            const b2vs_search_params& sp, float* out_d, int64_t* out_i,
                            cudaStream_t st) {
  IvfData* d = static_cast<IvfData*>(index->ivf);
  B2VS_CHECK(k >= 1 && k <= kMaxFusedK, B2VS_EUNSUP, "k=%d outside [1, %d]", k, kMaxFusedK);
  int n_probes = sp.n_probes > 0 ? sp.n_probes : 20;  // cuVS SearchParams default
  n_probes = std::min(n_probes, std::min(d->n_lists, kMaxProbes));
  // IVF-PQ refine: scan for k' = refine_ratio * k ADC candidates, then re-rank them exactly
  const bool refine = index->kind == B2VS_KIND_IVF_PQ && sp.refine_ratio > 1 && d->src_rows != nullptr;
  const int k_final = k;
  if (refine) k = std::min(kMaxFusedK, k * sp.refine_ratio);
  const int q_pad = static_cast<int>(round_up(nq, 128));
  B2VS_TRY(d->ws_probe_d.reserve(static_cast<size_t>(nq) * n_probes * sizeof(float)));
  B2VS_TRY(d->ws_probe_i.reserve(static_cast<size_t>(nq) * n_probes * sizeof(int64_t)));
  B2VS_TRY(d->ws_keys.reserve(static_cast<size_t>(n_probes) * q_pad * k * sizeof(u64)));
  B2VS_TRY(d->ws_qf.reserve(static_cast<size_t>(nq) * d->dp * sizeof(float)));
  B2VS_TRY(d->ws_qnorm.reserve(static_cast<size_t>(q_pad) * sizeof(float)));
  B2VS_TRY(d->ws_counter.reserve(2 * sizeof(unsigned long long)));  // scanned rows, candidates
  const int round16 = (index->kind == B2VS_KIND_IVF_FLAT && index->dtype != B2VS_F32) ? 1 : 0;
  DISPATCH_DTYPE(q_dtype, T, (queries_to_f32_kernel<T><<<static_cast<unsigned>(ceil_div(nq, 4)), 128, 0, st>>>(
                                 static_cast<const T*>(q), nq, index->dim, d->dp, d->fmt, round16,
                                 d->ws_qf.as<float>(), d->ws_qnorm.as<float>())));
  B2VS_CUDA(cudaGetLastError());
  // K4 coarse probe: top-n_probes centroids - on the tensor cores, or (a handful of queries) K4b
  int launches = 2;
  const int coarse_chunks = coarse_scan_chunks(index, d, nq, n_probes);
  if (coarse_chunks > 0) {
    B2VS_TRY(coarse_probe_scan(index, d, nq, n_probes, coarse_chunks, st));
  } else {
    B2VS_TRY(index->flat.search(q, q_dtype, nq, n_probes, 0, 0, d->ws_probe_d.as<float>(),
                                d->ws_probe_i.as<int64_t>(), nullptr, st));
    launches = index->flat.stats.launches;
  }
  B2VS_CUDA(cudaMemsetAsync(d->ws_counter.ptr, 0, 2 * sizeof(unsigned long long), st));
  launches += 2;
  const int items = nq * n_probes;
  const long long* probe_ids = reinterpret_cast<const long long*>(d->ws_probe_i.ptr);
  unsigned long long* counter = d->ws_counter.as<unsigned long long>();
  const float* qnorm_for_merge = nullptr;
  const bool timed = (sp.flags & B2VS_FLAG_TIME_KERNEL) != 0;
  if (timed) {
    if (!d->ev0) {
      B2VS_CUDA(cudaEventCreate(&d->ev0));
      B2VS_CUDA(cudaEventCreate(&d->ev1));
    }
    B2VS_CUDA(cudaEventRecord(d->ev0, st));
  }
  bool single_list = false;  // the scan left ONE sorted list per query (not one per probe)
  if (index->kind == B2VS_KIND_IVF_FLAT) {
    const int j = static_cast<int>(ceil_div(d->dp / 8, 32));
    const float alpha = index->metric == B2VS_METRIC_L2 ? -2.f : -1.f;
    const uint16_t* data = d->data.as<uint16_t>();
    const float* snorm = d->slot_norm.as<float>();
    const uint32_t* offs = d->offsets.as<uint32_t>();
    const float* qf = d->ws_qf.as<float>();
    const int ov = grouped_override();
    // The grouped tensor-core scan is the default at every batch size (measured faster than the
    // per-item scan from Q = 1 up); B2VS_IVF_GROUPED=0 selects the per-item kernels.
    const bool grouped = ov >= 0 ? ov == 1 : true;
    // fp32-source indexes keep fp32 queries against their bf16 rows: the query operand is split
    // into bf16 [hi | lo] halves multiplied against the same list tiles (2x the MMA work)
    const int q_split = index->dtype == B2VS_F32 ? 1 : 0;
    if (grouped) {
      const int cap = grouped_cap(k);
      // seed thresholds first (queries ordered by their nearest list), then group all the items
      B2VS_TRY(reserve_item_sort(d, items, kGroupRows));  // both sorts share these buffers
      const bool order_seeds = nq >= kSeedSortMinQueries;
      if (order_seeds) B2VS_TRY(sort_items_by_list(d, probe_ids, nq, n_probes, 1, st));
      B2VS_TRY(d->ws_g_tau.reserve(static_cast<size_t>(nq) * sizeof(float)));
      B2VS_TRY(d->ws_g_cand.reserve(static_cast<size_t>(nq) * cap * sizeof(u64)));
      B2VS_TRY(d->ws_g_cnt.reserve(static_cast<size_t>(nq) * sizeof(int)));
      B2VS_CUDA(cudaMemsetAsync(d->ws_g_cnt.ptr, 0, static_cast<size_t>(nq) * sizeof(int), st));
      FLAT_SCAN_DISPATCH(ivf_seed_tau_kernel, d->fmt, j, nq, st, data, snorm, offs, probe_ids, qf, d->dp,
                         n_probes, k, alpha, grouped_seed_rows(k), d->max_norm2,
                         index->metric == B2VS_METRIC_L2 ? 1 : 0, q_split ? 1e-5f : 0.f,
                         order_seeds ? d->ws_item_perm.as<uint32_t>() : nullptr, d->ws_g_tau.as<float>());
      B2VS_TRY(run_grouped_flat_scan(index, d, probe_ids, n_probes, nq, cap, counter, st));
      ivf_group_select_kernel<<<nq, kSelectThreads, static_cast<size_t>(cap) * sizeof(u64), st>>>(
          d->ws_g_cand.as<u64>(), d->ws_g_cnt.as<int>(), cap, k, d->ws_keys.as<u64>(), counter + 1);
      FLAT_SCAN_DISPATCH(ivf_flat_rescue_kernel, d->fmt, j, nq, st, data, snorm, offs, probe_ids, qf,
                         d->dp, n_probes, k, alpha, d->ws_g_cnt.as<int>(), cap, d->ws_keys.as<u64>());
      B2VS_CUDA(cudaGetLastError());
      launches += 16;
      single_list = true;
    } else {
      // Per-item scan.  Ordering the items by list pays once several queries share a list: the
      // batch then reads each probed list from HBM about once instead of once per probing query.
      const uint32_t* item_perm = nullptr;
      static const bool no_sort = getenv("B2VS_NO_ITEM_SORT") != nullptr;
      if (!no_sort && items >= 4 * d->n_lists) {
        B2VS_TRY(sort_items_by_list(d, probe_ids, items, 1, 1, st));
        item_perm = d->ws_item_perm.as<uint32_t>();
        launches += 5;
      }
      FLAT_SCAN_DISPATCH(ivf_flat_scan_kernel, d->fmt, j, items, st, data, snorm, offs, probe_ids, qf,
                         d->dp, n_probes, q_pad, k, alpha, d->ws_keys.as<u64>(), counter, item_perm);
      B2VS_CUDA(cudaGetLastError());
    }
    qnorm_for_merge = d->ws_qnorm.as<float>();
  } else if (d->pq_tc_ready && grouped_override() != 0) {
    // Grouped tensor-core scan (pq_tc.cuh): same pipeline as the IVF-Flat one, the list tiles are
    // decoded from the PQ codes instead of loaded.
    const int l2 = index->metric == B2VS_METRIC_L2 ? 1 : 0;
    const float alpha = l2 ? -2.f : -1.f;
    const int cap = grouped_cap(k);
    const uint32_t* offs = d->offsets.as<uint32_t>();
    const float* qf = d->ws_qf.as<float>();
    B2VS_TRY(reserve_item_sort(d, items, kGroupRows));
    const bool order_seeds = nq >= kSeedSortMinQueries;
    if (order_seeds) B2VS_TRY(sort_items_by_list(d, probe_ids, nq, n_probes, 1, st));
    B2VS_TRY(d->ws_g_tau.reserve(static_cast<size_t>(nq) * sizeof(float)));
    B2VS_TRY(d->ws_g_cand.reserve(static_cast<size_t>(nq) * cap * sizeof(u64)));
    B2VS_TRY(d->ws_g_cnt.reserve(static_cast<size_t>(nq) * sizeof(int)));
    B2VS_CUDA(cudaMemsetAsync(d->ws_g_cnt.ptr, 0, static_cast<size_t>(nq) * sizeof(int), st));
    const size_t lut_smem = (static_cast<size_t>(d->pq_dim) * 256 + index->dim) * sizeof(float);
    B2VS_CUDA(cudaFuncSetAttribute(ivf_pq_lut_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(lut_smem)));
    ivf_pq_lut_scan_kernel<<<nq, kScanThreads, lut_smem, st>>>(
        0, d->codes.as<uint8_t>(), d->row_ids.as<uint32_t>(), offs, probe_ids, qf,
        d->centroids.as<float>(), d->cb16.as<uint16_t>(), d->cbn.as<float>(), index->dim, d->dp,
        d->pq_dim, d->dsub, n_probes, k, index->metric, grouped_seed_rows(k), d->max_rhat2,
        order_seeds ? d->ws_item_perm.as<uint32_t>() : nullptr, nullptr, cap, d->ws_g_tau.as<float>(),
        nullptr);
    B2VS_TRY(run_grouped_pq_scan(index, d, probe_ids, n_probes, nq, cap, counter, st));
    ivf_group_select_kernel<<<nq, kSelectThreads, static_cast<size_t>(cap) * sizeof(u64), st>>>(
        d->ws_g_cand.as<u64>(), d->ws_g_cnt.as<int>(), cap, k, d->ws_keys.as<u64>(), counter + 1);
    ivf_pq_lut_scan_kernel<<<nq, kScanThreads, lut_smem, st>>>(
        1, d->codes.as<uint8_t>(), d->row_ids.as<uint32_t>(), offs, probe_ids, qf,
        d->centroids.as<float>(), d->cb16.as<uint16_t>(), d->cbn.as<float>(), index->dim, d->dp,
        d->pq_dim, d->dsub, n_probes, k, index->metric, 0u, d->max_rhat2, nullptr, d->ws_g_cnt.as<int>(),
        cap, nullptr, d->ws_keys.as<u64>());
    B2VS_CUDA(cudaGetLastError());
    launches += 16;
    single_list = true;
  } else {
    const size_t cb_floats = static_cast<size_t>(d->pq_dim) * 256 * d->dsub;
    const size_t smem_p = (cb_floats + static_cast<size_t>(d->mp) * 256 + 2 * index->dim) * sizeof(float);
    if (smem_p + 20 * 1024 <= 227 * 1024 && n_probes <= kMaxFusedK) {
      // codebooks fit next to the LUT: query-major persistent CTAs, one per SM
      const int grid = std::min(nq, sm_count(index->dev));
#define PQ_QUERY_LAUNCH(NCH, DSUB)                                                                      \
  do {                                                                                            \
    B2VS_CUDA(cudaFuncSetAttribute((ivf_pq_scan_query_kernel<NCH, DSUB>),                                 \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize,                   \
                                   static_cast<int>(smem_p)));                                    \
    ivf_pq_scan_query_kernel<NCH, DSUB><<<grid, kPqPersistThreads, smem_p, st>>>(                       \
        d->codes.as<uint8_t>(), d->row_ids.as<uint32_t>(), d->offsets.as<uint32_t>(), probe_ids,   \
        d->ws_qf.as<float>(), d->centroids.as<float>(), d->codebooks.as<float>(), index->dim,      \
        d->dp, d->pq_dim, d->mp, d->dsub, n_probes, nq, k, index->metric, d->ws_keys.as<u64>(),    \
        counter);                                                                                 \
  } while (0)
      const int nch = d->mp >> 4;
      if (nch == 4 && d->dsub == 2) PQ_QUERY_LAUNCH(4, 2);
      else if (nch == 2 && d->dsub == 4) PQ_QUERY_LAUNCH(2, 4);
      else if (nch == 1 && d->dsub == 8) PQ_QUERY_LAUNCH(1, 8);
      else PQ_QUERY_LAUNCH(0, 0);
#undef PQ_QUERY_LAUNCH
      single_list = true;
    } else {
      const size_t smem = (static_cast<size_t>(d->mp) * 256 + index->dim) * sizeof(float);
      B2VS_CUDA(cudaFuncSetAttribute(ivf_pq_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem)));
      ivf_pq_scan_kernel<<<items, kScanThreads, smem, st>>>(
          d->codes.as<uint8_t>(), d->row_ids.as<uint32_t>(), d->offsets.as<uint32_t>(), probe_ids,
          d->ws_qf.as<float>(), d->centroids.as<float>(), d->codebooks.as<float>(), index->dim, d->dp,
          d->pq_dim, d->mp, d->dsub, n_probes, q_pad, k, index->metric, d->ws_keys.as<u64>(), counter);
    }
    B2VS_CUDA(cudaGetLastError());
  }
  if (timed) B2VS_CUDA(cudaEventRecord(d->ev1, st));
  ++launches;
  if (!refine) {
    B2VS_TRY(launch_merge_splits(d->ws_keys.as<u64>(), single_list ? 1 : n_probes, q_pad, nq, k,
                                 index->metric, qnorm_for_merge, index->id_offset, out_d, out_i,
                                 nullptr, st, d->row_ids.as<uint32_t>()));
    ++launches;
  } else {
    B2VS_TRY(d->ws_ref_d.reserve(static_cast<size_t>(nq) * k * sizeof(float)));
    B2VS_TRY(d->ws_ref_i.reserve(static_cast<size_t>(nq) * k * sizeof(int64_t)));
    B2VS_TRY(launch_merge_splits(d->ws_keys.as<u64>(), single_list ? 1 : n_probes, q_pad, nq, k,
                                 index->metric, qnorm_for_merge, 0, d->ws_ref_d.as<float>(),
                                 d->ws_ref_i.as<int64_t>(), nullptr, st, d->row_ids.as<uint32_t>()));
    DISPATCH_DTYPE(index->dtype, T, (refine_kernel<T><<<static_cast<unsigned>(ceil_div(nq, 4)), 128, 0, st>>>(
                                        static_cast<const T*>(d->src_rows), index->dim,
                                        d->ws_qf.as<float>(), d->dp,
                                        reinterpret_cast<const long long*>(d->ws_ref_i.ptr), nq, k,
                                        k_final, index->metric, index->id_offset, out_d,
                                        reinterpret_cast<long long*>(out_i))));
    B2VS_CUDA(cudaGetLastError());
    launches += 2;
  }
  d->stats = b2vs_search_stats{};
  d->stats.launches = launches;
  d->stats.n_splits = n_probes;
  d->stats.grid = items;
  d->stats.algo_flops = 2.0 * nq * static_cast<double>(d->n_lists) * index->dim;
  d->counter_pending = true;
  d->last_nq = nq;
  d->timing_pending = timed;
  return B2VS_OK;
}

// ------------------------------------------------------------------------------------------
// Small batches as one CUDA graph.  B2VS_GRAPH=1 enables it for every eligible call, =0 disables
// it even when a call asks with B2VS_FLAG_GRAPH; unset = only calls that set the flag.
static int graph_override() {
  static const int v = [] {
    const char* e = std::getenv("B2VS_GRAPH");
    return e ? (e[0] == '0' ? 0 : 1) : -1;
  }();
  return v;
}

static void drop_graph_error() { cudaGetLastError(); }

// Replays (capturing first if needed) the search of this signature.  *handled = false means the
// caller must run the direct path (first sighting of the signature, or capture was refused).
static int ivf_search_graphed(b2vs_index* index, IvfData* d, const void* q, int q_dtype, int nq, int k,
                              const b2vs_search_params& sp, float* out_d, int64_t* out_i,
                              cudaStream_t st, bool* handled) {
  *handled = false;
  const int flags = sp.flags & ~B2VS_FLAG_GRAPH;
  SearchGraph* g = nullptr;
  for (SearchGraph& e : d->graphs)
    if (e.nq == nq && e.k == k && e.q_dtype == q_dtype && e.n_probes == sp.n_probes &&
        e.refine_ratio == sp.refine_ratio && e.flags == flags) { g = &e; break; }
  if (!g) {
    if (static_cast<int>(d->graphs.size()) >= kGraphMaxEntries) {
      size_t victim = 0;
      for (size_t i = 1; i < d->graphs.size(); ++i)
        if (d->graphs[i].last_use < d->graphs[victim].last_use) victim = i;
      // the victim's graph may still be running on the caller's stream
      if (d->graphs[victim].exec) cudaStreamSynchronize(st);
      d->graphs[victim].destroy();
      d->graphs.erase(d->graphs.begin() + static_cast<long>(victim));
    }
    d->graphs.emplace_back();
    g = &d->graphs.back();
    g->nq = nq; g->k = k; g->q_dtype = q_dtype; g->n_probes = sp.n_probes;
    g->refine_ratio = sp.refine_ratio; g->flags = flags;
  }
  g->last_use = ++d->graph_clock;
  if (g->failed) return B2VS_OK;
  if (g->exec && g->generation != realloc_generation()) {
    // some workspace moved since the capture: the graph's pointers may be stale
    cudaStreamSynchronize(st);
    cudaGraphExecDestroy(g->exec);
    g->exec = nullptr;
  }
  if (!g->exec) {
    // the first call of a signature runs directly and sizes every workspace, so that the
    // capture below allocates nothing
    if (g->seen++ == 0) return B2VS_OK;
    if (!d->cap_stream)
      B2VS_CUDA(cudaStreamCreateWithFlags(&d->cap_stream, cudaStreamNonBlocking));
    if (!g->io) {
      g->q_bytes = static_cast<size_t>(nq) * index->dim * elem_bytes(q_dtype);
      g->d_off = static_cast<size_t>(round_up(static_cast<int64_t>(g->q_bytes), 256));
      g->i_off = g->d_off + static_cast<size_t>(round_up(static_cast<int64_t>(nq) * k * sizeof(float), 256));
      void* p = nullptr;
      if (cudaMalloc(&p, g->i_off + static_cast<size_t>(nq) * k * sizeof(int64_t)) != cudaSuccess) {
        drop_graph_error();
        g->failed = true;
        return B2VS_OK;
      }
      g->io = static_cast<char*>(p);
    }
    b2vs_search_params spd = sp;
    spd.flags = flags;
    const uint64_t gen0 = realloc_generation();
    if (cudaStreamBeginCapture(d->cap_stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
      drop_graph_error();
      g->failed = true;
      return B2VS_OK;
    }
    const int rc = ivf_search_direct(index, g->io, q_dtype, nq, k, spd,
                                     reinterpret_cast<float*>(g->io + g->d_off),
                                     reinterpret_cast<int64_t*>(g->io + g->i_off), d->cap_stream);
    cudaGraph_t graph = nullptr;
    const cudaError_t e_end = cudaStreamEndCapture(d->cap_stream, &graph);
    size_t n_nodes = 0;
    bool ok = rc == B2VS_OK && e_end == cudaSuccess && graph != nullptr &&
              gen0 == realloc_generation() &&
              cudaGraphGetNodes(graph, nullptr, &n_nodes) == cudaSuccess &&
              cudaGraphInstantiate(&g->exec, graph, 0ull) == cudaSuccess;
    if (graph) cudaGraphDestroy(graph);
    if (!ok) {
      drop_graph_error();
      if (g->exec) cudaGraphExecDestroy(g->exec);
      g->exec = nullptr;
      // a workspace grew under the capture (another signature's sizes): try again next call;
      // anything else is a refusal
      if (gen0 == realloc_generation()) g->failed = true;
      return B2VS_OK;
    }
    g->generation = gen0;
    g->stats = d->stats;
    g->stats.launches = static_cast<int32_t>(n_nodes);
  }
  B2VS_CUDA(cudaMemcpyAsync(g->io, q, g->q_bytes, cudaMemcpyDeviceToDevice, st));
  B2VS_CUDA(cudaGraphLaunch(g->exec, st));
  B2VS_CUDA(cudaMemcpyAsync(out_d, g->io + g->d_off, static_cast<size_t>(nq) * k * sizeof(float),
                            cudaMemcpyDeviceToDevice, st));
  B2VS_CUDA(cudaMemcpyAsync(out_i, g->io + g->i_off, static_cast<size_t>(nq) * k * sizeof(int64_t),
                            cudaMemcpyDeviceToDevice, st));
  d->stats = g->stats;
  d->counter_pending = true;
  d->last_nq = nq;
  d->timing_pending = false;
  *handled = true;
  return B2VS_OK;
}

int ivf_search(b2vs_index* index, const void* q, int q_dtype, int nq, int k,
               const b2vs_search_params& sp, float* out_d, int64_t* out_i, cudaStream_t st) {
  IvfData* d = static_cast<IvfData*>(index->ivf);
  B2VS_CHECK(d != nullptr, B2VS_EINVAL, "IVF index has no list data");
  const int ov = graph_override();
  const bool want_graph = ov == 1 || (ov < 0 && (sp.flags & B2VS_FLAG_GRAPH) != 0);
  if (want_graph && nq <= kGraphMaxQueries && k >= 1 && (sp.flags & B2VS_FLAG_TIME_KERNEL) == 0 &&
      !uses_bigk_path(index, d, k, sp)) {
    bool handled = false;
    B2VS_TRY(ivf_search_graphed(index, d, q, q_dtype, nq, k, sp, out_d, out_i, st, &handled));
    if (handled) return B2VS_OK;
  }
  b2vs_search_params spd = sp;
  spd.flags &= ~B2VS_FLAG_GRAPH;
  return ivf_search_direct(index, q, q_dtype, nq, k, spd, out_d, out_i, st);
}

void ivf_fill_info(const b2vs_index* index, b2vs_index_info* info) {
  const IvfData* d = static_cast<const IvfData*>(index->ivf);
  if (!d) return;
  info->n_lists = d->n_lists;
  info->pq_dim = d->pq_dim;
  info->pq_bits = d->pq_bits;
  info->device_bytes += static_cast<int64_t>(d->owned_bytes());
}

void ivf_last_stats(const b2vs_index* index, b2vs_search_stats* stats) {
  IvfData* d = static_cast<IvfData*>(index->ivf);
  *stats = b2vs_search_stats{};
  if (!d) return;
  if (d->counter_pending && d->ws_counter.ptr) {
    unsigned long long cnt[2] = {0, 0};
    DeviceGuard guard(index->dev);
    if (cudaMemcpy(cnt, d->ws_counter.ptr, sizeof(cnt), cudaMemcpyDeviceToHost) == cudaSuccess) {
      d->stats.algo_bytes = static_cast<double>(cnt[0]) * d->row_bytes;
      d->stats.mean_candidates = static_cast<int32_t>(cnt[1] / static_cast<unsigned long long>(std::max(d->last_nq, 1)));
    }
    d->counter_pending = false;
  }
  if (d->timing_pending && d->ev1) {
    float ms = 0.f;
    DeviceGuard guard(index->dev);
    if (cudaEventSynchronize(d->ev1) == cudaSuccess &&
        cudaEventElapsedTime(&ms, d->ev0, d->ev1) == cudaSuccess)
      d->stats.kernel_ms = ms;
    else
      cudaGetLastError();
    d->timing_pending = false;
  }
  *stats = d->stats;
}

void ivf_destroy(b2vs_index* index) {
  IvfData* d = static_cast<IvfData*>(index->ivf);
  if (d) {
    d->destroy();
    delete d;
  }
  index->ivf = nullptr;
}

// ------------------------------------------------------------------------------------------
static int ivf_build(int kind, int dev, int metric, int dtype, int dim, const void* db, int64_t n,
                     int64_t id_offset, const b2vs_ivf_params* params, cudaStream_t st,
                     b2vs_index** out) {
  B2VS_TRY(check_matrix_args(dev, metric, dtype, dim, db, n, out));
  B2VS_CHECK(params != nullptr, B2VS_EINVAL, "IVF params are NULL");
  B2VS_CHECK(n >= 1 && n < (1ll << 31) - (1ll << 22), B2VS_EINVAL,
             "IVF build needs 1 <= n < 2^31 (n=%lld)", static_cast<long long>(n));
  B2VS_CHECK(dim <= 2048, B2VS_EUNSUP, "IVF supports dim <= 2048 (got %d)", dim);
  const int n_lists = params->n_lists;
  B2VS_CHECK(n_lists >= 1 && n_lists <= n, B2VS_EINVAL, "n_lists=%d must be in [1, n=%lld]", n_lists,
             static_cast<long long>(n));
  const int iters = params->kmeans_iters > 0 ? params->kmeans_iters : 20;
  const float frac = (params->train_fraction > 0.f && params->train_fraction <= 1.f)
                         ? params->train_fraction : 0.5f;
  int pq_dim = 0, dsub = 0;
  if (kind == B2VS_KIND_IVF_PQ) {
    pq_dim = params->pq_dim;
    B2VS_CHECK(params->pq_bits == 8 || params->pq_bits == 0, B2VS_EUNSUP,
               "only pq_bits=8 is supported (got %d)", params->pq_bits);
    B2VS_CHECK(pq_dim >= 1 && dim % pq_dim == 0, B2VS_EINVAL,
               "pq_dim=%d must divide dim=%d", pq_dim, dim);
    dsub = dim / pq_dim;
    B2VS_CHECK(dsub <= 16, B2VS_EUNSUP, "sub-vector length %d > 16 not supported", dsub);
    B2VS_CHECK(n >= 256, B2VS_EINVAL, "IVF-PQ needs at least 256 rows to train codebooks");
    B2VS_CHECK(pq_dim <= 200, B2VS_EUNSUP, "pq_dim=%d too large for the smem LUT", pq_dim);
  }
  DeviceGuard guard(dev);
  B2VS_CHECK(guard.ok, B2VS_ECUDA, "cannot select device %d", dev);

  b2vs_index* ix = new (std::nothrow) b2vs_index();
  IvfData* d = new (std::nothrow) IvfData();
  B2VS_CHECK(ix && d, B2VS_ENOMEM, "host allocation failed");
  ix->kind = kind; ix->dev = dev; ix->metric = metric; ix->dtype = dtype; ix->dim = dim;
  ix->n = n; ix->id_offset = id_offset; ix->ivf = d;
  d->n_lists = n_lists; d->n = n; d->pq_dim = pq_dim; d->pq_bits = pq_dim ? 8 : 0; d->dsub = dsub;
  d->mp = static_cast<int>(round_up(pq_dim, 16));
  d->dp = static_cast<int>(round_up(dim, 8));
  d->fmt = (dtype == B2VS_F16) ? 0 : 1;
  d->row_bytes = (kind == B2VS_KIND_IVF_FLAT) ? d->dp * 2 : pq_dim;
  if (kind == B2VS_KIND_IVF_PQ) d->src_rows = db;  // borrowed: only dereferenced by refine

  DevBuf train, labels, cursor, slot_of_row, slices;
  FlatEngine assign_eng;
  int rc = B2VS_OK;
  auto fail = [&](int code) {
    train.release(); labels.release(); cursor.release(); slot_of_row.release(); slices.release();
    assign_eng.destroy();
    ivf_destroy(ix);
    ix->flat.destroy();
    delete ix;
    return code;
  };
#define IB_TRY(expr) do { rc = (expr); if (rc != B2VS_OK) return fail(rc); } while (0)
#define IB_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); return fail(B2VS_ECUDA); } } while (0)

  // ---- 1. training subsample (every stride-th row, capped at 1024 rows per list)
  int64_t stride = std::max<int64_t>(1, static_cast<int64_t>(std::floor(1.0 / frac + 1e-6)));
  const int64_t cap = std::max<int64_t>(static_cast<int64_t>(n_lists) * 1024, 1);
  if (ceil_div(n, stride) > cap) stride = ceil_div(n, cap);
  int64_t n_train = ceil_div(n, stride);
  if (n_train < n_lists) { stride = 1; n_train = n; }
  const void* train_ptr = db;
  const int eb = elem_bytes(dtype);
  if (stride > 1) {
    IB_TRY(train.reserve(static_cast<size_t>(n_train) * dim * eb));
    const int blocks = static_cast<int>(std::min<int64_t>(ceil_div(n_train * dim, 256), 148 * 32));
    DISPATCH_DTYPE(dtype, T, (strided_rows_kernel<T><<<blocks, 256, 0, st>>>(
                                 static_cast<const T*>(db), train.as<T>(), n_train, stride, dim)));
    IB_CUDA(cudaGetLastError());
    train_ptr = train.ptr;
  }
  // ---- 2. coarse k-means
  IB_TRY(d->centroids.reserve(static_cast<size_t>(n_lists) * dim * sizeof(float)));
  IB_TRY(kmeans_fit_impl(dev, dtype, dim, train_ptr, n_train, n_lists, iters, params->seed,
                         d->centroids.as<float>(), nullptr, st));
  train.release();
  // ---- 3. assign every row (tensor-core GEMM + arg-min)
  const int force = (dtype == B2VS_F32) ? -1 : d->fmt;
  IB_TRY(assign_eng.init(dev, B2VS_METRIC_L2, B2VS_F32, dim, d->centroids.ptr, n_lists, st, force));
  IB_TRY(labels.reserve(static_cast<size_t>(n) * sizeof(int32_t)));
  IB_TRY(assign_eng.search(db, dtype, static_cast<int>(n), 1, 0, 0, nullptr, nullptr,
                           labels.as<int32_t>(), st));
  // ---- 4. lists: histogram -> scan -> scatter -> fill
  const int pad = 32;  // lists start on 32-slot boundaries (interleaved PQ groups; grouped flat scan)
  IB_TRY(d->sizes.reserve(static_cast<size_t>(n_lists) * sizeof(int)));
  IB_TRY(d->offsets.reserve(static_cast<size_t>(n_lists + 1) * sizeof(uint32_t)));
  IB_TRY(cursor.reserve(static_cast<size_t>(n_lists) * sizeof(int)));
  IB_TRY(slot_of_row.reserve(static_cast<size_t>(n) * sizeof(uint32_t)));
  IB_CUDA(cudaMemsetAsync(d->sizes.ptr, 0, static_cast<size_t>(n_lists) * sizeof(int), st));
  IB_CUDA(cudaMemsetAsync(cursor.ptr, 0, static_cast<size_t>(n_lists) * sizeof(int), st));
  const int eblocks = static_cast<int>(std::min<int64_t>(ceil_div(n, 256), 148 * 32));
  histogram_kernel<<<eblocks, 256, 0, st>>>(labels.as<int>(), n, d->sizes.as<int>());
  IB_CUDA(cudaGetLastError());
  scan_sizes_kernel<<<1, 1024, 0, st>>>(d->sizes.as<int>(), n_lists, pad, d->offsets.as<uint32_t>());
  IB_CUDA(cudaGetLastError());
  uint32_t total_slots = 0;
  IB_CUDA(cudaMemcpyAsync(&total_slots, d->offsets.as<uint32_t>() + n_lists, sizeof(uint32_t),
                          cudaMemcpyDeviceToHost, st));
  IB_CUDA(cudaStreamSynchronize(st));
  d->n_slots = total_slots;
  d->h_sizes.resize(n_lists);
  IB_CUDA(cudaMemcpy(d->h_sizes.data(), d->sizes.ptr, static_cast<size_t>(n_lists) * sizeof(int),
                     cudaMemcpyDeviceToHost));
  IB_TRY(build_list_ranks(d));
  IB_TRY(d->row_ids.reserve(std::max<size_t>(total_slots, 1) * sizeof(uint32_t)));
  IB_CUDA(cudaMemsetAsync(d->row_ids.ptr, 0xFF, std::max<size_t>(total_slots, 1) * sizeof(uint32_t), st));
  scatter_rows_kernel<<<eblocks, 256, 0, st>>>(labels.as<int>(), n, d->offsets.as<uint32_t>(),
                                               cursor.as<int>(), d->row_ids.as<uint32_t>(),
                                               slot_of_row.as<uint32_t>());
  IB_CUDA(cudaGetLastError());
  const int wblocks = static_cast<int>(std::min<int64_t>(ceil_div(n, 8), 148 * 16));
  if (kind == B2VS_KIND_IVF_FLAT) {
    IB_TRY(d->data.reserve(std::max<size_t>(total_slots, 1) * d->dp * 2));
    IB_CUDA(cudaMemsetAsync(cursor.ptr, 0, sizeof(unsigned int), st));  // re-used as the max-norm cell
    const size_t norm_floats = std::max<size_t>(total_slots, 1) + kNormSlack;
    IB_TRY(d->slot_norm.reserve(norm_floats * sizeof(float)));
    IB_CUDA(cudaMemsetAsync(d->data.ptr, 0, std::max<size_t>(total_slots, 1) * d->dp * 2, st));
    fill_f32_kernel<<<static_cast<unsigned>(ceil_div(norm_floats, 256)), 256, 0, st>>>(
        d->slot_norm.as<float>(), norm_floats, INFINITY);
    DISPATCH_DTYPE(dtype, T, (fill_flat_lists_kernel<T><<<wblocks, 256, 0, st>>>(
                                 static_cast<const T*>(db), n, dim, d->dp, d->fmt,
                                 slot_of_row.as<uint32_t>(), d->data.as<uint16_t>(),
                                 d->slot_norm.as<float>(), metric == B2VS_METRIC_L2 ? 1 : 0,
                                 cursor.as<unsigned int>())));
    IB_CUDA(cudaGetLastError());
    IB_CUDA(cudaMemcpyAsync(&d->max_norm2, cursor.ptr, sizeof(float), cudaMemcpyDeviceToHost, st));
    IB_CUDA(cudaStreamSynchronize(st));
  } else {
    // ---- 5. PQ codebooks on residual sub-vectors of a training subset, then encode all rows
    int64_t pstride = std::max<int64_t>(1, n / 131072);
    int64_t p_train = n / pstride;  // rows 0, pstride, ... all < n
    if (p_train < 256) { pstride = 1; p_train = n; }
    IB_TRY(slices.reserve(static_cast<size_t>(p_train) * dim * sizeof(float)));
    const int sblocks = static_cast<int>(std::min<int64_t>(ceil_div(p_train * dim, 256), 148 * 32));
    DISPATCH_DTYPE(dtype, T, (pq_train_slices_kernel<T><<<sblocks, 256, 0, st>>>(
                                 static_cast<const T*>(db), labels.as<int>(), d->centroids.as<float>(),
                                 p_train, pstride, dim, dsub, slices.as<float>())));
    IB_CUDA(cudaGetLastError());
    IB_TRY(d->codebooks.reserve(static_cast<size_t>(pq_dim) * 256 * dsub * sizeof(float)));
    const int pq_iters = std::min(iters, 10);
    KmWorkspace km_ws;   // one set of temporaries for all sub-codebooks
    for (int m = 0; m < pq_dim; ++m) {
      rc = kmeans_fit_impl(dev, B2VS_F32, dsub, slices.as<float>() + static_cast<size_t>(m) * p_train * dsub,
                           p_train, 256, pq_iters, params->seed + 31ull * (m + 1),
                           d->codebooks.as<float>() + static_cast<size_t>(m) * 256 * dsub, nullptr, st,
                           &km_ws);
      if (rc != B2VS_OK) break;
    }
    km_ws.release();
    if (rc != B2VS_OK) return fail(rc);
    slices.release();
    const size_t code_bytes = static_cast<size_t>(std::max<uint32_t>(total_slots, 32)) * d->mp;
    IB_TRY(d->codes.reserve(code_bytes));
    IB_CUDA(cudaMemsetAsync(d->codes.ptr, 0, code_bytes, st));
    dim3 grid(static_cast<unsigned>(std::min<int64_t>(ceil_div(n, 256), 148 * 8)), pq_dim);
    DISPATCH_DTYPE(dtype, T, (pq_encode_kernel<T><<<grid, 256, 256 * dsub * sizeof(float), st>>>(
                                 static_cast<const T*>(db), labels.as<int>(), d->centroids.as<float>(),
                                 d->codebooks.as<float>(), slot_of_row.as<uint32_t>(), n, dim, dsub,
                                 d->mp, d->codes.as<uint8_t>())));
    IB_CUDA(cudaGetLastError());
    IB_TRY(pq_prepare_grouped(ix, d, st));
  }
  // ---- 6. coarse quantizer used at search time (index metric)
  if (metric == B2VS_METRIC_L2) {
    ix->flat = assign_eng;          // take ownership of the buffers
    assign_eng = FlatEngine();
  } else {
    IB_TRY(ix->flat.init(dev, metric, B2VS_F32, dim, d->centroids.ptr, n_lists, st, force));
  }
  IB_CUDA(cudaStreamSynchronize(st));
  labels.release(); cursor.release(); slot_of_row.release();
  assign_eng.destroy();
#undef IB_TRY
#undef IB_CUDA
  *out = ix;
  return B2VS_OK;
}

}  // namespace b2vs

using namespace b2vs;

extern "C" int b2vs_ivfflat_build(int dev, int metric, int dtype, int dim, const void* db, int64_t n,
                                  int64_t id_offset, const b2vs_ivf_params* params, void* stream,
                                  b2vs_index** out) {
  return ivf_build(B2VS_KIND_IVF_FLAT, dev, metric, dtype, dim, db, n, id_offset, params,
                   static_cast<cudaStream_t>(stream), out);
}

extern "C" int b2vs_ivfpq_build(int dev, int metric, int dtype, int dim, const void* db, int64_t n,
                                int64_t id_offset, const b2vs_ivf_params* params, void* stream,
                                b2vs_index** out) {
  return ivf_build(B2VS_KIND_IVF_PQ, dev, metric, dtype, dim, db, n, id_offset, params,
                   static_cast<cudaStream_t>(stream), out);
}

extern "C" int b2vs_kmeans_fit(int dev, int dtype, int dim, const void* x, int64_t n, int n_clusters,
                               int iters, uint64_t seed, float* centroids, int32_t* labels,
                               void* stream) {
  B2VS_CHECK(x != nullptr && centroids != nullptr, B2VS_EINVAL, "NULL pointer passed to b2vs_kmeans_fit");
  B2VS_CHECK(dtype == B2VS_F32 || dtype == B2VS_F16 || dtype == B2VS_BF16, B2VS_EINVAL,
             "unknown dtype %d", dtype);
  B2VS_CHECK(dim >= 1 && dim <= 16384, B2VS_EINVAL, "dim=%d outside [1, 16384]", dim);
  B2VS_CHECK(iters >= 0, B2VS_EINVAL, "iters must be >= 0");
  DeviceGuard guard(dev);
  B2VS_CHECK(guard.ok, B2VS_ECUDA, "cannot select device %d", dev);
  return kmeans_fit_impl(dev, dtype, dim, x, n, n_clusters, iters, seed, centroids, labels,
                         static_cast<cudaStream_t>(stream));
}

extern "C" int b2vs_ivf_list_sizes_host(const b2vs_index* index, int32_t* sizes_host) {
  B2VS_CHECK(index && sizes_host, B2VS_EINVAL, "NULL argument");
  const IvfData* d = static_cast<const IvfData*>(index->ivf);
  B2VS_CHECK(d != nullptr, B2VS_EINVAL, "not an IVF index");
  std::copy(d->h_sizes.begin(), d->h_sizes.end(), sizes_host);
  return B2VS_OK;
}

extern "C" int b2vs_ivf_centroids_host(const b2vs_index* index, float* centroids_host) {
  B2VS_CHECK(index && centroids_host, B2VS_EINVAL, "NULL argument");
  const IvfData* d = static_cast<const IvfData*>(index->ivf);
  B2VS_CHECK(d != nullptr, B2VS_EINVAL, "not an IVF index");
  DeviceGuard guard(index->dev);
  B2VS_CUDA(cudaMemcpy(centroids_host, d->centroids.ptr,
                       static_cast<size_t>(d->n_lists) * index->dim * sizeof(float),
                       cudaMemcpyDeviceToHost));
  return B2VS_OK;
}

// ------------------------------------------------------------------------------------------
// Index persistence: the reference rebuilds every index on every run (it only
// saves raw embeddings as .pt files, cuvs-2gpu-main.ipynb cells 10/12).  An IVF index is written
// as one little-endian file: header + raw device arrays; loading re-creates the coarse engine.
namespace b2vs {

struct IndexFileHeader {
  char magic[8];       // "B2VSIDX2"
  int32_t kind, metric, dtype, dim, n_lists, pq_dim, pq_bits, dsub, mp, dp, fmt, row_bytes;
  int64_t n, id_offset, n_slots;
  float max_norm2;
  int32_t reserved;
  uint64_t bytes_centroids, bytes_offsets, bytes_sizes, bytes_row_ids, bytes_data, bytes_slot_norm,
      bytes_codebooks, bytes_codes;
};

static int write_section(FILE* f, const DevBuf& b, uint64_t bytes) {
  if (bytes == 0) return B2VS_OK;
  std::vector<char> host(std::min<uint64_t>(bytes, 64ull << 20));
  for (uint64_t off = 0; off < bytes; off += host.size()) {
    const uint64_t len = std::min<uint64_t>(host.size(), bytes - off);
    B2VS_CUDA(cudaMemcpy(host.data(), static_cast<const char*>(b.ptr) + off, len, cudaMemcpyDeviceToHost));
    B2VS_CHECK(fwrite(host.data(), 1, len, f) == len, B2VS_EINVAL, "short write while saving index");
  }
  return B2VS_OK;
}

static int read_section(FILE* f, DevBuf* b, uint64_t bytes) {
  if (bytes == 0) return B2VS_OK;
  B2VS_TRY(b->reserve(bytes));
  std::vector<char> host(std::min<uint64_t>(bytes, 64ull << 20));
  for (uint64_t off = 0; off < bytes; off += host.size()) {
    const uint64_t len = std::min<uint64_t>(host.size(), bytes - off);
    B2VS_CHECK(fread(host.data(), 1, len, f) == len, B2VS_EINVAL, "index file is truncated");
    B2VS_CUDA(cudaMemcpy(static_cast<char*>(b->ptr) + off, host.data(), len, cudaMemcpyHostToDevice));
  }
  return B2VS_OK;
}

}  // namespace b2vs

extern "C" int b2vs_index_save(const b2vs_index* index, const char* path) {
  B2VS_CHECK(index && path, B2VS_EINVAL, "NULL argument");
  B2VS_CHECK(index->kind != B2VS_KIND_FLAT, B2VS_EUNSUP,
             "flat indexes borrow their rows and hold no trained state: re-create them instead");
  const IvfData* d = static_cast<const IvfData*>(index->ivf);
  B2VS_CHECK(d != nullptr, B2VS_EINVAL, "index has no list data");
  DeviceGuard guard(index->dev);
  B2VS_CUDA(cudaDeviceSynchronize());
  IndexFileHeader h{};
  std::memcpy(h.magic, "B2VSIDX2", 8);
  h.kind = index->kind; h.metric = index->metric; h.dtype = index->dtype; h.dim = index->dim;
  h.n_lists = d->n_lists; h.pq_dim = d->pq_dim; h.pq_bits = d->pq_bits; h.dsub = d->dsub; h.mp = d->mp;
  h.dp = d->dp; h.fmt = d->fmt; h.row_bytes = d->row_bytes;
  h.n = index->n; h.id_offset = index->id_offset; h.n_slots = d->n_slots;
  h.max_norm2 = d->max_norm2;
  const uint64_t slots = std::max<int64_t>(d->n_slots, 1);
  h.bytes_centroids = static_cast<uint64_t>(d->n_lists) * index->dim * sizeof(float);
  h.bytes_offsets = static_cast<uint64_t>(d->n_lists + 1) * sizeof(uint32_t);
  h.bytes_sizes = static_cast<uint64_t>(d->n_lists) * sizeof(int);
  h.bytes_row_ids = slots * sizeof(uint32_t);
  if (index->kind == B2VS_KIND_IVF_FLAT) {
    h.bytes_data = slots * d->dp * 2;
    h.bytes_slot_norm = (slots + kNormSlack) * sizeof(float);
  } else {
    h.bytes_codebooks = static_cast<uint64_t>(d->pq_dim) * 256 * d->dsub * sizeof(float);
    h.bytes_codes = std::max<uint64_t>(d->n_slots, 32) * d->mp;
  }
  FILE* f = fopen(path, "wb");
  B2VS_CHECK(f != nullptr, B2VS_EINVAL, "cannot open %s for writing", path);
  int rc = fwrite(&h, sizeof(h), 1, f) == 1 ? B2VS_OK : B2VS_EINVAL;
  if (rc == B2VS_OK) rc = write_section(f, d->centroids, h.bytes_centroids);
  if (rc == B2VS_OK) rc = write_section(f, d->offsets, h.bytes_offsets);
  if (rc == B2VS_OK) rc = write_section(f, d->sizes, h.bytes_sizes);
  if (rc == B2VS_OK) rc = write_section(f, d->row_ids, h.bytes_row_ids);
  if (rc == B2VS_OK) rc = write_section(f, d->data, h.bytes_data);
  if (rc == B2VS_OK) rc = write_section(f, d->slot_norm, h.bytes_slot_norm);
  if (rc == B2VS_OK) rc = write_section(f, d->codebooks, h.bytes_codebooks);
  if (rc == B2VS_OK) rc = write_section(f, d->codes, h.bytes_codes);
  fclose(f);
  if (rc != B2VS_OK && b2vs_last_error()[0] == 0) set_error("write to %s failed", path);
  return rc;
}

extern "C" int b2vs_index_load(int dev, const char* path, const void* rows_for_refine, int64_t id_offset,
                               void* stream, b2vs_index** out) {
  B2VS_CHECK(path && out, B2VS_EINVAL, "NULL argument");
  *out = nullptr;
  int count = 0;
  B2VS_CUDA(cudaGetDeviceCount(&count));
  B2VS_CHECK(dev >= 0 && dev < count, B2VS_EINVAL, "device %d not in [0, %d)", dev, count);
  FILE* f = fopen(path, "rb");
  B2VS_CHECK(f != nullptr, B2VS_EINVAL, "cannot open %s", path);
  IndexFileHeader h{};
  if (fread(&h, sizeof(h), 1, f) != 1 || std::memcmp(h.magic, "B2VSIDX2", 8) != 0) {
    fclose(f);
    set_error("%s is not a b2vs index file", path);
    return B2VS_EINVAL;
  }
  DeviceGuard guard(dev);
  b2vs_index* ix = new (std::nothrow) b2vs_index();
  IvfData* d = new (std::nothrow) IvfData();
  if (!ix || !d) { fclose(f); set_error("host allocation failed"); return B2VS_ENOMEM; }
  ix->kind = h.kind; ix->dev = dev; ix->metric = h.metric; ix->dtype = h.dtype; ix->dim = h.dim;
  ix->n = h.n; ix->id_offset = id_offset >= 0 ? id_offset : h.id_offset; ix->ivf = d;
  d->n_lists = h.n_lists; d->pq_dim = h.pq_dim; d->pq_bits = h.pq_bits; d->dsub = h.dsub; d->mp = h.mp;
  d->dp = h.dp; d->fmt = h.fmt; d->row_bytes = h.row_bytes; d->n = h.n; d->n_slots = h.n_slots;
  d->max_norm2 = h.max_norm2;
  d->src_rows = rows_for_refine;
  int rc = read_section(f, &d->centroids, h.bytes_centroids);
  if (rc == B2VS_OK) rc = read_section(f, &d->offsets, h.bytes_offsets);
  if (rc == B2VS_OK) rc = read_section(f, &d->sizes, h.bytes_sizes);
  if (rc == B2VS_OK) rc = read_section(f, &d->row_ids, h.bytes_row_ids);
  if (rc == B2VS_OK) rc = read_section(f, &d->data, h.bytes_data);
  if (rc == B2VS_OK) rc = read_section(f, &d->slot_norm, h.bytes_slot_norm);
  if (rc == B2VS_OK) rc = read_section(f, &d->codebooks, h.bytes_codebooks);
  if (rc == B2VS_OK) rc = read_section(f, &d->codes, h.bytes_codes);
  fclose(f);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (rc == B2VS_OK) {
    d->h_sizes.resize(h.n_lists);
    if (cudaMemcpy(d->h_sizes.data(), d->sizes.ptr, h.bytes_sizes, cudaMemcpyDeviceToHost) != cudaSuccess)
      rc = B2VS_ECUDA;
    if (rc == B2VS_OK) rc = build_list_ranks(d);
  }
  if (rc == B2VS_OK) {
    const int force = (h.dtype == B2VS_F32) ? -1 : h.fmt;
    rc = ix->flat.init(dev, h.metric, B2VS_F32, h.dim, d->centroids.ptr, h.n_lists, st, force);
  }
  if (rc == B2VS_OK && h.kind == B2VS_KIND_IVF_PQ) rc = pq_prepare_grouped(ix, d, st);
  if (rc == B2VS_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = B2VS_ECUDA;
  if (rc != B2VS_OK) {
    ivf_destroy(ix);
    ix->flat.destroy();
    delete ix;
    return rc;
  }
  *out = ix;
  return B2VS_OK;
}
