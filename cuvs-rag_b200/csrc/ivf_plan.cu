// Planning of the grouped list scans: the (query, probe) items of a batch are counting-sorted by
// the list they probe (the K3 kernels of kmeans.cu), every list's group of queries is cut into
// 128-row blocks, and one work item = (query block, row range of the list) is emitted for the
// tensor-core scan kernels (bf_tc_kernel<1, true> / pq_tc_kernel).  Small batches are planned by
// ONE CTA (the Q <= 64 path is launch-bound).
#include "ivf_internal.cuh"

namespace b2vs {

// One thread per list: emits the list's work items.  group_off = exclusive scan (in size-rank
// order) of the per-list query counts rounded up to 128, in gathered-row units.
__global__ void build_group_work_kernel(const uint32_t* __restrict__ group_off,
                                        const uint32_t* __restrict__ offsets,
                                        const int* __restrict__ group_cnt,
                                        const int* __restrict__ list_of_rank, int n_lists,
                                        int chunk_rows, int slots, int row_limit,
                                        int4* __restrict__ work, int* __restrict__ n_work,
                                        unsigned long long* __restrict__ scanned_rows) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;   // size rank: items come out longest first
  if (r == 0) *n_work = static_cast<int>(group_off[n_lists] >> 7) * slots;
  unsigned long long scanned = 0, distinct = 0;
  if (r < n_lists) {
    const int l = list_of_rank[r];
    const int b0 = static_cast<int>(group_off[r] >> 7), b1 = static_cast<int>(group_off[r + 1] >> 7);
    const int begin = static_cast<int>(offsets[l]);
    int end = static_cast<int>(offsets[l + 1]);
    if (row_limit > 0) end = min(end, begin + row_limit);   // seed pass: the head of every list only
    // Small batches have fewer (list, query block) pairs than SMs: a list is then cut into `slots`
    // row ranges of chunk_rows (a multiple of the 256-row tile), one work item each - append mode
    // keeps no per-item state, so the pieces are independent.  Ranges past the list end are empty.
    for (int b = b0; b < b1; ++b)
      for (int c = 0; c < slots; ++c) {
        const int rb = min(end, begin + c * chunk_rows);
        const int re = c + 1 == slots ? end : min(end, rb + chunk_rows);
        // .w = real query rows of the block (the PQ scan sizes its UMMA N extent and loads by it)
        work[b * slots + c] = make_int4(b, rb, re, min(kGroupRows, group_cnt[r] - (b - b0) * kGroupRows));
      }
    if (group_cnt[r] > 0) {  // algorithmic work: every probing query sees every row
      scanned = static_cast<unsigned long long>(group_cnt[r]) * static_cast<unsigned>(end - begin);
      distinct = static_cast<unsigned long long>(end - begin);   // distinct list rows
    }
  }
  if (scanned_rows) {      // one pair of atomics per warp (16 K lists hammered two counters before)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      scanned += __shfl_xor_sync(0xffffffffu, scanned, o);
      distinct += __shfl_xor_sync(0xffffffffu, distinct, o);
    }
    if ((threadIdx.x & 31) == 0 && scanned) {
      atomicAdd(scanned_rows, scanned);
      atomicAdd(scanned_rows + 2, distinct);
    }
  }
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

// Counting sort of items by the size rank of their list (the same three kernels that build the
// lists).  Item i probes list probe_ids[i * stride].  group_pad = 1: dense permutation (row_item[i] = i-th item in list order);
// group_pad = 128: every list's group starts on a 128-row boundary, holes hold kNoRow.
size_t sorted_rows_cap(const IvfData* d, int items, int group_pad) {
  return group_pad == 1 ? static_cast<size_t>(items)
                        : static_cast<size_t>(items) +
                              static_cast<size_t>(group_pad) * std::min(d->n_lists, items);
}

int reserve_item_sort(IvfData* d, int items, int group_pad) {
  B2VS_TRY(d->ws_item_lab.reserve(static_cast<size_t>(items) * sizeof(int)));
  B2VS_TRY(d->ws_item_cnt.reserve(static_cast<size_t>(d->n_lists) * 2 * sizeof(int)));
  B2VS_TRY(d->ws_item_off.reserve((static_cast<size_t>(d->n_lists) + 1) * sizeof(uint32_t)));
  B2VS_TRY(d->ws_item_perm.reserve(sorted_rows_cap(d, items, group_pad) * sizeof(uint32_t)));
  B2VS_TRY(d->ws_item_slot.reserve(static_cast<size_t>(items) * sizeof(uint32_t)));
  return B2VS_OK;
}

int sort_items_by_list(IvfData* d, const long long* probe_ids, int items, int stride,
                              int group_pad, cudaStream_t st, int deal) {
  const size_t rows_cap = sorted_rows_cap(d, items, group_pad);
  B2VS_TRY(reserve_item_sort(d, items, group_pad));
  const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(ceil_div(items, 256), 2048));
  int* cnt = d->ws_item_cnt.as<int>();
  B2VS_CUDA(cudaMemsetAsync(cnt, 0, static_cast<size_t>(d->n_lists) * 2 * sizeof(int), st));
  if (group_pad > 1) B2VS_CUDA(cudaMemsetAsync(d->ws_item_perm.ptr, 0xFF, rows_cap * sizeof(uint32_t), st));
  probe_labels_kernel<<<blocks, 256, 0, st>>>(probe_ids, items, stride, d->rank_of_list.as<int>(),
                                              d->ws_item_lab.as<int>());
  B2VS_CUDA(cudaGetLastError());
  B2VS_TRY(launch_histogram(d->ws_item_lab.as<int>(), items, cnt, static_cast<int>(blocks), st));
  B2VS_TRY(launch_scan_sizes(cnt, d->n_lists, group_pad, d->ws_item_off.as<uint32_t>(), st));
  return launch_scatter_rows(d->ws_item_lab.as<int>(), items, d->ws_item_off.as<uint32_t>(),
                             cnt + d->n_lists, d->ws_item_perm.as<uint32_t>(),
                             d->ws_item_slot.as<uint32_t>(), static_cast<int>(blocks), st,
                             (group_pad == kGroupRows && deal != kDealNone) ? cnt : nullptr,
                             deal == kDealFour ? 1 : 0);
}

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
                      int row_limit, int deal_none, uint32_t* __restrict__ row_item, uint32_t* __restrict__ group_off,
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
    int rows = static_cast<int>(offsets[l + 1] - offsets[l]);
    if (row_limit > 0) rows = min(rows, row_limit);
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
  unsigned long long rows_scanned = 0, rows_distinct = 0;
  for (int i = lo; i < hi; ++i) {
    off[i] = run;
    group_off[i] = run;
    const int c = cnt[i];
    if (c == 0) continue;
    const int l = list_of_rank[i];
    const int begin = static_cast<int>(offsets[l]);
    int end = static_cast<int>(offsets[l + 1]);
    if (row_limit > 0) end = min(end, begin + row_limit);
    const int blocks = (c + kGroupRows - 1) / kGroupRows;
    const int b0 = static_cast<int>(run >> 7);
    for (int b = 0; b < blocks; ++b)
      for (int rb = begin; rb < end; rb += chunk_rows)
        work[wrun++] = make_int4(b0 + b, rb, min(end, rb + chunk_rows), min(kGroupRows, c - b * kGroupRows));
    rows_scanned += static_cast<unsigned long long>(c) * static_cast<unsigned>(end - begin);
    rows_distinct += static_cast<unsigned>(end - begin);
    run += static_cast<uint32_t>(blocks * kGroupRows);
  }
  if (scanned_rows && rows_scanned) {
    atomicAdd(scanned_rows, rows_scanned);
    atomicAdd(scanned_rows + 2, rows_distinct);
  }
  __syncthreads();
  for (int i = t; i < n_lists; i += kPlanThreads) cnt[i] = 0;   // now the scatter cursors
  __syncthreads();
  for (int i = t; i < items; i += kPlanThreads) {
    const long long l = probe_ids[i];
    const int r = rank_of_list[l < 0 ? 0 : l];
    const uint32_t idx = static_cast<uint32_t>(atomicAdd(&cnt[r], 1));
    row_item[off[r] + (deal_none ? idx : group_row_pos(idx, 0u))] = static_cast<uint32_t>(i);
  }
}

bool plan_is_small(const IvfData* d, int items) {
  return items <= kPlanMaxItems && d->n_lists <= kPlanMaxLists;
}

// Sort + work table of the grouped scans: the one-CTA plan for small batches, else the
// counting-sort kernels + build_group_work_kernel.
int plan_grouped_work(IvfData* d, const long long* probe_ids, int items, int chunk_rows,
                             int slots, int4* work, int* n_work, unsigned long long* counter,
                             cudaStream_t st, int row_limit, int deal) {
  if (plan_is_small(d, items)) {
    B2VS_TRY(reserve_item_sort(d, items, kGroupRows));
    B2VS_CUDA(cudaMemsetAsync(d->ws_item_perm.ptr, 0xFF,
                              sorted_rows_cap(d, items, kGroupRows) * sizeof(uint32_t), st));
    const size_t smem = (2 * static_cast<size_t>(d->n_lists) + 1) * sizeof(int);
    B2VS_CUDA(cudaFuncSetAttribute(ivf_plan_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem)));
    ivf_plan_small_kernel<<<1, kPlanThreads, smem, st>>>(
        probe_ids, items, d->rank_of_list.as<int>(), d->list_of_rank.as<int>(),
        d->offsets.as<uint32_t>(), d->n_lists, chunk_rows, slots, row_limit, deal == kDealNone ? 1 : 0,
        d->ws_item_perm.as<uint32_t>(),
        d->ws_item_off.as<uint32_t>(), work, n_work, counter);
    B2VS_CUDA(cudaGetLastError());
    return B2VS_OK;
  }
  B2VS_TRY(sort_items_by_list(d, probe_ids, items, 1, kGroupRows, st, deal));
  build_group_work_kernel<<<static_cast<unsigned>(ceil_div(d->n_lists, 256)), 256, 0, st>>>(
      d->ws_item_off.as<uint32_t>(), d->offsets.as<uint32_t>(), d->ws_item_cnt.as<int>(),
      d->list_of_rank.as<int>(), d->n_lists, chunk_rows, slots, row_limit, work, n_work, counter);
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}

// Row-range split of the grouped scan's work items (see build_group_work_kernel): large batches
// aim at two items per SM when the batch alone does not provide them.
void choose_work_split(const b2vs_index* index, const IvfData* d, int items, int* chunk_rows,
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
  if (plan_is_small(d, items)) {
    chunk_tiles = static_cast<int>(est_tiles / (16.0 * sms));
    chunk_tiles = std::max(1, std::min(chunk_tiles, std::min(4, max_tiles)));
  }
  // A/B switches: B2VS_WORK_CHUNK_TILES=n forces the chunk, B2VS_DEBUG_SPLIT prints it
  if (env().work_chunk_tiles > 0) chunk_tiles = std::min(env().work_chunk_tiles, max_tiles);
  *slots = static_cast<int>(ceil_div(max_tiles, chunk_tiles));
  *chunk_rows = chunk_tiles * 256;
  if (env().debug_split)
    std::fprintf(stderr, "[b2vs] work split: items=%d mean_rows=%.0f max_tiles=%d chunk_tiles=%d slots=%d\n",
                 items, mean_rows, max_tiles, chunk_tiles, *slots);
}


}  // namespace b2vs
