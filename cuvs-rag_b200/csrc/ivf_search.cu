// One IVF search batch (the work behind cuvs.neighbors.ivf_flat / ivf_pq .search at
// improved_multi_gpu_rag.py:225-233 and cuvs-2gpu-main.ipynb:L1801):
//   coarse probe (K4: the flat engine over the centroids, or K4b for a handful of queries)
//   -> per-query thresholds (seed) -> grouped tensor-core list scan (K5b / K7b) in append mode
//   -> per-query select (+ rescue of overflowed queries) -> merge / exact refine.
// Large k (128 < k <= 2048) takes the two-pass append + radix-select variants.
#include "ivf_internal.cuh"

namespace b2vs {

// ws_counter: [0] rows scanned counted per (query, probe), [1] appended candidates,
// [2] DISTINCT probed list rows (grouped scans), [3] large-k overflow flag (int)
constexpr size_t kCounterBytes = 4 * sizeof(unsigned long long);

// Candidate-buffer capacity (a power of two) and seed-sample length of the grouped scan, by k.
// B2VS_IVF_GROUPED_CAP shrinks the buffers so tests can drive the overflow-rescue path.
static int grouped_cap(int k) {
  if (env().grouped_cap > 0) return env().grouped_cap;
  (void)k;
  return 4096;   // 32 KB per query: the tail of the candidate-count distribution (5-10x the mean on
                 // structureless data) stays inside it; overflowing queries go to the sliced rescue
}
static uint32_t grouped_seed_rows(int k) {
  if (env().seed_rows > 0) return static_cast<uint32_t>(env().seed_rows);  // A/B knob
  return static_cast<uint32_t>(std::max(256, 16 * k));
}

// B2VS_IVF_GROUPED=0|1 forces the per-item / grouped IVF-Flat scan (A/B measurements, tests).
static int grouped_override() { return env().ivf_grouped; }

static int ivf_search_batch(b2vs_index* index, const void* q, int q_dtype, int nq, int k,
                            const b2vs_search_params& sp, float* out_d, int64_t* out_i,
                            cudaStream_t st);
static int ivf_pq_search_bigk(b2vs_index* index, const void* q, int q_dtype, int nq, int k,
                              const b2vs_search_params& sp, float* out_d, int64_t* out_i,
                              cudaStream_t st);

// Groups the nq * n_probes (query, probe) items by list and runs the tensor-core list scan in
// append mode against the thresholds in ws_g_tau; candidates land in ws_g_cand / ws_g_cnt (the
// caller sizes, zeroes and later selects from them).  probe_ids is [nq, n_probes] dense.
// phase: kPlanAndScan, or the two halves separately - kPlanOnly (sort, work table, gathered operand
// into the CURRENT planning buffers) and kScanOnly (the tensor-core kernel over buffers prepared by an
// earlier kPlanOnly call with the same arguments).
enum { kPlanAndScan = 0, kPlanOnly = 1, kScanOnly = 2 };
static int run_grouped_flat_scan(b2vs_index* index, IvfData* d, const long long* probe_ids,
                                 int n_probes, int nq, int cap, unsigned long long* counter,
                                 cudaStream_t st, bool timed = false, int row_limit = 0,
                                 int phase = kPlanAndScan) {
  const int items = nq * n_probes;
  const int q_split = index->dtype == B2VS_F32 ? 1 : 0;
  const int q_pitch = q_split ? 2 * static_cast<int>(round_up(d->dp, 64)) : d->dp;
  int chunk_rows = 0, slots = 1;
  choose_work_split(index, d, items, &chunk_rows, &slots);
  if (row_limit > 0) { chunk_rows = row_limit; slots = 1; }   // seed pass: one tile per list
  const int max_work = (items / kGroupRows + std::min(d->n_lists, items) + 1) * slots;
  const int64_t rows_cap = static_cast<int64_t>(sorted_rows_cap(d, items, kGroupRows));
  B2VS_TRY(d->ws_g_work.reserve(static_cast<size_t>(max_work) * sizeof(int4) + 16));
  B2VS_TRY(d->ws_g_q.reserve(static_cast<size_t>(rows_cap) * q_pitch * 2));
  B2VS_TRY(d->ws_g_rowq.reserve(static_cast<size_t>(rows_cap) * sizeof(int)));
  B2VS_TRY(d->ws_g_rowslot.reserve(static_cast<size_t>(rows_cap) * sizeof(int)));
  int* n_work = reinterpret_cast<int*>(d->ws_g_work.as<char>() + static_cast<size_t>(max_work) * sizeof(int4));
  if (phase != kScanOnly) {
    B2VS_TRY(plan_grouped_work(d, probe_ids, items, chunk_rows, slots, d->ws_g_work.as<int4>(), n_work,
                               counter, st, row_limit));
    B2VS_TRY(launch_gather_group_queries(d, rows_cap, n_probes, q_split, plan_is_small(d, items) ? 0 : items, st));
  }
  if (phase == kPlanOnly) return B2VS_OK;
  GroupedScanArgs ga{};
  ga.q_mat = d->ws_g_q.ptr; ga.q_rows = rows_cap;
  ga.x_mat = d->data.ptr; ga.x_rows = d->n_slots;
  ga.kdim = d->dp; ga.ab_format = d->fmt; ga.q_split = q_split;
  ga.beta = d->slot_norm.as<float>();
  ga.alpha = index->metric == B2VS_METRIC_L2 ? -2.f : -1.f;
  ga.work = d->ws_g_work.ptr; ga.n_work = n_work; ga.max_work = max_work;
  ga.row_query = d->ws_g_rowq.as<int>(); ga.tau = d->ws_g_tau.as<float>();
  ga.cand = d->ws_g_cand.as<u64>(); ga.count = d->ws_g_cnt.as<int>(); ga.cap = cap;
  ga.row_slot = d->ws_g_rowslot.as<int>(); ga.seed_all = row_limit > 0 ? 1 : 0;
  if (row_limit == 64 || row_limit == 128) ga.x_box_rows = row_limit;   // seed pass: load only the head rows
  // the counting-sort planner deals a block's rows over its leading ceil(w / 32) quarters (kDealPacked);
  // the one-CTA planner of small batches deals over all four
  ga.rows_in_work = plan_is_small(d, items) ? 0 : 1;
  if (timed) B2VS_CUDA(cudaEventRecord(d->ev0, st));
  B2VS_TRY(launch_grouped_scan(index->dev, ga, st));
  if (timed) B2VS_CUDA(cudaEventRecord(d->ev1, st));
  return B2VS_OK;
}

// Thresholds from a tensor-core SEED PASS (batches of kTcSeedMinQueries queries and more): the
// first kSeedTileRows rows of each query's m nearest lists (m = 2 for k <= 32, else 4) are scored by
// the same grouped kernel with no threshold - every score lands in a FIXED slot of the query's
// buffer, no atomics - and the k-th best of them becomes the query's threshold for the full pass.  Against the CUDA-core seed kernels (one CTA per query over
// the nearest list only) the sample is m times larger - on corpora without cluster structure the
// nearest list's head alone leaves thousands of candidates per query, overflowing the buffers
// into the exact rescue scan - it runs on the tensor cores, and it uses the full pass's own
// arithmetic, so the threshold carries no rounding cushion.
// B2VS_IVF_SEED_LISTS / B2VS_IVF_SEED_TILE: measurement knobs (lists per query, rows per list).
static int seed_tile_rows() {
  const int v = env().seed_tile;
  return v > 0 ? v : kSeedTileRows;
}
static int seed_lists(int n_probes, int cap, int k) {
  // two lists (512 sampled rows) for k <= 32, four for larger k; never more than a quarter of the buffer
  const int want = env().seed_lists > 0 ? env().seed_lists : (k <= 32 ? 2 : 4);
  return std::max(1, std::min(std::min(n_probes, want), cap / 4 / kSeedTileRows));
}
static bool use_tc_seed(int nq, int cap) {
  if (cap < 4 * kSeedTileRows) return false;   // shrunken test buffers cannot hold a seed tile
  const int m = env().seed_mode;
  return m >= 0 ? m == 1 : nq >= kTcSeedMinQueries;
}
template <typename ScanFn>
static int run_tc_seed(IvfData* d, const long long* probe_ids, int n_probes, int nq, int k, int cap,
                       cudaStream_t st, ScanFn scan) {
  const int m = seed_lists(n_probes, cap, k);
  B2VS_TRY(d->ws_seed_ids.reserve(static_cast<size_t>(nq) * m * sizeof(int64_t)));
  B2VS_CUDA(cudaMemcpy2DAsync(d->ws_seed_ids.ptr, static_cast<size_t>(m) * sizeof(int64_t), probe_ids,
                              static_cast<size_t>(n_probes) * sizeof(int64_t),
                              static_cast<size_t>(m) * sizeof(int64_t), nq, cudaMemcpyDeviceToDevice, st));
  // fixed slots: [query][seed list j][row of the list's first tile]; short lists leave kKeyInf
  B2VS_CUDA(cudaMemset2DAsync(d->ws_g_cand.ptr, static_cast<size_t>(cap) * sizeof(u64), 0xFF,
                              static_cast<size_t>(m) * kSeedTileRows * sizeof(u64), nq, st));
  B2VS_TRY(scan(reinterpret_cast<const long long*>(d->ws_seed_ids.ptr), m));
  B2VS_TRY(launch_seed_select(d, nq, m * kSeedTileRows, cap, k, st));
  return B2VS_OK;
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
  B2VS_TRY(d->ws_counter.reserve(kCounterBytes));
  B2VS_TRY(d->ws_g_tau.reserve(static_cast<size_t>(nq) * sizeof(float)));
  B2VS_TRY(d->ws_g_cand.reserve(static_cast<size_t>(nq) * kCapBig * sizeof(u64)));
  B2VS_TRY(d->ws_g_cnt.reserve(static_cast<size_t>(nq) * sizeof(int)));
  B2VS_TRY(reserve_item_sort(d, nq * n_probes, kGroupRows));
  B2VS_TRY(index->flat.search(q, q_dtype, nq, n_probes, 0, 0, d->ws_probe_d.as<float>(),
                              d->ws_probe_i.as<int64_t>(), nullptr, st));
  int launches = index->flat.stats.launches;
  const int round16 = index->dtype != B2VS_F32 ? 1 : 0;
  B2VS_TRY(launch_queries_to_f32(q, q_dtype, nq, index->dim, d->dp, d->fmt, round16,
                                 d->ws_qf.as<float>(), d->ws_qnorm.as<float>(), st));
  unsigned long long* counter = d->ws_counter.as<unsigned long long>();
  int* overflow = reinterpret_cast<int*>(counter + 3);
  B2VS_CUDA(cudaMemsetAsync(counter, 0, kCounterBytes, st));
  const long long* probe_ids = reinterpret_cast<const long long*>(d->ws_probe_i.ptr);
  // ---- pass 1: nearest m lists, no threshold
  B2VS_CUDA(cudaMemcpy2DAsync(d->ws_ref_i.ptr, static_cast<size_t>(m) * sizeof(int64_t), probe_ids,
                              static_cast<size_t>(n_probes) * sizeof(int64_t),
                              static_cast<size_t>(m) * sizeof(int64_t), nq, cudaMemcpyDeviceToDevice, st));
  B2VS_TRY(launch_fill_f32(d->ws_g_tau.as<float>(), static_cast<size_t>(nq), INFINITY, st));
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
constexpr size_t kMaxWorkspaceBytes = 6ull << 30;   // search workspaces of one sub-batch (grow-only)

// large k: k itself, or (IVF-PQ) the number of ADC candidates kept for the exact re-rank
bool uses_bigk_path(const b2vs_index* index, const IvfData* d, int k, const b2vs_search_params& sp) {
  const bool pq = index->kind == B2VS_KIND_IVF_PQ;
  const bool pq_refine = pq && sp.refine_ratio > 1 && d->src_rows != nullptr;
  return k > kMaxFusedK ||
         (pq_refine && d->pq_tc_ready && static_cast<int64_t>(k) * sp.refine_ratio > kMaxFusedK);
}

int ivf_search_direct(b2vs_index* index, const void* q, int q_dtype, int nq, int k,
                             const b2vs_search_params& sp, float* out_d, int64_t* out_i,
                             cudaStream_t st) {
  IvfData* d = static_cast<IvfData*>(index->ivf);
  B2VS_CHECK(d != nullptr, B2VS_EINVAL, "IVF index has no list data");
  int n_probes = sp.n_probes > 0 ? sp.n_probes : 20;
  n_probes = std::max(1, std::min(n_probes, std::min(d->n_lists, kMaxProbes)));
  int chunk = static_cast<int>(std::max<int64_t>(1024, kMaxItemsPerBatch / n_probes));
  const bool pq = index->kind == B2VS_KIND_IVF_PQ;
  // refine re-ranks against the caller's rows: without them (index loaded with
  // rows_for_refine = NULL) the request cannot be honoured - say so instead of silently
  // returning un-refined ADC results
  B2VS_CHECK(!(pq && sp.refine_ratio > 1 && d->src_rows == nullptr), B2VS_EINVAL,
             "refine_ratio=%d needs the shard's source rows: this IVF-PQ index was loaded without "
             "rows_for_refine", sp.refine_ratio);
  const bool bigk = uses_bigk_path(index, d, k, sp);
  if (bigk) {
    B2VS_CHECK(k <= kMaxBigK, B2VS_EUNSUP, "k=%d exceeds the large-k limit %d", k, kMaxBigK);
    chunk = std::min(chunk, 4096);   // 64 K-key candidate buffer per query
  } else {
    // the grouped scans hold, per query: a candidate buffer (cap keys), its gathered operand rows
    // (one per probe, + group padding) and the sort tables: bound the sub-batch by BYTES as well,
    // so few-probe searches of huge batches do not ask for tens of GB of workspace
    const int k_scan = (pq && sp.refine_ratio > 1) ? std::min(kMaxFusedK, k * sp.refine_ratio) : k;
    const size_t op_row = static_cast<size_t>(pq ? index->dim : d->dp) * 2 * (index->dtype == B2VS_F32 && !pq ? 2 : 1);
    const size_t per_query = static_cast<size_t>(grouped_cap(k_scan)) * sizeof(u64) +
                             2 * static_cast<size_t>(n_probes) * (op_row + 40) +   // two sets of planning buffers

                             static_cast<size_t>(d->dp) * sizeof(float) + static_cast<size_t>(k_scan) * 32;
    const int64_t by_bytes = static_cast<int64_t>(kMaxWorkspaceBytes / per_query);
    chunk = static_cast<int>(std::min<int64_t>(chunk, std::max<int64_t>(1024, by_bytes)));
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
                               int nq, int cap, unsigned long long* counter, cudaStream_t st,
                               bool timed = false, int row_limit = 0, int phase = kPlanAndScan) {
  const int items = nq * n_probes;
  const int l2 = index->metric == B2VS_METRIC_L2 ? 1 : 0;
  int chunk_rows = 0, slots = 1;
  choose_work_split(index, d, items, &chunk_rows, &slots);
  if (row_limit > 0) { chunk_rows = row_limit; slots = 1; }   // seed pass: one tile per list
  const int max_work = (items / kGroupRows + std::min(d->n_lists, items) + 1) * slots;
  const int64_t rows_cap = static_cast<int64_t>(sorted_rows_cap(d, items, kGroupRows));
  B2VS_TRY(d->ws_g_work.reserve(static_cast<size_t>(max_work) * sizeof(int4) + 16));
  B2VS_TRY(d->ws_g_q.reserve(static_cast<size_t>(rows_cap) * index->dim * 2));
  B2VS_TRY(d->ws_g_rowq.reserve(static_cast<size_t>(rows_cap) * sizeof(int)));
  B2VS_TRY(d->ws_g_rowslot.reserve(static_cast<size_t>(rows_cap) * sizeof(int)));
  B2VS_TRY(d->ws_g_bias.reserve(static_cast<size_t>(rows_cap) * sizeof(float)));
  int* n_work = reinterpret_cast<int*>(d->ws_g_work.as<char>() + static_cast<size_t>(max_work) * sizeof(int4));
  if (phase != kScanOnly) {
    B2VS_TRY(plan_grouped_work(d, probe_ids, items, chunk_rows, slots, d->ws_g_work.as<int4>(), n_work,
                               counter, st, row_limit, kDealNone));
    B2VS_TRY(launch_gather_group_residuals(index, d, rows_cap, probe_ids, n_probes,
                                           plan_is_small(d, items) ? 0 : items, st));
  }
  if (phase == kPlanOnly) return B2VS_OK;
  PqGroupedScanArgs ga{};
  ga.q_mat = d->ws_g_q.ptr; ga.q_rows = rows_cap;
  ga.dim = index->dim; ga.pq_dim = d->pq_dim; ga.dsub = d->dsub;
  ga.codes = d->codes.ptr;
  ga.n_groups = static_cast<uint32_t>(std::max<int64_t>(d->n_slots, 32) >> 5);
  ga.cb16t = d->cb16t.ptr;
  ga.beta = d->pq_norm.as<float>(); ga.alpha = l2 ? -2.f : -1.f;
  ga.work = d->ws_g_work.ptr; ga.n_work = n_work; ga.max_work = max_work;
  ga.row_query = d->ws_g_rowq.as<int>(); ga.row_bias = d->ws_g_bias.as<float>();
  ga.tau = d->ws_g_tau.as<float>();
  ga.cand = d->ws_g_cand.as<u64>(); ga.count = d->ws_g_cnt.as<int>(); ga.cap = cap;
  ga.row_slot = d->ws_g_rowslot.as<int>(); ga.seed_all = row_limit > 0 ? 1 : 0;
  if (timed) B2VS_CUDA(cudaEventRecord(d->ev0, st));
  B2VS_TRY(launch_pq_grouped_scan(index->dev, ga, st));
  if (timed) B2VS_CUDA(cudaEventRecord(d->ev1, st));
  return B2VS_OK;
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
  B2VS_TRY(d->ws_counter.reserve(kCounterBytes));
  B2VS_TRY(d->ws_g_tau.reserve(static_cast<size_t>(nq) * sizeof(float)));
  B2VS_TRY(d->ws_g_cand.reserve(static_cast<size_t>(nq) * kCapBig * sizeof(u64)));
  B2VS_TRY(d->ws_g_cnt.reserve(static_cast<size_t>(nq) * sizeof(int)));
  B2VS_TRY(reserve_item_sort(d, nq * n_probes, kGroupRows));
  B2VS_TRY(index->flat.search(q, q_dtype, nq, n_probes, 0, 0, d->ws_probe_d.as<float>(),
                              d->ws_probe_i.as<int64_t>(), nullptr, st));
  int launches = index->flat.stats.launches;
  B2VS_TRY(launch_queries_to_f32(q, q_dtype, nq, index->dim, d->dp, d->fmt, 0,
                                 d->ws_qf.as<float>(), d->ws_qnorm.as<float>(), st));
  unsigned long long* counter = d->ws_counter.as<unsigned long long>();
  int* overflow = reinterpret_cast<int*>(counter + 3);
  B2VS_CUDA(cudaMemsetAsync(counter, 0, kCounterBytes, st));
  const long long* probe_ids = reinterpret_cast<const long long*>(d->ws_probe_i.ptr);
  B2VS_CUDA(cudaMemcpy2DAsync(d->ws_keys.ptr, static_cast<size_t>(m) * sizeof(int64_t), probe_ids,
                              static_cast<size_t>(n_probes) * sizeof(int64_t),
                              static_cast<size_t>(m) * sizeof(int64_t), nq, cudaMemcpyDeviceToDevice, st));
  B2VS_TRY(launch_fill_f32(d->ws_g_tau.as<float>(), static_cast<size_t>(nq), INFINITY, st));
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
    B2VS_TRY(launch_refine_big(index, d, reinterpret_cast<const long long*>(d->ws_ref_i.ptr), nq, k_scan, k,
                               out_d, out_i, st));
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

// Number of pseudo-lists, or 0 when the tensor-core probe should run (B2VS_COARSE_SCAN=0 forces that).
static int coarse_scan_chunks(const b2vs_index* index, const IvfData* d, int nq, int n_probes) {
  const EnvConfig& e = env();
  if (e.coarse_scan == 0) return 0;
  const int maxq = e.coarse_scan_maxq < 0 ? kCoarseScanMaxQueries
                                          : std::max(1, std::min(e.coarse_scan_maxq, kCoarseScanProbeRows));
  if (nq > maxq || n_probes > kMaxFusedK) return 0;
  if (index->flat.split3 || index->flat.kdim != d->dp || (d->dp >> 3) > 32 * 8) return 0;
  const int chunks = static_cast<int>(ceil_div(d->n_lists, kCoarseScanRows));
  const int per_sm = e.coarse_scan_ctas < 0 ? kCoarseScanCtasPerSm : std::max(1, std::min(e.coarse_scan_ctas, 64));
  const int64_t max_ctas = static_cast<int64_t>(per_sm) * sm_count(index->dev);
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
  const float alpha = index->flat.metric == B2VS_METRIC_L2 ? -2.f : -1.f;
  B2VS_TRY(launch_flat_item_scan(index->flat.ab_format, d->dp, nq * chunks,
                                 static_cast<const uint16_t*>(index->flat.mat), index->flat.beta.as<float>(),
                                 d->cq_offsets.as<uint32_t>(), d->cq_probe.as<long long>(),
                                 d->ws_qf.as<float>(), chunks, nq, n_probes, alpha, d->ws_cq_keys.as<u64>(),
                                 nullptr, nullptr, st));
  return launch_merge_splits(d->ws_cq_keys.as<u64>(), chunks, nq, nq, n_probes, index->flat.metric,
                             d->ws_qnorm.as<float>(), 0, d->ws_probe_d.as<float>(),
                             d->ws_probe_i.as<int64_t>(), nullptr, st);
}

// Large batches: the main pass's planning + gather (0.14-0.2 ms at C3 / C4) depends only on the
// coarse probe, so it runs on a side stream - into the second set of planning buffers - while the
// seed pass (0.3-0.4 ms) runs on the caller's stream.  B2VS_PLAN_OVERLAP=0 keeps everything in line.
static bool plan_overlap_enabled(int nq) { return env().plan_overlap != 0 && nq >= kTcSeedMinQueries; }
static int plan_fork(IvfData* d, cudaStream_t st) {
  if (!d->side) {
    B2VS_CUDA(cudaStreamCreateWithFlags(&d->side, cudaStreamNonBlocking));
    B2VS_CUDA(cudaEventCreateWithFlags(&d->ev_fork, cudaEventDisableTiming));
    B2VS_CUDA(cudaEventCreateWithFlags(&d->ev_join, cudaEventDisableTiming));
  }
  B2VS_CUDA(cudaEventRecord(d->ev_fork, st));
  B2VS_CUDA(cudaStreamWaitEvent(d->side, d->ev_fork, 0));
  d->swap_plan_ws();      // the side stream's launches fill the other set
  return B2VS_OK;
}
static int plan_fork_done(IvfData* d) {
  B2VS_CUDA(cudaEventRecord(d->ev_join, d->side));
  d->swap_plan_ws();      // back: the seed pass plans into the first set
  return B2VS_OK;
}
static int plan_join(IvfData* d, cudaStream_t st) {
  B2VS_CUDA(cudaStreamWaitEvent(st, d->ev_join, 0));
  d->swap_plan_ws();      // the main scan reads what the side stream prepared
  return B2VS_OK;
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
  // the grouped scans leave ONE sorted list per query; the per-(query, probe) kernels one per probe
  const bool will_group = grouped_override() != 0 &&
                          (index->kind == B2VS_KIND_IVF_FLAT || d->pq_tc_ready);
  B2VS_TRY(d->ws_keys.reserve(static_cast<size_t>(will_group ? 1 : n_probes) * q_pad * k * sizeof(u64)));
  B2VS_TRY(d->ws_qf.reserve(static_cast<size_t>(nq) * d->dp * sizeof(float)));
  B2VS_TRY(d->ws_qnorm.reserve(static_cast<size_t>(q_pad) * sizeof(float)));
  B2VS_TRY(d->ws_counter.reserve(kCounterBytes));
  const int round16 = (index->kind == B2VS_KIND_IVF_FLAT && index->dtype != B2VS_F32) ? 1 : 0;
  B2VS_TRY(launch_queries_to_f32(q, q_dtype, nq, index->dim, d->dp, d->fmt, round16,
                                 d->ws_qf.as<float>(), d->ws_qnorm.as<float>(), st));
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
  B2VS_CUDA(cudaMemsetAsync(d->ws_counter.ptr, 0, kCounterBytes, st));
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
  }
  bool scan_timed_inside = false;   // grouped paths time the tensor-core scan kernel alone
  bool single_list = false;  // the scan left ONE sorted list per query (not one per probe)
  if (index->kind == B2VS_KIND_IVF_FLAT) {
    const float alpha = index->metric == B2VS_METRIC_L2 ? -2.f : -1.f;
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
      const bool tc_seed = use_tc_seed(nq, cap);
      const bool order_seeds = !tc_seed && nq >= kSeedSortMinQueries;
      if (order_seeds) B2VS_TRY(sort_items_by_list(d, probe_ids, nq, n_probes, 1, st));
      B2VS_TRY(d->ws_g_tau.reserve(static_cast<size_t>(nq) * sizeof(float)));
      B2VS_TRY(d->ws_g_cand.reserve(static_cast<size_t>(nq) * cap * sizeof(u64)));
      B2VS_TRY(d->ws_g_cnt.reserve(static_cast<size_t>(nq) * sizeof(int)));
      B2VS_CUDA(cudaMemsetAsync(d->ws_g_cnt.ptr, 0, static_cast<size_t>(nq) * sizeof(int), st));
      const bool overlap = tc_seed && plan_overlap_enabled(nq);
      if (overlap) {
        B2VS_TRY(plan_fork(d, st));
        B2VS_TRY(run_grouped_flat_scan(index, d, probe_ids, n_probes, nq, cap, counter, d->side, false, 0, kPlanOnly));
        B2VS_TRY(plan_fork_done(d));
      }
      if (tc_seed) {
        B2VS_TRY(run_tc_seed(d, probe_ids, n_probes, nq, k, cap, st, [&](const long long* ids, int m) {
          return run_grouped_flat_scan(index, d, ids, m, nq, cap, nullptr, st, false, seed_tile_rows());
        }));
        launches += 8;
      } else {
        B2VS_TRY(launch_flat_seed_tau(index, d, probe_ids, n_probes, nq, k, grouped_seed_rows(k),
                                      q_split ? 1e-5f : 0.f,
                                      order_seeds ? d->ws_item_perm.as<uint32_t>() : nullptr, st));
      }
      if (overlap) B2VS_TRY(plan_join(d, st));
      B2VS_TRY(run_grouped_flat_scan(index, d, probe_ids, n_probes, nq, cap, counter, st, timed, 0,
                                     overlap ? kScanOnly : kPlanAndScan));
      scan_timed_inside = true;
      B2VS_TRY(launch_group_select(d, nq, cap, k, counter + 1, st));
      B2VS_TRY(launch_flat_rescue(index, d, probe_ids, n_probes, nq, k, cap, st));
      launches += 16;
      single_list = true;
    } else {
      // Per-item scan.  Ordering the items by list pays once several queries share a list: the
      // batch then reads each probed list from HBM about once instead of once per probing query.
      const uint32_t* item_perm = nullptr;
      if (timed) B2VS_CUDA(cudaEventRecord(d->ev0, st));
      if (!env().no_item_sort && items >= 4 * d->n_lists) {
        B2VS_TRY(sort_items_by_list(d, probe_ids, items, 1, 1, st));
        item_perm = d->ws_item_perm.as<uint32_t>();
        launches += 5;
      }
      B2VS_TRY(launch_flat_item_scan(d->fmt, d->dp, items, d->data.as<uint16_t>(),
                                     d->slot_norm.as<float>(), d->offsets.as<uint32_t>(), probe_ids,
                                     d->ws_qf.as<float>(), n_probes, q_pad, k, alpha, d->ws_keys.as<u64>(),
                                     counter, item_perm, st));
    }
    qnorm_for_merge = d->ws_qnorm.as<float>();
  } else if (d->pq_tc_ready && grouped_override() != 0) {
    // Grouped tensor-core scan (pq_tc.cuh): same pipeline as the IVF-Flat one, the list tiles are
    // decoded from the PQ codes instead of loaded.
    const int cap = grouped_cap(k);
    B2VS_TRY(reserve_item_sort(d, items, kGroupRows));
    const bool tc_seed = use_tc_seed(nq, cap);
    const bool order_seeds = !tc_seed && nq >= kSeedSortMinQueries;
    if (order_seeds) B2VS_TRY(sort_items_by_list(d, probe_ids, nq, n_probes, 1, st));
    B2VS_TRY(d->ws_g_tau.reserve(static_cast<size_t>(nq) * sizeof(float)));
    B2VS_TRY(d->ws_g_cand.reserve(static_cast<size_t>(nq) * cap * sizeof(u64)));
    B2VS_TRY(d->ws_g_cnt.reserve(static_cast<size_t>(nq) * sizeof(int)));
    B2VS_CUDA(cudaMemsetAsync(d->ws_g_cnt.ptr, 0, static_cast<size_t>(nq) * sizeof(int), st));
    const bool overlap = tc_seed && plan_overlap_enabled(nq);
    if (overlap) {
      B2VS_TRY(plan_fork(d, st));
      B2VS_TRY(run_grouped_pq_scan(index, d, probe_ids, n_probes, nq, cap, counter, d->side, false, 0, kPlanOnly));
      B2VS_TRY(plan_fork_done(d));
    }
    if (tc_seed) {
      B2VS_TRY(run_tc_seed(d, probe_ids, n_probes, nq, k, cap, st, [&](const long long* ids, int m) {
        return run_grouped_pq_scan(index, d, ids, m, nq, cap, nullptr, st, false, seed_tile_rows());
      }));
      launches += 8;
    } else {
      B2VS_TRY(launch_pq_lut_scan(0, index, d, probe_ids, n_probes, nq, k, cap, grouped_seed_rows(k),
                                  order_seeds ? d->ws_item_perm.as<uint32_t>() : nullptr, st));
    }
    if (overlap) B2VS_TRY(plan_join(d, st));
    B2VS_TRY(run_grouped_pq_scan(index, d, probe_ids, n_probes, nq, cap, counter, st, timed, 0,
                                 overlap ? kScanOnly : kPlanAndScan));
    scan_timed_inside = true;
    B2VS_TRY(launch_group_select(d, nq, cap, k, counter + 1, st));
    B2VS_TRY(launch_pq_lut_scan(1, index, d, probe_ids, n_probes, nq, k, cap, 0u, nullptr, st));
    launches += 16;
    single_list = true;
  } else {
    if (timed) B2VS_CUDA(cudaEventRecord(d->ev0, st));
    B2VS_TRY(launch_pq_table_scan(index, d, probe_ids, n_probes, nq, q_pad, k, counter, &single_list, st));
  }
  if (timed && !scan_timed_inside) B2VS_CUDA(cudaEventRecord(d->ev1, st));
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
    B2VS_TRY(launch_refine(index, d, reinterpret_cast<const long long*>(d->ws_ref_i.ptr), nq, k, k_final,
                           out_d, out_i, st));
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


}  // namespace b2vs
