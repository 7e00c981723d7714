// K8 — top-k merges.
//  * merge_splits_kernel: folds the per-(split, query) sorted key lists produced by the fused
//    distance kernels into the final (distance, id) rows.
//  * merge_parts_kernel (b2vs_merge_topk): the cross-shard global top-k that replaces the
//    reference's host-side concatenate + np.argsort(...)[:k]
//    (improved_multi_gpu_rag.py:266-275, cuvs-2gpu-main.ipynb:L1806-1832,
//    test_search_result_aggregator.py:308-358 for the known answers).
// Both run one warp per query and keep a sorted 128-entry register list (4 keys per lane);
// a new sorted run is folded in with the classic "min(a[i], b[n-1-i]) is bitonic" step plus one
// bitonic merge pass.
#include <cmath>

#include "common.h"
#include "warp_select.cuh"

namespace b2vs {

constexpr int kMergeE = 4;  // 32 * 4 = 128 = kMaxFusedK
constexpr int kMergeAhead = 4;  // split lists in flight per warp in merge_splits_kernel

// The sorted keys of one query (4 per lane) -> its row of the answer: distances in the caller's
// metric, ids shifted into the global numbering (or remapped through the list layout).
__device__ __forceinline__ void emit_answer_row(const u64 (&acc)[kMergeE], int lane, int q, int k, int metric,
                                                const float* __restrict__ qnorm, long long id_offset,
                                                float* __restrict__ out_d, long long* __restrict__ out_i,
                                                int* __restrict__ out_label,
                                                const uint32_t* __restrict__ remap) {
  const float qn = (metric == B2VS_METRIC_L2 && qnorm) ? qnorm[q] : 0.f;
#pragma unroll
  for (int e = 0; e < kMergeE; ++e) {
    const int i = lane * kMergeE + e;
    if (i >= k) continue;
    const u64 key = acc[e];
    const bool valid = key != kKeyInf;
    const float sc = key_score(key);
    float d;
    if (metric == B2VS_METRIC_L2) d = valid ? fmaxf(sc + qn, 0.f) : INFINITY;
    else d = valid ? -sc : -INFINITY;
    const size_t o = static_cast<size_t>(q) * k + i;
    if (out_d) out_d[o] = d;
    uint32_t id = key_id(key);
    if (valid && remap) id = remap[id];
    if (out_i) out_i[o] = valid ? static_cast<long long>(id) + id_offset : -1ll;
    if (out_label) out_label[o] = valid ? static_cast<int>(id) : -1;
  }
}

// Sampled pass of a sharded search: the k best RAW scores of the query, ascending, +inf padded
// (comm.cu all-gathers them and takes the k-th best of the union as every shard's threshold).
__device__ __forceinline__ void emit_raw_scores(const u64 (&acc)[kMergeE], int lane, int q, int k,
                                                float* __restrict__ out_scores) {
#pragma unroll
  for (int e = 0; e < kMergeE; ++e) {
    const int i = lane * kMergeE + e;
    if (i < k) out_scores[static_cast<size_t>(q) * k + i] = acc[e] == kKeyInf ? INFINITY : key_score(acc[e]);
  }
}

__global__ void merge_splits_kernel(const u64* __restrict__ keys, int n_splits, int q_pad, int nq,
                                    int k, int metric, const float* __restrict__ qnorm,
                                    long long id_offset, float* __restrict__ out_d,
                                    long long* __restrict__ out_i, int* __restrict__ out_label,
                                    const uint32_t* __restrict__ remap,
                                    float* __restrict__ out_tau, float* __restrict__ out_scores) {
  const int lane = threadIdx.x & 31;
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= nq) return;
  u64 acc[kMergeE];
#pragma unroll
  for (int e = 0; e < kMergeE; ++e) acc[e] = kKeyInf;
  // The lists of kMergeAhead splits are loaded before any of them is folded in: with one warp
  // per query a small batch is a chain of dependent global loads otherwise (Q = 1 IVF-PQ coarse
  // probe, 64 splits: 65 us of a 236 us search).
  for (int s0 = 0; s0 < n_splits; s0 += kMergeAhead) {
    u64 ahead[kMergeAhead][kMergeE];
#pragma unroll
    for (int p = 0; p < kMergeAhead; ++p) {
      const bool live = s0 + p < n_splits;
      const u64* list = keys + (static_cast<size_t>(live ? s0 + p : s0) * q_pad + q) * k;
#pragma unroll
      for (int e = 0; e < kMergeE; ++e) {
        const int src = 32 * kMergeE - 1 - (lane * kMergeE + e);  // reversed run
        ahead[p][e] = (live && src < k) ? __ldcg(list + src) : kKeyInf;
      }
    }
#pragma unroll
    for (int p = 0; p < kMergeAhead; ++p) {
      if (s0 + p < n_splits) {   // warp-uniform
#pragma unroll
        for (int e = 0; e < kMergeE; ++e) acc[e] = acc[e] < ahead[p][e] ? acc[e] : ahead[p][e];
        warp_bitonic_merge<kMergeE>(acc, lane);
      }
    }
  }
  if (out_scores) {
    emit_raw_scores(acc, lane, q, k, out_scores);
    return;
  }
  if (out_tau) {
    // threshold-seeding pass: publish one ulp above the k-th best raw score (inclusive bound)
    u64 kth = kKeyInf;
#pragma unroll
    for (int e = 0; e < kMergeE; ++e)
      if (lane * kMergeE + e == k - 1) kth = acc[e];
    if (lane == (k - 1) / kMergeE)
      out_tau[q] = (kth == kKeyInf) ? INFINITY : nextafterf(key_score(kth), INFINITY);
    return;
  }
  emit_answer_row(acc, lane, q, k, metric, qnorm, id_offset, out_d, out_i, out_label, remap);
}

int launch_merge_splits(const u64* keys, int n_splits, int q_pad, int nq, int k, int metric,
                        const float* qnorm, int64_t id_offset, float* out_d, int64_t* out_i,
                        int32_t* out_label, cudaStream_t st, const uint32_t* remap, float* out_tau,
                        float* out_scores) {
  const int threads = 128;
  const int blocks = static_cast<int>(ceil_div(static_cast<int64_t>(nq) * 32, threads));
  merge_splits_kernel<<<blocks, threads, 0, st>>>(keys, n_splits, q_pad, nq, k, metric, qnorm,
                                                  id_offset, out_d,
                                                  reinterpret_cast<long long*>(out_i), out_label,
                                                  remap, out_tau, out_scores);
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}

// Two-pass selection over a small database (flat.cu: search_two_pass), the scalar side.
//
// chunk_tau_kernel: the first tensor-core pass left the minimum of every 32-row chunk of every
// query's score row.  The k-th smallest (minimum, chunk index) pair T bounds the answer exactly:
// an element qualifies iff (score, chunk) <= T.  The k chunks whose pair is <= T each hold an
// element that precedes any non-qualifying one in (score, id) order, so the true top-k all
// qualify; and every qualifying element lies in one of those k chunks, so at most 32 k elements
// do - ties included, because chunk indices are distinct.  One warp per query.
constexpr int kTwoPassThreads = 128;
__global__ void __launch_bounds__(kTwoPassThreads)
chunk_tau_kernel(const float* __restrict__ chunk_min, int chunk_ld, int n_chunks, int nq, int k,
                 float* __restrict__ tau, int* __restrict__ tau_chunk, int* __restrict__ count) {
  const int lane = threadIdx.x & 31;
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= nq) return;
  __shared__ u64 stage_mem[kTwoPassThreads / 32][kStageKeys];
  StagedTopK sel;
  sel.init(stage_mem[threadIdx.x >> 5]);
  const float* row = chunk_min + static_cast<size_t>(q) * chunk_ld;
  const float inf = __int_as_float(0x7f800000);
  for (int c0 = 0; c0 < n_chunks; c0 += 128) {     // four loads per lane in flight
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + 32 * j + lane;
      v[j] = c < n_chunks ? __ldcs(row + c) : inf;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool hit = v[j] <= sel.tk.tau && v[j] < inf;
      sel.push(hit ? pack_key(v[j], static_cast<uint32_t>(c0 + 32 * j + lane)) : kKeyInf, k, lane);
    }
  }
  sel.flush(k, lane);
  u64 kth = kKeyInf;
#pragma unroll
  for (int e = 0; e < kListE; ++e)
    if (lane * kListE + e == k - 1) kth = sel.tk.acc[e];
  if (lane == (k - 1) / kListE) {
    // fewer than k finite chunk minima: everything qualifies (n <= 32 * chunks < 32 k)
    tau[q] = kth == kKeyInf ? inf : key_score(kth);
    tau_chunk[q] = kth == kKeyInf ? 0x7fffffff : static_cast<int>(key_id(kth));
    count[q] = 0;
  }
}

// The k best of each query's qualifying elements (appended by the second pass), as answer rows.
__global__ void __launch_bounds__(kTwoPassThreads)
cand_select_kernel(const u64* __restrict__ cand, const int* __restrict__ count, int cap, int nq, int k,
                   int metric, const float* __restrict__ qnorm, long long id_offset,
                   float* __restrict__ out_d, long long* __restrict__ out_i, int* __restrict__ out_label) {
  static_assert(kMergeE == kListE, "answer rows are emitted from the 128-key list");
  const int lane = threadIdx.x & 31;
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= nq) return;
  __shared__ u64 stage_mem[kTwoPassThreads / 32][kStageKeys];
  StagedTopK sel;
  sel.init(stage_mem[threadIdx.x >> 5]);
  const int n = min(count[q], cap);    // count <= 32 k <= cap by construction
  const u64* src = cand + static_cast<size_t>(q) * cap;
  for (int i0 = 0; i0 < n; i0 += 64) {
    const int ia = i0 + lane, ib = i0 + 32 + lane;
    const u64 va = ia < n ? __ldcg(src + ia) : kKeyInf;
    const u64 vb = ib < n ? __ldcg(src + ib) : kKeyInf;
    sel.push((va != kKeyInf && key_score(va) <= sel.tk.tau) ? va : kKeyInf, k, lane);
    sel.push((vb != kKeyInf && key_score(vb) <= sel.tk.tau) ? vb : kKeyInf, k, lane);
  }
  sel.flush(k, lane);
  emit_answer_row(sel.tk.acc, lane, q, k, metric, qnorm, id_offset, out_d, out_i, out_label, nullptr);
}

int launch_chunk_tau(const float* chunk_min, int chunk_ld, int n_chunks, int nq, int k, float* tau,
                     int* tau_chunk, int* count, cudaStream_t st) {
  const int blocks = static_cast<int>(ceil_div(static_cast<int64_t>(nq) * 32, kTwoPassThreads));
  chunk_tau_kernel<<<blocks, kTwoPassThreads, 0, st>>>(chunk_min, chunk_ld, n_chunks, nq, k, tau, tau_chunk, count);
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}

int launch_cand_select(const u64* cand, const int* count, int cap, int nq, int k, int metric,
                       const float* qnorm, int64_t id_offset, float* out_d, int64_t* out_i,
                       int32_t* out_label, cudaStream_t st) {
  const int blocks = static_cast<int>(ceil_div(static_cast<int64_t>(nq) * 32, kTwoPassThreads));
  cand_select_kernel<<<blocks, kTwoPassThreads, 0, st>>>(cand, count, cap, nq, k, metric, qnorm, id_offset,
                                                         out_d, reinterpret_cast<long long*>(out_i), out_label);
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}

// Raw emission of the fused kernel (bf_tc.cuh: raw_count): every (split, query block) item left the
// UNSORTED contents of its rows' candidate buffers and their fill levels.  One warp per query streams
// the lists of all its splits through the staged selector (the items' final thresholds already cut
// them to ~k..256 keys each) and emits the answer row - or, for a sampled pass, the next threshold.
// Slot of (split s, query q): item = s * n_qblocks + q / (128 G); buffers ((item * G + rank) * E + e) * 128 + q % 128.
__global__ void __launch_bounds__(kTwoPassThreads)
merge_raw_kernel(const u64* __restrict__ cand, const int* __restrict__ count, int n_splits, int n_qblocks,
                 int group, int epi_groups, int nq, int k, int metric, const float* __restrict__ qnorm,
                 long long id_offset, float* __restrict__ out_d, long long* __restrict__ out_i,
                 int* __restrict__ out_label, float* __restrict__ out_tau, float* __restrict__ out_scores) {
  const int lane = threadIdx.x & 31;
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= nq) return;
  __shared__ u64 stage_mem[kTwoPassThreads / 32][kStageKeys];
  StagedTopK sel;
  sel.init(stage_mem[threadIdx.x >> 5]);
  const int qb = q / (128 * group), rank = (q / 128) % group, r = q % 128;
  for (int s = 0; s < n_splits; ++s) {
    for (int e = 0; e < epi_groups; ++e) {
      const size_t slot = ((static_cast<size_t>(s) * n_qblocks + qb) * group + rank) * epi_groups + e;
      const size_t row = slot * 128 + r;
      const int n = min(__ldg(count + row), kCap);
      const u64* src = cand + row * kCap;
      for (int i0 = 0; i0 < n; i0 += 64) {
        const int ia = i0 + lane, ib = i0 + 32 + lane;
        const u64 va = ia < n ? __ldcg(src + ia) : kKeyInf;
        const u64 vb = ib < n ? __ldcg(src + ib) : kKeyInf;
        sel.push((va != kKeyInf && key_score(va) <= sel.tk.tau) ? va : kKeyInf, k, lane);
        sel.push((vb != kKeyInf && key_score(vb) <= sel.tk.tau) ? vb : kKeyInf, k, lane);
      }
    }
  }
  sel.flush(k, lane);
  if (out_scores) {
    emit_raw_scores(sel.tk.acc, lane, q, k, out_scores);
    return;
  }
  if (out_tau) {
    // threshold-seeding pass: publish one ulp above the k-th best raw score (inclusive bound)
    u64 kth = kKeyInf;
#pragma unroll
    for (int e = 0; e < kListE; ++e)
      if (lane * kListE + e == k - 1) kth = sel.tk.acc[e];
    if (lane == (k - 1) / kListE)
      out_tau[q] = (kth == kKeyInf) ? INFINITY : nextafterf(key_score(kth), INFINITY);
    return;
  }
  emit_answer_row(sel.tk.acc, lane, q, k, metric, qnorm, id_offset, out_d, out_i, out_label, nullptr);
}

int launch_merge_raw(const u64* cand, const int* count, int n_splits, int n_qblocks, int group,
                     int epi_groups, int nq, int k, int metric, const float* qnorm, int64_t id_offset,
                     float* out_d, int64_t* out_i, int32_t* out_label, float* out_tau, cudaStream_t st,
                     float* out_scores) {
  const int blocks = static_cast<int>(ceil_div(static_cast<int64_t>(nq) * 32, kTwoPassThreads));
  merge_raw_kernel<<<blocks, kTwoPassThreads, 0, st>>>(cand, count, n_splits, n_qblocks, group, epi_groups,
                                                       nq, k, metric, qnorm, id_offset, out_d,
                                                       reinterpret_cast<long long*>(out_i), out_label, out_tau,
                                                       out_scores);
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}

// Sharded search, sampled pass: all_scores = [n_ranks][nq][k] raw scores (each rank's k best of ITS
// sample, ascending, +inf padded).  The k-th best of the union bounds the global k-th score from above
// like any rank's own k-th best does, but it is the k-th best of a sample n_ranks times larger: the
// ranks can sample n_ranks times more sparsely for the same threshold quality.  One warp per query.
__global__ void __launch_bounds__(kTwoPassThreads)
union_kth_kernel(const float* __restrict__ all_scores, int n_ranks, int nq, int k, float* __restrict__ tau) {
  const int lane = threadIdx.x & 31;
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= nq) return;
  __shared__ u64 stage_mem[kTwoPassThreads / 32][kStageKeys];
  StagedTopK sel;
  sel.init(stage_mem[threadIdx.x >> 5]);
  const float inf = __int_as_float(0x7f800000);
  for (int r = 0; r < n_ranks; ++r) {
    const float* src = all_scores + (static_cast<size_t>(r) * nq + q) * k;
    for (int i0 = 0; i0 < k; i0 += 32) {
      const int i = i0 + lane;
      const float v = i < k ? __ldcg(src + i) : inf;
      const bool hit = v <= sel.tk.tau && v < inf;
      sel.push(hit ? pack_key(v, static_cast<uint32_t>(r * k + i)) : kKeyInf, k, lane);
    }
  }
  sel.flush(k, lane);
  u64 kth = kKeyInf;
#pragma unroll
  for (int e = 0; e < kListE; ++e)
    if (lane * kListE + e == k - 1) kth = sel.tk.acc[e];
  if (lane == (k - 1) / kListE)
    tau[q] = (kth == kKeyInf) ? INFINITY : nextafterf(key_score(kth), INFINITY);
}

int launch_union_kth(const float* all_scores, int n_ranks, int nq, int k, float* tau, cudaStream_t st) {
  const int blocks = static_cast<int>(ceil_div(static_cast<int64_t>(nq) * 32, kTwoPassThreads));
  union_kth_kernel<<<blocks, kTwoPassThreads, 0, st>>>(all_scores, n_ranks, nq, k, tau);
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}

// Cross-shard merge.  Key = (orderable distance, position in the concatenated candidate row) so
// equal distances keep the lower part / lower rank first, exactly like a stable argsort over the
// concatenation.  Positions index d_all / i_all for the final gather.
__global__ void merge_parts_kernel(const float* __restrict__ d_all, const long long* __restrict__ i_all,
                                   int n_parts, int nq, int k_in, int k_out, int descending,
                                   float* __restrict__ out_d, long long* __restrict__ out_i) {
  const int lane = threadIdx.x & 31;
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= nq) return;
  u64 acc[kMergeE];
#pragma unroll
  for (int e = 0; e < kMergeE; ++e) acc[e] = kKeyInf;
  // Walk the concatenated candidate row in chunks of 128 (any order, any k_in): sort the chunk,
  // reverse it across the warp, min-combine with the running list, re-merge.
  const int total = n_parts * k_in;
  for (int c0 = 0; c0 < total; c0 += 32 * kMergeE) {
    u64 b[kMergeE];
#pragma unroll
    for (int e = 0; e < kMergeE; ++e) {
      const int pos = c0 + lane * kMergeE + e;
      b[e] = kKeyInf;
      if (pos < total) {
        const size_t src = (static_cast<size_t>(pos / k_in) * nq + q) * k_in + (pos % k_in);
        const float d = d_all[src];
        const long long id = i_all[src];
        if (id >= 0 && !isnan(d)) b[e] = pack_key(descending ? -d : d, static_cast<uint32_t>(pos));
      }
    }
    warp_bitonic_sort<kMergeE>(b, lane);
#pragma unroll
    for (int e = 0; e < kMergeE; ++e) {
      const u64 r = shfl_u64(b[kMergeE - 1 - e], 31 - lane);  // element 127 - i
      acc[e] = acc[e] < r ? acc[e] : r;
    }
    warp_bitonic_merge<kMergeE>(acc, lane);
  }
#pragma unroll
  for (int e = 0; e < kMergeE; ++e) {
    const int i = lane * kMergeE + e;
    if (i >= k_out) continue;
    const u64 key = acc[e];
    const size_t o = static_cast<size_t>(q) * k_out + i;
    if (key == kKeyInf) {
      out_d[o] = descending ? -INFINITY : INFINITY;
      out_i[o] = -1;
    } else {
      const uint32_t pos = key_id(key);
      const size_t src = (static_cast<size_t>(pos / k_in) * nq + q) * k_in + (pos % k_in);
      out_d[o] = d_all[src];
      out_i[o] = i_all[src];
    }
  }
}

}  // namespace b2vs

extern "C" int b2vs_merge_topk(int dev, const float* d_all, const int64_t* i_all, int n_parts,
                               int nq, int k_in, int k_out, int descending, float* out_d,
                               int64_t* out_i, void* stream) {
  using namespace b2vs;
  B2VS_CHECK(d_all && i_all && out_d && out_i, B2VS_EINVAL, "null pointer passed to b2vs_merge_topk");
  B2VS_CHECK(n_parts >= 1 && nq >= 1 && k_in >= 1, B2VS_EINVAL,
             "merge shape must be positive (n_parts=%d nq=%d k_in=%d)", n_parts, nq, k_in);
  B2VS_CHECK(k_out >= 1 && k_out <= kMaxBigK, B2VS_EUNSUP, "k_out=%d outside [1, %d]", k_out,
             kMaxBigK);
  DeviceGuard guard(dev);
  B2VS_CHECK(guard.ok, B2VS_ECUDA, "cannot select device %d", dev);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (k_out > kMaxFusedK)
    return launch_merge_parts_big(d_all, i_all, n_parts, nq, k_in, k_out, descending, out_d, out_i, st);
  const int threads = 128;
  const int blocks = static_cast<int>(ceil_div(static_cast<int64_t>(nq) * 32, threads));
  merge_parts_kernel<<<blocks, threads, 0, st>>>(d_all, reinterpret_cast<const long long*>(i_all),
                                                 n_parts, nq, k_in, k_out, descending, out_d,
                                                 reinterpret_cast<long long*>(out_i));
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}
