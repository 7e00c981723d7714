// Large-k selection (128 < k <= 2048): the reference's "top-2000" retrieval mode
// (improved_multi_gpu_rag.py:37-48 SearchConfig(top_k=2000); per-shard k' = 2k at :247).
//
// The fused distance kernel runs in append mode: every score below a fixed per-query threshold
// goes to the query's global candidate buffer (bf_tc.cuh, kModeAppend).  Here one CTA per query
//  * radix-selects the k-th smallest 64-bit key of the buffer (8 passes x 8 bits, smem histogram),
//  * either publishes it as the next pass's threshold, or gathers the k keys at or below it,
//    sorts them in shared memory (bitonic, <= 2048 keys) and writes (distance, id).
#include <cmath>

#include "common.h"
#include "topk.cuh"

namespace b2vs {

constexpr int kBigThreads = 512;
constexpr int kBigSortMax = 2048;

__global__ void __launch_bounds__(kBigThreads)
bigk_select_kernel(const u64* __restrict__ cand, const int* __restrict__ counts, int cap, int k,
                   int final_pass, int metric, const float* __restrict__ qnorm, long long id_offset,
                   float* __restrict__ out_tau, float* __restrict__ out_d,
                   long long* __restrict__ out_i, int* __restrict__ overflow,
                   const uint32_t* __restrict__ remap) {
  __shared__ int hist[256];
  __shared__ u64 s_prefix;
  __shared__ int s_rank;
  __shared__ int s_fill;
  __shared__ u64 sortbuf[kBigSortMax];
  const int q = blockIdx.x;
  const int n_raw = counts[q];
  if (n_raw > cap && threadIdx.x == 0 && overflow) atomicMax(overflow, n_raw);
  const int n = n_raw < cap ? n_raw : cap;
  const u64* keys = cand + static_cast<size_t>(q) * cap;
  u64 kth = kKeyInf;
  if (n >= k) {
    // ---- radix select of the k-th smallest key, most significant byte first
    if (threadIdx.x == 0) { s_prefix = 0; s_rank = k; }
    for (int pass = 0; pass < 8; ++pass) {
      const int shift = 56 - 8 * pass;
      for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
      __syncthreads();
      const u64 prefix = s_prefix;
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const u64 key = __ldcg(keys + i);
        const bool match = (pass == 0) || ((key >> (shift + 8)) == (prefix >> (shift + 8)));
        if (match) atomicAdd(&hist[static_cast<int>((key >> shift) & 0xFFu)], 1);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int rank = s_rank, b = 0;
        for (; b < 256; ++b) {
          if (rank <= hist[b]) break;
          rank -= hist[b];
        }
        s_rank = rank;
        s_prefix = prefix | (static_cast<u64>(b) << shift);
      }
      __syncthreads();
    }
    kth = s_prefix;
  }
  if (!final_pass) {
    if (threadIdx.x == 0)
      out_tau[q] = (kth == kKeyInf) ? INFINITY : nextafterf(key_score(kth), INFINITY);
    return;
  }
  // ---- gather the (at most k, keys are unique) keys <= kth, sort, emit
  int p2 = 1;
  while (p2 < k) p2 <<= 1;
  for (int i = threadIdx.x; i < p2; i += blockDim.x) sortbuf[i] = kKeyInf;
  if (threadIdx.x == 0) s_fill = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const u64 key = __ldcg(keys + i);
    if (key <= kth) {
      const int pos = atomicAdd(&s_fill, 1);
      if (pos < p2) sortbuf[pos] = key;
    }
  }
  __syncthreads();
  for (int size = 2; size <= p2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < p2; i += blockDim.x) {
        const int j = i ^ stride;
        if (j > i) {
          const bool up = (i & size) == 0;
          const u64 a = sortbuf[i], b = sortbuf[j];
          if ((a > b) == up) { sortbuf[i] = b; sortbuf[j] = a; }
        }
      }
      __syncthreads();
    }
  }
  const float qn = (metric == B2VS_METRIC_L2 && qnorm) ? qnorm[q] : 0.f;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const u64 key = sortbuf[i];
    const size_t o = static_cast<size_t>(q) * k + i;
    if (key == kKeyInf) {
      out_d[o] = metric == B2VS_METRIC_L2 ? INFINITY : -INFINITY;
      out_i[o] = -1;
    } else {
      const float sc = key_score(key);
      out_d[o] = metric == B2VS_METRIC_L2 ? fmaxf(sc + qn, 0.f) : -sc;
      const uint32_t id = key_id(key);
      out_i[o] = static_cast<long long>(remap ? remap[id] : id) + id_offset;   // remap: list slot -> row
    }
  }
}

int launch_bigk_select(const u64* cand, const int* counts, int cap, int nq, int k, int final_pass,
                       int metric, const float* qnorm, int64_t id_offset, float* out_tau,
                       float* out_d, int64_t* out_i, int* overflow, cudaStream_t st,
                       const uint32_t* remap) {
  B2VS_CHECK(k <= kBigSortMax, B2VS_EUNSUP, "k=%d exceeds the large-k limit %d", k, kBigSortMax);
  bigk_select_kernel<<<nq, kBigThreads, 0, st>>>(cand, counts, cap, k, final_pass, metric, qnorm,
                                                 id_offset, out_tau, out_d,
                                                 reinterpret_cast<long long*>(out_i), overflow, remap);
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}

// Cross-shard merge for large k: one CTA per query sorts all n_parts * k_in candidates in smem.
constexpr int kBigMergeMax = 16384;

__global__ void __launch_bounds__(kBigThreads)
merge_parts_big_kernel(const float* __restrict__ d_all, const long long* __restrict__ i_all,
                       int n_parts, int nq, int k_in, int k_out, int descending,
                       float* __restrict__ out_d, long long* __restrict__ out_i) {
  extern __shared__ u64 buf[];
  const int q = blockIdx.x;
  const int total = n_parts * k_in;
  int p2 = 1;
  while (p2 < total) p2 <<= 1;
  for (int pos = threadIdx.x; pos < p2; pos += blockDim.x) {
    u64 key = kKeyInf;
    if (pos < total) {
      const size_t src = (static_cast<size_t>(pos / k_in) * nq + q) * k_in + (pos % k_in);
      const float d = d_all[src];
      if (i_all[src] >= 0 && !isnan(d)) key = pack_key(descending ? -d : d, static_cast<uint32_t>(pos));
    }
    buf[pos] = key;
  }
  __syncthreads();
  for (int size = 2; size <= p2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < p2; i += blockDim.x) {
        const int j = i ^ stride;
        if (j > i) {
          const bool up = (i & size) == 0;
          const u64 a = buf[i], b = buf[j];
          if ((a > b) == up) { buf[i] = b; buf[j] = a; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < k_out; i += blockDim.x) {
    const size_t o = static_cast<size_t>(q) * k_out + i;
    const u64 key = i < p2 ? buf[i] : kKeyInf;
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

int launch_merge_parts_big(const float* d_all, const int64_t* i_all, int n_parts, int nq, int k_in,
                           int k_out, int descending, float* out_d, int64_t* out_i,
                           cudaStream_t st) {
  const long long total = static_cast<long long>(n_parts) * k_in;
  B2VS_CHECK(total <= kBigMergeMax, B2VS_EUNSUP,
             "merge of %lld candidates per query exceeds the limit %d", total, kBigMergeMax);
  int p2 = 1;
  while (p2 < total) p2 <<= 1;
  const size_t smem = static_cast<size_t>(p2) * sizeof(u64);
  B2VS_CUDA(cudaFuncSetAttribute(merge_parts_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem)));
  merge_parts_big_kernel<<<nq, kBigThreads, smem, st>>>(
      d_all, reinterpret_cast<const long long*>(i_all), n_parts, nq, k_in, k_out, descending, out_d,
      reinterpret_cast<long long*>(out_i));
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}

}  // namespace b2vs
