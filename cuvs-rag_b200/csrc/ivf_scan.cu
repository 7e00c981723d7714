// List-scan kernels of the IVF indexes and their launchers.
//   K5   ivf_flat_scan_kernel        one CTA per (query, probe), 128-bit loads, warp-resident top-k
//   K5b  helpers of the grouped tensor-core scan (bf_tc_kernel<1, true>, flat.cu): seed thresholds,
//        gathered query operand, per-query select, overflow rescue
//   K7   ivf_pq_scan_kernel / ivf_pq_scan_query_kernel   look-up-table ADC scans
//   K7b  helpers of the grouped PQ scan (pq_tc.cuh): LUT seed / rescue, gathered residual queries
//        refine_kernel / refine_big_kernel  exact re-rank of the ADC candidates
// Replaces what runs behind cuvs.neighbors.ivf_flat / ivf_pq .search at the reference call sites
// (improved_multi_gpu_rag.py:225-233, cuvs-2gpu-main.ipynb:L1801).
#include "ivf_internal.cuh"

namespace b2vs {

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

// The 16 codes of sub-spaces [m0, m0 + 16) of list slot (g << 5) + lane, packed into one uint4 (code
// of sub-space m0 + i in byte i).  Codes are stored sub-space major inside every 32-row group
// (pq_code_offset), the layout the grouped tensor-core scan's decoders want; the look-up-table
// scans below - lane = list row - collect their row's codes with 16 coalesced byte loads.
__device__ __forceinline__ uint4 ld_code_chunk(const uint8_t* __restrict__ codes, size_t g, int mp, int m0,
                                               int lane) {
  const uint8_t* p = codes + (g * static_cast<size_t>(mp) + static_cast<size_t>(m0)) * 32 + lane;
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    w[i] = static_cast<uint32_t>(__ldg(p + (4 * i) * 32)) | (static_cast<uint32_t>(__ldg(p + (4 * i + 1) * 32)) << 8) |
           (static_cast<uint32_t>(__ldg(p + (4 * i + 2) * 32)) << 16) |
           (static_cast<uint32_t>(__ldg(p + (4 * i + 3) * 32)) << 24);
  return make_uint4(w[0], w[1], w[2], w[3]);
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

// One warp per gathered row: row_item[v] = (query, probe) item or kNoRow on group padding.
__global__ void gather_group_queries_kernel(const uint32_t* __restrict__ row_item,
                                            const uint32_t* __restrict__ group_off, int n_lists,
                                            const float* __restrict__ qf, int dp, int n_probes,
                                            int fmt, int split, uint16_t* __restrict__ out,
                                            int* __restrict__ row_query, int* __restrict__ row_slot,
                                            const uint32_t* __restrict__ item_slot, int n_items) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  // by item (large batches): one warp per (query, probe) item, placed at the row the sort gave it -
  // the group-padding rows (2 M of them for 16 K lists) are never visited; row_query was pre-set to -1
  int64_t v = wid;
  uint32_t item;
  if (item_slot) {
    if (wid >= n_items) return;
    item = static_cast<uint32_t>(wid);
    v = item_slot[wid];
  } else {
    if (v >= static_cast<int64_t>(group_off[n_lists])) return;
    item = row_item[v];
  }
  const int q = item == kNoRow ? -1 : static_cast<int>(item / static_cast<uint32_t>(n_probes));
  if (lane == 0) {
    row_query[v] = q;
    row_slot[v] = item == kNoRow ? 0 : static_cast<int>(item % static_cast<uint32_t>(n_probes));
  }
  if (kSkipPaddingRows && q < 0) return;   // never qualifies (threshold -inf): bytes are don't-care
  if (!split) {
    // dp is a multiple of 8: 16-byte loads of four fp32 values, 8-byte stores of four 16-bit ones
    uint2* orow = reinterpret_cast<uint2*>(out + static_cast<size_t>(v) * dp);
    const float4* qrow = reinterpret_cast<const float4*>(qf + static_cast<size_t>(max(q, 0)) * dp);
    for (int j4 = lane; j4 < (dp >> 2); j4 += 32) {
      uint2 o = make_uint2(0u, 0u);
      if (q >= 0) {
        const float4 x = __ldg(qrow + j4);
        float back;
        o.x = static_cast<uint32_t>(to_op16(x.x, fmt, &back)) | (static_cast<uint32_t>(to_op16(x.y, fmt, &back)) << 16);
        o.y = static_cast<uint32_t>(to_op16(x.z, fmt, &back)) | (static_cast<uint32_t>(to_op16(x.w, fmt, &back)) << 16);
      }
      orow[j4] = o;
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

// Per query: the k best of the appended candidates, sorted.  One warp per query keeps a 128-key
// sorted list in registers (WarpTopK) and is offered the candidates 32 at a time, filtered by the
// running k-th score: after the first few offers almost every batch is rejected by one compare.
// (Round 1 sorted the whole buffer in shared memory, one CTA per query: 0.23 ms at C3.)
constexpr int kSelectThreads = 128;
__global__ void __launch_bounds__(kSelectThreads)
ivf_group_select_kernel(const u64* __restrict__ cand, const int* __restrict__ count, int cap, int k, int nq,
                        u64* __restrict__ out_keys, unsigned long long* __restrict__ total_cand,
                        int* __restrict__ over) {
  const int lane = threadIdx.x & 31;
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= nq) return;
  const int n = count[q];
  if (lane == 0 && total_cand) atomicAdd(total_cand, static_cast<unsigned long long>(n));
  if (n > cap) {   // overflow: queued for the rescue kernels (over[0] = how many, over[1..] = which)
    if (lane == 0) over[1 + atomicAdd(over, 1)] = q;
    return;
  }
  const u64* src = cand + static_cast<size_t>(q) * cap;
  __shared__ u64 stage_mem[kSelectThreads / 32][kStageKeys];
  StagedTopK sel;
  sel.init(stage_mem[threadIdx.x >> 5]);
  WarpTopK& tk = sel.tk;
  // two loads per lane in flight; keys are folded in 32 ACCEPTED keys at a time
  for (int i0 = 0; i0 < n; i0 += 64) {
    const int ia = i0 + lane, ib = i0 + 32 + lane;
    const u64 va = ia < n ? __ldcg(src + ia) : kKeyInf;
    const u64 vb = ib < n ? __ldcg(src + ib) : kKeyInf;
    // ties at the k-th score are settled by the key order in the merge
    sel.push((va != kKeyInf && key_score(va) <= tk.tau) ? va : kKeyInf, k, lane);
    sel.push((vb != kKeyInf && key_score(vb) <= tk.tau) ? vb : kKeyInf, k, lane);
  }
  sel.flush(k, lane);
  u64* out = out_keys + static_cast<size_t>(q) * k;
#pragma unroll
  for (int e = 0; e < kListE; ++e) {
    const int i = lane * kListE + e;
    if (i < k) out[i] = tk.acc[e];
  }
}

// Seed pass -> thresholds.  The tensor-core kernel appended the scores of the first rows of each
// query's nearest lists (no threshold) to the query's buffer; the k-th best of them is a valid
// upper bound of the query's final k-th score, computed by the SAME arithmetic as the full pass
// (identical MMA sequence over identical operands), so it needs no rounding cushion beyond the
// inclusive ulp.  One warp per query: a 128-key sorted list in registers, candidates offered 32
// at a time and filtered by the running k-th score.
__global__ void __launch_bounds__(128)
ivf_seed_select_kernel(const u64* __restrict__ cand, int n_keys, int cap, int k, int nq,
                       float* __restrict__ tau) {
  const int lane = threadIdx.x & 31;
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= nq) return;
  const int n = min(n_keys, cap);   // the seed pass fills fixed slots (kKeyInf where a list is short)
  const u64* src = cand + static_cast<size_t>(q) * cap;
  __shared__ u64 stage_mem[4][kStageKeys];
  StagedTopK sel;
  sel.init(stage_mem[threadIdx.x >> 5]);
  WarpTopK& tk = sel.tk;
  for (int i0 = 0; i0 < n; i0 += 128) {    // four loads per lane in flight
    u64 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = i0 + 32 * j + lane;
      v[j] = i < n ? __ldcg(src + i) : kKeyInf;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      sel.push((v[j] != kKeyInf && key_score(v[j]) < tk.tau) ? v[j] : kKeyInf, k, lane);
  }
  sel.flush(k, lane);
  // tk.tau = k-th best score (+inf while fewer than k candidates): publish it inclusively
  if (lane == 0) tau[q] = isinf(tk.tau) ? tk.tau : nextafterf(tk.tau, INFINITY);
}

// Rescue of the queries whose candidate buffer overflowed (threshold too loose for their
// neighbourhood): an exact scan of all their probes.  The select kernel queued them; every queued
// query is cut into kRescueSlices probe slices, one CTA each (a single CTA per query streamed
// n_probes whole lists - 240 MB at 64 probes of C3 - and ONE overflowing query cost the batch
// 5 ms), and a one-warp merge folds the slices' sorted lists.
constexpr int kRescueSlices = 8;
constexpr int kRescueCtas = 512;
template <int FMT, int J>
__global__ void __launch_bounds__(kScanThreads, 2)
ivf_flat_rescue_kernel(const uint16_t* __restrict__ data, const float* __restrict__ slot_norm,
                       const uint32_t* __restrict__ offsets, const long long* __restrict__ probe_ids,
                       const float* __restrict__ qf, int dp, int n_probes, int k, float alpha,
                       const int* __restrict__ over, u64* __restrict__ slice_keys) {
  __shared__ u64 lists[kScanWarps][32 * kListE];
  const int n_over = over[0];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = blockIdx.x; i < n_over; i += gridDim.x) {
    const int q = over[1 + i];
    float qr[J][8];
    load_query_regs<J>(qf, q, dp, lane, qr);
    WarpTopK tk;
    tk.init();
    for (int p = blockIdx.y; p < n_probes; p += kRescueSlices) {
      const long long list = probe_ids[static_cast<size_t>(q) * n_probes + p];
      if (list < 0) continue;
      scan_list_rows<FMT, J>(reinterpret_cast<const uint4*>(data), slot_norm, offsets[list],
                             offsets[list + 1], qr, dp >> 3, alpha, tk, k, warp, lane);
    }
    block_merge_and_store(tk, lists, k, warp, lane,
                          slice_keys + (static_cast<size_t>(i) * kRescueSlices + blockIdx.y) * k);
    __syncthreads();   // `lists` is reused by the next queued query
  }
}

// Folds the kRescueSlices sorted lists of every rescued query into its k answer keys (one warp each).
__global__ void __launch_bounds__(32)
ivf_rescue_merge_kernel(const int* __restrict__ over, const u64* __restrict__ slice_keys, int k,
                        u64* __restrict__ out_keys) {
  const int lane = threadIdx.x;
  const int n_over = over[0];
  for (int i = blockIdx.x; i < n_over; i += gridDim.x) {
    const int q = over[1 + i];
    const u64* src = slice_keys + static_cast<size_t>(i) * kRescueSlices * k;
    WarpTopK tk;
    tk.init();
    for (int j0 = 0; j0 < kRescueSlices * k; j0 += 32) {
      const int j = j0 + lane;
      const u64 v = j < kRescueSlices * k ? __ldcg(src + j) : kKeyInf;
      tk.offer((v != kKeyInf && key_score(v) <= tk.tau) ? v : kKeyInf, k, lane);
    }
#pragma unroll
    for (int e = 0; e < kListE; ++e) {
      const int idx = lane * kListE + e;
      if (idx < k) out_keys[static_cast<size_t>(q) * k + idx] = tk.acc[e];
    }
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
  for (uint32_t g0 = (begin >> 5) + warp; g0 < (end >> 5); g0 += kScanWarps) {
    const uint32_t slot = (g0 << 5) + lane;
    float s = bias;
    for (int ch = 0; ch < n_chunks; ++ch) {
      const uint4 v = ld_code_chunk(codes, g0, mp, ch * 16, lane);
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
          v[ch] = ld_code_chunk(codes, g_first, mp, ch * 16, lane);
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
              v[ch] = ld_code_chunk(codes, g0, mp, ch * 16, lane);
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
            const uint4 vv = ld_code_chunk(codes, g0, mp, cc * 16, lane);
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
  // rows and the fp32 query copy are 16-byte aligned when dim (and so dp) is a multiple of 4
  const bool vec4 = (dim & 3) == 0 && (dp & 3) == 0 && (reinterpret_cast<uintptr_t>(rows) & 15) == 0;
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
    if (vec4) {
      // four dimensions per lane and step: one 8-byte (16-bit rows) / 16-byte load per row
      for (int t = lane * 4; t < dim; t += 128) {
        const float4 qt = *reinterpret_cast<const float4*>(qv + t);
        float4 xv[kRefineAhead];
#pragma unroll
        for (int b = 0; b < kRefineAhead; ++b)
          xv[b] = row[b] >= 0 ? ld4_f32<T>(rows + static_cast<size_t>(row[b]) * dim + t)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int b = 0; b < kRefineAhead; ++b) {
          if (row[b] < 0) continue;
          if (metric == B2VS_METRIC_L2) {
            const float d0 = qt.x - xv[b].x, d1 = qt.y - xv[b].y, d2 = qt.z - xv[b].z, d3 = qt.w - xv[b].w;
            acc[b] = fmaf(d3, d3, fmaf(d2, d2, fmaf(d1, d1, fmaf(d0, d0, acc[b]))));
          } else {
            acc[b] = fmaf(-qt.w, xv[b].w, fmaf(-qt.z, xv[b].z, fmaf(-qt.y, xv[b].y, fmaf(-qt.x, xv[b].x, acc[b]))));
          }
        }
      }
    } else
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
  // mode 1: `count` is the overflow queue of the select kernel (count[0] = how many, count[1..] =
  // which queries); CTA (x, y) takes the queued queries x, x + gridDim.x, ... and of each the probe
  // slice y, y + kRescueSlices, ... -> out_keys[(queue position * kRescueSlices + y) * k]
  int i_over = blockIdx.x;
  const int n_over = mode == 1 ? count[0] : 0;
  const int l2 = metric == B2VS_METRIC_L2;
  const int n_chunks = pq_dim >> 4;
  const int mp = pq_dim;   // the grouped scan requires pq_dim % 16 == 0
  const int np = mode == 0 ? 1 : n_probes;
  const int p_first = mode == 0 ? 0 : static_cast<int>(blockIdx.y);
  const int p_step = mode == 0 ? 1 : kRescueSlices;
rescue_next:
  if (mode == 1 && i_over >= n_over) return;
  const int q = mode == 1 ? count[1 + i_over]
                          : (q_perm ? static_cast<int>(q_perm[blockIdx.x]) : static_cast<int>(blockIdx.x));
  WarpTopK tk;
  tk.init();
  float bias_first = 0.f;
  for (int p = p_first; p < np; p += p_step) {
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
    if (p == p_first) bias_first = bias;
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
        const uint4 v = ld_code_chunk(codes, g0, mp, ch * 16, lane);
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
    block_merge_and_store(tk, lists, k, warp, lane,
                          out_keys + (static_cast<size_t>(i_over) * kRescueSlices + blockIdx.y) * k);
    __syncthreads();   // lists / lut / rq are reused by the next queued query
    i_over += gridDim.x;
    goto rescue_next;
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
                                              float* __restrict__ row_bias, int* __restrict__ row_slot,
                                              const uint32_t* __restrict__ item_slot, int n_items) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  int64_t v = wid;     // by item / by row: see gather_group_queries_kernel
  uint32_t item;
  if (item_slot) {
    if (wid >= n_items) return;
    item = static_cast<uint32_t>(wid);
    v = item_slot[wid];
  } else {
    if (v >= static_cast<int64_t>(group_off[n_lists])) return;
    item = row_item[v];
  }
  uint16_t* orow = out + static_cast<size_t>(v) * dim;
  if (item == kNoRow) {
    // group padding: the row never qualifies (threshold -inf); its operand bytes are don't-care
    if (!kSkipPaddingRows)
      for (int j = lane; j < dim; j += 32) orow[j] = 0;
    if (lane == 0) { row_query[v] = -1; row_bias[v] = 0.f; row_slot[v] = 0; }
    return;
  }
  const int q = static_cast<int>(item / static_cast<uint32_t>(n_probes));
  if (lane == 0) row_slot[v] = static_cast<int>(item % static_cast<uint32_t>(n_probes));
  const long long list = probe_ids[item];
  // dim is a multiple of 64 on this path: four dimensions per lane and step (16-byte loads)
  const float4* c4 = reinterpret_cast<const float4*>(cent + static_cast<size_t>(list < 0 ? 0 : list) * dim);
  const float4* q4 = reinterpret_cast<const float4*>(qf + static_cast<size_t>(q) * dp);
  uint2* o2 = reinterpret_cast<uint2*>(orow);
  float acc = 0.f;
  for (int j4 = lane; j4 < (dim >> 2); j4 += 32) {
    const float4 qv = __ldg(q4 + j4);
    const float4 cv = __ldg(c4 + j4);
    const float in[4] = {l2 ? qv.x - cv.x : qv.x, l2 ? qv.y - cv.y : qv.y, l2 ? qv.z - cv.z : qv.z,
                         l2 ? qv.w - cv.w : qv.w};
    float back[4];
    uint16_t h[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = to_op16(in[e], 1, &back[e]);
    if (l2) acc += back[0] * back[0] + back[1] * back[1] + back[2] * back[2] + back[3] * back[3];
    else acc += qv.x * cv.x + qv.y * cv.y + qv.z * cv.z + qv.w * cv.w;
    o2[j4] = make_uint2(static_cast<uint32_t>(h[0]) | (static_cast<uint32_t>(h[1]) << 16),
                        static_cast<uint32_t>(h[2]) | (static_cast<uint32_t>(h[3]) << 16));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) { row_query[v] = q; row_bias[v] = l2 ? acc : -acc; }
}

// The same by item for 128-d indexes (one float4 per lane and row): a warp takes kGatherIlp
// consecutive items and issues all their loads before the first use - with one item per warp the
// kernel was a chain of three dependent loads per 256 bytes written (0.18 ms for the 640 K items
// of a C4 batch).
constexpr int kGatherIlp = 4;
__global__ void __launch_bounds__(256)
gather_group_residuals128_kernel(const long long* __restrict__ probe_ids, const float* __restrict__ qf,
                                 const float* __restrict__ cent, int n_probes, int l2,
                                 uint16_t* __restrict__ out, int* __restrict__ row_query,
                                 float* __restrict__ row_bias, int* __restrict__ row_slot,
                                 const uint32_t* __restrict__ item_slot, int n_items) {
  const int lane = threadIdx.x & 31;
  const int64_t i0 = ((static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5) * kGatherIlp;
  if (i0 >= n_items) return;
  uint32_t v[kGatherIlp];
  long long list[kGatherIlp];
  int q[kGatherIlp];
#pragma unroll
  for (int t = 0; t < kGatherIlp; ++t) {
    const int64_t it = min(i0 + t, static_cast<int64_t>(n_items) - 1);
    v[t] = __ldg(item_slot + it);
    list[t] = __ldg(probe_ids + it);
    q[t] = static_cast<int>(it / n_probes);
  }
  float4 qv[kGatherIlp], cv[kGatherIlp];
#pragma unroll
  for (int t = 0; t < kGatherIlp; ++t) {
    // consecutive items are consecutive probes of one query: its row is loaded once per warp
    if (t == 0 || q[t] != q[0])
      qv[t] = __ldg(reinterpret_cast<const float4*>(qf + static_cast<size_t>(q[t]) * 128) + lane);
    else
      qv[t] = qv[0];
    cv[t] = __ldg(reinterpret_cast<const float4*>(cent + static_cast<size_t>(list[t] < 0 ? 0 : list[t]) * 128) + lane);
  }
#pragma unroll
  for (int t = 0; t < kGatherIlp; ++t) {
    if (i0 + t >= n_items) break;
    const float in[4] = {l2 ? qv[t].x - cv[t].x : qv[t].x, l2 ? qv[t].y - cv[t].y : qv[t].y,
                         l2 ? qv[t].z - cv[t].z : qv[t].z, l2 ? qv[t].w - cv[t].w : qv[t].w};
    float back[4];
    uint16_t h[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = to_op16(in[e], 1, &back[e]);
    float acc = l2 ? back[0] * back[0] + back[1] * back[1] + back[2] * back[2] + back[3] * back[3]
                   : qv[t].x * cv[t].x + qv[t].y * cv[t].y + qv[t].z * cv[t].z + qv[t].w * cv[t].w;
    reinterpret_cast<uint2*>(out + static_cast<size_t>(v[t]) * 128)[lane] =
        make_uint2(static_cast<uint32_t>(h[0]) | (static_cast<uint32_t>(h[1]) << 16),
                   static_cast<uint32_t>(h[2]) | (static_cast<uint32_t>(h[3]) << 16));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      row_query[v[t]] = q[t];
      row_bias[v[t]] = l2 ? acc : -acc;
      row_slot[v[t]] = static_cast<int>((i0 + t) % n_probes);
    }
  }
}

// Instantiation table of the <FMT, J> scan kernels: J = 16-byte chunks of a row owned by a lane.
// (grid may be an int or a dim3)
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


// ------------------------------------------------------------------------------------------
// launchers
int launch_queries_to_f32(const void* q, int q_dtype, int nq, int dim, int dp, int fmt, int round16,
                          float* qf, float* qnorm, cudaStream_t st) {
  DISPATCH_DTYPE(q_dtype, T, (queries_to_f32_kernel<T><<<static_cast<unsigned>(ceil_div(nq, 4)), 128, 0, st>>>(
                                 static_cast<const T*>(q), nq, dim, dp, fmt, round16, qf, qnorm)));
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}

int launch_flat_item_scan(int fmt, int dp, int grid, const uint16_t* data, const float* slot_norm,
                          const uint32_t* offsets, const long long* probe_ids, const float* qf,
                          int n_probes, int q_pad, int k, float alpha, u64* out_keys,
                          unsigned long long* scanned_rows, const uint32_t* item_perm, cudaStream_t st) {
  const int j = static_cast<int>(ceil_div(dp / 8, 32));
  FLAT_SCAN_DISPATCH(ivf_flat_scan_kernel, fmt, j, grid, st, data, slot_norm, offsets, probe_ids, qf, dp,
                     n_probes, q_pad, k, alpha, out_keys, scanned_rows, item_perm);
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}

int launch_flat_seed_tau(const b2vs_index* index, IvfData* d, const long long* probe_ids, int n_probes,
                         int nq, int k, uint32_t seed_rows, float extra_eps, const uint32_t* q_perm,
                         cudaStream_t st) {
  const int j = static_cast<int>(ceil_div(d->dp / 8, 32));
  const float alpha = index->metric == B2VS_METRIC_L2 ? -2.f : -1.f;
  FLAT_SCAN_DISPATCH(ivf_seed_tau_kernel, d->fmt, j, nq, st, d->data.as<uint16_t>(),
                     d->slot_norm.as<float>(), d->offsets.as<uint32_t>(), probe_ids,
                     d->ws_qf.as<float>(), d->dp, n_probes, k, alpha, seed_rows, d->max_norm2,
                     index->metric == B2VS_METRIC_L2 ? 1 : 0, extra_eps, q_perm, d->ws_g_tau.as<float>());
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}

static int launch_rescue_merge(IvfData* d, int nq, int k, cudaStream_t st) {
  ivf_rescue_merge_kernel<<<std::min(nq, kRescueCtas), 32, 0, st>>>(d->ws_over.as<int>(), d->ws_rescue.as<u64>(), k,
                                                                   d->ws_keys.as<u64>());
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}

int launch_flat_rescue(const b2vs_index* index, IvfData* d, const long long* probe_ids, int n_probes,
                       int nq, int k, int cap, cudaStream_t st) {
  (void)cap;
  const int j = static_cast<int>(ceil_div(d->dp / 8, 32));
  const float alpha = index->metric == B2VS_METRIC_L2 ? -2.f : -1.f;
  B2VS_TRY(d->ws_rescue.reserve(static_cast<size_t>(nq) * kRescueSlices * k * sizeof(u64)));
  const dim3 grid(static_cast<unsigned>(std::min(nq, kRescueCtas)), kRescueSlices);
  FLAT_SCAN_DISPATCH(ivf_flat_rescue_kernel, d->fmt, j, grid, st, d->data.as<uint16_t>(),
                     d->slot_norm.as<float>(), d->offsets.as<uint32_t>(), probe_ids,
                     d->ws_qf.as<float>(), d->dp, n_probes, k, alpha, d->ws_over.as<int>(),
                     d->ws_rescue.as<u64>());
  B2VS_CUDA(cudaGetLastError());
  return launch_rescue_merge(d, nq, k, st);
}

int launch_group_select(IvfData* d, int nq, int cap, int k, unsigned long long* total_cand, cudaStream_t st) {
  // the overflow queue: [0] = count, [1 .. nq] = query ids
  B2VS_TRY(d->ws_over.reserve((static_cast<size_t>(nq) + 1) * sizeof(int)));
  B2VS_CUDA(cudaMemsetAsync(d->ws_over.ptr, 0, sizeof(int), st));
  ivf_group_select_kernel<<<static_cast<unsigned>(ceil_div(nq, kSelectThreads / 32)), kSelectThreads, 0, st>>>(
      d->ws_g_cand.as<u64>(), d->ws_g_cnt.as<int>(), cap, k, nq, d->ws_keys.as<u64>(), total_cand,
      d->ws_over.as<int>());
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}

int launch_seed_select(IvfData* d, int nq, int n_keys, int cap, int k, cudaStream_t st) {
  ivf_seed_select_kernel<<<static_cast<unsigned>(ceil_div(nq, 4)), 128, 0, st>>>(
      d->ws_g_cand.as<u64>(), n_keys, cap, k, nq, d->ws_g_tau.as<float>());
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}

// items > 0: the items were sorted by the counting-sort kernels (ws_item_slot holds every item's
// row): gather by item.  items == 0 (one-CTA planner): walk the rows.
static int prepare_by_item(IvfData* d, int64_t rows_cap, int items, cudaStream_t st) {
  if (items > 0)
    B2VS_CUDA(cudaMemsetAsync(d->ws_g_rowq.ptr, 0xFF, static_cast<size_t>(rows_cap) * sizeof(int), st));
  return B2VS_OK;
}

int launch_gather_group_queries(IvfData* d, int64_t rows_cap, int n_probes, int q_split, int items,
                                cudaStream_t st) {
  B2VS_TRY(prepare_by_item(d, rows_cap, items, st));
  const int64_t warps = items > 0 ? items : rows_cap;
  gather_group_queries_kernel<<<static_cast<unsigned>(ceil_div(warps, 8)), 256, 0, st>>>(
      d->ws_item_perm.as<uint32_t>(), d->ws_item_off.as<uint32_t>(), d->n_lists, d->ws_qf.as<float>(),
      d->dp, n_probes, d->fmt, q_split, d->ws_g_q.as<uint16_t>(), d->ws_g_rowq.as<int>(),
      d->ws_g_rowslot.as<int>(), items > 0 ? d->ws_item_slot.as<uint32_t>() : nullptr, items);
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}

int launch_gather_group_residuals(const b2vs_index* index, IvfData* d, int64_t rows_cap,
                                  const long long* probe_ids, int n_probes, int items, cudaStream_t st) {
  B2VS_TRY(prepare_by_item(d, rows_cap, items, st));
  if (items > 0 && index->dim == 128 && d->dp == 128) {
    const int64_t warps128 = ceil_div(static_cast<int64_t>(items), kGatherIlp);
    gather_group_residuals128_kernel<<<static_cast<unsigned>(ceil_div(warps128, 8)), 256, 0, st>>>(
        probe_ids, d->ws_qf.as<float>(), d->centroids.as<float>(), n_probes,
        index->metric == B2VS_METRIC_L2 ? 1 : 0, d->ws_g_q.as<uint16_t>(), d->ws_g_rowq.as<int>(),
        d->ws_g_bias.as<float>(), d->ws_g_rowslot.as<int>(), d->ws_item_slot.as<uint32_t>(), items);
    B2VS_CUDA(cudaGetLastError());
    return B2VS_OK;
  }
  const int64_t warps = items > 0 ? items : rows_cap;
  gather_group_residuals_kernel<<<static_cast<unsigned>(ceil_div(warps, 8)), 256, 0, st>>>(
      d->ws_item_perm.as<uint32_t>(), d->ws_item_off.as<uint32_t>(), d->n_lists, probe_ids,
      d->ws_qf.as<float>(), d->centroids.as<float>(), index->dim, d->dp, n_probes,
      index->metric == B2VS_METRIC_L2 ? 1 : 0, d->ws_g_q.as<uint16_t>(), d->ws_g_rowq.as<int>(),
      d->ws_g_bias.as<float>(), d->ws_g_rowslot.as<int>(),
      items > 0 ? d->ws_item_slot.as<uint32_t>() : nullptr, items);
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}

int launch_pq_lut_scan(int mode, const b2vs_index* index, IvfData* d, const long long* probe_ids,
                       int n_probes, int nq, int k, int cap, uint32_t seed_rows, const uint32_t* q_perm,
                       cudaStream_t st) {
  const size_t lut_smem = (static_cast<size_t>(d->pq_dim) * 256 + index->dim) * sizeof(float);
  B2VS_CUDA(cudaFuncSetAttribute(ivf_pq_lut_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(lut_smem)));
  dim3 grid(static_cast<unsigned>(nq));
  if (mode == 1) {   // rescue: queued queries x probe slices (see ivf_flat_rescue_kernel)
    B2VS_TRY(d->ws_rescue.reserve(static_cast<size_t>(nq) * kRescueSlices * k * sizeof(u64)));
    grid = dim3(static_cast<unsigned>(std::min(nq, kRescueCtas)), kRescueSlices);
  }
  ivf_pq_lut_scan_kernel<<<grid, kScanThreads, lut_smem, st>>>(
      mode, d->codes.as<uint8_t>(), d->row_ids.as<uint32_t>(), d->offsets.as<uint32_t>(), probe_ids,
      d->ws_qf.as<float>(), d->centroids.as<float>(), d->cb16.as<uint16_t>(), d->cbn.as<float>(),
      index->dim, d->dp, d->pq_dim, d->dsub, n_probes, k, index->metric, mode == 0 ? seed_rows : 0u,
      d->max_rhat2, mode == 0 ? q_perm : nullptr, mode == 0 ? nullptr : d->ws_over.as<int>(), cap,
      mode == 0 ? d->ws_g_tau.as<float>() : nullptr, mode == 0 ? nullptr : d->ws_rescue.as<u64>());
  B2VS_CUDA(cudaGetLastError());
  if (mode == 1) return launch_rescue_merge(d, nq, k, st);
  return B2VS_OK;
}

int launch_pq_table_scan(const b2vs_index* index, IvfData* d, const long long* probe_ids, int n_probes,
                         int nq, int q_pad, int k, unsigned long long* counter, bool* single_list,
                         cudaStream_t st) {
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
    *single_list = true;
  } else {
    const size_t smem = (static_cast<size_t>(d->mp) * 256 + index->dim) * sizeof(float);
    B2VS_CUDA(cudaFuncSetAttribute(ivf_pq_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem)));
    ivf_pq_scan_kernel<<<nq * n_probes, kScanThreads, smem, st>>>(
        d->codes.as<uint8_t>(), d->row_ids.as<uint32_t>(), d->offsets.as<uint32_t>(), probe_ids,
        d->ws_qf.as<float>(), d->centroids.as<float>(), d->codebooks.as<float>(), index->dim, d->dp,
        d->pq_dim, d->mp, d->dsub, n_probes, q_pad, k, index->metric, d->ws_keys.as<u64>(), counter);
    *single_list = false;
  }
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}

int launch_refine(const b2vs_index* index, IvfData* d, const long long* cand, int nq, int k_in, int k_out,
                  float* out_d, int64_t* out_i, cudaStream_t st) {
  DISPATCH_DTYPE(index->dtype, T, (refine_kernel<T><<<static_cast<unsigned>(ceil_div(nq, 4)), 128, 0, st>>>(
                                      static_cast<const T*>(d->src_rows), index->dim, d->ws_qf.as<float>(),
                                      d->dp, cand, nq, k_in, k_out, index->metric, index->id_offset, out_d,
                                      reinterpret_cast<long long*>(out_i))));
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}

int launch_refine_big(const b2vs_index* index, IvfData* d, const long long* cand, int nq, int k_in,
                      int k_out, float* out_d, int64_t* out_i, cudaStream_t st) {
  DISPATCH_DTYPE(index->dtype, T, (refine_big_kernel<T><<<nq, kRefineBigThreads, 0, st>>>(
                                      static_cast<const T*>(d->src_rows), index->dim, d->ws_qf.as<float>(),
                                      d->dp, cand, k_in, k_out, index->metric, index->id_offset, out_d,
                                      reinterpret_cast<long long*>(out_i))));
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}

}  // namespace b2vs
