// K0 — fused brute-force distance + top-k on the tcgen05 tensor cores (sm_100a).
//
// Replaces the distance GEMM + select_k that cuVS/FAISS run behind
// `ivf_flat.search` / `IndexFlat*.search` at the reference call sites
// (improved_multi_gpu_rag.py:225-233, cuvs-2gpu-main.ipynb:L1801, faiss-main.ipynb cell 9).
//
// Work decomposition: item = (db split s, query block qb).  A db tile is 256 rows (256 fp32 TMEM
// columns); two accumulator buffers fill the 512 TMEM columns so the epilogue of tile t overlaps
// the MMAs of tile t+1.  Items are ordered split-major so the CTAs resident at one time stream
// the SAME db rows against different query blocks and the db is read from HBM about once per
// batch (L2 serves the rest).
//
// Two instantiations:
//   G = 1  one CTA per item, query block = 128 rows, UMMA 128x256x16 (cta_group::1),
//          4 smem stages of (16 KB queries + 32 KB db).
//   G = 2  one CTA PAIR (cluster of 2, cta_group::2) per item, query block = 256 rows, UMMA
//          256x256x16: each CTA stages its own 128 query rows and HALF of the db tile, so the
//          L2->smem traffic and the smem operand reads per FLOP drop by a third / a half;
//          6 smem stages of (16 KB + 16 KB).  The leader CTA issues the MMAs; barriers that
//          gate it (smem full, accumulator empty) live in the leader and are signalled by both
//          CTAs, barriers it releases (smem empty, accumulator full) are multicast to both.
//
// Warp roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread),
// warp 2 = TMEM allocator, warps 4-7 = epilogue (thread i of warp w owns TMEM lane 32*(w-4)+i =
// one query row).  The epilogue never writes distances to HBM: score = alpha*acc + beta[col] is
// compared with the row's threshold and only candidates go to the row buffer (see topk.cuh).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "ptx.cuh"
#include "topk.cuh"

namespace b2vs {

constexpr int kBM = 128;   // query rows per CTA (TMEM lanes)
constexpr int kBN = 256;   // db rows per tile (TMEM columns)
#ifndef B2VS_BK
#define B2VS_BK 64
#endif
// K elements per smem stage: 64 = one 128-byte swizzle atom (default); 32 = 64-byte swizzle, half
// the stage size and more than twice the stage count (finer-grained operand prefetch)
constexpr int kBK = B2VS_BK;
static_assert(kBK == 64 || kBK == 32, "kBK must be 64 or 32");
constexpr int kNormBytes = kBN * 4;
// Epilogue warp groups: template parameter E of the kernel.  With E = 2 (warps 4-7 and 8-11) each
// group owns one of the two TMEM accumulator buffers, i.e. every other tile, with its own per-row
// candidate state; an item then leaves one sorted list per group and the split merge folds them.
//  * E = 1 for the long database passes: measured at C2 on the same box, 15 sustained steps,
//    2 groups 72.0 K QPS at ~1.2 GHz vs 1 group 73.6-74.3 K QPS at ~1.33 GHz - that kernel is
//    power-capped, and the extra warps cost more clock than the shorter accumulator hand-off wins;
//  * E = 2 was also tried where the epilogue is the whole kernel - the coarse probes of the IVF
//    indexes (top-64 of 16 384 centroids at C4: 79 CTAs, tensor pipe 5 % busy): slower too (732 vs
//    672 us), because two groups with private thresholds insert ~2 k ln(N/2k) candidates instead of
//    k ln(N/k).  The variant stays reachable (B2VS_FLAG_EPI2 / B2VS_EPI_GROUPS=2) for measurements.
constexpr int kMaxEpiGroups = 2;
// Work-table (grouped scan) epilogue: hits are staged in a small per-warp shared-memory queue and
// appended to the queries' global buffers in batches (one atomic round trip per batch instead of
// one per 32-column chunk on the epilogue's critical path).
constexpr int kQueueCap = 96;                                 // entries per epilogue warp
constexpr int kQueueWarpBytes = kQueueCap * (8 + 4);          // keys (u64) + query slots (int)
constexpr int kQueueBytes = 4 * kMaxEpiGroups * kQueueWarpBytes;   // one queue per epilogue warp

template <int G> struct TcCfg {
  static constexpr int kBRows = kBN / G;                     // db rows staged by one CTA
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = kBRows * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (kBK == 64) ? ((G == 1) ? 4 : 6) : ((G == 1) ? 9 : 13);
  static constexpr int kSmemBytes = kStages * kStageBytes + 2 * kNormBytes + 256 + kQueueBytes + 1024;
};

struct BfTcParams {
  const float* beta;    // [tiles_total*256] per db row additive term (||x||^2, 0, +inf on padding)
  u64* cand;            // [grid][E][128][kCap] candidate buffers (E = epilogue groups)
  u64* out_keys;        // [n_splits * E][q_pad][k] sorted ascending, kKeyInf padded
  // raw emission (k > 1, buffer mode): every item owns its candidate buffers,
  // cand[((item * G + cta_rank) * E + group) * 128 + row][kCap], and ends by storing the rows' fill
  // levels in raw_count (same indexing) - no final sort, no copy; merge_raw_kernel (merge.cu)
  // selects from the unsorted lists with the whole GPU instead of four warps per SM while the
  // tensor pipe waits (0.8 ms of an 18.3 ms search of a 1.25 M-row shard)
  int* raw_count;
  int n_qblocks;        // ceil(nq / (128*G))
  int q_pad;            // n_qblocks * 128 * G
  int nq;               // real query rows; rows >= nq are padding and never collect candidates
  int n_items;          // n_qblocks * n_splits
  int tiles_total;      // tiles visited by this launch: ceil(ceil(n_db / 256) / tile_stride)
  int tiles_per_split;
  int tile_stride;      // visit db tiles 0, stride, 2*stride, ... (threshold-seeding passes sample)
  const float* tau_init;  // [q_pad] starting threshold per query row (NULL = +inf)
  int k_blocks;         // ceil(K / 64)
  int k;                // 1..kMaxFusedK
  float alpha;          // -2 (L2) or -1 (inner product)
  uint32_t idesc;
  // large-k mode (k > kMaxFusedK): every score below the row's (fixed) threshold is appended to a
  // per-query global buffer shared by all splits; selection happens in bigk.cu
  u64* big_cand;        // [q_pad][big_cap] or NULL
  int* big_count;       // [q_pad] atomic append cursors
  int big_cap;
  // work-table mode (kWork kernels: the grouped IVF-Flat list scan, ivf.cu).  Item i is
  // work[i] = {query block, first db row, end db row, -}: the rows are one inverted list, the
  // query block holds (a slice of) the queries that probe it.  Always append mode; query row r
  // appends to the buffer of query row_query[r] (< 0: padding row) with threshold tau_init[query].
  const int4* work;
  const int* n_work;    // device scalar: number of work items
  const int* row_query; // [query rows]
  int x_kblocks;        // > 0: the query operand is [hi | lo] (2*x_kblocks k-blocks) against the SAME
                        // x_kblocks db k-blocks (fp32 queries split into two bf16 halves)
  // seed pass of the grouped scans (work mode): every score of the item's (single) tile is stored
  // at a FIXED place - big_cand[query][row_slot[row] * kSeedSlotRows + column] - no threshold, no
  // atomics; the caller pre-fills the buffers with kKeyInf
  int seed_all;
  const int* row_slot;  // [query rows] which of the query's seed lists this gathered row probes
  // two-pass selection over a small database (work mode; flat.cu: search_two_pass).
  //  seed_all == 2: only the MINIMUM of every 32-column chunk is stored, at
  //    chunk_min[query][column / 32] (row stride chunk_ld floats);
  //  tau_chunk != NULL: the threshold is a (score, chunk) pair - a score equal to tau_init[query]
  //    qualifies iff its chunk index is <= tau_chunk[query].
  float* chunk_min;
  int chunk_ld;
  const int* tau_chunk;
  // work mode: bytes one stage receives when the list-tile map's box is shorter than 256 rows (the
  // seed pass loads only the head of each list); 0 = a full stage
  uint32_t stage_tx;
  int debug_skip_emit;  // B2VS_K0_DEBUG=1 (timing experiments): items skip their final sort + emission
  int a_quarter_boxes;  // work mode: the query map's box is 32 rows; a stage loads only the block's leading quarters
                        // that hold queries (work.w rows, dealt over ceil(w / 32) quarters by the planner) - the
                        // seed pass's blocks hold ~5 queries: 3/4 of its query-operand traffic was padding
  int tail_boxes;       // work mode: use the TailMaps (128 / 64-row boxes) for the last tile of every item
};
constexpr int kSeedSlotRows = 256;   // = one tile: the seed pass scores the first tile of a list

constexpr bool kWorkNoHint = false;   // A/B switch for the evict-first hint of work-mode list tiles
constexpr int kModeBuffer = 0;   // per-(CTA,row) candidate buffer + warp compaction (k <= 128)
constexpr int kModeArgmin = 1;   // k == 1: running arg-min in a register
constexpr int kModeAppend = 2;   // large k: atomic append to the query's global buffer


// Scores one 32-column chunk of one query row held in registers: score = alpha*acc + beta.
// Fast path (almost always taken once the threshold has warmed up): a min-reduction and ONE
// compare, no per-element branches.  Slow path: append every score below the threshold to the
// row's candidate buffer (or track the arg-min when k == 1).
template <int kMode>
__device__ __forceinline__ void score_chunk(const uint32_t (&r)[32], const float4* __restrict__ nrm4,
                                            float alpha, uint32_t col, float& tau, int& cnt,
                                            u64& best, u64* __restrict__ my_cand,
                                            int* __restrict__ g_cnt = nullptr, int g_cap = 0,
                                            float bias = 0.f) {
  float sc[32];
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const float4 nb = nrm4[j4];
    sc[4 * j4 + 0] = fmaf(alpha, __uint_as_float(r[4 * j4 + 0]), nb.x);
    sc[4 * j4 + 1] = fmaf(alpha, __uint_as_float(r[4 * j4 + 1]), nb.y);
    sc[4 * j4 + 2] = fmaf(alpha, __uint_as_float(r[4 * j4 + 2]), nb.z);
    sc[4 * j4 + 3] = fmaf(alpha, __uint_as_float(r[4 * j4 + 3]), nb.w);
  }
  float m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
    m[i] = fminf(fminf(sc[4 * i], sc[4 * i + 1]), fminf(sc[4 * i + 2], sc[4 * i + 3]));
  const float mn = fminf(fminf(fminf(m[0], m[1]), fminf(m[2], m[3])),
                         fminf(fminf(m[4], m[5]), fminf(m[6], m[7])));
  if (mn < tau) {
    if (kMode == kModeArgmin) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (sc[j] < tau) {
          tau = sc[j];
          best = pack_key(sc[j], col + j);
        }
      }
    } else {
      // Hits are rare per row but common per warp (32 rows x 32 columns), so the cost here must
      // scale with the number of hits: build the hit mask without branches, then visit set bits,
      // pulling the score out of the register array with a 5-level select tree.
      uint32_t mask = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) mask |= (sc[j] < tau) ? (1u << j) : 0u;
      // append mode: ONE atomic reserves the slots of all of this chunk's hits
      int pos = 0;
      if (kMode == kModeAppend) pos = atomicAdd(g_cnt, __popc(mask));
      while (mask) {
        const int j = __ffs(mask) - 1;
        mask &= mask - 1;
        float v16[16], v8[8], v4[4], v2[2];
#pragma unroll
        for (int i = 0; i < 16; ++i) v16[i] = (j & 1) ? sc[2 * i + 1] : sc[2 * i];
#pragma unroll
        for (int i = 0; i < 8; ++i) v8[i] = (j & 2) ? v16[2 * i + 1] : v16[2 * i];
#pragma unroll
        for (int i = 0; i < 4; ++i) v4[i] = (j & 4) ? v8[2 * i + 1] : v8[2 * i];
#pragma unroll
        for (int i = 0; i < 2; ++i) v2[i] = (j & 8) ? v4[2 * i + 1] : v4[2 * i];
        const float v = (j & 16) ? v2[1] : v2[0];
        if (kMode == kModeAppend) {
          if (pos < g_cap) __stcg(my_cand + pos, pack_key(v + bias, col + j));  // bias: pq_tc.cuh
          ++pos;
        } else {
          __stcg(my_cand + cnt, pack_key(v, col + j));
          ++cnt;
        }
      }
    }
  }
}

// A chunk appends at most 32 keys per row: compact (warp-cooperatively) the rows that could
// overflow on the next chunk, which also tightens their thresholds.
__device__ __forceinline__ void compact_if_needed(int& cnt, float& tau, u64* cand_warp, int k, int lane) {
  uint32_t need = __ballot_sync(0xffffffffu, cnt > compact_trigger(k));
  while (need) {
    const int rr = __ffs(need) - 1;
    need &= need - 1;
    const int n_r = __shfl_sync(0xffffffffu, cnt, rr);
    u64* rb = cand_warp + static_cast<size_t>(rr) * kCap;
    const float nt = compact_row(rb, rb, n_r, k, lane);
    if (lane == rr) { cnt = min(n_r, k); tau = nt; }
  }
}

struct HitQueue {
  u64* keys;    // [kQueueCap] shared memory, private to one epilogue warp
  int* slots;   // [kQueueCap] query slot of each key
  int n;        // fill level (warp-uniform register)
};

// Appends the queued hits to their queries' global buffers: every lane takes entries, so the
// atomics of up to 32 entries are in flight together.  Warp-collective.
__device__ __forceinline__ void queue_drain(HitQueue& hq, u64* __restrict__ cand, int* __restrict__ count,
                                            int cap, int lane) {
  __syncwarp();
  constexpr int kPerLane = (kQueueCap + 31) / 32;
  u64 key[kPerLane];
  int slot[kPerLane], pos[kPerLane];
#pragma unroll
  for (int i = 0; i < kPerLane; ++i) {     // every atomic of the batch is issued before any is awaited
    const int e = lane + 32 * i;
    pos[i] = cap;
    if (e < hq.n) {
      key[i] = hq.keys[e];
      slot[i] = hq.slots[e];
      pos[i] = atomicAdd(count + slot[i], 1);
    }
  }
#pragma unroll
  for (int i = 0; i < kPerLane; ++i)
    if (pos[i] < cap) __stcg(cand + static_cast<size_t>(slot[i]) * cap + pos[i], key[i]);
  __syncwarp();
  hq.n = 0;
}

// Work-table epilogue: scores one 32-column chunk of the warp's 32 query rows and queues every
// score below its row's threshold.  Warp-collective (all lanes call it for the same chunk): the
// queue positions come from a warp prefix sum of the per-lane hit counts, no atomics.
__device__ __forceinline__ void score_chunk_queue(const uint32_t (&r)[32], const float4* __restrict__ nrm4,
                                                  float alpha, uint32_t col, float tau, float bias, int qslot,
                                                  HitQueue& hq, u64* __restrict__ cand,
                                                  int* __restrict__ count, int cap, int lane) {
  float sc[32];
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const float4 nb = nrm4[j4];
    sc[4 * j4 + 0] = fmaf(alpha, __uint_as_float(r[4 * j4 + 0]), nb.x);
    sc[4 * j4 + 1] = fmaf(alpha, __uint_as_float(r[4 * j4 + 1]), nb.y);
    sc[4 * j4 + 2] = fmaf(alpha, __uint_as_float(r[4 * j4 + 2]), nb.z);
    sc[4 * j4 + 3] = fmaf(alpha, __uint_as_float(r[4 * j4 + 3]), nb.w);
  }
  float m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
    m[i] = fminf(fminf(sc[4 * i], sc[4 * i + 1]), fminf(sc[4 * i + 2], sc[4 * i + 3]));
  const float mn = fminf(fminf(fminf(m[0], m[1]), fminf(m[2], m[3])),
                         fminf(fminf(m[4], m[5]), fminf(m[6], m[7])));
  if (!__any_sync(0xffffffffu, mn < tau)) return;      // the common case: nothing in this chunk
  uint32_t mask = 0;
#pragma unroll
  for (int j = 0; j < 32; ++j) mask |= (sc[j] < tau) ? (1u << j) : 0u;
  const int nh = __popc(mask);
  const int total = __reduce_add_sync(0xffffffffu, nh);
  // pulls score j out of the register array with a 5-level select tree
  auto pick = [&](int j) {
    float v16[16], v8[8], v4[4], v2[2];
#pragma unroll
    for (int i = 0; i < 16; ++i) v16[i] = (j & 1) ? sc[2 * i + 1] : sc[2 * i];
#pragma unroll
    for (int i = 0; i < 8; ++i) v8[i] = (j & 2) ? v16[2 * i + 1] : v16[2 * i];
#pragma unroll
    for (int i = 0; i < 4; ++i) v4[i] = (j & 4) ? v8[2 * i + 1] : v8[2 * i];
#pragma unroll
    for (int i = 0; i < 2; ++i) v2[i] = (j & 8) ? v4[2 * i + 1] : v4[2 * i];
    return (j & 16) ? v2[1] : v2[0];
  };
  if (total > kQueueCap) {
    // a flood (loose threshold): every lane reserves its slots in the query's buffer directly
    int pos = nh ? atomicAdd(count + qslot, nh) : 0;
    u64* const row_buf = cand + static_cast<size_t>(qslot) * cap;
    while (mask) {
      const int j = __ffs(mask) - 1;
      mask &= mask - 1;
      if (pos < cap) __stcg(row_buf + pos, pack_key(pick(j) + bias, col + j));
      ++pos;
    }
    return;
  }
  // Sparse hits (rarely two in one lane): every round takes each lane's lowest remaining hit, and
  // the queue positions come from the vote of the lanes that still have one - no prefix sum.
  if (hq.n + total > kQueueCap) queue_drain(hq, cand, count, cap, lane);
  while (true) {
    const bool has = mask != 0u;
    const uint32_t votes = __ballot_sync(0xffffffffu, has);
    if (votes == 0u) break;
    if (has) {
      const int j = __ffs(mask) - 1;
      mask &= mask - 1;
      const int pos = hq.n + __popc(votes & ((1u << lane) - 1u));
      hq.keys[pos] = pack_key(pick(j) + bias, col + j);
      hq.slots[pos] = qslot;
    }
    hq.n += __popc(votes);
  }
}

// Seed pass: all 32 scores of the chunk go to their fixed places (dst = the row's slot + column).
__device__ __forceinline__ void score_chunk_seed(const uint32_t (&r)[32], const float4* __restrict__ nrm4,
                                                 float alpha, uint32_t col, float bias,
                                                 u64* __restrict__ dst) {
  ulonglong2* d2 = reinterpret_cast<ulonglong2*>(dst);
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const float4 nb = nrm4[j4];
    const u64 k0 = pack_key(fmaf(alpha, __uint_as_float(r[4 * j4 + 0]), nb.x) + bias, col + 4 * j4 + 0);
    const u64 k1 = pack_key(fmaf(alpha, __uint_as_float(r[4 * j4 + 1]), nb.y) + bias, col + 4 * j4 + 1);
    const u64 k2 = pack_key(fmaf(alpha, __uint_as_float(r[4 * j4 + 2]), nb.z) + bias, col + 4 * j4 + 2);
    const u64 k3 = pack_key(fmaf(alpha, __uint_as_float(r[4 * j4 + 3]), nb.w) + bias, col + 4 * j4 + 3);
    __stcg(d2 + 2 * j4, make_ulonglong2(k0, k1));
    __stcg(d2 + 2 * j4 + 1, make_ulonglong2(k2, k3));
  }
}

// Two-pass selection, first pass: the minimum of the chunk's 32 scores.
__device__ __forceinline__ float score_chunk_min(const uint32_t (&r)[32], const float4* __restrict__ nrm4,
                                                 float alpha) {
  float m[8];
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const float4 nb = nrm4[j4];
    m[j4] = fminf(fminf(fmaf(alpha, __uint_as_float(r[4 * j4 + 0]), nb.x),
                        fmaf(alpha, __uint_as_float(r[4 * j4 + 1]), nb.y)),
                  fminf(fmaf(alpha, __uint_as_float(r[4 * j4 + 2]), nb.z),
                        fmaf(alpha, __uint_as_float(r[4 * j4 + 3]), nb.w)));
  }
  return fminf(fminf(fminf(m[0], m[1]), fminf(m[2], m[3])), fminf(fminf(m[4], m[5]), fminf(m[6], m[7])));
}

// Work mode: list-tile maps with 128- and 64-row boxes for the LAST tile of an item.  A 256-row box
// always reads 256 rows, i.e. on average 128 rows of the NEXT list behind every list's tail: 0.8 GB
// of a 15.4 GB C3 batch.  Rows the short box leaves stale in the stage only feed accumulator
// columns past the item's end, which the epilogue never scores.
struct alignas(64) TailMaps {
  CUtensorMap m128, m64;
};

constexpr int tc_threads(int epi_groups) { return 128 + 128 * epi_groups; }

template <int G, bool kWork = false, int E = 1>
__global__ void __launch_bounds__(tc_threads(E), 1)
bf_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_x,
             const BfTcParams p, const __grid_constant__ TailMaps tails) {
  static_assert(!kWork || G == 1, "work-table mode is single-CTA");
  static_assert(E == 1 || E == 2, "epilogue groups");
  constexpr int kEpiGroups = E;
  using Cfg = TcCfg<G>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kStageBytes = Cfg::kStageBytes;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  const uint32_t smem_base = (raw_addr + 1023u) & ~1023u;   // same offset in both CTAs of a pair
  uint8_t* smem = smem_raw + (smem_base - raw_addr);

  const uint32_t norm_base = smem_base + kStages * kStageBytes;
  const float* norm_ptr = reinterpret_cast<const float*>(smem + kStages * kStageBytes);
  const uint32_t bar_base = norm_base + 2 * kNormBytes;
  // barrier slots (8 bytes each)
  const uint32_t bar_full = bar_base;                     // [kStages] TMA -> MMA   (leader's is used)
  const uint32_t bar_empty = bar_base + 8 * kStages;      // [kStages] MMA -> TMA   (per CTA)
  const uint32_t bar_acc_full = bar_base + 16 * kStages;  // [2] MMA -> epilogue    (per CTA)
  const uint32_t bar_acc_empty = bar_acc_full + 16;       // [2] epilogue -> MMA    (leader's is used)
  const uint32_t bar_norm_full = bar_acc_full + 32;       // [2] TMA -> epilogue    (per CTA)
  const uint32_t bar_norm_empty = bar_acc_full + 48;      // [2] epilogue -> TMA    (per CTA)
  const uint32_t tmem_slot = bar_acc_full + 64;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
      smem + kStages * kStageBytes + 2 * kNormBytes + 16 * kStages + 64);

  uint8_t* const queue_mem = smem + kStages * kStageBytes + 2 * kNormBytes + 256;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (G == 2) ? ptx::cluster_ctarank() : 0u;
  const int unit = (G == 2) ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int n_units = (G == 2) ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(bar_full + 8 * i, G);          // one arrive(+tx) per CTA of the pair
      ptx::mbar_init(bar_empty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(bar_acc_full + 8 * i, 1);
      ptx::mbar_init(bar_acc_empty + 8 * i, 4 * G);  // every epilogue warp of the pair
      ptx::mbar_init(bar_norm_full + 8 * i, 1);
      ptx::mbar_init(bar_norm_empty + 8 * i, 4);
    }
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&tm_q);
    ptx::prefetch_tmap(&tm_x);
    if (kWork && p.tail_boxes) {
      ptx::prefetch_tmap(&tails.m128);
      ptx::prefetch_tmap(&tails.m64);
    }
  }
  if (warp == 2) {
    if (G == 2) {
      ptx::tmem_alloc_2sm(tmem_slot, 512);
      ptx::tmem_relinquish_2sm();
    } else {
      ptx::tmem_alloc(tmem_slot, 512);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  if (G == 2) ptx::cluster_sync_all(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // Whole warp walks the loops and waits; one elected lane issues the copies.
    {
      uint32_t stage = 0, phase = 0, tcount = 0;
      const uint32_t full_leader = (G == 2) ? ptx::mapa_cluster(bar_full, 0) : bar_full;
      const int n_items = kWork ? *p.n_work : p.n_items;
      for (int item = unit; item < n_items; item += n_units) {
        int qb, t0, t1, row_begin = 0, row_end = 0, a_quarters = 4;
        if (kWork) {
          const int4 w = __ldg(p.work + item);
          qb = w.x; row_begin = w.y; row_end = w.z; t0 = 0; t1 = (w.z - w.y + kBN - 1) / kBN;
          // .w = query rows of the block, which the planner placed in its first ceil(w / 32) quarters
          a_quarters = (w.w > 0 && w.w < kBM) ? (w.w + 31) >> 5 : 4;
        } else {
          qb = item % p.n_qblocks;
          t0 = (item / p.n_qblocks) * p.tiles_per_split;
          t1 = min(t0 + p.tiles_per_split, p.tiles_total);
        }
        const int q_row0 = qb * (kBM * G) + static_cast<int>(cta_rank) * kBM;
        for (int ti = t0; ti < t1; ++ti, ++tcount) {
          const int t = ti * p.tile_stride;
          const uint32_t as = tcount & 1u, aph = (tcount >> 1) & 1u;
          // first db row of the tile: tile-aligned, or (work mode) relative to the list start
          const int x_row0 = kWork ? row_begin + ti * kBN
                                   : t * kBN + static_cast<int>(cta_rank) * Cfg::kBRows;
          ptx::mbar_wait(bar_norm_empty + 8 * as, aph ^ 1u);
          if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(bar_norm_full + 8 * as, kNormBytes);
            ptx::bulk_load_1d(norm_base + as * kNormBytes,
                              p.beta + (kWork ? static_cast<size_t>(x_row0) : static_cast<size_t>(t) * kBN),
                              kNormBytes, bar_norm_full + 8 * as);
          }
          __syncwarp();
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
            const uint32_t a_dst = smem_base + stage * kStageBytes;
            if (ptx::elect_one()) {
              if (G == 2) {
                const uint32_t fb = full_leader + 8 * stage;
                ptx::mbar_arrive_expect_tx_cluster(fb, kStageBytes);
                ptx::tma_load_2d_2sm_hint(a_dst, &tm_q, fb, kb * kBK, q_row0, ptx::kEvictLast);
                ptx::tma_load_2d_2sm(a_dst + Cfg::kABytes, &tm_x, fb, kb * kBK, x_row0);
              } else {
                // work mode, last tile of the item: the shortest box that covers the rows left
                const CUtensorMap* tmx = &tm_x;
                uint32_t tx = (kWork && p.stage_tx) ? p.stage_tx : kStageBytes;
                if (kWork && p.tail_boxes) {
                  const int left = row_end - x_row0;
                  if (left <= 64) { tmx = &tails.m64; tx = Cfg::kABytes + 64 * kBK * 2; }
                  else if (left <= 128) { tmx = &tails.m128; tx = Cfg::kABytes + 128 * kBK * 2; }
                }
                if (kWork && p.a_quarter_boxes) {
                  // the query map's box is one 32-row quarter: only the quarters that hold queries
                  tx -= static_cast<uint32_t>(4 - a_quarters) * (32u * kBK * 2u);
                  ptx::mbar_arrive_expect_tx(bar_full + 8 * stage, tx);
                  for (int aq = 0; aq < a_quarters; ++aq)
                    ptx::tma_load_2d_hint(a_dst + aq * (32 * kBK * 2), &tm_q, bar_full + 8 * stage, kb * kBK,
                                          q_row0 + aq * 32, ptx::kEvictLast);
                } else {
                  ptx::mbar_arrive_expect_tx(bar_full + 8 * stage, tx);
                  // the query block is re-read for every db tile: ask L2 to keep it (evict-last)
                  ptx::tma_load_2d_hint(a_dst, &tm_q, bar_full + 8 * stage, kb * kBK, q_row0,
                                        ptx::kEvictLast);
                }
                const int xkb = (kWork && p.x_kblocks > 0 && kb >= p.x_kblocks) ? kb - p.x_kblocks : kb;
                // work mode streams every list once: keep it from evicting the query blocks in L2
                if (kWork && !kWorkNoHint)
                  ptx::tma_load_2d_hint(a_dst + Cfg::kABytes, tmx, bar_full + 8 * stage, xkb * kBK,
                                        x_row0, ptx::kEvictFirst);
                else
                  ptx::tma_load_2d(a_dst + Cfg::kABytes, &tm_x, bar_full + 8 * stage, xkb * kBK, x_row0);
              }
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA)
    // The whole warp walks the loops and waits on the barriers; one elected lane issues.
    if (cta_rank == 0) {
      uint32_t stage = 0, phase = 0, tcount = 0;
      const int n_items = kWork ? *p.n_work : p.n_items;
      for (int item = unit; item < n_items; item += n_units) {
        int t0, t1;
        if (kWork) {
          const int4 w = __ldg(p.work + item);
          t0 = 0; t1 = (w.z - w.y + kBN - 1) / kBN;
        } else {
          t0 = (item / p.n_qblocks) * p.tiles_per_split;
          t1 = min(t0 + p.tiles_per_split, p.tiles_total);
        }
        for (int t = t0; t < t1; ++t, ++tcount) {
          const uint32_t as = tcount & 1u, aph = (tcount >> 1) & 1u;
          ptx::mbar_wait(bar_acc_empty + 8 * as, aph ^ 1u);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + as * kBN;
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            ptx::mbar_wait(bar_full + 8 * stage, phase);
            ptx::tc_fence_after();
            const uint32_t a_addr = smem_base + stage * kStageBytes;
            const uint32_t b_addr = a_addr + Cfg::kABytes;
            // descriptors of the stage's first 16-wide K slice; +32 bytes per slice = +2 in the
            // (address >> 4) field (smem addresses stay below 2^18, no carry out of the field)
            const uint64_t adesc0 = ptx::make_kmajor_desc<kBK * 2>(a_addr);
            const uint64_t bdesc0 = ptx::make_kmajor_desc<kBK * 2>(b_addr);
            if (ptx::elect_one()) {
#pragma unroll
              for (int kk = 0; kk < kBK / 16; ++kk) {
                if (G == 2) ptx::umma_f16_2sm(d_tmem, adesc0 + 2u * kk, bdesc0 + 2u * kk, p.idesc, (kb | kk) != 0 ? 1u : 0u);
                else ptx::umma_f16(d_tmem, adesc0 + 2u * kk, bdesc0 + 2u * kk, p.idesc, (kb | kk) != 0 ? 1u : 0u);
              }
              // smem slot reusable (in both CTAs) once these MMAs retire
              if (G == 2) ptx::umma_commit_2sm(bar_empty + 8 * stage, 0x3);
              else ptx::umma_commit(bar_empty + 8 * stage);
              // last k-block: accumulator ready for the epilogue warps (of both CTAs)
              if (kb + 1 == p.k_blocks) {
                if (G == 2) ptx::umma_commit_2sm(bar_acc_full + 8 * as, 0x3);
                else ptx::umma_commit(bar_acc_full + 8 * as);
              }
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp & 3;                // TMEM lane quarter == warp % 4
    const uint32_t eg = static_cast<uint32_t>(warp - 4) >> 2;   // epilogue group = accumulator buffer
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
    const bool raw_out = !kWork && p.raw_count != nullptr;
    const uint32_t acc_empty_leader =
        (G == 2) ? ptx::mapa_cluster(bar_acc_empty, 0) : bar_acc_empty;
    const float inf = __int_as_float(0x7f800000);
    uint32_t tcount = 0;
    const int n_items = kWork ? *p.n_work : p.n_items;
    HitQueue hq;
    hq.keys = reinterpret_cast<u64*>(queue_mem + (warp - 4) * kQueueWarpBytes);
    hq.slots = reinterpret_cast<int*>(queue_mem + (warp - 4) * kQueueWarpBytes + kQueueCap * 8);
    hq.n = 0;
    for (int item = unit; item < n_items; item += n_units) {
      int qb, s = 0, t0, t1, row_begin = 0, row_end = 0;
      if (kWork) {
        const int4 w = __ldg(p.work + item);
        qb = w.x; row_begin = w.y; row_end = w.z; t0 = 0; t1 = (w.z - w.y + kBN - 1) / kBN;
      } else {
        qb = item % p.n_qblocks;
        s = item / p.n_qblocks;
        t0 = s * p.tiles_per_split;
        t1 = min(t0 + p.tiles_per_split, p.tiles_total);
      }
      // candidate buffers: the CTA's (re-used by its items), or - raw emission - the item's own
      const size_t cand_slot = raw_out ? (static_cast<size_t>(item) * G + cta_rank) * kEpiGroups + eg
                                       : static_cast<size_t>(blockIdx.x) * kEpiGroups + eg;
      u64* const cand_warp = p.cand + (cand_slot * kBM + ew * 32) * kCap;
      u64* const my_cand = cand_warp + static_cast<size_t>(lane) * kCap;
      size_t q_row = static_cast<size_t>(qb) * (kBM * G) + cta_rank * kBM + ew * 32 + lane;
      float tau = inf;
      int seed_slot = 0;
      bool real_row = true;
      bool quarter_live = true;   // work mode: false when none of the warp's 32 rows is a real query
      int tie_chunk = -1;         // (score, chunk) thresholds: chunks <= tie_chunk compare inclusively
      float tau_tie = -inf;
      if (kWork) {
        const int query = __ldg(p.row_query + q_row);
        if (p.seed_all == 1) seed_slot = __ldg(p.row_slot + q_row);
        tau = query >= 0 ? p.tau_init[query] : -inf;   // padding rows never qualify
        real_row = query >= 0;
        q_row = static_cast<size_t>(max(query, 0));
        quarter_live = __any_sync(0xffffffffu, real_row);
        if (p.tau_chunk != nullptr && real_row) {
          tie_chunk = __ldg(p.tau_chunk + query);
          tau_tie = nextafterf(tau, inf);
        }
      } else if (q_row >= static_cast<size_t>(p.nq)) {
        tau = -inf;            // padding row of the last query block: stays empty, costs nothing
      } else if (p.tau_init != nullptr) {
        tau = p.tau_init[q_row];
      }
      const int mode = (kWork || p.big_cand) ? kModeAppend : (p.k == 1 ? kModeArgmin : kModeBuffer);
      u64* const row_buf = (kWork || p.big_cand) ? p.big_cand + q_row * p.big_cap : my_cand;
      int* const row_cnt = (kWork || p.big_cand) ? p.big_count + q_row : nullptr;
      int cnt = 0;
      u64 best = kKeyInf;  // k == 1 fast path keeps the running arg-min in a register
      for (int ti = t0; ti < t1; ++ti, ++tcount) {
        const int t = ti * p.tile_stride;
        const uint32_t as = tcount & 1u, aph = (tcount >> 1) & 1u;
        if (kEpiGroups == 2 && as != eg) continue;   // the other group's accumulator buffer
        ptx::mbar_wait(bar_acc_full + 8 * as, aph);
        ptx::mbar_wait(bar_norm_full + 8 * as, aph);
        ptx::tc_fence_after();
        const float4* nrm4 = reinterpret_cast<const float4*>(norm_ptr + as * kBN);
        const uint32_t col0 = kWork ? static_cast<uint32_t>(row_begin + ti * kBN)
                                    : static_cast<uint32_t>(t) * kBN;
        // work mode: columns past the end of the list belong to the next list (lists are padded
        // to 32 rows, so validity is per 32-column chunk)
        const int nv = kWork ? row_end - static_cast<int>(col0) : kBN;
        if (kWork && !quarter_live) {   // nothing to score: keep the barrier protocol going
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            ptx::mbar_arrive(bar_acc_empty + 8 * as);
            ptx::mbar_arrive(bar_norm_empty + 8 * as);
          }
          continue;
        }
        // Two register buffers: the TMEM load of chunk c+1 is in flight while chunk c is scored.
        uint32_t ra[32], rb[32];
        const uint32_t tile_taddr = lane_taddr + as * kBN;
        ptx::tmem_ld_32x32b_x32(tile_taddr, ra);
#pragma unroll 1
        for (int c2 = 0; c2 < kBN / 64; ++c2) {
          ptx::tmem_ld_wait();
          ptx::tmem_ld_32x32b_x32(tile_taddr + c2 * 64 + 32, rb);
          if (kWork) {
            if (c2 * 64 < nv) {
              if (p.seed_all == 2) {
                const float mn = score_chunk_min(ra, nrm4 + c2 * 16, p.alpha);
                if (real_row) p.chunk_min[q_row * p.chunk_ld + ((col0 + c2 * 64) >> 5)] = mn;
              } else if (p.seed_all) {
                if (real_row)
                  score_chunk_seed(ra, nrm4 + c2 * 16, p.alpha, col0 + c2 * 64, 0.f,
                                   row_buf + seed_slot * kSeedSlotRows + (ti * kBN + c2 * 64));
              } else {
                const float t = static_cast<int>((col0 + c2 * 64) >> 5) <= tie_chunk ? tau_tie : tau;
                score_chunk_queue(ra, nrm4 + c2 * 16, p.alpha, col0 + c2 * 64, t, 0.f,
                                  static_cast<int>(q_row), hq, p.big_cand, p.big_count, p.big_cap, lane);
              }
            }
          } else if (mode == kModeArgmin)
            score_chunk<kModeArgmin>(ra, nrm4 + c2 * 16, p.alpha, col0 + c2 * 64, tau, cnt, best, row_buf);
          else if (mode == kModeAppend)
            score_chunk<kModeAppend>(ra, nrm4 + c2 * 16, p.alpha, col0 + c2 * 64, tau, cnt, best, row_buf,
                                     row_cnt, p.big_cap);
          else
            score_chunk<kModeBuffer>(ra, nrm4 + c2 * 16, p.alpha, col0 + c2 * 64, tau, cnt, best, row_buf);
          if (mode == kModeBuffer) compact_if_needed(cnt, tau, cand_warp, p.k, lane);
          ptx::tmem_ld_wait();
          if (c2 + 1 < kBN / 64) {
            ptx::tmem_ld_32x32b_x32(tile_taddr + c2 * 64 + 64, ra);
          } else {
            // every TMEM read of this tile has landed in registers: hand the accumulator back
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (G == 2) ptx::mbar_arrive_cluster(acc_empty_leader + 8 * as);
              else ptx::mbar_arrive(bar_acc_empty + 8 * as);
            }
          }
          if (kWork) {
            if (c2 * 64 + 32 < nv) {
              if (p.seed_all == 2) {
                const float mn = score_chunk_min(rb, nrm4 + c2 * 16 + 8, p.alpha);
                if (real_row) p.chunk_min[q_row * p.chunk_ld + ((col0 + c2 * 64 + 32) >> 5)] = mn;
              } else if (p.seed_all) {
                if (real_row)
                  score_chunk_seed(rb, nrm4 + c2 * 16 + 8, p.alpha, col0 + c2 * 64 + 32, 0.f,
                                   row_buf + seed_slot * kSeedSlotRows + (ti * kBN + c2 * 64 + 32));
              } else {
                const float t = static_cast<int>((col0 + c2 * 64 + 32) >> 5) <= tie_chunk ? tau_tie : tau;
                score_chunk_queue(rb, nrm4 + c2 * 16 + 8, p.alpha, col0 + c2 * 64 + 32, t, 0.f,
                                  static_cast<int>(q_row), hq, p.big_cand, p.big_count, p.big_cap, lane);
              }
            }
          } else if (mode == kModeArgmin)
            score_chunk<kModeArgmin>(rb, nrm4 + c2 * 16 + 8, p.alpha, col0 + c2 * 64 + 32, tau, cnt, best, row_buf);
          else if (mode == kModeAppend)
            score_chunk<kModeAppend>(rb, nrm4 + c2 * 16 + 8, p.alpha, col0 + c2 * 64 + 32, tau, cnt, best,
                                     row_buf, row_cnt, p.big_cap);
          else
            score_chunk<kModeBuffer>(rb, nrm4 + c2 * 16 + 8, p.alpha, col0 + c2 * 64 + 32, tau, cnt, best, row_buf);
          if (c2 + 1 == kBN / 64) {
            __syncwarp();   // all lanes are done with this tile's beta values
            if (lane == 0) ptx::mbar_arrive(bar_norm_empty + 8 * as);
          }
          if (mode == kModeBuffer) compact_if_needed(cnt, tau, cand_warp, p.k, lane);
        }
      }
      if (mode == kModeAppend) continue;  // candidates sit in the query's global buffer / the warp's queue
      if (p.debug_skip_emit) continue;
      if (raw_out && mode == kModeBuffer) {
        p.raw_count[cand_slot * kBM + ew * 32 + lane] = cnt;
        continue;
      }
      // ---- item done: emit this (split, query block)'s sorted top-k keys
      const size_t q_row0 = static_cast<size_t>(qb) * (kBM * G) + cta_rank * kBM + ew * 32;
      u64* out_blk = p.out_keys + ((static_cast<size_t>(s) * kEpiGroups + eg) * p.q_pad + q_row0) * p.k;
      if (mode == kModeArgmin) {
        out_blk[lane] = best;
      } else {
        __syncwarp();
        for (int rr = 0; rr < 32; ++rr) {
          const int n_r = __shfl_sync(0xffffffffu, cnt, rr);
          u64* dst = out_blk + static_cast<size_t>(rr) * p.k;
          if (n_r > 0) compact_row(cand_warp + static_cast<size_t>(rr) * kCap, dst, n_r, p.k, lane);
          for (int i = min(n_r, p.k) + lane; i < p.k; i += 32) dst[i] = kKeyInf;
        }
        __syncwarp();
      }
    }
    if (kWork) queue_drain(hq, p.big_cand, p.big_count, p.big_cap, lane);
  }

  __syncwarp();  // role branches above are per-lane: reconverge before the aligned barrier
  ptx::tc_fence_before();
  if (G == 2) ptx::cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    if (G == 2) ptx::tmem_dealloc_2sm(tmem_base, 512);
    else ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace b2vs
