// K7b — grouped IVF-PQ list scan on the tcgen05 tensor cores (sm_100a).
//
// Replaces the ADC look-up-table scan behind `cuvs.neighbors.ivf_pq.search`
// (index_building_coordinator.py:398-404, improved_multi_gpu_rag.py:131-138 + :225-233) for large
// batches.  ADC distance of query q to a row of list l with PQ code c:
//     L2:  ||(q - c_l) - r^(c)||^2 = ||rq||^2 - 2 rq.r^ + ||r^||^2        rq = q - c_l
//     IP:  -(q.c_l + q.r^)
// r^ = the row's decoded residual (concatenated codebook entries).  Instead of one look-up table
// per (query, list) and pq_dim shared-memory look-ups per (query, row), the (query, probe) items
// are grouped by list (ivf_plan.cu) and each list is DECODED ONCE per batch into a bf16 K-major
// tile in shared memory — by four decoder warps, straight into the 128-byte-swizzled layout the UMMA
// descriptors expect — and multiplied against the block of residual queries that probe the list.
//
// Round-2 layout: the LIST ROWS are the UMMA M dimension (128 TMEM lanes = 128 list rows), the
// queries of the group are the N dimension (columns).  At C4 a list is probed by ~40 queries: as
// the M operand they filled 30 % of a 128-row block and the epilogue still scored all 128 x 256
// entries of every tile; as columns only the ceil(c / 16) 16-column units that hold real queries
// are read back from TMEM at all, every epilogue lane owns a real list row (no idle lane quarters,
// no straggler warp), and a tile's accumulator is 128 columns, so four of them are in flight.
//
// The N extent follows the block: a work item carries the number of real query rows c of its block
// (work.w); the TMA producer loads ceil(c / 16) 16-row boxes of the query k-block instead of all 128
// rows, and the UMMAs run with N = 16 ceil(c / 16) - at C4 (c ~ 40) that is 3/8 of the tensor-pipe
// time, of the shared-memory reads of the N operand and of the TMA traffic of a full block.
//
// Resident query blocks (dim <= 128, i.e. at most two k-blocks): the query block of a work item
// is loaded ONCE (double-buffered across items) instead of once per tile, and the stages hold only
// the decoded list k-blocks.  (Measured neutral at C4 - the per-step loads were hidden - but it
// removes 2/3 of the kernel's TMA traffic.)
//
// Warp roles (768 threads):
//   warp 0        TMA producer: query blocks / k-blocks, the tile's ||r^||^2 vector (a ring of 16:
//                 those loads come from DRAM), and an L2 prefetch of the tile's codes
//   warp 1        MMA issuer
//   warp 2        TMEM allocator
//   warp 3        column-table builder: per work item the threshold-minus-bias, bias, query slot and
//                 seed slot of the block's 128 query rows, for all epilogue warps, a ring of four
//                 items ahead (each epilogue warp used to build its own copy at the start of every
//                 item: two dependent L2 round trips on its critical path)
//   warps 8-11, 16-19    two decoder sets taking alternate k-block stages; a warp owns one 32-row
//                 group of the stage
//   warps 4-7, 12-15, 20-23    three epilogue groups taking tiles round-robin
// A stage's "full" barrier takes one arrival per decoder warp of the set (plus the TMA transaction
// when query k-blocks are streamed).  bf16 codebooks of at most 64 KB stay resident in shared
// memory; larger ones (e.g. 768-d) are looked up in global memory, i.e. L2.
//
// Where the time goes (C4 main pass, B2VS_PQ_DEBUG role-skipping runs, profiles/r2_c4_pq_tc_roles.jsonl):
// bare pipeline 0.35 ms (~245 ns per k-block step; neither a deeper register prefetch of the codes
// nor whole-tile stages - one hand-off per tile - nor prefetched work entries moved it: DESIGN.md §9),
// + UMMAs 0.06, + decode 0.08, + TMEM loads / scoring / hit masks 0.14, + hit queueing 0.14 = 0.78 ms.
#pragma once
#include "bf_tc.cuh"

namespace b2vs {

constexpr int kPqTcThreads = 768;
constexpr int kPqM = 128;                                        // list rows per tile (UMMA M)
constexpr int kPqN = 128;                                        // query rows per block (UMMA N)
constexpr int kPqTcStages = 4;
constexpr int kPqListBytes = kPqM * kBK * 2;                     // 16 KB decoded list k-block
constexpr int kPqQueryBytes = kPqN * kBK * 2;                    // 16 KB query k-block
constexpr int kPqBoxRows = 16;                                   // TMA box: 16 query rows x 64 dims = one UMMA N unit
constexpr int kPqBoxBytes = kPqBoxRows * kBK * 2;                // 2 KB
constexpr int kPqTcStageBytes = kPqListBytes + kPqQueryBytes;    // 32 KB
constexpr int kPqAcc = 4;                                        // accumulator buffers (4 x 128 TMEM columns)
constexpr int kPqNormBytes = kPqM * 4;                           // ||r^||^2 of a tile's 128 rows
constexpr int kPqNormRing = 16;                                  // norm vectors in flight: the loads come from DRAM (~1.5 us),
                                                                 // four in flight capped the kernel at ~0.5 us per tile
constexpr int kPqTcMaxCbBytes = 64 * 1024;
// Warps 4-7, 12-15 and 20-23 are the epilogue groups, warps 8-11 and 16-19 the two decoder sets.
// Measured at C4 (main scan launch): 2 sets + 2 groups 0.85 ms, 1 set + 3 groups 0.82 ms (without
// the epilogue: 0.55 ms with two sets, 0.67 ms with one - both roles sit near the critical path).
constexpr int kPqEpiGroups = 3;
constexpr int kPqDecSets = 2;
constexpr int kPqTcQueueBytes = 4 * kPqEpiGroups * kQueueWarpBytes;   // one hit queue per epilogue warp (bf_tc.cuh)
constexpr int kPqColInfoBytes = kPqN * 16;                       // column table of one query block: tau', bias, query, seed slot
constexpr int kPqTabRing = 4;                                    // column tables in flight (built by warp 3, items ahead)
constexpr int kPqTcSmemBytes = kPqTcStages * kPqTcStageBytes + kPqNormRing * kPqNormBytes + 512 + kPqTcMaxCbBytes +
                               kPqTcQueueBytes + kPqTabRing * kPqColInfoBytes + 1024;
static_assert(kPqTcSmemBytes <= 227 * 1024, "pq_tc_kernel shared memory");
static_assert(kBK == 64, "the decoder writes 128-byte swizzled rows");
static_assert(kPqN == kBM, "query blocks are the gather kernels' 128-row groups");
static_assert(kPqTcStages % kPqDecSets == 0, "decoder sets take alternate stages");

struct PqTcParams {
  BfTcParams tc;            // work table, thresholds, append buffers (see bf_tc.cuh, work mode)
  const uint8_t* codes;     // PQ codes [n_groups][mp][32]: 32-row groups, sub-space major inside a group
  const uint32_t* cb16;     // bf16 codebooks, CODE-major [256][pq_dim][DSUB], viewed as 32-bit words
  const float* row_bias;    // [query rows] ||rq||^2 (L2) or -q.c_l (IP) of each gathered row
  int mp;                   // sub-spaces per row as stored (= pq_dim: the grouped scan needs pq_dim % 16 == 0)
  uint32_t n_groups;        // 32-row groups in `codes`
  int cb_words;             // pq_dim * 256 * DSUB / 2
  int qres;                 // 1: the query block of a work item stays resident in shared memory for all of the
                            // item's tiles (dim <= 128); 0: its k-blocks are re-loaded with every tile
  int debug;                // B2VS_PQ_DEBUG measurement bits: 1 = epilogue only releases the accumulators,
                            // 2 = decoders skip look-ups and stores, 4 = no UMMAs (results are garbage)
};

// DSUB = sub-vector length (2, 4 or 8: a code decodes to 4, 8 or 16 bytes of bf16).
// kCbSmem: the bf16 codebooks (dim * 512 bytes) fit in 64 KB and stay resident in shared memory;
// otherwise (e.g. 768-d) the decoders look entries up in global memory, where L2 holds them.
template <int DSUB, bool kCbSmem>
__global__ void __launch_bounds__(kPqTcThreads, 1)
pq_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const PqTcParams pp) {
  static_assert(DSUB == 2 || DSUB == 4 || DSUB == 8, "DSUB");
  const BfTcParams& p = pp.tc;
  constexpr int kStages = kPqTcStages;
  constexpr int kStageBytes = kPqTcStageBytes;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  const uint32_t smem_base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - raw_addr);

  constexpr int kOffNorm = kStages * kStageBytes;
  constexpr int kOffBar = kOffNorm + kPqNormRing * kPqNormBytes;
  constexpr int kOffCb = kOffBar + 512;
  constexpr int kOffQueue = kOffCb + kPqTcMaxCbBytes;
  constexpr int kOffCol = kOffQueue + kPqTcQueueBytes;
  const uint32_t norm_base = smem_base + kOffNorm;
  const float* norm_ptr = reinterpret_cast<const float*>(smem + kOffNorm);
  const uint32_t bar_base = smem_base + kOffBar;
  const uint32_t bar_full = bar_base;                                // [kStages] TMA + decoders -> MMA
  const uint32_t bar_empty = bar_base + 8 * kStages;                 // [kStages] MMA -> TMA, decoders
  const uint32_t bar_acc_full = bar_base + 16 * kStages;             // [kPqAcc] MMA -> epilogue
  const uint32_t bar_acc_empty = bar_acc_full + 8 * kPqAcc;          // [kPqAcc] epilogue -> MMA
  const uint32_t bar_norm_full = bar_base + 256;                     // [kPqNormRing] TMA -> epilogue
  const uint32_t bar_norm_empty = bar_norm_full + 8 * kPqNormRing;   // [kPqNormRing] epilogue -> TMA
  static_assert(16 * kPqNormRing <= 256, "norm barriers");
  const uint32_t bar_q_full = bar_acc_full + 16 * kPqAcc;            // [2] TMA -> MMA   (resident query blocks)
  const uint32_t bar_q_empty = bar_q_full + 16;                      // [2] MMA -> TMA
  const uint32_t tmem_slot = bar_q_full + 32;
  const uint32_t bar_tab_full = bar_base + 192;                      // [kPqTabRing] table builder -> epilogue warps
  const uint32_t bar_tab_empty = bar_tab_full + 8 * kPqTabRing;      // [kPqTabRing] epilogue warps -> table builder
  static_assert(16 * kPqTabRing <= 64 && 16 * kStages + 16 * kPqAcc + 32 + 8 <= 192, "barrier block");
  static_assert(16 * kStages + 16 * kPqAcc + 32 + 8 <= 256, "barrier block");
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + kOffBar + 16 * kStages + 16 * kPqAcc + 32);
  // resident mode: stages = decoded list k-blocks only (16 KB apart), the two query buffers behind them
  const bool qres = pp.qres != 0;
  const uint32_t stage_stride = qres ? kPqListBytes : kStageBytes;
  const uint32_t qbuf_base = smem_base + kStages * kPqListBytes;
  constexpr uint32_t kQBufBytes = 2 * kPqQueryBytes;                 // two k-blocks of 128 query rows
  static_assert(kPqTcStages * kPqListBytes + 2 * 2 * kPqQueryBytes <= kPqTcStages * kPqTcStageBytes,
                "resident query buffers fit in the streaming stages' footprint");
  uint32_t* cb_s = reinterpret_cast<uint32_t*>(smem + kOffCb);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int unit = static_cast<int>(blockIdx.x);
  const int n_units = static_cast<int>(gridDim.x);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(bar_full + 8 * i, qres ? 4 : 1 + 4);   // (TMA arrive(+tx) and) the four decoder warps of a set
      ptx::mbar_init(bar_empty + 8 * i, 1);
    }
    for (int i = 0; i < kPqNormRing; ++i) {
      ptx::mbar_init(bar_norm_full + 8 * i, 1);
      ptx::mbar_init(bar_norm_empty + 8 * i, 4);    // the four warps of the epilogue group that owns the tile
    }
    for (int i = 0; i < kPqTabRing; ++i) {
      ptx::mbar_init(bar_tab_full + 8 * i, 1);
      ptx::mbar_init(bar_tab_empty + 8 * i, 4 * kPqEpiGroups);   // every epilogue warp
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(bar_q_full + 8 * i, 1);
      ptx::mbar_init(bar_q_empty + 8 * i, 1);
    }
    for (int i = 0; i < kPqAcc; ++i) {
      ptx::mbar_init(bar_acc_full + 8 * i, 1);
      ptx::mbar_init(bar_acc_empty + 8 * i, 4);    // the four warps of the epilogue group that owns the tile
    }
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&tm_q);
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  if (kCbSmem)
    for (int i = threadIdx.x; i < pp.cb_words; i += kPqTcThreads) cb_s[i] = __ldg(pp.cb16 + i);
  const uint32_t* cb_w = kCbSmem ? cb_s : pp.cb16;   // codebook words the decoders read
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int n_items = *p.n_work;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    uint32_t stage = 0, phase = 0, tcount = 0, icount = 0;
    for (int item = unit; item < n_items; item += n_units) {
      const int4 w = __ldg(p.work + item);
      const int q_row0 = w.x * kPqN;
      const int t1 = (w.z - w.y + kPqM - 1) / kPqM;
      if (t1 <= 0) continue;
      const int n_box = (max(w.w, 1) + 15) >> 4;          // 16-row boxes that hold real queries
      if (qres) {
        // the item's whole query block, once: buffer icount & 1 (the MMA warp frees it after the item's last tile)
        const uint32_t ib = icount & 1u, iph = (icount >> 1) & 1u;
        ++icount;
        ptx::mbar_wait(bar_q_empty + 8 * ib, iph ^ 1u);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(bar_q_full + 8 * ib, static_cast<uint32_t>(p.k_blocks * n_box) * kPqBoxBytes);
          for (int kb = 0; kb < p.k_blocks; ++kb)
            for (int b = 0; b < n_box; ++b)
              ptx::tma_load_2d_hint(qbuf_base + ib * kQBufBytes + kb * kPqQueryBytes + b * kPqBoxBytes, &tm_q,
                                    bar_q_full + 8 * ib, kb * kBK, q_row0 + b * kPqBoxRows, ptx::kEvictLast);
        }
        __syncwarp();
      }
      for (int ti = 0; ti < t1; ++ti, ++tcount) {
        const uint32_t nb = tcount & (kPqNormRing - 1), nph = (tcount / kPqNormRing) & 1u;
        ptx::mbar_wait(bar_norm_empty + 8 * nb, nph ^ 1u);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(bar_norm_full + 8 * nb, kPqNormBytes);
          ptx::bulk_load_1d(norm_base + nb * kPqNormBytes, p.beta + static_cast<size_t>(w.y + ti * kPqM),
                            kPqNormBytes, bar_norm_full + 8 * nb);
        }
        // The tile's PQ codes (4 groups x mp x 32 bytes, contiguous) into L2 now: this warp runs a
        // norm ring ahead of the epilogue, i.e. several tiles ahead of the decoders, whose own
        // register prefetch (two k-block steps) does not cover the DRAM latency.
        {
          const size_t tile_off = (static_cast<size_t>(static_cast<uint32_t>(w.y) >> 5) + static_cast<size_t>(ti) * 4u) *
                                  static_cast<size_t>(pp.mp) * 32u;
          const size_t end_off = static_cast<size_t>(pp.n_groups) * static_cast<size_t>(pp.mp) * 32u;
          const uint32_t tile_bytes = 4u * static_cast<uint32_t>(pp.mp) * 32u;
          for (uint32_t o = static_cast<uint32_t>(lane) * 128u; o < tile_bytes; o += 32u * 128u)
            if (tile_off + o < end_off) ptx::prefetch_l2(pp.codes + tile_off + o);
        }
        __syncwarp();
        if (qres) continue;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
          if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(bar_full + 8 * stage, static_cast<uint32_t>(n_box) * kPqBoxBytes);
            // the query block is re-read for every tile of the list: keep it in L2
            for (int b = 0; b < n_box; ++b)
              ptx::tma_load_2d_hint(smem_base + stage * kStageBytes + kPqListBytes + b * kPqBoxBytes, &tm_q,
                                    bar_full + 8 * stage, kb * kBK, q_row0 + b * kPqBoxRows, ptx::kEvictLast);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    uint32_t stage = 0, phase = 0, tcount = 0, icount = 0;
    for (int item = unit; item < n_items; item += n_units) {
      const int4 w = __ldg(p.work + item);
      const int t1 = (w.z - w.y + kPqM - 1) / kPqM;
      if (t1 <= 0) continue;
      // N extent = the block's real query rows, in 16-column units (instruction descriptor bits [17,23) = N >> 3)
      const uint32_t n_ext = static_cast<uint32_t>((max(w.w, 1) + 15) & ~15);
      const uint32_t idesc = (p.idesc & ~(0x3Fu << 17)) | ((n_ext >> 3) << 17);
      const uint32_t ib = icount & 1u, iph = (icount >> 1) & 1u;
      ++icount;
      if (qres) {
        ptx::mbar_wait(bar_q_full + 8 * ib, iph);
        ptx::tc_fence_after();
      }
      for (int t = 0; t < t1; ++t, ++tcount) {
        const uint32_t ab = tcount & (kPqAcc - 1), aph = (tcount / kPqAcc) & 1u;
        ptx::mbar_wait(bar_acc_empty + 8 * ab, aph ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + ab * kPqN;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          ptx::mbar_wait(bar_full + 8 * stage, phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = smem_base + stage * stage_stride;          // decoded list rows: M operand
          const uint32_t b_addr = qres ? qbuf_base + ib * kQBufBytes + kb * kPqQueryBytes : a_addr + kPqListBytes;
          const uint64_t adesc0 = ptx::make_kmajor_desc<kBK * 2>(a_addr);
          const uint64_t bdesc0 = ptx::make_kmajor_desc<kBK * 2>(b_addr);    // queries: N operand
          if (ptx::elect_one()) {
            if (!(pp.debug & 4))
#pragma unroll
            for (int kk = 0; kk < kBK / 16; ++kk)
              ptx::umma_f16(d_tmem, adesc0 + 2u * kk, bdesc0 + 2u * kk, idesc, (kb | kk) != 0 ? 1u : 0u);
            ptx::umma_commit(bar_empty + 8 * stage);
            if (kb + 1 == p.k_blocks) {
              ptx::umma_commit(bar_acc_full + 8 * ab);
              // the item's last MMAs: the query buffer is free once they retire
              if (qres && t + 1 == t1) ptx::umma_commit(bar_q_empty + 8 * ib);
            }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ column-table builder
    // Per work item: threshold minus bias, bias, query slot and seed slot of each of the block's 128
    // query rows, for ALL epilogue warps, kPqTabRing items ahead of them (each epilogue warp used to
    // build its own copy at the start of every item: two dependent L2 round trips, ~1.5 us, on the
    // epilogue's critical path once per ~6 tiles).
    const float inf = __int_as_float(0x7f800000);
    uint32_t icount = 0;
    for (int item = unit; item < n_items; item += n_units) {
      const int4 w = __ldg(p.work + item);
      if (w.z - w.y <= 0) continue;
      const uint32_t slot = icount & (kPqTabRing - 1), tph = (icount / kPqTabRing) & 1u;
      ++icount;
      ptx::mbar_wait(bar_tab_empty + 8 * slot, tph ^ 1u);
      float* const t_tau = reinterpret_cast<float*>(smem + kOffCol + slot * kPqColInfoBytes);
      float* const t_bias = t_tau + kPqN;
      int* const t_q = reinterpret_cast<int*>(t_bias + kPqN);
      int* const t_slot = t_q + kPqN;
      int query[kPqN / 32];
#pragma unroll
      for (int j = 0; j < kPqN / 32; ++j)
        query[j] = __ldg(p.row_query + static_cast<size_t>(w.x) * kPqN + lane + 32 * j);
#pragma unroll
      for (int j = 0; j < kPqN / 32; ++j) {
        const int col = lane + 32 * j;
        const size_t v_row = static_cast<size_t>(w.x) * kPqN + col;
        float tq = -inf, bias = 0.f;
        if (query[j] >= 0) {
          bias = __ldg(pp.row_bias + v_row);
          const float t = p.tau_init[query[j]];
          // The threshold is on the full score (bias + alpha*acc + beta); the tile part is compared
          // against tau - bias, widened by a few ulps of the larger magnitude so that a key whose
          // rounded sum (v + bias) lies at the threshold is never lost to the rounding of
          // (tau - bias).  A slightly larger candidate set is harmless: the select step is exact.
          tq = (t - bias) + 4.f * 1.1920929e-7f * fmaxf(fabsf(t), fabsf(bias));
        }
        t_tau[col] = tq;
        t_bias[col] = bias;
        t_q[col] = max(query[j], 0);
        t_slot[col] = (p.seed_all && query[j] >= 0) ? __ldg(p.row_slot + v_row) : 0;
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_tab_full + 8 * slot);
    }
  } else if ((warp >= 8 && warp < 12) || (warp >= 16 && warp < 20)) {
    // ------------------------------------------------------------------ decoders
    // A stage's list k-block = 128 rows x 64 dims = LPR = 64 / DSUB sub-spaces per row; decoder warp
    // dw owns the tile's 32-row group dw.  LANES WALK THE SUB-SPACES of one list row (DSUB 2: 32
    // lanes = one row per instruction; DSUB 4 / 8: two / four rows per instruction), so with the
    // code-major codebook copy cb16t[code][sub-space]
    //  * a look-up instruction touches 32 consecutive banks whatever the codes are (the row-per-lane
    //    layout of round 1 hit ~3.5-way bank conflicts on its random 4-byte reads), and
    //  * the decoded pieces of one instruction fill whole 128-byte swizzled row lines: the stores
    //    are conflict-free too.
    // Codes are stored sub-space major inside 32-row groups (pq_code_offset), so the 32 codes a
    // lane needs for one (group, k-block) unit are ONE 32-byte piece; the pieces of the k-blocks two
    // steps ahead are already in flight while a k-block is decoded.
    const int dw = warp & 3;               // 32-row group of the stage
    const int ds = warp >= 16 ? 1 : 0;     // decoder set: k-block steps ds, ds + kPqDecSets, ...
    constexpr int LPR = 64 / DSUB;     // lanes per list row
    constexpr int RPI = 32 / LPR;      // list rows per warp instruction
    constexpr int WPC = DSUB / 2;      // 32-bit words per codebook entry
    const int sub = lane % LPR;
    const int rph = lane / LPR;
    const uint32_t wps = static_cast<uint32_t>(pp.mp) * WPC;        // words per code value in cb16t
    const uint32_t piece = static_cast<uint32_t>(sub) * WPC * 4u;   // lane's byte offset in a row line
    const uint32_t pc = piece >> 4, pw = piece & 15u;
    // Byte offset of the lane's piece inside an 8-row swizzle atom, for the row (k + rph) & 7: the
    // rows a lane writes are r + rph with r a multiple of RPI, so k = r & 7 is a compile-time
    // constant at every store and the stores need no address arithmetic beyond base + immediate.
    uint32_t lane_swz[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t sw = static_cast<uint32_t>(k + rph) & 7u;
      lane_swz[k] = sw * 128u + (((pc ^ sw) << 4) | pw);
    }
    // ONE iterator walks this set's k-block steps (every other step of the CTA's sequence) two
    // pieces ahead of the decode; a fetched piece carries its k-block index (-1 = past the end).
    struct KbIter { int item, ti, t1, kb; uint32_t g0; };
    KbIter it;
    auto seek = [&](int item) {               // first non-empty item >= item
      it.ti = 0; it.t1 = 0; it.g0 = 0;
      while (item < n_items) {
        const int4 w = __ldg(p.work + item);
        it.t1 = (w.z - w.y + kPqM - 1) / kPqM;
        it.g0 = static_cast<uint32_t>(w.y) >> 5;
        if (it.t1 > 0) break;
        item += n_units;
      }
      it.item = item;
    };
    auto advance = [&](int steps) {
      it.kb += steps;
      while (it.item < n_items && it.kb >= p.k_blocks) {
        it.kb -= p.k_blocks;
        if (++it.ti == it.t1) seek(it.item + n_units);
      }
    };
    struct Piece { uint4 c0, c1; int kb; };
    // the 32-byte code piece of this warp's group for the iterator's k-block, then two steps on
    auto fetch = [&](Piece& pcs) {
      pcs.c0 = make_uint4(0, 0, 0, 0);
      pcs.c1 = make_uint4(0, 0, 0, 0);
      pcs.kb = -1;
      if (it.item >= n_items) return;
      pcs.kb = it.kb;
      const uint32_t g = it.g0 + static_cast<uint32_t>(it.ti) * 4u + static_cast<uint32_t>(dw);
      if (g < pp.n_groups) {
        const uint4* src = reinterpret_cast<const uint4*>(
            pp.codes + (static_cast<size_t>(g) * pp.mp + static_cast<size_t>(it.kb * LPR + sub)) * 32);
        pcs.c0 = __ldg(src);
        pcs.c1 = __ldg(src + 1);
      }
      advance(kPqDecSets);
    };
    it.kb = 0;
    seek(unit);
    if (ds) advance(1);
    Piece p0, p1, p2;
    fetch(p0);
    fetch(p1);
    uint32_t stage = static_cast<uint32_t>(ds), phase = 0;
    while (p0.kb >= 0) {
      fetch(p2);
      ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
      const uint32_t unit_base = smem_base + stage * stage_stride + static_cast<uint32_t>(dw) * 4096u;  // 32 rows x 128 B
      uint32_t base8[8];
#pragma unroll
      for (int k = 0; k < 8; k += RPI) base8[k] = unit_base + lane_swz[k];
      const uint32_t* cb_kb = cb_w + static_cast<uint32_t>(p0.kb * LPR + sub) * WPC;
      const uint32_t cb_sa = smem_base + kOffCb + static_cast<uint32_t>(p0.kb * LPR + sub) * WPC * 4u;
      const uint32_t wps4 = wps * 4u;
      const uint32_t w8[8] = {p0.c0.x, p0.c0.y, p0.c0.z, p0.c0.w, p0.c1.x, p0.c1.y, p0.c1.z, p0.c1.w};
      // The look-ups of a batch of rows are all issued before the first store of the batch: the
      // store asm statements are ordering points for the compiler.
      constexpr int kBatch = 16 / WPC;           // rows per batch: 16 registers of look-up results
      if (!(pp.debug & 2))
#pragma unroll
      for (int r0 = 0; r0 < 32; r0 += RPI * kBatch) {
        uint32_t val[kBatch][WPC];
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
          const int r = r0 + b * RPI;
          const uint32_t code = __byte_perm(w8[r >> 2], 0u, 0x4440u | static_cast<uint32_t>((r & 3) + rph));   // byte (r & 3) + rph
          if (kCbSmem) {      // resident codebooks: one IMAD gives the shared-space byte address
            const uint32_t a = cb_sa + code * wps4;
            if (WPC == 1) {
              asm volatile("ld.shared.b32 %0, [%1];" : "=r"(val[b][0]) : "r"(a));
            } else if (WPC == 2) {
              asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(val[b][0]), "=r"(val[b][WPC - 1]) : "r"(a));
            } else {
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(val[b][0]), "=r"(val[b][1 % WPC]), "=r"(val[b][2 % WPC]), "=r"(val[b][3 % WPC])
                           : "r"(a));
            }
            continue;
          }
          const uint32_t* src = cb_kb + code * wps;
          if (WPC == 1) {
            val[b][0] = src[0];
          } else if (WPC == 2) {
            const uint2 v = *reinterpret_cast<const uint2*>(src);
            val[b][0] = v.x; val[b][WPC - 1] = v.y;
          } else {
            const uint4 v = *reinterpret_cast<const uint4*>(src);
            val[b][0] = v.x; val[b][1 % WPC] = v.y; val[b][2 % WPC] = v.z; val[b][3 % WPC] = v.w;
          }
        }
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
          const int r = r0 + b * RPI;                 // compile-time: the lane writes row r + rph of the group
          const uint32_t dst = base8[r & 7] + static_cast<uint32_t>(r >> 3) * 1024u;   // base + immediate
          if (WPC == 1) {
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst), "r"(val[b][0]));
          } else if (WPC == 2) {
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst), "r"(val[b][0]), "r"(val[b][WPC - 1]));
          } else {
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(val[b][0]),
                         "r"(val[b][1 % WPC]), "r"(val[b][2 % WPC]), "r"(val[b][3 % WPC]));
          }
        }
      }
      // generic-proxy writes -> visible to the tensor core's async-proxy reads
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_full + 8 * stage);
      stage += kPqDecSets;
      if (stage >= kStages) { stage -= kStages; phase ^= 1u; }
      p0 = p1;
      p1 = p2;
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (append mode)
    // Thread = one list row of the tile (TMEM lane 32 * ew + lane), columns = the group's queries.
    // The groups of four warps take tiles round-robin.  The column table of the item's query block
    // (threshold minus bias, bias, query slot, seed slot of its 128 rows) comes from warp 3; only
    // the 16-column units that hold real queries are read back from TMEM.
    const int ew = warp & 3;
    const uint32_t eg = warp < 8 ? 0u : (warp < 16 ? 1u : 2u);   // warps 4-7, 12-15, 20-23
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
    const float inf = __int_as_float(0x7f800000);
    const int wi = ew + 4 * static_cast<int>(eg);
    HitQueue hq;
    hq.keys = reinterpret_cast<u64*>(smem + kOffQueue + wi * kQueueWarpBytes);
    hq.slots = reinterpret_cast<int*>(smem + kOffQueue + wi * kQueueWarpBytes + kQueueCap * 8);
    hq.n = 0;
    uint32_t tcount = 0, icount = 0;
    for (int item = unit; item < n_items; item += n_units) {
      const int4 w = __ldg(p.work + item);
      const int row_begin = w.y, row_end = w.z;
      const int t1 = (w.z - w.y + kPqM - 1) / kPqM;
      if (t1 <= 0) continue;
      const uint32_t tslot = icount & (kPqTabRing - 1), tph = (icount / kPqTabRing) & 1u;
      ++icount;
      ptx::mbar_wait(bar_tab_full + 8 * tslot, tph);
      const float* const ci_tau = reinterpret_cast<const float*>(smem + kOffCol + tslot * kPqColInfoBytes);
      const float* const ci_bias = ci_tau + kPqN;
      const int* const ci_q = reinterpret_cast<const int*>(ci_bias + kPqN);
      const int* const ci_slot = ci_q + kPqN;
      const int n_real = max(w.w, 0);            // real queries are contiguous from column 0
      const int n_col_units = (n_real + 15) >> 4;
      for (int ti = 0; ti < t1; ++ti, ++tcount) {
        if (tcount % kPqEpiGroups != eg) continue;  // another epilogue group's tile
        const uint32_t ab = tcount & (kPqAcc - 1), aph = (tcount / kPqAcc) & 1u;
        ptx::mbar_wait(bar_acc_full + 8 * ab, aph);
        const uint32_t nb = tcount & (kPqNormRing - 1), nph = (tcount / kPqNormRing) & 1u;
        ptx::mbar_wait(bar_norm_full + 8 * nb, nph);
        ptx::tc_fence_after();
        if (pp.debug & 1) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            ptx::mbar_arrive(bar_acc_empty + 8 * ab);
            ptx::mbar_arrive(bar_norm_empty + 8 * nb);
          }
          continue;
        }
        const int slot = row_begin + ti * kPqM + ew * 32 + lane;     // this thread's list slot
        // rows past the end of the list (tile tail) belong to the next list: they never qualify
        const float beta_row = slot < row_end ? norm_ptr[nb * kPqM + ew * 32 + lane] : inf;
        const uint32_t tile_taddr = lane_taddr + ab * kPqN;
        // the TMEM load of unit u + 1 is in flight while unit u is scored
        uint32_t acc[16];
        if (n_col_units > 0) ptx::tmem_ld_32x32b_x16(tile_taddr, acc);
        for (int u = 0; u < n_col_units; ++u) {
          ptx::tmem_ld_wait();
          float s[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) s[j] = fmaf(p.alpha, __uint_as_float(acc[j]), beta_row);
          if (u + 1 < n_col_units) {
            ptx::tmem_ld_32x32b_x16(tile_taddr + (u + 1) * 16, acc);
          } else {
            // this warp's last TMEM read of the tile has landed: hand the accumulator back
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              ptx::mbar_arrive(bar_acc_empty + 8 * ab);
              ptx::mbar_arrive(bar_norm_empty + 8 * nb);
            }
          }
          if (p.seed_all) {
            // seed pass: every (real query, list row) score goes to its fixed place
            if (slot < row_end) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int col = u * 16 + j;
                if (col < n_real)
                  __stcg(p.big_cand + static_cast<size_t>(ci_q[col]) * p.big_cap + ci_slot[col] * kSeedSlotRows +
                             (slot - row_begin),
                         pack_key(s[j] + ci_bias[col], static_cast<uint32_t>(slot)));
              }
            }
            continue;
          }
          const float4* t4 = reinterpret_cast<const float4*>(ci_tau + u * 16);
          uint32_t mask = 0;
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 tv = t4[j4];
            mask |= (s[4 * j4 + 0] < tv.x ? 1u : 0u) << (4 * j4 + 0);
            mask |= (s[4 * j4 + 1] < tv.y ? 1u : 0u) << (4 * j4 + 1);
            mask |= (s[4 * j4 + 2] < tv.z ? 1u : 0u) << (4 * j4 + 2);
            mask |= (s[4 * j4 + 3] < tv.w ? 1u : 0u) << (4 * j4 + 3);
          }
          // Hits are sparse (a few per 32 x 16 block of scores, rarely two in one lane): every round
          // takes each lane's lowest remaining hit, and the queue positions come from the vote of
          // the lanes that still have one - no prefix sum over the warp.  The first vote is the
          // "nothing here" test.
          if (pp.debug & 16) mask = 0u;
          while (true) {
            const bool has = mask != 0u;
            const uint32_t votes = __ballot_sync(0xffffffffu, has);
            if (votes == 0u) break;
            const int nb = __popc(votes);
            if (hq.n + nb > kQueueCap) queue_drain(hq, p.big_cand, p.big_count, p.big_cap, lane);
            if (has) {
              const int j = __ffs(mask) - 1;
              mask &= mask - 1;
              float v8[8], v4[4], v2[2];
#pragma unroll
              for (int i = 0; i < 8; ++i) v8[i] = (j & 1) ? s[2 * i + 1] : s[2 * i];
#pragma unroll
              for (int i = 0; i < 4; ++i) v4[i] = (j & 2) ? v8[2 * i + 1] : v8[2 * i];
#pragma unroll
              for (int i = 0; i < 2; ++i) v2[i] = (j & 4) ? v4[2 * i + 1] : v4[2 * i];
              const float v = (j & 8) ? v2[1] : v2[0];
              const int col = u * 16 + j;
              const int pos = hq.n + __popc(votes & ((1u << lane) - 1u));
              hq.keys[pos] = pack_key(v + ci_bias[col], static_cast<uint32_t>(slot));
              hq.slots[pos] = ci_q[col];
            }
            hq.n += nb;
          }
        }
        if (n_col_units == 0) {      // a block without a real query (never planned, but keep the protocol sound)
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            ptx::mbar_arrive(bar_acc_empty + 8 * ab);
            ptx::mbar_arrive(bar_norm_empty + 8 * nb);
          }
        }
      }
      __syncwarp();      // every lane is done with the item's column table
      if (lane == 0) ptx::mbar_arrive(bar_tab_empty + 8 * tslot);
    }
    queue_drain(hq, p.big_cand, p.big_count, p.big_cap, lane);
  }

  __syncwarp();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace b2vs
