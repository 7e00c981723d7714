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
// Warp roles (512 threads): warp 0 = TMA producer (query k-blocks + the tile's ||r^||^2 vector),
// warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 8-11 = decoders (one 32-row group each),
// warps 4-7 / 12-15 = two epilogue groups taking alternate tiles.  A smem stage = 16 KB decoded
// list k-block (128 rows x 64 dims) + 16 KB query k-block (TMA); its "full" barrier takes the TMA
// transaction plus one arrival per decoder warp.  bf16 codebooks of at most 64 KB stay resident in
// shared memory; larger ones (e.g. 768-d) are looked up in global memory, i.e. L2.
#pragma once
#include "bf_tc.cuh"

namespace b2vs {

constexpr int kPqTcThreads = 512;
constexpr int kPqM = 128;                                        // list rows per tile (UMMA M)
constexpr int kPqN = 128;                                        // query rows per block (UMMA N)
constexpr int kPqTcStages = 4;
constexpr int kPqListBytes = kPqM * kBK * 2;                     // 16 KB decoded list k-block
constexpr int kPqQueryBytes = kPqN * kBK * 2;                    // 16 KB query k-block
constexpr int kPqTcStageBytes = kPqListBytes + kPqQueryBytes;    // 32 KB
constexpr int kPqAcc = 4;                                        // accumulator buffers (4 x 128 TMEM columns)
constexpr int kPqNormBytes = kPqM * 4;                           // ||r^||^2 of a tile's 128 rows
constexpr int kPqTcMaxCbBytes = 64 * 1024;
constexpr int kPqTcQueueBytes = 8 * kQueueWarpBytes;             // one hit queue per epilogue warp (bf_tc.cuh)
constexpr int kPqColInfoBytes = kPqN * 16;                       // per epilogue warp: tau', bias, query, seed slot
constexpr int kPqTcSmemBytes = kPqTcStages * kPqTcStageBytes + kPqAcc * kPqNormBytes + 256 + kPqTcMaxCbBytes +
                               kPqTcQueueBytes + 8 * kPqColInfoBytes + 1024;
static_assert(kPqTcSmemBytes <= 227 * 1024, "pq_tc_kernel shared memory");
static_assert(kBK == 64, "the decoder writes 128-byte swizzled rows");
static_assert(kPqN == kBM, "query blocks are the gather kernels' 128-row groups");

struct PqTcParams {
  BfTcParams tc;            // work table, thresholds, append buffers (see bf_tc.cuh, work mode)
  const uint8_t* codes;     // PQ codes [n_groups][mp][32]: 32-row groups, sub-space major inside a group
  const uint32_t* cb16;     // bf16 codebooks, CODE-major [256][pq_dim][DSUB], viewed as 32-bit words
  const float* row_bias;    // [query rows] ||rq||^2 (L2) or -q.c_l (IP) of each gathered row
  int mp;                   // sub-spaces per row as stored (= pq_dim: the grouped scan needs pq_dim % 16 == 0)
  uint32_t n_groups;        // 32-row groups in `codes`
  int cb_words;             // pq_dim * 256 * DSUB / 2
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
  constexpr int kOffBar = kOffNorm + kPqAcc * kPqNormBytes;
  constexpr int kOffCb = kOffBar + 256;
  constexpr int kOffQueue = kOffCb + kPqTcMaxCbBytes;
  constexpr int kOffCol = kOffQueue + kPqTcQueueBytes;
  const uint32_t norm_base = smem_base + kOffNorm;
  const float* norm_ptr = reinterpret_cast<const float*>(smem + kOffNorm);
  const uint32_t bar_base = smem_base + kOffBar;
  const uint32_t bar_full = bar_base;                                // [kStages] TMA + decoders -> MMA
  const uint32_t bar_empty = bar_base + 8 * kStages;                 // [kStages] MMA -> TMA, decoders
  const uint32_t bar_acc_full = bar_base + 16 * kStages;             // [kPqAcc] MMA -> epilogue
  const uint32_t bar_acc_empty = bar_acc_full + 8 * kPqAcc;          // [kPqAcc] epilogue -> MMA
  const uint32_t bar_norm_full = bar_acc_full + 16 * kPqAcc;         // [kPqAcc] TMA -> epilogue
  const uint32_t bar_norm_empty = bar_acc_full + 24 * kPqAcc;        // [kPqAcc] epilogue -> TMA
  const uint32_t tmem_slot = bar_acc_full + 32 * kPqAcc;
  static_assert(16 * kStages + 32 * kPqAcc + 8 <= 256, "barrier block");
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + kOffBar + 16 * kStages + 32 * kPqAcc);
  uint32_t* cb_s = reinterpret_cast<uint32_t*>(smem + kOffCb);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int unit = static_cast<int>(blockIdx.x);
  const int n_units = static_cast<int>(gridDim.x);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(bar_full + 8 * i, 1 + 4);   // TMA arrive(+tx) and the four decoder warps
      ptx::mbar_init(bar_empty + 8 * i, 1);
    }
    for (int i = 0; i < kPqAcc; ++i) {
      ptx::mbar_init(bar_acc_full + 8 * i, 1);
      ptx::mbar_init(bar_acc_empty + 8 * i, 4);    // the four warps of the epilogue group that owns the tile
      ptx::mbar_init(bar_norm_full + 8 * i, 1);
      ptx::mbar_init(bar_norm_empty + 8 * i, 4);
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
    uint32_t stage = 0, phase = 0, tcount = 0;
    for (int item = unit; item < n_items; item += n_units) {
      const int4 w = __ldg(p.work + item);
      const int q_row0 = w.x * kPqN;
      const int t1 = (w.z - w.y + kPqM - 1) / kPqM;
      for (int ti = 0; ti < t1; ++ti, ++tcount) {
        const uint32_t ab = tcount & (kPqAcc - 1), aph = (tcount / kPqAcc) & 1u;
        ptx::mbar_wait(bar_norm_empty + 8 * ab, aph ^ 1u);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(bar_norm_full + 8 * ab, kPqNormBytes);
          ptx::bulk_load_1d(norm_base + ab * kPqNormBytes, p.beta + static_cast<size_t>(w.y + ti * kPqM),
                            kPqNormBytes, bar_norm_full + 8 * ab);
        }
        __syncwarp();
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
          if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(bar_full + 8 * stage, kPqQueryBytes);
            // the query block is re-read for every tile of the list: keep it in L2
            ptx::tma_load_2d_hint(smem_base + stage * kStageBytes + kPqListBytes, &tm_q, bar_full + 8 * stage,
                                  kb * kBK, q_row0, ptx::kEvictLast);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    uint32_t stage = 0, phase = 0, tcount = 0;
    for (int item = unit; item < n_items; item += n_units) {
      const int4 w = __ldg(p.work + item);
      const int t1 = (w.z - w.y + kPqM - 1) / kPqM;
      for (int t = 0; t < t1; ++t, ++tcount) {
        const uint32_t ab = tcount & (kPqAcc - 1), aph = (tcount / kPqAcc) & 1u;
        ptx::mbar_wait(bar_acc_empty + 8 * ab, aph ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + ab * kPqN;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          ptx::mbar_wait(bar_full + 8 * stage, phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = smem_base + stage * kStageBytes;           // decoded list rows: M operand
          const uint64_t adesc0 = ptx::make_kmajor_desc<kBK * 2>(a_addr);
          const uint64_t bdesc0 = ptx::make_kmajor_desc<kBK * 2>(a_addr + kPqListBytes);   // queries: N operand
          if (ptx::elect_one()) {
#pragma unroll
            for (int kk = 0; kk < kBK / 16; ++kk)
              ptx::umma_f16(d_tmem, adesc0 + 2u * kk, bdesc0 + 2u * kk, p.idesc, (kb | kk) != 0 ? 1u : 0u);
            ptx::umma_commit(bar_empty + 8 * stage);
            if (kb + 1 == p.k_blocks) ptx::umma_commit(bar_acc_full + 8 * ab);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= 8 && warp < 12) {
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
    const int dw = warp - 8;
    constexpr int LPR = 64 / DSUB;     // lanes per list row
    constexpr int RPI = 32 / LPR;      // list rows per warp instruction
    constexpr int WPC = DSUB / 2;      // 32-bit words per codebook entry
    const int sub = lane % LPR;
    const int rph = lane / LPR;
    const uint32_t wps = static_cast<uint32_t>(pp.mp) * WPC;        // words per code value in cb16t
    const uint32_t piece = static_cast<uint32_t>(sub) * WPC * 4u;   // lane's byte offset in a row line
    const uint32_t pc = piece >> 4, pw = piece & 15u;
    struct KbIter { int item, ti, t1, kb; uint32_t g0; };
    auto seek = [&](KbIter& it, int item) {   // first k-block of the first non-empty item >= item
      it.ti = 0; it.kb = 0; it.t1 = 0; it.g0 = 0;
      while (item < n_items) {
        const int4 w = __ldg(p.work + item);
        it.t1 = (w.z - w.y + kPqM - 1) / kPqM;
        it.g0 = static_cast<uint32_t>(w.y) >> 5;
        if (it.t1 > 0) break;
        item += n_units;
      }
      it.item = item;
    };
    auto step = [&](KbIter& it) {
      if (++it.kb == p.k_blocks) {
        it.kb = 0;
        if (++it.ti == it.t1) seek(it, it.item + n_units);
      }
    };
    // the 32-byte code piece of this warp's group for k-block `it`
    auto fetch = [&](const KbIter& it, uint4 (&cv)[2]) {
      cv[0] = make_uint4(0, 0, 0, 0);
      cv[1] = make_uint4(0, 0, 0, 0);
      const uint32_t g = it.g0 + static_cast<uint32_t>(it.ti) * 4u + static_cast<uint32_t>(dw);
      if (it.item < n_items && g < pp.n_groups) {
        const uint4* src = reinterpret_cast<const uint4*>(
            pp.codes + (static_cast<size_t>(g) * pp.mp + static_cast<size_t>(it.kb * LPR + sub)) * 32);
        cv[0] = __ldg(src);
        cv[1] = __ldg(src + 1);
      }
    };
    KbIter cur, pre;
    seek(cur, unit);
    pre = cur;
    uint4 cv0[2], cv1[2], cv2[2];
    fetch(pre, cv0); step(pre);
    fetch(pre, cv1); step(pre);
    uint32_t stage = 0, phase = 0;
    while (cur.item < n_items) {
      fetch(pre, cv2); step(pre);
      ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
      const uint32_t unit_base = smem_base + stage * kStageBytes + static_cast<uint32_t>(dw) * 4096u;  // 32 rows x 128 B
      const uint32_t* cb_kb = cb_w + static_cast<uint32_t>(cur.kb * LPR + sub) * WPC;
      const uint32_t w8[8] = {cv0[0].x, cv0[0].y, cv0[0].z, cv0[0].w, cv0[1].x, cv0[1].y, cv0[1].z, cv0[1].w};
      // The look-ups of a batch of rows are all issued before the first store of the batch: the
      // store asm statements are ordering points for the compiler.
      constexpr int kBatch = 16 / WPC;           // rows per batch: 16 registers of look-up results
#pragma unroll
      for (int r0 = 0; r0 < 32; r0 += RPI * kBatch) {
        uint32_t val[kBatch][WPC];
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
          const int r = r0 + b * RPI;
          const uint32_t code = (w8[r >> 2] >> (8u * ((r & 3) + rph))) & 0xFFu;
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
          const uint32_t rr = static_cast<uint32_t>(r0 + b * RPI) + static_cast<uint32_t>(rph);   // row in the group
          const uint32_t swz = rr & 7u;
          const uint32_t dst = unit_base + (rr >> 3) * 1024u + swz * 128u + (((pc ^ swz) << 4) | pw);
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
      if (++stage == kStages) { stage = 0; phase ^= 1u; }
      cv0[0] = cv1[0]; cv0[1] = cv1[1];
      cv1[0] = cv2[0]; cv1[1] = cv2[1];
      step(cur);
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (append mode)
    // Thread = one list row of the tile (TMEM lane 32 * ew + lane), columns = the group's queries.
    // The two groups of four warps (4-7, 12-15) take alternate tiles.  Per work item every warp
    // builds its own copy of the column table (threshold minus bias, bias, query slot, seed slot of
    // each of the block's 128 query rows) in shared memory; only the 16-column units that hold real
    // queries are read back from TMEM.
    const int ew = warp & 3;
    const uint32_t eg = warp >= 12 ? 1u : 0u;
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
    const float inf = __int_as_float(0x7f800000);
    const int wi = ew + 4 * static_cast<int>(eg);
    HitQueue hq;
    hq.keys = reinterpret_cast<u64*>(smem + kOffQueue + wi * kQueueWarpBytes);
    hq.slots = reinterpret_cast<int*>(smem + kOffQueue + wi * kQueueWarpBytes + kQueueCap * 8);
    hq.n = 0;
    float* const ci_tau = reinterpret_cast<float*>(smem + kOffCol + wi * kPqColInfoBytes);
    float* const ci_bias = ci_tau + kPqN;
    int* const ci_q = reinterpret_cast<int*>(ci_bias + kPqN);
    int* const ci_slot = ci_q + kPqN;
    uint32_t tcount = 0;
    for (int item = unit; item < n_items; item += n_units) {
      const int4 w = __ldg(p.work + item);
      const int row_begin = w.y, row_end = w.z;
      const int t1 = (w.z - w.y + kPqM - 1) / kPqM;
      // ---- column table of this item's query block (real queries are contiguous from column 0)
      __syncwarp();
      int n_real = 0;
#pragma unroll
      for (int j = 0; j < kPqN / 32; ++j) {
        const int col = lane + 32 * j;
        const size_t v_row = static_cast<size_t>(w.x) * kPqN + col;
        const int query = __ldg(p.row_query + v_row);
        float tq = -inf, bias = 0.f;
        if (query >= 0) {
          bias = __ldg(pp.row_bias + v_row);
          const float t = p.tau_init[query];
          // The threshold is on the full score (bias + alpha*acc + beta); the tile part is compared
          // against tau - bias, widened by a few ulps of the larger magnitude so that a key whose
          // rounded sum (v + bias) lies at the threshold is never lost to the rounding of
          // (tau - bias).  A slightly larger candidate set is harmless: the select step is exact.
          tq = (t - bias) + 4.f * 1.1920929e-7f * fmaxf(fabsf(t), fabsf(bias));
        }
        ci_tau[col] = tq;
        ci_bias[col] = bias;
        ci_q[col] = max(query, 0);
        ci_slot[col] = p.seed_all ? __ldg(p.row_slot + v_row) : 0;
        n_real += __popc(__ballot_sync(0xffffffffu, query >= 0));
      }
      __syncwarp();
      const int n_col_units = (n_real + 15) >> 4;
      for (int ti = 0; ti < t1; ++ti, ++tcount) {
        if ((tcount & 1u) != eg) continue;          // the other epilogue group's tile
        const uint32_t ab = tcount & (kPqAcc - 1), aph = (tcount / kPqAcc) & 1u;
        ptx::mbar_wait(bar_acc_full + 8 * ab, aph);
        ptx::mbar_wait(bar_norm_full + 8 * ab, aph);
        ptx::tc_fence_after();
        const int slot = row_begin + ti * kPqM + ew * 32 + lane;     // this thread's list slot
        // rows past the end of the list (tile tail) belong to the next list: they never qualify
        const float beta_row = slot < row_end ? norm_ptr[ab * kPqM + ew * 32 + lane] : inf;
        const uint32_t tile_taddr = lane_taddr + ab * kPqN;
        for (int u = 0; u < n_col_units; ++u) {
          uint32_t acc[16];
          ptx::tmem_ld_32x32b_x16(tile_taddr + u * 16, acc);
          ptx::tmem_ld_wait();
          if (u + 1 == n_col_units) {
            // this warp's last TMEM read of the tile has landed: hand the accumulator back
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              ptx::mbar_arrive(bar_acc_empty + 8 * ab);
              ptx::mbar_arrive(bar_norm_empty + 8 * ab);
            }
          }
          float s[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) s[j] = fmaf(p.alpha, __uint_as_float(acc[j]), beta_row);
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
          if (!__any_sync(0xffffffffu, mask != 0u)) continue;      // the common case
          const int nh = __popc(mask);
          int inc = nh;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
          }
          const int total = __shfl_sync(0xffffffffu, inc, 31);
          const bool direct = total > kQueueCap;     // a flood (loose threshold): straight to global
          if (!direct && hq.n + total > kQueueCap) queue_drain(hq, p.big_cand, p.big_count, p.big_cap, lane);
          int pos = hq.n + inc - nh;
          while (mask) {
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
            const u64 key = pack_key(v + ci_bias[col], static_cast<uint32_t>(slot));
            const int qs = ci_q[col];
            if (direct) {
              const int gp = atomicAdd(p.big_count + qs, 1);
              if (gp < p.big_cap) __stcg(p.big_cand + static_cast<size_t>(qs) * p.big_cap + gp, key);
            } else {
              hq.keys[pos] = key;
              hq.slots[pos] = qs;
              ++pos;
            }
          }
          if (!direct) hq.n += total;
        }
        if (n_col_units == 0) {      // a block without a real query (never planned, but keep the protocol sound)
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            ptx::mbar_arrive(bar_acc_empty + 8 * ab);
            ptx::mbar_arrive(bar_norm_empty + 8 * ab);
          }
        }
      }
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
