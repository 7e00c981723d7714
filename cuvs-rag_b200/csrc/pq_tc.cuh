// K7b — grouped IVF-PQ list scan on the tcgen05 tensor cores (sm_100a).
//
// Replaces the ADC look-up-table scan behind `cuvs.neighbors.ivf_pq.search`
// (index_building_coordinator.py:398-404, improved_multi_gpu_rag.py:131-138 + :225-233) for large
// batches.  ADC distance of query q to a row of list l with PQ code c:
//     L2:  ||(q - c_l) - r^(c)||^2 = ||rq||^2 - 2 rq.r^ + ||r^||^2        rq = q - c_l
//     IP:  -(q.c_l + q.r^)
// r^ = the row's decoded residual (concatenated codebook entries).  Instead of one look-up table
// per (query, list) and pq_dim shared-memory look-ups per (query, row), the (query, probe) items
// are grouped by list (ivf.cu) and each list is DECODED ONCE per batch into a bf16 K-major tile
// in shared memory — by four decoder warps, straight into the 128-byte-swizzled layout the UMMA
// descriptors expect — and multiplied against the 128-row block of residual queries that probe
// the list.  The smem look-ups per row drop from (queries probing the list) x pq_dim to pq_dim,
// and with lanes walking the sub-spaces of a row over a code-major codebook copy every look-up
// and every store instruction is bank-conflict free (see the decoder role below).
//
// Same skeleton as bf_tc_kernel<1, true> (work-table + append mode): warp 0 = TMA producer (query
// k-blocks + the tile's ||r^||^2 vector), warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warps 4-7 and 12-15 = epilogue (column halves), warps 8-11 = decoders.  A smem stage = 16 KB query k-block (TMA) + 32 KB
// decoded list k-block (256 rows x 64 dims); its "full" barrier takes the TMA transaction plus
// one arrival per decoder warp.  bf16 codebooks of at most 64 KB stay resident in shared memory;
// larger ones (e.g. 768-d) are looked up in global memory, i.e. L2.
#pragma once
#include "bf_tc.cuh"

namespace b2vs {

constexpr int kPqTcThreads = 512;
constexpr int kPqTcStages = 3;
constexpr int kPqTcStageBytes = kBM * kBK * 2 + kBN * kBK * 2;   // 48 KB
constexpr int kPqTcMaxCbBytes = 64 * 1024;
constexpr int kPqTcQueueBytes = 8 * kQueueWarpBytes;   // one hit queue per epilogue warp (bf_tc.cuh)
constexpr int kPqTcSmemBytes =
    kPqTcStages * kPqTcStageBytes + 2 * kNormBytes + 256 + kPqTcMaxCbBytes + kPqTcQueueBytes + 1024;
static_assert(kPqTcSmemBytes <= 227 * 1024, "pq_tc_kernel shared memory");
static_assert(kBK == 64, "the decoder writes 128-byte swizzled rows");

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
  constexpr int kABytes = kBM * kBK * 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  const uint32_t smem_base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - raw_addr);

  const uint32_t norm_base = smem_base + kStages * kStageBytes;
  const float* norm_ptr = reinterpret_cast<const float*>(smem + kStages * kStageBytes);
  const uint32_t bar_base = norm_base + 2 * kNormBytes;
  const uint32_t bar_full = bar_base;                     // [kStages] TMA + decoders -> MMA
  const uint32_t bar_empty = bar_base + 8 * kStages;      // [kStages] MMA -> TMA, decoders
  const uint32_t bar_acc_full = bar_base + 16 * kStages;  // [2] MMA -> epilogue
  const uint32_t bar_acc_empty = bar_acc_full + 16;       // [2] epilogue -> MMA
  const uint32_t bar_norm_full = bar_acc_full + 32;       // [2] TMA -> epilogue
  const uint32_t bar_norm_empty = bar_acc_full + 48;      // [2] epilogue -> TMA
  const uint32_t tmem_slot = bar_acc_full + 64;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
      smem + kStages * kStageBytes + 2 * kNormBytes + 16 * kStages + 64);
  uint32_t* cb_s = reinterpret_cast<uint32_t*>(smem + kStages * kStageBytes + 2 * kNormBytes + 256);
  uint8_t* const queue_mem = smem + kStages * kStageBytes + 2 * kNormBytes + 256 + kPqTcMaxCbBytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int unit = static_cast<int>(blockIdx.x);
  const int n_units = static_cast<int>(gridDim.x);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(bar_full + 8 * i, 1 + 4);   // TMA arrive(+tx) and the four decoder warps
      ptx::mbar_init(bar_empty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(bar_acc_full + 8 * i, 1);
      ptx::mbar_init(bar_acc_empty + 8 * i, 8);    // eight epilogue warps
      ptx::mbar_init(bar_norm_full + 8 * i, 1);
      ptx::mbar_init(bar_norm_empty + 8 * i, 8);
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
    {
      uint32_t stage = 0, phase = 0, tcount = 0;
      for (int item = unit; item < n_items; item += n_units) {
        const int4 w = __ldg(p.work + item);
        const int q_row0 = w.x * kBM;
        const int t1 = (w.z - w.y + kBN - 1) / kBN;
        for (int ti = 0; ti < t1; ++ti, ++tcount) {
          const uint32_t as = tcount & 1u, aph = (tcount >> 1) & 1u;
          ptx::mbar_wait(bar_norm_empty + 8 * as, aph ^ 1u);
          if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(bar_norm_full + 8 * as, kNormBytes);
            ptx::bulk_load_1d(norm_base + as * kNormBytes, p.beta + static_cast<size_t>(w.y + ti * kBN),
                              kNormBytes, bar_norm_full + 8 * as);
          }
          __syncwarp();
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
            if (ptx::elect_one()) {
              ptx::mbar_arrive_expect_tx(bar_full + 8 * stage, kABytes);
              ptx::tma_load_2d_hint(smem_base + stage * kStageBytes, &tm_q, bar_full + 8 * stage,
                                    kb * kBK, q_row0, ptx::kEvictLast);
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    {
      uint32_t stage = 0, phase = 0, tcount = 0;
      for (int item = unit; item < n_items; item += n_units) {
        const int4 w = __ldg(p.work + item);
        const int t1 = (w.z - w.y + kBN - 1) / kBN;
        for (int t = 0; t < t1; ++t, ++tcount) {
          const uint32_t as = tcount & 1u, aph = (tcount >> 1) & 1u;
          ptx::mbar_wait(bar_acc_empty + 8 * as, aph ^ 1u);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + as * kBN;
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            ptx::mbar_wait(bar_full + 8 * stage, phase);
            ptx::tc_fence_after();
            const uint32_t a_addr = smem_base + stage * kStageBytes;
            const uint64_t adesc0 = ptx::make_kmajor_desc<kBK * 2>(a_addr);
            const uint64_t bdesc0 = ptx::make_kmajor_desc<kBK * 2>(a_addr + kABytes);
            if (ptx::elect_one()) {
#pragma unroll
              for (int kk = 0; kk < kBK / 16; ++kk)
                ptx::umma_f16(d_tmem, adesc0 + 2u * kk, bdesc0 + 2u * kk, p.idesc, (kb | kk) != 0 ? 1u : 0u);
              ptx::umma_commit(bar_empty + 8 * stage);
              if (kb + 1 == p.k_blocks) ptx::umma_commit(bar_acc_full + 8 * as);
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp >= 8 && warp < 12) {
    // ------------------------------------------------------------------ decoders
    // A stage's list k-block = 256 rows x 64 dims = LPR = 64 / DSUB sub-spaces per row.  LANES WALK
    // THE SUB-SPACES of one list row (DSUB 2: 32 lanes = one row per instruction; DSUB 4 / 8: two /
    // four rows per instruction), so with the code-major codebook copy cb16t[code][sub-space]
    //  * a look-up instruction touches 32 consecutive banks whatever the codes are (the row-per-lane
    //    layout of round 1 hit ~3.5-way bank conflicts on its random 4-byte reads), and
    //  * the decoded pieces of one instruction fill whole 128-byte swizzled row lines: the stores
    //    are conflict-free too.
    // Codes are stored sub-space major inside 32-row groups (pq_code_offset), so the 32 codes a
    // lane needs for one (group, k-block) unit are ONE 32-byte piece; the pieces of the k-blocks two
    // steps ahead are already in flight while a k-block is decoded (prefetch distance 2: the
    // exposed global-load latency of the decoders was the kernel's top stall).
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
        it.t1 = (w.z - w.y + kBN - 1) / kBN;
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
    // the two 32-byte code pieces (units = groups 2*dw, 2*dw + 1 of the tile) of k-block `it`
    auto fetch = [&](const KbIter& it, uint4 (&cv)[2][2]) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        cv[u][0] = make_uint4(0, 0, 0, 0);
        cv[u][1] = make_uint4(0, 0, 0, 0);
        const uint32_t g = it.g0 + static_cast<uint32_t>(it.ti) * 8u + static_cast<uint32_t>(2 * dw + u);
        if (it.item < n_items && g < pp.n_groups) {
          const uint4* src = reinterpret_cast<const uint4*>(
              pp.codes + (static_cast<size_t>(g) * pp.mp + static_cast<size_t>(it.kb * LPR + sub)) * 32);
          cv[u][0] = __ldg(src);
          cv[u][1] = __ldg(src + 1);
        }
      }
    };
    KbIter cur, pre;
    seek(cur, unit);
    pre = cur;
    uint4 cv0[2][2], cv1[2][2], cv2[2][2];
    fetch(pre, cv0); step(pre);
    fetch(pre, cv1); step(pre);
    uint32_t stage = 0, phase = 0;
    while (cur.item < n_items) {
      fetch(pre, cv2); step(pre);
      ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
      const uint32_t b_base = smem_base + stage * kStageBytes + kABytes;
      const uint32_t* cb_kb = cb_w + static_cast<uint32_t>(cur.kb * LPR + sub) * WPC;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const uint32_t w8[8] = {cv0[u][0].x, cv0[u][0].y, cv0[u][0].z, cv0[u][0].w,
                                cv0[u][1].x, cv0[u][1].y, cv0[u][1].z, cv0[u][1].w};
        const uint32_t unit_base = b_base + static_cast<uint32_t>(2 * dw + u) * 4096u;   // 32 rows x 128 B
        // The look-ups of a batch of rows are all issued before the first store of the batch: the
        // store asm statements are ordering points for the compiler, and one look-up latency per
        // row on the critical path (32 x 2 x ~40 cycles per k-block) was what the decoders cost.
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
      }
      // generic-proxy writes -> visible to the tensor core's async-proxy reads
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_full + 8 * stage);
      if (++stage == kStages) { stage = 0; phase ^= 1u; }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        cv0[u][0] = cv1[u][0]; cv0[u][1] = cv1[u][1];
        cv1[u][0] = cv2[u][0]; cv1[u][1] = cv2[u][1];
      }
      step(cur);
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (append mode)
    // Eight warps: warps 4-7 take columns 0-127 of a tile, warps 12-15 columns 128-255 (warp w may
    // only read TMEM lanes 32*(w%4)..+31, so both groups cover all four lane quarters).  With K as
    // short as 128 the MMA of a tile is over in ~1 us and this loop sets the pace; append mode
    // keeps no per-row state besides the threshold, so splitting a row's columns is free.
    const int ew = warp & 3;
    const int half = warp >= 12 ? 1 : 0;
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
    const float inf = __int_as_float(0x7f800000);
    uint32_t tcount = 0;
    HitQueue hq;
    {
      const int qi = ew + 4 * half;
      hq.keys = reinterpret_cast<u64*>(queue_mem + qi * kQueueWarpBytes);
      hq.slots = reinterpret_cast<int*>(queue_mem + qi * kQueueWarpBytes + kQueueCap * 8);
      hq.n = 0;
    }
    for (int item = unit; item < n_items; item += n_units) {
      const int4 w = __ldg(p.work + item);
      const int row_begin = w.y, row_end = w.z;
      const int t1 = (w.z - w.y + kBN - 1) / kBN;
      const size_t v_row = static_cast<size_t>(w.x) * kBM + ew * 32 + lane;
      const int query = __ldg(p.row_query + v_row);
      const float bias = query >= 0 ? __ldg(pp.row_bias + v_row) : 0.f;
      // The threshold is on the full score (bias + alpha*acc + beta); the tile part is compared
      // against tau - bias, widened by a few ulps of the larger magnitude so that a key whose
      // rounded sum (v + bias) lies at the threshold is never lost to the rounding of (tau - bias).
      // A slightly larger candidate set is harmless: the select step is exact.
      float tau = -inf;
      if (query >= 0) {
        const float tq = p.tau_init[query];
        tau = (tq - bias) + 4.f * 1.1920929e-7f * fmaxf(fabsf(tq), fabsf(bias));
      }
      const size_t qslot = static_cast<size_t>(max(query, 0));
      u64* const row_buf = p.big_cand + qslot * p.big_cap;
      const int seed_slot = p.seed_all ? __ldg(p.row_slot + v_row) : 0;
      // a quarter without a single real query row (group_row_pos packs the queries of a block into
      // as few quarters as possible) only keeps the barrier protocol going
      const bool quarter_live = __any_sync(0xffffffffu, query >= 0);
      for (int ti = 0; ti < t1; ++ti, ++tcount) {
        const uint32_t as = tcount & 1u, aph = (tcount >> 1) & 1u;
        ptx::mbar_wait(bar_acc_full + 8 * as, aph);
        ptx::mbar_wait(bar_norm_full + 8 * as, aph);
        ptx::tc_fence_after();
        const float4* nrm4 = reinterpret_cast<const float4*>(norm_ptr + as * kBN);
        const uint32_t col0 = static_cast<uint32_t>(row_begin + ti * kBN);
        const int nv = row_end - static_cast<int>(col0);
        const uint32_t tile_taddr = lane_taddr + as * kBN;
        uint32_t ra[32];
        if (!quarter_live) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            ptx::mbar_arrive(bar_acc_empty + 8 * as);
            ptx::mbar_arrive(bar_norm_empty + 8 * as);
          }
          continue;
        }
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const int col = half * (kBN / 2) + c * 32;
          ptx::tmem_ld_32x32b_x32(tile_taddr + col, ra);
          ptx::tmem_ld_wait();
          if (c == 3) {
            // this warp's last TMEM read of the tile has landed: hand the accumulator back
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(bar_acc_empty + 8 * as);
          }
          if (col < nv) {
            if (p.seed_all) {
              if (query >= 0)
                score_chunk_seed(ra, nrm4 + col / 4, p.alpha, col0 + col, bias,
                                 row_buf + seed_slot * kSeedSlotRows + (ti * kBN + col));
            } else {
              score_chunk_queue(ra, nrm4 + col / 4, p.alpha, col0 + col, tau, bias, static_cast<int>(qslot),
                                hq, p.big_cand, p.big_count, p.big_cap, lane);
            }
          }
        }
        __syncwarp();   // all lanes are done with this tile's ||r^||^2 values
        if (lane == 0) ptx::mbar_arrive(bar_norm_empty + 8 * as);
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
