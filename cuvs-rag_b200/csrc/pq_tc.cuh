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
// the list.  The smem look-ups per row drop from (queries probing the list) x pq_dim to pq_dim.
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
constexpr int kPqTcSmemBytes =
    kPqTcStages * kPqTcStageBytes + 2 * kNormBytes + 256 + kPqTcMaxCbBytes + 1024;
static_assert(kBK == 64, "the decoder writes 128-byte swizzled rows");

struct PqTcParams {
  BfTcParams tc;            // work table, thresholds, append buffers (see bf_tc.cuh, work mode)
  const uint4* codes4;      // PQ codes, 32-row groups interleaved by 16-byte chunks
  const uint32_t* cb16;     // bf16 codebooks [pq_dim][256][DSUB] viewed as 32-bit words
  const float* row_bias;    // [query rows] ||rq||^2 (L2) or -q.c_l (IP) of each gathered row
  int n_code_chunks;        // 16-byte code chunks per row (pq_dim / 16)
  uint32_t n_groups;        // 32-row groups in `codes4`
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
    // A stage's list k-block = 256 rows x 64 dims = 64 / DSUB codes per row.  Work unit of a lane:
    // one 16-byte code chunk of one row (coalesced across the warp by the interleaved layout)
    // -> 16 codebook look-ups -> 16 * DSUB bf16 values = (DSUB * 2) 16-byte stores into the
    // row's 128-byte swizzled line: chunk c of row r lives at (r/8)*1024 + (r%8)*128 + ((c^(r%8))*16).
    const int dw = warp - 8;
    // code chunks per k-block: 2 (DSUB 2) | 1 (DSUB 4) | half a chunk (DSUB 8: k-block kb uses
    // bytes [8*(kb&1), +8) of chunk kb/2)
    constexpr int kChunksPerKb = (DSUB == 2) ? 2 : 1;
    constexpr int kPairs = 8 * kChunksPerKb;                                     // (group, chunk) pairs per stage
    uint32_t stage = 0, phase = 0;
    for (int item = unit; item < n_items; item += n_units) {
      const int4 w = __ldg(p.work + item);
      const int t1 = (w.z - w.y + kBN - 1) / kBN;
      for (int ti = 0; ti < t1; ++ti) {
        const uint32_t g_tile0 = (static_cast<uint32_t>(w.y) >> 5) + static_cast<uint32_t>(ti) * 8u;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          // issue this warp's code loads before waiting for the smem slot
          uint4 cv[kPairs / 4];
#pragma unroll
          for (int u = 0; u < kPairs / 4; ++u) {
            const int pi = dw * (kPairs / 4) + u;
            const uint32_t g = g_tile0 + static_cast<uint32_t>(pi / kChunksPerKb);
            const int ch = (DSUB == 8) ? (kb >> 1) : kb * kChunksPerKb + (pi % kChunksPerKb);
            cv[u] = make_uint4(0, 0, 0, 0);
            if (g < pp.n_groups)
              cv[u] = __ldg(pp.codes4 + (static_cast<size_t>(g) * pp.n_code_chunks + ch) * 32 + lane);
          }
          ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
          const uint32_t b_base = smem_base + stage * kStageBytes + kABytes;
#pragma unroll
          for (int u = 0; u < kPairs / 4; ++u) {
            const int pi = dw * (kPairs / 4) + u;
            const int gl = pi / kChunksPerKb;
            const int chl = pi % kChunksPerKb;                  // chunk within the k-block
            const int ch = (DSUB == 8) ? (kb >> 1) : kb * kChunksPerKb + chl;
            const uint32_t r = static_cast<uint32_t>(gl) * 32u + lane;
            const uint32_t line = b_base + (r >> 3) * 1024u + (r & 7u) * 128u;
            const uint32_t wv[4] = {cv[u].x, cv[u].y, cv[u].z, cv[u].w};
            if (DSUB == 2) {
              // 16 codes -> 32 dims -> 4 output chunks; output chunk index = chl*4 + oc
#pragma unroll
              for (int oc = 0; oc < 4; ++oc) {
                uint32_t o[4];
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                  const int m = ch * 16 + oc * 4 + b;
                  o[b] = cb_w[m * 256 + ((wv[oc] >> (8 * b)) & 0xFFu)];
                }
                const uint32_t c = static_cast<uint32_t>(chl * 4 + oc);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(line + ((c ^ (r & 7u)) << 4)),
                             "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
              }
            } else if (DSUB == 8) {
              // this k-block's 8 codes (half of the chunk) -> 64 dims -> the row's 8 output chunks
              const uint4* cb4 = reinterpret_cast<const uint4*>(cb_w);
              const int h = kb & 1;
              uint4 e[8];
#pragma unroll
              for (int oc = 0; oc < 8; ++oc) {
                const int m = ch * 16 + h * 8 + oc;
                e[oc] = cb4[m * 256 + ((wv[2 * h + (oc >> 2)] >> (8 * (oc & 3))) & 0xFFu)];
              }
#pragma unroll
              for (int oc = 0; oc < 8; ++oc) {
                const uint32_t c = static_cast<uint32_t>(oc);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(line + ((c ^ (r & 7u)) << 4)),
                             "r"(e[oc].x), "r"(e[oc].y), "r"(e[oc].z), "r"(e[oc].w) : "memory");
              }
            } else {
              // 16 codes -> 64 dims -> the row's 8 output chunks
              const uint2* cb2 = reinterpret_cast<const uint2*>(cb_w);
#pragma unroll
              for (int oc = 0; oc < 8; ++oc) {
                const int i0 = oc * 2;
                const int m0 = ch * 16 + i0;
                const uint2 a = cb2[m0 * 256 + ((wv[i0 >> 2] >> (8 * (i0 & 3))) & 0xFFu)];
                const uint2 b = cb2[(m0 + 1) * 256 + ((wv[(i0 + 1) >> 2] >> (8 * ((i0 + 1) & 3))) & 0xFFu)];
                const uint32_t c = static_cast<uint32_t>(oc);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(line + ((c ^ (r & 7u)) << 4)),
                             "r"(a.x), "r"(a.y), "r"(b.x), "r"(b.y) : "memory");
              }
            }
          }
          // generic-proxy writes -> visible to the tensor core's async-proxy reads
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(bar_full + 8 * stage);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
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
      int* const row_cnt = p.big_count + qslot;
      int cnt = 0;
      u64 best = kKeyInf;
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
          if (col < nv)
            score_chunk<kModeAppend>(ra, nrm4 + col / 4, p.alpha, col0 + col, tau, cnt, best, row_buf,
                                     row_cnt, p.big_cap, bias);
        }
        __syncwarp();   // all lanes are done with this tile's ||r^||^2 values
        if (lane == 0) ptx::mbar_arrive(bar_norm_empty + 8 * as);
      }
    }
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
