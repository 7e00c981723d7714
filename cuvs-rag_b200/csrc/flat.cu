// Host side of the exact-search engine (K0): operand preparation, TMA descriptors, work
// decomposition and launch of bf_tc_kernel + the split merge.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "bf_tc.cuh"
#include "common.h"

namespace b2vs {

// ------------------------------------------------------------------------------------------
// Row preparation: one warp per row.  Produces the 16-bit K-major operand (optionally) and the
// per-row squared norm (optionally).
//   mode 0: convert/pad to the 16-bit operand format            out row = [x]            (kdim = dp)
//   mode 1: fp32 -> bf16 hi/lo split, database side             out row = [hi | hi | lo] (kdim = 3dp)
//   mode 2: fp32 -> bf16 hi/lo split, query side                out row = [hi | lo | hi] (kdim = 3dp)
// The norm is taken over the values the contraction will see: the rounded 16-bit values in
// mode 0, the original fp32 values in modes 1/2 (the split reproduces them to ~2^-16).
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) {
  return __bfloat162float(v);
}
__device__ __forceinline__ uint16_t f32_to_op(float v, int fmt, float* back) {
  if (fmt == 0) {
    __half h = __float2half_rn(v);
    *back = __half2float(h);
    return __half_as_ushort(h);
  }
  __nv_bfloat16 b = __float2bfloat16_rn(v);
  *back = __bfloat162float(b);
  return __bfloat16_as_ushort(b);
}

template <typename T>
__global__ void prep_rows_kernel(const T* __restrict__ src, int64_t n, int dim, int dp, int mode,
                                 int fmt, uint16_t* __restrict__ out, float* __restrict__ norm_out,
                                 int want_norm, int64_t n_pad, float pad_value) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int kdim = (mode == 0) ? dp : 3 * dp;
  for (int64_t r = warp0; r < n_pad; r += nwarps) {
    if (r >= n) {
      if (norm_out && lane == 0) norm_out[r] = pad_value;
      continue;
    }
    const T* row = src + r * dim;
    uint16_t* orow = out ? out + r * kdim : nullptr;
    float acc = 0.f;
    for (int j = lane; j < dp; j += 32) {
      const float v = (j < dim) ? to_f32<T>(row[j]) : 0.f;
      if (mode == 0) {
        float back;
        const uint16_t o = f32_to_op(v, fmt, &back);
        if (orow) orow[j] = o;
        acc = fmaf(back, back, acc);
      } else {
        float hi_f, lo_f;
        const uint16_t hi = f32_to_op(v, 1, &hi_f);
        const uint16_t lo = f32_to_op(v - hi_f, 1, &lo_f);
        if (orow) {
          orow[j] = hi;
          orow[dp + j] = (mode == 1) ? hi : lo;
          orow[2 * dp + j] = (mode == 1) ? lo : hi;
        }
        acc = fmaf(v, v, acc);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (norm_out && lane == 0) norm_out[r] = want_norm ? acc : 0.f;
  }
}

static int launch_prep(const void* src, int dtype, int64_t n, int dim, int dp, int mode, int fmt,
                       uint16_t* out, float* norm_out, int want_norm, int64_t n_pad,
                       cudaStream_t st) {
  if (n_pad <= 0) return B2VS_OK;
  const int threads = 256;
  const int64_t want_blocks = ceil_div(n_pad, threads / 32);
  const int blocks = static_cast<int>(std::min<int64_t>(want_blocks, 148 * 16));
  const float inf = INFINITY;
  switch (dtype) {
    case B2VS_F32:
      prep_rows_kernel<float><<<blocks, threads, 0, st>>>(static_cast<const float*>(src), n, dim,
                                                          dp, mode, fmt, out, norm_out, want_norm,
                                                          n_pad, inf);
      break;
    case B2VS_F16:
      prep_rows_kernel<__half><<<blocks, threads, 0, st>>>(static_cast<const __half*>(src), n, dim,
                                                           dp, mode, fmt, out, norm_out, want_norm,
                                                           n_pad, inf);
      break;
    case B2VS_BF16:
      prep_rows_kernel<__nv_bfloat16><<<blocks, threads, 0, st>>>(
          static_cast<const __nv_bfloat16*>(src), n, dim, dp, mode, fmt, out, norm_out, want_norm,
          n_pad, inf);
      break;
    default:
      set_error("unknown dtype %d", dtype);
      return B2VS_EINVAL;
  }
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}

// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 2-D map over a row-major [rows, cols] 16-bit matrix; box = [box_rows, 64], 128-byte swizzle,
// out-of-bounds elements read as zero.
int encode_tmap_2d(CUtensorMap* tm, const void* base, int ab_format, int64_t rows, int64_t cols,
                   int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  B2VS_CHECK(fn != nullptr, B2VS_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  B2VS_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, B2VS_EINVAL,
             "matrix base pointer must be 16-byte aligned");
  B2VS_CHECK((cols * 2) % 16 == 0, B2VS_EINVAL, "row pitch must be a multiple of 16 bytes");
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(cols) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kBK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, ab_format == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                  2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  kBK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B2VS_CHECK(r == CUDA_SUCCESS, B2VS_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d",
             static_cast<int>(r));
  return B2VS_OK;
}

int sm_count(int dev) {
  static int cache[64];
  if (dev < 0 || dev >= 64) return 148;
  if (cache[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)
      v = 148;
    cache[dev] = v;
  }
  return cache[dev];
}

// ------------------------------------------------------------------------------------------
int FlatEngine::init(int dev_, int metric_, int dtype, int dim_, const void* db, int64_t n_,
                     cudaStream_t st, int force_fmt) {
  dev = dev_;
  metric = metric_;
  src_dtype = dtype;
  dim = dim_;
  n = n_;
  const int dp = static_cast<int>(round_up(dim, 8));
  split3 = (dtype == B2VS_F32) && force_fmt < 0;
  ab_format = (dtype == B2VS_F16) ? 0 : 1;
  if (dtype == B2VS_F32 && force_fmt >= 0) ab_format = force_fmt;
  kdim = split3 ? 3 * dp : dp;
  const int64_t tiles = std::max<int64_t>(1, ceil_div(n, kBN));
  B2VS_TRY(beta.reserve(static_cast<size_t>(tiles) * kBN * sizeof(float)));
  // only a 16-bit source already in the operand format can be used in place
  const bool borrow =
      dtype != B2VS_F32 && dp == dim && (reinterpret_cast<uintptr_t>(db) & 15) == 0;
  uint16_t* out = nullptr;
  if (!borrow) {
    B2VS_TRY(owned.reserve(static_cast<size_t>(std::max<int64_t>(n, 1)) * kdim * 2));
    out = owned.as<uint16_t>();
    mat = owned.ptr;
  } else {
    mat = db;
  }
  B2VS_TRY(launch_prep(db, dtype, n, dim, dp, split3 ? 1 : 0, ab_format, out, beta.as<float>(),
                       metric == B2VS_METRIC_L2 ? 1 : 0, tiles * kBN, st));
  if (n > 0) {
    B2VS_TRY(encode_tmap_2d(&tm_x, mat, ab_format, n, kdim, kBN));           // G = 1: whole tile
    B2VS_TRY(encode_tmap_2d(&tm_x_half, mat, ab_format, n, kdim, kBN / 2));  // G = 2: half per CTA
  }
  return B2VS_OK;
}

void FlatEngine::resolve_timing() {
  if (!timing_pending || !ev1) return;
  float ms = 0.f;
  if (cudaEventSynchronize(ev1) == cudaSuccess && cudaEventElapsedTime(&ms, ev0, ev1) == cudaSuccess)
    stats.kernel_ms = ms;
  else
    cudaGetLastError();
  timing_pending = false;
}

void FlatEngine::destroy() {
  if (ev0) cudaEventDestroy(ev0);
  if (ev1) cudaEventDestroy(ev1);
  ev0 = ev1 = nullptr;
  owned.release();
  beta.release();
  ws_cand.release();
  ws_keys.release();
  ws_q.release();
  ws_qnorm.release();
  ws_tau.release();
  ws_big.release();
  ws_bigcnt.release();
  ws_chunk.release();
  ws_rawcnt.release();
  ws_samp.release();
  ws_work.release();
}

__global__ void fill_missing_kernel(float* out_d, int64_t* out_i, int32_t* out_label, int64_t total,
                                    float dval) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  if (out_d) out_d[i] = dval;
  if (out_i) out_i[i] = -1;
  if (out_label) out_label[i] = -1;
}

constexpr int kDefaultTcGroup = 1;
constexpr size_t kMaxRawBytes = 3ull << 30;   // raw emission: upper bound of the per-item candidate buffers
constexpr int kPairMaxK = 32;   // largest k served by the CTA-pair kernel by default

// B2VS_TC_GROUP=1|2 forces the single-CTA / CTA-pair kernel (bring-up and A/B measurements).
static int tc_group_override() { return env().tc_group; }

// Tile strides of the passes of one search (see FlatEngine::search).  B2VS_PASSES="16,1" etc.
// overrides the heuristic for A/B measurements.
static int pass_strides(int64_t tiles, int k, int* strides, int mult = 1) {
  const int env_n = env().pass_n;
  const int* env_s = env().pass_s;
  if (k == 1) { strides[0] = 1; return 1; }        // arg-min keeps its state in registers
  if (env_n > 0) {
    int m = 0;
    for (int i = 0; i < env_n; ++i)
      if (env_s[i] == 1 || tiles / env_s[i] >= 8) strides[m++] = env_s[i];
    return m;
  }
  // mult > 1 (sharded search, union exchange): the sampled passes of `mult` shards pool their
  // samples, so each shard samples `mult` times more sparsely; a pass whose sample would shrink
  // below 8 tiles is dropped
  int m = 0;
  if (tiles >= 8192 && tiles / (256 * mult) >= 8) strides[m++] = 256 * mult;
  if (tiles >= 256 && tiles / (16 * mult) >= 8) strides[m++] = 16 * mult;
  strides[m++] = 1;
  return m;
}

// Number of db splits: minimise waves * (tiles per split + fixed per-item cost in tile units).
static int choose_splits(int n_qblocks, int64_t tiles, int sms, int k, bool seeded = false) {
  const int64_t overhead = (k == 1) ? 1 : (seeded ? 16 : 40);
  int64_t best_cost = INT64_MAX;
  int best = 1;
  const int64_t smax = std::min<int64_t>(tiles, 512);
  for (int64_t s = 1; s <= smax; ++s) {
    const int64_t tps = ceil_div(tiles, s);
    const int64_t s_eff = ceil_div(tiles, tps);
    const int64_t waves = ceil_div(static_cast<int64_t>(n_qblocks) * s_eff, sms);
    const int64_t cost = waves * (tps + overhead);
    if (cost < best_cost) { best_cost = cost; best = static_cast<int>(s_eff); }
  }
  return best;
}

// One launch of the fused distance + top-k kernel over this engine's matrix: the single-CTA
// variant, or CTA pairs (cluster of 2, each CTA staging half of every db tile).
int FlatEngine::launch_fused(int group, int grid, const CUtensorMap& tm_q, const BfTcParams& p,
                             cudaStream_t st, int epi_groups) const {
  if (group == 1 && epi_groups == 2) {
    B2VS_CUDA(cudaFuncSetAttribute((bf_tc_kernel<1, false, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   TcCfg<1>::kSmemBytes));
    bf_tc_kernel<1, false, 2><<<grid, tc_threads(2), TcCfg<1>::kSmemBytes, st>>>(tm_q, tm_x, p, TailMaps{});
  } else if (group == 1) {
    B2VS_CUDA(cudaFuncSetAttribute(bf_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   TcCfg<1>::kSmemBytes));
    bf_tc_kernel<1><<<grid, tc_threads(1), TcCfg<1>::kSmemBytes, st>>>(tm_q, tm_x, p, TailMaps{});
  } else {
    B2VS_CUDA(cudaFuncSetAttribute(bf_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   TcCfg<2>::kSmemBytes));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(tc_threads(1));
    cfg.dynamicSmemBytes = TcCfg<2>::kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    B2VS_CUDA(cudaLaunchKernelEx(&cfg, bf_tc_kernel<2>, tm_q, tm_x_half, p, TailMaps{}));
  }
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}

bool flat_exchanges_tau(int64_t min_rows, int k, int stride_mult) {
  if (k <= 1 || k > kMaxFusedK) return false;
  int strides[3];
  return pass_strides(ceil_div(std::max<int64_t>(min_rows, 1), kBN), k, strides, std::max(1, stride_mult)) >= 2;
}

int FlatEngine::search(const void* q, int q_dtype, int nq, int k, int force_splits,
                       int64_t id_offset, float* out_d, int64_t* out_i, int32_t* out_label,
                       cudaStream_t st, int flags, const TauExchange* tau_exchange) {
  B2VS_CHECK(nq > 0, B2VS_EINVAL, "nq must be positive (got %d)", nq);
  B2VS_CHECK(k >= 1 && k <= kMaxBigK, B2VS_EUNSUP, "k=%d outside the supported range [1, %d]", k,
             kMaxBigK);
  B2VS_CHECK(k <= kMaxFusedK || out_label == nullptr, B2VS_EUNSUP, "labels need k <= %d",
             kMaxFusedK);
  stats = b2vs_search_stats{};
  const float missing = (metric == B2VS_METRIC_IP) ? -INFINITY : INFINITY;
  if (n == 0) {
    const int64_t total = static_cast<int64_t>(nq) * k;
    fill_missing_kernel<<<static_cast<unsigned>(ceil_div(total, 256)), 256, 0, st>>>(
        out_d, out_i, out_label, total, missing);
    B2VS_CUDA(cudaGetLastError());
    stats.launches = 1;
    return B2VS_OK;
  }
  const int dp = static_cast<int>(round_up(dim, 8));
  const int want_norm = (metric == B2VS_METRIC_L2) ? 1 : 0;
  // kernel variant: CTA pair (cta_group::2, 256-row query blocks) or single CTA (128-row blocks)
  int group = tc_group_override();
  // Default: the CTA-pair kernel for small k (sustained C2-size runs: k=1 +5.4 %, k=10 +3.2 % over
  // the single-CTA kernel), the single-CTA kernel otherwise (k=100: pair -3.5 %: both sit at the
  // power cap and the pair's 98 %-busy tensor pipe loses more clock than it gains).
  // (C5, 50M x 1024, k=10: the pair wins at Q = 256 (one 256-row block) and from Q = 4096 up,
  // the single-CTA kernel at Q = 1024 by 11 %.)
  if (group == 0)
    group = (k <= kPairMaxK && (nq >= 4096 || (nq > kBM && nq <= 2 * kBM))) ? 2 : kDefaultTcGroup;
  if (flags & B2VS_FLAG_TC_SINGLE) group = 1;
  if (flags & B2VS_FLAG_TC_PAIR) group = 2;
  // Epilogue warp groups: ONE by default everywhere.  Two groups (alternate tiles, private
  // thresholds) were tried for the coarse probes of the IVF indexes, whose selection work is the
  // whole kernel (top-64 of 16 384 centroids: tensor pipe 5 % busy): slower, 732 vs 672 us at C4,
  // 327 vs 289 us at C3 - each group sees half the columns with its own threshold, so together they
  // insert ~2 k ln(N / 2k) candidates instead of k ln(N / k).  B2VS_EPI_GROUPS=2 / B2VS_FLAG_EPI2 keep
  // the variant reachable for measurements.
  const int64_t tiles_all = ceil_div(n, kBN);
  (void)tiles_all;
  int epi_groups = 1;
  if (env().epi_groups > 0) epi_groups = env().epi_groups;
  if (flags & B2VS_FLAG_EPI2) epi_groups = 2;
  if (k == 1 || k > kMaxFusedK || (flags & B2VS_FLAG_TC_PAIR) || tc_group_override() == 2) epi_groups = 1;
  if (epi_groups == 2) group = 1;
  const int qrows = kBM * group;
  const int n_qblocks = static_cast<int>(ceil_div(nq, qrows));
  const int q_pad = n_qblocks * qrows;
  int launches = 0;

  // ---- query operand + norms
  const int op_dtype = ab_format == 0 ? B2VS_F16 : B2VS_BF16;
  // A query block whose rows run past the end of the query matrix is staged through TMA's
  // out-of-bounds fill, which measurably slows the load of that block (the HBM-bound small-batch
  // case loses ~30 %): unless the batch is an exact multiple of the block, queries are copied into
  // a zero-padded operand so every block is fully in bounds.
  const bool no_qpad = env().no_qpad;  // A/B switch
  const bool borrow_q = !split3 && q_dtype == op_dtype && dp == dim && (nq == q_pad || no_qpad) &&
                        (reinterpret_cast<uintptr_t>(q) & 15) == 0;
  B2VS_TRY(ws_qnorm.reserve(static_cast<size_t>(q_pad) * sizeof(float)));
  const void* q_mat = q;
  if (!borrow_q) {
    B2VS_TRY(ws_q.reserve(static_cast<size_t>(q_pad) * kdim * 2));
    q_mat = ws_q.ptr;
    if (q_pad > nq)
      B2VS_CUDA(cudaMemsetAsync(ws_q.as<uint16_t>() + static_cast<size_t>(nq) * kdim, 0,
                                static_cast<size_t>(q_pad - nq) * kdim * 2, st));
    B2VS_TRY(launch_prep(q, q_dtype, nq, dim, dp, split3 ? 2 : 0, ab_format, ws_q.as<uint16_t>(),
                         ws_qnorm.as<float>(), want_norm, nq, st));
    ++launches;
  } else if (want_norm) {
    B2VS_TRY(launch_prep(q, q_dtype, nq, dim, dp, 0, ab_format, nullptr, ws_qnorm.as<float>(), 1,
                         nq, st));
    ++launches;
  }
  CUtensorMap tm_q;
  B2VS_TRY(encode_tmap_2d(&tm_q, q_mat, ab_format, borrow_q ? nq : q_pad, kdim, kBM));

  if (!tau_exchange && force_splits <= 0 && two_pass_applies(nq, k, tc_group_override(), flags)) {
    B2VS_TRY(search_two_pass(q_mat, borrow_q ? nq : q_pad, nq, k, id_offset, out_d, out_i, out_label, st,
                          (flags & B2VS_FLAG_TIME_KERNEL) != 0, &launches));
    stats.launches = launches;
    stats.algo_flops = 2.0 * nq * static_cast<double>(n) * dim;
    return B2VS_OK;
  }
  if (k > kMaxFusedK) {
    B2VS_TRY(search_bigk(q_mat, nq, q_pad, group, k, id_offset, out_d, out_i, st, &launches));
    stats.launches = launches;
    stats.algo_flops = 2.0 * nq * static_cast<double>(n) * dim;
    return B2VS_OK;
  }

  // ---- passes.  A pass visits every `stride`-th db tile.  The last pass (stride 1) produces the
  // answer; earlier, sparser passes only seed each query's threshold with the k-th best score of
  // a 1/stride sample (a valid upper bound of the final k-th score), so the full pass starts with
  // a tight threshold: ~k*stride candidates per query in total instead of k*ln(N/k) per split,
  // and practically no buffer compactions.
  const int sms = sm_count(dev);
  const int units = std::max(1, sms / group);  // CTAs (G=1) or CTA pairs (G=2) that run at once
  const int64_t tiles = ceil_div(n, kBN);
  int strides[3];
  // sharded search with threshold exchange: the schedule follows the smallest shard of the job,
  // so that every rank runs the same number of passes (= collective calls)
  const int64_t sched_tiles = tau_exchange ? ceil_div(std::max<int64_t>(tau_exchange->schedule_rows, 1), kBN) : tiles;
  const bool union_mode = tau_exchange && tau_exchange->union_fn && tau_exchange->world > 1;
  int n_pass = pass_strides(sched_tiles, k, strides, union_mode ? tau_exchange->world : 1);
  if (union_mode) B2VS_TRY(ws_samp.reserve(static_cast<size_t>(q_pad) * k * sizeof(float)));
  B2VS_TRY(ws_tau.reserve(static_cast<size_t>(q_pad) * sizeof(float)));

  BfTcParams p{};
  p.beta = beta.as<float>();
  p.n_qblocks = n_qblocks;
  p.q_pad = q_pad;
  p.nq = nq;
  p.k_blocks = static_cast<int>(ceil_div(kdim, kBK));
  p.k = k;
  p.alpha = (metric == B2VS_METRIC_L2) ? -2.f : -1.f;
  p.idesc = ptx::make_idesc_f16(static_cast<uint32_t>(ab_format), qrows, kBN);
  p.debug_skip_emit = env().k0_debug & 1;

  const bool timed = (flags & B2VS_FLAG_TIME_KERNEL) != 0;
  if (timed) {
    if (!ev0) {
      B2VS_CUDA(cudaEventCreate(&ev0));
      B2VS_CUDA(cudaEventCreate(&ev1));
    }
    B2VS_CUDA(cudaEventRecord(ev0, st));
  }
  int n_splits = 1, grid = 1;
  for (int pass = 0; pass < n_pass; ++pass) {
    const int stride = strides[pass];
    const bool last = (pass == n_pass - 1);
    const int64_t ptiles = ceil_div(tiles, stride);
    n_splits = (force_splits > 0 && last)
                   ? static_cast<int>(std::min<int64_t>(force_splits, ptiles))
                   : choose_splits(n_qblocks, ptiles, units, k, /*seeded=*/pass > 0);
    const int tps = static_cast<int>(ceil_div(ptiles, n_splits));
    n_splits = static_cast<int>(ceil_div(ptiles, tps));
    const int n_items = n_qblocks * n_splits;
    grid = std::min(n_items, units) * group;
    // Raw emission (k > 1): per-ITEM candidate buffers that double as the pass's output - no final
    // sort inside the kernel; merge_raw_kernel selects from them.  Falls back to the sorted per-item
    // lists when the buffers would be out of proportion (2 KB per item row).
    const size_t raw_rows = static_cast<size_t>(n_items) * group * epi_groups * kBM;
    const bool raw = k > 1 && env().raw_emit != 0 && raw_rows * kCap * sizeof(u64) <= kMaxRawBytes;
    if (raw) {
      B2VS_TRY(ws_cand.reserve(raw_rows * kCap * sizeof(u64)));
      B2VS_TRY(ws_rawcnt.reserve(raw_rows * sizeof(int)));
    } else {
      B2VS_TRY(ws_cand.reserve(static_cast<size_t>(grid) * epi_groups * kBM * kCap * sizeof(u64)));
      B2VS_TRY(ws_keys.reserve(static_cast<size_t>(n_splits) * epi_groups * q_pad * k * sizeof(u64)));
    }
    p.cand = ws_cand.as<u64>();
    p.out_keys = ws_keys.as<u64>();
    p.raw_count = raw ? ws_rawcnt.as<int>() : nullptr;
    p.n_items = n_items;
    p.tiles_total = static_cast<int>(ptiles);
    p.tiles_per_split = tps;
    p.tile_stride = stride;
    p.tau_init = (pass > 0) ? ws_tau.as<float>() : nullptr;
    B2VS_TRY(launch_fused(group, grid, tm_q, p, st, epi_groups));
    ++launches;
    if (last && timed) B2VS_CUDA(cudaEventRecord(ev1, st));
    if (last) {
      if (raw)
        B2VS_TRY(launch_merge_raw(ws_cand.as<u64>(), ws_rawcnt.as<int>(), n_splits, n_qblocks, group, epi_groups,
                                  nq, k, metric, ws_qnorm.as<float>(), id_offset, out_d, out_i, out_label,
                                  nullptr, st));
      else
        B2VS_TRY(launch_merge_splits(ws_keys.as<u64>(), n_splits * epi_groups, q_pad, nq, k, metric,
                                     ws_qnorm.as<float>(), id_offset, out_d, out_i, out_label, st));
    } else {
      // sampled pass: only the k-th best raw score per query is kept, as the next pass's threshold
      // (union mode: the k best raw scores, for the exchange below)
      float* const samp = union_mode ? ws_samp.as<float>() : nullptr;
      if (raw)
        B2VS_TRY(launch_merge_raw(ws_cand.as<u64>(), ws_rawcnt.as<int>(), n_splits, n_qblocks, group, epi_groups,
                                  q_pad, k, metric, nullptr, 0, nullptr, nullptr, nullptr, ws_tau.as<float>(), st,
                                  samp));
      else
        B2VS_TRY(launch_merge_splits(ws_keys.as<u64>(), n_splits * epi_groups, q_pad, q_pad, k, metric,
                                     nullptr, 0, nullptr, nullptr, nullptr, st, nullptr, ws_tau.as<float>(), samp));
      if (union_mode) {
        // every shard continues with the k-th best score of the POOLED samples of all shards
        B2VS_TRY(tau_exchange->union_fn(tau_exchange->ctx, samp, q_pad, k, ws_tau.as<float>(), st));
        ++launches;
        continue;
      }
      // sharded search: every shard continues with the tightest bound any shard found (a shard's
      // k-th best sampled score bounds the GLOBAL k-th score from above, so the minimum does too)
      if (tau_exchange) B2VS_TRY(tau_exchange->fn(tau_exchange->ctx, ws_tau.as<float>(), q_pad, st));
    }
    ++launches;
  }
  timing_pending = timed;

  stats.launches = launches;
  stats.n_splits = n_splits;
  stats.grid = grid;
  stats.algo_flops = 2.0 * nq * static_cast<double>(n) * dim;
  stats.algo_bytes = 0;
  return B2VS_OK;
}

// ------------------------------------------------------------------------------------------
// Two-pass selection for a SMALL database searched with a LARGE k - the coarse probes of the IVF
// indexes (top-64 of 16 384 centroids at C4).  The fused kernel's selection costs ~k ln(N/k)
// insertions per query, each a divergent slow path of the lane that owns the row; there it is the
// whole kernel (0.67 ms at C4 with the tensor pipe 5 % busy, and only ceil(Q/128) = 79 of 148 SMs
// at work, since splitting the rows multiplies the insertions).  The GEMM itself is cheap here, so
// it runs twice, both times as the work-table kernel over (query block, tile range) items that fill
// every SM:
//   pass 1  stores only the minimum of every 32-row chunk (n/32 floats per query);
//   chunk_tau_kernel: k-th smallest (minimum, chunk) pair = an exact threshold that at most 32 k
//           elements pass, ties included (merge.cu);
//   pass 2  appends the elements that pass to the query's buffer (capacity 32 k: cannot overflow);
//   cand_select_kernel: top-k of those (~k + a few dozen in practice) -> answer rows.
// (A single pass that stored the whole score matrix - 663 MB at C4 - and selected from it was
// tried first: 0.31 ms of half-sector stores + 0.28 ms of selection, 0.60 ms in all.)
constexpr int64_t kTwoPassMaxRows = 65536;
constexpr int kTwoPassMinK = 16;
constexpr int kTwoPassMinQueries = 512;
constexpr size_t kTwoPassCandBytes = 1ull << 30;

bool FlatEngine::two_pass_applies(int nq, int k, int group_forced, int flags) const {
  if (k < 2 || k > kMaxFusedK || n < k) return false;
  if (group_forced == 2 || (flags & (B2VS_FLAG_TC_PAIR | B2VS_FLAG_EPI2))) return false;
  const int ov = env().two_pass;
  if (ov == 0) return false;
  if (n > 16 * kTwoPassMaxRows) return false;
  if (ov == 1) return true;
  return n <= kTwoPassMaxRows && k >= kTwoPassMinK && nq >= kTwoPassMinQueries;
}

// Work table of both passes: item = (query block, tile range), range-major so that the CTAs running
// at one time share database tiles; row_query = identity (-1 on the padding rows of the last block).
__global__ void two_pass_plan_kernel(int n_qblocks, int n_ranges, int tiles, int tiles_per_range,
                                     int nq, int4* __restrict__ work, int* __restrict__ n_work,
                                     int* __restrict__ row_query) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_qblocks * n_ranges) {
    const int r = i / n_qblocks, qb = i % n_qblocks;
    const int t0 = r * tiles_per_range, t1 = min(t0 + tiles_per_range, tiles);
    work[i] = make_int4(qb, t0 * kBN, t1 * kBN, 0);
  }
  if (i < n_qblocks * kBM) row_query[i] = i < nq ? i : -1;
  if (i == 0) *n_work = n_qblocks * n_ranges;
}

int FlatEngine::search_two_pass(const void* q_mat, int64_t q_rows, int nq, int k, int64_t id_offset,
                                float* out_d, int64_t* out_i, int32_t* out_label, cudaStream_t st,
                                bool timed, int* launches) {
  const int sms = sm_count(dev);
  const int tiles = static_cast<int>(ceil_div(n, kBN));
  const int n_chunks = tiles * (kBN / 32);
  const int cap = 32 * k;
  const size_t budget = env().two_pass_chunk_mb > 0 ? static_cast<size_t>(env().two_pass_chunk_mb) << 20
                                                    : kTwoPassCandBytes;
  const int64_t n_qblocks_all = ceil_div(nq, kBM);
  const int64_t blocks_per_batch = std::max<int64_t>(
      1, std::min<int64_t>(n_qblocks_all, static_cast<int64_t>(budget / (static_cast<size_t>(cap) * 8 * kBM))));
  const int64_t rows_b = blocks_per_batch * kBM;
  B2VS_TRY(ws_big.reserve(static_cast<size_t>(rows_b) * cap * sizeof(u64)));
  B2VS_TRY(ws_bigcnt.reserve(static_cast<size_t>(rows_b) * sizeof(int)));
  B2VS_TRY(ws_chunk.reserve(static_cast<size_t>(rows_b) * n_chunks * sizeof(float)));
  B2VS_TRY(ws_tau.reserve(static_cast<size_t>(rows_b) * 2 * sizeof(float)));
  const int max_work = static_cast<int>(blocks_per_batch) * tiles;
  B2VS_TRY(ws_work.reserve(static_cast<size_t>(max_work) * sizeof(int4) + 16 + static_cast<size_t>(rows_b) * sizeof(int)));
  int4* work = ws_work.as<int4>();
  int* n_work = reinterpret_cast<int*>(ws_work.as<char>() + static_cast<size_t>(max_work) * sizeof(int4));
  int* row_query = n_work + 4;
  float* tau = ws_tau.as<float>();
  int* tau_chunk = reinterpret_cast<int*>(tau + rows_b);
  if (timed) {
    if (!ev0) {
      B2VS_CUDA(cudaEventCreate(&ev0));
      B2VS_CUDA(cudaEventCreate(&ev1));
    }
    B2VS_CUDA(cudaEventRecord(ev0, st));
  }
  int n_ranges = 1;
  for (int64_t b0 = 0; b0 < n_qblocks_all; b0 += blocks_per_batch) {
    const int n_qblocks = static_cast<int>(std::min(blocks_per_batch, n_qblocks_all - b0));
    const int64_t q0 = b0 * kBM;
    const int nq_b = static_cast<int>(std::min<int64_t>(nq - q0, static_cast<int64_t>(n_qblocks) * kBM));
    // tile ranges cost nothing here (no selection state per item): fill the SMs
    n_ranges = choose_splits(n_qblocks, tiles, sms, /*k=*/1);
    const int tpr = static_cast<int>(ceil_div(tiles, n_ranges));
    n_ranges = static_cast<int>(ceil_div(tiles, tpr));
    const int n_items = n_qblocks * n_ranges;
    const int plan_threads = std::max(n_items, n_qblocks * kBM);
    two_pass_plan_kernel<<<static_cast<unsigned>(ceil_div(plan_threads, 256)), 256, 0, st>>>(
        n_qblocks, n_ranges, tiles, tpr, nq_b, work, n_work, row_query);
    B2VS_CUDA(cudaGetLastError());
    GroupedScanArgs ga{};
    ga.q_mat = static_cast<const uint16_t*>(q_mat) + static_cast<size_t>(q0) * kdim;
    ga.q_rows = q_rows - q0;
    ga.x_mat = mat; ga.x_rows = n;
    ga.kdim = kdim; ga.ab_format = ab_format; ga.q_split = 0;
    ga.beta = beta.as<float>();
    ga.alpha = (metric == B2VS_METRIC_L2) ? -2.f : -1.f;
    ga.work = work; ga.n_work = n_work; ga.max_work = n_items;
    ga.row_query = row_query;
    ga.tau = tau;
    ga.cand = ws_big.as<u64>(); ga.count = ws_bigcnt.as<int>(); ga.cap = cap;
    ga.seed_all = 2;
    ga.chunk_min = ws_chunk.as<float>(); ga.chunk_ld = n_chunks;
    ga.epi_groups = 2;      // both passes are epilogue-bound (the GEMM is a few percent of them)
    B2VS_TRY(launch_grouped_scan(dev, ga, st));
    B2VS_TRY(launch_chunk_tau(ws_chunk.as<float>(), n_chunks, n_chunks, nq_b, k, tau, tau_chunk,
                              ws_bigcnt.as<int>(), st));
    ga.seed_all = 0;
    ga.tau_chunk = tau_chunk;
    B2VS_TRY(launch_grouped_scan(dev, ga, st));
    const size_t o = static_cast<size_t>(q0) * k;
    B2VS_TRY(launch_cand_select(ws_big.as<u64>(), ws_bigcnt.as<int>(), cap, nq_b, k, metric,
                                ws_qnorm.as<float>() + q0, id_offset, out_d ? out_d + o : nullptr,
                                out_i ? out_i + o : nullptr, out_label ? out_label + o : nullptr, st));
    *launches += 5;
  }
  if (timed) B2VS_CUDA(cudaEventRecord(ev1, st));
  timing_pending = timed;
  stats.n_splits = n_ranges;
  stats.grid = sms;
  return B2VS_OK;
}

// ------------------------------------------------------------------------------------------
int launch_grouped_scan(int dev, const GroupedScanArgs& a, cudaStream_t st) {
  CUtensorMap tm_q, tm_xl;
  const int xkb = static_cast<int>(ceil_div(a.kdim, kBK));
  // query rows come in 32-row quarter boxes when the plan says how many rows a block holds
  const bool quarter_boxes = a.rows_in_work && env().a_quarters != 0;
  B2VS_TRY(encode_tmap_2d(&tm_q, a.q_mat, a.ab_format, a.q_rows, a.q_split ? 2 * xkb * kBK : a.kdim,
                          quarter_boxes ? 32 : kBM));
  const int box = (a.x_box_rows == 64 || a.x_box_rows == 128) ? a.x_box_rows : kBN;
  B2VS_TRY(encode_tmap_2d(&tm_xl, a.x_mat, a.ab_format, a.x_rows, a.kdim, box));
  BfTcParams p{};
  p.beta = a.beta;
  p.k_blocks = a.q_split ? 2 * xkb : xkb;
  p.x_kblocks = a.q_split ? xkb : 0;
  p.k = 1;
  p.tile_stride = 1;
  p.alpha = a.alpha;
  p.idesc = ptx::make_idesc_f16(static_cast<uint32_t>(a.ab_format), kBM, box);   // N = rows of a list tile
  p.stage_tx = box == kBN ? 0u : static_cast<uint32_t>(TcCfg<1>::kABytes + box * kBK * 2);
  // whole-tile launches: short boxes for the last tile of every item (B2VS_TAIL_BOXES=0: A/B switch)
  TailMaps tails{};
  p.tail_boxes = (box == kBN && !a.q_split && env().tail_boxes != 0) ? 1 : 0;
  if (p.tail_boxes) {
    B2VS_TRY(encode_tmap_2d(&tails.m128, a.x_mat, a.ab_format, a.x_rows, a.kdim, 128));
    B2VS_TRY(encode_tmap_2d(&tails.m64, a.x_mat, a.ab_format, a.x_rows, a.kdim, 64));
  }
  p.tau_init = a.tau;
  p.big_cand = a.cand;
  p.big_count = a.count;
  p.big_cap = a.cap;
  p.work = static_cast<const int4*>(a.work);
  p.n_work = a.n_work;
  p.row_query = a.row_query;
  p.row_slot = a.row_slot;
  p.seed_all = a.seed_all;
  p.a_quarter_boxes = quarter_boxes ? 1 : 0;
  p.chunk_min = a.chunk_min;
  p.chunk_ld = a.chunk_ld;
  p.tau_chunk = a.tau_chunk;
  const int grid = std::max(1, std::min(a.max_work, sm_count(dev)));
  const int epi = env().work_epi > 0 ? env().work_epi : (a.epi_groups == 2 ? 2 : 1);
  if (epi == 2) {
    B2VS_CUDA(cudaFuncSetAttribute((bf_tc_kernel<1, true, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   TcCfg<1>::kSmemBytes));
    bf_tc_kernel<1, true, 2><<<grid, tc_threads(2), TcCfg<1>::kSmemBytes, st>>>(tm_q, tm_xl, p, tails);
  } else {
    B2VS_CUDA(cudaFuncSetAttribute((bf_tc_kernel<1, true>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   TcCfg<1>::kSmemBytes));
    bf_tc_kernel<1, true><<<grid, tc_threads(1), TcCfg<1>::kSmemBytes, st>>>(tm_q, tm_xl, p, tails);
  }
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}

// ------------------------------------------------------------------------------------------
// Large k (128 < k <= 2048).  Passes are sparse-to-dense like the fused path, but the kernel only
// APPENDS every score below the query's threshold to a per-query global buffer (no in-kernel
// top-k); bigk_select_kernel then radix-selects the k-th key - the next pass's threshold, or,
// after the full pass, the answer.  Strides are chosen so a pass is expected to leave at most a
// quarter of the buffer per query; a buffer that overflows anyway (adversarial row order) is
// cut to a valid tighter threshold and the pass is repeated.
int FlatEngine::search_bigk(const void* q_mat, int nq, int q_pad, int group, int k,
                            int64_t id_offset, float* out_d, int64_t* out_i, cudaStream_t st,
                            int* launches) {
  constexpr int kCapBig = 65536;
  const int qrows = kBM * group;
  const int sms = sm_count(dev);
  const int units = std::max(1, sms / group);
  const int64_t tiles = ceil_div(n, kBN);
  const int chunk_rows = std::min(q_pad, 4096 / qrows * qrows);
  B2VS_TRY(ws_big.reserve(static_cast<size_t>(chunk_rows) * kCapBig * sizeof(u64)));
  B2VS_TRY(ws_bigcnt.reserve((static_cast<size_t>(chunk_rows) + 1) * sizeof(int)));
  B2VS_TRY(ws_tau.reserve(static_cast<size_t>(q_pad) * sizeof(float)));
  int* counts = ws_bigcnt.as<int>();
  int* overflow = counts + chunk_rows;
  // stride schedule: first pass sees <= cap/2 rows, later passes expect <= cap/4 candidates
  std::vector<int> strides;
  strides.push_back(static_cast<int>(std::max<int64_t>(1, ceil_div(tiles * kBN, kCapBig / 2))));
  while (strides.back() > 1) {
    const int64_t nxt = std::max<int64_t>(1, ceil_div(4ll * k * strides.back(), kCapBig));
    strides.push_back(static_cast<int>(std::min<int64_t>(nxt, strides.back() - 1)));
  }
  for (int q0 = 0; q0 < nq; q0 += chunk_rows) {
    const int rows_c = std::min(chunk_rows, q_pad - q0);
    const int valid_c = std::min(rows_c, nq - q0);
    CUtensorMap tm_qc;
    B2VS_TRY(encode_tmap_2d(&tm_qc,
                            static_cast<const uint16_t*>(q_mat) + static_cast<size_t>(q0) * kdim,
                            ab_format, rows_c, kdim, kBM));
    BfTcParams p{};
    p.beta = beta.as<float>();
    p.n_qblocks = rows_c / qrows;
    p.q_pad = rows_c;
    p.nq = valid_c;
    p.k_blocks = static_cast<int>(ceil_div(kdim, kBK));
    p.k = k;
    p.alpha = (metric == B2VS_METRIC_L2) ? -2.f : -1.f;
    p.idesc = ptx::make_idesc_f16(static_cast<uint32_t>(ab_format), qrows, kBN);
    p.big_cand = ws_big.as<u64>();
    p.big_count = counts;
    p.big_cap = kCapBig;
    for (size_t pi = 0; pi < strides.size(); ++pi) {
      const int stride = strides[pi];
      const bool last = (pi + 1 == strides.size());
      const int64_t ptiles = ceil_div(tiles, stride);
      bool seeded = pi > 0;
      for (int attempt = 0; attempt < 8; ++attempt) {
        int n_splits = choose_splits(p.n_qblocks, ptiles, units, 1);
        const int tps = static_cast<int>(ceil_div(ptiles, n_splits));
        n_splits = static_cast<int>(ceil_div(ptiles, tps));
        p.n_items = p.n_qblocks * n_splits;
        p.tiles_total = static_cast<int>(ptiles);
        p.tiles_per_split = tps;
        p.tile_stride = stride;
        p.tau_init = seeded ? ws_tau.as<float>() : nullptr;
        const int grid = std::min(p.n_items, units) * group;
        B2VS_CUDA(cudaMemsetAsync(counts, 0, (static_cast<size_t>(chunk_rows) + 1) * sizeof(int), st));
        B2VS_TRY(launch_fused(group, grid, tm_qc, p, st));
        // k-th key per query: the next pass's threshold, or (last pass) the sorted answer
        B2VS_TRY(launch_bigk_select(ws_big.as<u64>(), counts, kCapBig, last ? valid_c : rows_c, k,
                                    last ? 1 : 0, metric, ws_qnorm.as<float>() + q0, id_offset,
                                    ws_tau.as<float>(),
                                    last ? out_d + static_cast<size_t>(q0) * k : nullptr,
                                    last ? out_i + static_cast<size_t>(q0) * k : nullptr, overflow, st));
        *launches += 2;
        int h_over = 0;
        B2VS_CUDA(cudaMemcpyAsync(&h_over, overflow, sizeof(int), cudaMemcpyDeviceToHost, st));
        B2VS_CUDA(cudaStreamSynchronize(st));
        if (h_over <= kCapBig) break;
        // a buffer overflowed: the retained keys still give a valid, tighter bound; publish it as
        // the threshold of every query of the chunk and repeat this pass
        B2VS_CHECK(attempt < 7, B2VS_ECUDA, "large-k candidate buffers keep overflowing (%d keys)",
                   h_over);
        B2VS_TRY(launch_bigk_select(ws_big.as<u64>(), counts, kCapBig, rows_c, k, 0, metric, nullptr,
                                    0, ws_tau.as<float>(), nullptr, nullptr, nullptr, st));
        *launches += 1;
        seeded = true;
      }
    }
  }
  return B2VS_OK;
}

}  // namespace b2vs
