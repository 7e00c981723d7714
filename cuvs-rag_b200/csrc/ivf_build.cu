// IVF index construction (IndexBuildingCoordinator._create_index -> b2vs_ivfflat_build /
// b2vs_ivfpq_build; reference call sites index_building_coordinator.py:392-404,
// improved_multi_gpu_rag.py:126-138): training subsample -> k-means (kmeans.cu) -> assignment of
// every row on the tensor cores -> K3 list construction -> (IVF-PQ) K6 codebooks + encoding.
#include "ivf_internal.cuh"

namespace b2vs {

// one warp per row: copy the row into its list slot (16-bit storage) and record ||x||^2
__global__ void fill_f32_kernel(float* __restrict__ p, size_t n, float v) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

template <typename T>
__global__ void fill_flat_lists_kernel(const T* __restrict__ x, int64_t n, int dim, int dp, int fmt,
                                       const uint32_t* __restrict__ slot_of_row,
                                       uint16_t* __restrict__ data, float* __restrict__ slot_norm,
                                       int want_norm, unsigned int* __restrict__ max_norm_bits) {
  const int lane = threadIdx.x & 31;
  float warp_max = 0.f;
  const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = warp0; r < n; r += nwarps) {
    const uint32_t slot = slot_of_row[r];
    if (slot == kNoRow) continue;
    const T* row = x + r * dim;
    uint16_t* orow = data + static_cast<size_t>(slot) * dp;
    float acc = 0.f;
    for (int j = lane; j < dp; j += 32) {
      const float v = j < dim ? ld_f32<T>(row + j) : 0.f;
      float back;
      orow[j] = to_op16(v, fmt, &back);
      acc = fmaf(back, back, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) slot_norm[slot] = want_norm ? acc : 0.f;  // inner product: no additive term
    warp_max = fmaxf(warp_max, acc);
  }
  // largest ||x||^2 of the index (non-negative floats order like their bit patterns)
  if (lane == 0 && warp_max > 0.f) atomicMax(max_norm_bits, __float_as_uint(warp_max));
}

// ---- K6 / K7 IVF-PQ -----------------------------------------------------------------------
// residual sub-vectors of the training rows, laid out [pq_dim][n_train][dsub] fp32
template <typename T>
__global__ void pq_train_slices_kernel(const T* __restrict__ x, const int* __restrict__ labels,
                                       const float* __restrict__ cent, int64_t n_train,
                                       int64_t stride, int dim, int dsub, float* __restrict__ out) {
  const int64_t total = n_train * dim;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t t = i / dim;
    const int j = static_cast<int>(i - t * dim);
    const int64_t r = t * stride;
    const int c = labels[r];
    const float res = ld_f32<T>(x + r * dim + j) - cent[static_cast<size_t>(c) * dim + j];
    const int m = j / dsub, d = j - m * dsub;
    out[(static_cast<size_t>(m) * n_train + t) * dsub + d] = res;
  }
}

// grid (row blocks, pq_dim): each block holds one sub-codebook in smem; thread = row
template <typename T>
__global__ void __launch_bounds__(256)
pq_encode_kernel(const T* __restrict__ x, const int* __restrict__ labels,
                 const float* __restrict__ cent, const float* __restrict__ codebooks,
                 const uint32_t* __restrict__ slot_of_row, int64_t n, int dim, int dsub, int mp,
                 uint8_t* __restrict__ codes) {
  extern __shared__ float cb[];  // [256][dsub]
  const int m = blockIdx.y;
  for (int i = threadIdx.x; i < 256 * dsub; i += blockDim.x)
    cb[i] = codebooks[static_cast<size_t>(m) * 256 * dsub + i];
  __syncthreads();
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < n;
       r += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint32_t slot = slot_of_row[r];
    if (slot == kNoRow) continue;
    const int c = labels[r];
    float res[16];
    for (int d = 0; d < dsub; ++d)
      res[d] = ld_f32<T>(x + r * dim + m * dsub + d) - cent[static_cast<size_t>(c) * dim + m * dsub + d];
    float best = __int_as_float(0x7f800000);
    int best_j = 0;
    for (int j = 0; j < 256; ++j) {
      float s = 0.f;
      for (int d = 0; d < dsub; ++d) {
        const float t = res[d] - cb[j * dsub + d];
        s = fmaf(t, t, s);
      }
      if (s < best) { best = s; best_j = j; }
    }
    codes[pq_code_offset(slot, m, mp)] = static_cast<uint8_t>(best_j);
  }
}

// Lists in descending-size order (host; sizes are known after the build / load).
int build_list_ranks(IvfData* d) {
  std::vector<int> order(d->n_lists), rank(d->n_lists);
  for (int i = 0; i < d->n_lists; ++i) order[i] = i;
  if (!env().ivf_no_rank)  // A/B knob: keep list-id order
    std::stable_sort(order.begin(), order.end(),
                     [&](int a, int b) { return d->h_sizes[a] > d->h_sizes[b]; });
  for (int r = 0; r < d->n_lists; ++r) rank[order[r]] = r;
  d->max_list_rows = 0;
  for (int v : d->h_sizes) d->max_list_rows = std::max(d->max_list_rows, static_cast<int>(round_up(v, 32)));
  const size_t bytes = static_cast<size_t>(d->n_lists) * sizeof(int);
  B2VS_TRY(d->rank_of_list.reserve(bytes));
  B2VS_TRY(d->list_of_rank.reserve(bytes));
  B2VS_CUDA(cudaMemcpy(d->rank_of_list.ptr, rank.data(), bytes, cudaMemcpyHostToDevice));
  B2VS_CUDA(cudaMemcpy(d->list_of_rank.ptr, order.data(), bytes, cudaMemcpyHostToDevice));
  return B2VS_OK;
}

// ---- K7b grouped IVF-PQ scan: pieces around pq_tc_kernel (pq_tc.cuh) ------------------------
// bf16 copy of the codebooks + squared norm of every rounded entry
__global__ void pq_cb16_kernel(const float* __restrict__ codebooks, int entries, int dsub,
                               uint16_t* __restrict__ cb16, float* __restrict__ cbn) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= entries) return;
  float n2 = 0.f;
  for (int d = 0; d < dsub; ++d) {
    float back;
    cb16[static_cast<size_t>(e) * dsub + d] = to_op16(codebooks[static_cast<size_t>(e) * dsub + d], 1, &back);
    n2 = fmaf(back, back, n2);
  }
  cbn[e] = n2;
}

// transposed bf16 codebooks [256][pq_dim][dsub]: what the grouped scan's decoder warps look up
// (lanes = consecutive sub-spaces of one list row -> consecutive shared-memory banks)
__global__ void pq_cb16t_kernel(const uint16_t* __restrict__ cb16, int pq_dim, int dsub,
                                uint16_t* __restrict__ cb16t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // over [pq_dim][256][dsub]
  if (i >= pq_dim * 256 * dsub) return;
  const int d = i % dsub, j = (i / dsub) & 255, m = i / (dsub * 256);
  cb16t[(static_cast<size_t>(j) * pq_dim + m) * dsub + d] = cb16[i];
}

// ||decoded residual||^2 of every slot (thread = slot); +inf on padding slots and in the slack
__global__ void pq_slot_norms_kernel(const uint8_t* __restrict__ codes, const uint32_t* __restrict__ row_ids,
                                     uint32_t n_slots, size_t n_out, int pq_dim, int mp,
                                     const float* __restrict__ cbn, int l2, float* __restrict__ out,
                                     unsigned int* __restrict__ max_bits) {
  const size_t slot = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  float n2 = 0.f;
  const bool real = slot < n_slots && row_ids[slot] != kNoRow;
  if (real) {
    for (int m = 0; m < pq_dim; ++m)
      n2 += __ldg(cbn + m * 256 + __ldg(codes + pq_code_offset(slot, m, mp)));
  }
  if (slot < n_out) out[slot] = real ? (l2 ? n2 : 0.f) : INFINITY;
  const unsigned int wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(n2));
  if ((threadIdx.x & 31) == 0 && wmax != 0u) atomicMax(max_bits, wmax);
}

// Derives the grouped scan's operands from codebooks + codes (after a build or a load).
int pq_prepare_grouped(b2vs_index* index, IvfData* d, cudaStream_t st) {
  d->pq_tc_ready = false;
  if (!pq_grouped_supported(index->dim, d->dsub) || d->mp != d->pq_dim) return B2VS_OK;
  const int entries = d->pq_dim * 256;
  const size_t n_out = static_cast<size_t>(std::max<int64_t>(d->n_slots, 1)) + kNormSlack;
  B2VS_TRY(d->cb16.reserve(static_cast<size_t>(entries) * d->dsub * 2));
  B2VS_TRY(d->cb16t.reserve(static_cast<size_t>(entries) * d->dsub * 2));
  B2VS_TRY(d->cbn.reserve(static_cast<size_t>(entries) * sizeof(float)));
  B2VS_TRY(d->pq_norm.reserve(n_out * sizeof(float)));
  DevBuf cell;
  B2VS_TRY(cell.reserve(sizeof(unsigned int)));
  B2VS_CUDA(cudaMemsetAsync(cell.ptr, 0, sizeof(unsigned int), st));
  pq_cb16_kernel<<<static_cast<unsigned>(ceil_div(entries, 256)), 256, 0, st>>>(
      d->codebooks.as<float>(), entries, d->dsub, d->cb16.as<uint16_t>(), d->cbn.as<float>());
  pq_cb16t_kernel<<<static_cast<unsigned>(ceil_div(entries * d->dsub, 256)), 256, 0, st>>>(
      d->cb16.as<uint16_t>(), d->pq_dim, d->dsub, d->cb16t.as<uint16_t>());
  pq_slot_norms_kernel<<<static_cast<unsigned>(ceil_div(n_out, 256)), 256, 0, st>>>(
      d->codes.as<uint8_t>(), d->row_ids.as<uint32_t>(), static_cast<uint32_t>(d->n_slots), n_out,
      d->pq_dim, d->mp, d->cbn.as<float>(), index->metric == B2VS_METRIC_L2 ? 1 : 0,
      d->pq_norm.as<float>(), cell.as<unsigned int>());
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(&d->max_rhat2, cell.ptr, sizeof(float), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cell.release();
  B2VS_CHECK(e == cudaSuccess, B2VS_ECUDA, "preparing the grouped PQ scan failed: %s", cudaGetErrorString(e));
  d->pq_tc_ready = true;
  return B2VS_OK;
}

void ivf_fill_info(const b2vs_index* index, b2vs_index_info* info) {
  const IvfData* d = static_cast<const IvfData*>(index->ivf);
  if (!d) return;
  info->n_lists = d->n_lists;
  info->pq_dim = d->pq_dim;
  info->pq_bits = d->pq_bits;
  info->device_bytes += static_cast<int64_t>(d->owned_bytes());
}

void ivf_last_stats(const b2vs_index* index, b2vs_search_stats* stats) {
  IvfData* d = static_cast<IvfData*>(index->ivf);
  *stats = b2vs_search_stats{};
  if (!d) return;
  if (d->counter_pending && d->ws_counter.ptr) {
    unsigned long long cnt[3] = {0, 0, 0};
    DeviceGuard guard(index->dev);
    if (cudaMemcpy(cnt, d->ws_counter.ptr, sizeof(cnt), cudaMemcpyDeviceToHost) == cudaSuccess) {
      d->stats.algo_bytes = static_cast<double>(cnt[0]) * d->row_bytes;
      d->stats.distinct_bytes = static_cast<double>(cnt[2]) * d->row_bytes;
      d->stats.mean_candidates = static_cast<int32_t>(cnt[1] / static_cast<unsigned long long>(std::max(d->last_nq, 1)));
    }
    d->counter_pending = false;
  }
  if (d->timing_pending && d->ev1) {
    float ms = 0.f;
    DeviceGuard guard(index->dev);
    if (cudaEventSynchronize(d->ev1) == cudaSuccess &&
        cudaEventElapsedTime(&ms, d->ev0, d->ev1) == cudaSuccess)
      d->stats.kernel_ms = ms;
    else
      cudaGetLastError();
    d->timing_pending = false;
  }
  *stats = d->stats;
}

void ivf_destroy(b2vs_index* index) {
  IvfData* d = static_cast<IvfData*>(index->ivf);
  if (d) {
    d->destroy();
    delete d;
  }
  index->ivf = nullptr;
}

// ------------------------------------------------------------------------------------------
static int ivf_build(int kind, int dev, int metric, int dtype, int dim, const void* db, int64_t n,
                     int64_t id_offset, const b2vs_ivf_params* params, cudaStream_t st,
                     b2vs_index** out) {
  B2VS_TRY(check_matrix_args(dev, metric, dtype, dim, db, n, out));
  B2VS_CHECK(params != nullptr, B2VS_EINVAL, "IVF params are NULL");
  B2VS_CHECK(n >= 1 && n < (1ll << 31) - (1ll << 22), B2VS_EINVAL,
             "IVF build needs 1 <= n < 2^31 (n=%lld)", static_cast<long long>(n));
  B2VS_CHECK(dim <= 2048, B2VS_EUNSUP, "IVF supports dim <= 2048 (got %d)", dim);
  const int n_lists = params->n_lists;
  B2VS_CHECK(n_lists >= 1 && n_lists <= n, B2VS_EINVAL, "n_lists=%d must be in [1, n=%lld]", n_lists,
             static_cast<long long>(n));
  const int iters = params->kmeans_iters > 0 ? params->kmeans_iters : 20;
  const float frac = (params->train_fraction > 0.f && params->train_fraction <= 1.f)
                         ? params->train_fraction : 0.5f;
  int pq_dim = 0, dsub = 0;
  if (kind == B2VS_KIND_IVF_PQ) {
    pq_dim = params->pq_dim;
    B2VS_CHECK(params->pq_bits == 8 || params->pq_bits == 0, B2VS_EUNSUP,
               "only pq_bits=8 is supported (got %d)", params->pq_bits);
    B2VS_CHECK(pq_dim >= 1 && dim % pq_dim == 0, B2VS_EINVAL,
               "pq_dim=%d must divide dim=%d", pq_dim, dim);
    dsub = dim / pq_dim;
    B2VS_CHECK(dsub <= 16, B2VS_EUNSUP, "sub-vector length %d > 16 not supported", dsub);
    B2VS_CHECK(n >= 256, B2VS_EINVAL, "IVF-PQ needs at least 256 rows to train codebooks");
    B2VS_CHECK(pq_dim <= 200, B2VS_EUNSUP, "pq_dim=%d too large for the smem LUT", pq_dim);
  }
  DeviceGuard guard(dev);
  B2VS_CHECK(guard.ok, B2VS_ECUDA, "cannot select device %d", dev);

  b2vs_index* ix = new (std::nothrow) b2vs_index();
  IvfData* d = new (std::nothrow) IvfData();
  B2VS_CHECK(ix && d, B2VS_ENOMEM, "host allocation failed");
  ix->kind = kind; ix->dev = dev; ix->metric = metric; ix->dtype = dtype; ix->dim = dim;
  ix->n = n; ix->id_offset = id_offset; ix->ivf = d;
  d->n_lists = n_lists; d->n = n; d->pq_dim = pq_dim; d->pq_bits = pq_dim ? 8 : 0; d->dsub = dsub;
  d->mp = static_cast<int>(round_up(pq_dim, 16));
  d->dp = static_cast<int>(round_up(dim, 8));
  d->fmt = (dtype == B2VS_F16) ? 0 : 1;
  d->row_bytes = (kind == B2VS_KIND_IVF_FLAT) ? d->dp * 2 : pq_dim;
  if (kind == B2VS_KIND_IVF_PQ) d->src_rows = db;  // borrowed: only dereferenced by refine

  DevBuf train, labels, cursor, slot_of_row, slices;
  FlatEngine assign_eng;
  int rc = B2VS_OK;
  auto fail = [&](int code) {
    train.release(); labels.release(); cursor.release(); slot_of_row.release(); slices.release();
    assign_eng.destroy();
    ivf_destroy(ix);
    ix->flat.destroy();
    delete ix;
    return code;
  };
#define IB_TRY(expr) do { rc = (expr); if (rc != B2VS_OK) return fail(rc); } while (0)
#define IB_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); return fail(B2VS_ECUDA); } } while (0)

  // ---- 1. training subsample (every stride-th row, capped at 1024 rows per list)
  int64_t stride = std::max<int64_t>(1, static_cast<int64_t>(std::floor(1.0 / frac + 1e-6)));
  const int64_t cap = std::max<int64_t>(static_cast<int64_t>(n_lists) * 1024, 1);
  if (ceil_div(n, stride) > cap) stride = ceil_div(n, cap);
  int64_t n_train = ceil_div(n, stride);
  if (n_train < n_lists) { stride = 1; n_train = n; }
  const void* train_ptr = db;
  const int eb = elem_bytes(dtype);
  if (stride > 1) {
    IB_TRY(train.reserve(static_cast<size_t>(n_train) * dim * eb));
    IB_TRY(launch_strided_rows(db, train.ptr, dtype, n_train, stride, dim, st));
    train_ptr = train.ptr;
  }
  // ---- 2. coarse k-means
  IB_TRY(d->centroids.reserve(static_cast<size_t>(n_lists) * dim * sizeof(float)));
  IB_TRY(kmeans_fit_impl(dev, dtype, dim, train_ptr, n_train, n_lists, iters, params->seed,
                         d->centroids.as<float>(), nullptr, st));
  train.release();
  // ---- 3. assign every row (tensor-core GEMM + arg-min)
  const int force = (dtype == B2VS_F32) ? -1 : d->fmt;
  IB_TRY(assign_eng.init(dev, B2VS_METRIC_L2, B2VS_F32, dim, d->centroids.ptr, n_lists, st, force));
  IB_TRY(labels.reserve(static_cast<size_t>(n) * sizeof(int32_t)));
  IB_TRY(assign_eng.search(db, dtype, static_cast<int>(n), 1, 0, 0, nullptr, nullptr,
                           labels.as<int32_t>(), st));
  // ---- 4. lists: histogram -> scan -> scatter -> fill
  const int pad = 32;  // lists start on 32-slot boundaries (interleaved PQ groups; grouped flat scan)
  IB_TRY(d->sizes.reserve(static_cast<size_t>(n_lists) * sizeof(int)));
  IB_TRY(d->offsets.reserve(static_cast<size_t>(n_lists + 1) * sizeof(uint32_t)));
  IB_TRY(cursor.reserve(static_cast<size_t>(n_lists) * sizeof(int)));
  IB_TRY(slot_of_row.reserve(static_cast<size_t>(n) * sizeof(uint32_t)));
  IB_CUDA(cudaMemsetAsync(d->sizes.ptr, 0, static_cast<size_t>(n_lists) * sizeof(int), st));
  IB_CUDA(cudaMemsetAsync(cursor.ptr, 0, static_cast<size_t>(n_lists) * sizeof(int), st));
  const int eblocks = static_cast<int>(std::min<int64_t>(ceil_div(n, 256), 148 * 32));
  IB_TRY(launch_histogram(labels.as<int>(), n, d->sizes.as<int>(), eblocks, st));
  IB_TRY(launch_scan_sizes(d->sizes.as<int>(), n_lists, pad, d->offsets.as<uint32_t>(), st));
  uint32_t total_slots = 0;
  IB_CUDA(cudaMemcpyAsync(&total_slots, d->offsets.as<uint32_t>() + n_lists, sizeof(uint32_t),
                          cudaMemcpyDeviceToHost, st));
  IB_CUDA(cudaStreamSynchronize(st));
  d->n_slots = total_slots;
  d->h_sizes.resize(n_lists);
  IB_CUDA(cudaMemcpy(d->h_sizes.data(), d->sizes.ptr, static_cast<size_t>(n_lists) * sizeof(int),
                     cudaMemcpyDeviceToHost));
  IB_TRY(build_list_ranks(d));
  IB_TRY(d->row_ids.reserve(std::max<size_t>(total_slots, 1) * sizeof(uint32_t)));
  IB_CUDA(cudaMemsetAsync(d->row_ids.ptr, 0xFF, std::max<size_t>(total_slots, 1) * sizeof(uint32_t), st));
  IB_TRY(launch_scatter_rows(labels.as<int>(), n, d->offsets.as<uint32_t>(), cursor.as<int>(),
                             d->row_ids.as<uint32_t>(), slot_of_row.as<uint32_t>(), eblocks, st));
  const int wblocks = static_cast<int>(std::min<int64_t>(ceil_div(n, 8), 148 * 16));
  if (kind == B2VS_KIND_IVF_FLAT) {
    IB_TRY(d->data.reserve(std::max<size_t>(total_slots, 1) * d->dp * 2));
    IB_CUDA(cudaMemsetAsync(cursor.ptr, 0, sizeof(unsigned int), st));  // re-used as the max-norm cell
    const size_t norm_floats = std::max<size_t>(total_slots, 1) + kNormSlack;
    IB_TRY(d->slot_norm.reserve(norm_floats * sizeof(float)));
    IB_CUDA(cudaMemsetAsync(d->data.ptr, 0, std::max<size_t>(total_slots, 1) * d->dp * 2, st));
    fill_f32_kernel<<<static_cast<unsigned>(ceil_div(norm_floats, 256)), 256, 0, st>>>(
        d->slot_norm.as<float>(), norm_floats, INFINITY);
    DISPATCH_DTYPE(dtype, T, (fill_flat_lists_kernel<T><<<wblocks, 256, 0, st>>>(
                                 static_cast<const T*>(db), n, dim, d->dp, d->fmt,
                                 slot_of_row.as<uint32_t>(), d->data.as<uint16_t>(),
                                 d->slot_norm.as<float>(), metric == B2VS_METRIC_L2 ? 1 : 0,
                                 cursor.as<unsigned int>())));
    IB_CUDA(cudaGetLastError());
    IB_CUDA(cudaMemcpyAsync(&d->max_norm2, cursor.ptr, sizeof(float), cudaMemcpyDeviceToHost, st));
    IB_CUDA(cudaStreamSynchronize(st));
  } else {
    // ---- 5. PQ codebooks on residual sub-vectors of a training subset, then encode all rows
    int64_t pstride = std::max<int64_t>(1, n / 131072);
    int64_t p_train = n / pstride;  // rows 0, pstride, ... all < n
    if (p_train < 256) { pstride = 1; p_train = n; }
    IB_TRY(slices.reserve(static_cast<size_t>(p_train) * dim * sizeof(float)));
    const int sblocks = static_cast<int>(std::min<int64_t>(ceil_div(p_train * dim, 256), 148 * 32));
    DISPATCH_DTYPE(dtype, T, (pq_train_slices_kernel<T><<<sblocks, 256, 0, st>>>(
                                 static_cast<const T*>(db), labels.as<int>(), d->centroids.as<float>(),
                                 p_train, pstride, dim, dsub, slices.as<float>())));
    IB_CUDA(cudaGetLastError());
    IB_TRY(d->codebooks.reserve(static_cast<size_t>(pq_dim) * 256 * dsub * sizeof(float)));
    const int pq_iters = std::min(iters, 10);
    KmWorkspace km_ws;   // one set of temporaries for all sub-codebooks
    for (int m = 0; m < pq_dim; ++m) {
      rc = kmeans_fit_impl(dev, B2VS_F32, dsub, slices.as<float>() + static_cast<size_t>(m) * p_train * dsub,
                           p_train, 256, pq_iters, params->seed + 31ull * (m + 1),
                           d->codebooks.as<float>() + static_cast<size_t>(m) * 256 * dsub, nullptr, st,
                           &km_ws);
      if (rc != B2VS_OK) break;
    }
    km_ws.release();
    if (rc != B2VS_OK) return fail(rc);
    slices.release();
    const size_t code_bytes = static_cast<size_t>(std::max<uint32_t>(total_slots, 32)) * d->mp;
    IB_TRY(d->codes.reserve(code_bytes));
    IB_CUDA(cudaMemsetAsync(d->codes.ptr, 0, code_bytes, st));
    dim3 grid(static_cast<unsigned>(std::min<int64_t>(ceil_div(n, 256), 148 * 8)), pq_dim);
    DISPATCH_DTYPE(dtype, T, (pq_encode_kernel<T><<<grid, 256, 256 * dsub * sizeof(float), st>>>(
                                 static_cast<const T*>(db), labels.as<int>(), d->centroids.as<float>(),
                                 d->codebooks.as<float>(), slot_of_row.as<uint32_t>(), n, dim, dsub,
                                 d->mp, d->codes.as<uint8_t>())));
    IB_CUDA(cudaGetLastError());
    IB_TRY(pq_prepare_grouped(ix, d, st));
  }
  // ---- 6. coarse quantizer used at search time (index metric)
  if (metric == B2VS_METRIC_L2) {
    ix->flat = assign_eng;          // take ownership of the buffers
    assign_eng = FlatEngine();
  } else {
    IB_TRY(ix->flat.init(dev, metric, B2VS_F32, dim, d->centroids.ptr, n_lists, st, force));
  }
  IB_CUDA(cudaStreamSynchronize(st));
  labels.release(); cursor.release(); slot_of_row.release();
  assign_eng.destroy();
#undef IB_TRY
#undef IB_CUDA
  *out = ix;
  return B2VS_OK;
}


int launch_fill_f32(float* p, size_t n, float v, cudaStream_t st) {
  if (n == 0) return B2VS_OK;
  fill_f32_kernel<<<static_cast<unsigned>(ceil_div(static_cast<int64_t>(n), 256)), 256, 0, st>>>(p, n, v);
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}

}  // namespace b2vs

using namespace b2vs;

// B2VS_METRIC_COSINE: build the IP index over a unit-norm copy of the rows.  IVF-Flat copies what
// it needs into its lists, so the copy is dropped after the build; IVF-PQ keeps it (refine).
static int build_entry(int kind, int dev, int metric, int dtype, int dim, const void* db, int64_t n,
                       int64_t id_offset, const b2vs_ivf_params* params, void* stream, b2vs_index** out) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (metric != B2VS_METRIC_COSINE)
    return ivf_build(kind, dev, metric, dtype, dim, db, n, id_offset, params, st, out);
  B2VS_TRY(check_matrix_args(dev, metric, dtype, dim, db, n, out));
  DeviceGuard guard(dev);
  B2VS_CHECK(guard.ok, B2VS_ECUDA, "cannot select device %d", dev);
  DevBuf unit;
  int rc = cosine_rows(dtype, dim, db, n, st, &unit);
  if (rc == B2VS_OK) rc = ivf_build(kind, dev, B2VS_METRIC_IP, dtype, dim, unit.ptr, n, id_offset, params, st, out);
  if (rc != B2VS_OK) {
    cudaStreamSynchronize(st);
    unit.release();
    return rc;
  }
  (*out)->cosine = true;
  if (kind == B2VS_KIND_IVF_PQ) {
    (*out)->cos_rows = unit;   // ownership moves to the index (src_rows points into it)
  } else {
    cudaStreamSynchronize(st);
    unit.release();
  }
  return B2VS_OK;
}

extern "C" int b2vs_ivfflat_build(int dev, int metric, int dtype, int dim, const void* db, int64_t n,
                                  int64_t id_offset, const b2vs_ivf_params* params, void* stream,
                                  b2vs_index** out) {
  return build_entry(B2VS_KIND_IVF_FLAT, dev, metric, dtype, dim, db, n, id_offset, params, stream, out);
}

extern "C" int b2vs_ivfpq_build(int dev, int metric, int dtype, int dim, const void* db, int64_t n,
                                int64_t id_offset, const b2vs_ivf_params* params, void* stream,
                                b2vs_index** out) {
  return build_entry(B2VS_KIND_IVF_PQ, dev, metric, dtype, dim, db, n, id_offset, params, stream, out);
}

extern "C" int b2vs_ivf_list_sizes_host(const b2vs_index* index, int32_t* sizes_host) {
  B2VS_CHECK(index && sizes_host, B2VS_EINVAL, "NULL argument");
  const IvfData* d = static_cast<const IvfData*>(index->ivf);
  B2VS_CHECK(d != nullptr, B2VS_EINVAL, "not an IVF index");
  std::copy(d->h_sizes.begin(), d->h_sizes.end(), sizes_host);
  return B2VS_OK;
}

extern "C" int b2vs_ivf_centroids_host(const b2vs_index* index, float* centroids_host) {
  B2VS_CHECK(index && centroids_host, B2VS_EINVAL, "NULL argument");
  const IvfData* d = static_cast<const IvfData*>(index->ivf);
  B2VS_CHECK(d != nullptr, B2VS_EINVAL, "not an IVF index");
  DeviceGuard guard(index->dev);
  B2VS_CUDA(cudaMemcpy(centroids_host, d->centroids.ptr,
                       static_cast<size_t>(d->n_lists) * index->dim * sizeof(float),
                       cudaMemcpyDeviceToHost));
  return B2VS_OK;
}
