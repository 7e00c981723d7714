// GPU k-means (the coarse-quantizer / PQ-codebook trainer inside IndexBuildingCoordinator's build,
// reference call sites index_building_coordinator.py:392-404) and the label-grouping kernels.
//   K1 assign  = the fused tensor-core engine with k = 1 (arg-min kept in a register, flat.cu)
//   K2 update  = segmented reduction: rows grouped by label (histogram -> scan -> scatter), one CTA
//                column per cluster sums its segment; balancing (size-ranked pairing + cuVS's
//                adjust_centers re-seed position) on the device, no host round trip per iteration
//   K3 grouping kernels, shared with the IVF list construction and the (query, probe) item sort
#include "ivf_internal.cuh"

namespace b2vs {

// ---- K-means pieces -----------------------------------------------------------------------
template <typename T>
__global__ void strided_rows_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t n_out,
                                    int64_t stride, int dim) {
  const int64_t total = n_out * dim;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / dim;
    const int j = static_cast<int>(i - r * dim);
    dst[i] = src[r * stride * dim + j];
  }
}

template <typename T>
__global__ void seed_centroids_kernel(const T* __restrict__ x, int64_t n, int dim, int ncl,
                                      uint64_t seed, float* __restrict__ cent) {
  const int c = blockIdx.x;
  // one distinct stratum per centroid, random offset inside it
  const int64_t lo = static_cast<int64_t>((static_cast<double>(c) * n) / ncl);
  const int64_t hi = static_cast<int64_t>((static_cast<double>(c + 1) * n) / ncl);
  const int64_t span = hi > lo ? hi - lo : 1;
  const int64_t row = min(n - 1, lo + static_cast<int64_t>(mix64(seed ^ (0x51ull * (c + 1))) % span));
  for (int j = threadIdx.x; j < dim; j += blockDim.x)
    cent[static_cast<size_t>(c) * dim + j] = ld_f32<T>(x + row * dim + j);
}

template <typename T>
__global__ void segment_sum_kernel(const T* __restrict__ x, const uint32_t* __restrict__ offsets,
                                   const uint32_t* __restrict__ row_ids, int dim,
                                   float* __restrict__ sums) {
  const int c = blockIdx.x;
  const int j = blockIdx.y * blockDim.x + threadIdx.x;
  if (j >= dim) return;
  const uint32_t begin = offsets[c], end = offsets[c + 1];
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  uint32_t i = begin;
  for (; i + 4 <= end; i += 4) {
    const uint32_t r0 = row_ids[i], r1 = row_ids[i + 1], r2 = row_ids[i + 2], r3 = row_ids[i + 3];
    a0 += ld_f32<T>(x + static_cast<size_t>(r0) * dim + j);
    a1 += ld_f32<T>(x + static_cast<size_t>(r1) * dim + j);
    a2 += ld_f32<T>(x + static_cast<size_t>(r2) * dim + j);
    a3 += ld_f32<T>(x + static_cast<size_t>(r3) * dim + j);
  }
  for (; i < end; ++i) a0 += ld_f32<T>(x + static_cast<size_t>(row_ids[i]) * dim + j);
  sums[static_cast<size_t>(c) * dim + j] = (a0 + a1) + (a2 + a3);
}

// ---- K3 list construction -----------------------------------------------------------------
__global__ void histogram_kernel(const int* __restrict__ labels, int64_t n, int* __restrict__ sizes) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = labels[i];
    if (c >= 0) atomicAdd(sizes + c, 1);
  }
}

// Exclusive scan of list sizes rounded up to `pad`; single block of 1024 threads.  A WARP owns a
// contiguous range and walks it 32 elements at a time, lane-strided, so every load and store is one
// coalesced 128-byte request (a thread-contiguous split made each request 32 sectors: 32 K sector
// transactions through one SM's LSU = 18 us for 16 K lists, twice per search); up to 16 steps of
// a warp's range stay in registers between the two passes.
__global__ void __launch_bounds__(1024)
scan_sizes_kernel(const int* __restrict__ sizes, int n_lists, int pad, uint32_t* __restrict__ offsets) {
  __shared__ uint32_t warp_tot[32];
  constexpr int kRegSteps = 16;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_warps = blockDim.x >> 5;
  const int steps = (n_lists + 32 * n_warps - 1) / (32 * n_warps);   // 32-element steps per warp
  const int base = warp * steps * 32;
  // pad is 1, 32 or 128 at every call site: a mask instead of an integer division per element
  // (the kernel runs on ONE SM: 100 K warp instructions were 15 us)
  const bool pow2 = (pad & (pad - 1)) == 0;
  const uint32_t pm = static_cast<uint32_t>(pad - 1);
  auto padded = [&](int i) -> uint32_t {
    if (i >= n_lists) return 0u;
    const uint32_t v = static_cast<uint32_t>(sizes[i]);
    return pow2 ? (v + pm) & ~pm : (v + pm) / static_cast<uint32_t>(pad) * static_cast<uint32_t>(pad);
  };
  uint32_t x[kRegSteps];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kRegSteps; ++k) x[k] = k < steps ? padded(base + k * 32 + lane) : 0u;   // all loads in flight
#pragma unroll
  for (int k = 0; k < kRegSteps; ++k) s += x[k];
  for (int k = kRegSteps; k < steps; ++k) s += padded(base + k * 32 + lane);                     // > 16 K lists
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);                      // warp total
  if (lane == 0) warp_tot[warp] = s;
  __syncthreads();
  if (warp == 0) {
    const uint32_t w = lane < n_warps ? warp_tot[lane] : 0u;
    uint32_t winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += u;
    }
    warp_tot[lane] = winc - w;                     // exclusive prefix of the warp totals
    if (lane == 31) offsets[n_lists] = winc;
  }
  __syncthreads();
  uint32_t carry = warp_tot[warp];
  auto emit = [&](int k, uint32_t v) {
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    const int i = base + k * 32 + lane;
    if (i < n_lists) offsets[i] = carry + inc - v;
    carry += __shfl_sync(0xffffffffu, inc, 31);
  };
#pragma unroll
  for (int k = 0; k < kRegSteps; ++k)
    if (k < steps) emit(k, x[k]);          // warp-uniform guard, static register index
  for (int k = kRegSteps; k < steps; ++k) emit(k, padded(base + k * 32 + lane));
}

__global__ void scatter_rows_kernel(const int* __restrict__ labels, int64_t n,
                                    const uint32_t* __restrict__ offsets, int* __restrict__ cursor,
                                    uint32_t* __restrict__ row_ids, uint32_t* __restrict__ slot_of_row,
                                    const int* __restrict__ deal_sizes, int deal_four) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = labels[i];
    if (c < 0) { slot_of_row[i] = kNoRow; continue; }
    uint32_t idx = static_cast<uint32_t>(atomicAdd(cursor + c, 1));
    if (deal_sizes) idx = group_row_pos(idx, deal_four ? 0u : static_cast<uint32_t>(deal_sizes[c]));   // grouped scans
    const uint32_t slot = offsets[c] + idx;
    row_ids[slot] = static_cast<uint32_t>(i);
    slot_of_row[i] = slot;
  }
}


// K2 update, second half: mean of each cluster (a cluster left empty keeps its previous centre).
__global__ void finalize_centroids_kernel(int dim, const float* __restrict__ sums,
                                          const int* __restrict__ counts, float* __restrict__ cent) {
  const int c = blockIdx.x;
  const int cnt = counts[c];
  if (cnt <= 0) return;
  const float inv = 1.f / static_cast<float>(cnt);
  for (int j = threadIdx.x; j < dim; j += blockDim.x)
    cent[static_cast<size_t>(c) * dim + j] = sums[static_cast<size_t>(c) * dim + j] * inv;
}

// Balancing, step 1 - WHICH clusters move (one CTA).  Clusters are sorted by (size, id) ascending in
// global scratch;
// big cluster j (j-th from the top, size > 1.5 avg) wants quota_j = max(1, floor(size / avg) - 1)
// donors-in-reverse: the small clusters (size < 0.5 avg) at sorted positions
// [sum_{i<j} quota_i, + quota_j) are re-seeded inside it, as long as position < #small and
// < ncl-1-j - the closed form of oracle/ivf.py::balance_pairs' two-pointer walk.  No host round
// trip per Lloyd iteration.  (cuVS picks the large cluster through a random data row, i.e. with
// probability proportional to its size, and only moves clusters below a quarter of the average;
// ranking the clusters instead empties the over-full "hub" lists first: on a 1024-component
// mixture the largest list shrinks from 419 to 191 rows at 98 rows average, and the result is the
// same on structureless data - measured with the oracle)
constexpr int kBalanceThreads = 1024;
constexpr int kBalanceMaxClusters = 1 << 16;
__global__ void __launch_bounds__(kBalanceThreads)
balance_pairs_kernel(const int* __restrict__ counts, int ncl, long long n, u64* __restrict__ order,
                     int* __restrict__ qprefix, int* __restrict__ donor_of) {
  __shared__ int part[kBalanceThreads];
  __shared__ int s_small, s_big;
  const int t = threadIdx.x;
  int P = 1;
  while (P < ncl) P <<= 1;
  for (int i = t; i < P; i += kBalanceThreads)
    order[i] = i < ncl ? ((static_cast<u64>(static_cast<uint32_t>(counts[i])) << 32) | static_cast<uint32_t>(i))
                       : kKeyInf;
  for (int i = t; i < ncl; i += kBalanceThreads) donor_of[i] = -1;
  if (t == 0) { s_small = 0; s_big = 0; }
  __syncthreads();
  for (int size = 2; size <= P; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = t; i < (P >> 1); i += kBalanceThreads) {
        const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
        const bool up = (lo & size) == 0;
        const u64 a = order[lo], b = order[hi];
        if ((a > b) == up) { order[lo] = b; order[hi] = a; }
      }
      __syncthreads();
    }
  const double avg = static_cast<double>(n) / ncl;
  // #small = clusters below 0.5 avg (a prefix of the order), #big = clusters above 1.5 avg (a suffix)
  int my_small = 0, my_big = 0;
  for (int i = t; i < ncl; i += kBalanceThreads) {
    const int c = static_cast<int>(order[i] >> 32);
    my_small += (c < 0.5 * avg) ? 1 : 0;
    my_big += (c > 1.5 * avg) ? 1 : 0;
  }
  if (my_small) atomicAdd(&s_small, my_small);
  if (my_big) atomicAdd(&s_big, my_big);
  __syncthreads();
  const int n_small = s_small, n_big = s_big;
  if (n_small == 0 || n_big == 0) return;
  // exclusive prefix of the quotas of big clusters j = 0 .. n_big-1 (j-th largest)
  const int per = (n_big + kBalanceThreads - 1) / kBalanceThreads;
  const int j0 = min(n_big, t * per), j1 = min(n_big, j0 + per);
  int sum = 0;
  for (int j = j0; j < j1; ++j) {
    const int c = static_cast<int>(order[ncl - 1 - j] >> 32);
    int quota = static_cast<int>(c / avg) - 1;
    sum += quota < 1 ? 1 : quota;
  }
  part[t] = sum;
  __syncthreads();
  if (t == 0) {
    int run = 0;
    for (int i = 0; i < kBalanceThreads; ++i) { const int v = part[i]; part[i] = run; run += v; }
    qprefix[n_big] = run;
  }
  __syncthreads();
  int run = part[t];
  for (int j = j0; j < j1; ++j) {
    qprefix[j] = run;
    const int c = static_cast<int>(order[ncl - 1 - j] >> 32);
    int quota = static_cast<int>(c / avg) - 1;
    run += quota < 1 ? 1 : quota;
  }
  __syncthreads();
  for (int s = t; s < n_small; s += kBalanceThreads) {
    // big cluster whose quota range holds sorted position s: last j with qprefix[j] <= s
    int lo = 0, hi = n_big;   // qprefix[0] = 0 <= s
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (qprefix[mid] <= s) lo = mid; else hi = mid;
    }
    const int j = lo;
    if (s < qprefix[j + 1] && s < ncl - 1 - j)
      donor_of[static_cast<uint32_t>(order[s])] = static_cast<int>(static_cast<uint32_t>(order[ncl - 1 - j]));
  }
}

// Balancing, step 2 - WHERE they move: the re-seed position of cuVS / RAFT balanced k-means
// "adjust_centers" (raft/cluster/detail/kmeans_balanced.cuh, restated - same rule as
// oracle/ivf.py::adjust_centers): next to the large cluster l, nudged towards one of its members,
//     centre[small] = (wc * centre[l] + x[i]) / (wc + 1),    wc = min(size[l], kAdjustWeight),
// with i a random row of l (taken from the label-grouped row table of the K2 update, so no
// rejection sampling).  Restarting ON a data row (round 1) isolates the new centre in high
// dimensions - ||x||^2 dominates its score and only the row itself joins: 60 % singleton lists on
// iid Gaussian rows.  Large clusters are never moved themselves, so reading centre[l] races with
// nothing.
constexpr float kAdjustWeight = 7.0f;
template <typename T>
__global__ void adjust_centers_kernel(const T* __restrict__ x, int dim, const int* __restrict__ counts,
                                      const int* __restrict__ donor_of,
                                      const uint32_t* __restrict__ seg_off,
                                      const uint32_t* __restrict__ seg_rows, uint64_t seed,
                                      float* __restrict__ cent) {
  const int c = blockIdx.x;
  const int l = donor_of[c];
  if (l < 0 || l == c) return;
  const int cnt_l = counts[l];
  if (cnt_l <= 0) return;
  const uint32_t pick = static_cast<uint32_t>(mix64(seed ^ (0xA5ull * (c + 1))) % static_cast<uint64_t>(cnt_l));
  const size_t row = seg_rows[seg_off[l] + pick];
  const float wc = fminf(static_cast<float>(cnt_l), kAdjustWeight);
  const float inv = 1.f / (wc + 1.f);
  for (int j = threadIdx.x; j < dim; j += blockDim.x)
    cent[static_cast<size_t>(c) * dim + j] =
        (wc * cent[static_cast<size_t>(l) * dim + j] + ld_f32<T>(x + row * dim + j)) * inv;
}

int kmeans_fit_impl(int dev, int dtype, int dim, const void* x, int64_t n, int ncl, int iters,
                           uint64_t seed, float* cent, int32_t* labels_out, cudaStream_t st,
                           KmWorkspace* shared_ws) {
  B2VS_CHECK(n >= 1 && ncl >= 1 && ncl <= n, B2VS_EINVAL,
             "k-means needs 1 <= n_clusters <= n (n_clusters=%d, n=%lld)", ncl,
             static_cast<long long>(n));
  B2VS_CHECK(n < (1ll << 31), B2VS_EINVAL, "k-means input too large (n=%lld)", static_cast<long long>(n));
  KmWorkspace local_ws;
  KmWorkspace& w = shared_ws ? *shared_ws : local_ws;
  DevBuf &sums = w.sums, &counts = w.counts, &labels = w.labels,
         &seg_off = w.seg_off, &seg_cur = w.seg_cur, &seg_rows = w.seg_rows, &seg_slot = w.seg_slot;
  FlatEngine& eng = w.eng;
  int rc = B2VS_OK;
  auto cleanup = [&]() {
    if (!shared_ws) local_ws.release();   // a shared workspace is released by its owner
  };
#define KM_TRY(expr) do { rc = (expr); if (rc != B2VS_OK) { cleanup(); return rc; } } while (0)
#define KM_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); cleanup(); return B2VS_ECUDA; } } while (0)
  KM_TRY(sums.reserve(static_cast<size_t>(ncl) * dim * sizeof(float)));
  KM_TRY(counts.reserve(static_cast<size_t>(ncl) * sizeof(int)));
  KM_TRY(seg_off.reserve(static_cast<size_t>(ncl + 1) * sizeof(uint32_t)));
  KM_TRY(seg_cur.reserve(static_cast<size_t>(ncl) * sizeof(int)));
  KM_TRY(seg_rows.reserve(static_cast<size_t>(n) * sizeof(uint32_t)));
  KM_TRY(seg_slot.reserve(static_cast<size_t>(n) * sizeof(uint32_t)));
  {
    size_t p2 = 1;
    while (p2 < static_cast<size_t>(ncl)) p2 <<= 1;
    KM_TRY(w.order.reserve(p2 * sizeof(u64)));
    KM_TRY(w.donor_scratch.reserve((static_cast<size_t>(ncl) + 1) * sizeof(int)));
    KM_TRY(w.donors.reserve(static_cast<size_t>(ncl) * sizeof(int)));
  }
  int32_t* lab = labels_out;
  if (!lab) {
    KM_TRY(labels.reserve(static_cast<size_t>(n) * sizeof(int32_t)));
    lab = labels.as<int32_t>();
  }
  const int fmt = (dtype == B2VS_F16) ? 0 : 1;
  const int force = (dtype == B2VS_F32) ? -1 : fmt;
  DISPATCH_DTYPE(dtype, T, (seed_centroids_kernel<T><<<ncl, 128, 0, st>>>(
                               static_cast<const T*>(x), n, dim, ncl, seed, cent)));
  KM_CUDA(cudaGetLastError());
  const int acc_blocks = static_cast<int>(std::min<int64_t>(ceil_div(n, 8), 148 * 16));
  for (int it = 0; it < iters; ++it) {
    KM_TRY(eng.init(dev, B2VS_METRIC_L2, B2VS_F32, dim, cent, ncl, st, force));
    KM_TRY(eng.search(x, dtype, static_cast<int>(n), 1, 0, 0, nullptr, nullptr, lab, st));
    KM_CUDA(cudaMemsetAsync(counts.ptr, 0, static_cast<size_t>(ncl) * sizeof(int), st));
    KM_CUDA(cudaMemsetAsync(seg_cur.ptr, 0, static_cast<size_t>(ncl) * sizeof(int), st));
    histogram_kernel<<<acc_blocks, 256, 0, st>>>(lab, n, counts.as<int>());
    scan_sizes_kernel<<<1, 1024, 0, st>>>(counts.as<int>(), ncl, 1, seg_off.as<uint32_t>());
    scatter_rows_kernel<<<acc_blocks, 256, 0, st>>>(lab, n, seg_off.as<uint32_t>(), seg_cur.as<int>(),
                                                    seg_rows.as<uint32_t>(), seg_slot.as<uint32_t>(), nullptr, 0);
    KM_CUDA(cudaGetLastError());
    {
      const int tpb = dim >= 256 ? 256 : ((dim + 31) / 32) * 32;
      const dim3 grid(ncl, static_cast<unsigned>(ceil_div(dim, tpb)));
      DISPATCH_DTYPE(dtype, T, (segment_sum_kernel<T><<<grid, tpb, 0, st>>>(
                                   static_cast<const T*>(x), seg_off.as<uint32_t>(),
                                   seg_rows.as<uint32_t>(), dim, sums.as<float>())));
    }
    KM_CUDA(cudaGetLastError());
    finalize_centroids_kernel<<<ncl, 128, 0, st>>>(dim, sums.as<float>(), counts.as<int>(), cent);
    KM_CUDA(cudaGetLastError());
    if (it + 1 < iters && ncl > 1 && ncl <= kBalanceMaxClusters) {   // the last iteration ends with the M step
      balance_pairs_kernel<<<1, kBalanceThreads, 0, st>>>(counts.as<int>(), ncl, static_cast<long long>(n),
                                                         w.order.as<u64>(), w.donor_scratch.as<int>(),
                                                         w.donors.as<int>());
      DISPATCH_DTYPE(dtype, T, (adjust_centers_kernel<T><<<ncl, 128, 0, st>>>(
                                   static_cast<const T*>(x), dim, counts.as<int>(), w.donors.as<int>(),
                                   seg_off.as<uint32_t>(), seg_rows.as<uint32_t>(),
                                   seed + 977ull * (it + 1), cent)));
      KM_CUDA(cudaGetLastError());
    }
  }
  if (labels_out) {
    KM_TRY(eng.init(dev, B2VS_METRIC_L2, B2VS_F32, dim, cent, ncl, st, force));
    KM_TRY(eng.search(x, dtype, static_cast<int>(n), 1, 0, 0, nullptr, nullptr, labels_out, st));
  }
  KM_CUDA(cudaStreamSynchronize(st));  // temporaries below are freed; make sure nothing is in flight
  cleanup();
#undef KM_TRY
#undef KM_CUDA
  return B2VS_OK;
}


int launch_histogram(const int* labels, int64_t n, int* sizes, int blocks, cudaStream_t st) {
  histogram_kernel<<<blocks, 256, 0, st>>>(labels, n, sizes);
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}
int launch_scan_sizes(const int* sizes, int n_lists, int pad, uint32_t* offsets, cudaStream_t st) {
  scan_sizes_kernel<<<1, 1024, 0, st>>>(sizes, n_lists, pad, offsets);
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}
int launch_scatter_rows(const int* labels, int64_t n, const uint32_t* offsets, int* cursor,
                        uint32_t* row_ids, uint32_t* slot_of_row, int blocks, cudaStream_t st,
                        const int* deal_sizes, int deal_four) {
  scatter_rows_kernel<<<blocks, 256, 0, st>>>(labels, n, offsets, cursor, row_ids, slot_of_row, deal_sizes,
                                              deal_four);
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}
int launch_strided_rows(const void* src, void* dst, int dtype, int64_t n_out, int64_t stride, int dim,
                        cudaStream_t st) {
  const int blocks = static_cast<int>(std::min<int64_t>(ceil_div(n_out * dim, 256), 148 * 32));
  DISPATCH_DTYPE(dtype, T, (strided_rows_kernel<T><<<blocks, 256, 0, st>>>(
                               static_cast<const T*>(src), static_cast<T*>(dst), n_out, stride, dim)));
  B2VS_CUDA(cudaGetLastError());
  return B2VS_OK;
}

}  // namespace b2vs

using namespace b2vs;

extern "C" int b2vs_kmeans_fit(int dev, int dtype, int dim, const void* x, int64_t n, int n_clusters,
                               int iters, uint64_t seed, float* centroids, int32_t* labels,
                               void* stream) {
  B2VS_CHECK(x != nullptr && centroids != nullptr, B2VS_EINVAL, "NULL pointer passed to b2vs_kmeans_fit");
  B2VS_CHECK(dtype == B2VS_F32 || dtype == B2VS_F16 || dtype == B2VS_BF16, B2VS_EINVAL,
             "unknown dtype %d", dtype);
  B2VS_CHECK(dim >= 1 && dim <= 16384, B2VS_EINVAL, "dim=%d outside [1, 16384]", dim);
  B2VS_CHECK(iters >= 0, B2VS_EINVAL, "iters must be >= 0");
  DeviceGuard guard(dev);
  B2VS_CHECK(guard.ok, B2VS_ECUDA, "cannot select device %d", dev);
  return kmeans_fit_impl(dev, dtype, dim, x, n, n_clusters, iters, seed, centroids, labels,
                         static_cast<cudaStream_t>(stream));
}
