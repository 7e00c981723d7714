// C ABI entry points (include/b2vs.h): error state, flat index, dispatch, host-buffer search.
#include <atomic>
#include <cstring>
#include <new>

#include "common.h"
#include "ivf.h"

namespace b2vs {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<uint64_t> g_realloc_generation{1};
uint64_t realloc_generation() { return g_realloc_generation.load(std::memory_order_acquire); }
void note_realloc() { g_realloc_generation.fetch_add(1, std::memory_order_acq_rel); }

static bool valid_dtype(int d) { return d == B2VS_F32 || d == B2VS_F16 || d == B2VS_BF16; }
static bool valid_metric(int m) { return m == B2VS_METRIC_L2 || m == B2VS_METRIC_IP; }

int check_matrix_args(int dev, int metric, int dtype, int dim, const void* db, int64_t n,
                      b2vs_index** out) {
  B2VS_CHECK(out != nullptr, B2VS_EINVAL, "out index pointer is NULL");
  *out = nullptr;
  B2VS_CHECK(valid_metric(metric), B2VS_EINVAL, "unknown metric %d", metric);
  B2VS_CHECK(valid_dtype(dtype), B2VS_EINVAL, "unknown dtype %d", dtype);
  B2VS_CHECK(dim >= 1 && dim <= 16384, B2VS_EINVAL, "dim=%d outside [1, 16384]", dim);
  B2VS_CHECK(n >= 0 && n < (1ll << 32), B2VS_EINVAL, "n=%lld outside [0, 2^32)",
             static_cast<long long>(n));
  B2VS_CHECK(n == 0 || db != nullptr, B2VS_EINVAL, "database pointer is NULL");
  int count = 0;
  B2VS_CUDA(cudaGetDeviceCount(&count));
  B2VS_CHECK(dev >= 0 && dev < count, B2VS_EINVAL, "device %d not in [0, %d)", dev, count);
  return B2VS_OK;
}

}  // namespace b2vs

using namespace b2vs;

extern "C" const char* b2vs_last_error(void) { return g_err; }
extern "C" int b2vs_version(void) { return B2VS_VERSION; }

extern "C" int b2vs_device_count(int* count) {
  B2VS_CHECK(count != nullptr, B2VS_EINVAL, "count pointer is NULL");
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) {
    cudaGetLastError();
    c = 0;
  }
  *count = c;
  return B2VS_OK;
}

extern "C" int b2vs_bf_create(int dev, int metric, int dtype, int dim, const void* db, int64_t n,
                              int64_t id_offset, void* stream, b2vs_index** out) {
  B2VS_TRY(check_matrix_args(dev, metric, dtype, dim, db, n, out));
  DeviceGuard guard(dev);
  B2VS_CHECK(guard.ok, B2VS_ECUDA, "cannot select device %d", dev);
  b2vs_index* ix = new (std::nothrow) b2vs_index();
  B2VS_CHECK(ix != nullptr, B2VS_ENOMEM, "host allocation failed");
  ix->kind = B2VS_KIND_FLAT;
  ix->dev = dev;
  ix->metric = metric;
  ix->dtype = dtype;
  ix->dim = dim;
  ix->n = n;
  ix->id_offset = id_offset;
  int rc = ix->flat.init(dev, metric, dtype, dim, db, n, static_cast<cudaStream_t>(stream));
  if (rc != B2VS_OK) {
    ix->flat.destroy();
    delete ix;
    return rc;
  }
  *out = ix;
  return B2VS_OK;
}

extern "C" int b2vs_search(b2vs_index* index, const void* queries, int q_dtype, int nq, int k,
                           const b2vs_search_params* params, float* out_d, int64_t* out_i,
                           void* stream) {
  B2VS_CHECK(index != nullptr, B2VS_EINVAL, "index is NULL");
  B2VS_CHECK(queries != nullptr && out_d != nullptr && out_i != nullptr, B2VS_EINVAL,
             "queries / output pointer is NULL");
  B2VS_CHECK(valid_dtype(q_dtype), B2VS_EINVAL, "unknown query dtype %d", q_dtype);
  B2VS_CHECK(nq >= 1, B2VS_EINVAL, "nq must be positive (got %d)", nq);
  B2VS_CHECK(k >= 1, B2VS_EINVAL, "k must be positive (got %d)", k);
  DeviceGuard guard(index->dev);
  B2VS_CHECK(guard.ok, B2VS_ECUDA, "cannot select device %d", index->dev);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  b2vs_search_params sp{};
  if (params) sp = *params;
  switch (index->kind) {
    case B2VS_KIND_FLAT:
      return index->flat.search(queries, q_dtype, nq, k, sp.n_splits, index->id_offset, out_d,
                                out_i, nullptr, st, sp.flags);
    case B2VS_KIND_IVF_FLAT:
    case B2VS_KIND_IVF_PQ:
      return ivf_search(index, queries, q_dtype, nq, k, sp, out_d, out_i, st);
    default:
      set_error("unknown index kind %d", index->kind);
      return B2VS_EINVAL;
  }
}

extern "C" int b2vs_search_host(b2vs_index* index, const void* queries_host, int q_dtype, int nq,
                                int k, const b2vs_search_params* params, float* out_d_host,
                                int64_t* out_i_host, void* stream) {
  B2VS_CHECK(index != nullptr, B2VS_EINVAL, "index is NULL");
  B2VS_CHECK(queries_host && out_d_host && out_i_host, B2VS_EINVAL,
             "queries / output pointer is NULL");
  B2VS_CHECK(valid_dtype(q_dtype), B2VS_EINVAL, "unknown query dtype %d", q_dtype);
  B2VS_CHECK(nq >= 1 && k >= 1, B2VS_EINVAL, "nq and k must be positive (nq=%d k=%d)", nq, k);
  DeviceGuard guard(index->dev);
  B2VS_CHECK(guard.ok, B2VS_ECUDA, "cannot select device %d", index->dev);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static thread_local DevBuf* io = nullptr;  // per-thread staging, one per device in practice
  static thread_local int io_dev = -1;
  if (io == nullptr || io_dev != index->dev) {
    if (io) { io->release(); delete io; }
    io = new DevBuf();
    io_dev = index->dev;
  }
  const size_t qb = static_cast<size_t>(nq) * index->dim * elem_bytes(q_dtype);
  const size_t qb_al = static_cast<size_t>(round_up(static_cast<int64_t>(qb), 256));
  const size_t db = static_cast<size_t>(nq) * k * sizeof(float);
  const size_t db_al = static_cast<size_t>(round_up(static_cast<int64_t>(db), 256));
  const size_t ib = static_cast<size_t>(nq) * k * sizeof(int64_t);
  B2VS_TRY(io->reserve(qb_al + db_al + ib));
  char* base = io->as<char>();
  B2VS_CUDA(cudaMemcpyAsync(base, queries_host, qb, cudaMemcpyHostToDevice, st));
  float* d_dev = reinterpret_cast<float*>(base + qb_al);
  int64_t* i_dev = reinterpret_cast<int64_t*>(base + qb_al + db_al);
  B2VS_TRY(b2vs_search(index, base, q_dtype, nq, k, params, d_dev, i_dev, stream));
  B2VS_CUDA(cudaMemcpyAsync(out_d_host, d_dev, db, cudaMemcpyDeviceToHost, st));
  B2VS_CUDA(cudaMemcpyAsync(out_i_host, i_dev, ib, cudaMemcpyDeviceToHost, st));
  B2VS_CUDA(cudaStreamSynchronize(st));
  return B2VS_OK;
}

extern "C" int b2vs_index_info_get(const b2vs_index* index, b2vs_index_info* info) {
  B2VS_CHECK(index && info, B2VS_EINVAL, "NULL argument");
  std::memset(info, 0, sizeof(*info));
  info->kind = index->kind;
  info->device = index->dev;
  info->metric = index->metric;
  info->dtype = index->dtype;
  info->dim = index->dim;
  info->n_rows = index->n;
  info->id_offset = index->id_offset;
  info->device_bytes = static_cast<int64_t>(index->flat.owned_bytes());
  if (index->kind != B2VS_KIND_FLAT) ivf_fill_info(index, info);
  return B2VS_OK;
}

extern "C" int b2vs_index_last_stats(const b2vs_index* index, b2vs_search_stats* stats) {
  B2VS_CHECK(index && stats, B2VS_EINVAL, "NULL argument");
  if (index->kind == B2VS_KIND_FLAT) {
    DeviceGuard guard(index->dev);
    const_cast<b2vs_index*>(index)->flat.resolve_timing();
    *stats = index->flat.stats;
  } else {
    ivf_last_stats(index, stats);
  }
  return B2VS_OK;
}

extern "C" int b2vs_index_destroy(b2vs_index* index) {
  if (!index) return B2VS_OK;
  DeviceGuard guard(index->dev);
  if (index->kind != B2VS_KIND_FLAT) ivf_destroy(index);
  index->flat.destroy();
  delete index;
  return B2VS_OK;
}
