// C ABI entry points (include/b2vs.h): error state, flat index, dispatch, host-buffer search.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <vector>

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

static int env_int_or(const char* name, int fallback) {
  const char* e = std::getenv(name);
  return (e && *e) ? std::atoi(e) : fallback;
}
static EnvConfig read_env() {
  EnvConfig c;
  if (const char* e = std::getenv("B2VS_TC_GROUP")) c.tc_group = (e[0] == '1' || e[0] == '2') ? e[0] - '0' : 0;
  if (const char* e = std::getenv("B2VS_PASSES")) {
    int v[3] = {0, 0, 0};
    const int got = std::sscanf(e, "%d,%d,%d", &v[0], &v[1], &v[2]);
    for (int i = 0; i < got && i < 3; ++i)
      if (v[i] >= 1) c.pass_s[c.pass_n++] = v[i];
    if (c.pass_n > 0 && c.pass_s[c.pass_n - 1] != 1) c.pass_n = 0;  // the last pass must be the full one
  }
  c.no_qpad = std::getenv("B2VS_NO_QPAD") != nullptr;
  c.ivf_no_rank = std::getenv("B2VS_IVF_NO_RANK") != nullptr;
  {
    const int v = env_int_or("B2VS_IVF_GROUPED_CAP", 0);
    if (v >= 32 && v <= 4096 && (v & (v - 1)) == 0) c.grouped_cap = v;
  }
  {
    const int v = env_int_or("B2VS_IVF_SEED_ROWS", 0);
    if (v >= 32) c.seed_rows = v;
  }
  if (const char* e = std::getenv("B2VS_IVF_GROUPED")) c.ivf_grouped = (e[0] == '0' || e[0] == '1') ? e[0] - '0' : -1;
  {
    const int v = env_int_or("B2VS_WORK_CHUNK_TILES", 0);
    if (v > 0) c.work_chunk_tiles = v;
  }
  c.debug_split = std::getenv("B2VS_DEBUG_SPLIT") != nullptr;
  if (const char* e = std::getenv("B2VS_COARSE_SCAN")) c.coarse_scan = e[0] == '0' ? 0 : 1;
  c.coarse_scan_maxq = env_int_or("B2VS_COARSE_SCAN_MAXQ", -1);
  c.coarse_scan_ctas = env_int_or("B2VS_COARSE_SCAN_CTAS", -1);
  c.no_item_sort = std::getenv("B2VS_NO_ITEM_SORT") != nullptr;
  if (const char* e = std::getenv("B2VS_GRAPH")) c.graph = e[0] == '0' ? 0 : 1;
  c.graph_maxq = std::max(0, env_int_or("B2VS_GRAPH_MAXQ", 0));
  if (const char* e = std::getenv("B2VS_TAIL_BOXES")) c.tail_boxes = e[0] == '0' ? 0 : 1;
  if (const char* e = std::getenv("B2VS_PLAN_OVERLAP")) c.plan_overlap = e[0] == '0' ? 0 : 1;
  if (const char* e = std::getenv("B2VS_IVF_SEED")) c.seed_mode = e[0] == '0' ? 0 : 1;
  {
    const int v = env_int_or("B2VS_IVF_SEED_LISTS", 0);
    if (v >= 1 && v <= 16) c.seed_lists = v;
    const int t = env_int_or("B2VS_IVF_SEED_TILE", 0);
    if (t >= 32 && t <= 256 && t % 32 == 0) c.seed_tile = t;
  }
  c.pq_debug = std::max(0, env_int_or("B2VS_PQ_DEBUG", 0));
  c.k0_debug = std::max(0, env_int_or("B2VS_K0_DEBUG", 0));
  if (const char* e = std::getenv("B2VS_SAMPLE_UNION")) c.sample_union = e[0] == '0' ? 0 : 1;
  if (const char* e = std::getenv("B2VS_A_QUARTERS")) c.a_quarters = e[0] == '0' ? 0 : 1;
  if (const char* e = std::getenv("B2VS_RAW_EMIT")) c.raw_emit = e[0] == '0' ? 0 : 1;
  if (const char* e = std::getenv("B2VS_WORK_EPI")) c.work_epi = (e[0] == '1' || e[0] == '2') ? e[0] - '0' : 0;
  if (const char* e = std::getenv("B2VS_TWO_PASS")) c.two_pass = e[0] == '0' ? 0 : 1;
  c.two_pass_chunk_mb = std::max(0, env_int_or("B2VS_TWO_PASS_CHUNK_MB", 0));
  if (const char* e = std::getenv("B2VS_CANARY")) c.canary = e[0] == '1';
  if (const char* e = std::getenv("B2VS_EPI_GROUPS")) c.epi_groups = (e[0] == '1' || e[0] == '2') ? e[0] - '0' : 0;
  return c;
}
static EnvConfig g_env;
static std::once_flag g_env_once;
static std::mutex g_env_mutex;
const EnvConfig& env() {
  std::call_once(g_env_once, [] { g_env = read_env(); });
  return g_env;
}

// ---- guard-zone registry (B2VS_CANARY=1) -------------------------------------------------------
static std::mutex g_canary_mutex;
static std::map<void*, size_t>& canary_map() {
  static std::map<void*, size_t> m;
  return m;
}
bool canary_enabled() {
  static const bool on = [] { const char* e = std::getenv("B2VS_CANARY"); return e && e[0] == '1'; }();
  return on;   // fixed for the life of the process: allocations and frees must agree on the layout
}
void canary_register(void* user_ptr, size_t bytes) {
  std::lock_guard<std::mutex> lock(g_canary_mutex);
  canary_map()[user_ptr] = bytes;
}
void canary_unregister(void* user_ptr) {
  std::lock_guard<std::mutex> lock(g_canary_mutex);
  canary_map().erase(user_ptr);
}

static bool valid_dtype(int d) { return d == B2VS_F32 || d == B2VS_F16 || d == B2VS_BF16; }
static bool valid_metric(int m) {
  return m == B2VS_METRIC_L2 || m == B2VS_METRIC_IP || m == B2VS_METRIC_COSINE;
}

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

// B2VS_METRIC_COSINE build step: unit-norm copy of the rows (owned by `buf`).
int cosine_rows(int dtype, int dim, const void* db, int64_t n, cudaStream_t st, DevBuf* buf) {
  B2VS_TRY(buf->reserve(static_cast<size_t>(std::max<int64_t>(n, 1)) * dim * elem_bytes(dtype)));
  return launch_unit_rows(db, buf->ptr, dtype, n, dim, st);
}

}  // namespace b2vs

using namespace b2vs;

extern "C" const char* b2vs_last_error(void) { return g_err; }
extern "C" int b2vs_version(void) { return B2VS_VERSION; }

extern "C" int b2vs_reload_env(void) {
  env();  // make sure the once-flag is spent before overwriting
  std::lock_guard<std::mutex> lock(g_env_mutex);
  g_env = read_env();
  note_realloc();   // captured search graphs baked the old switches in: have them re-captured
  return B2VS_OK;
}

extern "C" int b2vs_debug_check_canaries(int* n_buffers, int* n_corrupt) {
  B2VS_CHECK(n_buffers && n_corrupt, B2VS_EINVAL, "NULL argument");
  *n_buffers = 0;
  *n_corrupt = 0;
  B2VS_CHECK(canary_enabled(), B2VS_EUNSUP, "guard zones are off: start the process with B2VS_CANARY=1");
  std::lock_guard<std::mutex> lock(g_canary_mutex);
  std::vector<unsigned char> host(2 * kCanaryBytes);
  for (const auto& kv : canary_map()) {
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, kv.first) != cudaSuccess) { cudaGetLastError(); continue; }
    DeviceGuard guard(attr.device);
    cudaDeviceSynchronize();
    char* user = static_cast<char*>(kv.first);
    if (cudaMemcpy(host.data(), user - kCanaryBytes, kCanaryBytes, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(host.data() + kCanaryBytes, user + kv.second, kCanaryBytes, cudaMemcpyDeviceToHost) !=
            cudaSuccess) {
      cudaGetLastError();
      ++*n_corrupt;
      continue;
    }
    ++*n_buffers;
    bool bad = false;
    for (unsigned char b : host) bad = bad || b != 0xA5;
    if (bad) {
      ++*n_corrupt;
      set_error("guard zone of a %zu-byte device buffer at %p was overwritten", kv.second, kv.first);
    }
  }
  return B2VS_OK;
}

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
  ix->cosine = metric == B2VS_METRIC_COSINE;
  ix->metric = ix->cosine ? B2VS_METRIC_IP : metric;
  ix->dtype = dtype;
  ix->dim = dim;
  ix->n = n;
  ix->id_offset = id_offset;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = B2VS_OK;
  if (ix->cosine) {
    rc = cosine_rows(dtype, dim, db, n, st, &ix->cos_rows);
    db = ix->cos_rows.ptr;
  }
  if (rc == B2VS_OK) rc = ix->flat.init(dev, ix->metric, dtype, dim, db, n, st);
  if (rc != B2VS_OK) {
    ix->flat.destroy();
    ix->cos_rows.release();
    delete ix;
    return rc;
  }
  *out = ix;
  return B2VS_OK;
}

extern "C" int b2vs_search(b2vs_index* index, const void* queries, int q_dtype, int nq, int dim, int k,
                           const b2vs_search_params* params, float* out_d, int64_t* out_i,
                           void* stream) {
  B2VS_CHECK(index != nullptr, B2VS_EINVAL, "index is NULL");
  B2VS_CHECK(queries != nullptr && out_d != nullptr && out_i != nullptr, B2VS_EINVAL,
             "queries / output pointer is NULL");
  B2VS_CHECK(valid_dtype(q_dtype), B2VS_EINVAL, "unknown query dtype %d", q_dtype);
  B2VS_CHECK(nq >= 1, B2VS_EINVAL, "nq must be positive (got %d)", nq);
  B2VS_CHECK(k >= 1, B2VS_EINVAL, "k must be positive (got %d)", k);
  B2VS_CHECK(dim == index->dim, B2VS_EINVAL, "queries have dim %d, the index has dim %d", dim, index->dim);
  DeviceGuard guard(index->dev);
  B2VS_CHECK(guard.ok, B2VS_ECUDA, "cannot select device %d", index->dev);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  b2vs_search_params sp{};
  if (params) sp = *params;
  if (index->cosine) {
    // unit-norm copy of the batch (same dtype), then the IP engine; 1 - similarity at the end
    B2VS_TRY(index->cos_q.reserve(static_cast<size_t>(nq) * dim * elem_bytes(q_dtype)));
    B2VS_TRY(launch_unit_rows(queries, index->cos_q.ptr, q_dtype, nq, dim, st));
    queries = index->cos_q.ptr;
  }
  int rc;
  switch (index->kind) {
    case B2VS_KIND_FLAT:
      rc = index->flat.search(queries, q_dtype, nq, k, sp.n_splits, index->id_offset, out_d, out_i,
                              nullptr, st, sp.flags);
      break;
    case B2VS_KIND_IVF_FLAT:
    case B2VS_KIND_IVF_PQ:
      rc = ivf_search(index, queries, q_dtype, nq, k, sp, out_d, out_i, st);
      break;
    default:
      set_error("unknown index kind %d", index->kind);
      return B2VS_EINVAL;
  }
  B2VS_TRY(rc);
  if (index->cosine) B2VS_TRY(launch_cosine_fixup(out_d, static_cast<int64_t>(nq) * k, st));
  return B2VS_OK;
}

extern "C" int b2vs_search_host(b2vs_index* index, const void* queries_host, int q_dtype, int nq,
                                int dim, int k, const b2vs_search_params* params, float* out_d_host,
                                int64_t* out_i_host, void* stream) {
  B2VS_CHECK(index != nullptr, B2VS_EINVAL, "index is NULL");
  B2VS_CHECK(queries_host && out_d_host && out_i_host, B2VS_EINVAL,
             "queries / output pointer is NULL");
  B2VS_CHECK(valid_dtype(q_dtype), B2VS_EINVAL, "unknown query dtype %d", q_dtype);
  B2VS_CHECK(nq >= 1 && k >= 1, B2VS_EINVAL, "nq and k must be positive (nq=%d k=%d)", nq, k);
  B2VS_CHECK(dim == index->dim, B2VS_EINVAL, "queries have dim %d, the index has dim %d", dim, index->dim);
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
  B2VS_TRY(b2vs_search(index, base, q_dtype, nq, dim, k, params, d_dev, i_dev, stream));
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
  info->metric = index->cosine ? B2VS_METRIC_COSINE : index->metric;
  info->dtype = index->dtype;
  info->dim = index->dim;
  info->n_rows = index->n;
  info->id_offset = index->id_offset;
  info->device_bytes = static_cast<int64_t>(index->flat.owned_bytes() + index->cos_rows.bytes);
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
  index->cos_rows.release();
  index->cos_q.release();
  delete index;
  return B2VS_OK;
}
