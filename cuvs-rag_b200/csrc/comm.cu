// Cross-shard exchange behind the C ABI (include/b2vs.h, "Cross-shard exchange"): NCCL
// communicator + the collectives of the sharded search, replacing the reference's host-side
// concatenation of per-GPU results (improved_multi_gpu_rag.py:251-277,
// cuvs-2gpu-main.ipynb:L1806-1832) and its query replication (`query.to(device)` per GPU,
// improved_multi_gpu_rag.py:217).
//
// NCCL is bound at run time (dlopen of libnccl.so.2 - the copy PyTorch already loaded is reused
// when the host process has one), so libb2vs.so carries no link-time dependency on it.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <mutex>
#include <new>

#include "common.h"
#include "ivf.h"
#include "topk.cuh"

namespace b2vs {

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

const NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // RTLD_NOLOAD first: share the host process's NCCL (one set of NVLink/NVLS resources)
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    api.handle = h;
#define B2VS_NCCL_SYM(field, name)                                             \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name));           \
  if (!api.field) return;
    B2VS_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    B2VS_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    B2VS_NCCL_SYM(CommInitAll, "ncclCommInitAll")
    B2VS_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    B2VS_NCCL_SYM(AllGather, "ncclAllGather")
    B2VS_NCCL_SYM(AllReduce, "ncclAllReduce")
    B2VS_NCCL_SYM(Send, "ncclSend")
    B2VS_NCCL_SYM(Recv, "ncclRecv")
    B2VS_NCCL_SYM(GroupStart, "ncclGroupStart")
    B2VS_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    B2VS_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef B2VS_NCCL_SYM
    api.ok = true;
  });
  return api;
}

#define B2VS_NCCL(call)                                                                     \
  do {                                                                                      \
    ncclResult_t r__ = (call);                                                              \
    if (r__ != ncclSuccess) {                                                               \
      ::b2vs::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, nccl().GetErrorString(r__)); \
      return B2VS_ECUDA;                                                                    \
    }                                                                                       \
  } while (0)

int require_nccl() {
  B2VS_CHECK(nccl().ok, B2VS_EUNSUP,
             "NCCL (libnccl.so.2) could not be loaded: the cross-shard exchange is unavailable");
  return B2VS_OK;
}

ncclDataType_t nccl_dtype(int dtype) {
  return dtype == B2VS_F32 ? ncclFloat32 : (dtype == B2VS_F16 ? ncclFloat16 : ncclBfloat16);
}

void part_even(int64_t n, int parts, int r, int64_t* b, int64_t* e) {
  const int64_t base = n / parts, rem = n % parts;
  *b = r * base + (r < rem ? r : rem);
  *e = *b + base + (r < rem ? 1 : 0);
}

}  // namespace

}  // namespace b2vs

struct b2vs_comm {
  ncclComm_t comm = nullptr;
  int n_ranks = 1, rank = 0, dev = 0;
  // grow-only workspaces of the exchange calls
  b2vs::DevBuf recv_d, recv_i, q_all, loc_d, loc_i, io, samp_all;
};

using namespace b2vs;

extern "C" int b2vs_partition_even(int64_t n, int n_parts, int rank, int64_t* begin, int64_t* end) {
  B2VS_CHECK(begin && end, B2VS_EINVAL, "NULL argument");
  B2VS_CHECK(n >= 0 && n_parts >= 1 && rank >= 0 && rank < n_parts, B2VS_EINVAL,
             "bad partition request (n=%lld, parts=%d, rank=%d)", static_cast<long long>(n), n_parts, rank);
  part_even(n, n_parts, rank, begin, end);
  return B2VS_OK;
}

extern "C" int b2vs_comm_unique_id(void* id128) {
  B2VS_CHECK(id128 != nullptr, B2VS_EINVAL, "NULL argument");
  B2VS_TRY(require_nccl());
  static_assert(sizeof(ncclUniqueId) == B2VS_UNIQUE_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  B2VS_NCCL(nccl().GetUniqueId(&id));
  std::memcpy(id128, &id, sizeof(id));
  return B2VS_OK;
}

extern "C" int b2vs_comm_init_rank(int dev, int n_ranks, int rank, const void* id128, b2vs_comm** out) {
  B2VS_CHECK(out != nullptr && id128 != nullptr, B2VS_EINVAL, "NULL argument");
  *out = nullptr;
  B2VS_CHECK(n_ranks >= 1 && rank >= 0 && rank < n_ranks, B2VS_EINVAL, "rank %d not in [0, %d)", rank, n_ranks);
  B2VS_TRY(require_nccl());
  DeviceGuard guard(dev);
  B2VS_CHECK(guard.ok, B2VS_ECUDA, "cannot select device %d", dev);
  b2vs_comm* c = new (std::nothrow) b2vs_comm();
  B2VS_CHECK(c != nullptr, B2VS_ENOMEM, "host allocation failed");
  c->n_ranks = n_ranks; c->rank = rank; c->dev = dev;
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  ncclResult_t r = nccl().CommInitRank(&c->comm, n_ranks, id, rank);
  if (r != ncclSuccess) {
    set_error("ncclCommInitRank(rank %d of %d, device %d) -> %s", rank, n_ranks, dev, nccl().GetErrorString(r));
    delete c;
    return B2VS_ECUDA;
  }
  *out = c;
  return B2VS_OK;
}

extern "C" int b2vs_comm_init_all(int n, const int* devs, b2vs_comm** comms) {
  B2VS_CHECK(comms != nullptr && devs != nullptr && n >= 1 && n <= 64, B2VS_EINVAL, "bad arguments");
  B2VS_TRY(require_nccl());
  ncclComm_t raw[64];
  B2VS_NCCL(nccl().CommInitAll(raw, n, devs));
  for (int i = 0; i < n; ++i) {
    b2vs_comm* c = new (std::nothrow) b2vs_comm();
    if (!c) {
      for (int j = 0; j < i; ++j) { delete comms[j]; comms[j] = nullptr; }
      for (int j = 0; j < n; ++j) nccl().CommDestroy(raw[j]);
      set_error("host allocation failed");
      return B2VS_ENOMEM;
    }
    c->comm = raw[i]; c->n_ranks = n; c->rank = i; c->dev = devs[i];
    comms[i] = c;
  }
  return B2VS_OK;
}

extern "C" int b2vs_comm_info(const b2vs_comm* comm, int* n_ranks, int* rank, int* dev) {
  B2VS_CHECK(comm != nullptr, B2VS_EINVAL, "comm is NULL");
  if (n_ranks) *n_ranks = comm->n_ranks;
  if (rank) *rank = comm->rank;
  if (dev) *dev = comm->dev;
  return B2VS_OK;
}

extern "C" int b2vs_comm_destroy(b2vs_comm* comm) {
  if (!comm) return B2VS_OK;
  DeviceGuard guard(comm->dev);
  cudaDeviceSynchronize();
  if (comm->comm && nccl().ok) nccl().CommDestroy(comm->comm);
  for (DevBuf* b : {&comm->recv_d, &comm->recv_i, &comm->q_all, &comm->loc_d, &comm->loc_i, &comm->io,
                    &comm->samp_all})
    b->release();
  delete comm;
  return B2VS_OK;
}

extern "C" int b2vs_allgather_queries(b2vs_comm* comm, const void* q_local, int q_dtype, int nq_total,
                                      int dim, void* q_all, void* stream) {
  B2VS_CHECK(comm && q_all, B2VS_EINVAL, "NULL argument");
  B2VS_CHECK(nq_total >= 1 && dim >= 1, B2VS_EINVAL, "nq_total and dim must be positive");
  DeviceGuard guard(comm->dev);
  B2VS_CHECK(guard.ok, B2VS_ECUDA, "cannot select device %d", comm->dev);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int G = comm->n_ranks;
  const size_t row = static_cast<size_t>(dim) * elem_bytes(q_dtype);
  int64_t b = 0, e = 0;
  part_even(nq_total, G, comm->rank, &b, &e);
  B2VS_CHECK(e == b || q_local != nullptr, B2VS_EINVAL, "q_local is NULL");
  if (G == 1) {
    if (q_local != q_all)
      B2VS_CUDA(cudaMemcpyAsync(q_all, q_local, static_cast<size_t>(nq_total) * row, cudaMemcpyDeviceToDevice, st));
    return B2VS_OK;
  }
  if (nq_total % G == 0) {
    B2VS_NCCL(nccl().AllGather(q_local, q_all, static_cast<size_t>(e - b) * row, ncclInt8, comm->comm, st));
    return B2VS_OK;
  }
  // ragged slices: one group of sends / receives straight into place
  B2VS_NCCL(nccl().GroupStart());
  for (int p = 0; p < G; ++p) {
    int64_t pb = 0, pe = 0;
    part_even(nq_total, G, p, &pb, &pe);
    if (e > b) B2VS_NCCL(nccl().Send(q_local, static_cast<size_t>(e - b) * row, ncclInt8, p, comm->comm, st));
    if (pe > pb)
      B2VS_NCCL(nccl().Recv(static_cast<char*>(q_all) + static_cast<size_t>(pb) * row,
                            static_cast<size_t>(pe - pb) * row, ncclInt8, p, comm->comm, st));
  }
  B2VS_NCCL(nccl().GroupEnd());
  return B2VS_OK;
}

extern "C" int b2vs_allgather_topk(b2vs_comm* comm, const float* d_local, const int64_t* i_local, int nq,
                                   int k, float* d_all, int64_t* i_all, void* stream) {
  B2VS_CHECK(comm && d_local && i_local && d_all && i_all, B2VS_EINVAL, "NULL argument");
  B2VS_CHECK(nq >= 1 && k >= 1, B2VS_EINVAL, "nq and k must be positive");
  DeviceGuard guard(comm->dev);
  B2VS_CHECK(guard.ok, B2VS_ECUDA, "cannot select device %d", comm->dev);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t cnt = static_cast<size_t>(nq) * k;
  if (comm->n_ranks == 1) {
    B2VS_CUDA(cudaMemcpyAsync(d_all, d_local, cnt * sizeof(float), cudaMemcpyDeviceToDevice, st));
    B2VS_CUDA(cudaMemcpyAsync(i_all, i_local, cnt * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    return B2VS_OK;
  }
  B2VS_NCCL(nccl().GroupStart());
  B2VS_NCCL(nccl().AllGather(d_local, d_all, cnt, ncclFloat32, comm->comm, st));
  B2VS_NCCL(nccl().AllGather(i_local, i_all, cnt, ncclInt64, comm->comm, st));
  B2VS_NCCL(nccl().GroupEnd());
  return B2VS_OK;
}

extern "C" int b2vs_allgather_merge_topk(b2vs_comm* comm, const float* d_local, const int64_t* i_local,
                                         int nq, int k, int k_out, int descending, float* out_d,
                                         int64_t* out_i, void* stream) {
  B2VS_CHECK(comm && out_d && out_i, B2VS_EINVAL, "NULL argument");
  B2VS_CHECK(nq >= 1 && k >= 1, B2VS_EINVAL, "nq and k must be positive");
  const size_t cnt = static_cast<size_t>(comm->n_ranks) * nq * k;
  {
    DeviceGuard guard(comm->dev);
    B2VS_CHECK(guard.ok, B2VS_ECUDA, "cannot select device %d", comm->dev);
    B2VS_TRY(comm->recv_d.reserve(cnt * sizeof(float)));
    B2VS_TRY(comm->recv_i.reserve(cnt * sizeof(int64_t)));
  }
  B2VS_TRY(b2vs_allgather_topk(comm, d_local, i_local, nq, k, comm->recv_d.as<float>(),
                               comm->recv_i.as<int64_t>(), stream));
  return b2vs_merge_topk(comm->dev, comm->recv_d.as<float>(), comm->recv_i.as<int64_t>(), comm->n_ranks, nq,
                         k, k_out, descending, out_d, out_i, stream);
}

extern "C" int b2vs_exchange_merge_topk(b2vs_comm* comm, const float* d_local, const int64_t* i_local,
                                        int nq, int k, int k_out, int descending, float* out_d,
                                        int64_t* out_i, void* stream) {
  B2VS_CHECK(comm && d_local && i_local, B2VS_EINVAL, "NULL argument");
  B2VS_CHECK(nq >= 1 && k >= 1, B2VS_EINVAL, "nq and k must be positive");
  DeviceGuard guard(comm->dev);
  B2VS_CHECK(guard.ok, B2VS_ECUDA, "cannot select device %d", comm->dev);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int G = comm->n_ranks;
  int64_t b = 0, e = 0;
  part_even(nq, G, comm->rank, &b, &e);
  const int rows = static_cast<int>(e - b);
  if (rows > 0) B2VS_CHECK(out_d && out_i, B2VS_EINVAL, "output pointer is NULL");
  if (G == 1)
    return b2vs_merge_topk(comm->dev, d_local, i_local, 1, nq, k, k_out, descending, out_d, out_i, stream);
  const size_t cnt = static_cast<size_t>(G) * std::max(rows, 1) * k;
  B2VS_TRY(comm->recv_d.reserve(cnt * sizeof(float)));
  B2VS_TRY(comm->recv_i.reserve(cnt * sizeof(int64_t)));
  float* rd = comm->recv_d.as<float>();
  int64_t* ri = comm->recv_i.as<int64_t>();
  B2VS_NCCL(nccl().GroupStart());
  for (int p = 0; p < G; ++p) {
    int64_t pb = 0, pe = 0;
    part_even(nq, G, p, &pb, &pe);
    const size_t scnt = static_cast<size_t>(pe - pb) * k;   // what rank p merges: its slice of MY lists
    if (scnt) {
      B2VS_NCCL(nccl().Send(d_local + static_cast<size_t>(pb) * k, scnt, ncclFloat32, p, comm->comm, st));
      B2VS_NCCL(nccl().Send(i_local + static_cast<size_t>(pb) * k, scnt, ncclInt64, p, comm->comm, st));
    }
    const size_t rcnt = static_cast<size_t>(rows) * k;
    if (rcnt) {
      B2VS_NCCL(nccl().Recv(rd + static_cast<size_t>(p) * rcnt, rcnt, ncclFloat32, p, comm->comm, st));
      B2VS_NCCL(nccl().Recv(ri + static_cast<size_t>(p) * rcnt, rcnt, ncclInt64, p, comm->comm, st));
    }
  }
  B2VS_NCCL(nccl().GroupEnd());
  if (rows == 0) return B2VS_OK;
  return b2vs_merge_topk(comm->dev, rd, ri, G, rows, k, k_out, descending, out_d, out_i, stream);
}

extern "C" int b2vs_allreduce_min_f32(b2vs_comm* comm, float* values, int64_t n, void* stream) {
  B2VS_CHECK(comm && values, B2VS_EINVAL, "NULL argument");
  if (comm->n_ranks == 1 || n <= 0) return B2VS_OK;
  DeviceGuard guard(comm->dev);
  B2VS_CHECK(guard.ok, B2VS_ECUDA, "cannot select device %d", comm->dev);
  B2VS_NCCL(nccl().AllReduce(values, values, static_cast<size_t>(n), ncclFloat32, ncclMin, comm->comm,
                             static_cast<cudaStream_t>(stream)));
  return B2VS_OK;
}

namespace b2vs {
namespace {
int tau_exchange_cb(void* ctx, float* tau, int64_t n, cudaStream_t st) {
  return b2vs_allreduce_min_f32(static_cast<b2vs_comm*>(ctx), tau, n, st);
}
// Union exchange of a sampled pass: all-gather every rank's k best sampled scores per query, then the
// k-th best of the union is everybody's threshold (merge.cu: union_kth_kernel).
int tau_union_cb(void* ctx, const float* scores, int64_t n, int k, float* tau, cudaStream_t st) {
  b2vs_comm* comm = static_cast<b2vs_comm*>(ctx);
  const size_t cnt = static_cast<size_t>(n) * k;
  B2VS_TRY(comm->samp_all.reserve(cnt * comm->n_ranks * sizeof(float)));
  B2VS_NCCL(nccl().AllGather(scores, comm->samp_all.ptr, cnt, ncclFloat32, comm->comm, st));
  return launch_union_kth(comm->samp_all.as<float>(), comm->n_ranks, static_cast<int>(n), k, tau, st);
}
bool union_exchange_enabled() { return env().sample_union != 0; }
}  // namespace
}  // namespace b2vs

extern "C" int b2vs_comm_register_index(b2vs_comm* comm, b2vs_index* index, void* stream) {
  B2VS_CHECK(comm && index, B2VS_EINVAL, "NULL argument");
  B2VS_CHECK(index->dev == comm->dev, B2VS_EINVAL, "index on device %d, communicator on device %d",
             index->dev, comm->dev);
  DeviceGuard guard(comm->dev);
  B2VS_CHECK(guard.ok, B2VS_ECUDA, "cannot select device %d", comm->dev);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  index->sharded_min_rows = 0;
  if (comm->n_ranks == 1) { index->sharded_min_rows = index->n; return B2VS_OK; }
  B2VS_TRY(comm->io.reserve(sizeof(float)));
  // shard sizes are < 2^32 rows: exact in the float MIN after >> 8 (256-row tile units)
  const float tiles = static_cast<float>(index->n >> 8);
  B2VS_CUDA(cudaMemcpyAsync(comm->io.ptr, &tiles, sizeof(float), cudaMemcpyHostToDevice, st));
  B2VS_TRY(b2vs_allreduce_min_f32(comm, comm->io.as<float>(), 1, stream));
  float min_tiles = 0.f;
  B2VS_CUDA(cudaMemcpyAsync(&min_tiles, comm->io.ptr, sizeof(float), cudaMemcpyDeviceToHost, st));
  B2VS_CUDA(cudaStreamSynchronize(st));
  index->sharded_min_rows = static_cast<int64_t>(min_tiles) << 8;
  return B2VS_OK;
}

extern "C" int b2vs_search_sharded(b2vs_comm* comm, b2vs_index* index, const void* q_local, int q_dtype,
                                   int nq_total, int dim, int k, const b2vs_search_params* params,
                                   float* out_d, int64_t* out_i, void* stream) {
  B2VS_CHECK(comm && index, B2VS_EINVAL, "NULL argument");
  B2VS_CHECK(index->dev == comm->dev, B2VS_EINVAL, "index on device %d, communicator on device %d",
             index->dev, comm->dev);
  B2VS_CHECK(nq_total >= 1 && k >= 1, B2VS_EINVAL, "nq_total and k must be positive");
  B2VS_CHECK(dim == index->dim, B2VS_EINVAL, "queries have dim %d, the index has dim %d", dim, index->dim);
  DeviceGuard guard(comm->dev);
  B2VS_CHECK(guard.ok, B2VS_ECUDA, "cannot select device %d", comm->dev);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t row = static_cast<size_t>(dim) * elem_bytes(q_dtype);
  const size_t cnt = static_cast<size_t>(nq_total) * k;
  B2VS_TRY(comm->q_all.reserve(static_cast<size_t>(nq_total) * row));
  B2VS_TRY(comm->loc_d.reserve(cnt * sizeof(float)));
  B2VS_TRY(comm->loc_i.reserve(cnt * sizeof(int64_t)));
  B2VS_TRY(b2vs_allgather_queries(comm, q_local, q_dtype, nq_total, dim, comm->q_all.ptr, stream));
  b2vs_search_params sp{};
  if (params) sp = *params;
  int rc;
  // Threshold exchange needs every rank to run the same pass schedule: it is derived from the
  // smallest shard's size, agreed once by b2vs_comm_register_index (unregistered indexes - and
  // jobs with an empty shard - keep private thresholds: no collective inside the search).
  // pooled samples (union exchange) when the sparser schedule still has a sampled pass, else the MIN
  // exchange of k-th scores on the single-shard schedule
  const bool pooled = union_exchange_enabled() && flat_exchanges_tau(index->sharded_min_rows, k, comm->n_ranks);
  if (index->kind == B2VS_KIND_FLAT && comm->n_ranks > 1 && index->n > 0 &&
      (pooled || flat_exchanges_tau(index->sharded_min_rows, k, 1))) {
    TauExchange tx{tau_exchange_cb, comm, index->sharded_min_rows, comm->n_ranks,
                   pooled ? tau_union_cb : nullptr};
    rc = index->flat.search(comm->q_all.ptr, q_dtype, nq_total, k, sp.n_splits, index->id_offset,
                            comm->loc_d.as<float>(), comm->loc_i.as<int64_t>(), nullptr, st, sp.flags, &tx);
  } else {
    rc = b2vs_search(index, comm->q_all.ptr, q_dtype, nq_total, dim, k, &sp, comm->loc_d.as<float>(),
                     comm->loc_i.as<int64_t>(), stream);
  }
  B2VS_TRY(rc);
  const int descending = (index->metric == B2VS_METRIC_IP && !index->cosine) ? 1 : 0;
  return b2vs_exchange_merge_topk(comm, comm->loc_d.as<float>(), comm->loc_i.as<int64_t>(), nq_total, k, k,
                                  descending, out_d, out_i, stream);
}

extern "C" int b2vs_search_sharded_host(b2vs_comm* comm, b2vs_index* index, const void* q_local_host,
                                        int q_dtype, int nq_total, int dim, int k,
                                        const b2vs_search_params* params, float* out_d_host,
                                        int64_t* out_i_host, void* stream) {
  B2VS_CHECK(comm && index, B2VS_EINVAL, "NULL argument");
  B2VS_CHECK(nq_total >= 1 && k >= 1 && dim >= 1, B2VS_EINVAL, "nq_total, dim and k must be positive");
  DeviceGuard guard(comm->dev);
  B2VS_CHECK(guard.ok, B2VS_ECUDA, "cannot select device %d", comm->dev);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int64_t b = 0, e = 0;
  part_even(nq_total, comm->n_ranks, comm->rank, &b, &e);
  const size_t rows = static_cast<size_t>(e - b);
  const size_t qb = rows * dim * elem_bytes(q_dtype);
  const size_t qb_al = static_cast<size_t>(round_up(static_cast<int64_t>(std::max<size_t>(qb, 16)), 256));
  const size_t db = rows * k * sizeof(float);
  const size_t db_al = static_cast<size_t>(round_up(static_cast<int64_t>(std::max<size_t>(db, 16)), 256));
  const size_t ib = rows * k * sizeof(int64_t);
  // a second staging buffer (not comm->io, which the agreement step of search_sharded uses)
  static thread_local DevBuf* stage = nullptr;
  static thread_local int stage_dev = -1;
  if (stage == nullptr || stage_dev != comm->dev) {
    if (stage) { stage->release(); delete stage; }
    stage = new DevBuf();
    stage_dev = comm->dev;
  }
  B2VS_TRY(stage->reserve(qb_al + db_al + std::max<size_t>(ib, 16)));
  char* base = stage->as<char>();
  if (qb) {
    B2VS_CHECK(q_local_host && out_d_host && out_i_host, B2VS_EINVAL, "host pointer is NULL");
    B2VS_CUDA(cudaMemcpyAsync(base, q_local_host, qb, cudaMemcpyHostToDevice, st));
  }
  float* d_dev = reinterpret_cast<float*>(base + qb_al);
  int64_t* i_dev = reinterpret_cast<int64_t*>(base + qb_al + db_al);
  B2VS_TRY(b2vs_search_sharded(comm, index, base, q_dtype, nq_total, dim, k, params, d_dev, i_dev, stream));
  if (db) {
    B2VS_CUDA(cudaMemcpyAsync(out_d_host, d_dev, db, cudaMemcpyDeviceToHost, st));
    B2VS_CUDA(cudaMemcpyAsync(out_i_host, i_dev, ib, cudaMemcpyDeviceToHost, st));
  }
  B2VS_CUDA(cudaStreamSynchronize(st));
  return B2VS_OK;
}
