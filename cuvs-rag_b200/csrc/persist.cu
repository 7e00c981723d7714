// Index persistence: the reference rebuilds every index on every run (it only
// saves raw embeddings as .pt files, cuvs-2gpu-main.ipynb cells 10/12).
#include "ivf_internal.cuh"

// ------------------------------------------------------------------------------------------
// Index persistence: the reference rebuilds every index on every run (it only
// saves raw embeddings as .pt files, cuvs-2gpu-main.ipynb cells 10/12).  An IVF index is written
// as one little-endian file: header + raw device arrays; loading re-creates the coarse engine.
namespace b2vs {

struct IndexFileHeader {
  char magic[8];       // "B2VSIDX3" (3: header_bytes in `reserved`, every size re-derived on load)
  int32_t kind, metric, dtype, dim, n_lists, pq_dim, pq_bits, dsub, mp, dp, fmt, row_bytes;
  int64_t n, id_offset, n_slots;
  float max_norm2;
  int32_t reserved;
  uint64_t bytes_centroids, bytes_offsets, bytes_sizes, bytes_row_ids, bytes_data, bytes_slot_norm,
      bytes_codebooks, bytes_codes;
};

static int write_section(FILE* f, const DevBuf& b, uint64_t bytes) {
  if (bytes == 0) return B2VS_OK;
  std::vector<char> host(std::min<uint64_t>(bytes, 64ull << 20));
  for (uint64_t off = 0; off < bytes; off += host.size()) {
    const uint64_t len = std::min<uint64_t>(host.size(), bytes - off);
    B2VS_CUDA(cudaMemcpy(host.data(), static_cast<const char*>(b.ptr) + off, len, cudaMemcpyDeviceToHost));
    B2VS_CHECK(fwrite(host.data(), 1, len, f) == len, B2VS_EINVAL, "short write while saving index");
  }
  return B2VS_OK;
}

static int read_section(FILE* f, DevBuf* b, uint64_t bytes) {
  if (bytes == 0) return B2VS_OK;
  B2VS_TRY(b->reserve(bytes));
  std::vector<char> host(std::min<uint64_t>(bytes, 64ull << 20));
  for (uint64_t off = 0; off < bytes; off += host.size()) {
    const uint64_t len = std::min<uint64_t>(host.size(), bytes - off);
    B2VS_CHECK(fread(host.data(), 1, len, f) == len, B2VS_EINVAL, "index file is truncated");
    B2VS_CUDA(cudaMemcpy(static_cast<char*>(b->ptr) + off, host.data(), len, cudaMemcpyHostToDevice));
  }
  return B2VS_OK;
}

}  // namespace b2vs


namespace b2vs {

// Section sizes implied by the header scalars - the single definition used by save AND load, so a
// file whose bytes_* fields disagree with its own shape (truncated, corrupted, other build) is
// rejected before any kernel can index past an undersized buffer.
static void expected_sections(const IndexFileHeader& h, uint64_t out[8]) {
  const uint64_t slots = static_cast<uint64_t>(std::max<int64_t>(h.n_slots, 1));
  out[0] = static_cast<uint64_t>(h.n_lists) * h.dim * sizeof(float);       // centroids
  out[1] = (static_cast<uint64_t>(h.n_lists) + 1) * sizeof(uint32_t);      // offsets
  out[2] = static_cast<uint64_t>(h.n_lists) * sizeof(int);                 // sizes
  out[3] = slots * sizeof(uint32_t);                                       // row_ids
  out[4] = out[5] = out[6] = out[7] = 0;
  if (h.kind == B2VS_KIND_IVF_FLAT) {
    out[4] = slots * h.dp * 2;                                             // data
    out[5] = (slots + kNormSlack) * sizeof(float);                         // slot_norm
  } else {
    out[6] = static_cast<uint64_t>(h.pq_dim) * 256 * h.dsub * sizeof(float);                 // codebooks
    out[7] = static_cast<uint64_t>(std::max<int64_t>(h.n_slots, 32)) * h.mp;                 // codes
  }
}

// Same limits as ivf_build (ivf_build.cu) + internal consistency of the derived fields.
static int validate_header(const IndexFileHeader& h, const char* path) {
  B2VS_CHECK(h.reserved == static_cast<int32_t>(sizeof(IndexFileHeader)), B2VS_EINVAL,
             "%s: header size %d != %zu (file written by another build)", path, h.reserved,
             sizeof(IndexFileHeader));
  B2VS_CHECK(h.kind == B2VS_KIND_IVF_FLAT || h.kind == B2VS_KIND_IVF_PQ, B2VS_EINVAL,
             "%s: bad index kind %d", path, h.kind);
  B2VS_CHECK(h.metric == B2VS_METRIC_L2 || h.metric == B2VS_METRIC_IP || h.metric == B2VS_METRIC_COSINE, B2VS_EINVAL,
             "%s: bad metric %d", path, h.metric);
  B2VS_CHECK(h.dtype == B2VS_F32 || h.dtype == B2VS_F16 || h.dtype == B2VS_BF16, B2VS_EINVAL,
             "%s: bad dtype %d", path, h.dtype);
  B2VS_CHECK(h.dim >= 1 && h.dim <= 2048, B2VS_EINVAL, "%s: dim=%d outside [1, 2048]", path, h.dim);
  B2VS_CHECK(h.n >= 1 && h.n < (1ll << 31) - (1ll << 22), B2VS_EINVAL, "%s: n=%lld out of range", path,
             static_cast<long long>(h.n));
  B2VS_CHECK(h.n_lists >= 1 && h.n_lists <= h.n, B2VS_EINVAL, "%s: n_lists=%d outside [1, n]", path,
             h.n_lists);
  B2VS_CHECK(h.dp == static_cast<int32_t>(round_up(h.dim, 8)), B2VS_EINVAL, "%s: dp=%d inconsistent with dim=%d",
             path, h.dp, h.dim);
  B2VS_CHECK(h.fmt == ((h.dtype == B2VS_F16) ? 0 : 1), B2VS_EINVAL, "%s: fmt=%d inconsistent with dtype=%d",
             path, h.fmt, h.dtype);
  // every list is padded to 32 slots: n <= n_slots <= n + 32 * n_lists, and a multiple of 32
  B2VS_CHECK(h.n_slots >= h.n && h.n_slots <= h.n + 32ll * h.n_lists && h.n_slots % 32 == 0 &&
                 h.n_slots < (1ll << 32),
             B2VS_EINVAL, "%s: n_slots=%lld inconsistent with n=%lld, n_lists=%d", path,
             static_cast<long long>(h.n_slots), static_cast<long long>(h.n), h.n_lists);
  if (h.kind == B2VS_KIND_IVF_PQ) {
    B2VS_CHECK(h.pq_bits == 8, B2VS_EINVAL, "%s: pq_bits=%d (only 8)", path, h.pq_bits);
    B2VS_CHECK(h.pq_dim >= 1 && h.pq_dim <= 200 && h.dim % h.pq_dim == 0, B2VS_EINVAL,
               "%s: pq_dim=%d invalid for dim=%d", path, h.pq_dim, h.dim);
    B2VS_CHECK(h.dsub == h.dim / h.pq_dim && h.dsub <= 16, B2VS_EINVAL, "%s: dsub=%d inconsistent", path, h.dsub);
    B2VS_CHECK(h.mp == static_cast<int32_t>(round_up(h.pq_dim, 16)), B2VS_EINVAL, "%s: mp=%d inconsistent", path, h.mp);
    B2VS_CHECK(h.row_bytes == h.pq_dim, B2VS_EINVAL, "%s: row_bytes=%d inconsistent", path, h.row_bytes);
  } else {
    B2VS_CHECK(h.pq_dim == 0 && h.pq_bits == 0 && h.dsub == 0 && h.mp == 0, B2VS_EINVAL,
               "%s: PQ fields set on an IVF-Flat index", path);
    B2VS_CHECK(h.row_bytes == h.dp * 2, B2VS_EINVAL, "%s: row_bytes=%d inconsistent", path, h.row_bytes);
  }
  uint64_t want[8];
  expected_sections(h, want);
  const uint64_t got[8] = {h.bytes_centroids, h.bytes_offsets, h.bytes_sizes, h.bytes_row_ids,
                           h.bytes_data, h.bytes_slot_norm, h.bytes_codebooks, h.bytes_codes};
  static const char* names[8] = {"centroids", "offsets", "sizes", "row_ids", "data", "slot_norm",
                                 "codebooks", "codes"};
  for (int i = 0; i < 8; ++i)
    B2VS_CHECK(got[i] == want[i], B2VS_EINVAL, "%s: section %s is %llu bytes, the header's shape implies %llu",
               path, names[i], static_cast<unsigned long long>(got[i]), static_cast<unsigned long long>(want[i]));
  return B2VS_OK;
}

// offsets must be the exclusive scan of the 32-padded sizes and end at n_slots; sizes sum to n
static int validate_lists(const IndexFileHeader& h, const std::vector<int32_t>& sizes,
                          const std::vector<uint32_t>& offsets, const char* path) {
  uint64_t run = 0, total = 0;
  for (int i = 0; i < h.n_lists; ++i) {
    B2VS_CHECK(sizes[i] >= 0 && offsets[i] == run, B2VS_EINVAL, "%s: list table corrupt at list %d", path, i);
    total += static_cast<uint64_t>(sizes[i]);
    run += static_cast<uint64_t>(round_up(sizes[i], 32));
  }
  B2VS_CHECK(offsets[h.n_lists] == run && run == static_cast<uint64_t>(h.n_slots) &&
                 total == static_cast<uint64_t>(h.n),
             B2VS_EINVAL, "%s: list table does not add up (slots %llu / %lld, rows %llu / %lld)", path,
             static_cast<unsigned long long>(run), static_cast<long long>(h.n_slots),
             static_cast<unsigned long long>(total), static_cast<long long>(h.n));
  return B2VS_OK;
}

}  // namespace b2vs

using namespace b2vs;

extern "C" int b2vs_index_save(const b2vs_index* index, const char* path) {
  B2VS_CHECK(index && path, B2VS_EINVAL, "NULL argument");
  B2VS_CHECK(index->kind != B2VS_KIND_FLAT, B2VS_EUNSUP,
             "flat indexes borrow their rows and hold no trained state: re-create them instead");
  const IvfData* d = static_cast<const IvfData*>(index->ivf);
  B2VS_CHECK(d != nullptr, B2VS_EINVAL, "index has no list data");
  DeviceGuard guard(index->dev);
  B2VS_CHECK(guard.ok, B2VS_ECUDA, "cannot select device %d", index->dev);
  B2VS_CUDA(cudaDeviceSynchronize());
  IndexFileHeader h{};
  std::memcpy(h.magic, "B2VSIDX3", 8);
  h.kind = index->kind; h.metric = index->cosine ? B2VS_METRIC_COSINE : index->metric;
  h.dtype = index->dtype; h.dim = index->dim;
  h.n_lists = d->n_lists; h.pq_dim = d->pq_dim; h.pq_bits = d->pq_bits; h.dsub = d->dsub;
  h.mp = d->pq_dim ? d->mp : 0;
  h.dp = d->dp; h.fmt = d->fmt; h.row_bytes = d->row_bytes;
  h.n = index->n; h.id_offset = index->id_offset; h.n_slots = d->n_slots;
  h.max_norm2 = d->max_norm2;
  h.reserved = static_cast<int32_t>(sizeof(IndexFileHeader));
  uint64_t sec[8];
  expected_sections(h, sec);
  h.bytes_centroids = sec[0]; h.bytes_offsets = sec[1]; h.bytes_sizes = sec[2]; h.bytes_row_ids = sec[3];
  h.bytes_data = sec[4]; h.bytes_slot_norm = sec[5]; h.bytes_codebooks = sec[6]; h.bytes_codes = sec[7];
  FILE* f = fopen(path, "wb");
  B2VS_CHECK(f != nullptr, B2VS_EINVAL, "cannot open %s for writing", path);
  int rc = fwrite(&h, sizeof(h), 1, f) == 1 ? B2VS_OK : B2VS_EINVAL;
  if (rc == B2VS_OK) rc = write_section(f, d->centroids, h.bytes_centroids);
  if (rc == B2VS_OK) rc = write_section(f, d->offsets, h.bytes_offsets);
  if (rc == B2VS_OK) rc = write_section(f, d->sizes, h.bytes_sizes);
  if (rc == B2VS_OK) rc = write_section(f, d->row_ids, h.bytes_row_ids);
  if (rc == B2VS_OK) rc = write_section(f, d->data, h.bytes_data);
  if (rc == B2VS_OK) rc = write_section(f, d->slot_norm, h.bytes_slot_norm);
  if (rc == B2VS_OK) rc = write_section(f, d->codebooks, h.bytes_codebooks);
  if (rc == B2VS_OK) rc = write_section(f, d->codes, h.bytes_codes);
  fclose(f);
  if (rc != B2VS_OK && b2vs_last_error()[0] == 0) set_error("write to %s failed", path);
  return rc;
}

extern "C" int b2vs_index_load(int dev, const char* path, const void* rows_for_refine, int64_t id_offset,
                               void* stream, b2vs_index** out) {
  B2VS_CHECK(path && out, B2VS_EINVAL, "NULL argument");
  *out = nullptr;
  int count = 0;
  B2VS_CUDA(cudaGetDeviceCount(&count));
  B2VS_CHECK(dev >= 0 && dev < count, B2VS_EINVAL, "device %d not in [0, %d)", dev, count);
  FILE* f = fopen(path, "rb");
  B2VS_CHECK(f != nullptr, B2VS_EINVAL, "cannot open %s", path);
  IndexFileHeader h{};
  if (fread(&h, sizeof(h), 1, f) != 1 || std::memcmp(h.magic, "B2VSIDX3", 8) != 0) {
    fclose(f);
    set_error("%s is not a b2vs index file (version 3)", path);
    return B2VS_EINVAL;
  }
  int rc = validate_header(h, path);
  if (rc == B2VS_OK) {
    // the file must hold exactly the sections the header promises
    uint64_t total = sizeof(h) + h.bytes_centroids + h.bytes_offsets + h.bytes_sizes + h.bytes_row_ids +
                     h.bytes_data + h.bytes_slot_norm + h.bytes_codebooks + h.bytes_codes;
    if (fseek(f, 0, SEEK_END) != 0 || static_cast<uint64_t>(ftell(f)) != total ||
        fseek(f, static_cast<long>(sizeof(h)), SEEK_SET) != 0) {
      set_error("%s: file size does not match its header (truncated or trailing bytes)", path);
      rc = B2VS_EINVAL;
    }
  }
  if (rc != B2VS_OK) { fclose(f); return rc; }
  DeviceGuard guard(dev);
  if (!guard.ok) { fclose(f); set_error("cannot select device %d", dev); return B2VS_ECUDA; }
  b2vs_index* ix = new (std::nothrow) b2vs_index();
  IvfData* d = new (std::nothrow) IvfData();
  if (!ix || !d) {
    fclose(f);
    delete ix;
    delete d;
    set_error("host allocation failed");
    return B2VS_ENOMEM;
  }
  ix->kind = h.kind; ix->dev = dev; ix->dtype = h.dtype; ix->dim = h.dim;
  ix->cosine = h.metric == B2VS_METRIC_COSINE;
  ix->metric = ix->cosine ? B2VS_METRIC_IP : h.metric;   // a cosine index IS the IP engine on unit rows
  ix->n = h.n; ix->id_offset = id_offset >= 0 ? id_offset : h.id_offset; ix->ivf = d;
  d->n_lists = h.n_lists; d->pq_dim = h.pq_dim; d->pq_bits = h.pq_bits; d->dsub = h.dsub;
  d->mp = static_cast<int>(round_up(h.pq_dim, 16));
  d->dp = h.dp; d->fmt = h.fmt; d->row_bytes = h.row_bytes; d->n = h.n; d->n_slots = h.n_slots;
  d->max_norm2 = h.max_norm2;
  d->src_rows = rows_for_refine;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  rc = B2VS_OK;
  if (ix->cosine && rows_for_refine && h.kind == B2VS_KIND_IVF_PQ) {
    // refine compares unit vectors: keep an owned unit-norm copy of the caller's rows
    rc = cosine_rows(h.dtype, h.dim, rows_for_refine, h.n, st, &ix->cos_rows);
    d->src_rows = ix->cos_rows.ptr;
  }
  if (rc == B2VS_OK) rc = read_section(f, &d->centroids, h.bytes_centroids);
  if (rc == B2VS_OK) rc = read_section(f, &d->offsets, h.bytes_offsets);
  if (rc == B2VS_OK) rc = read_section(f, &d->sizes, h.bytes_sizes);
  if (rc == B2VS_OK) rc = read_section(f, &d->row_ids, h.bytes_row_ids);
  if (rc == B2VS_OK) rc = read_section(f, &d->data, h.bytes_data);
  if (rc == B2VS_OK) rc = read_section(f, &d->slot_norm, h.bytes_slot_norm);
  if (rc == B2VS_OK) rc = read_section(f, &d->codebooks, h.bytes_codebooks);
  if (rc == B2VS_OK) rc = read_section(f, &d->codes, h.bytes_codes);
  fclose(f);
  if (rc == B2VS_OK) {
    d->h_sizes.resize(h.n_lists);
    std::vector<uint32_t> h_off(static_cast<size_t>(h.n_lists) + 1);
    if (cudaMemcpy(d->h_sizes.data(), d->sizes.ptr, h.bytes_sizes, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(h_off.data(), d->offsets.ptr, h.bytes_offsets, cudaMemcpyDeviceToHost) != cudaSuccess) {
      set_error("%s: reading the list table back failed", path);
      rc = B2VS_ECUDA;
    }
    if (rc == B2VS_OK) rc = validate_lists(h, d->h_sizes, h_off, path);
    if (rc == B2VS_OK) rc = build_list_ranks(d);
  }
  if (rc == B2VS_OK) {
    const int force = (h.dtype == B2VS_F32) ? -1 : h.fmt;
    rc = ix->flat.init(dev, ix->metric, B2VS_F32, h.dim, d->centroids.ptr, h.n_lists, st, force);
  }
  if (rc == B2VS_OK && h.kind == B2VS_KIND_IVF_PQ) rc = pq_prepare_grouped(ix, d, st);
  if (rc == B2VS_OK && cudaStreamSynchronize(st) != cudaSuccess) {
    set_error("%s: device error while loading", path);
    rc = B2VS_ECUDA;
  }
  if (rc != B2VS_OK) {
    ivf_destroy(ix);
    ix->flat.destroy();
    ix->cos_rows.release();
    delete ix;
    return rc;
  }
  *out = ix;
  return B2VS_OK;
}
