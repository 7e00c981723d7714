// Small IVF batches replayed as one CUDA graph, and the ivf_search entry point (api.cu ->
// b2vs_search).  The interactive single-query case of the reference
// (improved_multi_gpu_rag.py:239-277: Q = 1, k' = 2k) is a chain of ~15 short kernels.
#include "ivf_internal.cuh"

namespace b2vs {

// ------------------------------------------------------------------------------------------
// Small batches as one CUDA graph.  B2VS_GRAPH=1 enables it for every eligible call, =0 disables
// it even when a call asks with B2VS_FLAG_GRAPH; unset = only calls that set the flag.
static int graph_override() { return env().graph; }

static void drop_graph_error() { cudaGetLastError(); }

// Replays (capturing first if needed) the search of this signature.  *handled = false means the
// caller must run the direct path (first sighting of the signature, or capture was refused).
static int ivf_search_graphed(b2vs_index* index, IvfData* d, const void* q, int q_dtype, int nq, int k,
                              const b2vs_search_params& sp, float* out_d, int64_t* out_i,
                              cudaStream_t st, bool* handled) {
  *handled = false;
  const int flags = sp.flags & ~B2VS_FLAG_GRAPH;
  SearchGraph* g = nullptr;
  for (SearchGraph& e : d->graphs)
    if (e.nq == nq && e.k == k && e.q_dtype == q_dtype && e.n_probes == sp.n_probes &&
        e.refine_ratio == sp.refine_ratio && e.flags == flags) { g = &e; break; }
  if (!g) {
    if (static_cast<int>(d->graphs.size()) >= kGraphMaxEntries) {
      size_t victim = 0;
      for (size_t i = 1; i < d->graphs.size(); ++i)
        if (d->graphs[i].last_use < d->graphs[victim].last_use) victim = i;
      // the victim's graph may still be running on the caller's stream
      if (d->graphs[victim].exec) cudaStreamSynchronize(st);
      d->graphs[victim].destroy();
      d->graphs.erase(d->graphs.begin() + static_cast<long>(victim));
    }
    d->graphs.emplace_back();
    g = &d->graphs.back();
    g->nq = nq; g->k = k; g->q_dtype = q_dtype; g->n_probes = sp.n_probes;
    g->refine_ratio = sp.refine_ratio; g->flags = flags;
  }
  g->last_use = ++d->graph_clock;
  if (g->failed) return B2VS_OK;
  if (g->exec && g->generation != realloc_generation()) {
    // some workspace moved since the capture: the graph's pointers may be stale
    cudaStreamSynchronize(st);
    cudaGraphExecDestroy(g->exec);
    g->exec = nullptr;
  }
  if (!g->exec) {
    // the first call of a signature runs directly and sizes every workspace, so that the
    // capture below allocates nothing
    if (g->seen++ == 0) return B2VS_OK;
    if (!d->cap_stream)
      B2VS_CUDA(cudaStreamCreateWithFlags(&d->cap_stream, cudaStreamNonBlocking));
    if (!g->io) {
      g->q_bytes = static_cast<size_t>(nq) * index->dim * elem_bytes(q_dtype);
      g->d_off = static_cast<size_t>(round_up(static_cast<int64_t>(g->q_bytes), 256));
      g->i_off = g->d_off + static_cast<size_t>(round_up(static_cast<int64_t>(nq) * k * sizeof(float), 256));
      void* p = nullptr;
      if (cudaMalloc(&p, g->i_off + static_cast<size_t>(nq) * k * sizeof(int64_t)) != cudaSuccess) {
        drop_graph_error();
        g->failed = true;
        return B2VS_OK;
      }
      g->io = static_cast<char*>(p);
    }
    b2vs_search_params spd = sp;
    spd.flags = flags;
    const uint64_t gen0 = realloc_generation();
    if (cudaStreamBeginCapture(d->cap_stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
      drop_graph_error();
      g->failed = true;
      return B2VS_OK;
    }
    const int rc = ivf_search_direct(index, g->io, q_dtype, nq, k, spd,
                                     reinterpret_cast<float*>(g->io + g->d_off),
                                     reinterpret_cast<int64_t*>(g->io + g->i_off), d->cap_stream);
    cudaGraph_t graph = nullptr;
    const cudaError_t e_end = cudaStreamEndCapture(d->cap_stream, &graph);
    size_t n_nodes = 0;
    bool ok = rc == B2VS_OK && e_end == cudaSuccess && graph != nullptr &&
              gen0 == realloc_generation() &&
              cudaGraphGetNodes(graph, nullptr, &n_nodes) == cudaSuccess &&
              cudaGraphInstantiate(&g->exec, graph, 0ull) == cudaSuccess;
    if (graph) cudaGraphDestroy(graph);
    if (!ok) {
      drop_graph_error();
      if (g->exec) cudaGraphExecDestroy(g->exec);
      g->exec = nullptr;
      // a workspace grew under the capture (another signature's sizes): try again next call;
      // anything else is a refusal
      if (gen0 == realloc_generation()) g->failed = true;
      return B2VS_OK;
    }
    g->generation = gen0;
    g->stats = d->stats;
    g->stats.launches = static_cast<int32_t>(n_nodes);
  }
  B2VS_CUDA(cudaMemcpyAsync(g->io, q, g->q_bytes, cudaMemcpyDeviceToDevice, st));
  B2VS_CUDA(cudaGraphLaunch(g->exec, st));
  B2VS_CUDA(cudaMemcpyAsync(out_d, g->io + g->d_off, static_cast<size_t>(nq) * k * sizeof(float),
                            cudaMemcpyDeviceToDevice, st));
  B2VS_CUDA(cudaMemcpyAsync(out_i, g->io + g->i_off, static_cast<size_t>(nq) * k * sizeof(int64_t),
                            cudaMemcpyDeviceToDevice, st));
  d->stats = g->stats;
  d->counter_pending = true;
  d->last_nq = nq;
  d->timing_pending = false;
  *handled = true;
  return B2VS_OK;
}

int ivf_search(b2vs_index* index, const void* q, int q_dtype, int nq, int k,
               const b2vs_search_params& sp, float* out_d, int64_t* out_i, cudaStream_t st) {
  IvfData* d = static_cast<IvfData*>(index->ivf);
  B2VS_CHECK(d != nullptr, B2VS_EINVAL, "IVF index has no list data");
  const int ov = graph_override();
  // Graph replay: small batches (nq <= 64) when the call asks with B2VS_FLAG_GRAPH, and large
  // batches (nq >= 2048, fixed shapes in serving) by default - the ~27 launches of a 10K-query
  // batch leave ~2 us gaps that add up to 0.03-0.06 ms of a 1.7-3.7 ms step (measured at C3 / C4).
  // B2VS_GRAPH=1 / =0 force it for every eligible call / never; B2VS_GRAPH_MAXQ moves the upper bound.
  const bool small = nq <= kGraphMaxQueries, large = nq >= kGraphAutoMinQueries;
  const bool want_graph = ov == 1 || (ov < 0 && ((small && (sp.flags & B2VS_FLAG_GRAPH) != 0) || large));
  const int graph_maxq = env().graph_maxq > 0 ? env().graph_maxq : kGraphAutoMaxQueries;
  if (want_graph && (small || large || ov == 1) && nq <= graph_maxq && k >= 1 &&
      (sp.flags & B2VS_FLAG_TIME_KERNEL) == 0 &&
      !uses_bigk_path(index, d, k, sp)) {
    bool handled = false;
    B2VS_TRY(ivf_search_graphed(index, d, q, q_dtype, nq, k, sp, out_d, out_i, st, &handled));
    if (handled) return B2VS_OK;
  }
  b2vs_search_params spd = sp;
  spd.flags &= ~B2VS_FLAG_GRAPH;
  return ivf_search_direct(index, q, q_dtype, nq, k, spd, out_d, out_i, st);
}


}  // namespace b2vs
