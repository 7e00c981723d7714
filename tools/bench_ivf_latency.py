"""Small-batch latency of the IVF paths (the reference's interactive RAG case: Q=1, k'=2k)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cuvs_rag_b200 as b2

kind = sys.argv[1] if len(sys.argv) > 1 else "flat"
use_graph = len(sys.argv) > 2 and sys.argv[2] == "graph"   # also time B2VS_FLAG_GRAPH on Q <= 64
n, d, nlist, nprobe = 10_000_000, 768, 4096, 32
if kind == "pq":
    n, d, nlist, nprobe = 12_500_000, 128, 16384, 64
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(7)
cent = torch.randn(nlist, d, generator=g, device=dev)
x = torch.empty((n, d), dtype=torch.float16, device=dev)
for s in range(0, n, 1 << 19):
    e = min(n, s + (1 << 19))
    lab = torch.randint(0, nlist, (e - s,), generator=g, device=dev)
    x[s:e] = (cent[lab] + 0.42 * torch.randn((e - s, d), generator=g, device=dev)).to(torch.float16)
ix = (b2.NativeIndex.ivf_flat(x, nlist, kmeans_iters=10) if kind == "flat"
      else b2.NativeIndex.ivf_pq(x, nlist, 64, kmeans_iters=10))
sizes = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024]
if len(sys.argv) > 3:
    sizes = [int(v) for v in sys.argv[3].split(",")]
rr = 4 if kind == "pq" else 0
for nq in sizes:
    qi = torch.randint(0, n, (nq,), generator=g, device=dev)
    q = (x[qi].float() + 0.1 * torch.randn((nq, d), generator=g, device=dev)).to(torch.float16)
    for graph in ([False, True] if use_graph and nq <= 64 else [False]):
        for _ in range(3):
            ix.search(q, 20, n_probes=nprobe, refine_ratio=rr, graph=graph)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            ix.search(q, 20, n_probes=nprobe, refine_ratio=rr, graph=graph)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 50
        print(json.dumps({"kind": kind, "graph": graph, "Q": nq, "ms": round(ms, 4),
                          "qps": round(nq / ms * 1e3), "nodes": ix.last_stats().launches}), flush=True)
