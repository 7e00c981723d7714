"""Q = 1..8 IVF latency against the row-range chunk of the grouped scan's work items
(B2VS_WORK_CHUNK_TILES; 0 = the heuristic).  usage: sweep_work_split.py flat|pq [Q,Q,...] [chunk,chunk,...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cuvs_rag_b200 as b2

kind = sys.argv[1] if len(sys.argv) > 1 else "flat"
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
rr = 4 if kind == "pq" else 0
sizes = ix.list_sizes()
print(json.dumps({"kind": kind, "mean_list": float(sizes.float().mean()), "max_list": int(sizes.max())}), flush=True)
q_list = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 4, 8]
c_list = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1, 2, 3, 4, 5, 6, 8, 10, 16]
for nq in q_list:
    qi = torch.randint(0, n, (nq,), generator=g, device=dev)
    q = (x[qi].float() + 0.1 * torch.randn((nq, d), generator=g, device=dev)).to(torch.float16)
    os.environ["B2VS_DEBUG_SPLIT"] = "1"
    b2._native.reload_env()
    os.environ.pop("B2VS_WORK_CHUNK_TILES", None)
    b2._native.reload_env()
    ix.search(q, 20, n_probes=nprobe, refine_ratio=rr)
    os.environ.pop("B2VS_DEBUG_SPLIT")
    b2._native.reload_env()
    for c in c_list:
        if c:
            os.environ["B2VS_WORK_CHUNK_TILES"] = str(c)
            b2._native.reload_env()
        else:
            os.environ.pop("B2VS_WORK_CHUNK_TILES", None)
            b2._native.reload_env()
        for _ in range(5):
            ix.search(q, 20, n_probes=nprobe, refine_ratio=rr)
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                ix.search(q, 20, n_probes=nprobe, refine_ratio=rr)
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 50)
        print(json.dumps({"kind": kind, "Q": nq, "chunk_tiles": c, "ms": round(best, 4)}), flush=True)
