"""C3 / C4-shard recall@10 and ms per 10K-query batch over (n_probes, refine_ratio) on bench.py's corpus:
picks the cheapest setting that meets recall@10 >= 0.95 (BASELINE metric)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cuvs_rag_b200 as b2
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "C4"
latent = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = torch.device("cuda:0")
if name == "C3":
    n, d, nl = 10_000_000, 768, 4096
    grid = [(16, 0), (32, 0), (48, 0), (64, 0), (96, 0)]
    if len(sys.argv) > 3:
        grid = [(int(v), 0) for v in sys.argv[3].split(",")]
else:
    n, d, nl = 12_500_000, 128, 16384
    grid = [(32, 4), (64, 2), (64, 4), (64, 8), (96, 4), (96, 8), (128, 4), (128, 8)]
x = bench.ivf_corpus(n, d, latent, torch.float16, dev, seed=5000)
q = bench.ivf_corpus(10_000, d, latent, torch.float16, dev, seed=99)
flat = b2.NativeIndex.flat(x)
_, truth = flat.search(q, 10)
flat.destroy()
ix = b2.NativeIndex.ivf_flat(x, nl, kmeans_iters=20) if name == "C3" else b2.NativeIndex.ivf_pq(x, nl, 64, kmeans_iters=20)
for npb, rr in grid:
    for _ in range(3):
        _, ids = ix.search(q, 10, n_probes=npb, refine_ratio=rr)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ix.search(q, 10, n_probes=npb, refine_ratio=rr)
    e1.record(); torch.cuda.synchronize()
    rec = float((ids.unsqueeze(2) == truth.unsqueeze(1)).any(2).float().mean())
    print(json.dumps({"config": name, "n_probes": npb, "refine_ratio": rr, "recall_at_10": round(rec, 4),
                      "ms_per_batch": round(e0.elapsed_time(e1) / 10, 3),
                      "mean_candidates": ix.last_stats().mean_candidates}), flush=True)
