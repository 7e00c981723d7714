"""A/B of B2VS_* switches on one C3 / C4-shard index (bench.py's corpus): ms per 10K-query batch,
recall@10 against the exact index, mean candidates per query - one JSON line per setting.
usage: ivf_knob_sweep.py C3|C4 "A=1,B=2;A=3;..."   (an empty setting = defaults)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cuvs_rag_b200 as b2
from cuvs_rag_b200 import _native
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
settings = sys.argv[2].split(";") if len(sys.argv) > 2 else [""]
npb_override = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rr_override = int(sys.argv[4]) if len(sys.argv) > 4 else -1
dev = torch.device("cuda:0")
if name == "C3":
    n, d, nl, npb, rr = 10_000_000, 768, 4096, 32, 0
else:
    n, d, nl, npb, rr = 12_500_000, 128, 16384, 64, 4
if npb_override:
    npb = npb_override
if rr_override >= 0:
    rr = rr_override
x = bench.ivf_corpus(n, d, 16, torch.float16, dev, seed=5000)
q = bench.ivf_corpus(10_000, d, 16, torch.float16, dev, seed=99)
flat = b2.NativeIndex.flat(x)
_, gt = flat.search(q, 10)
torch.cuda.synchronize()
del flat
ix = b2.NativeIndex.ivf_flat(x, nl, kmeans_iters=20) if name == "C3" else b2.NativeIndex.ivf_pq(x, nl, 64, kmeans_iters=20)
touched = set()
for s in settings:
    for kname in touched:
        os.environ.pop(kname, None)
    touched = set()
    for kv in filter(None, s.split(",")):
        kname, v = kv.split("=")
        os.environ[kname] = v
        touched.add(kname)
    _native.reload_env()
    for _ in range(3):
        dd, ii = ix.search(q, 10, n_probes=npb, refine_ratio=rr)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ix.search(q, 10, n_probes=npb, refine_ratio=rr)
    e1.record(); torch.cuda.synchronize()
    ix.search(q, 10, n_probes=npb, refine_ratio=rr, time_kernel=True)
    torch.cuda.synchronize()
    kernel_ms = ix.last_stats().kernel_ms
    hit = (ii.unsqueeze(2) == gt.unsqueeze(1)).any(2).float().mean().item()
    print(json.dumps({"config": name, "n_probes": npb, "refine": rr, "setting": s,
                      "ms_per_batch": round(e0.elapsed_time(e1) / 20, 4), "scan_kernel_ms": round(kernel_ms, 4), "recall@10": round(hit, 4),
                      "mean_candidates": round(ix.last_stats().mean_candidates, 1)}), flush=True)
