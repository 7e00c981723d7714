"""One C3 / C4-shard search batch under `ncu --profile-from-start off --metrics gpu__time_duration.sum`:
the launch list of a single 10K-query IVF search on bench.py's corpus (profiles/r2_*_launches.csv).
usage: ivf_phase_probe.py C3|C4 [latent_dim] [scale]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cuvs_rag_b200 as b2
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
latent = int(sys.argv[2]) if len(sys.argv) > 2 else 16
scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
npb_override = int(sys.argv[4]) if len(sys.argv) > 4 else 0
dev = torch.device("cuda:0")
if name == "C3":
    n, d, nl, npb, rr = int(10_000_000 * scale), 768, 4096, 32, 0
else:
    n, d, nl, npb, rr = int(12_500_000 * scale), 128, 16384, 64, 4
if npb_override:
    npb = npb_override
x = bench.ivf_corpus(n, d, latent, torch.float16, dev, seed=5000)
q = bench.ivf_corpus(10_000, d, latent, torch.float16, dev, seed=99)
ix = b2.NativeIndex.ivf_flat(x, nl, kmeans_iters=20) if name == "C3" else b2.NativeIndex.ivf_pq(x, nl, 64, kmeans_iters=20)
for _ in range(3):
    ix.search(q, 10, n_probes=npb, refine_ratio=rr)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ix.search(q, 10, n_probes=npb, refine_ratio=rr)
e1.record(); torch.cuda.synchronize()
print(name, "ms per batch", e0.elapsed_time(e1) / 10, "mean candidates", ix.last_stats().mean_candidates, flush=True)
torch.cuda.profiler.start()
ix.search(q, 10, n_probes=npb, refine_ratio=rr)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
