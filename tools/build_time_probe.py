"""Wall-clock of the C3 build and of its k-means alone (b2vs_kmeans_fit) for a few iteration counts."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cuvs_rag_b200 as b2
import bench
dev = torch.device("cuda:0")
x = bench.ivf_corpus(10_000_000, 768, 16, torch.float16, dev, seed=5000)
train = x[::3].contiguous()
def t(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return time.perf_counter() - t0, r
for it in (1, 5, 20, 20):
    dt, _ = t(lambda: b2.kmeans_fit(train, 4096, iters=it, seed=0))
    print(f"kmeans_fit {train.shape[0]} x 768, 4096 clusters, iters={it}: {dt:.3f} s", flush=True)
for rep in range(2):
    dt, ix = t(lambda: b2.NativeIndex.ivf_flat(x, 4096, kmeans_iters=20))
    print(f"ivf_flat build: {dt:.3f} s", flush=True)
    ix.destroy()
