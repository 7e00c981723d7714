"""Launch list of one IVF build (ncu --profile-from-start off): where the build time goes.
usage: build_phase_probe.py C3|C4 [latent]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cuvs_rag_b200 as b2
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
latent = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = torch.device("cuda:0")
n, d, nl = (10_000_000, 768, 4096) if name == "C3" else (12_500_000, 128, 16384)
x = bench.ivf_corpus(n, d, latent, torch.float16, dev, seed=5000)
w = b2.NativeIndex.ivf_flat(x[:200000], 64, kmeans_iters=2); w.destroy()
torch.cuda.synchronize(); t0 = time.perf_counter()
torch.cuda.profiler.start()
ix = b2.NativeIndex.ivf_flat(x, nl, kmeans_iters=20) if name == "C3" else b2.NativeIndex.ivf_pq(x, nl, 64, kmeans_iters=20)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(name, "build s", time.perf_counter() - t0, flush=True)
