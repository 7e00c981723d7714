"""One Q = 1 IVF search under the profiler (per-kernel durations of the interactive case):
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/q1_flat_launches.csv python tools/q1_probe.py flat
"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuvs_rag_b200 as b2
kind = sys.argv[1] if len(sys.argv) > 1 else "flat"
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(7)
n, d, nlist, nprobe = 10_000_000, 768, 4096, 32
if kind == "pq":
    n, d, nlist, nprobe = 12_500_000, 128, 16384, 64
cent = torch.randn(nlist, d, generator=g, device=dev)
x = torch.empty((n, d), dtype=torch.float16, device=dev)
for s in range(0, n, 1 << 19):
    e = min(n, s + (1 << 19))
    lab = torch.randint(0, nlist, (e - s,), generator=g, device=dev)
    x[s:e] = (cent[lab] + 0.42 * torch.randn((e - s, d), generator=g, device=dev)).to(torch.float16)
ix = (b2.NativeIndex.ivf_flat(x, nlist, kmeans_iters=4) if kind == "flat"
      else b2.NativeIndex.ivf_pq(x, nlist, 64, kmeans_iters=4))
rr = 4 if kind == "pq" else 0
q = x[12345:12346].clone()
for _ in range(3):
    ix.search(q, 20, n_probes=nprobe, refine_ratio=rr)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ix.search(q, 20, n_probes=nprobe, refine_ratio=rr)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
