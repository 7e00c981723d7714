import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuvs_rag_b200 as b2
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(7)
n, d, nlist = 10_000_000, 768, 4096
cent = torch.randn(nlist, d, generator=g, device=dev)
x = torch.empty((n, d), dtype=torch.float16, device=dev)
for s in range(0, n, 1 << 19):
    e = min(n, s + (1 << 19))
    lab = torch.randint(0, nlist, (e - s,), generator=g, device=dev)
    x[s:e] = (cent[lab] + 0.42 * torch.randn((e - s, d), generator=g, device=dev)).to(torch.float16)
ix = b2.NativeIndex.ivf_flat(x, nlist, kmeans_iters=4)
q = x[12345:12346].clone()
for _ in range(3):
    ix.search(q, 20, n_probes=32)
torch.cuda.synchronize()
