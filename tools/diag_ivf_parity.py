"""Where does an IVF recall gap GPU-vs-oracle come from: the trained centroids or the search?
Builds the GPU index and the oracle on structureless data, then a second oracle whose centroids are
the GPU's, and prints recall@10 per probe count plus list-size statistics and k-means inertia."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cuvs_rag_b200 as b2
from oracle.exact import exact_knn
from oracle.ivf import IvfFlatOracle, assign, recall, kmeans

n, d, nlist, k, nq = 196_608, 768, 1024, 10, 400
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 8
g = torch.Generator().manual_seed(31)
x = torch.randn(n, d, generator=g).to(torch.float16)
g = torch.Generator().manual_seed(32)
q = torch.randn(nq, d, generator=g).to(torch.float16)
ix = b2.NativeIndex.ivf_flat(x.cuda(), nlist, kmeans_iters=iters)
_, truth = exact_knn(x.float(), q.float(), k)
t0 = time.time()
o1 = IvfFlatOracle(x.float(), nlist, iters=iters)
print("oracle build s", time.time() - t0)
o2 = IvfFlatOracle.__new__(IvfFlatOracle)
o2.metric, o2.db = "sqeuclidean", x.float()
o2.cent = ix.centroids().clone()
o2.labels = assign(o2.db, o2.cent)
o2.order = torch.argsort(o2.labels, stable=True)
cnt = torch.bincount(o2.labels, minlength=nlist)
o2.offsets = torch.zeros(nlist + 1, dtype=torch.int64); o2.offsets[1:] = torch.cumsum(cnt, 0)
# oracle without balancing, and GPU kmeans_fit alone
o3 = IvfFlatOracle(x.float(), nlist, iters=iters, balance=False)
probes = [1, 8, 32, 128]
r1, r2, r3 = o1.search_many(q.float(), k, probes), o2.search_many(q.float(), k, probes), o3.search_many(q.float(), k, probes)
for p in probes:
    _, gi = ix.search(q.cuda(), k, n_probes=p)
    print(f"nprobe {p:4d}: gpu {recall(gi.cpu(), truth):.4f} | oracle(own kmeans) {recall(r1[p], truth):.4f} | "
          f"oracle(GPU centroids) {recall(r2[p], truth):.4f} | oracle(no balancing) {recall(r3[p], truth):.4f}")
def stats(name, sizes, cent, labels):
    sizes = sizes.float()
    inertia = ((x.float() - cent[labels]) ** 2).sum(1).mean().item()
    print(f"{name}: sizes mean {sizes.mean():.1f} std {sizes.std():.1f} min {sizes.min():.0f} max {sizes.max():.0f} "
          f"empty {(sizes == 0).sum().item()} | inertia {inertia:.3f} | mean ||c||^2 {(cent * cent).sum(1).mean():.4f}")
stats("gpu   ", ix.list_sizes(), o2.cent, o2.labels)
stats("oracle", torch.bincount(o1.labels, minlength=nlist), o1.cent, o1.labels)
stats("nobal ", torch.bincount(o3.labels, minlength=nlist), o3.cent, o3.labels)
print("gpu labels == assign(GPU centroids):", (ix.list_sizes().long() == cnt).float().mean().item())
