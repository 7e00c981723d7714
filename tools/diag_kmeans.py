"""GPU k-means (b2vs_kmeans_fit) vs the oracle's on structureless data: cluster-size spread for
several seeds and input dtypes; which balancing moves the GPU made (checked against the rule)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cuvs_rag_b200 as b2
from oracle.ivf import kmeans, assign, balance_pairs

n, d, ncl, iters = 98_304, 768, 1024, 8
g = torch.Generator().manual_seed(31)
x = torch.randn(2 * n, d, generator=g).to(torch.float16)[::2].contiguous()   # the IVF test's training rows
def stats(tag, cent, xs):
    lab = assign(xs.float(), cent)
    cnt = torch.bincount(lab, minlength=ncl).float()
    print(f"{tag:34s} std {cnt.std():6.1f} max {cnt.max():5.0f} min {cnt.min():3.0f} <=48 {(cnt <= 48).sum():4d} "
          f">=288 {(cnt >= 288).sum():4d} mean||c||^2 {(cent * cent).sum(1).mean():.2f}", flush=True)
for seed in (0, 1, 2):
    for dt in (torch.float16, torch.float32, torch.bfloat16):
        c, _ = b2.kmeans_fit(x.to(dt).cuda(), ncl, iters=iters, seed=seed)
        stats(f"gpu {str(dt)[6:]:8s} seed {seed}", c.cpu(), x)
    stats(f"oracle            seed {seed}", kmeans(x.float(), ncl, iters, seed), x)
# one balancing step in isolation: iters = 2 from the GPU's own seeds
for it in (1, 2, 3):
    c, _ = b2.kmeans_fit(x.cuda(), ncl, iters=it, seed=0)
    stats(f"gpu fp16 iters={it}", c.cpu(), x)
    stats(f"oracle   iters={it}", kmeans(x.float(), ncl, it, 0), x)
