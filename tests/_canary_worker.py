"""Worker of tests/test_gpu_canary.py: started with B2VS_CANARY=1, drives every search path of the
library once and then has the guard zones of all live device buffers checked."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cuvs_rag_b200 as b2  # noqa: E402

n = b2._native
g = torch.Generator().manual_seed(3)
keep = []


def check(tag):
    torch.cuda.synchronize()
    bufs, bad = n.check_canaries()
    print(f"CANARY {tag}: {bufs} buffers, {bad} corrupt", flush=True)
    assert bad == 0, (tag, n.lib().b2vs_last_error())
    return bufs


# exact search: fused k, k = 1 labels (k-means), large k, fp32 split, ragged sizes, both kernel variants
for (rows, d, nq, k, dt) in ((70_001, 128, 257, 10, torch.bfloat16), (300_000, 96, 33, 100, torch.float16),
                            (5_000, 40, 7, 5, torch.float32), (90_000, 64, 300, 500, torch.bfloat16)):
    x = torch.randn(rows, d, generator=g).to(dt).cuda()
    q = torch.randn(nq, d, generator=g).to(dt).cuda()
    ix = b2.NativeIndex.flat(x, id_offset=7)
    ix.search(q, k)
    keep.append((ix, x))
check("flat")
x = torch.randn(80_000, 128, generator=g).to(torch.float16).cuda()
for nq in (1, 33, 300, 700):          # one-CTA planner, legacy seed, tensor-core seed, counting-sort planner
    q = torch.randn(nq, 128, generator=g).to(torch.float16).cuda()
    for build, kw in ((lambda: b2.NativeIndex.ivf_flat(x, 128, kmeans_iters=4), {}),
                      (lambda: b2.NativeIndex.ivf_pq(x, 128, 64, kmeans_iters=4), {"refine_ratio": 4})):
        ix = build()
        ix.search(q, 10, n_probes=16, **kw)
        ix.search(q, 10, n_probes=128, **kw)
        if nq >= 33:
            ix.search(q, 300, n_probes=32)          # large-k two-pass paths
        keep.append((ix, x))
check("ivf")
for dsub_dim, m in ((256, 64), (256, 32)):          # dsub 4 and 8 decoders
    xx = torch.randn(40_000, dsub_dim, generator=g).to(torch.float16).cuda()
    ix = b2.NativeIndex.ivf_pq(xx, 64, m, kmeans_iters=3)
    ix.search(xx[:300].contiguous(), 10, n_probes=8, refine_ratio=2)
    keep.append((ix, xx))
check("pq dsub 4/8")
comm = n.Comm.init_rank("cuda:0", 1, 0, n.Comm.unique_id())
comm.search_sharded(keep[0][0], torch.randn(64, 128, generator=g).to(torch.bfloat16).cuda(), 64, 10)
comm.search_sharded(keep[0][0], torch.randn(64, 128, generator=g).to(torch.bfloat16), 64, 10)
ci = b2.NativeIndex.flat(torch.randn(3000, 32, generator=g).cuda(), metric="cosine")
ci.search(torch.randn(9, 32, generator=g).cuda(), 4)
b2.merge_topk(torch.rand(3, 50, 20).cuda(), torch.randint(0, 1000, (3, 50, 20)).cuda(), 20)
total = check("comm + cosine + merge")
assert total >= 40, total
print("CANARY_OK", total, flush=True)
