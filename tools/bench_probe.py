"""Coarse-probe shaped flat search: nq queries vs n_lists centroids, k = n_probes; sweeps n_splits."""
import json, sys, torch
sys.path.insert(0, ".")
import cuvs_rag_b200 as b2

n_lists, dim, nq, k = (int(a) for a in sys.argv[1:5])
g = torch.Generator().manual_seed(0)
x = torch.randn(n_lists, dim, generator=g).to(torch.bfloat16).cuda()
q = torch.randn(nq, dim, generator=g).to(torch.bfloat16).cuda()
ix = b2.NativeIndex.flat(x, metric="sqeuclidean")
out = {}
for splits in (0, 1, 2, 3, 4, 8):
    for _ in range(3):
        ix.search(q, k, n_splits=splits)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ix.search(q, k, n_splits=splits)
    e1.record(); torch.cuda.synchronize()
    out[f"splits{splits}_ms"] = round(e0.elapsed_time(e1) / 10, 4)
print(json.dumps(out))
