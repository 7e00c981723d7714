"""Large-k exact search timing (the reference's top-2000 mode)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cuvs_rag_b200 as b2
n, d, nq = 2_000_000, 768, 2000
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
db = torch.randn(n, d, generator=g, device=dev).to(torch.bfloat16)
q = torch.randn(nq, d, generator=g, device=dev).to(torch.bfloat16)
ix = b2.NativeIndex.flat(db)
for k in (100, 500, 2000):
    ix.search(q, k); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ix.search(q, k); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"n": n, "dim": d, "nq": nq, "k": k, "ms": round(ms, 2), "qps": round(nq / ms * 1e3),
                      "tflops": round(2.0 * nq * n * d / ms / 1e9, 1), "launches": ix.last_stats().launches}), flush=True)
