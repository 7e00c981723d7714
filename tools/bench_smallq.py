"""HBM-bound regime of the exact search: small query batches over a large bf16 database."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cuvs_rag_b200 as b2
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
d = 1024
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(5)
db = torch.empty((n, d), dtype=torch.bfloat16, device=dev)
for s in range(0, n, 1 << 20):
    e = min(n, s + (1 << 20))
    db[s:e] = torch.randn((e - s, d), generator=g, device=dev).to(torch.bfloat16)
ix = b2.NativeIndex.flat(db)
for nq in [1, 8, 16, 32, 48, 64, 96, 128]:
    q = torch.randn((nq, d), generator=g, device=dev).to(torch.bfloat16)
    ix.search(q, 10); torch.cuda.synchronize()
    ms = []
    for _ in range(3):
        ix.search(q, 10, time_kernel=True)
        st = ix.last_stats(); ms.append(st.kernel_ms)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ix.search(q, 10); e1.record(); torch.cuda.synchronize()
    print(json.dumps({"Q": nq, "fused_ms": round(min(ms), 3), "call_ms": round(e0.elapsed_time(e1), 3),
                      "db_GBs": round(n * d * 2 / min(ms) / 1e6, 1), "splits": st.n_splits, "grid": st.grid,
                      "launches": st.launches}), flush=True)
