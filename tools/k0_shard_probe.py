"""Exact search (K0) on one shard of the 8-GPU C2 run (1.25M x 768 bf16, 10K queries, k = 100) and
on the 1-GPU size, under B2VS_* settings: ms per batch and the full pass alone.
usage: k0_shard_probe.py "A=1&B=2;C=3;..." [rows]   (settings separated by ;, variables of a setting by &)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cuvs_rag_b200 as b2
from cuvs_rag_b200 import _native

settings = sys.argv[1].split(";") if len(sys.argv) > 1 else [""]
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 1_250_000
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(rows, 768, device="cuda", generator=g).to(torch.bfloat16)
q = torch.randn(10_000, 768, device="cuda", generator=g).to(torch.bfloat16)
ix = b2.NativeIndex.flat(x)
touched = set()
for s in settings:
    for kname in touched:
        os.environ.pop(kname, None)
    touched = set()
    for kv in filter(None, s.split("&")):
        kname, v = kv.split("=", 1)
        os.environ[kname] = v
        touched.add(kname)
    _native.reload_env()
    for _ in range(3):
        ix.search(q, 100)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ix.search(q, 100)
    e1.record(); torch.cuda.synchronize()
    ix.search(q, 100, time_kernel=True)
    torch.cuda.synchronize()
    st = ix.last_stats()
    print(json.dumps({"rows": rows, "setting": s, "ms_per_batch": round(e0.elapsed_time(e1) / 20, 3),
                      "full_pass_ms": round(st.kernel_ms, 3), "n_splits": st.n_splits, "grid": st.grid}), flush=True)
