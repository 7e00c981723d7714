"""Times the fused exact-search kernel for a few (k, kernel variant) combinations."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cuvs_rag_b200 as b2

n = int(os.environ.get("SWEEP_N", 2_000_000)); d = int(os.environ.get("SWEEP_D", 768)); nq = int(os.environ.get("SWEEP_Q", 10_000))
ks = [int(x) for x in os.environ.get("SWEEP_K", "1,10,100").split(",")]
reps = int(os.environ.get("SWEEP_REPS", 3))
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
db = torch.randn(n, d, generator=g, device=dev).to(torch.bfloat16)
q = torch.randn(nq, d, generator=g, device=dev).to(torch.bfloat16)
ix = b2.NativeIndex.flat(db)
N = b2._native
for k in ks:
    for name, flag in (("single", 2), ("pair", 4)):
        out_d = torch.empty((nq, k), dtype=torch.float32, device=dev); out_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
        sp = N.SearchParams(0, 0, int(os.environ.get("SWEEP_SPLITS", 0)), flag | 1)
        ms = []
        for r in range(reps + 1):
            rc = N.lib().b2vs_search(ix._h, q.data_ptr(), N.BF16, nq, k, ctypes.byref(sp), out_d.data_ptr(), out_i.data_ptr(), torch.cuda.current_stream().cuda_stream)
            assert rc == 0, N.lib().b2vs_last_error()
            st = ix.last_stats()
            if r > 0: ms.append(st.kernel_ms)
        best = min(ms)
        print(json.dumps({"k": k, "variant": name, "kernel_ms": round(best, 3), "tflops": round(2.0 * nq * n * d / best / 1e9, 1), "splits": st.n_splits, "grid": st.grid}), flush=True)
