"""Measures the non-headline BASELINE.json configs on one GPU: C1 (exact IP, 100K x 384 fp32, 1K
queries, k=10) and the C5 query-batch sweep (exact L2, N x 1024 bf16, Q = 1 .. 65536, k=10)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cuvs_rag_b200 as b2

dev = torch.device("cuda:0")
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists("MEASURED_PEAKS.json") else {"bf16_tflops_sustained": 1403.3, "hbm_gbs": 6459.0}


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


which = sys.argv[1] if len(sys.argv) > 1 else "c1"
if which == "c1":
    g = torch.Generator(device=dev).manual_seed(1234)
    db = torch.nn.functional.normalize(torch.randn(100_000, 384, generator=g, device=dev), dim=1)
    q = torch.nn.functional.normalize(torch.randn(1000, 384, generator=g, device=dev), dim=1)
    ix = b2.NativeIndex.flat(db, metric="inner_product")
    ms = timed(lambda: ix.search(q, 10), 50)
    flops = 2.0 * 1000 * 100_000 * 384
    print(json.dumps({"config": "C1 exact IP k=10 100Kx384 fp32 Q=1000", "ms": round(ms, 4), "qps": round(1000 / ms * 1e3),
                      "tflops_algorithmic": round(flops / ms / 1e9, 2), "note": "fp32 via bf16 hi/lo split = 3x MMA work"}))
else:
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 50_000_000
    d = 1024
    g = torch.Generator(device=dev).manual_seed(5)
    db = torch.empty((n, d), dtype=torch.bfloat16, device=dev)
    for s in range(0, n, 1 << 20):
        e = min(n, s + (1 << 20))
        db[s:e] = torch.randn((e - s, d), generator=g, device=dev).to(torch.bfloat16)
    ix = b2.NativeIndex.flat(db)
    for nq in [1, 4, 16, 64, 256, 1024, 4096, 16384, 65536]:
        q = torch.randn((nq, d), generator=g, device=dev).to(torch.bfloat16)
        reps = 3 if nq <= 4096 else 1
        ms = timed(lambda: ix.search(q, 10), reps)
        flops = 2.0 * nq * n * d
        byts = float(n) * d * 2
        print(json.dumps({"config": f"C5 exact L2 k=10 {n}x{d} bf16", "Q": nq, "ms": round(ms, 3), "qps": round(nq / ms * 1e3, 1),
                          "tflops": round(flops / ms / 1e9, 1), "db_GBs": round(byts / ms / 1e6, 1),
                          "frac_tensor": round(flops / ms / 1e9 / peaks["bf16_tflops_sustained"], 3),
                          "frac_hbm": round(byts / ms / 1e6 / peaks["hbm_gbs"], 3)}), flush=True)
