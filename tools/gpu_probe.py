"""GPU bring-up probe: runs each exact-search case in its own subprocess (a CUDA fault in one
case cannot poison the next) and prints one PASS/FAIL line per case.  Dev tool, not a test."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [
    # name, N, D, Q, k, dtype, metric, n_splits
    ("tiny_bf16_l2", 1000, 64, 10, 5, "bf16", "sqeuclidean", 0),
    ("one_tile_k1", 256, 64, 128, 1, "bf16", "sqeuclidean", 0),
    ("mid_bf16_l2_k100", 20000, 768, 300, 100, "bf16", "sqeuclidean", 0),
    ("mid_fp16_ip_k10", 30000, 384, 257, 10, "fp16", "inner_product", 0),
    ("fp32_ip_k10", 20000, 384, 200, 10, "fp32", "inner_product", 0),
    ("fp32_l2_k10", 5000, 96, 64, 10, "fp32", "sqeuclidean", 0),
    ("odd_dim_100", 3000, 100, 33, 7, "bf16", "sqeuclidean", 0),
    ("forced_splits", 50000, 128, 512, 32, "bf16", "sqeuclidean", 7),
    ("k128", 9000, 256, 130, 128, "fp16", "sqeuclidean", 0),
    ("k_gt_n", 50, 64, 5, 100, "bf16", "sqeuclidean", 0),
    ("big_200k", 200000, 768, 1024, 100, "bf16", "sqeuclidean", 0),
    ("manyq_k1_d64", 256, 64, 50000, 1, "fp16", "sqeuclidean", 0),
    ("manyq_k1_d128", 256, 128, 100000, 1, "fp16", "sqeuclidean", 0),
    ("manyq_k1_d768", 512, 768, 60000, 1, "fp16", "sqeuclidean", 0),
    ("manyq_k10_d128", 5000, 128, 40000, 10, "bf16", "sqeuclidean", 0),
    ("multi_item_k100", 300000, 256, 2048, 100, "bf16", "sqeuclidean", 40),
]


def run_case(name, n, d, q, k, dtype, metric, n_splits):
    import torch
    import cuvs_rag_b200 as b2
    from oracle.exact import topk_parity_report

    tdt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[dtype]
    g = torch.Generator().manual_seed(1234)
    db = torch.randn(n, d, generator=g)
    qs = torch.randn(q, d, generator=g)
    if metric == "inner_product":
        db = torch.nn.functional.normalize(db, dim=1)
        qs = torch.nn.functional.normalize(qs, dim=1)
    db_t, qs_t = db.to(tdt), qs.to(tdt)
    dev = torch.device("cuda:0")
    db_g, qs_g = db_t.to(dev), qs_t.to(dev)
    t0 = time.time()
    ix = b2.NativeIndex.flat(db_g, metric=metric, id_offset=1000)
    dd, ii = ix.search(qs_g, k, n_splits=n_splits)
    torch.cuda.synchronize()
    t1 = time.time()
    st = ix.last_stats()
    # timing of a second call
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ix.search(qs_g, k, n_splits=n_splits)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ii_local = ii.cpu() - 1000
    ii_local[ii.cpu() < 0] = -1
    kk = min(k, n)
    rep = topk_parity_report(dd.cpu(), ii_local, db_t.float(), qs_t.float(), k, metric)
    tail_ok = True
    if k > n:
        tail_ok = bool((ii.cpu()[:, n:] == -1).all())
    rep.update(name=name, ms=round(ms, 3), first_s=round(t1 - t0, 3), splits=st.n_splits, grid=st.grid,
               tflops=round(st.algo_flops / (ms * 1e-3) / 1e12, 2), tail_ok=tail_ok)
    rep["ok"] = rep["ok"] and tail_ok
    print("RESULT " + json.dumps(rep))
    return 0 if rep["ok"] else 1


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--case":
        c = [c for c in CASES if c[0] == sys.argv[2]][0]
        sys.exit(run_case(*c))
    only = sys.argv[1:] if len(sys.argv) > 1 else None
    fails = 0
    for c in CASES:
        if only and c[0] not in only:
            continue
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", c[0]],
                               capture_output=True, text=True, timeout=240)
            out = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            status = "PASS" if r.returncode == 0 else "FAIL"
            print(f"{status} {c[0]} rc={r.returncode} {out[-1] if out else ''}")
            if r.returncode != 0:
                fails += 1
                print("  stdout-tail:", r.stdout[-1500:].replace("\n", "\n    "))
                print("  stderr-tail:", r.stderr[-2500:].replace("\n", "\n    "))
        except subprocess.TimeoutExpired:
            fails += 1
            print(f"TIMEOUT {c[0]}")
        sys.stdout.flush()
    print(f"probe done, {fails} failing")
    sys.exit(1 if fails else 0)
