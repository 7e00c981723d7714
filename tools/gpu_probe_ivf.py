"""GPU bring-up probe for k-means / IVF-Flat / IVF-PQ (each case in its own subprocess)."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def clustered(n, d, n_clusters, seed, sigma=0.6):
    import torch
    g = torch.Generator().manual_seed(seed)
    cent = torch.randn(n_clusters, d, generator=g)
    lab = torch.randint(0, n_clusters, (n,), generator=g)
    return cent[lab] + sigma * torch.randn(n, d, generator=g), cent


def recall(ids, truth):
    hits = 0
    for a, b in zip(ids.tolist(), truth.tolist()):
        hits += len(set(a) & set(b))
    return hits / float(truth.numel())


def case_kmeans():
    import torch, cuvs_rag_b200 as b2
    x, cent = clustered(50000, 64, 32, 1, sigma=0.2)
    xg = x.to("cuda:0").to(torch.float16)
    c, lab = b2.kmeans_fit(xg, 32, iters=15, seed=3)
    torch.cuda.synchronize()
    c = c.cpu()
    d = torch.cdist(cent, c).min(dim=1).values
    inertia = ((x - c[lab.cpu().long()]) ** 2).sum(1).mean().item()
    ok = inertia < 64 * 0.2 * 0.2 * 2.5
    print("RESULT " + json.dumps({"name": "kmeans", "max_center_err": d.max().item(), "inertia": inertia,
                                  "ideal": 64 * 0.04, "ok": ok}))
    return 0 if ok else 1


def case_ivf(kind, dtype, metric, n=100000, d=128, nlist=256, nprobe=16, k=10, pq_dim=64):
    import torch, cuvs_rag_b200 as b2
    from oracle.exact import exact_knn
    tdt = {"fp16": torch.float16, "bf16": torch.bfloat16, "fp32": torch.float32}[dtype]
    x, _ = clustered(n, d, 200, 5)
    q, _ = clustered(500, d, 200, 5)
    q = x[torch.randperm(n, generator=torch.Generator().manual_seed(9))[:500]] + 0.1 * torch.randn(500, d, generator=torch.Generator().manual_seed(10))
    xt, qt = x.to(tdt), q.to(tdt)
    xg, qg = xt.to("cuda:0"), qt.to("cuda:0")
    t0 = time.time()
    if kind == "ivf_flat":
        ix = b2.NativeIndex.ivf_flat(xg, nlist, metric=metric, id_offset=7, kmeans_iters=10)
    else:
        ix = b2.NativeIndex.ivf_pq(xg, nlist, pq_dim, metric=metric, id_offset=7, kmeans_iters=10)
    torch.cuda.synchronize()
    tb = time.time() - t0
    sizes = ix.list_sizes()
    dd, ii = ix.search(qg, k, n_probes=nprobe, time_kernel=True)
    torch.cuda.synchronize()
    st = ix.last_stats()
    td, ti = exact_knn(xt.float(), qt.float(), k, metric)
    r = recall(ii.cpu() - 7, ti)
    # full probe == exact for ivf_flat
    r_full = None
    if kind == "ivf_flat":
        d2, i2 = ix.search(qg, k, n_probes=min(nlist, 128))
        r_full = recall(i2.cpu() - 7, ti)
    ok = int(sizes.sum()) == n and r > (0.85 if kind == "ivf_flat" else 0.6)
    gbs = st.algo_bytes / (st.kernel_ms * 1e-3) / 1e9 if st.kernel_ms > 0 else 0
    print("RESULT " + json.dumps({"name": f"{kind}_{dtype}_{metric}", "recall": r, "recall_probe128": r_full,
                                  "build_s": round(tb, 2), "lists_sum": int(sizes.sum()),
                                  "max_list": int(sizes.max()), "min_list": int(sizes.min()),
                                  "scan_ms": st.kernel_ms, "scan_GBs": round(gbs, 1), "ok": ok}))
    return 0 if ok else 1


CASES = {
    "kmeans": case_kmeans,
    "ivf_flat_fp16_l2": lambda: case_ivf("ivf_flat", "fp16", "sqeuclidean"),
    "ivf_flat_bf16_ip": lambda: case_ivf("ivf_flat", "bf16", "inner_product"),
    "ivf_flat_fp32_l2": lambda: case_ivf("ivf_flat", "fp32", "sqeuclidean", d=96),
    "ivf_flat_768": lambda: case_ivf("ivf_flat", "fp16", "sqeuclidean", n=200000, d=768, nlist=512, nprobe=32),
    "ivf_pq_fp16_l2": lambda: case_ivf("ivf_pq", "fp16", "sqeuclidean"),
    "ivf_pq_fp16_ip": lambda: case_ivf("ivf_pq", "fp16", "inner_product"),
    "ivf_pq_96": lambda: case_ivf("ivf_pq", "fp32", "sqeuclidean", d=768, pq_dim=96, n=60000, nlist=64, nprobe=16),
}

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--case":
        sys.exit(CASES[sys.argv[2]]())
    fails = 0
    for name in CASES:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", name],
                               capture_output=True, text=True, timeout=300)
            out = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            print(("PASS" if r.returncode == 0 else "FAIL"), name, out[-1] if out else "")
            if r.returncode != 0:
                fails += 1
                print("  stdout-tail:", r.stdout[-800:].replace("\n", "\n    "))
                print("  stderr-tail:", r.stderr[-2000:].replace("\n", "\n    "))
        except subprocess.TimeoutExpired:
            fails += 1
            print("TIMEOUT", name)
        sys.stdout.flush()
    print(f"ivf probe done, {fails} failing")
