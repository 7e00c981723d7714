"""IVF-Flat / IVF-PQ at config scale (C3 / C4-per-shard): build time, scan bandwidth, recall.
Usage: python tools/bench_ivf.py [flat|pq] [n] [dim] [n_lists] [n_probes] [nq] [pq_dim]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cuvs_rag_b200 as b2

kind = sys.argv[1] if len(sys.argv) > 1 else "flat"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
d = int(sys.argv[3]) if len(sys.argv) > 3 else 768
nlist = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
nprobe = int(sys.argv[5]) if len(sys.argv) > 5 else 32
nq = int(sys.argv[6]) if len(sys.argv) > 6 else 10_000
pq_dim = int(sys.argv[7]) if len(sys.argv) > 7 else 64
k = 10
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(7)
# clustered mixture (iid Gaussians in 768-d have no list structure)
ncomp = nlist
cent = torch.randn(ncomp, d, generator=g, device=dev)
x = torch.empty((n, d), dtype=torch.float16, device=dev)
chunk = 1 << 19
for s in range(0, n, chunk):
    e = min(n, s + chunk)
    lab = torch.randint(0, ncomp, (e - s,), generator=g, device=dev)
    x[s:e] = (cent[lab] + 0.3 * 1.4142 * torch.randn((e - s, d), generator=g, device=dev)).to(torch.float16)
qi = torch.randint(0, n, (nq,), generator=g, device=dev)
q = (x[qi].float() + 0.1 * torch.randn((nq, d), generator=g, device=dev)).to(torch.float16)
torch.cuda.synchronize()
t0 = time.time()
if kind == "flat":
    ix = b2.NativeIndex.ivf_flat(x, nlist, kmeans_iters=int(os.environ.get("KM_ITERS", 20)))
else:
    ix = b2.NativeIndex.ivf_pq(x, nlist, pq_dim, kmeans_iters=int(os.environ.get("KM_ITERS", 20)))
torch.cuda.synchronize()
build_s = time.time() - t0
sizes = ix.list_sizes()
out = {"refine": int(os.environ.get("REFINE", 0)), "kind": kind, "n": n, "dim": d, "n_lists": nlist, "n_probes": nprobe, "nq": nq, "build_s": round(build_s, 2),
       "list_min": int(sizes.min()), "list_max": int(sizes.max()), "list_mean": float(sizes.float().mean())}
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dd, ii = ix.search(q, k, n_probes=nprobe, time_kernel=True, refine_ratio=int(os.environ.get('REFINE', 0)))
    e1.record(); torch.cuda.synchronize()
    st = ix.last_stats()
    out[f"search_ms_{rep}"] = round(e0.elapsed_time(e1), 3)
    out[f"scan_ms_{rep}"] = round(st.kernel_ms, 3)
out["qps"] = round(nq / (out["search_ms_2"] * 1e-3), 1)
out["scan_GBs"] = round(st.algo_bytes / (st.kernel_ms * 1e-3) / 1e9, 1)
out["algo_bytes"] = st.algo_bytes
out["mean_candidates"] = st.mean_candidates
# recall@10 against the exact index on a query subset
nchk = min(nq, 1000)
flat = b2.NativeIndex.flat(x)
td, ti = flat.search(q[:nchk], k)
hits = 0
for a, b in zip(ii[:nchk].tolist(), ti.tolist()):
    hits += len(set(a) & set(b))
out["recall_at_10"] = hits / float(nchk * k)
print(json.dumps(out))
