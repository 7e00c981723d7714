"""Randomised differential test of the IVF search paths (run on a GPU box):
grouped tensor-core scan vs per-(query, probe) scan of the SAME index, full-probe vs the flat index,
large k vs the flat index — over random shapes, dtypes, metrics, k and batch sizes."""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cuvs_rag_b200 as b2

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rnd = random.Random(seed)
bad = 0
for case in range(n_cases):
    n = rnd.choice([900, 3000, 12000, 40000])
    d = rnd.choice([8, 24, 64, 72, 96, 128, 200, 256, 384])
    nlist = rnd.choice([1, 4, 16, 64, 200])
    nlist = min(nlist, n // 4)
    nprobe = rnd.choice([1, 3, nlist, min(nlist, 20)])
    nq = rnd.choice([1, 5, 64, 129, 700])
    k = rnd.choice([1, 5, 10, 33, 100, 128, 300, 2000])
    dtype = rnd.choice([torch.float16, torch.bfloat16, torch.float32])
    metric = rnd.choice(["sqeuclidean", "inner_product"])
    kind = rnd.choice(["flat", "flat", "pq"])
    g = torch.Generator().manual_seed(seed * 1000 + case)
    cent = torch.randn(max(4, nlist), d, generator=g)
    x = (cent[torch.randint(0, cent.shape[0], (n,), generator=g)] + 0.5 * torch.randn(n, d, generator=g)).to(dtype).cuda()
    q = (x[torch.randint(0, n, (nq,), generator=g)].float().cpu() + 0.1 * torch.randn(nq, d, generator=g)).to(dtype).cuda()
    tag = f"case {case}: {kind} n={n} d={d} nlist={nlist} nprobe={nprobe} nq={nq} k={k} {dtype} {metric}"
    try:
        if kind == "flat":
            ix = b2.NativeIndex.ivf_flat(x, nlist, metric=metric, kmeans_iters=4, id_offset=3)
        else:
            m = rnd.choice([mm for mm in (d // 2, d // 4, d // 8) if mm >= 1 and d % mm == 0] or [1])
            if n < 256:
                continue
            ix = b2.NativeIndex.ivf_pq(x, nlist, m, metric=metric, kmeans_iters=4, id_offset=3)
        if kind == "pq" and k > 128:
            try:
                dd, ii = ix.search(q, k, n_probes=nprobe)
            except RuntimeError as e:
                assert "grouped scan" in str(e) or "exceed" in str(e), e
                print("ok (unsupported, reported)", tag); continue
        os.environ["B2VS_IVF_GROUPED"] = "1"
        b2._native.reload_env()
        d1, i1 = ix.search(q, k, n_probes=nprobe)
        torch.cuda.synchronize()
        if k <= 128:
            os.environ["B2VS_IVF_GROUPED"] = "0"
            b2._native.reload_env()
            d0, i0 = ix.search(q, k, n_probes=nprobe)
            torch.cuda.synchronize()
            inter = sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip(i1.cpu(), i0.cpu()))
            valid = int((i0 >= 0).sum())
            need = 0.998 if kind == "flat" else 0.85
            ok = inter >= need * max(valid, 1) - 2
        else:
            ok = True
        # sortedness + id range
        fin = torch.isfinite(d1)
        srt = (d1[:, 1:] >= d1[:, :-1] - 1e-4) if metric == "sqeuclidean" else (d1[:, 1:] <= d1[:, :-1] + 1e-4)
        ok = ok and bool((srt | ~fin[:, 1:]).all()) and int(i1.max()) < n + 3 and bool(((i1 >= 3) | (i1 == -1)).all())
        if kind == "flat" and nprobe == nlist:
            flat = b2.NativeIndex.flat(x, metric=metric, id_offset=3)
            fd, fi = flat.search(q, min(k, n))
            kk = min(k, n)
            inter = sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip(i1[:, :kk].cpu(), fi.cpu()))
            ok = ok and inter >= 0.985 * nq * kk
        os.environ.pop("B2VS_IVF_GROUPED", None)
        b2._native.reload_env()
        print("ok " if ok else "BAD", tag, flush=True)
        bad += 0 if ok else 1
    except Exception as e:
        os.environ.pop("B2VS_IVF_GROUPED", None)
        b2._native.reload_env()
        print("EXC", tag, repr(e)[:300], flush=True)
        bad += 1
print("bad cases:", bad)
sys.exit(1 if bad else 0)
