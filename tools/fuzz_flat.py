"""Randomised check of the exact search (both kernel variants, every k range) against a torch
fp32 evaluation of the same rounded operands."""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cuvs_rag_b200 as b2

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 60
rnd = random.Random(seed)
bad = 0
for case in range(n_cases):
    n = rnd.choice([1, 37, 255, 257, 5000, 70001, 400000])
    d = rnd.choice([1, 7, 8, 33, 64, 100, 128, 384, 770, 1024])
    nq = rnd.choice([1, 2, 127, 128, 129, 256, 300, 4100])
    k = rnd.choice([1, 2, 10, 16, 17, 48, 49, 100, 128, 129, 1000, 2048])
    dtype = rnd.choice([torch.float16, torch.bfloat16, torch.float32])
    metric = rnd.choice(["sqeuclidean", "inner_product"])
    group = rnd.choice(["", "1", "2"])
    if n * d > 3e8 or nq * n > 3e9:
        continue
    tag = f"case {case}: n={n} d={d} nq={nq} k={k} {dtype} {metric} group={group or 'auto'}"
    g = torch.Generator(device="cuda").manual_seed(seed * 1000 + case)
    x = torch.randn(n, d, generator=g, device="cuda").to(dtype)
    q = torch.randn(nq, d, generator=g, device="cuda").to(dtype)
    if group:
        os.environ["B2VS_TC_GROUP"] = group
        b2._native.reload_env()
    else:
        os.environ.pop("B2VS_TC_GROUP", None)
        b2._native.reload_env()
    try:
        ix = b2.NativeIndex.flat(x, metric=metric, id_offset=11)
        dd, ii = ix.search(q, k)
        torch.cuda.synchronize()
        xf, qf = x.double(), q.double()
        kk = min(k, n)
        if metric == "sqeuclidean":
            full = (xf * xf).sum(1)[None, :] - 2.0 * qf @ xf.T + (qf * qf).sum(1)[:, None]
            td, ti = torch.topk(full, kk, dim=1, largest=False)
        else:
            full = qf @ xf.T
            td, ti = torch.topk(full, kk, dim=1, largest=True)
        scale = float(full.abs().max()) + 1e-6
        err = float((dd[:, :kk].double() - td).abs().max()) / scale
        inter = sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip((ii[:, :kk] - 11).cpu(), ti.cpu()))
        tail_ok = bool((ii[:, kk:] == -1).all())
        ok = err < 2e-3 and inter >= 0.995 * nq * kk and tail_ok
        print("ok " if ok else "BAD", tag, f"err={err:.2e} overlap={inter / (nq * kk):.4f}", flush=True)
        bad += 0 if ok else 1
    except Exception as e:
        print("EXC", tag, repr(e)[:300], flush=True)
        bad += 1
os.environ.pop("B2VS_TC_GROUP", None)
b2._native.reload_env()
print("bad cases:", bad)
sys.exit(1 if bad else 0)
